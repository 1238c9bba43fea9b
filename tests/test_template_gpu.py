"""GPU parity of MakeTemplate (pulses.py:239-427) against the literal NumPy restatement.
Tolerance parity: NumPy evaluates arctan2 / unwrap in float32 with libm, the kernels with CUDA's atan2f
(<= 2 ulp): per-sample phase differences of ~1e-5 deg, accept / reject decisions identical on data with margins."""
import numpy as np
import pytest

from oracle import template as otpl

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ctx():
    from mkids_sdr_b200 import _lib
    return _lib.default_context(0)


def test_make_template_matches_reference_loop(ctx):
    from mkids_sdr_b200.template import MakeTemplate
    I, Q = otpl.fake_pulses(1300, seed=1)                      # > 1000: the first pass only uses 1000 pulses
    Ir, Qr = I.copy(), Q.copy()
    ref = otpl.make_template(Ir, Qr)
    got = MakeTemplate(I, Q, ctx=ctx)
    assert ref['count1'] > 500 and ref['count'] > 600 and ref['flag'] == 0
    assert got['accepted1'] == ref['accepted1'] and got['accepted2'] == ref['accepted2']
    assert (got['count1'], got['count'], got['flag'], got['pstart']) == (ref['count1'], ref['count'], ref['flag'], ref['pstart'])
    assert abs(got['pm'] - ref['pm']) < 1e-4 and abs(got['pdev'] - ref['pdev']) < 1e-4
    assert np.max(np.abs(got['tP'] - ref['tP'])) < 2e-6        # templates are normalised to a peak of ~1
    assert np.max(np.abs(got['tPf'] - ref['tPf'])) < 2e-6
    assert np.max(np.abs(got['noise'] - ref['noise']) / ref['noise']) < 1e-3
    assert np.array_equal(got['noiseidx'], ref['noiseidx'])
    # the in-place shift of the table rows (float32): identical
    assert np.array_equal(I, Ir) and np.array_equal(Q, Qr)
