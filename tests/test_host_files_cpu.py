"""Host-side formats around the decode path that need no GPU: the beammap dataset-name parse (PacketMaster.c:880-904), the
per-second text image PacketMaster leaves for the dashboard (write_quicklook_image_v2, PacketMaster.c:679-727; read back the
way ArconsDashboard.py:635 does, numpy.loadtxt) and the oracle of pulses.QuickLook (lib/pulses.py:210-236)."""
import os

import numpy as np

from mkids_sdr_b200.decode import parse_beammap, write_quicklook_file
from oracle import decode as odec


def test_beammap_names():
    npix = 253
    names = [['/r0/p0/', '/r7/p252/t1319000000'], ['r3/p17', '/rX/p3/'], ['/r2/', '']]
    adr = parse_beammap(names, npix)
    assert adr.dtype == np.int32 and adr.shape == (3, 2)
    # atoi past the leading letter: digits only, 0 when there are none; a missing token counts as 0
    assert adr.tolist() == [[0, 7 * npix + 252], [3 * npix + 17, 3], [2 * npix, 0]]


def test_quicklook_file_format(tmp_path):
    img = np.array([[0, 1, 2499], [65535, 7, 8]], dtype=np.uint16)
    path = write_quicklook_file(str(tmp_path / 'obs_20110726-114310.h5'), img, 12)
    assert path == str(tmp_path / 'bin' / 'obs_20110726-114310_12.txt')
    text = open(path).read()
    assert text == '0 1 2499 \n65535 7 8 \n'                          # "%d " after every value, a line per image row
    assert np.array_equal(np.loadtxt(path), img)
    assert os.listdir(tmp_path / 'bin') == ['obs_20110726-114310_12.txt']     # the lock file is gone again


def test_quicklook_skysub_oracle():
    rng = np.random.default_rng(3)
    counts = rng.integers(0, 2500, (5, 12))
    adr = rng.permutation(12).reshape(3, 4)
    out = odec.quicklook_skysub(counts, adr, 1, 4)
    image = counts[1:4].sum(axis=0)[adr].astype(np.float64)
    assert out.dtype == np.float32 and np.array_equal(out, np.float32(image - np.median(image)))
    assert np.array_equal(odec.quicklook_skysub(counts, adr, 2, 2), np.zeros((3, 4), np.float32))


def test_quicklook_oracle_matches_reference_run():
    """oracle.decode.quicklook_skysub against pulses.QuickLook itself (lib/pulses.py:210-236), executed on a stand-in
    observation file by tests/golden/make_golden_refrun.py::run_pulses_quicklook."""
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'quicklook_golden.npz'))
    for i in range(3):
        t0, t1 = (int(v) for v in g['ql_span_%d' % i])
        want = g['ql_skysub_%d' % i]
        got = odec.quicklook_skysub(g['ql_counts'], g['ql_adr'], t0, t1)
        assert want.dtype == np.float32 and got.dtype == np.float32 and np.array_equal(got, want), (t0, t1)
