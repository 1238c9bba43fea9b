"""The reference's literal `freqCombLUT` run ONCE at the BASELINE size (256 tones x 2^19 samples).

    python tests/golden/make_golden_refrun_fullsize.py [/root/reference]          (about 5-10 minutes of CPython loops)

Same mechanism as make_golden_refrun.py (method source read from the reference tree at run time, Python-2 -> 3 edits,
executed against a mock `self`); only the sha256 of the int16 images of I and Q, the scale factor and a few samples are
stored (tests/golden/refrun_fullsize_lut.json).  Pins row a1 at the size the benchmark runs:
  DataReadout/ChannelizerControls/ROACH_Setup.py  freqCombLUT :416-475
"""
import hashlib
import json
import os
import sys
import time
import types
import warnings

import numpy

here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, here)
import make_golden_refrun as g      # noqa: E402  (its __main__ part does not run on import)


def main():
    Ref, printed, ns = g.build_class(os.path.join(g.CC, 'ROACH_Setup.py'), ['freqCombLUT'])
    fs, N, T = 512e6, 2 ** 19, 256
    k = numpy.sort(numpy.random.default_rng(0).choice(numpy.arange(-N // 2 + 1, N // 2), T, replace=False))
    f = [float(v) for v in (k % N) * fs / N]
    attens = numpy.random.default_rng(1).integers(0, 20, T).astype(float)
    amin = attens.min()
    amps = [10 ** (+(amin - a) / 20.) for a in attens]             # define_DAC_LUT, ROACH_Setup.py:499-501
    s = g.Self()
    s.freqCombLUT = types.MethodType(Ref.freqCombLUT, s)
    s.sampleRate, s.freqRes = fs, fs / N
    s.minimumAttenuation, s.previous_scale_factor, s.last_scale_factor = 10, 5000.0, None
    t0 = time.time()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        I, Q = s.freqCombLUT('yes', f, fs, fs / N, amps)
    dt = time.time() - t0
    I = numpy.asarray(I, dtype=numpy.int64); Q = numpy.asarray(Q, dtype=numpy.int64)
    assert I.size == N and Q.size == N and abs(I).max() <= 32767 and abs(Q).max() <= 32767
    out = dict(N=N, T=T, seconds=dt, scale_factor=repr(float(s.scale_factor)),
               sha256_I=hashlib.sha256(I.astype('<i2').tobytes()).hexdigest(),
               sha256_Q=hashlib.sha256(Q.astype('<i2').tobytes()).hexdigest(),
               I_first=[int(v) for v in I[:8]], Q_first=[int(v) for v in Q[:8]],
               I_sum=int(I.sum()), Q_sum=int(Q.sum()), I_absmax=int(abs(I).max()), Q_absmax=int(abs(Q).max()),
               source='ROACH_Setup.py freqCombLUT(echo="yes") executed from the reference tree; tones: '
                      'k=sort(default_rng(0).choice(arange(-N/2+1,N/2),256)), f=(k mod N) fs/N; attens default_rng(1).integers(0,20,256)')
    json.dump(out, open(os.path.join(here, 'refrun_fullsize_lut.json'), 'w'), indent=1)
    print(out)


if __name__ == '__main__':
    main()
