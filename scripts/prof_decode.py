"""Small decode run for ncu: python scripts/prof_decode.py [field n_bins]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mkids_sdr_b200 import _lib, synth
from mkids_sdr_b200.decode import PhotonDecoder
field = sys.argv[1] if len(sys.argv) > 1 else 'none'
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 0
field = None if field == 'none' else field
ctx = _lib.default_context(0)
R, npix, secs = 8, 253, 10
streams, _ = synth.photon_streams(10**7, R, npix, secs, seed=1234)
lens = [len(s) for s in streams]
reps = 4
words = np.tile(np.concatenate(streams), reps)
offs = np.concatenate([[0], np.cumsum(lens * reps)]).astype(np.int64)
roach = np.tile(np.arange(R), reps)
dw = ctx.to_device(words)
lut = (np.arange(4096) * 10 // 4096) if nb == 10 else None
dec = PhotonDecoder(R, npix, secs, 2500, field, max(nb, 1), lut, ctx=ctx)
for it in range(3):
    dec.decode_words(dw, offs, roach, want_stats=False)
ctx.sync()
ctx.record(0)
for it in range(3):
    dec.decode_words(dw, offs, roach, want_stats=False)
ctx.record(1)
ms = ctx.elapsed_ms(0, 1) / 3
print(field, nb, 'ms', ms, 'GB/s', words.size * 8 / ms / 1e6)
