"""Merged-list run (timing / ncu): python scripts/prof_merged.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mkids_sdr_b200 import _lib, synth
from mkids_sdr_b200.decode import PhotonDecoder
ctx = _lib.default_context(0)
R, npix, secs = 8, 253, 10
streams, _ = synth.photon_streams(10**7, R, npix, secs, seed=1234)
lens = [len(s) for s in streams]
reps = 4
words = np.tile(np.concatenate(streams), reps)
offs = np.concatenate([[0], np.cumsum(lens * reps)]).astype(np.int64)
roach = np.tile(np.arange(R), reps)
sec0 = np.repeat(np.arange(reps) * secs, R).astype(np.int32)
dw = ctx.to_device(words)
dec = PhotonDecoder(R, npix, secs * reps, 2500, None, 1, None, ctx=ctx)
for it in range(2):
    lw, lo, so = dec.decode_merged(dw, offs, roach, sec0, n_words=words.size)
print('total merged', lo[-1], 'of', words.size)
