"""Partitioned 4096-bin histogram (decode.cu HIST == 3) against the in-place form on the bench's decode input.
    MKID_DEC_TIMING=1 python scripts/prof_part.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mkids_sdr_b200 import _lib, synth
from mkids_sdr_b200.decode import PhotonDecoder
ctx = _lib.default_context(0)
R, npix, secs = 8, 253, 10
streams, _ = synth.photon_streams(10 ** 7, R, npix, secs, seed=1234)
lens = [len(s) for s in streams]; reps = 16
words = np.tile(np.concatenate(streams), reps)
offs = np.concatenate([[0], np.cumsum(lens * reps)]).astype(np.int64)
roach = np.tile(np.arange(R), reps)
dw = ctx.to_device(words)
ref = None
for mode in sys.argv[1:] or ("0", "1"):
    os.environ["MKID_DEC_PART"] = mode
    dec = PhotonDecoder(R, npix, secs, 2500, "peak", 4096, None, ctx=ctx)
    for _ in range(3):
        dec.decode_words(dw, offs, roach, want_stats=False, want_sec=False)
    h = dec.hist()
    ref = h if ref is None else ref
    print("mode", mode, "hist sum", int(h.sum()), "same as first:", bool(np.array_equal(h, ref)), flush=True)
