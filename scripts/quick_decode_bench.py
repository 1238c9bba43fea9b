import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np
from mkids_sdr_b200 import _lib, synth
from mkids_sdr_b200.decode import PhotonDecoder
ctx = _lib.default_context(0)
R, npix, secs = 8, 253, 10
streams, _ = synth.photon_streams(10**7, R, npix, secs, seed=1234)
lens = [len(s) for s in streams]
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
words = np.concatenate(streams)
n = words.size
reps = 16   # replicate to exceed L2
big = np.tile(words, reps)
offs = np.concatenate([[0], np.cumsum(lens * reps)]).astype(np.int64)
roach = np.tile(np.arange(R), reps)
dw = ctx.to_device(big)
for field, nb in ((None, 0), ('p1', 10), ('peak', 4096)):
    lut = (np.arange(4096) * 10 // 4096) if field == 'p1' else None
    dec = PhotonDecoder(R, npix, secs, 2500, field, max(nb, 1), lut, ctx=ctx)
    for it in range(3):
        dec.decode_words(dw, offs, roach, want_stats=False, want_sec=False)
    ctx.sync()
    ctx.record(0)
    K = 5
    t0 = time.perf_counter()
    for it in range(K):
        dec.decode_words(dw, offs, roach, want_stats=False, want_sec=False)
    t_host = (time.perf_counter() - t0) / K * 1e3
    ctx.record(1)
    ms = ctx.elapsed_ms(0, 1) / K
    print(field, nb, 'host ms/call', round(t_host, 4), 'ms/pass', ms, 'Gwords/s', big.size / ms / 1e6, 'GB/s', big.size * 8 / ms / 1e6, 'frac', big.size * 8 / ms / 1e6 / 6552)
# calibration: plain read of the same number of bytes (torch reduction) on this box
import torch
x = torch.zeros(big.size, dtype=torch.int64, device='cuda')
for it in range(3):
    x.sum()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(5):
    x.sum()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print('calibration torch sum ms', ms, 'GB/s', big.size * 8 / ms / 1e6)
