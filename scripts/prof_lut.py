"""LUT synthesis run for ncu (BASELINE config 1 at batch 64: comb, DDS tables, DRAM image; tables left in HBM): python scripts/prof_lut.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mkids_sdr_b200 import _lib, lut
ctx = _lib.default_context(0)
N, T, FS = 2 ** 19, 256, 512e6
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
k = np.sort(np.random.default_rng(0).choice(np.arange(-N // 2 + 1, N // 2), T, replace=False))
f = (k % N) * FS / N
amps = 10 ** (-(np.random.default_rng(1).integers(0, 20, T)) / 20.)
ff = np.tile(f, (batch, 1)); aa = np.tile(amps, (batch, 1))
oi, oq = ctx.alloc(batch * N * 2), ctx.alloc(batch * N * 2)
for it in range(2):
    lut.comb_lut(ff, FS, N, aa, ctx=ctx, out_I=oi, out_Q=oq)
res = FS / N
resid = np.rint((f - np.rint(f * 512 / FS) * FS / 512) / res) * res
rr = np.tile(resid, (batch, 1))
di, dq, img = ctx.alloc(batch * N * 2), ctx.alloc(batch * N * 2), ctx.alloc(batch * N * 8)
for it in range(2):
    lut.dds_lut(rr, np.zeros_like(rr), FS, N, ctx=ctx, out_I=di, out_Q=dq)
    lut.pack_dram(oi, oq, di, dq, ctx=ctx, n=batch * N, out=img)
ctx.sync()
print('comb_lut + dds_lut + pack_dram batch', batch, 'done')
