"""K4 (channelize_kernel) micro-benchmark at the bench.py configuration: 8 boards x 2^25 samples, 253 driven tones,
N_lut 2^19, matched_30us.  Prints the mean device time of the kernel (CUDA events recorded around every launch on the
context stream) and a sha256 over phase rows + photon words of a shorter call, so that kernel variants can be checked
for bit-identical output against each other.

    [MKIDGPU_LIB=path/to/variant.so] python scripts/k4_bench.py [steps]
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mkids_sdr_b200 import _lib  # noqa: E402
from mkids_sdr_b200.chain import ReadoutChain  # noqa: E402
from mkids_sdr_b200.channelizer import synth_adc  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B, n, N_LUT = 8, 1 << 25, 2 ** 19
ctx = _lib.default_context(0)
chain, boards = ReadoutChain.synthetic(B, N_LUT, 253, seed0=42, ctx=ctx, exptime=64, threshold=-4000)
tone_bins = np.stack([bd['tone_bins'] for bd in boards])
iq = torch.empty((B, n, 2), dtype=torch.int16, device='cuda')
synth_adc(B, n, tone_bins, n_lut=N_LUT, pulse_rate=1000.0, seed=1000, out=iq, ctx=ctx)
ctx.sync()
# bit-identity probe: first 2^22 samples of every board, phase rows + words
chain.reset()
words, ph = chain.chan.process(iq[:, :1 << 22].contiguous(), detect=True, want_phase=True)
h = hashlib.sha256(ph.tobytes())
for w in words:
    h.update(w.tobytes())
chain.reset()
for _ in range(3):
    chain.process_async(iq, n=n)
chain.sync_state()
for _ in range(steps):
    chain.process_async(iq, n=n)
chain.sync_state()
ms = chain.chan.kernel_ms_sum(min(steps, 64)) / min(steps, 64)
print('%s  K4 %.4f ms per 8 x 2^25 samples  (%.1f GS/s, %.1f GB/s)  sha %s  words %d'
      % (os.path.basename(os.environ.get('MKIDGPU_LIB', 'libmkidgpu.so')), ms, B * n / ms / 1e6, B * n * 4 / ms / 1e6,
         h.hexdigest()[:16], sum(len(w) for w in words)))
