"""One board on one GPU (the per-GPU share of config 4 at 8 GPUs): device time per step, host time to queue a step,
single-stream vs two-context pipeline.  python scripts/strong_probe.py [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mkids_sdr_b200 import _lib
from mkids_sdr_b200.chain import ReadoutChain
from mkids_sdr_b200.channelizer import synth_adc
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
B, n, N_LUT = int(os.environ.get('PROBE_B', '1')), 1 << 25, 2 ** 19
ctx = _lib.default_context(0)
for pipelined in (False, True):
    chain, boards = ReadoutChain.synthetic(B, N_LUT, 253, seed0=42, ctx=ctx, exptime=16, n_roaches_total=8, n_bins=4096,
                                           want_merged=True, pipelined=pipelined)
    chain.derive_thresholds(boards)
    tb = np.stack([bd['tone_bins'] for bd in boards])
    iq = torch.empty((B, n, 2), dtype=torch.int16, device='cuda')
    synth_adc(B, n, tb, n_lut=N_LUT, pulse_rate=float(os.environ.get('PROBE_RATE', '1000')), seed=1000, out=iq, ctx=ctx)
    ctx.sync()
    for _ in range(5):
        chain.process_async(iq, n=n)
    chain.sync_state()
    ctx.record(0)
    t0 = time.time()
    for _ in range(steps):
        chain.process_async(iq, n=n)
    t_host = time.time() - t0
    chain.join()
    ctx.record(1)
    chain.sync_state()
    dev = ctx.elapsed_ms(0, 1)
    k4 = chain.chan.kernel_ms_sum(min(steps, 64)) / min(steps, 64)
    print('boards %d pipelined %d: device %.4f ms/step, host enqueue %.4f ms/step, K4 %.4f ms  -> %.1f GS/s per GPU'
          % (B, pipelined, dev / steps, t_host * 1e3 / steps, k4, B * n / (dev / steps) / 1e6))
    del chain, iq
