"""Where the detection tail of the two-context pipeline spends its time when it runs beside the next channelizer kernel:
device time of detect (resolve + scan + emit), decode and merged list on the second context, per batch.
    [MKID_TAIL_DELAY_US=20] [PROBE_B=8] python scripts/tail_probe.py [steps]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mkids_sdr_b200 import _lib
from mkids_sdr_b200.chain import ReadoutChain
from mkids_sdr_b200.channelizer import synth_adc
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
B, n, N_LUT = int(os.environ.get('PROBE_B', '8')), 1 << 25, 2 ** 19
ctx = _lib.default_context(0)
chain, boards = ReadoutChain.synthetic(B, N_LUT, 253, seed0=42, ctx=ctx, exptime=16, n_roaches_total=8, n_bins=4096,
                                       want_merged=True, pipelined=True)
chain.derive_thresholds(boards)
tb = np.stack([bd['tone_bins'] for bd in boards])
iq = torch.empty((B, n, 2), dtype=torch.int16, device='cuda')
synth_adc(B, n, tb, n_lut=N_LUT, pulse_rate=1000.0, seed=1000, out=iq, ctx=ctx)
ctx.sync()
for _ in range(4):
    chain.process_async(iq, n=n)
chain.sync_state()
# the body of ReadoutChain.process_async (pipelined branch) with event marks on the second context
c, b, self = chain.ctx, chain.ctx2, chain
start = np.arange(B, dtype=np.int64) * self._cap
caps = np.full(B, self._cap, dtype=np.int64)
acc = np.zeros(4)
for k in range(steps):
    par = self._k & 1
    if self._k >= 2:
        c.wait_event(b, 40 + par)
    c._check(c.lib.mkid_chan_process(c.h, self.chan.h, _lib.ptr(iq), int(n), 2, None, 0, None, None))
    self.chan.t_consumed += n // 512
    c.record(42 + par)
    b.wait_event(c, 42 + par)
    e0 = 20 if k == steps - 3 else 10
    b.record(e0)
    b._check(b.lib.mkid_chan_detect_pending(b.h, self.chan.h, _lib.ptr(self._words_dev), int(self._cap), None))
    b.record(e0 + 1)
    self.dec.decode_words_dev(self._words_dev, start, caps, self.chan.n_words_dev(), self.roach0 + np.arange(B),
                              self._sec_dev[self._sec_cur], self._sec_dev[1 - self._sec_cur], B * self._cap)
    b.record(e0 + 2)
    b._check(b.lib.mkid_merge_words_dev(b.h, _lib.ptr(self._words_dev), _lib.ptr(start), _lib.ptr(caps), self.chan.n_words_dev(),
                                        _lib.ptr(self._sec_dev[self._sec_cur]), B, ctypes.byref(self.dec.cfg),
                                        _lib.ptr(self.merged_words_dev), B * self._cap, _lib.ptr(self.merged_offsets_dev)))
    b.record(e0 + 3)
    b.record(40 + par)
    self._k += 1
    self._sec_cur = 1 - self._sec_cur
chain.sync_state()
b.sync()
acc += [b.elapsed_ms(20, 21), b.elapsed_ms(21, 22), b.elapsed_ms(22, 23), 1]      # batch steps - 3: two more K4 launches follow it
print('boards %d delay %s: detect %.3f ms, decode %.3f ms, merged list %.3f ms per batch (second context, beside K4); K4 %.4f ms'
      % (B, os.environ.get('MKID_TAIL_DELAY_US', 'default'), acc[0] / acc[3], acc[1] / acc[3], acc[2] / acc[3],
         chain.chan.kernel_ms_sum(8) / 8))
