"""Host side of the software channelizer + pulse detection (K4/K5).

`Channelizer` holds the boards (feedlines) one GPU processes and their streaming state.  Its
configuration comes straight from the reference's control plane (see include/mkidgpu.h):
bins/residuals from `select_bins`, the DDS LUT from `define_DDS_LUT`, the quantised taps of
`loadFIRcoeffs`, the centres of `loadIQcenters` and the thresholds of `loadThresholds`.
"""
import ctypes

import numpy as np

from . import _lib

N_CH = 256
FIR_TAPS = 26
MIN_CALL_SAMPLES = 59392       # input history length: a call must bring at least this many samples


class Channelizer:
    def __init__(self, n_boards, n_lut, mean_len=20, holdoff=1000, peak_win=32, ctx=None):
        self.ctx = ctx or _lib.default_context()
        self.n_boards, self.n_lut = int(n_boards), int(n_lut)
        self.M, self.L, self.W = int(mean_len), int(holdoff), int(peak_win)
        prm = _lib.ChanParams(self.n_boards, self.n_lut, self.M, self.L, self.W, 0)
        h = ctypes.c_void_p()
        c = self.ctx
        c._check(c.lib.mkid_chan_create(c.h, ctypes.byref(prm), ctypes.byref(h)))
        self.h = h
        self.t_consumed = 0

    # ------------------------------------------------------------------ configuration
    def set_fir(self, fir_int):
        """fir_int: the 26 quantised taps c = int(tap*2047) (ROACH_Pulses.py:69,88-89)."""
        f = np.ascontiguousarray(fir_int, dtype=np.int32)
        assert f.size == FIR_TAPS
        c = self.ctx
        c._check(c.lib.mkid_chan_set_fir(c.h, self.h, _lib.ptr(f)))

    def set_window(self, h):
        h = np.ascontiguousarray(h, dtype=np.float32)
        assert h.size == 2048
        c = self.ctx
        c._check(c.lib.mkid_chan_set_window(c.h, self.h, _lib.ptr(h)))

    def set_board(self, board, bins, I_dds, Q_dds, zero_ch=None, centers_i=None, centers_q=None, thresholds=None):
        a = lambda x, dt: None if x is None else np.ascontiguousarray(x, dtype=dt)
        bins = a(bins, np.int32)
        I = a(I_dds, np.int16); Q = a(Q_dds, np.int16)
        assert bins.size == N_CH and I.size == self.n_lut and Q.size == self.n_lut
        z = a(zero_ch, np.uint8); ci = a(centers_i, np.int32); cq = a(centers_q, np.int32); th = a(thresholds, np.int32)
        c = self.ctx
        c._check(c.lib.mkid_chan_set_board(c.h, self.h, int(board), _lib.ptr(bins), _lib.ptr(I), _lib.ptr(Q),
                                           _lib.ptr(z), _lib.ptr(ci), _lib.ptr(cq), _lib.ptr(th)))

    def set_thresholds(self, board, thresholds):
        th = np.ascontiguousarray(thresholds, dtype=np.int32)
        c = self.ctx
        c._check(c.lib.mkid_chan_set_thresholds(c.h, self.h, int(board), _lib.ptr(th)))

    def set_f32_phase_out(self, dev_buffer):
        """Test hook: DeviceBuffer (float32 [B][n/512][256]) receiving the unquantised phase, or None."""
        c = self.ctx
        c._check(c.lib.mkid_chan_set_f32_phase_out(c.h, self.h, _lib.ptr(dev_buffer)))

    def reset(self):
        c = self.ctx
        c._check(c.lib.mkid_chan_reset(c.h, self.h))
        self.t_consumed = 0

    # ------------------------------------------------------------------ streaming
    def words_capacity(self, n):
        """Upper bound of words per board for a call of n samples."""
        T = n // 512
        return N_CH * (T // self.L + 2) + T // 10 ** 6 + 2

    def process(self, iq, n=None, detect=True, want_phase=False, words_out=None, words_cap=None):
        """iq: int16 [n_boards][n][2] (numpy / torch / DeviceBuffer).  Returns (words, phase):
        words = list of u64 arrays (one per board) or None, phase = int16 [B][n/512][256] or None."""
        if n is None:
            n = iq.shape[-2]
        T = n // 512
        c = self.ctx
        cap = int(words_cap or self.words_capacity(n))
        words = n_words = None
        if detect:
            words = words_out if words_out is not None else np.empty((self.n_boards, cap), dtype=np.uint64)
            n_words = np.zeros(self.n_boards, dtype=np.int32)
        phase = np.empty((self.n_boards, T, N_CH), dtype=np.int16) if want_phase else None
        rc = c.lib.mkid_chan_process(c.h, self.h, _lib.ptr(iq), int(n), 1 if detect else 0, _lib.ptr(words), cap,
                                     _lib.ptr(n_words), _lib.ptr(phase))
        self.t_consumed += T if rc in (0, _lib.MKID_EINVAL) and n % 2048 == 0 else 0      # (a too small word buffer still consumes the call)
        c._check(rc)
        if detect and words_out is None:
            words = [words[b, :n_words[b]].copy() for b in range(self.n_boards)]
        elif detect:
            words = (words, n_words)
        return words, phase

    def process_async(self, iq, words_out, words_cap, n=None):
        """Asynchronous process(): iq and words_out in device memory; nothing is copied back and the call does not
        synchronise.  The per-board word counts stay on the device (n_words_dev())."""
        if n is None:
            n = iq.shape[-2]
        c = self.ctx
        c._check(c.lib.mkid_chan_process(c.h, self.h, _lib.ptr(iq), int(n), 1, _lib.ptr(words_out), int(words_cap), None, None))
        self.t_consumed += n // 512

    def n_words_dev(self):
        """Device address of the int32 [n_boards] word counts of the last process call."""
        out = ctypes.c_void_p()
        c = self.ctx
        c._check(c.lib.mkid_chan_n_words_dev(c.h, self.h, ctypes.byref(out)))
        return out.value

    def n_words(self):
        """Word counts of the last process call (synchronises)."""
        nw = np.zeros(self.n_boards, dtype=np.int32)
        c = self.ctx
        c._check(c.lib.mkid_memcpy(c.h, _lib.ptr(nw), self.n_words_dev(), nw.nbytes))
        c.sync()
        return nw

    def overflowed(self, clear=True):
        """True if some process call since the last clear produced more words than its word buffer could hold (the
        only way an asynchronous call can report it).  Synchronises."""
        flag = ctypes.c_int32()
        c = self.ctx
        c._check(c.lib.mkid_chan_overflowed(c.h, self.h, ctypes.byref(flag), 1 if clear else 0))
        return bool(flag.value)

    def kernel_ms_sum(self, last_n):
        """Summed device time of the channelize kernel (K4) over the last `last_n` (<= 64) process calls."""
        ms = ctypes.c_float()
        c = self.ctx
        c._check(c.lib.mkid_chan_kernel_ms_sum(c.h, self.h, int(last_n), ctypes.byref(ms)))
        return float(ms.value)

    def last_kernel_ms(self):
        """Device time of the channelize kernel (K4) of the last process() call."""
        ms = ctypes.c_float()
        c = self.ctx
        c._check(c.lib.mkid_chan_last_kernel_ms(c.h, self.h, ctypes.byref(ms)))
        return float(ms.value)

    def detect(self, phase, t_abs0=0, t_next=None):
        """K5 alone on int16 phase rows [B][rows][256]; resolves rows [M, rows-W-1)."""
        ph = np.ascontiguousarray(phase, dtype=np.int16)
        assert ph.shape[0] == self.n_boards and ph.shape[2] == N_CH
        rows = ph.shape[1]
        tn = np.zeros((self.n_boards, N_CH), dtype=np.int64) if t_next is None else np.ascontiguousarray(t_next, np.int64)
        cap = N_CH * (rows // self.L + 2) + 4
        words = np.empty((self.n_boards, cap), dtype=np.uint64)
        n_words = np.zeros(self.n_boards, dtype=np.int32)
        c = self.ctx
        c._check(c.lib.mkid_chan_detect(c.h, self.h, _lib.ptr(ph), rows, int(t_abs0), _lib.ptr(tn), _lib.ptr(words), cap,
                                        _lib.ptr(n_words)))
        return [words[b, :n_words[b]].copy() for b in range(self.n_boards)], tn

    def close(self):
        if getattr(self, 'h', None):
            self.ctx.lib.mkid_chan_destroy(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def synth_adc(n_boards, n, tone_bin, tone_amp=None, tone_phase=None, n_lut=2 ** 19, full_scale=1800.0, noise_lsb=8.0,
              pulse_rate=1000.0, tau_us=30.0, deg=(20.0, 120.0), seed=42, t_abs0_us=0, out=None, ctx=None):
    """Synthetic ADC streams int16 [n_boards][n][2] generated on the GPU:
    tone_bin [n_boards][n_tones] fine bins (f = bin*fs/n_lut, taken mod n_lut)."""
    ctx = ctx or _lib.default_context()
    tb = np.ascontiguousarray(np.asarray(tone_bin).reshape(n_boards, -1) % n_lut, dtype=np.int32)
    nt = tb.shape[1]
    ta = np.ones((n_boards, nt), np.float32) if tone_amp is None else np.ascontiguousarray(tone_amp, np.float32).reshape(n_boards, nt)
    tp = np.zeros((n_boards, nt), np.float32) if tone_phase is None else np.ascontiguousarray(tone_phase, np.float32).reshape(n_boards, nt)
    prm = _lib.SynthParams(nt, int(n_lut), float(full_scale), float(noise_lsb), float(pulse_rate), float(tau_us),
                           float(deg[0]), float(deg[1]), int(seed))
    ret = out if out is not None else np.empty((n_boards, n, 2), dtype=np.int16)
    ctx._check(ctx.lib.mkid_synth_adc(ctx.h, ctypes.byref(prm), n_boards, _lib.ptr(tb), _lib.ptr(ta), _lib.ptr(tp),
                                      int(n), int(t_abs0_us), _lib.ptr(ret)))
    return ret
