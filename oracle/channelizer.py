"""Oracle: software model of the firmware channelizer + pulse detection (TEST INFRASTRUCTURE).

PARITY UNPINNED.  The FPGA data plane (PFB/FFT-512 -> bin select -> DDS mix -> 26-tap FIR
-> /2 -> centre subtract -> atan2 -> trigger -> 64-bit photon words) is NOT in the reference
repository (README.md:28-31 describes a Firmware/ directory that does not exist,
.MISSING_LARGE_BLOBS lists the stripped .bof files).  This float64 NumPy model, written
from the control-plane facts the reference does hold, IS the parity definition for those
stages.  What pins each stage:

  stage                    pinned by (reference file:line)
  ------------------------ -------------------------------------------------------------
  FFT length 512, 2x over- fft_len=2**9, channel rate sampleRate/fft_len*2 = 2 MS/s
  sampled (hop 256)          ChannelizerControls/ROACH_Setup.py:507,525,535
  bin per channel          select_bins ROACH_Setup.py:534-550 (bins >= 256 = negative freq.)
  DDS LUT + layout/lag     define_DDS_LUT ROACH_Setup.py:506-532, ch_shift=154 (:508)
  FIR 26 taps, 12-bit      loadFIRcoeffs ROACH_Pulses.py:59-111 (c = int(tap*2047))
  output rate 1 MS/s       timestamps in us ROACH_Pulses.py:855
  centre subtract, /8      loadIQcenters ROACH_Pulses.py:948-956
  phase = atan2(Q,I)       readouttesterIQ.py:82-83 ; raw = phase*2**13 (Fix16_13)
                           ROACH_Pulses.py:374-378
  threshold (raw, <0)      loadThresholds ROACH_Pulses.py:259-288 ("threshold adjusting firmware")
  trigger rule             rolling mean of the previous M samples, hold-off L
                           pulse_triggering_v2.py:104-174
  photon word layout       ROACH_Pulses.py:805-811 ; Fix12_9 offset binary Utils/bin.py:5-11
  parabola peak            Utils/bin.py:12-16 peakfit
  end-of-second word       PacketMaster.c:329-333

Model choices the reference does not determine (stated, and identical in the CUDA path):
  * PFB prototype: P=4 taps/branch, Hamming-windowed sinc, sum(h)=1, stored as float32.
  * frame f is causal: it uses samples 256 (f+1) - 512 P ... 256 (f+1) - 1; channel sample
    z_c[f] = X_f[bin_c] * (-1)**(bin_c*(f+1))  (removes the half-frame hop phase of odd bins).
  * y = z * conj(dds) / 32767 ; FIR is applied as a CORRELATION (tap 0 on the oldest of the
    26 samples: the shipped matched_30us.txt decays from tap 0, and `lpf = lpf[::-1]` is
    commented out at ROACH_Pulses.py:78,85): w[t] = sum_k c_k y[2t+1-25+k] / 2047.
  * phase raw = rint(atan2(Im w - 8 Qc, Re w - 8 Ic) * 2**13) as int16.
  * trigger in the integer domain on raw values (exact): candidate iff
    M*raw[t] - sum_{k=1..M} raw[t-k] < M*thr ; accepted iff t >= t_next ; t_next = t + L.
  * peak = first minimum of raw over [t, t+W) ; parabola through its neighbours (peakfit,
    float64) ; 12-bit fields = (trunc(v/16) + 2048) & 0xfff (Fix12_9 offset binary of v/2**13
    rad, truncation toward zero as castBin 'Truncate') ; baseline v = sum/M.
  * timestamp = absolute output index mod 1e6 (us), words ordered by (time, channel), one
    all-ones word at every second boundary.
  * host-link format of the ADC stream (no counterpart in the reference: its ADC feeds the FPGA directly): either
    int16 [n][2] or, for a 12-bit ADC, 3 bytes per complex sample = little-endian 24-bit group
    I[11:0] | Q[11:0] << 12, two's complement (adc_pack12 / adc_unpack12 below).
"""
import numpy as np

NFFT = 512
HOP = 256
FIR_TAPS = 26
T_START = 64                 # first output index at which a trigger may fire
CH_SHIFT = 154


def pfb_window(P=4, nfft=NFFT):
    m = np.arange(nfft * P, dtype=np.float64)
    x = (m - (nfft * P - 1) / 2.0) / nfft
    h = np.sinc(x) * np.hamming(nfft * P)
    h = h / h.sum()
    return h.astype(np.float32)


class ChanConfig:
    """Channelizer parameters of one board (all from the reference's control plane)."""

    def __init__(self, bins, I_dds, Q_dds, fir_int, centers_i=None, centers_q=None, thresholds=None,
                 zero_ch=None, M=20, L=1000, W=32, P=4, n_ch=256):
        self.n_ch = n_ch
        self.bins = np.asarray(bins, dtype=np.int64) % NFFT
        self.I_dds = np.asarray(I_dds, dtype=np.int64)
        self.Q_dds = np.asarray(Q_dds, dtype=np.int64)
        self.N_lut = len(self.I_dds)
        self.Ld = self.N_lut // 256                      # DDS samples per channel
        self.fir_int = np.asarray(fir_int, dtype=np.int64)
        assert self.fir_int.size == FIR_TAPS
        self.centers_i = np.zeros(n_ch, np.int64) if centers_i is None else np.asarray(centers_i, np.int64)
        self.centers_q = np.zeros(n_ch, np.int64) if centers_q is None else np.asarray(centers_q, np.int64)
        self.thresholds = np.full(n_ch, -3000, np.int64) if thresholds is None else np.asarray(thresholds, np.int64)
        self.zero_ch = np.zeros(n_ch, bool) if zero_ch is None else np.asarray(zero_ch, bool)
        self.M, self.L, self.W, self.P = int(M), int(L), int(W), int(P)
        self.h = pfb_window(P)

    def dds_table(self):
        """complex [Ld][n_ch]: DDS sample t of channel m = lut[(t//2)*512 + 2*((m+154)%256) + t%2]
        (layout of define_DDS_LUT, ROACH_Setup.py:526-530)."""
        t = np.arange(self.Ld)
        m = np.arange(self.n_ch)
        idx = (t[:, None] // 2) * 512 + 2 * ((m[None, :] + CH_SHIFT) % 256) + (t[:, None] % 2)
        return self.I_dds[idx] + 1j * self.Q_dds[idx]

    @property
    def halo_frames(self):
        """frames of input history one output block needs before its first frame"""
        return FIR_TAPS - 1 + 2 * self.M


def channelize_phase(iq, cfg, f0=0, history=None, return_w=False):
    """iq: int16 [n][2] (n multiple of 512) new samples of one board; f0 = absolute index of the
    first new frame; history: complex or int16[h][2] samples preceding iq (zeros if None).
    Returns (phase_rad float64 [T][C], raw int16 [T][C]) for the T = n/512 new outputs; outputs
    whose FIR window reaches before the start of the stream see zeros there."""
    iq = np.asarray(iq)
    n = iq.shape[0]
    assert n % (2 * HOP) == 0
    x_new = iq[:, 0].astype(np.float64) + 1j * iq[:, 1].astype(np.float64)
    P = cfg.P
    need = (FIR_TAPS - 2) * HOP + NFFT * P
    if history is None:
        hist = np.zeros(need, dtype=np.complex128)
    else:
        history = np.asarray(history)
        hist = history[:, 0].astype(np.float64) + 1j * history[:, 1].astype(np.float64)
        if hist.size < need:
            hist = np.concatenate([np.zeros(need - hist.size, np.complex128), hist])
        hist = hist[-need:]
    x = np.concatenate([hist, x_new])
    F = n // HOP
    n_fr = F + FIR_TAPS - 1                                  # frames -(25) .. F-1
    h = cfg.h.astype(np.float64)
    # frame j (j=0 is local frame -25) starts at x index j*HOP and spans NFFT*P samples
    U = np.zeros((n_fr, NFFT), dtype=np.complex128)
    for p in range(P):
        seg = np.lib.stride_tricks.sliding_window_view(x[p * NFFT:], NFFT)[::HOP][:n_fr]
        U += seg * h[None, p * NFFT:(p + 1) * NFFT]
    X = np.fft.fft(U, axis=1)[:, cfg.bins]                   # [n_fr][C]
    f_abs = f0 - (FIR_TAPS - 1) + np.arange(n_fr)
    sign = 1.0 - 2.0 * ((cfg.bins[None, :] * (f_abs[:, None] + 1)) & 1)
    z = X * sign
    dds = cfg.dds_table()
    d = dds[f_abs % cfg.Ld, :]
    y = z * np.conj(d) / 32767.0
    y[f_abs < 0, :] = 0.0                                    # before the start of the stream
    c = cfg.fir_int.astype(np.float64)
    T = F // 2
    w = np.zeros((T, cfg.n_ch), dtype=np.complex128)
    for k in range(FIR_TAPS):
        # local frame 2t+1-25+k  -> row index (2t+1-25+k) + 25 = 2t+1+k
        w += c[k] * y[1 + k:1 + k + 2 * T:2, :]
    w = w / 2047.0
    w[:, cfg.zero_ch] = 0.0
    a = w.real - 8.0 * cfg.centers_i[None, :]
    b = w.imag - 8.0 * cfg.centers_q[None, :]
    phase = np.arctan2(b, a)
    raw = np.rint(phase * 8192.0).astype(np.int16)
    if return_w:
        return phase, raw, w
    return phase, raw


# ---------------------------------------------------------------- integer detection / emission
def _code12(v_trunc16):
    return (int(v_trunc16) + 2048) & 0xFFF


def _trunc_div(a, b):
    """C integer division (toward zero)."""
    q = abs(int(a)) // int(b)
    return q if a >= 0 else -q


def peakfit_f64(y1, y2, y3):
    """Utils/bin.py:12-16 on float64."""
    y1, y2, y3 = float(y1), float(y2), float(y3)
    d = y3 + y1 - 2 * y2
    if d == 0:
        return y2
    return y2 - 0.125 * ((y3 - y1) ** 2) / d


def detect_emit(raw, cfg, t_abs0, t_next, n_resolve):
    """raw: int [rows][C] phase rows, row r = absolute output index t_abs0 + r.
    Triggers are resolved for rows r in [r_lo, r_lo + n_resolve) where r_lo = max(cfg.M, ...)
    is chosen by the caller through t_abs0 so that rows r-M..r+W exist.  Here: resolves rows
    r in [cfg.M, cfg.M + n_resolve).  t_next: int array [C] (updated in place).
    Returns the board's word list (python ints) in emission order."""
    raw = np.asarray(raw, dtype=np.int64)
    M, L, W = cfg.M, cfg.L, cfg.W
    C = raw.shape[1]
    r_lo, r_hi = M, M + n_resolve
    assert r_hi + W < raw.shape[0] + 1
    cs = np.concatenate([np.zeros((1, C), np.int64), np.cumsum(raw, axis=0)])
    events = []
    for c in range(C):
        S = cs[r_lo:r_hi, c] - cs[r_lo - M:r_hi - M, c]        # sum of raw[r-M..r-1]
        cand = (M * raw[r_lo:r_hi, c] - S) < M * int(cfg.thresholds[c])
        idx = np.nonzero(cand)[0]
        for i in idx:
            r = r_lo + int(i)
            t = t_abs0 + r
            if t < T_START or t < t_next[c]:
                continue
            t_next[c] = t + L
            win = raw[r:r + W, c]
            tp = r + int(np.argmin(win))
            y4 = peakfit_f64(raw[tp - 1, c], raw[tp, c], raw[tp + 1, c])
            peak = _code12(int(y4 / 16.0))
            p1 = _code12(_trunc_div(raw[tp, c], 16))
            base = _code12(_trunc_div(S[i], 16 * M))
            word = (c << 56) | (peak << 44) | (p1 << 32) | (base << 20) | (t % 10 ** 6)
            events.append((t, c, word))
    # second boundaries inside the resolved time range
    ta, tb = t_abs0 + r_lo, t_abs0 + r_hi
    B = ((ta + 10 ** 6 - 1) // 10 ** 6) * 10 ** 6
    while B < tb:
        if B > 0:
            events.append((B, -1, 0xFFFFFFFFFFFFFFFF))
        B += 10 ** 6
    events.sort(key=lambda e: (e[0], e[1]))
    return [e[2] for e in events]


# ---------------------------------------------------------------- synthetic board input (small sizes)
def synth_board(n, tone_bins_fine, N_lut, amps=None, seed=42, pulse_rate=1000.0, tau_us=30.0,
                pulse_deg=(20.0, 120.0), noise_lsb=8.0, fs=512e6, full_scale=1800.0, phases=None):
    """int16 [n][2] ADC stream: comb of tones at fine bins k (f = k*fs/N_lut), each phase-modulated
    by exponential pulses, plus white noise, scaled to a 12-bit range."""
    rng = np.random.default_rng(seed)
    T = len(tone_bins_fine)
    amps = np.ones(T) if amps is None else np.asarray(amps, float)
    phases = rng.uniform(0, 2 * np.pi, T) if phases is None else phases
    t = np.arange(n, dtype=np.float64)
    n_us = int(np.ceil(n / 512.0)) + 1
    x = np.zeros(n, dtype=np.complex128)
    decay = np.exp(-1.0 / tau_us)
    pulses = []
    for i, k in enumerate(tone_bins_fine):
        theta_us = np.zeros(n_us)
        npul = rng.poisson(pulse_rate * n_us * 1e-6)
        t0s = np.sort(rng.integers(0, n_us, npul))
        amps_deg = rng.uniform(pulse_deg[0], pulse_deg[1], npul)
        imp = np.zeros(n_us)
        np.add.at(imp, t0s, -np.deg2rad(amps_deg))
        acc = 0.0
        for u in range(n_us):
            acc = acc * decay + imp[u]
            theta_us[u] = acc
        pulses.append((t0s, amps_deg))
        theta = np.repeat(theta_us, 512)[:n]
        x += amps[i] * np.exp(1j * (2 * np.pi * ((int(k) * t) % N_lut) / N_lut + phases[i] + theta))
    scale = full_scale / np.abs(x).max()
    x = x * scale
    x = x + rng.normal(0, noise_lsb, n) + 1j * rng.normal(0, noise_lsb, n)
    iq = np.empty((n, 2), dtype=np.int16)
    iq[:, 0] = np.clip(np.rint(x.real), -2047, 2047)
    iq[:, 1] = np.clip(np.rint(x.imag), -2047, 2047)
    return iq, pulses


def adc_pack12(iq):
    """int16 [..., n, 2] (values in [-2048, 2047]) -> uint8 [..., 3 n]: 24-bit little-endian groups I | Q << 12."""
    iq = np.asarray(iq).astype(np.int64)
    if iq.min() < -2048 or iq.max() > 2047:
        raise ValueError('sample does not fit 12 bits')
    g = (iq[..., 0] & 0xFFF) | ((iq[..., 1] & 0xFFF) << 12)
    out = np.empty(g.shape + (3,), dtype=np.uint8)
    out[..., 0] = g & 0xFF
    out[..., 1] = (g >> 8) & 0xFF
    out[..., 2] = (g >> 16) & 0xFF
    return out.reshape(g.shape[:-1] + (3 * g.shape[-1],))


def adc_unpack12(packed):
    """uint8 [..., 3 n] -> int16 [..., n, 2] (sign-extended 12-bit I and Q)."""
    b = np.asarray(packed, dtype=np.uint8).astype(np.int64)
    b = b.reshape(b.shape[:-1] + (b.shape[-1] // 3, 3))
    g = b[..., 0] | (b[..., 1] << 8) | (b[..., 2] << 16)
    i = g & 0xFFF
    q = (g >> 12) & 0xFFF
    i = np.where(i >= 2048, i - 4096, i)
    q = np.where(q >= 2048, q - 4096, q)
    return np.stack([i, q], axis=-1).astype(np.int16)
