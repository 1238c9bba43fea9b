"""Oracle: DAC frequency-comb / DDS LUT synthesis (TEST INFRASTRUCTURE).

NumPy restatement of DataReadout/ChannelizerControls/ROACH_Setup.py:416-578 and
its multi-tone twin ROACH_Setup_DAC.py:396-558, keeping the reference's float64
operation order (so the int16 results are the reference's, truncation flips
included) and its Python-2 semantics (round half away from zero, `/` on ints).

Pinned: `freq_comb_lut('yes',[412e6],512e6,7812.5,[1.0])` reproduces the
reference's own dump ChannelizerControls/dac.npy.npz bit-exactly
(tests/test_oracle_golden.py).
"""
import struct

import numpy as np

from .fixed import py2_round

AMP_FULL_SCALE = 2 ** 15 - 1          # ROACH_Setup.py:420
SCALE_FUDGE = 1.1                     # ROACH_Setup.py:453
FFT_LEN = 2 ** 9                      # ROACH_Setup.py:507,535
CH_SHIFT = 154                        # ROACH_Setup.py:508 (= DDS_LAG setEnvironment.sh:24)


def random_phases(n):
    """numpy.random.seed(1000) then one uniform(0,2pi) per tone, in tone order
    (ROACH_Setup.py:426-429)."""
    rs = np.random.RandomState(1000)
    return np.array([rs.uniform(0, 2 * np.pi) for _ in range(n)])


def freq_comb_lut_literal(echo, freq, sampleRate, resolution, amplitude=None, phase=None,
                          random_phase='yes', offset=0):
    """Line-by-line ROACH_Setup.py:416-475 (python list comprehensions; slow).
    Returns (I, Q, scale_factor)."""
    N_freqs = len(freq)
    amplitude = [1.] * 256 if amplitude is None else amplitude
    phase = [0.] * 256 if phase is None else list(phase)
    size = int(sampleRate / resolution)
    I, Q = np.array([0.] * size), np.array([0.] * size)
    np.random.seed(1000)
    for n in range(N_freqs):
        if random_phase == 'yes':
            phase[n] = np.random.uniform(0, 2 * np.pi)
        x = [2 * np.pi * freq[n] * (t + offset) / sampleRate + phase[n] for t in range(size)]
        y = [2 * np.pi * freq[n] * t / sampleRate + phase[n] for t in range(size)]
        single_I = amplitude[n] * np.cos(x)
        single_Q = amplitude[n] * np.sin(y)
        I = I + single_I
        Q = Q + single_Q
    a = np.array([abs(I).max(), abs(Q).max()])
    scale_factor = a.max()
    if echo == 'yes':
        scale_factor = SCALE_FUDGE * scale_factor
    I = np.array([int(i * AMP_FULL_SCALE / scale_factor) for i in I])
    Q = np.array([int(q * AMP_FULL_SCALE / scale_factor) for q in Q])
    return I, Q, scale_factor


def comb_float(freq, sampleRate, size, amplitude, phase, offset=0):
    """The float64 I/Q sums of ROACH_Setup.py:439-446, vectorised with the SAME
    per-element operation order: ((2*pi*f)*(t+offset))/fs + phi, a*cos(.),
    sequential accumulation over tones in list order."""
    t = np.arange(size, dtype=np.float64)
    I = np.zeros(size)
    Q = np.zeros(size)
    for n in range(len(freq)):
        w = 2 * np.pi * float(freq[n])
        x = w * (t + offset) / sampleRate + phase[n]
        y = w * t / sampleRate + phase[n] if offset != 0 else x
        I = I + amplitude[n] * np.cos(x)
        Q = Q + amplitude[n] * np.sin(y)
    return I, Q


def freq_comb_lut(echo, freq, sampleRate, resolution, amplitude=None, phase=None,
                  random_phase='yes', offset=0):
    """Vectorised ROACH_Setup.py:416-475.  Returns (I, Q, scale_factor, phases_used)."""
    N_freqs = len(freq)
    amplitude = [1.] * max(256, N_freqs) if amplitude is None else amplitude
    phase = [0.] * max(256, N_freqs) if phase is None else list(phase)
    size = int(sampleRate / resolution)
    if random_phase == 'yes':
        ph = random_phases(N_freqs)
        phase[:N_freqs] = list(ph)
    I, Q = comb_float(freq, sampleRate, size, amplitude, phase, offset)
    a = np.array([abs(I).max(), abs(Q).max()])
    scale_factor = a.max()
    if echo == 'yes':
        scale_factor = SCALE_FUDGE * scale_factor
    Ii = np.trunc(I * AMP_FULL_SCALE / scale_factor).astype(np.int64)
    Qi = np.trunc(Q * AMP_FULL_SCALE / scale_factor).astype(np.int64)
    return Ii, Qi, float(scale_factor), np.array(phase[:N_freqs], dtype=np.float64)


def dac_freqs_single(dac_freq, lo_freq, sampleRate, freqRes):
    """define_DAC_LUT of ROACH_Setup.py:477-498 (single-tone GUI): mirror about
    LO, +fs if below LO, snap to the freqRes grid with py2 round."""
    freqs = [float(dac_freq)]
    f_base = float(lo_freq)
    freqs = [f_base + (f_base - f) for f in freqs]
    freqs = [f + sampleRate if f < f_base else f for f in freqs]
    return [py2_round((f - f_base) / freqRes) * freqRes for f in freqs]


def dac_freqs_multi(freqs, lo_freq, freqRes, sampleRate=512e6):
    """define_DAC_LUT of ROACH_Setup_DAC.py:457-478 (multi-tone GUI): exactly the single-tone recipe applied to the
    whole list -- mirror about LO (:464-467), +fs if below LO (:473-475), snap to the freqRes grid (py2 round, :478).
    (Pinned by the reference's own method executed in the dev container: tests/golden/refrun_golden.npz.)"""
    f_base = float(lo_freq)
    freqs = [f_base + (f_base - float(f)) for f in freqs]
    freqs = [f + sampleRate if f < f_base else f for f in freqs]
    return [py2_round((f - f_base) / freqRes) * freqRes for f in freqs]


def dac_amplitudes(attens):
    """ROACH_Setup.py:499-501."""
    attens = np.asarray(attens, dtype=np.float64)
    atten_min = attens.min()
    return [10 ** (+(atten_min - a) / 20.) for a in attens]


def select_bins(readout_freqs, sampleRate, freqRes):
    """ROACH_Setup.py:534-550.  Returns (fft_bins, residuals)."""
    bins, residuals = [], []
    for f in readout_freqs:
        fft_bin = int(py2_round(f * FFT_LEN / sampleRate))
        fft_freq = fft_bin * sampleRate / FFT_LEN
        freq_residual = py2_round((f - fft_freq) / freqRes) * freqRes
        residuals.append(freq_residual)
        bins.append(fft_bin)
    return bins, residuals


def dds_freqs(freqs, lo_freq, sampleRate, freqRes):
    """define_DDS_LUT ROACH_Setup.py:510-518: +fs if below LO, snap, pad to 256."""
    f_base = float(lo_freq)
    freqs = [float(f) + sampleRate if float(f) < f_base else float(f) for f in freqs]
    out = [0 for _ in range(256)]
    for n in range(len(freqs)):
        out[n] = py2_round((freqs[n] - f_base) / freqRes) * freqRes
    return out


def define_dds_lut(freq_residuals, sampleRate, freqRes, phase=None):
    """ROACH_Setup.py:520-530: 256 single-tone tables at fs/512*2, scale = max
    (no fudge), scattered to [j*512 + 2*((m+154)%256) + s]."""
    phase = [0.] * 256 if phase is None else phase
    L = int(sampleRate / freqRes)
    I_dds = np.zeros(L, dtype=np.int64)
    Q_dds = np.zeros(L, dtype=np.int64)
    fs2 = sampleRate / FFT_LEN * 2
    scales = []
    for m in range(256):
        I, Q, sc, _ = freq_comb_lut('no', [freq_residuals[m]], fs2, freqRes, [1.], [phase[m]], 'no')
        scales.append(sc)
        half = len(I) // 2
        j = np.arange(half)
        slot = 2 * ((m + CH_SHIFT) % 256)
        I_dds[j * 512 + slot] = I[2 * j]
        I_dds[j * 512 + slot + 1] = I[2 * j + 1]
        Q_dds[j * 512 + slot] = Q[2 * j]
        Q_dds[j * 512 + slot + 1] = Q[2 * j + 1]
    return I_dds, Q_dds, np.array(scales)


def pack_dram(I_dac, Q_dac, I_dds, Q_dds):
    """write_LUTs ROACH_Setup.py:560-569: per sample pair 8 big-endian int16
    q_dds1 q_dds0 q_dac1 q_dac0 i_dds1 i_dds0 i_dac1 i_dac0."""
    n = len(I_dac) // 2
    out = np.empty((n, 8), dtype='>i2')
    I_dac, Q_dac, I_dds, Q_dds = (np.asarray(a) for a in (I_dac, Q_dac, I_dds, Q_dds))
    out[:, 0] = Q_dds[1::2][:n]
    out[:, 1] = Q_dds[0::2][:n]
    out[:, 2] = Q_dac[1::2][:n]
    out[:, 3] = Q_dac[0::2][:n]
    out[:, 4] = I_dds[1::2][:n]
    out[:, 5] = I_dds[0::2][:n]
    out[:, 6] = I_dac[1::2][:n]
    out[:, 7] = I_dac[0::2][:n]
    return out.tobytes()


def pack_dram_literal(I_dac, Q_dac, I_dds, Q_dds):
    """Literal struct.pack loop of ROACH_Setup.py:560-569 (small inputs)."""
    binaryData = b''
    for n in range(len(I_dac) // 2):
        binaryData += (struct.pack('>h', int(Q_dds[2 * n + 1])) + struct.pack('>h', int(Q_dds[2 * n])) +
                       struct.pack('>h', int(Q_dac[2 * n + 1])) + struct.pack('>h', int(Q_dac[2 * n])) +
                       struct.pack('>h', int(I_dds[2 * n + 1])) + struct.pack('>h', int(I_dds[2 * n])) +
                       struct.pack('>h', int(I_dac[2 * n + 1])) + struct.pack('>h', int(I_dac[2 * n])))
    return binaryData
