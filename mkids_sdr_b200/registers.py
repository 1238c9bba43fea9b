"""The register protocol of the channelizer firmware and the control-plane arithmetic around it, as data and
vectorised NumPy -- written from the protocol facts, not from the GUI code.

What the firmware expects (register names as the ROACH designs export them; cited where the reference uses them):
  * per-channel values are written to a VALUE register and latched by a STROBE register that carries the channel index:
    high = (ch << 1) | 1, then low.  Three loaders return to (ch << 1), the centre loader returns to 0
    (ROACH_Pulses.py:93-95, 953-956, 965-967, 286-288).
  * FIR: 26 taps of 12-bit two's complement, two per register, 13 registers FIR_b0b1 .. FIR_b24b25 (ROACH_Pulses.py:59-111).
  * IQ centres: (int(I/8) << 16) + int(Q/8) in one register (ROACH_Pulses.py:948-956).
  * tone placement: fs/512 coarse bins + a residual on the LUT frequency grid (ROACH_Setup.py:534-550); DAC tones are
    mirrored about the LO (ROACH_Setup.py:484-498).
  * Python-2 arithmetic the numbers were produced with: round() = half away from zero, int() = truncation.
"""
import struct

import numpy as np

FFT_LEN = 512
N_CHANNELS = 256
FIR_TAPS = 26
FIR_REGISTERS = tuple('FIR_b%db%d' % (2 * n, 2 * n + 1) for n in range(FIR_TAPS // 2))
THRESHOLD_FLOOR = -25736                      # -pi in Fix16_13 (ROACH_Pulses.py:275)
PHASE_LSB_DEG = 360. / 2 ** 16 * 4 / np.pi    # Fix16_13 raw -> degrees (ROACH_Pulses.py:374-378)

# loader -> (strobe register, what the strobe returns to)
LOADERS = {
    'fir': ('FIR_load_coeff', 'index'),
    'centers': ('conv_phase_load_centers', 'zero'),
    'bins': ('load_bins', 'index'),
    'threshold': ('capture_load_thresh', 'index'),
}


def strobe(roach, loader, ch):
    reg, low = LOADERS[loader]
    roach.write_int(reg, (ch << 1) | 1)
    roach.write_int(reg, (ch << 1) if low == 'index' else 0)


def round_half_away(x):
    """Python-2 round() on an array (or scalar): half away from zero."""
    x = np.asarray(x, dtype=np.float64)
    return np.where(x >= 0, np.floor(x + 0.5), -np.floor(-x + 0.5))


# ---------------------------------------------------------------- tone placement
def baseband_tones(freqs_hz, lo_hz, fs, res, mirror):
    """Frequencies relative to the LO on the `res` grid, in [0, fs): tones below the LO alias to fs - delta.
    mirror=True is the DAC comb (reflected about the LO first), mirror=False the DDS / readout side."""
    f = np.asarray(freqs_hz, dtype=np.float64)
    lo = float(lo_hz)
    if mirror:
        f = lo + (lo - f)
    f = np.where(f < lo, f + fs, f)
    return round_half_away((f - lo) / res) * res


def coarse_fine(freqs, fs, res):
    """(fft_bin int64, residual float64) per tone: nearest fs/512 bin and what is left, on the `res` grid."""
    f = np.asarray(freqs, dtype=np.float64)
    fft_bin = round_half_away(f * FFT_LEN / fs).astype(np.int64)
    resid = round_half_away((f - fft_bin * fs / FFT_LEN) / res) * res
    return fft_bin, resid


def write_bins(roach, fft_bins):
    for ch, b in enumerate(fft_bins):
        roach.write_int('bins', int(b))
        strobe(roach, 'bins', ch)


# ---------------------------------------------------------------- FIR
def fir_taps_int(taps):
    """12-bit tap values: truncation of tap * 2047."""
    return np.trunc(np.asarray(taps, dtype=np.float64) * (2 ** 11 - 1)).astype(np.int64)


def fir_register_words(taps_int):
    """The 13 register values: tap 2n in bits 0..11, tap 2n+1 in bits 12..23 (two's complement fields)."""
    c = np.asarray(taps_int, dtype=np.int64) & 0xFFF
    return (c[1::2] << 12) | c[0::2]


def write_fir(roach, words, channels):
    payload = [struct.pack('>l', int(w)) for w in words]
    for ch in channels:
        for name, data in zip(FIR_REGISTERS, payload):
            roach.write(name, data)
            strobe(roach, 'fir', ch)


# ---------------------------------------------------------------- IQ centres
def center_words(iq_centers):
    """[(I_c, Q_c, register word)] for the 256 centres: components / 8 truncated toward zero."""
    c = np.asarray(iq_centers, dtype=np.complex128)
    i_c = np.trunc(c.real / 2 ** 3).astype(np.int64)
    q_c = np.trunc(c.imag / 2 ** 3).astype(np.int64)
    return i_c, q_c, (i_c << 16) + q_c


def write_centers(roach, words):
    for ch, w in enumerate(words):
        roach.write_int('conv_phase_centers', int(w))
        strobe(roach, 'centers', ch)


# ---------------------------------------------------------------- DAC start / stop
def dac_start(roach):
    """Start the LUT playback; the firmware may need several tries until DRAM_LUT_rd_valid drops (ROACH_Setup.py:881-894)."""
    roach.write_int('startDAC', 1)
    while roach.read_int('DRAM_LUT_rd_valid') != 0:
        roach.write_int('startDAC', 0)
        roach.write_int('startDAC', 1)


def dac_stop(roach):
    roach.write_int('startDAC', 0)


# ---------------------------------------------------------------- phase snapshots / thresholds
def snapshot_phase(roach, ch, steps, words=2 ** 10):
    """`steps` snapshots of `words` 32-bit words of channel ch from snapPhase_bram; every word holds two int16 samples,
    the later one in the high half (ROACH_Pulses.py:241-253).  Returns int16 [steps * words * 2] in time order."""
    parts = []
    for _ in range(steps):
        roach.write_int('ch_we', ch)
        roach.write_int('startSnap', 0)
        roach.write_int('snapPhase_ctrl', 1)
        roach.write_int('snapPhase_ctrl', 0)
        roach.write_int('startSnap', 1)
        parts.append(roach.read('snapPhase_bram', 4 * words))
    a = np.frombuffer(b''.join(parts), dtype='>i2').reshape(-1, 2)
    return np.ascontiguousarray(a[:, ::-1]).reshape(-1).astype(np.int16)


def write_threshold(roach, ch, threshold_raw):
    roach.write_int('capture_threshold', int(threshold_raw))
    strobe(roach, 'threshold', ch)


def custom_threshold_raw(custom_deg):
    """A user threshold in degrees as raw Fix16_13 (truncated by the register write), not below -pi."""
    t = custom_deg / PHASE_LSB_DEG
    return THRESHOLD_FLOOR if t < THRESHOLD_FLOOR else t


# ---------------------------------------------------------------- averaged IQ point of every channel
def snapshot_avg_iq(roach):
    """One accumulator read-out: avgIQ_bram holds 256 big-endian int32 I values, then 256 Q values (ROACH_Setup.py:654-662,
    775-778).  Returns (I int64 [256], Q int64 [256])."""
    roach.write_int('startAccumulator', 0)
    roach.write_int('avgIQ_ctrl', 1)
    roach.write_int('avgIQ_ctrl', 0)
    roach.write_int('startAccumulator', 1)
    a = np.frombuffer(roach.read('avgIQ_bram', 4 * 2 * N_CHANNELS), dtype='>i4').astype(np.int64)
    return a[:N_CHANNELS], a[N_CHANNELS:]
