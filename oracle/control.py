"""Oracle: channelizer control-plane quantisation (TEST INFRASTRUCTURE).

Restates, without Qt/KATCP:
  * loadFIRcoeffs   ROACH_Pulses.py:59-111   (12-bit tap quantisation + register packing)
  * loadIQcenters   ROACH_Pulses.py:948-956  (centre quantisation + packing)
  * findIQcenters   ROACH_Setup.py:621-625
  * loadThresholds  ROACH_Pulses.py:211-299  (histogram-CDF threshold)
  * IQ snapshot decode pulse_triggering_IQ.py:121-147 (App. A.3) and phase model :152
  * phase snapshot decode ROACH_Pulses.py:251-253 / pulse_triggering.py:92-93 (App. A.4)
"""
import struct

import numpy as np

SCALE_TO_ANGLE = 360. / 2 ** 16 * 4 / np.pi        # ROACH_Pulses.py:222 (deg per Fix16_13 LSB)
THRESH_FLOOR = -25736                              # ROACH_Pulses.py:275


def fir_quantise(fir, taps=26):
    """c = int(tap*(2**11-1)) truncation toward zero (ROACH_Pulses.py:69,88-89)."""
    lpf = np.array(fir, dtype=np.float64) * (2 ** 11 - 1)
    return [int(lpf[k]) for k in range(taps)]


def fir_registers(fir, taps=26):
    """13 registers FIR_b{2n}b{2n+1} = int(binary_repr(c[2n+1],12)+binary_repr(c[2n],12),2)
    (ROACH_Pulses.py:87-93); returns list of (name, u32 value, packed '>l' bytes)."""
    c = fir_quantise(fir, taps)
    regs = []
    for n in range(taps // 2):
        coeff0 = np.binary_repr(c[2 * n], 12)
        coeff1 = np.binary_repr(c[2 * n + 1], 12)
        coeffs = int(coeff1 + coeff0, 2)
        regs.append(('FIR_b' + str(2 * n) + 'b' + str(2 * n + 1), coeffs, struct.pack('>l', coeffs)))
    return regs


def iq_center_word(center):
    """(I_c<<16)+(Q_c<<0) with I_c=int(re/2**3), Q_c=int(im/2**3) (ROACH_Pulses.py:949-953).
    Python ints: a negative Q_c borrows from the I field exactly as the reference does."""
    I_c = int(center.real / 2 ** 3)
    Q_c = int(center.imag / 2 ** 3)
    return (I_c << 16) + (Q_c << 0), I_c, Q_c


def find_iq_center(I, Q):
    """ROACH_Setup.py:621-625: midpoint of min/max."""
    I_c = (np.max(I) + np.min(I)) / 2.
    Q_c = (np.max(Q) + np.min(Q)) / 2.
    return complex(I_c, Q_c)


def threshold_from_phase(phase_raw, Nsigma=2.5):
    """ROACH_Pulses.py:259-277 on raw Fix16_13 samples. Returns (threshold_raw int,
    med (bin edge, float), thresh-edge p5 (float))."""
    phase = np.asarray(phase_raw)
    n, bins = np.histogram(phase, bins=100)
    n = np.array(n, dtype='float32') / np.sum(n)
    tot = np.zeros(len(bins))
    for i in range(len(bins)):
        tot[i] = np.sum(n[:i])
    med = bins[(np.abs(tot - 0.5)).argmin()]
    thresh = bins[(np.abs(tot - 0.05)).argmin()]
    threshold = int(-Nsigma * abs(med - thresh))
    if threshold < THRESH_FLOOR:
        threshold = THRESH_FLOOR
    return threshold, med, thresh


def twos_comp(val, bits):
    """pulse_triggering.py:22-26."""
    if (val & (1 << (bits - 1))) != 0:
        val = val - (1 << bits)
    return val


def decode_iq_snapshot(buf):
    """pulse_triggering_IQ.py:121-147 without the hex-string detour:
    per 16 bytes two samples, I = low 16 bits of the first 20-bit field, Q = low 16
    of the second.  Returns (Iraw, Qraw) int arrays of length len(buf)//8."""
    b = np.frombuffer(buf, dtype=np.uint8).reshape(-1, 16).astype(np.int64)
    I0 = ((b[:, 6] & 0xF) << 12) | (b[:, 7] << 4) | (b[:, 8] >> 4)
    Q0 = (b[:, 9] << 8) | b[:, 10]
    I1 = ((b[:, 11] & 0xF) << 12) | (b[:, 12] << 4) | (b[:, 13] >> 4)
    Q1 = (b[:, 14] << 8) | b[:, 15]

    def tc(v):
        return np.where(v & 0x8000, v - 0x10000, v)
    Iraw = np.stack([tc(I0), tc(I1)], axis=1).reshape(-1)
    Qraw = np.stack([tc(Q0), tc(Q1)], axis=1).reshape(-1)
    return Iraw, Qraw


def decode_iq_snapshot_literal(buf):
    """Literal hex-string nibble surgery of pulse_triggering_IQ.py:121-147."""
    hx = ["0x{:02x}".format(c) for c in bytearray(buf)]
    Iraw, Qraw = [], []
    for k in range(len(buf) // 16):
        I0 = hx[6 + 16 * k][3] + hx[7 + 16 * k][2:4] + hx[8 + 16 * k][2]
        Iraw.append(twos_comp(int(I0, 16), 16))
        I1 = hx[11 + 16 * k][3] + hx[12 + 16 * k][2:4] + hx[13 + 16 * k][2]
        Iraw.append(twos_comp(int(I1, 16), 16))
        Q0 = hx[9 + 16 * k][2:4] + hx[10 + 16 * k][2:4]
        Qraw.append(twos_comp(int(Q0, 16), 16))
        Q1 = hx[14 + 16 * k][2:4] + hx[15 + 16 * k][2:4]
        Qraw.append(twos_comp(int(Q1, 16), 16))
    return np.array(Iraw), np.array(Qraw)


def phase_deg_from_iq(Iraw, Qraw, Ic=0, Qc=0):
    """pulse_triggering_IQ.py:152."""
    return -360 * (np.arctan2((np.asarray(Qraw) - Qc), (np.asarray(Iraw) - Ic))) / (2 * np.pi)


def decode_phase_snapshot(buf, two_per_word=True):
    """Phase BRAM (App. A.4).  two_per_word: ROACH_Pulses.py:251-253 (first sample =
    bytes [2:4], second = bytes [0:2]); else one int16 in bytes [2:4]
    (pulse_triggering.py:92-93)."""
    a = np.frombuffer(buf, dtype='>i2').reshape(-1, 2).astype(np.int64)
    if two_per_word:
        return np.stack([a[:, 1], a[:, 0]], axis=1).reshape(-1)
    return a[:, 1].copy()
