"""The quantiser of K1 flags a sample for exact re-evaluation when its scaled bulk value is within
eps = 8 * sigma of an integer, sigma = 2.2e-16 * 2 pi N * sqrt(sum a^2) * 32767 / scale (csrc/lut.cu, comb_prep_kernel /
comb_scale_kernel).  This test measures what sigma models - the difference between the reference-order float64 sum
(ROACH_Setup.py:439-440) and the exact value - on BASELINE config 1 and checks the margin."""
import os
import re

import numpy as np

from oracle import lut as olut

FS = 512e6


def test_flag_distance_covers_reference_rounding():
    src = open(os.path.join(os.path.dirname(__file__), '..', 'mkids_sdr_b200', 'csrc', 'lut.cu')).read()
    m = re.search(r'p\.sigma\[\w+\] = ([0-9.e-]+) \* 6\.283185307179586 \* \(double\)p\.N \* sqrt\(ss\);', src)
    assert m, 'sigma model not found in lut.cu'
    coeff = float(m.group(1))
    assert re.search(r'p\.eps\[b\] = fmax\(1e-7, 8\.0 \* p\.sigma\[b\] \* 32767\.0 / sc\);', src)
    N, T = 2 ** 19, 256
    k = np.sort(np.random.default_rng(0).choice(np.arange(-N // 2 + 1, N // 2), T, replace=False))
    f = (k % N) * FS / N
    amps = np.asarray(olut.dac_amplitudes(np.random.default_rng(1).integers(0, 20, T)))
    ph = olut.random_phases(T)
    # the largest arguments (end of the table) have the largest rounding errors
    ts = np.concatenate([np.random.default_rng(5).integers(0, N, 2000), np.arange(N - 2000, N)])
    w = 2 * np.pi * f
    ref = np.zeros(len(ts))
    for n in range(T):
        ref += amps[n] * np.cos((w[n] * ts) / FS + ph[n])                    # reference order, float64
    ld = np.longdouble
    two_pi = ld(8) * np.arctan(ld(1))
    acc = np.zeros(len(ts), dtype=ld)
    kk = (k % N).astype(np.int64)
    for n in range(T):
        acc += ld(amps[n]) * np.cos(two_pi * ((kk[n] * ts) % N).astype(ld) / N + ld(ph[n]))   # exact phase reduction
    scale = 1.1 * max(np.abs(ref).max(), 20.0)                                # ~ the table's scale (30.15)
    dev_lsb = np.abs((ref - acc).astype(np.float64)) * 32767 / scale
    eps = 8 * coeff * 2 * np.pi * N * np.sqrt((amps ** 2).sum()) * 32767 / scale
    assert 1e-6 < eps < 1e-4
    assert dev_lsb.max() * 8 < eps, (dev_lsb.max(), eps)                      # measured: max 4e-6 LSB against eps 4.8e-5
