"""PINNED: bit-identical to the outputs of the reference's own image_Worker methods (ArconsDashboard.py:1282-1384), executed
in the dev container by tests/golden/make_golden_analysis.py -> tests/golden/analysis_golden.npz
(tests/test_oracle_golden.py::test_spectra_oracle_matches_reference_run).

Oracle: per-pixel spectra products of the dashboard's image_Worker (TEST INFRASTRUCTURE).

Restates DataReadout/ReadoutControls/ArconsDashboard.py:1282-1504 without Qt / files:
  * bin centres E0..E9              setup_thread :1299-1322
  * per-bin medians over the pixels :1338-1340 (numpy.median)
  * sky subtraction  x - int(med)   subtract_sky :1338-1355
  * pc[p] = C0[p] + ... + C9[p]     run :1453-1455
  * mean energy / wavelength        calc_mean_energy :1356-1363
  * SNR of a pixel selection        calculate_SNR :1364-1384
Plain Python loops in the reference's own operation order.
"""
import math

import numpy as np

H_PLANCK = 4.13567E-15      # [eV*s]  ArconsDashboard.py:62
C_LIGHT = 3.0E17            # [nm/s]  ArconsDashboard.py:61


def bin_centres(Emin=0.92, Emax=3.18, bintype='wavelength', h=H_PLANCK, c=C_LIGHT):
    if bintype == 'energy':
        binmin, binmax = Emin, Emax
    else:
        binmax = h * c / Emin
        binmin = h * c / Emax
    dE = (binmax - binmin) / 10.
    E = [binmin + dE / 2.]
    for _ in range(9):
        E.append(E[-1] + dE)
    return E, binmin, binmax, dE


def image_worker(darray, Emin=0.92, Emax=3.18, bintype='wavelength', sky_subtraction=False, spectrum_pixel=(),
                 h=H_PLANCK, c=C_LIGHT):
    """darray: int [total_pix][10] (data.bin).  Returns dict(medians, pc, me, totalcounts, SNR, integrated_SNR)."""
    darray = np.asarray(darray)
    total_pix = darray.shape[0]
    E, _, _, _ = bin_centres(Emin, Emax, bintype, h, c)
    C = [[int(v) for v in darray[:, i]] for i in range(10)]
    medians = [float(np.median(darray[:, i])) for i in range(10)]
    if sky_subtraction:
        for i in range(10):
            C[i] = [x - int(medians[i]) for x in C[i]]
    pc = [sum(C[i][m] for i in range(10)) for m in range(total_pix)]
    me = [0.0] * total_pix
    for p in range(total_pix):
        num = C[0][p] * E[0]
        for i in range(1, 10):
            num = num + C[i][p] * E[i]
        if pc[p] != 0:
            me[p] = num / pc[p]
        else:                                   # numpy int division by zero in the reference: inf / nan
            me[p] = math.copysign(math.inf, num) if num != 0 else math.nan
    if bintype == 'wavelength':
        me = [(h * c / e) if e != 0 else math.inf for e in me]
    out = dict(medians=medians, pc=np.array(pc, dtype=np.int64), me=np.array(me, dtype=np.float64), E=E)
    if len(spectrum_pixel):
        totalcounts = [0] * 10
        for p in spectrum_pixel:
            for i in range(10):
                totalcounts[i] += C[i][p]
        npix = len(spectrum_pixel)
        total_signal = 0
        total_n = 0
        SNR = [0] * 10
        for i in range(10):
            if sky_subtraction:
                signal = totalcounts[i]
                noise = npix * medians[i]
            else:
                signal = totalcounts[i] - npix * medians[i]
                noise = npix * medians[i]
            if signal < 0:
                signal = 0
            if noise == 0:
                noise = 1
            total_signal += signal
            total_n += noise
            SNR[i] = signal / (math.sqrt(noise))
        out.update(totalcounts=totalcounts, SNR=SNR, integrated_SNR=total_signal / (math.sqrt(total_n)))
    return out
