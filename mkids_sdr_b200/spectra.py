"""Per-pixel spectra products of the dashboard's image_Worker on the GPU
(DataReadout/ReadoutControls/ArconsDashboard.py:1282-1504): per-bin medians, sky subtraction, photon counts,
mean energy / wavelength per pixel, SNR of a pixel selection.  Input: the [n_pix][10] spectrum histogram the
photon decoder accumulates (`PhotonDecoder(..., hist_field='peak', n_bins=10, bin_lut=...)`, i.e. data.bin)."""
import math

import numpy as np

from . import _lib

h = 4.13567E-15      # [eV*s]  ArconsDashboard.py:62
c = 3.0E17           # [nm/s]  ArconsDashboard.py:61


class ImageWorker:
    """Headless twin of image_Worker: same attribute names (Emin, Emax, bintype, sky_subtraction, spectrum_pixel,
    medians, pc, me, SNR, integrated_SNR, E0..E9)."""

    def __init__(self, nxpix=44, nypix=46, Emin=0.92, Emax=3.18, ctx=None):
        self.ctx = ctx or _lib.default_context()
        self.nxpix, self.nypix, self.Emin, self.Emax = int(nxpix), int(nypix), Emin, Emax
        self.total_pix = self.nxpix * self.nypix
        self.sky_subtraction = False
        self.spectrum_pixel = []
        self.bintype = 'wavelength'
        self.setup_thread()

    def setup_thread(self):                                   # :1299-1322
        self.SNR = [0] * 10
        if self.bintype == 'energy':
            self.binmin, self.binmax = self.Emin, self.Emax
        else:
            self.binmax = h * c / self.Emin
            self.binmin = h * c / self.Emax
        self.dE = (self.binmax - self.binmin) / 10.
        E = [self.binmin + self.dE / 2.]
        for _ in range(9):
            E.append(E[-1] + self.dE)
        self.E = E
        for i, e in enumerate(E):
            setattr(self, 'E%d' % i, e)

    def run(self, darray):
        """darray: u32 [total_pix][10], host array or device buffer (e.g. PhotonDecoder.hist_dev)."""
        n_pix = self.total_pix
        med = np.empty(10, np.float64); pc = np.empty(n_pix, np.int64); me = np.empty(n_pix, np.float64)
        E = np.array(self.E, dtype=np.float64)
        cx = self.ctx
        d = np.ascontiguousarray(darray, dtype=np.uint32) if isinstance(darray, np.ndarray) else darray
        cx._check(cx.lib.mkid_spectra_products(cx.h, _lib.ptr(d), n_pix, _lib.ptr(E), h * c,
                                               1 if self.bintype == 'wavelength' else 0, 1 if self.sky_subtraction else 0,
                                               _lib.ptr(med), _lib.ptr(pc), _lib.ptr(me)))
        cx.sync()
        self.medians, self.pc, self.me = [float(m) for m in med], pc, me
        if len(self.spectrum_pixel):
            sel = np.asarray(self.spectrum_pixel, dtype=np.int64)
            host = d if isinstance(d, np.ndarray) else None
            if host is None:
                host = np.empty((n_pix, 10), np.uint32)
                cx._check(cx.lib.mkid_memcpy(cx.h, _lib.ptr(host), _lib.ptr(d), host.nbytes)); cx.sync()
            C = host.reshape(n_pix, 10)[sel].astype(np.int64)
            if self.sky_subtraction:
                C = C - np.array([int(m) for m in self.medians], dtype=np.int64)
            self.calculate_SNR([int(v) for v in C.sum(axis=0)], self.medians, len(sel))
        return self.pc, self.me

    def calculate_SNR(self, totalcounts, medians, npix):       # :1364-1384
        total_signal = 0
        total_n = 0
        self.SNR = [0] * 10
        for i in range(len(medians)):
            if self.sky_subtraction:
                signal = totalcounts[i]
            else:
                signal = totalcounts[i] - npix * medians[i]
            noise = npix * medians[i]
            if signal < 0:
                signal = 0
            if noise == 0:
                noise = 1
            total_signal += signal
            total_n += noise
            self.SNR[i] = signal / (math.sqrt(noise))
        self.totalcounts = totalcounts
        self.integrated_SNR = total_signal / (math.sqrt(total_n))
