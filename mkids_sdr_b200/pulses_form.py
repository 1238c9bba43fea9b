"""Headless twin of the "channelizer" GUI's hot-path methods (class AppForm of
DataReadout/ChannelizerControls/ROACH_Pulses.py): FIR / centre / bin / threshold loading and the
photon read-out.  Same method names and attributes; `self.roach` is injectable.
"""
import math
import os
import struct

import numpy

from . import decode as _decode
from .fake_roach import FakeRoach


def _py2_round(x):
    x = float(x)
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


class PulsesForm:
    def __init__(self, roach=None, sampleRate=512e6, N_lut_entries=2 ** 16, ctx=None):
        self.ctx = ctx
        self.roach = roach if roach is not None else FakeRoach()
        self.sampleRate = sampleRate
        self.N_lut_entries = N_lut_entries
        self.freqRes = sampleRate / N_lut_entries            # ROACH_Pulses.py:40
        self.dacStatus = 'off'
        self.dac_freqs = []                                  # textedit_DACfreqs (Hz)
        self.lo_freq = 0.0                                   # textbox_loFreq
        self.lutDir = './'                                   # textbox_lutDir
        self.fir = [0.] * 26                                 # importFIRcoeffs :1088-1103
        self.zeroChannels = [0] * 256
        self.customThresholds = numpy.array([360.0] * 256)   # 360.0 = "none" (:45)
        self.thresholds = numpy.array([0.] * 256)
        self.medians = numpy.array([0.] * 256)
        self.iq_centers = numpy.array([0. + 0j] * 256)
        self.scale_factor = 1.
        self.channel = 0                                     # textbox_channel
        self.seconds = 1                                     # textbox_seconds

    # ------------------------------------------------------------------ a6
    def importFIRcoeffs(self, path):
        """ROACH_Pulses.py:1088-1103."""
        self.fir = list(numpy.loadtxt(path))

    def loadFIRcoeffs(self):
        """ROACH_Pulses.py:59-111: 12-bit taps, pairs packed c[2n+1]<<12 | c[2n] into 13 registers
        per channel; deleted and inactive channels get zero taps.  Sets self.fir_int."""
        N_freqs = len(self.dac_freqs)
        taps = 26
        self.fir_int = [int(v) for v in numpy.array(self.fir) * (2 ** 11 - 1)]
        for ch in range(N_freqs):
            if self.zeroChannels[ch]:
                lpf = numpy.array([0.] * taps) * (2 ** 11 - 1)
            else:
                lpf = numpy.array(self.fir) * (2 ** 11 - 1)
            for n in range(taps // 2):
                coeff0 = numpy.binary_repr(int(lpf[2 * n]), 12)
                coeff1 = numpy.binary_repr(int(lpf[2 * n + 1]), 12)
                coeffs = int(coeff1 + coeff0, 2)
                coeffs_bin = struct.pack('>l', coeffs)
                register_name = 'FIR_b' + str(2 * n) + 'b' + str(2 * n + 1)
                self.roach.write(register_name, coeffs_bin)
                self.roach.write_int('FIR_load_coeff', (ch << 1) + (1 << 0))
                self.roach.write_int('FIR_load_coeff', (ch << 1) + (0 << 0))
        lpf = numpy.array([0.] * taps)
        for ch in range(N_freqs, 256):
            for n in range(taps // 2):
                coeffs = struct.pack('>h', int(lpf[2 * n + 1])) + struct.pack('>h', int(lpf[2 * n]))
                register_name = 'FIR_b' + str(2 * n) + 'b' + str(2 * n + 1)
                self.roach.write(register_name, coeffs)
                self.roach.write_int('FIR_load_coeff', (ch << 1) + (1 << 0))
                self.roach.write_int('FIR_load_coeff', (ch << 1) + (0 << 0))

    # ------------------------------------------------------------------ a7
    def loadIQcenters(self):
        """ROACH_Pulses.py:948-956."""
        self.centers_int = []
        for ch in range(256):
            I_c = int(self.iq_centers[ch].real / 2 ** 3)
            Q_c = int(self.iq_centers[ch].imag / 2 ** 3)
            self.centers_int.append((I_c, Q_c))
            center = (I_c << 16) + (Q_c << 0)
            self.roach.write_int('conv_phase_centers', center)
            self.roach.write_int('conv_phase_load_centers', (ch << 1) + (1 << 0))
            self.roach.write_int('conv_phase_load_centers', 0)

    # ------------------------------------------------------------------ a3
    def select_bins(self, readout_freqs):
        """ROACH_Pulses.py:958-974."""
        fft_len = 2 ** 9
        i = 0
        residuals = []
        self.fft_bins = []
        for f in readout_freqs:
            fft_bin = int(_py2_round(f * fft_len / self.sampleRate))
            fft_freq = fft_bin * self.sampleRate / fft_len
            freq_residual = _py2_round((f - fft_freq) / self.freqRes) * self.freqRes
            residuals.append(freq_residual)
            self.fft_bins.append(fft_bin)
            self.roach.write_int('bins', fft_bin)
            self.roach.write_int('load_bins', (i << 1) + (1 << 0))
            self.roach.write_int('load_bins', (i << 1) + (0 << 0))
            i = i + 1
        return residuals

    def toggleDAC(self):
        """ROACH_Pulses.py:927-946 without the sleeps."""
        if self.dacStatus == 'off':
            self.roach.write_int('startDAC', 1)
            while self.roach.read_int('DRAM_LUT_rd_valid') != 0:
                self.roach.write_int('startDAC', 0)
                self.roach.write_int('startDAC', 1)
            self.dacStatus = 'on'
        else:
            self.roach.write_int('startDAC', 0)
            self.dacStatus = 'off'

    def loadLUTs(self):
        """ROACH_Pulses.py:976-1011: luts.dat -> dram_memory, centers.dat -> iq_centers, bins."""
        self.scale_factor = 1.
        self.iq_centers = numpy.array([0. + 0j] * 256)
        if self.dacStatus == 'off':
            self.roach.write_int('startDAC', 0)
        else:
            self.toggleDAC()
        saveDir = str(self.lutDir)
        f = open(os.path.join(saveDir, 'luts.dat'), 'rb')
        binaryData = f.read()
        f.close()
        self.binaryData = binaryData
        self.roach.write('dram_memory', binaryData)
        x = numpy.atleast_2d(numpy.loadtxt(os.path.join(saveDir, 'centers.dat')))
        N_freqs = len(x[:, 0])
        for n in range(N_freqs):
            self.iq_centers[n] = complex(x[n, 0], x[n, 1])
        freqs = [float(v) for v in self.dac_freqs]
        f_base = float(self.lo_freq)
        for n in range(len(freqs)):
            if freqs[n] < f_base:
                freqs[n] = freqs[n] + 512e6
        freqs_dds = [0 for j in range(256)]
        for n in range(len(freqs)):
            freqs_dds[n] = _py2_round((freqs[n] - f_base) / self.freqRes) * self.freqRes
        self.freq_residuals = self.select_bins(freqs_dds)
        self.loadIQcenters()
        self.toggleDAC()

    def dds_from_luts(self):
        """I_dds, Q_dds (int16 [N]) recovered from the DRAM image (inverse of write_LUTs' packing)."""
        a = numpy.frombuffer(self.binaryData, dtype='>i2').reshape(-1, 8)
        n = a.shape[0] * 2
        I_dds = numpy.empty(n, numpy.int16); Q_dds = numpy.empty(n, numpy.int16)
        Q_dds[1::2], Q_dds[0::2] = a[:, 0], a[:, 1]
        I_dds[1::2], I_dds[0::2] = a[:, 4], a[:, 5]
        return I_dds, Q_dds

    # ------------------------------------------------------------------ a10
    def loadThresholds(self, steps=10):
        """ROACH_Pulses.py:211-299 without raw_input/pickle: per channel `steps` phase snapshots of
        1024 words, histogram CDF, threshold = int(-2.5*|med - p5|) clamped at -25736."""
        Nsigma = 2.5
        N_freqs = len(self.dac_freqs)
        self.thresholds, self.medians = numpy.array([0.] * N_freqs), numpy.array([0.] * N_freqs)
        self.thresholds_raw = numpy.zeros(N_freqs, dtype=numpy.int64)
        L = 2 ** 10
        scale_to_angle = 360. / 2 ** 16 * 4 / numpy.pi
        for ch in range(N_freqs):
            bin_data_phase = b''
            for n in range(steps):
                self.roach.write_int('ch_we', ch)
                self.roach.write_int('startSnap', 0)
                self.roach.write_int('snapPhase_ctrl', 1)
                self.roach.write_int('snapPhase_ctrl', 0)
                self.roach.write_int('startSnap', 1)
                bin_data_phase = bin_data_phase + self.roach.read('snapPhase_bram', 4 * L)
            a = numpy.frombuffer(bin_data_phase, dtype='>i2').reshape(-1, 2)
            phase = numpy.stack([a[:, 1], a[:, 0]], axis=1).reshape(-1).astype(numpy.int64)   # :251-253
            threshold, med = self._threshold(phase, Nsigma)
            self.thresholds[ch] = scale_to_angle * threshold
            self.medians[ch] = scale_to_angle * med
            if self.customThresholds[ch] != 360.0:
                threshold = self.customThresholds[ch] / scale_to_angle
                if threshold < -25736:
                    threshold = -25736
            self.thresholds_raw[ch] = int(threshold)
            self.roach.write_int('capture_threshold', int(threshold))
            self.roach.write_int('capture_load_thresh', (ch << 1) + (1 << 0))
            self.roach.write_int('capture_load_thresh', (ch << 1) + (0 << 0))

    @staticmethod
    def _threshold(phase, Nsigma=2.5):
        n, bins = numpy.histogram(phase, bins=100)
        n = numpy.array(n, dtype='float32') / numpy.sum(n)
        tot = numpy.zeros(len(bins))
        for i in range(len(bins)):
            tot[i] = numpy.sum(n[:i])
        med = bins[(numpy.abs(tot - 0.5)).argmin()]
        thresh = bins[(numpy.abs(tot - 0.05)).argmin()]
        threshold = int(-Nsigma * abs(med - thresh))
        if threshold < -25736:
            threshold = -25736
        return threshold, med

    # ------------------------------------------------------------------ a14
    def readPulses(self, steps=None):
        """ROACH_Pulses.py:782-889: read both photon BRAMs between write pointers (ring wrap at
        2**14), unpack on the GPU, per-channel lists, Fix12_9 -> degrees, three 40-bin histograms."""
        scale_to_degrees = 360. / 2 ** 12 * 4 / numpy.pi
        channel_count = numpy.zeros(256, dtype=numpy.int64)
        p1 = [[] for n in range(256)]
        timestamp = [[] for n in range(256)]
        baseline = [[] for n in range(256)]
        peaks = [[] for n in range(256)]
        seconds = int(self.seconds)
        nStepsPerSec = 10.
        steps = int(seconds * nStepsPerSec) if steps is None else steps
        self.roach.write_int('startBuffer', 1)
        self.total_counts = []
        for n in range(steps):
            addr0 = self.roach.read_int('pulses_addr')
            addr1 = self.roach.read_int('pulses_addr')
            bin_data_0 = self.roach.read('pulses_bram0', 4 * 2 ** 14)
            bin_data_1 = self.roach.read('pulses_bram1', 4 * 2 ** 14)
            lo = numpy.frombuffer(bin_data_0, dtype='>u4').astype(numpy.uint64)
            hi = numpy.frombuffer(bin_data_1, dtype='>u4').astype(numpy.uint64)
            if addr1 >= addr0:
                idx = numpy.arange(addr0, addr1)
                wrap = False
                self.total_counts.append(addr1 - addr0)
            else:
                idx = numpy.concatenate([numpy.arange(addr0, 2 ** 14), numpy.arange(0, addr1)])
                wrap = True
                self.total_counts.append(addr1 + 2 ** 14 - addr0)
            if idx.size == 0:
                continue
            words = (hi[idx] << numpy.uint64(32)) | lo[idx]
            ch, ts, base, peak, p1f = _decode.unpack_fields(words, ctx=self.ctx)          # GPU (:805-811)
            channel_count += numpy.bincount(ch, minlength=256)
            order = numpy.argsort(ch, kind='stable')
            bounds = numpy.searchsorted(ch[order], numpy.arange(257))
            for c in numpy.nonzero(numpy.diff(bounds))[0]:
                sel = order[bounds[c]:bounds[c + 1]]
                timestamp[c].extend(ts[sel].tolist())
                baseline[c].extend(base[sel].tolist())
                peaks[c].extend(peak[sel].tolist())
                if wrap:     # the reference appends p1 only in the wrap-around branch (:820,:828)
                    p1[c].extend(((p1f[sel].astype(numpy.int64) - 2 ** 11) * scale_to_degrees).tolist())
        self.roach.write_int('startBuffer', 0)
        self.channel_count = channel_count
        self.timestamp, self.baseline, self.peaks, self.p1 = timestamp, baseline, peaks, p1
        ch = int(self.channel)
        base = numpy.array(baseline[ch], dtype='float')
        base = base / 2.0 ** 9 - 4.0
        base = base * 180.0 / numpy.pi
        times = numpy.array(timestamp[ch], dtype='float') / 1e6
        peaksCh = numpy.array(peaks[ch], dtype='float')
        peaksCh = peaksCh / 2.0 ** 9 - 4.0
        peaksCh = peaksCh * 180.0 / numpy.pi
        peaksSubBase = peaksCh - base
        r = (-150, 10)
        nBin = 40
        self.hgBase, self.bins = numpy.histogram(base, nBin, range=r, density=False)
        self.hgPeak, self.bins = numpy.histogram(peaksCh, nBin, range=r, density=False)
        self.hgPeakSubBase, self.bins = numpy.histogram(peaksSubBase, nBin, range=r, density=False)
        self.base_deg, self.peak_deg, self.times = base, peaksCh, times
        return channel_count
