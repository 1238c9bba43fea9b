"""Headless twin of the "channelizer" GUI's hot-path methods (class AppForm of
DataReadout/ChannelizerControls/ROACH_Pulses.py): FIR / centre / bin / threshold loading and the photon read-out.
Same method names, attributes and register traffic as the reference; `self.roach` is injectable.  The register
protocol lives in registers.py; the arithmetic is vectorised, the thresholds are derived on the GPU
(mkid_thresholds_from_phase) and the photon words are unpacked on the GPU (mkid_unpack_fields).
"""
import os

import numpy

from . import _lib
from . import decode as _decode
from . import registers as _regs
from .fake_roach import FakeRoach


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
        self.fir = [0.] * _regs.FIR_TAPS                     # importFIRcoeffs :1088-1103
        self.zeroChannels = [0] * 256
        self.customThresholds = numpy.array([360.0] * 256)   # 360.0 = "none" (:45)
        self.thresholds = numpy.array([0.] * 256)
        self.medians = numpy.array([0.] * 256)
        self.iq_centers = numpy.array([0. + 0j] * 256)
        self.scale_factor = 1.
        self.channel = 0                                     # textbox_channel
        self.seconds = 1                                     # textbox_seconds
        self.Nsigma = 2.5                                    # (:216)

    # ------------------------------------------------------------------ a6  (ROACH_Pulses.py:59-111, 1088-1103)
    def importFIRcoeffs(self, path):
        self.fir = list(numpy.loadtxt(path))

    def loadFIRcoeffs(self):
        """The 13 FIR registers of every channel: the quantised taps for the driven channels, zeros for deleted
        (zeroChannels) and undriven ones.  Sets self.fir_int (the 26 integers the data path runs with)."""
        n_driven = len(self.dac_freqs)
        self.fir_int = [int(v) for v in _regs.fir_taps_int(self.fir)]
        live = _regs.fir_register_words(self.fir_int)
        dead = numpy.zeros_like(live)
        for ch in range(_regs.N_CHANNELS):
            zeroed = ch >= n_driven or self.zeroChannels[ch]
            _regs.write_fir(self.roach, dead if zeroed else live, [ch])

    # ------------------------------------------------------------------ a7  (ROACH_Pulses.py:948-956)
    def loadIQcenters(self):
        i_c, q_c, words = _regs.center_words(self.iq_centers)
        self.centers_int = list(zip(i_c.tolist(), q_c.tolist()))
        _regs.write_centers(self.roach, words)

    # ------------------------------------------------------------------ a3  (ROACH_Pulses.py:958-974)
    def select_bins(self, readout_freqs):
        fft_bins, resid = _regs.coarse_fine(readout_freqs, self.sampleRate, self.freqRes)
        self.fft_bins = fft_bins.tolist()
        _regs.write_bins(self.roach, self.fft_bins)
        return resid.tolist()

    def toggleDAC(self):
        """ROACH_Pulses.py:927-946 (no sleeps: nothing here waits for hardware)."""
        if self.dacStatus == 'off':
            _regs.dac_start(self.roach)
            self.dacStatus = 'on'
        else:
            _regs.dac_stop(self.roach)
            self.dacStatus = 'off'

    def loadLUTs(self):
        """The hand-over from the set-up GUI (ROACH_Pulses.py:976-1011): luts.dat -> dram_memory, centers.dat -> centres,
        coarse bins from the tone list, DAC restarted."""
        self.scale_factor = 1.
        if self.dacStatus == 'off':
            _regs.dac_stop(self.roach)
        else:
            self.toggleDAC()
        with open(os.path.join(str(self.lutDir), 'luts.dat'), 'rb') as f:
            self.binaryData = f.read()
        self.roach.write('dram_memory', self.binaryData)
        centres = numpy.atleast_2d(numpy.loadtxt(os.path.join(str(self.lutDir), 'centers.dat')))
        self.iq_centers = numpy.zeros(256, dtype=complex)
        self.iq_centers[:len(centres)] = centres[:, 0] + 1j * centres[:, 1]
        tones = numpy.zeros(256)
        tones[:len(self.dac_freqs)] = _regs.baseband_tones(self.dac_freqs, self.lo_freq, 512e6, self.freqRes, mirror=False)
        self.freq_residuals = self.select_bins(tones.tolist())
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

    # ------------------------------------------------------------------ a10  (ROACH_Pulses.py:211-353)
    def _thresholds_gpu(self, phase_cols):
        """phase_cols: int16 [n_samples][n_ch] raw Fix16_13 snapshots -> (threshold_raw int, median edge) per channel, by
        the histogram-CDF rule of ROACH_Pulses.py:259-277 evaluated on the GPU (all channels in one launch)."""
        from . import triggers
        ctx = self.ctx or _lib.default_context()
        n, n_ch = phase_cols.shape
        dev = ctx.to_device(numpy.ascontiguousarray(phase_cols, dtype=numpy.int16))
        thr, med, _ = triggers.thresholds_from_phase(dev, 1, n, n, n_ch=n_ch, Nsigma=self.Nsigma, ctx=ctx)
        dev.free()
        return thr[0], med[0]

    def _apply_threshold(self, ch, thr_raw, med_edge):
        self.thresholds[ch] = _regs.PHASE_LSB_DEG * thr_raw
        self.medians[ch] = _regs.PHASE_LSB_DEG * med_edge
        value = thr_raw
        if self.customThresholds[ch] != 360.0:
            value = _regs.custom_threshold_raw(self.customThresholds[ch])
        self.thresholds_raw[ch] = int(value)
        _regs.write_threshold(self.roach, ch, int(value))

    def loadThresholds(self, steps=10, batched=False):
        """Per driven channel: `steps` phase snapshots, threshold = int(-Nsigma * |median - 5 % point|) of the 100-bin
        histogram CDF, not below -pi; custom thresholds override.  batched=False keeps the reference's register order
        (snapshots of a channel, then its threshold); batched=True takes all snapshots first and derives every
        threshold in one kernel launch."""
        n_driven = len(self.dac_freqs)
        self.thresholds, self.medians = numpy.zeros(n_driven), numpy.zeros(n_driven)
        self.thresholds_raw = numpy.zeros(n_driven, dtype=numpy.int64)
        if batched:
            cols = numpy.stack([_regs.snapshot_phase(self.roach, ch, steps) for ch in range(n_driven)], axis=1)
            thr, med = self._thresholds_gpu(cols)
            for ch in range(n_driven):
                self._apply_threshold(ch, int(thr[ch]), float(med[ch]))
            return
        for ch in range(n_driven):
            thr, med = self._thresholds_gpu(_regs.snapshot_phase(self.roach, ch, steps)[:, None])
            self._apply_threshold(ch, int(thr[0]), float(med[0]))

    def loadSingleThreshold(self, ch, steps=10):
        """ROACH_Pulses.py:301-353: the same rule for one channel (self.thresholds / medians keep their length)."""
        if not hasattr(self, 'thresholds_raw') or len(self.thresholds_raw) <= ch:
            self.thresholds_raw = numpy.zeros(max(ch + 1, len(self.thresholds)), dtype=numpy.int64)
        thr, med = self._thresholds_gpu(_regs.snapshot_phase(self.roach, ch, steps)[:, None])
        self._apply_threshold(ch, int(thr[0]), float(med[0]))

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
