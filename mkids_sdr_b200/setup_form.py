"""Headless twin of the "templar" GUI's LUT path (class AppForm of
DataReadout/ChannelizerControls/ROACH_Setup.py and its multi-tone variant ROACH_Setup_DAC.py).

Same method names, argument order, attributes and register traffic as the reference; Qt widgets become plain
attributes, `self.roach` is injectable (FakeRoach by default).  The tone placement and the register protocol are the
vectorised helpers of registers.py; the hot loops (comb, DDS tables, DRAM image) run on the GPU.
"""
import os

import numpy

from . import lut as _lut
from . import registers as _regs
from .fake_roach import FakeRoach


class SetupForm:
    def __init__(self, roach=None, sampleRate=512e6, N_lut_entries=2 ** 16, multi_tone=True, ctx=None,
                 LUT_saveDir='.'):
        self.ctx = ctx
        self.roach = roach if roach is not None else FakeRoach()
        self.sampleRate = sampleRate                       # ROACH_Setup.py:82
        self.N_lut_entries = N_lut_entries                 # :83
        self.freqRes = self.sampleRate / self.N_lut_entries  # :84
        self.multi_tone = multi_tone                       # True: ROACH_Setup_DAC.py semantics
        self.dacStatus = 'off'
        self.LUT_saveDir = LUT_saveDir
        self.last_scale_factor = None
        self.scale_factor = None
        self.previous_scale_factor = 1.0
        self.iq_centers = numpy.array([0. + 0j] * 256)
        # former GUI widgets
        self.dac_freqs = []            # textedit_DACfreqs / spinBox_DACfreq (Hz)
        self.lo_freq = 0.0             # textbox_loFreq / spinBox_loFreq
        self.offset = 0                # textbox_offset           (ROACH_Setup_DAC.py:397)
        self.keep_old_scale = False    # cbox_keepScaleFactor     (:398)
        self.use_custom_scale = False  # cbox_useScaleFactor      (:399)
        self.custom_scale = 1.0        # textbox_customScale
        self.attens = numpy.array([0.])
        self.minimumAttenuation = 0.0
        self.save_npz = True

    # ------------------------------------------------------------------ a1
    def freqCombLUT(self, echo, freq, sampleRate, resolution, amplitude=[1.] * 256, phase=[0.] * 256,
                    random_phase='yes'):
        """ROACH_Setup.py:416-475.  Returns (I, Q) integer numpy arrays; sets scale_factor /
        last_scale_factor when echo == 'yes'; overwrites phase[n] when random_phase == 'yes'."""
        N_freqs = len(freq)
        size = int(sampleRate / resolution)
        override = None
        if echo == 'yes':
            if self.keep_old_scale and self.last_scale_factor is not None:
                override = self.last_scale_factor
            if self.use_custom_scale:
                override = float(self.custom_scale)
        I, Q, scale, ph = _lut.comb_lut([list(freq)], sampleRate, size, [list(amplitude[:N_freqs])],
                                        [list(phase[:N_freqs])], echo, random_phase, self.offset,
                                        override, ctx=self.ctx)
        if random_phase == 'yes':
            for n in range(N_freqs):
                phase[n] = float(ph[0, n])            # the reference mutates the caller's list (:429)
        if echo == 'yes':
            self.scale_factor = float(scale[0])
            self.last_scale_factor = self.scale_factor
        return I[0].astype(numpy.int64), Q[0].astype(numpy.int64)

    # ------------------------------------------------------------------ a2  (ROACH_Setup.py:477-504, ROACH_Setup_DAC.py:457-483)
    def define_DAC_LUT(self):
        """DAC comb: tones mirrored about the LO onto the freqRes grid, amplitudes from the attenuations relative to
        the smallest one, comb synthesised by freqCombLUT('yes', ...)."""
        self.freqs_dac = _regs.baseband_tones(self.dac_freqs, self.lo_freq, self.sampleRate, self.freqRes, mirror=True).tolist()
        attens = numpy.asarray(self.attens, dtype=float)
        loudest = attens.min()
        amplitudes = [10 ** (+(loudest - a) / 20.) for a in attens]
        self.I_dac, self.Q_dac = self.freqCombLUT('yes', self.freqs_dac, self.sampleRate, self.freqRes, amplitudes)

    # ------------------------------------------------------------------ a3  (ROACH_Setup.py:534-550)
    def select_bins(self, readout_freqs):
        """Coarse fs/512 bin + residual per channel; the bins go to the firmware (three register writes each)."""
        fft_bins, resid = _regs.coarse_fine(readout_freqs, self.sampleRate, self.freqRes)
        self.fft_bins = fft_bins.tolist()
        _regs.write_bins(self.roach, self.fft_bins)
        return resid.tolist()

    # ------------------------------------------------------------------ a4  (ROACH_Setup.py:506-532)
    def define_DDS_LUT(self, phase=[0.] * 256):
        """256 fine-mixer tables (residual tone of every channel at 2 MS/s, start phase `phase[m]`), interleaved with
        the channel lag of 154; one GPU launch."""
        tones = numpy.zeros(_regs.N_CHANNELS)
        tones[:len(self.dac_freqs)] = _regs.baseband_tones(self.dac_freqs, self.lo_freq, self.sampleRate, self.freqRes,
                                                           mirror=False)
        self.freq_residuals = self.select_bins(tones.tolist())
        n_lut = int(self.sampleRate / self.freqRes)
        I, Q, sc = _lut.dds_lut([self.freq_residuals], [list(phase[:256])], self.sampleRate, n_lut, 154, self.offset,
                                ctx=self.ctx)
        self.I_dds, self.Q_dds = I[0].astype(numpy.int64), Q[0].astype(numpy.int64)
        self.dds_scales = sc[0]

    def define_LUTs(self):
        """ROACH_Setup.py:395-414."""
        self.iq_centers = numpy.zeros(256, dtype=complex)
        self.define_DAC_LUT()
        self.define_DDS_LUT()
        self.write_LUTs()

    # ------------------------------------------------------------------ a5  (ROACH_Setup.py:552-578)
    def write_LUTs(self):
        """dac.npy.npz, the 8*N-byte DRAM image to 'dram_memory' (packed on the GPU) and to luts.dat."""
        if self.dacStatus == 'off':
            _regs.dac_stop(self.roach)
        else:
            self.toggleDAC()
        if self.save_npz:
            numpy.savez(os.path.join(self.LUT_saveDir, 'dac.npy'), I_dac=self.I_dac, Q_dac=self.Q_dac, I_dds=self.I_dds,
                        Q_dds=self.Q_dds)
        self.binaryData = _lut.pack_dram(self.I_dac, self.Q_dac, self.I_dds, self.Q_dds, ctx=self.ctx)
        self.roach.write('dram_memory', self.binaryData)
        with open(os.path.join(self.LUT_saveDir, 'luts.dat'), 'wb') as f:
            f.write(self.binaryData)

    def toggleDAC(self):
        """ROACH_Setup.py:881-914 (no sleeps)."""
        if self.dacStatus == 'off':
            _regs.dac_start(self.roach)
            self.dacStatus = 'on'
        else:
            _regs.dac_stop(self.roach)
            self.dacStatus = 'off'

    # ------------------------------------------------------------------ a7  (ROACH_Setup.py:595-625)
    def findIQcenters(self, I, Q):
        """Centre of an IQ loop: midpoint of the extreme values of each component."""
        return complex((I.max() + I.min()) / 2., (Q.max() + Q.min()) / 2.)

    def loadIQcenters(self):
        """Centres to the firmware and to centers.dat (:611)."""
        _regs.write_centers(self.roach, _regs.center_words(self.iq_centers)[2])
        numpy.savetxt(os.path.join(self.LUT_saveDir, 'centers.dat'),
                      numpy.column_stack([self.iq_centers.real, self.iq_centers.imag]))

    def rotateLoopsReady(self):
        """ROACH_Setup.py:645-671 (without the LO sweep that follows it in the GUI): read the averaged IQ point of every
        channel, rotate the loops, restart the DAC that write_LUTs stopped."""
        I_avg, Q_avg = _regs.snapshot_avg_iq(self.roach)
        phase = self.rotateLoops(I_avg, Q_avg)
        self.toggleDAC()
        return phase

    def rotateLoops(self, I_avg, Q_avg):
        """The DDS start phase of every driven channel becomes the angle of its averaged IQ point about its centre
        (arctan2(Q - Qc, I - Ic), ROACH_Setup.py:661-664); the DDS tables and the DRAM image are rebuilt."""
        n = len(self.dac_freqs)
        c = numpy.asarray(self.iq_centers[:n])
        phase = [0.] * 256
        phase[:n] = numpy.arctan2(numpy.asarray(Q_avg, float)[:n] - c.imag, numpy.asarray(I_avg, float)[:n] - c.real).tolist()
        self.define_DDS_LUT(phase)
        self.write_LUTs()
        return phase

    def loadFreqsAttens(self, freqFile, f_offset=0.0):
        """Frequency file of ROACH_Setup_DAC.py:892-927: row 0 = previous scale factor, then f_GHz I_c Q_c atten."""
        x = numpy.atleast_2d(numpy.loadtxt(freqFile))
        self.previous_scale_factor = x[0, 0]
        body = x[1:]
        self.dac_freqs = [f_ghz * 1e9 + f_offset for f_ghz in body[:, 0]]
        self.iq_centers = numpy.zeros(256, dtype=complex)
        self.iq_centers[:len(body)] = body[:, 1] + 1j * body[:, 2]
        self.attens = body[:, 3]
        self.minimumAttenuation = numpy.array(body[:, 3]).min()
