"""Headless twin of the "templar" GUI's LUT path (class AppForm of
DataReadout/ChannelizerControls/ROACH_Setup.py and its multi-tone variant ROACH_Setup_DAC.py).

Same method names, argument order and attributes as the reference; Qt widgets become plain
attributes, `self.roach` is injectable (FakeRoach by default) and the hot loops run on the GPU.
"""
import math
import os
import struct

import numpy

from . import lut as _lut
from .fake_roach import FakeRoach


def _py2_round(x):
    """Python-2 round(): half away from zero (ROACH_Setup.py:498,517,540,542 were written for it)."""
    x = float(x)
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


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

    # ------------------------------------------------------------------ a2
    def define_DAC_LUT(self):
        f_base = float(self.lo_freq)
        freqs = [float(f) for f in self.dac_freqs]
        # both GUIs (ROACH_Setup.py:484-495, ROACH_Setup_DAC.py:464-475): mirror about LO, +fs if below
        freqs = [f_base + (f_base - f) for f in freqs]
        freqs = [f + self.sampleRate if f < f_base else f for f in freqs]
        self.freqs_dac = [_py2_round((f - f_base) / self.freqRes) * self.freqRes for f in freqs]
        atten_min = numpy.asarray(self.attens, dtype=float).min()
        amplitudes = [10 ** (+(atten_min - a) / 20.) for a in numpy.asarray(self.attens, dtype=float)]
        self.I_dac, self.Q_dac = self.freqCombLUT('yes', self.freqs_dac, self.sampleRate, self.freqRes, amplitudes)

    # ------------------------------------------------------------------ a3
    def select_bins(self, readout_freqs):
        """ROACH_Setup.py:534-550: fft bin + residual per channel, three register writes each."""
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

    # ------------------------------------------------------------------ a4
    def define_DDS_LUT(self, phase=[0.] * 256):
        """ROACH_Setup.py:506-532 (256 tables on the GPU in one launch)."""
        ch_shift = 154
        freqs = [float(f) for f in self.dac_freqs]
        f_base = float(self.lo_freq)
        for n in range(len(freqs)):
            if freqs[n] < f_base:
                freqs[n] = freqs[n] + self.sampleRate
        freqs_dds = [0 for j in range(256)]
        for n in range(len(freqs)):
            freqs_dds[n] = _py2_round((freqs[n] - f_base) / self.freqRes) * self.freqRes
        freq_residuals = self.select_bins(freqs_dds)
        self.freq_residuals = freq_residuals
        L = int(self.sampleRate / self.freqRes)
        I, Q, sc = _lut.dds_lut([freq_residuals], [list(phase[:256])], self.sampleRate, L, ch_shift, self.offset,
                                ctx=self.ctx)
        self.I_dds, self.Q_dds = I[0].astype(numpy.int64), Q[0].astype(numpy.int64)
        self.dds_scales = sc[0]

    def define_LUTs(self):
        """ROACH_Setup.py:395-414."""
        self.iq_centers = numpy.array([0. + 0j] * 256)
        self.define_DAC_LUT()
        self.define_DDS_LUT()
        self.write_LUTs()

    # ------------------------------------------------------------------ a5
    def write_LUTs(self):
        """ROACH_Setup.py:552-578: dac.npy.npz, DRAM image to 'dram_memory', luts.dat."""
        if self.dacStatus == 'off':
            self.roach.write_int('startDAC', 0)
        else:
            self.toggleDAC()
        if self.save_npz:
            numpy.savez(os.path.join(self.LUT_saveDir, 'dac.npy'), I_dac=self.I_dac, Q_dac=self.Q_dac, I_dds=self.I_dds,
                        Q_dds=self.Q_dds)
        binaryData = _lut.pack_dram(self.I_dac, self.Q_dac, self.I_dds, self.Q_dds, ctx=self.ctx)
        self.binaryData = binaryData
        self.roach.write('dram_memory', binaryData)
        f = open(os.path.join(self.LUT_saveDir, 'luts.dat'), 'wb')
        f.write(binaryData)
        f.close()

    def toggleDAC(self):
        """ROACH_Setup.py:881-914 without the sleeps."""
        if self.dacStatus == 'off':
            self.roach.write_int('startDAC', 1)
            while self.roach.read_int('DRAM_LUT_rd_valid') != 0:
                self.roach.write_int('startDAC', 0)
                self.roach.write_int('startDAC', 1)
            self.dacStatus = 'on'
        else:
            self.roach.write_int('startDAC', 0)
            self.dacStatus = 'off'

    # ------------------------------------------------------------------ a7
    def findIQcenters(self, I, Q):
        """ROACH_Setup.py:621-625."""
        I_0 = (I.max() + I.min()) / 2.
        Q_0 = (Q.max() + Q.min()) / 2.
        return complex(I_0, Q_0)

    def loadIQcenters(self):
        """ROACH_Setup.py:595-606, plus centers.dat (:611)."""
        for ch in range(256):
            I_c = int(self.iq_centers[ch].real / 2 ** 3)
            Q_c = int(self.iq_centers[ch].imag / 2 ** 3)
            center = (I_c << 16) + (Q_c << 0)
            self.roach.write_int('conv_phase_centers', center)
            self.roach.write_int('conv_phase_load_centers', (ch << 1) + (1 << 0))
            self.roach.write_int('conv_phase_load_centers', 0)
        numpy.savetxt(os.path.join(self.LUT_saveDir, 'centers.dat'),
                      numpy.column_stack([self.iq_centers.real, self.iq_centers.imag]))

    def loadFreqsAttens(self, freqFile, f_offset=0.0):
        """Freq-file format of ROACH_Setup_DAC.py:892-927: row 0 = previous scale factor, rows =
        f_GHz I_c Q_c atten."""
        x = numpy.atleast_2d(numpy.loadtxt(freqFile))
        self.previous_scale_factor = x[0, 0]
        N_freqs = len(x[1:, 0])
        self.dac_freqs = [l * 1e9 + f_offset for l in x[1:, 0]]
        self.iq_centers = numpy.array([0. + 0j] * 256)
        for n in range(N_freqs):
            self.iq_centers[n] = complex(x[n + 1, 1], x[n + 1, 2])
        self.attens = x[1:, 3]
        self.minimumAttenuation = numpy.array(x[1:, 3]).min()
