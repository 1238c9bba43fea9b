"""GPU: the headless GUI twins (SetupForm / PulsesForm) behave like the reference's AppForm methods:
same register traffic, same files, same numbers (oracle = restated reference lines)."""
import os
import struct

import numpy as np
import pytest

from oracle import control, decode as odec, lut as olut

pytestmark = pytest.mark.gpu
FS = 512e6


@pytest.fixture(scope='module')
def ctx():
    from mkids_sdr_b200 import _lib
    return _lib.default_context(0)


def test_multitone_setup_to_pulses_roundtrip(ctx, tmp_path, golden_dir):
    from mkids_sdr_b200.pulses_form import PulsesForm
    from mkids_sdr_b200.setup_form import SetupForm
    N = 2 ** 16
    rng = np.random.default_rng(4)
    lo = 5.0e9
    k = np.sort(rng.choice(np.arange(-N // 2 + 100, N // 2 - 100), 60, replace=False))
    freqs = [lo + float(v) * FS / N for v in k]
    sf = SetupForm(N_lut_entries=N, multi_tone=True, ctx=ctx, LUT_saveDir=str(tmp_path))
    sf.dac_freqs, sf.lo_freq = freqs, lo
    sf.attens = rng.integers(0, 15, 60).astype(float)
    sf.iq_centers = np.array([0. + 0j] * 256)
    sf.iq_centers[:60] = rng.normal(0, 400, 60) + 1j * rng.normal(0, 400, 60)
    sf.define_DAC_LUT()
    sf.define_DDS_LUT([0.] * 256)
    sf.write_LUTs()
    sf.loadIQcenters()
    # oracle: the same through the restated reference
    fd = olut.dac_freqs_multi(freqs, lo, FS / N)
    Io, Qo, so, _ = olut.freq_comb_lut('yes', fd, FS, FS / N, olut.dac_amplitudes(sf.attens))
    assert sf.freqs_dac == fd and sf.scale_factor == so
    assert np.array_equal(sf.I_dac, Io) and np.array_equal(sf.Q_dac, Qo)
    bins, resid = olut.select_bins(olut.dds_freqs(freqs, lo, FS, FS / N), FS, FS / N)
    Id, Qd, _ = olut.define_dds_lut(resid, FS, FS / N)
    assert sf.fft_bins == bins and sf.freq_residuals == resid
    assert np.array_equal(sf.I_dds, Id) and np.array_equal(sf.Q_dds, Qd)
    assert sf.binaryData == olut.pack_dram(Io, Qo, Id, Qd)
    assert sf.roach.writes('bins') == bins
    assert sf.roach.writes('load_bins')[:4] == [1, 0, 3, 2]
    # the channelizer GUI reloads the files
    pf = PulsesForm(N_lut_entries=N, ctx=ctx)
    pf.dac_freqs, pf.lo_freq, pf.lutDir = freqs, lo, str(tmp_path)
    pf.loadLUTs()
    assert pf.roach.mem['dram_memory'] == sf.binaryData
    assert pf.fft_bins == bins
    Id2, Qd2 = pf.dds_from_luts()
    assert np.array_equal(Id2, Id) and np.array_equal(Qd2, Qd)
    want = [control.iq_center_word(c)[0] for c in sf.iq_centers]
    assert pf.roach.writes('conv_phase_centers') == want
    # FIR registers (known answers of the shipped matched filter)
    pf.importFIRcoeffs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'mkids_sdr_b200', 'data',
                                    'matched_30us.txt'))
    pf.zeroChannels = [0] * 256
    pf.zeroChannels[3] = 1
    pf.loadFIRcoeffs()
    regs = [r[2] for r in control.fir_registers(pf.fir)]
    log = [(n, v) for op, n, v in pf.roach.log if op == 'write' and n.startswith('FIR_b')]
    assert len(log) == 256 * 13
    assert pf.fir_int == control.fir_quantise(pf.fir)
    assert struct.unpack('>l', regs[0])[0] == 0x9B0A0


def test_load_thresholds_golden_snapshot(ctx, golden_dir):
    from mkids_sdr_b200.pulses_form import PulsesForm
    deg = np.load(os.path.join(golden_dir, 'ch_snap_0.npy'))
    raw = np.round(deg / control.SCALE_TO_ANGLE).astype(np.int64)            # 2048 Fix16_13 samples
    pf = PulsesForm(ctx=ctx)
    pf.dac_freqs = [1.0]
    # snapPhase_bram: 1024 words, second sample in bytes [0:2], first in bytes [2:4] (ROACH_Pulses.py:251-253)
    words = np.empty((1024, 2), dtype='>i2')
    words[:, 1] = raw[0::2]; words[:, 0] = raw[1::2]
    pf.roach.mem['snapPhase_bram'] = words.tobytes()
    pf.loadThresholds(steps=1)
    assert pf.thresholds_raw[0] == -5913
    assert abs(pf.thresholds[0] - (-41.356194)) < 1e-5 and abs(pf.medians[0] - 132.619587) < 1e-5
    assert pf.roach.writes('capture_threshold') == [-5913]
    assert np.array_equal(control.decode_phase_snapshot(words.tobytes()), raw)


def test_read_pulses_matches_reference_loop(ctx):
    from mkids_sdr_b200.pulses_form import PulsesForm
    rng = np.random.default_rng(11)
    n = 2 ** 14
    ch = rng.integers(0, 256, n)
    w = odec.pack_word(ch, rng.integers(0, 4096, n), rng.integers(0, 4096, n), rng.integers(0, 4096, n),
                       rng.integers(0, 2 ** 20, n))
    b0 = (w & np.uint64(0xFFFFFFFF)).astype('>u4').tobytes()
    b1 = (w >> np.uint64(32)).astype('>u4').tobytes()

    class Roach:
        def __init__(self, addrs): self.addrs = list(addrs)
        def write_int(self, *a): pass
        def read_int(self, name): return self.addrs.pop(0)
        def read(self, name, size): return b0 if name == 'pulses_bram0' else b1

    pairs = [(100, 5000), (16000, 300), (7, 7)]
    pf = PulsesForm(roach=Roach([a for p in pairs for a in p]), ctx=ctx)
    pf.channel = 3
    cc = pf.readPulses(steps=3)
    ref = odec.read_pulses([b0] * 3, [b1] * 3, pairs, sel_ch=3)
    assert np.array_equal(cc, ref['channel_count'])
    assert pf.total_counts == ref['total_counts']
    for c in (0, 3, 77, 255):
        assert pf.timestamp[c] == ref['timestamp'][c] and pf.baseline[c] == ref['baseline'][c]
        assert pf.peaks[c] == ref['peaks'][c] and pf.p1[c] == ref['p1'][c]
    assert np.array_equal(pf.hgBase, ref['hgBase']) and np.array_equal(pf.hgPeak, ref['hgPeak'])
    assert np.array_equal(pf.hgPeakSubBase, ref['hgPeakSubBase'])


def test_setup_form_matches_reference_run(ctx, tmp_path, golden_dir):
    """SetupForm against the outputs of the reference's own ROACH_Setup_DAC.py methods executed in the dev container
    (tests/golden/refrun_golden.npz): define_DAC_LUT (mirror about LO!), freqCombLUT with seed-1000 phases,
    define_DDS_LUT with non-zero phases, select_bins register traffic, write_LUTs DRAM image."""
    from mkids_sdr_b200.setup_form import SetupForm
    g = np.load(os.path.join(golden_dir, 'refrun_golden.npz'))
    N = int(g['setup_N'])
    sf = SetupForm(N_lut_entries=N, multi_tone=True, ctx=ctx, LUT_saveDir=str(tmp_path))
    sf.dac_freqs, sf.lo_freq = [float(f) for f in g['setup_dac_freqs']], float(g['setup_lo'])
    sf.attens = np.array(g['setup_attens'])
    sf.define_DAC_LUT()
    sf.define_DDS_LUT(list(g['setup_dds_phase']))
    sf.write_LUTs()
    assert np.array_equal(np.array(sf.freqs_dac), g['setup_freqs_dac'])
    assert sf.scale_factor == float(g['setup_scale_factor'])
    assert np.array_equal(sf.I_dac, g['setup_I_dac']) and np.array_equal(sf.Q_dac, g['setup_Q_dac'])
    assert np.array_equal(sf.I_dds, g['setup_I_dds']) and np.array_equal(sf.Q_dds, g['setup_Q_dds'])
    assert sf.roach.writes('bins') == [int(v) for v in g['setup_bins']]
    assert np.array_equal(np.frombuffer(sf.binaryData, dtype=np.uint8), g['setup_dram'])


def test_pulses_form_matches_reference_run(ctx, golden_dir):
    """PulsesForm.loadThresholds / readPulses against the reference's own methods executed in the dev container."""
    from mkids_sdr_b200.pulses_form import PulsesForm
    g = np.load(os.path.join(golden_dir, 'refrun_golden.npz'))
    pf = PulsesForm(ctx=ctx)
    pf.dac_freqs = [1.0, 2.0]
    snaps = []
    for ch in range(2):
        raw = g['thr_raw_phase'][ch].reshape(10, 2048)
        for st in range(10):
            words = np.empty((1024, 2), dtype='>i2')
            words[:, 1] = raw[st, 0::2]; words[:, 0] = raw[st, 1::2]
            snaps.append(words.tobytes())

    class Roach:
        def __init__(self): self.snaps, self.caps = list(snaps), []
        def write_int(self, name, v, *a):
            if name == 'capture_threshold': self.caps.append(int(v))
        def read(self, name, size): return self.snaps.pop(0)
    pf.roach = Roach()
    pf.loadThresholds(steps=10)
    assert pf.roach.caps == [int(v) for v in g['thr_capture_threshold']]
    assert np.array_equal(pf.thresholds, g['thr_thresholds_deg']) and np.array_equal(pf.medians, g['thr_medians_deg'])
    # readPulses
    w = g['rp_words']
    b0 = (w & np.uint64(0xFFFFFFFF)).astype('>u4').tobytes()
    b1 = (w >> np.uint64(32)).astype('>u4').tobytes()
    pairs = [tuple(int(v) for v in p) for p in g['rp_pairs']]

    class Roach2:
        def __init__(self): self.addrs = [a for p in pairs for a in p]
        def write_int(self, *a): pass
        def read_int(self, name): return self.addrs.pop(0)
        def read(self, name, size): return b0 if name == 'pulses_bram0' else b1
    pf2 = PulsesForm(roach=Roach2(), ctx=ctx)
    pf2.channel = 3
    cc = pf2.readPulses(steps=len(pairs))
    assert np.array_equal(cc, g['rp_channel_count'])
    assert np.array_equal(pf2.hgBase, g['rp_hgBase']) and np.array_equal(pf2.hgPeak, g['rp_hgPeak'])
    assert np.array_equal(pf2.hgPeakSubBase, g['rp_hgPeakSubBase'])
    assert np.array_equal(pf2.peak_deg, g['rp_peaksCh_deg']) and np.array_equal(pf2.times, g['rp_timesCh'])


def test_thresholds_batched_and_single_equal_the_reference_order(ctx, golden_dir):
    """loadThresholds(batched=True) (all snapshots first, one kernel launch) and loadSingleThreshold
    (ROACH_Pulses.py:301-353) give the thresholds of the reference-run golden data; custom thresholds override."""
    from mkids_sdr_b200.pulses_form import PulsesForm
    g = np.load(os.path.join(golden_dir, 'refrun_golden.npz'))

    def snaps_of(ch):
        raw = g['thr_raw_phase'][ch].reshape(10, 2048)
        out = []
        for st in range(10):
            words = np.empty((1024, 2), dtype='>i2')
            words[:, 1] = raw[st, 0::2]; words[:, 0] = raw[st, 1::2]
            out.append(words.tobytes())
        return out

    class Roach:
        def __init__(self, snaps): self.snaps, self.caps, self.log = list(snaps), [], []
        def write_int(self, name, v, *a):
            self.log.append((name, int(v)))
            if name == 'capture_threshold': self.caps.append(int(v))
        def read(self, name, size): return self.snaps.pop(0)
    pf = PulsesForm(roach=Roach(snaps_of(0) + snaps_of(1)), ctx=ctx)
    pf.dac_freqs = [1.0, 2.0]
    pf.loadThresholds(steps=10, batched=True)
    assert pf.roach.caps == [int(v) for v in g['thr_capture_threshold']]
    assert np.array_equal(pf.thresholds, g['thr_thresholds_deg']) and np.array_equal(pf.medians, g['thr_medians_deg'])
    # single channel, with and without a custom threshold
    pf.roach = Roach(snaps_of(1) + snaps_of(1))
    pf.loadSingleThreshold(1)
    assert pf.roach.caps == [int(g['thr_capture_threshold'][1])]
    assert pf.roach.log[-2:] == [('capture_load_thresh', 3), ('capture_load_thresh', 2)]
    pf.customThresholds[1] = -30.0
    pf.loadSingleThreshold(1)
    assert pf.roach.caps[-1] == int(-30.0 / control.SCALE_TO_ANGLE)
    assert pf.thresholds[1] == g['thr_thresholds_deg'][1]                     # the derived value is still reported


def test_rotate_loops_ready(ctx):
    """rotateLoopsReady (ROACH_Setup.py:645-671): phases = arctan2(Q - Qc, I - Ic) of the averaged IQ read-out, DDS tables
    rebuilt with them (== the oracle's define_dds_lut), DRAM image rewritten, DAC restarted."""
    from mkids_sdr_b200.setup_form import SetupForm
    N = 2 ** 12
    sf = SetupForm(N_lut_entries=N, multi_tone=True, ctx=ctx)
    sf.save_npz = False
    sf.LUT_saveDir = os.environ.get('TMPDIR', '/tmp')
    sf.lo_freq = 5.0e9
    rng = np.random.default_rng(5)
    k = np.sort(rng.choice(np.arange(-N // 2 + 8, N // 2 - 8), 6, replace=False))
    sf.dac_freqs = [sf.lo_freq + float(v) * FS / N for v in k]
    sf.attens = np.zeros(6)
    sf.iq_centers = np.zeros(256, dtype=complex)
    sf.iq_centers[:6] = rng.integers(-50, 50, 6) + 1j * rng.integers(-50, 50, 6)
    sf.define_LUTs()
    I_avg = rng.integers(-10 ** 5, 10 ** 5, 256); Q_avg = rng.integers(-10 ** 5, 10 ** 5, 256)
    sf.roach.mem['avgIQ_bram'] = np.concatenate([I_avg, Q_avg]).astype('>i4').tobytes()
    phase = sf.rotateLoopsReady()
    exp = np.arctan2(Q_avg[:6] - sf.iq_centers[:6].imag, I_avg[:6] - sf.iq_centers[:6].real)
    assert np.array_equal(np.array(phase[:6]), exp) and phase[6:] == [0.] * 250
    Io, Qo, _ = olut.define_dds_lut(sf.freq_residuals, FS, FS / N, phase)
    assert np.array_equal(sf.I_dds, Io) and np.array_equal(sf.Q_dds, Qo)
    assert sf.dacStatus == 'on' and sf.roach.writes('startDAC')[-1] == 1
    assert sf.binaryData == olut.pack_dram(sf.I_dac, sf.Q_dac, sf.I_dds, sf.Q_dds)
