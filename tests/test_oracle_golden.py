"""Pin the oracle against the reference's own fixtures / known answers (CPU only).

Fixtures under tests/golden/ were condensed from /root/reference by
tests/golden/make_golden.py."""
import hashlib
import os

import numpy as np
import pytest

from oracle import control, decode, fixed, lut, trigger


# ------------------------------------------------------------------ LUT (a1,a2,a4,a5)
def test_dac_golden_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, 'dac_golden.npz'))
    I, Q, scale, ph = lut.freq_comb_lut('yes', [412e6], 512e6, 7812.5, [1.0])
    assert ph[0] == 4.106624480316831
    assert abs(scale - 1.1) < 1e-3
    assert np.array_equal(I, g['I_dac'].astype(np.int64))
    assert np.array_equal(Q, g['Q_dac'].astype(np.int64))
    assert hashlib.sha256(I.astype('<i2').tobytes()).hexdigest() == str(g['I_dac_sha256'])
    assert str(g['I_dac_sha256']).startswith('c449906788665797857b3b505edc5ebc')
    assert str(g['Q_dac_sha256']).startswith('bbb728b827f6d4043cca562c646323c2')


def test_dac_golden_through_define_dac_lut(golden_dir):
    g = np.load(os.path.join(golden_dir, 'dac_golden.npz'))
    f = lut.dac_freqs_single(4.75e9, 4.65e9, 512e6, 7812.5)      # LO 4.65 GHz, tone 4.75 GHz
    assert f == [412e6]
    I, Q, _, _ = lut.freq_comb_lut('yes', f, 512e6, 7812.5, lut.dac_amplitudes([5.0]))
    assert np.array_equal(I, g['I_dac'].astype(np.int64))


def test_dds_golden_and_dram_image(golden_dir):
    g = np.load(os.path.join(golden_dir, 'dac_golden.npz'))
    freqs_dds = lut.dds_freqs([4.75e9], 4.65e9, 512e6, 7812.5)
    bins, resid = lut.select_bins(freqs_dds, 512e6, 7812.5)
    assert bins[0] == 100 and resid[0] == 0.0
    I_dds, Q_dds, _ = lut.define_dds_lut(resid, 512e6, 7812.5)
    assert np.array_equal(I_dds, g['I_dds'].astype(np.int64))
    assert np.array_equal(Q_dds, g['Q_dds'].astype(np.int64))
    img = lut.pack_dram(g['I_dac'], g['Q_dac'], g['I_dds'], g['Q_dds'])
    assert len(img) == 524288
    assert img[:16].hex() == '000000001e28a0557fff7fff8f9abdbd'
    assert hashlib.sha256(img).hexdigest() == \
        '44b66622c49f414ceeae34d9011391c081bdf1ab06902a14de6243df3ff270d1'
    assert lut.pack_dram_literal(g['I_dac'][:64], g['Q_dac'][:64], g['I_dds'][:64], g['Q_dds'][:64]) == img[:512]


def test_literal_vs_vectorised_comb():
    f = [k * 512e6 / 1024 for k in (3, 700, 129, 513)]
    a = [1.0, 0.5, 0.25, 0.8]
    I0, Q0, s0 = lut.freq_comb_lut_literal('yes', f, 512e6, 512e6 / 1024, a)
    I1, Q1, s1, _ = lut.freq_comb_lut('yes', f, 512e6, 512e6 / 1024, a)
    assert s0 == s1 and np.array_equal(I0, I1) and np.array_equal(Q0, Q1)


def test_dds_known_answer():
    # N=2^16 -> size 256, residual 3*7812.5, phase 0.3, echo='no'
    I, Q, sc, _ = lut.freq_comb_lut('no', [23437.5], 2e6, 7812.5, [1.], [0.3], 'no')
    assert sc == 0.9999850084539461
    assert list(I[:6]) == [31303, 30506, 29544, 28421, 27145, 25721]
    assert list(Q[:6]) == [9683, 11960, 14171, 16306, 18353, 20300]


def test_select_bins_known_answers():
    bins, res = lut.select_bins([412e6, 100023437.5, 255.5e6, 0], 512e6, 7812.5)
    assert bins == [412, 100, 256, 0]
    assert res == [0.0, 23437.5, -500000.0, 0.0]


def test_seed_1000_phases():
    ph = lut.random_phases(4)
    assert list(ph) == [4.106624480316831, 0.7226099352629045, 5.9708033309423225, 3.029697928700732]


# ------------------------------------------------------------------ fixed point (a15)
def test_fixed_against_reference_py3_outputs(golden_dir):
    g = np.load(os.path.join(golden_dir, 'utils_bin_py3.npz'))
    v = g['values']
    for nb, bp, key in ((12, 9, 'reinterpret_12_9'), (16, 13, 'reinterpret_16_13'), (18, 16, 'reinterpret_18_16')):
        assert np.array_equal(fixed.reinterpretBin(v, nb, bp), g[key])
        # scalar extractBin (py2 semantics) agrees with the reference's vectorised routine
        for x, want in zip(v[:300].tolist() + v[-50:].tolist(), list(g[key][:300]) + list(g[key][-50:])):
            assert fixed.extractBin(x, nb, bp) == want
    assert np.array_equal([fixed.bin12_9ToRad(x) for x in range(4096)], g['bin12_9ToRad'])
    assert np.array_equal([fixed.bin12_9ToDeg(x) for x in range(4096)], g['bin12_9ToDeg'])
    assert [fixed.binMask(n) for n in range(1, 33)] == g['binMask'].tolist()
    assert [fixed.castBin(x) for x in g['castBin_in']] == g['castBin_trunc_12_9'].tolist()
    assert [fixed.castBin(x, quantization='Round') for x in g['castBin_in']] == g['castBin_round_12_9'].tolist()
    assert [fixed.peakfit(1, 3, 2), fixed.peakfit(1, 2, 3), fixed.peakfit(-5., -9., -6.)] == g['peakfit'].tolist()


def test_fixed_known_answers():
    # SURVEY App. C; call sites lib/set_alpha.py:11, set_base_thresh.py:10, set_svf.py:24-25
    ka = {0x000: 0.0, 0x001: 0.001953125, 0x7FF: 3.998046875, 0x800: -4.0,
          0x801: -3.998046875, 0xFFF: -0.001953125, 0x923: -3.431640625}
    for k, v in ka.items():
        assert fixed.extractBin(k) == v
        assert fixed.bin12_9ToRad(k) == fixed.extractBin(k ^ 0x800)
    assert fixed.castBin(0.08) == 40
    assert fixed.castBin(0.08, quantization='Round') == 41
    assert fixed.castBin(-0.08) == 4056
    assert fixed.castBin(1., quantization='Round', nBits=16, binaryPoint=13) == 8192
    assert fixed.castBin(2 * np.sin(np.pi * 200 / 1e6), quantization='Round', nBits=18, binaryPoint=16) == 82
    assert fixed.castBin(1 / 0.7, quantization='Round', nBits=18, binaryPoint=16) == 93623
    assert fixed.castBin(1.5 / 512, quantization='Round') == 2
    assert fixed.castBin(2.5 / 512, quantization='Round') == 3      # py2 half away from zero
    assert fixed.peakfit(1, 3, 2) == 3.0416666666666665
    assert fixed.peakfit(1, 2, 3) == 2


# ------------------------------------------------------------------ control plane (a6,a7,a10)
def test_fir_quantisation_known_answers(golden_dir):
    t = np.load(os.path.join(golden_dir, 'fir_taps.npz'))
    assert control.fir_quantise(t['matched_30us']) == [160, 155, 150, 144, 140, 135, 130, 126, 122, 118, 115,
                                                        111, 107, 104, 101, 97, 94, 90, 88, 84, 82, 79, 77, 74, 72, 69]
    regs = [r[1] for r in control.fir_registers(t['matched_30us'])]
    assert regs == [0x9B0A0, 0x90096, 0x8708C, 0x7E082, 0x7607A, 0x6F073, 0x6806B, 0x61065, 0x5A05E,
                    0x54058, 0x4F052, 0x4A04D, 0x45048]
    assert control.fir_quantise(t['BlackmanFilter_250kHz']) == [0, 0, 3, 8, 17, 32, 53, 80, 111, 142, 172, 194, 206,
                                                                 206, 194, 172, 142, 111, 80, 53, 32, 17, 8, 3, 0, 0]


def _snap_raw(golden_dir):
    deg = np.load(os.path.join(golden_dir, 'ch_snap_0.npy'))
    raw = deg / control.SCALE_TO_ANGLE
    assert np.max(np.abs(raw - np.round(raw))) < 1e-6
    return deg, np.round(raw).astype(np.int64)


def test_snapshot_scaling_and_threshold(golden_dir):
    deg, raw = _snap_raw(golden_dir)
    assert len(raw) == 2048 and list(raw[:4]) == [18480, 18640, 19056, 18976]
    thr, med, p5 = control.threshold_from_phase(raw)
    assert thr == -5913
    assert abs(med - 18961.6) < 1e-6 and abs(p5 - 16596.16) < 1e-6
    assert abs(control.SCALE_TO_ANGLE * thr - (-41.356194)) < 1e-5
    assert abs(control.SCALE_TO_ANGLE * med - 132.619587) < 1e-5


def test_iq_snapshot_decode_matches_literal():
    rng = np.random.default_rng(3)
    buf = rng.integers(0, 256, 16 * 200, dtype=np.uint8).tobytes()
    I0, Q0 = control.decode_iq_snapshot_literal(buf)
    I1, Q1 = control.decode_iq_snapshot(buf)
    assert np.array_equal(I0, I1) and np.array_equal(Q0, Q1)


# ------------------------------------------------------------------ triggers (a11,a12)
def test_trigger_known_answers(golden_dir):
    deg, _ = _snap_raw(golden_dir)
    assert trigger.trigger_rolling_literal(deg, 20, 1000, 10.) == [153]
    assert trigger.trigger_rolling_literal(deg, 20, 1000, 15.) == [237]
    assert trigger.trigger_rolling_literal(deg, 20, 1000, 20.) == []
    assert trigger.trigger_block_literal(deg, 128, 10.) == [100, 309, 512, 731, 970, 1170, 1405, 1625]
    assert trigger.trigger_block_literal(deg, 128, 15.) == [237, 440, 662, 912, 1112, 1708]
    assert trigger.trigger_block_literal(deg, 128, 20.) == [239]


def test_trigger_fast_form_equals_literal(golden_dir):
    deg, _ = _snap_raw(golden_dir)
    rng = np.random.default_rng(0)
    x = np.concatenate([deg, deg[::-1] + rng.normal(0, 3, deg.size)])
    for thr in (8., 12., 15.):
        for L in (50, 300, 1000):
            lit = trigger.trigger_rolling_literal(x, 20, L, thr)
            d = trigger.rolling_candidates(x, 20)
            fast = trigger.greedy_holdoff(d > thr, 120, L, len(x), L)
            assert lit == fast


# ------------------------------------------------------------------ decode (a14,a16)
def test_word_layout_known_answer():
    w = decode.pack_word(37, 0x5A3, 0x5A1, 0x7F0, 123456)
    assert int(w) == 0x255A35A17F01E240
    ch, ts, base, peak, p1 = decode.unpack_fields([w])
    assert (ch[0], ts[0], base[0], peak[0], p1[0]) == (37, 123456, 0x7F0, 0x5A3, 0x5A1)
    wire = decode.words_to_wire(np.full(8192, w, dtype=np.uint64))
    assert wire[:4].hex() == '7f01e240' and wire[32768:32772].hex() == '255a35a1'
    assert np.array_equal(decode.wire_to_words(wire), np.full(8192, w, dtype=np.uint64))


def _mini_streams(seed=5, R=3, npix=7, secs=4, per_sec=300, cap=20):
    rng = np.random.default_rng(seed)
    streams = []
    for r in range(R):
        parts = []
        for s in range(secs + 1):                      # one second more than exptime -> ignored tail
            n = per_sec + int(rng.integers(0, 50))
            ch = rng.integers(0, npix + 2, n)          # some non-pixel addresses
            ch[rng.random(n) < 0.3] = 2                # hot pixel -> exceeds cap
            ts = np.sort(rng.integers(0, 10 ** 6, n))
            w = decode.pack_word(ch, rng.integers(0, 4096, n), rng.integers(0, 4096, n),
                                 rng.integers(0, 4096, n), ts)
            eos = np.array([0xFFFFFFFFFFFFFFFF if (s != 1 or r != 0) else 0xFF00000000000001], dtype=np.uint64)
            parts += [w, eos]
        streams.append(np.concatenate(parts))
    return streams, npix, secs, cap


def test_packetmaster_vectorised_equals_literal():
    streams, npix, secs, cap = _mini_streams()
    lit = decode.packetmaster_bin_literal(streams, npix, secs, cap)
    vec = decode.packetmaster_bin(streams, npix, secs, cap, want_lists=True)
    assert np.array_equal(lit['counts'], vec['counts'])
    for k in ('n_eos', 'n_corrupt_eos', 'n_nonpixel', 'n_ignored'):
        assert lit[k] == vec[k], k
    assert lit['n_corrupt_eos'] == 1 and lit['n_nonpixel'] > 0 and lit['n_ignored'] > 0
    assert vec['counts'].max() == cap - 1               # cap quirk: max_events-1 counted
    off = vec['list_offsets']
    npt = len(streams) * npix
    for (sec, pix), words in lit['lists'].items():
        k = sec * npt + pix
        assert list(vec['list_words'][off[k]:off[k + 1]]) == words


def test_merged_list_equals_literal_walk():
    """oracle.decode.merged_list against a word-by-word walk in PacketMaster's loop order (PacketMaster.c:304-342)."""
    streams, npix, secs, cap = _mini_streams(seed=4)
    R = len(streams)
    per_key = {}
    for r, st in enumerate(streams):
        sec = 0
        for w in st.tolist():
            adr = w >> 56
            if adr == 255:
                sec += 1
            elif adr < npix and sec < secs:
                per_key.setdefault(sec * R + r, []).append(w)
    ww, off = decode.merged_list(streams, npix, secs)
    assert off[-1] == ww.size == sum(len(v) for v in per_key.values())
    for k in range(secs * R):
        assert ww[off[k]:off[k + 1]].tolist() == per_key.get(k, [])


def test_packetmaster_c_core_equals_numpy():
    import ctypes
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_build', 'libpm_core.so')
    if not os.path.exists(so):
        import subprocess
        subprocess.check_call(['make', '-C', os.path.dirname(os.path.dirname(so))])
    lib = ctypes.CDLL(so)
    streams, npix, secs, cap = _mini_streams(seed=9)
    R = len(streams)
    counts = np.zeros((secs, R * npix), dtype=np.int32)
    hist = np.zeros((R * npix, 4096), dtype=np.uint32)
    st = (ctypes.c_int64 * 5)()
    for r, w in enumerate(streams):
        w = np.ascontiguousarray(w)
        rc = lib.pm_core_words(w.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(w.size), r, npix, R * npix, secs, cap,
                               counts.ctypes.data_as(ctypes.c_void_p), None, hist.ctypes.data_as(ctypes.c_void_p),
                               4096, 44, None, st)
        assert rc == 0
    vec = decode.packetmaster_bin(streams, npix, secs, cap)
    assert np.array_equal(counts, vec['counts'])
    assert (st[0], st[1], st[2], st[3]) == (vec['n_eos'], vec['n_corrupt_eos'], vec['n_nonpixel'], vec['n_ignored'])
    assert np.array_equal(hist, decode.pixel_field_hist(streams, npix, secs, 'peak'))


def test_read_pulses_wrap_and_hist():
    rng = np.random.default_rng(11)
    n = 2 ** 14
    ch = rng.integers(0, 256, n)
    w = decode.pack_word(ch, rng.integers(0, 4096, n), rng.integers(0, 4096, n), rng.integers(0, 4096, n),
                         rng.integers(0, 2 ** 20, n))
    b0 = (w & np.uint64(0xFFFFFFFF)).astype('>u4').tobytes()
    b1 = (w >> np.uint64(32)).astype('>u4').tobytes()
    out = decode.read_pulses([b0, b0], [b1, b1], [(100, 5000), (16000, 300)], sel_ch=3)
    assert out['total_counts'] == [4900, 684]
    idx = list(range(100, 5000)) + list(range(16000, n)) + list(range(0, 300))
    assert np.array_equal(out['channel_count'], np.bincount(ch[idx], minlength=256))
    lutb, edges = decode.deg_bin_lut()
    sel = [i for i in idx if ch[i] == 3]
    base_raw = ((w[sel] >> np.uint64(20)) & np.uint64(0xFFF)).astype(np.int64)
    h = np.bincount(lutb[base_raw][lutb[base_raw] < 40], minlength=40)
    assert np.array_equal(h, out['hgBase'])


def test_template_oracle_matches_reference_run(golden_dir):
    """oracle/template.py against the outputs of the reference's own MakeTemplate (lib/pulses.py:239-427), executed
    in the dev container by tests/golden/make_golden_analysis.py on the same synthetic iqpulses table."""
    import hashlib
    from oracle import template as otpl
    g = np.load(os.path.join(golden_dir, 'analysis_golden.npz'))
    n, seed = (int(v) for v in g['tpl_params'])
    I, Q = otpl.fake_pulses(n, seed=seed)
    assert hashlib.sha256(I.tobytes() + Q.tobytes()).hexdigest() == str(g['tpl_input_sha256'])
    r = otpl.make_template(I, Q)
    assert r['count'] == float(g['tpl_count']) and r['flag'] == int(g['tpl_flag']) and r['pstart'] == int(g['tpl_pstart'][0])
    assert np.array_equal(r['tPf'], g['tpl_phasetemplate'])
    assert np.array_equal(r['noise'], g['tpl_phasenoise'])
    assert np.array_equal(r['noiseidx'], g['tpl_phasenoiseidx'])
    assert hashlib.sha256(I.tobytes() + Q.tobytes()).hexdigest() == str(g['tpl_shifted_rows_sha256'])   # in-place shifts


def test_spectra_oracle_matches_reference_run(golden_dir):
    """oracle/spectra.py against the reference's own image_Worker methods (ArconsDashboard.py:1282-1384)."""
    from oracle import spectra as ospec
    g = np.load(os.path.join(golden_dir, 'analysis_golden.npz'))
    darray = g['iw_darray'].astype(np.int64)
    sel = [int(v) for v in g['iw_spectrum_pixel']]
    for sky, bt in ((False, 'wavelength'), (True, 'wavelength'), (True, 'energy')):
        r = ospec.image_worker(darray, bintype=bt, sky_subtraction=sky, spectrum_pixel=sel)
        key = 'iw_%d_%s_' % (int(sky), bt)
        assert np.array_equal(np.array(r['E']), g[key + 'E'])
        assert np.array_equal(np.array(r['medians']), g[key + 'medians'])
        assert np.array_equal(r['pc'], g[key + 'pc'])
        ref_me = g[key + 'me']
        ok = np.isfinite(ref_me)
        assert np.array_equal(r['me'][ok], ref_me[ok]) and np.array_equal(np.isfinite(r['me']), ok)
        assert np.array_equal(np.array(r['totalcounts']), g[key + 'totalcounts'])
        assert np.array_equal(np.array(r['SNR'], dtype=np.float64), g[key + 'SNR'])
        assert r['integrated_SNR'] == float(g[key + 'integrated_SNR'])


def test_lut_oracle_matches_reference_run_multitone(golden_dir):
    """oracle/lut.py against the reference's own define_DAC_LUT / freqCombLUT / define_DDS_LUT / select_bins /
    write_LUTs (ROACH_Setup_DAC.py:396-557), executed in the dev container on a 12-tone comb at N = 2^12 with
    seed-1000 random phases and non-zero DDS phases (tests/golden/make_golden_refrun.py)."""
    g = np.load(os.path.join(golden_dir, 'refrun_golden.npz'))
    fs, N, lo = 512e6, int(g['setup_N']), float(g['setup_lo'])
    res = fs / N
    dac_freqs = [float(f) for f in g['setup_dac_freqs']]
    # define_DAC_LUT of the multi-tone GUI: mirror about LO, +fs if below LO, snap (py2 round), amplitudes from attens
    freqs_dac = lut.dac_freqs_multi(dac_freqs, lo, res, fs)
    assert np.array_equal(np.array(freqs_dac), g['setup_freqs_dac'])
    amps = lut.dac_amplitudes(g['setup_attens'])
    I, Q, scale, _ = lut.freq_comb_lut('yes', freqs_dac, fs, res, amps)
    assert scale == float(g['setup_scale_factor'])
    assert np.array_equal(I, g['setup_I_dac']) and np.array_equal(Q, g['setup_Q_dac'])
    # define_DDS_LUT + select_bins (on the ORIGINAL frequency list, :488-496)
    fd = lut.dds_freqs(dac_freqs, lo, fs, res)
    bins, resid = lut.select_bins(fd, fs, res)
    assert np.array_equal(np.array(bins), g['setup_bins'])
    I_dds, Q_dds, _ = lut.define_dds_lut(resid, fs, res, list(g['setup_dds_phase']))
    assert np.array_equal(I_dds, g['setup_I_dds']) and np.array_equal(Q_dds, g['setup_Q_dds'])
    assert np.array_equal(np.frombuffer(lut.pack_dram(I, Q, I_dds, Q_dds), dtype=np.uint8), g['setup_dram'])


def test_control_oracle_matches_reference_run(golden_dir):
    """loadFIRcoeffs / loadIQcenters / loadThresholds of ROACH_Pulses.py executed in the dev container."""
    g = np.load(os.path.join(golden_dir, 'refrun_golden.npz'))
    taps = np.load(os.path.join(golden_dir, 'fir_taps.npz'))['matched_30us']
    regs = control.fir_registers(taps)
    zero = control.fir_registers([0.] * 26)
    for ch, rr in enumerate((regs, zero, regs)):                 # channel 1 is deleted (zeroChannels)
        for n, (name, val, packed) in enumerate(rr):
            assert str(g['fir_reg_names'][ch * 13 + n]) == name
            assert bytes(g['fir_reg_bytes'][ch * 13 + n]) == packed
    assert list(g['fir_load_coeff'][:4]) == [1, 0, 1, 0] and list(g['fir_load_coeff'][26:28]) == [3, 2]
    assert bytes(g['fir_inactive_bytes']) == b'\x00\x00\x00\x00'
    for c, w in zip(g['iq_centers_in'], g['iq_center_writes']):
        assert control.iq_center_word(complex(c))[0] == int(w)
    for ch in range(2):
        thr, med, _ = control.threshold_from_phase(g['thr_raw_phase'][ch])
        assert thr == int(g['thr_capture_threshold'][ch])
        assert control.SCALE_TO_ANGLE * thr == g['thr_thresholds_deg'][ch]
        assert control.SCALE_TO_ANGLE * med == g['thr_medians_deg'][ch]


def test_read_pulses_oracle_matches_reference_run(golden_dir):
    """oracle/decode.read_pulses against the reference's own readPulses (ROACH_Pulses.py:782-889): 10 steps,
    two ring wraps, channel 3 selected."""
    g = np.load(os.path.join(golden_dir, 'refrun_golden.npz'))
    w = g['rp_words']
    b0 = (w & np.uint64(0xFFFFFFFF)).astype('>u4').tobytes()
    b1 = (w >> np.uint64(32)).astype('>u4').tobytes()
    pairs = [tuple(int(v) for v in p) for p in g['rp_pairs']]
    r = decode.read_pulses([b0] * len(pairs), [b1] * len(pairs), pairs, sel_ch=3)
    assert np.array_equal(r['channel_count'], g['rp_channel_count'])
    assert np.array_equal(r['hgBase'], g['rp_hgBase']) and np.array_equal(r['hgPeak'], g['rp_hgPeak'])
    assert np.array_equal(r['hgPeakSubBase'], g['rp_hgPeakSubBase'])
    assert np.array_equal(r['peak_deg'], g['rp_peaksCh_deg']) and np.array_equal(r['times'], g['rp_timesCh'])


def test_trigger_oracle_matches_reference_run(golden_dir):
    """oracle/trigger.py and the I/Q snapshot decode against the reference's own trigger loop
    (pulse_triggering_v2.py:102-174, executed in the dev container)."""
    g = np.load(os.path.join(golden_dir, 'refrun_golden.npz'))
    x = g['trig_phase']
    assert trigger.trigger_rolling_literal(x, 20, 1000, 25.0) == [int(v) for v in g['trig_hits_20_1000']]
    assert trigger.trigger_rolling_literal(x, 10, 300, 15.0) == [int(v) for v in g['trig_hits_10_300']]
    I, Q = control.decode_iq_snapshot(g['trig_iq_snapshot'].tobytes())
    assert np.array_equal(np.stack([I, Q], axis=1), g['trig_iq_decoded'])


def _utils_cases(g, tag):
    vals = [int(v) for v in g['ub_values']]
    fl = [float(v) for v in g['ub_floats']]
    for key in g.files:
        if not key.startswith(tag + '_'):
            continue
        parts = key[len(tag) + 1:].split('_')
        yield key, parts, vals, fl


def test_fixed_oracle_matches_reference_run(golden_dir):
    """oracle/fixed.py against the reference's Utils/bin.py and Utils/binTools.py executed with Python-2 division in the
    dev container (under Python 3 the modules import but extractBin / castBin are silently wrong)."""
    g = np.load(os.path.join(golden_dir, 'refrun_golden.npz'))
    n_checked = 0
    for tag in ('bin', 'binTools'):
        for key, parts, vals, fl in _utils_cases(g, tag):
            if parts[0] == 'extract':
                nb, bp, after = int(parts[1]), int(parts[2]), int(parts[3])
                fmt = parts[4] if len(parts) > 4 else 'rad'
                got = np.array([fixed.extractBin(v, nb, bp, after, fmt) for v in vals], dtype=np.float64)
            elif parts[0] == 'cast':
                nb, bp, q, fmt = int(parts[1]), int(parts[2]), parts[3], parts[4]
                got = np.array([fixed.castBin(v, nb, bp, q, fmt) for v in fl], dtype=np.float64)
            elif parts[0] in ('binMask', 'bitmask'):
                got = np.array([fixed.binMask(k) for k in range(1, 40)], dtype=np.uint64)
            elif parts[0] == 'peakfit':
                got = np.array([fixed.peakfit(*y) for y in g['ub_peakfit_in']])
            elif parts[0] == 'bin12':
                got = np.array([fixed.bin12_9ToRad(v & 0xFFF) for v in vals], dtype=np.float64)
            else:
                continue
            assert np.array_equal(got, g[key]), key
            n_checked += 1
    assert n_checked >= 40


def test_block_trigger_oracle_matches_reference_run(golden_dir):
    """oracle/trigger.trigger_block_literal against the reference's own block-mean trigger loop
    (pulse_triggering.py:104-208, executed in the dev container)."""
    g = np.load(os.path.join(golden_dir, 'refrun_golden.npz'))
    for tag in 'abc':
        A, thr = g['btrig_%s_params' % tag]
        hits = trigger.trigger_block_literal(g['btrig_%s_phase' % tag], int(A), float(thr))
        assert hits == [int(v) for v in g['btrig_%s_hits' % tag]], tag
    assert len(g['btrig_c_hits']) > 5


def test_product_utils_scalar_paths_match_reference_run(golden_dir):
    """The host-side scalar paths of mkids_sdr_b200/Utils/{bin,binTools}.py (drop-in for `from Utils.bin import *`,
    lib/set_alpha.py:7) against the reference modules executed with Python-2 division."""
    from mkids_sdr_b200.Utils import bin as pbin, binTools as ptools
    g = np.load(os.path.join(golden_dir, 'refrun_golden.npz'))
    n_checked = 0
    for tag, mod in (('bin', pbin), ('binTools', ptools)):
        for key, parts, vals, fl in _utils_cases(g, tag):
            if parts[0] == 'extract':
                nb, bp, after = int(parts[1]), int(parts[2]), int(parts[3])
                fmt = parts[4] if len(parts) > 4 else 'rad'
                got = np.array([mod.extractBin(v, nb, bp, after, fmt) for v in vals], dtype=np.float64)
            elif parts[0] == 'cast':
                nb, bp, q, fmt = int(parts[1]), int(parts[2]), parts[3], parts[4]
                got = np.array([mod.castBin(v, nb, bp, q, fmt) for v in fl], dtype=np.float64)
            elif parts[0] == 'peakfit':
                got = np.array([mod.peakfit(*y) for y in g['ub_peakfit_in']])
            elif parts[0] == 'bin12':
                got = np.array([mod.bin12_9ToRad(v & 0xFFF) for v in vals], dtype=np.float64)
            elif parts[0] == 'binMask':
                got = np.array([mod.binMask(k) for k in range(1, 40)], dtype=np.uint64)
            elif parts[0] == 'bitmask':
                got = np.array([mod.bitmask(k) for k in range(1, 40)], dtype=np.uint64)
            else:
                continue
            assert np.array_equal(got, g[key]), key
            n_checked += 1
    assert n_checked >= 40


def test_contsnapshot_trigger_oracle_matches_reference_run(golden_dir):
    """oracle/trigger.trigger_contsnapshot_literal against the trigger loop of AppForm.contsnapshot
    (ROACH_Pulses.py:625-725) executed in the dev container."""
    g = np.load(os.path.join(golden_dir, 'refrun_golden.npz'))
    for tag in 'abc':
        A, thr = g['ctrig_%s_params' % tag]
        hits = trigger.trigger_contsnapshot_literal(g['ctrig_%s_phase' % tag], int(A), float(thr))
        assert hits == [int(v) for v in g['ctrig_%s_hits' % tag]], tag
        assert len(hits) > 5


def test_lut_oracle_matches_reference_run_at_full_size(golden_dir):
    """Row a1 at the BASELINE size: the oracle's vectorised freqCombLUT (256 tones x 2^19 samples) against the sha256 of
    the reference's own literal freqCombLUT, executed once from the reference tree (make_golden_refrun_fullsize.py)."""
    import hashlib
    import json
    g = json.load(open(os.path.join(golden_dir, 'refrun_fullsize_lut.json')))
    N, T, FS = g['N'], g['T'], 512e6
    k = np.sort(np.random.default_rng(0).choice(np.arange(-N // 2 + 1, N // 2), T, replace=False))
    f = (k % N) * FS / N
    amps = lut.dac_amplitudes(np.random.default_rng(1).integers(0, 20, T))
    I, Q, scale, _ = lut.freq_comb_lut('yes', list(f), FS, FS / N, amps)
    assert repr(float(scale)) == g['scale_factor']
    assert hashlib.sha256(np.asarray(I).astype('<i2').tobytes()).hexdigest() == g['sha256_I']
    assert hashlib.sha256(np.asarray(Q).astype('<i2').tobytes()).hexdigest() == g['sha256_Q']
    assert [int(v) for v in I[:8]] == g['I_first'] and int(np.asarray(I).sum()) == g['I_sum']


def test_adc_pack12_known_answer():
    """Host-link format of a 12-bit ADC stream (oracle/channelizer.py header): I = 0x123, Q = -2 (0xFFE) -> group 0xFFE123,
    bytes 23 E1 FF; four samples are three little-endian 32-bit words."""
    from oracle import channelizer as oc
    iq = np.array([[0x123, -2], [-2048, 2047], [0, -1], [5, 6]], dtype=np.int16)
    pk = oc.adc_pack12(iq)
    assert pk.dtype == np.uint8 and pk.shape == (12,)
    assert pk[:3].tolist() == [0x23, 0xE1, 0xFF]
    assert pk[3:6].tolist() == [0x00, 0xF8, 0x7F]                 # I = 0x800, Q = 0x7FF -> 0x7FF800
    assert np.array_equal(oc.adc_unpack12(pk), iq)
    with pytest.raises(ValueError):
        oc.adc_pack12(np.array([[2048, 0]], dtype=np.int16))
