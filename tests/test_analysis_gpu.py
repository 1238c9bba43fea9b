"""GPU parity of the analysis kernels around the hot path: snapshot decoders,
float64 software triggers (bit-identical hit lists vs the literal NumPy loops) and loadThresholds."""
import os

import numpy as np
import pytest

from oracle import control, trigger as otrig

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ctx():
    from mkids_sdr_b200 import _lib
    return _lib.default_context(0)


def _pulse_stream(rng, n, rate=0.002, depth=(20., 120.), sigma=4.0, offset=0.0):
    x = rng.normal(offset, sigma, n)
    t = np.arange(n)
    for t0 in np.nonzero(rng.random(n) < rate)[0]:
        x[t0:] -= rng.uniform(*depth) * np.exp(-(t[t0:] - t0) / 30.0)
    return x


def test_rolling_trigger_matches_literal_loop(ctx, golden_dir):
    from mkids_sdr_b200 import triggers
    rng = np.random.default_rng(3)
    streams = np.stack([_pulse_stream(rng, 16384) for _ in range(6)])
    for M, L, thr in ((20, 1000, 25.), (10, 1000, 15.), (20, 300, 100.), (7, 50, 12.5)):
        got = triggers.trigger_rolling(streams, meanlength=M, pulselength=L, phase_threshold=thr, ctx=ctx)
        for s in range(streams.shape[0]):
            assert got[s] == otrig.trigger_rolling_literal(streams[s], M, L, thr), (M, L, thr, s)
    assert sum(len(g) for g in got) > 10
    # hit lists produced by the reference's own script loop (tests/golden/refrun_golden.npz)
    g = np.load(os.path.join(golden_dir, 'refrun_golden.npz'))
    assert triggers.trigger_rolling(g['trig_phase'], 20, 1000, 25.0, ctx=ctx) == [int(v) for v in g['trig_hits_20_1000']]
    assert triggers.trigger_rolling(g['trig_phase'], 10, 300, 15.0, ctx=ctx) == [int(v) for v in g['trig_hits_10_300']]
    Ig, Qg = triggers.decode_iq_snapshot(g['trig_iq_snapshot'].tobytes(), ctx=ctx)
    assert np.array_equal(np.stack([Ig, Qg], axis=1), g['trig_iq_decoded'])
    # the reference's own snapshot (ch_snap_0.txt, degrees)
    deg = np.load(os.path.join(golden_dir, 'ch_snap_0.npy'))
    for thr in (2.0, 5.0, 15.0):
        assert triggers.trigger_rolling(deg, 20, 300, thr, ctx=ctx) == otrig.trigger_rolling_literal(deg, 20, 300, thr)


def test_rolling_trigger_numpy16_summation_order(ctx):
    """sum_order=1: the plain left-to-right sum NumPy 1.6 (the reference's EPD 7.3) used for np.mean."""
    from mkids_sdr_b200 import triggers
    rng = np.random.default_rng(8)
    x = _pulse_stream(rng, 8192)
    M, L, thr = 20, 500, 20.
    hits, bob = [], 100 + M
    while bob < len(x):
        if bob + L > len(x):
            break
        acc = 0.0
        for v in x[bob - M:bob]:
            acc += v
        if abs(acc / M - x[bob]) > thr:
            hits.append(bob); bob += L
        else:
            bob += 1
    assert triggers.trigger_rolling(x, M, L, thr, numpy16_sum=True, ctx=ctx) == hits


def test_block_triggers_match_literal_loops(ctx):
    from mkids_sdr_b200 import triggers
    rng = np.random.default_rng(5)
    streams = np.stack([_pulse_stream(rng, 20000, offset=o) for o in (0.0, 3.0, -2.0, 150.0)])
    for A, thr in ((128, 25.), (64, 15.), (100, 40.), (300, 30.)):
        got = triggers.trigger_block(streams, averagelength=A, phase_threshold=thr, ctx=ctx)
        for s in range(streams.shape[0]):
            assert got[s] == otrig.trigger_block_literal(streams[s], A, thr), (A, thr, s)
    # hit lists produced by the reference's own loops (tests/golden/refrun_golden.npz)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'refrun_golden.npz'))
    for tag in 'abc':
        A, thr = g['btrig_%s_params' % tag]
        assert triggers.trigger_block(g['btrig_%s_phase' % tag], int(A), float(thr), ctx=ctx) == [int(v) for v in g['btrig_%s_hits' % tag]]
        A, thr = g['ctrig_%s_params' % tag]
        assert triggers.trigger_contsnapshot(g['ctrig_%s_phase' % tag], int(A), float(thr), ctx=ctx) == \
            [int(v) for v in g['ctrig_%s_hits' % tag]]
    big = _pulse_stream(rng, 2 ** 18, rate=0.0005)
    for A in (64, 32, 256):
        assert triggers.trigger_contsnapshot(big, A, 25., ctx=ctx) == otrig.trigger_contsnapshot_literal(big, A, 25.)


def test_iq_snapshot_and_phase_model(ctx):
    from mkids_sdr_b200 import triggers
    rng = np.random.default_rng(11)
    buf = rng.integers(0, 256, 4 * 16384, dtype=np.uint8).tobytes()       # L_IQ = 16384 words of 4 bytes
    I, Q = triggers.decode_iq_snapshot(buf, ctx=ctx)
    I0, Q0 = control.decode_iq_snapshot(buf)
    assert np.array_equal(I, I0) and np.array_equal(Q, Q0)
    Il, Ql = control.decode_iq_snapshot_literal(buf[:1024])
    assert np.array_equal(I[:128], Il) and np.array_equal(Q[:128], Ql)
    deg = triggers.phase_deg_from_iq(I, Q, 12.5, -3.0, ctx=ctx)
    ref = control.phase_deg_from_iq(I0, Q0, 12.5, -3.0)
    assert np.max(np.abs(deg - ref)) <= 1e-12            # float64 atan2: CUDA (<= 2 ulp) vs libm


def test_thresholds_from_device_phase_stream(ctx):
    from mkids_sdr_b200 import triggers
    rng = np.random.default_rng(2)
    B, rows, C, n, row0 = 2, 20480 + 96, 256, 20480, 64
    ph = np.empty((B, rows, C), dtype=np.int16)
    for b in range(B):
        for c in range(C):
            sig = rng.uniform(5, 400)
            x = rng.normal(rng.uniform(-3000, 3000), sig, rows)
            if c % 7 == 0:                                   # pulses: a long negative tail
                x -= np.abs(rng.normal(0, 6 * sig, rows)) * (rng.random(rows) < 0.03)
            ph[b, :, c] = np.clip(np.round(x), -32768, 32767)
    ph[0, :, 5] = 1234                                       # flat channel: outer edges widened by 0.5
    ph[1, :, 9] = np.where(np.arange(rows) % 2, -7, 8)       # two values only
    dev = ctx.to_device(ph)
    thr, med, p5 = triggers.thresholds_from_phase(dev, B, rows, n, row0=row0, ctx=ctx)
    for b in range(B):
        for c in range(C):
            t0, m0, q0 = control.threshold_from_phase(ph[b, row0:row0 + n, c].astype(np.int64))
            assert (thr[b, c], med[b, c], p5[b, c]) == (t0, m0, q0), (b, c)
    assert thr.min() >= -25736 and (thr < 0).sum() > 400


def test_longsnapshot_noise_spectrum(ctx):
    """ROACH_Pulses.py:521-537 with the reference's sizes: 2^20 samples, 100 averages of 10485-point FFTs."""
    from mkids_sdr_b200 import triggers
    rng = np.random.default_rng(4)
    n = 2 ** 20
    t = np.arange(n)
    x = 3.0 * np.sin(2 * np.pi * 0.0123 * t) + rng.normal(0, 2.0, n) + 40.0
    got, freqs = triggers.noise_spectrum(x, 100, 50.0, ctx=ctx)
    nper = n // 100
    ref = np.zeros(nper)
    for i in range(100):
        noise = np.abs(np.fft.fft(x[i * nper:(i + 1) * nper]))
        ref += 20 * (np.log10(noise / 50.0 / 1e-6))
    ref /= 100
    assert got.shape == (nper,) and nper == 10485
    assert np.max(np.abs(got - ref)) < 1e-8                     # dB; float64 DFT vs pocketfft
    assert np.array_equal(freqs, np.fft.fftfreq(nper))


@pytest.mark.parametrize('sky,bintype', [(False, 'wavelength'), (True, 'wavelength'), (True, 'energy')])
def test_image_worker_spectra_products(ctx, sky, bintype):
    """ArconsDashboard.py image_Worker: medians, sky subtraction, pc, mean energy, SNR -- bit-exact."""
    from oracle import spectra as ospec
    from mkids_sdr_b200.spectra import ImageWorker
    rng = np.random.default_rng(12)
    n_pix = 44 * 46
    darray = rng.poisson(rng.uniform(5, 60, (n_pix, 1)) * np.linspace(1.5, 0.5, 10)[None, :]).astype(np.uint32)
    darray[7] = 0                                              # an empty pixel
    iw = ImageWorker(44, 46, ctx=ctx)
    iw.bintype = bintype
    iw.setup_thread()
    iw.sky_subtraction = sky
    iw.spectrum_pixel = [3, 50, 51, 900, 2001]
    pc, me = iw.run(ctx.to_device(darray))
    ref = ospec.image_worker(darray, bintype=bintype, sky_subtraction=sky, spectrum_pixel=iw.spectrum_pixel)
    assert iw.E == ref['E']
    assert iw.medians == ref['medians']
    assert np.array_equal(pc, ref['pc'])
    ok = np.isfinite(ref['me'])
    assert np.array_equal(me[ok], ref['me'][ok]) and np.array_equal(np.isnan(me), np.isnan(ref['me']))
    assert iw.SNR == ref['SNR'] and iw.integrated_SNR == ref['integrated_SNR']
