"""K4 + K5 parity ON THE BENCHMARKED CONFIGURATION: N_lut = 2^19, 253 driven tones per board
(about 28 ADC counts per tone under sigma = 8 noise), matched_30us FIR, thresholds from `derive_thresholds`
(loadThresholds on a pulse-free stretch), pulse rate 1000 /s/channel, 2^22 samples of the first boards of the bench
input (same seeds as bench.py rank 0), against oracle.channelizer (float64 model = the parity definition of the absent
firmware) and its integer detection.

What is asserted, with the MEASURED numbers printed (pytest -s) and recorded in DESIGN.md section 2:
  * max |dphase| <= 1e-5 rad on driven channels (north_star tolerance), fast path == f32-hook path bit for bit;
  * Fix16_13 raw samples: |d| <= 1 LSB everywhere, fraction of +-1 LSB samples <= RAW_FLIP_MAX;
  * photon words: emission is bit-exact GIVEN the phase rows (oracle detection on the GPU's rows == GPU words);
    against the float64 model end to end every trigger that differs is explained by a +-1 LSB flip: its trigger
    quantity M*raw[t] - sum(raw[t-M..t-1]) lies within 2*M LSB of M*thr in the oracle's rows, or it lies in the
    hold-off shadow (L rows) of such a marginal trigger; words of common triggers differ only by +-1 code per field.
"""
import numpy as np
import pytest

from oracle import channelizer as oc

pytestmark = pytest.mark.gpu

N_LUT = 2 ** 19
N_ACTIVE = 253
RAW_FLIP_MAX = 0.02          # measured 0.4 % (see the printed line); generous margin
TRIG_DIFF_MAX = 0.01         # share of triggers allowed to differ (all of them must be explained)


@pytest.fixture(scope='module')
def ctx():
    from mkids_sdr_b200 import _lib
    return _lib.default_context(0)


@pytest.fixture(scope='module')
def bench_setup(ctx):
    """Two boards configured exactly as bench.py configures rank 0 (seed0 = 42), thresholds derived the same way."""
    from mkids_sdr_b200.chain import ReadoutChain
    from mkids_sdr_b200.channelizer import synth_adc
    B = 2
    chain, boards = ReadoutChain.synthetic(B, N_LUT, N_ACTIVE, seed0=42, ctx=ctx, exptime=4)
    thr = chain.derive_thresholds(boards)
    n = 2 ** 22
    tone_bins = np.stack([bd['tone_bins'] for bd in boards])
    iq = synth_adc(B, n, tone_bins, n_lut=N_LUT, pulse_rate=1000.0, seed=1000, ctx=ctx)
    cfgs = [oc.ChanConfig(bd['bins'], bd['I_dds'], bd['Q_dds'], chain.fir_int, thresholds=thr[b],
                          zero_ch=bd['zero_ch'].astype(bool), M=20, L=1000, W=32) for b, bd in enumerate(boards)]
    return chain, boards, thr, iq, cfgs, n


def _explain(t, c, raw, cfg, marginal_rows):
    """Is a differing trigger (row t, channel c) explained by +-1 LSB flips?  q = M*raw[t] - sum(raw[t-M..t-1])."""
    M, L = cfg.M, cfg.L
    q = M * int(raw[t, c]) - int(raw[t - M:t, c].sum())
    if abs(q - M * int(cfg.thresholds[c])) <= 2 * M:
        marginal_rows.setdefault(c, []).append(t)
        return 'marginal'
    return None


def test_bench_config_phase_and_words(ctx, bench_setup):
    chain, boards, thr, iq, cfgs, n = bench_setup
    B = len(boards)
    T = n // 512
    ch = chain.chan
    # ---- run 1: the benchmarked fast path (no f32 hook): raw phase rows + photon words
    ch.reset()
    words, raw_gpu = ch.process(iq, detect=True, want_phase=True)
    # ---- run 2: f32 hook (slow-path template) for the phase tolerance; its int16 rows must equal run 1
    f32 = ctx.alloc(B * T * 256 * 4)
    ch.set_f32_phase_out(f32)
    ch.reset()
    _, raw_gpu2 = ch.process(iq, detect=False, want_phase=True)
    ch.set_f32_phase_out(None)
    ch.reset()
    assert np.array_equal(raw_gpu, raw_gpu2), 'fast path and f32-hook path differ'
    ph_gpu = f32.download(np.float32).reshape(B, T, 256).astype(np.float64)
    f32.free()
    tot_words = tot_trig_diff = tot_common = tot_word_diff = 0
    for b in range(B):
        cfg = cfgs[b]
        ph_ref, raw_ref, w = oc.channelize_phase(iq[b], cfg, return_w=True)
        act = ~cfg.zero_ch
        amp = np.abs(w)[64:, act]
        d = np.angle(np.exp(1j * (ph_gpu[b] - ph_ref)))[64:, act]
        dmax = float(np.abs(d).max())
        dq = (raw_gpu[b].astype(np.int64) - raw_ref.astype(np.int64))[64:, act]
        dq = (dq + 25736) % 51472 - 25736
        flips = float((dq != 0).mean())
        print('\n[bench-config parity] board %d: max|dphase| = %.3e rad (tolerance 1e-5), rms %.3e; |w| min %.1f median %.1f; '
              '+-1 LSB raw samples %.4f %% of %d (max |d| = %d LSB)'
              % (b, dmax, float(np.sqrt((d ** 2).mean())), float(amp.min()), float(np.median(amp)), 100 * flips, dq.size,
                 int(np.abs(dq).max())))
        assert dmax <= 1e-5, dmax
        assert np.abs(dq).max() <= 1
        assert flips <= RAW_FLIP_MAX, flips
        assert np.array_equal(raw_gpu[b][64:, ~act], raw_ref[64:, ~act])          # idle channels: identical
        # ---- K5 given the phase rows: bit-exact
        ref_on_gpu_rows = oc.detect_emit(raw_gpu[b], cfg, 0, np.zeros(256, np.int64), T - 64 - cfg.M)
        assert np.array_equal(words[b], np.array(ref_on_gpu_rows, dtype=np.uint64)), 'emission not exact on identical rows'
        # ---- end to end against the float64 model
        ref = oc.detect_emit(raw_ref, cfg, 0, np.zeros(256, np.int64), T - 64 - cfg.M)
        key = lambda x: (int(x) >> 56, int(x) & 0xFFFFF)
        ga = {key(x): int(x) for x in words[b] if int(x) != 2 ** 64 - 1}
        rb = {key(x): int(x) for x in ref if int(x) != 2 ** 64 - 1}
        only = sorted((set(ga) ^ set(rb)), key=lambda k: (k[0], k[1]))
        marginal = {}
        unexplained = []
        for (c, ts) in only:                    # ts == absolute row here (less than one second of stream)
            raw_side = raw_ref if (c, ts) in rb else raw_gpu[b].astype(np.int64)
            if _explain(ts, c, raw_side, cfg, marginal) is None:
                unexplained.append((c, ts))
        # cascades: a trigger inside the hold-off shadow of a marginal one of the same channel
        still = [(c, ts) for (c, ts) in unexplained
                 if not any(0 < abs(ts - m) <= cfg.L for m in marginal.get(c, []))]
        common = set(ga) & set(rb)
        wdiff = 0
        for kk in common:
            if ga[kk] != rb[kk]:
                wdiff += 1
                for sh in (44, 32, 20):         # peak, p1, baseline codes: at most one code apart
                    assert abs(((ga[kk] >> sh) & 0xFFF) - ((rb[kk] >> sh) & 0xFFF)) <= 1, (hex(ga[kk]), hex(rb[kk]))
        print('[bench-config parity] board %d: %d oracle words, %d GPU words; triggers only on one side: %d '
              '(%d marginal within 2M LSB of M*thr, %d in their hold-off shadow, %d unexplained); common triggers %d, '
              'of which %d words differ by one code in a field'
              % (b, len(rb), len(ga), len(only), sum(len(v) for v in marginal.values()), len(unexplained) - len(still),
                 len(still), len(common), wdiff))
        assert len(rb) > 500
        assert not still, still
        assert len(only) <= TRIG_DIFF_MAX * len(rb), (len(only), len(rb))
        tot_words += len(rb); tot_trig_diff += len(only); tot_common += len(common); tot_word_diff += wdiff
    print('[bench-config parity] total: %d words, %d differing triggers (%.3f %%), %d of %d common words differ in a code (%.3f %%)'
          % (tot_words, tot_trig_diff, 100.0 * tot_trig_diff / tot_words, tot_word_diff, tot_common,
             100.0 * tot_word_diff / max(tot_common, 1)))
