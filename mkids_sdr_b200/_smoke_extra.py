"""Extra smoke checks (called from __graft_entry__.smoke): LUT synthesis and the full chain on
small inputs, each checked against the oracle."""
import numpy as np


def run(ctx):
    from oracle import channelizer as oc
    from oracle import decode as odec
    from oracle import lut as olut
    from . import lut
    from .chain import ReadoutChain
    from .channelizer import synth_adc
    # K1: 64-tone comb at N = 2^14 must equal the NumPy restatement of freqCombLUT
    N, T, FS = 2 ** 14, 64, 512e6
    k = np.sort(np.random.default_rng(0).choice(np.arange(1, N // 2), T, replace=False))
    f = k * FS / N
    I, Q, sc, ph = lut.comb_lut(f, FS, N, np.ones(T), ctx=ctx)
    Io, Qo, so, _ = olut.freq_comb_lut('yes', list(f), FS, FS / N, [1.] * T)
    assert sc[0] == so and np.array_equal(I[0], Io) and np.array_equal(Q[0], Qo), 'comb LUT differs from the oracle'
    # K4+K5+K6: one board, 2^19 samples
    chain, boards = ReadoutChain.synthetic(1, 2 ** 16, 32, seed0=3, threshold=-2500, holdoff=100, ctx=ctx, exptime=2)
    n = 2 ** 19
    iq = synth_adc(1, n, boards[0]['tone_bins'][None, :], n_lut=2 ** 16, pulse_rate=5000., seed=5, ctx=ctx)
    words_host = np.zeros((1, chain.chan.words_capacity(n)), dtype=np.uint64)
    nw = chain.process(iq, words_host=words_host)
    _, ph16 = (lambda c: (c.reset(), c.process(iq, detect=False, want_phase=True))[1])(chain.chan)
    cfg = oc.ChanConfig(boards[0]['bins'], boards[0]['I_dds'], boards[0]['Q_dds'], chain.fir_int,
                        thresholds=np.full(256, -2500), zero_ch=boards[0]['zero_ch'].astype(bool), M=20, L=100, W=32)
    ref = oc.detect_emit(ph16[0], cfg, 0, np.zeros(256, np.int64), n // 512 - 64 - cfg.M)
    assert nw[0] == len(ref) and np.array_equal(words_host[0, :nw[0]], np.array(ref, dtype=np.uint64)), \
        'photon words differ from the oracle'
    res = odec.packetmaster_bin([np.array(ref, dtype=np.uint64)], 253, 2)
    assert np.array_equal(chain.dec.counts(), res['counts']), 'per-pixel counts differ from the oracle'
