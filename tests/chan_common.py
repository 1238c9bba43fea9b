"""Shared set-up for the channelizer tests: a board configuration built with the ORACLE's
restatement of the reference control plane, fed identically to the oracle model and the GPU."""
import os

import numpy as np

from oracle import channelizer as oc
from oracle import control, lut as olut

FS = 512e6


def board_config(n_lut=2 ** 16, n_tones=32, seed=0, M=20, L=100, W=32, thr=-2500, fir='matched_30us',
                 centers=False):
    """Returns (oracle ChanConfig, tone fine bins [n_tones])."""
    res = FS / n_lut
    rng = np.random.default_rng(seed)
    ks = np.sort(rng.choice(np.arange(-n_lut // 2 + 2000, n_lut // 2 - 2000), n_tones, replace=False))
    freqs = [float(k) * res for k in ks]
    freqs_pos = [f if f >= 0 else f + FS for f in freqs]
    bins, resid = olut.select_bins(freqs_pos + [0.0] * (256 - n_tones), FS, res)
    phase = [0.] * 256
    I_dds, Q_dds, _ = olut.define_dds_lut(resid, FS, res, phase)
    taps = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'fir_taps.npz'))[fir]
    cfg = oc.ChanConfig(bins, I_dds, Q_dds, control.fir_quantise(taps), M=M, L=L, W=W,
                        thresholds=np.full(256, thr))
    cfg.zero_ch[n_tones:] = True
    if centers:
        cfg.centers_i[:n_tones] = rng.integers(-3, 4, n_tones)
        cfg.centers_q[:n_tones] = rng.integers(-3, 4, n_tones)
    return cfg, ks % n_lut


def make_gpu_channelizer(cfgs, ctx):
    from mkids_sdr_b200.channelizer import Channelizer
    c0 = cfgs[0]
    ch = Channelizer(len(cfgs), c0.N_lut, c0.M, c0.L, c0.W, ctx=ctx)
    ch.set_fir(c0.fir_int)
    ch.set_window(c0.h)
    for b, cfg in enumerate(cfgs):
        ch.set_board(b, cfg.bins, cfg.I_dds, cfg.Q_dds, cfg.zero_ch, cfg.centers_i, cfg.centers_q, cfg.thresholds)
    return ch


def compare_words_with_model(words_gpu, raw_gpu, raw_ref, cfg, T, t_abs0=0):
    """GPU photon words of one board against the float64 model end to end.  Emission is bit-exact given the phase rows;
    against the model's own rows a trigger can differ only through a +-1 LSB rounding flip: either its trigger quantity
    q = M*raw[t] - sum(raw[t-M..t-1]) lies within 2*M LSB of M*thr on the side that fired ("marginal"), or it lies in the
    hold-off shadow (L rows) of a marginal trigger of the same channel.  Returns a dict of counts; `unexplained` must be
    empty.  (timestamps are rows here: use less than one second of stream)"""
    ref = oc.detect_emit(raw_ref, cfg, t_abs0, np.zeros(256, np.int64), T - 64 - cfg.M)
    key = lambda x: (int(x) >> 56, int(x) & 0xFFFFF)
    ga = {key(x): int(x) for x in words_gpu if int(x) != 2 ** 64 - 1}
    rb = {key(x): int(x) for x in ref if int(x) != 2 ** 64 - 1}
    only = sorted(set(ga) ^ set(rb))
    M, L = cfg.M, cfg.L
    marginal, rest = {}, []
    for (c, ts) in only:
        raw = np.asarray(raw_ref if (c, ts) in rb else raw_gpu, dtype=np.int64)
        q = M * int(raw[ts, c]) - int(raw[ts - M:ts, c].sum())
        if abs(q - M * int(cfg.thresholds[c])) <= 2 * M:
            marginal.setdefault(c, []).append(ts)
        else:
            rest.append((c, ts))
    unexplained = [(c, ts) for (c, ts) in rest if not any(0 < abs(ts - m) <= L for m in marginal.get(c, []))]
    common = set(ga) & set(rb)
    wdiff = 0
    for kk in common:
        if ga[kk] != rb[kk]:
            wdiff += 1
            for sh in (44, 32, 20):             # peak, p1, baseline codes: at most one code apart
                assert abs(((ga[kk] >> sh) & 0xFFF) - ((rb[kk] >> sh) & 0xFFF)) <= 1, (hex(ga[kk]), hex(rb[kk]))
    return dict(n_ref=len(rb), n_gpu=len(ga), only=len(only), marginal=sum(len(v) for v in marginal.values()),
                shadow=len(rest) - len(unexplained), unexplained=unexplained, common=len(common), word_diff=wdiff)
