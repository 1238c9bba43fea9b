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
