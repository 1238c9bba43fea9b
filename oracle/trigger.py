"""Oracle: software pulse triggers (TEST INFRASTRUCTURE).

Restates the only CPU models of phase -> trigger the reference holds:
  * rolling-mean trigger  pulse_triggering_v2.py:104-174 (= pulse_triggering_IQ.py:160-200)
  * block-mean trigger    pulse_triggering.py:114-208 and contsnapshot
                          ROACH_Pulses.py:614-725
Both literal (np.mean per sample, the reference's hot loop) and an equivalent
fast form (candidate mask + greedy hold-off) used at large sizes.
"""
import numpy as np


def trigger_rolling_literal(phasevalues, meanlength=20, pulselength=1000, phase_threshold=25.,
                            pre=100):
    """pulse_triggering_v2.py:104-174 verbatim control flow.  Returns hit indices."""
    x = np.asarray(phasevalues, dtype=np.float64)
    n = len(x)
    hits = []
    bob = pre + meanlength                                   # :105
    while bob < n:                                           # :109
        if bob + pulselength > n:                            # :112
            break
        rollingaverage = np.mean(x[bob - meanlength:bob])    # :115
        if abs(rollingaverage - x[bob]) > phase_threshold:   # :119
            hits.append(bob)
            bob = bob + pulselength                          # :168
        else:
            bob = bob + 1                                    # :171
    return hits


def rolling_candidates(x, meanlength):
    """cand[t] = |mean(x[t-M:t]) - x[t]| for t >= M, using np.mean on each window
    exactly like the literal loop (float64, numpy's summation order)."""
    x = np.asarray(x, dtype=np.float64)
    win = np.lib.stride_tricks.sliding_window_view(x, meanlength)     # win[i] = x[i:i+M]
    means = np.array([np.mean(w) for w in win[:-1]]) if len(x) <= 4096 else win[:-1].mean(axis=1)
    d = np.full(len(x), 0.0)
    d[meanlength:] = np.abs(means - x[meanlength:])
    return d


def greedy_holdoff(cand_mask, start, holdoff, n, tail):
    """First candidate >= start, then first candidate >= prev+holdoff, stopping when
    bob+tail > n (the literal loops' break test)."""
    idx = np.nonzero(cand_mask)[0]
    hits = []
    bob = start
    while True:
        k = np.searchsorted(idx, bob, side='left')
        if k >= len(idx):
            break
        t = int(idx[k])
        if t + tail > n:
            break
        hits.append(t)
        bob = t + holdoff
    return hits


def trigger_block_literal(phasevalues, averagelength=128, phase_threshold=25., start=100,
                          pre=100, post=300, holdoff=200, wrap_negative=True):
    """pulse_triggering.py:114-208: add 360 to negatives (:110-112), block means
    (:117-120), |mean[bob//A]-x[bob]|>thr, window [bob-100,bob+300), hold-off 200."""
    x = np.array(phasevalues, dtype=np.float64)
    n = len(x)
    if wrap_negative:
        for k in range(n):
            if x[k] < 0:
                x[k] = x[k] + 360
    numberofaverages = n // averagelength
    phase_means = np.zeros(numberofaverages)
    for j in range(numberofaverages):
        phase_means[j] = np.mean(x[(averagelength * (j + 1) - averagelength):averagelength * (j + 1)])
    hits = []
    bob = start
    while bob < n:
        whichmean = bob // averagelength
        if bob + post > n:
            break
        if abs(phase_means[whichmean] - x[bob]) > phase_threshold:
            hits.append(bob)
            bob = bob + holdoff
        else:
            bob = bob + 1
    return hits


def trigger_contsnapshot_literal(qdr_phase_values, averagelength=64, phase_threshold=25.):
    """ROACH_Pulses.py:614-725: start 500, break if bob+1500 > len, hold-off 1000,
    window [bob-500, bob+1500)."""
    return trigger_block_literal(qdr_phase_values, averagelength, phase_threshold, start=500,
                                 pre=500, post=1500, holdoff=1000, wrap_negative=False)
