"""PINNED: bit-identical to the outputs of the reference's own MakeTemplate (lib/pulses.py:239-427), executed in the dev
container by tests/golden/make_golden_analysis.py on the same synthetic iqpulses table -> tests/golden/analysis_golden.npz
(tests/test_oracle_golden.py::test_template_oracle_matches_reference_run).

Oracle: matched-filter template builder (TEST INFRASTRUCTURE).

Literal restatement of MakeTemplate, DataReadout/ReadoutControls/lib/pulses.py:239-427, without PyTables /
matplotlib: the `iqpulses` table is passed in as two float32 arrays I, Q [n_pulses][2000] (2 us per sample).
NOTE `I += ...` (:283,:343) shifts the table rows IN PLACE: pulses seen by the first pass are shifted a second
time in the second pass (by ~0).  Kept as is: the arrays passed in are modified.
"""
import numpy as np


def make_template(I_all, Q_all):
    dat_I, dat_Q = I_all, Q_all
    tP = np.zeros(2000, dtype='float64')
    tPf = np.zeros(2000, dtype='float64')
    noise = np.zeros(800, dtype='float64')
    N = len(dat_I)
    count = 0.0
    peaklist = []
    idx = np.arange(2000) * 2.0
    fitidx = np.concatenate((idx[:900], idx[1800:]))
    xc = 0.0
    yc = 0.0
    I1m = np.median(dat_I[:100, :900])                                  # :273
    Q1m = np.median(dat_Q[:100, :900])
    if N > 1000:
        N = 1000
    accepted1 = []
    for j in range(N):                                                  # first pass :281-323
        I = dat_I[j]
        Q = dat_Q[j]
        I += (I1m - np.median(I[1:900]))
        Q += (Q1m - np.median(Q[1:900]))
        P1 = np.arctan2(Q - yc, I - xc)
        P2 = np.rad2deg(np.unwrap(P1))
        fit = np.poly1d(np.polyfit(fitidx, np.concatenate((P2[:900], P2[1800:])), 1))
        P3 = P2 - fit(idx)
        stdev = np.std(P3[:100])
        if np.abs(np.mean(P3[:100]) - np.mean(P3[1900:])) > stdev * 2.0:
            continue
        peak = np.max(P3[980:1050])
        peaklist.append(peak)
        if peak < 15.0 or peak > 120.0:
            continue
        ploc = int((np.where(P3 == peak))[0][0])
        if ploc < 980 or ploc > 1020:
            continue
        P4 = np.roll(P3, 1000 - ploc)
        tP += P4 / np.max(P4)
        count += 1
        accepted1.append(j)
    count1 = int(count)
    tP /= count
    peaklist = np.asarray(peaklist)
    pm = np.median(peaklist[np.where(peaklist > 15)])
    pdev = np.std(peaklist[np.where(peaklist > 15)])
    N = len(dat_I)
    count = 0.0
    accepted2 = []
    for j in range(N):                                                  # second pass :337-385
        I = dat_I[j]
        Q = dat_Q[j]
        I += (I1m - np.median(I[1:900]))
        Q += (Q1m - np.median(Q[1:900]))
        P1 = np.arctan2(Q - yc, I - xc)
        P2 = np.rad2deg(np.unwrap(P1))
        fit = np.poly1d(np.polyfit(fitidx, np.concatenate((P2[:900], P2[1800:])), 1))
        P3 = P2 - fit(idx)
        stdev = np.std(P3[:100])
        if np.abs(np.mean(P3[:100]) - np.mean(P3[1900:])) > stdev * 2.0:
            continue
        conv = np.convolve(tP[900:1500], P3)
        ploc = int((np.where(conv == np.max(conv)))[0][0] - 1160.0)
        peak = np.max(P3[1000 + ploc])
        if peak < pm - 4.0 * pdev or peak > pm + 4.0 * pdev:
            continue
        if ploc < -30 or ploc > 30:
            continue
        P4 = np.roll(P3, -ploc)
        tPf += P4 / np.max(P4)
        count += 1
        accepted2.append(j)
        noise += np.abs(np.fft.fft(np.deg2rad(P4[50:850]))) ** 2
    tPf /= count
    noise /= count
    noiseidx = np.fft.fftfreq(len(noise), d=0.000002)
    flag = 1 if (count < 500 or pm < 10 or pm > 150) else 0
    return dict(tP=tP, tPf=tPf, noise=noise, noiseidx=noiseidx, count1=count1, count=int(count), pm=float(pm),
                pdev=float(pdev), flag=flag, pstart=int((np.where(tPf == np.max(tPf)))[0][0]),
                accepted1=accepted1, accepted2=accepted2, peaklist=peaklist)


def fake_pulses(n_pulses=300, seed=0, radius=2000.0, sigma=6.0, double_frac=0.05):
    """Synthetic iqpulses table in the spirit of FakeTemplateData (pulses.py:429+): a resonator loop point at
    (radius, 0) rotated by an exponential phase pulse starting near sample 1000, white noise, slow drifts."""
    rng = np.random.default_rng(seed)
    t = np.arange(2000)
    I = np.empty((n_pulses, 2000), dtype=np.float32)
    Q = np.empty((n_pulses, 2000), dtype=np.float32)
    for j in range(n_pulses):
        t0 = 1000 + rng.integers(-8, 9)
        amp = np.deg2rad(rng.uniform(25., 100.)) if rng.random() > 0.1 else np.deg2rad(rng.uniform(2., 10.))
        ph = np.where(t >= t0, amp * (np.exp(-(t - t0) / 60.0) - np.exp(-(t - t0) / 2.0)), 0.0)
        if rng.random() < double_frac:
            t1 = rng.integers(200, 800)
            ph = ph + np.where(t >= t1, 0.6 * amp * np.exp(-(t - t1) / 60.0), 0.0)
        ph = ph + rng.normal(0, 0.004) * (t / 2000.0) + rng.normal(0, 0.01)
        off_i, off_q = rng.normal(0, 30.0), rng.normal(0, 30.0)
        I[j] = radius * np.cos(ph) + rng.normal(0, sigma, 2000) + off_i
        Q[j] = radius * np.sin(ph) + rng.normal(0, sigma, 2000) + off_q
    return I, Q
