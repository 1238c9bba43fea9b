"""MakeTemplate (DataReadout/ReadoutControls/lib/pulses.py:239-427) with the per-pulse arithmetic on the GPU.

    out = MakeTemplate(I, Q)        # I, Q: float32 [n_pulses][2000] (the `iqpulses` table of one resonator)

Returns the fields the reference writes to the `opt` table (`phasetemplate`, `phasenoise`, `phasenoiseidx`,
`count`, `pstart`, `flag`) plus the intermediate results.  Like the reference (`I += ...` on the table rows, :283,
:343) the input arrays are shifted IN PLACE.  The loop structure and every accept / reject comparison are the
reference's; only the array arithmetic of each pulse runs in CUDA (csrc/template.cu)."""
import numpy as np

from . import _lib


def MakeTemplate(I_all, Q_all, ctx=None):
    ctx = ctx or _lib.default_context()
    assert I_all.dtype == np.float32 and Q_all.dtype == np.float32 and I_all.shape == Q_all.shape
    assert I_all.ndim == 2 and I_all.shape[1] == 2000 and I_all.flags['C_CONTIGUOUS'] and Q_all.flags['C_CONTIGUOUS']
    n_all = I_all.shape[0]
    lib, h = ctx.lib, ctx.h
    dI, dQ = ctx.to_device(I_all), ctx.to_device(Q_all)
    dP3 = ctx.alloc(n_all * 2000 * 8)
    med = np.zeros(1, np.float32)
    rows = min(100, n_all)
    ctx._check(lib.mkid_tpl_median(h, _lib.ptr(dI), rows, 900, 2000, _lib.ptr(med))); ctx.sync()
    I1m = float(med[0])                                                  # :273
    ctx._check(lib.mkid_tpl_median(h, _lib.ptr(dQ), rows, 900, 2000, _lib.ptr(med))); ctx.sync()
    Q1m = float(med[0])

    def prepare(n):
        st = np.zeros((n, 6), np.float64)
        ctx._check(lib.mkid_tpl_prepare(h, _lib.ptr(dI), _lib.ptr(dQ), n, I1m, Q1m, _lib.ptr(dP3), _lib.ptr(st)))
        ctx.sync()
        return st, st[:, 5].copy().view(np.int32)[0::2]

    def accumulate(pulse, shift, norm, want_noise):
        tm = np.zeros(2000, np.float64)
        noise = np.zeros(800, np.float64) if want_noise else None
        if len(pulse):
            pu, sh, no = np.asarray(pulse, np.int32), np.asarray(shift, np.int32), np.asarray(norm, np.float64)
            ctx._check(lib.mkid_tpl_accumulate(h, _lib.ptr(dP3), _lib.ptr(pu), _lib.ptr(sh), _lib.ptr(no), len(pulse),
                                               _lib.ptr(tm), _lib.ptr(noise)))
            ctx.sync()
        return tm, noise

    # ---------------- first pass: preliminary template from (at most) 1000 pulses (:277-323)
    N = min(n_all, 1000)
    st, ploc = prepare(N)
    count = 0.0
    peaklist, acc, shifts, norms = [], [], [], []
    for j in range(N):
        mean_first, mean_last, stdev, peak, max_all = st[j, :5]
        if np.abs(mean_first - mean_last) > stdev * 2.0:                 # bad baseline subtraction (:298-300)
            continue
        peaklist.append(peak)
        if peak < 15.0 or peak > 120.0:
            continue
        if ploc[j] < 980 or ploc[j] > 1020:
            continue
        acc.append(j); shifts.append(1000 - int(ploc[j])); norms.append(max_all)
        count += 1
    count1 = int(count)
    tP, _ = accumulate(acc, shifts, norms, False)
    tP /= count
    accepted1 = acc
    peaklist = np.asarray(peaklist)
    pm = np.median(peaklist[np.where(peaklist > 15)])
    pdev = np.std(peaklist[np.where(peaklist > 15)])

    # ---------------- second pass: all pulses, start time from the correlation with the preliminary template (:337-385)
    st, _ = prepare(n_all)
    amax = np.zeros(n_all, np.int32); p3_at = np.zeros(n_all, np.float64)
    kern = np.ascontiguousarray(tP[900:1500])
    ctx._check(lib.mkid_tpl_convpeak(h, _lib.ptr(dP3), n_all, _lib.ptr(kern), _lib.ptr(amax), _lib.ptr(p3_at))); ctx.sync()
    count = 0.0
    acc, shifts, norms = [], [], []
    for j in range(n_all):
        mean_first, mean_last, stdev, _, max_all = st[j, :5]
        if np.abs(mean_first - mean_last) > stdev * 2.0:
            continue
        ploc2 = int(amax[j] - 1160.0)
        peak = p3_at[j]
        if peak < pm - 4.0 * pdev or peak > pm + 4.0 * pdev:
            continue
        if ploc2 < -30 or ploc2 > 30:
            continue
        acc.append(j); shifts.append(-ploc2); norms.append(max_all)
        count += 1
    tPf, noise = accumulate(acc, shifts, norms, True)
    tPf /= count
    noise /= count
    noiseidx = np.fft.fftfreq(len(noise), d=0.000002)
    # the table rows were shifted in place (twice for the pulses of the first pass)
    ctx._check(lib.mkid_memcpy(h, _lib.ptr(I_all), _lib.ptr(dI), I_all.nbytes))
    ctx._check(lib.mkid_memcpy(h, _lib.ptr(Q_all), _lib.ptr(dQ), Q_all.nbytes))
    ctx.sync()
    dI.free(); dQ.free(); dP3.free()
    flag = 1 if (count < 500 or pm < 10 or pm > 150) else 0
    return dict(tP=tP, tPf=tPf, phasetemplate=tPf, noise=noise, phasenoise=noise, noiseidx=noiseidx, phasenoiseidx=noiseidx,
                count1=count1, count=int(count), pm=float(pm), pdev=float(pdev), flag=flag,
                pstart=int((np.where(tPf == np.max(tPf)))[0][0]), accepted1=accepted1, accepted2=acc, peaklist=peaklist)
