"""Software pulse triggers, snapshot decoders and threshold derivation of the reference's analysis
scripts, on the GPU.

  trigger_rolling        DataReadout/ChannelizerControls/pulse_triggering_v2.py:104-174
                         (= pulse_triggering_IQ.py:160-200, pulse_triggering_just_phase_v2.py:91-119)
  trigger_block          pulse_triggering.py:114-208
  trigger_contsnapshot   ROACH_Pulses.py:614-725 (contsnapshot)
  decode_iq_snapshot     pulse_triggering_IQ.py:121-147
  phase_deg_from_iq      pulse_triggering_IQ.py:152
  thresholds_from_phase  ROACH_Pulses.py:259-288 (loadThresholds) for all channels of a phase stream at once

Parameter names follow the scripts.  Hit lists are bit-identical to the NumPy loops (np.mean is evaluated
in NumPy's own float64 summation order).
"""
import ctypes

import numpy as np

from . import _lib


def _soft_trigger(phase, mode, mean_len, start, holdoff, tail, threshold, wrap_negative=False, numpy16_sum=False,
                  max_hits=None, ctx=None):
    ctx = ctx or _lib.default_context()
    x = np.ascontiguousarray(phase, dtype=np.float64)
    single = x.ndim == 1
    x2 = x.reshape(1, -1) if single else x
    s, n = x2.shape
    if max_hits is None:
        max_hits = int(n // max(holdoff, 1) + 2)
    hits = np.zeros((s, max_hits), dtype=np.int32)
    n_hits = np.zeros(s, dtype=np.int32)
    cfg = _lib.TriggerCfg(mode, int(mean_len), int(start), int(holdoff), int(tail), 1 if wrap_negative else 0,
                          1 if numpy16_sum else 0, 0, float(threshold))
    ctx._check(ctx.lib.mkid_soft_trigger(ctx.h, _lib.ptr(x2), s, n, ctypes.byref(cfg), _lib.ptr(hits), max_hits,
                                         _lib.ptr(n_hits)))
    ctx.sync()
    out = [hits[i, :min(int(n_hits[i]), max_hits)].tolist() for i in range(s)]
    return out[0] if single else out


def trigger_rolling(phasevalues, meanlength=20, pulselength=1000, phase_threshold=25., pre=100, **kw):
    """pulse_triggering_v2.py:104-174: returns the trigger indices `bob` (1-D input) or a list per stream."""
    return _soft_trigger(phasevalues, 0, meanlength, pre + meanlength, pulselength, pulselength, phase_threshold, **kw)


def trigger_block(phasevalues, averagelength=128, phase_threshold=25., start=100, post=300, holdoff=200,
                  wrap_negative=True, **kw):
    """pulse_triggering.py:114-208: block means of `averagelength`, window [bob-100, bob+300), hold-off 200."""
    return _soft_trigger(phasevalues, 1, averagelength, start, holdoff, post, phase_threshold,
                         wrap_negative=wrap_negative, **kw)


def trigger_contsnapshot(qdr_phase_values, averagelength=64, phase_threshold=25., **kw):
    """ROACH_Pulses.py:614-725: start 500, window [bob-500, bob+1500), hold-off 1000."""
    return trigger_block(qdr_phase_values, averagelength, phase_threshold, start=500, post=1500, holdoff=1000,
                         wrap_negative=False, **kw)


def decode_iq_snapshot(buf, ctx=None):
    """pulse_triggering_IQ.py:121-147 -> (Iraw, Qraw) int16 arrays of length len(buf)//8."""
    ctx = ctx or _lib.default_context()
    b = np.frombuffer(bytes(buf), dtype=np.uint8) if not isinstance(buf, np.ndarray) else np.ascontiguousarray(buf, np.uint8)
    n = b.size // 8
    I = np.empty(n, np.int16); Q = np.empty(n, np.int16)
    ctx._check(ctx.lib.mkid_iq_snapshot_decode(ctx.h, _lib.ptr(b), b.size, _lib.ptr(I), _lib.ptr(Q)))
    ctx.sync()
    return I, Q


def phase_deg_from_iq(Iraw, Qraw, Ic=0., Qc=0., ctx=None):
    """pulse_triggering_IQ.py:152."""
    ctx = ctx or _lib.default_context()
    I = np.ascontiguousarray(Iraw, dtype=np.int16); Q = np.ascontiguousarray(Qraw, dtype=np.int16)
    deg = np.empty(I.size, np.float64)
    ctx._check(ctx.lib.mkid_phase_deg_from_iq(ctx.h, _lib.ptr(I), _lib.ptr(Q), I.size, float(Ic), float(Qc), _lib.ptr(deg)))
    ctx.sync()
    return deg


def thresholds_from_phase(phase_dev, n_boards, rows, n_samples, row0=0, n_ch=256, Nsigma=2.5, ctx=None):
    """loadThresholds on a device-resident raw phase stream int16 [n_boards][rows][n_ch] (samples
    row0 .. row0+n_samples of every channel).  Returns (thr_raw int32 [B][n_ch], med, p5 float64)."""
    ctx = ctx or _lib.default_context()
    assert 0 <= row0 and row0 + n_samples <= rows, 'sample window outside the phase stream'
    thr = np.empty((n_boards, n_ch), np.int32)
    med = np.empty((n_boards, n_ch), np.float64); p5 = np.empty((n_boards, n_ch), np.float64)
    base = _lib.ptr(phase_dev).value + int(row0) * n_ch * 2
    ctx._check(ctx.lib.mkid_thresholds_from_phase(ctx.h, base, n_boards, rows * n_ch, n_ch, n_ch, int(n_samples),
                                                  float(Nsigma), _lib.ptr(thr), _lib.ptr(med), _lib.ptr(p5)))
    ctx.sync()
    return thr, med, p5


def noise_spectrum(qdr_phase_values, nFFTAverages=100, norm1=50.0, ctx=None):
    """longsnapshot noise spectrum (ROACH_Pulses.py:521-537).  Returns (noiseFFT [dB], noiseFFTFreqs) for a
    1-D phase stream in degrees, or noiseFFT [streams][nSamplesPerFFT] for a 2-D input."""
    ctx = ctx or _lib.default_context()
    x = np.ascontiguousarray(qdr_phase_values, dtype=np.float64)
    single = x.ndim == 1
    x2 = x.reshape(1, -1) if single else x
    s, n = x2.shape
    nper = n // int(nFFTAverages)
    out = np.empty((s, nper), np.float64)
    ctx._check(ctx.lib.mkid_noise_spectrum(ctx.h, _lib.ptr(x2), s, n, int(nFFTAverages), float(norm1), _lib.ptr(out)))
    ctx.sync()
    freqs = np.fft.fftfreq(nper)
    return (out[0], freqs) if single else (out, freqs)
