"""CPU arm of bench.py (`cpu_baseline` leg and `--impl reference`): TEST INFRASTRUCTURE, never the product path.

Everything the CPU arm needs is made on the CPU, so that it runs without libmkidgpu.so and without a GPU:
  * the board configuration of ReadoutChain.synthetic_boards, restated with the oracle's control plane
    (dds_freqs / select_bins / define_dds_lut = ROACH_Setup.py:506-550);
  * a synthetic ADC stream of the bench's kind (comb of the driven tones on the fs/N_lut grid, each
    phase-modulated by exponential pulses, white noise, 12-bit range).  The comb is periodic in N_lut samples (one IFFT);
    the pulses are a sparse correction a_i e^{j(w_i t + phi_i)} (e^{j theta_i(t)} - 1) on the samples they touch;
  * thresholds by loadThresholds (ROACH_Pulses.py:259-288) on a pulse-free stretch, as ReadoutChain.derive_thresholds.
One worker process per core keeps its board, stream and thresholds RESIDENT; a timed step is one `pool.map` over the
workers, each running the float64 model (channelize_phase), the integer detection (detect_emit) and PacketMaster's
binning (packetmaster_bin) over its stream.  Pool start-up, imports, input synthesis and threshold derivation are outside
the clock.
"""
import os
import time

import numpy as np

from . import channelizer as oc
from . import control, decode as odec
from . import lut as olut

FS = 512e6
_W = {}


def make_board(seed, n_lut=2 ** 19, n_active=253, lo_freq=5.0e9):
    """Same recipe as mkids_sdr_b200.chain.ReadoutChain.synthetic_boards (seed = seed0 + board)."""
    rng = np.random.default_rng(seed)
    k = np.sort(rng.choice(np.arange(-n_lut // 2 + 4096, n_lut // 2 - 4096), 256, replace=False))
    res = FS / n_lut
    dac_freqs = [lo_freq + float(v) * res for v in k]
    freqs_dds = olut.dds_freqs(dac_freqs, lo_freq, FS, res)
    bins, resid = olut.select_bins(freqs_dds, FS, res)
    I_dds, Q_dds, _ = olut.define_dds_lut(resid, FS, res, [0.] * 256)
    zero = np.zeros(256, bool)
    zero[n_active:] = True
    return dict(tone_bins=(k % n_lut)[:n_active].astype(np.int64), bins=np.array(bins), I_dds=I_dds, Q_dds=Q_dds, zero_ch=zero)


def synth_fast(n, tone_bins, n_lut, seed, pulse_rate=1000.0, tau_us=30.0, deg=(20.0, 120.0), noise_lsb=8.0,
               full_scale=1800.0):
    """int16 [n][2]; see the module docstring.  n must be a multiple of 512."""
    rng = np.random.default_rng(seed)
    T = len(tone_bins)
    phases = rng.uniform(0, 2 * np.pi, T)
    X = np.zeros(n_lut, dtype=np.complex128)
    np.add.at(X, np.asarray(tone_bins) % n_lut, np.exp(1j * phases))
    period = np.fft.ifft(X) * n_lut
    reps = -(-n // n_lut)
    x = np.tile(period, reps)[:n].copy()
    n_us = n // 512
    if pulse_rate > 0:
        decay = np.exp(-1.0 / tau_us)
        cut = int(np.ceil(tau_us * np.log(np.deg2rad(deg[1]) / 1e-3)))        # us until a pulse is below 1 mrad
        for i, kb in enumerate(tone_bins):
            npul = rng.poisson(pulse_rate * n_us * 1e-6)
            if npul == 0:
                continue
            t0s = np.sort(rng.integers(0, n_us, npul))
            depth = np.deg2rad(rng.uniform(deg[0], deg[1], npul))
            theta = np.zeros(n_us)
            for t0, dpt in zip(t0s, depth):
                u1 = min(n_us, t0 + cut)
                theta[t0:u1] -= dpt * decay ** np.arange(u1 - t0)
            us = np.nonzero(theta)[0]
            if us.size == 0:
                continue
            idx = (us[:, None] * 512 + np.arange(512)[None, :]).reshape(-1)
            carrier = np.exp(1j * (2 * np.pi * ((int(kb) * idx) % n_lut) / n_lut + phases[i]))
            x[idx] += carrier * (np.exp(1j * np.repeat(theta[us], 512)) - 1.0)
    x *= full_scale / (4.0 * np.sqrt(T)) if T > 1 else full_scale
    x += rng.normal(0, noise_lsb, n) + 1j * rng.normal(0, noise_lsb, n)
    iq = np.empty((n, 2), dtype=np.int16)
    iq[:, 0] = np.clip(np.rint(x.real), -2047, 2047)
    iq[:, 1] = np.clip(np.rint(x.imag), -2047, 2047)
    return iq


def _channelize(iq, cfg, slab=1 << 20):
    """channelize_phase in slabs (bounded memory), carrying the input history."""
    n = iq.shape[0]
    raws = []
    for a in range(0, n, slab):
        b = min(n, a + slab)
        h = iq[max(0, a - 2 * 8192):a] if a else None
        _, raw = oc.channelize_phase(iq[a:b], cfg, f0=a // 256, history=h)
        raws.append(raw)
    return np.concatenate(raws)


def worker_init(spec):
    """spec: dict(seed0, n_boards, n_lut, n_active, n_each, fir_int, index via a shared counter file is avoided: the
    worker index comes from its process name)."""
    import multiprocessing as mp
    ident = mp.current_process()._identity
    widx = (ident[0] - 1) if ident else 0
    b = widx % spec['n_boards']
    bd = make_board(spec['seed0'] + b, spec['n_lut'], spec['n_active'])
    cfg = oc.ChanConfig(bd['bins'], bd['I_dds'], bd['Q_dds'], spec['fir_int'], zero_ch=bd['zero_ch'], M=20, L=1000, W=32)
    # thresholds: loadThresholds on a pulse-free stretch (2^22 samples)
    quiet = synth_fast(1 << 22, bd['tone_bins'], spec['n_lut'], seed=7 + b, pulse_rate=0.0)
    raw = _channelize(quiet, cfg)
    thr = np.full(256, -25736, dtype=np.int64)
    nsamp = min(20480, raw.shape[0] - 64)
    for c in range(spec['n_active']):
        thr[c] = control.threshold_from_phase(raw[64:64 + nsamp, c])[0]
    cfg.thresholds = thr
    _W['cfg'] = cfg
    _W['iq'] = synth_fast(spec['n_each'], bd['tone_bins'], spec['n_lut'], seed=1000 + widx)
    _W['npix'] = spec.get('npix_per_roach', 253)


def worker_step(_):
    cfg, iq = _W['cfg'], _W['iq']
    raw = _channelize(iq, cfg)
    words = oc.detect_emit(raw, cfg, 0, np.zeros(256, np.int64), raw.shape[0] - 64 - cfg.M)
    res = odec.packetmaster_bin([np.array(words, dtype=np.uint64)], _W['npix'], 4)
    return len(words), int(res['counts'].sum())


class CpuArm:
    """Persistent worker pool: `CpuArm(cores, n_each, ...)` builds everything (untimed), `step()` returns the seconds of one
    pass of all workers over their resident streams."""

    def __init__(self, cores, n_each, fir_int, seed0=42, n_boards=8, n_lut=2 ** 19, n_active=253):
        import multiprocessing as mp
        self.cores, self.n_each = cores, n_each
        spec = dict(seed0=seed0, n_boards=n_boards, n_lut=n_lut, n_active=n_active, n_each=n_each,
                    fir_int=np.asarray(fir_int, dtype=np.int64))
        t0 = time.time()
        self.pool = mp.get_context('spawn').Pool(cores, initializer=worker_init, initargs=(spec,))
        self.words = self.pool.map(worker_step, range(cores), chunksize=1)           # warm-up pass (also waits for the initialisers)
        self.setup_seconds = time.time() - t0

    def step(self):
        t0 = time.time()
        self.words = self.pool.map(worker_step, range(self.cores), chunksize=1)
        return time.time() - t0

    def describe(self):
        return '%d resident board streams x 2^%d samples, one process per core (float64 model + integer detection + PacketMaster binning)' % (
            self.cores, int(np.log2(self.n_each)))

    def close(self):
        self.pool.close()
        self.pool.join()


def fir_int_default():
    """The quantised taps of matched_30us.txt (ROACH_Pulses.py:69,88-89), from the test fixture copy."""
    here = os.path.dirname(os.path.abspath(__file__))
    taps = np.load(os.path.join(os.path.dirname(here), 'tests', 'golden', 'fir_taps.npz'))['matched_30us']
    return control.fir_quantise(taps)
