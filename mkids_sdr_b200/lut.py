"""Host wrappers of the LUT synthesis kernels (K1-K3): thin ctypes calls, numpy in / numpy out."""
import numpy as np

from . import _lib

AMP_FULL_SCALE = 2 ** 15 - 1     # ROACH_Setup.py:420
SCALE_FUDGE = 1.1                # ROACH_Setup.py:453
FFT_LEN = 2 ** 9                 # ROACH_Setup.py:507
CH_SHIFT = 154                   # ROACH_Setup.py:508


def random_phases(n, seed=1000):
    """numpy.random.seed(1000); uniform(0, 2*pi) per tone, from the library's own MT19937."""
    out = np.empty(n, dtype=np.float64)
    lib = _lib.load()
    assert lib.mkid_random_phases(seed, n, _lib.ptr(out)) == 0
    return out


def comb_lut(freqs, sample_rate, n_samples, amplitudes, phases=None, echo='yes', random_phase='yes', offset=0,
             scale_override=None, ctx=None, out_I=None, out_Q=None):
    """Batched freqCombLUT.  freqs/amplitudes/phases: [batch][T] (or [T]).  Returns
    (I int16 [batch][N], Q int16 [batch][N], scale [batch], phases_used [batch][T]).  out_I / out_Q: optional
    device buffers (int16 [batch][N]) that receive the tables instead of host arrays (the LUT then never leaves HBM,
    e.g. on its way into mkid_pack_dram / the channelizer)."""
    ctx = ctx or _lib.default_context()
    f = np.atleast_2d(np.asarray(freqs, dtype=np.float64))
    batch, T = f.shape
    a = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(np.asarray(amplitudes, dtype=np.float64))[:, :T], (batch, T)))
    ph = np.zeros((batch, T)) if phases is None else np.atleast_2d(np.asarray(phases, dtype=np.float64))[:, :T]
    ph = np.ascontiguousarray(np.broadcast_to(ph, (batch, T))).copy()
    f = np.ascontiguousarray(f)
    I = out_I if out_I is not None else np.empty((batch, n_samples), dtype=np.int16)
    Q = out_Q if out_Q is not None else np.empty((batch, n_samples), dtype=np.int16)
    scale = np.empty(batch, dtype=np.float64)
    fudge = SCALE_FUDGE if echo == 'yes' else 1.0
    ctx._check(ctx.lib.mkid_comb_lut(ctx.h, _lib.ptr(f), _lib.ptr(a), _lib.ptr(ph), T, float(sample_rate), int(n_samples),
                                     int(offset), fudge, 1 if random_phase == 'yes' else 0,
                                     float(scale_override) if scale_override else 0.0, batch, _lib.ptr(I), _lib.ptr(Q),
                                     _lib.ptr(scale)))
    return I, Q, scale, ph


def dds_lut(residuals, phases, sample_rate, n_lut, ch_shift=CH_SHIFT, offset=0, ctx=None, out_I=None, out_Q=None,
            want_scales=True):
    """Batched define_DDS_LUT tables.  residuals/phases [batch][256] -> (I_dds, Q_dds int16 [batch][n_lut], scales).
    out_I / out_Q: optional device buffers (int16 [batch][n_lut]) that receive the tables instead of host arrays.
    want_scales=False (device outputs only): the call does not synchronise and returns no scales."""
    ctx = ctx or _lib.default_context()
    r = np.ascontiguousarray(np.atleast_2d(np.asarray(residuals, dtype=np.float64)))
    batch = r.shape[0]
    p = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(np.asarray(phases, dtype=np.float64)), r.shape))
    assert r.shape[1] == 256
    I = out_I if out_I is not None else np.empty((batch, n_lut), dtype=np.int16)
    Q = out_Q if out_Q is not None else np.empty((batch, n_lut), dtype=np.int16)
    sc = np.empty((batch, 256), dtype=np.float64) if want_scales else None
    ctx._check(ctx.lib.mkid_dds_lut(ctx.h, _lib.ptr(r), _lib.ptr(p), float(sample_rate), int(n_lut), int(ch_shift), int(offset), batch,
                                    _lib.ptr(I), _lib.ptr(Q), _lib.ptr(sc)))
    return I, Q, sc


def pack_dram(I_dac, Q_dac, I_dds, Q_dds, ctx=None, n=None, out=None):
    """write_LUTs byte image (ROACH_Setup.py:560-569) -> bytes of length 8*N.  With n (int16 samples per table) the
    four tables and out may be device buffers (the image of a batch of sets is the images of the sets back to back);
    out is then returned instead of bytes."""
    ctx = ctx or _lib.default_context()
    if n is not None:
        assert out is not None
        ctx._check(ctx.lib.mkid_pack_dram(ctx.h, _lib.ptr(I_dac), _lib.ptr(Q_dac), _lib.ptr(I_dds), _lib.ptr(Q_dds), int(n),
                                          _lib.ptr(out)))
        return out
    arrs = [np.ascontiguousarray(np.asarray(x).astype(np.int16)) for x in (I_dac, Q_dac, I_dds, Q_dds)]
    n = arrs[0].size
    assert all(x.size == n for x in arrs)
    out = np.empty(8 * n, dtype=np.uint8)
    ctx._check(ctx.lib.mkid_pack_dram(ctx.h, _lib.ptr(arrs[0]), _lib.ptr(arrs[1]), _lib.ptr(arrs[2]), _lib.ptr(arrs[3]), n,
                                      _lib.ptr(out)))
    return out.tobytes()


def sincos_cr(x, ctx=None):
    ctx = ctx or _lib.default_context()
    x = np.ascontiguousarray(x, dtype=np.float64)
    s = np.empty_like(x); c = np.empty_like(x)
    ctx._check(ctx.lib.mkid_sincos_cr(ctx.h, _lib.ptr(x), x.size, _lib.ptr(s), _lib.ptr(c)))
    return s, c
