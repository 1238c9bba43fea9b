"""ctypes binding of libmkidgpu.so (the C ABI declared in include/mkidgpu.h).

The product path has NO CPU fallback: if the shared library is missing, or no
sm_100 device is present, the calls raise.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int32, c_int64, c_size_t, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('MKIDGPU_LIB') or os.path.join(_HERE, 'libmkidgpu.so')     # (override: kernel experiments)

MERGE_MAX_SEC = 4            # MKID_MERGE_MAX_SEC of include/mkidgpu.h
MKID_OK, MKID_ENODEV, MKID_EINVAL, MKID_ENOMEM, MKID_ECUDA, MKID_ENCCL = 0, -1, -2, -3, -4, -5


class MkidError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('mkidgpu error %d: %s' % (code, msg))
        self.code = code


class DecodeCfg(Structure):
    _fields_ = [('n_roaches', c_int32), ('npix_per_roach', c_int32), ('exptime', c_int32),
                ('max_events', c_int32), ('hist_field_shift', c_int32), ('n_bins', c_int32),
                ('bin_lut', c_void_p)]


class ChanParams(Structure):
    _fields_ = [('n_boards', c_int32), ('n_lut', c_int32), ('mean_len', c_int32), ('holdoff', c_int32),
                ('peak_win', c_int32), ('reserved', c_int32)]


class SynthParams(Structure):
    _fields_ = [('n_tones', c_int32), ('n_lut', c_int32), ('full_scale', c_float), ('noise_lsb', c_float),
                ('pulse_rate', c_float), ('tau_us', c_float), ('deg_lo', c_float), ('deg_hi', c_float),
                ('seed', ctypes.c_uint64)]


class TriggerCfg(Structure):
    _fields_ = [('mode', c_int32), ('mean_len', c_int32), ('start', c_int32), ('holdoff', c_int32), ('tail', c_int32),
                ('wrap_negative', c_int32), ('sum_order', c_int32), ('reserved', c_int32), ('threshold', c_double)]


class DecodeStats(Structure):
    _fields_ = [('n_eos', c_int64), ('n_corrupt_eos', c_int64), ('n_nonpixel', c_int64),
                ('n_ignored', c_int64), ('n_valid', c_int64)]


# name -> (restype, argtypes); kept in one table so the CPU test-suite can check that the
# library exports every symbol the header declares.
_SIGNATURES = {
    'mkid_version': (c_char_p, []),
    'mkid_init': (c_int32, [c_int32, POINTER(c_void_p)]),
    'mkid_destroy': (None, [c_void_p]),
    'mkid_last_error': (c_char_p, [c_void_p]),
    'mkid_sync': (c_int32, [c_void_p]),
    'mkid_stream': (c_void_p, [c_void_p]),
    'mkid_merge_words_dev': (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int64, c_void_p]),
    'mkid_chan_overflowed': (c_int32, [c_void_p, c_void_p, c_void_p, c_int32]),
    'mkid_chan_set_pipelined': (c_int32, [c_void_p, c_void_p, c_int32]),
    'mkid_chan_detect_pending': (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    'mkid_stream_wait_event': (c_int32, [c_void_p, c_void_p, c_int32]),
    'mkid_wait_stream': (c_int32, [c_void_p, c_void_p]),
    'mkid_stream_wait_ctx': (c_int32, [c_void_p, c_void_p]),
    'mkid_nccl_version': (c_int32, [c_void_p, c_void_p]),
    'mkid_nccl_unique_id': (c_int32, [c_void_p, c_void_p]),
    'mkid_nccl_init': (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    'mkid_nccl_destroy': (c_int32, [c_void_p, c_void_p]),
    'mkid_hist_allreduce': (c_int32, [c_void_p, c_void_p, c_void_p, c_size_t]),
    'mkid_hist_reduce': (c_int32, [c_void_p, c_void_p, c_void_p, c_size_t, c_int32]),
    'mkid_launch_count': (c_int64, [c_void_p]),
    'mkid_event_record': (c_int32, [c_void_p, c_int32]),
    'mkid_event_elapsed_ms': (c_int32, [c_void_p, c_int32, c_int32, POINTER(c_float)]),
    'mkid_host_alloc': (c_int32, [c_void_p, c_size_t, POINTER(c_void_p)]),
    'mkid_host_free': (c_int32, [c_void_p, c_void_p]),
    'mkid_dev_alloc': (c_int32, [c_void_p, c_size_t, POINTER(c_void_p)]),
    'mkid_dev_free': (c_int32, [c_void_p, c_void_p]),
    'mkid_memcpy': (c_int32, [c_void_p, c_void_p, c_void_p, c_size_t]),
    'mkid_memset': (c_int32, [c_void_p, c_void_p, c_int32, c_size_t]),
    'mkid_flush_l2': (c_int32, [c_void_p]),
    'mkid_upload_async': (c_int32, [c_void_p, c_void_p, c_void_p, c_size_t, c_int32]),
    'mkid_upload_wait': (c_int32, [c_void_p, c_int32]),
    'mkid_upload_consumed': (c_int32, [c_void_p, c_int32]),
    'mkid_decode_words': (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                    POINTER(DecodeCfg), c_void_p, c_void_p, c_void_p]),
    'mkid_decode_words_seg': (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_int32, POINTER(DecodeCfg), c_void_p, c_void_p, c_void_p]),
    'mkid_decode_words_dev': (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_int32, POINTER(DecodeCfg), c_void_p, c_void_p]),
    'mkid_decode_lists': (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                    POINTER(DecodeCfg), c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    'mkid_decode_merged': (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                     POINTER(DecodeCfg), c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    'mkid_decode_wire_lists': (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                         POINTER(DecodeCfg), c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_void_p]),
    'mkid_decode_wire': (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                   POINTER(DecodeCfg), c_void_p, c_void_p, c_void_p]),
    'mkid_counts_cap': (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int32]),
    'mkid_unpack_fields': (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'mkid_reinterpret_bin': (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p]),
    'mkid_quicklook_image': (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p]),
    'mkid_chan_create': (c_int32, [c_void_p, POINTER(ChanParams), POINTER(c_void_p)]),
    'mkid_chan_destroy': (None, [c_void_p, c_void_p]),
    'mkid_chan_set_fir': (c_int32, [c_void_p, c_void_p, c_void_p]),
    'mkid_chan_set_window': (c_int32, [c_void_p, c_void_p, c_void_p]),
    'mkid_chan_set_board': (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    'mkid_chan_set_thresholds': (c_int32, [c_void_p, c_void_p, c_int32, c_void_p]),
    'mkid_chan_reset': (c_int32, [c_void_p, c_void_p]),
    'mkid_chan_set_f32_phase_out': (c_int32, [c_void_p, c_void_p, c_void_p]),
    'mkid_chan_process': (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_int64, c_void_p,
                                    c_void_p]),
    'mkid_chan_last_kernel_ms': (c_int32, [c_void_p, c_void_p, POINTER(c_float)]),
    'mkid_chan_kernel_ms_sum': (c_int32, [c_void_p, c_void_p, c_int32, POINTER(c_float)]),
    'mkid_chan_n_words_dev': (c_int32, [c_void_p, c_void_p, POINTER(c_void_p)]),
    'mkid_chan_detect': (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64,
                                   c_void_p]),
    'mkid_iq_snapshot_decode': (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    'mkid_phase_deg_from_iq': (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_void_p]),
    'mkid_soft_trigger': (c_int32, [c_void_p, c_void_p, c_int32, c_int64, POINTER(TriggerCfg), c_void_p, c_int32, c_void_p]),
    'mkid_thresholds_from_phase': (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_int64, c_int32, c_int64, c_double,
                                             c_void_p, c_void_p, c_void_p]),
    'mkid_noise_spectrum': (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_int32, c_double, c_void_p]),
    'mkid_spectra_products': (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_double, c_int32, c_int32, c_void_p,
                                        c_void_p, c_void_p]),
    'mkid_tpl_median': (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int64, c_void_p]),
    'mkid_tpl_prepare': (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_float, c_float, c_void_p, c_void_p]),
    'mkid_tpl_convpeak': (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    'mkid_tpl_accumulate': (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    'mkid_dashboard_image': (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                       c_void_p, c_void_p, c_void_p, c_void_p]),
    'mkid_random_phases': (c_int32, [ctypes.c_uint32, c_int32, c_void_p]),
    'mkid_comb_lut': (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_double, c_int32, c_int32, c_double,
                                c_int32, c_double, c_int32, c_void_p, c_void_p, c_void_p]),
    'mkid_dds_lut': (c_int32, [c_void_p, c_void_p, c_void_p, c_double, c_int32, c_int32, c_int32, c_int32, c_void_p,
                               c_void_p, c_void_p]),
    'mkid_pack_dram': (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    'mkid_sincos_cr': (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    'mkid_adc_unpack12': (c_int32, [c_void_p, c_void_p, c_int64, c_void_p]),
    'mkid_adc_pack12': (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    'mkid_synth_adc': (c_int32, [c_void_p, POINTER(SynthParams), c_int32, c_void_p, c_void_p, c_void_p, c_int64,
                                 c_int64, c_void_p]),
}

_lib = None


def header_symbols():
    """Names of all functions declared in include/mkidgpu.h (parsed from the header)."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), 'include', 'mkidgpu.h')
    txt = open(hdr).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(mkid_[a-z0-9_]+)\s*\(', txt)))


def load():
    """dlopen libmkidgpu.so and declare the prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MkidError(MKID_ENODEV, 'libmkidgpu.so not built (%s); run `python -m mkids_sdr_b200.build`. '
                                     'There is no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(x):
    """Address of a numpy array / torch tensor / DeviceBuffer / int / None as c_void_p."""
    if x is None:
        return None
    if isinstance(x, int):
        return c_void_p(x)
    if isinstance(x, np.ndarray):
        assert x.flags['C_CONTIGUOUS'], 'array must be C-contiguous'
        return c_void_p(x.ctypes.data)
    if hasattr(x, 'data_ptr'):          # torch tensor (host or device): optional carrier
        return c_void_p(x.data_ptr())
    if hasattr(x, 'ptr'):
        return c_void_p(x.ptr)
    raise TypeError('cannot take the address of %r' % type(x))


class DeviceBuffer:
    """Device memory owned through the C ABI (for callers that do not use torch)."""

    def __init__(self, ctx, nbytes):
        self.ctx, self.nbytes = ctx, int(nbytes)
        p = c_void_p()
        ctx._check(ctx.lib.mkid_dev_alloc(ctx.h, max(self.nbytes, 16), ctypes.byref(p)))
        self.ptr = p.value

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        self.ctx._check(self.ctx.lib.mkid_memcpy(self.ctx.h, c_void_p(self.ptr), ptr(arr), arr.nbytes))
        self.ctx.sync()
        return self

    def download(self, dtype, count=None, offset_bytes=0):
        dtype = np.dtype(dtype)
        if count is None:
            count = (self.nbytes - offset_bytes) // dtype.itemsize
        out = np.empty(count, dtype=dtype)
        self.ctx._check(self.ctx.lib.mkid_memcpy(self.ctx.h, ptr(out), c_void_p(self.ptr + offset_bytes), out.nbytes))
        self.ctx.sync()
        return out

    def zero(self):
        self.ctx._check(self.ctx.lib.mkid_memset(self.ctx.h, c_void_p(self.ptr), 0, self.nbytes))
        return self

    def free(self):
        if self.ptr:
            self.ctx.lib.mkid_dev_free(self.ctx.h, c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedBuffer:
    """Pinned host memory owned through the C ABI, exposed as a numpy array."""

    def __init__(self, ctx, nbytes):
        self.ctx, self.nbytes = ctx, int(nbytes)
        p = c_void_p()
        ctx._check(ctx.lib.mkid_host_alloc(ctx.h, max(self.nbytes, 16), ctypes.byref(p)))
        self.ptr = p.value
        self._buf = (ctypes.c_uint8 * max(self.nbytes, 16)).from_address(self.ptr)

    def view(self, dtype, count=None):
        a = np.frombuffer(self._buf, dtype=np.uint8, count=self.nbytes).view(dtype)
        return a if count is None else a[:count]

    def free(self):
        if self.ptr:
            self._buf = None
            self.ctx.lib.mkid_host_free(self.ctx.h, c_void_p(self.ptr))
            self.ptr = 0


class Context:
    """One GPU + one CUDA stream (mkid_ctx)."""

    def __init__(self, device=0):
        self.lib = load()
        h = c_void_p()
        rc = self.lib.mkid_init(int(device), ctypes.byref(h))
        if rc != MKID_OK:
            raise MkidError(rc, (self.lib.mkid_last_error(None) or b'').decode())
        self.h = h
        self.device = int(device)

    def _check(self, rc):
        if rc != MKID_OK:
            raise MkidError(rc, (self.lib.mkid_last_error(self.h) or b'').decode())

    def sync(self):
        self._check(self.lib.mkid_sync(self.h))

    @property
    def launches(self):
        return int(self.lib.mkid_launch_count(self.h))

    @property
    def stream(self):
        return self.lib.mkid_stream(self.h)

    def record(self, slot):
        self._check(self.lib.mkid_event_record(self.h, slot))

    def wait_event(self, owner, slot):
        """This context's stream waits for event `slot` recorded on another context of the same GPU."""
        self._check(self.lib.mkid_stream_wait_event(self.h, owner.h, int(slot)))

    def elapsed_ms(self, a, b):
        ms = c_float()
        self._check(self.lib.mkid_event_elapsed_ms(self.h, a, b, ctypes.byref(ms)))
        return float(ms.value)

    def upload_async(self, dst_dev, src_host, nbytes, slot):
        """Host -> device copy on the copy stream (overlaps the context stream); see upload_wait / upload_consumed."""
        self._check(self.lib.mkid_upload_async(self.h, ptr(dst_dev), ptr(src_host), int(nbytes), int(slot)))

    def upload_wait(self, slot):
        self._check(self.lib.mkid_upload_wait(self.h, int(slot)))

    def upload_consumed(self, slot):
        self._check(self.lib.mkid_upload_consumed(self.h, int(slot)))

    def flush_l2(self):
        self._check(self.lib.mkid_flush_l2(self.h))

    def adc_unpack12(self, packed_dev, n_samples, iq_dev):
        """12-bit packed ADC samples (3 bytes per complex sample) -> int16 [n][2]; device buffers, asynchronous."""
        self._check(self.lib.mkid_adc_unpack12(self.h, ptr(packed_dev), int(n_samples), ptr(iq_dev)))

    def adc_pack12(self, iq_dev, n_samples, packed_dev, count_clipped=True):
        """int16 [n][2] -> 12-bit packed; returns the number of values that did not fit 12 bits (clipped)."""
        n_bad = ctypes.c_int64(0)
        self._check(self.lib.mkid_adc_pack12(self.h, ptr(iq_dev), int(n_samples), ptr(packed_dev),
                                             ctypes.byref(n_bad) if count_clipped else None))
        return int(n_bad.value)

    def alloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    def pinned(self, nbytes):
        return PinnedBuffer(self, nbytes)

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        return DeviceBuffer(self, arr.nbytes).upload(arr)

    def close(self):
        if getattr(self, 'h', None):
            self.lib.mkid_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=None):
    """Per-device default context (LOCAL_RANK selects the GPU under torchrun)."""
    if device is None:
        device = int(os.environ.get('LOCAL_RANK', '0'))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
