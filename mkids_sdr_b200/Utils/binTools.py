"""Drop-in for the reference's Utils/binTools.py.  reinterpretBin runs on the GPU."""
import numpy as np

from .. import _lib


def bitmask(nBits):
    # Utils/binTools.py:2-3
    return (1 << nBits) - 1


def bin12_9ToDeg(binOffset12_9):
    x = binOffset12_9 / 2.0 ** 9 - 4.0
    return x * 180.0 / np.pi


def bin12_9ToRad(binOffset12_9):
    x = binOffset12_9 / 2.0 ** 9 - 4.0
    return x


def peakfit(y1, y2, y3):
    if y3 + y1 - 2 * y2 == 0:
        return y2
    y4 = y2 - 0.125 * ((y3 - y1) ** 2) / (y3 + y1 - 2 * y2)
    return y4


def reinterpretBin(values, nBits=12, binaryPoint=9, nBitsAfterEnd=0, ctx=None):
    """Utils/binTools.py:50-64 on the GPU (mkid_reinterpret_bin): u64 -> f64."""
    ctx = ctx or _lib.default_context()
    v = np.ascontiguousarray(np.asarray(values).astype(np.uint64))
    out = np.empty(v.shape, dtype=np.float64)
    if v.size:
        ctx._check(ctx.lib.mkid_reinterpret_bin(ctx.h, _lib.ptr(v), v.size, nBits, binaryPoint, nBitsAfterEnd,
                                                _lib.ptr(out)))
    return out


def extractBin(value, nBits=12, binaryPoint=9, nBitsAfterEnd=0, format='rad'):
    # Utils/binTools.py:18-29 (identical to Utils/bin.py)
    from . import bin as _bin
    return _bin.extractBin(value, nBits, binaryPoint, nBitsAfterEnd, format)


def castBin(value, nBits=12, binaryPoint=9, quantization='Truncate', format='uint'):
    # Utils/binTools.py:31-48 (identical to Utils/bin.py)
    from . import bin as _bin
    return _bin.castBin(value, nBits, binaryPoint, quantization, format)
