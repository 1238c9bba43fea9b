"""Drop-in for the reference's Utils/bin.py (function-for-function, Python-2 results).

Scalars are control-plane helpers (register values) and are evaluated on the host with
the reference's Python-2 semantics (`/` on ints floors, round() is half away from zero);
array inputs go to the GPU through Utils.binTools.reinterpretBin.
"""
import math

import numpy as np

from . import binTools as _bt


def _round_half_away(x):
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


def binMask(nBits):
    # Utils/bin.py:2-3
    return (1 << nBits) - 1


def bin12_9ToDeg(binOffset12_9):
    # Utils/bin.py:5-7
    x = binOffset12_9 / 2.0 ** 9 - 4.0
    return x * 180.0 / np.pi


def bin12_9ToRad(binOffset12_9):
    # Utils/bin.py:9-11
    x = binOffset12_9 / 2.0 ** 9 - 4.0
    return x


def peakfit(y1, y2, y3):
    # Utils/bin.py:12-16
    if y3 + y1 - 2 * y2 == 0:
        return y2
    y4 = y2 - 0.125 * ((y3 - y1) ** 2) / (y3 + y1 - 2 * y2)
    return y4


def extractBin(value, nBits=12, binaryPoint=9, nBitsAfterEnd=0, format='rad'):
    # Utils/bin.py:18-29.  Arrays are decoded on the GPU.
    if isinstance(value, np.ndarray):
        out = _bt.reinterpretBin(value, nBits, binaryPoint, nBitsAfterEnd=nBitsAfterEnd)
        return out * 180.0 / np.pi if format == 'deg' else out
    value = int(value) >> nBitsAfterEnd
    bitMask = (1 << nBits) - 1
    value = value & bitMask
    signBit = value // 2 ** (nBits - 1)          # py2 integer division
    if signBit != 0:
        value = ((~value) & bitMask) + 1
        value = -value
    value = float(value) / 2.0 ** binaryPoint
    if format == 'deg':
        value = value * 180.0 / np.pi
    return value


def castBin(value, nBits=12, binaryPoint=9, quantization='Truncate', format='uint'):
    # Utils/bin.py:31-48
    if format == 'deg':
        value = value * np.pi / 180.0
    value = value * 2 ** binaryPoint
    if quantization == 'Truncate':
        value = int(value)
    else:
        value = int(_round_half_away(value))     # py2 round()
    bitMask = (1 << nBits) - 1
    if value < 0:
        value = -value
        value = ((~value) & bitMask) + 1
    value = value & bitMask
    if format != 'uint':
        value = extractBin(value, nBits=nBits, binaryPoint=binaryPoint)
        if format == 'deg':
            value = value * 180.0 / np.pi
    return value
