"""Oracle: fixed-point bitfield helpers (TEST INFRASTRUCTURE, not product code).

Restates reference ``Utils/bin.py:1-48`` and ``Utils/binTools.py:1-64`` with the
Python-2 semantics the reference was written for:
  * ``/`` on ints is floor division           (bin.py:22  signBit)
  * ``round()`` rounds half AWAY from zero    (bin.py:38)
  * ``int()`` truncates toward zero
"""
import math

import numpy as np


def py2_round(x):
    """Python-2 ``round``: half away from zero (py3 rounds half to even)."""
    x = float(x)
    if x >= 0.0:
        return math.floor(x + 0.5)
    return -math.floor(-x + 0.5)


def binMask(nBits):
    # Utils/bin.py:2-3
    return int('1' * nBits, 2)


bitmask = binMask  # Utils/binTools.py:2-3


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
    # Utils/bin.py:18-29 ; signBit uses py2 integer division
    value = int(value) >> nBitsAfterEnd
    bitMask = int('1' * nBits, 2)
    value = value & bitMask
    signBit = int(value) // 2 ** (nBits - 1)
    if signBit != 0:
        value = ((~value) & bitMask) + 1
        value = -value
    value = float(value) / 2.0 ** binaryPoint
    if format == 'deg':
        value = value * 180.0 / np.pi
    return value


def castBin(value, nBits=12, binaryPoint=9, quantization='Truncate', format='uint'):
    # Utils/bin.py:31-48 ; round() is py2 half-away-from-zero
    if format == 'deg':
        value = value * np.pi / 180.0
    value = value * 2 ** binaryPoint
    if quantization == 'Truncate':
        value = int(value)
    else:
        value = int(py2_round(value))
    bitMask = int('1' * nBits, 2)
    if value < 0:
        value = -value
        value = ((~value) & bitMask) + 1
    value = value & bitMask
    if format != 'uint':
        value = extractBin(value, nBits=nBits, binaryPoint=binaryPoint)
        if format == 'deg':
            value = value * 180.0 / np.pi
    return value


def reinterpretBin(values, nBits=12, binaryPoint=9):
    # Utils/binTools.py:50-64 (vectorised extractBin, u64 -> f64)
    mask = int('1' * nBits, 2)
    values = np.asarray(values).astype(np.uint64) & np.uint64(mask)
    values = np.array(values, dtype=np.uint64)
    signBits = np.array(values // np.uint64(2 ** (nBits - 1)), dtype=bool)
    values[signBits] = ((~values[signBits]) & np.uint64(mask)) + np.uint64(1)
    values = np.array(values, dtype=np.double)
    values[signBits] = -values[signBits]
    values = values / 2. ** binaryPoint
    return values
