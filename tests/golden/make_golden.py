"""Generate tests/golden/* from the reference tree (run in the dev container only).

    python tests/golden/make_golden.py [/root/reference]

/root/reference does not exist on the GPU box, so the fixtures this writes are
committed.  Everything here is DATA copied/condensed from the reference's own
fixtures (no source code):
  dac_golden.npz       ChannelizerControls/dac.npy.npz as int16 + sha256 of <i2 images
  ch_snap_0.npy        ChannelizerControls/ch_snap_0.txt (2048 phase samples, degrees)
  fir_taps.npz         ChannelizerControls/LUT/{matched_30us,Blackman,Hamming,Rect}*.txt
  utils_bin_py3.npz    outputs of the reference's Utils/binTools.reinterpretBin and the
                       py3-safe functions of Utils/bin.py, imported from the reference
"""
import hashlib
import importlib.util
import os
import sys

import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
here = os.path.dirname(os.path.abspath(__file__))
cc = os.path.join(ref, 'DataReadout', 'ChannelizerControls')

d = np.load(os.path.join(cc, 'dac.npy.npz'))
out = {}
for k in ('I_dac', 'Q_dac', 'I_dds', 'Q_dds'):
    a = np.asarray(d[k])
    assert a.min() >= -32768 and a.max() <= 32767
    out[k] = a.astype('<i2')
    out[k + '_sha256'] = np.array(hashlib.sha256(a.astype('<i2').tobytes()).hexdigest())
np.savez_compressed(os.path.join(here, 'dac_golden.npz'), **out)

snap = np.loadtxt(os.path.join(cc, 'ch_snap_0.txt'))
np.save(os.path.join(here, 'ch_snap_0.npy'), snap)

firs = {}
for name in ('matched_30us', 'BlackmanFilter_250kHz', 'HammingFilter_250kHz', 'RectFilter_250kHz'):
    firs[name] = np.loadtxt(os.path.join(cc, 'LUT', name + '.txt'))
np.savez(os.path.join(here, 'fir_taps.npz'), **firs)


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


bt = load(os.path.join(ref, 'Utils', 'binTools.py'), 'ref_binTools')
b = load(os.path.join(ref, 'Utils', 'bin.py'), 'ref_bin')
rng = np.random.default_rng(7)
vals = np.concatenate([np.arange(4096, dtype=np.uint64),
                       rng.integers(0, 2 ** 63, 4096, dtype=np.uint64)])
res = {'values': vals,
       'reinterpret_12_9': bt.reinterpretBin(vals.copy(), 12, 9),
       'reinterpret_16_13': bt.reinterpretBin(vals.copy(), 16, 13),
       'reinterpret_18_16': bt.reinterpretBin(vals.copy(), 18, 16),
       'bin12_9ToRad': np.array([b.bin12_9ToRad(int(v)) for v in range(4096)]),
       'bin12_9ToDeg': np.array([b.bin12_9ToDeg(int(v)) for v in range(4096)]),
       'binMask': np.array([b.binMask(n) for n in range(1, 33)], dtype=np.uint64),
       # castBin(format='uint') never reaches the py2-only `/` in extractBin, and for
       # these inputs round() has no exact .5 ties -> py3 run == py2 run
       'castBin_in': np.array([0.08, -0.08, 1.0, 3.99, -4.0, 0.3, -1.7]),
       'castBin_trunc_12_9': np.array([b.castBin(v) for v in [0.08, -0.08, 1.0, 3.99, -4.0, 0.3, -1.7]]),
       'castBin_round_12_9': np.array([b.castBin(v, quantization='Round') for v in [0.08, -0.08, 1.0, 3.99, -4.0, 0.3, -1.7]]),
       'peakfit': np.array([b.peakfit(1, 3, 2), b.peakfit(1, 2, 3), b.peakfit(-5., -9., -6.)]),
       }
np.savez_compressed(os.path.join(here, 'utils_bin_py3.npz'), **res)
print('golden fixtures written to', here)
