"""Golden outputs of the reference's OWN analysis code, run in the dev container (tests/golden/analysis_golden.npz).

    python tests/golden/make_golden_analysis.py [/root/reference]

The modules cannot be imported as they are (Python-2 syntax; PyTables / PyQt4 / matplotlib at module level), so this
script reads the source text of the functions from the reference tree AT RUN TIME, applies the two mechanical
Python-2 -> 3 edits they need (`print x` -> `print(x)`, `xrange` -> `range`), and executes them against in-memory
stand-ins for the HDF5 file objects / the Qt base class.  No reference source is copied into this repository; only
the numerical OUTPUTS are stored.  /root/reference does not exist on the GPU box, so the fixtures are committed.

  MakeTemplate            DataReadout/ReadoutControls/lib/pulses.py:239-427
  image_Worker methods    DataReadout/ReadoutControls/ArconsDashboard.py:1282-1384 (setup_thread, subtract_sky,
                          calc_mean_energy, calculate_SNR) driven as image_Worker.run does (:1442-1490)
"""
import hashlib
import os
import re
import sys
import textwrap
import time
import warnings

import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(here)))
from oracle import template as otpl                    # only for the synthetic INPUT (fake_pulses)


def py2to3(src):
    out = []
    for line in src.splitlines():
        m = re.match(r'^(\s*)print\s+(.*)$', line)
        if m and not m.group(2).startswith('('):
            line = '%sprint(%s)' % (m.group(1), m.group(2))
        out.append(line.replace('xrange', 'range'))
    return '\n'.join(out) + '\n'


def extract(path, start_pat, stop_pat):
    lines = open(path).read().splitlines()
    i0 = next(i for i, l in enumerate(lines) if re.match(start_pat, l))
    i1 = next(i for i in range(i0 + 1, len(lines)) if re.match(stop_pat, lines[i]))
    return '\n'.join(lines[i0:i1]) + '\n'


# ------------------------------------------------------------------ MakeTemplate
class _Row(dict):
    def append(self):
        self.table.rows.append(dict(self))


class _Table:
    def __init__(self):
        self.rows = []
        self.row = _Row()
        self.row.table = self


class _Node:
    pass


class _InFile:
    def __init__(self, I, Q):
        dat = np.zeros(len(I), dtype=[('I', 'f4', (2000,)), ('Q', 'f4', (2000,))])
        dat['I'], dat['Q'] = I, Q
        self.dat = dat
        grp = _Node()
        grp._v_name = 'r1p0'
        grp.iqpulses = _Node(); grp.iqpulses.read = lambda: dat
        grp.iqsweep = _Node(); grp.iqsweep.copy = lambda newparent=None: None
        r1 = _Node()
        r1._f_walkGroups = lambda: iter([r1, grp])
        self.root = _Node(); self.root.r1 = r1

    def close(self):
        pass


class _OutFile:
    def __init__(self):
        self.tables = []

    def createGroup(self, *a, **k):
        return _Node()

    def createTable(self, *a, **k):
        t = _Table()
        self.tables.append(t)
        return t

    def close(self):
        pass


def run_make_template(I, Q):
    src = py2to3(extract(os.path.join(ref, 'DataReadout', 'ReadoutControls', 'lib', 'pulses.py'),
                         r'^def MakeTemplate\(', r'^def FakeTemplateData\('))
    infile, outfile = _InFile(I, Q), _OutFile()
    ns = dict(np=np, time=time, plot=lambda *a, **k: None, show=lambda *a, **k: None,
              Filters=lambda **k: None, PulseAnalysis=object,
              openFile=lambda name, mode='r', title='': infile if mode == 'r' else outfile)
    exec(compile(src, 'pulses.py:MakeTemplate', 'exec'), ns)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ns['MakeTemplate']('fake.h5')
    row = outfile.tables[0].rows[0]
    return row, infile.dat


# ------------------------------------------------------------------ image_Worker
def run_image_worker(darray, sky, bintype, spectrum_pixel):
    src = py2to3(extract(os.path.join(ref, 'DataReadout', 'ReadoutControls', 'ArconsDashboard.py'),
                         r'^class image_Worker\(', r'^class timer_Worker\('))
    cons = extract(os.path.join(ref, 'DataReadout', 'ReadoutControls', 'ArconsDashboard.py'), r'^c = ', r'^h = ') + \
        extract(os.path.join(ref, 'DataReadout', 'ReadoutControls', 'ArconsDashboard.py'), r'^h = ', r'^\S*$|^[a-zA-Z#]')

    class QThread:
        def __init__(self, parent=None):
            pass

    ns = dict(QThread=QThread, SIGNAL=lambda s: s, numXPixel=44, numYPixel=46, ndarray=np.ndarray, median=np.median,
              sqrt=np.sqrt, arange=np.arange, reshape=np.reshape)
    exec(cons, ns)
    exec(compile(src, 'ArconsDashboard.py:image_Worker', 'exec'), ns)
    w = ns['image_Worker'](None, 44, 46)
    w.emit = lambda *a: None
    w.bintype = bintype
    # setup_thread hard-codes bintype = "wavelength" (:1301): re-run its arithmetic for the other type by hand-set attributes
    w.setup_thread()
    if bintype == 'energy':
        w.bintype = 'energy'
        w.binmin, w.binmax = w.Emin, w.Emax
        w.dE = (w.binmax - w.binmin) / 10.
        w.E0 = w.binmin + w.dE / 2.
        for i in range(1, 10):
            setattr(w, 'E%d' % i, getattr(w, 'E%d' % (i - 1)) + w.dE)
    w.sky_subtraction = sky
    w.spectrum_pixel = list(spectrum_pixel)
    # the body of image_Worker.run between unpack_file and the image output (:1453-1477, :1490-1500)
    for i in range(10):
        setattr(w, 'C%d' % i, darray[:, i])
    w.medians = [np.median(getattr(w, 'C%d' % i)) for i in range(10)]
    if w.sky_subtraction:
        w.subtract_sky(w.medians)
    for m in range(w.total_pix):
        w.pc[m] = sum(getattr(w, 'C%d' % i)[m] for i in range(10))
    with np.errstate(all='ignore'):
        w.calc_mean_energy()
    totalcounts = [0] * 10
    for p in w.spectrum_pixel:
        for i in range(10):
            totalcounts[i] += getattr(w, 'C%d' % i)[p]
    w.calculate_SNR(totalcounts, w.medians, len(w.spectrum_pixel))
    E = [getattr(w, 'E%d' % i) for i in range(10)]
    return dict(E=np.array(E), medians=np.array(w.medians, dtype=np.float64), pc=np.array(w.pc, dtype=np.int64),
                me=np.array(w.me, dtype=np.float64), totalcounts=np.array(totalcounts, dtype=np.int64),
                SNR=np.array(w.SNR, dtype=np.float64), integrated_SNR=float(w.integrated_SNR))


if __name__ == '__main__':
    out = {}
    N_PULSES, SEED = 1300, 1
    I, Q = otpl.fake_pulses(N_PULSES, seed=SEED)
    out['tpl_input_sha256'] = np.array(hashlib.sha256(I.tobytes() + Q.tobytes()).hexdigest())
    out['tpl_params'] = np.array([N_PULSES, SEED])
    row, dat = run_make_template(I.copy(), Q.copy())
    out['tpl_count'] = np.array(row['count'])
    out['tpl_flag'] = np.array(row['flag'])
    out['tpl_pstart'] = np.asarray(row['pstart']).reshape(-1)
    out['tpl_phasetemplate'] = np.asarray(row['phasetemplate'])
    out['tpl_phasenoise'] = np.asarray(row['phasenoise'])
    out['tpl_phasenoiseidx'] = np.asarray(row['phasenoiseidx'])
    out['tpl_shifted_rows_sha256'] = np.array(hashlib.sha256(dat['I'].tobytes() + dat['Q'].tobytes()).hexdigest())
    rng = np.random.default_rng(12)
    n_pix = 44 * 46
    darray = rng.poisson(rng.uniform(5, 60, (n_pix, 1)) * np.linspace(1.5, 0.5, 10)[None, :]).astype(np.int64)
    darray[7] = 0
    out['iw_darray'] = darray.astype(np.uint32)
    sel = [3, 50, 51, 900, 2001]
    out['iw_spectrum_pixel'] = np.array(sel)
    for sky, bt in ((False, 'wavelength'), (True, 'wavelength'), (True, 'energy')):
        r = run_image_worker(darray.copy(), sky, bt, sel)
        for k, v in r.items():
            out['iw_%d_%s_%s' % (int(sky), bt, k)] = np.asarray(v)
    np.savez_compressed(os.path.join(here, 'analysis_golden.npz'), **out)
    print('wrote analysis_golden.npz:', {k: np.asarray(v).shape for k, v in out.items() if not k.startswith('iw_darray')})
