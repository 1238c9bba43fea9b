"""Golden outputs of the reference's OWN AppForm methods, executed in the dev container (tests/golden/refrun_golden.npz).

    python tests/golden/make_golden_refrun.py [/root/reference]

The GUI modules cannot be imported (Python-2 syntax, PyQt4 / corr / matplotlib at module level), so this script reads
the source text of individual METHODS from the reference tree AT RUN TIME and executes them:

  * text edits: tabs expanded, `print x` -> `print(x)`, `xrange` -> `range`;
  * AST edit:   every `/` becomes a call to a Python-2 division (integer floor division on ints, true division else);
  * name space: Python-2 `round` (half away from zero, float result), `map` returning a list, `unicode`, `raw_input`,
                a `struct` whose pack / unpack work on Python-2 style byte strings (latin-1 text);
  * `self`:     a plain object holding the attributes the GUI would hold; unknown attributes (widgets, plot axes)
                are mocks; `self.roach` records every write / serves canned reads.

No reference source is stored in this repository, only the numerical OUTPUTS of the runs.

  ROACH_Setup_DAC.py   freqCombLUT :396-455, define_DAC_LUT :457-483, define_DDS_LUT :485-511, select_bins :513-529,
                       write_LUTs :531-557                                   (multi-tone comb, 12 tones, N = 2^12)
  pulse_triggering_v2.py  twos_comp :22-26 and the trigger loop :102-174 (rolling mean + I/Q snapshot decode)
  ArconsDashboard.py   StartQt4.make_image :633-723 (7 seconds: sky taking, sky subtraction, integration, flat field)
  pulse_triggering.py  the block-mean trigger :104-208
  Utils/bin.py, Utils/binTools.py  whole modules (extractBin, castBin, peakfit, masks) with Python-2 division
  ROACH_Pulses.py      the trigger loop of contsnapshot :625-725
  ROACH_Pulses.py      loadFIRcoeffs :59-111, loadIQcenters :948-956, loadThresholds :211-299, find_nearest, readPulses :782-919
"""
import ast
import builtins
import math
import os
import re
import struct as _struct
import sys
import tempfile
import types
import warnings
from unittest import mock

import numpy

ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
here = os.path.dirname(os.path.abspath(__file__))
CC = os.path.join(ref, 'DataReadout', 'ChannelizerControls')


# ------------------------------------------------------------------ source -> executable methods
def method_source(path, name):
    lines = open(path).read().expandtabs(8).splitlines()
    i0 = next(i for i, l in enumerate(lines) if re.match(r'^    def %s\(' % name, l))
    i1 = next((i for i in range(i0 + 1, len(lines)) if re.match(r'^    def |^\S', lines[i])), len(lines))
    out = []
    for line in lines[i0:i1]:
        m = re.match(r'^(\s*)print\s+(.*)$', line)
        if m and not m.group(2).startswith('('):
            line = '%sprint(%s)' % (m.group(1), m.group(2))
        elif re.match(r'^\s*print\s*$', line):
            line = line.replace('print', 'print()')
        out.append(line.replace('xrange', 'range'))
    return '\n'.join(out) + '\n'


class _Py2Div(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(ast.Call(func=ast.Name(id='_py2div', ctx=ast.Load()), args=[node.left, node.right],
                                              keywords=[]), node)
        return node


def _py2div(a, b):
    if isinstance(a, (int, numpy.integer)) and isinstance(b, (int, numpy.integer)) and not isinstance(a, bool):
        return a // b
    return a / b


def _py2round(x, n=0):
    m = 10 ** n
    v = x * m
    r = math.floor(abs(v) + 0.5) * (1 if v >= 0 else -1)
    return float(r) / m


class _Struct:                       # Python-2 byte strings as latin-1 text
    @staticmethod
    def pack(fmt, *a):
        return _struct.pack(fmt, *[int(v) for v in a]).decode('latin-1')

    @staticmethod
    def unpack(fmt, s):
        return _struct.unpack(fmt, s.encode('latin-1') if isinstance(s, str) else s)


def build_class(path, names, extra_ns=None):
    body = ''.join(method_source(path, n) for n in names)
    tree = ast.parse('class Ref(object):\n' + body)
    tree = ast.fix_missing_locations(_Py2Div().visit(tree))
    printed = []
    ns = dict(numpy=numpy, struct=_Struct, os=os, time=mock.MagicMock(), math=math, _py2div=_py2div, round=_py2round,
              map=lambda f, *a: list(builtins.map(f, *a)), unicode=str, raw_input=lambda *a: 'y',
              print=lambda *a, **k: printed.append(a), datetime=mock.MagicMock(), pickle=mock.MagicMock(),
              roachNo=0, __file__=os.path.join(CC, 'x.py'))
    ns.update(extra_ns or {})
    exec(compile(tree, os.path.basename(path), 'exec'), ns)
    return ns['Ref'], printed, ns


class Roach:
    def __init__(self, reads=None, read_ints=None):
        self.log, self.reads, self.read_ints = [], dict(reads or {}), dict(read_ints or {})

    def write_int(self, name, value, *a, **k):
        self.log.append(('write_int', name, int(value)))

    def write(self, name, data, *a, **k):
        data = data.encode('latin-1') if isinstance(data, str) else bytes(data)
        self.log.append(('write', name, data))

    def read(self, name, size, *a, **k):
        v = self.reads[name]
        d = v.pop(0) if isinstance(v, list) else v
        return d[:size].decode('latin-1')

    def read_int(self, name, *a, **k):
        return self.read_ints[name].pop(0)


class Self:
    """attribute bag: real values where set, mocks for widgets / axes"""

    def __getattr__(self, k):
        m = mock.MagicMock()
        object.__setattr__(self, k, m)
        return m


def text_widget(s):
    w = mock.MagicMock()
    w.text.return_value = s
    w.toPlainText.return_value = s
    return w


# ------------------------------------------------------------------ the runs
def run_setup_dac():
    Ref, printed, ns = build_class(os.path.join(CC, 'ROACH_Setup_DAC.py'),
                                   ['freqCombLUT', 'define_DAC_LUT', 'define_DDS_LUT', 'select_bins', 'write_LUTs'])
    fs, N = 512e6, 2 ** 12
    res = fs / N
    lo = 5.0e9
    rng = numpy.random.default_rng(7)
    k = numpy.sort(rng.choice(numpy.arange(-N // 2 + 8, N // 2 - 8), 12, replace=False))
    dac_freqs = [lo + float(v) * res for v in k]
    attens = rng.integers(0, 20, 12).astype(float)
    s = Self()
    for m in Ref.__dict__:
        if not m.startswith('__'):
            setattr(s, m, types.MethodType(getattr(Ref, m), s))
    s.sampleRate, s.freqRes = fs, res
    s.attens = numpy.array(attens)
    s.minimumAttenuation, s.previous_scale_factor, s.last_scale_factor = 10, 5000.0, None
    s.textbox_offset = text_widget('0')
    s.textbox_customScale = text_widget('1')
    s.cbox_keepScaleFactor = mock.MagicMock(); s.cbox_keepScaleFactor.isChecked.return_value = False
    s.cbox_useScaleFactor = mock.MagicMock(); s.cbox_useScaleFactor.isChecked.return_value = False
    s.textedit_DACfreqs = text_widget(' '.join(repr(f) for f in dac_freqs))
    s.textbox_loFreq = text_widget(repr(lo))
    s.roach = Roach()
    s.dacStatus = 'off'
    tmp = tempfile.mkdtemp()
    s.textbox_saveDir = text_widget(tmp)
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            s.define_DAC_LUT()
            phase = list(numpy.random.default_rng(3).uniform(-numpy.pi, numpy.pi, 256))
            s.define_DDS_LUT(phase)
            # write_LUTs writes the byte string through a text-mode file: give it a binary-safe open
            ns['open'] = lambda p, mode='r': open(p, mode, encoding='latin-1', newline='')
            s.write_LUTs()
    finally:
        os.chdir(cwd)
    dram = [d for op, n, d in s.roach.log if op == 'write' and n == 'dram_memory'][0]
    bins = [v for op, n, v in s.roach.log if op == 'write_int' and n == 'bins']
    return dict(setup_N=numpy.array(N), setup_lo=numpy.array(lo), setup_dac_freqs=numpy.array(dac_freqs),
                setup_attens=attens, setup_dds_phase=numpy.array(phase), setup_freqs_dac=numpy.array(s.freqs_dac),
                setup_I_dac=numpy.asarray(s.I_dac, dtype=numpy.int64), setup_Q_dac=numpy.asarray(s.Q_dac, dtype=numpy.int64),
                setup_I_dds=numpy.asarray(s.I_dds, dtype=numpy.int64), setup_Q_dds=numpy.asarray(s.Q_dds, dtype=numpy.int64),
                setup_scale_factor=numpy.array(s.scale_factor), setup_bins=numpy.array(bins),
                setup_dram=numpy.frombuffer(dram, dtype=numpy.uint8).copy())


def run_pulses():
    out = {}
    path = os.path.join(CC, 'ROACH_Pulses.py')
    Ref, printed, ns = build_class(path, ['loadFIRcoeffs', 'loadIQcenters', 'find_nearest', 'loadThresholds', 'readPulses'])

    def new_self():
        s = Self()
        for m in Ref.__dict__:
            if not m.startswith('__'):
                setattr(s, m, types.MethodType(getattr(Ref, m), s))
        return s
    # ---- loadFIRcoeffs: 3 driven channels (one deleted), the reference's 30 us matched filter
    s = new_self()
    s.textedit_DACfreqs = text_widget('1.0 2.0 3.0')
    s.fir = list(numpy.loadtxt(os.path.join(CC, 'LUT', 'matched_30us.txt')))
    s.zeroChannels = [0, 1, 0] + [0] * 253
    s.roach = Roach()
    s.loadFIRcoeffs()
    regs = [(n, d) for op, n, d in s.roach.log if op == 'write']
    out['fir_reg_names'] = numpy.array([n for n, d in regs[:39]])
    out['fir_reg_bytes'] = numpy.array([numpy.frombuffer(d, dtype=numpy.uint8) for n, d in regs[:39]])
    out['fir_load_coeff'] = numpy.array([v for op, n, v in s.roach.log if op == 'write_int' and n == 'FIR_load_coeff'][:78])
    out['fir_inactive_bytes'] = numpy.frombuffer(regs[39][1], dtype=numpy.uint8).copy()
    # ---- loadIQcenters
    s = new_self()
    rng = numpy.random.default_rng(5)
    centers = (rng.integers(-30000, 30000, 256) + 1j * rng.integers(-30000, 30000, 256)) * 1.0 + (0.37 - 0.81j)
    s.iq_centers = numpy.array(centers)
    s.roach = Roach()
    s.loadIQcenters()
    out['iq_centers_in'] = numpy.array(centers)
    out['iq_center_writes'] = numpy.array([v for op, n, v in s.roach.log if op == 'write_int' and n == 'conv_phase_centers'])
    out['iq_center_regs'] = numpy.array(sorted(set(n for op, n, v in s.roach.log)))
    # ---- loadThresholds: 2 channels x 10 snapshots of 1024 words
    s = new_self()
    s.textedit_DACfreqs = text_widget('1.0 2.0')
    s.customThresholds = numpy.array([360.0] * 256)
    snaps = []
    raw_all = []
    for chn in range(2):
        for st in range(10):
            raw = numpy.clip(numpy.round(rng.normal(1500 * (chn + 1), 300 + 200 * chn, 2048) -
                                         numpy.abs(rng.normal(0, 2500, 2048)) * (rng.random(2048) < 0.05)), -32768, 32767)
            words = numpy.empty((1024, 2), dtype='>i2')
            words[:, 1] = raw[0::2]; words[:, 0] = raw[1::2]
            snaps.append(words.tobytes())
            raw_all.append(raw.astype(numpy.int64))
    s.roach = Roach(reads={'snapPhase_bram': snaps})
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        s.loadThresholds()
    out['thr_raw_phase'] = numpy.array(raw_all).reshape(2, -1)
    out['thr_thresholds_deg'] = numpy.array(s.thresholds, dtype=numpy.float64)
    out['thr_medians_deg'] = numpy.array(s.medians, dtype=numpy.float64)
    out['thr_capture_threshold'] = numpy.array([v for op, n, v in s.roach.log if op == 'write_int' and n == 'capture_threshold'])
    # ---- readPulses: 3 steps incl. a ring wrap
    s = new_self()
    n = 2 ** 14
    ch = rng.integers(0, 256, n)
    peak = rng.integers(0, 4096, n); p1 = rng.integers(0, 4096, n); base = rng.integers(0, 4096, n); ts = rng.integers(0, 2 ** 20, n)
    w = (ch.astype(numpy.uint64) << numpy.uint64(56)) | (peak.astype(numpy.uint64) << numpy.uint64(44)) | \
        (p1.astype(numpy.uint64) << numpy.uint64(32)) | (base.astype(numpy.uint64) << numpy.uint64(20)) | ts.astype(numpy.uint64)
    b0 = (w & numpy.uint64(0xFFFFFFFF)).astype('>u4').tobytes()
    b1 = (w >> numpy.uint64(32)).astype('>u4').tobytes()
    pairs = [(100, 5000), (16000, 300), (7, 7), (300, 900), (900, 2000), (2000, 2100), (16383, 1), (1, 40), (40, 4000), (4000, 9000)]
    s.textedit_DACfreqs = text_widget('1.0 2.0')
    s.textbox_seconds = text_widget('1')            # steps = int(seconds*10) = 10
    s.textbox_channel = text_widget('3')
    s.roach = Roach(reads={'pulses_bram0': b0, 'pulses_bram1': b1}, read_ints={'pulses_addr': [a for pr in pairs for a in pr]})
    saved = []
    ns['numpy'] = types.SimpleNamespace(**{k: getattr(numpy, k) for k in dir(numpy) if not k.startswith('__')})
    ns['numpy'].savetxt = lambda name, arr, **k: saved.append(numpy.array(arr, dtype=numpy.float64))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        s.readPulses()
    bars = [c.args for c in s.axes1.bar.call_args_list]
    out['rp_words'] = w
    out['rp_pairs'] = numpy.array(pairs)
    out['rp_hgBase'] = numpy.asarray(bars[0][1]); out['rp_hgPeak'] = numpy.asarray(bars[1][1])
    out['rp_hgPeakSubBase'] = numpy.asarray(bars[2][1])
    out['rp_peaksCh_deg'] = saved[0]; out['rp_timesCh'] = saved[1]
    cc = [a for a in printed if a and a[0] == 'total counts by channel: ']
    out['rp_channel_count'] = numpy.array(cc[-1][1])
    return out


def run_trigger_script():
    """The trigger loop of pulse_triggering_v2.py:102-174 (rolling-mean trigger with the 40-bit I/Q snapshot decode inside
    the hit branch), dedented and executed with the script's own variable names."""
    path = os.path.join(CC, 'pulse_triggering_v2.py')
    lines = open(path).read().expandtabs(8).splitlines()
    t0 = next(i for i, l in enumerate(lines) if re.match(r'^def twos_comp\(', l))
    t1 = next(i for i in range(t0 + 1, len(lines)) if lines[i].strip() and not lines[i].startswith(' '))
    i0 = next(i for i, l in enumerate(lines) if re.match(r'^\s+george = 0\s*$', l))
    i1 = next(i for i in range(i0, len(lines)) if re.match(r'^\s+bob = bob \+ 1\s*$', lines[i]))
    ind = len(lines[i0]) - len(lines[i0].lstrip())
    body = '\n'.join(l[ind:] if l.strip() else '' for l in lines[i0:i1 + 1])
    src = '\n'.join(lines[t0:t1]) + '\n' + body + '\n'
    tree = ast.fix_missing_locations(_Py2Div().visit(ast.parse(src)))
    rng = numpy.random.default_rng(21)
    n = 16384
    x = rng.normal(0, 4.0, n)
    t = numpy.arange(n)
    for p0 in numpy.nonzero(rng.random(n) < 0.0015)[0]:
        x[p0:] -= rng.uniform(20, 120) * numpy.exp(-(t[p0:] - p0) / 30.0)
    L_IQ = 64
    snap = rng.integers(0, 256, 4 * L_IQ, dtype=numpy.uint8).tobytes()
    hits, saved_iq = [], []
    out = {}
    for (M, L, thr) in ((20, 1000, 25.0), (10, 300, 15.0)):
        ns = dict(np=numpy, _py2div=_py2div, datetime=__import__('datetime').datetime, ord=ord,
                  meanlength=M, pulselength=L, phase_threshold=thr, number_of_phase_values=n, phasevalues=list(x),
                  steps_IQ=1, L_IQ=L_IQ, bin_data_IQ='', bin_data_IQ_ord=[], bin_data_IQ_hex=[], total_pulses=0,
                  final_pulse_count=0, savedirIQ='', savedirPhase='', roach=Roach(reads={'conv_phase_snapIQ_bram': snap}))
        cur = []

        def savetxt(name, arr, fmt=None, _ns=ns, _cur=cur):
            if name.startswith('pulse_'):
                _cur.append(_ns['bob'])
            elif not saved_iq:
                saved_iq.append(numpy.array(arr))
        ns['np'] = types.SimpleNamespace(mean=numpy.mean, savetxt=savetxt, column_stack=numpy.column_stack)
        exec(compile(tree, 'pulse_triggering_v2.py', 'exec'), ns)
        out['trig_hits_%d_%d' % (M, L)] = numpy.array(cur)
    out['trig_phase'] = x
    out['trig_iq_snapshot'] = numpy.frombuffer(snap, dtype=numpy.uint8).copy()
    out['trig_iq_decoded'] = saved_iq[0]
    return out


def run_block_trigger_script():
    """The block-mean trigger of pulse_triggering.py:104-208 (+360 wrap of negative samples, means of fixed blocks,
    window [bob-100, bob+300), hold-off 200), dedented and executed with the script's own variable names."""
    path = os.path.join(CC, 'pulse_triggering.py')
    lines = open(path).read().expandtabs(8).splitlines()
    t0 = next(i for i, l in enumerate(lines) if re.match(r'^def twos_comp\(', l))
    t1 = next(i for i in range(t0 + 1, len(lines)) if lines[i].strip() and not lines[i].startswith(' '))
    i0 = next(i for i, l in enumerate(lines) if re.match(r'^\s+for k in range\(number_of_phase_values\):\s*$', l))
    i1 = next(i for i in range(i0, len(lines)) if re.match(r'^\s+bob = bob \+ 1\s*$', lines[i]))
    ind = len(lines[i0]) - len(lines[i0].lstrip())
    body = '\n'.join(l[ind:] if l.strip() else '' for l in lines[i0:i1 + 1])
    src = '\n'.join(lines[t0:t1]) + '\n' + body + '\n'
    tree = ast.fix_missing_locations(_Py2Div().visit(ast.parse(src)))
    rng = numpy.random.default_rng(22)
    n = 20000
    out = {}
    L_IQ = 64
    snap = rng.integers(0, 256, 4 * L_IQ, dtype=numpy.uint8).tobytes()
    for tag, offset, A, thr in (('a', 3.0, 128, 25.0), ('b', -2.0, 64, 15.0), ('c', 150.0, 100, 40.0)):
        x = rng.normal(offset, 4.0, n)
        t = numpy.arange(n)
        for p0 in numpy.nonzero(rng.random(n) < 0.002)[0]:
            x[p0:] -= rng.uniform(20, 120) * numpy.exp(-(t[p0:] - p0) / 30.0)
        ns = dict(_py2div=_py2div, datetime=__import__('datetime').datetime, ord=ord, averagelength=A,
                  numberofaverages=n // A, phase_threshold=thr, number_of_phase_values=n, phasevalues=list(x),
                  steps_IQ=1, L_IQ=L_IQ, bin_data_IQ='', bin_data_IQ_ord=[], bin_data_IQ_hex=[], total_pulses=0,
                  final_pulse_count=0, savedirIQ='', savedirPhase='', roach=Roach(reads={'conv_phase_snapIQ_bram': snap}))
        cur = []

        def savetxt(name, arr, fmt=None, _ns=ns, _cur=cur):
            if name.startswith('pulse_'):
                _cur.append(_ns['bob'])
        ns['np'] = types.SimpleNamespace(mean=numpy.mean, zeros=numpy.zeros, ones=numpy.ones, savetxt=savetxt,
                                         column_stack=numpy.column_stack)
        exec(compile(tree, 'pulse_triggering.py', 'exec'), ns)
        out['btrig_%s_phase' % tag] = x
        out['btrig_%s_params' % tag] = numpy.array([A, thr])
        out['btrig_%s_hits' % tag] = numpy.array(cur)
    return out


def run_utils_bin():
    """Utils/bin.py and Utils/binTools.py executed with Python-2 division (under Python 3 extractBin / castBin are
    silently wrong because of `int(value)/2**(nBits-1)`)."""
    out = {}
    rng = numpy.random.default_rng(31)
    vals = [int(v) for v in rng.integers(0, 2 ** 40, 200)]
    fl = [float(v) for v in numpy.concatenate([rng.uniform(-4, 4, 150), rng.uniform(-40, 40, 50)])]
    ys = rng.uniform(-100, 100, (50, 3))
    for fname in ('bin.py', 'binTools.py'):
        src = open(os.path.join(ref, 'Utils', fname)).read().expandtabs(8)
        src = '\n'.join(re.sub(r'^(\s*)print\s+(.*)$', r'\1print(\2)', l) for l in src.splitlines()) + '\n'
        tree = ast.fix_missing_locations(_Py2Div().visit(ast.parse(src)))
        ns = dict(_py2div=_py2div, round=_py2round, __name__='refutils')
        exec(compile(tree, fname, 'exec'), ns)
        tag = fname[:-3]
        out['ub_values'] = numpy.array(vals, dtype=numpy.uint64)
        out['ub_floats'] = numpy.array(fl)
        for (nb, bp, after) in ((12, 9, 0), (12, 9, 20), (16, 13, 4), (18, 16, 0)):
            if 'extractBin' in ns:
                for fmt in ('rad', 'deg'):
                    try:
                        out['%s_extract_%d_%d_%d_%s' % (tag, nb, bp, after, fmt)] = numpy.array(
                            [ns['extractBin'](v, nb, bp, after, fmt) for v in vals], dtype=numpy.float64)
                    except TypeError:
                        out['%s_extract_%d_%d_%d' % (tag, nb, bp, after)] = numpy.array(
                            [ns['extractBin'](v, nb, bp, after) for v in vals], dtype=numpy.float64)
                        break
            if 'castBin' in ns and after == 0:
                for q in ('Truncate', 'Round'):
                    for fmt in ('uint', 'rad', 'deg'):
                        try:
                            out['%s_cast_%d_%d_%s_%s' % (tag, nb, bp, q, fmt)] = numpy.array(
                                [ns['castBin'](v, nb, bp, q, fmt) for v in fl], dtype=numpy.float64)
                        except Exception:
                            pass
        for name in ('binMask', 'bitmask'):
            if name in ns:
                out['%s_%s' % (tag, name)] = numpy.array([ns[name](k) for k in range(1, 40)], dtype=numpy.uint64)
        if 'peakfit' in ns:
            out['ub_peakfit_in'] = ys
            out['%s_peakfit' % tag] = numpy.array([ns['peakfit'](*y) for y in ys])
        if 'bin12_9ToRad' in ns:
            out['%s_bin12_9ToRad' % tag] = numpy.array([ns['bin12_9ToRad'](v & 0xFFF) for v in vals], dtype=numpy.float64)
    return out


def run_dashboard_make_image():
    """StartQt4.make_image of ReadoutControls/ArconsDashboard.py:633-723 over 7 consecutive seconds: sky taking for
    the first two, then sky subtraction, a 3-second integration window and the flat field."""
    path = os.path.join(ref, 'DataReadout', 'ReadoutControls', 'ArconsDashboard.py')
    figs = []
    plt = mock.MagicMock()
    plt.figimage.side_effect = lambda img, **k: figs.append(numpy.array(img, dtype=numpy.float64))
    Ref, printed, ns = build_class(path, ['make_image'], extra_ns=dict(
        loadtxt=lambda f: f, shape=numpy.shape, flipud=numpy.flipud, reshape=numpy.reshape, sum=numpy.sum, sort=numpy.sort,
        where=numpy.where, plt=plt, numXPixel=44, numYPixel=46))
    rng = numpy.random.default_rng(17)
    rows, cols, secs = 46, 44, 7
    n_pix = rows * cols
    pixel_adr = rng.permutation(n_pix).reshape(rows, cols)
    counts = rng.poisson(rng.uniform(20, 900, n_pix), (secs, n_pix)).astype(numpy.int64)
    counts[:, 11] = 2600                                  # above the 2500-event cap
    counts[:, 12] = 2300
    flat = rng.uniform(0.8, 1.2, (rows, cols))
    s = Self()
    s.make_image = types.MethodType(Ref.make_image, s)
    s.nxpix, s.nypix = cols, rows
    s.counts = numpy.zeros((secs + 1, n_pix)); s.rotated_counts = numpy.zeros((secs + 1, rows, cols))
    s.image_time = 0
    s.taking_sky, s.skytime, s.skycount = True, 2, numpy.zeros((rows, cols))
    s.skyrate = numpy.zeros((rows, cols))
    s.sky_subtraction = False
    s.flatFactors = flat
    s.ui = mock.MagicMock()
    s.ui.int_time_spinBox.value.return_value = 3
    s.ui.contrast_mode.isChecked.return_value = False
    s.ui.brightpix.value.return_value = 5
    out = dict(dash_pixel_adr=pixel_adr, dash_counts=counts, dash_flat=flat)
    for t in range(secs):
        s.sky_subtraction = t >= 3
        s.ui.flat_field_radioButton.isChecked.return_value = t >= 5
        capped = numpy.minimum(counts[t], 2499)
        s.binfile = capped[pixel_adr].astype(numpy.uint16).astype(numpy.float64)      # the text image of second t
        s.make_image()
        out['dash_frame_%d' % t] = figs[-1]
        out['dash_vmax_%d' % t] = numpy.array(s.vmax)
        out['dash_redpix_%d' % t] = numpy.array(s.redpix)
    out['dash_skyrate'] = numpy.array(s.skyrate)
    return out


def run_contsnapshot_loop():
    """The trigger part of AppForm.contsnapshot, ROACH_Pulses.py:~625-725 (block means of 2^k samples, start 500, window
    [bob-500, bob+1500), hold-off 1000), dedented and executed with the method's own variable names."""
    path = os.path.join(CC, 'ROACH_Pulses.py')
    lines = open(path).read().expandtabs(8).splitlines()
    m0 = next(i for i, l in enumerate(lines) if re.match(r'^    def contsnapshot\(', l))
    i0 = next(i for i in range(m0, len(lines)) if re.match(r'^\s+phase_threshold = float\(self\.textbox_phasethreshold', lines[i]))
    i1 = next(i for i in range(i0, len(lines)) if re.match(r'^\s+bob = bob \+ 1\s*$', lines[i]))
    ind = len(lines[i0]) - len(lines[i0].lstrip())
    body = '\n'.join(l[ind:] if l.strip() else '' for l in lines[i0:i1 + 1]) + '\n'
    tree = ast.fix_missing_locations(_Py2Div().visit(ast.parse(body)))
    rng = numpy.random.default_rng(23)
    out = {}
    for tag, k, thr in (('a', 6, 25.0), ('b', 5, 40.0), ('c', 8, 15.0)):
        n = 2 ** 15
        x = rng.normal(60.0, 4.0, n)
        t = numpy.arange(n)
        for p0 in numpy.nonzero(rng.random(n) < 0.0012)[0]:
            x[p0:] -= rng.uniform(20, 120) * numpy.exp(-(t[p0:] - p0) / 30.0)
        hits = []
        ns = {}

        class Rec(list):
            def extend(self, it, _ns=ns):
                if not hits or hits[-1] != _ns['bob']:
                    hits.append(_ns['bob'])
        selfo = Self()
        selfo.textbox_phasethreshold = text_widget(repr(thr))
        selfo.textbox_averagelength = text_widget(str(k))
        selfo.textbox_pulsesavepath = text_widget('/tmp/x/')
        osm = mock.MagicMock(); osm.path.exists.return_value = True
        ns.update(dict(_py2div=_py2div, numpy=numpy, datetime=__import__('datetime').datetime, os=osm, self=selfo,
                       qdr_phase_values=x, nContsnapSamples=n, total_pulses=0, finalphasearray=Rec(), pulsenumberarray=Rec()))
        exec(compile(tree, 'ROACH_Pulses.py:contsnapshot', 'exec'), ns)
        out['ctrig_%s_phase' % tag] = x
        out['ctrig_%s_params' % tag] = numpy.array([2 ** k, thr])
        out['ctrig_%s_hits' % tag] = numpy.array(hits)
    return out


def function_source(path, name):
    """A module-level function of a Python-2 file as Python-3 text (print statements, xrange)."""
    lines = open(path).read().expandtabs(8).splitlines()
    i0 = next(i for i, l in enumerate(lines) if re.match(r'^def %s\(' % name, l))
    i1 = next((i for i in range(i0 + 1, len(lines)) if re.match(r'^\S', lines[i])), len(lines))
    out = []
    for line in lines[i0:i1]:
        m = re.match(r'^(\s*)print\s+(.*)$', line)
        if m and not m.group(2).startswith('('):
            line = '%sprint(%s)' % (m.group(1), m.group(2))
        out.append(line.replace('xrange', 'range'))
    return '\n'.join(out) + '\n'


def run_pulses_quicklook():
    """QuickLook of ReadoutControls/lib/pulses.py:210-236 executed on a stand-in for the observation file: a 32 x 32
    beammap of dataset names and, per pixel, one photon list per second whose length is what PacketMaster stored (the
    capped count).  The image handed to imshow is the golden output."""
    path = os.path.join(ref, 'DataReadout', 'ReadoutControls', 'lib', 'pulses.py')
    rng = numpy.random.default_rng(23)
    secs, npix = 6, 1024
    counts = numpy.minimum(rng.poisson(rng.uniform(5, 400, npix), (secs, npix)), 2499).astype(numpy.int64)
    counts[:, 77] = 2499                                   # a pixel at the cap
    adr = rng.permutation(npix).reshape(32, 32)
    names = [['/r%d/p%d/' % (a // 256, a % 256) for a in row] for row in adr]

    class Node:
        def __init__(self, value):
            self.value = value

        def read(self):
            return self.value

    class H5:
        def __init__(self):
            self.root = self
            self.beammap = self
            self.beamimage = Node(names)

        def _f_getChild(self, name):
            r, p_ = re.match(r'/r(\d+)/p(\d+)/', name).groups()
            pix = int(r) * 256 + int(p_)
            return Node([numpy.zeros(int(c), dtype=numpy.uint64) for c in counts[:, pix]])

        def close(self):
            pass

    shown = []
    plt = mock.MagicMock()
    plt.figure.return_value.add_subplot.return_value.imshow.side_effect = lambda img, **k: shown.append(numpy.array(img))
    ns = dict(np=numpy, openFile=lambda *a, **k: H5(), plt=plt)
    exec(compile(function_source(path, 'QuickLook'), 'pulses.py:QuickLook', 'exec'), ns)
    out = dict(ql_counts=counts, ql_adr=adr.astype(numpy.int32))
    for i, (t0, t1) in enumerate(((0, 6), (2, 5), (3, 4))):
        ns['QuickLook']('obs.h5', t0, t1)
        out['ql_span_%d' % i] = numpy.array([t0, t1])
        out['ql_skysub_%d' % i] = shown[-1]
    return out


if __name__ == '__main__':
    ql = run_pulses_quicklook()
    numpy.savez_compressed(os.path.join(here, 'quicklook_golden.npz'), **ql)
    print('wrote quicklook_golden.npz:', {k: numpy.asarray(v).shape for k, v in ql.items()})
    out = {}
    out.update(run_setup_dac())
    out.update(run_pulses())
    out.update(run_trigger_script())
    out.update(run_block_trigger_script())
    out.update(run_utils_bin())
    out.update(run_dashboard_make_image())
    out.update(run_contsnapshot_loop())
    numpy.savez_compressed(os.path.join(here, 'refrun_golden.npz'), **out)
    print('wrote refrun_golden.npz:', {k: numpy.asarray(v).shape for k, v in out.items()})
