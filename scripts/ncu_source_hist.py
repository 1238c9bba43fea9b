"""Per-opcode instruction histogram and shared-memory / stall hot spots of ONE kernel from an .ncu-rep source page.

    python scripts/ncu_source_hist.py REP.ncu-rep UNITS [--top N]

UNITS = work units the captured launch processed (ADC samples, photon words ...): the histogram is printed as warp-level
SASS instructions x 32 lanes per unit, i.e. thread instructions per unit.  Needs `ncu --import-source on` at capture time.
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def main():
    rep, units = sys.argv[1], float(sys.argv[2])
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 12
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    head = rows[hi]
    col = {k: i for i, k in enumerate(head)}
    body = [r for r in rows[hi + 1:] if len(r) == len(head)]
    f = lambda r, k: float(r[col[k]] or 0)
    by_op = defaultdict(lambda: [0.0, 0.0, 0.0])
    tot_inst = tot_samp = 0.0
    for r in body:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[col['Source']])
        op = m.group(2) if m else '?'
        base = op.split('.')[0]
        if base in ('LDS', 'STS', 'LDG', 'STG', 'ATOMS', 'RED', 'ATOMG'):
            base = '.'.join(op.split('.')[:1]) + ('.' + op.split('.')[-1] if op.split('.')[-1] in ('64', '128', 'U16', 'S16', 'U8') else '')
        ex = f(r, 'Instructions Executed')
        by_op[base][0] += ex
        by_op[base][1] += f(r, '# Samples')
        by_op[base][2] += f(r, 'L1 Wavefronts Shared')
        tot_inst += ex
        tot_samp += f(r, '# Samples')
    print('# %s: %.0f warp instructions, %.2f thread instructions per unit' % (rep, tot_inst, tot_inst * 32 / units))
    print('%-14s %12s %9s %9s %12s' % ('opcode', 'thread/unit', 'share', 'samples%', 'smem wf/unit'))
    for op, (ex, sm, wf) in sorted(by_op.items(), key=lambda kv: -kv[1][0]):
        if ex * 32 / units < 0.05:
            continue
        print('%-14s %12.2f %8.1f%% %8.1f%% %12.3f' % (op, ex * 32 / units, 100 * ex / tot_inst, 100 * sm / max(tot_samp, 1), wf / units))
    stall_cols = [k for k in head if k.startswith('stall_') and 'Not Issued' not in k]
    tot = {k: sum(f(r, k) for r in body) for k in stall_cols}
    s = sum(tot.values())
    print('# stall samples by reason: ' + ', '.join('%s %.1f%%' % (k[6:], 100 * v / s) for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v / s > 0.01))
    print('# instructions with excessive shared wavefronts (bank conflicts):')
    exc = sorted(body, key=lambda r: -f(r, 'L1 Wavefronts Shared Excessive'))[:top]
    for r in exc:
        if f(r, 'L1 Wavefronts Shared Excessive') > 0:
            print('  %s  %-44s wf %.3g ideal %.3g excessive %.3g' % (r[col['Address']][-5:], r[col['Source']][:44], f(r, 'L1 Wavefronts Shared'),
                                                                f(r, 'L1 Wavefronts Shared Ideal'), f(r, 'L1 Wavefronts Shared Excessive')))
    print('# top stall instructions:')
    for r in sorted(body, key=lambda r: -f(r, '# Samples'))[:top]:
        why = sorted(((f(r, k), k[6:]) for k in stall_cols), reverse=True)[:2]
        print('  %s  %-44s samples %.0f (%.1f%%)  %s' % (r[col['Address']][-5:], r[col['Source']][:44], f(r, '# Samples'), 100 * f(r, '# Samples') / tot_samp,
                                                   ', '.join('%s %.0f' % (k, v) for v, k in why)))


if __name__ == '__main__':
    main()
