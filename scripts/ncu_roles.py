"""Per-ROLE view of the warp-specialised channelize_ws_kernel from an .ncu-rep captured with --import-source on: the SASS is
cut at the setmaxnreg instructions (FFT | PFB | CHAN + out-of-line wait loops); for every role the stall-reason mix of
the sampled warps, the instructions executed per warp and block of 8 frames by opcode, and the hottest instructions.

    python scripts/ncu_roles.py REP.ncu-rep BLOCKS_PER_CTA [--top N]
(BLOCKS_PER_CTA = blocks of 8 frames one CTA processes: 921 for bench.py's 8 boards x 2^25 samples on 148 SMs)
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def main():
    rep, blocks = sys.argv[1], float(sys.argv[2])
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 6
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    head = rows[hi]
    col = {k: i for i, k in enumerate(head)}
    body = [r for r in rows[hi + 1:] if len(r) == len(head)]
    f = lambda r, k: float(r[col[k]] or 0)
    cuts = [i for i, r in enumerate(body) if 'USETMAXREG' in r[col['Source']]]
    names = ['FFT (2 groups, 56 regs)', 'PFB (40 regs)', 'CHAN (96 regs) + the out-of-line wait loops of all roles']
    bounds = cuts + [len(body)]
    stall_cols = [k for k in head if k.startswith('stall_') and 'Not Issued' not in k]
    grid = sum(f(r, 'Instructions Executed') for r in body[:cuts[0]] if 'S2R' in r[col['Source']] and 'TID' in r[col['Source']])
    n_cta = grid / 32.0 if grid else 144.0
    for n in range(len(cuts)):
        reg = body[bounds[n]:bounds[n + 1]]
        tot = sum(f(r, '# Samples') for r in reg)
        t = {k: sum(f(r, k) for r in reg) for k in stall_cols}
        s = sum(t.values()) or 1
        print('== %s: %d stall samples' % (names[n], tot))
        print('   stall mix: ' + ', '.join('%s %.1f%%' % (k[6:], 100 * v / s) for k, v in sorted(t.items(), key=lambda kv: -kv[1]) if v / s > 0.02))
        by = defaultdict(float)
        for r in reg:
            src = r[col['Source']]
            m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.64|\.U16|\.S16)?)', src)
            by[m.group(2) if m else '?'] += f(r, 'Instructions Executed')
        div = n_cta * 8 * blocks * (2 if n == 0 else 1)          # FFT: two groups share the blocks
        print('   instructions per warp and block: ' + ', '.join('%s %.1f' % (k, v / div) for k, v in sorted(by.items(), key=lambda kv: -kv[1])[:22]))
        for r in sorted(reg, key=lambda r: -f(r, '# Samples'))[:top]:
            why = sorted(((f(r, k), k[6:]) for k in stall_cols), reverse=True)[:2]
            print('   %s %-46s %6.0f (%.1f%%) %s' % (r[col['Address']][-5:], r[col['Source']][:46], f(r, '# Samples'), 100 * f(r, '# Samples') / max(tot, 1),
                                                  ', '.join('%s %.0f' % (k, v) for v, k in why)))


if __name__ == '__main__':
    main()
