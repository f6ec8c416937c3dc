#!/usr/bin/env python3
"""
Developer tool: dynamic instruction mix of one kernel launch from an `ncu --set full
--import-source on` report (read here, no GPU needed).

    ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv
    python tools/sass_mix.py src.csv [launch_index] [warp_steps]

Prints executed warp-instructions and stall samples per opcode (divided by `warp_steps` when
given: e.g. warps x steps of a rollout).
"""
import collections
import csv
import re
import sys


def main(path, launch=0, per=1.0):
    rows = list(csv.reader(open(path)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'hdr': None, 'rows': []}
            blocks.append(cur)
        elif cur is not None and cur['hdr'] is None:
            cur['hdr'] = r
        elif cur is not None and r:
            cur['rows'].append(r)
    b = blocks[launch]
    h = b['hdr']
    ci, cs, ce = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
    ex, st = collections.Counter(), collections.Counter()
    for r in b['rows']:
        m = re.match(r'\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+)', r[ci])
        if not m:
            continue
        op = m.group(1)
        if op in ('MUFU', 'HMMA', 'LDS', 'STS', 'LDG', 'STG', 'F2FP', 'LDSM'):
            op = r[ci].split()[0] if not r[ci].strip().startswith('@') else r[ci].split()[1]
            op = '.'.join(op.split('.')[:2])
        ex[op] += int(r[ce] or 0)
        st[op] += int(r[cs] or 0)
    tot, tots = sum(ex.values()), sum(st.values())
    print('# %s' % b['name'])
    print('# executed warp-instructions: %d (%.1f per unit), stall samples: %d' % (tot, tot / per, tots))
    print('%-14s %14s %10s %8s %8s' % ('opcode', 'executed', 'per unit', 'exec %', 'stall %'))
    for op, n in ex.most_common(40):
        print('%-14s %14d %10.2f %8.2f %8.2f' % (op, n, n / per, 100.0 * n / tot, 100.0 * st[op] / max(tots, 1)))
    # where in the kernel the instructions are executed: runs of consecutive SASS lines with the same execution count
    # (= basic blocks executed together), largest first -- "rows a..b  xN executions  instructions  per unit  first opcodes"
    runs, cur = [], None
    for k, r in enumerate(b['rows']):
        n = int(r[ce] or 0)
        if cur is not None and cur[2] == n:
            cur[1] = k
        else:
            cur = [k, k, n]
            runs.append(cur)
    print('# hottest straight-line regions (SASS rows, executions of the region, warp-instructions, per unit)')
    for a0, a1, n in sorted(runs, key=lambda c: -(c[1] - c[0] + 1) * c[2])[:24]:
        ops = ' '.join(re.sub(r'\s+', ' ', b['rows'][k][ci]).strip().split(' ')[0] for k in range(a0, min(a1 + 1, a0 + 6)))
        print('%5d..%-5d x%-11d %13d %8.2f   %s' % (a0, a1, n, (a1 - a0 + 1) * n, (a1 - a0 + 1) * n / per, ops))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, float(sys.argv[3]) if len(sys.argv) > 3 else 1.0)
