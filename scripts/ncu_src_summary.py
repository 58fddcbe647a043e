#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: executed instructions, stall-reason totals, hottest SASS lines.

    ncu -i prof.ncu-rep --page source --csv > src.csv ; python scripts/ncu_src_summary.py src.csv [top]
"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[hi]
    col = {n: i for i, n in enumerate(hdr)}
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    stall_cols = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]

    def num(r, n):
        try:
            return float(r[col[n]])
        except (ValueError, KeyError):
            return 0.0
    tot_inst = sum(num(r, 'Instructions Executed') for r in body)
    tot_samp = sum(num(r, '# Samples') for r in body)
    print(f'SASS lines {len(body)}  warp-instructions executed {tot_inst:.0f}  samples {tot_samp:.0f}')
    st = {n: sum(num(r, n) for r in body) for n in stall_cols}
    tot = sum(st.values()) or 1
    print('stall reasons (all samples):')
    for n, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]:
        print(f'  {n:28s} {v:10.0f}  {100 * v / tot:5.1f}%')
    print(f'hottest {top} SASS lines by samples:')
    for r in sorted(body, key=lambda r: -num(r, '# Samples'))[:top]:
        main_st = max(stall_cols, key=lambda n: num(r, n))
        print(f"  {r[col['Address']][-6:]} {num(r, '# Samples'):7.0f} {num(r, 'Instructions Executed'):9.0f} {main_st:18s} {r[col['Source']][:90]}")
    # opcode histogram by executed count
    ops = {}
    for r in body:
        op = r[col['Source']].split()
        if not op:
            continue
        name = op[1] if op[0].startswith('@') and len(op) > 1 else op[0]
        name = name.split('.')[0]
        ops[name] = ops.get(name, 0) + num(r, 'Instructions Executed')
    print('opcode histogram (warp-instructions executed):')
    for n, v in sorted(ops.items(), key=lambda kv: -kv[1])[:22]:
        print(f'  {n:12s} {v:12.0f} {100 * v / (tot_inst or 1):5.1f}%')


if __name__ == '__main__':
    main()
