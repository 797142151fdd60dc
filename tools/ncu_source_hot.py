#!/usr/bin/env python
"""Hot spots of an `ncu --page source --csv` dump (SASS view): samples and executed instructions
per opcode, stall-reason totals, and the top-N instructions by stall samples.
usage: tools/ncu_source_hot.py source.csv [N]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot_s = sum(int(r[ix['# Samples']]) for r in body)
tot_i = sum(int(r[ix['Instructions Executed']]) for r in body)
print(f"{rows[0][1][:80]}\ninstructions {len(body)} static, {tot_i} executed (warp-level), samples {tot_s}")
op_s, op_i = collections.Counter(), collections.Counter()
for r in body:
    toks = r[ix['Source']].split()
    op = (toks[1] if toks[0].startswith('@') else toks[0]).rstrip(';')
    op = '.'.join(op.split('.')[:2]) if op.startswith(('LDG', 'STG', 'RED', 'ATOM', 'MUFU', 'SHFL', 'LDL', 'STL')) else op.split('.')[0]
    op_s[op] += int(r[ix['# Samples']]); op_i[op] += int(r[ix['Instructions Executed']])
print("opcode            exec%   samples%")
for op, c in op_i.most_common(28):
    print(f"  {op:14s} {100*c/tot_i:6.2f}  {100*op_s[op]/tot_s:6.2f}")
st = [h for h in hdr if h.startswith('stall_')]
tots = {h: sum(int(r[ix[h]]) for r in body) for h in st}
print("stalls:", {k[6:]: round(100 * v / tot_s, 1) for k, v in sorted(tots.items(), key=lambda kv: -kv[1]) if v})
print(f"top {N} instructions by samples:")
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix['# Samples']]))[:N]
for i in sorted(order):
    r = body[i]
    top = max(st, key=lambda h: int(r[ix[h]]))
    print(f"  #{i:4d} {100*int(r[ix['# Samples']])/tot_s:5.2f}%  {r[ix['Source']].strip()[:70]:70s} {top[6:]}")
