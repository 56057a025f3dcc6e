"""Summarise `ncu --page source --csv` output: stall totals and the hottest SASS instructions.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_stalls.py src.csv [kernel-index] [ntop]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 40
heads = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = heads[which]
end = heads[which + 1] - 1 if which + 1 < len(heads) else len(rows)
h = rows[hi]
print("kernel:", rows[hi - 1][1] if hi > 0 else "?")
ia, isrc, isamp, iex = h.index('Address'), h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
def num(x):
    try: return int(x)
    except Exception: return 0
data = [r for r in rows[hi + 1:end] if len(r) > isamp and r[ia].startswith('0x') or (len(r) > isamp and r[ia].isdigit())]
tot = sum(num(r[isamp]) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for r in data:
    for i, c in stall_cols:
        agg[c] = agg.get(c, 0) + num(r[i])
print("stall totals:", [(c, v, f"{100*v/max(tot,1):.1f}%") for c, v in sorted(agg.items(), key=lambda x: -x[1])[:9]])
pipes = {}
for r in data:
    op = r[isrc].split()[0] if r[isrc].split() else '?'
    if op.startswith('@'): op = r[isrc].split()[1]
    op = op.split('.')[0]
    pipes[op] = pipes.get(op, 0) + num(r[iex])
print("executed by opcode:", sorted(pipes.items(), key=lambda x: -x[1])[:16])
for r in sorted(data, key=lambda r: -num(r[isamp]))[:ntop]:
    st = sorted(((c, num(r[i])) for i, c in stall_cols if num(r[i]) > 0), key=lambda x: -x[1])[:3]
    print(r[ia][-5:], r[isamp].rjust(5), r[iex].rjust(8), r[isrc][:64].ljust(64), st)
