"""Instruction mix of every loop (backward branch) in a cuobjdump -sass listing:
python tools/sass_loops.py file.sass [min_instructions]"""
import re, sys, collections
lines = open(sys.argv[1]).read().splitlines()
minlen = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ins = []  # (addr, text)
for l in lines:
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.U)?\b.*?\b0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_index:
            body = ins[addr_index[tgt]:i + 1]
            if len(body) < minlen:
                continue
            c = collections.Counter()
            for _, x in body:
                x = re.sub(r"^@!?U?P\d+\s+", "", x)
                op = x.split()[0].split(".")[0]
                c[op] += 1
            fp64 = sum(v for k, v in c.items() if k in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX"))
            print(f"loop {tgt:#x}..{a:#x}: {len(body)} instr, FP64-pipe {fp64}, " +
                  ", ".join(f"{k} {v}" for k, v in c.most_common(14)))
