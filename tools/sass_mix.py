#!/usr/bin/env python3
"""Opcode mix of one kernel in a `cuobjdump -sass` dump: whole function and its largest loop (the particle sweep).
  cuobjdump -sass lib.so > all.sass; python tools/sass_mix.py all.sass <substring of the mangled name>"""
import re, sys, collections
path, key = sys.argv[1], sys.argv[2]
lines = open(path).read().splitlines()
start = next(i for i, l in enumerate(lines) if "Function :" in l and key in l)
end = next((i for i in range(start + 1, len(lines)) if "Function :" in lines[i]), len(lines))
ins = []  # (addr, opcode, text)
for l in lines[start:end]:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        txt = m.group(2).strip()
        t = txt.split()
        op = t[1] if t[0].startswith("@") else t[0]
        ins.append((int(m.group(1), 16), op, txt))
print(lines[start].strip()[:120]); print("instructions:", len(ins))
addr_idx = {a: i for i, (a, _, _) in enumerate(ins)}
loops = []
for i, (a, op, txt) in enumerate(ins):
    if op.startswith("BRA"):
        m = re.search(r"0x([0-9a-f]+)", txt)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr_idx:
                loops.append((i - addr_idx[tgt], addr_idx[tgt], i))
loops.sort(reverse=True)
def cls(op):
    b = op.split(".")[0]
    if b in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"): return "fp64:" + b
    if b == "MUFU": return "MUFU" + ("." + op.split(".")[1] if "." in op else "")
    if b in ("LDS", "STS", "LDG", "STG", "LDGSTS", "LD", "ST", "LDL", "STL", "ATOMG", "ATOMS", "RED", "LDC", "LDCU"): return "mem:" + b
    if b in ("BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "BAR", "WARPSYNC", "NOP", "BREAK", "BMOV"): return "ctl:" + b
    return "int/other:" + b
def mix(sub, title):
    c = collections.Counter(cls(op) for _, op, _ in sub)
    tot = len(sub)
    fp64 = sum(v for k, v in c.items() if k.startswith("fp64"))
    print(f"--- {title}: {tot} instructions, fp64 {fp64} ({100*fp64/tot:.1f} %)")
    for k, v in c.most_common(28): print(f"   {k:24s} {v:6d} {100*v/tot:5.1f} %")
mix(ins, "whole function")
for size, a, b in loops[:3]:
    mix(ins[a:b + 1], f"loop of {size} instructions at 0x{ins[a][0]:x}")
