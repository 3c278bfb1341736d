#!/usr/bin/env python3
"""Instruction mix of the steady-state inner loops (one wavefront step) of the sweep kernels, read
from the SASS -- no GPU needed.  usage: tools/sass_steady.py [kalign_b200/libkalign_b200.so]"""
import re
import subprocess
import sys
from collections import Counter

obj = sys.argv[1] if len(sys.argv) > 1 else "kalign_b200/libkalign_b200.so"
out = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
for f in re.split(r"\n\s*Function : ", out)[1:]:
    name = f.split("\n", 1)[0]
    if "sweep_kernel" not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    idx = {a: i for i, (a, _) in enumerate(ins)}
    print(re.search(r"kb_sweep_kernelILi(\d)E", name).group(0))
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA\s+(?:\w+,\s*)?0x([0-9a-f]+)", t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in idx and 40 <= i - idx[tgt] + 1 < 640:
            body = ins[idx[tgt]:i + 1]
            c = Counter()
            for _, tt in body:
                op = tt.split()[1] if tt.startswith("@") else tt.split()[0]
                c[op.split(".")[0]] += 1
            if c["SHFL"] in (3, 6) and c["VOTE"] == 0:      # innermost steady loops only
                keys = ["FADD", "FADD2", "FMUL", "FMNMX", "FMNMX3", "FSEL", "MOV", "IMAD", "LDS", "LDG", "LDL", "STL", "ISETP", "BRA"]
                print("  @%05x %4d instr  " % (ins[idx[tgt]][0], len(body)) + " ".join("%s=%d" % (k, c[k]) for k in keys if c[k]))
