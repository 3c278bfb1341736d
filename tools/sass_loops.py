#!/usr/bin/env python3
"""Static look at the loops of a kernel's SASS (no GPU needed): for every backward branch, the
instruction count of the loop body and its mix (FP, shuffles, shared/global/local accesses).
usage: tools/sass_loops.py kalign_b200/csrc/kb_dp.o kb_sweep_kernelILi0E [min_len]"""
import re
import subprocess
import sys
from collections import Counter


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    out = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)
    for f in funcs[1:]:
        name = f.split("\n", 1)[0]
        if pat not in name:
            continue
        ins = []
        for line in f.split("\n"):
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2)))
        addr_idx = {a: i for i, (a, _) in enumerate(ins)}
        print("==", name[:110], len(ins), "instructions")
        loops = []
        for i, (a, t) in enumerate(ins):
            m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\w+,\s*)?`?\(?\.?(?:L_x_\d+|0x([0-9a-f]+))", t)
            if m and m.group(1):
                tgt = int(m.group(1), 16)
                if tgt <= a and tgt in addr_idx:
                    loops.append((addr_idx[tgt], i))
        for (s, e) in sorted(loops, key=lambda x: x[0]):
            n = e - s + 1
            if n < min_len:
                continue
            c = Counter()
            for _, t in ins[s:e + 1]:
                op = t.split()[1] if t.startswith("@") else t.split()[0]
                op = op.split(".")[0]
                c[op] += 1
            keys = ["FADD", "FMUL", "FMNMX", "FMNMX3", "FSEL", "SHFL", "LDS", "STS", "LDG", "STG", "LD", "ST", "LDL", "STL", "BRA", "ISETP",
                    "LDGSTS", "WARPSYNC", "CALL", "VOTE", "BSSY", "BSYNC"]
            mix = " ".join("%s=%d" % (k, c[k]) for k in keys if c[k])
            print("  loop @%05x..%05x  %4d instr  %s" % (ins[s][0], ins[e][0], n, mix))


if __name__ == "__main__":
    main()
