#!/usr/bin/env python3
"""Sum a KB200_TRACE=1 log of one bench step: sweep time per Hirschberg round (anchor batch and
tree levels), small-box kernel, per-level phases.  usage: tools/trace_sum.py <trace.err> [njobs_anchor]"""
import collections
import re
import sys

lines = open(sys.argv[1]).read().splitlines()
starts = [i for i, l in enumerate(lines) if re.search(r"jobs=\d+ round=0 ", l) and int(re.search(r"jobs=(\d+)", l)[1]) > 20000]
L = lines[starts[-1]:] if starts else lines
anchor = collections.defaultdict(float)
tree = collections.defaultdict(float)
small_a = small_t = 0.0
big = int(re.search(r"jobs=(\d+)", L[0])[1]) if starts else -1
for l in L:
    m = re.search(r"jobs=(\d+) round=(\d+) .*sweep_ms=([\d.]+)", l)
    if m:
        (anchor if int(m[1]) == big else tree)[int(m[2])] += float(m[3])
    m = re.search(r"jobs=(\d+) small.*small_ms=([\d.]+)", l)
    if m:
        if int(m[1]) == big:
            small_a += float(m[2])
        else:
            small_t += float(m[2])
ph = collections.defaultdict(float)
for l in L:
    m = re.search(r"prep ([\d.]+) bonus ([\d.]+) dp ([\d.]+) post ([\d.]+) weave ([\d.]+)", l)
    if m:
        for k, v in zip(["prep", "bonus", "dp", "post", "weave"], m.groups()):
            ph[k] += float(v)
print("anchor sweeps", {k: round(v, 1) for k, v in sorted(anchor.items())}, "small %.1f" % small_a, "sum %.1f" % (sum(anchor.values()) + small_a))
print("tree sweeps  ", {k: round(v, 1) for k, v in sorted(tree.items())}, "small %.1f" % small_t, "sum %.1f" % (sum(tree.values()) + small_t))
print("tree phases  ", {k: round(v, 1) for k, v in ph.items()})
