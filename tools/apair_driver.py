"""one kb200_aln_pairwise_dist call on a C3-shaped random alignment (for ncu): n rows x alnlen columns, 30 % gaps"""
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kalign_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = int(sys.argv[2]) if len(sys.argv) > 2 else 5136
rng = np.random.default_rng(1)
a = rng.choice(np.frombuffer(b"ACGU", dtype=np.uint8), size=(n, L))
a[rng.random((n, L)) < 0.3] = ord("-")
rows = [bytes(r).decode() for r in a]
ctx = _lib.Context(0)
dm = ctx.aln_pairwise_dist(rows)
s = ctx.stats()
print("n %d alnlen %d kernel %.3f ms, %.3e column-pairs/s, dm[0,1] %.6f" % (n, L, 1e3 * s["apair_seconds"], s["apair_col_pairs"] / s["apair_seconds"], dm[0, 1]))
ctx.close()
