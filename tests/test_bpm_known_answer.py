"""Known-answer test of the Myers distance, after the reference's own bpm_utest (lib/src/bpm_test.c:308-347): mutate a
63-symbol sequence position by position and require the dynamic-programming distance, the one-word bit-parallel
routine and the blocked routine to agree.  Here: a numpy restatement of dyn_256's recurrence (lib/src/bpm.c:28-90 --
free start in the text, free text extension after the pattern's last symbol, i.e. the distance of the pattern to
its best-matching substring) against the oracle's ko_bpm_block (oracle/kalign_oracle.c), and -- where oracle/_ref
exists -- against the reference's dyn_256, bpm and bpm_block themselves.  The GPU kernel is pinned to the same
values through tests/golden/bpm.npz and test_gpu_pipeline.py (distance matrices equal to d_estimation's)."""
import ctypes as C

import numpy as np
import pytest

import kbind

pytestmark = pytest.mark.skipif(not kbind.have_oracle(), reason="oracle lib not built")


def dyn(t, p):
    """min over end positions in t of the edit distance between p and a substring of t ending there"""
    m = len(p)
    prev = np.arange(m + 1, dtype=np.int64)
    for ch in t:
        cur = np.empty_like(prev)
        cur[0] = 0
        for j in range(1, m + 1):
            c = 0 if ch == p[j - 1] else 1
            cur[j] = min(prev[j - 1] + c, prev[j] + (1 if j < m else 0), cur[j - 1] + 1)
        prev = cur
    return int(prev[m])


def cases():
    rng = np.random.default_rng(63)
    out = []
    for A in (4, 13):
        a = rng.integers(0, A, size=63).astype(np.uint8)
        for k in range(0, 63, 3):                      # k mutated positions, as bpm_utest's outer loop
            for _ in range(4):
                b = a.copy()
                pos = rng.choice(63, size=k, replace=False)
                b[pos] = rng.integers(0, A, size=k)
                out.append((a, b))
        # unequal lengths and patterns longer than one word (blocked routine only)
        for n, m in ((200, 63), (130, 64), (300, 129), (90, 90)):
            t = rng.integers(0, A, size=n).astype(np.uint8)
            p = t[int(rng.integers(0, n - m + 1)):][:m].copy()
            p[rng.choice(m, size=m // 10, replace=False)] = rng.integers(0, A, size=m // 10)
            out.append((t, p))
    return out


def test_oracle_agrees_with_dynamic_programming():
    o = kbind.oracle()
    n = 0
    for t, p in cases():
        assert o.ko_bpm_block(t, p, len(t), len(p)) == dyn(t, p), (len(t), len(p))
        n += 1
    assert n > 150


@pytest.mark.skipif(not kbind.have_ref(), reason="oracle/_ref not built")
def test_reference_four_way_agreement():
    ref = kbind.ref()
    u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
    for f in ("dyn_256", "bpm"):
        getattr(ref, f).argtypes = [u8p, u8p, C.c_int, C.c_int]
        getattr(ref, f).restype = C.c_uint8
    o = kbind.oracle()
    for t, p in cases():
        want = dyn(t, p)
        assert ref.bpm_block(t, p, len(t), len(p)) == want
        assert o.ko_bpm_block(t, p, len(t), len(p)) == want
        if len(p) <= 255:
            assert ref.dyn_256(t, p, len(t), len(p)) == want
        if len(p) <= 63:
            assert ref.bpm(t, p, len(t), len(p)) == want
