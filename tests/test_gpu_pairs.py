"""GPU parity of the batched Hirschberg engine (kb200_pair_align_batch) against the oracle:
raw paths bit-identical, top-level meet-up score equal (tolerance 1e-5 relative, north star)."""
import numpy as np
import pytest

import kbind
from test_oracle_vs_ref import mutate, pfasum_like

pytestmark = pytest.mark.gpu


def oracle_profile(rng, A, L, subm, gpo, gpe, tgpe, depth):
    """profile of 2**depth related sequences built with the oracle's profile ops + alignments"""
    o = kbind.oracle()
    root = rng.integers(0, A, size=L).astype(np.uint8)

    def rec(d):
        if d == 0:
            s = mutate(rng, root, A)
            p = np.zeros((len(s) + 2) * 64, dtype=np.float32)
            o.ko_make_profile(s, len(s), subm, gpo, gpe, tgpe, 0.0, p)
            return ("leaf", s, p, len(s), 1)
        a, b = rec(d - 1), rec(d - 1)
        la, lb, na, nb = a[3], b[3], a[4], b[4]
        if a[0] == "leaf":
            if la < lb:
                path, _ = kbind.oracle_align(0, la, lb, subm, gpo, gpe, tgpe, seq1=a[1], seq2=b[1]); mirror = 0
            else:
                path, _ = kbind.oracle_align(0, lb, la, subm, gpo, gpe, tgpe, seq1=b[1], seq2=a[1]); mirror = 1
        else:
            pa, pb = a[2].copy(), b[2].copy()
            o.ko_set_gap_penalties(pa, la, nb)
            o.ko_set_gap_penalties(pb, lb, na)
            a = (a[0], a[1], pa, la, na); b = (b[0], b[1], pb, lb, nb)
            if la < lb:
                path, _ = kbind.oracle_align(2, la, lb, subm, gpo, gpe, tgpe, prof1=pa, prof2=pb); mirror = 0
            else:
                path, _ = kbind.oracle_align(2, lb, la, subm, gpo, gpe, tgpe, prof1=pb, prof2=pa); mirror = 1
        full = np.zeros(la + lb + 2, dtype=np.int32)
        full[:len(path)] = path
        o.ko_code_path(full, la, lb, mirror)
        newp = np.zeros((full[0] + 2) * 64, dtype=np.float32)
        o.ko_update(a[2], b[2], newp, full, na, nb, gpo, gpe, tgpe)
        return ("prof", None, newp, int(full[0]), na + nb)

    r = rec(depth)
    return r[2], r[3], r[4]


def check(ctx, prm, jobs, oracle_kwargs):
    paths, scores = ctx.pair_align_batch(prm, jobs)
    for i, (j, kw) in enumerate(zip(jobs, oracle_kwargs)):
        po, so = kbind.oracle_align(**kw)
        la = j["len_a"]
        assert np.array_equal(paths[i][1:la + 1], po[1:la + 1]), ("path", i, j["kind"], la, j["len_b"])
        ref = so["top_score"]
        assert abs(scores[i] - ref) <= 1e-5 * max(1.0, abs(ref)), ("score", i, scores[i], ref)


@pytest.fixture(scope="module")
def ctx():
    from kalign_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


@pytest.fixture(params=["auto", "thick", "thick_nosmall"])
def strips(request):
    """the engine picks thin (32-row) strips for small batches; KB200_THIN=0 forces the thick-strip
    code paths (K = 8 / 4 / 2 rows per lane) the big batches use, KB200_NO_SMALL keeps every box
    in the warp kernel down to the last recursion level"""
    import os
    old = {k: os.environ.get(k) for k in ("KB200_THIN", "KB200_NO_SMALL")}
    if request.param != "auto":
        os.environ["KB200_THIN"] = "0"
    if request.param == "thick_nosmall":
        os.environ["KB200_NO_SMALL"] = "1"
    yield request.param
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("name,A,gpo,gpe,tgpe", [("protein", 20, 7.0, 1.25, 1.0),
                                                  ("rna", 4, 217.0, 39.4, 292.6),
                                                  ("dna", 4, 8.0, 6.0, 0.0)])
def test_seqseq_batch(ctx, strips, name, A, gpo, gpe, tgpe):
    from kalign_b200 import _lib
    rng = np.random.default_rng(101)
    subm = pfasum_like(rng) if name == "protein" else np.ascontiguousarray(
        rng.integers(-4, 6, size=(23, 23)).astype(np.float32) * (50.0 if name == "rna" else 1.0))
    prm = _lib.params_from(subm, gpo, gpe, tgpe, nalpha=23 if name == "protein" else 5)
    jobs, kws = [], []
    sizes = [1, 2, 3, 5, 17, 31, 32, 33, 64, 65, 100, 127, 128, 129, 200, 257, 300, 450, 600, 1100]
    for t, la in enumerate(sizes * 2):
        s1 = rng.integers(0, A, size=la).astype(np.uint8)
        s2 = mutate(rng, s1, A) if t % 3 else rng.integers(0, A, size=int(rng.integers(la, 2 * la + 2))).astype(np.uint8)
        if len(s2) < len(s1):
            s1, s2 = s2, s1
        soff = float(np.float32(rng.random() * 2)) if (name == "protein" and t % 2) else 0.0
        jobs.append(dict(kind=0, len_a=len(s1), len_b=len(s2), seq_rows=s1, seq_cols=s2, soff=soff))
        kws.append(dict(kind=0, len_a=len(s1), len_b=len(s2), subm=subm, gpo=gpo, gpe=gpe, tgpe=tgpe,
                        soff=soff, seq1=s1, seq2=s2))
    check(ctx, prm, jobs, kws)


def test_seqseq_bonus(ctx, strips):
    from kalign_b200 import _lib
    rng = np.random.default_rng(7)
    subm = pfasum_like(rng)
    prm = _lib.params_from(subm, 7.0, 1.25, 1.0)
    jobs, kws = [], []
    for la in (2, 9, 40, 130, 260):
        s1 = rng.integers(0, 20, size=la).astype(np.uint8)
        s2 = mutate(rng, s1, 20)
        if len(s2) < len(s1):
            s1, s2 = s2, s1
        la, lb = len(s1), len(s2)
        bonus = np.zeros(la * lb, dtype=np.float32)
        idx = rng.integers(0, la * lb, size=2 * la)
        bonus[idx] = (rng.random(2 * la) * 2).astype(np.float32)
        jobs.append(dict(kind=0, len_a=la, len_b=lb, seq_rows=s1, seq_cols=s2, bonus=bonus))
        kws.append(dict(kind=0, len_a=la, len_b=lb, subm=subm, gpo=7.0, gpe=1.25, tgpe=1.0, seq1=s1, seq2=s2, bonus=bonus))
    check(ctx, prm, jobs, kws)


@pytest.mark.parametrize("name,A,gpo,gpe,tgpe", [("protein", 20, 7.0, 1.25, 1.0),
                                                  ("rna", 4, 217.0, 39.4, 292.6)])
def test_profile_batch(ctx, strips, name, A, gpo, gpe, tgpe):
    from kalign_b200 import _lib
    rng = np.random.default_rng(55)
    subm = pfasum_like(rng) if name == "protein" else np.ascontiguousarray(
        rng.integers(-4, 6, size=(23, 23)).astype(np.float32) * 30.0)
    prm = _lib.params_from(subm, gpo, gpe, tgpe, nalpha=23 if name == "protein" else 5)
    o = kbind.oracle()
    jobs, kws = [], []
    for t, L in enumerate([4, 20, 40, 70, 140, 180, 300, 560]):
        p1, l1, n1 = oracle_profile(rng, A, L, subm, gpo, gpe, tgpe, depth=1 + t % 3)
        p2, l2, n2 = oracle_profile(rng, A, L, subm, gpo, gpe, tgpe, depth=1 + (t + 1) % 3)
        if l1 >= l2:
            p1, l1, n1, p2, l2, n2 = p2, l2, n2, p1, l1, n1
        a, b = p1.copy(), p2.copy()
        o.ko_set_gap_penalties(a, l1, n2)
        o.ko_set_gap_penalties(b, l2, n1)
        jobs.append(dict(kind=2, len_a=l1, len_b=l2, prof_rows=a, prof_cols=b))
        kws.append(dict(kind=2, len_a=l1, len_b=l2, subm=subm, gpo=gpo, gpe=gpe, tgpe=tgpe, prof1=a, prof2=b))
        s = mutate(rng, rng.integers(0, A, size=L).astype(np.uint8), A)
        a1 = p1.copy()
        o.ko_set_gap_penalties(a1, l1, 1)
        jobs.append(dict(kind=1, len_a=l1, len_b=len(s), prof_rows=a1, seq_cols=s, sip=n1))
        kws.append(dict(kind=1, len_a=l1, len_b=len(s), subm=subm, gpo=gpo, gpe=gpe, tgpe=tgpe, prof1=a1, seq2=s, sip=n1))
    check(ctx, prm, jobs, kws)
