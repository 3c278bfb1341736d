"""Pin the CPU restatement (oracle/kalign_oracle.c) against the REAL reference (oracle/_ref):
raw Hirschberg paths bit-identical, scores equal, bpm_block equal, profile ops equal.
CPU only; skipped when oracle/_ref has not been built."""
import numpy as np
import pytest

import kbind
from kalign_b200 import synth

pytestmark = pytest.mark.skipif(not (kbind.have_ref() and kbind.have_oracle()),
                                reason="oracle/_ref or oracle lib not built")


def rand_subm(rng, sym=True, scale=5.0, integer=True):
    m = rng.normal(0, scale, size=(23, 23))
    if integer:
        m = np.round(m)
    if sym:
        m = (m + m.T) / 2 if not integer else np.round((m + m.T) / 2)
    return np.ascontiguousarray(m, dtype=np.float32)


def pfasum_like(rng):
    m = rand_subm(rng)
    m[np.arange(23), np.arange(23)] = np.abs(m[np.arange(23), np.arange(23)]) + 4
    return m


def mutate(rng, s, A, sub=0.2, indel=0.05):
    out = []
    for c in s:
        r = rng.random()
        if r < indel:
            continue
        if r < 2 * indel:
            out.append(rng.integers(0, A))
        out.append(rng.integers(0, A) if rng.random() < sub else c)
    if not out:
        out = [0]
    return np.array(out, dtype=np.uint8)


def make_profile(seq, subm, gpo, gpe, tgpe, soff=0.0):
    L = len(seq)
    p = np.zeros((L + 2) * 64, dtype=np.float32)
    kbind.oracle().ko_make_profile(seq, L, subm, gpo, gpe, tgpe, soff, p)
    return p


def merged_profile(rng, A, L, subm, gpo, gpe, tgpe, depth=2):
    """a profile of 2**depth related sequences built with the REFERENCE's own ops"""
    lib = kbind.refh()
    root = rng.integers(0, A, size=L).astype(np.uint8)

    def rec(d):
        if d == 0:
            s = mutate(rng, root, A)
            p = np.zeros((len(s) + 2) * 64, dtype=np.float32)
            lib.refh_make_profile(s, len(s), subm, gpo, gpe, tgpe, 0.0, p)
            return ("leaf", s, p, len(s), 1)
        a = rec(d - 1)
        b = rec(d - 1)
        la, lb = a[3], b[3]
        na, nb = a[4], b[4]
        if a[0] == "leaf" and b[0] == "leaf":
            # orientation as do_align: shorter on rows, tie -> b rows (aln_run.c:300)
            if la < lb:
                path, _ = kbind.ref_align(0, la, lb, subm, gpo, gpe, tgpe, seq1=a[1], seq2=b[1])
                mirror = 0
            else:
                path, _ = kbind.ref_align(0, lb, la, subm, gpo, gpe, tgpe, seq1=b[1], seq2=a[1])
                mirror = 1
        else:
            pa, pb = a[2].copy(), b[2].copy()
            lib.refh_set_gap_penalties(pa, la, nb)
            lib.refh_set_gap_penalties(pb, lb, na)
            a = (a[0], a[1], pa, la, na)
            b = (b[0], b[1], pb, lb, nb)
            if la < lb:
                path, _ = kbind.ref_align(2, la, lb, subm, gpo, gpe, tgpe, prof1=pa, prof2=pb)
                mirror = 0
            else:
                path, _ = kbind.ref_align(2, lb, la, subm, gpo, gpe, tgpe, prof1=pb, prof2=pa)
                mirror = 1
        full = np.zeros(la + lb + 2, dtype=np.int32)
        full[:len(path)] = path
        lib.refh_code_path(full, la, lb, mirror)
        newp = np.zeros((full[0] + 2) * 64, dtype=np.float32)
        lib.refh_update(a[2], b[2], newp, full, na, nb, gpo, gpe, tgpe)
        return ("prof", None, newp, int(full[0]), na + nb)

    r = rec(depth)
    return r[2], r[3], r[4]


PARAMS = [
    ("protein", 20, 7.0, 1.25, 1.0),
    ("rna", 4, 217.0, 39.4, 292.6),
    ("dna", 4, 8.0, 6.0, 0.0),
]


@pytest.mark.parametrize("name,A,gpo,gpe,tgpe", PARAMS)
def test_seqseq_paths_identical(name, A, gpo, gpe, tgpe):
    rng = np.random.default_rng(11)
    for trial in range(40):
        subm = pfasum_like(rng) if name == "protein" else np.ascontiguousarray(
            rng.integers(-4, 6, size=(23, 23)).astype(np.float32) * (50.0 if name == "rna" else 1.0))
        la = int(rng.integers(1, 90))
        s1 = rng.integers(0, A, size=la).astype(np.uint8)
        s2 = mutate(rng, s1, A) if trial % 3 else rng.integers(0, A, size=int(rng.integers(la, 2 * la + 2))).astype(np.uint8)
        if len(s2) < la:
            s1, s2 = s2, s1
        la, lb = len(s1), len(s2)
        soff = float(rng.random()) * 2 if name == "protein" and trial % 2 else 0.0
        pr, sr = kbind.ref_align(0, la, lb, subm, gpo, gpe, tgpe, soff=soff, seq1=s1, seq2=s2)
        po, so = kbind.oracle_align(0, la, lb, subm, gpo, gpe, tgpe, soff=soff, seq1=s1, seq2=s2)
        assert np.array_equal(pr[1:la + 1], po[1:la + 1]), (trial, la, lb)
        assert sr["margin_count"] == so["margin_count"]
        assert sr["margin_sum"] == so["margin_sum"]
        _, sc = kbind.ref_align(0, la, lb, subm, gpo, gpe, tgpe, soff=soff, seq1=s1, seq2=s2, score_only=True)
        assert sc["score"] == so["top_score"]


def test_seqseq_bonus_identical():
    rng = np.random.default_rng(5)
    subm = pfasum_like(rng)
    for trial in range(15):
        la = int(rng.integers(2, 60))
        s1 = rng.integers(0, 20, size=la).astype(np.uint8)
        s2 = mutate(rng, s1, 20)
        if len(s2) < la:
            s1, s2 = s2, s1
        la, lb = len(s1), len(s2)
        bonus = np.zeros(la * lb, dtype=np.float32)
        idx = rng.integers(0, la * lb, size=la)
        bonus[idx] = rng.random(la).astype(np.float32) * 2
        pr, sr = kbind.ref_align(0, la, lb, subm, 7.0, 1.25, 1.0, seq1=s1, seq2=s2, bonus=bonus)
        po, so = kbind.oracle_align(0, la, lb, subm, 7.0, 1.25, 1.0, seq1=s1, seq2=s2, bonus=bonus)
        assert np.array_equal(pr[1:la + 1], po[1:la + 1])
        assert sr["margin_sum"] == so["margin_sum"]


@pytest.mark.parametrize("name,A,gpo,gpe,tgpe", PARAMS)
def test_profile_kernels_identical(name, A, gpo, gpe, tgpe):
    rng = np.random.default_rng(23)
    subm = pfasum_like(rng) if name == "protein" else np.ascontiguousarray(
        rng.integers(-4, 6, size=(23, 23)).astype(np.float32) * (30.0 if name == "rna" else 1.0))
    lib = kbind.refh()
    for trial in range(8):
        L = int(rng.integers(5, 70))
        p1, l1, n1 = merged_profile(rng, A, L, subm, gpo, gpe, tgpe, depth=1 + trial % 2)
        p2, l2, n2 = merged_profile(rng, A, L, subm, gpo, gpe, tgpe, depth=1 + (trial + 1) % 2)
        # profile-profile, shorter on rows (aln_run.c:361-386)
        if l1 >= l2:
            p1, l1, n1, p2, l2, n2 = p2, l2, n2, p1, l1, n1
        a = p1.copy(); b = p2.copy()
        lib.refh_set_gap_penalties(a, l1, n2)
        lib.refh_set_gap_penalties(b, l2, n1)
        a2 = p1.copy(); b2 = p2.copy()
        kbind.oracle().ko_set_gap_penalties(a2, l1, n2)
        kbind.oracle().ko_set_gap_penalties(b2, l2, n1)
        assert np.array_equal(a, a2) and np.array_equal(b, b2)
        pr, sr = kbind.ref_align(2, l1, l2, subm, gpo, gpe, tgpe, prof1=a, prof2=b)
        po, so = kbind.oracle_align(2, l1, l2, subm, gpo, gpe, tgpe, prof1=a, prof2=b)
        assert np.array_equal(pr[1:l1 + 1], po[1:l1 + 1])
        assert sr["margin_sum"] == so["margin_sum"] and sr["margin_count"] == so["margin_count"]
        # profile (rows) - sequence (cols)
        s = mutate(rng, rng.integers(0, A, size=L).astype(np.uint8), A)
        pr, sr = kbind.ref_align(1, l1, len(s), subm, gpo, gpe, tgpe, prof1=a, seq2=s, sip=n1)
        po, so = kbind.oracle_align(1, l1, len(s), subm, gpo, gpe, tgpe, prof1=a, seq2=s, sip=n1)
        assert np.array_equal(pr[1:l1 + 1], po[1:l1 + 1])
        assert sr["margin_sum"] == so["margin_sum"]
        # coded path + merge
        for mirror in (0, 1):
            la, lb = (l1, l2) if not mirror else (l2, l1)
            rows = l1
            praw, _ = kbind.ref_align(2, l1, l2, subm, gpo, gpe, tgpe, prof1=a, prof2=b)
            full_r = np.zeros(l1 + l2 + 2, dtype=np.int32); full_r[:rows + 2] = praw
            full_o = full_r.copy()
            lib.refh_code_path(full_r, la, lb, mirror)
            kbind.oracle().ko_code_path(full_o, la, lb, mirror)
            assert np.array_equal(full_r[:full_r[0] + 2], full_o[:full_o[0] + 2])
            pa_, pb_, na_, nb_ = (a, b, n1, n2) if not mirror else (b, a, n2, n1)
            newr = np.zeros((full_r[0] + 2) * 64, dtype=np.float32)
            newo = np.zeros_like(newr)
            lib.refh_update(pa_, pb_, newr, full_r, na_, nb_, gpo, gpe, tgpe)
            kbind.oracle().ko_update(pa_, pb_, newo, full_o, na_, nb_, gpo, gpe, tgpe)
            assert np.array_equal(newr, newo)


def test_make_profile_identical():
    rng = np.random.default_rng(3)
    subm = pfasum_like(rng)
    for L in (1, 2, 17, 64):
        s = rng.integers(0, 23, size=L).astype(np.uint8)
        a = np.zeros((L + 2) * 64, dtype=np.float32)
        b = np.ones((L + 2) * 64, dtype=np.float32)
        kbind.refh().refh_make_profile(s, L, subm, 7.0, 1.25, 1.0, 0.37, a)
        kbind.oracle().ko_make_profile(s, L, subm, 7.0, 1.25, 1.0, 0.37, b)
        assert np.array_equal(a, b)


def test_bpm_block_matches_reference():
    rng = np.random.default_rng(9)
    ref = kbind.ref()
    o = kbind.oracle()
    cases = []
    for A in (4, 13):
        for _ in range(150):
            n = int(rng.integers(1, 400))
            t = rng.integers(0, A, size=n).astype(np.uint8)
            if rng.random() < 0.6:
                p = mutate(rng, t, A, sub=rng.random() * 0.5, indel=rng.random() * 0.1)
            else:
                p = rng.integers(0, A, size=int(rng.integers(1, n + 1))).astype(np.uint8)
            if len(p) > len(t):
                t, p = p, t
            cases.append((t, p))
    # long patterns: > 1024 is truncated (bpm.c:369-371), multiples of 64, exact 1024
    for m in (63, 64, 65, 128, 1023, 1024, 1025, 1500):
        t = rng.integers(0, 4, size=m + int(rng.integers(0, 300))).astype(np.uint8)
        p = mutate(rng, t, 4, sub=0.2, indel=0.02)[:m]
        if len(p) > len(t):
            t, p = p, t
        cases.append((t, p))
    for t, p in cases:
        a = ref.bpm_block(t, p, len(t), len(p))
        b = o.ko_bpm_block(t, p, len(t), len(p))
        assert a == b, (len(t), len(p), a, b)


def test_posmap_and_pipeline_small():
    """whole reference pipeline on a small family: oracle re-derives every anchor position map
    and every task path from the reference's own intermediate state."""
    seqs = synth.family(24, 60, synth.PROTEIN, seed=7)
    run = kbind.RefRun(seqs, n_threads=1, consistency=5, weight=2.0)
    subm, gp = run.params()
    subm = np.ascontiguousarray(subm)
    anchors = run.anchor_ids()
    assert len(anchors) == 5
    for i in range(0, run.n, 5):
        si = run.codes(i)
        for k, ak in enumerate(anchors):
            want = run.posmap(i, k)
            if i == ak:
                assert np.array_equal(want, np.arange(len(si)))
                continue
            sj = run.codes(int(ak))
            li, lj = len(si), len(sj)
            # pairwise_align_map orientation: i on rows when len_i <= len_j (anchor_consistency.c:47)
            if li <= lj:
                p, _ = kbind.oracle_align(0, li, lj, subm, gp[0], gp[1], gp[2], seq1=si, seq2=sj)
                mirror = 0
            else:
                p, _ = kbind.oracle_align(0, lj, li, subm, gp[0], gp[1], gp[2], seq1=sj, seq2=si)
                mirror = 1
            full = np.zeros(li + lj + 2, dtype=np.int32)
            full[:len(p)] = p
            kbind.oracle().ko_code_path(full, li, lj, mirror)
            got = np.zeros(li, dtype=np.int32)
            kbind.oracle().ko_posmap_from_path(full, li, got)
            assert np.array_equal(want, got), (i, k)
    run.close()
