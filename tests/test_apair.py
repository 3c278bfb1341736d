"""compute_aln_pairwise_dist (lib/src/aln_apair_dist.c:9), the N x N identity distances of the realign loop.

tests/golden/apair.npz holds matrices written by the UNMODIFIED reference (tools/gen_golden_apair.py).
CPU: the numpy restatement in kbind reproduces them (and the reference itself where oracle/_ref exists).
GPU: kb200_aln_pairwise_dist through the C ABI reproduces them bit for bit, and equals the restatement on
sizes that span many tiles."""
import os

import numpy as np
import pytest

import kbind

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "apair.npz")


def cases():
    z = np.load(G)
    for k in z["names"]:
        k = str(k)
        yield k, [str(s) for s in z["rows_" + k]], z["dm_" + k]


def test_restatement_reproduces_golden():
    n = 0
    for k, rows, dm in cases():
        got = kbind.oracle_aln_pairwise_dist(rows)
        assert got.dtype == np.float32 and np.array_equal(got, dm), k
        assert np.array_equal(dm, dm.T) and not dm.diagonal().any()
        n += 1
    assert n == 12


@pytest.mark.skipif(not kbind.have_ref(), reason="oracle/_ref not built")
def test_reference_reproduces_golden():
    for k, rows, dm in cases():
        assert np.array_equal(kbind.ref_aln_pairwise_dist(rows), dm), k


@pytest.fixture(scope="module")
def ctx():
    from kalign_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
def test_gpu_reproduces_golden(ctx):
    for k, rows, dm in cases():
        got = ctx.aln_pairwise_dist(rows)
        assert np.array_equal(got, dm), k


@pytest.mark.gpu
@pytest.mark.parametrize("n,L,gap", [(300, 2001, 0.25), (1000, 700, 0.6), (130, 5000, 0.02)])
def test_gpu_equals_restatement_many_tiles(ctx, n, L, gap):
    rng = np.random.default_rng(n * 7 + L)
    a = rng.choice(np.frombuffer(b"ACGUacgu", dtype=np.uint8), size=(n, L))
    a[rng.random((n, L)) < gap] = ord("-")
    a[n // 2] = ord("-")                      # a row without residues
    rows = [bytes(r).decode() for r in a]
    got = ctx.aln_pairwise_dist(rows)
    want = kbind.oracle_aln_pairwise_dist(rows)
    assert np.array_equal(got, want)
    assert np.array_equal(got, got.T)


@pytest.mark.gpu
def test_gpu_alignment_of_the_product_round_trip(ctx):
    """distances of an alignment the product itself produced: identical sequences are at distance 0,
    every entry lies in [0, 1]"""
    from kalign_b200 import synth
    seqs = synth.family(50, 200, synth.RNA, seed=9)
    seqs[11] = seqs[10]
    rows = ctx.kalign(seqs, n_threads=2, type_=2, consistency=0)
    dm = ctx.aln_pairwise_dist(rows)
    assert dm[10, 11] == 0.0 and dm.min() >= 0.0 and dm.max() <= 1.0
    assert np.array_equal(dm, kbind.oracle_aln_pairwise_dist(rows))
