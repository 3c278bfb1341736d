"""kalign_run_seeded's guide-tree noise and the ensemble's independent runs.

tests/golden/seeded.npz was written by the UNMODIFIED reference (tools/gen_golden_seeded.py): noise factors of
build_tree_kmeans_noisy (lib/src/bisectingKmeans.c:104-116 with the generator of lib/src/tlrng.c), the per-run
parameters of resolve_run_params (lib/src/ensemble.c:55-76) and alignments of kalign_run_seeded with noisy trees.
CPU: the product's host-side restatements (kb200_tree_noise, kb200_ensemble_run_params) reproduce them bit for bit.
GPU: kb200_kalign_seeded / kb200_ensemble_run reproduce the alignments through the C ABI."""
import hashlib
import os
import sys

import numpy as np
import pytest

import kbind
from kalign_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_golden_seeded as G  # noqa: E402  (the case tables; nothing is generated at import)

Z = np.load(os.path.join(ROOT, "tests", "golden", "seeded.npz"))


def test_noise_factors_reproduce_golden():
    for i, (seed, sigma) in enumerate(G.NOISE):
        f = _lib.tree_noise(seed, sigma, 100000)
        assert np.array_equal(f[:64], Z["noise%d_head" % i])
        assert hashlib.sha256(f.tobytes()).hexdigest() == str(Z["noise%d_sha" % i])
        assert f.min() >= np.float32(0.1)
    with pytest.raises(RuntimeError):
        _lib.tree_noise(0, 0.2, 4)                # seed 0 means "no noise" in the reference, never a stream


def test_run_params_reproduce_golden():
    want = Z["run_params"]
    i = 0
    for base in G.BASES:
        for k in range(26):
            got = _lib.ensemble_run_params(*base, k, 42)
            assert np.array_equal(np.array(got, dtype=np.float64), want[i]), (base, k)
            i += 1
    # run 0 is the deterministic default run; later runs cycle through entries 1..11, 0 of the table
    assert _lib.ensemble_run_params(5.5, 2.0, 1.0, 0, 42) == (5.5, 2.0, 1.0, 0, 0.0)
    assert _lib.ensemble_run_params(5.5, 2.0, 1.0, 12, 42)[:3] == (5.5, 2.0, 1.0)


@pytest.mark.skipif(not kbind.have_ref(), reason="oracle/_ref not built")
def test_reference_agrees_live():
    for seed, sigma in G.NOISE:
        assert np.array_equal(_lib.tree_noise(seed, sigma, 5000), kbind.ref_tree_noise(seed, sigma, 5000))
    for k in range(40):
        assert _lib.ensemble_run_params(55.0, 8.5, 4.25, k, 7) == kbind.ref_resolve_run_params(55.0, 8.5, 4.25, k, 7)


def test_runs_shard_without_overlap():
    """rank r of `world` takes the runs k % world == r: every run exactly once"""
    for n_runs in (1, 5, 8):
        for world in (1, 2, 3, 8):
            seen = sorted(k for r in range(world) for k in range(r, n_runs, world))
            assert seen == list(range(n_runs))


@pytest.fixture(scope="module")
def ctx():
    c = _lib.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(G.SEEDED))
def test_kalign_seeded_equals_reference(ctx, name):
    fk, kw = G.SEEDED[name]
    seqs, type_ = G.families()[fk]
    rows = ctx.kalign_seeded(seqs, n_threads=2, type_=type_, **kw)
    assert rows == [str(x) for x in Z["rows_" + name]]


@pytest.mark.gpu
@pytest.mark.parametrize("fk", sorted(G.ENSEMBLE))
def test_ensemble_runs_equal_reference(ctx, fk):
    n_runs, seed = G.ENSEMBLE[fk]
    seqs, type_ = G.families()[fk]
    # two "ranks" of a two-GPU job, computed one after the other on this GPU: together all runs, each identical to
    # the reference's kalign_run_seeded with the parameters the reference's resolve_run_params gives that run
    got = {}
    for rank in range(2):
        part = ctx.ensemble_runs(seqs, n_runs, seed=seed, rank=rank, world=2, n_threads=2, type_=type_)
        assert sorted(part) == list(range(rank, n_runs, 2))
        got.update(part)
    for k in range(n_runs):
        assert got[k] == [str(x) for x in Z["ens_%s_%d" % (fk, k)]], k
    assert any(got[k] != got[0] for k in range(1, n_runs))


def test_creation_order_equals_reference_golden():
    """create_tasks (lib/src/bisectingKmeans.c:1084) fills the list in pre-order; tests/golden/full_T3.npz holds the
    reference's list of the C3 tree (N = 10 000) before any sort: sorting it by c and restoring the creation order
    with kb200_tasks_creation_order must give it back"""
    g = np.load(os.path.join(ROOT, "tests", "golden", "full_T3.npz"))
    t, n = g["tasks"], int(g["n"])
    srt = t[np.argsort(t[:, 2], kind="stable")]
    assert not np.array_equal(srt, t)
    assert np.array_equal(_lib.tasks_creation_order(srt, n), t)
    assert t[0, 2] == 2 * n - 2                       # the root comes first
    bad = srt.copy()
    bad[5, 2] += 1
    with pytest.raises(RuntimeError):
        _lib.tasks_creation_order(bad, n)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["rna_default", "dna_default"])
def test_guide_tree_entry_point(ctx, tag):
    """kb200_guide_tree on the tree-alphabet codes of a committed reference run: the reference's task list and
    msa->seq_distances (for nucleotides the tree alphabet is the alignment alphabet the fixture stores)"""
    z = np.load(os.path.join(ROOT, "tests", "golden", "msa_%s.npz" % tag))
    n = len(z["seqs"])
    flat, offs, lens = _lib.pack([np.ascontiguousarray(z["codes%d" % i]) for i in range(n)])
    abc, sd = ctx.guide_tree(flat, offs, lens)
    assert np.array_equal(sd, z["seq_distances"])
    want = z["tasks"]
    want = want[np.argsort(want[:, 2], kind="stable")]
    assert np.array_equal(abc[np.argsort(abc[:, 2], kind="stable")], want)
    assert np.array_equal(abc, _lib.tasks_creation_order(want, n))
    # a noisy tree of the same sequences is another tree, deterministically
    a1, _ = ctx.guide_tree(flat, offs, lens, tree_seed=11, tree_noise=0.4)
    a2, _ = ctx.guide_tree(flat, offs, lens, tree_seed=11, tree_noise=0.4)
    assert np.array_equal(a1, a2)
