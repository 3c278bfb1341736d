"""GPU parity of the three seam entry points and of the whole kalign() pipeline against the REAL
reference (oracle/_ref travels to the GPU box as built files): distance matrix exact, anchor
position maps exact, per-sequence gaps exact, final MSA strings byte-identical."""
import numpy as np
import pytest

import kbind
from kalign_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not kbind.have_ref(), reason="oracle/_ref missing")]


@pytest.fixture(scope="module")
def ctx():
    from kalign_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


@pytest.fixture(params=["auto", "thick"])
def strips(request):
    """KB200_THIN=0: the thick-strip code paths of the big batches, on the small test families"""
    import os
    old = os.environ.get("KB200_THIN")
    if request.param == "thick":
        os.environ["KB200_THIN"] = "0"
    yield request.param
    if old is None:
        os.environ.pop("KB200_THIN", None)
    else:
        os.environ["KB200_THIN"] = old


def ref_state(seqs, consistency, type_=8):
    run = kbind.RefRun(seqs, n_threads=2, type_=type_, consistency=consistency, weight=2.0)
    return run


FAMILIES = [
    ("protein_small", lambda: synth.family(24, 60, synth.PROTEIN, seed=7), 8),
    ("protein_mid", lambda: synth.family(120, 150, synth.PROTEIN, seed=8), 8),
    ("rna", lambda: synth.family(70, 200, synth.RNA, seed=9), 2),
    ("dna_type", lambda: synth.family(40, 120, synth.DNA, seed=10, sub=0.05, ins=0.01, dele=0.01), 0),
]


@pytest.mark.parametrize("name,gen,type_", FAMILIES)
@pytest.mark.parametrize("consistency", [0, 5])
def test_seams_and_msa(ctx, strips, name, gen, type_, consistency):
    from kalign_b200 import _lib
    seqs = gen()
    run = ref_state(seqs, consistency, type_)
    try:
        subm, gp = run.params()
        biotype = run.biotype()
        prm = _lib.make_params(biotype, type_)
        assert np.array_equal(np.array(prm.subm[:]).reshape(23, 23), subm)
        assert (prm.gpo, prm.gpe, prm.tgpe) == (gp[0], gp[1], gp[2])
        codes = [run.codes(i) for i in range(run.n)]      # alignment alphabet, sorted order
        flat, offs, lens = _lib.pack(codes)
        tasks = run.tasks()
        sd = run.seq_distances()
        posmaps = None
        K = 0
        if consistency:
            anchors = run.anchor_ids()
            K = len(anchors)
            posmaps = ctx.anchor_posmaps(prm, flat, offs, lens, anchors)
            for i in range(run.n):
                for k in range(K):
                    assert np.array_equal(_lib.posmap_view(posmaps, offs, lens, K, i, k), run.posmap(i, k)), (i, k)
        gaps = ctx.align_tree(prm, flat, offs, lens, tasks, sd, posmaps, K, 2.0)
        for i in range(run.n):
            assert np.array_equal(gaps[i], run.gaps(i)), (name, i)
        want = run.aligned()
    finally:
        run.close()
    got = ctx.kalign(seqs, n_threads=2, type_=type_, consistency=consistency, weight=2.0)
    assert got == want


def test_distance_matrix_exact(ctx):
    from kalign_b200 import _lib
    for seqs in (synth.family(90, 130, synth.PROTEIN, seed=3), synth.family(50, 1300, synth.RNA, seed=4)):
        dm_ref, anchors = kbind.ref_distance_matrix(seqs)
        run = kbind.RefRun(seqs, stop_after=1)
        # tree alphabet codes are gone after the pipeline re-encodes proteins; rebuild them from
        # the oracle-independent product path instead: compare through kb200_kalign elsewhere and
        # check the raw kernel here on the alignment alphabet restricted to codes < 13
        codes = [np.minimum(run.codes(i), 12).astype(np.uint8) for i in range(run.n)]
        run.close()
        flat, offs, lens = _lib.pack(codes)
        rows = np.arange(len(codes), dtype=np.int32)
        cols = np.arange(0, len(codes), max(1, len(codes) // 32), dtype=np.int32)[:32]
        dm = ctx.distances(flat, offs, lens, rows, cols)
        o = kbind.oracle()
        for i in range(0, len(codes), 7):
            for c, j in enumerate(cols):
                assert dm[i, c] == o.ko_pair_distance(codes[i], len(codes[i]), codes[int(j)], len(codes[int(j)])), (i, j)


def test_shuffle_and_thread_invariance(ctx):
    """the reference's DSSIM property (tests/dssim_test.c:52-70): same set, shuffled input order /
    different host thread counts -> identical alignment of every sequence."""
    seqs = synth.family(60, 90, synth.PROTEIN, seed=21)
    a = ctx.kalign(seqs, n_threads=1, consistency=5)
    b = ctx.kalign(seqs, n_threads=4, consistency=5)
    assert a == b


@pytest.mark.parametrize("name,gen,type_", FAMILIES)
@pytest.mark.parametrize("consistency", [0, 5])
def test_task_confidence(ctx, name, gen, type_, consistency):
    """task->confidence (mean meet-up margin, aln_run.c:390-394): the float sum is accumulated in the
    reference's recursion order, so it is compared for equality, not closeness"""
    from kalign_b200 import _lib
    seqs = gen()
    run = ref_state(seqs, consistency, type_)
    try:
        prm = _lib.make_params(run.biotype(), type_)
        codes = [run.codes(i) for i in range(run.n)]
        flat, offs, lens = _lib.pack(codes)
        posmaps, K = None, 0
        if consistency:
            anchors = run.anchor_ids()
            K = len(anchors)
            posmaps = ctx.anchor_posmaps(prm, flat, offs, lens, anchors)
        gaps, conf = ctx.align_tree(prm, flat, offs, lens, run.tasks(), run.seq_distances(), posmaps, K, 2.0, confidence=True)
        want = run.task_confidence()
        assert np.array_equal(conf, want), np.flatnonzero(conf != want)[:5]
        for i in range(run.n):
            assert np.array_equal(gaps[i], run.gaps(i))
    finally:
        run.close()


@pytest.mark.parametrize("name,gen,type_", FAMILIES)
@pytest.mark.parametrize("dist_scale,usw", [(0.5, 0.0), (0.0, 2.0), (0.8, 1.0)])
def test_dist_scale_and_seq_weights(ctx, name, gen, type_, dist_scale, usw):
    """the non-default aln_param fields kalign_run_seeded / kalign_run_dist_scale pass through the
    unchanged signature (kalign.h:51-57): per-task gap scaling (compute_gap_scale, aln_run.c:126-164)
    and the balanced profile merge (update_n with use_seq_weights, aln_setup.c:237-300)"""
    from kalign_b200 import _lib
    seqs = gen()
    run = kbind.RefRun(seqs, n_threads=2, type_=type_, consistency=5, weight=2.0, dist_scale=dist_scale, use_seq_weights=usw)
    try:
        subm, gp = run.params()
        prm = _lib.make_params(run.biotype(), type_)
        prm.dist_scale = dist_scale
        prm.use_seq_weights = usw
        codes = [run.codes(i) for i in range(run.n)]
        flat, offs, lens = _lib.pack(codes)
        anchors = run.anchor_ids()
        K = len(anchors)
        posmaps = ctx.anchor_posmaps(prm, flat, offs, lens, anchors)
        gaps, conf = ctx.align_tree(prm, flat, offs, lens, run.tasks(), run.seq_distances(), posmaps, K, 2.0, confidence=True)
        for i in range(run.n):
            assert np.array_equal(gaps[i], run.gaps(i)), (name, i)
        assert np.array_equal(conf, run.task_confidence())
    finally:
        run.close()


def test_bad_task_list_is_rejected(ctx):
    """unsorted / duplicated tasks fail loudly instead of dereferencing a missing profile"""
    from kalign_b200 import _lib
    seqs = synth.family(8, 40, synth.PROTEIN, seed=5)
    run = ref_state(seqs, 0, 8)
    try:
        prm = _lib.make_params(run.biotype(), 8)
        codes = [run.codes(i) for i in range(run.n)]
        flat, offs, lens = _lib.pack(codes)
        tasks = run.tasks().copy()
        bad = tasks[::-1].copy()
        with pytest.raises(RuntimeError):
            ctx.align_tree(prm, flat, offs, lens, bad, run.seq_distances())
        dup = tasks.copy()
        dup[1, 0] = dup[0, 0]
        with pytest.raises(RuntimeError):
            ctx.align_tree(prm, flat, offs, lens, dup, run.seq_distances())
        ctx.align_tree(prm, flat, offs, lens, tasks, run.seq_distances())
    finally:
        run.close()


def test_device_resident_distances(ctx):
    """kb200_seqs_upload / kb200_distances_on (the d_estimation seam keeps the msa on the device between its
    calls): rectangle and explicit pair list equal the one-shot kb200_distances"""
    from kalign_b200 import _lib
    rng = np.random.default_rng(5)
    codes = [rng.integers(0, 13, size=int(n)).astype(np.uint8) for n in rng.integers(30, 400, size=70)]
    flat, offs, lens = _lib.pack(codes)
    rows = np.arange(len(codes), dtype=np.int32)
    cols = np.array([3, 17, 40, 69, 5], dtype=np.int32)
    want = ctx.distances(flat, offs, lens, rows, cols)
    d = _lib.DeviceSeqs(ctx, flat, offs, lens)
    try:
        assert np.array_equal(d.distances(rows, cols), want)
        a = np.array([10, 11, 69, 0, 33], dtype=np.int32)
        b = np.array([3, 3, 17, 40, 5], dtype=np.int32)
        got = d.pair_distances(a, b)
        colidx = {int(c): k for k, c in enumerate(cols)}
        assert np.array_equal(got, np.array([want[int(x), colidx[int(y)]] for x, y in zip(a, b)], dtype=np.float32))
    finally:
        d.close()
