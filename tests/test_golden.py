"""Committed golden vectors (tests/golden/*.npz, generated from the real reference by
tools/gen_golden.py): the oracle restatement must reproduce them on CPU; the GPU path must
reproduce them through the C ABI (-m gpu).  Needs neither /root/reference nor oracle/_ref."""
import glob
import os

import numpy as np
import pytest

import kbind

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_oracle = pytest.mark.skipif(not kbind.have_oracle(), reason="oracle lib not built")


def _jobs():
    z = np.load(os.path.join(G, "pairs.npz"))
    for n in range(int(z["njobs"])):
        kind = int(z["j%d_kind" % n])
        gp = z["j%d_gp" % n]
        d = dict(kind=kind, subm=np.ascontiguousarray(z["j%d_subm" % n]), gpo=float(gp[0]), gpe=float(gp[1]),
                 tgpe=float(gp[2]), soff=float(gp[3]), path=z["j%d_path" % n], score=float(z["j%d_score" % n]),
                 margin=z["j%d_margin" % n])
        if kind == 0:
            d.update(seq1=z["j%d_seq1" % n], seq2=z["j%d_seq2" % n])
            d.update(len_a=len(d["seq1"]), len_b=len(d["seq2"]), sip=0)
        elif kind == 1:
            l = z["j%d_len" % n]
            d.update(prof1=np.ascontiguousarray(z["j%d_prof1" % n]), seq2=z["j%d_seq2" % n], len_a=int(l[0]), len_b=int(l[1]), sip=int(l[2]))
        else:
            l = z["j%d_len" % n]
            d.update(prof1=np.ascontiguousarray(z["j%d_prof1" % n]), prof2=np.ascontiguousarray(z["j%d_prof2" % n]),
                     len_a=int(l[0]), len_b=int(l[1]), sip=0)
        yield d


@needs_oracle
def test_oracle_reproduces_golden_pairs():
    n = 0
    for j in _jobs():
        p, st = kbind.oracle_align(j["kind"], j["len_a"], j["len_b"], j["subm"], j["gpo"], j["gpe"], j["tgpe"], soff=j["soff"],
                                   seq1=j.get("seq1"), seq2=j.get("seq2"), prof1=j.get("prof1"), prof2=j.get("prof2"), sip=j["sip"])
        la = j["len_a"]
        assert np.array_equal(p[1:la + 1], j["path"][1:la + 1])
        assert np.float32(st["top_score"]) == np.float32(j["score"])
        assert np.float32(st["margin_sum"]) == np.float32(j["margin"][0]) and st["margin_count"] == int(j["margin"][1])
        n += 1
    assert n == 22


@needs_oracle
def test_oracle_reproduces_golden_bpm():
    z = np.load(os.path.join(G, "bpm.npz"))
    o = kbind.oracle()
    for i in range(int(z["n"])):
        t, p = np.ascontiguousarray(z["t%d" % i]), np.ascontiguousarray(z["p%d" % i])
        assert o.ko_bpm_block(t, p, len(t), len(p)) == int(z["d%d" % i])


def test_golden_msa_files_are_consistent():
    files = sorted(glob.glob(os.path.join(G, "msa_*.npz")))
    assert len(files) == 4
    for f in files:
        z = np.load(f)
        seqs = [str(s) for s in z["seqs"]]
        rows = [str(s) for s in z["aligned"]]
        assert len(set(len(r) for r in rows)) == 1
        assert [r.replace("-", "") for r in rows] == seqs
        assert z["tasks"].shape == (len(seqs) - 1, 3)


@pytest.mark.gpu
def test_gpu_reproduces_golden_pairs():
    from kalign_b200 import _lib
    ctx = _lib.Context(0)
    for j in _jobs():
        prm = _lib.params_from(j["subm"], j["gpo"], j["gpe"], j["tgpe"], nalpha=23 if j["gpo"] == 7.0 else 5)
        job = dict(kind=j["kind"], len_a=j["len_a"], len_b=j["len_b"], sip=j["sip"], soff=j["soff"])
        if j["kind"] == 0:
            job.update(seq_rows=j["seq1"], seq_cols=j["seq2"])
        elif j["kind"] == 1:
            job.update(prof_rows=j["prof1"], seq_cols=j["seq2"])
        else:
            job.update(prof_rows=j["prof1"], prof_cols=j["prof2"])
        paths, scores = ctx.pair_align_batch(prm, [job])
        la = j["len_a"]
        assert np.array_equal(paths[0][1:la + 1], j["path"][1:la + 1]), j["kind"]
        assert abs(float(scores[0]) - j["score"]) <= 1e-5 * max(1.0, abs(j["score"]))
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["c1_protein_default", "protein_fast", "rna_default", "dna_default"])
def test_gpu_reproduces_golden_msa(tag):
    from kalign_b200 import _lib
    z = np.load(os.path.join(G, "msa_%s.npz" % tag))
    seqs = [str(s) for s in z["seqs"]]
    ctx = _lib.Context(0)
    # seam level: the reference's own guide tree / distances in, gaps out
    n = len(seqs)
    codes = [np.ascontiguousarray(z["codes%d" % i]) for i in range(n)]
    flat, offs, lens = _lib.pack(codes)
    biotype = 0 if tag.startswith(("c1", "protein")) else 1
    prm = _lib.make_params(biotype, int(z["type"]))
    assert np.array_equal(np.array(prm.subm[:], dtype=np.float32).reshape(23, 23), z["subm"])
    K = int(z["consistency"])
    posmaps = None
    if K:
        posmaps = ctx.anchor_posmaps(prm, flat, offs, lens, z["anchor_ids"])
        K = len(z["anchor_ids"])
    gaps = ctx.align_tree(prm, flat, offs, lens, z["tasks"], z["seq_distances"], posmaps, K, 2.0)
    for i in range(n):
        assert np.array_equal(gaps[i], z["gaps%d" % i]), (tag, i)
    # whole pipeline (own guide tree): byte-identical rows
    got = ctx.kalign(seqs, n_threads=2, type_=int(z["type"]), consistency=int(z["consistency"]), weight=2.0)
    assert got == [str(s) for s in z["aligned"]]
    ctx.close()
