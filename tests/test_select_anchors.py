"""CPU: kb200_select_anchors (host helper of the C ABI) against the reference's static select_anchors
(lib/src/anchor_consistency.c:124-198), via the reference pipeline's own seq_distances / anchor ids."""
import numpy as np
import pytest

import kbind
from kalign_b200 import _lib, synth

pytestmark = pytest.mark.skipif(not kbind.have_ref(), reason="oracle/_ref missing")


@pytest.mark.parametrize("n,length,alphabet,type_,K", [(40, 80, synth.PROTEIN, 8, 5), (25, 150, synth.RNA, 2, 5),
                                                     (12, 60, synth.PROTEIN, 8, 8), (6, 90, synth.DNA, 0, 3)])
def test_select_anchors_matches_reference(n, length, alphabet, type_, K):
    seqs = synth.family(n, length, alphabet, seed=100 + n)
    run = kbind.RefRun(seqs, n_threads=2, type_=type_, consistency=K, weight=2.0)
    try:
        sd = run.seq_distances()
        want = run.anchor_ids()
    finally:
        run.close()
    assert len(want) == min(K, n)
    got = np.zeros(len(want), dtype=np.int32)
    assert _lib.load().kb200_select_anchors(np.ascontiguousarray(sd, dtype=np.float32), n, len(want), got) == 0
    assert np.array_equal(got, want)


def test_select_anchors_rejects_bad_arguments():
    lib = _lib.load()
    sd = np.zeros(4, dtype=np.float32)
    out = np.zeros(8, dtype=np.int32)
    assert lib.kb200_select_anchors(sd, 4, 0, out) != 0
    assert lib.kb200_select_anchors(sd, 4, 5, out) != 0
