"""GPU: edge cases of the whole alignment call against the unmodified reference (oracle/_ref):
ragged and unrelated inputs, length-1 sequences, duplicates, the smallest families (N = 2, 3: below /
at the consistency minimum), ambiguity codes and lower case, one long sequence among short ones."""
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


def _rand(rng, alphabet, n):
    a = np.frombuffer(alphabet.encode(), dtype=np.uint8)
    return a[rng.integers(0, len(a), size=n)].tobytes().decode()


def _cases():
    rng = np.random.default_rng(77)
    ragged = [_rand(rng, synth.PROTEIN, int(l)) for l in [1, 1, 2, 3, 5, 8, 13, 31, 32, 33, 64, 65, 100, 127, 128, 129, 200, 257, 300, 400]]
    ragged_nt = [_rand(rng, "ACGU", int(l)) for l in rng.integers(1, 500, size=30)]
    dup = [synth.family(1, 150, synth.PROTEIN, seed=5)[0]] * 20
    two = synth.family(2, 220, synth.PROTEIN, seed=6)
    three = synth.family(3, 180, synth.RNA, seed=7)
    fam = synth.family(12, 90, synth.PROTEIN, seed=8)
    ambig = [s[:10] + "XBZ" + s[10:40].lower() + "x" + s[40:] for s in fam]
    fam_nt = synth.family(12, 160, synth.DNA, seed=9)
    ambig_nt = [s[:20] + "NNRY" + s[20:60].lower() + s[60:] for s in fam_nt]
    long_short = [_rand(rng, synth.PROTEIN, 2000)] + synth.family(15, 30, synth.PROTEIN, seed=10)
    return [
        ("ragged_protein", ragged, 8), ("ragged_rna", ragged_nt, 2), ("duplicates", dup, 8), ("two", two, 8),
        ("three_rna", three, 2), ("ambiguity_protein", ambig, 8), ("ambiguity_dna", ambig_nt, 0), ("long_and_short", long_short, 8),
    ]


@pytest.mark.parametrize("consistency", [0, 5])
@pytest.mark.parametrize("name,seqs,type_", _cases(), ids=[c[0] for c in _cases()])
def test_edge_case_identical_to_reference(ctx, name, seqs, type_, consistency):
    got = ctx.kalign(seqs, n_threads=3, type_=type_, consistency=consistency, weight=2.0)
    run = kbind.RefRun(seqs, n_threads=3, type_=type_, consistency=consistency, weight=2.0)
    try:
        want = run.aligned()
    finally:
        run.close()
    assert got == want
    assert all(r.replace("-", "") == s for r, s in zip(got, seqs))
