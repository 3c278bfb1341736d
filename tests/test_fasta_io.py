"""FASTA in / out: kb200_fasta_read / kb200_fasta_write against the reference's read_fasta
and write_msa_fasta (lib/src/msa_io.c:412,668).

Host code: everything here runs without a GPU.  Three arms: the product (libkalign_b200.so through the C ABI),
a pure-python restatement of the reference's rules (kbind.oracle_read_fasta / oracle_write_fasta), and -- where
oracle/_ref exists -- the unmodified reference itself (kalign_read_input / kalign_write_msa through
oracle/ref_harness.c).  tests/golden/fasta_io.npz pins the restatement to outputs the reference produced in the
build container (tools/gen_golden_fasta.py)."""
import os

import numpy as np
import pytest

import kbind
from kalign_b200 import _lib

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fasta_io.npz")

# files that exercise every rule of read_file_stdin / read_fasta
CASES = {
    "plain": b">a\nACDEFG\nHIKL\n>b desc here\nMNPQ\n",
    "crlf": b">a\r\nACGT\r\nAC\r\n>b\r\nGGTT\r\n",
    "no_final_newline": b">a\nACGT\n>b\nTTGA",
    "blank_lines": b">a\n\nAC\n\n\nGT\n>b\n\nTT\n\n",
    "lower_digits_spaces": b">a\n  1 acgt acgt 10\n 11 ACgt\n>b\n60 tt aa\n",
    "gaps_and_stop": b">a\nAC--GT..A*\n--AC\n>b\n-\nA-C-\n",
    "tab_truncates": b">a\tcomment\nACGT\tTTTT\nGG\n>b\nAC\x01GT\nT\x7fA\n",
    # (a one-character FIRST line makes kalign_read_input report "no input", msa_io.c:105-116: the empty name comes second)
    "empty_name_and_record": b">w\nACGT\n>\nAC\n>x\n>y\nAA\n>z\n",
    "gt_inside_line": b">a\nAC>GT\nA>\n>b\nTT\n",
    "long_name": b">" + b"n" * 700 + b" tail\nACGT\n>b\nAC\n",
    "long_line": b">a\n" + b"ACGTN" * 5000 + b"\n>b\n" + b"acgt-" * 3000 + b"\n",
    "header_only_spaces": b">   \nAC\n>\t\nGT\n",
    "leading_blank": b"\n\n>a\nACGT\n>b\nAA\n",
    "nul_inside": b">a\nAC\x00GT\nTT\n>b\nGG\n",
}
FAILS = {
    "residues_before_header": b"ACGT\n>a\nAC\n>b\nGG\n",
}


def write(tmp_path, name, data):
    p = str(tmp_path / (name + ".fa"))
    with open(p, "wb") as f:
        f.write(data)
    return p


def same(a, b):
    ra, fa = a
    rb, fb = b
    assert len(ra) == len(rb)
    for (na, sa, ga), (nb, sb, gb) in zip(ra, rb):
        assert na == nb and sa == sb and np.array_equal(ga, gb), (na, nb)
    assert np.array_equal(fa, fb)


def product(path, n_threads=0):
    f = _lib.Fasta(path, n_threads)
    try:
        return f.records(), f.letter_freq()
    finally:
        f.close()


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("threads", [1, 3])
def test_read_equals_restatement(tmp_path, name, threads):
    p = write(tmp_path, name, CASES[name])
    same(product(p, threads), kbind.oracle_read_fasta(CASES[name]))


def test_restatement_reproduces_golden():
    z = np.load(G)
    for name in sorted(CASES):
        recs, freq = kbind.oracle_read_fasta(CASES[name])
        assert bytes(z["file_" + name]) == CASES[name], "golden generated from a different file: " + name
        assert [r[0] for r in recs] == [bytes(x) for x in z["names_" + name]], name
        assert [r[1] for r in recs] == [bytes(x) for x in z["seqs_" + name]], name
        assert np.array_equal(np.concatenate([r[2] for r in recs]) if recs else np.zeros(0, np.int32), z["gaps_" + name]), name
        assert np.array_equal(freq, z["freq_" + name]), name
    assert kbind.oracle_write_fasta([bytes(x) for x in z["w_names"]], [bytes(x) for x in z["w_rows"]]) == bytes(z["w_file"])


@pytest.mark.skipif(not kbind.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", sorted(CASES))
def test_read_equals_reference(tmp_path, name):
    p = write(tmp_path, name, CASES[name])
    ref = kbind.ref_read_fasta(p)
    assert ref is not None
    same(product(p), ref)
    same(kbind.oracle_read_fasta(CASES[name]), ref)


@pytest.mark.parametrize("name", sorted(FAILS))
def test_read_failures(tmp_path, name):
    p = write(tmp_path, name, FAILS[name])
    assert kbind.oracle_read_fasta(FAILS[name]) is None
    with pytest.raises(RuntimeError):
        _lib.Fasta(p)
    with pytest.raises(RuntimeError):
        _lib.Fasta(str(tmp_path / "does_not_exist.fa"))


def test_empty_file(tmp_path):
    p = write(tmp_path, "empty", b"")
    recs, freq = product(p)
    assert recs == [] and not freq.any()


def random_file(rng):
    pieces = [b">", b"\n", b"\r\n", b"\t", b" ", b"-", b".", b"*", b"1", b"\x01", b"ACGT", b"acgu", b"MKV", b"n", b">x y", b"\n>",
              b"\n\n", b"WYV" * 30]
    out = bytearray(b">first\n" if rng.random() < 0.8 else b"\n>f\n")
    for _ in range(int(rng.integers(0, 120))):
        out += pieces[int(rng.integers(0, len(pieces)))]
    return bytes(out)


def test_read_fuzz(tmp_path):
    rng = np.random.default_rng(7)
    use_ref = kbind.have_ref()
    for it in range(300):
        data = random_file(rng)
        want = kbind.oracle_read_fasta(data)
        p = write(tmp_path, "fuzz", data)
        if want is None:
            with pytest.raises(RuntimeError):
                _lib.Fasta(p)
            continue
        same(product(p, 1 + it % 4), want)
        # the reference treats a one-character first line as "no input" and wants >= 2 records (msa_io.c:105-116,176)
        if use_ref and len(want[0]) >= 2 and len(data.split(b"\n")[0].split(b"\r")[0]) != 1:
            ref = kbind.ref_read_fasta(p)
            if ref is not None:
                same(want, ref)


def rows_for(n, alnlen, seed):
    rng = np.random.default_rng(seed)
    a = rng.choice(np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY-", dtype=np.uint8), size=(n, max(alnlen, 1)))[:, :alnlen]
    names = [("seq_%d some description" % i).encode() for i in range(n)]
    if n:
        names[0] = b""
    return names, [bytes(r) for r in a]


@pytest.mark.parametrize("n,alnlen", [(1, 1), (3, 59), (3, 60), (4, 61), (5, 120), (2, 0), (40, 1234), (0, 0)])
def test_write_equals_restatement_and_reference(tmp_path, n, alnlen):
    names, rows = rows_for(n, alnlen, n * 100 + alnlen)
    p = str(tmp_path / "out.afa")
    _lib.fasta_write(p, names, rows, n_threads=3)
    got = open(p, "rb").read()
    assert got == kbind.oracle_write_fasta(names, rows)
    if kbind.have_ref() and n > 0:
        q = str(tmp_path / "ref.afa")
        kbind.ref_write_fasta(q, names, rows)
        assert got == open(q, "rb").read()


def test_round_trip_large(tmp_path):
    """write -> read returns the rows (gap characters as gap counts), multi-threaded on a file of ~6 MB"""
    names, rows = rows_for(3000, 2000, 5)
    names[0] = b"first"
    p = str(tmp_path / "big.afa")
    _lib.fasta_write(p, names, rows)
    recs, freq = product(p)
    assert len(recs) == 3000
    for (nm, sq, gp), name, row in zip(recs, names, rows):
        assert nm == name and sq == row.replace(b"-", b"")
        assert int(gp.sum()) == row.count(b"-") and len(gp) == len(sq) + 1
    same((recs, freq), kbind.oracle_read_fasta(open(p, "rb").read()))
