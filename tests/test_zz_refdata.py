"""The input files of the reference's own test suite (kalign_itest_*, kalign_api_test, kalign_ensemble_test and the
reader fixtures of /root/reference/tests/data; tests/CMakeLists.txt:55-127) as golden vectors.

tests/golden/refdata.npz (tools/gen_golden_refdata.py) holds each file's bytes, the records the UNMODIFIED reference
reads from it, and the files the reference CLI writes for it in default mode and with --fast.
CPU: kb200_fasta_read and the python restatement read the same records.
GPU: kb200_kalign_file (the product's reader, alignment and writer in one call) and the drop-in CLI write the
reference's output files byte for byte."""
import os
import subprocess

import numpy as np
import pytest

import kbind
from kalign_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Z = np.load(os.path.join(ROOT, "tests", "golden", "refdata.npz"))
FASTA = [str(x) for x in Z["fasta"]]
ALIGN = [str(x) for x in Z["align"]]
CLI = os.path.join(ROOT, "integration", "_out", "kalign")


def _write(tmp_path, name):
    p = str(tmp_path / name)
    with open(p, "wb") as f:
        f.write(bytes(Z["file_" + name]))
    return p


@pytest.mark.parametrize("name", FASTA)
def test_reader_on_reference_test_files(tmp_path, name):
    p = _write(tmp_path, name)
    f = _lib.Fasta(p, 2)
    try:
        recs, freq = f.records(), f.letter_freq()
    finally:
        f.close()
    ora = kbind.oracle_read_fasta(bytes(Z["file_" + name]))
    for got in ((recs, freq), ora):
        assert [r[0] for r in got[0]] == [bytes(x) for x in Z["names_" + name]]
        assert [r[1] for r in got[0]] == [bytes(x) for x in Z["seqs_" + name]]
        assert np.array_equal(np.concatenate([r[2] for r in got[0]]), Z["gaps_" + name])
        assert np.array_equal(got[1], Z["freq_" + name])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["default", "fast"])
@pytest.mark.parametrize("name", ALIGN)
def test_alignments_of_reference_test_files(tmp_path, name, mode):
    p = _write(tmp_path, name)
    want = bytes(Z["out_%s_%s" % (mode, name)])
    ctx = _lib.Context(0)
    try:
        out = str(tmp_path / "gpu.afa")
        ctx.kalign_file(p, out, n_threads=4, type_=8, consistency=5 if mode == "default" else 0, weight=2.0)
    finally:
        ctx.close()
    assert open(out, "rb").read() == want
    if os.path.exists(CLI):
        out2 = str(tmp_path / "cli.afa")
        r = subprocess.run([CLI, "-i", p, "-o", out2, "-n", "4"] + (["--fast"] if mode == "fast" else []),
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-2000:]
        assert open(out2, "rb").read() == want
