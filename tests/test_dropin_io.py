"""The drop-in library's kalign_read_input / kalign_write_msa (integration/kalign_gpu_seams.c over
kb200_fasta_read / kb200_fasta_write) against the unmodified reference's (oracle/_ref): the same C driver
(tests/io_driver.c) is linked against either library and must print the same struct msa -- names, residues,
gap counts, letter frequencies, detected alphabet and alignment state, allocation sizes -- and write the same
file.  Host code only: runs without a GPU.  Needs the reference's headers (struct msa), so it is skipped
where /root/reference is absent."""
import os
import re
import subprocess

import numpy as np
import pytest

import kbind
from test_fasta_io import CASES, random_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/lib/src"
DROPIN = os.path.join(ROOT, "integration", "_out")
REFDIR = os.path.join(ROOT, "oracle", "_ref")

pytestmark = pytest.mark.skipif(not (os.path.exists(os.path.join(REF_SRC, "msa_struct.h")) and
                                     os.path.exists(os.path.join(DROPIN, "libkalign.so.3")) and kbind.have_ref()),
                                reason="reference headers / integration/_out / oracle/_ref missing")


@pytest.fixture(scope="module")
def drivers(tmp_path_factory):
    d = tmp_path_factory.mktemp("io_driver")
    out = {}
    for tag, libdir, lib, extra in (("gpu", DROPIN, "kalign", [os.path.join(ROOT, "kalign_b200")]), ("ref", REFDIR, "kalign_ref", [])):
        exe = str(d / ("io_" + tag))
        cmd = ["/usr/bin/gcc", "-O1", "-w", "-I" + REF_SRC, os.path.join(ROOT, "tests", "io_driver.c"), "-o", exe,
               "-L" + libdir, "-l" + lib, "-Wl,-rpath," + libdir] + ["-Wl,-rpath-link," + e for e in extra]
        subprocess.run(cmd, check=True)
        out[tag] = exe
    return out


def run_both(drivers, tmp_path, data, fmt=None):
    fa = str(tmp_path / "in.fa")
    with open(fa, "wb") as f:
        f.write(data)
    res = {}
    for tag, exe in drivers.items():
        os.makedirs(str(tmp_path / tag), exist_ok=True)
        outp = str(tmp_path / tag / "out.aln")          # same file name: the MSF header quotes it
        if os.path.exists(outp):
            os.remove(outp)
        p = subprocess.run([exe, fa, outp] + ([fmt] if fmt else []), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
        res[tag] = (p.returncode, p.stdout, open(outp, "rb").read() if os.path.exists(outp) else None)
    return res


@pytest.mark.parametrize("name", sorted(CASES))
def test_struct_msa_identical(drivers, tmp_path, name):
    r = run_both(drivers, tmp_path, CASES[name])
    assert r["gpu"][0] == 0 and r["gpu"] == r["ref"]
    assert b"numseq" in r["gpu"][1]


def aligned_file(n, alnlen, seed, width):
    rng = np.random.default_rng(seed)
    a = rng.choice(np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY---", dtype=np.uint8), size=(n, alnlen))
    a[:, 0] = ord("M")
    out = bytearray()
    for i in range(n):
        out += b">seq%d\n" % i
        row = bytes(a[i])
        for j in range(0, alnlen, width):
            out += row[j:j + width] + b"\n"
    return bytes(out)


@pytest.mark.parametrize("n,alnlen,width,fmt", [(5, 130, 60, None), (40, 61, 1000, "fasta"), (3, 60, 7, "fa"),
                                                 (6, 90, 60, "msf"), (6, 90, 60, "clu")])
def test_alignment_read_finalise_write_identical(drivers, tmp_path, n, alnlen, width, fmt):
    """an aligned FASTA file in, the same alignment out: FASTA output goes through kb200_fasta_write,
    MSF / Clustal output through the reference's writers -- all identical to the reference build"""
    r = run_both(drivers, tmp_path, aligned_file(n, alnlen, n + alnlen, width), fmt)
    if fmt == "msf":
        # the MSF header carries the time of writing ("October 17, 2026 11:15"): the two runs may straddle a minute
        stamp = re.compile(rb"[A-Z][a-z]+ +\d+, \d{4} +\d\d:\d\d")
        r = {k: (v[0], v[1], stamp.sub(b"<time>", v[2]) if v[2] is not None else None) for k, v in r.items()}
    assert r["gpu"] == r["ref"]
    assert r["gpu"][2] is not None and len(r["gpu"][2]) > n * alnlen and b"write 0" in r["gpu"][1]


def test_other_formats_and_quirks_identical(drivers, tmp_path):
    """inputs the seam must hand to the reference's reader: Clustal, MSF markers inside a FASTA header,
    a one-character first line, an empty file"""
    clustal = b"CLUSTAL W (1.83) multiple sequence alignment\n\n" + b"s1   ACGT-ACGT\ns2   ACGTTACGT\n\n"
    cases = [clustal, b">a MSF: x\nACGT\n>b\nAC\n", b">\nACGT\n>b\nAC\n", b"", b">only\nACGT\n", b"no fasta here\n"]
    for data in cases:
        r = run_both(drivers, tmp_path, data)
        assert r["gpu"] == r["ref"], data[:30]


def test_fuzz_identical(drivers, tmp_path):
    rng = np.random.default_rng(11)
    n = 0
    for _ in range(120):
        data = random_file(rng)
        want = kbind.oracle_read_fasta(data)
        # punctuation before the first header makes the reference dereference a NULL record (msa_io.c:467)
        if want is None:
            continue
        r = run_both(drivers, tmp_path, data)
        assert r["gpu"] == r["ref"], data
        n += 1
    assert n > 60
