"""GPU: the drop-in library and CLI of integration/ (reference host code + GPU seams) against the
unmodified reference (oracle/_ref): the `kalign` executable writes byte-identical alignment files,
and kalign() called through lib/include/kalign/kalign.h's signature returns identical rows."""
import os
import subprocess

import pytest

import kbind
from kalign_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "integration", "_out")
LIB = os.path.join(OUT, "libkalign.so.3")
CLI = os.path.join(OUT, "kalign")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "kalign_ref")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (os.path.exists(LIB) and os.path.exists(REF_CLI)), reason="integration/_out or oracle/_ref missing")]

CASES = [
    ("protein_default", lambda: synth.family(60, 120, synth.PROTEIN, seed=21), []),
    ("protein_fast", lambda: synth.family(60, 120, synth.PROTEIN, seed=22), ["--fast"]),
    ("rna_default", lambda: synth.family(80, 300, synth.RNA, seed=23), ["--type", "rna"]),
    ("dna_fast", lambda: synth.family(30, 400, synth.DNA, seed=24, sub=0.05, ins=0.01, dele=0.01), ["--type", "dna", "--fast"]),
    ("two_sequences", lambda: synth.family(2, 90, synth.PROTEIN, seed=25), []),
    # enough sequences for dozens of bisections: d_estimation(pair = 1) is called once per leaf cluster
    # on the device-resident sequences, --refine confident reads the task->confidence the seam wrote
    ("protein_3000_fast", lambda: synth.family(3000, 100, synth.PROTEIN, seed=26), ["--fast"]),
    ("protein_refine_confident", lambda: synth.family(40, 90, synth.PROTEIN, seed=27), ["--refine", "confident"]),
    ("rna_refine_all", lambda: synth.family(24, 150, synth.RNA, seed=28), ["--type", "rna", "--refine", "all"]),
    # callers that loop over the path: the ensemble runs (kalign_ensemble, ensemble.c:286-340: per-run
    # gap penalties + noisy guide trees through the same seams) and the realign loop (kalign_run_realign,
    # aln_wrap.c:455-490: create_msa_tree -> compute_aln_pairwise_dist on the GPU -> new tree -> create_msa_tree)
    ("protein_ensemble3", lambda: synth.family(30, 100, synth.PROTEIN, seed=41), ["--ensemble", "3"]),
    ("rna_ensemble4_refine", lambda: synth.family(24, 160, synth.RNA, seed=42), ["--type", "rna", "--ensemble", "4", "--refine", "confident"]),
    ("protein_realign2", lambda: synth.family(40, 110, synth.PROTEIN, seed=43), ["--realign", "2"]),
    ("dna_realign1", lambda: synth.family(70, 260, synth.DNA, seed=44, sub=0.06, ins=0.01, dele=0.01), ["--type", "dna", "--realign", "1"]),
    ("protein_precise", lambda: synth.family(30, 100, synth.PROTEIN, seed=45), ["--precise"]),
]


def _write_fasta(path, seqs):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">s%d\n%s\n" % (i, s))


@pytest.mark.parametrize("name,gen,flags", CASES)
def test_cli_output_identical(tmp_path, name, gen, flags):
    fa = str(tmp_path / "in.fa")
    _write_fasta(fa, gen())
    outs = {}
    for tag, exe in (("gpu", CLI), ("ref", REF_CLI)):
        out = str(tmp_path / (tag + ".afa"))
        p = subprocess.run([exe, "-i", fa, "-o", out, "-n", "4"] + flags, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert p.returncode == 0, p.stdout[-2000:]
        outs[tag] = open(out, "rb").read()
    assert len(outs["gpu"]) > 0
    assert outs["gpu"] == outs["ref"]


_CALL_KALIGN = r"""
import ctypes as C, json, sys
lib = C.CDLL(sys.argv[1])
seqs = json.load(open(sys.argv[2]))
lib.kalign.restype = C.c_int
lib.kalign.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int,
                       C.c_float, C.c_float, C.c_float, C.POINTER(C.POINTER(C.c_char_p)), C.POINTER(C.c_int)]
arr = (C.c_char_p * len(seqs))(*[s.encode() for s in seqs])
lens = (C.c_int * len(seqs))(*[len(s) for s in seqs])
out = C.POINTER(C.c_char_p)()
alen = C.c_int(0)
rc = lib.kalign(arr, lens, len(seqs), 2, 8, -1.0, -1.0, -1.0, C.byref(out), C.byref(alen))
assert rc == 0, rc
json.dump([out[i][:alen.value].decode() for i in range(len(seqs))], open(sys.argv[3], "w"))
"""


def test_kalign_entry_point_identical(tmp_path):
    """int kalign(char**, int*, int, int, int, float, float, float, char***, int*)  (kalign.h:45).
    Each library is loaded in its own process: both export the same symbol names."""
    import json
    import sys
    seqs = synth.family(40, 100, synth.PROTEIN, seed=31)
    inp = str(tmp_path / "seqs.json")
    json.dump(seqs, open(inp, "w"))
    rows = {}
    for tag, path in (("gpu", LIB), ("ref", kbind.REF_SO)):
        outp = str(tmp_path / (tag + ".json"))
        p = subprocess.run([sys.executable, "-c", _CALL_KALIGN, path, inp, outp], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert p.returncode == 0, p.stdout[-2000:]
        rows[tag] = json.load(open(outp))
    assert rows["gpu"] == rows["ref"]
    assert all(r.replace("-", "") == s for r, s in zip(rows["gpu"], seqs))


@pytest.mark.parametrize("kind,flags,kw", [("protein", [], dict(type_=8, consistency=5)), ("rna", ["--type", "rna", "--fast"], dict(type_=2, consistency=0))])
def test_kalign_file_identical_to_reference_cli(tmp_path, kind, flags, kw):
    """kb200_kalign_file (FASTA file in, FASTA file out through the product's own reader / writer) against the
    reference CLI on a file whose record names decide the order of equal-length sequences, with gap characters,
    lower case, an empty record and CRLF line ends in the input"""
    from kalign_b200 import _lib
    seqs = synth.family(60, 90, synth.PROTEIN if kind == "protein" else synth.RNA, seed=51)
    seqs = [s[:80] if i % 3 == 0 else s for i, s in enumerate(seqs)]          # many equal lengths
    fa = str(tmp_path / "in.fa")
    with open(fa, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">%c%d some text\r\n" % (b"zyxw"[i % 4], 1000 - i))
            body = s if i % 5 else s.lower()
            if i == 7:
                body = body[:30] + "--" + body[30:]
            for j in range(0, len(body), 50):
                f.write(body[j:j + 50].encode() + b"\r\n")
            if i == 11:
                f.write(b">empty record\r\n")
    ref_out = str(tmp_path / "ref.afa")
    p = subprocess.run([REF_CLI, "-i", fa, "-o", ref_out, "-n", "4"] + flags, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout[-2000:]
    ctx = _lib.Context(0)
    try:
        out = str(tmp_path / "gpu.afa")
        ctx.kalign_file(fa, out, n_threads=4, weight=2.0, **kw)
    finally:
        ctx.close()
    assert open(out, "rb").read() == open(ref_out, "rb").read()
