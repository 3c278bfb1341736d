"""The input files of the reference's OWN test suite (tests/CMakeLists.txt:55-127: kalign_itest_*, kalign_api_test,
kalign_ensemble_test run on tests/data/BB11001.tfa, BB12006.tfa, BB30014.tfa; tests/data/*.good.*, small.fa, tiny.fa
are the reader's fixtures) as golden vectors: every file's bytes, what the UNMODIFIED reference reads from it
(kalign_read_input, lib/src/msa_io.c:80) and what the reference CLI writes for it (oracle/_ref/kalign_ref, default
mode and --fast) -> tests/golden/refdata.npz.  /root/reference does not travel to the GPU box; the vectors do.
Run once in the build container:  python tools/gen_golden_refdata.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kbind  # noqa: E402

DATA = "/root/reference/tests/data"
FASTA = ["BB11001.tfa", "BB12006.tfa", "BB30014.tfa", "small.fa", "tiny.fa", "tiny_internal.fa",
         "afa.good.1", "afa.good.2", "afa.good.3", "a2m.good.1", "a2m.good.2"]
ALIGN = ["BB11001.tfa", "BB12006.tfa", "BB30014.tfa", "small.fa"]          # what the reference's itests align
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "kalign_ref")


def main():
    rec = {"fasta": np.array(FASTA), "align": np.array(ALIGN)}
    with tempfile.TemporaryDirectory() as d:
        for name in FASTA:
            p = os.path.join(DATA, name)
            data = open(p, "rb").read()
            rec["file_" + name] = np.frombuffer(data, dtype=np.uint8)
            ref = kbind.ref_read_fasta(p)
            assert ref is not None, name
            recs, freq = ref
            rec["names_" + name] = np.array([r[0] for r in recs], dtype=object).astype("S")
            rec["seqs_" + name] = np.array([r[1] for r in recs], dtype=object).astype("S")
            rec["gaps_" + name] = np.concatenate([r[2] for r in recs])
            rec["freq_" + name] = freq
        for name in ALIGN:
            for tag, flags in (("default", []), ("fast", ["--fast"])):
                out = os.path.join(d, name + "." + tag)
                p = subprocess.run([REF_CLI, "-i", os.path.join(DATA, name), "-o", out, "-n", "4"] + flags,
                                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                assert p.returncode == 0, p.stdout[-1000:]
                rec["out_%s_%s" % (tag, name)] = np.frombuffer(open(out, "rb").read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "refdata.npz"), **rec)
    print("refdata:", len(FASTA), "files read,", 2 * len(ALIGN), "alignments")


if __name__ == "__main__":
    main()
