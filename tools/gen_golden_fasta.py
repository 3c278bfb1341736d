"""Golden vectors for the FASTA reader / writer from the UNMODIFIED reference (oracle/_ref: kalign_read_input,
kalign_write_msa through oracle/ref_harness.c) -> tests/golden/fasta_io.npz.  The input files are the CASES of
tests/test_fasta_io.py.  Run once in the build container:  python tools/gen_golden_fasta.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kbind  # noqa: E402
import test_fasta_io as T  # noqa: E402


def main():
    rec = {}
    with tempfile.TemporaryDirectory() as d:
        for name, data in sorted(T.CASES.items()):
            p = os.path.join(d, name + ".fa")
            open(p, "wb").write(data)
            ref = kbind.ref_read_fasta(p)
            assert ref is not None, name
            recs, freq = ref
            rec["file_" + name] = np.frombuffer(data, dtype=np.uint8)
            rec["names_" + name] = np.array([r[0] for r in recs], dtype=object).astype("S")
            rec["seqs_" + name] = np.array([r[1] for r in recs], dtype=object).astype("S")
            rec["gaps_" + name] = np.concatenate([r[2] for r in recs]) if recs else np.zeros(0, np.int32)
            rec["freq_" + name] = freq
        names, rows = T.rows_for(7, 133, 99)
        names[0] = b"first record"
        q = os.path.join(d, "w.afa")
        kbind.ref_write_fasta(q, names, rows)
        rec["w_names"] = np.array(names, dtype="S")
        rec["w_rows"] = np.array(rows, dtype="S")
        rec["w_file"] = np.frombuffer(open(q, "rb").read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fasta_io.npz"), **rec)
    print("fasta_io:", len(T.CASES), "read cases + 1 write case")


if __name__ == "__main__":
    main()
