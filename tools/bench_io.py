"""FASTA in / out timing, host code only -- runs without a GPU.

The product's kb200_fasta_read / kb200_fasta_write against the unmodified reference's kalign_read_input /
kalign_write_msa (oracle/_ref through oracle/ref_harness.c) on the same files: the input of a BASELINE
workload written as FASTA (60 residues per line) and an alignment-shaped output (n rows x alnlen).  Only the C
calls are timed (argument marshalling is done before).  usage: python tools/bench_io.py [C4|C3|C2] [alnlen]"""
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kbind  # noqa: E402
from kalign_b200 import _lib, synth  # noqa: E402


def best(fn, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C4"
    seqs = synth.config(wl)
    n = len(seqs)
    alnlen = int(sys.argv[2]) if len(sys.argv) > 2 else int(1.6 * max(len(s) for s in seqs))
    lib = _lib.load()
    out = {"workload": wl, "nseq": n, "threads": None, "where": "host cores of the machine this ran on (no GPU involved)"}
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "in.fa")
        with open(fa, "w") as f:
            for i, s in enumerate(seqs):
                f.write(">s%d\n" % i)
                for j in range(0, len(s), 60):
                    f.write(s[j:j + 60] + "\n")
        out["input_bytes"] = os.path.getsize(fa)
        path = os.fsencode(fa)

        def prod_read():
            h = C.c_void_p()
            assert lib.kb200_fasta_read(path, 0, C.byref(h)) == 0
            assert lib.kb200_fasta_numseq(h) == n
            lib.kb200_fasta_free(h)

        out["read_seconds"] = best(prod_read, 5)
        if kbind.have_ref():
            rl = kbind.refh()

            def ref_read():
                h = rl.refh_read_input(path)
                assert h and rl.refh_msa_numseq(h) == n
                rl.refh_msa_free(h)

            out["read_seconds_reference"] = best(ref_read, 2)
            # same content
            a = kbind.ref_read_fasta(fa)
            f = _lib.Fasta(fa)
            b = (f.records(), f.letter_freq())
            f.close()
            out["read_identical"] = bool(len(a[0]) == len(b[0]) and all(x[0] == y[0] and x[1] == y[1] and np.array_equal(x[2], y[2])
                                                                         for x, y in zip(a[0], b[0])) and np.array_equal(a[1], b[1]))
        # alignment-shaped output
        rng = np.random.default_rng(1)
        rows = rng.choice(np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY-", dtype=np.uint8), size=(n, alnlen))
        rows_b = [bytes(r) for r in rows]
        names_b = [b"s%d" % i for i in range(n)]
        nm = (C.c_char_p * n)(*names_b)
        rw = (C.c_char_p * n)(*rows_b)
        o1 = os.fsencode(os.path.join(d, "prod.afa"))
        o2 = os.fsencode(os.path.join(d, "ref.afa"))
        out["alnlen"] = alnlen

        def prod_write():
            assert lib.kb200_fasta_write(o1, nm, rw, n, alnlen, 0) == 0

        out["write_seconds"] = best(prod_write, 3)
        out["output_bytes"] = os.path.getsize(o1)
        if kbind.have_ref():
            def ref_write():
                assert rl.refh_write_rows(nm, rw, n, alnlen, o2, b"fasta") == 0

            out["write_seconds_reference"] = best(ref_write, 2)
            out["write_identical"] = open(o1, "rb").read() == open(o2, "rb").read()
    import ctypes.util  # noqa: F401
    out["threads"] = int(os.environ.get("OMP_NUM_THREADS", 0)) or os.cpu_count()
    for k in ("read", "write"):
        if k + "_seconds_reference" in out:
            out[k + "_speedup"] = out[k + "_seconds_reference"] / out[k + "_seconds"]
    out["read_GBps"] = out["input_bytes"] / out["read_seconds"] / 1e9
    out["write_GBps"] = out["output_bytes"] / out["write_seconds"] / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
