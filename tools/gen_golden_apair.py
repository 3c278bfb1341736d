"""Golden vectors of compute_aln_pairwise_dist (lib/src/aln_apair_dist.c:9) from the UNMODIFIED reference
(oracle/_ref, through oracle/ref_harness.c refh_aln_pairwise_dist) -> tests/golden/apair.npz.
Run once in the build container:  python tools/gen_golden_apair.py"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kbind  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def random_rows(rng, n, L, alphabet, gap_p):
    rows = []
    for _ in range(n):
        r = rng.choice(list(alphabet), size=L)
        g = rng.random(L) < gap_p
        r = np.where(g, "-", r)
        rows.append("".join(r))
    return rows


def main():
    rng = np.random.default_rng(2024)
    cases = {}
    # the aligned rows of the committed reference alignments
    for f in sorted(glob.glob(os.path.join(OUT, "msa_*.npz"))):
        z = np.load(f)
        cases["msa_" + os.path.basename(f)[4:-4]] = [str(s) for s in z["aligned"]]
    # lengths around the 4-column word, the 128-column chunk and the 64-row tile of the kernel
    for n, L in ((2, 1), (3, 3), (5, 127), (64, 128), (65, 129), (129, 513), (7, 1031)):
        cases["rnd_%dx%d" % (n, L)] = random_rows(rng, n, L, "ACGTN", 0.3)
    # case matters ('a' != 'A'), an all-gap row (distance 1 to everything), two rows without a common column
    rows = random_rows(rng, 20, 200, "ACDEFGHIKLMNPQRSTVWYacdx", 0.4)
    rows[3] = "-" * 200
    rows[7] = "A" * 100 + "-" * 100
    rows[8] = "-" * 100 + "A" * 100
    cases["special"] = rows
    rec = {"names": np.array(sorted(cases))}
    for k, rows in cases.items():
        rec["rows_" + k] = np.array(rows)
        rec["dm_" + k] = kbind.ref_aln_pairwise_dist(rows)
        assert np.array_equal(rec["dm_" + k], kbind.oracle_aln_pairwise_dist(rows)), k
    np.savez_compressed(os.path.join(OUT, "apair.npz"), **rec)
    print("apair:", len(cases), "cases")


if __name__ == "__main__":
    main()
