"""Seeded synthetic sequence families: a root of length L drawn i.i.d.
uniform over the alphabet; every member is the root with per-site substitution / insertion /
deletion (star phylogeny).  Names are s0..s{N-1} (the reference sorts by (len desc, name asc))."""
import numpy as np

PROTEIN = "ACDEFGHIKLMNPQRSTVWY"
DNA = "ACGT"
RNA = "ACGU"

# BASELINE.json configs: name -> (N, L, alphabet, seed, sub, indel)
CONFIGS = {
    "C1": (16, 200, PROTEIN, 1, 0.15, 0.015),
    "C2": (1000, 400, PROTEIN, 2, 0.15, 0.015),
    "C3": (10000, 1500, RNA, 3, 0.15, 0.015),
    "C4": (100000, 300, PROTEIN, 4, 0.15, 0.015),
    "C5": (1000, 30000, DNA, 5, 0.01, 0.001),
}


def family(n, length, alphabet=PROTEIN, seed=1, sub=0.15, ins=0.015, dele=0.015):
    """returns a list of n python strings"""
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(alphabet.encode(), dtype=np.uint8)
    A = len(alpha)
    root = rng.integers(0, A, size=length)
    out = []
    for _ in range(n):
        r = rng.random(length)
        keep = r >= dele
        s = root.copy()
        m = rng.random(length) < sub
        s[m] = rng.integers(0, A, size=int(m.sum()))
        insm = rng.random(length) < ins
        # build: for each kept site emit it, and after sites flagged for insertion emit a random letter
        n_ins = int(insm.sum())
        ins_letters = rng.integers(0, A, size=n_ins)
        pieces = np.empty(length * 2, dtype=np.int64)
        valid = np.zeros(length * 2, dtype=bool)
        pieces[0::2] = s
        valid[0::2] = keep
        tmp = np.zeros(length, dtype=np.int64)
        tmp[insm] = ins_letters
        pieces[1::2] = tmp
        valid[1::2] = insm
        seq = alpha[pieces[valid]]
        if seq.size == 0:
            seq = alpha[root[:1]]
        out.append(seq.tobytes().decode())
    return out


def config(name, n=None):
    N, L, alphabet, seed, sub, indel = CONFIGS[name]
    if n is not None:
        N = n
    return family(N, L, alphabet, seed, sub, indel, indel)


def write_fasta(path, seqs):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">s%d\n" % i)
            for k in range(0, len(s), 60):
                f.write(s[k:k + 60] + "\n")


def msa_sha256(rows):
    """SHA-256 of an alignment: the rows in input order joined by a newline.  bench.py prints it in
    every line and the full-size parity tests compare it with tests/golden/full_*.npz, which
    tools/gen_golden_full.py wrote from the unmodified reference."""
    import hashlib
    h = hashlib.sha256()
    first = True
    for r in rows:
        if not first:
            h.update(b"\n")
        first = False
        h.update(r.encode() if isinstance(r, str) else r)
    return h.hexdigest()
