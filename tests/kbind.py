"""ctypes bindings used by the tests only: the oracle restatement (oracle/libkalign_oracle.so)
and the real reference behind oracle/_ref/libref_harness.so.  Never imported by the product."""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libkalign_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libkalign_ref.so")
REFH_SO = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")

NEGF = -np.finfo(np.float32).max


class KoJob(C.Structure):
    _fields_ = [("kind", C.c_int),
                ("seq1", C.c_void_p), ("seq2", C.c_void_p),
                ("prof1", C.c_void_p), ("prof2", C.c_void_p),
                ("len_a", C.c_int), ("len_b", C.c_int), ("sip", C.c_int),
                ("subm", C.c_void_p),
                ("gpo", C.c_float), ("gpe", C.c_float), ("tgpe", C.c_float),
                ("soff", C.c_float),
                ("bonus", C.c_void_p)]


class KoStats(C.Structure):
    _fields_ = [("cells", C.c_double), ("margin_sum", C.c_float), ("margin_count", C.c_int),
                ("top_score", C.c_float), ("n_boxes", C.c_int)]


_oracle = None
_refh = None
_ref = None


def have_oracle():
    return os.path.exists(ORACLE_SO)


def have_ref():
    return os.path.exists(REFH_SO) and os.path.exists(REF_SO)


def oracle():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(ORACLE_SO)
        lib.ko_align.argtypes = [C.POINTER(KoJob), i32p, C.POINTER(KoStats)]
        lib.ko_align.restype = C.c_int
        lib.ko_bpm_block.argtypes = [u8p, u8p, C.c_int, C.c_int]
        lib.ko_bpm_block.restype = C.c_int
        lib.ko_pair_distance.argtypes = [u8p, C.c_int, u8p, C.c_int]
        lib.ko_pair_distance.restype = C.c_float
        lib.ko_make_profile.argtypes = [u8p, C.c_int, f32p, C.c_float, C.c_float, C.c_float, C.c_float, f32p]
        lib.ko_make_profile.restype = None
        lib.ko_set_gap_penalties.argtypes = [f32p, C.c_int, C.c_int]
        lib.ko_set_gap_penalties.restype = None
        lib.ko_update.argtypes = [f32p, f32p, f32p, i32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
        lib.ko_update.restype = None
        lib.ko_code_path.argtypes = [i32p, C.c_int, C.c_int, C.c_int]
        lib.ko_code_path.restype = None
        lib.ko_posmap_from_path.argtypes = [i32p, C.c_int, i32p]
        lib.ko_posmap_from_path.restype = None
        _oracle = lib
    return _oracle


def ref():
    """The unmodified reference library (internal symbols are exported)."""
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO, mode=C.RTLD_GLOBAL)
        lib.bpm_block.argtypes = [u8p, u8p, C.c_int, C.c_int]
        lib.bpm_block.restype = C.c_int
        _ref = lib
    return _ref


def refh():
    global _refh
    if _refh is None:
        ref()
        lib = C.CDLL(REFH_SO)
        lib.refh_pair_align.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_int, C.c_int,
                                        f32p, C.c_float, C.c_float, C.c_float, C.c_float,
                                        C.c_void_p, C.c_int,
                                        i32p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]
        lib.refh_pair_align.restype = C.c_int
        lib.refh_code_path.argtypes = [i32p, C.c_int, C.c_int, C.c_int]
        lib.refh_make_profile.argtypes = [u8p, C.c_int, f32p, C.c_float, C.c_float, C.c_float, C.c_float, f32p]
        lib.refh_set_gap_penalties.argtypes = [f32p, C.c_int, C.c_int]
        lib.refh_update.argtypes = [f32p, f32p, f32p, i32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
        lib.refh_run_pipeline.argtypes = [C.POINTER(C.c_char_p), i32p, C.c_int, C.c_int, C.c_int,
                                          C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_int]
        lib.refh_run_pipeline.restype = C.c_void_p
        lib.refh_run_pipeline_ex.argtypes = [C.POINTER(C.c_char_p), i32p, C.c_int, C.c_int, C.c_int,
                                             C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_int,
                                             C.c_float, C.c_float]
        lib.refh_run_pipeline_ex.restype = C.c_void_p
        for name in ("refh_numseq", "refh_biotype", "refh_alnlen"):
            getattr(lib, name).argtypes = [C.c_void_p]
            getattr(lib, name).restype = C.c_int
        lib.refh_times.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(dtype=np.float64)]
        lib.refh_times.restype = None
        lib.refh_get_order.argtypes = [C.c_void_p, i32p, i32p]
        lib.refh_get_order.restype = None
        lib.refh_get_codes.argtypes = [C.c_void_p, C.c_int, u8p]
        lib.refh_get_codes.restype = None
        lib.refh_get_tasks.argtypes = [C.c_void_p, i32p]
        lib.refh_get_tasks.restype = C.c_int
        lib.refh_get_task_confidence.argtypes = [C.c_void_p, f32p]
        lib.refh_get_task_confidence.restype = None
        lib.refh_get_seq_distances.argtypes = [C.c_void_p, f32p]
        lib.refh_get_seq_distances.restype = None
        lib.refh_get_params.argtypes = [C.c_void_p, f32p, f32p]
        lib.refh_get_params.restype = None
        lib.refh_get_anchor_ids.argtypes = [C.c_void_p, i32p]
        lib.refh_get_anchor_ids.restype = C.c_int
        lib.refh_get_posmap.argtypes = [C.c_void_p, C.c_int, C.c_int, i32p]
        lib.refh_get_posmap.restype = C.c_int
        lib.refh_get_gaps.argtypes = [C.c_void_p, C.c_int, i32p]
        lib.refh_get_gaps.restype = None
        lib.refh_get_aligned.argtypes = [C.c_void_p, C.c_char_p]
        lib.refh_get_aligned.restype = C.c_int
        lib.refh_distance_matrix.argtypes = [C.POINTER(C.c_char_p), i32p, C.c_int, f32p, i32p, C.POINTER(C.c_int)]
        lib.refh_distance_matrix.restype = C.c_int
        lib.refh_free.argtypes = [C.c_void_p]
        lib.refh_free.restype = None
        lib.refh_aln_pairwise_dist.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_int, f32p]
        lib.refh_aln_pairwise_dist.restype = C.c_int
        lib.refh_read_input.argtypes = [C.c_char_p]
        lib.refh_read_input.restype = C.c_void_p
        for f in ("refh_msa_numseq", "refh_msa_biotype", "refh_msa_aligned"):
            getattr(lib, f).argtypes = [C.c_void_p]
            getattr(lib, f).restype = C.c_int
        for f in ("refh_msa_seq_len", "refh_msa_name_len"):
            getattr(lib, f).argtypes = [C.c_void_p, C.c_int]
            getattr(lib, f).restype = C.c_int
        lib.refh_msa_record.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, i32p]
        lib.refh_msa_record.restype = None
        lib.refh_msa_letter_freq.argtypes = [C.c_void_p, i32p]
        lib.refh_msa_letter_freq.restype = None
        lib.refh_msa_free.argtypes = [C.c_void_p]
        lib.refh_msa_free.restype = None
        lib.refh_write_rows.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_char_p, C.c_char_p]
        lib.refh_write_rows.restype = C.c_int
        lib.refh_tree_noise.argtypes = [C.c_uint64, C.c_float, C.c_longlong, f32p]
        lib.refh_tree_noise.restype = C.c_int
        lib.refh_run_seeded.argtypes = [C.POINTER(C.c_char_p), i32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                        C.c_uint64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_char_p, C.c_int]
        lib.refh_run_seeded.restype = C.c_int
        lib.refh_time_public_api.argtypes = [C.POINTER(C.c_char_p), i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
        lib.refh_time_public_api.restype = C.c_double
        _refh = lib
    return _refh


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def oracle_align(kind, len_a, len_b, subm, gpo, gpe, tgpe, soff=0.0, seq1=None, seq2=None,
                 prof1=None, prof2=None, sip=0, bonus=None):
    """returns (raw path [len_a+2], stats dict)"""
    job = KoJob()
    job.kind = kind
    job.seq1 = _ptr(seq1); job.seq2 = _ptr(seq2)
    job.prof1 = _ptr(prof1); job.prof2 = _ptr(prof2)
    job.len_a = len_a; job.len_b = len_b; job.sip = sip
    job.subm = _ptr(subm)
    job.gpo = gpo; job.gpe = gpe; job.tgpe = tgpe; job.soff = soff
    job.bonus = _ptr(bonus)
    path = np.full(len_a + 2, -7, dtype=np.int32)
    st = KoStats()
    rc = oracle().ko_align(C.byref(job), path, C.byref(st))
    assert rc == 0
    return path, dict(cells=st.cells, margin_sum=st.margin_sum, margin_count=st.margin_count,
                      top_score=st.top_score, n_boxes=st.n_boxes)


def ref_align(kind, len_a, len_b, subm, gpo, gpe, tgpe, soff=0.0, seq1=None, seq2=None,
              prof1=None, prof2=None, sip=0, bonus=None, score_only=False):
    path = np.full(len_a + 2, -7, dtype=np.int32)
    score = C.c_float(0); ms = C.c_float(0); mc = C.c_int(0)
    rc = refh().refh_pair_align(kind, _ptr(seq1), _ptr(seq2), _ptr(prof1), _ptr(prof2),
                                len_a, len_b, sip, np.ascontiguousarray(subm, dtype=np.float32),
                                gpo, gpe, tgpe, soff, _ptr(bonus), int(score_only),
                                path, C.byref(score), C.byref(ms), C.byref(mc))
    assert rc == 0
    return path, dict(score=score.value, margin_sum=ms.value, margin_count=mc.value)


class RefRun:
    """One run of the reference pipeline with access to its intermediate state."""

    def __init__(self, seqs, n_threads=1, type_=8, gpo=-1.0, gpe=-1.0, tgpe=-1.0,
                 consistency=0, weight=2.0, stop_after=0, dist_scale=0.0, use_seq_weights=0.0):
        lib = refh()
        n = len(seqs)
        arr = (C.c_char_p * n)(*[s.encode() if isinstance(s, str) else s for s in seqs])
        lens = np.array([len(s) for s in seqs], dtype=np.int32)
        self._h = lib.refh_run_pipeline_ex(arr, lens, n, n_threads, type_, gpo, gpe, tgpe,
                                           consistency, weight, stop_after, dist_scale, use_seq_weights)
        if not self._h:
            raise RuntimeError("reference pipeline failed")
        self.lib = lib
        self.n = lib.refh_numseq(self._h)
        self.rank = np.zeros(self.n, dtype=np.int32)
        self.lens = np.zeros(self.n, dtype=np.int32)
        lib.refh_get_order(self._h, self.rank, self.lens)

    def times(self):
        t = np.zeros(4, dtype=np.float64)
        self.lib.refh_times(self._h, t)
        return dict(dist_tree=t[0], anchor=t[1], tree_aln=t[2], total=t[3])

    def codes(self, i):
        out = np.zeros(int(self.lens[i]), dtype=np.uint8)
        self.lib.refh_get_codes(self._h, i, out)
        return out

    def tasks(self):
        out = np.zeros(3 * (self.n - 1), dtype=np.int32)
        nt = self.lib.refh_get_tasks(self._h, out)
        return out.reshape(-1, 3)[:nt].copy()

    def task_confidence(self):
        out = np.zeros(self.n - 1, dtype=np.float32)
        self.lib.refh_get_task_confidence(self._h, out)
        return out

    def seq_distances(self):
        out = np.zeros(self.n, dtype=np.float32)
        self.lib.refh_get_seq_distances(self._h, out)
        return out

    def params(self):
        subm = np.zeros(23 * 23, dtype=np.float32)
        gp = np.zeros(4, dtype=np.float32)
        self.lib.refh_get_params(self._h, subm, gp)
        return subm.reshape(23, 23), gp

    def biotype(self):
        return self.lib.refh_biotype(self._h)

    def anchor_ids(self):
        out = np.zeros(64, dtype=np.int32)
        k = self.lib.refh_get_anchor_ids(self._h, out)
        return out[:k].copy()

    def posmap(self, i, k):
        out = np.zeros(int(self.lens[i]), dtype=np.int32)
        assert self.lib.refh_get_posmap(self._h, i, k, out) == 0
        return out

    def gaps(self, i):
        out = np.zeros(int(self.lens[i]) + 1, dtype=np.int32)
        self.lib.refh_get_gaps(self._h, i, out)
        return out

    def aligned(self):
        L = self.lib.refh_alnlen(self._h)
        buf = C.create_string_buffer(self.n * (L + 1))
        self.lib.refh_get_aligned(self._h, buf)
        raw = buf.raw
        return [raw[i * (L + 1): i * (L + 1) + L].decode() for i in range(self.n)]

    def close(self):
        if self._h:
            self.lib.refh_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ref_distance_matrix(seqs):
    lib = refh()
    n = len(seqs)
    arr = (C.c_char_p * n)(*[s.encode() for s in seqs])
    lens = np.array([len(s) for s in seqs], dtype=np.int32)
    na = min(32, n)
    dm = np.zeros(n * na, dtype=np.float32)
    anchors = np.zeros(na, dtype=np.int32)
    nout = C.c_int(0)
    assert lib.refh_distance_matrix(arr, lens, n, dm, anchors, C.byref(nout)) == 0
    return dm.reshape(n, na), anchors


def ref_aln_pairwise_dist(rows):
    """the reference's compute_aln_pairwise_dist (lib/src/aln_apair_dist.c:9) on aligned strings"""
    lib = refh()
    n = len(rows)
    keep = [r.encode() for r in rows]
    arr = (C.c_char_p * n)(*keep)
    dm = np.zeros(n * n, dtype=np.float32)
    assert lib.refh_aln_pairwise_dist(arr, n, len(rows[0]), dm) == 0
    return dm.reshape(n, n)


def oracle_aln_pairwise_dist(rows):
    """numpy restatement of pairwise_identity_dist (lib/src/aln_apair_dist.c:62-82): columns where both
    rows hold a residue are counted, equal characters among them are matches, d = 1 - m / a in float32
    (1 when a == 0), 0 on the diagonal (aln_apair_dist.c:25).  TEST INFRASTRUCTURE ONLY."""
    a = np.frombuffer("".join(rows).encode(), dtype=np.uint8).reshape(len(rows), -1)
    res = a != ord("-")
    n = len(rows)
    dm = np.zeros((n, n), dtype=np.float32)
    for i in range(n):
        both = res[i][None, :] & res
        al = both.sum(axis=1).astype(np.int64)
        m = (both & (a == a[i][None, :])).sum(axis=1).astype(np.int64)
        with np.errstate(divide="ignore", invalid="ignore"):
            d = np.float32(1.0) - m.astype(np.float32) / al.astype(np.float32)
        d = np.where(al == 0, np.float32(1.0), d).astype(np.float32)
        d[i] = 0.0
        dm[i] = d
    return dm


def ref_read_fasta(path):
    """the reference's kalign_read_input (lib/src/msa_io.c:80) on a file:
    (records [(name bytes, residues bytes, gaps int32[len + 1])], letter_freq int32[128]) or None when it fails"""
    lib = refh()
    h = lib.refh_read_input(os.fsencode(path))
    if not h:
        return None
    try:
        recs = []
        for i in range(lib.refh_msa_numseq(h)):
            ln = lib.refh_msa_seq_len(h, i)
            name = C.create_string_buffer(lib.refh_msa_name_len(h, i) + 1)
            seq = C.create_string_buffer(ln + 1)
            gaps = np.zeros(ln + 1, dtype=np.int32)
            lib.refh_msa_record(h, i, name, seq, gaps)
            recs.append((name.value, seq.raw[:ln], gaps))
        freq = np.zeros(128, dtype=np.int32)
        lib.refh_msa_letter_freq(h, freq)
        return recs, freq
    finally:
        lib.refh_msa_free(h)


def ref_write_fasta(path, names, rows):
    """the reference's kalign_write_msa(..., "fasta") (lib/src/msa_io.c:193,668) on names and finished rows"""
    lib = refh()
    n = len(rows)
    nm = (C.c_char_p * max(1, n))(*[x if isinstance(x, bytes) else x.encode() for x in names])
    rw = (C.c_char_p * max(1, n))(*[x if isinstance(x, bytes) else x.encode() for x in rows])
    assert lib.refh_write_rows(nm, rw, n, len(rows[0]) if n else 0, os.fsencode(path), b"fasta") == 0


_ALPHA = set(range(ord("A"), ord("Z") + 1)) | set(range(ord("a"), ord("z") + 1))
_PUNCT = set(range(33, 48)) | set(range(58, 65)) | set(range(91, 97)) | set(range(123, 127))


def oracle_read_fasta(data):
    """pure-python restatement of read_file_stdin + read_fasta (lib/src/msa_io.c:348-482) on the bytes of a
    file: lines split at newline, a line's content ends at its first control character (< 32 or 127), '>' as the
    first character opens a record, letters are residues, punctuation is counted in gaps[len], all characters
    of sequence lines are counted in letter_freq.  Returns (records, letter_freq) or None where the
    reference fails (sequence data before the first header).  TEST INFRASTRUCTURE ONLY."""
    recs = []
    freq = np.zeros(128, dtype=np.int32)
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()                      # getline yields no line after a final newline
    cur = None
    for raw in lines:
        cut = len(raw)
        for i, ch in enumerate(raw):
            if ch < 32 or ch == 127:
                cut = i
                break
        line = raw[:cut]
        if line[:1] == b">":
            cur = [bytes(line[1:]), bytearray(), [0]]
            recs.append(cur)
            continue
        for ch in line:
            if ch < 128:
                freq[ch] += 1
            if ch in _ALPHA:
                if cur is None:
                    return None
                cur[1].append(ch)
                cur[2].append(0)
            elif ch in _PUNCT:
                if cur is None:
                    return None
                cur[2][-1] += 1
    return [(r[0], bytes(r[1]), np.array(r[2], dtype=np.int32)) for r in recs], freq


def oracle_write_fasta(names, rows):
    """restatement of write_msa_fasta (lib/src/msa_io.c:668-717): the bytes of the output file"""
    out = bytearray()
    for nm, r in zip(names, rows):
        nm = nm if isinstance(nm, bytes) else nm.encode()
        r = r if isinstance(r, bytes) else r.encode()
        out += b">" + nm + b"\n"
        for j in range(0, len(r), 60):
            out += r[j:j + 60] + b"\n"
    return bytes(out)


def ref_tree_noise(seed, sigma, n):
    """noise factors of build_tree_kmeans_noisy from the reference's generator (lib/src/tlrng.c)"""
    out = np.zeros(n, dtype=np.float32)
    assert refh().refh_tree_noise(seed, sigma, n, out) == 0
    return out


def ref_run_seeded(seqs, n_threads=2, type_=8, gpo=-1.0, gpe=-1.0, tgpe=-1.0, tree_seed=0, tree_noise=0.0, dist_scale=0.0,
                   vsm_amax=-1.0, use_seq_weights=-1.0, consistency=0, weight=2.0):
    """the reference's kalign_run_seeded (lib/src/aln_wrap.c:133) on strings named s0..; aligned rows in input order"""
    lib = refh()
    n = len(seqs)
    keep = [s.encode() for s in seqs]
    arr = (C.c_char_p * n)(*keep)
    lens = np.array([len(s) for s in keep], dtype=np.int32)
    cap = 8 * max(len(s) for s in seqs) + 64
    buf = C.create_string_buffer(n * (cap + 1))
    L = lib.refh_run_seeded(arr, lens, n, n_threads, type_, gpo, gpe, tgpe, tree_seed, tree_noise, dist_scale, vsm_amax,
                            use_seq_weights, consistency, weight, buf, cap)
    assert L >= 0
    raw = buf.raw
    return [raw[i * (L + 1):i * (L + 1) + L].decode() for i in range(n)]


_refens = None


def ref_resolve_run_params(base_gpo, base_gpe, base_tgpe, k, seed):
    """the reference's static resolve_run_params (lib/src/ensemble.c:55), through oracle/ref_ensemble_params.c"""
    global _refens
    if _refens is None:
        _refens = C.CDLL(os.path.join(os.path.dirname(REF_SO), "libref_ensemble.so"))
        _refens.refh_resolve_run_params.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, C.c_uint64, C.POINTER(C.c_float),
                                                    C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.POINTER(C.c_float)]
    g, e, t, nz = C.c_float(), C.c_float(), C.c_float(), C.c_float()
    rs = C.c_uint64()
    _refens.refh_resolve_run_params(base_gpo, base_gpe, base_tgpe, k, seed, C.byref(g), C.byref(e), C.byref(t), C.byref(rs), C.byref(nz))
    return g.value, e.value, t.value, rs.value, nz.value
