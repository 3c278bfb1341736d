"""ctypes binding of the product C-ABI (include/kalign_b200.h -> kalign_b200/libkalign_b200.so).

The extension is the ONLY implementation: if it is missing or no CUDA device is usable every call
raises -- there is no CPU fallback."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libkalign_b200.so")

KIND_SS, KIND_SP, KIND_PP = 0, 1, 2

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


class Params(C.Structure):
    _fields_ = [("subm", C.c_float * (23 * 23)),
                ("gpo", C.c_float), ("gpe", C.c_float), ("tgpe", C.c_float),
                ("vsm_amax", C.c_float), ("nalpha", C.c_int),
                ("dist_scale", C.c_float), ("use_seq_weights", C.c_float)]


class Pair(C.Structure):
    _fields_ = [("kind", C.c_int),
                ("seq_rows", C.c_void_p), ("seq_cols", C.c_void_p),
                ("prof_rows", C.c_void_p), ("prof_cols", C.c_void_p),
                ("len_a", C.c_int), ("len_b", C.c_int), ("sip", C.c_int), ("soff", C.c_float),
                ("bonus", C.c_void_p), ("path_out", C.c_void_p), ("score_out", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("dp_cells", C.c_double), ("dp_seconds", C.c_double), ("sweep_seconds", C.c_double),
                ("n_boxes", C.c_longlong), ("n_launches", C.c_longlong),
                ("bpm_seconds", C.c_double), ("bpm_pairs", C.c_double),
                ("h2d_bytes", C.c_double), ("d2h_bytes", C.c_double),
                ("cells_ss", C.c_double), ("cells_sp", C.c_double), ("cells_pp", C.c_double),
                ("cells_bonus", C.c_double), ("align_seconds", C.c_double), ("small_seconds", C.c_double),
                ("small_ss", C.c_double), ("small_sp", C.c_double), ("small_pp", C.c_double),
                ("n_collectives", C.c_double), ("collective_bytes", C.c_double),
                ("apair_seconds", C.c_double), ("apair_col_pairs", C.c_double)]


# every symbol include/kalign_b200.h declares
EXPORTS = ["kb200_device_count", "kb200_ctx_create", "kb200_ctx_destroy", "kb200_get_stats",
           "kb200_version", "kb200_params_init", "kb200_pair_align_batch", "kb200_distances",
           "kb200_seqs_upload", "kb200_distances_on", "kb200_seqs_free", "kb200_aln_pairwise_dist",
           "kb200_anchor_posmaps", "kb200_select_anchors", "kb200_align_tree", "kb200_align_tree_conf", "kb200_kalign",
           "kb200_msa_create", "kb200_msa_align", "kb200_msa_result", "kb200_msa_info", "kb200_msa_tree", "kb200_msa_free",
           "kb200_alphabet", "kb200_guide_tree", "kb200_tasks_creation_order", "kb200_kalign_seeded", "kb200_tree_noise", "kb200_ensemble_run_params", "kb200_ensemble_run",
           "kb200_fasta_read", "kb200_fasta_numseq", "kb200_fasta_get", "kb200_fasta_letter_freq", "kb200_fasta_arrays",
           "kb200_fasta_free", "kb200_fasta_write", "kb200_kalign_file",
           "kb200_comm_unique_id", "kb200_ctx_comm_init", "kb200_ctx_comm_destroy", "kb200_partition"]

_lib = None


def _preload_nccl():
    """The library links libnccl.so.2.  If PyTorch is (or will be) imported in this process its bundled
    NCCL must be the one that gets loaded -- an older system libnccl loaded first breaks
    `import torch` (undefined symbol in libtorch_cuda).  So load torch's copy first when it exists."""
    import glob
    import sys
    for base in sys.path:
        for cand in glob.glob(os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so.2")):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                return cand
            except OSError:
                pass
    return None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError("kalign_b200: %s is missing (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                           "there is no CPU fallback" % SO_PATH)
    _preload_nccl()
    lib = C.CDLL(SO_PATH)
    lib.kb200_device_count.restype = C.c_int
    lib.kb200_version.restype = C.c_char_p
    lib.kb200_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.kb200_ctx_create.restype = C.c_int
    lib.kb200_ctx_destroy.argtypes = [C.c_void_p]
    lib.kb200_ctx_destroy.restype = None
    lib.kb200_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    lib.kb200_get_stats.restype = C.c_int
    lib.kb200_params_init.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
    lib.kb200_params_init.restype = C.c_int
    lib.kb200_pair_align_batch.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Pair), C.c_int]
    lib.kb200_pair_align_batch.restype = C.c_int
    lib.kb200_distances.argtypes = [C.c_void_p, u8p, i64p, i32p, C.c_int, i32p, C.c_int, i32p, C.c_int, f32p]
    lib.kb200_distances.restype = C.c_int
    lib.kb200_seqs_upload.argtypes = [C.c_void_p, u8p, i64p, i32p, C.c_int, C.POINTER(C.c_void_p)]
    lib.kb200_seqs_upload.restype = C.c_int
    lib.kb200_distances_on.argtypes = [C.c_void_p, i32p, C.c_int, i32p, C.c_int, C.c_int, f32p]
    lib.kb200_distances_on.restype = C.c_int
    lib.kb200_seqs_free.argtypes = [C.c_void_p]
    lib.kb200_seqs_free.restype = None
    lib.kb200_aln_pairwise_dist.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.kb200_aln_pairwise_dist.restype = C.c_int
    lib.kb200_kalign_seeded.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), i32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                        C.c_ulonglong, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float,
                                        C.POINTER(C.POINTER(C.c_void_p)), C.POINTER(C.c_int)]
    lib.kb200_kalign_seeded.restype = C.c_int
    lib.kb200_alphabet.argtypes = [C.c_int, np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS"), C.POINTER(C.c_int)]
    lib.kb200_alphabet.restype = C.c_int
    lib.kb200_guide_tree.argtypes = [C.c_void_p, u8p, i64p, i32p, C.c_int, C.c_int, C.c_ulonglong, C.c_float, i32p, f32p]
    lib.kb200_guide_tree.restype = C.c_int
    lib.kb200_tasks_creation_order.argtypes = [i32p, C.c_int, C.c_int, i32p]
    lib.kb200_tasks_creation_order.restype = C.c_int
    lib.kb200_tree_noise.argtypes = [C.c_ulonglong, C.c_float, C.c_longlong, f32p]
    lib.kb200_tree_noise.restype = C.c_int
    lib.kb200_ensemble_run_params.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, C.c_ulonglong, C.POINTER(C.c_float),
                                              C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_ulonglong), C.POINTER(C.c_float)]
    lib.kb200_ensemble_run_params.restype = C.c_int
    lib.kb200_ensemble_run.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), i32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                       C.c_int, C.c_ulonglong, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float,
                                       C.POINTER(C.POINTER(C.c_void_p)), C.POINTER(C.c_int)]
    lib.kb200_ensemble_run.restype = C.c_int
    lib.kb200_fasta_read.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.kb200_fasta_read.restype = C.c_int
    lib.kb200_fasta_numseq.argtypes = [C.c_void_p]
    lib.kb200_fasta_numseq.restype = C.c_int
    lib.kb200_fasta_get.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int),
                                    C.POINTER(C.c_void_p)]
    lib.kb200_fasta_get.restype = C.c_int
    lib.kb200_fasta_letter_freq.argtypes = [C.c_void_p]
    lib.kb200_fasta_letter_freq.restype = C.POINTER(C.c_int)
    lib.kb200_fasta_arrays.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.kb200_fasta_arrays.restype = C.c_int
    lib.kb200_fasta_free.argtypes = [C.c_void_p]
    lib.kb200_fasta_free.restype = None
    lib.kb200_fasta_write.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int]
    lib.kb200_fasta_write.restype = C.c_int
    lib.kb200_kalign_file.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                      C.c_int, C.c_float]
    lib.kb200_kalign_file.restype = C.c_int
    lib.kb200_anchor_posmaps.argtypes = [C.c_void_p, C.POINTER(Params), u8p, i64p, i32p, C.c_int,
                                         i32p, C.c_int, C.c_longlong, C.c_longlong, i32p]
    lib.kb200_anchor_posmaps.restype = C.c_int
    lib.kb200_select_anchors.argtypes = [f32p, C.c_int, C.c_int, i32p]
    lib.kb200_select_anchors.restype = C.c_int
    lib.kb200_align_tree.argtypes = [C.c_void_p, C.POINTER(Params), u8p, i64p, i32p, C.c_int,
                                     i32p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float, i32p]
    lib.kb200_align_tree.restype = C.c_int
    lib.kb200_align_tree_conf.argtypes = [C.c_void_p, C.POINTER(Params), u8p, i64p, i32p, C.c_int,
                                          i32p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float, i32p, C.c_void_p, C.c_void_p]
    lib.kb200_align_tree_conf.restype = C.c_int
    lib.kb200_kalign.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), i32p, C.c_int, C.c_int, C.c_int,
                                 C.c_float, C.c_float, C.c_float, C.c_int, C.c_float,
                                 C.POINTER(C.POINTER(C.c_void_p)), C.POINTER(C.c_int)]
    lib.kb200_kalign.restype = C.c_int
    lib.kb200_msa_create.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), i32p, C.c_int, C.c_int, C.c_int,
                                     C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.POINTER(C.c_void_p)]
    lib.kb200_msa_create.restype = C.c_int
    lib.kb200_msa_align.argtypes = [C.c_void_p]
    lib.kb200_msa_align.restype = C.c_int
    lib.kb200_msa_result.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_void_p)), C.POINTER(C.c_int)]
    lib.kb200_msa_result.restype = C.c_int
    lib.kb200_msa_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.kb200_msa_info.restype = C.c_int
    lib.kb200_msa_tree.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.kb200_msa_tree.restype = C.c_int
    lib.kb200_msa_free.argtypes = [C.c_void_p]
    lib.kb200_msa_free.restype = None
    lib.kb200_comm_unique_id.argtypes = [C.c_void_p, C.c_int]
    lib.kb200_comm_unique_id.restype = C.c_int
    lib.kb200_ctx_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.kb200_ctx_comm_init.restype = C.c_int
    lib.kb200_ctx_comm_destroy.argtypes = [C.c_void_p]
    lib.kb200_ctx_comm_destroy.restype = None
    lib.kb200_partition.argtypes = [np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS"), C.c_int, C.c_int, i32p]
    lib.kb200_partition.restype = C.c_int
    _lib = lib
    return lib


def make_params(biotype, type_=8, gpo=-1.0, gpe=-1.0, tgpe=-1.0):
    p = Params()
    if load().kb200_params_init(C.byref(p), biotype, type_, gpo, gpe, tgpe) != 0:
        raise RuntimeError("kb200_params_init failed")
    return p


def params_from(subm, gpo, gpe, tgpe, nalpha=23, vsm_amax=0.0, dist_scale=0.0, use_seq_weights=0.0):
    p = Params()
    flat = np.ascontiguousarray(subm, dtype=np.float32).reshape(-1)
    for i in range(23 * 23):
        p.subm[i] = float(flat[i])
    p.gpo, p.gpe, p.tgpe = float(gpo), float(gpe), float(tgpe)
    p.nalpha = nalpha
    p.vsm_amax = vsm_amax
    p.dist_scale = dist_scale
    p.use_seq_weights = use_seq_weights
    return p


class Context:
    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        if self.lib.kb200_ctx_create(device, C.byref(h)) != 0:
            raise RuntimeError("kalign_b200: no usable CUDA device %d (no CPU fallback)" % device)
        self.h = h

    def close(self):
        if self.h:
            self.lib.kb200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self):
        s = Stats()
        self.lib.kb200_get_stats(self.h, C.byref(s))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def pair_align_batch(self, prm, jobs):
        """jobs: list of dicts(kind, len_a, len_b, seq_rows, seq_cols, prof_rows, prof_cols, sip, soff, bonus).
        returns (list of raw paths, np.array of top scores)"""
        n = len(jobs)
        arr = (Pair * n)()
        paths, keep = [], []
        scores = np.zeros(n, dtype=np.float32)
        for i, j in enumerate(jobs):
            p = arr[i]
            p.kind = j["kind"]
            for name in ("seq_rows", "seq_cols", "prof_rows", "prof_cols", "bonus"):
                a = j.get(name)
                if a is not None:
                    a = np.ascontiguousarray(a)
                    keep.append(a)
                    setattr(p, name, a.ctypes.data)
            p.len_a = j["len_a"]
            p.len_b = j["len_b"]
            p.sip = j.get("sip", 0)
            p.soff = j.get("soff", 0.0)
            path = np.full(j["len_a"] + 2, -7, dtype=np.int32)
            paths.append(path)
            p.path_out = path.ctypes.data
            p.score_out = scores.ctypes.data + 4 * i
        rc = self.lib.kb200_pair_align_batch(self.h, C.byref(prm), arr, n)
        if rc != 0:
            raise RuntimeError("kb200_pair_align_batch failed")
        return paths, scores


def pack(seqs_codes):
    """list of uint8 arrays -> (concatenated codes, offs int64, lens int32)"""
    lens = np.array([len(s) for s in seqs_codes], dtype=np.int32)
    offs = np.zeros(len(seqs_codes), dtype=np.int64)
    if len(seqs_codes) > 1:
        offs[1:] = np.cumsum(lens[:-1], dtype=np.int64)
    flat = np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqs_codes]) if len(seqs_codes) else np.zeros(0, np.uint8)
    return np.ascontiguousarray(flat), offs, lens


def _distances(self, flat, offs, lens, rows, cols):
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    dm = np.zeros(len(rows) * len(cols), dtype=np.float32)
    if self.lib.kb200_distances(self.h, flat, offs, lens, len(lens), rows, len(rows), cols, len(cols), dm) != 0:
        raise RuntimeError("kb200_distances failed")
    return dm.reshape(len(rows), len(cols))


def _aln_pairwise_dist(self, rows):
    """kb200_aln_pairwise_dist: rows = equally long aligned strings ('-' = gap) -> (n, n) float32"""
    n = len(rows)
    alnlen = len(rows[0]) if n else 0
    if any(len(r) != alnlen for r in rows):
        raise ValueError("aligned rows must have one length")
    enc = [r.encode("ascii") if isinstance(r, str) else bytes(r) for r in rows]
    arr = (C.c_char_p * n)(*enc)
    dm = np.full((n, n), np.nan, dtype=np.float32)
    ptrs = (C.c_void_p * n)(*[dm.ctypes.data + i * n * 4 for i in range(n)])
    if self.lib.kb200_aln_pairwise_dist(self.h, arr, n, alnlen, ptrs) != 0:
        raise RuntimeError("kb200_aln_pairwise_dist failed")
    return dm


def alphabet(letters):
    """kb200_alphabet -> (to_internal int8[128], L)"""
    t = np.zeros(128, dtype=np.int8)
    L = C.c_int(0)
    if load().kb200_alphabet(letters, t, C.byref(L)) != 0:
        raise RuntimeError("kb200_alphabet failed")
    return t, L.value


def tasks_creation_order(tasks_sorted, nseq):
    """kb200_tasks_creation_order: a task list sorted by c -> the order create_tasks filled it in"""
    t = np.ascontiguousarray(tasks_sorted, dtype=np.int32).reshape(-1, 3)
    out = np.zeros_like(t)
    if load().kb200_tasks_creation_order(t.reshape(-1), len(t), nseq, out.reshape(-1)) != 0:
        raise RuntimeError("kb200_tasks_creation_order failed")
    return out


def _guide_tree(self, flat, offs, lens, n_threads=2, tree_seed=0, tree_noise=0.0):
    """kb200_guide_tree on tree-alphabet codes -> (tasks (n-1) x 3 in creation order, seq_distances)"""
    n = len(lens)
    abc = np.zeros(3 * (n - 1), dtype=np.int32)
    sd = np.zeros(n, dtype=np.float32)
    if self.lib.kb200_guide_tree(self.h, flat, offs, lens, n, n_threads, tree_seed, tree_noise, abc, sd) != 0:
        raise RuntimeError("kb200_guide_tree failed")
    return abc.reshape(-1, 3), sd


def tree_noise(seed, sigma, n):
    """kb200_tree_noise: the factors build_tree_kmeans_noisy multiplies the anchor distances with (host code)"""
    out = np.zeros(n, dtype=np.float32)
    if load().kb200_tree_noise(seed, sigma, n, out) != 0:
        raise RuntimeError("kb200_tree_noise failed")
    return out


def ensemble_run_params(base_gpo, base_gpe, base_tgpe, run, seed):
    """kb200_ensemble_run_params -> (gpo, gpe, tgpe, tree_seed, tree_noise) of run `run`"""
    g, e, t, n = C.c_float(), C.c_float(), C.c_float(), C.c_float()
    ts = C.c_ulonglong()
    if load().kb200_ensemble_run_params(base_gpo, base_gpe, base_tgpe, run, seed, C.byref(g), C.byref(e), C.byref(t), C.byref(ts), C.byref(n)) != 0:
        raise RuntimeError("kb200_ensemble_run_params failed")
    return g.value, e.value, t.value, ts.value, n.value


def _rows_out(lib_, out, alen, nrows):
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    rows = []
    for i in range(nrows):
        rows.append(C.string_at(out[i], alen.value).decode())
        libc.free(out[i])
    libc.free(C.cast(out, C.c_void_p))
    return rows


def _kalign_seeded(self, seqs, n_threads=1, type_=8, gpo=-1.0, gpe=-1.0, tgpe=-1.0, tree_seed=0, tree_noise=0.0,
                   dist_scale=0.0, vsm_amax=-1.0, use_seq_weights=-1.0, consistency=0, weight=2.0):
    """kb200_kalign_seeded: kalign_run_seeded on python strings"""
    n = len(seqs)
    keep = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    arr = (C.c_char_p * n)(*keep)
    lens = np.array([len(s) for s in keep], dtype=np.int32)
    out = C.POINTER(C.c_void_p)()
    alen = C.c_int(0)
    if self.lib.kb200_kalign_seeded(self.h, arr, lens, n, n_threads, type_, gpo, gpe, tgpe, tree_seed, tree_noise, dist_scale,
                                    vsm_amax, use_seq_weights, consistency, weight, C.byref(out), C.byref(alen)) != 0:
        raise RuntimeError("kb200_kalign_seeded failed")
    return _rows_out(self.lib, out, alen, int((lens > 0).sum()))


def _ensemble_runs(self, seqs, n_runs, seed=42, rank=0, world=1, n_threads=1, type_=8, gpo=-1.0, gpe=-1.0, tgpe=-1.0,
                   dist_scale=0.0, vsm_amax=-1.0, use_seq_weights=-1.0, consistency=0, weight=2.0):
    """the runs k of a kalign_ensemble with k % world == rank (one process per GPU, no collective): {k: rows}"""
    n = len(seqs)
    keep = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    arr = (C.c_char_p * n)(*keep)
    lens = np.array([len(s) for s in keep], dtype=np.int32)
    res = {}
    for k in range(rank, n_runs, world):
        out = C.POINTER(C.c_void_p)()
        alen = C.c_int(0)
        if self.lib.kb200_ensemble_run(self.h, arr, lens, n, n_threads, type_, gpo, gpe, tgpe, k, seed, dist_scale, vsm_amax,
                                       use_seq_weights, consistency, weight, C.byref(out), C.byref(alen)) != 0:
            raise RuntimeError("kb200_ensemble_run %d failed" % k)
        res[k] = _rows_out(self.lib, out, alen, int((lens > 0).sum()))
    return res


class Fasta:
    """kb200_fasta_read: records of a FASTA file parsed with the semantics of the reference's read_fasta
    (lib/src/msa_io.c:412).  Host code -- needs no GPU."""

    def __init__(self, path, n_threads=0):
        self.lib = load()
        h = C.c_void_p()
        if self.lib.kb200_fasta_read(os.fsencode(path), n_threads, C.byref(h)) != 0:
            raise RuntimeError("kb200_fasta_read failed: %s" % path)
        self.h = h
        self.n = self.lib.kb200_fasta_numseq(h)

    def record(self, i):
        """(name bytes, residues bytes, gaps int32[len + 1])"""
        name = C.c_char_p()
        seq = C.c_void_p()
        ln = C.c_int()
        gaps = C.c_void_p()
        if self.lib.kb200_fasta_get(self.h, i, C.byref(name), C.byref(seq), C.byref(ln), C.byref(gaps)) != 0:
            raise IndexError(i)
        s = C.string_at(seq.value, ln.value)
        g = np.ctypeslib.as_array(C.cast(gaps.value, C.POINTER(C.c_int)), shape=(ln.value + 1,)).copy()
        return name.value, s, g

    def records(self):
        return [self.record(i) for i in range(self.n)]

    def letter_freq(self):
        return np.ctypeslib.as_array(self.lib.kb200_fasta_letter_freq(self.h), shape=(128,)).copy()

    def close(self):
        if self.h:
            self.lib.kb200_fasta_free(self.h)
            self.h = None


def fasta_write(path, names, rows, n_threads=0):
    """kb200_fasta_write: what write_msa_fasta (lib/src/msa_io.c:668) writes for these names and rows"""
    lib = load()
    n = len(rows)
    alnlen = len(rows[0]) if n else 0
    nm = (C.c_char_p * n)(*[x if isinstance(x, bytes) else x.encode() for x in names])
    rw = (C.c_char_p * n)(*[x if isinstance(x, bytes) else x.encode() for x in rows])
    if lib.kb200_fasta_write(os.fsencode(path), nm, rw, n, alnlen, n_threads) != 0:
        raise RuntimeError("kb200_fasta_write failed: %s" % path)


def _kalign_file(self, infile, outfile, n_threads=1, type_=8, gpo=-1.0, gpe=-1.0, tgpe=-1.0, consistency=0, weight=2.0):
    if self.lib.kb200_kalign_file(self.h, os.fsencode(infile), os.fsencode(outfile), n_threads, type_, gpo, gpe, tgpe,
                                  consistency, weight) != 0:
        raise RuntimeError("kb200_kalign_file failed")


class DeviceSeqs:
    """sequences resident on the device between distance calls (kb200_seqs_upload / kb200_distances_on)"""

    def __init__(self, ctx, flat, offs, lens):
        self.lib = ctx.lib
        h = C.c_void_p()
        if self.lib.kb200_seqs_upload(ctx.h, flat, offs, lens, len(lens), C.byref(h)) != 0:
            raise RuntimeError("kb200_seqs_upload failed")
        self.h = h

    def distances(self, rows, cols):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        dm = np.zeros(len(rows) * len(cols), dtype=np.float32)
        if self.lib.kb200_distances_on(self.h, rows, len(rows), cols, len(cols), 0, dm) != 0:
            raise RuntimeError("kb200_distances_on failed")
        return dm.reshape(len(rows), len(cols))

    def pair_distances(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.int32)
        b = np.ascontiguousarray(b, dtype=np.int32)
        dm = np.zeros(len(a), dtype=np.float32)
        if self.lib.kb200_distances_on(self.h, a, len(a), b, 0, 1, dm) != 0:
            raise RuntimeError("kb200_distances_on failed")
        return dm

    def close(self):
        if self.h:
            self.lib.kb200_seqs_free(self.h)
            self.h = None


def _anchor_posmaps(self, prm, flat, offs, lens, anchor_ids, begin=0, end=None, out=None):
    anchor_ids = np.ascontiguousarray(anchor_ids, dtype=np.int32)
    K = len(anchor_ids)
    n = len(lens)
    if end is None:
        end = n * K
    if out is None:
        out = np.full(int(lens.sum()) * K, -9, dtype=np.int32)
    if self.lib.kb200_anchor_posmaps(self.h, C.byref(prm), flat, offs, lens, n, anchor_ids, K, begin, end, out) != 0:
        raise RuntimeError("kb200_anchor_posmaps failed")
    return out


def posmap_view(posmaps, offs, lens, K, i, k):
    o = K * int(offs[i]) + k * int(lens[i])
    return posmaps[o:o + int(lens[i])]


def _align_tree(self, prm, flat, offs, lens, tasks, seq_distances=None, posmaps=None, K=0, weight=2.0, confidence=False):
    tasks = np.ascontiguousarray(tasks, dtype=np.int32).reshape(-1)
    n = len(lens)
    gaps = np.zeros(int(lens.sum()) + n, dtype=np.int32)
    sd = None if seq_distances is None else np.ascontiguousarray(seq_distances, dtype=np.float32)
    pm = None if posmaps is None else np.ascontiguousarray(posmaps, dtype=np.int32)
    conf = np.zeros(len(tasks) // 3, dtype=np.float32)
    rc = self.lib.kb200_align_tree_conf(self.h, C.byref(prm), flat, offs, lens, n, tasks, len(tasks) // 3,
                                        None if sd is None else sd.ctypes.data,
                                        None if pm is None else pm.ctypes.data, K, weight, gaps,
                                        conf.ctypes.data if confidence else None, None)
    if rc != 0:
        raise RuntimeError("kb200_align_tree failed")
    out = [gaps[int(offs[i]) + i: int(offs[i]) + i + int(lens[i]) + 1] for i in range(n)]
    return (out, conf) if confidence else out


def _kalign(self, seqs, n_threads=1, type_=8, gpo=-1.0, gpe=-1.0, tgpe=-1.0, consistency=0, weight=2.0, timing=None):
    """kb200_kalign on python strings.  timing: optional list; the wall-clock seconds of the C call alone (host
    char** in, malloc'd host rows out -- what a C caller of kalign() sees) are appended to it."""
    import time
    n = len(seqs)
    keep = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    arr = (C.c_char_p * n)(*keep)
    lens = np.array([len(s) for s in keep], dtype=np.int32)
    out = C.POINTER(C.c_void_p)()
    alen = C.c_int(0)
    t0 = time.perf_counter()
    rc = self.lib.kb200_kalign(self.h, arr, lens, n, n_threads, type_, gpo, gpe, tgpe, consistency, weight,
                               C.byref(out), C.byref(alen))
    if timing is not None:
        timing.append(time.perf_counter() - t0)
    if rc != 0:
        raise RuntimeError("kb200_kalign failed")
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    nrows = int((lens > 0).sum())
    rows = []
    for i in range(nrows):
        rows.append(C.string_at(out[i], alen.value).decode())
        libc.free(out[i])
    libc.free(C.cast(out, C.c_void_p))
    return rows


Context.distances = _distances
Context.anchor_posmaps = _anchor_posmaps
Context.aln_pairwise_dist = _aln_pairwise_dist
Context.kalign_file = _kalign_file
Context.kalign_seeded = _kalign_seeded
Context.guide_tree = _guide_tree
Context.ensemble_runs = _ensemble_runs
Context.align_tree = _align_tree
Context.kalign = _kalign


class Msa:
    """staged pipeline (kb200_msa_*): create -> align (repeatable, device-resident inputs) -> result"""

    def __init__(self, ctx, seqs, n_threads=1, type_=8, gpo=-1.0, gpe=-1.0, tgpe=-1.0, consistency=0, weight=2.0):
        self.ctx = ctx
        self.lib = ctx.lib
        n = len(seqs)
        self._keep = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
        self._arr = (C.c_char_p * n)(*self._keep)
        self._lens = np.array([len(s) for s in self._keep], dtype=np.int32)
        h = C.c_void_p()
        if self.lib.kb200_msa_create(ctx.h, self._arr, self._lens, n, n_threads, type_, gpo, gpe, tgpe,
                                     consistency, weight, C.byref(h)) != 0:
            raise RuntimeError("kb200_msa_create failed")
        self.h = h
        self.nrows = int((self._lens > 0).sum())

    def align(self):
        if self.lib.kb200_msa_align(self.h) != 0:
            raise RuntimeError("kb200_msa_align failed")

    def result(self):
        out = C.POINTER(C.c_void_p)()
        alen = C.c_int(0)
        if self.lib.kb200_msa_result(self.h, C.byref(out), C.byref(alen)) != 0:
            raise RuntimeError("kb200_msa_result failed")
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        rows = []
        for i in range(self.nrows):
            rows.append(C.string_at(out[i], alen.value).decode())
            libc.free(out[i])
        libc.free(C.cast(out, C.c_void_p))
        return rows

    def tree(self):
        """(tasks (N-1) x 3 int32, seq_distances float32[N]) in the sorted index space"""
        n = C.c_int(0)
        self.lib.kb200_msa_info(self.h, C.byref(n), None, None)
        abc = np.zeros((n.value - 1, 3), dtype=np.int32)
        sd = np.zeros(n.value, dtype=np.float32)
        if self.lib.kb200_msa_tree(self.h, abc.ctypes.data, sd.ctypes.data) != 0:
            raise RuntimeError("kb200_msa_tree failed")
        return abc, sd

    def close(self):
        if self.h:
            self.lib.kb200_msa_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
