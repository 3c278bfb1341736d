/*
 * kalign_b200.h -- C ABI of the B200-native kalign alignment hot path.
 *
 * Plain pointers and sizes only.  All entry points return KB200_OK (0) or KB200_FAIL (1), the
 * reference's own convention (lib/src/tldevel.h:29-30); diagnostics go to stderr.  There is NO
 * CPU fallback: without a usable CUDA device every compute entry point fails.
 *
 * The three "seam" entry points sit at exactly the granularity at which the reference's run
 * wrapper (lib/src/aln_wrap.c:133-261) calls into its hot path:
 *
 *   kb200_distances      <->  d_estimation()              lib/src/sequence_distance.c:37
 *                              (called from build_tree_kmeans, lib/src/bisectingKmeans.c:205,294)
 *   kb200_anchor_posmaps <->  anchor_consistency_build()  lib/src/anchor_consistency.c:200
 *                              (the serial N x K pairwise_align_map loop, :246-267)
 *   kb200_align_tree     <->  create_msa_tree()           lib/src/aln_run.c:43
 *                              (recursive_aln / do_align / aln_runner, the aln_task scheduler)
 *
 * kb200_pair_align_batch exposes the batched Hirschberg engine itself (aln_runner,
 * lib/src/aln_controller.c:21); kb200_kalign / kb200_kalign_seeded mirror the public
 * kalign() / kalign_run_seeded() of lib/include/kalign/kalign.h:45,51 on plain arrays
 * (kb200_msa_* is the same call sequence in stages).  On either side of the path:
 *
 *   kb200_guide_tree        <->  build_tree_kmeans[_noisy]()   lib/src/bisectingKmeans.c:177,76
 *   kb200_aln_pairwise_dist <->  compute_aln_pairwise_dist()   lib/src/aln_apair_dist.c:9   (realign loop)
 *   kb200_ensemble_run      <->  one run of kalign_ensemble()  lib/src/ensemble.c:286-340
 *   kb200_fasta_read/_write <->  read_fasta / write_msa_fasta  lib/src/msa_io.c:412,668     (host code)
 *
 * Index space: like the reference after msa_sort_len_name (lib/src/msa_sort.c:14) all per-sequence
 * arrays below are in the caller's order; "seqs" is the concatenation of the internal residue
 * codes (uint8, lib/src/alphabet.c) with offs[i] the start of sequence i.
 */
#ifndef KALIGN_B200_H
#define KALIGN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KB200_OK 0
#define KB200_FAIL 1

/* same numeric values as lib/include/kalign/kalign.h:18-26 */
#define KB200_TYPE_DNA 0
#define KB200_TYPE_DNA_INTERNAL 1
#define KB200_TYPE_RNA 2
#define KB200_TYPE_PROTEIN 3
#define KB200_TYPE_PROTEIN_DIVERGENT 4
#define KB200_TYPE_PROTEIN_PFASUM43 5
#define KB200_TYPE_PROTEIN_PFASUM60 6
#define KB200_TYPE_PROTEIN_PFASUM_AUTO 7
#define KB200_TYPE_UNDEFINED 8

#define KB200_KIND_SS 0   /* sequence x sequence   (lib/src/aln_seqseq.c)         */
#define KB200_KIND_SP 1   /* profile(rows) x seq   (lib/src/aln_seqprofile.c)     */
#define KB200_KIND_PP 2   /* profile x profile     (lib/src/aln_profileprofile.c) */

typedef struct kb200_ctx kb200_ctx;

/* scoring parameters: struct aln_param of lib/src/aln_param.h:24-39, flattened */
typedef struct kb200_params {
        float subm[23 * 23];      /* row-major substitution matrix */
        float gpo, gpe, tgpe;
        float vsm_amax;           /* variable scoring matrix amax (aln_param.c:96) */
        int   nalpha;             /* 5 (nucleotide codes) or 23 (protein codes) */
        float dist_scale;         /* aln_param.dist_scale: per-task gap penalty scale (compute_gap_scale,
                                     lib/src/aln_run.c:126-164); 0 = off (kalign_run_seeded's default) */
        float use_seq_weights;    /* aln_param.use_seq_weights: pseudo-count of the balanced profile merge
                                     (update_n, lib/src/aln_setup.c:237-300); 0 = off */
} kb200_params;

/* one pairwise job for kb200_pair_align_batch (HOST pointers) */
typedef struct kb200_pair {
        int kind;                 /* KB200_KIND_* */
        const uint8_t* seq_rows;  /* SS: row residues */
        const uint8_t* seq_cols;  /* SS, SP: column residues */
        const float* prof_rows;   /* SP, PP: (len_a+2)*64 floats, gap terms [27..29] already set */
        const float* prof_cols;   /* PP: (len_b+2)*64 floats */
        int len_a;                /* DP rows */
        int len_b;                /* DP cols */
        int sip;                  /* SP: sequences in the row profile (aln_seqprofile.c:18) */
        float soff;               /* SS: subm_offset (aln_seqseq.c:38) */
        const float* bonus;       /* optional dense len_a*len_b consistency bonus, or NULL */
        int* path_out;            /* raw path, len_a+2 ints (aln_controller.c:200-201) */
        float* score_out;         /* optional: top-level meet-up score */
} kb200_pair;

/* statistics of the last engine call on a context */
typedef struct kb200_stats {
        double dp_cells;          /* sum over all sweeps of rows * (endb-startb)   */
        double dp_seconds;        /* device time of the DP rounds (CUDA events) */
        double sweep_seconds;     /* device time of the sweep kernels only */
        long long n_boxes;        /* Hirschberg boxes processed */
        long long n_launches;     /* kernels launched by the call */
        double bpm_seconds;       /* device time of the bpm kernel */
        double bpm_pairs;
        double h2d_bytes, d2h_bytes;
        double cells_ss, cells_sp, cells_pp;   /* dp_cells split by kernel kind */
        double cells_bonus;       /* cells that also read the consistency bonus */
        double align_seconds;     /* device-timed span of kb200_msa_align calls (CUDA events) */
        double small_seconds;     /* device time of the small-box kernel */
        double small_ss, small_sp, small_pp;   /* cells handled by the small-box kernel (subset of cells_*) */
        double n_collectives, collective_bytes; /* NCCL all-gathers issued / bytes gathered (multi-GPU) */
        double apair_seconds;     /* device time of the kb200_aln_pairwise_dist kernels */
        double apair_col_pairs;   /* (row pairs) x (alignment columns) they compared */
} kb200_stats;

int  kb200_device_count(void);
int  kb200_ctx_create(int device, kb200_ctx** out);
void kb200_ctx_destroy(kb200_ctx* ctx);
int  kb200_get_stats(kb200_ctx* ctx, kb200_stats* out);
const char* kb200_version(void);

/* default scoring parameters of aln_param_init (lib/src/aln_param.c:17-109).
   biotype: 0 protein, 1 nucleotide (msa_struct.h:19-20); negative gpo/gpe/tgpe keep defaults. */
int kb200_params_init(kb200_params* p, int biotype, int type, float gpo, float gpe, float tgpe);

/* Batched Hirschberg alignments; every job is independent (aln_runner, aln_controller.c:21). */
int kb200_pair_align_batch(kb200_ctx* ctx, const kb200_params* prm,
                           const kb200_pair* jobs, int njobs);

/* dm[r*ncols + c] = calc_distance(seq rows[r], seq cols[c]) + length term
   (sequence_distance.c:117-123,153-162).  Codes must be < 13. */
int kb200_distances(kb200_ctx* ctx, const uint8_t* seqs, const int64_t* offs, const int* lens,
                    int nseq, const int* rows, int nrows, const int* cols, int ncols, float* dm);

/* The same on sequences that stay resident on the device between calls (the reference calls
   d_estimation once per leaf cluster of the guide tree, bisectingKmeans.c:294: thousands of small
   calls on the same msa).  explicit_pairs != 0: nrows pairs (rows[p], cols[p]) -> dm[p], ncols ignored. */
typedef struct kb200_seqs kb200_seqs;
int  kb200_seqs_upload(kb200_ctx* ctx, const uint8_t* seqs, const int64_t* offs, const int* lens, int nseq, kb200_seqs** out);
int  kb200_distances_on(kb200_seqs* s, const int* rows, int nrows, const int* cols, int ncols, int explicit_pairs, float* dm);
void kb200_seqs_free(kb200_seqs* s);

/* compute_aln_pairwise_dist (lib/src/aln_apair_dist.c:9), the N x N matrix of the realign loop
   (kalign_run_realign, lib/src/aln_wrap.c:455-490): rows[i] = alnlen characters of a finished alignment
   ('-' = gap, as written by finalise_alignment); dm_rows[i][j] = dm_rows[j][i] = 1 - matches / aligned over
   the columns where both rows hold a residue, 1 when there is none, 0 on the diagonal
   (pairwise_identity_dist, aln_apair_dist.c:62-82).  dm_rows[i] are the caller's n row allocations of n
   floats each, the reference's float** layout. */
int kb200_aln_pairwise_dist(kb200_ctx* ctx, const char* const* rows, int n, int alnlen, float* const* dm_rows);

/* posmaps: concatenated over i of K maps of len_i ints: map (i,k) starts at
   K*offs[i] + k*lens[i]  (anchor_consistency.c:246-267). Only pairs p in [pair_begin,pair_end)
   of the flattened (i*K+k) list are computed (multi-GPU shard); others are left untouched. */
int kb200_anchor_posmaps(kb200_ctx* ctx, const kb200_params* prm,
                         const uint8_t* seqs, const int64_t* offs, const int* lens, int nseq,
                         const int* anchor_ids, int K,
                         long long pair_begin, long long pair_end, int* posmaps);

/* Anchor choice of anchor_consistency_build (static select_anchors, anchor_consistency.c:124-198):
   K diverse sequences by farthest-point sampling on seq_distances.  Host-only. */
int kb200_select_anchors(const float* seq_distances, int nseq, int K, int* anchor_ids);

/* Progressive alignment over a guide tree (create_msa_tree, aln_run.c:43).
   tasks_abc: ntasks x 3 (a, b, c) sorted by c (task.c:151-161); seq_distances: nseq floats
   (bisectingKmeans.c:247-256) or NULL; posmaps/K/weight: consistency table or NULL/0.
   gaps_out: concatenated gaps[] of every sequence (len_i+1 ints each, at offs[i]+i). */
int kb200_align_tree(kb200_ctx* ctx, const kb200_params* prm,
                     const uint8_t* seqs, const int64_t* offs, const int* lens, int nseq,
                     const int* tasks_abc, int ntasks, const float* seq_distances,
                     const int* posmaps, int K, float weight,
                     int* gaps_out);
/* the same, also returning per task (task order; either may be NULL):
   task_confidence: task->confidence, the mean meet-up margin margin_sum / margin_count accumulated
                    in the reference's recursion order (lib/src/aln_seqseq.c:375-385, aln_run.c:390-394);
   task_plen:       msa->plen[c], the length of the merged profile (aln_run.c:424) */
int kb200_align_tree_conf(kb200_ctx* ctx, const kb200_params* prm,
                          const uint8_t* seqs, const int64_t* offs, const int* lens, int nseq,
                          const int* tasks_abc, int ntasks, const float* seq_distances,
                          const int* posmaps, int K, float weight,
                          int* gaps_out, float* task_confidence, int* task_plen);

/* kalign() of lib/include/kalign/kalign.h:45 on the GPU: same arguments, same ownership
   (rows and the array are malloc'd; caller frees).  consistency_anchors = 0 reproduces
   kalign()/kalign_run(); 5 with weight 2.0 reproduces the CLI default mode.
   Like kalign()'s array mode (kalign_arr_to_msa, lib/src/msa_op.c:440) the characters are taken as
   they are: '-' / '.' are NOT stripped (gaps[] start at zero; only file input de-aligns), any
   character outside the alphabet is coded as residue 0 with a warning (msa_op.c:358-362).
   Equal-length sequences are ordered by the names "s<i>" (i = input position); the reference leaves
   the names of an array-mode msa uninitialised, so its own order of equal-length sequences is
   undefined there -- the FASTA path (names from the file) is the one tests/ and bench.py pin. */
int kb200_kalign(kb200_ctx* ctx, char** seq, int* len, int numseq, int n_threads, int type,
                 float gpo, float gpe, float tgpe, int consistency_anchors, float consistency_weight,
                 char*** aligned, int* out_aln_len);

/* kalign_run_seeded (lib/include/kalign/kalign.h:51, lib/src/aln_wrap.c:133) on plain arrays, refine = none:
   kb200_kalign plus
     tree_seed, tree_noise : tree_seed != 0 && tree_noise > 0 builds the guide tree from anchor distances
                             multiplied by max(0.1, gaussian(1, tree_noise)) drawn from the reference's generator
                             (build_tree_kmeans_noisy, lib/src/bisectingKmeans.c:76-176; lib/src/tlrng.c)
     dist_scale            : ap->dist_scale (compute_gap_scale, lib/src/aln_run.c:126)
     vsm_amax, use_seq_weights : >= 0 override the defaults of aln_param_init (aln_wrap.c:193-199) */
int kb200_kalign_seeded(kb200_ctx* ctx, char** seq, int* len, int numseq, int n_threads, int type,
                        float gpo, float gpe, float tgpe, unsigned long long tree_seed, float tree_noise,
                        float dist_scale, float vsm_amax, float use_seq_weights,
                        int consistency_anchors, float consistency_weight, char*** aligned, int* out_aln_len);
/* build_tree_kmeans / build_tree_kmeans_noisy (lib/src/bisectingKmeans.c:177,76) on plain arrays: anchor distances,
   bisecting k-means and the UPGMA of the leaf clusters on the GPU.  seqs: codes of the TREE alphabet (13-letter
   reduced protein / nucleotide, what msa->sequences[i]->s holds when the reference calls it), sorted order.
   tasks_abc: (nseq - 1) x 3 ints (a, b, c) in the order create_tasks (:1084) fills struct aln_tasks -- a node, its left
   subtree, its right subtree; seq_distances (may be NULL): msa->seq_distances (:247-256). */
int kb200_guide_tree(kb200_ctx* ctx, const uint8_t* seqs, const int64_t* offs, const int* lens, int nseq, int n_threads,
                     unsigned long long tree_seed, float tree_noise, int* tasks_abc, float* seq_distances);
/* the code tables convert_msa_to_internal uses (create_alphabet, lib/src/alphabet.c:140-302): letters = 5 (nucleotide:
   A C G T/U and N for every IUPAC ambiguity code), 13 (reduced protein alphabet of the guide tree), 23 (protein
   alphabet of the alignment, ARNDCQEGHILKMFPSTWYVBZX, U as X); to_internal[128] maps a character to its code, -1 when it
   is not in the alphabet (such characters are coded 0 with a warning, msa_op.c:358-362); *L = number of codes.  Host code. */
int kb200_alphabet(int letters, signed char* to_internal, int* L);
/* creation order of a task list sorted by c (node ids nseq + index), host code */
int kb200_tasks_creation_order(const int* tasks_sorted, int ntasks, int nseq, int* tasks_out);
/* the n noise factors of build_tree_kmeans_noisy for (seed, sigma), in the order they are applied
   (row by row over the N x 32 anchor distances).  Host code. */
int kb200_tree_noise(unsigned long long seed, float sigma, long long n, float* out);

/* Ensemble (kalign_ensemble, lib/src/ensemble.c:221-340): n_runs independent alignments of the
   same sequences -- run 0 with the base penalties, run k > 0 with penalties scaled by entry k % 12 of the
   reference's table and a noisy guide tree seeded seed + k (resolve_run_params, ensemble.c:55-76).
   kb200_ensemble_run_params resolves run k's parameters, kb200_ensemble_run computes run k (gpo / gpe / tgpe
   < 0: the defaults of the detected alphabet; use_seq_weights < 0 means 0 here, ensemble.c:248).  The runs
   share nothing, so they shard over GPUs with no collective: rank r of `world` processes (one per GPU)
   computes the runs k with k % world == r.  POAR consensus / scoring of the collected runs is host code and
   stays the reference's. */
int kb200_ensemble_run_params(float base_gpo, float base_gpe, float base_tgpe, int run, unsigned long long seed,
                              float* gpo, float* gpe, float* tgpe, unsigned long long* tree_seed, float* tree_noise);
int kb200_ensemble_run(kb200_ctx* ctx, char** seq, int* len, int numseq, int n_threads, int type,
                       float gpo, float gpe, float tgpe, int run, unsigned long long seed,
                       float dist_scale, float vsm_amax, float use_seq_weights,
                       int consistency_anchors, float consistency_weight, char*** aligned, int* out_aln_len);

/* FASTA in / out, host code: the file is mapped and parsed by n_threads threads at once
   (n_threads <= 0: all the process may use) with the semantics of read_file_stdin + read_fasta
   (lib/src/msa_io.c:348,412): a line's content ends at its first control character; '>' at the start of a
   line opens a record whose name is the rest of the line; on the other lines letters are residues,
   punctuation counts as gap characters in front of the next residue (gaps[len]++), everything else is
   dropped; letter_freq[128] counts every character of the sequence lines (what detect_alphabet reads).
   Views returned by kb200_fasta_get / _arrays / _letter_freq stay valid until kb200_fasta_free. */
typedef struct kb200_fasta kb200_fasta;
int  kb200_fasta_read(const char* path, int n_threads, kb200_fasta** out);
int  kb200_fasta_numseq(const kb200_fasta* f);
/* record i: name, residues (NUL-terminated), length, len + 1 gap counts; any out pointer may be NULL */
int  kb200_fasta_get(const kb200_fasta* f, int i, const char** name, const char** seq, int* len, const int** gaps);
const int* kb200_fasta_letter_freq(const kb200_fasta* f);
/* the (seq, len) arrays kb200_kalign / kalign() take */
int  kb200_fasta_arrays(const kb200_fasta* f, const char* const** seqs, const int** lens);
void kb200_fasta_free(kb200_fasta* f);
/* write_msa_fasta (lib/src/msa_io.c:668): ">name", the row in lines of 60 characters; one buffer, one write */
int  kb200_fasta_write(const char* path, const char* const* names, const char* const* rows, int n, int alnlen, int n_threads);
/* the CLI's main path for one FASTA file (src/run_kalign.c: kalign_read_input -> kalign_run_seeded ->
   kalign_write_msa with the default output format): record names break the ties of the (length, name)
   sort, empty records are dropped, rows are written in input order */
int  kb200_kalign_file(kb200_ctx* ctx, const char* infile, const char* outfile, int n_threads, int type,
                       float gpo, float gpe, float tgpe, int consistency_anchors, float consistency_weight);

/* Multi-GPU (one process per GPU of one node, NCCL over NVLink): rank 0 creates a unique id,
   the caller distributes it (e.g. torch.distributed broadcast), every rank attaches its context.
   With a communicator attached, kb200_msa_align shards the N x K anchor pairs and the tasks of every
   guide-tree level across the ranks (the reference's independent OpenMP tasks,
   lib/src/aln_run.c:95-109) and all-gathers position maps / coded paths / merged profiles; every
   rank ends with the complete result.  kb200_partition is the (host-only) contiguous cost-balanced
   partition used for both. */
#define KB200_COMM_ID_BYTES 128
int  kb200_comm_unique_id(void* id_out, int nbytes);
int  kb200_ctx_comm_init(kb200_ctx* ctx, int rank, int world, const void* id, int nbytes);
void kb200_ctx_comm_destroy(kb200_ctx* ctx);
int  kb200_partition(const double* cost, int n, int world, int* bounds);

/* The same pipeline in stages, so that the DP stages can be run (and timed) on sequences that
   are already resident in HBM:
     kb200_msa_create : everything of kalign_run_seeded that precedes the DP stages
                        (aln_wrap.c:144-205: check, sort, encode, upload, distances, guide tree,
                         parameters, anchor selection)
     kb200_msa_align  : anchor_consistency_build + create_msa_tree (aln_wrap.c:208-226), repeatable
     kb200_msa_result : finalise_alignment + msa_sort_rank + kalign_msa_to_arr (aln_wrap.c:240-242)
   The input strings must stay alive until kb200_msa_free. */
typedef struct kb200_msa kb200_msa;
int  kb200_msa_create(kb200_ctx* ctx, char** seq, int* len, int numseq, int n_threads, int type,
                      float gpo, float gpe, float tgpe, int consistency_anchors, float consistency_weight,
                      kb200_msa** out);
int  kb200_msa_align(kb200_msa* m);
int  kb200_msa_result(kb200_msa* m, char*** aligned, int* out_aln_len);
int  kb200_msa_info(kb200_msa* m, int* numseq, int* biotype, int* n_anchors);
/* the guide tree kb200_msa_create built (build_tree_kmeans, lib/src/bisectingKmeans.c:177-271):
   tasks_abc receives (numseq-1) x 3 ints (a, b, c) in task order (sorted by c), seq_distances
   numseq floats (msa->seq_distances, :247-256), both in the sorted index space; either may be NULL */
int  kb200_msa_tree(kb200_msa* m, int* tasks_abc, float* seq_distances);
void kb200_msa_free(kb200_msa* m);

#ifdef __cplusplus
}
#endif
#endif
