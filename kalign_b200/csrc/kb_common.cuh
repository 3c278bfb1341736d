// kb_common.cuh -- shared device/host structures of the B200 Hirschberg engine.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <float.h>
#include <vector>

#include "../../include/kalign_b200.h"

#define KB_NEGF (-FLT_MAX)
// most anchors (entries per row of the sparse consistency bonus) the device kernels handle
#define KB_BONUS_KMAX 8

#define KB_CUDA(call)                                                                          \
        do {                                                                                   \
                cudaError_t e__ = (call);                                                      \
                if (e__ != cudaSuccess) {                                                      \
                        fprintf(stderr, "[kalign_b200] CUDA error %s at %s:%d: %s\n",         \
                                cudaGetErrorName(e__), __FILE__, __LINE__, cudaGetErrorString(e__)); \
                        return KB200_FAIL;                                                     \
                }                                                                              \
        } while (0)

#define KB_RUN(call)                                                                           \
        do {                                                                                   \
                if ((call) != KB200_OK) {                                                      \
                        fprintf(stderr, "[kalign_b200] failure at %s:%d\n", __FILE__, __LINE__); \
                        return KB200_FAIL;                                                     \
                }                                                                              \
        } while (0)

// One pairwise alignment ("struct aln_mem" of lib/src/aln_struct.h:16-59, device-resident).
struct KbJob {
        const uint8_t* seq_r;   // SS: row residues
        const uint8_t* seq_c;   // SS, SP: column residues
        const float* prof_r;    // SP, PP: row profile  (len_a+2)*64
        const float* prof_c;    // PP: column profile   (len_b+2)*64
        int len_a;              // DP rows
        int len_b;              // DP cols
        int kind;               // KB200_KIND_*
        int nalpha;             // PP: residues that can have non-zero counts (5 / 23)
        float o, e, t;          // SS: -gpo,-gpe,-tgpe; SP: -(gpo*sip), -(gpe*sip), -(tgpe*sip)
        float nsoff;            // SS: -subm_offset
        int* path;              // raw path, len_a+2
        float4* rowF;           // final forward rows of the job's boxes, (len_a+len_b+2) entries
        float4* rowB;           // final backward rows
        const float* cpack;     // PP: packed column records (len_b+2), built by the engine
        const float* bonus;     // dense bonus (flat i*len_b + j) or nullptr
        const int* bkey;        // sparse bonus: sorted flat keys (i*len_b + j) ...
        const float* bval;      // ... and values; nb entries
        int nb;
        float* score;           // optional: top-level meet-up score
        // optional: per-box meet-up margins (max - max2, aln_seqseq.c:375-385) indexed by the box's
        // position in the recursion tree (root 1, children 2i / 2i+1); -1 = box without a margin,
        // all-ones bits = no such box.  Summed in the reference's recursion order by
        // kb_confidence_kernel (task->confidence, aln_run.c:390-394).
        float* margins;
        unsigned margin_cap;
};

// One Hirschberg box = one (forward sweep, backward sweep, meet-up) triple
// (aln_runner / aln_continue, lib/src/aln_controller.c:21,194).
struct __align__(16) KbBox {
        int job;
        int sa, ea, sb, eb;
        float f0a, f0ga, f0gb;  // injected forward boundary state  (m->f[0])
        float b0a, b0ga, b0gb;  // injected backward boundary state (m->b[0])
        unsigned hid;           // position in the job's recursion tree: root 1, children 2*hid (upper-left) / 2*hid+1
};

// growable device buffer
struct KbDevBuf {
        void* p = nullptr;
        size_t cap = 0;
        int ensure(size_t bytes)
        {
                if (bytes <= cap) {
                        return KB200_OK;
                }
                if (p) {
                        cudaFree(p);
                        p = nullptr;
                        cap = 0;
                }
                size_t want = bytes + bytes / 8 + 256;
                cudaError_t e = cudaMalloc(&p, want);
                if (e != cudaSuccess) {
                        fprintf(stderr, "[kalign_b200] cudaMalloc(%zu) failed: %s\n", want, cudaGetErrorString(e));
                        return KB200_FAIL;
                }
                cap = want;
                return KB200_OK;
        }
        void release()
        {
                if (p) {
                        cudaFree(p);
                }
                p = nullptr;
                cap = 0;
        }
        template <typename T> T* as() const { return (T*)p; }
};

// pinned host staging: descriptors are written here and copied with cudaMemcpyAsync, so that no
// host-to-device copy has to wait for (or be waited on by) anything; reset() at the points where
// the stream is known to be idle (one per guide-tree level)
struct KbPinned {
        std::vector<void*> blocks;
        std::vector<size_t> caps;
        size_t cur = 0, used = 0;
        void* get(size_t bytes);
        void reset() { cur = 0; used = 0; }
        void release();
};

// per Hirschberg round (depth) of one engine call: everything the kernels of the round exchange
// lives on the device, so the host enqueues all rounds without reading anything back
struct KbRound {
        unsigned cursor;        // sweep kernel: next work unit
        unsigned nunits;        // plan kernel -> sweep kernel
        unsigned nboxes;        // boxes of this round (written by the meet-up of the previous one)
        unsigned thin;          // plan kernel -> sweep kernel: thin (32-row) strips this round
        unsigned maxrows;       // plan kernel -> sweep kernel: most rows any sweep of this round has
        unsigned kinds;         // plan kernel -> sweep kernel: bit k set = the round has a box of job kind k
        unsigned pad0, pad1;
};
constexpr int KB_MAX_ROUNDS = 48;

// device-side statistics / error flags, accumulated over engine calls, read back by kb_collect()
struct KbDevStats {
        unsigned long long cells[8];     // sweep ss/sp/pp, bonus, small ss/sp/pp
        unsigned long long nboxes;
        unsigned flags;                  // KB_FLAG_*
        unsigned pad;
};
enum { KB_FLAG_BOX_OVERFLOW = 1, KB_FLAG_UNIT_OVERFLOW = 2, KB_FLAG_SMALL_OVERFLOW = 4, KB_FLAG_ROUNDS = 8, KB_FLAG_STACK = 16,
       KB_FLAG_MARGIN = 32 };

// chunked bump arena for device-resident profiles; chunks are kept across calls (reset())
struct KbArena {
        std::vector<void*> chunks;
        std::vector<size_t> caps;
        size_t chunk_bytes = (size_t)1 << 30;
        size_t cur = 0;       // chunk being filled
        size_t used = 0;      // bytes used in chunk `cur`
        float* alloc_floats(size_t n);
        void reset() { cur = 0; used = 0; }
        void release();
};

struct kb200_ctx {
        int rank = 0, world = 1;     // multi-GPU: one process per GPU (kb200_ctx_comm_init)
        void* comm = nullptr;        // ncclComm_t
        int device = 0;
        unsigned tag_counter = 8192u;  // row-buffer hand-off tags (kb_dp.cu)
        int sm_count = 148;
        cudaStream_t stream = nullptr;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
        kb200_stats stats;
        // timed spans whose events are resolved at the next kb_collect() (no host wait per launch)
        std::vector<cudaEvent_t> ev_pool;
        std::vector<int> ev_kind;        // kind of span i (events 2i, 2i+1): KB_SPAN_*
        size_t ev_used = 0;
        KbPinned pinned;
        KbDevBuf d_stats;                // KbDevStats
        unsigned rows_tag_floor = 0;     // row buffer holds no tag >= this from an earlier owner (0: must be cleared)
        const void* tbl_src = nullptr;   // score table currently staged in d_tbl (host pointer identity + stride)
        int tbl_stride = 0;
        // engine scratch
        KbDevBuf d_jobs, d_boxA, d_boxB, d_boxS, d_counters, d_rows, d_tbl, d_units, d_prog, d_pack, d_ppidx;
        // device buffers of released sequence sets, reused by the next upload (cudaMalloc / cudaFree
        // per public call cost up to hundreds of ms next to a 10 GB profile arena)
        std::vector<KbDevBuf> seq_pool;
        // staging for the host-pointer entry points
        KbDevBuf d_stage0, d_stage1, d_stage2, d_stage3, d_stage4, d_stage5;
        // progressive alignment (kb_tree.cu)
        KbDevBuf t_subm, t_leaf, t_gapset, t_prefix, t_raw, t_coded, t_scr, t_pjobs, t_mjobs, t_src, t_bonus, t_bidx, t_bval, t_posmaps,
                 t_gaps, t_colof, t_aoff, t_bpos, t_bconf, t_binv, t_bdesc, t_wp, t_wdesc, t_alen, t_bvote, t_margin, t_conf;
        const void* posmaps_tag = nullptr;   // host array currently mirrored in t_posmaps
        size_t posmaps_n = 0;
        KbArena arena;
        // pinned host blocks handed to msa objects for their gaps[] result (the 60 MB device-to-host copy at
        // the end of a C3 alignment runs at PCIe speed only into page-locked memory); kept across calls
        struct PinnedBlock { void* p; size_t cap; bool used; };
        std::vector<PinnedBlock> host_pool;
        // device k-means of the guide tree (kb_kmeans.cu): own stream, scratch kept across calls
        cudaStream_t stream2 = nullptr;
        KbDevBuf km_rowsA, km_rowsB, km_ordA, km_ordB, km_side, km_best, km_dmin, km_desc;
};

enum { KB_SPAN_SWEEP = 0, KB_SPAN_DP = 1, KB_SPAN_SMALL = 2 };
// timed span on the context's stream: begin returns the span id (or -1), end closes it
int kb_span_begin(kb200_ctx* ctx, int kind);
void kb_span_end(kb200_ctx* ctx, int span);
// asynchronous host->device copy through the pinned staging area
int kb_h2d(kb200_ctx* ctx, void* dst, const void* src, size_t bytes);
// wait for the stream, fold device statistics and timed spans into ctx->stats, report device-side
// error flags (work-list overflow ...) as KB200_FAIL
int kb_collect(kb200_ctx* ctx);

// DP engine (kb_dp.cu): enqueue all jobs; NO host synchronisation -- results, statistics and error
// flags are valid after the next kb_collect() / stream synchronisation.
// run all jobs (device-resident descriptors are built from `jobs`, whose
// pointers are device pointers; rowF/rowB are assigned here).  subm: 23*23 floats (host).
int kb_run_hirschberg(kb200_ctx* ctx, const float* subm_host, std::vector<KbJob>& jobs);

// task->confidence of every job that carries a margin array: conf_out[j] = margin_sum / margin_count
// accumulated in the reference's recursion order (0 when no meet-up produced a margin)
int kb_confidences(kb200_ctx* ctx, int njobs, float* d_conf_out);
// entries of a job's margin array
unsigned kb_margin_cap(int len_a);

// page-locked host block of at least `bytes` from the context's pool (nullptr on failure) / give it back
void* kb_host_take(kb200_ctx* ctx, size_t bytes);
void kb_host_give(kb200_ctx* ctx, void* p);

// bpm (kb_bpm.cu)
int kb_bpm_pairs(kb200_ctx* ctx, const uint8_t* d_seqs, const int64_t* d_offs, const int* d_lens,
                 const int* d_rows, int nrows, const int* d_cols, int ncols, float* d_dm);

// multi-GPU helpers (kb_comm.cu)
// contiguous partition of n items with the given costs into `world` shards: bounds[0..world]
void kb_partition(const double* cost, int n, int world, int* bounds);
// all-gather of variable-size contiguous segments of one device buffer:
// segment r = bytes [seg[r], seg[r+1]) is owned by rank r and ends up on every rank
int kb_allgatherv(kb200_ctx* ctx, void* dbuf, const size_t* seg);
