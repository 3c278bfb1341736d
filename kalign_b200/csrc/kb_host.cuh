// kb_host.cuh -- internal host-side interfaces shared by the C-ABI entry points.
#pragma once
#include "kb_common.cuh"
#include "kb_profile.cuh"

#include <vector>

// sequences resident on the device (codes concatenated) + the host view they were uploaded from
struct KbSeqs {
        const uint8_t* h_seqs = nullptr;
        const int64_t* h_offs = nullptr;
        const int* h_lens = nullptr;
        int n = 0;
        int64_t total = 0;
        KbDevBuf d_seqs, d_offs, d_lens;
        kb200_ctx* owner = nullptr;      // buffers go back to owner->seq_pool on release
        int upload(kb200_ctx* ctx, const uint8_t* seqs, const int64_t* offs, const int* lens, int nseq);
        // device buffers + offsets / lengths only: the codes are written by a kernel (seqs may be nullptr)
        int alloc(kb200_ctx* ctx, const int64_t* offs, const int* lens, int nseq);
        void release();
        const uint8_t* dseq(int i) const { return d_seqs.as<uint8_t>() + h_offs[i]; }
};

// smallest pooled device buffer of the context that fits (none: the caller's ensure() allocates) / hand one back
void kb_take_pooled(kb200_ctx* ctx, KbDevBuf& b, size_t bytes);
void kb_give_pooled(kb200_ctx* ctx, KbDevBuf& b);

// d_estimation replacement on device-resident sequences.
// explicit == 0: rows x cols rectangle; explicit == 1: nrows pairs (rows[p], cols[p]).
int kb_distances_dev(kb200_ctx* ctx, KbSeqs& S, const int* rows, int nrows, const int* cols, int ncols,
                     int explicit_pairs, float* dm_host);

int kb_anchor_posmaps_dev(kb200_ctx* ctx, const kb200_params* prm, KbSeqs& S, const int* anchor_ids, int K,
                          long long pair_begin, long long pair_end, int* posmaps_host, int* d_full = nullptr);
int kb_anchor_posmaps_sharded(kb200_ctx* ctx, const kb200_params* prm, KbSeqs& S, const int* anchor_ids, int K, int* posmaps_host,
                              int host_copy);
// the progressive phase needs the position maps on the HOST only in the A/B mode KB200_HOST_BONUS
// (or when K exceeds the device kernels' compile-time bound)
int kb_bonus_on_host(int K);

int kb_align_tree_dev(kb200_ctx* ctx, const kb200_params* prm, KbSeqs& S,
                      const int* tasks_abc, int ntasks, const float* seq_distances,
                      const int* posmaps, int K, float weight, int n_threads, int* gaps_out, int posmaps_on_device,
                      float* conf_out = nullptr, int* plen_out = nullptr);

// host threads worth using: min(OpenMP max threads, cgroup CPU quota, 16)
int kb_default_threads();
