// kb_profile.cuh -- descriptors of the streaming profile / path kernels (kb_profile.cu)
#pragma once
#include "kb_common.cuh"

struct KbLeafProfile {
        const uint8_t* seq;
        float* prof;            // (len+2)*64 floats
        int len;
        float nsoff;            // -subm_offset of the consuming task (aln_run.c:239-242)
        float ngpo, ngpe, ntgpe;
};

struct KbGapSet {
        float* prof;
        int len;                // columns 0..len+1 are touched
        int nsip;               // size of the OTHER operand (aln_run.c:244,252)
};

struct KbPathJob {
        const int* raw;         // raw Hirschberg path (DP orientation)
        int* coded;             // out: len_a+len_b+2 ints
        int* scratch;           // len_a+2 ints, only when mirror
        int* posmap;            // optional, len_a ints
        int len_a, len_b;       // UN-swapped lengths (a = rows of the coded path)
        int mirror;
};

struct KbMergeJob {
        const float* pa;
        const float* pb;
        float* newp;            // (alnlen+2)*64
        const int* path;        // coded path
        int2* src;              // alnlen+2 entries
        int alnlen;
        int sipa, sipb;
        float gpo, gpe, tgpe;
        // balanced merge (use_seq_weights > 0, aln_setup.c:237-300): residue counts of the match columns
        // are rescaled per side, the summed substitution scores corrected by the matching delta
        int rebalance;
        float scaleA, scaleB;
        const float* subm;      // 23 x 23, device
};

int kb_make_profiles(kb200_ctx* ctx, const KbLeafProfile* d_leaves, int nleaves,
                     const long long* d_prefix, long long total_cols, const float* d_subm);
int kb_set_gap_penalties(kb200_ctx* ctx, const KbGapSet* d_sets, int nsets,
                         const long long* d_prefix, long long total_cols);
int kb_code_paths(kb200_ctx* ctx, const KbPathJob* d_pj, int njobs);
int kb_merge_index(kb200_ctx* ctx, const KbMergeJob* d_mj, int njobs);
int kb_merge_profiles(kb200_ctx* ctx, const KbMergeJob* d_mj, int njobs,
                      const long long* d_prefix, long long total_cols);
int kb_bpm_pairs_words(kb200_ctx* ctx, int max_words, const uint8_t* d_seqs, const int64_t* d_offs, const int* d_lens,
                       const int* d_rows, int nrows, const int* d_cols, int ncols, float* d_dm);
