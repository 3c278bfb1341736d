// kb_bonus.cuh -- descriptors of the device-side gap weaving / consistency bonus (kb_bonus.cu)
#pragma once
#include "kb_host.cuh"


struct KbWeaveTask {
        const int* path;     // coded path (device)
        int* Pa;             // la+lb+3 ints of scratch
        int* Pb;
        int alnlen;          // unused by the kernel when < 0: read from path[0]
        int* out_len;        // optional: receives the alignment length (the only thing the host needs of the path)
};

struct KbWeaveMember {
        int seq;
        const int* P;        // prefix array of the member's side of its task
};

struct KbBonusOperand {
        int* pos;            // [K][len]
        float* conf;
        int len;
        int m0, m1;          // member range in the level's member list (sip order)
};

struct KbBonusTask {
        const int* pos_a; const float* conf_a; int len_a;    // DP rows operand
        const int* pos_b; const float* conf_b; int len_b;    // DP cols operand
        int* inv;            // sum_k anchor_len ints, initialised to -1
        int* bcol;           // out: len_a*K sorted bonus columns per row (INT_MAX = unused)
        float* bval;         // out: len_a*K values
};

int kb_bonus_init_state(kb200_ctx* ctx, KbSeqs& S, int* d_gaps, int* d_colof);
int kb_weave_level(kb200_ctx* ctx, KbSeqs& S, const KbWeaveTask* d_tasks, int ntasks,
                   const KbWeaveMember* d_members, int nmembers, int* d_gaps, int* d_colof);
// one many-member operand of the sliced vote kernels (kb_bonus_votes_kernel)
struct KbVoteOp {
        int op;              // index into the level's operand array
        int nchunks;         // 32-column chunks
        int nslices;         // slices of KB_BONUS_VOTE_SLICE members
        int pad;
        long long unit0;     // first (chunk, slice) work unit of this operand
        long long vote0;     // first slot of the vote scratch (K * len slots)
        long long col0;      // first column in the concatenated column space of these operands
};
#define KB_BONUS_VOTE_SLICE 64

// operands with at most KB_BONUS_SMALL_NMEM members go to the thread-per-column kernel (small list),
// the others to the sliced warp kernels (vops)
#define KB_BONUS_SMALL_NMEM 16
int kb_bonus_level(kb200_ctx* ctx, KbSeqs& S, int K, float paw,
                   const KbBonusOperand* d_ops,
                   const int* d_small_list, const long long* d_small_prefix, int n_small, long long small_cols,
                   const KbVoteOp* d_vops, int n_large, long long large_units, long long large_cols, long long vote_slots,
                   const int* d_memb, const int* d_colof, const int* d_posmaps,
                   const KbBonusTask* d_tasks, const long long* d_colb_prefix, long long colb_total,
                   const long long* d_row_prefix, long long row_total, int ntasks, const int* d_aoff);
