// kb_bpm.cu -- batched bit-parallel Myers/Hyyro block edit distance.
//
// Replaces bpm_block (lib/src/bpm.c:356-580) as used by calc_distance / d_estimation
// (lib/src/sequence_distance.c:37-162).  Value restated (see oracle/kalign_oracle.c): with
// maxd = m every 64-bit block stays active for the whole text, so
//     k = min(m, min_i score[last block] after text symbol i),  i over n + W symbols,
// W = 64*ceil(m/64) - m zero symbols appended to the text, pattern positions >= m match anything,
// pattern truncated at 1024 (bpm.c:369-371).
//
// Parallel shape: LP lanes (LP = 1..16, one per 64-bit block) own one pair; lane b processes text
// symbol t-b at step t (wavefront over blocks), the horizontal carry (hout -> hin) moves one lane
// per step through a segmented warp shuffle.  Peq[13][LP] lives in shared memory.
#include "kb_common.cuh"

#include <algorithm>

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int BPM_THREADS = 128;

template <int LP>
__global__ void __launch_bounds__(BPM_THREADS)
kb_bpm_kernel(const uint8_t* __restrict__ seqs, const int64_t* __restrict__ offs, const int* __restrict__ lens,
              const int* __restrict__ rows, const int nrows, const int* __restrict__ cols, const int ncols,
              float* __restrict__ dm)
{
        // ncols > 0: rectangular rows x cols matrix, pair = r*ncols + c  -> (rows[r], cols[c])
        // ncols == 0: explicit pair list of nrows pairs             -> (rows[p], cols[p])
        __shared__ unsigned long long s_peq[(BPM_THREADS / LP) * 13 * LP];
        const int tid = threadIdx.x;
        const int grp = tid / LP;
        const int b = tid % LP;
        const long long npairs = (ncols > 0) ? (long long)nrows * (long long)ncols : (long long)nrows;
        const long long pair = (long long)blockIdx.x * (BPM_THREADS / LP) + grp;
        const bool live = pair < npairs;
        unsigned long long* peq = s_peq + grp * 13 * LP;
        const uint8_t* tx = seqs;
        const uint8_t* pt = seqs;
        int n = 0, m = 0, l1 = 0, l2 = 0;
        if (live) {
                int s1, s2;
                if (ncols > 0) {
                        s1 = rows[(int)(pair / ncols)];
                        s2 = cols[(int)(pair % ncols)];
                } else {
                        s1 = rows[pair];
                        s2 = cols[pair];
                }
                l1 = lens[s1];
                l2 = lens[s2];
                // calc_distance (sequence_distance.c:153-162): longer is the text; on equal
                // length the second argument is the text
                if (l1 > l2) {
                        tx = seqs + offs[s1]; n = l1;
                        pt = seqs + offs[s2]; m = l2;
                } else {
                        tx = seqs + offs[s2]; n = l2;
                        pt = seqs + offs[s1]; m = l1;
                }
                if (m > 1024) {
                        m = 1024;
                }
        }
        const int bmax = (m == 0) ? 1 : ((m + 63) >> 6);
        const int W = 64 * bmax - m;
        // ---- Peq for my block ----
#pragma unroll
        for (int c = 0; c < 13; c++) {
                peq[c * LP + b] = 0ull;
        }
        if (live && b < bmax) {
                unsigned long long wild = 0ull;
                for (int i = 0; i < 64; i++) {
                        const int idx = b * 64 + i;
                        const unsigned long long bit = 1ull << i;
                        if (idx >= m) {
                                wild |= bit;
                        } else {
                                const int ch = pt[idx];
                                peq[ch * LP + b] |= bit;
                        }
                }
                if (wild) {
#pragma unroll
                        for (int c = 0; c < 13; c++) {
                                peq[c * LP + b] |= wild;
                        }
                }
        }
        __syncwarp();
        unsigned long long Pv = ~0ull, Mv = 0ull;
        int score = (b + 1) * 64;
        int k = m;
        int hout = 0;
        // all groups of a warp run the same number of steps (max over the warp)
        int steps = live ? (n + W + bmax - 1) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
                steps = max(steps, __shfl_xor_sync(FULL, steps, o));
        }
        const int nW = n + W;
        for (int t = 0; t < steps; t++) {
                int hin = __shfl_up_sync(FULL, hout, 1, LP);
                if (b == 0) {
                        hin = 0;
                }
                const int i = t - b;
                if (live && b < bmax && i >= 0 && i < nW) {
                        const int ch = (i < n) ? (int)tx[i] : 0;
                        unsigned long long Eq = peq[ch * LP + b];
                        const unsigned long long Xv = Eq | Mv;
                        if (hin < 0) {
                                Eq |= 1ull;
                        }
                        const unsigned long long Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
                        unsigned long long Ph = Mv | ~(Xh | Pv);
                        unsigned long long Mh = Pv & Xh;
                        hout = (int)(Ph >> 63) - (int)(Mh >> 63);
                        Ph <<= 1;
                        Mh <<= 1;
                        if (hin < 0) {
                                Mh |= 1ull;
                        } else if (hin > 0) {
                                Ph |= 1ull;
                        }
                        Pv = Mh | ~(Xv | Ph);
                        Mv = Ph & Xv;
                        score += hout;
                        if (b == bmax - 1 && score < k) {
                                k = score;
                        }
                }
        }
        if (live && b == bmax - 1) {
                // sequence_distance.c:120-123: length term in double, narrowed to float, added in float
                float dist = (float)(unsigned int)k;
                const int s = (l1 + l2) / 2;
                const double sd = (double)s;
                const float add = (float)(((10000.0 < sd) ? 10000.0 : sd) / 10000.0);
                dist += add;
                dm[pair] = dist;
        }
}

template <int LP>
int launch(kb200_ctx* ctx, const uint8_t* d_seqs, const int64_t* d_offs, const int* d_lens,
           const int* d_rows, int nrows, const int* d_cols, int ncols, float* d_dm)
{
        const long long npairs = (ncols > 0) ? (long long)nrows * (long long)ncols : (long long)nrows;
        const int per = BPM_THREADS / LP;
        const long long grid = (npairs + per - 1) / per;
        if (grid > 0x7fffffffLL) {
                fprintf(stderr, "[kalign_b200] bpm: too many pairs for one launch\n");
                return KB200_FAIL;
        }
        kb_bpm_kernel<LP><<<(unsigned)grid, BPM_THREADS, 0, ctx->stream>>>(d_seqs, d_offs, d_lens, d_rows, nrows, d_cols, ncols, d_dm);
        KB_CUDA(cudaGetLastError());
        return KB200_OK;
}

} // namespace

// max_words: largest number of 64-bit pattern blocks any pair of this launch needs (1..16)
int kb_bpm_pairs_words(kb200_ctx* ctx, int max_words, const uint8_t* d_seqs, const int64_t* d_offs, const int* d_lens,
                       const int* d_rows, int nrows, const int* d_cols, int ncols, float* d_dm)
{
        if (nrows <= 0 || ncols < 0) {
                return KB200_OK;
        }
        KB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
        int rc;
        if (max_words <= 1) rc = launch<1>(ctx, d_seqs, d_offs, d_lens, d_rows, nrows, d_cols, ncols, d_dm);
        else if (max_words <= 2) rc = launch<2>(ctx, d_seqs, d_offs, d_lens, d_rows, nrows, d_cols, ncols, d_dm);
        else if (max_words <= 4) rc = launch<4>(ctx, d_seqs, d_offs, d_lens, d_rows, nrows, d_cols, ncols, d_dm);
        else if (max_words <= 8) rc = launch<8>(ctx, d_seqs, d_offs, d_lens, d_rows, nrows, d_cols, ncols, d_dm);
        else rc = launch<16>(ctx, d_seqs, d_offs, d_lens, d_rows, nrows, d_cols, ncols, d_dm);
        if (rc != KB200_OK) {
                return rc;
        }
        KB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
        KB_CUDA(cudaStreamSynchronize(ctx->stream));
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        ctx->stats.bpm_seconds += 1e-3 * (double)ms;
        ctx->stats.bpm_pairs += (ncols > 0) ? (double)nrows * (double)ncols : (double)nrows;
        ctx->stats.n_launches += 1;
        return KB200_OK;
}

int kb_bpm_pairs(kb200_ctx* ctx, const uint8_t* d_seqs, const int64_t* d_offs, const int* d_lens,
                 const int* d_rows, int nrows, const int* d_cols, int ncols, float* d_dm)
{
        return kb_bpm_pairs_words(ctx, 16, d_seqs, d_offs, d_lens, d_rows, nrows, d_cols, ncols, d_dm);
}
