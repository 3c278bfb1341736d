// kb_profile.cu -- streaming profile kernels and path post-processing (device resident between
// guide-tree levels).
//
// Replaces (behaviour cited, nothing copied):
//   make_profile_n         lib/src/aln_setup.c:40    leaf profile, (len+2) x 64 floats
//   set_gap_penalties_n    lib/src/aln_setup.c:101   [27..29] = [55..57] * nsip(other operand)
//   update_n               lib/src/aln_setup.c:230   merge along the coded path (incl. the use_seq_weights rebalance)
//   mirror_path_n          lib/src/aln_setup.c:438
//   add_gap_info_to_path_n lib/src/aln_setup.c:121
//   pairwise_align_map     lib/src/anchor_consistency.c:85-111 (coded path -> position map)
#include "kb_common.cuh"
#include "kb_profile.cuh"
#include <algorithm>

namespace {

// one warp per profile column, 2 floats per lane
__global__ void kb_make_profiles_kernel(const KbLeafProfile* __restrict__ leaves, const int nleaves,
                                        const long long* __restrict__ col_prefix, const long long total_cols,
                                        const float* __restrict__ subm /* 23x23 */)
{
        const int lane = threadIdx.x & 31;
        const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
        for (long long gc = warp; gc < total_cols; gc += nwarps) {
                // find the leaf owning global column gc (binary search on the prefix)
                int lo = 0, hi = nleaves - 1;
                while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (col_prefix[mid] <= gc) lo = mid; else hi = mid - 1;
                }
                const KbLeafProfile L = leaves[lo];
                const int col = (int)(gc - col_prefix[lo]);      // 0..len+1
                float* out = L.prof + ((size_t)col << 6);
                float v0 = 0.0f, v1 = 0.0f;                        // elements lane, lane+32
                if (col >= 1 && col <= L.len) {
                        const int c = L.seq[col - 1];
                        if (lane == c) {
                                v0 = 1.0f;                         // prof[c] += weight (weight = 1)
                        }
                        if (lane < 23) {
                                v1 = subm[c * 23 + lane] + L.nsoff; // subm[c][j] - soff
                        }
                }
                if (lane == 23) v1 = L.ngpo;
                if (lane == 24) v1 = L.ngpe;
                if (lane == 25) v1 = L.ntgpe;
                out[lane] = v0;
                out[lane + 32] = v1;
        }
}

__global__ void kb_set_gap_kernel(const KbGapSet* __restrict__ sets, const int nsets,
                                  const long long* __restrict__ col_prefix, const long long total_cols)
{
        const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        const long long nth = (long long)gridDim.x * blockDim.x;
        for (long long gc = tid; gc < total_cols; gc += nth) {
                int lo = 0, hi = nsets - 1;
                while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (col_prefix[mid] <= gc) lo = mid; else hi = mid - 1;
                }
                const KbGapSet S = sets[lo];
                float* col = S.prof + ((size_t)(gc - col_prefix[lo]) << 6);
                const float f = (float)S.nsip;
                col[27] = col[55] * f;
                col[28] = col[56] * f;
                col[29] = col[57] * f;
        }
}

// ---- warp-parallel path post-processing -------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v, const int lane)
{
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
        }
        return v;
}

// entries a row contributes to the coded path: `skip` gap-in-A entries, then one 0 / 2
// (add_gap_info_to_path_n, aln_setup.c:141-180: the skip is only emitted when the previous row
// was matched, b != -1, and for row 1 it is path[1]-1)
__device__ __forceinline__ int row_skip(const int i, const int r, const int rp)
{
        if (r == -1) return 0;
        if (i == 1) return r - 1;
        return (r - 1 != rp && rp != -1) ? (r - rp - 1) : 0;
}

// raw path -> (mirror) -> coded path (+ optional position map).  One WARP per job: the walk of
// the reference is a prefix sum over per-row entry counts, so it is done with warp scans.
__global__ void kb_code_path_kernel(const KbPathJob* __restrict__ pj, const int njobs)
{
        const int lane = threadIdx.x & 31;
        const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        if (j >= njobs) {
                return;
        }
        const KbPathJob P = pj[j];
        const int len_a = P.len_a, len_b = P.len_b;
        const int* raw = P.raw;
        int* __restrict__ o = P.coded;
        if (P.mirror) {
                // raw was produced with rows = b (len_b entries); scratch receives the mirrored path
                int* mr = P.scratch;
                for (int i = lane; i < len_a + 2; i += 32) {
                        mr[i] = -1;
                }
                __syncwarp();
                for (int i = 1 + lane; i <= len_b; i += 32) {
                        const int c = raw[i];
                        if (c != -1) {
                                mr[c] = i;
                        }
                }
                __syncwarp();
                raw = mr;
        }
        // pass A: total length, first / last matched row and the index of their 0-entry
        int carry = 1;                       // next free index of o
        int first_pz = 0x7fffffff, last_pz = -1;
        for (int base = 1; base <= len_a; base += 32) {
                const int i = base + lane;
                int cnt = 0, skip = 0, r = -1;
                if (i <= len_a) {
                        r = raw[i];
                        const int rp = (i > 1) ? raw[i - 1] : -1;
                        skip = row_skip(i, r, rp);
                        cnt = skip + 1;
                }
                const int incl = warp_incl_scan(cnt, lane);
                const int start = carry + incl - cnt;
                if (i <= len_a && r != -1) {
                        const int pz = start + skip;
                        first_pz = min(first_pz, pz);
                        last_pz = max(last_pz, pz);
                }
                carry += __shfl_sync(0xffffffffu, incl, 31);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
                first_pz = min(first_pz, __shfl_xor_sync(0xffffffffu, first_pz, off));
                last_pz = max(last_pz, __shfl_xor_sync(0xffffffffu, last_pz, off));
        }
        int total = carry;                   // index after the last row entry
        const int lastr = raw[len_a];
        const int tail = (lastr != -1 && lastr < len_b) ? (len_b - lastr) : 0;
        const int end = total + tail;        // index of the terminator
        // pass B: write the entries (terminal bit 32 on the leading / trailing gap runs,
        // aln_setup.c:212-222; the open/ext/close loop of the reference never runs, :191-195)
        carry = 1;
        int carry_b = 0;                     // residues of b consumed so far (position map)
        for (int base = 1; base <= len_a; base += 32) {
                const int i = base + lane;
                int cnt = 0, skip = 0, r = -1, adv = 0;
                if (i <= len_a) {
                        r = raw[i];
                        const int rp = (i > 1) ? raw[i - 1] : -1;
                        skip = row_skip(i, r, rp);
                        cnt = skip + 1;
                        adv = skip + ((r != -1) ? 1 : 0);
                }
                const int incl = warp_incl_scan(cnt, lane);
                const int inclb = warp_incl_scan(adv, lane);
                const int start = carry + incl - cnt;
                if (i <= len_a) {
                        for (int q = 0; q < skip; q++) {
                                const int idx = start + q;
                                o[idx] = 1 | ((idx < first_pz || idx > last_pz) ? 32 : 0);
                        }
                        const int idx = start + skip;
                        if (r == -1) {
                                o[idx] = 2 | ((idx < first_pz || idx > last_pz) ? 32 : 0);
                        } else {
                                o[idx] = 0;
                        }
                        if (P.posmap) {
                                // anchor_consistency.c:85-111: a match maps row i-1 to the b position
                                P.posmap[i - 1] = (r != -1) ? (carry_b + inclb - adv + skip) : -1;
                        }
                }
                carry += __shfl_sync(0xffffffffu, incl, 31);
                carry_b += __shfl_sync(0xffffffffu, inclb, 31);
        }
        for (int q = lane; q < tail; q += 32) {
                const int idx = total + q;
                o[idx] = 1 | ((idx < first_pz || idx > last_pz) ? 32 : 0);
        }
        if (lane == 0) {
                o[0] = end - 1;
                o[end] = 3;
        }
}

// per merge: source column indices (1-based) of every output column = exclusive counts of the
// a-consuming / b-consuming entries before it (update_n's profa/profb pointer walk)
__global__ void kb_merge_index_kernel(const KbMergeJob* __restrict__ mj, const int njobs)
{
        const int lane = threadIdx.x & 31;
        const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        if (j >= njobs) {
                return;
        }
        const KbMergeJob M = mj[j];
        const int* __restrict__ path = M.path;
        int2* __restrict__ src = M.src;
        const int alnlen = M.alnlen;
        int ca = 1, cb = 1;
        for (int base = 1; base <= alnlen + 1; base += 32) {
                const int c = base + lane;
                int fa = 0, fb = 0;
                if (c <= alnlen) {
                        const int p = path[c];
                        fa = (!p || (p & 2)) ? 1 : 0;
                        fb = (!p || (p & 1)) ? 1 : 0;
                }
                const int ia = warp_incl_scan(fa, lane);
                const int ib = warp_incl_scan(fb, lane);
                if (c <= alnlen + 1) {
                        src[c] = make_int2(ca + ia - fa, cb + ib - fb);
                }
                ca += __shfl_sync(0xffffffffu, ia, 31);
                cb += __shfl_sync(0xffffffffu, ib, 31);
        }
}

__device__ __forceinline__ float gap_adjust_val(float v, const int lane_el, const int p, const float sip,
                                                const float gpo, const float gpe, const float tgpe)
{
        // lane_el: element index 0..63.  update_n, aln_setup.c:321-365 / :374-417
        if (!(p & 20)) {
                float gp;
                if (p & 32) {
                        if (lane_el == 25) v += sip;
                        gp = tgpe * sip;
                } else {
                        if (lane_el == 24) v += sip;
                        gp = gpe * sip;
                }
                if (lane_el >= 32 && lane_el < 55) v -= gp;
                return v;
        }
#pragma unroll
        for (int rep = 0; rep < 2; rep++) {
                if (!(p & (rep == 0 ? 16 : 4))) {
                        continue;
                }
                float gp;
                if (p & 32) {
                        if (lane_el == 25) v += sip;
                        gp = tgpe * sip;
                        if (lane_el == 23) v += sip;
                        gp += gpo * sip;
                } else {
                        if (lane_el == 23) v += sip;
                        gp = gpo * sip;
                }
                if (lane_el >= 32 && lane_el < 55) v -= gp;
        }
        return v;
}

__global__ void kb_merge_kernel(const KbMergeJob* __restrict__ mj, const int njobs,
                                const long long* __restrict__ col_prefix, const long long total_cols)
{
        const int lane = threadIdx.x & 31;
        const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
        for (long long gc = warp; gc < total_cols; gc += nwarps) {
                int lo = 0, hi = njobs - 1;
                while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (col_prefix[mid] <= gc) lo = mid; else hi = mid - 1;
                }
                const KbMergeJob M = mj[lo];
                const int c = (int)(gc - col_prefix[lo]);           // output column 0..alnlen+1
                const int alnlen = M.alnlen;
                float* np = M.newp + ((size_t)c << 6);
                float v0, v1;
                // balanced merge of a match / boundary column: counts [0..22] = a*scaleA + b*scaleB
                // (two rounded products, one rounded sum), everything else a + b
                auto balanced0 = [&](const float* a, const float* b) -> float {
                        if (M.rebalance && lane < 23) {
                                return __fadd_rn(__fmul_rn(a[lane], M.scaleA), __fmul_rn(b[lane], M.scaleB));
                        }
                        return a[lane] + b[lane];
                };
                if (c == 0) {
                        v0 = balanced0(M.pa, M.pb);
                        v1 = M.pa[lane + 32] + M.pb[lane + 32];
                } else {
                        const int2 s = M.src[c];
                        const float* a = M.pa + ((size_t)s.x << 6);
                        const float* b = M.pb + ((size_t)s.y << 6);
                        const int p = (c <= alnlen) ? M.path[c] : 0;
                        if (c > alnlen || !p) {
                                v0 = balanced0(a, b);
                                v1 = a[lane + 32] + b[lane + 32];
                                if (M.rebalance && c <= alnlen && lane < 23) {
                                        // newp[32+j] += sum_aa (a[aa]*dA + b[aa]*dB) * subm[aa][j], aa ascending from 0.0f
                                        const float dA = __fadd_rn(M.scaleA, -1.0f);
                                        const float dB = __fadd_rn(M.scaleB, -1.0f);
                                        float delta = 0.0f;
                                        for (int aa = 0; aa < 23; aa++) {
                                                const float t = __fadd_rn(__fmul_rn(a[aa], dA), __fmul_rn(b[aa], dB));
                                                delta = __fadd_rn(delta, __fmul_rn(t, M.subm[aa * 23 + lane]));
                                        }
                                        v1 = __fadd_rn(v1, delta);
                                }
                        } else if (p & 1) {
                                v0 = gap_adjust_val(b[lane], lane, p, (float)M.sipa, M.gpo, M.gpe, M.tgpe);
                                v1 = gap_adjust_val(b[lane + 32], lane + 32, p, (float)M.sipa, M.gpo, M.gpe, M.tgpe);
                        } else {
                                v0 = gap_adjust_val(a[lane], lane, p, (float)M.sipb, M.gpo, M.gpe, M.tgpe);
                                v1 = gap_adjust_val(a[lane + 32], lane + 32, p, (float)M.sipb, M.gpo, M.gpe, M.tgpe);
                        }
                }
                np[lane] = v0;
                np[lane + 32] = v1;
        }
}

} // namespace

int kb_make_profiles(kb200_ctx* ctx, const KbLeafProfile* d_leaves, int nleaves,
                     const long long* d_prefix, long long total_cols, const float* d_subm)
{
        if (nleaves <= 0) return KB200_OK;
        const long long warps = total_cols;
        int grid = (int)std::min<long long>((warps + 3) / 4, (long long)ctx->sm_count * 32);
        kb_make_profiles_kernel<<<grid, 128, 0, ctx->stream>>>(d_leaves, nleaves, d_prefix, total_cols, d_subm);
        KB_CUDA(cudaGetLastError());
        ctx->stats.n_launches++;
        return KB200_OK;
}

int kb_set_gap_penalties(kb200_ctx* ctx, const KbGapSet* d_sets, int nsets,
                         const long long* d_prefix, long long total_cols)
{
        if (nsets <= 0) return KB200_OK;
        int grid = (int)std::min<long long>((total_cols + 255) / 256, (long long)ctx->sm_count * 16);
        kb_set_gap_kernel<<<grid, 256, 0, ctx->stream>>>(d_sets, nsets, d_prefix, total_cols);
        KB_CUDA(cudaGetLastError());
        ctx->stats.n_launches++;
        return KB200_OK;
}

int kb_code_paths(kb200_ctx* ctx, const KbPathJob* d_pj, int njobs)
{
        if (njobs <= 0) return KB200_OK;
        kb_code_path_kernel<<<(njobs * 32 + 127) / 128, 128, 0, ctx->stream>>>(d_pj, njobs);
        KB_CUDA(cudaGetLastError());
        ctx->stats.n_launches++;
        return KB200_OK;
}

int kb_merge_index(kb200_ctx* ctx, const KbMergeJob* d_mj, int njobs)
{
        if (njobs <= 0) return KB200_OK;
        kb_merge_index_kernel<<<(njobs * 32 + 127) / 128, 128, 0, ctx->stream>>>(d_mj, njobs);
        KB_CUDA(cudaGetLastError());
        ctx->stats.n_launches++;
        return KB200_OK;
}

int kb_merge_profiles(kb200_ctx* ctx, const KbMergeJob* d_mj, int njobs,
                      const long long* d_prefix, long long total_cols)
{
        if (njobs <= 0) return KB200_OK;
        int grid = (int)std::min<long long>((total_cols + 3) / 4, (long long)ctx->sm_count * 32);
        kb_merge_kernel<<<grid, 128, 0, ctx->stream>>>(d_mj, njobs, d_prefix, total_cols);
        KB_CUDA(cudaGetLastError());
        ctx->stats.n_launches++;
        return KB200_OK;
}
