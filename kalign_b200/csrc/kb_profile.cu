// kb_profile.cu -- streaming profile kernels and path post-processing (device resident between
// guide-tree levels).
//
// Replaces (behaviour cited, nothing copied):
//   make_profile_n         lib/src/aln_setup.c:40    leaf profile, (len+2) x 64 floats
//   set_gap_penalties_n    lib/src/aln_setup.c:101   [27..29] = [55..57] * nsip(other operand)
//   update_n               lib/src/aln_setup.c:230   merge along the coded path (no seq weights)
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

// raw path -> (mirror) -> coded path (+ optional position map).  One thread per job: the walk is
// inherently serial and O(len); jobs are independent.
__global__ void kb_code_path_kernel(const KbPathJob* __restrict__ pj, const int njobs)
{
        const int j = blockIdx.x * blockDim.x + threadIdx.x;
        if (j >= njobs) {
                return;
        }
        const KbPathJob P = pj[j];
        const int len_a = P.len_a, len_b = P.len_b;
        const int n = len_a + len_b + 2;
        const int* raw = P.raw;
        int* o = P.coded;
        if (P.mirror) {
                // raw was produced with rows = b (len_b entries); scratch receives the mirrored path
                int* mr = P.scratch;
                for (int i = 0; i < len_a + 2; i++) {
                        mr[i] = -1;
                }
                for (int i = 1; i <= len_b; i++) {
                        const int c = raw[i];
                        if (c != -1) {
                                mr[c] = i;
                        }
                }
                raw = mr;
        }
        for (int i = 0; i < n; i++) {
                o[i] = 0;
        }
        int jj = 1;
        int b = -1;
        for (int i = 1; i <= len_a; i++) {
                const int r = raw[i];
                if (r == -1) {
                        o[jj++] = 2;
                } else {
                        int skip;
                        if (i == 1) {
                                skip = r - 1;
                        } else if (r - 1 != b && b != -1) {
                                skip = r - b - 1;
                        } else {
                                skip = 0;
                        }
                        for (int a = 0; a < skip; a++) {
                                o[jj++] = 1;
                        }
                        o[jj++] = 0;
                }
                b = r;
        }
        {
                const int last = raw[len_a];
                if (last < len_b && last != -1) {
                        for (int a = 0; a < len_b - last; a++) {
                                o[jj++] = 1;
                        }
                }
        }
        o[0] = jj - 1;
        o[jj] = 3;
        // (the reference's open/ext/close flag loop never runs, aln_setup.c:191-195)
        int i = 1;
        while (i < n && o[i] != 0) {
                o[i] |= 32;
                i++;
        }
        i = o[0];
        while (i > 0 && o[i] != 0) {
                o[i] |= 32;
                i--;
        }
        if (P.posmap) {
                int* pm = P.posmap;
                const int len_i = len_a;
                for (int c = 0; c < len_i; c++) {
                        pm[c] = -1;
                }
                int pos_a = 0, pos_b = 0;
                for (int c = 1; o[c] != 3; c++) {
                        const int v = o[c];
                        if (v == 0) {
                                if (pos_a < len_i) pm[pos_a] = pos_b;
                                pos_a++; pos_b++;
                        } else if (v & 1) {
                                pos_b++;
                        } else if (v & 2) {
                                if (pos_a < len_i) pm[pos_a] = -1;
                                pos_a++;
                        }
                }
        }
}

// per merge: source column indices of every output column (prefix over the coded path), one
// thread per job (serial, O(len)), then one warp per output column does the 64-float merge.
__global__ void kb_merge_index_kernel(const KbMergeJob* __restrict__ mj, const int njobs)
{
        const int j = blockIdx.x * blockDim.x + threadIdx.x;
        if (j >= njobs) {
                return;
        }
        const KbMergeJob M = mj[j];
        const int* path = M.path;
        int2* src = M.src;
        int ia = 1, ib = 1;
        int c = 1;
        for (; path[c] != 3; c++) {
                const int p = path[c];
                src[c] = make_int2(ia, ib);
                if (!p) {
                        ia++; ib++;
                } else {
                        if (p & 1) ib++;
                        if (p & 2) ia++;
                }
        }
        src[c] = make_int2(ia, ib);      // last boundary column
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
                if (c == 0) {
                        v0 = M.pa[lane] + M.pb[lane];
                        v1 = M.pa[lane + 32] + M.pb[lane + 32];
                } else {
                        const int2 s = M.src[c];
                        const float* a = M.pa + ((size_t)s.x << 6);
                        const float* b = M.pb + ((size_t)s.y << 6);
                        const int p = (c <= alnlen) ? M.path[c] : 0;
                        if (c > alnlen || !p) {
                                v0 = a[lane] + b[lane];
                                v1 = a[lane + 32] + b[lane + 32];
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
        kb_code_path_kernel<<<(njobs + 63) / 64, 64, 0, ctx->stream>>>(d_pj, njobs);
        KB_CUDA(cudaGetLastError());
        ctx->stats.n_launches++;
        return KB200_OK;
}

int kb_merge_index(kb200_ctx* ctx, const KbMergeJob* d_mj, int njobs)
{
        if (njobs <= 0) return KB200_OK;
        kb_merge_index_kernel<<<(njobs + 63) / 64, 64, 0, ctx->stream>>>(d_mj, njobs);
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
