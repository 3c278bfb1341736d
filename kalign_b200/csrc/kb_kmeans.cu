// kb_kmeans.cu -- bisecting k-means of the guide tree on the device.
//
// Replaces the host loop of (behaviour cited, nothing copied):
//   bisecting_kmeans   lib/src/bisectingKmeans.c:273-406   (seeds in batches of four, early stop)
//   split2             lib/src/bisectingKmeans.c:766-971   (2-means on the anchor-distance rows)
//   edist_256          lib/src/euclidean_dist.c:161-206    (8-lane summation order)
//   cmp_floats         lib/src/bisectingKmeans.c:63-73
// and produces the same tree as kb_kmeans.h (the host restatement, kept for the CPU tests and for
// inputs with fewer than 50 sequences, where no bisection happens at all).
//
// What has to stay sequential for bit-identity is small: the three float sums of a Lloyd iteration
// (score, left centroid, right centroid) add their terms in sample order.  Everything else is
// parallel.  One CTA owns one (cluster, seed) pair and iterates to convergence without the host:
//   phase 1  every thread takes samples: both distances (one thread = one sample, so the 8-lane
//            order of edist_256 is just the order of a scalar loop), side, min distance
//   phase 2  warp 0: lane j accumulates dimension j of both centroids over the samples in order
//            (a coalesced 128-byte row per sample, loads software-pipelined, two independent add
//            chains); warp 1, lane 0: the score, in order
//   then the centroid update and the reference's epsilon convergence test.
// All clusters of one tree level x all their 40 seeds run in ONE launch; the host reads the scores
// back (one synchronisation per level), replays the reference's batches of four with its "stop when
// no seed of a batch improves" rule per cluster, keeps the winner's assignment, and stable-partitions
// every cluster's sample range (ids and distance rows) on the device: children keep sample order, as
// split2's sl / sr lists do, and every cluster's rows stay contiguous.
#include "kb_host.cuh"
#include "kb_kmeans_dev.h"

#include <string.h>
#include <algorithm>
#include <vector>

namespace {

constexpr int KM_DIM = 32;            // anchors = dimensions (bisection only happens for N >= 50 > 32)
constexpr int KM_THREADS = 1024;
constexpr int KM_SEEDS = 40;            // bisecting_kmeans: tries = MIN(40, num_samples)

struct KmJob {
        int begin, end;               // the cluster's range in order[]
        int seed;                     // local index of the seed sample
        int slot;                     // 0..39: which assignment buffer this seed writes
        float score;                  // out
        int num_l;                    // out
        int iters;                    // out (diagnostics)
        int pad;
};

__device__ __forceinline__ int km_cmp(const float a, const float b)       // cmp_floats
{
        const float epsilon = 1e-6;
        if (fabsf(__fadd_rn(a, -b)) < epsilon) return 0;
        return (a > b) ? 1 : -1;
}

// edist_256's order: 8 lanes accumulate (a-b)^2 over the 4 chunks of 8, then
// ((l0+l4)+(l1+l5)) + ((l2+l6)+(l3+l7)), then sqrtf
__device__ __forceinline__ float km_edist(const float* __restrict__ a, const float* __restrict__ b)
{
        float r[8];
#pragma unroll
        for (int l = 0; l < 8; l++) r[l] = 0.0f;
#pragma unroll
        for (int c = 0; c < KM_DIM; c += 8) {
#pragma unroll
                for (int l = 0; l < 8; l++) {
                        const float t = __fadd_rn(a[c + l], -b[c + l]);
                        r[l] = __fadd_rn(r[l], __fmul_rn(t, t));
                }
        }
        const float s0 = __fadd_rn(r[0], r[4]), s1 = __fadd_rn(r[1], r[5]);
        const float s2 = __fadd_rn(r[2], r[6]), s3 = __fadd_rn(r[3], r[7]);
        return __fsqrt_rn(__fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3)));
}

// The ordered sums of one pass over the cluster: rows (and, for a Lloyd iteration, the per-sample
// side / min distance) are streamed through shared memory in tiles of KM_TILE samples, loaded by
// warps 2.. while warp 0 (32 centroid dimensions, two add chains) and warp 1 (score) consume the
// previous tile in sample order -- the loads never sit on the add chains.
constexpr int KM_TILE = 256;
struct KmSmem {
        float rows[2][KM_TILE][KM_DIM];
        float dmn[2][KM_TILE];
        unsigned char side[2][KM_TILE];
};

template <bool WITH_SIDES>
__device__ __forceinline__ void km_ordered_sums(KmSmem& S, const float* __restrict__ rowsP, const unsigned char* sd, const float* dmn,
                                                const int ns, float* s_wl, float* s_wr, int* s_numl, float* s_score)
{
        const int tid = threadIdx.x;
        const int lane = tid & 31;
        const int warp = tid >> 5;
        const int ntiles = (ns + KM_TILE - 1) / KM_TILE;
        auto load_tile = [&](const int t, const int first_thread, const int nthreads) {
                const int i0 = t * KM_TILE;
                const int cnt = min(KM_TILE, ns - i0);
                const float4* __restrict__ src = reinterpret_cast<const float4*>(rowsP + (size_t)i0 * KM_DIM);
                float4* dst = reinterpret_cast<float4*>(&S.rows[t & 1][0][0]);
                for (int x = tid - first_thread; x < cnt * (KM_DIM / 4); x += nthreads) dst[x] = __ldg(src + x);
                if constexpr (WITH_SIDES) {
                        for (int x = tid - first_thread; x < cnt; x += nthreads) {
                                S.side[t & 1][x] = sd[i0 + x];
                                S.dmn[t & 1][x] = dmn[i0 + x];
                        }
                }
        };
        load_tile(0, 0, blockDim.x);
        float wl = 0.0f, wr = 0.0f, sc = 0.0f;
        int nl = 0;
        for (int t = 0; t < ntiles; t++) {
                __syncthreads();                 // tile t is in shared memory, tile t-1 has been consumed
                if (warp >= 2) {
                        if (t + 1 < ntiles) load_tile(t + 1, 64, blockDim.x - 64);
                } else if (warp == 0) {
                        const int cnt = min(KM_TILE, ns - t * KM_TILE);
                        const float* rp = &S.rows[t & 1][0][lane];
                        const unsigned char* sp = S.side[t & 1];
                        int i = 0;
                        for (; i + 8 <= cnt; i += 8) {
                                float v[8];
#pragma unroll
                                for (int q = 0; q < 8; q++) v[q] = rp[(i + q) * KM_DIM];
                                if constexpr (WITH_SIDES) {
                                        const unsigned long long m = *reinterpret_cast<const unsigned long long*>(sp + i);
#pragma unroll
                                        for (int q = 0; q < 8; q++) {
                                                if ((m >> (8 * q)) & 1ull) wr = __fadd_rn(wr, v[q]); else { wl = __fadd_rn(wl, v[q]); nl++; }
                                        }
                                } else {
#pragma unroll
                                        for (int q = 0; q < 8; q++) wl = __fadd_rn(wl, v[q]);
                                }
                        }
                        for (; i < cnt; i++) {
                                const float v = rp[i * KM_DIM];
                                if (WITH_SIDES && sp[i]) wr = __fadd_rn(wr, v); else { wl = __fadd_rn(wl, v); nl++; }
                        }
                } else if (WITH_SIDES && lane == 0) {
                        const int cnt = min(KM_TILE, ns - t * KM_TILE);
                        const float* dp = S.dmn[t & 1];
                        int i = 0;
                        for (; i + 8 <= cnt; i += 8) {
                                float v[8];
#pragma unroll
                                for (int q = 0; q < 8; q++) v[q] = dp[i + q];
#pragma unroll
                                for (int q = 0; q < 8; q++) sc = __fadd_rn(sc, v[q]);
                        }
                        for (; i < cnt; i++) sc = __fadd_rn(sc, dp[i]);
                }
        }
        if (warp == 0) {
                s_wl[lane] = wl;
                s_wr[lane] = wr;
                if (lane == 0) *s_numl = nl;
        } else if (warp == 1 && lane == 0) {
                *s_score = sc;
        }
        __syncthreads();
}

__global__ void __launch_bounds__(KM_THREADS)
kb_kmeans_seed_kernel(const float* __restrict__ rows_all /* N x 32, in order[] order */, KmJob* __restrict__ jobs,
                      const int N, unsigned char* __restrict__ side_all, float* __restrict__ dmin_all)
{
        extern __shared__ unsigned char km_raw[];
        KmSmem& S = *reinterpret_cast<KmSmem*>(km_raw);
        __shared__ float s_cl[KM_DIM], s_cr[KM_DIM], s_wl[KM_DIM], s_wr[KM_DIM];
        __shared__ int s_numl, s_stop;
        __shared__ float s_score;
        KmJob J = jobs[blockIdx.x];
        const int ns = J.end - J.begin;
        const float* __restrict__ rowsP = rows_all + (size_t)J.begin * KM_DIM;
        unsigned char* sd = side_all + (size_t)J.slot * (size_t)N + J.begin;
        float* dmn = dmin_all + (size_t)J.slot * (size_t)N + J.begin;
        const int tid = threadIdx.x;
        const int lane = tid & 31;
        const int warp = tid >> 5;
        // mean of the cluster's rows, summed in sample order (split2: w[j] += row[j]; w[j] /= ns)
        km_ordered_sums<false>(S, rowsP, sd, dmn, ns, s_wl, s_wr, &s_numl, &s_score);
        if (warp == 0) {
                const float w = __fdiv_rn(s_wl[lane], (float)ns);
                const float cl = __ldg(rowsP + (size_t)J.seed * KM_DIM + lane);
                s_cl[lane] = cl;
                s_cr[lane] = __fadd_rn(w, -__fadd_rn(cl, -w));           // w - (cl - w)
        }
        if (tid == 0) s_stop = 0;
        __syncthreads();
        int iters = 0;
        for (int it = 0; it < 500; it++) {
                iters = it + 1;
                // ---- phase 1: distances, side, min distance (parallel over samples) ----
                for (int i = tid; i < ns; i += (int)blockDim.x) {
                        const float4* __restrict__ rp = reinterpret_cast<const float4*>(rowsP + (size_t)i * KM_DIM);
                        float row[KM_DIM];
#pragma unroll
                        for (int q = 0; q < KM_DIM / 4; q++) {
                                const float4 x = __ldg(rp + q);
                                row[4 * q] = x.x; row[4 * q + 1] = x.y; row[4 * q + 2] = x.z; row[4 * q + 3] = x.w;
                        }
                        const float dl = km_edist(row, s_cl);
                        const float dr = km_edist(row, s_cr);
                        dmn[i] = (dl < dr) ? dl : dr;
                        const int r = km_cmp(dr, dl);
                        // dr < dl -> right; dr > dl -> left; tie -> odd samples right, even left (:879-898)
                        sd[i] = (unsigned char)((r == -1) ? 1 : ((r == 1) ? 0 : (i & 1)));
                }
                __syncthreads();
                // ---- phase 2: the three ordered sums ----
                km_ordered_sums<true>(S, rowsP, sd, dmn, ns, s_wl, s_wr, &s_numl, &s_score);
                // ---- centroid update and convergence (split2 :930-960) ----
                if (warp == 0) {
                        const int nl = s_numl, nr = ns - nl;
                        int stop = 0;
                        if (nl == 0 || nr == 0) {
                                stop = 2;                 // degenerate: halves by index, score 0
                        } else {
                                const float wl = __fdiv_rn(s_wl[lane], (float)nl);
                                const float wr = __fdiv_rn(s_wr[lane], (float)nr);
                                const int ch = (km_cmp(wl, s_cl[lane]) != 0) || (km_cmp(wr, s_cr[lane]) != 0);
                                const unsigned any = __ballot_sync(0xffffffffu, ch);
                                if (any == 0u) {
                                        stop = 1;
                                } else {
                                        s_cl[lane] = wl;
                                        s_cr[lane] = wr;
                                }
                        }
                        if (lane == 0) s_stop = stop;
                }
                __syncthreads();
                if (s_stop) break;
        }
        if (s_stop == 2) {
                for (int i = tid; i < ns; i += (int)blockDim.x) sd[i] = (unsigned char)((i < ns / 2) ? 0 : 1);
                if (tid == 0) { s_score = 0.0f; s_numl = ns / 2; }
                __syncthreads();
        }
        if (tid == 0) {
                jobs[blockIdx.x].score = s_score;
                jobs[blockIdx.x].num_l = s_numl;
                jobs[blockIdx.x].iters = iters;
        }
}

struct KmCopy { int begin, end, slot; };

// keep the winning seed's assignment of a cluster
__global__ void kb_kmeans_keep_kernel(const KmCopy* __restrict__ cp, const int N, const unsigned char* __restrict__ side_all,
                                      unsigned char* __restrict__ best)
{
        const KmCopy c = cp[blockIdx.x];
        const unsigned char* __restrict__ src = side_all + (size_t)c.slot * (size_t)N;
        for (int i = c.begin + threadIdx.x; i < c.end; i += blockDim.x) best[i] = src[i];
}

struct KmPart { int begin, end, num_l; };

// stable partition of a cluster's range: left samples in order, then right samples in order; the
// samples' distance rows move with them, so that every cluster's rows stay contiguous
__global__ void __launch_bounds__(256)
kb_kmeans_partition_kernel(const KmPart* __restrict__ parts, const unsigned char* __restrict__ best,
                           const int* __restrict__ in, int* __restrict__ out,
                           const float* __restrict__ rows_in, float* __restrict__ rows_out)
{
        __shared__ int s_warp[8];
        __shared__ int s_base;
        __shared__ int s_dst[256];
        const KmPart P = parts[blockIdx.x];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (threadIdx.x == 0) s_base = 0;
        __syncthreads();
        for (int t0 = P.begin; t0 < P.end; t0 += 256) {
                const int i = t0 + threadIdx.x;
                const bool in_range = i < P.end;
                const int isl = (in_range && best[i] == 0) ? 1 : 0;
                int incl = isl;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += v;
                }
                if (lane == 31) s_warp[warp] = incl;
                __syncthreads();
                int woff = 0;
                for (int w = 0; w < warp; w++) woff += s_warp[w];
                const int base = s_base;
                const int left_before = base + woff + incl - isl;          // left samples before i in the range
                int dst = -1;
                if (in_range) {
                        const int local = i - P.begin;
                        dst = isl ? (P.begin + left_before) : (P.begin + P.num_l + (local - left_before));
                        out[dst] = in[i];
                }
                s_dst[threadIdx.x] = dst;
                __syncthreads();
                if (threadIdx.x == 255) s_base = base + woff + incl;
                // rows: one warp per sample, one coalesced 128-byte move each
                const int cnt = min(256, P.end - t0);
                for (int x = warp; x < cnt; x += 8) {
                        rows_out[(size_t)s_dst[x] * KM_DIM + lane] = rows_in[(size_t)(t0 + x) * KM_DIM + lane];
                }
                __syncthreads();
        }
}

// UPGMA of the leaf clusters (upgma, bisectingKmeans.c:974-1053), one warp per cluster (n < 50):
// the n x n matrix lives in shared memory; every step finds the first minimum of the active upper
// triangle in row-major scan order (strict <, so ties go to the smallest i*n+j), records the merge and
// averages row / column node_a exactly as the reference does ((x + y) * 0.5F + 0.001F, two roundings).
// pd: distances of the cluster's pairs (i < j, row-major), as the batched bpm launch wrote them.
constexpr int UPGMA_MAXN = 50;
__global__ void __launch_bounds__(32)
kb_upgma_kernel(const float* __restrict__ pd, const long long* __restrict__ pair0, const int* __restrict__ csize,
                const long long* __restrict__ merge0, const int ncl, int2* __restrict__ merges)
{
        __shared__ float dm[UPGMA_MAXN * UPGMA_MAXN];
        const int c = blockIdx.x;
        if (c >= ncl) return;
        const int lane = threadIdx.x;
        const int n = csize[c];
        const float* __restrict__ src = pd + pair0[c];
        for (int x = lane; x < n * n; x += 32) {
                const int i = x / n, j = x % n;
                float v = 0.0f;
                if (i != j) {
                        const int a = (i < j) ? i : j, b = (i < j) ? j : i;
                        v = src[(long long)a * n - (long long)a * (a + 1) / 2 + (b - a - 1)];
                }
                dm[x] = v;
        }
        __syncwarp();
        unsigned long long active = (n >= 64) ? ~0ull : ((1ull << n) - 1ull);
        int2* __restrict__ out = merges + merge0[c];
        for (int step = 0; step < n - 1; step++) {
                float best = FLT_MAX;
                int bidx = 0x7fffffff;
                for (int x = lane; x < n * n; x += 32) {
                        const int i = x / n, j = x % n;
                        if (i < j && ((active >> i) & 1ull) && ((active >> j) & 1ull)) {
                                const float v = dm[x];
                                if (v < best) { best = v; bidx = x; }       // ascending x per lane: first minimum kept
                        }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
                        if (ov < best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
                }
                // the reference keeps node_a / node_b of the previous step when nothing is below FLT_MAX; with
                // finite distances there always is a minimum
                const int na = bidx / n, nb = bidx % n;
                if (lane == 0) out[step] = make_int2(na, nb);
                for (int j = lane; j < n; j += 32) {
                        if (j != nb) {
                                dm[na * n + j] = __fadd_rn(__fmul_rn(__fadd_rn(dm[na * n + j], dm[nb * n + j]), 0.5F), 0.001F);
                        }
                }
                __syncwarp();
                if (lane == 0) dm[na * n + na] = 0.0F;
                __syncwarp();
                for (int j = lane; j < n; j += 32) {
                        dm[j * n + na] = dm[na * n + j];
                }
                __syncwarp();
                active &= ~(1ull << nb);
        }
}

struct HCluster {
        int begin, end;
        int node;              // index in the builder's node vector
        // seed search state
        int next_i = 0;
        bool have_best = false;
        float best_score = 0.0f;
        int best_numl = 0;
        bool done = false;
};

} // namespace

// UPGMA of all leaf clusters in one launch.  d_pd: the pair distances on the device (cluster after
// cluster, i < j row-major); csize / pair0 / merge0 per cluster (host); merges_out: sum (n-1) int2.
int kb_upgma_dev(kb200_ctx* ctx, const float* d_pd, const std::vector<int>& csize, const std::vector<long long>& pair0,
                 const std::vector<long long>& merge0, long long nmerges, int* merges_out)
{
        const int ncl = (int)csize.size();
        if (ncl == 0 || nmerges == 0) return KB200_OK;
        for (int n : csize) {
                if (n >= UPGMA_MAXN) {
                        fprintf(stderr, "[kalign_b200] UPGMA: cluster of %d sequences (max %d)\n", n, UPGMA_MAXN - 1);
                        return KB200_FAIL;
                }
        }
        cudaStream_t st = ctx->stream;
        const size_t bytes = sizeof(int) * (size_t)ncl + sizeof(long long) * 2 * (size_t)ncl + 64;
        KB_RUN(ctx->km_desc.ensure(bytes + sizeof(int2) * (size_t)nmerges + 64));
        char* base = ctx->km_desc.as<char>();
        long long* d_pair0 = (long long*)base;
        long long* d_merge0 = d_pair0 + ncl;
        int* d_csize = (int*)(d_merge0 + ncl);
        int2* d_merges = (int2*)(base + ((bytes + 15) & ~(size_t)15));
        KB_RUN(kb_h2d(ctx, d_pair0, pair0.data(), sizeof(long long) * (size_t)ncl));
        KB_RUN(kb_h2d(ctx, d_merge0, merge0.data(), sizeof(long long) * (size_t)ncl));
        KB_RUN(kb_h2d(ctx, d_csize, csize.data(), sizeof(int) * (size_t)ncl));
        kb_upgma_kernel<<<ncl, 32, 0, st>>>(d_pd, d_pair0, d_csize, d_merge0, ncl, d_merges);
        KB_CUDA(cudaGetLastError());
        ctx->stats.n_launches++;
        KB_CUDA(cudaMemcpyAsync(merges_out, d_merges, sizeof(int2) * (size_t)nmerges, cudaMemcpyDeviceToHost, st));
        KB_CUDA(cudaStreamSynchronize(st));
        ctx->pinned.reset();
        return KB200_OK;
}

// Device version of kb_tree_bisect (kb_msa.cu): the k-means tree as plain arrays (KbKmeansTree).
// dm_host: N x 32 floats.  Uses its own stream, so that it can run beside the anchor batch of the
// engine's stream.
int kb_kmeans_bisect_dev(kb200_ctx* ctx, const float* dm_host, int N, KbKmeansTree& B)
{
        KB_CUDA(cudaSetDevice(ctx->device));
        int root_store = -1;
        int* root_out = &root_store;
        // scratch and stream live in the context (no cudaMalloc / cudaFree per call)
        if (!ctx->stream2) KB_CUDA(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
        cudaStream_t st = ctx->stream2;
        KB_RUN(ctx->km_rowsA.ensure(sizeof(float) * (size_t)N * KM_DIM));
        KB_RUN(ctx->km_rowsB.ensure(sizeof(float) * (size_t)N * KM_DIM));
        KB_RUN(ctx->km_ordA.ensure(sizeof(int) * (size_t)N));
        KB_RUN(ctx->km_ordB.ensure(sizeof(int) * (size_t)N));
        KB_RUN(ctx->km_side.ensure((size_t)KM_SEEDS * (size_t)N));
        KB_RUN(ctx->km_best.ensure((size_t)N));
        KB_RUN(ctx->km_dmin.ensure(sizeof(float) * KM_SEEDS * (size_t)N));
        float *d_dm = ctx->km_rowsA.as<float>(), *d_dmB = ctx->km_rowsB.as<float>();     // distance rows in order[] order (ping-pong with the ids)
        int *d_ordA = ctx->km_ordA.as<int>(), *d_ordB = ctx->km_ordB.as<int>();
        unsigned char *d_side = ctx->km_side.as<unsigned char>(), *d_best = ctx->km_best.as<unsigned char>();
        float* d_dmin = ctx->km_dmin.as<float>();
        void* d_desc = nullptr;
        int rc = KB200_OK;
        auto fail = [&](const char* what) {
                fprintf(stderr, "[kalign_b200] device k-means: %s failed: %s\n", what, cudaGetErrorString(cudaGetLastError()));
                rc = KB200_FAIL;
        };
#define KM_TRY(x, what) do { if (rc == KB200_OK && (x) != cudaSuccess) fail(what); } while (0)
        KM_TRY(cudaFuncSetAttribute(kb_kmeans_seed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KmSmem)), "cudaFuncSetAttribute");
        std::vector<int> iota((size_t)N);
        for (int i = 0; i < N; i++) iota[(size_t)i] = i;
        KM_TRY(cudaMemcpyAsync(d_dm, dm_host, sizeof(float) * (size_t)N * KM_DIM, cudaMemcpyHostToDevice, st), "H2D");
        KM_TRY(cudaMemcpyAsync(d_ordA, iota.data(), sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, st), "H2D");
        auto ensure_desc = [&](size_t bytes) {
                if (ctx->km_desc.ensure(bytes * 2 + 4096) != KB200_OK) rc = KB200_FAIL;
                d_desc = ctx->km_desc.p;
        };
        // the k-means tree: internal nodes get their children when the level below is known
        B.left.clear();
        B.right.clear();
        B.leaf_node.clear();
        B.leaf_begin.clear();
        B.leaf_end.clear();
        struct Leaf { int begin, end, node; };
        std::vector<Leaf> leaves;
        auto new_node = [&]() -> int {
                B.left.push_back(-1);
                B.right.push_back(-1);
                return (int)B.left.size() - 1;
        };
        std::vector<HCluster> level;
        long long launches = 0;
        {
                HCluster r;
                r.begin = 0; r.end = N;
                r.node = new_node();
                *root_out = r.node;
                if (N < 50) {
                        leaves.push_back({0, N, r.node});
                } else {
                        level.push_back(r);
                }
        }
        int* cur = d_ordA;
        int* nxt = d_ordB;
        float* rcur = d_dm;
        float* rnxt = d_dmB;
        std::vector<KmJob> jobs;
        std::vector<KmCopy> copies;
        std::vector<KmPart> parts;
        while (rc == KB200_OK && !level.empty()) {
                // ---- seeds.  The reference tries them in batches of four and stops after the first batch
                //      that does not improve (bisecting_kmeans :319-362).  Running later batches
                //      speculatively does not change which seed wins (the host replays the batches in
                //      order), so a level with few, large clusters -- where the latency of one seed is what
                //      costs -- gets all 40 seeds in ONE launch; a level with many small clusters gets two
                //      batches per launch and stops per cluster, because there the total work is what costs.
                int largest = 0;
                for (const HCluster& C : level) largest = std::max(largest, C.end - C.begin);
                const int per_launch = (level.size() <= 16) ? KM_SEEDS : 8;
                const int threads = (largest > 4096) ? KM_THREADS : 256;
                long long it_all = 0, it_max = 0, njobs_all = 0;
                while (rc == KB200_OK) {
                        jobs.clear();
                        std::vector<int> owner;
                        for (size_t c = 0; c < level.size(); c++) {
                                HCluster& C = level[c];
                                if (C.done) continue;
                                const int ns = C.end - C.begin;
                                const int tries = std::min(KM_SEEDS, ns);
                                const int step = ns / tries;
                                for (int j = 0; j < per_launch && C.next_i + j < tries; j++) {
                                        KmJob J;
                                        J.begin = C.begin; J.end = C.end;
                                        J.seed = (C.next_i + j) * step;
                                        J.slot = j;
                                        J.score = 0.0f; J.num_l = 0; J.iters = 0; J.pad = 0;
                                        jobs.push_back(J);
                                        owner.push_back((int)c);
                                }
                        }
                        if (jobs.empty()) break;
                        ensure_desc(sizeof(KmJob) * jobs.size());
                        KM_TRY(cudaMemcpyAsync(d_desc, jobs.data(), sizeof(KmJob) * jobs.size(), cudaMemcpyHostToDevice, st), "H2D");
                        if (rc != KB200_OK) break;
                        kb_kmeans_seed_kernel<<<(unsigned)jobs.size(), threads, sizeof(KmSmem), st>>>(rcur, (KmJob*)d_desc, N, d_side, d_dmin);
                        launches++;
                        KM_TRY(cudaGetLastError(), "seed kernel launch");
                        KM_TRY(cudaMemcpyAsync(jobs.data(), d_desc, sizeof(KmJob) * jobs.size(), cudaMemcpyDeviceToHost, st), "D2H");
                        KM_TRY(cudaStreamSynchronize(st), "seed kernel");
                        if (rc != KB200_OK) break;
                        for (const KmJob& J : jobs) { it_all += J.iters; it_max = std::max<long long>(it_max, J.iters); }
                        njobs_all += (long long)jobs.size();
                        copies.clear();
                        size_t q = 0;
                        while (q < jobs.size()) {
                                HCluster& C = level[(size_t)owner[q]];
                                size_t q1 = q;
                                while (q1 < jobs.size() && owner[q1] == owner[q]) q1++;
                                const int tries = std::min(KM_SEEDS, C.end - C.begin);
                                int winner = -1;
                                for (size_t b0 = q; b0 < q1 && !C.done; b0 += 4) {
                                        int change = 0;
                                        for (size_t x = b0; x < b0 + 4 && x < q1; x++) {
                                                const KmJob& J = jobs[x];
                                                if (!C.have_best) {
                                                        C.have_best = true;
                                                        C.best_score = J.score; C.best_numl = J.num_l;
                                                        winner = J.slot; change++;
                                                } else if (C.best_score > J.score) {
                                                        C.best_score = J.score; C.best_numl = J.num_l;
                                                        winner = J.slot; change++;
                                                }
                                        }
                                        C.next_i += 4;
                                        if (!change || C.next_i >= tries) C.done = true;
                                }
                                if (winner >= 0) copies.push_back({C.begin, C.end, winner});
                                q = q1;
                        }
                        if (!copies.empty()) {
                                ensure_desc(sizeof(KmCopy) * copies.size());
                                KM_TRY(cudaMemcpyAsync(d_desc, copies.data(), sizeof(KmCopy) * copies.size(), cudaMemcpyHostToDevice, st), "H2D");
                                if (rc != KB200_OK) break;
                                kb_kmeans_keep_kernel<<<(unsigned)copies.size(), 256, 0, st>>>((const KmCopy*)d_desc, N, d_side, d_best);
                                launches++;
                                KM_TRY(cudaGetLastError(), "keep kernel launch");
                                KM_TRY(cudaStreamSynchronize(st), "keep kernel");      // d_desc is reused below
                        }
                }
                if (rc == KB200_OK && getenv("KB200_TRACE")) {
                        fprintf(stderr, "[kb200 trace] k-means level: %zu clusters (largest %d), %lld seed jobs, iterations mean %.1f max %lld\n",
                                level.size(), largest, njobs_all, njobs_all ? (double)it_all / (double)njobs_all : 0.0, it_max);
                }
                if (rc != KB200_OK) break;
                // ---- partition every cluster; children become the next level or leaf clusters ----
                parts.clear();
                for (const HCluster& C : level) parts.push_back({C.begin, C.end, C.best_numl});
                ensure_desc(sizeof(KmPart) * parts.size());
                KM_TRY(cudaMemcpyAsync(d_desc, parts.data(), sizeof(KmPart) * parts.size(), cudaMemcpyHostToDevice, st), "H2D");
                // ranges outside this level's clusters (finished leaves) keep their order
                KM_TRY(cudaMemcpyAsync(nxt, cur, sizeof(int) * (size_t)N, cudaMemcpyDeviceToDevice, st), "D2D");
                KM_TRY(cudaMemcpyAsync(rnxt, rcur, sizeof(float) * (size_t)N * KM_DIM, cudaMemcpyDeviceToDevice, st), "D2D");
                if (rc != KB200_OK) break;
                kb_kmeans_partition_kernel<<<(unsigned)parts.size(), 256, 0, st>>>((const KmPart*)d_desc, d_best, cur, nxt, rcur, rnxt);
                launches++;
                KM_TRY(cudaGetLastError(), "partition kernel launch");
                KM_TRY(cudaStreamSynchronize(st), "partition kernel");
                std::swap(cur, nxt);
                std::swap(rcur, rnxt);
                std::vector<HCluster> next;
                for (const HCluster& C : level) {
                        const int mids = C.begin + C.best_numl;
                        const int rng[2][2] = {{C.begin, mids}, {mids, C.end}};
                        int child[2];
                        for (int s = 0; s < 2; s++) {
                                child[s] = new_node();
                                if (rng[s][1] - rng[s][0] < 50) {
                                        leaves.push_back({rng[s][0], rng[s][1], child[s]});
                                } else {
                                        HCluster H;
                                        H.begin = rng[s][0]; H.end = rng[s][1]; H.node = child[s];
                                        next.push_back(H);
                                }
                        }
                        B.left[(size_t)C.node] = child[0];
                        B.right[(size_t)C.node] = child[1];
                }
                level.swap(next);
        }
        B.order.assign((size_t)N, 0);
        KM_TRY(cudaMemcpyAsync(B.order.data(), cur, sizeof(int) * (size_t)N, cudaMemcpyDeviceToHost, st), "D2H");
        KM_TRY(cudaStreamSynchronize(st), "final copy");
        if (rc == KB200_OK) {
                for (const Leaf& L : leaves) {
                        B.leaf_node.push_back(L.node);
                        B.leaf_begin.push_back(L.begin);
                        B.leaf_end.push_back(L.end);
                }
                B.root = root_store;
        }
#undef KM_TRY
        (void)launches;
        return rc;
}
