// kb_apair.cu -- N x N pairwise identity distances of a finished alignment
//
//   kb200_aln_pairwise_dist <-> compute_aln_pairwise_dist   lib/src/aln_apair_dist.c:9
//                               (pairwise_identity_dist     lib/src/aln_apair_dist.c:62)
//
// called by the realign loop (kalign_run_realign, lib/src/aln_wrap.c:455-490: align -> distances of
// the alignment -> new guide tree -> align again).  The reference walks every pair of aligned rows
// column by column on one thread: N^2/2 * alnlen byte compares (C3: 2e11).  Here the rows go to the
// device once; a CTA owns a 64 x 64 tile of pairs, streams the two row groups through shared memory
// 128 columns at a time and every thread keeps a 4 x 4 block of (matches, aligned) counters in
// registers.  Four columns are compared per 32-bit operation (zero-byte test of a ^ b), the
// "both rows hold a residue" count comes from one-bit-per-column residue masks (32 columns per
// AND + POPC).  The counters are exact integers; the distance is the reference's expression
// 1.0f - (float)matches / (float)aligned with an IEEE division, so the matrix is bit-identical.
// Integer / byte work, issue-bound on the LOP/POPC pipes; no tensor cores.
#include "kb_host.cuh"

#include <string.h>
#include <algorithm>

namespace {

constexpr int AP_TILE = 64;           // rows per tile side
constexpr int AP_WORDS = 32;          // 4-byte words per staged chunk (128 columns)
constexpr int AP_THREADS = 256;       // 16 x 16 threads, 4 x 4 pairs each
constexpr int AP_PITCH = AP_WORDS + 1;

// one bit per column: the row holds a residue there (anything but '-', aln_apair_dist.c:70)
__global__ void kb_apair_mask_kernel(const uint32_t* __restrict__ rows, int n, int words, int mwords, uint32_t* __restrict__ mask)
{
        const long long total = (long long)n * mwords;
        for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
                const int r = (int)(t / mwords);
                const int m = (int)(t % mwords);
                const uint32_t* src = rows + (size_t)r * (size_t)words + (size_t)m * 8;
                uint32_t bits = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                        const uint32_t w = (m * 8 + k < words) ? src[k] : 0x2d2d2d2du;
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                                if (((w >> (8 * b)) & 0xffu) != 0x2du) {
                                        bits |= 1u << (4 * k + b);
                                }
                        }
                }
                mask[t] = bits;
        }
}

// 0x80 in every byte of x that is zero (exact: no borrow between bytes)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t x)
{
        const uint32_t t = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;
        return ~(t | x | 0x7f7f7f7fu);
}

// tiles (bi <= bj) of the upper triangle; every thread writes dm[i][j] and dm[j][i]
__global__ void __launch_bounds__(AP_THREADS)
kb_apair_tile_kernel(const uint32_t* __restrict__ rows, const uint32_t* __restrict__ mask, int n, int words, int mwords,
                     int ntile, float* __restrict__ dm)
{
        __shared__ uint32_t s_a[AP_TILE][AP_PITCH];
        __shared__ uint32_t s_b[AP_TILE][AP_PITCH];
        __shared__ uint32_t s_ma[AP_TILE][AP_WORDS / 8 + 1];
        __shared__ uint32_t s_mb[AP_TILE][AP_WORDS / 8 + 1];

        // linear tile id -> (bi, bj), bi <= bj
        int bi = 0;
        {
                long long t = blockIdx.x;
                // row bi of the triangle holds ntile - bi tiles
                // solve by a short search from the closed form (float error corrected by the loops)
                double disc = (2.0 * ntile + 1.0) * (2.0 * ntile + 1.0) - 8.0 * (double)t;
                bi = (int)(((2.0 * ntile + 1.0) - sqrt(disc)) * 0.5);
                if (bi < 0) bi = 0;
                if (bi > ntile - 1) bi = ntile - 1;
                while (bi > 0 && (long long)bi * ntile - (long long)bi * (bi - 1) / 2 > t) bi--;
                while ((long long)(bi + 1) * ntile - (long long)(bi + 1) * bi / 2 <= t) bi++;
        }
        const int bj = bi + (int)((long long)blockIdx.x - ((long long)bi * ntile - (long long)bi * (bi - 1) / 2));
        const int i0 = bi * AP_TILE;
        const int j0 = bj * AP_TILE;
        const int tx = threadIdx.x & 15;
        const int ty = threadIdx.x >> 4;

        int matches[4][4];
        int aligned[4][4];
#pragma unroll
        for (int p = 0; p < 4; p++) {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                        matches[p][q] = 0;
                        aligned[p][q] = 0;
                }
        }

        for (int w0 = 0; w0 < words; w0 += AP_WORDS) {
                __syncthreads();
                // stage 64 x 32 words of either side; rows beyond n and words beyond the row read as gaps
                for (int t = threadIdx.x; t < AP_TILE * AP_WORDS; t += AP_THREADS) {
                        const int r = t / AP_WORDS;
                        const int w = t % AP_WORDS;
                        const bool inw = w0 + w < words;
                        s_a[r][w] = (inw && i0 + r < n) ? rows[(size_t)(i0 + r) * (size_t)words + (size_t)(w0 + w)] : 0x2d2d2d2du;
                        s_b[r][w] = (inw && j0 + r < n) ? rows[(size_t)(j0 + r) * (size_t)words + (size_t)(w0 + w)] : 0x2d2d2d2du;
                }
                for (int t = threadIdx.x; t < AP_TILE * (AP_WORDS / 8); t += AP_THREADS) {
                        const int r = t / (AP_WORDS / 8);
                        const int m = t % (AP_WORDS / 8);
                        const bool inm = w0 / 8 + m < mwords;
                        s_ma[r][m] = (inm && i0 + r < n) ? mask[(size_t)(i0 + r) * (size_t)mwords + (size_t)(w0 / 8 + m)] : 0u;
                        s_mb[r][m] = (inm && j0 + r < n) ? mask[(size_t)(j0 + r) * (size_t)mwords + (size_t)(w0 / 8 + m)] : 0u;
                }
                __syncthreads();
#pragma unroll
                for (int m = 0; m < AP_WORDS / 8; m++) {
                        uint32_t ma[4], mb[4];
#pragma unroll
                        for (int p = 0; p < 4; p++) {
                                ma[p] = s_ma[ty + 16 * p][m];
                                mb[p] = s_mb[tx + 16 * p][m];
                        }
#pragma unroll
                        for (int p = 0; p < 4; p++) {
#pragma unroll
                                for (int q = 0; q < 4; q++) {
                                        aligned[p][q] += __popc(ma[p] & mb[q]);
                                }
                        }
                }
#pragma unroll 4
                for (int w = 0; w < AP_WORDS; w++) {
                        uint32_t a[4], na[4], b[4];
#pragma unroll
                        for (int p = 0; p < 4; p++) {
                                a[p] = s_a[ty + 16 * p][w];
                                b[p] = s_b[tx + 16 * p][w];
                                // 0x80 where row a holds a residue: a == b there implies b holds one too
                                na[p] = ~zero_bytes(a[p] ^ 0x2d2d2d2du) & 0x80808080u;
                        }
#pragma unroll
                        for (int p = 0; p < 4; p++) {
#pragma unroll
                                for (int q = 0; q < 4; q++) {
                                        matches[p][q] += __popc(zero_bytes(a[p] ^ b[q]) & na[p]);
                                }
                        }
                }
        }
#pragma unroll
        for (int p = 0; p < 4; p++) {
                const int i = i0 + ty + 16 * p;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                        const int j = j0 + tx + 16 * q;
                        if (i < n && j < n) {
                                float d;
                                if (i == j) {
                                        d = 0.0f;                       // aln_apair_dist.c:25
                                } else if (aligned[p][q] == 0) {
                                        d = 1.0f;                       // aln_apair_dist.c:78-80
                                } else {
                                        d = __fsub_rn(1.0f, __fdiv_rn((float)matches[p][q], (float)aligned[p][q]));
                                }
                                dm[(size_t)i * (size_t)n + (size_t)j] = d;
                                dm[(size_t)j * (size_t)n + (size_t)i] = d;
                        }
                }
        }
}

}  // namespace

int kb200_aln_pairwise_dist(kb200_ctx* ctx, const char* const* rows, int n, int alnlen, float* const* dm_rows)
{
        if (!ctx || !rows || !dm_rows || n <= 0 || alnlen < 0) {
                return KB200_FAIL;
        }
        KB_CUDA(cudaSetDevice(ctx->device));
        const int words = std::max(1, (alnlen + 3) / 4);
        const int mwords = (words + 7) / 8;
        const size_t row_bytes = (size_t)words * 4;
        const size_t in_bytes = row_bytes * (size_t)n;
        const size_t out_bytes = (size_t)n * (size_t)n * sizeof(float);
        size_t free_b = 0, total_b = 0;
        KB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        if (out_bytes + in_bytes + (size_t)n * mwords * 4 > free_b - std::min(free_b, (size_t)1 << 30) + ctx->d_stage5.cap) {
                fprintf(stderr, "[kalign_b200] kb200_aln_pairwise_dist: %d x %d floats do not fit on the device\n", n, n);
                return KB200_FAIL;
        }
        // rows -> one pinned block (row pitch = words * 4, the tail padded with gap characters), one copy
        char* stage = (char*)kb_host_take(ctx, in_bytes);
        if (!stage) {
                return KB200_FAIL;
        }
        for (int i = 0; i < n; i++) {
                if (!rows[i] || !dm_rows[i]) {
                        kb_host_give(ctx, stage);
                        return KB200_FAIL;
                }
        }
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(kb_default_threads())
#endif
        for (int i = 0; i < n; i++) {
                memcpy(stage + (size_t)i * row_bytes, rows[i], (size_t)alnlen);
                memset(stage + (size_t)i * row_bytes + (size_t)alnlen, '-', row_bytes - (size_t)alnlen);
        }
        int rc = KB200_OK;
        if (ctx->d_stage3.ensure(in_bytes) != KB200_OK || ctx->d_stage4.ensure((size_t)n * mwords * 4) != KB200_OK ||
            ctx->d_stage5.ensure(out_bytes) != KB200_OK) {
                kb_host_give(ctx, stage);
                return KB200_FAIL;
        }
        cudaError_t e = cudaMemcpyAsync(ctx->d_stage3.p, stage, in_bytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) {
                cudaEventRecord(ctx->ev0, ctx->stream);
                const long long nm = (long long)n * mwords;
                const int mgrid = (int)std::min<long long>((nm + 255) / 256, (long long)ctx->sm_count * 8);
                kb_apair_mask_kernel<<<std::max(1, mgrid), 256, 0, ctx->stream>>>(ctx->d_stage3.as<uint32_t>(), n, words, mwords,
                                                                                  ctx->d_stage4.as<uint32_t>());
                const int ntile = (n + AP_TILE - 1) / AP_TILE;
                const long long tiles = (long long)ntile * (ntile + 1) / 2;
                if (tiles > 0x7fffffffLL) {
                        rc = KB200_FAIL;
                } else {
                        kb_apair_tile_kernel<<<(unsigned)tiles, AP_THREADS, 0, ctx->stream>>>(ctx->d_stage3.as<uint32_t>(),
                                                                                              ctx->d_stage4.as<uint32_t>(), n, words, mwords,
                                                                                              ntile, ctx->d_stage5.as<float>());
                }
                cudaEventRecord(ctx->ev1, ctx->stream);
                e = cudaGetLastError();
        }
        if (e == cudaSuccess) {
                e = cudaStreamSynchronize(ctx->stream);
        }
        if (e == cudaSuccess && rc == KB200_OK) {
                float ms = 0.0f;
                if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) {
                        ctx->stats.apair_seconds += 1e-3 * (double)ms;
                }
                ctx->stats.apair_col_pairs += 0.5 * (double)n * (double)(n - 1) * (double)alnlen;
                ctx->stats.n_launches += 2;
                ctx->stats.h2d_bytes += (double)in_bytes;
                ctx->stats.d2h_bytes += (double)out_bytes;
        }
        kb_host_give(ctx, stage);
        if (e != cudaSuccess || rc != KB200_OK) {
                fprintf(stderr, "[kalign_b200] kb200_aln_pairwise_dist: %s\n", e != cudaSuccess ? cudaGetErrorString(e) : "too many tiles");
                return KB200_FAIL;
        }
        // the caller's rows are separate allocations (float** dm of the reference): bands of rows come back
        // through two pinned blocks, copied out while the next band is in flight
        const size_t band_rows = std::max<size_t>(1, std::min<size_t>((size_t)n, ((size_t)64 << 20) / ((size_t)n * sizeof(float))));
        const size_t band_bytes = band_rows * (size_t)n * sizeof(float);
        float* pin[2] = {(float*)kb_host_take(ctx, band_bytes), nullptr};
        pin[1] = pin[0] ? (float*)kb_host_take(ctx, band_bytes) : nullptr;
        if (!pin[0] || !pin[1]) {
                if (pin[0]) kb_host_give(ctx, pin[0]);
                return KB200_FAIL;
        }
        const size_t nbands = ((size_t)n + band_rows - 1) / band_rows;
        auto band_copy = [&](size_t b) -> cudaError_t {
                const size_t r0 = b * band_rows;
                const size_t nr = std::min(band_rows, (size_t)n - r0);
                return cudaMemcpyAsync(pin[b & 1], ctx->d_stage5.as<float>() + r0 * (size_t)n, nr * (size_t)n * sizeof(float),
                                       cudaMemcpyDeviceToHost, ctx->stream);
        };
        e = band_copy(0);
        for (size_t b = 0; b < nbands && e == cudaSuccess; b++) {
                e = cudaStreamSynchronize(ctx->stream);
                if (e == cudaSuccess && b + 1 < nbands) {
                        e = band_copy(b + 1);
                }
                if (e == cudaSuccess) {
                        const size_t r0 = b * band_rows;
                        const size_t nr = std::min(band_rows, (size_t)n - r0);
                        const float* src = pin[b & 1];
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(kb_default_threads())
#endif
                        for (long long r = 0; r < (long long)nr; r++) {
                                memcpy(dm_rows[r0 + (size_t)r], src + (size_t)r * (size_t)n, (size_t)n * sizeof(float));
                        }
                }
        }
        if (e == cudaSuccess) {
                e = cudaStreamSynchronize(ctx->stream);
        }
        kb_host_give(ctx, pin[0]);
        kb_host_give(ctx, pin[1]);
        if (e != cudaSuccess) {
                fprintf(stderr, "[kalign_b200] kb200_aln_pairwise_dist: %s\n", cudaGetErrorString(e));
                return KB200_FAIL;
        }
        return KB200_OK;
}
