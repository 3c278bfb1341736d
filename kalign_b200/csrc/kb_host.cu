// kb_host.cu -- host orchestration of the distance matrix and the anchor-consistency batch.
//
//   kb200_distances      <-> d_estimation             lib/src/sequence_distance.c:37
//   kb200_anchor_posmaps <-> anchor_consistency_build lib/src/anchor_consistency.c:200
//                            (pairwise_align_map :19, the serial N x K loop :246-267)
#include "kb_host.cuh"

#include <string.h>
#include <algorithm>
#include <chrono>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int kb_default_threads()
{
        // hardware threads visible to the process, NOT omp_get_max_threads(): launchers such as
        // torchrun export OMP_NUM_THREADS=1; the num_threads() clauses used here override that.
        int n = 1;
#ifdef _OPENMP
        n = omp_get_num_procs();
#endif
        FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r");
        if (f) {
                char q[64];
                long long period = 0;
                if (fscanf(f, "%63s %lld", q, &period) == 2 && period > 0 && strcmp(q, "max") != 0) {
                        const long long quota = atoll(q);
                        if (quota > 0) {
                                n = std::min<long long>(n, std::max<long long>(1, quota / period));
                        }
                }
                fclose(f);
        }
        return std::max(1, std::min(n, 16));
}

// smallest pooled buffer that is large enough -- and not absurdly larger, so that a 4 KB request does not
// walk away with the 60 MB buffer the next request needs.  When nothing fits the pool is left alone and
// the caller's ensure() allocates: after the first call or two every size a workload uses is in the
// pool, and no later call pays for cudaMalloc / cudaFree (both synchronise the device; measured stalls of
// 50-400 ms per call when buffers were re-allocated every time).
void kb_take_pooled(kb200_ctx* ctx, KbDevBuf& b, size_t bytes)
{
        if (b.p || ctx->seq_pool.empty()) {
                return;
        }
        int best = -1;
        for (size_t i = 0; i < ctx->seq_pool.size(); i++) {
                const size_t c = ctx->seq_pool[i].cap;
                if (c >= bytes && c <= 4 * bytes + ((size_t)1 << 20) && (best < 0 || c < ctx->seq_pool[(size_t)best].cap)) {
                        best = (int)i;
                }
        }
        if (best < 0) {
                return;
        }
        b = ctx->seq_pool[(size_t)best];
        ctx->seq_pool.erase(ctx->seq_pool.begin() + (long)best);
}

// hand a device buffer back to the context's pool
void kb_give_pooled(kb200_ctx* ctx, KbDevBuf& b)
{
        if (!b.p) {
                return;
        }
        if (ctx && ctx->seq_pool.size() < 64) {
                ctx->seq_pool.push_back(b);
                b.p = nullptr;
                b.cap = 0;
        } else {
                b.release();
        }
}

void KbSeqs::release()
{
        KbDevBuf* bufs[3] = {&d_seqs, &d_offs, &d_lens};
        for (KbDevBuf* b : bufs) {
                kb_give_pooled(owner, *b);
        }
}

int KbSeqs::alloc(kb200_ctx* ctx, const int64_t* offs, const int* lens, int nseq)
{
        h_seqs = nullptr;
        h_offs = offs;
        h_lens = lens;
        n = nseq;
        total = 0;
        for (int i = 0; i < nseq; i++) {
                total = std::max<int64_t>(total, offs[i] + lens[i]);
        }
        owner = ctx;
        kb_take_pooled(ctx, d_seqs, (size_t)total + 16);
        kb_take_pooled(ctx, d_offs, sizeof(int64_t) * (size_t)nseq);
        kb_take_pooled(ctx, d_lens, sizeof(int) * (size_t)nseq);
        KB_RUN(d_seqs.ensure((size_t)total + 16));
        KB_RUN(d_offs.ensure(sizeof(int64_t) * (size_t)nseq));
        KB_RUN(d_lens.ensure(sizeof(int) * (size_t)nseq));
        KB_CUDA(cudaMemcpyAsync(d_offs.p, offs, sizeof(int64_t) * (size_t)nseq, cudaMemcpyHostToDevice, ctx->stream));
        KB_CUDA(cudaMemcpyAsync(d_lens.p, lens, sizeof(int) * (size_t)nseq, cudaMemcpyHostToDevice, ctx->stream));
        KB_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->stats.h2d_bytes += 12.0 * nseq;
        return KB200_OK;
}

int KbSeqs::upload(kb200_ctx* ctx, const uint8_t* seqs, const int64_t* offs, const int* lens, int nseq)
{
        h_seqs = seqs;
        h_offs = offs;
        h_lens = lens;
        n = nseq;
        total = 0;
        for (int i = 0; i < nseq; i++) {
                total = std::max<int64_t>(total, offs[i] + lens[i]);
        }
        owner = ctx;
        kb_take_pooled(ctx, d_seqs, (size_t)total + 16);
        kb_take_pooled(ctx, d_offs, sizeof(int64_t) * (size_t)nseq);
        kb_take_pooled(ctx, d_lens, sizeof(int) * (size_t)nseq);
        KB_RUN(d_seqs.ensure((size_t)total + 16));
        KB_RUN(d_offs.ensure(sizeof(int64_t) * (size_t)nseq));
        KB_RUN(d_lens.ensure(sizeof(int) * (size_t)nseq));
        KB_CUDA(cudaMemcpyAsync(d_seqs.p, seqs, (size_t)total, cudaMemcpyHostToDevice, ctx->stream));
        KB_CUDA(cudaMemcpyAsync(d_offs.p, offs, sizeof(int64_t) * (size_t)nseq, cudaMemcpyHostToDevice, ctx->stream));
        KB_CUDA(cudaMemcpyAsync(d_lens.p, lens, sizeof(int) * (size_t)nseq, cudaMemcpyHostToDevice, ctx->stream));
        KB_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->stats.h2d_bytes += (double)total + 12.0 * nseq;
        return KB200_OK;
}

void* kb_host_take(kb200_ctx* ctx, size_t bytes)
{
        int best = -1;
        for (size_t i = 0; i < ctx->host_pool.size(); i++) {
                const kb200_ctx::PinnedBlock& b = ctx->host_pool[i];
                if (!b.used && b.cap >= bytes && (best < 0 || b.cap < ctx->host_pool[(size_t)best].cap)) best = (int)i;
        }
        if (best >= 0) {
                ctx->host_pool[(size_t)best].used = true;
                return ctx->host_pool[(size_t)best].p;
        }
        // nothing fits: allocate (cudaFreeHost / cudaMallocHost synchronise the device, so free blocks are only
        // recycled when the pool has grown unreasonably)
        if (ctx->host_pool.size() >= 32) {
                for (size_t i = 0; i < ctx->host_pool.size(); i++) {
                        if (!ctx->host_pool[i].used) {
                                cudaFreeHost(ctx->host_pool[i].p);
                                ctx->host_pool.erase(ctx->host_pool.begin() + (long)i);
                                break;
                        }
                }
        }
        void* p = nullptr;
        const size_t want = bytes + bytes / 8 + 4096;
        if (cudaMallocHost(&p, want) != cudaSuccess) {
                cudaGetLastError();
                fprintf(stderr, "[kalign_b200] cudaMallocHost(%zu) failed\n", want);
                return nullptr;
        }
        ctx->host_pool.push_back({p, want, true});
        return p;
}

void kb_host_give(kb200_ctx* ctx, void* p)
{
        for (kb200_ctx::PinnedBlock& b : ctx->host_pool) {
                if (b.p == p) b.used = false;
        }
}

float* KbArena::alloc_floats(size_t n)
{
        const size_t bytes = (n * sizeof(float) + 255) & ~(size_t)255;
        while (true) {
                if (cur < chunks.size()) {
                        if (used + bytes <= caps[cur]) {
                                float* r = (float*)((char*)chunks[cur] + used);
                                used += bytes;
                                return r;
                        }
                        cur++;
                        used = 0;
                        continue;
                }
                const size_t want = std::max(chunk_bytes, bytes);
                void* p = nullptr;
                if (cudaMalloc(&p, want) != cudaSuccess) {
                        fprintf(stderr, "[kalign_b200] arena: cudaMalloc(%zu) failed\n", want);
                        cudaGetLastError();
                        return nullptr;
                }
                chunks.push_back(p);
                caps.push_back(want);
        }
}

void KbArena::release()
{
        for (void* p : chunks) {
                cudaFree(p);
        }
        chunks.clear();
        caps.clear();
        cur = used = 0;
}

// ---------------------------------------------------------------------------------------------

int kb_distances_dev(kb200_ctx* ctx, KbSeqs& S, const int* rows, int nrows, const int* cols, int ncols,
                     int explicit_pairs, float* dm_host)
{
        if (nrows <= 0) {
                return KB200_OK;
        }
        const size_t ncol_items = explicit_pairs ? (size_t)nrows : (size_t)ncols;
        const size_t npairs = explicit_pairs ? (size_t)nrows : (size_t)nrows * (size_t)ncols;
        if (npairs == 0) {
                return KB200_OK;
        }
        // longest pattern any pair can have: min(longest row, longest col), capped at 1024
        int maxr = 0, maxc = 0;
        for (int i = 0; i < nrows; i++) {
                maxr = std::max(maxr, S.h_lens[rows[i]]);
        }
        for (size_t i = 0; i < ncol_items; i++) {
                maxc = std::max(maxc, S.h_lens[cols[i]]);
        }
        int m = std::min(std::min(maxr, maxc), 1024);
        if (explicit_pairs) {
                m = 0;
                for (int i = 0; i < nrows; i++) {
                        m = std::max(m, std::min(S.h_lens[rows[i]], S.h_lens[cols[i]]));
                }
                m = std::min(m, 1024);
        }
        const int words = std::max(1, (m + 63) / 64);
        KB_RUN(ctx->d_stage4.ensure(sizeof(int) * ((size_t)nrows + ncol_items)));
        KB_RUN(ctx->d_stage5.ensure(sizeof(float) * npairs));
        int* d_rows = ctx->d_stage4.as<int>();
        int* d_cols = d_rows + nrows;
        KB_CUDA(cudaMemcpyAsync(d_rows, rows, sizeof(int) * (size_t)nrows, cudaMemcpyHostToDevice, ctx->stream));
        KB_CUDA(cudaMemcpyAsync(d_cols, cols, sizeof(int) * ncol_items, cudaMemcpyHostToDevice, ctx->stream));
        if (ctx->world > 1 && !explicit_pairs && nrows >= 8 * ctx->world) {
                // multi-GPU: the rows of the matrix are split contiguously across the ranks
                // (sequence_distance.c:108 is an `omp parallel for` over the same rows), every rank writes
                // its rows at their final offsets, one all-gather-v makes the matrix complete everywhere
                std::vector<size_t> seg((size_t)ctx->world + 1);
                for (int r = 0; r <= ctx->world; r++) {
                        seg[(size_t)r] = (size_t)((long long)nrows * r / ctx->world) * (size_t)ncols * sizeof(float);
                }
                const int r0 = (int)((long long)nrows * ctx->rank / ctx->world);
                const int r1 = (int)((long long)nrows * (ctx->rank + 1) / ctx->world);
                KB_RUN(kb_bpm_pairs_words(ctx, words, S.d_seqs.as<uint8_t>(), S.d_offs.as<int64_t>(), S.d_lens.as<int>(),
                                          d_rows + r0, r1 - r0, d_cols, ncols, ctx->d_stage5.as<float>() + (size_t)r0 * (size_t)ncols));
                KB_RUN(kb_allgatherv(ctx, ctx->d_stage5.p, seg.data()));
        } else {
                KB_RUN(kb_bpm_pairs_words(ctx, words, S.d_seqs.as<uint8_t>(), S.d_offs.as<int64_t>(), S.d_lens.as<int>(),
                                          d_rows, nrows, d_cols, explicit_pairs ? 0 : ncols, ctx->d_stage5.as<float>()));
        }
        if (dm_host) {
                KB_CUDA(cudaMemcpyAsync(dm_host, ctx->d_stage5.p, sizeof(float) * npairs, cudaMemcpyDeviceToHost, ctx->stream));
                ctx->stats.d2h_bytes += (double)(sizeof(float) * npairs);
        }
        // dm_host == nullptr: the distances stay in ctx->d_stage5 for a consumer on the device (UPGMA)
        KB_CUDA(cudaStreamSynchronize(ctx->stream));
        return KB200_OK;
}

// ---------------------------------------------------------------------------------------------

// multi-GPU: cost-balanced shard of the N x K pair list, local compute, one all-gather
int kb_anchor_posmaps_sharded(kb200_ctx* ctx, const kb200_params* prm, KbSeqs& S, const int* anchor_ids, int K, int* posmaps,
                              int host_copy)
{
        // Position maps are produced into ctx->t_posmaps and STAY there: the progressive phase reads
        // them on the device.  `posmaps` (host) is only filled when host_copy != 0 (the host
        // restatement of the bonus, KB200_HOST_BONUS); otherwise it merely tags the device copy.
        const int N = S.n;
        const long long np = (long long)N * K;
        ctx->posmaps_tag = nullptr;
        ctx->posmaps_n = 0;
        cudaStream_t st = ctx->stream;
        auto map_off = [&](long long p) -> size_t {
                if (p >= np) return (size_t)K * (size_t)S.total;
                const int i = (int)(p / K), k = (int)(p % K);
                return (size_t)K * (size_t)S.h_offs[i] + (size_t)k * (size_t)S.h_lens[i];
        };
        const size_t full = (size_t)K * (size_t)S.total;
        KB_RUN(ctx->t_posmaps.ensure(sizeof(int) * (full + 8)));
        int* d_maps = ctx->t_posmaps.as<int>();
        if (ctx->world <= 1) {
                KB_RUN(kb_anchor_posmaps_dev(ctx, prm, S, anchor_ids, K, 0, np, nullptr, d_maps));
        } else {
                std::vector<double> cost((size_t)np);
                for (long long p = 0; p < np; p++) {
                        const int i = (int)(p / K), k = (int)(p % K);
                        cost[(size_t)p] = (i == anchor_ids[k]) ? 1.0 : (double)S.h_lens[i] * (double)S.h_lens[anchor_ids[k]] + 1.0;
                }
                std::vector<int> b((size_t)ctx->world + 1);
                kb_partition(cost.data(), (int)np, ctx->world, b.data());
                // local shard: results land at their final offsets of the full device array
                KB_RUN(kb_anchor_posmaps_dev(ctx, prm, S, anchor_ids, K, b[(size_t)ctx->rank], b[(size_t)ctx->rank + 1], nullptr, d_maps));
                std::vector<size_t> seg((size_t)ctx->world + 1);
                for (int r = 0; r <= ctx->world; r++) seg[(size_t)r] = map_off(b[(size_t)r]) * sizeof(int);
                KB_RUN(kb_allgatherv(ctx, ctx->t_posmaps.p, seg.data()));
        }
        // identity maps of the anchors themselves (anchor_consistency.c:252-258)
        int maxlen = 0;
        for (int k = 0; k < K; k++) maxlen = std::max(maxlen, S.h_lens[anchor_ids[k]]);
        std::vector<int> iota((size_t)maxlen + 1);
        for (int q = 0; q <= maxlen; q++) iota[(size_t)q] = q;
        for (int k = 0; k < K; k++) {
                const int i = anchor_ids[k];
                const size_t o = map_off((long long)i * K + k);
                KB_CUDA(cudaMemcpyAsync(d_maps + o, iota.data(), sizeof(int) * (size_t)S.h_lens[i], cudaMemcpyHostToDevice, st));
        }
        if (host_copy && posmaps) {
                KB_CUDA(cudaMemcpyAsync(posmaps, d_maps, sizeof(int) * full, cudaMemcpyDeviceToHost, st));
                ctx->stats.d2h_bytes += (double)(sizeof(int) * full);
        }
        KB_CUDA(cudaStreamSynchronize(st));
        ctx->posmaps_tag = (const void*)posmaps;
        ctx->posmaps_n = full;
        return KB200_OK;
}

int kb_anchor_posmaps_dev(kb200_ctx* ctx, const kb200_params* prm, KbSeqs& S, const int* anchor_ids, int K,
                          long long pair_begin, long long pair_end, int* posmaps, int* d_full)
{
        const int N = S.n;
        if (pair_begin < 0) pair_begin = 0;
        if (pair_end > (long long)N * K) pair_end = (long long)N * K;
        if (pair_end <= pair_begin) {
                return KB200_OK;
        }
        cudaStream_t st = ctx->stream;
        auto map_off = [&](long long p) -> size_t {
                const int i = (int)(p / K), k = (int)(p % K);
                return (size_t)K * (size_t)S.h_offs[i] + (size_t)k * (size_t)S.h_lens[i];
        };
        const size_t out_begin = map_off(pair_begin);
        size_t out_end;
        {
                const long long last = pair_end - 1;
                out_end = map_off(last) + (size_t)S.h_lens[(int)(last / K)];
        }
        const size_t out_n = out_end - out_begin;
        // sizes
        size_t n_raw = 0, n_coded = 0, n_scr = 0;
        std::vector<KbJob> jobs;
        std::vector<KbPathJob> pjobs;
        jobs.reserve((size_t)(pair_end - pair_begin));
        pjobs.reserve((size_t)(pair_end - pair_begin));
        for (long long p = pair_begin; p < pair_end; p++) {
                const int i = (int)(p / K), k = (int)(p % K);
                const int ak = anchor_ids[k];
                if (i == ak) {
                        continue;
                }
                const int li = S.h_lens[i], lj = S.h_lens[ak];
                const int rows = (li <= lj) ? li : lj;      // anchor_consistency.c:47-63
                n_raw += (size_t)rows + 2;
                n_coded += (size_t)li + (size_t)lj + 2;
                n_scr += (size_t)li + 2;
        }
        KB_RUN(ctx->d_stage0.ensure(sizeof(int) * (n_raw + 8)));
        KB_RUN(ctx->d_stage1.ensure(sizeof(int) * (n_coded + 8)));
        KB_RUN(ctx->d_stage2.ensure(sizeof(int) * (n_scr + 8)));
        if (!d_full) {
                KB_RUN(ctx->d_stage3.ensure(sizeof(int) * (out_n + 8)));
        }
        int* d_raw = ctx->d_stage0.as<int>();
        int* d_coded = ctx->d_stage1.as<int>();
        int* d_scr = ctx->d_stage2.as<int>();
        // d_full: results go to their final offsets of a full-size device array (multi-GPU gather)
        int* d_out = d_full ? (d_full + out_begin) : ctx->d_stage3.as<int>();
        size_t o_raw = 0, o_coded = 0, o_scr = 0;
        for (long long p = pair_begin; p < pair_end; p++) {
                const int i = (int)(p / K), k = (int)(p % K);
                const int ak = anchor_ids[k];
                if (i == ak) {
                        continue;
                }
                const int li = S.h_lens[i], lj = S.h_lens[ak];
                const bool swapped = !(li <= lj);
                KbJob j;
                memset(&j, 0, sizeof(j));
                j.kind = KB200_KIND_SS;
                j.nalpha = prm->nalpha;
                if (!swapped) {
                        j.seq_r = S.dseq(i); j.len_a = li;
                        j.seq_c = S.dseq(ak); j.len_b = lj;
                } else {
                        j.seq_r = S.dseq(ak); j.len_a = lj;
                        j.seq_c = S.dseq(i); j.len_b = li;
                }
                // pairwise_align_map runs with the UNSCALED ap and subm_offset = 0
                j.o = -prm->gpo; j.e = -prm->gpe; j.t = -prm->tgpe; j.nsoff = -0.0f;
                j.path = d_raw + o_raw;
                jobs.push_back(j);
                KbPathJob pj;
                pj.raw = d_raw + o_raw;
                pj.coded = d_coded + o_coded;
                pj.scratch = d_scr + o_scr;
                pj.posmap = d_out + (map_off(p) - out_begin);
                pj.len_a = li;
                pj.len_b = lj;
                pj.mirror = swapped ? 1 : 0;
                pjobs.push_back(pj);
                o_raw += (size_t)j.len_a + 2;
                o_coded += (size_t)li + (size_t)lj + 2;
                o_scr += (size_t)li + 2;
        }
        const auto t0 = std::chrono::steady_clock::now();
        auto t1 = t0, t2 = t0;
        if (!jobs.empty()) {
                KB_CUDA(cudaMemsetAsync(d_raw, 0xFF, sizeof(int) * n_raw, st));
                KB_RUN(kb_run_hirschberg(ctx, prm->subm, jobs));
                t1 = std::chrono::steady_clock::now();
                KB_RUN(ctx->d_stage4.ensure(sizeof(KbPathJob) * pjobs.size()));
                KB_RUN(kb_h2d(ctx, ctx->d_stage4.p, pjobs.data(), sizeof(KbPathJob) * pjobs.size()));
                KB_RUN(kb_code_paths(ctx, ctx->d_stage4.as<KbPathJob>(), (int)pjobs.size()));
        }
        if (getenv("KB200_TRACE")) {
                KB_CUDA(cudaStreamSynchronize(st));
                t2 = std::chrono::steady_clock::now();
        }
        if (!posmaps) {
                KB_RUN(kb_collect(ctx));
                return KB200_OK;
        }
        KB_CUDA(cudaMemcpyAsync(posmaps + out_begin, d_out, sizeof(int) * out_n, cudaMemcpyDeviceToHost, st));
        KB_RUN(kb_collect(ctx));
        if (getenv("KB200_TRACE")) {
                const auto t3 = std::chrono::steady_clock::now();
                auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
                        return std::chrono::duration<double, std::milli>(b - a).count();
                };
                fprintf(stderr, "[kb200 trace] anchor: %zu jobs dp %.2f code_paths %.2f d2h %.2f ms\n", jobs.size(), ms(t0, t1), ms(t1, t2), ms(t2, t3));
        }
        ctx->stats.d2h_bytes += (double)(sizeof(int) * out_n);
        // identity maps for the anchors themselves (anchor_consistency.c:252-258)
        for (long long p = pair_begin; p < pair_end; p++) {
                const int i = (int)(p / K), k = (int)(p % K);
                if (i == anchor_ids[k]) {
                        int* m = posmaps + map_off(p);
                        for (int q = 0; q < S.h_lens[i]; q++) {
                                m[q] = q;
                        }
                }
        }
        return KB200_OK;
}

// ---------------------------------------------------------------------------------------------
extern "C" {

int kb200_distances(kb200_ctx* ctx, const uint8_t* seqs, const int64_t* offs, const int* lens,
                    int nseq, const int* rows, int nrows, const int* cols, int ncols, float* dm)
{
        if (!ctx || !seqs || !offs || !lens || !rows || !cols || !dm || nseq <= 0) {
                return KB200_FAIL;
        }
        KB_CUDA(cudaSetDevice(ctx->device));
        KbSeqs S;
        int rc = S.upload(ctx, seqs, offs, lens, nseq);
        if (rc == KB200_OK) {
                rc = kb_distances_dev(ctx, S, rows, nrows, cols, ncols, 0, dm);
        }
        S.release();
        return rc;
}

struct kb200_seqs {
        kb200_ctx* ctx = nullptr;
        KbSeqs S;
        std::vector<int64_t> offs;
        std::vector<int> lens;
};

int kb200_seqs_upload(kb200_ctx* ctx, const uint8_t* seqs, const int64_t* offs, const int* lens, int nseq, kb200_seqs** out)
{
        if (!ctx || !seqs || !offs || !lens || !out || nseq <= 0) {
                return KB200_FAIL;
        }
        *out = nullptr;
        KB_CUDA(cudaSetDevice(ctx->device));
        kb200_seqs* h = new kb200_seqs();
        h->ctx = ctx;
        h->offs.assign(offs, offs + nseq);
        h->lens.assign(lens, lens + nseq);
        if (h->S.upload(ctx, seqs, h->offs.data(), h->lens.data(), nseq) != KB200_OK) {
                h->S.release();
                delete h;
                return KB200_FAIL;
        }
        h->S.h_seqs = nullptr;       // the caller's code array is not referenced after the upload
        *out = h;
        return KB200_OK;
}

int kb200_distances_on(kb200_seqs* h, const int* rows, int nrows, const int* cols, int ncols, int explicit_pairs, float* dm)
{
        if (!h || !rows || !cols || !dm || nrows < 0) {
                return KB200_FAIL;
        }
        KB_CUDA(cudaSetDevice(h->ctx->device));
        const int n = h->S.n;
        const int ncheck = explicit_pairs ? nrows : ncols;
        for (int i = 0; i < nrows; i++) if (rows[i] < 0 || rows[i] >= n) return KB200_FAIL;
        for (int i = 0; i < ncheck; i++) if (cols[i] < 0 || cols[i] >= n) return KB200_FAIL;
        return kb_distances_dev(h->ctx, h->S, rows, nrows, cols, ncols, explicit_pairs ? 1 : 0, dm);
}

void kb200_seqs_free(kb200_seqs* h)
{
        if (!h) return;
        cudaSetDevice(h->ctx->device);
        h->S.release();
        delete h;
}

int kb200_anchor_posmaps(kb200_ctx* ctx, const kb200_params* prm,
                         const uint8_t* seqs, const int64_t* offs, const int* lens, int nseq,
                         const int* anchor_ids, int K, long long pair_begin, long long pair_end, int* posmaps)
{
        if (!ctx || !prm || !seqs || !offs || !lens || !anchor_ids || !posmaps || nseq <= 0 || K <= 0) {
                return KB200_FAIL;
        }
        KB_CUDA(cudaSetDevice(ctx->device));
        KbSeqs S;
        int rc = S.upload(ctx, seqs, offs, lens, nseq);
        if (rc == KB200_OK) {
                rc = kb_anchor_posmaps_dev(ctx, prm, S, anchor_ids, K, pair_begin, pair_end, posmaps);
        }
        S.release();
        return rc;
}

} // extern "C"
