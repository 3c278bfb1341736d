// kb_api.cu -- C-ABI entry points (include/kalign_b200.h): context, parameters, batched pairwise engine.
#include "kb_common.cuh"

#include <string.h>
#include <stdlib.h>
#include <vector>

#include "kb_subm_tables.inc"

extern "C" {

const char* kb200_version(void) { return "kalign_b200 0.1 (sm_100a)"; }

int kb200_device_count(void)
{
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) {
                cudaGetLastError();
                return 0;
        }
        return n;
}

int kb200_ctx_create(int device, kb200_ctx** out)
{
        if (!out) {
                return KB200_FAIL;
        }
        *out = nullptr;
        int n = kb200_device_count();
        if (n <= 0 || device < 0 || device >= n) {
                fprintf(stderr, "[kalign_b200] no usable CUDA device (count=%d, requested=%d); there is no CPU fallback\n", n, device);
                return KB200_FAIL;
        }
        kb200_ctx* ctx = new kb200_ctx();
        ctx->device = device;
        KB_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        KB_CUDA(cudaGetDeviceProperties(&prop, device));
        ctx->sm_count = prop.multiProcessorCount;
        KB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        KB_CUDA(cudaEventCreate(&ctx->ev0));
        KB_CUDA(cudaEventCreate(&ctx->ev1));
        KB_CUDA(cudaEventCreate(&ctx->ev2));
        KB_CUDA(cudaEventCreate(&ctx->ev3));
        memset(&ctx->stats, 0, sizeof(ctx->stats));
        *out = ctx;
        return KB200_OK;
}

void kb200_ctx_destroy(kb200_ctx* ctx)
{
        if (!ctx) {
                return;
        }
        cudaSetDevice(ctx->device);
        KbDevBuf* bufs[] = {&ctx->d_jobs, &ctx->d_boxA, &ctx->d_boxB, &ctx->d_counters, &ctx->d_rows, &ctx->d_tbl,
                            &ctx->d_units, &ctx->d_prog, &ctx->d_pack, &ctx->d_ppidx, &ctx->d_boxS,
                            &ctx->d_stage0, &ctx->d_stage1, &ctx->d_stage2, &ctx->d_stage3, &ctx->d_stage4, &ctx->d_stage5};
        for (KbDevBuf* b : bufs) {
                b->release();
        }
        KbDevBuf* tbufs[] = {&ctx->t_subm, &ctx->t_leaf, &ctx->t_gapset, &ctx->t_prefix, &ctx->t_raw, &ctx->t_coded, &ctx->t_scr,
                             &ctx->t_pjobs, &ctx->t_mjobs, &ctx->t_src, &ctx->t_bonus, &ctx->t_bidx, &ctx->t_bval, &ctx->t_posmaps,
                             &ctx->t_gaps, &ctx->t_colof, &ctx->t_aoff, &ctx->t_bpos, &ctx->t_bconf, &ctx->t_binv, &ctx->t_bdesc, &ctx->t_wp, &ctx->t_wdesc, &ctx->t_alen, &ctx->t_bvote, &ctx->t_margin, &ctx->t_conf};
        for (KbDevBuf* b : tbufs) {
                b->release();
        }
        ctx->arena.release();
        KbDevBuf* kbufs[] = {&ctx->km_rowsA, &ctx->km_rowsB, &ctx->km_ordA, &ctx->km_ordB, &ctx->km_side, &ctx->km_best, &ctx->km_dmin, &ctx->km_desc};
        for (KbDevBuf* b : kbufs) {
                b->release();
        }
        if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
        for (kb200_ctx::PinnedBlock& b : ctx->host_pool) {
                cudaFreeHost(b.p);
        }
        ctx->host_pool.clear();
        ctx->pinned.release();
        ctx->d_stats.release();
        for (cudaEvent_t e : ctx->ev_pool) {
                cudaEventDestroy(e);
        }
        ctx->ev_pool.clear();
        for (KbDevBuf& b : ctx->seq_pool) {
                b.release();
        }
        ctx->seq_pool.clear();
        if (ctx->ev0) cudaEventDestroy(ctx->ev0);
        if (ctx->ev1) cudaEventDestroy(ctx->ev1);
        if (ctx->ev2) cudaEventDestroy(ctx->ev2);
        if (ctx->ev3) cudaEventDestroy(ctx->ev3);
        if (ctx->stream) cudaStreamDestroy(ctx->stream);
        delete ctx;
}

int kb200_get_stats(kb200_ctx* ctx, kb200_stats* out)
{
        if (!ctx || !out) {
                return KB200_FAIL;
        }
        *out = ctx->stats;
        return KB200_OK;
}

// aln_param_init, lib/src/aln_param.c:17-109.  Matrices are data (PFASUM43/60: Keul et al. 2017,
// gon250, the kalign nucleotide sets); the tables in kb_subm_tables.inc are generated from the
// compiled reference by tools/gen_subm_tables.py.
int kb200_params_init(kb200_params* p, int biotype, int type, float gpo, float gpe, float tgpe)
{
        if (!p) {
                return KB200_FAIL;
        }
        const KbSubmTable* t = nullptr;
        if (biotype == 1) {
                switch (type) {
                case KB200_TYPE_DNA: t = &KB_TBL_DNA; break;
                case KB200_TYPE_DNA_INTERNAL: t = &KB_TBL_DNA_INTERNAL; break;
                case KB200_TYPE_RNA: t = &KB_TBL_RNA; break;
                case KB200_TYPE_PROTEIN:
                        fprintf(stderr, "[kalign_b200] Detected DNA sequences but --type protein option was selected.\n");
                        return KB200_FAIL;
                default: t = &KB_TBL_RNA; break;
                }
                p->nalpha = 5;
                p->vsm_amax = 0.0f;
        } else if (biotype == 0) {
                switch (type) {
                case KB200_TYPE_PROTEIN: t = &KB_TBL_PFASUM43; break;
                case KB200_TYPE_PROTEIN_DIVERGENT: t = &KB_TBL_GON250; break;
                case KB200_TYPE_PROTEIN_PFASUM43: t = &KB_TBL_PFASUM43; break;
                case KB200_TYPE_PROTEIN_PFASUM60: t = &KB_TBL_PFASUM60; break;
                case KB200_TYPE_DNA:
                case KB200_TYPE_DNA_INTERNAL:
                case KB200_TYPE_RNA:
                        fprintf(stderr, "[kalign_b200] Detected protein sequences but a nucleotide --type was selected.\n");
                        return KB200_FAIL;
                default: t = &KB_TBL_PFASUM43; break;
                }
                p->nalpha = 23;
                p->vsm_amax = 2.0f;
        } else {
                fprintf(stderr, "[kalign_b200] Unable to determine what alphabet to use.\n");
                return KB200_FAIL;
        }
        for (int i = 0; i < 23 * 23; i++) {
                p->subm[i] = t->subm[i];
        }
        p->gpo = t->gpo;
        p->gpe = t->gpe;
        p->tgpe = t->tgpe;
        p->dist_scale = 0.0f;          // aln_param.c:95,99
        p->use_seq_weights = 0.0f;
        if (gpo >= 0.0) p->gpo = gpo;
        if (gpe >= 0.0) p->gpe = gpe;
        if (tgpe >= 0.0) p->tgpe = tgpe;
        return KB200_OK;
}

int kb200_pair_align_batch(kb200_ctx* ctx, const kb200_params* prm, const kb200_pair* pairs, int njobs)
{
        if (!ctx || !prm || (!pairs && njobs > 0) || njobs < 0) {
                return KB200_FAIL;
        }
        if (njobs == 0) {
                return KB200_OK;
        }
        KB_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        // ---- stage inputs: bytes (sequences), floats (profiles, bonus) ----
        size_t nbytes = 0, nfloats = 0, npath = 0;
        for (int i = 0; i < njobs; i++) {
                const kb200_pair& p = pairs[i];
                if (p.len_a < 0 || p.len_b < 0) {
                        return KB200_FAIL;
                }
                if (p.kind == KB200_KIND_SS) {
                        nbytes += (size_t)p.len_a + (size_t)p.len_b;
                } else if (p.kind == KB200_KIND_SP) {
                        nbytes += (size_t)p.len_b;
                        nfloats += (size_t)(p.len_a + 2) * 64;
                } else if (p.kind == KB200_KIND_PP) {
                        nfloats += (size_t)(p.len_a + 2) * 64 + (size_t)(p.len_b + 2) * 64;
                } else {
                        return KB200_FAIL;
                }
                if (p.bonus) {
                        nfloats += (size_t)p.len_a * (size_t)p.len_b;
                }
                npath += (size_t)p.len_a + 2;
        }
        std::vector<uint8_t> hb(nbytes + 16);
        std::vector<float> hf(nfloats + 16);
        KB_RUN(ctx->d_stage0.ensure(nbytes + 16));
        KB_RUN(ctx->d_stage1.ensure((nfloats + 16) * sizeof(float)));
        KB_RUN(ctx->d_stage2.ensure(npath * sizeof(int)));
        KB_RUN(ctx->d_stage3.ensure((size_t)njobs * sizeof(float)));
        uint8_t* db = ctx->d_stage0.as<uint8_t>();
        float* df = ctx->d_stage1.as<float>();
        int* dpath = ctx->d_stage2.as<int>();
        float* dscore = ctx->d_stage3.as<float>();
        std::vector<KbJob> jobs((size_t)njobs);
        size_t ob = 0, of = 0, op = 0;
        for (int i = 0; i < njobs; i++) {
                const kb200_pair& p = pairs[i];
                KbJob j;
                memset(&j, 0, sizeof(j));
                j.kind = p.kind;
                j.len_a = p.len_a;
                j.len_b = p.len_b;
                j.nalpha = prm->nalpha;
                if (p.kind == KB200_KIND_SS) {
                        memcpy(hb.data() + ob, p.seq_rows, (size_t)p.len_a);
                        j.seq_r = db + ob; ob += (size_t)p.len_a;
                        memcpy(hb.data() + ob, p.seq_cols, (size_t)p.len_b);
                        j.seq_c = db + ob; ob += (size_t)p.len_b;
                        j.o = -prm->gpo; j.e = -prm->gpe; j.t = -prm->tgpe;
                        j.nsoff = -p.soff;
                } else if (p.kind == KB200_KIND_SP) {
                        memcpy(hb.data() + ob, p.seq_cols, (size_t)p.len_b);
                        j.seq_c = db + ob; ob += (size_t)p.len_b;
                        const size_t w = (size_t)(p.len_a + 2) * 64;
                        memcpy(hf.data() + of, p.prof_rows, w * sizeof(float));
                        j.prof_r = df + of; of += w;
                        // aln_seqprofile.c:31-33: open = gpo * sip (float * int)
                        j.o = -(prm->gpo * (float)p.sip);
                        j.e = -(prm->gpe * (float)p.sip);
                        j.t = -(prm->tgpe * (float)p.sip);
                } else {
                        const size_t wa = (size_t)(p.len_a + 2) * 64, wb = (size_t)(p.len_b + 2) * 64;
                        memcpy(hf.data() + of, p.prof_rows, wa * sizeof(float));
                        j.prof_r = df + of; of += wa;
                        memcpy(hf.data() + of, p.prof_cols, wb * sizeof(float));
                        j.prof_c = df + of; of += wb;
                }
                if (p.bonus) {
                        const size_t w = (size_t)p.len_a * (size_t)p.len_b;
                        memcpy(hf.data() + of, p.bonus, w * sizeof(float));
                        j.bonus = df + of; of += w;
                }
                j.path = p.path_out ? (dpath + op) : nullptr;
                op += (size_t)p.len_a + 2;
                j.score = dscore + i;
                jobs[(size_t)i] = j;
        }
        KB_CUDA(cudaMemcpyAsync(db, hb.data(), nbytes, cudaMemcpyHostToDevice, st));
        KB_CUDA(cudaMemcpyAsync(df, hf.data(), nfloats * sizeof(float), cudaMemcpyHostToDevice, st));
        KB_CUDA(cudaMemsetAsync(dpath, 0xFF, npath * sizeof(int), st));   // init_alnmem: path = -1
        KB_CUDA(cudaMemsetAsync(dscore, 0, (size_t)njobs * sizeof(float), st));
        ctx->stats.h2d_bytes += (double)(nbytes + nfloats * sizeof(float));
        KB_RUN(kb_run_hirschberg(ctx, prm->subm, jobs));
        std::vector<int> hpath(npath);
        std::vector<float> hscore((size_t)njobs);
        KB_CUDA(cudaMemcpyAsync(hpath.data(), dpath, npath * sizeof(int), cudaMemcpyDeviceToHost, st));
        KB_CUDA(cudaMemcpyAsync(hscore.data(), dscore, (size_t)njobs * sizeof(float), cudaMemcpyDeviceToHost, st));
        KB_RUN(kb_collect(ctx));       // waits for the stream; engine error flags -> KB200_FAIL
        ctx->stats.d2h_bytes += (double)(npath * sizeof(int) + (size_t)njobs * sizeof(float));
        op = 0;
        for (int i = 0; i < njobs; i++) {
                const kb200_pair& p = pairs[i];
                if (p.path_out) {
                        memcpy(p.path_out, hpath.data() + op, ((size_t)p.len_a + 2) * sizeof(int));
                }
                if (p.score_out) {
                        *p.score_out = hscore[(size_t)i];
                }
                op += (size_t)p.len_a + 2;
        }
        return KB200_OK;
}

} // extern "C"
