/*
 * kalign_gpu_seams.c -- the reference-side binding of INTEGRATION.md, built for real.
 *
 * This file is the ONE host-side C file a kalign maintainer adds.  It is compiled together with the
 * reference's own, unmodified C files of lib/src (in place, see integration/Makefile) and linked with
 *
 *     -Wl,--wrap=d_estimation -Wl,--wrap=anchor_consistency_build -Wl,--wrap=create_msa_tree
 *     -Wl,--wrap=compute_aln_pairwise_dist -Wl,--wrap=build_tree_kmeans -Wl,--wrap=build_tree_kmeans_noisy
 *
 * so that the three calls kalign_run_seeded() makes into its hot path
 *
 *     d_estimation()              lib/src/sequence_distance.c:37   (from bisectingKmeans.c:205,294)
 *     anchor_consistency_build()  lib/src/anchor_consistency.c:200 (from aln_wrap.c:211)
 *     create_msa_tree()           lib/src/aln_run.c:43             (from aln_wrap.c:225)
 *
 * and the N x N identity distances of the callers that loop over that path (kalign_run_realign,
 * kalign_post_realign; the ensemble runs reach the same seams through kalign_run_seeded / _realign)
 *
 *     compute_aln_pairwise_dist() lib/src/aln_apair_dist.c:9       (from aln_wrap.c:458,604)
 *
 * and, on the two sides of the path, the FASTA reader and writer of the public API (kalign.h:36-37): the
 * reference's lib sources are compiled with -Dkalign_read_input=kalign_read_input_ref
 * -Dkalign_write_msa=kalign_write_msa_ref (its own functions stay available under those names, used for
 * every other format and for standard input / output) and this file defines the public names on top of
 * kb200_fasta_read / kb200_fasta_write (mapped file, all host threads, one write)
 *
 *     kalign_read_input()         lib/src/msa_io.c:80   (read_file_stdin :348, read_fasta :412)
 *     kalign_write_msa()          lib/src/msa_io.c:193  (write_msa_fasta :668)
 *
 * The guide tree as a whole (build_tree_kmeans / build_tree_kmeans_noisy, lib/src/bisectingKmeans.c:177,76, from
 * aln_wrap.c:171,173,298,396) goes to kb200_guide_tree: anchor distances, bisecting k-means and the UPGMA of the leaf
 * clusters run on the GPU, so d_estimation is only reached by callers outside the run wrappers.
 *
 * resolve to the functions below, which flatten struct msa into plain arrays and call the C ABI of
 * libkalign_b200.so (include/kalign_b200.h).  Everything else -- kalign.h, struct msa, I/O, the
 * guide tree, parameter handling, finalise_alignment, the CLI -- is the reference's code, untouched.
 * The result is a libkalign.so.3 + kalign executable that are drop-in for the alignment path.
 *
 * d_estimation keeps the msa's sequences on the device between its calls (kb200_seqs_upload /
 * kb200_distances_on) and asks for exactly the pairs the reference's loops leave in the matrix.
 *
 * There is no CPU fallback: when no CUDA device is usable every seam returns FAIL, and
 * kalign_run()/kalign() fail with it.  Environment: KALIGN_B200_DEVICE selects the GPU (default 0).
 *
 * Every aln_param field the wrapped functions read is honoured: dist_scale (compute_gap_scale,
 * aln_run.c:126), use_seq_weights (update_n, aln_setup.c:237), vsm_amax, the consistency table.
 * create_msa_tree's post-conditions are complete: gaps[], task->confidence (read by --refine
 * confident, aln_refine.c:70), msa->plen / nsip / sip of the internal nodes (aln_run.c:424-436).
 * The refinement passes themselves (refine_alignment, create_msa_tree_inline_refine) stay the
 * reference's CPU code: they rebuild everything from the leaves and only consume task->confidence.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef HAVE_AVX2
#include <xmmintrin.h>
#include <mm_malloc.h>
#endif

#include "tldevel.h"
#include "msa_struct.h"
#include "task.h"
#include "aln_param.h"
#include "anchor_consistency.h"
#include "msa_alloc.h"
#include "msa_op.h"

#include "kalign_b200.h"

#ifdef HAVE_OPENMP
#include <omp.h>
#endif

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;   /* d_estimation(pair=1) is called from OpenMP tasks */
static kb200_ctx* g_ctx = NULL;

/* flat copy of the last consistency table (host): handed to kb200_align_tree without re-packing;
   keyed on the table it was made from, touched only under g_lock */
static struct consistency_table* g_ct_owner = NULL;
static int* g_posmaps = NULL;

/* sequences resident on the device between the d_estimation calls of one guide tree: bisecting_kmeans
   calls d_estimation(pair = 1) once per leaf cluster (bisectingKmeans.c:294), thousands of times on the
   same msa; flattening and uploading the whole msa every time would dominate.  Keyed on (msa, alphabet,
   numseq); released when create_msa_tree runs (the sequences are re-encoded before it). */
static struct msa* g_dist_msa = NULL;
static int g_dist_L = -1;
static int g_dist_n = 0;
static kb200_seqs* g_dist_seqs = NULL;

static void dist_cache_drop(void)
{
        if(g_dist_seqs){
                kb200_seqs_free(g_dist_seqs);
        }
        g_dist_seqs = NULL;
        g_dist_msa = NULL;
        g_dist_L = -1;
        g_dist_n = 0;
}

static kb200_ctx* seam_ctx(void)
{
        if(!g_ctx){
                int dev = 0;
                const char* e = getenv("KALIGN_B200_DEVICE");
                if(e){
                        dev = atoi(e);
                }
                if(kb200_ctx_create(dev, &g_ctx) != KB200_OK){
                        g_ctx = NULL;
                        fprintf(stderr, "[kalign/b200] no usable CUDA device %d: the alignment path has no CPU fallback\n", dev);
                }
        }
        return g_ctx;
}

struct flat_msa{
        uint8_t* seqs;
        int64_t* offs;
        int* lens;
        int n;
};

static void flat_free(struct flat_msa* f)
{
        free(f->seqs);
        free(f->offs);
        free(f->lens);
        f->seqs = NULL; f->offs = NULL; f->lens = NULL;
}

/* msa->sequences[i]->s / ->len (sorted, encoded by convert_msa_to_internal) as one code array */
static int flatten(struct msa* msa, struct flat_msa* f)
{
        int64_t total = 0;
        int i;
        f->n = msa->numseq;
        f->offs = malloc(sizeof(int64_t) * (size_t)f->n);
        f->lens = malloc(sizeof(int) * (size_t)f->n);
        f->seqs = NULL;
        if(!f->offs || !f->lens){
                flat_free(f);
                return FAIL;
        }
        for(i = 0; i < f->n; i++){
                f->offs[i] = total;
                f->lens[i] = msa->sequences[i]->len;
                total += f->lens[i];
        }
        f->seqs = malloc((size_t)total + 1);
        if(!f->seqs){
                flat_free(f);
                return FAIL;
        }
        for(i = 0; i < f->n; i++){
                memcpy(f->seqs + f->offs[i], msa->sequences[i]->s, (size_t)f->lens[i]);
        }
        return OK;
}

static void fill_params(struct msa* msa, struct aln_param* ap, kb200_params* prm)
{
        int i, j;
        for(i = 0; i < 23; i++){
                for(j = 0; j < 23; j++){
                        prm->subm[i * 23 + j] = ap->subm[i][j];
                }
        }
        prm->gpo = ap->gpo;
        prm->gpe = ap->gpe;
        prm->tgpe = ap->tgpe;
        prm->vsm_amax = ap->vsm_amax;
        prm->nalpha = (msa->biotype == ALN_BIOTYPE_DNA) ? 5 : 23;
        prm->dist_scale = ap->dist_scale;
        prm->use_seq_weights = ap->use_seq_weights;
}

/* ------------------------------------------------------------------------------------------------
 * d_estimation (sequence_distance.c:37): pair == 0 -> numseq rows x num_samples (rows padded to a
 * multiple of 8 floats, 32-byte aligned under AVX2, freed row by row by the caller); pair == 1 ->
 * galloc'd num_samples x num_samples matrix (freed with gfree), where the reference's double loop
 * leaves dm[i][j] = dm[j][i] = calc_distance(seq[max(i,j)], seq[min(i,j)]) + length term. */
float** __wrap_d_estimation(struct msa* msa, int* samples, int num_samples, int pair)
{
        float** dm = NULL;
        float* flat = NULL;
        int* rows = NULL;
        int* cols = NULL;
        kb200_ctx* ctx = NULL;
        int numseq = msa->numseq;
        int npairs = 0;
        int i, j;
        int rc = KB200_OK;

        pthread_mutex_lock(&g_lock);
        ctx = seam_ctx();
        if(!ctx){
                pthread_mutex_unlock(&g_lock);
                goto ERROR;
        }
        if(!(g_dist_seqs && g_dist_msa == msa && g_dist_L == msa->L && g_dist_n == numseq)){
                struct flat_msa f = {NULL, NULL, NULL, 0};
                dist_cache_drop();
                if(flatten(msa, &f) != OK || kb200_seqs_upload(ctx, f.seqs, f.offs, f.lens, numseq, &g_dist_seqs) != KB200_OK){
                        flat_free(&f);
                        g_dist_seqs = NULL;
                        pthread_mutex_unlock(&g_lock);
                        goto ERROR;
                }
                flat_free(&f);
                g_dist_msa = msa;
                g_dist_L = msa->L;
                g_dist_n = numseq;
        }
        if(pair){
                /* the reference's double loop leaves dm[i][j] = dm[j][i] = calc_distance(seq[max(i,j)], seq[min(i,j)]):
                   one explicit pair per i >= j */
                npairs = num_samples * (num_samples + 1) / 2;
                rows = malloc(sizeof(int) * (size_t)(npairs > 0 ? npairs : 1));
                cols = malloc(sizeof(int) * (size_t)(npairs > 0 ? npairs : 1));
                flat = malloc(sizeof(float) * (size_t)(npairs > 0 ? npairs : 1));
                if(rows && cols && flat){
                        int p = 0;
                        for(i = 0; i < num_samples; i++){
                                for(j = 0; j <= i; j++){
                                        rows[p] = samples[i];
                                        cols[p] = samples[j];
                                        p++;
                                }
                        }
                        rc = kb200_distances_on(g_dist_seqs, rows, npairs, cols, 0, 1, flat);
                }else{
                        rc = KB200_FAIL;
                }
        }else{
                rows = malloc(sizeof(int) * (size_t)numseq);
                flat = malloc(sizeof(float) * (size_t)numseq * (size_t)num_samples);
                if(rows && flat){
                        for(i = 0; i < numseq; i++){
                                rows[i] = i;
                        }
                        rc = kb200_distances_on(g_dist_seqs, rows, numseq, samples, num_samples, 0, flat);
                }else{
                        rc = KB200_FAIL;
                }
        }
        pthread_mutex_unlock(&g_lock);
        if(rc != KB200_OK){
                goto ERROR;
        }
        if(pair){
                int p = 0;
                RUN(galloc(&dm, num_samples, num_samples));
                for(i = 0; i < num_samples; i++){
                        for(j = 0; j <= i; j++){
                                dm[i][j] = flat[p];
                                dm[j][i] = flat[p];
                                p++;
                        }
                }
        }else{
                int a = num_samples / 8;
                if(num_samples % 8){
                        a++;
                }
                a = a << 3;
                MMALLOC(dm, sizeof(float*) * numseq);
                for(i = 0; i < numseq; i++){
                        dm[i] = NULL;
                }
                for(i = 0; i < numseq; i++){
#ifdef HAVE_AVX2
                        dm[i] = _mm_malloc(sizeof(float) * a, 32);
#else
                        MMALLOC(dm[i], sizeof(float) * a);
#endif
                        if(!dm[i]){
                                goto ERROR;
                        }
                        for(j = 0; j < num_samples; j++){
                                dm[i][j] = flat[(size_t)i * (size_t)num_samples + (size_t)j];
                        }
                        for(j = num_samples; j < a; j++){
                                dm[i][j] = 0.0F;
                        }
                }
        }
        free(flat);
        free(rows);
        free(cols);
        return dm;
ERROR:
        free(flat);
        free(rows);
        free(cols);
        return NULL;
}

/* ------------------------------------------------------------------------------------------------
 * anchor_consistency_build (anchor_consistency.c:200): same table, same ownership (every map is
 * its own allocation, released by anchor_consistency_free); the N x K pairwise alignments run as
 * one batch on the GPU. */
int __wrap_anchor_consistency_build(struct msa* msa, struct aln_param* ap, int n_anchors, float weight,
                                    struct consistency_table** ct_out)
{
        struct consistency_table* ct = NULL;
        struct flat_msa f = {NULL, NULL, NULL, 0};
        kb200_params prm;
        kb200_ctx* ctx = NULL;
        int* posmaps = NULL;
        int N = msa->numseq;
        int K = n_anchors;
        int i, k;
        int64_t total = 0;

        if(K <= 0 || N < 3){
                *ct_out = NULL;
                return OK;
        }
        if(K > N){
                K = N;
        }
        if(msa->seq_distances == NULL){
                *ct_out = NULL;
                return OK;
        }
        ctx = seam_ctx();
        if(!ctx){
                *ct_out = NULL;
                return FAIL;
        }
        MMALLOC(ct, sizeof(struct consistency_table));
        ct->pos_maps = NULL;
        ct->map_lengths = NULL;
        ct->anchor_ids = NULL;
        ct->n_anchors = K;
        ct->numseq = N;
        ct->weight = weight;
        MMALLOC(ct->anchor_ids, sizeof(int) * K);
        MMALLOC(ct->pos_maps, sizeof(int*) * N * K);
        MMALLOC(ct->map_lengths, sizeof(int) * N * K);
        for(i = 0; i < N * K; i++){
                ct->pos_maps[i] = NULL;
                ct->map_lengths[i] = 0;
        }
        if(kb200_select_anchors(msa->seq_distances, N, K, ct->anchor_ids) != KB200_OK){
                goto ERROR;
        }
        if(!msa->quiet){
                LOG_MSG("Anchor consistency: K=%d, weight=%.1f (GPU batch)", K, weight);
        }
        RUN(flatten(msa, &f));
        fill_params(msa, ap, &prm);
        for(i = 0; i < N; i++){
                total += f.lens[i];
        }
        posmaps = malloc(sizeof(int) * (size_t)total * (size_t)K + sizeof(int));
        if(!posmaps){
                goto ERROR;
        }
        if(kb200_anchor_posmaps(ctx, &prm, f.seqs, f.offs, f.lens, N, ct->anchor_ids, K, 0, (long long)N * K, posmaps) != KB200_OK){
                goto ERROR;
        }
        for(i = 0; i < N; i++){
                for(k = 0; k < K; k++){
                        const int len_i = f.lens[i];
                        ct->map_lengths[i * K + k] = len_i;
                        MMALLOC(ct->pos_maps[i * K + k], sizeof(int) * (len_i > 0 ? len_i : 1));
                        memcpy(ct->pos_maps[i * K + k], posmaps + (size_t)K * (size_t)f.offs[i] + (size_t)k * (size_t)len_i,
                               sizeof(int) * (size_t)len_i);
                }
        }
        /* keep the flat copy for create_msa_tree */
        pthread_mutex_lock(&g_lock);
        free(g_posmaps);
        g_posmaps = posmaps;
        g_ct_owner = ct;
        pthread_mutex_unlock(&g_lock);
        flat_free(&f);
        *ct_out = ct;
        return OK;
ERROR:
        free(posmaps);
        flat_free(&f);
        anchor_consistency_free(ct);
        *ct_out = NULL;
        return FAIL;
}

/* ------------------------------------------------------------------------------------------------
 * create_msa_tree (aln_run.c:43): same arguments; post-condition used by finalise_alignment
 * (msa_op.c:546): every sequences[i]->gaps[0..len] filled. */
int __wrap_create_msa_tree(struct msa* msa, struct aln_param* ap, struct aln_tasks* t)
{
        struct flat_msa f = {NULL, NULL, NULL, 0};
        struct consistency_table* ct = (struct consistency_table*)msa->consistency_table;
        kb200_params prm;
        kb200_ctx* ctx = NULL;
        int* abc = NULL;
        int* gaps = NULL;
        int* posmaps = NULL;
        float* conf = NULL;
        int* plen = NULL;
        int own_posmaps = 0;
        int j, g;
        int K = 0;
        float weight = 0.0F;
        int N = msa->numseq;
        int i, k;
        int64_t total = 0;

        pthread_mutex_lock(&g_lock);
        ctx = seam_ctx();
        dist_cache_drop();                               /* the guide tree is built; the sequences were re-encoded */
        pthread_mutex_unlock(&g_lock);
        if(!ctx){
                return FAIL;
        }
        RUN(sort_tasks(t, TASK_ORDER_TREE));             /* as aln_run.c:48 */
        RUN(flatten(msa, &f));
        fill_params(msa, ap, &prm);
        for(i = 0; i < N; i++){
                total += f.lens[i];
        }
        abc = malloc(sizeof(int) * 3 * (size_t)(t->n_tasks > 0 ? t->n_tasks : 1));
        gaps = malloc(sizeof(int) * ((size_t)total + (size_t)N));
        conf = malloc(sizeof(float) * (size_t)(t->n_tasks > 0 ? t->n_tasks : 1));
        plen = malloc(sizeof(int) * (size_t)(t->n_tasks > 0 ? t->n_tasks : 1));
        if(!abc || !gaps || !conf || !plen){
                goto ERROR;
        }
        for(i = 0; i < t->n_tasks; i++){
                abc[3 * i] = t->list[i]->a;
                abc[3 * i + 1] = t->list[i]->b;
                abc[3 * i + 2] = t->list[i]->c;
        }
        if(ct){
                K = ct->n_anchors;
                weight = ct->weight;
                pthread_mutex_lock(&g_lock);
                if(ct == g_ct_owner && g_posmaps){
                        posmaps = g_posmaps;             /* ownership moves to this call */
                        own_posmaps = 1;
                        g_posmaps = NULL;
                        g_ct_owner = NULL;
                }
                pthread_mutex_unlock(&g_lock);
                if(!posmaps){
                        /* a table built elsewhere (e.g. by the reference's CPU path): pack it */
                        posmaps = malloc(sizeof(int) * (size_t)total * (size_t)K + sizeof(int));
                        if(!posmaps){
                                goto ERROR;
                        }
                        own_posmaps = 1;
                        for(i = 0; i < N; i++){
                                for(k = 0; k < K; k++){
                                        memcpy(posmaps + (size_t)K * (size_t)f.offs[i] + (size_t)k * (size_t)f.lens[i],
                                               ct->pos_maps[i * K + k], sizeof(int) * (size_t)f.lens[i]);
                                }
                        }
                }
        }
        if(kb200_align_tree_conf(ctx, &prm, f.seqs, f.offs, f.lens, N, abc, t->n_tasks, msa->seq_distances,
                                 posmaps, K, weight, gaps, conf, plen) != KB200_OK){
                goto ERROR;
        }
        for(i = 0; i < N; i++){
                memcpy(msa->sequences[i]->gaps, gaps + f.offs[i] + i, sizeof(int) * (size_t)(f.lens[i] + 1));
        }
        /* the rest of do_align's bookkeeping (aln_run.c:390-436): confidence, plen, nsip, sip */
        for(i = 0; i < t->n_tasks; i++){
                const int a = t->list[i]->a;
                const int b = t->list[i]->b;
                const int c = t->list[i]->c;
                t->list[i]->confidence = conf[i];
                msa->plen[c] = plen[i];
                msa->nsip[c] = msa->nsip[a] + msa->nsip[b];
                MREALLOC(msa->sip[c], sizeof(int) * (size_t)(msa->nsip[a] + msa->nsip[b]));
                g = 0;
                for(j = msa->nsip[a]; j--;){
                        msa->sip[c][g] = msa->sip[a][j];
                        g++;
                }
                for(j = msa->nsip[b]; j--;){
                        msa->sip[c][g] = msa->sip[b][j];
                        g++;
                }
        }
        if(own_posmaps){
                free(posmaps);
        }
        free(abc);
        free(gaps);
        free(conf);
        free(plen);
        flat_free(&f);
        return OK;
ERROR:
        if(own_posmaps){
                free(posmaps);
        }
        free(abc);
        free(gaps);
        free(conf);
        free(plen);
        flat_free(&f);
        return FAIL;
}

/* ------------------------------------------------------------------------------------------------
 * compute_aln_pairwise_dist (aln_apair_dist.c:9): same float** layout (n rows of n floats, each its
 * own allocation, released by free_aln_dm); the N^2/2 row comparisons run on the GPU. */
int __wrap_compute_aln_pairwise_dist(struct msa* msa, float*** dm_ptr)
{
        float** dm = NULL;
        const char** rows = NULL;
        kb200_ctx* ctx = NULL;
        int n = 0;
        int i;
        int rc;

        ASSERT(msa != NULL, "No MSA");
        ASSERT(msa->aligned == ALN_STATUS_FINAL, "MSA must be finalized");
        n = msa->numseq;
        pthread_mutex_lock(&g_lock);
        ctx = seam_ctx();
        pthread_mutex_unlock(&g_lock);
        if(!ctx){
                return FAIL;
        }
        MMALLOC(dm, sizeof(float*) * n);
        for(i = 0; i < n; i++){
                dm[i] = NULL;
        }
        for(i = 0; i < n; i++){
                MMALLOC(dm[i], sizeof(float) * n);
        }
        rows = malloc(sizeof(char*) * (size_t)(n > 0 ? n : 1));
        if(!rows){
                goto ERROR;
        }
        for(i = 0; i < n; i++){
                rows[i] = msa->sequences[i]->seq;
        }
        rc = kb200_aln_pairwise_dist(ctx, rows, n, msa->alnlen, dm);
        free(rows);
        rows = NULL;
        if(rc != KB200_OK){
                goto ERROR;
        }
        *dm_ptr = dm;
        return OK;
ERROR:
        free(rows);
        if(dm){
                for(i = 0; i < n; i++){
                        if(dm[i]){
                                MFREE(dm[i]);
                        }
                }
                MFREE(dm);
        }
        return FAIL;
}

/* ------------------------------------------------------------------------------------------------
 * kalign_read_input (msa_io.c:80) for FASTA files.  The format is decided as detect_alignment_format
 * does (msa_io.c:248-346: the first 100 lines; any Clustal / MSF marker wins over '>'); everything that
 * is not an unambiguous FASTA file -- standard input, MSF, Clustal, a second input file merged into an
 * existing msa, a file whose first line has one character (the reference calls that "no input") --
 * goes to the reference's own reader. */
int kalign_read_input_ref(char* infile, struct msa** msa, int quiet);
int kalign_write_msa_ref(struct msa* msa, char* outfile, char* format);

static int fasta_fast_path(const char* infile)
{
        static const char* markers[6] = {"multiple sequence alignment", "CLUSTAL W", "CLUSTAL O",
                                         "!!AA_MULTIPLE_ALIGNMENT", "!!NA_MULTIPLE_ALIGNMENT", "MSF:"};
        enum { HEAD = 1 << 20 };
        FILE* f = fopen(infile, "r");
        char* buf = NULL;
        size_t n, pos = 0;
        int lines = 0;
        int fasta = 0;
        int ok = 1;
        if(!f){
                return 0;
        }
        buf = malloc(HEAD + 1);
        if(!buf){
                fclose(f);
                return 0;
        }
        n = fread(buf, 1, HEAD, f);
        while(ok && lines < 100 && pos < n){
                char* nl = memchr(buf + pos, 10, n - pos);
                size_t end = nl ? (size_t)(nl - buf) : n;
                size_t cut = pos;
                int m;
                if(!nl && n == HEAD){
                        ok = 0;                  /* 100 lines do not fit in the head of the file: let the reference decide */
                        break;
                }
                while(cut < end && !((unsigned char)buf[cut] < 32 || buf[cut] == 127)){
                        cut++;
                }
                {
                        char keep = buf[cut];
                        buf[cut] = 0;            /* the line's content, as read_file_stdin cuts it (msa_io.c:381-386) */
                        if(lines == 0 && cut - pos == 1){
                                ok = 0;
                        }
                        if(buf[pos] == 62){
                                fasta = 1;
                        }
                        for(m = 0; m < 6; m++){
                                if(strstr(buf + pos, markers[m])){
                                        ok = 0;
                                }
                        }
                        buf[cut] = keep;
                }
                pos = end + 1;
                lines++;
        }
        free(buf);
        fclose(f);
        return ok && fasta;
}

int kalign_read_input(char* infile, struct msa** msa, int quiet)
{
        struct msa* m = NULL;
        kb200_fasta* f = NULL;
        const int* freq = NULL;
        int n, i, alloc;

        if(!infile || !msa || *msa != NULL || !fasta_fast_path(infile)){
                return kalign_read_input_ref(infile, msa, quiet);
        }
        if(kb200_fasta_read(infile, 0, &f) != KB200_OK){
                return kalign_read_input_ref(infile, msa, quiet);       /* the reference reports what is wrong with the file */
        }
        n = kb200_fasta_numseq(f);
        alloc = 512 * ((n + 511) / 512 > 0 ? (n + 511) / 512 : 1);     /* alloc_msa(512) + resize_msa steps (msa_io.c:423-431) */
        RUN(alloc_msa(&m, alloc));
        for(i = 0; i < n; i++){
                struct msa_seq* s = m->sequences[i];
                const char* name = NULL;
                const char* seq = NULL;
                const int* gaps = NULL;
                int len = 0;
                int a, j;
                kb200_fasta_get(f, i, &name, &seq, &len, &gaps);
                MFREE(s->name);
                MMALLOC(s->name, strlen(name) + 1);
                strcpy(s->name, name);
                a = 512 * (len / 512 + 1);                              /* resize_msa_seq steps (msa_alloc.c:141) */
                if(a != s->alloc_len){
                        s->alloc_len = a;
                        MREALLOC(s->seq, sizeof(char) * a);
                        MREALLOC(s->s, sizeof(uint8_t) * a);
                        MREALLOC(s->gaps, sizeof(int) * (a + 1));
                }
                memcpy(s->seq, seq, (size_t)len);
                s->seq[len] = 0;
                memcpy(s->gaps, gaps, sizeof(int) * (size_t)(len + 1));
                for(j = len + 1; j < a + 1; j++){
                        s->gaps[j] = 0;
                }
                s->len = len;
        }
        m->numseq = n;
        freq = kb200_fasta_letter_freq(f);
        for(i = 0; i < 128; i++){
                m->letter_freq[i] = freq[i];
        }
        kb200_fasta_free(f);
        f = NULL;
        m->quiet = quiet;
        RUN(detect_alphabet(m));
        RUN(detect_aligned(m));
        RUN(set_sip_nsip(m));
        if(!quiet){
                LOG_MSG("Read %d sequences from %s.", m->numseq, infile);
        }
        *msa = m;
        m = NULL;
        /* check_for_sequences (msa_io.c:176-191) */
        if((*msa)->numseq == 0){
                ERROR_MSG("No sequences were found in the input files or standard input.");
        }else if((*msa)->numseq == 1){
                ERROR_MSG("Only 1 sequence was found in the input files or standard input");
        }
        return OK;
ERROR:
        if(f){
                kb200_fasta_free(f);
        }
        if(m){
                kalign_free_msa(m);
        }
        return FAIL;
}

/* ------------------------------------------------------------------------------------------------
 * kalign_write_msa (msa_io.c:193) for FASTA output to a file; other formats and standard output stay
 * with the reference's writers (parse_format_argument, msa_io.c:223-246: "msf" and "clu" win). */
int kalign_write_msa(struct msa* msa, char* outfile, char* format)
{
        const char** names = NULL;
        const char** rows = NULL;
        int i, rc;
        if(!msa || !outfile || msa->aligned != ALN_STATUS_FINAL ||
           (format && (strstr(format, "msf") || strstr(format, "clu") || !strstr(format, "fa")))){
                return kalign_write_msa_ref(msa, outfile, format);
        }
        names = malloc(sizeof(char*) * (size_t)(msa->numseq > 0 ? msa->numseq : 1));
        rows = malloc(sizeof(char*) * (size_t)(msa->numseq > 0 ? msa->numseq : 1));
        if(!names || !rows){
                free(names);
                free(rows);
                return FAIL;
        }
        for(i = 0; i < msa->numseq; i++){
                names[i] = msa->sequences[i]->name;
                rows[i] = msa->sequences[i]->seq;
        }
        rc = kb200_fasta_write(outfile, names, rows, msa->numseq, msa->alnlen, 0);
        free(names);
        free(rows);
        return rc == KB200_OK ? OK : FAIL;
}

/* ------------------------------------------------------------------------------------------------
 * build_tree_kmeans / build_tree_kmeans_noisy (bisectingKmeans.c:177,76): same post-conditions -- the task
 * list in create_tasks' order (a node, its left subtree, its right subtree) and msa->seq_distances. */
static int guide_tree_on_gpu(struct msa* msa, struct aln_tasks** tasks, uint64_t seed, float noise_sigma)
{
        struct aln_tasks* t = NULL;
        struct flat_msa f = {NULL, NULL, NULL, 0};
        kb200_ctx* ctx = NULL;
        int* abc = NULL;
        int numseq;
        int n_threads = 1;
        int i;

        ASSERT(msa != NULL, "No alignment.");
        t = *tasks;
        if(!t){
                RUN(alloc_tasks(&t, msa->numseq));
                *tasks = t;
        }
        numseq = msa->numseq;
        pthread_mutex_lock(&g_lock);
        ctx = seam_ctx();
        pthread_mutex_unlock(&g_lock);
        if(!ctx){
                return FAIL;
        }
#ifdef HAVE_OPENMP
        n_threads = omp_get_max_threads();
#endif
        if(!msa->quiet){
                LOG_MSG("Calculating pairwise distances and building the guide tree (GPU)");
        }
        if(msa->seq_distances == NULL){
                MMALLOC(msa->seq_distances, sizeof(float) * numseq);
        }
        RUN(flatten(msa, &f));
        abc = malloc(sizeof(int) * 3 * (size_t)(numseq > 1 ? numseq - 1 : 1));
        if(!abc){
                goto ERROR;
        }
        if(kb200_guide_tree(ctx, f.seqs, f.offs, f.lens, numseq, n_threads, (unsigned long long)seed, noise_sigma, abc,
                            msa->seq_distances) != KB200_OK){
                goto ERROR;
        }
        for(i = 0; i < numseq - 1; i++){
                struct task* task = t->list[t->n_tasks];
                task->a = abc[3 * i];
                task->b = abc[3 * i + 1];
                task->c = abc[3 * i + 2];
                t->n_tasks++;
        }
        free(abc);
        flat_free(&f);
        return OK;
ERROR:
        free(abc);
        flat_free(&f);
        return FAIL;
}

int __wrap_build_tree_kmeans(struct msa* msa, struct aln_tasks** tasks)
{
        return guide_tree_on_gpu(msa, tasks, 0, 0.0f);
}

int __wrap_build_tree_kmeans_noisy(struct msa* msa, struct aln_tasks** tasks, uint64_t seed, float noise_sigma)
{
        return guide_tree_on_gpu(msa, tasks, seed, noise_sigma);
}
