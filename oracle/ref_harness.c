/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin flat-C entry points around the UNMODIFIED reference (TimoLassmann/kalign 3.5.1),
 * compiled by oracle/Makefile from the sources where they lie under /root/reference into
 * oracle/_ref/libref_harness.so.  Nothing in the product (kalign_b200/) links or loads this.
 * It exists so that tests / golden generators / bench.py --impl reference can drive the
 * reference's own functions stage by stage:
 *
 *   aln_runner            lib/src/aln_controller.c:21
 *   init_alnmem           lib/src/aln_setup.c:13
 *   make_profile_n        lib/src/aln_setup.c:40
 *   set_gap_penalties_n   lib/src/aln_setup.c:101
 *   add_gap_info_to_path_n lib/src/aln_setup.c:121
 *   update_n              lib/src/aln_setup.c:230
 *   mirror_path_n         lib/src/aln_setup.c:438
 *   kalign_run_seeded     lib/src/aln_wrap.c:133  (re-stated call sequence with dumps)
 *   compute_aln_pairwise_dist lib/src/aln_apair_dist.c:9
 *   kalign_read_input     lib/src/msa_io.c:80   (record dumps for the FASTA parser's parity tests)
 *
 * All functions return 0 on success.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <float.h>
#include <time.h>

#ifdef HAVE_OPENMP
#include <omp.h>
#endif
#ifdef HAVE_AVX2
#include <xmmintrin.h>
#include <mm_malloc.h>
#endif

#include "tldevel.h"
#include "msa_struct.h"
#include "msa_op.h"
#include "msa_alloc.h"
#include "msa_check.h"
#include "msa_sort.h"
#include "alphabet.h"
#include "task.h"
#include "bisectingKmeans.h"
#include "sequence_distance.h"
#include "pick_anchor.h"
#include "aln_param.h"
#include "aln_struct.h"
#include "aln_mem.h"
#include "aln_setup.h"
#include "aln_controller.h"
#include "aln_run.h"
#include "anchor_consistency.h"
#include "bpm.h"
#include "aln_apair_dist.h"
#include "tlrng.h"
#include "kalign/kalign.h"

static double now_s(void)
{
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Build an aln_param with a caller supplied 23x23 matrix. */
static struct aln_param* make_ap(const float* subm, float gpo, float gpe, float tgpe, float soff)
{
        struct aln_param* ap = calloc(1, sizeof(struct aln_param));
        ap->subm = malloc(sizeof(float*) * 23);
        for(int i = 0; i < 23; i++){
                ap->subm[i] = malloc(sizeof(float) * 23);
                for(int j = 0; j < 23; j++){
                        ap->subm[i][j] = subm[i * 23 + j];
                }
        }
        ap->nthreads = 1;
        ap->gpo = gpo;
        ap->gpe = gpe;
        ap->tgpe = tgpe;
        ap->subm_offset = soff;
        return ap;
}

static void free_ap(struct aln_param* ap)
{
        for(int i = 0; i < 23; i++){
                free(ap->subm[i]);
        }
        free(ap->subm);
        free(ap);
}

/* One pairwise Hirschberg alignment through the reference's aln_runner.
   kind: 0 seq-seq (rows=seq1, cols=seq2), 1 profile(rows)-seq(cols), 2 profile-profile.
   bonus: optional dense len_a*len_b matrix (may be NULL).
   path_out: raw path, len_a+2 ints (index 1..len_a valid).
   score_only != 0 -> only the top level meet-up score is produced. */
int refh_pair_align(int kind,
                    const uint8_t* seq1, const uint8_t* seq2,
                    const float* prof1, const float* prof2,
                    int len_a, int len_b, int sip,
                    const float* subm, float gpo, float gpe, float tgpe, float soff,
                    const float* bonus,
                    int score_only,
                    int* path_out, float* score_out,
                    float* margin_sum_out, int* margin_count_out)
{
        struct aln_mem* m = NULL;
        struct aln_param* ap = make_ap(subm, gpo, gpe, tgpe, soff);
        if(alloc_aln_mem(&m, 256) != OK){
                return 1;
        }
        m->ap = ap;
        m->mode = score_only ? ALN_MODE_SCORE_ONLY : ALN_MODE_FULL;
        m->run_parallel = 0;
        m->len_a = len_a;
        m->len_b = len_b;
        init_alnmem(m);
        m->margin_sum = 0.0F;
        m->margin_count = 0;
        m->seq1 = NULL; m->seq2 = NULL; m->prof1 = NULL; m->prof2 = NULL;
        m->sip = sip;
        if(kind == 0){
                m->seq1 = seq1; m->seq2 = seq2;
        }else if(kind == 1){
                m->prof1 = prof1; m->seq2 = seq2;
        }else{
                m->prof1 = prof1; m->prof2 = prof2;
        }
        m->consistency = (float*)bonus;
        m->consistency_stride = len_b;
        aln_runner(m);
        if(score_out){ *score_out = m->score; }
        if(margin_sum_out){ *margin_sum_out = m->margin_sum; }
        if(margin_count_out){ *margin_count_out = m->margin_count; }
        if(path_out && !score_only){
                for(int i = 0; i < len_a + 2; i++){
                        path_out[i] = m->path[i];
                }
        }
        m->consistency = NULL;
        free_aln_mem(m);
        free_ap(ap);
        return 0;
}

/* raw path (+ optional mirror) -> coded path. path_io has len_a+len_b+2 ints.
   If mirror != 0 the raw path was produced in swapped orientation (rows = b): it is mirrored
   first exactly as do_align does (aln_run.c:323), with (len_a,len_b) the un-swapped lengths. */
int refh_code_path(int* path_io, int len_a, int len_b, int mirror)
{
        struct aln_mem* m = NULL;
        if(alloc_aln_mem(&m, 256) != OK){
                return 1;
        }
        m->len_a = len_a;
        m->len_b = len_b;
        resize_aln_mem(m);
        int n = (mirror ? len_b : len_a) + 2;
        for(int i = 0; i < n; i++){
                m->path[i] = path_io[i];
        }
        if(mirror){
                mirror_path_n(m, len_a, len_b);
        }
        add_gap_info_to_path_n(m);
        for(int i = 0; i < len_a + len_b + 2; i++){
                path_io[i] = m->path[i];
        }
        free_aln_mem(m);
        return 0;
}

int refh_make_profile(const uint8_t* seq, int len, const float* subm,
                      float gpo, float gpe, float tgpe, float soff, float* prof_out)
{
        struct aln_param* ap = make_ap(subm, gpo, gpe, tgpe, soff);
        float* p = NULL;
        make_profile_n(ap, seq, len, 1.0f, &p);
        memcpy(prof_out, p, sizeof(float) * 64 * (len + 2));
        MFREE(p);
        free_ap(ap);
        return 0;
}

int refh_set_gap_penalties(float* prof, int len, int nsip)
{
        return set_gap_penalties_n(prof, len, nsip);
}

int refh_update(const float* profa, const float* profb, float* newp, int* path,
                int sipa, int sipb, float gpo, float gpe, float tgpe)
{
        float zero[23 * 23];
        memset(zero, 0, sizeof(zero));
        struct aln_param* ap = make_ap(zero, gpo, gpe, tgpe, 0.0f);
        ap->use_seq_weights = 0.0f;
        update_n(profa, profb, newp, ap, path, sipa, sipb);
        free_ap(ap);
        return 0;
}

/* ---- whole pipeline with stage dumps ------------------------------------------------- */

struct refh_run {
        struct msa* msa;
        struct aln_tasks* tasks;
        struct aln_param* ap;
        double t_dist_tree;
        double t_anchor;
        double t_tree_aln;
        double t_total;
};

static struct msa* seqs_to_msa(char** seqs, int* lens, int n)
{
        struct msa* msa = NULL;
        if(kalign_arr_to_msa(seqs, lens, n, &msa) != OK){
                return NULL;
        }
        /* names matter for the (len desc, name asc) sort: s0..s{n-1}, as in the FASTA generator */
        for(int i = 0; i < n; i++){
                snprintf(msa->sequences[i]->name, MSA_NAME_LEN, "s%d", i);
        }
        msa->quiet = 1;
        return msa;
}

/* Run the reference exactly as kalign_run_seeded (aln_wrap.c:133-261) does, refine=NONE,
   default tree (no noise), recording stage wall-clock times.  Returns an opaque handle. */
void* refh_run_pipeline_ex(char** seqs, int* lens, int n, int n_threads, int type,
                           float gpo, float gpe, float tgpe,
                           int consistency_anchors, float consistency_weight,
                           int stop_after, float dist_scale, float use_seq_weights);

void* refh_run_pipeline(char** seqs, int* lens, int n, int n_threads, int type,
                        float gpo, float gpe, float tgpe,
                        int consistency_anchors, float consistency_weight,
                        int stop_after /* 0 all, 1 after tree, 2 after anchor */)
{
        return refh_run_pipeline_ex(seqs, lens, n, n_threads, type, gpo, gpe, tgpe, consistency_anchors, consistency_weight,
                                    stop_after, 0.0f, 0.0f);
}

/* dist_scale / use_seq_weights as kalign_run_seeded sets them on the aln_param (aln_wrap.c:193-198);
   compute_tree_weights (static, aln_wrap.c:70) only fills msa->seq_weights, which the DP path never reads */
void* refh_run_pipeline_ex(char** seqs, int* lens, int n, int n_threads, int type,
                           float gpo, float gpe, float tgpe,
                           int consistency_anchors, float consistency_weight,
                           int stop_after, float dist_scale, float use_seq_weights)
{
        struct refh_run* r = calloc(1, sizeof(struct refh_run));
        double t0, t1;
        struct msa* msa = seqs_to_msa(seqs, lens, n);
        if(!msa){ free(r); return NULL; }
        r->msa = msa;
        double tstart = now_s();
        if(kalign_essential_input_check(msa, 0) != OK){ goto ERROR; }
        if(msa->aligned != ALN_STATUS_UNALIGNED){
                dealign_msa(msa);
        }
        msa_sort_len_name(msa);
        if(msa->biotype == ALN_BIOTYPE_DNA){
                msa->L = ALPHA_defDNA;
                convert_msa_to_internal(msa, ALPHA_defDNA);
        }else if(msa->biotype == ALN_BIOTYPE_PROTEIN){
                msa->L = ALPHA_redPROTEIN;
                convert_msa_to_internal(msa, ALPHA_redPROTEIN);
        }else{
                goto ERROR;
        }
        alloc_tasks(&r->tasks, msa->numseq);
#ifdef HAVE_OPENMP
        omp_set_num_threads(n_threads);
#endif
        t0 = now_s();
        if(build_tree_kmeans(msa, &r->tasks) != OK){ goto ERROR; }
        t1 = now_s();
        r->t_dist_tree = t1 - t0;
        if(msa->biotype == ALN_BIOTYPE_PROTEIN){
                convert_msa_to_internal(msa, ALPHA_ambigiousPROTEIN);
        }
        if(type == KALIGN_TYPE_PROTEIN_PFASUM_AUTO){
                type = KALIGN_TYPE_PROTEIN_PFASUM43; /* harness does not exercise AUTO */
        }
        if(aln_param_init(&r->ap, msa->biotype, n_threads, type, gpo, gpe, tgpe) != OK){ goto ERROR; }
        if(use_seq_weights >= 0.0f){ r->ap->use_seq_weights = use_seq_weights; }
        if(dist_scale > 0.0f){ r->ap->dist_scale = dist_scale; }
        if(stop_after == 1){
                r->t_total = now_s() - tstart;
                return r;
        }
        if(consistency_anchors > 0){
                r->ap->consistency_anchors = consistency_anchors;
                r->ap->consistency_weight = consistency_weight;
                t0 = now_s();
                if(anchor_consistency_build(msa, r->ap, consistency_anchors, consistency_weight,
                                            (struct consistency_table**)&msa->consistency_table) != OK){ goto ERROR; }
                t1 = now_s();
                r->t_anchor = t1 - t0;
        }
        if(stop_after == 2){
                r->t_total = now_s() - tstart;
                return r;
        }
        t0 = now_s();
        if(create_msa_tree(msa, r->ap, r->tasks) != OK){ goto ERROR; }
        t1 = now_s();
        r->t_tree_aln = t1 - t0;
        msa->aligned = ALN_STATUS_ALIGNED;
        /* keep the consistency table alive for dumps; freed in refh_free */
        finalise_alignment(msa);
        /* NOTE: msa_sort_rank is applied lazily in refh_get_aligned so that the internal
           (sorted) index space stays valid for the dump functions. */
        r->t_total = now_s() - tstart;
        return r;
ERROR:
        r->msa = NULL;
        return NULL;
}

int refh_numseq(void* h){ return ((struct refh_run*)h)->msa->numseq; }
int refh_biotype(void* h){ return ((struct refh_run*)h)->msa->biotype; }
int refh_alnlen(void* h){ return ((struct refh_run*)h)->msa->alnlen; }

void refh_times(void* h, double* out4)
{
        struct refh_run* r = h;
        out4[0] = r->t_dist_tree; out4[1] = r->t_anchor; out4[2] = r->t_tree_aln; out4[3] = r->t_total;
}

/* sorted-index -> original input rank, and sorted sequence lengths */
void refh_get_order(void* h, int* rank_out, int* len_out)
{
        struct refh_run* r = h;
        for(int i = 0; i < r->msa->numseq; i++){
                rank_out[i] = r->msa->sequences[i]->rank;
                len_out[i] = r->msa->sequences[i]->len;
        }
}

/* internal codes of sorted sequence i (whatever alphabet is current) */
void refh_get_codes(void* h, int i, uint8_t* out)
{
        struct refh_run* r = h;
        memcpy(out, r->msa->sequences[i]->s, r->msa->sequences[i]->len);
}

int refh_get_tasks(void* h, int* abc_out)
{
        struct refh_run* r = h;
        for(int i = 0; i < r->tasks->n_tasks; i++){
                abc_out[3 * i + 0] = r->tasks->list[i]->a;
                abc_out[3 * i + 1] = r->tasks->list[i]->b;
                abc_out[3 * i + 2] = r->tasks->list[i]->c;
        }
        return r->tasks->n_tasks;
}

void refh_get_task_confidence(void* h, float* out)
{
        struct refh_run* r = h;
        for(int i = 0; i < r->tasks->n_tasks; i++){
                out[i] = r->tasks->list[i]->confidence;
        }
}

void refh_get_seq_distances(void* h, float* out)
{
        struct refh_run* r = h;
        memcpy(out, r->msa->seq_distances, sizeof(float) * r->msa->numseq);
}

void refh_get_params(void* h, float* subm_out, float* gp_out)
{
        struct refh_run* r = h;
        for(int i = 0; i < 23; i++){
                for(int j = 0; j < 23; j++){
                        subm_out[i * 23 + j] = r->ap->subm[i][j];
                }
        }
        gp_out[0] = r->ap->gpo; gp_out[1] = r->ap->gpe; gp_out[2] = r->ap->tgpe; gp_out[3] = r->ap->vsm_amax;
}

int refh_get_anchor_ids(void* h, int* out)
{
        struct refh_run* r = h;
        struct consistency_table* ct = r->msa->consistency_table;
        if(!ct){ return 0; }
        for(int k = 0; k < ct->n_anchors; k++){ out[k] = ct->anchor_ids[k]; }
        return ct->n_anchors;
}

/* position map of sorted sequence i against anchor slot k (len_i ints) */
int refh_get_posmap(void* h, int i, int k, int* out)
{
        struct refh_run* r = h;
        struct consistency_table* ct = r->msa->consistency_table;
        if(!ct){ return 1; }
        memcpy(out, ct->pos_maps[i * ct->n_anchors + k], sizeof(int) * r->msa->sequences[i]->len);
        return 0;
}

/* gaps[] of sorted sequence i (len_i+1 ints) */
void refh_get_gaps(void* h, int i, int* out)
{
        struct refh_run* r = h;
        memcpy(out, r->msa->sequences[i]->gaps, sizeof(int) * (r->msa->sequences[i]->len + 1));
}

/* aligned rows in ORIGINAL input order; each row alnlen chars + NUL, rows concatenated */
int refh_get_aligned(void* h, char* out)
{
        struct refh_run* r = h;
        int L = r->msa->alnlen;
        for(int i = 0; i < r->msa->numseq; i++){
                int rank = r->msa->sequences[i]->rank;
                memcpy(out + (size_t)rank * (L + 1), r->msa->sequences[i]->seq, L);
                out[(size_t)rank * (L + 1) + L] = 0;
        }
        return 0;
}

/* N x num_anchors distance matrix exactly as build_tree_kmeans computes it
   (bisectingKmeans.c:203-205); codes must be the tree alphabet => call with stop_after=1 handle
   BEFORE the protein re-encode is not possible, so this re-encodes a scratch msa. */
int refh_distance_matrix(char** seqs, int* lens, int n, float* dm_out, int* anchors_out, int* n_anchor_out)
{
        struct msa* msa = seqs_to_msa(seqs, lens, n);
        if(!msa){ return 1; }
        kalign_essential_input_check(msa, 0);
        msa_sort_len_name(msa);
        if(msa->biotype == ALN_BIOTYPE_DNA){
                convert_msa_to_internal(msa, ALPHA_defDNA);
        }else{
                convert_msa_to_internal(msa, ALPHA_redPROTEIN);
        }
        int na = 0;
        int* anchors = pick_anchor(msa, &na);
        float** dm = d_estimation(msa, anchors, na, 0);
        for(int i = 0; i < msa->numseq; i++){
                for(int j = 0; j < na; j++){
                        dm_out[i * na + j] = dm[i][j];
                }
#ifdef HAVE_AVX2
                _mm_free(dm[i]);
#else
                MFREE(dm[i]);
#endif
        }
        MFREE(dm);
        for(int j = 0; j < na; j++){ anchors_out[j] = anchors[j]; }
        *n_anchor_out = na;
        MFREE(anchors);
        kalign_free_msa(msa);
        return 0;
}

void refh_free(void* h)
{
        struct refh_run* r = h;
        if(!r){ return; }
        if(r->msa){
                if(r->msa->consistency_table){
                        anchor_consistency_free((struct consistency_table*)r->msa->consistency_table);
                        r->msa->consistency_table = NULL;
                }
                kalign_free_msa(r->msa);
        }
        if(r->ap){ aln_param_free(r->ap); }
        if(r->tasks){ free_tasks(r->tasks); }
        free(r);
}

/* Plain public-API run: kalign_run_seeded end to end (what the CLI does), timing only. */
double refh_time_public_api(char** seqs, int* lens, int n, int n_threads, int type,
                            int consistency_anchors, float consistency_weight)
{
        struct msa* msa = seqs_to_msa(seqs, lens, n);
        if(!msa){ return -1.0; }
        double t0 = now_s();
        int rc = kalign_run_seeded(msa, n_threads, type, -1.0f, -1.0f, -1.0f, KALIGN_REFINE_NONE, 0,
                                   0, 0.0f, 0.0f, -1.0f, -1.0f, consistency_anchors, consistency_weight);
        double t1 = now_s();
        kalign_free_msa(msa);
        return rc == OK ? (t1 - t0) : -1.0;
}

/* compute_aln_pairwise_dist (aln_apair_dist.c:9) on plain aligned rows: dm_out = n x n floats */
int refh_aln_pairwise_dist(char** rows, int n, int alnlen, float* dm_out)
{
        struct msa m;
        struct msa_seq* seqs = calloc((size_t)n, sizeof(struct msa_seq));
        struct msa_seq** ptr = malloc(sizeof(struct msa_seq*) * (size_t)n);
        float** dm = NULL;
        int rc = 1;
        if(!seqs || !ptr){ free(seqs); free(ptr); return 1; }
        memset(&m, 0, sizeof(m));
        for(int i = 0; i < n; i++){
                seqs[i].seq = rows[i];
                ptr[i] = &seqs[i];
        }
        m.sequences = ptr;
        m.numseq = n;
        m.alnlen = alnlen;
        m.aligned = ALN_STATUS_FINAL;
        if(compute_aln_pairwise_dist(&m, &dm) == OK){
                for(int i = 0; i < n; i++){
                        memcpy(dm_out + (size_t)i * (size_t)n, dm[i], sizeof(float) * (size_t)n);
                }
                free_aln_dm(dm, n);
                rc = 0;
        }
        free(seqs);
        free(ptr);
        return rc;
}

/* kalign_read_input (msa_io.c:80) on a file; the msa is returned as an opaque handle (NULL on failure) */
void* refh_read_input(const char* path)
{
        struct msa* msa = NULL;
        if(kalign_read_input((char*)path, &msa, 1) != OK){
                return NULL;
        }
        return msa;
}

int refh_msa_numseq(void* h){ return ((struct msa*)h)->numseq; }
int refh_msa_biotype(void* h){ return ((struct msa*)h)->biotype; }
int refh_msa_aligned(void* h){ return ((struct msa*)h)->aligned; }
int refh_msa_seq_len(void* h, int i){ return ((struct msa*)h)->sequences[i]->len; }
int refh_msa_name_len(void* h, int i){ return (int)strlen(((struct msa*)h)->sequences[i]->name); }

void refh_msa_record(void* h, int i, char* name_out, char* seq_out, int* gaps_out)
{
        struct msa_seq* s = ((struct msa*)h)->sequences[i];
        strcpy(name_out, s->name);
        memcpy(seq_out, s->seq, (size_t)s->len);
        memcpy(gaps_out, s->gaps, sizeof(int) * (size_t)(s->len + 1));
}

void refh_msa_letter_freq(void* h, int* out128)
{
        memcpy(out128, ((struct msa*)h)->letter_freq, sizeof(int) * 128);
}

void refh_msa_free(void* h)
{
        kalign_free_msa((struct msa*)h);
}

/* kalign_write_msa (msa_io.c:193) on plain names and finished rows */
int refh_write_rows(char** names, char** rows, int n, int alnlen, const char* path, const char* format)
{
        struct msa m;
        struct msa_seq* seqs = calloc((size_t)(n > 0 ? n : 1), sizeof(struct msa_seq));
        struct msa_seq** ptr = malloc(sizeof(struct msa_seq*) * (size_t)(n > 0 ? n : 1));
        int rc;
        if(!seqs || !ptr){ free(seqs); free(ptr); return 1; }
        memset(&m, 0, sizeof(m));
        for(int i = 0; i < n; i++){
                seqs[i].name = names[i];
                seqs[i].seq = rows[i];
                seqs[i].len = alnlen;
                ptr[i] = &seqs[i];
        }
        m.sequences = ptr;
        m.numseq = n;
        m.alnlen = alnlen;
        m.aligned = ALN_STATUS_FINAL;
        rc = kalign_write_msa(&m, (char*)path, (char*)format) == OK ? 0 : 1;
        free(seqs);
        free(ptr);
        return rc;
}

/* the noise factors of build_tree_kmeans_noisy (bisectingKmeans.c:104-116), from the reference's own generator */
int refh_tree_noise(uint64_t seed, float sigma, long long n, float* out)
{
        struct rng_state* rng = init_rng(seed);
        if(!rng){ return 1; }
        for(long long i = 0; i < n; i++){
                double noise = tl_random_gaussian(rng, 1.0, (double)sigma);
                if(noise < 0.1) noise = 0.1;
                out[i] = (float)noise;
        }
        free_rng(rng);
        return 0;
}

/* kalign_run_seeded (aln_wrap.c:133) end to end on plain arrays; rows in input order, concatenated with NULs.
   out must hold n * (cap + 1) bytes; returns the alignment length or -1 */
int refh_run_seeded(char** seqs, int* lens, int n, int n_threads, int type, float gpo, float gpe, float tgpe,
                    uint64_t tree_seed, float tree_noise, float dist_scale, float vsm_amax, float use_seq_weights,
                    int consistency_anchors, float consistency_weight, char* out, int cap)
{
        struct msa* msa = seqs_to_msa(seqs, lens, n);
        int L = -1;
        if(!msa){ return -1; }
        if(kalign_run_seeded(msa, n_threads, type, gpo, gpe, tgpe, KALIGN_REFINE_NONE, 0, tree_seed, tree_noise,
                             dist_scale, vsm_amax, use_seq_weights, consistency_anchors, consistency_weight) == OK
           && msa->alnlen <= cap){
                L = msa->alnlen;
                for(int i = 0; i < msa->numseq; i++){
                        memcpy(out + (size_t)i * (size_t)(L + 1), msa->sequences[i]->seq, (size_t)L);
                        out[(size_t)i * (size_t)(L + 1) + (size_t)L] = 0;
                }
        }
        kalign_free_msa(msa);
        return L;
}
