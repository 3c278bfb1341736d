/*
 * kalign_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's alignment hot path, written from the algorithm,
 * not from the reference's text: one direction-generic row sweep replaces the reference's six
 * forward/backward functions, one generic meet-up replaces its three.  Parity of this restatement
 * against the real reference (oracle/_ref, built from /root/reference) is pinned by
 * tests/test_oracle_vs_ref.py and the committed fixtures in tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (kalign_b200/) never does.
 */
#ifndef KALIGN_ORACLE_H
#define KALIGN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KO_KIND_SS 0   /* rows = sequence, cols = sequence   (aln_seqseq.c)        */
#define KO_KIND_SP 1   /* rows = profile,  cols = sequence   (aln_seqprofile.c)    */
#define KO_KIND_PP 2   /* rows = profile,  cols = profile    (aln_profileprofile.c)*/

typedef struct ko_job {
        int kind;
        const uint8_t* seq1;   /* row residues   (SS)            */
        const uint8_t* seq2;   /* column residues (SS, SP)       */
        const float* prof1;    /* row profile, (len_a+2)*64 (SP, PP) */
        const float* prof2;    /* col profile, (len_b+2)*64 (PP)     */
        int len_a;             /* DP rows */
        int len_b;             /* DP cols */
        int sip;               /* SP: number of sequences in the profile (aln_seqprofile.c:31-33) */
        const float* subm;     /* 23x23 row-major (SS) */
        float gpo, gpe, tgpe;  /* SS, SP */
        float soff;            /* subm_offset (SS) */
        const float* bonus;    /* dense len_a*len_b or NULL; flat index row*len_b + state_col */
} ko_job;

typedef struct ko_stats {
        double cells;          /* sum over sweeps of rows*(endb-startb) */
        float margin_sum;      /* aln_seqseq.c:376-383, accumulated in the reference's DFS order */
        int margin_count;
        float top_score;       /* score of the first (top level) meet-up */
        int n_boxes;
} ko_stats;

/* Full Hirschberg alignment (aln_controller.c:21-436).  path: len_a+2 ints, raw path
   (path[r], r=1..len_a = matched column 1..len_b or -1). */
int ko_align(const ko_job* job, int* path, ko_stats* st);

/* Myers/Hyyro block edit distance exactly as bpm_block (bpm.c:356-580). */
int ko_bpm_block(const uint8_t* t, const uint8_t* p, int n, int m);
/* calc_distance + length term (sequence_distance.c:120-123,153-162) */
float ko_pair_distance(const uint8_t* s1, int l1, const uint8_t* s2, int l2);

/* profile ops (aln_setup.c) */
void ko_make_profile(const uint8_t* seq, int len, const float* subm,
                     float gpo, float gpe, float tgpe, float soff, float* prof /* (len+2)*64 */);
void ko_set_gap_penalties(float* prof, int len, int nsip);
void ko_update(const float* profa, const float* profb, float* newp, const int* path,
               int sipa, int sipb, float gpo, float gpe, float tgpe);
/* raw path -> coded path (mirror_path_n + add_gap_info_to_path_n). path_io: len_a+len_b+2 ints. */
void ko_code_path(int* path_io, int len_a, int len_b, int mirror);
/* coded path -> position map of sequence a onto b (anchor_consistency.c:85-111) */
void ko_posmap_from_path(const int* path, int len_i, int* posmap);

#ifdef __cplusplus
}
#endif
#endif
