/*
 * kalign_oracle.c -- TEST INFRASTRUCTURE ONLY (see kalign_oracle.h).
 *
 * CPU restatement of the kalign 3.5.1 alignment hot path.  Every function cites the reference
 * file:line whose behaviour it restates.  The formulation is ours:
 *
 *   - ONE direction-generic sweep.  The reference has six functions
 *     (aln_seqseq.c:15,121  aln_seqprofile.c:13,125  aln_profileprofile.c:17,158).  Here a sweep
 *     runs over "logical" rows v=0..R-1 and logical columns u=0..C (C = endb-startb); a forward
 *     sweep maps (v,u) -> (starta+v, startb+u), a backward sweep maps (v,u) -> (enda-1-v, endb-u).
 *     Per-row and per-column gap terms are fetched through small accessors so that seq-seq,
 *     profile-seq and profile-profile share the recurrence
 *         A [v][u] = max3(A[v-1][u-1], GA[v-1][u-1]+CO(u-1), GB[v-1][u-1]+RO(v-1)) + match(v,u) [+bonus]
 *         GA[v][u] = max (GA[v][u-1]+CE(u),  A[v][u-1]+CO(u))
 *         GB[v][u] = max (GB[v-1][u]+RE(v),  A[v-1][u]+RO(v))
 *     with the terminal-gap and boundary-column rules of the reference.
 *   - x - y is evaluated as x + (-y): identical in IEEE-754.
 *   - MAX(a,b) is (a > b ? a : b) exactly as the reference's macro (aln_seqseq.c:12).
 *   - compiled with -ffp-contract=off: one rounded multiply, one rounded add per term
 *     (reference: -mavx2 without -mfma, CMakeLists.txt:192-193).
 */
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <math.h>

#include "kalign_oracle.h"

#define NEGF (-FLT_MAX)

typedef struct { float a, ga, gb; } kst;

static inline float mx(float a, float b) { return a > b ? a : b; }

/* ---- accessors ------------------------------------------------------------------------- */

typedef struct { float o, e, t, oprev; } gapterm;  /* open, extend, terminal, open of previous index */

/* gap terms attached to logical column u (1..C) of a box [sb,eb]; aln_seqseq.c:34-36,
   aln_seqprofile.c:31-33, aln_profileprofile.c:46,54,99,117 (fwd) / :188,196,242,260 (bwd). */
static inline gapterm col_terms(const ko_job* J, int bwd, int sb, int eb, int u)
{
        gapterm g;
        if(J->kind == KO_KIND_PP){
                int pc = bwd ? (eb - u + 1) : (sb + u);      /* profile column of this state column */
                int pp = bwd ? (pc + 1) : (pc - 1);          /* the one visited just before */
                const float* q = J->prof2 + ((size_t)pc << 6);
                g.o = q[27]; g.e = q[28]; g.t = q[29];
                g.oprev = J->prof2[((size_t)pp << 6) + 27];
        }else if(J->kind == KO_KIND_SP){
                float open = J->gpo * (float)J->sip;
                float ext  = J->gpe * (float)J->sip;
                float text = J->tgpe * (float)J->sip;
                g.o = -open; g.e = -ext; g.t = -text; g.oprev = -open;
        }else{
                g.o = -J->gpo; g.e = -J->gpe; g.t = -J->tgpe; g.oprev = -J->gpo;
        }
        return g;
}

/* gap terms attached to DP row i (0-based); profile column i+1 (aln_seqprofile.c:36,60). */
static inline gapterm row_terms(const ko_job* J, int bwd, int i)
{
        gapterm g;
        if(J->kind == KO_KIND_SS){
                g.o = -J->gpo; g.e = -J->gpe; g.t = -J->tgpe; g.oprev = -J->gpo;
        }else{
                int pr = i + 1;
                int pp = bwd ? (pr + 1) : (pr - 1);           /* prof1[91] / prof1[-37] */
                const float* p = J->prof1 + ((size_t)pr << 6);
                g.o = p[27]; g.e = p[28]; g.t = p[29];
                g.oprev = J->prof1[((size_t)pp << 6) + 27];
        }
        return g;
}

/* pa + match(i, state column j) [+ bonus]; residue index of state column j is j-1 in a forward
   sweep (seq2-- at aln_seqseq.c:60) and j in a backward sweep (aln_seqseq.c:198). */
static inline float add_match(const ko_job* J, int bwd, int i, int j, float pa)
{
        int r = bwd ? j : j - 1;
        if(J->kind == KO_KIND_SS){
                pa += J->subm[J->seq1[i] * 23 + J->seq2[r]] - J->soff;       /* aln_seqseq.c:82 */
        }else if(J->kind == KO_KIND_SP){
                pa += J->prof1[((size_t)(i + 1) << 6) + 32 + J->seq2[r]];    /* aln_seqprofile.c:81 */
        }else{
                const float* p = J->prof1 + ((size_t)(i + 1) << 6);
                const float* q = J->prof2 + ((size_t)(r + 1) << 6) + 32;
                for(int c = 22; c >= 0; c--){                                /* aln_profileprofile.c:71-76,102-106 */
                        if(p[c]){
                                pa += p[c] * q[c];
                        }
                }
        }
        if(J->bonus){
                pa += J->bonus[(size_t)i * J->len_b + j];                     /* aln_seqseq.c:83-85 */
        }
        return pa;
}

/* ---- the sweep --------------------------------------------------------------------------- */

/* Sweep rows [r0,r1) (forward: ascending, backward: descending) over state columns [sb,eb].
   S has C+1 entries indexed by LOGICAL column.  `in` is the injected boundary state
   (f[0] / b[0] of the reference, aln_controller.c:43-48). */
static void sweep(const ko_job* J, int bwd, int r0, int r1, int sb, int eb, kst in, kst* S)
{
        const int C = eb - sb;
        const int R = r1 - r0;
        const int first_term = bwd ? (eb == J->len_b) : (sb == 0);   /* aln_seqseq.c:43,155 */
        const int last_term  = bwd ? (sb == 0) : (eb == J->len_b);   /* aln_seqseq.c:112,231 */
        int u, v;

        S[0] = in;
        for(u = 1; u < C; u++){
                gapterm g = col_terms(J, bwd, sb, eb, u);
                S[u].a = NEGF;
                if(first_term){
                        S[u].ga = mx(S[u-1].ga, S[u-1].a) + g.t;
                }else{
                        S[u].ga = mx(S[u-1].ga + g.e, S[u-1].a + g.o);
                }
                S[u].gb = NEGF;
        }
        S[C].a = NEGF; S[C].ga = NEGF; S[C].gb = NEGF;

        for(v = 0; v < R; v++){
                const int i = bwd ? (r1 - 1 - v) : (r0 + v);
                const gapterm rt = row_terms(J, bwd, i);
                float pa = S[0].a, pga = S[0].ga, pgb = S[0].gb;
                float xa = NEGF, xga = NEGF, ca;
                S[0].a = NEGF;
                S[0].ga = NEGF;
                if(first_term){
                        S[0].gb = mx(pgb, pa) + rt.t;
                }else{
                        S[0].gb = mx(pgb + rt.e, pa + rt.o);
                }
                for(u = 1; u <= C; u++){
                        const int j = bwd ? (eb - u) : (sb + u);
                        const gapterm ct = col_terms(J, bwd, sb, eb, u);
                        ca = S[u].a;
                        pa = mx(mx(pa, pga + ct.oprev), pgb + rt.oprev);
                        pa = add_match(J, bwd, i, j, pa);
                        S[u].a = pa;
                        pga = S[u].ga;
                        S[u].ga = (u < C) ? mx(xga + ct.e, xa + ct.o) : NEGF;
                        pgb = S[u].gb;
                        if(u == C && last_term){
                                S[u].gb = mx(pgb, ca) + rt.t;
                        }else{
                                S[u].gb = mx(pgb + rt.e, ca + rt.o);
                        }
                        pa = ca;
                        xa = S[u].a;
                        xga = S[u].ga;
                }
        }
}

/* ---- meet-up (aln_seqseq.c:241, aln_seqprofile.c:232, aln_profileprofile.c:301) ---------- */

typedef struct { float max, max2; int c, t; } meet;

static inline void offer(meet* m, float s, int col, int t)
{
        if(s > m->max){
                m->max2 = m->max;
                m->max = s; m->c = col; m->t = t;
        }else if(s > m->max2){
                m->max2 = s;
        }
}

static meet meetup(const ko_job* J, int mid, int sb, int eb, const kst* F, const kst* B)
{
        meet m = { NEGF, NEGF, -1, -1 };
        const float middle = (float)(eb - sb) / 2.0F + (float)sb;
        float x2, x3, x5, x6, x6last, x7;
        const float* P = NULL;
        int i;
        if(J->kind == KO_KIND_SS){
                x2 = x3 = x5 = x7 = -J->gpo;
                x6 = (sb == 0) ? -J->tgpe : -J->gpe;
                x6last = (eb == J->len_b) ? -J->tgpe : -J->gpe;
        }else{
                P = J->prof1 + ((size_t)(mid + 1) << 6);
                x3 = P[27];
                x7 = P[-37];
                x6 = (sb == 0) ? P[29] : P[28];
                x6last = (eb == J->len_b) ? P[29] : P[28];
                x2 = x5 = -(J->gpo * (float)J->sip);      /* SP; overwritten per column for PP */
        }
        for(i = sb; i <= eb; i++){
                const kst f = F[i - sb];
                const kst b = B[eb - i];
                float sub = fabsf(middle - (float)i);
                sub /= 1000.0F;
                if(i < eb){
                        if(J->kind == KO_KIND_PP){
                                x2 = J->prof2[((size_t)(i + 1) << 6) + 27];
                                x5 = J->prof2[((size_t)i << 6) + 27];
                        }
                        offer(&m, f.a + b.a - sub, i, 1);
                        offer(&m, f.a + b.ga + x2 - sub, i, 2);
                        offer(&m, f.a + b.gb + x3 - sub, i, 3);
                        offer(&m, f.ga + b.a + x5 - sub, i, 5);
                        offer(&m, f.gb + b.gb + x6 - sub, i, 6);
                        offer(&m, f.gb + b.a + x7 - sub, i, 7);
                }else{
                        offer(&m, f.a + b.gb + x3 - sub, i, 3);
                        offer(&m, f.gb + b.gb + x6last - sub, i, 6);
                }
        }
        return m;
}

/* ---- Hirschberg controller (aln_controller.c:21-436) -------------------------------------- */

typedef struct {
        const ko_job* J;
        int* path;
        kst* F;
        kst* B;
        ko_stats* st;
        int depth;
} hctx;

static const kst K_A  = { 0.0F, NEGF, NEGF };
static const kst K_GA = { NEGF, 0.0F, NEGF };
static const kst K_GB = { NEGF, NEGF, 0.0F };

static void hirsch(hctx* h, int sa, int ea, int sb, int eb, kst fin, kst bin)
{
        if(sa >= ea || sb >= eb){
                return;
        }
        const int mid = (ea - sa) / 2 + sa;
        /* the reference keeps one f[] and one b[] array indexed by absolute column; children are
           processed strictly after the parent's meet-up, so private scratch per call is equivalent */
        kst* F = h->F + sb;     /* logical u -> F[u]         */
        kst* B = h->B + sb;     /* logical u -> B[u]  (own array, no aliasing with F) */
        sweep(h->J, 0, sa, mid, sb, eb, fin, F);
        sweep(h->J, 1, mid, ea, sb, eb, bin, B);
        meet m = meetup(h->J, mid, sb, eb, F, B);
        if(h->st){
                h->st->cells += (double)(ea - sa) * (double)(eb - sb);
                h->st->n_boxes++;
                if(h->depth == 0){
                        h->st->top_score = m.max;
                }
                if(m.max2 > NEGF){
                        h->st->margin_sum += m.max - m.max2;
                        h->st->margin_count++;
                }
        }
        h->depth++;
        int* path = h->path;
        const int c = m.c;
        switch(m.t){
        case 1:
                path[mid] = c; path[mid + 1] = c + 1;
                hirsch(h, sa, mid - 1, sb, c - 1, fin, K_A);
                hirsch(h, mid + 1, ea, c + 1, eb, K_A, bin);
                break;
        case 2:
                path[mid] = c;
                hirsch(h, sa, mid - 1, sb, c - 1, fin, K_A);
                hirsch(h, mid, ea, c + 1, eb, K_GA, bin);
                break;
        case 3:
                path[mid] = c;
                hirsch(h, sa, mid - 1, sb, c - 1, fin, K_A);
                hirsch(h, mid + 1, ea, c, eb, K_GB, bin);
                break;
        case 5:
                path[mid + 1] = c + 1;
                hirsch(h, sa, mid, sb, c - 1, fin, K_GA);
                hirsch(h, mid + 1, ea, c + 1, eb, K_A, bin);
                break;
        case 6:
                hirsch(h, sa, mid - 1, sb, c, fin, K_GB);
                hirsch(h, mid + 1, ea, c, eb, K_GB, bin);
                break;
        case 7:
                path[mid + 1] = c + 1;
                hirsch(h, sa, mid - 1, sb, c, fin, K_GB);
                hirsch(h, mid + 1, ea, c + 1, eb, K_A, bin);
                break;
        default:
                break;
        }
        h->depth--;
}

int ko_align(const ko_job* job, int* path, ko_stats* st)
{
        hctx h;
        int g = (job->len_a > job->len_b ? job->len_a : job->len_b) + 2;
        int i;
        h.J = job;
        h.path = path;
        h.st = st;
        h.depth = 0;
        /* every recursion level re-initialises the slice it uses, but a child may still be
           reading its parent's... no: parent rows are dead once the meet-up is taken. */
        h.F = malloc(sizeof(kst) * (size_t)(job->len_b + 2));
        h.B = malloc(sizeof(kst) * (size_t)(job->len_b + 2));
        if(!h.F || !h.B){
                free(h.F); free(h.B);
                return 1;
        }
        if(st){
                st->cells = 0.0; st->margin_sum = 0.0F; st->margin_count = 0;
                st->top_score = 0.0F; st->n_boxes = 0;
        }
        for(i = 0; i < g && i < job->len_a + 2; i++){   /* init_alnmem, aln_setup.c:31-34 */
                path[i] = -1;
        }
        hirsch(&h, 0, job->len_a, 0, job->len_b, K_A, K_A);
        free(h.F);
        free(h.B);
        return 0;
}

/* ---- Myers / Hyyro block edit distance (bpm.c:356-580) ----------------------------------- */
/*
 * The reference carries Edlib-style band bookkeeping (y may shrink / grow), but it starts with
 * every block active (y = ceil(maxd/64)-1 with maxd = m, bpm.c:443) and the shrink test
 * score[y] >= maxd + 64 (bpm.c:556) can never fire because score[b] <= 64(b+1) < m + 64 for the
 * last block; the grow test needs y < b_max-1 (bpm.c:510).  So the value is the plain multi-word
 * recurrence with all blocks live: min(m, min_i score[last block] after text char i), text
 * extended by W = 64*b_max - m zero symbols, pattern positions >= m matching everything.
 * tests/test_oracle_vs_ref.py checks this claim against the real bpm_block.
 */
int ko_bpm_block(const uint8_t* t, const uint8_t* p, int n, int m)
{
        uint64_t Peq[13][16];
        uint64_t P[16], M[16];
        int score[16];
        int b, i, c;
        if(m > 1024){
                m = 1024;
        }
        const int bmax = (m == 0) ? 1 : (m / 64 + ((m % 64) ? 1 : 0));
        const int W = 64 * bmax - m;
        int k = m;
        memset(Peq, 0, sizeof(Peq));
        for(c = 0; c < 13; c++){
                for(i = 0; i < 64 * bmax; i++){
                        if(i >= m || p[i] == c){
                                Peq[c][i >> 6] |= ((uint64_t)1) << (i & 63);
                        }
                }
        }
        for(b = 0; b < bmax; b++){
                P[b] = ~(uint64_t)0;
                M[b] = 0;
                score[b] = (b + 1) * 64;
        }
        for(i = 0; i < n + W; i++){
                const int ch = (i < n) ? t[i] : 0;
                int hin = 0;
                for(b = 0; b < bmax; b++){
                        uint64_t Pv = P[b], Mv = M[b], Eq = Peq[ch][b];
                        const uint64_t Xv = Eq | Mv;
                        if(hin < 0){
                                Eq |= 1;
                        }
                        const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
                        uint64_t Ph = Mv | ~(Xh | Pv);
                        uint64_t Mh = Pv & Xh;
                        const int hout = (int)(Ph >> 63) - (int)(Mh >> 63);
                        Ph <<= 1;
                        Mh <<= 1;
                        if(hin < 0){
                                Mh |= 1;
                        }else if(hin > 0){
                                Ph |= 1;
                        }
                        P[b] = Mh | ~(Xv | Ph);
                        M[b] = Ph & Xv;
                        score[b] += hout;
                        hin = hout;
                }
                if(score[bmax - 1] < k){
                        k = score[bmax - 1];
                }
        }
        return k;
}

/* sequence_distance.c:153-162 (longer = text; equal length: second argument = text) and
   :120-123 (length term, computed in double, added in float). */
float ko_pair_distance(const uint8_t* s1, int l1, const uint8_t* s2, int l2)
{
        int d;
        if(l1 > l2){
                d = ko_bpm_block(s1, s2, l1, l2);
        }else{
                d = ko_bpm_block(s2, s1, l2, l1);
        }
        float dist = (float)(uint32_t)d;
        int s = (l1 + l2) / 2;
        float add = (float)(((10000.0 < (double)s) ? 10000.0 : (double)s) / 10000.0);
        dist += add;
        return dist;
}

/* ---- profile ops (aln_setup.c) ------------------------------------------------------------ */

/* make_profile_n, aln_setup.c:40-99 (weight is always 1.0, aln_run.c:207-211) */
void ko_make_profile(const uint8_t* seq, int len, const float* subm,
                     float gpo, float gpe, float tgpe, float soff, float* prof)
{
        int i, j;
        memset(prof, 0, sizeof(float) * 64 * (size_t)(len + 2));
        for(i = 0; i <= len + 1; i++){
                float* col = prof + ((size_t)i << 6);
                if(i >= 1 && i <= len){
                        const int c = seq[i - 1];
                        col[c] += 1.0f;
                        for(j = 0; j < 23; j++){
                                col[32 + j] = subm[c * 23 + j] - soff;
                        }
                }
                col[55] = -gpo;
                col[56] = -gpe;
                col[57] = -tgpe;
        }
}

/* set_gap_penalties_n, aln_setup.c:101-119: every column 0..len+1 */
void ko_set_gap_penalties(float* prof, int len, int nsip)
{
        for(int i = 0; i <= len + 1; i++){
                float* col = prof + ((size_t)i << 6);
                col[27] = col[55] * (float)nsip;
                col[28] = col[56] * (float)nsip;
                col[29] = col[57] * (float)nsip;
        }
}

/* gap column bookkeeping of update_n (aln_setup.c:321-365 and its mirror :374-417) */
static void gap_adjust(float* np, int p, float sip, float gpo, float gpe, float tgpe)
{
        float gp;
        int j, rep;
        if(!(p & 20)){
                if(p & 32){
                        np[25] += sip;
                        gp = tgpe * sip;
                }else{
                        np[24] += sip;
                        gp = gpe * sip;
                }
                for(j = 32; j < 55; j++){
                        np[j] -= gp;
                }
                return;
        }
        for(rep = 0; rep < 2; rep++){
                if(!(p & (rep == 0 ? 16 : 4))){
                        continue;
                }
                if(p & 32){
                        np[25] += sip;
                        gp = tgpe * sip;
                        np[23] += sip;
                        gp += gpo * sip;
                }else{
                        np[23] += sip;
                        gp = gpo * sip;
                }
                for(j = 32; j < 55; j++){
                        np[j] -= gp;
                }
        }
}

/* update_n without sequence-weight rebalancing (use_seq_weights = 0, aln_param.c:100) */
void ko_update(const float* pa, const float* pb, float* np, const int* path,
               int sipa, int sipb, float gpo, float gpe, float tgpe)
{
        int c, i;
        for(i = 0; i < 64; i++){
                np[i] = pa[i] + pb[i];
        }
        pa += 64; pb += 64; np += 64;
        for(c = 1; path[c] != 3; c++){
                const int p = path[c];
                if(!p){
                        for(i = 0; i < 64; i++){
                                np[i] = pa[i] + pb[i];
                        }
                        pa += 64; pb += 64;
                }
                if(p & 1){
                        memcpy(np, pb, sizeof(float) * 64);
                        pb += 64;
                        gap_adjust(np, p, (float)sipa, gpo, gpe, tgpe);
                }
                if(p & 2){
                        memcpy(np, pa, sizeof(float) * 64);
                        pa += 64;
                        gap_adjust(np, p, (float)sipb, gpo, gpe, tgpe);
                }
                np += 64;
        }
        for(i = 0; i < 64; i++){
                np[i] = pa[i] + pb[i];
        }
}

/* mirror_path_n (aln_setup.c:438-462) then add_gap_info_to_path_n (aln_setup.c:121-228).
   NB: the reference's "add gap info" loop tests o_path[j] (the terminator it has just written,
   aln_setup.c:191-195) so it never runs: open/extend/close bits 4/8/16 are never set.  Only the
   terminal bit 32 is applied to leading and trailing gap entries (aln_setup.c:212-222). */
void ko_code_path(int* path_io, int len_a, int len_b, int mirror)
{
        const int n = len_a + len_b + 2;
        int* raw = malloc(sizeof(int) * (size_t)n);
        int* o = calloc((size_t)n + 1, sizeof(int));
        int i, j, a, b;
        if(mirror){
                for(i = 0; i < len_a + 2; i++){
                        raw[i] = -1;
                }
                for(i = 1; i <= len_b; i++){
                        if(path_io[i] != -1){
                                raw[path_io[i]] = i;
                        }
                }
        }else{
                memcpy(raw, path_io, sizeof(int) * (size_t)(len_a + 2));
        }
        j = 1;
        b = -1;
        for(i = 1; i <= len_a; i++){
                if(raw[i] == -1){
                        o[j++] = 2;
                }else{
                        int skip;
                        if(i == 1){
                                skip = raw[1] - 1;                   /* aln_setup.c:147-157 */
                        }else if(raw[i] - 1 != b && b != -1){
                                skip = raw[i] - b - 1;               /* aln_setup.c:167-173 */
                        }else{
                                skip = 0;
                        }
                        for(a = 0; a < skip; a++){
                                o[j++] = 1;
                        }
                        o[j++] = 0;
                }
                b = raw[i];
        }
        if(raw[len_a] < len_b && raw[len_a] != -1){
                for(a = 0; a < len_b - raw[len_a]; a++){
                        o[j++] = 1;
                }
        }
        o[0] = j - 1;
        o[j] = 3;
        i = 1;
        while(i < n && o[i] != 0){
                o[i] |= 32;
                i++;
        }
        i = o[0];
        while(i > 0 && o[i] != 0){
                o[i] |= 32;
                i--;
        }
        memcpy(path_io, o, sizeof(int) * (size_t)n);
        free(raw);
        free(o);
}

/* anchor_consistency.c:85-111 */
void ko_posmap_from_path(const int* path, int len_i, int* posmap)
{
        int pos_a = 0, pos_b = 0, c;
        for(c = 0; c < len_i; c++){
                posmap[c] = -1;
        }
        for(c = 1; path[c] != 3; c++){
                if(path[c] == 0){
                        if(pos_a < len_i){
                                posmap[pos_a] = pos_b;
                        }
                        pos_a++; pos_b++;
                }else if(path[c] & 1){
                        pos_b++;
                }else if(path[c] & 2){
                        if(pos_a < len_i){
                                posmap[pos_a] = -1;
                        }
                        pos_a++;
                }
        }
}
