// kb_dp.cu -- batched Hirschberg engine: anti-diagonal wavefront sweeps + meet-up + box recursion.
//
// Replaces (behaviour cited, nothing copied):
//   aln_seqseq_foward/backward          lib/src/aln_seqseq.c:15,121
//   aln_seqprofile_foward/backward      lib/src/aln_seqprofile.c:13,125
//   aln_profileprofile_foward/backward  lib/src/aln_profileprofile.c:17,158
//   aln_*_meetup                        lib/src/aln_seqseq.c:241, aln_seqprofile.c:232, aln_profileprofile.c:301
//   aln_runner / aln_continue           lib/src/aln_controller.c:21,194
//
// Formulation (identical to oracle/kalign_oracle.c): a sweep runs over logical rows v and logical
// columns u = 0..C of a box; forward maps (v,u)->(starta+v, startb+u), backward maps
// (v,u)->(enda-1-v, endb-u).  Cell recurrence, operands and operation order are exactly the
// reference's (x-y is evaluated as x+(-y); compiled with -fmad=false so every multiply and add is
// rounded separately; max is evaluated with fmaxf, value-identical to the reference's
// (a>b?a:b) because no NaN can occur and the sign of a zero never reaches a comparison; the
// profile-profile dot product runs densely over the alphabet, adding exact zeros for the residues
// the reference's sparse list skips -- value-identical for the same reason).
//
// Parallel shape.  The rows of a (box, direction) sweep are cut into STRIPS of 32*K rows; a strip
// is swept by one warp: lane l owns K consecutive rows and walks the columns with a skew of one
// column per lane (anti-diagonal wavefront); the in-diagonal hand-off of the row above is a warp
// shuffle.  The row that leaves a strip (H/E/F = a/ga/gb, one float4 per column) is streamed
// through the job's row buffer in global memory; the NEXT strip of the same sweep -- run
// concurrently by another warp, possibly on another SM -- consumes it a few columns behind
// (each float4 carries a per-launch strip tag in its 4th word: data and "ready" flag travel in ONE
// 16-byte store, no fences, no L1 invalidation), so one big box is spread over many SMs
// (pipelined multi-CTA wavefront) while a batch of many boxes simply fills the machine with
// independent strips.  Work units (box, direction, strip) are handed out in order by an atomic
// cursor to a persistent grid: a strip only ever waits for a unit that was handed out earlier.
// All boxes of one Hirschberg depth of ALL jobs of a batch form one round: plan -> sweep ->
// meet-up (which emits the child boxes of aln_continue into the next round's work-list).
#include "kb_common.cuh"

#include <algorithm>
#include <stdlib.h>
#include <type_traits>

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int WARPS_PER_CTA = 4;
constexpr int TBL_MAX = 23 * 32;     // shared table capacity (floats)

// kernel variants
// V_SP5: profile(rows) x sequence over a 5-letter alphabet -- the rows' score vectors are staged in
// shared memory (lane-private, conflict-free) instead of being gathered from global memory per cell
// V_SS5: sequence x sequence over a 5-letter alphabet -- each row's five (score - offset) values are
// staged the same way, which removes the per-cell table-address arithmetic and the offset add
enum { V_SS = 0, V_SP = 1, V_PP5 = 2, V_PP23 = 3, V_SP5 = 4, V_SS5 = 5 };
template <int V> __host__ __device__ constexpr bool is_ss() { return V == V_SS || V == V_SS5; }
constexpr int PACK5 = 8;             // packed column record, 5-letter alphabets: s0..s4, [27],[28],[29]
constexpr int PACK23 = 28;           // 23-letter: s0..s22, [27],[28],[29], pad, pad

template <int V> struct VTraits;
template <> struct VTraits<V_SS> { static constexpr int NA = 0; };
template <> struct VTraits<V_SP> { static constexpr int NA = 0; };
template <> struct VTraits<V_SP5> { static constexpr int NA = 0; };
template <> struct VTraits<V_SS5> { static constexpr int NA = 0; };
template <> struct VTraits<V_PP5> { static constexpr int NA = 5; };
template <> struct VTraits<V_PP23> { static constexpr int NA = 23; };

struct Trip {
        float a, ga, gb;
};

struct KbUnit {
        int item;    // box * 2 + direction
        int strip;
};

__device__ __forceinline__ float kmax(float a, float b) { return fmaxf(a, b); }

// MODE_EDGE: one routine for the fill / drain steps of a strip, where the lanes of a warp sit on
// different kinds of column (first / interior / last): the column kind is a per-lane flag resolved
// with selects instead of three divergent code paths.
enum { MODE_FIRST = 0, MODE_MID = 1, MODE_LAST = 2, MODE_EDGE = 3 };
// consistency bonus of a batch: none / sparse per-row lists (tree levels) / caller-supplied dense matrix
enum { BONUS_NONE = 0, BONUS_SPARSE = 1, BONUS_DENSE = 2 };

// rows per strip of a job in a round: "wide" rounds (few big boxes) use thin strips so that one
// box spreads over many warps; otherwise thick strips amortise the per-step overhead.
__device__ __host__ __forceinline__ int rows_per_strip(int kind, int nalpha, int thin, bool has_bonus)
{
        if (thin) {
                return 32;
        }
        if (kind == KB200_KIND_PP && nalpha > 5) {
                return 64;
        }
        if (kind == KB200_KIND_SS && !has_bonus) {
                return 256;          // K = 8 rows per lane; the bonus variants keep K = 4 (register budget)
        }
        return 128;
}

// ---------------------------------------------------------------------------------------------
template <int V, int K> struct RowCtx {
        int rbase[K];                                   // SS
        const float* prow[K];                           // SP
        const float* sprow;                             // SP5: this lane's slot of the staged score vectors [k][letter][lane]
        float cnt[K][VTraits<V>::NA > 0 ? VTraits<V>::NA : 1];   // PP: residue counts of the row
        float RO[K], RE[K], RT[K], ROp[K];              // SP, PP
        int irow[K];
};

template <int V> struct ColCtx {
        int cres;                                       // SS, SP
        float qs[VTraits<V>::NA > 0 ? VTraits<V>::NA : 1];       // PP: scores of the column
        float CO, CE, COp;
        int jcol;
};

// Sparse consistency bonus of a row: <= nb (column, value) entries sorted by column.  The warp
// strips stage the lists of their rows in shared memory (lane-private slots, [entry][row][lane]:
// conflict-free), so that stepping to the next entry after a hit is an LDS instead of a dependent
// L2 access that would stall the whole warp; the thread-per-box kernel reads them from global.
constexpr int BON_SLOTS = KB_BONUS_KMAX;
constexpr int BON_KMAX_ROWS = 4;      // rows per lane of the widest strip that carries a bonus

template <int K, bool BSM>
__device__ __forceinline__ void bonus_entry(const KbJob& J, const int2* __restrict__ s_bon, const int irow, const int k, const int e,
                                            int& c, float& v)
{
        if constexpr (BSM) {
                const int2 q = s_bon[(e * K + k) * 32];
                c = q.x;
                v = __int_as_float(q.y);
        } else {
                const size_t o = (size_t)irow * (size_t)J.nb + (size_t)e;
                c = __ldg(J.bkey + o);
                v = __ldg(J.bval + o);
        }
}

// Packed fp32x2 addition (Blackwell FADD2): two independent IEEE round-to-nearest
// additions per instruction -- the same values as two scalar operations, at half the issue slots
// and half the load on the FP pipe that bounds the sweep.
__device__ __forceinline__ float2 add2(const float2 a, const float2 b)
{
        unsigned long long ra, rb, rd;
        asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
        asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
        float2 d;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
        return d;
}

// NOTE: there is deliberately no packed multiply here.  ptxas contracts mul.rn.f32x2 (and even
// fma.rn.f32x2 with a -0 addend) followed by add.rn.f32x2 into ONE FFMA2 -- a single rounding --
// regardless of -fmad=false, which would break bit-identity with the reference's separately rounded
// multiply and add (aln_profileprofile.c:99-106).  Products are therefore scalar __fmul_rn; only
// the additions are packed.  tests/test_sass_contract.py checks the SASS for fused multiply-adds.

template <int V, int K, bool TAIL, int MODE, int BONUS, bool BSM>
__device__ __forceinline__ void cells(const KbJob& J, const RowCtx<V, K>& rc, const unsigned vmask,
                                      const bool first_term, const bool last_term,
                                      const ColCtx<V>& cc, const float (&bon)[K],
                                      const bool sparse, const int bdir, int (&sp_i)[K], int (&sp_c)[K], float (&sp_v)[K],
                                      const float (&sp_wrap)[K], const int2* __restrict__ s_bon,
                                      const float* __restrict__ s_tbl,
                                      float (&sA)[K], float (&sGA)[K], float (&sGB)[K],
                                      Trip d, Trip& u, const bool e_first = false, const bool e_last = false);

// Interior columns, rows taken two at a time: the same operations on the same operands as the
// scalar routine below (every add / multiply is still rounded on its own), issued as packed pairs:
//   (oGA_k + COp, oGA_k + CE)            -> diagonal term of row k+1 and the row's own ga
//   (oA_k + CO, oA_k+1 + CO)             -> ga of both rows
//   (oGB_k + ROp_k+1, oGB_k+1 + ROp_k+2) -> diagonal terms of rows k+1, k+2
//   (m_k + x_k, m_k+1 + x_k+1) and the profile-profile dot product of both rows
// Only the gb chain down the column (row k+1 needs the new a / gb of row k) stays scalar.
template <int V, int K, bool TAIL, int BONUS, bool BSM, bool EDGE>
__device__ __forceinline__ void cells_mid2(const KbJob& J, const RowCtx<V, K>& rc, const unsigned vmask,
                                           const ColCtx<V>& cc,
                                           const int bdir, int (&sp_i)[K], int (&sp_c)[K], float (&sp_v)[K],
                                           const int2* __restrict__ s_bon, const float* __restrict__ s_tbl,
                                           float (&sA)[K], float (&sGA)[K], float (&sGB)[K],
                                           const Trip d, Trip& u,
                                           const bool first_term, const bool last_term, const float (&sp_wrap)[K],
                                           const bool e_first, const bool e_last)
{
        // EDGE: e_first / e_last flag the lane's column as the first / last one of the box
        const bool e_any = EDGE && (e_first || e_last);
        const bool e_term = EDGE && ((e_first && first_term) || (e_last && last_term));
        static_assert(K % 2 == 0, "rows in pairs");
        constexpr int NA = VTraits<V>::NA;
        unsigned hits = 0;
        const float2 colGA = make_float2(cc.COp, cc.CE);
        const float2 colCO = make_float2(cc.CO, cc.CO);
        float ROp0;
        if constexpr (is_ss<V>()) {
                ROp0 = J.o;
        } else {
                ROp0 = rc.ROp[0];
        }
        // diagonal of row 0: the lane above, one column back
        float dA = d.a;
        float tGA = d.ga + cc.COp;
        float tGB = d.gb + ROp0;
#pragma unroll
        for (int k = 0; k < K; k += 2) {
                const float oA0 = sA[k], oA1 = sA[k + 1];
                const float oGA0 = sGA[k], oGA1 = sGA[k + 1];
                const float oGB0 = sGB[k], oGB1 = sGB[k + 1];
                float RO0, RE0, RO1, RE1, ROpA, ROpB;
                if constexpr (is_ss<V>()) {
                        RO0 = RO1 = J.o; RE0 = RE1 = J.e; ROpA = ROpB = J.o;
                } else {
                        RO0 = rc.RO[k]; RE0 = rc.RE[k]; RO1 = rc.RO[k + 1]; RE1 = rc.RE[k + 1];
                        ROpA = rc.ROp[k + 1];
                        ROpB = rc.ROp[(k + 2 < K) ? (k + 2) : k];        // last pair: second half unused
                }
                const float2 g0 = add2(make_float2(oGA0, oGA0), colGA);      // .x: + COp (row k+1), .y: + CE (own ga)
                const float2 g1 = add2(make_float2(oGA1, oGA1), colGA);
                const float2 h = add2(make_float2(oA0, oA1), colCO);
                const float2 b = add2(make_float2(oGB0, oGB1), make_float2(ROpA, ROpB));
                const float m0 = kmax(kmax(dA, tGA), tGB);
                const float m1 = kmax(kmax(oA0, g0.x), b.x);
                float2 a01 = make_float2(m0, m1);
                if constexpr (V == V_SS) {
                        const float2 x = add2(make_float2(s_tbl[rc.rbase[k] + cc.cres], s_tbl[rc.rbase[k + 1] + cc.cres]),
                                              make_float2(J.nsoff, J.nsoff));
                        a01 = add2(a01, x);
                } else if constexpr (V == V_SP) {
                        a01 = add2(a01, make_float2(__ldg(rc.prow[k] + 32 + cc.cres), __ldg(rc.prow[k + 1] + 32 + cc.cres)));
                } else if constexpr (V == V_SP5 || V == V_SS5) {
                        const float* sv = rc.sprow + cc.cres * 32;
                        a01 = add2(a01, make_float2(sv[k * 160], sv[(k + 1) * 160]));
                } else {
#pragma unroll
                        for (int c = NA - 1; c >= 0; c--) {
                                const float2 p = make_float2(__fmul_rn(rc.cnt[k][c], cc.qs[c]), __fmul_rn(rc.cnt[k + 1][c], cc.qs[c]));
                                a01 = add2(a01, p);
                        }
                }
                if constexpr (BONUS == BONUS_SPARSE) {
                        const bool hit0 = (cc.jcol == sp_c[k]) && !(EDGE && e_first);
                        const bool hit1 = (cc.jcol == sp_c[k + 1]) && !(EDGE && e_first);
                        a01 = add2(a01, make_float2(hit0 ? sp_v[k] : 0.0f, hit1 ? sp_v[k + 1] : 0.0f));
                        hits |= (hit0 ? (1u << k) : 0u) | (hit1 ? (2u << k) : 0u);
                        if constexpr (EDGE) {
                                if (e_last) {
                                        a01 = add2(a01, make_float2(sp_wrap[k], sp_wrap[k + 1]));   // j == len_b wraps to (i+1, 0)
                                }
                        }
                } else if constexpr (BONUS == BONUS_DENSE) {
                        if (J.bonus && !(EDGE && e_first)) {
                                a01 = add2(a01, make_float2(__ldg(J.bonus + (size_t)rc.irow[k] * (size_t)J.len_b + (size_t)cc.jcol),
                                                            __ldg(J.bonus + (size_t)rc.irow[k + 1] * (size_t)J.len_b + (size_t)cc.jcol)));
                        }
                }
                float a0 = a01.x, a1 = a01.y;
                float ga0 = kmax(g0.y, h.x);
                float ga1 = kmax(g1.y, h.y);
                float gb0 = kmax(u.gb + RE0, u.a + RO0);
                if constexpr (EDGE) {
                        float RT0;
                        if constexpr (is_ss<V>()) {
                                RT0 = J.t;
                        } else {
                                RT0 = rc.RT[k];
                        }
                        const float gt0 = kmax(u.gb, u.a) + RT0;
                        gb0 = e_term ? gt0 : gb0;
                        a0 = e_first ? KB_NEGF : a0;
                        ga0 = e_any ? KB_NEGF : ga0;
                }
                if constexpr (TAIL) {
                        if (!((vmask >> k) & 1u)) {
                                a0 = u.a; ga0 = u.ga; gb0 = u.gb;
                        }
                }
                float gb1 = kmax(gb0 + RE1, a0 + RO1);
                if constexpr (EDGE) {
                        float RT1;
                        if constexpr (is_ss<V>()) {
                                RT1 = J.t;
                        } else {
                                RT1 = rc.RT[k + 1];
                        }
                        const float gt1 = kmax(gb0, a0) + RT1;
                        gb1 = e_term ? gt1 : gb1;
                        a1 = e_first ? KB_NEGF : a1;
                        ga1 = e_any ? KB_NEGF : ga1;
                }
                if constexpr (TAIL) {
                        if (!((vmask >> (k + 1)) & 1u)) {
                                a1 = a0; ga1 = ga0; gb1 = gb0;
                        }
                }
                sA[k] = a0; sGA[k] = ga0; sGB[k] = gb0;
                sA[k + 1] = a1; sGA[k + 1] = ga1; sGB[k + 1] = gb1;
                u.a = a1; u.ga = ga1; u.gb = gb1;
                dA = oA1; tGA = g1.x; tGB = b.y;
        }
        if constexpr (BONUS == BONUS_SPARSE) {
                if (hits) {
#pragma unroll
                        for (int k = 0; k < K; k++) {
                                if ((hits >> k) & 1u) {
                                        sp_i[k] += bdir;
                                        const int e = sp_i[k];
                                        const bool ok = (e >= 0) && (e < J.nb);
                                        int nc;
                                        float nv;
                                        bonus_entry<K, BSM>(J, s_bon, rc.irow[k], k, ok ? e : 0, nc, nv);
                                        sp_c[k] = ok ? nc : ((bdir > 0) ? 0x7fffffff : -1);
                                        sp_v[k] = ok ? nv : 0.0f;
                                }
                        }
                }
        }
}

template <int V, int K, bool TAIL, int MODE, int BONUS, bool BSM>
__device__ __forceinline__ void cells(const KbJob& J, const RowCtx<V, K>& rc, const unsigned vmask,
                                      const bool first_term, const bool last_term,
                                      const ColCtx<V>& cc, const float (&bon)[K],
                                      const bool sparse, const int bdir, int (&sp_i)[K], int (&sp_c)[K], float (&sp_v)[K],
                                      const float (&sp_wrap)[K], const int2* __restrict__ s_bon,
                                      const float* __restrict__ s_tbl,
                                      float (&sA)[K], float (&sGA)[K], float (&sGB)[K],
                                      Trip d, Trip& u /* in: up at column u; out: bottom row */,
                                      const bool e_first, const bool e_last)
{
        if constexpr ((MODE == MODE_MID || MODE == MODE_EDGE) && (K % 2) == 0) {
                cells_mid2<V, K, TAIL, BONUS, BSM, MODE == MODE_EDGE>(J, rc, vmask, cc, bdir, sp_i, sp_c, sp_v, s_bon, s_tbl, sA, sGA, sGB,
                                                                      d, u, first_term, last_term, sp_wrap, e_first, e_last);
                return;
        }
        const bool e_any = (MODE == MODE_EDGE) && (e_first || e_last);
        const bool e_term = (MODE == MODE_EDGE) && ((e_first && first_term) || (e_last && last_term));
        unsigned hits = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
                const float oA = sA[k], oGA = sGA[k], oGB = sGB[k];
                float RO, RE, RT, ROp;
                if constexpr (is_ss<V>()) {
                        RO = J.o; RE = J.e; RT = J.t; ROp = J.o;
                } else {
                        RO = rc.RO[k]; RE = rc.RE[k]; RT = rc.RT[k]; ROp = rc.ROp[k];
                }
                float a, ga, gb;
                if constexpr (MODE == MODE_FIRST) {
                        a = KB_NEGF;
                        ga = KB_NEGF;
                        gb = first_term ? (kmax(u.gb, u.a) + RT) : kmax(u.gb + RE, u.a + RO);
                } else {
                        a = kmax(kmax(d.a, d.ga + cc.COp), d.gb + ROp);
                        if constexpr (V == V_SS) {
                                const float x = s_tbl[rc.rbase[k] + cc.cres] + J.nsoff;
                                a = a + x;
                        } else if constexpr (V == V_SP) {
                                a = a + __ldg(rc.prow[k] + 32 + cc.cres);
                        } else if constexpr (V == V_SP5 || V == V_SS5) {
                                a = a + rc.sprow[k * 160 + cc.cres * 32];
                        } else {
#pragma unroll
                                for (int c = VTraits<V>::NA - 1; c >= 0; c--) {
                                        a = __fadd_rn(a, __fmul_rn(rc.cnt[k][c], cc.qs[c]));
                                }
                        }
                        if constexpr (BONUS == BONUS_SPARSE) {
                                // sorted per-row list walked in sweep direction: at most nb hits per row.
                                // The reference adds its dense matrix entry to EVERY cell
                                // (aln_seqseq.c:83-85): + 0.0f where the list has no entry.  A job
                                // without a list keeps the never-matching sentinel in sp_c.
                                const bool hit = (cc.jcol == sp_c[k]) && !(MODE == MODE_EDGE && e_first);
                                a = a + (hit ? sp_v[k] : 0.0f);
                                hits |= hit ? (1u << k) : 0u;
                                if constexpr (MODE == MODE_LAST) {
                                        a = a + sp_wrap[k];   // forward sweep, j == len_b: flat index wraps to (i+1, 0)
                                } else if constexpr (MODE == MODE_EDGE) {
                                        if (e_last) {
                                                a = a + sp_wrap[k];
                                        }
                                }
                        } else if constexpr (BONUS == BONUS_DENSE) {
                                if (J.bonus && !(MODE == MODE_EDGE && e_first)) {
                                        // dense matrix supplied by the caller (kb200_pair_align_batch)
                                        a = a + __ldg(J.bonus + (size_t)rc.irow[k] * (size_t)J.len_b + (size_t)cc.jcol);
                                }
                        }
                        if constexpr (MODE == MODE_MID) {
                                ga = kmax(oGA + cc.CE, oA + cc.CO);
                                gb = kmax(u.gb + RE, u.a + RO);
                        } else if constexpr (MODE == MODE_EDGE) {
                                const float gm = kmax(u.gb + RE, u.a + RO);
                                const float gt = kmax(u.gb, u.a) + RT;
                                ga = e_any ? KB_NEGF : kmax(oGA + cc.CE, oA + cc.CO);
                                gb = e_term ? gt : gm;
                                a = e_first ? KB_NEGF : a;
                        } else {
                                ga = KB_NEGF;
                                gb = last_term ? (kmax(u.gb, u.a) + RT) : kmax(u.gb + RE, u.a + RO);
                        }
                }
                if constexpr (TAIL) {
                        if (!((vmask >> k) & 1u)) {
                                a = u.a; ga = u.ga; gb = u.gb;
                        }
                }
                sA[k] = a; sGA[k] = ga; sGB[k] = gb;
                d.a = oA; d.ga = oGA; d.gb = oGB;
                u.a = a; u.ga = ga; u.gb = gb;
        }
        if constexpr (BONUS == BONUS_SPARSE && MODE != MODE_FIRST) {
                // step the lists that were hit to their next entry: ONE divergent region per step
                if (hits) {
#pragma unroll
                        for (int k = 0; k < K; k++) {
                                if ((hits >> k) & 1u) {
                                        sp_i[k] += bdir;
                                        const int e = sp_i[k];
                                        const bool ok = (e >= 0) && (e < J.nb);
                                        int nc;
                                        float nv;
                                        bonus_entry<K, BSM>(J, s_bon, rc.irow[k], k, ok ? e : 0, nc, nv);
                                        sp_c[k] = ok ? nc : ((bdir > 0) ? 0x7fffffff : -1);
                                        sp_v[k] = ok ? nv : 0.0f;
                                }
                        }
                }
        }
}

// rows first .. first+K-1 of a sweep (logical numbering: 0 is the row next to the init row); rows
// past the end repeat the last one and are masked out (pass-through) by the returned bit mask
template <int V, int K>
__device__ __forceinline__ unsigned load_rows(const KbJob& J, const int bwd, const int r0, const int r1, const int first,
                                               const int tstride, RowCtx<V, K>& rc)
{
        constexpr int NA = VTraits<V>::NA;
        const int R = r1 - r0;
        unsigned vmask = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
                int g = first + k;
                const bool valid = g < R;
                if (valid) {
                        vmask |= (1u << k);
                }
                if (!valid) {
                        g = (R > 0) ? (R - 1) : 0;
                }
                int i = bwd ? (r1 - 1 - g) : (r0 + g);
                if (R == 0) {
                        i = r0;      // pass-through rows: values never survive
                        if (i >= J.len_a) i = J.len_a - 1;
                        if (i < 0) i = 0;
                }
                rc.irow[k] = i;
                if constexpr (is_ss<V>()) {
                        rc.rbase[k] = (int)J.seq_r[i] * tstride;
                } else {
                        const float* p = J.prof_r + ((size_t)(i + 1) << 6);
                        const float* pp = bwd ? (p + 64) : (p - 64);
                        rc.prow[k] = p;
                        rc.RO[k] = __ldg(p + 27);
                        rc.RE[k] = __ldg(p + 28);
                        rc.RT[k] = __ldg(p + 29);
                        rc.ROp[k] = __ldg(pp + 27);
                        if constexpr (NA > 0) {
#pragma unroll
                                for (int c = 0; c < NA; c++) {
                                        rc.cnt[k][c] = __ldg(p + c);
                                }
                        }
                }
        }
        return vmask;
}

// sparse consistency bonus: per row the index / column / value of the next entry in sweep
// direction, and the value the forward sweep picks up at j == len_b (flat index (i+1, 0))
template <int V, int K, bool BSM>
__device__ __forceinline__ void sparse_init(const KbJob& J, const int bwd, const int sb, const int eb, const RowCtx<V, K>& rc,
                                            int2* __restrict__ s_bon,
                                            int (&sp_i)[K], int (&sp_c)[K], float (&sp_v)[K], float (&sp_wrap)[K])
{
        const int KS = J.nb;
        if constexpr (BSM) {
                // stage the lists of this lane's rows (lane-private slots; s_bon is already offset by the lane)
#pragma unroll
                for (int k = 0; k < K; k++) {
                        const int* __restrict__ bc = J.bkey + (size_t)rc.irow[k] * (size_t)KS;
                        const float* __restrict__ bv = J.bval + (size_t)rc.irow[k] * (size_t)KS;
                        for (int e = 0; e < KS; e++) {
                                s_bon[(e * K + k) * 32] = make_int2(__ldg(bc + e), __float_as_int(__ldg(bv + e)));
                        }
                }
                __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
                int e;
                int c = 0;
                float v = 0.0f;
                if (!bwd) {
                        // cells visit j = sb+1 .. eb ascending
                        e = 0;
                        while (e < KS) {
                                bonus_entry<K, BSM>(J, s_bon, rc.irow[k], k, e, c, v);
                                if (c >= sb + 1) break;
                                e++;
                        }
                        if (e < KS) { sp_c[k] = c; sp_v[k] = v; }
                        if (eb == J.len_b && rc.irow[k] + 1 < J.len_a) {
                                // flat index (i, len_b) of the reference's dense matrix is (i+1, 0)
                                const size_t o = (size_t)(rc.irow[k] + 1) * (size_t)KS;
                                if (__ldg(J.bkey + o) == 0) sp_wrap[k] = __ldg(J.bval + o);
                        }
                } else {
                        // cells visit j = eb-1 .. sb descending
                        e = KS - 1;
                        while (e >= 0) {
                                bonus_entry<K, BSM>(J, s_bon, rc.irow[k], k, e, c, v);
                                if (c <= eb - 1) break;
                                e--;
                        }
                        if (e >= 0) { sp_c[k] = c; sp_v[k] = v; }
                }
                sp_i[k] = e;
        }
}

// One strip of 32*K rows starting at logical row `row0` of the sweep.
//   in_tag  : tag the row above must carry (0 for strip 0: the init row is generated)
//   out_tag : tag this strip stamps on the row it emits
// Hand-off protocol: the producer writes {a, ga, gb, tag} with one 16-byte store; the consumer
// re-reads the slot (ld.volatile.v4, L2) until the tag matches.  Tags are unique per launch and
// strip, so a slot still holding an older row (an earlier strip, an earlier round) never matches.
template <int V, int K, bool TAIL, int BONUS>
__device__ void sweep_strip(const KbJob& J, const int bwd, const int sb, const int eb,
                            const int r0, const int r1, const int row0,
                            const bool first_term, const bool last_term,
                            const Trip in, float4* __restrict__ rowbuf,
                            const unsigned in_tag, const unsigned out_tag,
                            const float* __restrict__ s_tbl, const int tstride, const int lane, float4* s_ring, float4* s_rec,
                            int2* s_bon)
{
        static_assert(BONUS != BONUS_SPARSE || K <= BON_KMAX_ROWS, "bonus staging area is sized for K <= 4");
        constexpr int NA = VTraits<V>::NA;
        constexpr int PW4 = (V == V_PP5) ? (PACK5 / 4) : (PACK23 / 4);
        const int C = eb - sb;
        const int R = r1 - r0;
        RowCtx<V, K> rc;
        const unsigned vmask = load_rows<V, K>(J, bwd, r0, r1, row0 + lane * K, tstride, rc);
        if constexpr (V == V_SP5 || V == V_SS5) {
                // stage the score vectors of this lane's rows: the per-cell lookup becomes a conflict-free
                // LDS (every lane reads its own bank) with an immediate row offset -- instead of an
                // uncoalesced gather that misses L1 (profile rows) or two address multiply-adds per
                // cell plus the offset add (sequence rows: (subm[r][c] - offset) is formed here, once)
                static_assert(K <= ((BONUS == BONUS_NONE) ? 8 : 4), "score-vector staging area");
                float* sp = reinterpret_cast<float*>(s_rec) + lane;
                __syncwarp();
#pragma unroll
                for (int k = 0; k < K; k++) {
#pragma unroll
                        for (int c = 0; c < 5; c++) {
                                if constexpr (V == V_SP5) {
                                        sp[(k * 5 + c) * 32] = __ldg(rc.prow[k] + 32 + c);
                                } else {
                                        sp[(k * 5 + c) * 32] = s_tbl[rc.rbase[k] + c] + J.nsoff;
                                }
                        }
                }
                rc.sprow = sp;
        }
        float sA[K], sGA[K], sGB[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
                sA[k] = KB_NEGF; sGA[k] = KB_NEGF; sGB[k] = KB_NEGF;
        }
        Trip d = {KB_NEGF, KB_NEGF, KB_NEGF};
        Trip bot = {KB_NEGF, KB_NEGF, KB_NEGF};
        float genA = in.a, genGA = in.ga;   // init-row generator (strip 0, lane 0)
        float prevCO = 0.0f;                // PP: [27] of the column visited one step earlier
        const bool gen = (in_tag == 0u);
        // Hand-off read side.  The row above is pulled in BLOCKS of HB columns, well ahead of its use:
        // column c is loaded by lane (c & 31) (ld.volatile.v4 into a register: a coalesced 128-byte
        // request per block), validated by its tag every HB steps (warp-uniform poll), and
        // committed to a 16-column shared-memory ring from which lane 0 takes one entry per step
        // (broadcast LDS).  Four blocks are in flight (24..32 steps of read-ahead), so that neither
        // the L2 latency nor the poll sits on the per-step path -- which is what bounds a warp that
        // runs alone on its scheduler (few big boxes: top of the guide tree, long sequences).
        constexpr int HB = 8;
        float4 blk = make_float4(0.f, 0.f, 0.f, 0.f);
        auto peek_above = [&](const int col) -> float4 {
                float4 v;
                const float4* p = rowbuf + col;
                asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                             : "l"(p)
                             : "memory");
                return v;
        };
        if (!gen && lane <= C) {
                blk = peek_above(lane);
        }
        // Column records of 5-letter profile-profile sweeps go through a 64-column shared-memory
        // ring as well (cp.async, 16 columns at a time, a block ahead): the per-step record read is
        // an LDS.128 pair instead of an L1-missing global load every fourth column.
        constexpr bool RECRING = (NA > 0 && NA <= 5);
        float4* const s_recA = s_rec;
        float4* const s_recB = s_rec + 64;
        auto rec_issue = [&](const int col0, const int ncols) {
                if constexpr (RECRING) {
                        const int col = col0 + (lane >> 1);
                        const int half = lane & 1;
                        if ((lane >> 1) < ncols && col <= C) {
                                const long long ridx = bwd ? (long long)(eb - col + 1) : (long long)(sb + col);
                                const float4* src = reinterpret_cast<const float4*>(J.cpack) + ridx * 2 + half;
                                const unsigned dst = (unsigned)__cvta_generic_to_shared((half ? s_recB : s_recA) + (col & 63));
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                        }
                }
        };
        auto rec_wait = [&]() {
                if constexpr (RECRING) {
                        asm volatile("cp.async.wait_all;" ::: "memory");
                        __syncwarp();
                }
        };
        // every HB steps (t % HB == 0): commit the hand-off block of columns t .. t+HB-1 and start
        // the load of the block 32 columns further; every 16 steps the same for the record ring
        auto boundary = [&](const int t) {
                if (!gen && t <= C) {
                        const bool mine = (lane >> 3) == ((t >> 3) & 3);
                        const int col = t + (lane & 7);
                        const bool need = mine && (col <= C);
                        while (true) {
                                const bool ok = !need || (__float_as_uint(blk.w) == in_tag);
                                if (__all_sync(FULL, ok)) {
                                        break;
                                }
                                if (!ok) {
                                        blk = peek_above(col);
                                }
                        }
                        if (need) {
                                s_ring[col & 15] = blk;
                        }
                        __syncwarp();
                        if (mine && col + 32 <= C) {
                                blk = peek_above(col + 32);
                        }
                }
                if constexpr (RECRING) {
                        if ((t & 15) == 0 && t > 0) {
                                rec_wait();                      // columns t+1 .. t+16 (issued 16 steps ago)
                                rec_issue(t + 17, 16);           // slots of columns t-47 .. t-32: last read at step t-2
                        }
                }
        };
        static_assert(HB == 8, "lane groups of 8");
        const int steps = C + 32;
        // column input of the current step (filled one step ahead); lane 0 starts on column 0 at t=0
        float4 curA = make_float4(0.f, 0.f, 0.f, 0.f), curB = curA;
        int cur_cres = 0;
        float cur_bon[K];
#pragma unroll
        for (int k = 0; k < K; k++) cur_bon[k] = 0.0f;
        // sparse consistency bonus: per row the index / column / value of the next entry in sweep
        // direction, and the value the forward sweep picks up at j == len_b (flat index (i+1, 0))
        const bool sparse = (BONUS == BONUS_SPARSE) && (J.bkey != nullptr);
        int2* const my_bon = s_bon + lane;       // lane-private slots [entry][row k][lane]
        const int bdir = bwd ? -1 : 1;
        int sp_i[K], sp_c[K];
        float sp_v[K], sp_wrap[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
                sp_i[k] = 0; sp_c[k] = bwd ? -1 : 0x7fffffff; sp_v[k] = 0.0f; sp_wrap[k] = 0.0f;
        }
        if constexpr (BONUS) {
                if (sparse) {
                        sparse_init<V, K, true>(J, bwd, sb, eb, rc, my_bon, sp_i, sp_c, sp_v, sp_wrap);
                }
        }
        if constexpr (RECRING) {
                static_assert(PACK5 == 8, "PP5 record is two float4");
                rec_issue(0, 16);
                rec_issue(16, 16);
                rec_issue(32, 1);
                rec_wait();
                // lane 0 is on column 0 at step 0; the other lanes read their column 0 one step ahead
                curA = s_recA[0];
                curB = s_recB[0];
        }
        // running pointers instead of per-step index arithmetic: the column visited NEXT (pu = u+1;
        // 1-lane at t=0) and the state column of the current u
        const int dstep = bwd ? -1 : 1;
        int jcur = bwd ? (eb + lane) : (sb - lane);
        const long long pr_first = bwd ? (long long)(eb - (1 - lane)) : (long long)(sb + (1 - lane)) - 1;
        const float4* recp = nullptr;          // packed record of the next column (PP)
        const uint8_t* seqp = nullptr;         // residue of the next column (SS, SP)
        if constexpr (NA > 5) {
                recp = reinterpret_cast<const float4*>(J.cpack) + (pr_first + 1) * PW4;
        } else if constexpr (NA == 0) {
                seqp = J.seq_c + pr_first;
        }
        // one step of the wavefront.  STEADY (32 <= t <= C-1): every lane is on an interior column
        // (1 <= u <= C-1), so the activity test and the boundary-column dispatch disappear.
        auto step = [&](auto steady_tag, const int t) {
                constexpr bool STEADY = decltype(steady_tag)::value;
                const int u = t - lane;
                // row above, column t, for lane 0 (consumer strips): committed by boundary()
                float4 hin = make_float4(0.f, 0.f, 0.f, 0.f);
                if (!gen) {
                        hin = s_ring[t & 15];
                }
                Trip up;
                up.a = __shfl_up_sync(FULL, bot.a, 1);
                up.ga = __shfl_up_sync(FULL, bot.ga, 1);
                up.gb = __shfl_up_sync(FULL, bot.gb, 1);
                const bool act = STEADY ? true : ((u >= 0) && (u <= C));
                // ---- column input of the NEXT step (software prefetch: the load latency is off the
                //      recurrence's critical path, which matters when few warps are resident) ----
                constexpr bool PREF = (NA <= 5);   // 23-letter records are loaded in-step (register budget)
                float4 nxtA = make_float4(0.f, 0.f, 0.f, 0.f), nxtB = nxtA;   // PP5 record = 2 x float4
                int ncres = 0;
                if constexpr (RECRING) {
                        // ring slot of column u+1 (lanes outside the box read a slot they never use)
                        nxtA = s_recA[(u + 1) & 63];
                        nxtB = s_recB[(u + 1) & 63];
                } else if constexpr (PREF) {
                        const int pu = u + 1;
                        if (STEADY || (pu >= 1 && pu <= C)) {
                                ncres = (int)__ldg(seqp);
                        }
                }
                if (act) {
                        // ---- column context (prefetched during the previous step) ----
                        ColCtx<V> cc;
                        const int j = jcur;                             // state column: eb-u / sb+u
                        cc.jcol = j;
                        cc.cres = cur_cres;
                        float CT;
                        if constexpr (NA > 0 && !PREF) {
                                const float4* __restrict__ rec = recp - dstep * PW4;   // record of the current column
                                float buf[PW4 * 4];
#pragma unroll
                                for (int w = 0; w < PW4; w++) {
                                        const float4 v = __ldg(rec + w);
                                        buf[4 * w] = v.x; buf[4 * w + 1] = v.y; buf[4 * w + 2] = v.z; buf[4 * w + 3] = v.w;
                                }
#pragma unroll
                                for (int c = 0; c < NA; c++) {
                                        cc.qs[c] = buf[c];
                                }
                                cc.CO = buf[NA];
                                cc.CE = buf[NA + 1];
                                CT = buf[NA + 2];
                                cc.COp = prevCO;
                                prevCO = cc.CO;
                        } else if constexpr (NA > 0) {
                                // PACK5: s0..s3 | s4, [27], [28], [29]
                                cc.qs[0] = curA.x; cc.qs[1] = curA.y; cc.qs[2] = curA.z; cc.qs[3] = curA.w;
                                cc.qs[4] = curB.x;
                                cc.CO = curB.y;
                                cc.CE = curB.z;
                                CT = curB.w;
                                cc.COp = prevCO;
                                prevCO = cc.CO;
                        } else {
                                cc.CO = J.o; cc.CE = J.e; CT = J.t; cc.COp = J.o;
                        }
                        // ---- lane 0: take the row above from the source ----
                        if constexpr (STEADY) {
                                // branch-free: every lane evaluates the init-row generator (three
                                // operations), lane 0 keeps the result -- no divergent region per step
                                const float nga = first_term ? (kmax(genGA, genA) + CT) : kmax(genGA + cc.CE, genA + cc.CO);
                                const bool l0 = (lane == 0);
                                const float sa = gen ? KB_NEGF : hin.x;
                                const float sga = gen ? nga : hin.y;
                                const float sgb = gen ? KB_NEGF : hin.z;
                                up.a = l0 ? sa : up.a;
                                up.ga = l0 ? sga : up.ga;
                                up.gb = l0 ? sgb : up.gb;
                                genA = KB_NEGF;
                                genGA = nga;
                        } else if (lane == 0) {
                                if (gen) {
                                        if (!STEADY && u == 0) {
                                                up = in;
                                        } else if (STEADY || u < C) {
                                                float nga;
                                                if (first_term) {
                                                        nga = kmax(genGA, genA) + CT;
                                                } else {
                                                        nga = kmax(genGA + cc.CE, genA + cc.CO);
                                                }
                                                up.a = KB_NEGF; up.ga = nga; up.gb = KB_NEGF;
                                                genA = KB_NEGF; genGA = nga;
                                        } else {
                                                up.a = KB_NEGF; up.ga = KB_NEGF; up.gb = KB_NEGF;
                                        }
                                } else {
                                        up.a = hin.x; up.ga = hin.y; up.gb = hin.z;
                                }
                        }
                        const Trip got = up;
                        if constexpr (STEADY) {
                                cells<V, K, TAIL, MODE_MID, BONUS, true>(J, rc, vmask, first_term, last_term, cc, cur_bon, sparse, bdir, sp_i, sp_c, sp_v, sp_wrap, my_bon, s_tbl, sA, sGA, sGB, d, up);
                        } else {
                                cells<V, K, TAIL, MODE_EDGE, BONUS, true>(J, rc, vmask, first_term, last_term, cc, cur_bon, sparse, bdir, sp_i, sp_c, sp_v, sp_wrap, my_bon, s_tbl, sA, sGA, sGB, d, up, u == 0, u == C);
                        }
                        d = got;
                        bot = up;
                }
                cur_cres = ncres;
                if constexpr (NA > 0 && PREF) {
                        curA = nxtA;
                        curB = nxtB;
                }
                jcur += dstep;
                if constexpr (NA > 5) {
                        recp += dstep * PW4;
                } else if constexpr (NA == 0) {
                        seqp += dstep;
                }
                if (lane == 31) {
                        const int uo = t - 31;
                        if (STEADY || (uo >= 0 && uo <= C)) {
                                rowbuf[uo] = make_float4(bot.a, bot.ga, bot.gb, __uint_as_float(out_tag));
                        }
                }
        };
        {
                int t = 0;
                const int t_fill = (steps < 32) ? steps : 32;
                for (; t < t_fill; t++) {
                        if ((t & (HB - 1)) == 0) {
                                boundary(t);
                        }
                        step(std::false_type{}, t);
                }
                while (t + HB <= C) {           // 32 <= t <= C-1, in runs of HB steps
                        boundary(t);
                        // two steps per iteration: the register rotation of the software pipeline
                        // (current <- next column, diagonal <- row above) becomes renaming
                        // (not in the sparse-bonus family: its list cursors already fill the register file)
                        if constexpr (BONUS == BONUS_SPARSE) {
#pragma unroll 1
                                for (int q = 0; q < HB; q++) {
                                        step(std::true_type{}, t);
                                        t++;
                                }
                        } else {
#pragma unroll 1
                                for (int q = 0; q < HB; q += 2) {
                                        step(std::true_type{}, t);
                                        step(std::true_type{}, t + 1);
                                        t += 2;
                                }
                        }
                }
                if (t < C) {
                        boundary(t);
                        for (; t < C; t++) {
                                step(std::true_type{}, t);
                        }
                }
                for (; t < steps; t++) {
                        if ((t & (HB - 1)) == 0) {
                                boundary(t);
                        }
                        step(std::false_type{}, t);
                }
        }
        rec_wait();      // no copy into the record ring may outlive the strip (the area is reused)
        __syncwarp();
}

template <int V, int BONUS>
__device__ void sweep_unit(const KbJob& J, const KbBox& bx, const int bwd, const int strip, const int thin,
                           const unsigned tag_base, const float* __restrict__ s_tbl, const int tstride, const int lane,
                           float4* s_ring, float4* s_rec, int2* s_bon)
{
        const int mid = (bx.ea - bx.sa) / 2 + bx.sa;
        const int r0 = bwd ? mid : bx.sa;
        const int r1 = bwd ? bx.ea : mid;
        const int R = r1 - r0;
        const int sb = bx.sb, eb = bx.eb;
        const bool first_term = bwd ? (eb == J.len_b) : (sb == 0);
        const bool last_term = bwd ? (sb == 0) : (eb == J.len_b);
        float4* rowbuf = (bwd ? J.rowB : J.rowF) + (bx.sa + bx.sb);
        Trip in;
        if (bwd) {
                in.a = bx.b0a; in.ga = bx.b0ga; in.gb = bx.b0gb;
        } else {
                in.a = bx.f0a; in.ga = bx.f0ga; in.gb = bx.f0gb;
        }
        const int rps = rows_per_strip(J.kind, J.nalpha, thin, BONUS != BONUS_NONE);
        const int nstr = (R + rps - 1) / rps > 0 ? (R + rps - 1) / rps : 1;
        const int row0 = strip * rps;
        (void)nstr;
        const unsigned prev = (strip > 0) ? (tag_base + (unsigned)strip) : 0u;      // tag written by strip-1
        const unsigned mine = tag_base + (unsigned)strip + 1u;
        const int rem = R - row0;
        // rows per lane of this strip: the full width of the kind's strips, or -- last strip of a
        // sweep, boxes of the deeper rounds -- the smallest K that covers the remaining rows, so that
        // a 187-row half box runs with 6 rows per lane (97 % of the lanes' rows live) instead of 8
#define KB_STRIP(KK, TT) sweep_strip<V, KK, TT, BONUS>(J, bwd, sb, eb, r0, r1, row0, first_term, last_term, in, rowbuf, prev, mine, s_tbl, tstride, lane, s_ring, s_rec, s_bon)
        const int kneed = (rem + 31) >> 5;
        if (rps == 32 || kneed <= 1) {
                KB_STRIP(1, true);
        } else if constexpr (V == V_PP23) {
                if (rem >= 64) KB_STRIP(2, false);
                else KB_STRIP(2, true);
        } else if constexpr (is_ss<V>() && !BONUS) {
                if (rem >= 256) KB_STRIP(8, false);
                else if (kneed > 6) KB_STRIP(8, true);
                else if (kneed > 4) KB_STRIP(6, true);
                else if (kneed == 4) KB_STRIP(4, true);
                else if (kneed == 3) KB_STRIP(3, true);
                else KB_STRIP(2, true);
        } else {
                if (rem >= 128) KB_STRIP(4, false);
                else if (kneed == 4) KB_STRIP(4, true);
                else if (kneed == 3) KB_STRIP(3, true);
                else KB_STRIP(2, true);
        }
#undef KB_STRIP
}

// BONUS is a kernel-level template parameter: a batch either carries consistency bonuses for all
// of its jobs (tree levels in default mode) or for none (anchor batch, --fast), and the two
// families get independent register allocation / code size.
template <int BONUS>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 4)
kb_sweep_kernel(const KbJob* __restrict__ jobs, const KbBox* __restrict__ boxes,
                const KbUnit* __restrict__ units, const unsigned* __restrict__ nunits_p,
                unsigned int* __restrict__ cursor, const unsigned tag_base,
                const float* __restrict__ tbl, const int thin, const int tstride)
{
        __shared__ float s_tbl[TBL_MAX];
        __shared__ float4 s_ring_all[WARPS_PER_CTA][16];     // hand-off ring (row above), one per warp
        // per warp: 5-letter profile-profile column records (two 64-column rings, 2 KB) or the staged
        // score vectors of a 5-letter sequence / profile-sequence strip ([K rows][5 letters][32 lanes] floats)
        constexpr int REC_F4 = (BONUS == BONUS_NONE) ? 320 : 160;     // K = 8 rows per lane without a bonus
        __shared__ float4 s_rec_all[WARPS_PER_CTA][REC_F4];
        for (int i = threadIdx.x; i < TBL_MAX; i += blockDim.x) {
                s_tbl[i] = tbl[i];
        }
        __syncthreads();
        const int lane = threadIdx.x & 31;
        // sparse bonus lists of the rows of the strip a warp is sweeping (bonus kernel family only)
        __shared__ int2 s_bon_all[(BONUS == BONUS_SPARSE) ? WARPS_PER_CTA * BON_SLOTS * BON_KMAX_ROWS * 32 : 1];
        float4* s_ring = s_ring_all[threadIdx.x >> 5];
        float4* s_rec = s_rec_all[threadIdx.x >> 5];
        int2* s_bon = (BONUS == BONUS_SPARSE) ? (s_bon_all + (threadIdx.x >> 5) * (BON_SLOTS * BON_KMAX_ROWS * 32)) : s_bon_all;
        const unsigned total = *nunits_p;
        while (true) {
                unsigned unit = 0;
                if (lane == 0) {
                        unit = atomicAdd(cursor, 1u);
                }
                unit = __shfl_sync(FULL, unit, 0);
                if (unit >= total) {
                        break;
                }
                const KbUnit un = units[unit];
                const KbBox bx = boxes[un.item >> 1];
                const int bwd = un.item & 1;
                const KbJob J = jobs[bx.job];
                const unsigned ps = tag_base;
                if (J.kind == KB200_KIND_SS) {
                        if (tstride == 5) {
                                sweep_unit<V_SS5, BONUS>(J, bx, bwd, un.strip, thin, ps, s_tbl, tstride, lane, s_ring, s_rec, s_bon);
                        } else {
                                sweep_unit<V_SS, BONUS>(J, bx, bwd, un.strip, thin, ps, s_tbl, tstride, lane, s_ring, s_rec, s_bon);
                        }
                } else if (J.kind == KB200_KIND_SP) {
                        if (J.nalpha <= 5) {
                                sweep_unit<V_SP5, BONUS>(J, bx, bwd, un.strip, thin, ps, s_tbl, tstride, lane, s_ring, s_rec, s_bon);
                        } else {
                                sweep_unit<V_SP, BONUS>(J, bx, bwd, un.strip, thin, ps, s_tbl, tstride, lane, s_ring, s_rec, s_bon);
                        }
                } else if (J.nalpha <= 5) {
                        sweep_unit<V_PP5, BONUS>(J, bx, bwd, un.strip, thin, ps, s_tbl, tstride, lane, s_ring, s_rec, s_bon);
                } else {
                        sweep_unit<V_PP23, BONUS>(J, bx, bwd, un.strip, thin, ps, s_tbl, tstride, lane, s_ring, s_rec, s_bon);
                }
        }
}

// ---------------------------------------------------------------------------------------------
// plan: cut every (box, direction) into strips, reserve a contiguous, ordered unit range
__global__ void kb_plan_kernel(const KbJob* __restrict__ jobs, const KbBox* __restrict__ boxes, const int nboxes,
                               const int thin, const int batch_bonus, KbUnit* __restrict__ units,
                               unsigned* __restrict__ nunits)
{
        const int lane = threadIdx.x & 31;
        const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        int nstr = 0;
        if (item < 2LL * nboxes) {
                const KbBox bx = boxes[item >> 1];
                const int bwd = (int)(item & 1);
                const int mid = (bx.ea - bx.sa) / 2 + bx.sa;
                const int R = bwd ? (bx.ea - mid) : (mid - bx.sa);
                const int rps = rows_per_strip(jobs[bx.job].kind, jobs[bx.job].nalpha, thin, batch_bonus != 0);
                nstr = (R + rps - 1) / rps;
                if (nstr < 1) nstr = 1;
        }
        // warp-inclusive scan of nstr
        int incl = nstr;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
        }
        const int warp_total = __shfl_sync(FULL, incl, 31);
        unsigned base = 0;
        if (lane == 31 && warp_total > 0) {
                base = atomicAdd(nunits, (unsigned)warp_total);
        }
        base = __shfl_sync(FULL, base, 31);
        const unsigned mine = base + (unsigned)(incl - nstr);
        for (int s = 0; s < nstr; s++) {
                KbUnit un;
                un.item = (int)item;
                un.strip = s;
                units[mine + s] = un;
        }
}

// PP jobs: pack the column profile into compact records (scores of the alphabet, [27],[28],[29])
__global__ void kb_pack_kernel(const KbJob* __restrict__ jobs, const int* __restrict__ pp_jobs, const int npp,
                               const long long* __restrict__ col_prefix, const long long total_cols)
{
        const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        const long long nth = (long long)gridDim.x * blockDim.x;
        for (long long gc = gid; gc < total_cols; gc += nth) {
                int lo = 0, hi = npp - 1;
                while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (col_prefix[mid] <= gc) lo = mid; else hi = mid - 1;
                }
                const KbJob& J = jobs[pp_jobs[lo]];
                const int col = (int)(gc - col_prefix[lo]);
                const float* q = J.prof_c + ((size_t)col << 6);
                const int na = (J.nalpha <= 5) ? 5 : 23;
                const int pw = (J.nalpha <= 5) ? PACK5 : PACK23;
                float* out = const_cast<float*>(J.cpack) + (size_t)col * pw;
                for (int c = 0; c < na; c++) {
                        out[c] = q[32 + c];
                }
                out[na] = q[27];
                out[na + 1] = q[28];
                out[na + 2] = q[29];
                for (int c = na + 3; c < pw; c++) {
                        out[c] = 0.0f;
                }
        }
}

// ---------------------------------------------------------------------------------------------
// meet-up + aln_continue

struct Best {
        float max, max2;
        int key;
};

__device__ __forceinline__ void offer(Best& m, const float s, const int key)
{
        if (s > m.max) {
                m.max2 = m.max;
                m.max = s;
                m.key = key;
        } else if (s > m.max2) {
                m.max2 = s;
        }
}

__device__ __forceinline__ void put_child(KbBox* __restrict__ next, int slot, int job, int depth,
                                          int sa, int ea, int sb, int eb, Trip f0, Trip b0)
{
        KbBox c;
        c.job = job; c.sa = sa; c.ea = ea; c.sb = sb; c.eb = eb;
        c.f0a = f0.a; c.f0ga = f0.ga; c.f0gb = f0.gb;
        c.b0a = b0.a; c.b0ga = b0.ga; c.b0gb = b0.gb;
        c.depth = depth;
        next[slot] = c;
}

__global__ void __launch_bounds__(128)
kb_meetup_kernel(const KbJob* __restrict__ jobs, const KbBox* __restrict__ boxes, const int nboxes,
                 KbBox* __restrict__ next, unsigned int* __restrict__ next_count,
                 KbBox* __restrict__ small, unsigned int* __restrict__ small_count, const int small_rows, const int small_cols,
                 unsigned long long* __restrict__ cells)
{
        const int lane = threadIdx.x & 31;
        const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const int nwarps = (gridDim.x * blockDim.x) >> 5;
        for (int b = warp; b < nboxes; b += nwarps) {
                const KbBox bx = boxes[b];
                const KbJob J = jobs[bx.job];
                const int sa = bx.sa, ea = bx.ea, sb = bx.sb, eb = bx.eb;
                const int mid = (ea - sa) / 2 + sa;
                const float4* __restrict__ F = J.rowF + (sa + sb);
                const float4* __restrict__ B = J.rowB + (sa + sb);
                const float middle = (float)(eb - sb) / 2.0F + (float)sb;
                float x2, x3, x5, x6, x6last, x7;
                if (J.kind == KB200_KIND_SS) {
                        x2 = x3 = x5 = x7 = J.o;
                        x6 = (sb == 0) ? J.t : J.e;
                        x6last = (eb == J.len_b) ? J.t : J.e;
                } else {
                        const float* P = J.prof_r + ((size_t)(mid + 1) << 6);
                        x3 = P[27];
                        x7 = P[-37];
                        x6 = (sb == 0) ? P[29] : P[28];
                        x6last = (eb == J.len_b) ? P[29] : P[28];
                        x2 = x5 = J.o;
                }
                Best m;
                m.max = KB_NEGF; m.max2 = KB_NEGF; m.key = 0x7fffffff;
                for (int i = sb + lane; i <= eb; i += 32) {
                        const float4 f = __ldcg(F + (i - sb));
                        const float4 bb = __ldcg(B + (eb - i));
                        float sub = fabsf(middle - (float)i);
                        sub = __fdiv_rn(sub, 1000.0F);
                        const int kb = (i - sb) * 8;
                        if (i < eb) {
                                if (J.kind == KB200_KIND_PP) {
                                        x2 = J.prof_c[((size_t)(i + 1) << 6) + 27];
                                        x5 = J.prof_c[((size_t)i << 6) + 27];
                                }
                                offer(m, f.x + bb.x - sub, kb + 1);
                                offer(m, f.x + bb.y + x2 - sub, kb + 2);
                                offer(m, f.x + bb.z + x3 - sub, kb + 3);
                                offer(m, f.y + bb.x + x5 - sub, kb + 5);
                                offer(m, f.z + bb.z + x6 - sub, kb + 6);
                                offer(m, f.z + bb.x + x7 - sub, kb + 7);
                        } else {
                                offer(m, f.x + bb.z + x3 - sub, kb + 3);
                                offer(m, f.z + bb.z + x6last - sub, kb + 6);
                        }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                        const float omax = __shfl_xor_sync(FULL, m.max, off);
                        const float omax2 = __shfl_xor_sync(FULL, m.max2, off);
                        const int okey = __shfl_xor_sync(FULL, m.key, off);
                        const bool other_wins = (omax > m.max) || (omax == m.max && okey < m.key);
                        const float lose = other_wins ? m.max : omax;
                        const float w2 = kmax(m.max2, omax2);
                        if (other_wins) {
                                m.max = omax;
                                m.key = okey;
                        }
                        m.max2 = kmax(lose, w2);
                }
                if (lane == 0) {
                        const unsigned long long nc = (unsigned long long)(ea - sa) * (unsigned long long)(eb - sb);
                        atomicAdd(cells + J.kind, nc);
                        if (J.bonus || J.bkey) {
                                atomicAdd(cells + 3, nc);
                        }
                        if (bx.depth == 0 && J.score) {
                                *J.score = m.max;
                        }
                        if (m.key != 0x7fffffff && J.path) {
                                const int c = sb + (m.key >> 3);
                                const int t = m.key & 7;
                                int* __restrict__ path = J.path;
                                const Trip fin = {bx.f0a, bx.f0ga, bx.f0gb};
                                const Trip bin = {bx.b0a, bx.b0ga, bx.b0gb};
                                const Trip KA = {0.0F, KB_NEGF, KB_NEGF};
                                const Trip KGA = {KB_NEGF, 0.0F, KB_NEGF};
                                const Trip KGB = {KB_NEGF, KB_NEGF, 0.0F};
                                // children of aln_continue (aln_controller.c:198-431)
                                int lsa = sa, lea, lsb = sb, leb, rsa, rea = ea, rsb, reb = eb;
                                Trip lb0, rf0;
                                switch (t) {
                                case 1:
                                        path[mid] = c; path[mid + 1] = c + 1;
                                        lea = mid - 1; leb = c - 1; lb0 = KA;
                                        rsa = mid + 1; rsb = c + 1; rf0 = KA;
                                        break;
                                case 2:
                                        path[mid] = c;
                                        lea = mid - 1; leb = c - 1; lb0 = KA;
                                        rsa = mid; rsb = c + 1; rf0 = KGA;
                                        break;
                                case 3:
                                        path[mid] = c;
                                        lea = mid - 1; leb = c - 1; lb0 = KA;
                                        rsa = mid + 1; rsb = c; rf0 = KGB;
                                        break;
                                case 5:
                                        path[mid + 1] = c + 1;
                                        lea = mid; leb = c - 1; lb0 = KGA;
                                        rsa = mid + 1; rsb = c + 1; rf0 = KA;
                                        break;
                                case 6:
                                        lea = mid - 1; leb = c; lb0 = KGB;
                                        rsa = mid + 1; rsb = c; rf0 = KGB;
                                        break;
                                default: /* 7 */
                                        path[mid + 1] = c + 1;
                                        lea = mid - 1; leb = c; lb0 = KGB;
                                        rsa = mid + 1; rsb = c + 1; rf0 = KA;
                                        break;
                                }
                                const bool hasL = (lsa < lea) && (lsb < leb);
                                const bool hasR = (rsa < rea) && (rsb < reb);
                                const bool smL = hasL && small && ((lea - lsa) <= small_rows) && ((leb - lsb) <= small_cols);
                                const bool smR = hasR && small && ((rea - rsa) <= small_rows) && ((reb - rsb) <= small_cols);
                                const int nbig = ((hasL && !smL) ? 1 : 0) + ((hasR && !smR) ? 1 : 0);
                                const int nsml = (smL ? 1 : 0) + (smR ? 1 : 0);
                                if (nbig) {
                                        unsigned slot = atomicAdd(next_count, (unsigned)nbig);
                                        if (hasL && !smL) {
                                                put_child(next, (int)slot, bx.job, bx.depth + 1, lsa, lea, lsb, leb, fin, lb0);
                                                slot++;
                                        }
                                        if (hasR && !smR) {
                                                put_child(next, (int)slot, bx.job, bx.depth + 1, rsa, rea, rsb, reb, rf0, bin);
                                        }
                                }
                                if (nsml) {
                                        unsigned slot = atomicAdd(small_count, (unsigned)nsml);
                                        if (smL) {
                                                put_child(small, (int)slot, bx.job, bx.depth + 1, lsa, lea, lsb, leb, fin, lb0);
                                                slot++;
                                        }
                                        if (smR) {
                                                put_child(small, (int)slot, bx.job, bx.depth + 1, rsa, rea, rsb, reb, rf0, bin);
                                        }
                                }
                        }
                }
        }
}


// ---------------------------------------------------------------------------------------------
// small boxes: one THREAD finishes the whole remaining Hirschberg recursion of a box with few rows
// and columns (serial sweeps exactly as the reference runs them, explicit DFS stack), so the deep
// recursion levels -- millions of boxes of a handful of cells -- cost one launch instead of one
// round each.  Arithmetic is the same cell recurrence in the same order.
constexpr int SMALL_ROWS = 16;        // default thresholds (few jobs: deeper sweeps stay parallel)
constexpr int SMALL_COLS = 48;
constexpr int SMALL_ROWS_MAX = 128;   // many jobs: the machine is full anyway, skip the lane-starved rounds
constexpr int SMALL_COLS_MAX = 256;
constexpr int SMALL_STACK = 12;

struct SBox {
        int sa, ea, sb, eb;
        Trip f0, b0;
};

// One sweep of a small box by ONE thread: the rows are taken K at a time through the shared cell
// routine (cells<>, the same code the warp strips run: K cells of a column chained in registers),
// so the row array S -- thread-local memory -- is read and written once per K rows.
template <int V, int K, int BONUS>
__device__ void small_sweep(const KbJob& J, const int bwd, const int r0, const int r1, const int sb, const int eb,
                            const Trip in, float4* __restrict__ S, const float* __restrict__ s_tbl, const int tstride)
{
        constexpr int NA = VTraits<V>::NA;
        constexpr int PW = (V == V_PP5) ? PACK5 : PACK23;
        const int C = eb - sb;
        const int R = r1 - r0;
        const bool first_term = bwd ? (eb == J.len_b) : (sb == 0);
        const bool last_term = bwd ? (sb == 0) : (eb == J.len_b);
        const int dstep = bwd ? -1 : 1;
        // record / residue of state column u: r = bwd ? eb-u : sb+u-1
        const long long rfirst = bwd ? (long long)eb : (long long)sb - 1;
        {
                // init row
                float pa = in.a, pga = in.ga;
                S[0] = make_float4(in.a, in.ga, in.gb, 0.0f);
                for (int u = 1; u < C; u++) {
                        float CO, CE, CT;
                        if constexpr (NA > 0) {
                                const float* rec = J.cpack + (size_t)(rfirst + (long long)dstep * u + 1) * PW;
                                CO = __ldg(rec + NA); CE = __ldg(rec + NA + 1); CT = __ldg(rec + NA + 2);
                        } else {
                                CO = J.o; CE = J.e; CT = J.t;
                        }
                        const float ga = first_term ? (kmax(pga, pa) + CT) : kmax(pga + CE, pa + CO);
                        S[u] = make_float4(KB_NEGF, ga, KB_NEGF, 0.0f);
                        pa = KB_NEGF;
                        pga = ga;
                }
                S[C] = make_float4(KB_NEGF, KB_NEGF, KB_NEGF, 0.0f);
        }
        const bool sparse = (BONUS == BONUS_SPARSE) && (J.bkey != nullptr);
        for (int v0 = 0; v0 < R; v0 += K) {
                RowCtx<V, K> rc;
                const unsigned vmask = load_rows<V, K>(J, bwd, r0, r1, v0, tstride, rc);
                float sA[K], sGA[K], sGB[K], bon[K];
                int sp_i[K], sp_c[K];
                float sp_v[K], sp_wrap[K];
#pragma unroll
                for (int k = 0; k < K; k++) {
                        sA[k] = KB_NEGF; sGA[k] = KB_NEGF; sGB[k] = KB_NEGF; bon[k] = 0.0f;
                        sp_i[k] = 0; sp_c[k] = bwd ? -1 : 0x7fffffff; sp_v[k] = 0.0f; sp_wrap[k] = 0.0f;
                }
                if constexpr (BONUS) {
                        if (sparse) {
                                sparse_init<V, K, false>(J, bwd, sb, eb, rc, nullptr, sp_i, sp_c, sp_v, sp_wrap);
                        }
                }
                Trip d = {KB_NEGF, KB_NEGF, KB_NEGF};
                ColCtx<V> cc;
                cc.cres = 0;
                cc.CO = J.o; cc.CE = J.e; cc.COp = J.o;
                float prevCO = 0.0f;
                auto column = [&](const int u) {
                        cc.jcol = bwd ? (eb - u) : (sb + u);
                        const long long r = rfirst + (long long)dstep * u;
                        if constexpr (NA > 0) {
                                const float* __restrict__ rec = J.cpack + (size_t)(r + 1) * PW;
                                if constexpr (PW % 4 == 0) {
                                        float buf[PW];
#pragma unroll
                                        for (int w = 0; w < PW / 4; w++) {
                                                const float4 q = __ldg(reinterpret_cast<const float4*>(rec) + w);
                                                buf[4 * w] = q.x; buf[4 * w + 1] = q.y; buf[4 * w + 2] = q.z; buf[4 * w + 3] = q.w;
                                        }
#pragma unroll
                                        for (int c = 0; c < NA; c++) cc.qs[c] = buf[c];
                                        cc.CO = buf[NA]; cc.CE = buf[NA + 1];
                                }
                                cc.COp = prevCO;
                                prevCO = cc.CO;
                        } else {
                                if (u >= 1) {
                                        cc.cres = (int)__ldg(J.seq_c + r);
                                }
                        }
                };
                // column 0
                {
                        column(0);
                        const float4 q = S[0];
                        Trip up = {q.x, q.y, q.z};
                        const Trip got = up;
                        cells<V, K, true, MODE_FIRST, BONUS, false>(J, rc, vmask, first_term, last_term, cc, bon, sparse, dstep, sp_i, sp_c, sp_v, sp_wrap, nullptr, s_tbl, sA, sGA, sGB, d, up);
                        d = got;
                        S[0] = make_float4(up.a, up.ga, up.gb, 0.0f);
                }
                for (int u = 1; u < C; u++) {
                        column(u);
                        const float4 q = S[u];
                        Trip up = {q.x, q.y, q.z};
                        const Trip got = up;
                        cells<V, K, true, MODE_MID, BONUS, false>(J, rc, vmask, first_term, last_term, cc, bon, sparse, dstep, sp_i, sp_c, sp_v, sp_wrap, nullptr, s_tbl, sA, sGA, sGB, d, up);
                        d = got;
                        S[u] = make_float4(up.a, up.ga, up.gb, 0.0f);
                }
                {
                        column(C);
                        const float4 q = S[C];
                        Trip up = {q.x, q.y, q.z};
                        cells<V, K, true, MODE_LAST, BONUS, false>(J, rc, vmask, first_term, last_term, cc, bon, sparse, dstep, sp_i, sp_c, sp_v, sp_wrap, nullptr, s_tbl, sA, sGA, sGB, d, up);
                        S[C] = make_float4(up.a, up.ga, up.gb, 0.0f);
                }
        }
}

template <int V, int MAXC, int BONUS>
__device__ void small_box_run(const KbJob& J, const KbBox& root, const float* __restrict__ s_tbl, const int tstride,
                              unsigned long long& ncells)
{
        // rows per pass (register budget as in the strips: 8 for plain seq-seq, 4, 2 for 23-letter profiles)
        constexpr int KR = (V == V_PP23) ? 2 : ((V == V_SS && BONUS == BONUS_NONE) ? 8 : 4);
        float4 F[MAXC + 1], B[MAXC + 1];
        SBox stack[SMALL_STACK];
        int sp = 0;
        {
                SBox b;
                b.sa = root.sa; b.ea = root.ea; b.sb = root.sb; b.eb = root.eb;
                b.f0.a = root.f0a; b.f0.ga = root.f0ga; b.f0.gb = root.f0gb;
                b.b0.a = root.b0a; b.b0.ga = root.b0ga; b.b0.gb = root.b0gb;
                stack[sp++] = b;
        }
        int* __restrict__ path = J.path;
        const Trip KA = {0.0F, KB_NEGF, KB_NEGF};
        const Trip KGA = {KB_NEGF, 0.0F, KB_NEGF};
        const Trip KGB = {KB_NEGF, KB_NEGF, 0.0F};
        while (sp > 0) {
                const SBox bx = stack[--sp];
                const int sa = bx.sa, ea = bx.ea, sb = bx.sb, eb = bx.eb;
                const int mid = (ea - sa) / 2 + sa;
                small_sweep<V, KR, BONUS>(J, 0, sa, mid, sb, eb, bx.f0, F, s_tbl, tstride);
                small_sweep<V, KR, BONUS>(J, 1, mid, ea, sb, eb, bx.b0, B, s_tbl, tstride);
                ncells += (unsigned long long)(ea - sa) * (unsigned long long)(eb - sb);
                // meet-up
                const float middle = (float)(eb - sb) / 2.0F + (float)sb;
                float x2, x3, x5, x6, x6last, x7;
                if constexpr (is_ss<V>()) {
                        x2 = x3 = x5 = x7 = J.o;
                        x6 = (sb == 0) ? J.t : J.e;
                        x6last = (eb == J.len_b) ? J.t : J.e;
                } else {
                        const float* P = J.prof_r + ((size_t)(mid + 1) << 6);
                        x3 = P[27];
                        x7 = P[-37];
                        x6 = (sb == 0) ? P[29] : P[28];
                        x6last = (eb == J.len_b) ? P[29] : P[28];
                        x2 = x5 = J.o;
                }
                Best m;
                m.max = KB_NEGF; m.max2 = KB_NEGF; m.key = 0x7fffffff;
                for (int i = sb; i <= eb; i++) {
                        const float4 fq = F[i - sb];
                        const float4 bq = B[eb - i];
                        const Trip f = {fq.x, fq.y, fq.z};
                        const Trip b = {bq.x, bq.y, bq.z};
                        float sub = fabsf(middle - (float)i);
                        sub = __fdiv_rn(sub, 1000.0F);
                        const int kb = (i - sb) * 8;
                        if (i < eb) {
                                if constexpr (V == V_PP5 || V == V_PP23) {
                                        x2 = J.prof_c[((size_t)(i + 1) << 6) + 27];
                                        x5 = J.prof_c[((size_t)i << 6) + 27];
                                }
                                offer(m, f.a + b.a - sub, kb + 1);
                                offer(m, f.a + b.ga + x2 - sub, kb + 2);
                                offer(m, f.a + b.gb + x3 - sub, kb + 3);
                                offer(m, f.ga + b.a + x5 - sub, kb + 5);
                                offer(m, f.gb + b.gb + x6 - sub, kb + 6);
                                offer(m, f.gb + b.a + x7 - sub, kb + 7);
                        } else {
                                offer(m, f.a + b.gb + x3 - sub, kb + 3);
                                offer(m, f.gb + b.gb + x6last - sub, kb + 6);
                        }
                }
                if (m.key == 0x7fffffff) {
                        continue;
                }
                const int c = sb + (m.key >> 3);
                const int t = m.key & 7;
                SBox L, Rr;
                L.sa = sa; L.sb = sb; L.f0 = bx.f0;
                Rr.ea = ea; Rr.eb = eb; Rr.b0 = bx.b0;
                switch (t) {
                case 1:
                        path[mid] = c; path[mid + 1] = c + 1;
                        L.ea = mid - 1; L.eb = c - 1; L.b0 = KA;
                        Rr.sa = mid + 1; Rr.sb = c + 1; Rr.f0 = KA;
                        break;
                case 2:
                        path[mid] = c;
                        L.ea = mid - 1; L.eb = c - 1; L.b0 = KA;
                        Rr.sa = mid; Rr.sb = c + 1; Rr.f0 = KGA;
                        break;
                case 3:
                        path[mid] = c;
                        L.ea = mid - 1; L.eb = c - 1; L.b0 = KA;
                        Rr.sa = mid + 1; Rr.sb = c; Rr.f0 = KGB;
                        break;
                case 5:
                        path[mid + 1] = c + 1;
                        L.ea = mid; L.eb = c - 1; L.b0 = KGA;
                        Rr.sa = mid + 1; Rr.sb = c + 1; Rr.f0 = KA;
                        break;
                case 6:
                        L.ea = mid - 1; L.eb = c; L.b0 = KGB;
                        Rr.sa = mid + 1; Rr.sb = c; Rr.f0 = KGB;
                        break;
                default:
                        path[mid + 1] = c + 1;
                        L.ea = mid - 1; L.eb = c; L.b0 = KGB;
                        Rr.sa = mid + 1; Rr.sb = c + 1; Rr.f0 = KA;
                        break;
                }
                if (Rr.sa < Rr.ea && Rr.sb < Rr.eb && sp < SMALL_STACK) stack[sp++] = Rr;
                if (L.sa < L.ea && L.sb < L.eb && sp < SMALL_STACK) stack[sp++] = L;
        }
}

template <int MAXC, int BONUS>
__global__ void __launch_bounds__(128, 4)
kb_small_kernel(const KbJob* __restrict__ jobs, const KbBox* __restrict__ boxes, const unsigned* __restrict__ nsmall_p,
                unsigned long long* __restrict__ cells, const float* __restrict__ tbl, const int tstride)
{
        __shared__ float s_tbl[TBL_MAX];
        for (int i = threadIdx.x; i < TBL_MAX; i += blockDim.x) {
                s_tbl[i] = tbl[i];
        }
        __syncthreads();
        const unsigned n = *nsmall_p;
        const unsigned nth = gridDim.x * blockDim.x;
        for (unsigned b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += nth) {
                const KbBox bx = boxes[b];
                const KbJob J = jobs[bx.job];
                unsigned long long nc = 0;
                if (J.kind == KB200_KIND_SS) small_box_run<V_SS, MAXC, BONUS>(J, bx, s_tbl, tstride, nc);
                else if (J.kind == KB200_KIND_SP) small_box_run<V_SP, MAXC, BONUS>(J, bx, s_tbl, tstride, nc);
                else if (J.nalpha <= 5) small_box_run<V_PP5, MAXC, BONUS>(J, bx, s_tbl, tstride, nc);
                else small_box_run<V_PP23, MAXC, BONUS>(J, bx, s_tbl, tstride, nc);
                atomicAdd(cells + 4 + J.kind, nc);
                if (J.bonus || J.bkey) {
                        atomicAdd(cells + 3, nc);
                }
        }
}

} // namespace

// ---------------------------------------------------------------------------------------------

int kb_run_hirschberg(kb200_ctx* ctx, const float* subm_host, std::vector<KbJob>& jobs)
{
        const int n = (int)jobs.size();
        if (n == 0) {
                return KB200_OK;
        }
        cudaStream_t st = ctx->stream;
        // row buffers, packed column records
        size_t total_cols = 0;
        size_t box_cap = 0;
        size_t unit_cap = 0;
        size_t pack_floats = 0;
        std::vector<int> pp_jobs;
        std::vector<long long> pp_prefix;
        long long pp_cols = 0;
        for (int i = 0; i < n; i++) {
                total_cols += (size_t)(jobs[i].len_a + jobs[i].len_b + 2);
                box_cap += (size_t)std::max(1, jobs[i].len_a);
                // every box contributes <= ceil(rows/32) + 2 units, rows of same-depth boxes are disjoint
                unit_cap += (size_t)jobs[i].len_a / 32 + 2 * (size_t)std::max(1, jobs[i].len_a) + 4;
                if (jobs[i].kind == KB200_KIND_PP) {
                        const int pw = (jobs[i].nalpha <= 5) ? PACK5 : PACK23;
                        pp_jobs.push_back(i);
                        pp_prefix.push_back(pp_cols);
                        pp_cols += jobs[i].len_b + 2;
                        pack_floats += (size_t)(jobs[i].len_b + 2) * pw;
                }
        }
        KB_RUN(ctx->d_rows.ensure(2 * total_cols * sizeof(float4)));
        // stale rows (an earlier batch, an earlier context) must never carry a live hand-off tag
        KB_CUDA(cudaMemsetAsync(ctx->d_rows.p, 0, 2 * total_cols * sizeof(float4), st));
        KB_RUN(ctx->d_pack.ensure(pack_floats * sizeof(float) + 64));
        {
                float4* base = ctx->d_rows.as<float4>();
                float* pbase = ctx->d_pack.as<float>();
                size_t off = 0, poff = 0;
                for (int i = 0; i < n; i++) {
                        const size_t w = (size_t)(jobs[i].len_a + jobs[i].len_b + 2);
                        jobs[i].rowF = base + off;
                        jobs[i].rowB = base + total_cols + off;
                        off += w;
                        if (jobs[i].kind == KB200_KIND_PP) {
                                const int pw = (jobs[i].nalpha <= 5) ? PACK5 : PACK23;
                                jobs[i].cpack = pbase + poff;
                                poff += (size_t)(jobs[i].len_b + 2) * pw;
                        }
                }
        }
        KB_RUN(ctx->d_jobs.ensure(sizeof(KbJob) * (size_t)n));
        KB_CUDA(cudaMemcpyAsync(ctx->d_jobs.p, jobs.data(), sizeof(KbJob) * (size_t)n, cudaMemcpyHostToDevice, st));
        KB_RUN(ctx->d_boxA.ensure(sizeof(KbBox) * box_cap));
        KB_RUN(ctx->d_boxB.ensure(sizeof(KbBox) * box_cap));
        KB_RUN(ctx->d_boxS.ensure(sizeof(KbBox) * box_cap));
        KB_RUN(ctx->d_units.ensure(sizeof(KbUnit) * unit_cap));
        KB_RUN(ctx->d_counters.ensure(128));
        KB_RUN(ctx->d_tbl.ensure(sizeof(float) * TBL_MAX));
        // shared-memory score table: row stride = alphabet size, so that a 5-letter table (25
        // entries) puts every entry in its own bank -- lanes reading different entries never conflict
        int max_alpha = 5;
        for (int i = 0; i < n; i++) {
                max_alpha = std::max(max_alpha, jobs[i].nalpha);
        }
        const int tstride = (max_alpha <= 5) ? 5 : 23;
        {
                std::vector<float> tbl(TBL_MAX, 0.0f);
                for (int i = 0; i < tstride; i++) {
                        for (int j = 0; j < tstride; j++) {
                                tbl[i * tstride + j] = subm_host[i * 23 + j];
                        }
                }
                KB_CUDA(cudaMemcpyAsync(ctx->d_tbl.p, tbl.data(), sizeof(float) * tbl.size(), cudaMemcpyHostToDevice, st));
                KB_CUDA(cudaStreamSynchronize(st));   // tbl is a stack-lifetime vector
        }
        if (!pp_jobs.empty()) {
                KB_RUN(ctx->d_ppidx.ensure(sizeof(int) * pp_jobs.size() + sizeof(long long) * pp_jobs.size() + 64));
                long long* d_pref = ctx->d_ppidx.as<long long>();
                int* d_idx = (int*)(d_pref + pp_jobs.size());
                KB_CUDA(cudaMemcpyAsync(d_pref, pp_prefix.data(), sizeof(long long) * pp_jobs.size(), cudaMemcpyHostToDevice, st));
                KB_CUDA(cudaMemcpyAsync(d_idx, pp_jobs.data(), sizeof(int) * pp_jobs.size(), cudaMemcpyHostToDevice, st));
                const int grid = (int)std::min<long long>((pp_cols + 255) / 256, (long long)ctx->sm_count * 16);
                kb_pack_kernel<<<grid, 256, 0, st>>>(ctx->d_jobs.as<KbJob>(), d_idx, (int)pp_jobs.size(), d_pref, pp_cols);
                KB_CUDA(cudaGetLastError());
                KB_CUDA(cudaStreamSynchronize(st));
                ctx->stats.n_launches += 1;
        }
        std::vector<KbBox> init;
        init.reserve(n);
        size_t rows_total = 0;
        for (int i = 0; i < n; i++) {
                if (jobs[i].len_a <= 0 || jobs[i].len_b <= 0) {
                        continue;
                }
                KbBox b;
                b.job = i; b.sa = 0; b.ea = jobs[i].len_a; b.sb = 0; b.eb = jobs[i].len_b;
                b.f0a = 0.0F; b.f0ga = KB_NEGF; b.f0gb = KB_NEGF;
                b.b0a = 0.0F; b.b0ga = KB_NEGF; b.b0gb = KB_NEGF;
                b.depth = 0;
                init.push_back(b);
                rows_total += (size_t)jobs[i].len_a;
        }
        if (init.empty()) {
                return KB200_OK;
        }
        KB_CUDA(cudaMemcpyAsync(ctx->d_boxA.p, init.data(), sizeof(KbBox) * init.size(), cudaMemcpyHostToDevice, st));
        KB_CUDA(cudaStreamSynchronize(st));
        // counters: [0] sweep cursor, [1] next box count, [2] unit count, [4..11] cells (u64 x4)
        KB_CUDA(cudaMemsetAsync(ctx->d_counters.p, 0, 128, st));
        unsigned count = (unsigned)init.size();
        KbBox* cur = ctx->d_boxA.as<KbBox>();
        KbBox* nxt = ctx->d_boxB.as<KbBox>();
        unsigned int* d_cursor = ctx->d_counters.as<unsigned int>();
        unsigned int* d_next = d_cursor + 1;
        unsigned int* d_nunits = d_cursor + 2;
        unsigned long long* d_cells = (unsigned long long*)(d_cursor + 4);   // [ss, sp, pp, bonus]
        unsigned int* d_nsmall = d_cursor + 3;
        KbBox* d_small = ctx->d_boxS.as<KbBox>();
        const bool use_small = getenv("KB200_NO_SMALL") == nullptr;
        // thread-per-box threshold: with enough jobs to fill the machine the deep rounds (boxes of a
        // few dozen rows: 1-2 live lanes per warp in a sweep) are cheaper as serial per-thread work
        int small_rows = SMALL_ROWS, small_cols = SMALL_COLS;
        {
                // largest power of two T such that the boxes of T rows (about rows_total / T of them)
                // still give every resident thread of half the machine a box of its own
                const size_t want = (size_t)ctx->sm_count * 256;
                while (small_rows * 2 <= SMALL_ROWS_MAX && rows_total / (size_t)(small_rows * 2) >= want) {
                        small_rows *= 2;
                }
                if (small_rows > SMALL_ROWS) small_cols = std::min(2 * small_rows + small_rows / 2, SMALL_COLS_MAX);
        }
        if (const char* e = getenv("KB200_SMALL_ROWS")) small_rows = std::min(std::max(atoi(e), 1), SMALL_ROWS_MAX);
        if (const char* e = getenv("KB200_SMALL_COLS")) small_cols = std::min(std::max(atoi(e), 4), SMALL_COLS_MAX);
        // KB200_THIN=0|1 forces thick / thin strips (tests run every batch both ways)
        int force_thin = -1;
        if (const char* e = getenv("KB200_THIN")) force_thin = (atoi(e) != 0) ? 1 : 0;
        float sweep_ms = 0.0f;
        const bool trace = getenv("KB200_TRACE") != nullptr;
        int round = 0;
        // all-or-nothing: a job without bonus in a bonus batch simply has an empty list / null dense
        bool batch_bonus = false, batch_dense = false;
        for (int i = 0; i < n; i++) {
                if (jobs[i].bonus || jobs[i].bkey) batch_bonus = true;
                if (jobs[i].bonus) batch_dense = true;
                if (jobs[i].bkey && jobs[i].nb > KB_BONUS_KMAX) {
                        fprintf(stderr, "[kalign_b200] sparse bonus with %d entries per row (max %d)\n", jobs[i].nb, KB_BONUS_KMAX);
                        return KB200_FAIL;
                }
        }
        {
                // 4 resident CTAs per SM: 15 KB of rings and tables per CTA, + 32 KB of staged bonus lists
                // in the bonus family (4 x 48 KB: the whole shared-memory carve-out)
                static bool carveout_set = false;
                if (!carveout_set) {
                        cudaFuncSetAttribute(kb_sweep_kernel<BONUS_SPARSE>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
                        cudaFuncSetAttribute(kb_sweep_kernel<BONUS_NONE>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
                        cudaFuncSetAttribute(kb_sweep_kernel<BONUS_DENSE>, cudaFuncAttributePreferredSharedMemoryCarveout, 40);
                        carveout_set = true;
                }
        }
        // resident warps of the persistent sweep grid
        const int sweep_ctas = ctx->sm_count * 4;
        const size_t resident_warps = (size_t)sweep_ctas * WARPS_PER_CTA;
        KB_CUDA(cudaEventRecord(ctx->ev0, st));
        while (count > 0) {
                KB_CUDA(cudaMemsetAsync(d_cursor, 0, 12, st));
                // thin strips when thick ones could not occupy the machine: the rows still alive
                // at this depth are at most rows_total, spread over `count` boxes
                const size_t thick_units = rows_total / 128 + 2 * (size_t)count;
                int thin = (thick_units < 2 * resident_warps) ? 1 : 0;
                if (force_thin >= 0) thin = force_thin;
                const unsigned items = 2u * count;
                // tags: unique per launch (8192 strips per sweep at most: 256k rows), never 0
                ctx->tag_counter += 65536u;
                if (ctx->tag_counter > 0xfff00000u) ctx->tag_counter = 65536u;
                const unsigned tag_base = ctx->tag_counter;
                kb_plan_kernel<<<(items + 127) / 128, 128, 0, st>>>(ctx->d_jobs.as<KbJob>(), cur, (int)count, thin, batch_bonus ? 1 : 0,
                                                                     ctx->d_units.as<KbUnit>(), d_nunits);
                KB_CUDA(cudaEventRecord(ctx->ev2, st));
                {
                        const KbJob* dj = ctx->d_jobs.as<KbJob>();
                        const KbUnit* du = ctx->d_units.as<KbUnit>();
                        const float* dt = ctx->d_tbl.as<float>();
                        const int thr = WARPS_PER_CTA * 32;
                        if (batch_dense) {
                                kb_sweep_kernel<BONUS_DENSE><<<sweep_ctas, thr, 0, st>>>(dj, cur, du, d_nunits, d_cursor, tag_base, dt, thin, tstride);
                        } else if (batch_bonus) {
                                kb_sweep_kernel<BONUS_SPARSE><<<sweep_ctas, thr, 0, st>>>(dj, cur, du, d_nunits, d_cursor, tag_base, dt, thin, tstride);
                        } else {
                                kb_sweep_kernel<BONUS_NONE><<<sweep_ctas, thr, 0, st>>>(dj, cur, du, d_nunits, d_cursor, tag_base, dt, thin, tstride);
                        }
                }
                KB_CUDA(cudaEventRecord(ctx->ev3, st));
                int mgrid = (int)std::min<unsigned>((count + 3) / 4, (unsigned)(ctx->sm_count * 16));
                kb_meetup_kernel<<<mgrid, 128, 0, st>>>(ctx->d_jobs.as<KbJob>(), cur, (int)count, nxt, d_next,
                                                        use_small ? d_small : nullptr, d_nsmall, small_rows, small_cols, d_cells);
                KB_CUDA(cudaGetLastError());
                unsigned host_counts[4] = {0, 0, 0, 0};
                KB_CUDA(cudaMemcpyAsync(host_counts, d_cursor, sizeof(host_counts), cudaMemcpyDeviceToHost, st));
                KB_CUDA(cudaStreamSynchronize(st));
                const unsigned next_count = host_counts[1];
                float ms = 0.0f;
                cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3);
                sweep_ms += ms;
                if (trace) {
                        fprintf(stderr, "[kb200 trace] jobs=%d round=%d boxes=%u units=%u thin=%d sweep_ms=%.3f\n", n, round, count,
                                host_counts[2], thin, ms);
                }
                round++;
                ctx->stats.n_boxes += count;
                ctx->stats.n_launches += 3;
                if ((size_t)next_count > box_cap || (size_t)host_counts[2] > unit_cap || (size_t)host_counts[3] > box_cap) {
                        fprintf(stderr, "[kalign_b200] work-list overflow (boxes %u > %zu or units %u > %zu)\n", next_count, box_cap,
                                host_counts[2], unit_cap);
                        return KB200_FAIL;
                }
                count = next_count;
                std::swap(cur, nxt);
        }
        {
                // every box that became small during the rounds: finish its recursion in one launch
                KB_CUDA(cudaEventRecord(ctx->ev2, st));
                const int sgrid = ctx->sm_count * 4;
                const KbJob* dj = ctx->d_jobs.as<KbJob>();
                const float* dt = ctx->d_tbl.as<float>();
                const int fam = batch_dense ? BONUS_DENSE : (batch_bonus ? BONUS_SPARSE : BONUS_NONE);
                if (small_cols <= SMALL_COLS) {
                        if (fam == BONUS_DENSE) kb_small_kernel<SMALL_COLS, BONUS_DENSE><<<sgrid, 128, 0, st>>>(dj, d_small, d_nsmall, d_cells, dt, tstride);
                        else if (fam == BONUS_SPARSE) kb_small_kernel<SMALL_COLS, BONUS_SPARSE><<<sgrid, 128, 0, st>>>(dj, d_small, d_nsmall, d_cells, dt, tstride);
                        else kb_small_kernel<SMALL_COLS, BONUS_NONE><<<sgrid, 128, 0, st>>>(dj, d_small, d_nsmall, d_cells, dt, tstride);
                } else {
                        if (fam == BONUS_DENSE) kb_small_kernel<SMALL_COLS_MAX, BONUS_DENSE><<<sgrid, 128, 0, st>>>(dj, d_small, d_nsmall, d_cells, dt, tstride);
                        else if (fam == BONUS_SPARSE) kb_small_kernel<SMALL_COLS_MAX, BONUS_SPARSE><<<sgrid, 128, 0, st>>>(dj, d_small, d_nsmall, d_cells, dt, tstride);
                        else kb_small_kernel<SMALL_COLS_MAX, BONUS_NONE><<<sgrid, 128, 0, st>>>(dj, d_small, d_nsmall, d_cells, dt, tstride);
                }
                KB_CUDA(cudaGetLastError());
                KB_CUDA(cudaEventRecord(ctx->ev3, st));
                KB_CUDA(cudaEventSynchronize(ctx->ev3));
                float sms = 0.0f;
                cudaEventElapsedTime(&sms, ctx->ev2, ctx->ev3);
                ctx->stats.small_seconds += 1e-3 * (double)sms;
                ctx->stats.n_launches += 1;
                if (trace) {
                        unsigned ns = 0;
                        cudaMemcpy(&ns, d_nsmall, sizeof(unsigned), cudaMemcpyDeviceToHost);
                        fprintf(stderr, "[kb200 trace] jobs=%d small(<=%dx%d) boxes=%u small_ms=%.3f\n", n, small_rows, small_cols, ns, sms);
                }
        }
        KB_CUDA(cudaEventRecord(ctx->ev1, st));
        unsigned long long cells[7] = {0, 0, 0, 0, 0, 0, 0};   // sweep ss/sp/pp, bonus, small ss/sp/pp
        KB_CUDA(cudaMemcpyAsync(cells, d_cells, sizeof(cells), cudaMemcpyDeviceToHost, st));
        KB_CUDA(cudaStreamSynchronize(st));
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        ctx->stats.dp_seconds += 1e-3 * (double)ms;
        ctx->stats.sweep_seconds += 1e-3 * (double)sweep_ms;
        for (int k = 0; k < 3; k++) {
                ctx->stats.dp_cells += (double)cells[k] + (double)cells[4 + k];
        }
        ctx->stats.cells_ss += (double)cells[0] + (double)cells[4];
        ctx->stats.cells_sp += (double)cells[1] + (double)cells[5];
        ctx->stats.cells_pp += (double)cells[2] + (double)cells[6];
        ctx->stats.cells_bonus += (double)cells[3];
        ctx->stats.small_ss += (double)cells[4];
        ctx->stats.small_sp += (double)cells[5];
        ctx->stats.small_pp += (double)cells[6];
        return KB200_OK;
}
