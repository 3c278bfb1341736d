// kb_sweep.cuh -- the wavefront sweep of the batched Hirschberg engine: cell routines, strips, the
// persistent sweep kernel template.  Included by one translation unit per kernel family
// (kb_sweep_none.cu / kb_sweep_sparse.cu / kb_sweep_dense.cu instantiate kb_sweep_kernel<BONUS>,
// so that the three families compile in parallel) and by kb_dp.cu (the thread-per-box kernel shares
// the cell routines).  See kb_dp.cu for the formulation and the references into lib/src.
#pragma once
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

// score-table read by 32-bit shared address: one integer add + LDS per cell (the table is read-only
// while a kernel runs)
__device__ __forceinline__ float lds_f32(const unsigned addr)
{
        float v;
        asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
        return v;
}

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
                // thin regime: `thin` rows per lane (1, 2 or 4; 23-letter profile-profile strips have at most 2)
                return (kind == KB200_KIND_PP && nalpha > 5 && thin > 2) ? 64 : 32 * thin;
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
        int rbase[K];                                   // SS: row residue * table stride
        unsigned taddr[K];                              // SS: shared-memory byte address of the row's table line
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
                                      const ColCtx<V>& cc,
                                      const int bdir, int (&sp_i)[K], int (&sp_c)[K], float (&sp_v)[K],
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
                        const unsigned c4 = (unsigned)cc.cres * 4u;
                        const float2 x = add2(make_float2(lds_f32(rc.taddr[k] + c4), lds_f32(rc.taddr[k + 1] + c4)),
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
                                      const ColCtx<V>& cc,
                                      const int bdir, int (&sp_i)[K], int (&sp_c)[K], float (&sp_v)[K],
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
                                const float x = lds_f32(rc.taddr[k] + (unsigned)cc.cres * 4u) + J.nsoff;
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

// 23-letter seq-seq rows: shared-memory byte address of every row's line of the score table
template <int V, int K>
__device__ __forceinline__ void set_table_addr(RowCtx<V, K>& rc, const float* __restrict__ s_tbl)
{
        if constexpr (V == V_SS) {
                const unsigned tb = (unsigned)__cvta_generic_to_shared(s_tbl);
#pragma unroll
                for (int k = 0; k < K; k++) {
                        rc.taddr[k] = tb + (unsigned)rc.rbase[k] * 4u;
                }
        }
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
        if constexpr (!BSM) {
                // thread-per-box kernel: the lists are read from global memory.  All keys of the K rows are
                // loaded independently (one latency, not one per probed entry): the lists are sorted, so
                // the first entry in sweep direction is a count of keys below / up to the bound.
                int e0[K];
#pragma unroll
                for (int k = 0; k < K; k++) {
                        const int* __restrict__ bc = J.bkey + (size_t)rc.irow[k] * (size_t)KS;
                        int cnt = 0;
#pragma unroll
                        for (int x = 0; x < KB_BONUS_KMAX; x++) {
                                if (x < KS) {
                                        const int key = __ldg(bc + x);
                                        cnt += bwd ? ((key <= eb - 1) ? 1 : 0) : ((key < sb + 1) ? 1 : 0);
                                }
                        }
                        e0[k] = bwd ? (cnt - 1) : cnt;
                }
#pragma unroll
                for (int k = 0; k < K; k++) {
                        const int e = e0[k];
                        if (e >= 0 && e < KS) {
                                const size_t o = (size_t)rc.irow[k] * (size_t)KS + (size_t)e;
                                sp_c[k] = __ldg(J.bkey + o);
                                sp_v[k] = __ldg(J.bval + o);
                        }
                        sp_i[k] = e;
                        if (!bwd && eb == J.len_b && rc.irow[k] + 1 < J.len_a) {
                                // flat index (i, len_b) of the reference's dense matrix is (i+1, 0)
                                const size_t o = (size_t)(rc.irow[k] + 1) * (size_t)KS;
                                if (__ldg(J.bkey + o) == 0) sp_wrap[k] = __ldg(J.bval + o);
                        }
                }
                return;
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

// ---- sparse consistency bonus of the warp strips: EVENTS instead of a compare in every cell -----
// A row carries at most nb (<= 8) bonus entries over thousands of columns.  The strip therefore
// runs the bonus-free cell routine and keeps, per lane, ONE sorted queue of the (column, row, value)
// entries of its K rows in shared memory (lane-private slots, conflict-free).  A step costs one
// integer compare (column of the next event == column of the lane); on a hit -- lane-divergent,
// rare -- the value is added to the cell's `a` exactly where the reference adds its matrix entry
// (after the substitution / profile term: aln_seqseq.c:83-85, aln_profileprofile.c:107-109) and the
// gb chain down the lane's rows of this column is recomputed from the corrected value.  All other
// cells add the reference's +0.0f, which never changes a value.
constexpr int BON_QCAP = BON_SLOTS * BON_KMAX_ROWS;          // events per lane
constexpr int BON_QSLOTS = BON_QCAP + 2 + BON_KMAX_ROWS;     // + two sentinels + the rows' wrap values
constexpr int BON_DYN_SMEM = WARPS_PER_CTA * BON_QSLOTS * 32 * 8;   // bytes of dynamic shared memory, sparse family

template <int V, int K, bool TAIL, bool EDGE, bool ALLROWS>
__device__ __forceinline__ void bonus_fix(const KbJob& J, const RowCtx<V, K>& rc, const unsigned vmask, const Trip got,
                                          const int k0, const float val, const int2* __restrict__ wrapv, const bool e_term,
                                          float (&sA)[K], float (&sGA)[K], float (&sGB)[K], Trip& u)
{
        Trip p = got;
#pragma unroll
        for (int k = 0; k < K; k++) {
                float RO, RE, RT;
                if constexpr (is_ss<V>()) {
                        RO = J.o; RE = J.e; RT = J.t;
                } else {
                        RO = rc.RO[k]; RE = rc.RE[k]; RT = rc.RT[k];
                }
                float a = sA[k];
                if constexpr (ALLROWS) {
                        a = a + __int_as_float(wrapv[k * 32].x);     // forward sweep, j == len_b: flat index wraps to (i+1, 0)
                } else {
                        if (k == k0) a = a + val;
                }
                float gb = kmax(p.gb + RE, p.a + RO);
                if constexpr (EDGE) {
                        const float gt = kmax(p.gb, p.a) + RT;
                        gb = e_term ? gt : gb;
                }
                float ga = sGA[k];
                if constexpr (TAIL) {
                        if (!((vmask >> k) & 1u)) {
                                a = p.a; ga = p.ga; gb = p.gb;
                        }
                }
                sA[k] = a; sGA[k] = ga; sGB[k] = gb;
                p.a = a; p.ga = ga; p.gb = gb;
        }
        u = p;
}

// build the lane's event queue: entries of the valid rows whose column the sweep visits, sorted by
// column.  q[0] / q[n+1] are sentinels that never match a column; returns the index of the first
// event in sweep direction.  Also stages the rows' wrap values (forward sweep ending at len_b).
template <int V, int K>
__device__ __forceinline__ int bonus_queue_init(const KbJob& J, const int bwd, const int sb, const int eb, const RowCtx<V, K>& rc,
                                                const unsigned vmask, int2* __restrict__ q, bool& wrap_on)
{
        static_assert(K <= BON_KMAX_ROWS, "event queue capacity");
        const int nb = J.nb;
        // forward cells visit j = sb+1 .. eb, backward cells j = eb-1 .. sb (the first column of a sweep takes no bonus)
        const int lo = bwd ? sb : sb + 1;
        const int hi = bwd ? eb - 1 : eb;
        int n = 0;
        __syncwarp();
        q[0] = make_int2((int)0x80000000, 0);
        wrap_on = false;
        if (J.bkey != nullptr) {
#pragma unroll
                for (int k = 0; k < K; k++) {
                        if ((vmask >> k) & 1u) {
                                const int* __restrict__ bc = J.bkey + (size_t)rc.irow[k] * (size_t)nb;
                                const float* __restrict__ bv = J.bval + (size_t)rc.irow[k] * (size_t)nb;
                                for (int e = 0; e < nb; e++) {
                                        const int c = __ldg(bc + e);
                                        if (c >= lo && c <= hi) {
                                                // insertion from the back: rows are diagonal-correlated, the queue is nearly sorted already
                                                const int key = (c << 3) | k;
                                                const int vb = __float_as_int(__ldg(bv + e));
                                                int y = n;
                                                while (y >= 1 && q[y * 32].x > key) {
                                                        q[(y + 1) * 32] = q[y * 32];
                                                        y--;
                                                }
                                                q[(y + 1) * 32] = make_int2(key, vb);
                                                n++;
                                        }
                                }
                        }
                }
                if (!bwd && eb == J.len_b) {
                        float anyw = 0.0f;
#pragma unroll
                        for (int k = 0; k < K; k++) {
                                float w = 0.0f;
                                if (((vmask >> k) & 1u) && rc.irow[k] + 1 < J.len_a) {
                                        const size_t o = (size_t)(rc.irow[k] + 1) * (size_t)nb;
                                        if (__ldg(J.bkey + o) == 0) w = __ldg(J.bval + o);
                                }
                                q[(BON_QCAP + 2 + k) * 32] = make_int2(__float_as_int(w), 0);
                                anyw = (w != 0.0f) ? 1.0f : anyw;
                        }
                        wrap_on = (anyw != 0.0f);
                }
        }
        q[(n + 1) * 32] = make_int2(0x7fffffff, 0);
        __syncwarp();
        return bwd ? n : 1;
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
                            int2* s_bon, unsigned long long* s_mbar, unsigned& rec_phase)
{
        static_assert(BONUS != BONUS_SPARSE || K <= BON_KMAX_ROWS, "bonus event queue is sized for K <= 4");
        constexpr int NA = VTraits<V>::NA;
        constexpr int PW4 = (V == V_PP5) ? (PACK5 / 4) : (PACK23 / 4);
        const int C = eb - sb;
        const int R = r1 - r0;
        RowCtx<V, K> rc;
        const unsigned vmask = load_rows<V, K>(J, bwd, r0, r1, row0 + lane * K, tstride, rc);
        set_table_addr<V, K>(rc, s_tbl);
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
        // Record ring: 64 columns x 32 B, four blocks of 16 columns.  Ring position of column u is
        // u + 15 (block 0 holds column 0 only), so that the block a sweep prefetches into never
        // overlaps the three blocks the 32 lanes are reading.  A block is ONE bulk asynchronous copy
        // (cp.async.bulk, the TMA engine's linear form) issued by one elected lane: 16 packed column
        // records are contiguous in the job's cpack array in either sweep direction; completion is
        // signalled on the block's mbarrier (expect_tx / complete_tx), which the warp waits on
        // one block ahead of use.  Backward sweeps visit the columns in descending order: the same
        // contiguous run, read back to front (slot index XOR 15).
        float4* const s_ring_rec = s_rec;
        const int rec_flip = bwd ? 15 : 0;
        auto rec_slot = [&](const int col) -> const float4* {
                return s_ring_rec + ((((col + 15) & 63) ^ rec_flip) << 1);
        };
        auto rec_issue = [&](const int j) {          // block j: columns max(0, 16j-15) .. min(C, 16j)
                if constexpr (RECRING) {
                        const int c0 = (j == 0) ? 0 : (16 * j - 15);
                        const int c1 = (16 * j < C) ? (16 * j) : C;
                        if (c0 <= c1) {
                                __syncwarp();        // every lane is done reading the block this copy overwrites
                                if (lane == 0) {
                                        const int n = c1 - c0 + 1;
                                        const long long r0 = bwd ? (long long)(eb - c1 + 1) : (long long)(sb + c0);
                                        const int s0 = bwd ? (15 - ((c1 + 15) & 15)) : ((c0 + 15) & 15);
                                        const float* src = J.cpack + r0 * PACK5;
                                        const unsigned dst = (unsigned)__cvta_generic_to_shared(s_ring_rec + (((j & 3) * 16 + s0) << 1));
                                        const unsigned bar = (unsigned)__cvta_generic_to_shared(s_mbar + (j & 3));
                                        const unsigned bytes = (unsigned)n * (unsigned)(PACK5 * sizeof(float));
                                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                                     ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
                                }
                        }
                }
        };
        auto rec_wait = [&](const int j) {           // block j has landed (no-op for a block without columns)
                if constexpr (RECRING) {
                        const int c0 = (j == 0) ? 0 : (16 * j - 15);
                        if (c0 <= C) {
                                const unsigned bar = (unsigned)__cvta_generic_to_shared(s_mbar + (j & 3));
                                const unsigned parity = (rec_phase >> (j & 3)) & 1u;
                                unsigned done = 0;
                                do {
                                        asm volatile("{\n\t.reg .pred p;\n\t"
                                                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                                                     "selp.u32 %0, 1, 0, p;\n\t}"
                                                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
                                } while (!done);
                                rec_phase ^= 1u << (j & 3);
                        }
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
                                rec_wait((t >> 4) + 1);          // columns t+1 .. t+16 (issued 16 steps ago)
                                rec_issue((t >> 4) + 2);         // slots of columns t-47 .. t-32: last read at step t-2
                        }
                }
        };
        static_assert(HB == 8, "lane groups of 8");
        const int steps = C + 32;
        // column input of the current step (filled one step ahead); lane 0 starts on column 0 at t=0
        float4 curA = make_float4(0.f, 0.f, 0.f, 0.f), curB = curA;
        int cur_cres = 0;
        // sparse consistency bonus: the lane's event queue (see bonus_fix / bonus_queue_init above)
        int2* const my_bon = s_bon + lane;       // lane-private slots [slot][lane]
        const int bdir = bwd ? -1 : 1;
        int ev_i = 0;                            // queue index of the next event in sweep direction
        int ev_col = 0x7fffffff;                 // its column (sentinel: never equal to a lane's column)
        bool wrap_on = false;
        if constexpr (BONUS == BONUS_SPARSE) {
                ev_i = bonus_queue_init<V, K>(J, bwd, sb, eb, rc, vmask, my_bon, wrap_on);
                ev_col = my_bon[ev_i * 32].x >> 3;
        }
        // the cell routine runs bonus-free in the sparse family (events) -- dummies for its list arguments
        constexpr int CB = (BONUS == BONUS_SPARSE) ? BONUS_NONE : BONUS;
        int sp_i[K], sp_c[K];
        float sp_v[K], sp_wrap[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
                sp_i[k] = 0; sp_c[k] = 0; sp_v[k] = 0.0f; sp_wrap[k] = 0.0f;
        }
        if constexpr (RECRING) {
                static_assert(PACK5 == 8, "PP5 record is two float4");
                rec_issue(0);
                rec_issue(1);
                rec_issue(2);
                rec_wait(0);
                rec_wait(1);
                // lane 0 is on column 0 at step 0; the other lanes read their column 0 one step ahead
                curA = rec_slot(0)[0];
                curB = rec_slot(0)[1];
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
                        const float4* rp = rec_slot(u + 1);
                        nxtA = rp[0];
                        nxtB = rp[1];
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
                                cells<V, K, TAIL, MODE_MID, CB, true>(J, rc, vmask, first_term, last_term, cc, bdir, sp_i, sp_c, sp_v, sp_wrap, my_bon, s_tbl, sA, sGA, sGB, d, up);
                        } else {
                                cells<V, K, TAIL, MODE_EDGE, CB, true>(J, rc, vmask, first_term, last_term, cc, bdir, sp_i, sp_c, sp_v, sp_wrap, my_bon, s_tbl, sA, sGA, sGB, d, up, u == 0, u == C);
                        }
                        if constexpr (BONUS == BONUS_SPARSE) {
                                if (j == ev_col) {
                                        // rare, lane-divergent: bonus entries of this lane's rows in this column
                                        const bool e_term = !STEADY && ((u == 0 && first_term) || (u == C && last_term));
                                        do {
                                                const int2 q = my_bon[ev_i * 32];
                                                bonus_fix<V, K, TAIL, !STEADY, false>(J, rc, vmask, got, q.x & 7, __int_as_float(q.y), my_bon, e_term,
                                                                                      sA, sGA, sGB, up);
                                                ev_i += bdir;
                                                ev_col = my_bon[ev_i * 32].x >> 3;
                                        } while (j == ev_col);
                                }
                                if constexpr (!STEADY) {
                                        if (wrap_on && u == C) {
                                                bonus_fix<V, K, TAIL, true, true>(J, rc, vmask, got, 0, 0.0f, my_bon + (BON_QCAP + 2) * 32, last_term,
                                                                                  sA, sGA, sGB, up);
                                        }
                                }
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
#pragma unroll 1
                        for (int q = 0; q < HB; q += 2) {
                                step(std::true_type{}, t);
                                step(std::true_type{}, t + 1);
                                t += 2;
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
        // every block that was issued has been waited for inside the loop (block j is waited at step
        // 16(j-1) < C + 32 whenever it has a column), so no copy outlives the strip
        __syncwarp();
}

template <int V, int BONUS>
__device__ void sweep_unit(const KbJob& J, const KbBox& bx, const int bwd, const int strip, const int thin,
                           const unsigned tag_base, const float* __restrict__ s_tbl, const int tstride, const int lane,
                           float4* s_ring, float4* s_rec, int2* s_bon, unsigned long long* s_mbar, unsigned& rec_phase)
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
        const int rem = (R - row0 < rps) ? (R - row0) : rps;     // rows of THIS strip
        // rows per lane of this strip: the full width of the kind's strips, or -- last strip of a
        // sweep, boxes of the deeper rounds -- the smallest K that covers the remaining rows, so that
        // a 187-row half box runs with 6 rows per lane (97 % of the lanes' rows live) instead of 8
#define KB_STRIP(KK, TT) sweep_strip<V, KK, TT, BONUS>(J, bwd, sb, eb, r0, r1, row0, first_term, last_term, in, rowbuf, prev, mine, s_tbl, tstride, lane, s_ring, s_rec, s_bon, s_mbar, rec_phase)
        const int kneed = (rem + 31) >> 5;
        if (kneed <= 1) {
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
                else if (rem == 64) KB_STRIP(2, false);      // full strips of the thin regime (two rows per lane)
                else KB_STRIP(2, true);
        }
#undef KB_STRIP
}

// ---- sub-warp groups for rounds of small seq-seq boxes -------------------------------------------
// A sweep of R <= 128 rows leaves most of a 32-lane strip idle (R = 46: two rows per lane at 72 %
// lane use, 32 fill/drain steps on ~190 columns).  In such rounds a warp is cut into groups of W = 8
// or 16 lanes; every group sweeps its OWN (box, direction) as a single strip -- W x K rows, K the
// smallest of {2,3,4,6,8} that covers the warp's tallest sweep -- with a skew of only W columns.  No
// hand-off: the init row is generated, the last lane of the group writes the bottom row for the
// meet-up.  Same cell routine, same operands, same order as the strips.
template <int V, int K, int W>
__device__ void sweep_group(const KbJob* __restrict__ jobs, const KbBox* __restrict__ boxes, const KbUnit* __restrict__ units,
                            const unsigned unit, const bool valid, const unsigned out_tag,
                            const float* __restrict__ s_tbl, const int tstride, const int lane, float4* s_rec)
{
        static_assert(is_ss<V>(), "groups sweep sequence-sequence boxes");
        const int glane = lane & (W - 1);
        // per-lane view of the group's unit (an invalid group mirrors unit 0 of the launch and computes nothing)
        const KbUnit un = units[valid ? unit : 0u];
        const KbBox bx = boxes[un.item >> 1];
        const int bwd = un.item & 1;
        const KbJob J = jobs[bx.job];
        const int mid = (bx.ea - bx.sa) / 2 + bx.sa;
        const int r0 = bwd ? mid : bx.sa;
        const int r1 = bwd ? bx.ea : mid;
        const int sb = bx.sb, eb = bx.eb;
        const int C = eb - sb;
        const bool first_term = bwd ? (eb == J.len_b) : (sb == 0);
        const bool last_term = bwd ? (sb == 0) : (eb == J.len_b);
        float4* __restrict__ rowbuf = (bwd ? J.rowB : J.rowF) + (bx.sa + bx.sb);
        Trip in;
        if (bwd) {
                in.a = bx.b0a; in.ga = bx.b0ga; in.gb = bx.b0gb;
        } else {
                in.a = bx.f0a; in.ga = bx.f0ga; in.gb = bx.f0gb;
        }
        RowCtx<V, K> rc;
        const unsigned vmask = load_rows<V, K>(J, bwd, r0, r1, glane * K, tstride, rc);
        set_table_addr<V, K>(rc, s_tbl);
        if constexpr (V == V_SS5) {
                float* sp = reinterpret_cast<float*>(s_rec) + lane;
                __syncwarp();
#pragma unroll
                for (int k = 0; k < K; k++) {
#pragma unroll
                        for (int c = 0; c < 5; c++) {
                                sp[(k * 5 + c) * 32] = s_tbl[rc.rbase[k] + c] + J.nsoff;
                        }
                }
                rc.sprow = sp;
                __syncwarp();
        }
        float sA[K], sGA[K], sGB[K];
        int sp_i[K], sp_c[K];
        float sp_v[K], sp_wrap[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
                sA[k] = KB_NEGF; sGA[k] = KB_NEGF; sGB[k] = KB_NEGF;
                sp_i[k] = 0; sp_c[k] = 0; sp_v[k] = 0.0f; sp_wrap[k] = 0.0f;
        }
        Trip d = {KB_NEGF, KB_NEGF, KB_NEGF};
        Trip bot = {KB_NEGF, KB_NEGF, KB_NEGF};
        float genA = in.a, genGA = in.ga;
        const int dstep = bwd ? -1 : 1;
        int jcur = bwd ? (eb + glane) : (sb - glane);
        const long long pr_first = bwd ? (long long)(eb - (1 - glane)) : (long long)(sb + (1 - glane)) - 1;
        const uint8_t* seqp = J.seq_c + pr_first;
        int cur_cres = 0;
        // the warp runs the steps of its longest group; steps in which every lane of every valid group
        // sits on an interior column take the cheaper interior-column routine
        const int cmax = (int)__reduce_max_sync(FULL, (unsigned)(valid ? C : 0));
        const int cmin = (int)__reduce_min_sync(FULL, (unsigned)(valid ? C : 0x7fffffff));
        const int steps = cmax + W;
        ColCtx<V> cc;
        cc.CO = J.o; cc.CE = J.e; cc.COp = J.o;
        const float CT = J.t;
        auto step = [&](auto steady_tag, const int t) {
                constexpr bool STEADY = decltype(steady_tag)::value;
                const int u = t - glane;
                Trip up;
                up.a = __shfl_up_sync(FULL, bot.a, 1, W);
                up.ga = __shfl_up_sync(FULL, bot.ga, 1, W);
                up.gb = __shfl_up_sync(FULL, bot.gb, 1, W);
                const bool act = valid && (STEADY || ((u >= 0) && (u <= C)));
                int ncres = 0;
                {
                        const int pu = u + 1;
                        if (valid && (STEADY || (pu >= 1 && pu <= C))) {
                                ncres = (int)__ldg(seqp);
                        }
                }
                if (act) {
                        cc.jcol = jcur;
                        cc.cres = cur_cres;
                        if (glane == 0) {
                                // the init row (aln_seqseq.c:43-58), one column per step
                                if (!STEADY && u == 0) {
                                        up = in;
                                } else if (STEADY || u < C) {
                                        const float nga = first_term ? (kmax(genGA, genA) + CT) : kmax(genGA + cc.CE, genA + cc.CO);
                                        up.a = KB_NEGF; up.ga = nga; up.gb = KB_NEGF;
                                        genA = KB_NEGF; genGA = nga;
                                } else {
                                        up.a = KB_NEGF; up.ga = KB_NEGF; up.gb = KB_NEGF;
                                }
                        }
                        const Trip got = up;
                        if constexpr (STEADY) {
                                cells<V, K, true, MODE_MID, BONUS_NONE, true>(J, rc, vmask, first_term, last_term, cc, dstep, sp_i, sp_c, sp_v, sp_wrap, nullptr, s_tbl, sA, sGA, sGB, d, up);
                        } else {
                                cells<V, K, true, MODE_EDGE, BONUS_NONE, true>(J, rc, vmask, first_term, last_term, cc, dstep, sp_i, sp_c, sp_v, sp_wrap, nullptr, s_tbl, sA, sGA, sGB, d, up, u == 0, u == C);
                        }
                        d = got;
                        bot = up;
                        if (glane == W - 1) {
                                rowbuf[u] = make_float4(bot.a, bot.ga, bot.gb, __uint_as_float(out_tag));
                        }
                }
                cur_cres = ncres;
                jcur += dstep;
                seqp += dstep;
        };
        int t = 0;
        const int t_fill = (steps < W) ? steps : W;
        for (; t < t_fill; t++) step(std::false_type{}, t);
        // W <= t <= cmin - 1: every lane of every valid group has 1 <= u <= C - 1
        for (; t + 2 <= cmin; t += 2) {
                step(std::true_type{}, t);
                step(std::true_type{}, t + 1);
        }
        for (; t < steps; t++) step(std::false_type{}, t);
        __syncwarp();
}

template <int V, int W>
__device__ __forceinline__ void sweep_group_k(const int kneed, const KbJob* __restrict__ jobs, const KbBox* __restrict__ boxes,
                                              const KbUnit* __restrict__ units, const unsigned unit, const bool valid, const unsigned out_tag,
                                              const float* __restrict__ s_tbl, const int tstride, const int lane, float4* s_rec)
{
#define KB_GROUP(KK) sweep_group<V, KK, W>(jobs, boxes, units, unit, valid, out_tag, s_tbl, tstride, lane, s_rec)
        if constexpr (W == 16) {
                // 64 < rows <= 128
                if (kneed <= 6) KB_GROUP(6); else KB_GROUP(8);
        } else {
                if (kneed <= 2) KB_GROUP(2);
                else if (kneed == 3) KB_GROUP(3);
                else if (kneed == 4) KB_GROUP(4);
                else if (kneed <= 6) KB_GROUP(6);
                else KB_GROUP(8);
        }
#undef KB_GROUP
}

// BONUS is a kernel-level template parameter: a batch either carries consistency bonuses for all
// of its jobs (tree levels in default mode) or for none (anchor batch, --fast), and the two
// families get independent register allocation / code size.
template <int BONUS>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 4)
kb_sweep_kernel(const KbJob* __restrict__ jobs, const KbBox* __restrict__ boxes,
                const KbUnit* __restrict__ units, KbRound* __restrict__ rnd, const unsigned tag_base,
                const float* __restrict__ tbl, const int tstride)
{
        // the round's unit count and strip regime were written by kb_plan_kernel (device memory:
        // the host enqueues every round of a call without reading anything back)
        const unsigned total = rnd->nunits;
        if (total == 0u) {
                return;
        }
        const int thin = (int)rnd->thin;
        unsigned int* const cursor = &rnd->cursor;
        __shared__ float s_tbl[TBL_MAX];
        __shared__ float4 s_ring_all[WARPS_PER_CTA][16];     // hand-off ring (row above), one per warp
        // per warp: 5-letter profile-profile column records (two 64-column rings, 2 KB) or the staged
        // score vectors of a 5-letter sequence / profile-sequence strip ([K rows][5 letters][32 lanes] floats)
        constexpr int REC_F4 = (BONUS == BONUS_NONE) ? 320 : 160;     // K = 8 rows per lane without a bonus
        __shared__ float4 s_rec_all[WARPS_PER_CTA][REC_F4];
        for (int i = threadIdx.x; i < TBL_MAX; i += blockDim.x) {
                s_tbl[i] = tbl[i];
        }
        // one mbarrier per block of a warp's record ring (bulk-copy completion, see sweep_strip)
        __shared__ unsigned long long s_mbar_all[WARPS_PER_CTA][4];
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                        const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_mbar_all[threadIdx.x >> 5][b]);
                        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
                }
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        unsigned long long* s_mbar = s_mbar_all[threadIdx.x >> 5];
        unsigned rec_phase = 0;          // phase parity of the warp's four mbarriers
        const int lane = threadIdx.x & 31;
        // sparse bonus lists of the rows of the strip a warp is sweeping (bonus kernel family only)
        // (dynamic shared memory: with it the CTA exceeds the 48 KB static limit)
        extern __shared__ int2 s_bon_all[];
        float4* s_ring = s_ring_all[threadIdx.x >> 5];
        float4* s_rec = s_rec_all[threadIdx.x >> 5];
        int2* s_bon = (BONUS == BONUS_SPARSE) ? (s_bon_all + (threadIdx.x >> 5) * (BON_QSLOTS * 32)) : s_bon_all;
        if constexpr (BONUS == BONUS_NONE) {
                // rounds of small sequence-sequence boxes: sub-warp groups (sweep_group above).  Every box
                // of such a round is a single strip (maxrows <= 128 < rows per strip), so a unit IS a sweep.
                const unsigned maxrows = rnd->maxrows;
                if (!thin && rnd->kinds == (1u << KB200_KIND_SS) && maxrows <= 128u && maxrows >= 1u) {
                        const int W = (maxrows <= 64u) ? 8 : 16;
                        const unsigned G = 32u / (unsigned)W;
                        while (true) {
                                unsigned base = 0;
                                if (lane == 0) {
                                        base = atomicAdd(cursor, G);
                                }
                                base = __shfl_sync(FULL, base, 0);
                                if (base >= total) {
                                        break;
                                }
                                const unsigned unit = base + (unsigned)(lane / W);
                                const bool valid = unit < total;
                                // rows of my group's sweep -> rows per lane the warp needs
                                int R = 0;
                                if (valid) {
                                        const KbUnit un = units[unit];
                                        const KbBox bx = boxes[un.item >> 1];
                                        const int mid = (bx.ea - bx.sa) / 2 + bx.sa;
                                        R = (un.item & 1) ? (bx.ea - mid) : (mid - bx.sa);
                                }
                                const int kneed = (int)__reduce_max_sync(FULL, (unsigned)((R + W - 1) / W));
                                const unsigned tag = tag_base + 1u;
                                if (W == 8) {
                                        if (tstride == 5) sweep_group_k<V_SS5, 8>(kneed, jobs, boxes, units, unit, valid, tag, s_tbl, tstride, lane, s_rec);
                                        else sweep_group_k<V_SS, 8>(kneed, jobs, boxes, units, unit, valid, tag, s_tbl, tstride, lane, s_rec);
                                } else {
                                        if (tstride == 5) sweep_group_k<V_SS5, 16>(kneed, jobs, boxes, units, unit, valid, tag, s_tbl, tstride, lane, s_rec);
                                        else sweep_group_k<V_SS, 16>(kneed, jobs, boxes, units, unit, valid, tag, s_tbl, tstride, lane, s_rec);
                                }
                        }
                        return;
                }
        }
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
                                sweep_unit<V_SS5, BONUS>(J, bx, bwd, un.strip, thin, ps, s_tbl, tstride, lane, s_ring, s_rec, s_bon, s_mbar, rec_phase);
                        } else {
                                sweep_unit<V_SS, BONUS>(J, bx, bwd, un.strip, thin, ps, s_tbl, tstride, lane, s_ring, s_rec, s_bon, s_mbar, rec_phase);
                        }
                } else if (J.kind == KB200_KIND_SP) {
                        if (J.nalpha <= 5) {
                                sweep_unit<V_SP5, BONUS>(J, bx, bwd, un.strip, thin, ps, s_tbl, tstride, lane, s_ring, s_rec, s_bon, s_mbar, rec_phase);
                        } else {
                                sweep_unit<V_SP, BONUS>(J, bx, bwd, un.strip, thin, ps, s_tbl, tstride, lane, s_ring, s_rec, s_bon, s_mbar, rec_phase);
                        }
                } else if (J.nalpha <= 5) {
                        sweep_unit<V_PP5, BONUS>(J, bx, bwd, un.strip, thin, ps, s_tbl, tstride, lane, s_ring, s_rec, s_bon, s_mbar, rec_phase);
                } else {
                        sweep_unit<V_PP23, BONUS>(J, bx, bwd, un.strip, thin, ps, s_tbl, tstride, lane, s_ring, s_rec, s_bon, s_mbar, rec_phase);
                }
        }
}


} // namespace

// launchers, one per kernel family / translation unit (units: KbUnit array written by kb_plan_kernel)
cudaError_t kb_sweep_launch_none(int grid, int block, cudaStream_t st, const KbJob* jobs, const KbBox* boxes, const void* units,
                                 KbRound* rnd, unsigned tag_base, const float* tbl, int tstride);
cudaError_t kb_sweep_launch_sparse(int grid, int block, cudaStream_t st, const KbJob* jobs, const KbBox* boxes, const void* units,
                                   KbRound* rnd, unsigned tag_base, const float* tbl, int tstride);
cudaError_t kb_sweep_launch_dense(int grid, int block, cudaStream_t st, const KbJob* jobs, const KbBox* boxes, const void* units,
                                  KbRound* rnd, unsigned tag_base, const float* tbl, int tstride);
