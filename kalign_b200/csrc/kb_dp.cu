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
// (a>b?a:b) because no NaN can occur and the sign of a zero never reaches a comparison).
//
// Parallel shape: one warp sweeps one (box, direction).  Lane l owns K consecutive rows of a strip
// of 32*K rows and walks the columns with a skew of one column per lane (anti-diagonal wavefront);
// the in-diagonal hand-off of the row above is a warp shuffle; the row that leaves a strip
// (H/E/F = a/ga/gb) is streamed through a per-job row buffer in global memory (float4 per column)
// and read back by the next strip / the meet-up kernel.  Boxes of one Hirschberg depth of ALL jobs
// of a batch are processed by one sweep launch and one meet-up launch (level-synchronous
// work-list); the meet-up emits the child boxes of aln_continue into the next work-list.
#include "kb_common.cuh"

#include <algorithm>

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int WARPS_PER_CTA = 4;
constexpr int TBL_STRIDE = 32;

struct Trip {
        float a, ga, gb;
};

__device__ __forceinline__ float kmax(float a, float b) { return fmaxf(a, b); }

enum { MODE_FIRST = 0, MODE_MID = 1, MODE_LAST = 2 };

// ---------------------------------------------------------------------------------------------
// per-strip row context
template <int KIND, int K> struct RowCtx {
        // SS: table row offsets; SP/PP: profile column pointers and gap terms
        int rbase[K];           // SS: residue * TBL_STRIDE
        const float* prow[K];   // SP/PP: profile column of the row
        float RO[K], RE[K], RT[K], ROp[K];
        int irow[K];            // absolute DP row index (bonus)
};

template <int KIND, int K, bool TAIL, int MODE, bool BONUS>
__device__ __forceinline__ void cells(const KbJob& J, const RowCtx<KIND, K>& rc, unsigned vmask,
                                      bool first_term, bool last_term,
                                      // column context
                                      int cres, const float* __restrict__ q, float CO, float CE, float COp, int jcol,
                                      const float* __restrict__ s_tbl,
                                      float (&sA)[K], float (&sGA)[K], float (&sGB)[K],
                                      Trip d, Trip& u /* in: up at column u; out: bottom row */)
{
#pragma unroll
        for (int k = 0; k < K; k++) {
                const float oA = sA[k], oGA = sGA[k], oGB = sGB[k];
                float RO, RE, RT, ROp;
                if constexpr (KIND == KB200_KIND_SS) {
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
                        a = kmax(kmax(d.a, d.ga + COp), d.gb + ROp);
                        if constexpr (KIND == KB200_KIND_SS) {
                                const float x = s_tbl[rc.rbase[k] + cres] + J.nsoff;
                                a = a + x;
                        } else if constexpr (KIND == KB200_KIND_SP) {
                                a = a + __ldg(rc.prow[k] + 32 + cres);
                        } else {
                                const float* __restrict__ p = rc.prow[k];
                                for (int c = J.nalpha - 1; c >= 0; c--) {
                                        const float pr = __ldg(p + c);
                                        const float prod = __fmul_rn(pr, __ldg(q + 32 + c));
                                        a = __fadd_rn(a, prod);
                                }
                        }
                        if constexpr (BONUS) {
                                a = a + __ldg(J.bonus + (size_t)rc.irow[k] * (size_t)J.len_b + (size_t)jcol);
                        }
                        if constexpr (MODE == MODE_MID) {
                                ga = kmax(oGA + CE, oA + CO);
                                gb = kmax(u.gb + RE, u.a + RO);
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
}

// One strip of 32*K rows starting at logical row `row0` of the sweep.
template <int KIND, int K, bool TAIL, bool BONUS>
__device__ void sweep_strip(const KbJob& J, const int bwd, const int sb, const int eb,
                            const int r0, const int r1, const int row0,
                            const bool first_term, const bool last_term,
                            const Trip in, float4* __restrict__ rowbuf,
                            const float* __restrict__ s_tbl, const int lane)
{
        const int C = eb - sb;
        const int R = r1 - r0;
        RowCtx<KIND, K> rc;
        unsigned vmask = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
                int g = row0 + lane * K + k;
                const bool valid = g < R;
                if (valid) {
                        vmask |= (1u << k);
                }
                if (!valid) {
                        g = (R > 0) ? (R - 1) : 0;
                }
                int i = bwd ? (r1 - 1 - g) : (r0 + g);
                if (R == 0) {
                        i = r0;      // never used for arithmetic that survives (pass-through rows)
                        if (i >= J.len_a) i = J.len_a - 1;
                        if (i < 0) i = 0;
                }
                rc.irow[k] = i;
                if constexpr (KIND == KB200_KIND_SS) {
                        rc.rbase[k] = (int)J.seq_r[i] * TBL_STRIDE;
                } else {
                        const float* p = J.prof_r + ((size_t)(i + 1) << 6);
                        const float* pp = bwd ? (p + 64) : (p - 64);
                        rc.prow[k] = p;
                        rc.RO[k] = __ldg(p + 27);
                        rc.RE[k] = __ldg(p + 28);
                        rc.RT[k] = __ldg(p + 29);
                        rc.ROp[k] = __ldg(pp + 27);
                }
        }
        float sA[K], sGA[K], sGB[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
                sA[k] = KB_NEGF; sGA[k] = KB_NEGF; sGB[k] = KB_NEGF;
        }
        Trip d = {KB_NEGF, KB_NEGF, KB_NEGF};
        Trip bot = {KB_NEGF, KB_NEGF, KB_NEGF};
        // init-row generator (strip 0, lane 0 only): previous column's (a, ga)
        float genA = in.a, genGA = in.ga;
        float prevCO = 0.0f;   // PP: [27] of the column visited one step earlier
        const bool gen = (row0 == 0);
        float4 pre = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!gen && lane == 0) {
                pre = __ldcg(rowbuf);
        }
        const int steps = C + 32;
        for (int t = 0; t < steps; t++) {
                const int u = t - lane;
                Trip up;
                up.a = __shfl_up_sync(FULL, bot.a, 1);
                up.ga = __shfl_up_sync(FULL, bot.ga, 1);
                up.gb = __shfl_up_sync(FULL, bot.gb, 1);
                const bool act = (u >= 0) && (u <= C);
                if (act) {
                        // ---- column context ----
                        const int j = bwd ? (eb - u) : (sb + u);        // state column
                        const int r = bwd ? j : (j - 1);                // residue / profile index
                        int cres = 0;
                        const float* q = nullptr;
                        float CO, CE, CT, COp;
                        if constexpr (KIND == KB200_KIND_PP) {
                                // profile column of state column j: r+1; for u==0 this is the
                                // boundary column visited "before" u==1 (only its [27] is used)
                                q = J.prof_c + ((size_t)(r + 1) << 6);
                                CO = __ldg(q + 27);
                                CE = __ldg(q + 28);
                                CT = first_term ? __ldg(q + 29) : 0.0f;
                                COp = prevCO;
                                prevCO = CO;
                        } else {
                                CO = J.o; CE = J.e; CT = J.t; COp = J.o;
                                if (u >= 1) {
                                        cres = (int)__ldg(J.seq_c + r);
                                }
                        }
                        // ---- lane 0: take the row above from the source ----
                        if (lane == 0) {
                                if (gen) {
                                        if (u == 0) {
                                                up = in;
                                        } else if (u < C) {
                                                float nga;
                                                if (first_term) {
                                                        nga = kmax(genGA, genA) + CT;
                                                } else {
                                                        nga = kmax(genGA + CE, genA + CO);
                                                }
                                                up.a = KB_NEGF; up.ga = nga; up.gb = KB_NEGF;
                                                genA = KB_NEGF; genGA = nga;
                                        } else {
                                                up.a = KB_NEGF; up.ga = KB_NEGF; up.gb = KB_NEGF;
                                        }
                                } else {
                                        up.a = pre.x; up.ga = pre.y; up.gb = pre.z;
                                        if (u < C) {
                                                pre = __ldcg(rowbuf + u + 1);
                                        }
                                }
                        }
                        const Trip got = up;
                        if (u == 0) {
                                cells<KIND, K, TAIL, MODE_FIRST, BONUS>(J, rc, vmask, first_term, last_term, cres, q, CO, CE, COp, j,
                                                                        s_tbl, sA, sGA, sGB, d, up);
                        } else if (u < C) {
                                cells<KIND, K, TAIL, MODE_MID, BONUS>(J, rc, vmask, first_term, last_term, cres, q, CO, CE, COp, j,
                                                                      s_tbl, sA, sGA, sGB, d, up);
                        } else {
                                cells<KIND, K, TAIL, MODE_LAST, BONUS>(J, rc, vmask, first_term, last_term, cres, q, CO, CE, COp, j,
                                                                       s_tbl, sA, sGA, sGB, d, up);
                        }
                        d = got;
                        bot = up;
                }
                if (lane == 31) {
                        const int uo = t - 31;
                        if (uo >= 0 && uo <= C) {
                                rowbuf[uo] = make_float4(bot.a, bot.ga, bot.gb, 0.0f);
                        }
                }
        }
        __syncwarp();
}

template <int KIND, bool BONUS>
__device__ void sweep_box(const KbJob& J, const KbBox& bx, const int bwd, const float* __restrict__ s_tbl, const int lane)
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
        int row0 = 0;
        do {
                const int rem = R - row0;
                if (rem >= 128) {
                        sweep_strip<KIND, 4, false, BONUS>(J, bwd, sb, eb, r0, r1, row0, first_term, last_term, in, rowbuf, s_tbl, lane);
                        row0 += 128;
                } else if (rem > 32) {
                        sweep_strip<KIND, 4, true, BONUS>(J, bwd, sb, eb, r0, r1, row0, first_term, last_term, in, rowbuf, s_tbl, lane);
                        row0 += 128;
                } else {
                        sweep_strip<KIND, 1, true, BONUS>(J, bwd, sb, eb, r0, r1, row0, first_term, last_term, in, rowbuf, s_tbl, lane);
                        row0 += 32;
                }
        } while (row0 < R);
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
kb_sweep_kernel(const KbJob* __restrict__ jobs, const KbBox* __restrict__ boxes, const int nboxes,
                unsigned int* __restrict__ cursor, const float* __restrict__ tbl)
{
        __shared__ float s_tbl[23 * TBL_STRIDE];
        for (int i = threadIdx.x; i < 23 * TBL_STRIDE; i += blockDim.x) {
                s_tbl[i] = tbl[i];
        }
        __syncthreads();
        const int lane = threadIdx.x & 31;
        const unsigned total = 2u * (unsigned)nboxes;
        while (true) {
                unsigned item = 0;
                if (lane == 0) {
                        item = atomicAdd(cursor, 1u);
                }
                item = __shfl_sync(FULL, item, 0);
                if (item >= total) {
                        break;
                }
                const KbBox bx = boxes[item >> 1];
                const int bwd = (int)(item & 1u);
                const KbJob J = jobs[bx.job];
                const bool bonus = (J.bonus != nullptr);
                if (J.kind == KB200_KIND_SS) {
                        if (bonus) sweep_box<KB200_KIND_SS, true>(J, bx, bwd, s_tbl, lane);
                        else sweep_box<KB200_KIND_SS, false>(J, bx, bwd, s_tbl, lane);
                } else if (J.kind == KB200_KIND_SP) {
                        if (bonus) sweep_box<KB200_KIND_SP, true>(J, bx, bwd, s_tbl, lane);
                        else sweep_box<KB200_KIND_SP, false>(J, bx, bwd, s_tbl, lane);
                } else {
                        if (bonus) sweep_box<KB200_KIND_PP, true>(J, bx, bwd, s_tbl, lane);
                        else sweep_box<KB200_KIND_PP, false>(J, bx, bwd, s_tbl, lane);
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
                        if (J.bonus) {
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
                                const int nchild = (hasL ? 1 : 0) + (hasR ? 1 : 0);
                                if (nchild) {
                                        unsigned slot = atomicAdd(next_count, (unsigned)nchild);
                                        if (hasL) {
                                                put_child(next, (int)slot, bx.job, bx.depth + 1, lsa, lea, lsb, leb, fin, lb0);
                                                slot++;
                                        }
                                        if (hasR) {
                                                put_child(next, (int)slot, bx.job, bx.depth + 1, rsa, rea, rsb, reb, rf0, bin);
                                        }
                                }
                        }
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
        // row buffers
        size_t total_cols = 0;
        size_t box_cap = 0;
        for (int i = 0; i < n; i++) {
                total_cols += (size_t)(jobs[i].len_a + jobs[i].len_b + 2);
                box_cap += (size_t)std::max(1, jobs[i].len_a);
        }
        KB_RUN(ctx->d_rows.ensure(2 * total_cols * sizeof(float4)));
        {
                float4* base = ctx->d_rows.as<float4>();
                size_t off = 0;
                for (int i = 0; i < n; i++) {
                        const size_t w = (size_t)(jobs[i].len_a + jobs[i].len_b + 2);
                        jobs[i].rowF = base + off;
                        jobs[i].rowB = base + total_cols + off;
                        off += w;
                }
        }
        KB_RUN(ctx->d_jobs.ensure(sizeof(KbJob) * (size_t)n));
        KB_CUDA(cudaMemcpyAsync(ctx->d_jobs.p, jobs.data(), sizeof(KbJob) * (size_t)n, cudaMemcpyHostToDevice, st));
        KB_RUN(ctx->d_boxA.ensure(sizeof(KbBox) * box_cap));
        KB_RUN(ctx->d_boxB.ensure(sizeof(KbBox) * box_cap));
        KB_RUN(ctx->d_counters.ensure(64));
        KB_RUN(ctx->d_tbl.ensure(sizeof(float) * 23 * TBL_STRIDE));
        {
                std::vector<float> tbl(23 * TBL_STRIDE, 0.0f);
                for (int i = 0; i < 23; i++) {
                        for (int j = 0; j < 23; j++) {
                                tbl[i * TBL_STRIDE + j] = subm_host[i * 23 + j];
                        }
                }
                KB_CUDA(cudaMemcpyAsync(ctx->d_tbl.p, tbl.data(), sizeof(float) * tbl.size(), cudaMemcpyHostToDevice, st));
                KB_CUDA(cudaStreamSynchronize(st));   // tbl is a stack-lifetime vector
        }
        {
                std::vector<KbBox> init;
                init.reserve(n);
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
                }
                if (init.empty()) {
                        return KB200_OK;
                }
                KB_CUDA(cudaMemcpyAsync(ctx->d_boxA.p, init.data(), sizeof(KbBox) * init.size(), cudaMemcpyHostToDevice, st));
                KB_CUDA(cudaStreamSynchronize(st));
                box_cap = std::max(box_cap, init.size());
                // counters layout: [0] sweep cursor (u32), [1] next count (u32), [2..3] cells (u64)
                KB_CUDA(cudaMemsetAsync(ctx->d_counters.p, 0, 64, st));
                unsigned count = (unsigned)init.size();
                KbBox* cur = ctx->d_boxA.as<KbBox>();
                KbBox* nxt = ctx->d_boxB.as<KbBox>();
                unsigned int* d_cursor = ctx->d_counters.as<unsigned int>();
                unsigned int* d_next = d_cursor + 1;
                unsigned long long* d_cells = (unsigned long long*)(d_cursor + 2);   // [ss, sp, pp, bonus]
                float sweep_ms = 0.0f;
                KB_CUDA(cudaEventRecord(ctx->ev0, st));
                while (count > 0) {
                        KB_CUDA(cudaMemsetAsync(d_cursor, 0, 8, st));
                        const unsigned items = 2u * count;
                        int grid = (int)std::min<unsigned>((items + WARPS_PER_CTA - 1) / WARPS_PER_CTA,
                                                           (unsigned)(ctx->sm_count * 12));
                        KB_CUDA(cudaEventRecord(ctx->ev2, st));
                        kb_sweep_kernel<<<grid, WARPS_PER_CTA * 32, 0, st>>>(ctx->d_jobs.as<KbJob>(), cur, (int)count, d_cursor,
                                                                            ctx->d_tbl.as<float>());
                        KB_CUDA(cudaEventRecord(ctx->ev3, st));
                        int mgrid = (int)std::min<unsigned>((count + 3) / 4, (unsigned)(ctx->sm_count * 16));
                        kb_meetup_kernel<<<mgrid, 128, 0, st>>>(ctx->d_jobs.as<KbJob>(), cur, (int)count, nxt, d_next, d_cells);
                        KB_CUDA(cudaGetLastError());
                        unsigned next_count = 0;
                        KB_CUDA(cudaMemcpyAsync(&next_count, d_next, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
                        KB_CUDA(cudaStreamSynchronize(st));
                        float ms = 0.0f;
                        cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3);
                        sweep_ms += ms;
                        ctx->stats.n_boxes += count;
                        ctx->stats.n_launches += 2;
                        if ((size_t)next_count > box_cap) {
                                fprintf(stderr, "[kalign_b200] box list overflow (%u > %zu)\n", next_count, box_cap);
                                return KB200_FAIL;
                        }
                        count = next_count;
                        std::swap(cur, nxt);
                }
                KB_CUDA(cudaEventRecord(ctx->ev1, st));
                unsigned long long cells[4] = {0, 0, 0, 0};
                KB_CUDA(cudaMemcpyAsync(cells, d_cells, sizeof(cells), cudaMemcpyDeviceToHost, st));
                KB_CUDA(cudaStreamSynchronize(st));
                float ms = 0.0f;
                cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
                ctx->stats.dp_seconds += 1e-3 * (double)ms;
                ctx->stats.sweep_seconds += 1e-3 * (double)sweep_ms;
                ctx->stats.dp_cells += (double)cells[0] + (double)cells[1] + (double)cells[2];
                ctx->stats.cells_ss += (double)cells[0];
                ctx->stats.cells_sp += (double)cells[1];
                ctx->stats.cells_pp += (double)cells[2];
                ctx->stats.cells_bonus += (double)cells[3];
        }
        return KB200_OK;
}
