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
#include "kb_sweep.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// plan: cut every (box, direction) into strips, reserve a contiguous, ordered unit range
__global__ void kb_plan_kernel(const KbJob* __restrict__ jobs, const KbBox* __restrict__ boxes, KbRound* __restrict__ rnd,
                               const unsigned long long rows_total, const unsigned resident_warps, const int force_thin, const int thin_k,
                               const int batch_bonus, KbUnit* __restrict__ units, const unsigned unit_cap,
                               KbDevStats* __restrict__ dstats)
{
        const unsigned nboxes = rnd->nboxes;
        if (nboxes == 0u) {
                return;
        }
        // thin strips when thick ones could not occupy the machine: the rows still alive at this
        // depth are at most rows_total, spread over nboxes boxes (every thread computes the same value)
        const unsigned long long thick_units = rows_total / 128ull + 2ull * (unsigned long long)nboxes;
        int thin = (thick_units < 2ull * (unsigned long long)resident_warps) ? thin_k : 0;
        if (force_thin >= 0) thin = force_thin ? thin_k : 0;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
                rnd->thin = (unsigned)thin;
                atomicAdd(&dstats->nboxes, (unsigned long long)nboxes);
        }
        const int lane = threadIdx.x & 31;
        const long long items = 2LL * (long long)nboxes;
        const long long nth = (long long)gridDim.x * blockDim.x;
        for (long long base = (long long)blockIdx.x * blockDim.x + (threadIdx.x - lane); base < items; base += nth) {
                const long long item = base + lane;
                int nstr = 0;
                int myrows = 0;
                unsigned mykind = 0u;
                if (item < items) {
                        const KbBox bx = boxes[item >> 1];
                        const int bwd = (int)(item & 1);
                        const int mid = (bx.ea - bx.sa) / 2 + bx.sa;
                        const int R = bwd ? (bx.ea - mid) : (mid - bx.sa);
                        const int rps = rows_per_strip(jobs[bx.job].kind, jobs[bx.job].nalpha, thin, batch_bonus != 0);
                        nstr = (R + rps - 1) / rps;
                        if (nstr < 1) nstr = 1;
                        myrows = R;
                        mykind = 1u << jobs[bx.job].kind;
                }
                // what the sweep kernel needs to choose sub-warp groups for a round of small boxes
                {
                        const unsigned wr = __reduce_max_sync(FULL, (unsigned)myrows);
                        const unsigned wk = __reduce_or_sync(FULL, mykind);
                        if (lane == 0) {
                                atomicMax(&rnd->maxrows, wr);
                                atomicOr(&rnd->kinds, wk);
                        }
                }
                // warp-inclusive scan of nstr
                int incl = nstr;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(FULL, incl, o);
                        if (lane >= o) incl += v;
                }
                const int warp_total = __shfl_sync(FULL, incl, 31);
                unsigned ubase = 0;
                if (lane == 31 && warp_total > 0) {
                        ubase = atomicAdd(&rnd->nunits, (unsigned)warp_total);
                }
                ubase = __shfl_sync(FULL, ubase, 31);
                if ((unsigned long long)ubase + (unsigned long long)warp_total > (unsigned long long)unit_cap) {
                        if (lane == 0) atomicOr(&dstats->flags, (unsigned)KB_FLAG_UNIT_OVERFLOW);
                        continue;       // the sweep leaves these boxes alone; the call is reported as failed
                }
                const unsigned mine = ubase + (unsigned)(incl - nstr);
                for (int s2 = 0; s2 < nstr; s2++) {
                        KbUnit un;
                        un.item = (int)item;
                        un.strip = s2;
                        units[mine + s2] = un;
                }
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

__device__ __forceinline__ void put_child(KbBox* __restrict__ next, int slot, int job, unsigned hid,
                                          int sa, int ea, int sb, int eb, Trip f0, Trip b0)
{
        KbBox c;
        c.job = job; c.sa = sa; c.ea = ea; c.sb = sb; c.eb = eb;
        c.f0a = f0.a; c.f0ga = f0.ga; c.f0gb = f0.gb;
        c.b0a = b0.a; c.b0ga = b0.ga; c.b0gb = b0.gb;
        c.hid = hid;
        next[slot] = c;
}

__global__ void __launch_bounds__(128)
kb_meetup_kernel(const KbJob* __restrict__ jobs, const KbBox* __restrict__ boxes, const KbRound* __restrict__ rnd,
                 KbBox* __restrict__ next, unsigned int* __restrict__ next_count,
                 KbBox* __restrict__ small, unsigned int* __restrict__ small_count, const int small_rows, const int small_cols,
                 const unsigned box_cap, KbDevStats* __restrict__ dstats)
{
        const int nboxes = (int)rnd->nboxes;
        if (nboxes == 0) {
                return;
        }
        unsigned long long* const cells = dstats->cells;
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
                        if (bx.hid == 1u && J.score) {
                                *J.score = m.max;
                        }
                        if (J.margins) {
                                if (bx.hid < J.margin_cap) {
                                        J.margins[bx.hid] = (m.max2 > KB_NEGF) ? (m.max - m.max2) : -1.0f;
                                } else {
                                        atomicOr(&dstats->flags, (unsigned)KB_FLAG_MARGIN);
                                }
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
                                        if (slot + (unsigned)nbig > box_cap) {
                                                atomicOr(&dstats->flags, (unsigned)KB_FLAG_BOX_OVERFLOW);
                                                continue;
                                        }
                                        if (hasL && !smL) {
                                                put_child(next, (int)slot, bx.job, 2u * bx.hid, lsa, lea, lsb, leb, fin, lb0);
                                                slot++;
                                        }
                                        if (hasR && !smR) {
                                                put_child(next, (int)slot, bx.job, 2u * bx.hid + 1u, rsa, rea, rsb, reb, rf0, bin);
                                        }
                                }
                                if (nsml) {
                                        unsigned slot = atomicAdd(small_count, (unsigned)nsml);
                                        if (slot + (unsigned)nsml > box_cap) {
                                                atomicOr(&dstats->flags, (unsigned)KB_FLAG_SMALL_OVERFLOW);
                                                continue;
                                        }
                                        if (smL) {
                                                put_child(small, (int)slot, bx.job, 2u * bx.hid, lsa, lea, lsb, leb, fin, lb0);
                                                slot++;
                                        }
                                        if (smR) {
                                                put_child(small, (int)slot, bx.job, 2u * bx.hid + 1u, rsa, rea, rsb, reb, rf0, bin);
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
        unsigned hid;
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
                set_table_addr<V, K>(rc, s_tbl);
                float sA[K], sGA[K], sGB[K];
                int sp_i[K], sp_c[K];
                float sp_v[K], sp_wrap[K];
#pragma unroll
                for (int k = 0; k < K; k++) {
                        sA[k] = KB_NEGF; sGA[k] = KB_NEGF; sGB[k] = KB_NEGF;
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
                        cells<V, K, true, MODE_FIRST, BONUS, false>(J, rc, vmask, first_term, last_term, cc, dstep, sp_i, sp_c, sp_v, sp_wrap, nullptr, s_tbl, sA, sGA, sGB, d, up);
                        d = got;
                        S[0] = make_float4(up.a, up.ga, up.gb, 0.0f);
                }
                for (int u = 1; u < C; u++) {
                        column(u);
                        const float4 q = S[u];
                        Trip up = {q.x, q.y, q.z};
                        const Trip got = up;
                        cells<V, K, true, MODE_MID, BONUS, false>(J, rc, vmask, first_term, last_term, cc, dstep, sp_i, sp_c, sp_v, sp_wrap, nullptr, s_tbl, sA, sGA, sGB, d, up);
                        d = got;
                        S[u] = make_float4(up.a, up.ga, up.gb, 0.0f);
                }
                {
                        column(C);
                        const float4 q = S[C];
                        Trip up = {q.x, q.y, q.z};
                        cells<V, K, true, MODE_LAST, BONUS, false>(J, rc, vmask, first_term, last_term, cc, dstep, sp_i, sp_c, sp_v, sp_wrap, nullptr, s_tbl, sA, sGA, sGB, d, up);
                        S[C] = make_float4(up.a, up.ga, up.gb, 0.0f);
                }
        }
}

template <int V, int MAXC, int BONUS>
__device__ void small_box_run(const KbJob& J, const KbBox& root, const float* __restrict__ s_tbl, const int tstride,
                              unsigned long long& ncells, unsigned& err)
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
                b.hid = root.hid;
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
                if (J.margins) {
                        if (bx.hid < J.margin_cap) {
                                J.margins[bx.hid] = (m.max2 > KB_NEGF) ? (m.max - m.max2) : -1.0f;
                        } else {
                                err |= (unsigned)KB_FLAG_MARGIN;
                        }
                }
                if (m.key == 0x7fffffff) {
                        continue;
                }
                const int c = sb + (m.key >> 3);
                const int t = m.key & 7;
                SBox L, Rr;
                L.hid = 2u * bx.hid;
                Rr.hid = 2u * bx.hid + 1u;
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
                // a full stack is an error of the call (reported through KbDevStats::flags), never a
                // silently dropped box: the depth is bounded by log2(SMALL_ROWS_MAX) + 2 < SMALL_STACK
                if (Rr.sa < Rr.ea && Rr.sb < Rr.eb) {
                        if (sp < SMALL_STACK) stack[sp++] = Rr; else err |= (unsigned)KB_FLAG_STACK;
                }
                if (L.sa < L.ea && L.sb < L.eb) {
                        if (sp < SMALL_STACK) stack[sp++] = L; else err |= (unsigned)KB_FLAG_STACK;
                }
        }
}

// ONLY_SS: every job of the batch is sequence-sequence (the anchor batch, tree level 1): the kernel
// carries no profile code
template <int MAXC, int BONUS, bool ONLY_SS, int MINB>
__global__ void __launch_bounds__(128, MINB)
kb_small_kernel(const KbJob* __restrict__ jobs, const KbBox* __restrict__ boxes, const unsigned* __restrict__ nsmall_p,
                const unsigned box_cap, KbDevStats* __restrict__ dstats, const float* __restrict__ tbl, const int tstride)
{
        unsigned n = *nsmall_p;
        if (n == 0u) {
                return;
        }
        if (n > box_cap) n = box_cap;        // overflow was flagged by the meet-up
        __shared__ float s_tbl[TBL_MAX];
        for (int i = threadIdx.x; i < TBL_MAX; i += blockDim.x) {
                s_tbl[i] = tbl[i];
        }
        __syncthreads();
        unsigned long long* const cells = dstats->cells;
        const unsigned nth = gridDim.x * blockDim.x;
        for (unsigned b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += nth) {
                const KbBox bx = boxes[b];
                const KbJob J = jobs[bx.job];
                unsigned long long nc = 0;
                unsigned err = 0;
                if constexpr (ONLY_SS) {
                        small_box_run<V_SS, MAXC, BONUS>(J, bx, s_tbl, tstride, nc, err);
                } else {
                        if (J.kind == KB200_KIND_SS) small_box_run<V_SS, MAXC, BONUS>(J, bx, s_tbl, tstride, nc, err);
                        else if (J.kind == KB200_KIND_SP) small_box_run<V_SP, MAXC, BONUS>(J, bx, s_tbl, tstride, nc, err);
                        else if (J.nalpha <= 5) small_box_run<V_PP5, MAXC, BONUS>(J, bx, s_tbl, tstride, nc, err);
                        else small_box_run<V_PP23, MAXC, BONUS>(J, bx, s_tbl, tstride, nc, err);
                }
                atomicAdd(cells + 4 + J.kind, nc);
                if (J.bonus || J.bkey) {
                        atomicAdd(cells + 3, nc);
                }
                if (err) {
                        atomicOr(&dstats->flags, err);
                }
        }
}

static_assert(SMALL_STACK >= 7 + 3, "explicit DFS stack: log2(SMALL_ROWS_MAX) + 3 frames");

} // namespace

// ---------------------------------------------------------------------------------------------

// ---- pinned staging, timed spans, deferred statistics ------------------------------------------

void* KbPinned::get(size_t bytes)
{
        bytes = (bytes + 255) & ~(size_t)255;
        while (true) {
                if (cur < blocks.size()) {
                        if (used + bytes <= caps[cur]) {
                                void* r = (char*)blocks[cur] + used;
                                used += bytes;
                                return r;
                        }
                        cur++;
                        used = 0;
                        continue;
                }
                const size_t want = std::max<size_t>((size_t)16 << 20, bytes);
                void* p = nullptr;
                if (cudaMallocHost(&p, want) != cudaSuccess) {
                        cudaGetLastError();
                        fprintf(stderr, "[kalign_b200] cudaMallocHost(%zu) failed\n", want);
                        return nullptr;
                }
                blocks.push_back(p);
                caps.push_back(want);
        }
}

void KbPinned::release()
{
        for (void* p : blocks) {
                cudaFreeHost(p);
        }
        blocks.clear();
        caps.clear();
        cur = used = 0;
}

int kb_h2d(kb200_ctx* ctx, void* dst, const void* src, size_t bytes)
{
        if (bytes == 0) {
                return KB200_OK;
        }
        void* stage = ctx->pinned.get(bytes);
        if (!stage) {
                return KB200_FAIL;
        }
        memcpy(stage, src, bytes);
        KB_CUDA(cudaMemcpyAsync(dst, stage, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return KB200_OK;
}

int kb_span_begin(kb200_ctx* ctx, int kind)
{
        const size_t id = ctx->ev_used;
        if (ctx->ev_pool.size() < 2 * (id + 1)) {
                cudaEvent_t a = nullptr, b = nullptr;
                if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) {
                        cudaGetLastError();
                        return -1;
                }
                ctx->ev_pool.push_back(a);
                ctx->ev_pool.push_back(b);
                ctx->ev_kind.push_back(kind);
        }
        ctx->ev_kind[id] = kind;
        ctx->ev_used = id + 1;
        cudaEventRecord(ctx->ev_pool[2 * id], ctx->stream);
        return (int)id;
}

void kb_span_end(kb200_ctx* ctx, int span)
{
        if (span >= 0) {
                cudaEventRecord(ctx->ev_pool[2 * (size_t)span + 1], ctx->stream);
        }
}

static int ensure_dstats(kb200_ctx* ctx)
{
        if (!ctx->d_stats.p) {
                KB_RUN(ctx->d_stats.ensure(sizeof(KbDevStats)));
                KB_CUDA(cudaMemsetAsync(ctx->d_stats.p, 0, sizeof(KbDevStats), ctx->stream));
        }
        return KB200_OK;
}

int kb_collect(kb200_ctx* ctx)
{
        cudaStream_t st = ctx->stream;
        KbDevStats hs;
        memset(&hs, 0, sizeof(hs));
        if (ctx->d_stats.p) {
                KB_CUDA(cudaMemcpyAsync(&hs, ctx->d_stats.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
                KB_CUDA(cudaMemsetAsync(ctx->d_stats.p, 0, sizeof(KbDevStats), st));
        }
        KB_CUDA(cudaStreamSynchronize(st));
        ctx->pinned.reset();     // every staged copy has completed
        for (int k = 0; k < 3; k++) {
                ctx->stats.dp_cells += (double)hs.cells[k] + (double)hs.cells[4 + k];
        }
        ctx->stats.cells_ss += (double)hs.cells[0] + (double)hs.cells[4];
        ctx->stats.cells_sp += (double)hs.cells[1] + (double)hs.cells[5];
        ctx->stats.cells_pp += (double)hs.cells[2] + (double)hs.cells[6];
        ctx->stats.cells_bonus += (double)hs.cells[3];
        ctx->stats.small_ss += (double)hs.cells[4];
        ctx->stats.small_sp += (double)hs.cells[5];
        ctx->stats.small_pp += (double)hs.cells[6];
        ctx->stats.n_boxes += (long long)hs.nboxes;
        for (size_t i = 0; i < ctx->ev_used; i++) {
                float ms = 0.0f;
                if (cudaEventElapsedTime(&ms, ctx->ev_pool[2 * i], ctx->ev_pool[2 * i + 1]) != cudaSuccess) {
                        cudaGetLastError();
                        continue;
                }
                const double sec = 1e-3 * (double)ms;
                if (ctx->ev_kind[i] == KB_SPAN_SWEEP) ctx->stats.sweep_seconds += sec;
                else if (ctx->ev_kind[i] == KB_SPAN_DP) ctx->stats.dp_seconds += sec;
                else ctx->stats.small_seconds += sec;
        }
        ctx->ev_used = 0;
        if (hs.flags) {
                fprintf(stderr, "[kalign_b200] engine error flags 0x%x:%s%s%s%s%s%s\n", hs.flags,
                        (hs.flags & KB_FLAG_BOX_OVERFLOW) ? " box work-list overflow" : "",
                        (hs.flags & KB_FLAG_UNIT_OVERFLOW) ? " unit list overflow" : "",
                        (hs.flags & KB_FLAG_SMALL_OVERFLOW) ? " small-box list overflow" : "",
                        (hs.flags & KB_FLAG_ROUNDS) ? " boxes left after the last round" : "",
                        (hs.flags & KB_FLAG_STACK) ? " small-box recursion stack overflow" : "",
                        (hs.flags & KB_FLAG_MARGIN) ? " recursion deeper than the margin array" : "");
                return KB200_FAIL;
        }
        return KB200_OK;
}

namespace {
// task->confidence: the reference adds the margin of every meet-up to margin_sum in the order its
// recursion visits the boxes -- the box, then everything below its upper-left child, then
// everything below the lower-right child (aln_runner / aln_continue, aln_controller.c:21,194).
// A float sum is not associative, so one thread per job replays exactly that order over the
// job's margin array (explicit stack; the tree is at most ~20 deep).
__global__ void kb_confidence_kernel(const KbJob* __restrict__ jobs, const int njobs, float* __restrict__ out)
{
        const int j = blockIdx.x * blockDim.x + threadIdx.x;
        if (j >= njobs) {
                return;
        }
        const float* __restrict__ mg = jobs[j].margins;
        const unsigned cap = jobs[j].margin_cap;
        if (!mg) {
                out[j] = 0.0f;
                return;
        }
        unsigned stack[40];
        int sp = 0;
        stack[sp++] = 1u;
        float sum = 0.0f;
        int count = 0;
        while (sp > 0) {
                const unsigned id = stack[--sp];
                if (id >= cap) continue;
                const float v = mg[id];
                if (__float_as_uint(v) == 0xffffffffu) continue;      // no such box
                if (v >= 0.0f) {
                        sum = __fadd_rn(sum, v);
                        count++;
                }
                if (sp + 2 <= 40) {
                        stack[sp++] = 2u * id + 1u;
                        stack[sp++] = 2u * id;
                }
        }
        out[j] = (count > 0) ? __fdiv_rn(sum, (float)count) : 0.0f;
}

// boxes that survive the last enqueued round would be lost: flag them
__global__ void kb_rounds_check_kernel(const KbRound* __restrict__ last, KbDevStats* __restrict__ dstats)
{
        if (last->nboxes != 0u) {
                atomicOr(&dstats->flags, (unsigned)KB_FLAG_ROUNDS);
        }
}
} // namespace

unsigned kb_margin_cap(int len_a)
{
        unsigned cap = 16;
        while (cap < 8u * (unsigned)(len_a + 1)) cap <<= 1;
        return cap;
}

int kb_confidences(kb200_ctx* ctx, int njobs, float* d_conf_out)
{
        if (njobs <= 0) return KB200_OK;
        kb_confidence_kernel<<<(njobs + 63) / 64, 64, 0, ctx->stream>>>(ctx->d_jobs.as<KbJob>(), njobs, d_conf_out);
        KB_CUDA(cudaGetLastError());
        ctx->stats.n_launches++;
        return KB200_OK;
}

int kb_run_hirschberg(kb200_ctx* ctx, const float* subm_host, std::vector<KbJob>& jobs)
{
        const int n = (int)jobs.size();
        if (n == 0) {
                return KB200_OK;
        }
        cudaStream_t st = ctx->stream;
        KB_RUN(ensure_dstats(ctx));
        KbDevStats* d_stats = ctx->d_stats.as<KbDevStats>();
        // row buffers, packed column records
        size_t total_cols = 0;
        size_t box_cap = 0;
        size_t unit_cap = 0;
        size_t pack_floats = 0;
        std::vector<int> pp_jobs;
        std::vector<long long> pp_prefix;
        long long pp_cols = 0;
        int max_rows = 1;
        for (int i = 0; i < n; i++) {
                total_cols += (size_t)(jobs[i].len_a + jobs[i].len_b + 2);
                box_cap += (size_t)std::max(1, jobs[i].len_a);
                max_rows = std::max(max_rows, jobs[i].len_a);
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
        if (box_cap > 0x7fffffffull || unit_cap > 0xfffffff0ull) {
                fprintf(stderr, "[kalign_b200] batch too large for the 32-bit work lists (%zu rows)\n", box_cap);
                return KB200_FAIL;
        }
        {
                // Row buffers carry the hand-off tags.  Tags only grow within a context, so a buffer
                // that was cleared when it was allocated never holds a tag a later launch could
                // mistake for its own; it is cleared again only when it is re-allocated (fresh memory
                // may hold anything) or when the tag counter is about to wrap.
                const size_t need = 2 * total_cols * sizeof(float4);
                const void* before = ctx->d_rows.p;
                const size_t cap_before = ctx->d_rows.cap;
                KB_RUN(ctx->d_rows.ensure(need));
                const unsigned tags_needed = 65536u * (unsigned)(KB_MAX_ROUNDS + 2);
                const bool wrap = ctx->tag_counter > 0xfff00000u - tags_needed;
                if (wrap) {
                        ctx->tag_counter = 65536u;
                }
                if (ctx->d_rows.p != before || ctx->d_rows.cap != cap_before || wrap || ctx->rows_tag_floor == 0u) {
                        KB_CUDA(cudaMemsetAsync(ctx->d_rows.p, 0, ctx->d_rows.cap, st));
                        ctx->rows_tag_floor = 1u;
                }
        }
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
        KB_RUN(kb_h2d(ctx, ctx->d_jobs.p, jobs.data(), sizeof(KbJob) * (size_t)n));
        KB_RUN(ctx->d_boxA.ensure(sizeof(KbBox) * box_cap));
        KB_RUN(ctx->d_boxB.ensure(sizeof(KbBox) * box_cap));
        KB_RUN(ctx->d_boxS.ensure(sizeof(KbBox) * box_cap));
        KB_RUN(ctx->d_units.ensure(sizeof(KbUnit) * unit_cap));
        KB_RUN(ctx->d_counters.ensure(sizeof(KbRound) * (KB_MAX_ROUNDS + 2) + 64));
        KB_RUN(ctx->d_tbl.ensure(sizeof(float) * TBL_MAX));
        // shared-memory score table: row stride = alphabet size, so that a 5-letter table (25
        // entries) puts every entry in its own bank -- lanes reading different entries never conflict
        int max_alpha = 5;
        for (int i = 0; i < n; i++) {
                max_alpha = std::max(max_alpha, jobs[i].nalpha);
        }
        const int tstride = (max_alpha <= 5) ? 5 : 23;
        {
                float* tbl = (float*)ctx->pinned.get(sizeof(float) * TBL_MAX);
                if (!tbl) return KB200_FAIL;
                memset(tbl, 0, sizeof(float) * TBL_MAX);
                for (int i = 0; i < tstride; i++) {
                        for (int j = 0; j < tstride; j++) {
                                tbl[i * tstride + j] = subm_host[i * 23 + j];
                        }
                }
                KB_CUDA(cudaMemcpyAsync(ctx->d_tbl.p, tbl, sizeof(float) * TBL_MAX, cudaMemcpyHostToDevice, st));
        }
        if (!pp_jobs.empty()) {
                KB_RUN(ctx->d_ppidx.ensure(sizeof(int) * pp_jobs.size() + sizeof(long long) * pp_jobs.size() + 64));
                long long* d_pref = ctx->d_ppidx.as<long long>();
                int* d_idx = (int*)(d_pref + pp_jobs.size());
                KB_RUN(kb_h2d(ctx, d_pref, pp_prefix.data(), sizeof(long long) * pp_jobs.size()));
                KB_RUN(kb_h2d(ctx, d_idx, pp_jobs.data(), sizeof(int) * pp_jobs.size()));
                const int grid = (int)std::min<long long>((pp_cols + 255) / 256, (long long)ctx->sm_count * 16);
                kb_pack_kernel<<<grid, 256, 0, st>>>(ctx->d_jobs.as<KbJob>(), d_idx, (int)pp_jobs.size(), d_pref, pp_cols);
                KB_CUDA(cudaGetLastError());
                ctx->stats.n_launches += 1;
        }
        size_t rows_total = 0;
        unsigned ninit = 0;
        {
                KbBox* init = (KbBox*)ctx->pinned.get(sizeof(KbBox) * (size_t)n);
                if (!init) return KB200_FAIL;
                for (int i = 0; i < n; i++) {
                        if (jobs[i].len_a <= 0 || jobs[i].len_b <= 0) {
                                continue;
                        }
                        KbBox b;
                        b.job = i; b.sa = 0; b.ea = jobs[i].len_a; b.sb = 0; b.eb = jobs[i].len_b;
                        b.f0a = 0.0F; b.f0ga = KB_NEGF; b.f0gb = KB_NEGF;
                        b.b0a = 0.0F; b.b0ga = KB_NEGF; b.b0gb = KB_NEGF;
                        b.hid = 1u;
                        init[ninit++] = b;
                        rows_total += (size_t)jobs[i].len_a;
                }
                if (ninit == 0) {
                        return KB200_OK;
                }
                KB_CUDA(cudaMemcpyAsync(ctx->d_boxA.p, init, sizeof(KbBox) * (size_t)ninit, cudaMemcpyHostToDevice, st));
        }
        // per-round control words: [r] = {cursor, nunits, nboxes, thin}; then the small-box count
        KbRound* d_rounds = ctx->d_counters.as<KbRound>();
        unsigned* d_nsmall = (unsigned*)(d_rounds + KB_MAX_ROUNDS + 1);
        KB_CUDA(cudaMemsetAsync(ctx->d_counters.p, 0, sizeof(KbRound) * (KB_MAX_ROUNDS + 2), st));
        {
                unsigned* h = (unsigned*)ctx->pinned.get(sizeof(unsigned));
                if (!h) return KB200_FAIL;
                *h = ninit;
                KB_CUDA(cudaMemcpyAsync(&d_rounds[0].nboxes, h, sizeof(unsigned), cudaMemcpyHostToDevice, st));
        }
        KbBox* cur = ctx->d_boxA.as<KbBox>();
        KbBox* nxt = ctx->d_boxB.as<KbBox>();
        KbBox* d_small = ctx->d_boxS.as<KbBox>();
        const bool use_small = getenv("KB200_NO_SMALL") == nullptr;
        // thread-per-box threshold: with enough jobs to fill the machine the deep rounds (boxes of a
        // few dozen rows: 1-2 live lanes per warp in a sweep) are cheaper as serial per-thread work
        int small_rows = SMALL_ROWS, small_cols = SMALL_COLS;
        {
                // largest power of two T such that the boxes of T rows (about rows_total / T of them)
                // still give every resident thread of half the machine a box of its own
                const size_t want = (size_t)ctx->sm_count * 256;
                // 23-letter profile operands: a thread-per-box sweep re-reads 28-float column records and
                // keeps 23 residue counts per row, it runs 3-4x slower per cell than the warp strips
                // (C4 level 2: 0.06 vs 0.21 T cells/s): boxes of more than 64 rows stay with the strips (measured optimum, C4 257 -> 246 ms)
                int rows_cap = SMALL_ROWS_MAX;
                for (int i = 0; i < n; i++) {
                        if (jobs[i].kind != KB200_KIND_SS && jobs[i].nalpha > 5) { rows_cap = 64; break; }
                }
                if (const char* e = getenv("KB200_SMALL_ROWS_PROF")) rows_cap = std::min(std::max(atoi(e), SMALL_ROWS), SMALL_ROWS_MAX);
                // all-seq-seq batches without bonus (anchor batch, --fast level 1): rounds of small boxes run as
                // sub-warp groups (sweep_group), which beat the thread-per-box kernel down to ~64-row boxes
                {
                        bool all_ss_plain = true;
                        for (int i = 0; i < n; i++) {
                                if (jobs[i].kind != KB200_KIND_SS || jobs[i].bonus || jobs[i].bkey) { all_ss_plain = false; break; }
                        }
                        if (all_ss_plain) {
                                rows_cap = 64;        // measured on C3: 128 -> 390, 64 -> 379, 32 -> 382, 16 -> 391 ms per step
                                if (const char* e = getenv("KB200_SMALL_ROWS_SS")) rows_cap = std::min(std::max(atoi(e), SMALL_ROWS), SMALL_ROWS_MAX);
                        }
                }
                while (small_rows * 2 <= rows_cap && rows_total / (size_t)(small_rows * 2) >= want) {
                        small_rows *= 2;
                }
                if (small_rows > SMALL_ROWS) small_cols = std::min(2 * small_rows + small_rows / 2, SMALL_COLS_MAX);
        }
        // few boxes (top of the guide tree: a handful of big profile pairs per GPU): a single thread finishing
        // a 16 x 48 box is a 0.3 ms serial tail of the level; hand over only boxes of <= 8 rows there
        if (small_rows == SMALL_ROWS && rows_total / (size_t)SMALL_ROWS < (size_t)ctx->sm_count * 32 && getenv("KB200_THIN_SMALL_OFF") == nullptr) {
                small_rows = 8;
                small_cols = 24;
        }
        if (const char* e = getenv("KB200_SMALL_ROWS")) small_rows = std::min(std::max(atoi(e), 1), SMALL_ROWS_MAX);
        if (const char* e = getenv("KB200_SMALL_COLS")) small_cols = std::min(std::max(atoi(e), 4), SMALL_COLS_MAX);
        // KB200_THIN=0|1 forces thick / thin strips (tests run every batch both ways)
        int force_thin = -1;
        if (const char* e = getenv("KB200_THIN")) force_thin = (atoi(e) != 0) ? 1 : 0;
        // rows per lane of the thin regime.  A lone warp advances one wavefront step in ~320 cycles
        // whatever its row count (measured, profiles/r02_thin_k.txt), so two rows per lane halve the
        // number of strips -- and with it the pipeline-fill part of a sweep -- for free; four rows
        // start to cost per step.  KB200_THIN_K=1|2|4 overrides.
        int thin_k = 2;
        if (const char* e = getenv("KB200_THIN_K")) thin_k = std::min(std::max(atoi(e), 1), 4);
        if (thin_k == 3) thin_k = 2;
        const bool trace = getenv("KB200_TRACE") != nullptr;
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
        // resident warps of the persistent sweep grid
        const int sweep_ctas = ctx->sm_count * 4;
        const unsigned resident_warps = (unsigned)(sweep_ctas * WARPS_PER_CTA);
        // Rounds.  A box of R rows has children of at most R/2 + 1 rows and the recursion ends below
        // two rows, so ceil(log2(rows)) + 2 rounds always suffice.  Every round is enqueued without
        // looking at the device: a round whose box count turned out to be zero costs three empty
        // launches.  kb_rounds_check_kernel flags boxes that would be left over.
        int nrounds = 2;
        while ((1 << (nrounds - 2)) < max_rows && nrounds < KB_MAX_ROUNDS) nrounds++;
        nrounds = std::min(nrounds + 1, KB_MAX_ROUNDS);
        const int span_dp = kb_span_begin(ctx, KB_SPAN_DP);
        unsigned long long trace_cells = 0;      // KB200_TRACE: cells of the rounds so far (since the last kb_collect)
        if (trace) {
                KbDevStats hs;
                KB_CUDA(cudaMemcpyAsync(&hs, d_stats, sizeof(hs), cudaMemcpyDeviceToHost, st));
                KB_CUDA(cudaStreamSynchronize(st));
                trace_cells = hs.cells[0] + hs.cells[1] + hs.cells[2];
        }
        for (int round = 0; round < nrounds; round++) {
                KbRound* rnd = d_rounds + round;
                // boxes this round can hold at most: twice the previous round's, never more than the rows
                const unsigned long long bound = std::min<unsigned long long>((unsigned long long)box_cap,
                                                                              (unsigned long long)ninit << std::min(round, 40));
                // tags: unique per launch (8192 strips per sweep at most: 256k rows), never 0
                ctx->tag_counter += 65536u;
                const unsigned tag_base = ctx->tag_counter;
                const int pgrid = (int)std::min<unsigned long long>((2 * bound + 127) / 128, (unsigned long long)ctx->sm_count * 8);
                kb_plan_kernel<<<pgrid, 128, 0, st>>>(ctx->d_jobs.as<KbJob>(), cur, rnd, (unsigned long long)rows_total, resident_warps,
                                                       force_thin, thin_k, batch_bonus ? 1 : 0, ctx->d_units.as<KbUnit>(), (unsigned)unit_cap, d_stats);
                KB_CUDA(cudaGetLastError());
                const int span = kb_span_begin(ctx, KB_SPAN_SWEEP);
                {
                        const KbJob* dj = ctx->d_jobs.as<KbJob>();
                        const KbUnit* du = ctx->d_units.as<KbUnit>();
                        const float* dt = ctx->d_tbl.as<float>();
                        const int thr = WARPS_PER_CTA * 32;
                        // no more CTAs than the round can have units (one warp per unit)
                        const unsigned long long ubound = std::min<unsigned long long>((unsigned long long)unit_cap,
                                                                                       2 * bound + (unsigned long long)rows_total / 32 + 2);
                        const int grid = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((unsigned long long)sweep_ctas,
                                                                                                           (ubound + WARPS_PER_CTA - 1) / WARPS_PER_CTA));
                        if (batch_dense) {
                                KB_CUDA(kb_sweep_launch_dense(grid, thr, st, dj, cur, du, rnd, tag_base, dt, tstride));
                        } else if (batch_bonus) {
                                KB_CUDA(kb_sweep_launch_sparse(grid, thr, st, dj, cur, du, rnd, tag_base, dt, tstride));
                        } else {
                                KB_CUDA(kb_sweep_launch_none(grid, thr, st, dj, cur, du, rnd, tag_base, dt, tstride));
                        }
                }
                kb_span_end(ctx, span);
                const int mgrid = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((bound + 3) / 4, (unsigned long long)ctx->sm_count * 16));
                kb_meetup_kernel<<<mgrid, 128, 0, st>>>(ctx->d_jobs.as<KbJob>(), cur, rnd, nxt, &d_rounds[round + 1].nboxes,
                                                        use_small ? d_small : nullptr, d_nsmall, small_rows, small_cols, (unsigned)box_cap, d_stats);
                KB_CUDA(cudaGetLastError());
                ctx->stats.n_launches += 3;
                if (trace) {
                        // debugging aid only: per-round read-back (serialises the rounds)
                        KbRound hr;
                        KbDevStats hs;
                        KB_CUDA(cudaMemcpyAsync(&hr, rnd, sizeof(hr), cudaMemcpyDeviceToHost, st));
                        KB_CUDA(cudaMemcpyAsync(&hs, d_stats, sizeof(hs), cudaMemcpyDeviceToHost, st));
                        KB_CUDA(cudaStreamSynchronize(st));
                        float ms = 0.0f;
                        cudaEventElapsedTime(&ms, ctx->ev_pool[2 * (size_t)span], ctx->ev_pool[2 * (size_t)span + 1]);
                        const unsigned long long csum = hs.cells[0] + hs.cells[1] + hs.cells[2];
                        if (hr.nboxes) {
                                fprintf(stderr, "[kb200 trace] jobs=%d round=%d boxes=%u units=%u thin=%u sweep_ms=%.3f cells=%llu\n", n, round, hr.nboxes,
                                        hr.nunits, hr.thin, ms, csum - trace_cells);
                        }
                        trace_cells = csum;
                }
                std::swap(cur, nxt);
        }
        kb_rounds_check_kernel<<<1, 1, 0, st>>>(d_rounds + nrounds, d_stats);
        {
                // every box that became small during the rounds: finish its recursion in one launch
                const int span = kb_span_begin(ctx, KB_SPAN_SMALL);
                // 4 CTAs per SM (fewer were measured slower: 50 / 57 / 63 / 85 ms at 4 / 3 / 2 / 1 on C3)
                const int sgrid = (int)std::max<size_t>(1, std::min<size_t>((size_t)ctx->sm_count * 4, (box_cap + 127) / 128));
                const KbJob* dj = ctx->d_jobs.as<KbJob>();
                const float* dt = ctx->d_tbl.as<float>();
                const unsigned bc = (unsigned)box_cap;
                const int fam = batch_dense ? BONUS_DENSE : (batch_bonus ? BONUS_SPARSE : BONUS_NONE);
                bool only_ss = true;
                for (int i = 0; i < n; i++) {
                        if (jobs[i].kind != KB200_KIND_SS) { only_ss = false; break; }
                }
                // (6 or 8 CTAs per SM for the all-seq-seq variant were measured: no gain -- the kernel is bound by its
                //  instruction count and divergence, not by occupancy or DRAM; profiles/r02_small_kernel_full_C3.csv)
#define KB_SMALL_LAUNCH(MC, FAM) \
        do { \
                if (only_ss) kb_small_kernel<MC, FAM, true, 4><<<sgrid, 128, 0, st>>>(dj, d_small, d_nsmall, bc, d_stats, dt, tstride); \
                else kb_small_kernel<MC, FAM, false, 4><<<sgrid, 128, 0, st>>>(dj, d_small, d_nsmall, bc, d_stats, dt, tstride); \
        } while (0)
                if (small_cols <= SMALL_COLS) {
                        if (fam == BONUS_DENSE) KB_SMALL_LAUNCH(SMALL_COLS, BONUS_DENSE);
                        else if (fam == BONUS_SPARSE) KB_SMALL_LAUNCH(SMALL_COLS, BONUS_SPARSE);
                        else KB_SMALL_LAUNCH(SMALL_COLS, BONUS_NONE);
                } else {
                        if (fam == BONUS_DENSE) KB_SMALL_LAUNCH(SMALL_COLS_MAX, BONUS_DENSE);
                        else if (fam == BONUS_SPARSE) KB_SMALL_LAUNCH(SMALL_COLS_MAX, BONUS_SPARSE);
                        else KB_SMALL_LAUNCH(SMALL_COLS_MAX, BONUS_NONE);
                }
#undef KB_SMALL_LAUNCH
                KB_CUDA(cudaGetLastError());
                kb_span_end(ctx, span);
                ctx->stats.n_launches += 2;
                if (trace) {
                        unsigned ns = 0;
                        KB_CUDA(cudaMemcpyAsync(&ns, d_nsmall, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
                        KB_CUDA(cudaStreamSynchronize(st));
                        float sms = 0.0f;
                        cudaEventElapsedTime(&sms, ctx->ev_pool[2 * (size_t)span], ctx->ev_pool[2 * (size_t)span + 1]);
                        fprintf(stderr, "[kb200 trace] jobs=%d small(<=%dx%d) boxes=%u small_ms=%.3f\n", n, small_rows, small_cols, ns, sms);
                }
        }
        kb_span_end(ctx, span_dp);
        return KB200_OK;
}
