// kb_bonus.cu -- device-resident gap weaving and anchor-consistency bonus.
//
// Replaces, with the same arithmetic and the same tie rules (nothing copied):
//   make_seq / update_gaps                 lib/src/weave_alignment.c:41,96
//   get_node_anchor_positions              lib/src/anchor_consistency.c:352-467
//   anchor_consistency_get_bonus_profile   lib/src/anchor_consistency.c:469-561
//
// State kept on the device for the whole tree: gaps[] of every sequence (len+1 ints, the layout of
// kb200_align_tree's gaps_out) and colof[p] = p + sum_{q<=p} gaps[q], the profile column of
// residue p in the node the sequence currently belongs to.
//
// Weaving: a member's gaps[i] grows by the number of new gap columns that fall into its i-th gap
// run; with P the prefix sum of the task's new-gap vector that is P[end_i+1] - P[rel_i] with
// rel_i = colof_old[i-1]+1 and end_i = rel_i + gaps_old[i] -- independent per (sequence, i).
//
// Votes: one thread per profile column visits the members in sip[] order ("first seen position
// wins", anchor_consistency.c:440-445), finds the member's residue in that column by binary search
// in colof, and counts agreement.  Inverse map: the largest column j wins (atomicMax ==
// "last write wins" of the ascending loop, :519-524).  Scatter: one thread per DP row adds its
// <=K terms in anchor order k into the dense matrix, (per_anchor_weight*conf_a)*conf_b, :532-533.
#include "kb_host.cuh"
#include "kb_bonus.cuh"

#include <algorithm>

namespace {

constexpr int KMAX = KB_BONUS_KMAX;

// ---- weave -----------------------------------------------------------------------------------
// per task, one warp: P[x] = number of new gap columns inserted before profile column x of the
// operand = number of gap entries that precede the x-th operand-consuming entry of the coded path
// (make_seq's gap_a / gap_b, weave_alignment.c:57-70, as prefix sums).
__global__ void kb_weave_prefix_kernel(const KbWeaveTask* __restrict__ tasks, const int ntasks)
{
        const int lane = threadIdx.x & 31;
        const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        if (q >= ntasks) return;
        const KbWeaveTask T = tasks[q];
        const int* __restrict__ path = T.path;
        int* __restrict__ Pa = T.Pa;
        int* __restrict__ Pb = T.Pb;
        const int alnlen = (T.alnlen >= 0) ? T.alnlen : path[0];
        int ka = 0, kb = 0, ga = 0, gb = 0;      // running counts: consumed columns / gap entries
        for (int base = 1; base <= alnlen; base += 32) {
                const int c = base + lane;
                int p = 0;
                const bool in = c <= alnlen;
                if (in) p = path[c];
                const int fa = (in && (!p || (p & 2))) ? 1 : 0;   // consumes a column of a
                const int fb = (in && (!p || (p & 1))) ? 1 : 0;   // consumes a column of b
                const int xa = (in && (p & 1)) ? 1 : 0;           // gap column in a
                const int xb = (in && (p & 2)) ? 1 : 0;           // gap column in b
                int ia = fa, ib = fb, ja = xa, jb = xb;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                        const int t0 = __shfl_up_sync(0xffffffffu, ia, o);
                        const int t1 = __shfl_up_sync(0xffffffffu, ib, o);
                        const int t2 = __shfl_up_sync(0xffffffffu, ja, o);
                        const int t3 = __shfl_up_sync(0xffffffffu, jb, o);
                        if (lane >= o) { ia += t0; ib += t1; ja += t2; jb += t3; }
                }
                if (fa) Pa[ka + ia] = ga + ja - xa;       // gap entries before this a-consuming entry
                if (fb) Pb[kb + ib] = gb + jb - xb;
                ka += __shfl_sync(0xffffffffu, ia, 31);
                kb += __shfl_sync(0xffffffffu, ib, 31);
                ga += __shfl_sync(0xffffffffu, ja, 31);
                gb += __shfl_sync(0xffffffffu, jb, 31);
        }
        if (lane == 0) {
                if (T.out_len) *T.out_len = alnlen;
                Pa[0] = 0; Pb[0] = 0;
                Pa[ka + 1] = ga;
                Pb[kb + 1] = gb;
        }
}

// one warp per member sequence: update gaps from the old colof, then rebuild colof
__global__ void kb_weave_apply_kernel(const KbWeaveMember* __restrict__ members, const int nmembers,
                                      const int64_t* __restrict__ offs, const int* __restrict__ lens,
                                      int* __restrict__ gaps, int* __restrict__ colof)
{
        const int lane = threadIdx.x & 31;
        const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        if (w >= nmembers) return;
        const KbWeaveMember M = members[w];
        const int si = M.seq;
        const int len = lens[si];
        int* __restrict__ g = gaps + offs[si] + si;
        int* __restrict__ co = colof + offs[si];
        const int* __restrict__ P = M.P;
        // pass 1: adds (reads old colof / old gaps only)
        for (int i = lane; i <= len; i += 32) {
                const int rel = (i == 0) ? 0 : (co[i - 1] + 1);
                const int gi = g[i];
                const int end = rel + gi;
                const int add = P[end + 1] - P[rel];
                g[i] = gi + add;
        }
        __syncwarp();
        // pass 2: colof[p] = p + inclusive prefix of gaps
        int carry = 0;
        for (int base = 0; base < len; base += 32) {
                const int p = base + lane;
                int v = (p < len) ? (g[p] + 1) : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, v, o);
                        if (lane >= o) v += t;
                }
                if (p < len) {
                        co[p] = carry + v - 1;
                }
                carry += __shfl_sync(0xffffffffu, v, 31);
        }
}

__global__ void kb_init_colof_kernel(const int64_t* __restrict__ offs, const int* __restrict__ lens, const int nseq,
                                     int* __restrict__ colof)
{
        const int lane = threadIdx.x & 31;
        const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        if (w >= nseq) return;
        int* co = colof + offs[w];
        for (int p = lane; p < lens[w]; p += 32) {
                co[p] = p;
        }
}

// ---- votes -----------------------------------------------------------------------------------
// Operands with many members.  Work unit = (chunk of 32 profile columns, slice of VOTE_SLICE
// members), one WARP each, so that the top tree levels (a handful of operands with thousands of
// members) still spread over the whole machine.  Members are taken 32 at a time: first every lane
// finds, for its own member, the first residue at or after the chunk's first column (binary search
// in colof); then the 32 members are visited in sip[] order, lane j reading the member's residue
// a_lo+j -- the residues that fall into the chunk are consecutive -- and its K anchor positions
// (coalesced); the owner lane of a column picks them up by shuffle: residue columns are strictly
// increasing, so the j-th residue in the chunk is the j-th set bit of the chunk's occupancy mask.
//   pass 0: "first seen position wins" (anchor_consistency.c:440-445) = the vote of the valid
//           member with the smallest index: atomicMin of (member index, position) per column;
//           total votes by atomicAdd
//   pass 1: votes agreeing with the winner, atomicAdd
//   finalize: positions / confidence = agree / total (:449-457)
constexpr int VOTE_SLICE = 64;

template <int PASS>
__global__ void __launch_bounds__(128)
kb_bonus_votes_kernel(const KbBonusOperand* __restrict__ ops, const KbVoteOp* __restrict__ vops, const int nvops,
                      const long long total_units, const int K,
                      const int* __restrict__ memb, const int64_t* __restrict__ offs,
                      const int* __restrict__ lens, const int* __restrict__ colof,
                      const int* __restrict__ posmaps,
                      unsigned long long* __restrict__ vkey, int* __restrict__ vtotal, int* __restrict__ vagree)
{
        const int lane = threadIdx.x & 31;
        const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        if (gid >= total_units) return;
        int lo = 0, hi = nvops - 1;
        while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (vops[mid].unit0 <= gid) lo = mid; else hi = mid - 1;
        }
        const KbVoteOp VO = vops[lo];
        const KbBonusOperand O = ops[VO.op];
        const long long lu = gid - VO.unit0;
        const int chunk = (int)(lu / VO.nslices);
        const int slice = (int)(lu % VO.nslices);
        const int c0 = chunk * 32;
        const int c = c0 + lane;                 // the column this lane owns
        const int nmem = O.m1 - O.m0;
        const int m_begin = slice * VOTE_SLICE;
        const int m_end = min(nmem, m_begin + VOTE_SLICE);
        int best[KMAX], first_m[KMAX], count[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; k++) {
                best[k] = -1; first_m[k] = -1; count[k] = 0;
                if (PASS == 1 && k < K && c < O.len) {
                        const unsigned long long key = vkey[VO.vote0 + (long long)k * O.len + c];
                        best[k] = (key == ~0ull) ? -1 : (int)(unsigned)(key & 0xffffffffull);
                }
        }
        const unsigned lt_mask = (1u << lane) - 1u;
        for (int mb = m_begin; mb < m_end; mb += 32) {
                // ---- lane = member: locate the chunk in the member's residues ----
                int my_len = 0, my_alo = 0;
                long long my_off = 0;
                if (mb + lane < m_end) {
                        const int si = memb[O.m0 + mb + lane];
                        my_len = lens[si];
                        my_off = (long long)offs[si];
                        const int* __restrict__ co = colof + my_off;
                        int a = 0, b = my_len;
                        while (a < b) {
                                const int mid = (a + b) >> 1;
                                if (co[mid] < c0) a = mid + 1; else b = mid;
                        }
                        my_alo = a;
                }
                const int cnt = min(32, m_end - mb);
                // ---- members in sip order; lane = residue a_lo + lane, then lane = column ----
#pragma unroll 4
                for (int i = 0; i < cnt; i++) {
                        const int len = __shfl_sync(0xffffffffu, my_len, i);
                        const int alo = __shfl_sync(0xffffffffu, my_alo, i);
                        const long long off = __shfl_sync(0xffffffffu, my_off, i);
                        const int idx = alo + lane;
                        int rel = 32;
                        if (idx < len) {
                                rel = colof[off + idx] - c0;
                        }
                        const bool in = rel < 32;
                        const unsigned occ = __reduce_or_sync(0xffffffffu, in ? (1u << rel) : 0u);
                        if (occ == 0u) continue;
                        const int src = __popc(occ & lt_mask);
                        const bool has = (occ >> lane) & 1u;
                        const int* __restrict__ map0 = posmaps + (size_t)K * (size_t)off + idx;
#pragma unroll
                        for (int k = 0; k < KMAX; k++) {
                                if (k < K) {
                                        const int mine = in ? map0[(size_t)k * len] : -1;
                                        const int apos = __shfl_sync(0xffffffffu, mine, src);
                                        if (has && apos >= 0) {
                                                if (PASS == 0) {
                                                        if (count[k] == 0) { best[k] = apos; first_m[k] = mb + i; }
                                                        count[k]++;
                                                } else {
                                                        if (apos == best[k]) count[k]++;
                                                }
                                        }
                                }
                        }
                }
        }
        if (c < O.len) {
#pragma unroll
                for (int k = 0; k < KMAX; k++) {
                        if (k < K && count[k] > 0) {
                                const long long slot = VO.vote0 + (long long)k * O.len + c;
                                if (PASS == 0) {
                                        atomicMin(vkey + slot, ((unsigned long long)(unsigned)first_m[k] << 32) | (unsigned long long)(unsigned)best[k]);
                                        atomicAdd(vtotal + slot, count[k]);
                                } else {
                                        atomicAdd(vagree + slot, count[k]);
                                }
                        }
                }
        }
}

__global__ void kb_bonus_votes_final_kernel(const KbBonusOperand* __restrict__ ops, const KbVoteOp* __restrict__ vops, const int nvops,
                                            const long long total_cols, const int K,
                                            const unsigned long long* __restrict__ vkey, const int* __restrict__ vtotal,
                                            const int* __restrict__ vagree)
{
        const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (gid >= total_cols) return;
        int lo = 0, hi = nvops - 1;
        while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (vops[mid].col0 <= gid) lo = mid; else hi = mid - 1;
        }
        const KbVoteOp VO = vops[lo];
        const KbBonusOperand O = ops[VO.op];
        const int c = (int)(gid - VO.col0);
        for (int k = 0; k < K; k++) {
                const long long slot = VO.vote0 + (long long)k * O.len + c;
                const unsigned long long key = vkey[slot];
                const int total = vtotal[slot], agree = vagree[slot];
                const bool ok = (key != ~0ull) && total > 0 && agree > 0;
                O.pos[(size_t)k * O.len + c] = ok ? (int)(unsigned)(key & 0xffffffffull) : -1;
                O.conf[(size_t)k * O.len + c] = ok ? ((float)agree / (float)total) : 0.0f;
        }
}

// Operands with few members (the bulk of the low tree levels: leaves and profiles of a handful of
// sequences): one THREAD per profile column walks the members serially in sip[] order -- first seen
// position, total and agreeing votes in ONE pass, exactly the reference's loop order -- finding the
// member's residue in its column by binary search in colof.  Adjacent threads are adjacent columns
// of one operand.

__global__ void kb_bonus_positions_small_kernel(const KbBonusOperand* __restrict__ ops, const int* __restrict__ op_list,
                                                const long long* __restrict__ col_prefix,
                                                const int nops, const long long total_cols, const int K,
                                                const int* __restrict__ memb, const int64_t* __restrict__ offs,
                                                const int* __restrict__ lens, const int* __restrict__ colof,
                                                const int* __restrict__ posmaps)
{
        const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (gid >= total_cols) return;
        int lo = 0, hi = nops - 1;
        while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (col_prefix[mid] <= gid) lo = mid; else hi = mid - 1;
        }
        const KbBonusOperand O = ops[op_list[lo]];
        const int c = (int)(gid - col_prefix[lo]);
        int* __restrict__ pos = O.pos;        // [K][len]
        float* __restrict__ conf = O.conf;
        const int nmem = O.m1 - O.m0;
        if (nmem == 1) {
                // leaf: direct lookup (anchor_consistency.c:360-378)
                const int si = memb[O.m0];
                const int seq_len = lens[si];
                for (int k = 0; k < K; k++) {
                        int v = -1;
                        if (c < seq_len) {
                                v = posmaps[(size_t)K * (size_t)offs[si] + (size_t)k * (size_t)seq_len + c];
                        }
                        pos[(size_t)k * O.len + c] = v;
                        conf[(size_t)k * O.len + c] = (v >= 0) ? 1.0f : 0.0f;
                }
                return;
        }
        int best[KMAX], total[KMAX], agree[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; k++) {
                best[k] = -1; total[k] = 0; agree[k] = 0;
        }
#pragma unroll 1
        for (int m = 0; m < nmem; m++) {
                const int si = memb[O.m0 + m];
                const int seq_len = lens[si];
                const int* __restrict__ co = colof + offs[si];
                int a = 0, b = seq_len;
                while (a < b) {
                        const int mid = (a + b) >> 1;
                        if (co[mid] < c) a = mid + 1; else b = mid;
                }
                const bool hit = (a < seq_len) && (co[a] == c);
                if (!hit) continue;
                const int* __restrict__ map0 = posmaps + (size_t)K * (size_t)offs[si] + a;
#pragma unroll
                for (int k = 0; k < KMAX; k++) {
                        if (k < K) {
                                const int apos = map0[(size_t)k * seq_len];
                                if (apos >= 0) {
                                        if (total[k] == 0) best[k] = apos;       // first seen position wins
                                        total[k]++;
                                        if (apos == best[k]) agree[k]++;          // best is final once set
                                }
                        }
                }
        }
#pragma unroll
        for (int k = 0; k < KMAX; k++) {
                if (k < K) {
                        const bool ok = total[k] > 0 && agree[k] > 0;
                        pos[(size_t)k * O.len + c] = ok ? best[k] : -1;
                        conf[(size_t)k * O.len + c] = ok ? ((float)agree[k] / (float)total[k]) : 0.0f;
                }
        }
}

// inverse map of the column operand: inv[k][anchor_pos] = largest column j mapping there
__global__ void kb_bonus_inverse_kernel(const KbBonusTask* __restrict__ tasks, const long long* __restrict__ col_prefix,
                                        const int ntasks, const long long total_cols, const int K,
                                        const int* __restrict__ aoff)
{
        const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (gid >= total_cols) return;
        int lo = 0, hi = ntasks - 1;
        while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (col_prefix[mid] <= gid) lo = mid; else hi = mid - 1;
        }
        const KbBonusTask T = tasks[lo];
        const int j = (int)(gid - col_prefix[lo]);
        for (int k = 0; k < K; k++) {
                const int ap = T.pos_b[(size_t)k * T.len_b + j];
                if (ap >= 0) {
                        atomicMax(T.inv + aoff[k] + ap, j);
                }
        }
}

// one thread per DP row: the row's <=K bonus terms, merged per column in anchor order k (a cell hit
// by several anchors sums its terms in k order, exactly like the reference's += into a zeroed
// matrix) and sorted by column: bcol[row*K + e] ascending, unused slots = INT_MAX.
__global__ void kb_bonus_scatter_kernel(const KbBonusTask* __restrict__ tasks, const long long* __restrict__ row_prefix,
                                        const int ntasks, const long long total_rows, const int K,
                                        const int* __restrict__ aoff, const float paw)
{
        const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (gid >= total_rows) return;
        int lo = 0, hi = ntasks - 1;
        while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (row_prefix[mid] <= gid) lo = mid; else hi = mid - 1;
        }
        const KbBonusTask T = tasks[lo];
        const int i = (int)(gid - row_prefix[lo]);
        int cols[KMAX];
        float vals[KMAX];
        int n = 0;
        for (int k = 0; k < K; k++) {
                const int ak = T.pos_a[(size_t)k * T.len_a + i];
                if (ak < 0) continue;
                const int bj = T.inv[aoff[k] + ak];
                if (bj < 0) continue;
                const float ca = T.conf_a[(size_t)k * T.len_a + i];
                const float cb = T.conf_b[(size_t)k * T.len_b + bj];
                const float v = __fmul_rn(__fmul_rn(paw, ca), cb);
                int e = 0;
                for (; e < n; e++) {
                        if (cols[e] == bj) break;
                }
                if (e < n) {
                        vals[e] = __fadd_rn(vals[e], v);           // later anchor, same cell
                } else {
                        cols[n] = bj;
                        vals[n] = __fadd_rn(0.0f, v);               // 0 + v, as += into the zeroed matrix
                        n++;
                }
        }
        // insertion sort by column (n <= K <= 8)
        for (int x = 1; x < n; x++) {
                const int c = cols[x];
                const float v = vals[x];
                int y = x - 1;
                while (y >= 0 && cols[y] > c) {
                        cols[y + 1] = cols[y];
                        vals[y + 1] = vals[y];
                        y--;
                }
                cols[y + 1] = c;
                vals[y + 1] = v;
        }
        int* __restrict__ oc = T.bcol + (size_t)i * K;
        float* __restrict__ ov = T.bval + (size_t)i * K;
        for (int e = 0; e < K; e++) {
                oc[e] = (e < n) ? cols[e] : 0x7fffffff;
                ov[e] = (e < n) ? vals[e] : 0.0f;
        }
}

} // namespace

int kb_bonus_init_state(kb200_ctx* ctx, KbSeqs& S, int* d_gaps, int* d_colof)
{
        KB_CUDA(cudaMemsetAsync(d_gaps, 0, sizeof(int) * ((size_t)S.total + (size_t)S.n), ctx->stream));
        const int warps = S.n;
        kb_init_colof_kernel<<<(warps * 32 + 127) / 128, 128, 0, ctx->stream>>>(S.d_offs.as<int64_t>(), S.d_lens.as<int>(), S.n, d_colof);
        KB_CUDA(cudaGetLastError());
        ctx->stats.n_launches++;
        return KB200_OK;
}

int kb_weave_level(kb200_ctx* ctx, KbSeqs& S, const KbWeaveTask* d_tasks, int ntasks,
                   const KbWeaveMember* d_members, int nmembers, int* d_gaps, int* d_colof)
{
        if (ntasks <= 0) return KB200_OK;
        kb_weave_prefix_kernel<<<(ntasks * 32 + 127) / 128, 128, 0, ctx->stream>>>(d_tasks, ntasks);
        KB_CUDA(cudaGetLastError());
        if (nmembers > 0) {
                kb_weave_apply_kernel<<<(nmembers * 32 + 127) / 128, 128, 0, ctx->stream>>>(d_members, nmembers, S.d_offs.as<int64_t>(),
                                                                                             S.d_lens.as<int>(), d_gaps, d_colof);
                KB_CUDA(cudaGetLastError());
        }
        ctx->stats.n_launches += 2;
        return KB200_OK;
}

int kb_bonus_level(kb200_ctx* ctx, KbSeqs& S, int K, float paw,
                   const KbBonusOperand* d_ops,
                   const int* d_small_list, const long long* d_small_prefix, int n_small, long long small_cols,
                   const KbVoteOp* d_vops, int n_large, long long large_units, long long large_cols, long long vote_slots,
                   const int* d_memb, const int* d_colof, const int* d_posmaps,
                   const KbBonusTask* d_tasks, const long long* d_colb_prefix, long long colb_total,
                   const long long* d_row_prefix, long long row_total, int ntasks, const int* d_aoff)
{
        if (ntasks <= 0) return KB200_OK;
        cudaStream_t st = ctx->stream;
        if (n_small > 0) {
                kb_bonus_positions_small_kernel<<<(unsigned)((small_cols + 127) / 128), 128, 0, st>>>(
                        d_ops, d_small_list, d_small_prefix, n_small, small_cols, K, d_memb, S.d_offs.as<int64_t>(), S.d_lens.as<int>(),
                        d_colof, d_posmaps);
                KB_CUDA(cudaGetLastError());
                ctx->stats.n_launches++;
        }
        if (n_large > 0) {
                // vote scratch: [keys u64 | totals int | agreeing int], one slot per (operand column, anchor)
                KB_RUN(ctx->t_bvote.ensure((size_t)vote_slots * 16 + 64));
                unsigned long long* vkey = ctx->t_bvote.as<unsigned long long>();
                int* vtotal = (int*)(vkey + vote_slots);
                int* vagree = vtotal + vote_slots;
                KB_CUDA(cudaMemsetAsync(vkey, 0xFF, (size_t)vote_slots * 8, st));
                KB_CUDA(cudaMemsetAsync(vtotal, 0, (size_t)vote_slots * 8, st));
                const unsigned grid = (unsigned)((large_units * 32 + 127) / 128);
                kb_bonus_votes_kernel<0><<<grid, 128, 0, st>>>(d_ops, d_vops, n_large, large_units, K, d_memb, S.d_offs.as<int64_t>(),
                                                                 S.d_lens.as<int>(), d_colof, d_posmaps, vkey, vtotal, vagree);
                KB_CUDA(cudaGetLastError());
                kb_bonus_votes_kernel<1><<<grid, 128, 0, st>>>(d_ops, d_vops, n_large, large_units, K, d_memb, S.d_offs.as<int64_t>(),
                                                                 S.d_lens.as<int>(), d_colof, d_posmaps, vkey, vtotal, vagree);
                KB_CUDA(cudaGetLastError());
                kb_bonus_votes_final_kernel<<<(unsigned)((large_cols + 127) / 128), 128, 0, st>>>(d_ops, d_vops, n_large, large_cols, K, vkey, vtotal, vagree);
                KB_CUDA(cudaGetLastError());
                ctx->stats.n_launches += 3;
        }
        kb_bonus_inverse_kernel<<<(unsigned)((colb_total + 127) / 128), 128, 0, st>>>(d_tasks, d_colb_prefix, ntasks, colb_total, K, d_aoff);
        KB_CUDA(cudaGetLastError());
        kb_bonus_scatter_kernel<<<(unsigned)((row_total + 127) / 128), 128, 0, st>>>(d_tasks, d_row_prefix, ntasks, row_total, K, d_aoff, paw);
        KB_CUDA(cudaGetLastError());
        ctx->stats.n_launches += 2;
        return KB200_OK;
}
