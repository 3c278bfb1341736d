// kb_tree.cu -- progressive alignment over the guide tree, one batched launch sequence per level.
//
// Replaces the aln_task scheduler and do_align:
//   create_msa_tree / recursive_aln   lib/src/aln_run.c:43,81   (post-order OpenMP task recursion)
//   do_align                          lib/src/aln_run.c:213-441
//   compute_subm_offset               lib/src/aln_run.c:166-203
//   make_seq / update_gaps            lib/src/weave_alignment.c:41,96
//   anchor_consistency_get_bonus_profile / get_node_anchor_positions
//                                     lib/src/anchor_consistency.c:469,352
//
// All tasks whose children are finished (same "level": 1 + max level of the children) are
// mutually independent (aln_run.c:95-109).  Per level: leaf profiles / gap-penalty rescale
// (streaming kernels) -> one batched Hirschberg run over every task of the level -> path coding
// on device -> profile merge on device -> the coded paths come back to the host for the gap
// weaving bookkeeping (msa->gaps, sip, nsip, plen).  Profiles never leave the device.
#include "kb_host.cuh"
#include "kb_bonus.cuh"

#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <utility>
#include <chrono>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

__global__ void kb_scatter_kernel(float* __restrict__ dst, const long long* __restrict__ idx,
                                  const float* __restrict__ val, const long long n)
{
        const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (i < n) {
                dst[idx[i]] = val[i];
        }
}

struct Tree {
        int N = 0;
        const KbSeqs* S = nullptr;
        std::vector<int> nsip, plen, level;
        std::vector<std::vector<int>> sip;
        std::vector<float*> prof;
        std::vector<std::vector<int>> gaps;   // per sequence, len+1
        const int* posmaps = nullptr;
        int K = 0;
        float weight = 0.0f;
};

// ---- consistency bonus, level-wide ------------------------------------------------------------
// get_node_anchor_positions (anchor_consistency.c:352-467) for every operand of a tree level and
// all K anchors at once.  Parallel decomposition (results are order independent per column):
// work item = (operand node, block of profile columns); inside a block the members are visited
// in sip[] order, which is all the "first-seen position wins" vote needs (:440-445).
struct NodePos {
        int len = 0;
        std::vector<int> pos;      // K x len, k-major
        std::vector<float> conf;
};

struct PosChunk {
        int slot;                  // index into the level's NodePos array
        int node;
        int c0, c1;
};

// column of residue p of sequence si in its current profile: p + sum_{q<=p} gaps[q]
void fill_colof(const Tree& T, int si, std::vector<int>& co)
{
        const std::vector<int>& g = T.gaps[(size_t)si];
        const int len = T.S->h_lens[si];
        co.resize((size_t)len);
        int col = 0;
        for (int p = 0; p < len; p++) {
                col += g[(size_t)p];
                co[(size_t)p] = col;
                col++;
        }
}

void pos_chunk(const Tree& T, const std::vector<std::vector<int>>& colof, const PosChunk& w, NodePos& out)
{
        const KbSeqs& S = *T.S;
        const int K = T.K;
        const int len = out.len;
        if (T.nsip[w.node] == 1) {
                const int seq_len = S.h_lens[w.node];
                for (int k = 0; k < K; k++) {
                        const int* map = T.posmaps + (size_t)K * (size_t)S.h_offs[w.node] + (size_t)k * (size_t)seq_len;
                        for (int i = w.c0; i < w.c1; i++) {
                                if (i < seq_len) {
                                        out.pos[(size_t)k * len + i] = map[i];
                                        out.conf[(size_t)k * len + i] = (map[i] >= 0) ? 1.0f : 0.0f;
                                } else {
                                        out.pos[(size_t)k * len + i] = -1;
                                        out.conf[(size_t)k * len + i] = 0.0f;
                                }
                        }
                }
                return;
        }
        const int W = w.c1 - w.c0;
        std::vector<int> best((size_t)K * W, -1), agree((size_t)K * W, 0), total((size_t)K * W, 0);
        for (int si : T.sip[(size_t)w.node]) {
                const std::vector<int>& co = colof[(size_t)si];
                const int seq_len = S.h_lens[si];
                const int* map0 = T.posmaps + (size_t)K * (size_t)S.h_offs[si];
                int p = (int)(std::lower_bound(co.begin(), co.end(), w.c0) - co.begin());
                for (; p < seq_len && co[(size_t)p] < w.c1; p++) {
                        const int c = co[(size_t)p] - w.c0;
                        for (int k = 0; k < K; k++) {
                                const int apos = map0[(size_t)k * seq_len + p];
                                if (apos < 0) continue;
                                const size_t e = (size_t)k * W + c;
                                total[e]++;
                                if (best[e] < 0) {
                                        best[e] = apos;
                                        agree[e] = 1;
                                } else if (apos == best[e]) {
                                        agree[e]++;
                                }
                        }
                }
        }
        for (int k = 0; k < K; k++) {
                for (int c = 0; c < W; c++) {
                        const size_t e = (size_t)k * W + c;
                        const size_t o = (size_t)k * len + (w.c0 + c);
                        if (total[e] > 0 && agree[e] > 0) {
                                out.pos[o] = best[e];
                                out.conf[o] = (float)agree[e] / (float)total[e];
                        } else {
                                out.pos[o] = -1;
                                out.conf[o] = 0.0f;
                        }
                }
        }
}

// anchor_consistency_get_bonus_profile, anchor_consistency.c:469-561, as a sorted sparse list of
// (flat index i*len_b + bj, value); contributions to one cell are summed in anchor order k.
void combine_bonus(const Tree& T, const NodePos& A, const NodePos& B, std::vector<std::pair<long long, float>>& out)
{
        out.clear();
        const int K = T.K;
        const int len_a = A.len, len_b = B.len;
        const float paw = T.weight / (float)K;
        std::vector<int> inv_b;
        std::vector<float> inv_conf_b;
        std::vector<std::pair<long long, float>> ent;
        for (int k = 0; k < K; k++) {
                const int* apos_a = A.pos.data() + (size_t)k * len_a;
                const int* apos_b = B.pos.data() + (size_t)k * len_b;
                const float* conf_a = A.conf.data() + (size_t)k * len_a;
                const float* conf_b = B.conf.data() + (size_t)k * len_b;
                int anchor_len = 0;
                for (int i = 0; i < len_a; i++) {
                        if (apos_a[i] >= anchor_len) anchor_len = apos_a[i] + 1;
                }
                for (int j = 0; j < len_b; j++) {
                        if (apos_b[j] >= anchor_len) anchor_len = apos_b[j] + 1;
                }
                if (anchor_len == 0) continue;
                inv_b.assign((size_t)anchor_len, -1);
                inv_conf_b.assign((size_t)anchor_len, 0.0f);
                for (int j = 0; j < len_b; j++) {
                        if (apos_b[j] >= 0 && apos_b[j] < anchor_len) {
                                inv_b[(size_t)apos_b[j]] = j;
                                inv_conf_b[(size_t)apos_b[j]] = conf_b[j];
                        }
                }
                for (int i = 0; i < len_a; i++) {
                        const int ak = apos_a[i];
                        if (ak >= 0 && ak < anchor_len) {
                                const int bj = inv_b[(size_t)ak];
                                if (bj >= 0) {
                                        const float v = paw * conf_a[i] * inv_conf_b[(size_t)ak];
                                        ent.emplace_back((long long)i * (long long)len_b + (long long)bj, v);
                                }
                        }
                }
        }
        std::stable_sort(ent.begin(), ent.end(),
                         [](const std::pair<long long, float>& x, const std::pair<long long, float>& y) { return x.first < y.first; });
        for (size_t i = 0; i < ent.size();) {
                float v = 0.0f;
                size_t j = i;
                while (j < ent.size() && ent[j].first == ent[i].first) {
                        v += ent[j].second;
                        j++;
                }
                out.emplace_back(ent[i].first, v);
                i = j;
        }
}

void level_bonus(Tree& T, int q0, int q1, const std::vector<int>& rown, const std::vector<int>& rlen,
                 const std::vector<int>& coln, const std::vector<int>& clen, int n_threads,
                 std::vector<std::vector<int>>& colof,
                 std::vector<std::vector<std::pair<long long, float>>>& lists)
{
        const int K = T.K;
        const int nt = (int)rown.size();
        const int nm = q1 - q0;
        (void)n_threads;
        // members of every profile operand of my tasks of this level
        std::vector<int> members;
        for (int q = q0; q < q1; q++) {
                const int nodes[2] = {rown[(size_t)q], coln[(size_t)q]};
                for (int s = 0; s < 2; s++) {
                        if (T.nsip[(size_t)nodes[s]] > 1) {
                                members.insert(members.end(), T.sip[(size_t)nodes[s]].begin(), T.sip[(size_t)nodes[s]].end());
                        }
                }
        }
        std::vector<NodePos> np((size_t)2 * std::max(nm, 0));
        std::vector<PosChunk> chunks;
        const int BLK = 512;
        for (int q = q0; q < q1; q++) {
                const int nodes[2] = {rown[(size_t)q], coln[(size_t)q]};
                const int lens2[2] = {rlen[(size_t)q], clen[(size_t)q]};
                for (int s = 0; s < 2; s++) {
                        NodePos& P = np[(size_t)2 * (q - q0) + s];
                        P.len = lens2[s];
                        P.pos.resize((size_t)K * P.len);
                        P.conf.resize((size_t)K * P.len);
                        for (int c0 = 0; c0 < P.len; c0 += BLK) {
                                PosChunk w;
                                w.slot = 2 * (q - q0) + s; w.node = nodes[s]; w.c0 = c0; w.c1 = std::min(P.len, c0 + BLK);
                                chunks.push_back(w);
                        }
                }
        }
        lists.assign((size_t)nt, {});
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
#endif
        {
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
                for (long long m = 0; m < (long long)members.size(); m++) {
                        fill_colof(T, members[(size_t)m], colof[(size_t)members[(size_t)m]]);
                }
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
                for (long long w = 0; w < (long long)chunks.size(); w++) {
                        pos_chunk(T, colof, chunks[(size_t)w], np[(size_t)chunks[(size_t)w].slot]);
                }
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
                for (int q = q0; q < q1; q++) {
                        combine_bonus(T, np[(size_t)2 * (q - q0)], np[(size_t)2 * (q - q0) + 1], lists[(size_t)q]);
                }
        }
}

// make_seq + update_gaps, weave_alignment.c:41-112, level-wide: gap vectors per task, then one
// independent update per member sequence.
struct WeaveItem {
        int si;
        const std::vector<int>* ng;
};

void level_weave(Tree& T, int nt, const std::vector<int>& tl, const int* tasks_abc, const int* hcoded,
                 const std::vector<size_t>& coded_off, int n_threads)
{
        std::vector<std::vector<int>> gap_a((size_t)nt), gap_b((size_t)nt);
        std::vector<WeaveItem> items;
        for (int q = 0; q < nt; q++) {
                const int* path = hcoded + coded_off[(size_t)q];
                gap_a[(size_t)q].assign((size_t)path[0] + 1, 0);
                gap_b[(size_t)q].assign((size_t)path[0] + 1, 0);
        }
        for (int q = 0; q < nt; q++) {
                const int t = tl[(size_t)q];
                for (int si : T.sip[(size_t)tasks_abc[3 * t]]) items.push_back({si, &gap_a[(size_t)q]});
                for (int si : T.sip[(size_t)tasks_abc[3 * t + 1]]) items.push_back({si, &gap_b[(size_t)q]});
        }
        (void)n_threads;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
#endif
        {
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
                for (int q = 0; q < nt; q++) {
                        const int* path = hcoded + coded_off[(size_t)q];
                        std::vector<int>& ga = gap_a[(size_t)q];
                        std::vector<int>& gb = gap_b[(size_t)q];
                        int posa = 0, posb = 0;
                        for (int c = 1; path[c] != 3; c++) {
                                const int p = path[c];
                                if (!p) {
                                        posa++; posb++;
                                } else if (p & 1) {
                                        ga[(size_t)posa] += 1; posb++;
                                } else if (p & 2) {
                                        gb[(size_t)posb] += 1; posa++;
                                }
                        }
                }
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 8)
#endif
                for (long long m = 0; m < (long long)items.size(); m++) {
                        const WeaveItem& it = items[(size_t)m];
                        std::vector<int>& gis = T.gaps[(size_t)it.si];
                        const std::vector<int>& ng = *it.ng;
                        const int old_len = T.S->h_lens[it.si];
                        int rel = 0;
                        for (int i = 0; i <= old_len; i++) {
                                int add = 0;
                                for (int j = rel; j <= rel + gis[(size_t)i]; j++) {
                                        add += ng[(size_t)j];
                                }
                                rel += gis[(size_t)i] + 1;
                                gis[(size_t)i] += add;
                        }
                }
        }
}

} // namespace

// KB200_HOST_BONUS=1 selects the host restatement of the bonus / weave stages -- an explicit A/B
// switch for debugging, never chosen silently (more anchors than the device kernels handle is an error)
int kb_bonus_on_host(int K)
{
        (void)K;
        return getenv("KB200_HOST_BONUS") != nullptr;
}

int kb_align_tree_dev(kb200_ctx* ctx, const kb200_params* prm, KbSeqs& S,
                      const int* tasks_abc, int ntasks, const float* seq_distances,
                      const int* posmaps, int K, float weight, int n_threads, int* gaps_out, int posmaps_on_device,
                      float* conf_out, int* plen_out)
{
        const int N = S.n;
        if (ntasks != N - 1 || N < 2) {
                fprintf(stderr, "[kalign_b200] align_tree: need exactly N-1 tasks (N=%d, ntasks=%d)\n", N, ntasks);
                return KB200_FAIL;
        }
        if (K > KB_BONUS_KMAX && posmaps && !kb_bonus_on_host(K)) {
                fprintf(stderr, "[kalign_b200] align_tree: %d consistency anchors, the device kernels handle at most %d\n", K, KB_BONUS_KMAX);
                return KB200_FAIL;
        }
        cudaStream_t st = ctx->stream;
        const int NP = 2 * N - 1;
        Tree T;
        T.N = N; T.S = &S;
        T.nsip.assign(NP, 0); T.plen.assign(NP, 0); T.level.assign(NP, 0);
        T.sip.resize(NP); T.prof.assign(NP, nullptr);
        T.gaps.resize(N);
        T.posmaps = (K > 0 && N >= 3) ? posmaps : nullptr;     // anchor_consistency_build: N<3 -> no table
        T.K = std::min(K, N); T.weight = weight;
        for (int i = 0; i < N; i++) {
                T.nsip[i] = 1;
                T.sip[i].assign(1, i);
                T.plen[i] = 0;
                T.gaps[i].assign((size_t)S.h_lens[i] + 1, 0);
        }
        int maxlevel = 0;
        {
                // the task list must be sort_tasks' order (lib/src/task.c:114): node c = N + t, both
                // operands produced earlier and consumed exactly once
                std::vector<char> used((size_t)NP, 0);
                for (int t = 0; t < ntasks; t++) {
                        const int a = tasks_abc[3 * t], b = tasks_abc[3 * t + 1], c = tasks_abc[3 * t + 2];
                        if (c != N + t || a < 0 || b < 0 || a >= c || b >= c || a == b || used[(size_t)a] || used[(size_t)b]) {
                                fprintf(stderr, "[kalign_b200] align_tree: task %d (%d, %d -> %d) is not part of a guide tree sorted by node id\n", t, a, b, c);
                                return KB200_FAIL;
                        }
                        used[(size_t)a] = 1;
                        used[(size_t)b] = 1;
                }
        }
        for (int t = 0; t < ntasks; t++) {
                const int a = tasks_abc[3 * t], b = tasks_abc[3 * t + 1], c = tasks_abc[3 * t + 2];
                if (a < 0 || b < 0 || c < N || a >= NP || b >= NP || c >= NP) {
                        return KB200_FAIL;
                }
                T.level[c] = 1 + std::max(T.level[a], T.level[b]);
                maxlevel = std::max(maxlevel, T.level[c]);
        }
        std::vector<std::vector<int>> by_level((size_t)maxlevel + 1);
        for (int t = 0; t < ntasks; t++) {
                by_level[(size_t)T.level[tasks_abc[3 * t + 2]]].push_back(t);
        }
        // device scratch lives in the context and is reused by later calls
        KbArena& arena = ctx->arena;
        arena.reset();
        KbDevBuf &d_subm = ctx->t_subm, &d_leaf = ctx->t_leaf, &d_gapset = ctx->t_gapset, &d_prefix = ctx->t_prefix,
                 &d_raw = ctx->t_raw, &d_coded = ctx->t_coded, &d_scr = ctx->t_scr, &d_pjobs = ctx->t_pjobs,
                 &d_mjobs = ctx->t_mjobs, &d_src = ctx->t_src, &d_bonus = ctx->t_bonus, &d_bidx = ctx->t_bidx, &d_bval = ctx->t_bval;
        int rc = KB200_OK;
        auto cleanup = [&]() {};
#define TR(x) do { if ((x) != KB200_OK) { fprintf(stderr, "[kalign_b200] align_tree failure at %s:%d\n", __FILE__, __LINE__); cleanup(); return KB200_FAIL; } } while (0)
#define TC(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { fprintf(stderr, "[kalign_b200] CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); cleanup(); return KB200_FAIL; } } while (0)
        TR(d_subm.ensure(sizeof(float) * 23 * 23));
        TR(kb_h2d(ctx, d_subm.p, prm->subm, sizeof(float) * 23 * 23));

        std::vector<std::vector<std::pair<long long, float>>> bonus_lists;
        std::vector<std::vector<int>> colof((size_t)N);
        // ---- device-resident gaps / colof / position maps (kb_bonus.cu); KB200_HOST_BONUS=1 keeps the
        //      host implementation of weave_alignment.c / anchor_consistency.c for A/B checks
        const bool dev_state = !kb_bonus_on_host(T.K);
        int* d_gaps = nullptr;
        int* d_colof = nullptr;
        const int* d_posmaps = nullptr;
        int maxlen = 0;
        for (int i = 0; i < N; i++) maxlen = std::max(maxlen, S.h_lens[i]);
        if (dev_state) {
                TR(ctx->t_gaps.ensure(sizeof(int) * ((size_t)S.total + (size_t)N + 8)));
                TR(ctx->t_colof.ensure(sizeof(int) * ((size_t)S.total + 8)));
                d_gaps = ctx->t_gaps.as<int>();
                d_colof = ctx->t_colof.as<int>();
                TR(kb_bonus_init_state(ctx, S, d_gaps, d_colof));
                if (T.posmaps) {
                        const size_t full = (size_t)T.K * (size_t)S.total;
                        if (!posmaps_on_device || ctx->posmaps_tag != (const void*)T.posmaps || ctx->posmaps_n != full) {
                                TR(ctx->t_posmaps.ensure(sizeof(int) * (full + 8)));
                                TR(kb_h2d(ctx, ctx->t_posmaps.p, T.posmaps, sizeof(int) * full));
                                ctx->stats.h2d_bytes += (double)(sizeof(int) * full);
                                ctx->posmaps_tag = nullptr;      // a caller-owned array may change behind our back
                                ctx->posmaps_n = 0;
                        }
                        d_posmaps = ctx->t_posmaps.as<int>();
                        std::vector<int> aoff((size_t)T.K);
                        for (int k = 0; k < T.K; k++) aoff[(size_t)k] = k * (maxlen + 1);
                        TR(ctx->t_aoff.ensure(sizeof(int) * (size_t)T.K + 16));
                        TR(kb_h2d(ctx, ctx->t_aoff.p, aoff.data(), sizeof(int) * (size_t)T.K));
                }
        }
        const size_t inv_per_task = (size_t)T.K * (size_t)(maxlen + 1);
        for (int L = 1; L <= maxlevel; L++) {
                const std::vector<int>& tl = by_level[(size_t)L];
                const int nt = (int)tl.size();
                if (nt == 0) continue;
                const bool trace = getenv("KB200_TRACE") != nullptr;
                auto tnow = []() { return std::chrono::steady_clock::now(); };
                auto tms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
                        return std::chrono::duration<double, std::milli>(b - a).count();
                };
                // KB200_TRACE: phase times need the stream drained at the phase boundaries (debugging aid)
                auto tsync = [&]() { if (trace) cudaStreamSynchronize(st); };
                const auto t_level0 = tnow();
                // ---- per task: scoring offset, operand lengths ----
                std::vector<float> soff((size_t)nt, 0.0f);
                // scaled gap penalties of the task (do_align, aln_run.c:227-237: gpo/gpe/tgpe *= gap_scale)
                std::vector<float> tgpo((size_t)nt, prm->gpo), tgpe_((size_t)nt, prm->gpe), ttgpe((size_t)nt, prm->tgpe);
                std::vector<int> la((size_t)nt), lb((size_t)nt);
                for (int q = 0; q < nt; q++) {
                        const int a = tasks_abc[3 * tl[q]], b = tasks_abc[3 * tl[q] + 1];
                        la[q] = (T.nsip[a] == 1) ? S.h_lens[a] : T.plen[a];
                        lb[q] = (T.nsip[b] == 1) ? S.h_lens[b] : T.plen[b];
                        // compute_subm_offset, aln_run.c:166-203
                        // compute_gap_scale, aln_run.c:126-164
                        if (prm->dist_scale > 0.0f && seq_distances) {
                                float sum = 0.0f;
                                int count = 0;
                                for (int si : T.sip[a]) { sum += seq_distances[si]; count++; }
                                for (int si : T.sip[b]) { sum += seq_distances[si]; count++; }
                                if (count) {
                                        const float avg_div = sum / (float)count;
                                        float scale = 1.0f - prm->dist_scale * avg_div;
                                        if (scale < 0.3f) scale = 0.3f;
                                        if (scale > 1.0f) scale = 1.0f;
                                        if (scale < 1.0f) {
                                                tgpo[q] = prm->gpo * scale;
                                                tgpe_[q] = prm->gpe * scale;
                                                ttgpe[q] = prm->tgpe * scale;
                                        }
                                }
                        }
                        const float amax = prm->vsm_amax;
                        if (amax > 0.0f && seq_distances) {
                                float sum = 0.0f;
                                int count = 0;
                                for (int si : T.sip[a]) { sum += seq_distances[si]; count++; }
                                for (int si : T.sip[b]) { sum += seq_distances[si]; count++; }
                                if (count) {
                                        const float avg = sum / (float)count;
                                        float off = amax - avg;
                                        if (off < 0.0f) off = 0.0f;
                                        soff[q] = off;
                                }
                        }
                }
                // ---- multi-GPU: contiguous shard of this level's task list, balanced by la*lb ----
                int q0 = 0, q1 = nt;
                std::vector<int> tb((size_t)ctx->world + 1, 0);
                {
                        std::vector<double> cost((size_t)nt);
                        for (int q = 0; q < nt; q++) cost[(size_t)q] = (double)la[q] * (double)lb[q] + 1.0;
                        kb_partition(cost.data(), nt, ctx->world, tb.data());
                        q0 = tb[(size_t)ctx->rank];
                        q1 = tb[(size_t)ctx->rank + 1];
                }
                const int ntm = q1 - q0;      // my tasks
                // ---- leaf profiles (make_profile_n with THIS task's offset) and gap rescale ----
                std::vector<KbLeafProfile> leaves;
                std::vector<long long> leaf_prefix;
                std::vector<KbGapSet> gsets;
                std::vector<long long> gs_prefix;
                long long leaf_cols = 0, gs_cols = 0;
                for (int q = q0; q < q1; q++) {
                        const int a = tasks_abc[3 * tl[q]], b = tasks_abc[3 * tl[q] + 1];
                        const int nodes[2] = {a, b};
                        const int other[2] = {b, a};
                        for (int s = 0; s < 2; s++) {
                                const int nd = nodes[s];
                                if (T.nsip[nd] == 1) {
                                        const int len = S.h_lens[nd];
                                        float* p = arena.alloc_floats((size_t)(len + 2) * 64);
                                        if (!p) { cleanup(); return KB200_FAIL; }
                                        T.prof[nd] = p;
                                        KbLeafProfile lp;
                                        lp.seq = S.dseq(nd); lp.prof = p; lp.len = len;
                                        lp.nsoff = -soff[q];
                                        lp.ngpo = -tgpo[q]; lp.ngpe = -tgpe_[q]; lp.ntgpe = -ttgpe[q];
                                        leaves.push_back(lp);
                                        leaf_prefix.push_back(leaf_cols);
                                        leaf_cols += len + 2;
                                } else {
                                        KbGapSet gs;
                                        gs.prof = T.prof[nd]; gs.len = T.plen[nd]; gs.nsip = T.nsip[other[s]];
                                        gsets.push_back(gs);
                                        gs_prefix.push_back(gs_cols);
                                        gs_cols += T.plen[nd] + 2;
                                }
                        }
                }
                TR(d_leaf.ensure(sizeof(KbLeafProfile) * leaves.size() + 16));
                TR(d_gapset.ensure(sizeof(KbGapSet) * gsets.size() + 16));
                TR(d_prefix.ensure(sizeof(long long) * (leaves.size() + gsets.size() + (size_t)nt) + 64));
                long long* d_pref_leaf = d_prefix.as<long long>();
                long long* d_pref_gs = d_pref_leaf + leaves.size();
                long long* d_pref_mg = d_pref_gs + gsets.size();
                if (!leaves.empty()) {
                        TR(kb_h2d(ctx, d_leaf.p, leaves.data(), sizeof(KbLeafProfile) * leaves.size()));
                        TR(kb_h2d(ctx, d_pref_leaf, leaf_prefix.data(), sizeof(long long) * leaves.size()));
                        TR(kb_make_profiles(ctx, d_leaf.as<KbLeafProfile>(), (int)leaves.size(), d_pref_leaf, leaf_cols, d_subm.as<float>()));
                }
                if (!gsets.empty()) {
                        TR(kb_h2d(ctx, d_gapset.p, gsets.data(), sizeof(KbGapSet) * gsets.size()));
                        TR(kb_h2d(ctx, d_pref_gs, gs_prefix.data(), sizeof(long long) * gsets.size()));
                        TR(kb_set_gap_penalties(ctx, d_gapset.as<KbGapSet>(), (int)gsets.size(), d_pref_gs, gs_cols));
                }

                // ---- jobs ----
                std::vector<KbJob> jobs((size_t)nt);
                std::vector<KbPathJob> pjobs((size_t)nt);
                std::vector<char> mirror((size_t)nt, 0);
                std::vector<int> rown((size_t)nt), coln((size_t)nt), rlen((size_t)nt), clen((size_t)nt);
                size_t n_raw = 0, n_coded = 0, n_scr = 0;
                for (int q = 0; q < nt; q++) {
                        const int a = tasks_abc[3 * tl[q]], b = tasks_abc[3 * tl[q] + 1];
                        const bool leaf_a = T.nsip[a] == 1, leaf_b = T.nsip[b] == 1;
                        KbJob j;
                        memset(&j, 0, sizeof(j));
                        j.nalpha = prm->nalpha;
                        // kernel choice and orientation, aln_run.c:297-388
                        if (leaf_a && leaf_b) {
                                j.kind = KB200_KIND_SS;
                                if (la[q] < lb[q]) { rown[q] = a; coln[q] = b; } else { rown[q] = b; coln[q] = a; mirror[q] = 1; }
                                j.seq_r = S.dseq(rown[q]); j.seq_c = S.dseq(coln[q]);
                                j.o = -tgpo[q]; j.e = -tgpe_[q]; j.t = -ttgpe[q];
                                j.nsoff = -soff[q];
                        } else if (leaf_a) {
                                j.kind = KB200_KIND_SP;
                                rown[q] = b; coln[q] = a; mirror[q] = 1;
                                j.prof_r = T.prof[b]; j.seq_c = S.dseq(a);
                                const float sipf = (float)T.nsip[b];
                                j.o = -(tgpo[q] * sipf); j.e = -(tgpe_[q] * sipf); j.t = -(ttgpe[q] * sipf);
                        } else if (leaf_b) {
                                j.kind = KB200_KIND_SP;
                                rown[q] = a; coln[q] = b;
                                j.prof_r = T.prof[a]; j.seq_c = S.dseq(b);
                                const float sipf = (float)T.nsip[a];
                                j.o = -(tgpo[q] * sipf); j.e = -(tgpe_[q] * sipf); j.t = -(ttgpe[q] * sipf);
                        } else {
                                j.kind = KB200_KIND_PP;
                                if (la[q] < lb[q]) { rown[q] = a; coln[q] = b; } else { rown[q] = b; coln[q] = a; mirror[q] = 1; }
                                j.prof_r = T.prof[rown[q]]; j.prof_c = T.prof[coln[q]];
                        }
                        rlen[q] = mirror[q] ? lb[q] : la[q];
                        clen[q] = mirror[q] ? la[q] : lb[q];
                        j.len_a = rlen[q];
                        j.len_b = clen[q];
                        jobs[(size_t)q] = j;
                        n_raw += (size_t)rlen[q] + 2;
                        n_coded += (size_t)la[q] + (size_t)lb[q] + 2;
                        n_scr += (size_t)la[q] + 2;
                }
                tsync();
                const auto t_prep = tnow();
                // ---- consistency bonus (default mode), dense on device ----
                if (T.posmaps && dev_state && ntm > 0) {
                        const int K = T.K;
                        std::vector<int> memb;
                        std::vector<KbBonusOperand> ops((size_t)2 * ntm);
                        std::vector<long long> op_prefix((size_t)2 * ntm), colb_prefix((size_t)ntm), row_prefix((size_t)ntm);
                        std::vector<KbBonusTask> btasks((size_t)ntm);
                        size_t pos_items = 0, sparse_items = 0;
                        long long op_cols = 0, colb_total = 0, row_total = 0;
                        for (int q = q0; q < q1; q++) {
                                pos_items += (size_t)K * ((size_t)rlen[q] + (size_t)clen[q]);
                                sparse_items += (size_t)K * (size_t)rlen[q];
                        }
                        TR(ctx->t_bpos.ensure(sizeof(int) * (pos_items + 16)));
                        TR(ctx->t_bconf.ensure(sizeof(float) * (pos_items + 16)));
                        TR(ctx->t_binv.ensure(sizeof(int) * (inv_per_task * (size_t)ntm + 16)));
                        TR(d_bidx.ensure(sizeof(int) * (sparse_items + 16)));
                        TR(d_bval.ensure(sizeof(float) * (sparse_items + 16)));
                        size_t po = 0, doff = 0;
                        for (int q = q0; q < q1; q++) {
                                const int nodes[2] = {rown[q], coln[q]};
                                const int lens2[2] = {rlen[q], clen[q]};
                                for (int sd = 0; sd < 2; sd++) {
                                        KbBonusOperand& O = ops[(size_t)2 * (q - q0) + sd];
                                        O.pos = ctx->t_bpos.as<int>() + po;
                                        O.conf = ctx->t_bconf.as<float>() + po;
                                        O.len = lens2[sd];
                                        O.m0 = (int)memb.size();
                                        memb.insert(memb.end(), T.sip[(size_t)nodes[sd]].begin(), T.sip[(size_t)nodes[sd]].end());
                                        O.m1 = (int)memb.size();
                                        op_prefix[(size_t)2 * (q - q0) + sd] = op_cols;
                                        op_cols += lens2[sd];
                                        po += (size_t)K * (size_t)lens2[sd];
                                }
                                KbBonusTask& B = btasks[(size_t)(q - q0)];
                                const KbBonusOperand& A = ops[(size_t)2 * (q - q0)];
                                const KbBonusOperand& Bo = ops[(size_t)2 * (q - q0) + 1];
                                B.pos_a = A.pos; B.conf_a = A.conf; B.len_a = A.len;
                                B.pos_b = Bo.pos; B.conf_b = Bo.conf; B.len_b = Bo.len;
                                B.inv = ctx->t_binv.as<int>() + inv_per_task * (size_t)(q - q0);
                                B.bcol = d_bidx.as<int>() + doff;
                                B.bval = d_bval.as<float>() + doff;
                                jobs[(size_t)q].bkey = B.bcol;
                                jobs[(size_t)q].bval = B.bval;
                                jobs[(size_t)q].nb = K;
                                doff += (size_t)K * (size_t)rlen[q];
                                colb_prefix[(size_t)(q - q0)] = colb_total; colb_total += clen[q];
                                row_prefix[(size_t)(q - q0)] = row_total; row_total += rlen[q];
                        }
                        // operands by member count: thread-per-column kernel (few members) / sliced warp kernels
                        std::vector<int> small_list;
                        std::vector<long long> small_prefix;
                        std::vector<KbVoteOp> vops;
                        long long small_cols = 0, large_units = 0, large_cols = 0, vote_slots = 0;
                        for (size_t o = 0; o < ops.size(); o++) {
                                const int nmem = ops[o].m1 - ops[o].m0;
                                if (nmem <= KB_BONUS_SMALL_NMEM) {
                                        small_list.push_back((int)o); small_prefix.push_back(small_cols); small_cols += ops[o].len;
                                } else {
                                        KbVoteOp v;
                                        v.op = (int)o;
                                        v.nchunks = (ops[o].len + 31) / 32;
                                        v.nslices = (nmem + KB_BONUS_VOTE_SLICE - 1) / KB_BONUS_VOTE_SLICE;
                                        v.pad = 0;
                                        v.unit0 = large_units; large_units += (long long)v.nchunks * v.nslices;
                                        v.vote0 = vote_slots; vote_slots += (long long)K * ops[o].len;
                                        v.col0 = large_cols; large_cols += ops[o].len;
                                        vops.push_back(v);
                                }
                        }
                        (void)op_cols;
                        TR(ctx->t_bdesc.ensure(sizeof(int) * (memb.size() + ops.size()) + sizeof(KbBonusOperand) * ops.size() + sizeof(KbBonusTask) * btasks.size() +
                                               sizeof(KbVoteOp) * vops.size() +
                                               sizeof(long long) * (op_prefix.size() + colb_prefix.size() + row_prefix.size()) + 256));
                        char* base = ctx->t_bdesc.as<char>();
                        KbBonusOperand* d_ops = (KbBonusOperand*)base; base += sizeof(KbBonusOperand) * ops.size();
                        KbBonusTask* d_bt = (KbBonusTask*)base; base += sizeof(KbBonusTask) * btasks.size();
                        KbVoteOp* d_vops = (KbVoteOp*)base; base += sizeof(KbVoteOp) * vops.size();
                        long long* d_smp = (long long*)base; base += sizeof(long long) * small_prefix.size();
                        long long* d_cbp = (long long*)base; base += sizeof(long long) * colb_prefix.size();
                        long long* d_rwp = (long long*)base; base += sizeof(long long) * row_prefix.size();
                        int* d_sml = (int*)base; base += sizeof(int) * small_list.size();
                        int* d_memb = (int*)base;
                        TR(kb_h2d(ctx, d_ops, ops.data(), sizeof(KbBonusOperand) * ops.size()));
                        TR(kb_h2d(ctx, d_bt, btasks.data(), sizeof(KbBonusTask) * btasks.size()));
                        if (!small_list.empty()) {
                                TR(kb_h2d(ctx, d_smp, small_prefix.data(), sizeof(long long) * small_prefix.size()));
                                TR(kb_h2d(ctx, d_sml, small_list.data(), sizeof(int) * small_list.size()));
                        }
                        if (!vops.empty()) {
                                TR(kb_h2d(ctx, d_vops, vops.data(), sizeof(KbVoteOp) * vops.size()));
                        }
                        TR(kb_h2d(ctx, d_cbp, colb_prefix.data(), sizeof(long long) * colb_prefix.size()));
                        TR(kb_h2d(ctx, d_rwp, row_prefix.data(), sizeof(long long) * row_prefix.size()));
                        TR(kb_h2d(ctx, d_memb, memb.data(), sizeof(int) * memb.size()));
                        TC(cudaMemsetAsync(ctx->t_binv.p, 0xFF, sizeof(int) * inv_per_task * (size_t)ntm, st));
                        TR(kb_bonus_level(ctx, S, K, T.weight / (float)K, d_ops,
                                          d_sml, d_smp, (int)small_list.size(), small_cols,
                                          d_vops, (int)vops.size(), large_units, large_cols, vote_slots,
                                          d_memb, d_colof, d_posmaps,
                                          d_bt, d_cbp, colb_total, d_rwp, row_total, ntm, ctx->t_aoff.as<int>()));
                } else if (T.posmaps && !dev_state) {
                        level_bonus(T, q0, q1, rown, rlen, coln, clen, n_threads, colof, bonus_lists);
                        size_t dense = 0, nent = 0;
                        for (int q = q0; q < q1; q++) {
                                dense += (size_t)rlen[q] * (size_t)clen[q];
                                nent += bonus_lists[(size_t)q].size();
                        }
                        TR(d_bonus.ensure(sizeof(float) * (dense + 16)));
                        TR(d_bidx.ensure(sizeof(long long) * (nent + 16)));
                        TR(d_bval.ensure(sizeof(float) * (nent + 16)));
                        std::vector<long long> hidx(nent);
                        std::vector<float> hval(nent);
                        size_t off = 0, e = 0;
                        for (int q = q0; q < q1; q++) {
                                jobs[(size_t)q].bonus = d_bonus.as<float>() + off;
                                for (const auto& pr : bonus_lists[(size_t)q]) {
                                        hidx[e] = (long long)off + pr.first;
                                        hval[e] = pr.second;
                                        e++;
                                }
                                off += (size_t)rlen[q] * (size_t)clen[q];
                        }
                        TC(cudaMemsetAsync(d_bonus.p, 0, sizeof(float) * dense, st));
                        if (nent) {
                                TR(kb_h2d(ctx, d_bidx.p, hidx.data(), sizeof(long long) * nent));
                                TR(kb_h2d(ctx, d_bval.p, hval.data(), sizeof(float) * nent));
                                kb_scatter_kernel<<<(unsigned)((nent + 255) / 256), 256, 0, st>>>(d_bonus.as<float>(), d_bidx.as<long long>(), d_bval.as<float>(), (long long)nent);
                                TC(cudaGetLastError());
                                ctx->stats.n_launches++;
                        }
                        TC(cudaStreamSynchronize(st));
                        ctx->stats.h2d_bytes += 12.0 * (double)nent;
                }
                tsync();
                const auto t_bonus = tnow();
                TR(d_raw.ensure(sizeof(int) * (n_raw + 16)));
                TR(d_coded.ensure(sizeof(int) * (n_coded + 16)));
                TR(d_scr.ensure(sizeof(int) * (n_scr + 16)));
                {
                        size_t o_raw = 0, o_coded = 0, o_scr = 0;
                        for (int q = 0; q < nt; q++) {
                                jobs[(size_t)q].path = d_raw.as<int>() + o_raw;
                                KbPathJob pj;
                                pj.raw = d_raw.as<int>() + o_raw;
                                pj.coded = d_coded.as<int>() + o_coded;
                                pj.scratch = d_scr.as<int>() + o_scr;
                                pj.posmap = nullptr;
                                pj.len_a = la[q]; pj.len_b = lb[q];
                                pj.mirror = mirror[q];
                                pjobs[(size_t)q] = pj;
                                o_raw += (size_t)rlen[q] + 2;
                                o_coded += (size_t)la[q] + (size_t)lb[q] + 2;
                                o_scr += (size_t)la[q] + 2;
                        }
                }
                std::vector<size_t> coded_off((size_t)nt);
                {
                        size_t o = 0;
                        for (int q = 0; q < nt; q++) {
                                coded_off[(size_t)q] = o;
                                o += (size_t)la[q] + (size_t)lb[q] + 2;
                        }
                }
                TC(cudaMemsetAsync(d_raw.p, 0xFF, sizeof(int) * n_raw, st));
                if (conf_out && ntm > 0) {
                        // per-box meet-up margins of my tasks (task->confidence, aln_run.c:390-394)
                        size_t nm = 0;
                        for (int q = q0; q < q1; q++) nm += kb_margin_cap(rlen[q]);
                        TR(ctx->t_margin.ensure(sizeof(float) * (nm + 16)));
                        TC(cudaMemsetAsync(ctx->t_margin.p, 0xFF, sizeof(float) * nm, st));
                        size_t o = 0;
                        for (int q = q0; q < q1; q++) {
                                jobs[(size_t)q].margins = ctx->t_margin.as<float>() + o;
                                jobs[(size_t)q].margin_cap = kb_margin_cap(rlen[q]);
                                o += jobs[(size_t)q].margin_cap;
                        }
                }
                {
                        std::vector<KbJob> mine(jobs.begin() + q0, jobs.begin() + q1);
                        TR(kb_run_hirschberg(ctx, prm->subm, mine));
                }
                if (conf_out) {
                        TR(ctx->t_conf.ensure(sizeof(float) * ((size_t)nt + 16)));
                        if (ntm > 0) TR(kb_confidences(ctx, ntm, ctx->t_conf.as<float>() + q0));
                        if (ctx->world > 1) {
                                std::vector<size_t> seg((size_t)ctx->world + 1, 0);
                                for (int r = 0; r <= ctx->world; r++) seg[(size_t)r] = (size_t)tb[(size_t)r] * sizeof(float);
                                TR(kb_allgatherv(ctx, ctx->t_conf.p, seg.data()));
                        }
                }
                tsync();
                const auto t_dp = tnow();
                TR(d_pjobs.ensure(sizeof(KbPathJob) * (size_t)nt + 16));
                if (ntm > 0) {
                        TR(kb_h2d(ctx, d_pjobs.p, pjobs.data() + q0, sizeof(KbPathJob) * (size_t)ntm));
                        TR(kb_code_paths(ctx, d_pjobs.as<KbPathJob>(), ntm));
                }
                if (ctx->world > 1) {
                        // coded paths of every rank's shard (segments are contiguous in task order)
                        std::vector<size_t> seg((size_t)ctx->world + 1, 0);
                        size_t o = 0;
                        int r = 0;
                        for (int q = 0; q <= nt; q++) {
                                while (r <= ctx->world && q == tb[(size_t)r]) { seg[(size_t)r] = o * sizeof(int); r++; }
                                if (q < nt) o += (size_t)la[q] + (size_t)lb[q] + 2;
                        }
                        TR(kb_allgatherv(ctx, d_coded.p, seg.data()));
                }
                if (dev_state) {
                        // gap weaving of EVERY task of the level on the device (all ranks keep the full state)
                        std::vector<KbWeaveTask> wt((size_t)nt);
                        std::vector<KbWeaveMember> wm;
                        size_t pneed = 0;
                        for (int q = 0; q < nt; q++) pneed += 2 * ((size_t)la[q] + (size_t)lb[q] + 4);
                        TR(ctx->t_wp.ensure(sizeof(int) * (pneed + 16)));
                        TR(ctx->t_alen.ensure(sizeof(int) * ((size_t)nt + 16)));
                        size_t po = 0;
                        for (int q = 0; q < nt; q++) {
                                const int t = tl[q];
                                KbWeaveTask& W = wt[(size_t)q];
                                W.path = d_coded.as<int>() + coded_off[(size_t)q];
                                W.Pa = ctx->t_wp.as<int>() + po; po += (size_t)la[q] + (size_t)lb[q] + 4;
                                W.Pb = ctx->t_wp.as<int>() + po; po += (size_t)la[q] + (size_t)lb[q] + 4;
                                W.alnlen = -1;
                                W.out_len = ctx->t_alen.as<int>() + q;
                                for (int si : T.sip[(size_t)tasks_abc[3 * t]]) wm.push_back({si, W.Pa});
                                for (int si : T.sip[(size_t)tasks_abc[3 * t + 1]]) wm.push_back({si, W.Pb});
                        }
                        TR(ctx->t_wdesc.ensure(sizeof(KbWeaveTask) * wt.size() + sizeof(KbWeaveMember) * wm.size() + 64));
                        KbWeaveTask* d_wt = ctx->t_wdesc.as<KbWeaveTask>();
                        KbWeaveMember* d_wm = (KbWeaveMember*)(d_wt + wt.size());
                        TR(kb_h2d(ctx, d_wt, wt.data(), sizeof(KbWeaveTask) * wt.size()));
                        TR(kb_h2d(ctx, d_wm, wm.data(), sizeof(KbWeaveMember) * wm.size()));
                        TR(kb_weave_level(ctx, S, d_wt, nt, d_wm, (int)wm.size(), d_gaps, d_colof));
                }
                // the host only needs every task's alignment length (plen bookkeeping, merged profile
                // sizes); the coded paths themselves stay on the device (KB200_HOST_BONUS: full copy)
                std::vector<int> hcoded;
                std::vector<int> alen((size_t)nt);
                if (dev_state) {
                        // THE host synchronisation of the level: alignment lengths + engine statistics / error flags
                        TC(cudaMemcpyAsync(alen.data(), ctx->t_alen.p, sizeof(int) * (size_t)nt, cudaMemcpyDeviceToHost, st));
                        std::vector<float> lconf;
                        if (conf_out) {
                                lconf.resize((size_t)nt);
                                TC(cudaMemcpyAsync(lconf.data(), ctx->t_conf.p, sizeof(float) * (size_t)nt, cudaMemcpyDeviceToHost, st));
                        }
                        TR(kb_collect(ctx));
                        if (conf_out) {
                                for (int q = 0; q < nt; q++) conf_out[tl[(size_t)q]] = lconf[(size_t)q];
                        }
                        ctx->stats.d2h_bytes += (double)(sizeof(int) * (size_t)nt);
                } else {
                        hcoded.resize(n_coded);
                        TC(cudaMemcpyAsync(hcoded.data(), d_coded.p, sizeof(int) * n_coded, cudaMemcpyDeviceToHost, st));
                        std::vector<float> lconf;
                        if (conf_out) {
                                lconf.resize((size_t)nt);
                                TC(cudaMemcpyAsync(lconf.data(), ctx->t_conf.p, sizeof(float) * (size_t)nt, cudaMemcpyDeviceToHost, st));
                        }
                        TR(kb_collect(ctx));
                        if (conf_out) {
                                for (int q = 0; q < nt; q++) conf_out[tl[(size_t)q]] = lconf[(size_t)q];
                        }
                        ctx->stats.d2h_bytes += (double)(sizeof(int) * n_coded);
                        for (int q = 0; q < nt; q++) alen[(size_t)q] = hcoded[coded_off[(size_t)q]];
                }
                // ---- merge profiles on device (update_n), skipped for the root task ----
                std::vector<KbMergeJob> mjobs;
                std::vector<long long> mprefix;
                long long mcols = 0;
                size_t nsrc = 0;
                std::vector<size_t> prof_off((size_t)nt + 1, 0);     // float offsets inside the level block
                for (int q = 0; q < nt; q++) {
                        size_t w = 0;
                        if (tl[q] != ntasks - 1) {
                                w = ((size_t)alen[(size_t)q] + 2) * 64;
                                if (q >= q0 && q < q1) nsrc += (size_t)alen[(size_t)q] + 2;
                        }
                        prof_off[(size_t)q + 1] = prof_off[(size_t)q] + w;
                }
                TR(d_src.ensure(sizeof(int2) * (nsrc + 16)));
                float* block = nullptr;
                if (prof_off[(size_t)nt] > 0) {
                        block = arena.alloc_floats(prof_off[(size_t)nt]);
                        if (!block) { cleanup(); return KB200_FAIL; }
                }
                {
                        size_t osrc = 0;
                        for (int q = 0; q < nt; q++) {
                                const int t = tl[q];
                                const int a = tasks_abc[3 * t], b = tasks_abc[3 * t + 1], c = tasks_abc[3 * t + 2];
                                const int alnlen = alen[(size_t)q];
                                if (t == ntasks - 1) continue;
                                T.prof[c] = block + prof_off[(size_t)q];
                                if (q < q0 || q >= q1) continue;
                                KbMergeJob m;
                                m.pa = T.prof[a]; m.pb = T.prof[b]; m.newp = T.prof[c];
                                m.path = d_coded.as<int>() + coded_off[(size_t)q];
                                m.src = d_src.as<int2>() + osrc;
                                m.alnlen = alnlen;
                                m.sipa = T.nsip[a]; m.sipb = T.nsip[b];
                                m.gpo = prm->gpo; m.gpe = prm->gpe; m.tgpe = prm->tgpe;      // update_n runs with the UNSCALED ap (aln_run.c:407)
                                m.rebalance = 0; m.scaleA = 1.0f; m.scaleB = 1.0f; m.subm = d_subm.as<float>();
                                if (prm->use_seq_weights > 0.0f && m.sipa > 0 && m.sipb > 0) {
                                        // aln_setup.c:252-259
                                        const float pseudo = prm->use_seq_weights;
                                        const float total = (float)(m.sipa + m.sipb);
                                        const float denom = total + 2.0f * pseudo;
                                        m.scaleA = total * ((float)m.sipa + pseudo) / (denom * (float)m.sipa);
                                        m.scaleB = total * ((float)m.sipb + pseudo) / (denom * (float)m.sipb);
                                        m.rebalance = 1;
                                }
                                mjobs.push_back(m);
                                mprefix.push_back(mcols);
                                mcols += alnlen + 2;
                                osrc += (size_t)alnlen + 2;
                        }
                }
                if (!mjobs.empty()) {
                        TR(d_mjobs.ensure(sizeof(KbMergeJob) * mjobs.size()));
                        TR(kb_h2d(ctx, d_mjobs.p, mjobs.data(), sizeof(KbMergeJob) * mjobs.size()));
                        TR(kb_h2d(ctx, d_pref_mg, mprefix.data(), sizeof(long long) * mprefix.size()));
                        TR(kb_merge_index(ctx, d_mjobs.as<KbMergeJob>(), (int)mjobs.size()));
                        TR(kb_merge_profiles(ctx, d_mjobs.as<KbMergeJob>(), (int)mjobs.size(), d_pref_mg, mcols));
                }
                if (ctx->world > 1 && block) {
                        // the single all-gather of merged sub-profiles of this level (NCCL over NVLink)
                        std::vector<size_t> seg((size_t)ctx->world + 1, 0);
                        for (int r = 0; r <= ctx->world; r++) seg[(size_t)r] = prof_off[(size_t)tb[(size_t)r]] * sizeof(float);
                        TR(kb_allgatherv(ctx, block, seg.data()));
                }
                tsync();
                const auto t_post = tnow();
                // ---- host bookkeeping while the merge kernels run: gaps, sip, nsip, plen ----
                if (!dev_state) {
                        level_weave(T, nt, tl, tasks_abc, hcoded.data(), coded_off, n_threads);
                }
                for (int q = 0; q < nt; q++) {
                        const int t = tl[q];
                        const int a = tasks_abc[3 * t], b = tasks_abc[3 * t + 1], c = tasks_abc[3 * t + 2];
                        T.plen[c] = alen[(size_t)q];
                        T.nsip[c] = T.nsip[a] + T.nsip[b];
                        std::vector<int>& sc = T.sip[c];
                        sc.clear();
                        sc.reserve((size_t)T.nsip[c]);
                        for (int j = T.nsip[a]; j--;) sc.push_back(T.sip[a][(size_t)j]);     // aln_run.c:428-436
                        for (int j = T.nsip[b]; j--;) sc.push_back(T.sip[b][(size_t)j]);
                        if (!T.posmaps) {
                                std::vector<int>().swap(T.sip[a]);   // children are dead; keep memory bounded
                                std::vector<int>().swap(T.sip[b]);
                        } else {
                                std::vector<int>().swap(T.sip[a]);
                                std::vector<int>().swap(T.sip[b]);
                        }
                }
                if (trace) {
                        TC(cudaStreamSynchronize(st));
                        const auto t_end = tnow();
                        fprintf(stderr, "[kb200 trace] tree level %d: %d tasks prep %.2f bonus %.2f dp %.2f post %.2f weave %.2f ms\n", L, nt,
                                tms(t_level0, t_prep), tms(t_prep, t_bonus), tms(t_bonus, t_dp), tms(t_dp, t_post), tms(t_post, t_end));
                }
        }
        if (dev_state) {
                // gaps_out == nullptr: the caller keeps working on the device copy (ctx->t_gaps / t_colof)
                if (gaps_out) {
                        TC(cudaMemcpyAsync(gaps_out, d_gaps, sizeof(int) * ((size_t)S.total + (size_t)N), cudaMemcpyDeviceToHost, st));
                        ctx->stats.d2h_bytes += (double)(sizeof(int) * ((size_t)S.total + (size_t)N));
                }
                TR(kb_collect(ctx));
        } else if (gaps_out) {
                for (int i = 0; i < N; i++) {
                        memcpy(gaps_out + S.h_offs[i] + i, T.gaps[i].data(), sizeof(int) * ((size_t)S.h_lens[i] + 1));
                }
        }
        if (plen_out) {
                for (int t = 0; t < ntasks; t++) plen_out[t] = T.plen[(size_t)tasks_abc[3 * t + 2]];
        }
        cleanup();
        (void)rc;
        return KB200_OK;
#undef TR
#undef TC
}

extern "C" int kb200_align_tree(kb200_ctx* ctx, const kb200_params* prm,
                                const uint8_t* seqs, const int64_t* offs, const int* lens, int nseq,
                                const int* tasks_abc, int ntasks, const float* seq_distances,
                                const int* posmaps, int K, float weight, int* gaps_out)
{
        return kb200_align_tree_conf(ctx, prm, seqs, offs, lens, nseq, tasks_abc, ntasks, seq_distances, posmaps, K, weight, gaps_out, nullptr, nullptr);
}

extern "C" int kb200_align_tree_conf(kb200_ctx* ctx, const kb200_params* prm,
                                     const uint8_t* seqs, const int64_t* offs, const int* lens, int nseq,
                                     const int* tasks_abc, int ntasks, const float* seq_distances,
                                     const int* posmaps, int K, float weight, int* gaps_out, float* task_confidence, int* task_plen)
{
        if (!ctx || !prm || !seqs || !offs || !lens || !tasks_abc || !gaps_out || nseq < 2) {
                return KB200_FAIL;
        }
        KB_CUDA(cudaSetDevice(ctx->device));
        KbSeqs S;
        int rc = S.upload(ctx, seqs, offs, lens, nseq);
        if (rc == KB200_OK) {
                const int nthr = kb_default_threads();
                rc = kb_align_tree_dev(ctx, prm, S, tasks_abc, ntasks, seq_distances, posmaps, K, weight, nthr, gaps_out, 0, task_confidence, task_plen);
        }
        S.release();
        return rc;
}
