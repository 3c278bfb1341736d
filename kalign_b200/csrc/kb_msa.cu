// kb_msa.cu -- the kalign() / kalign_run_seeded() call sequence on plain arrays, with the three
// hot-path seams running on the GPU.
//
// Host-side mirror of (behaviour cited, nothing copied):
//   kalign / kalign_run_seeded          lib/src/aln_wrap.c:110,133
//   kalign_arr_to_msa / detect_alphabet lib/src/msa_op.c:440,142
//   kalign_essential_input_check        lib/src/msa_check.c:66
//   msa_sort_len_name / msa_sort_rank   lib/src/msa_sort.c:14,25
//   create_alphabet (+ merge/clean)     lib/src/alphabet.c:66-437
//   pick_anchor / select_seqs           lib/src/pick_anchor.c:17,33
//   build_tree_kmeans / bisecting_kmeans / split2 / upgma / label_internal / create_tasks
//                                       lib/src/bisectingKmeans.c:177,273,766,974,1067,1084
//   edist_256 (AVX lane order)          lib/src/euclidean_dist.c:161-206
//   select_anchors                      lib/src/anchor_consistency.c:124-198
//   finalise_alignment                  lib/src/msa_op.c:546
// The guide tree is still host code (SURVEY.md section 8f-1 "next"); its two distance-matrix
// calls (d_estimation pair=0 / pair=1) run on the GPU as two batched bpm launches.
#include "kb_host.cuh"
#include "kb_kmeans.h"
#include "kb_kmeans_dev.h"

#include <math.h>
#include <time.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <thread>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// host-stage timing for KB200_TRACE=1
inline double kb_now()
{
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
inline bool kb_trace_on()
{
        static const bool on = getenv("KB200_TRACE") != nullptr;
        return on;
}

// ---- alphabets (alphabet.c) ------------------------------------------------------------------
enum { ALPHA_DNA = 5, ALPHA_RED = 13, ALPHA_AMB = 23 };

struct Alphabet {
        int8_t to_internal[128];
        int L;
};

void merge2(Alphabet& a, int x, int y)
{
        const int8_t m = std::min(a.to_internal[x], a.to_internal[y]);
        a.to_internal[x] = m;
        a.to_internal[y] = m;
}

// compress the codes that are in use to 0..L-1 in increasing order and mirror to lower case
// (clean_and_set_to_extern, alphabet.c:393-437)
void finish(Alphabet& a)
{
        int8_t trans[32];
        for (int i = 0; i < 32; i++) trans[i] = -1;
        for (int i = 64; i < 96; i++) {
                if (a.to_internal[i] != -1) trans[a.to_internal[i]] = 1;
        }
        int code = 0;
        for (int i = 0; i < 32; i++) {
                if (trans[i] == 1) trans[i] = (int8_t)code++;
        }
        a.L = code;
        for (int i = 64; i < 96; i++) {
                if (a.to_internal[i] != -1) {
                        a.to_internal[i] = trans[a.to_internal[i]];
                        a.to_internal[i + 32] = a.to_internal[i];
                }
        }
}

Alphabet make_alphabet(int type)
{
        Alphabet a;
        for (int i = 0; i < 128; i++) a.to_internal[i] = -1;
        a.L = 0;
        if (type == ALPHA_AMB) {
                const char* aa = "ARNDCQEGHILKMFPSTWYVBZX";                       // alphabet.c:179-203
                for (int i = 0; i < 23; i++) a.to_internal[(int)aa[i]] = (int8_t)i;
                a.to_internal[(int)'U'] = 22;
        } else if (type == ALPHA_DNA) {
                const char* nt = "ACGTUNRYSWKMBDHV";                              // alphabet.c:206-245
                for (int i = 0; i < 16; i++) a.to_internal[(int)nt[i]] = (int8_t)i;
                merge2(a, 'U', 'T');
                const char* amb = "RYSWKMBDHV";
                for (int i = 0; i < 10; i++) merge2(a, 'N', amb[i]);
        } else {
                const char* aa = "ACDEFGHIKLMNPQRSTVWY";                          // alphabet.c:248-302
                for (int i = 0; i < 20; i++) a.to_internal[(int)aa[i]] = (int8_t)i;
                a.to_internal[(int)'B'] = 20;
                a.to_internal[(int)'Z'] = 21;
                a.to_internal[(int)'X'] = 22;
                merge2(a, 'L', 'M'); merge2(a, 'I', 'V'); merge2(a, 'K', 'R'); merge2(a, 'E', 'Q');
                merge2(a, 'A', 'S'); merge2(a, 'A', 'T'); merge2(a, 'S', 'T'); merge2(a, 'N', 'D');
                merge2(a, 'F', 'Y'); merge2(a, 'B', 'N'); merge2(a, 'B', 'D'); merge2(a, 'Z', 'E');
                merge2(a, 'Z', 'Q');
                a.to_internal[(int)'U'] = a.to_internal[(int)'C'];
        }
        finish(a);
        return a;
}

// ---- sequences -------------------------------------------------------------------------------
struct Seq {
        const char* seq;
        std::string name;
        int len;
        int rank;
};

int cmp_len_name(const void* a, const void* b)       // msa_sort.c:62-81
{
        const Seq* one = *(Seq* const*)a;
        const Seq* two = *(Seq* const*)b;
        if (one->len > two->len) return -1;
        if (one->len == two->len) {
                return strncmp(one->name.c_str(), two->name.c_str(), 256) < 0 ? -1 : 1;
        }
        return 1;
}

struct LenId {
        int len;
        int id;
};

int cmp_len_only(const void* a, const void* b)       // pick_anchor.c:74-84 (never returns 0)
{
        const LenId* one = *(LenId* const*)a;
        const LenId* two = *(LenId* const*)b;
        return (one->len > two->len) ? -1 : 1;
}

// ---- guide tree: data structures, split2, bisect -> kb_kmeans.h (plain C++, CPU-testable) ----

// upgma, bisectingKmeans.c:974-1053.  dm: n x n (modified in place).  Returns sub-tree root.
// The sub-tree's 2n-1 nodes are written to B.nodes[base ..] (reserved by the caller), so that the
// leaf clusters can be processed concurrently.
int upgma(TreeBuilder& B, float* dm, const std::vector<int>& samples, const int base)
{
        const int n = (int)samples.size();
        int next_node = base;
        std::vector<int> as(n), tree(n);
        for (int i = 0; i < n; i++) {
                as[i] = i + 1;
                Node nd;
                nd.id = samples[i];
                tree[i] = next_node;
                B.nodes[(size_t)next_node++] = nd;
        }
        int node_a = 0, node_b = 0;
        for (int cnode = n; cnode != 2 * n - 1; cnode++) {
                float mn = FLT_MAX;
                for (int i = 0; i < n - 1; i++) {
                        if (!as[i]) continue;
                        for (int j = i + 1; j < n; j++) {
                                if (as[j] && dm[i * n + j] < mn) {
                                        mn = dm[i * n + j];
                                        node_a = i;
                                        node_b = j;
                                }
                        }
                }
                Node nd;
                nd.left = tree[node_a];
                nd.right = tree[node_b];
                tree[node_a] = next_node;
                B.nodes[(size_t)next_node++] = nd;
                tree[node_b] = -1;
                as[node_a] = cnode + 1;
                as[node_b] = 0;
                for (int j = n; j--;) {
                        if (j != node_b) {
                                dm[node_a * n + j] = (dm[node_a * n + j] + dm[node_b * n + j]) * 0.5F + 0.001F;
                        }
                }
                dm[node_a * n + node_a] = 0.0F;
                for (int j = n; j--;) {
                        dm[j * n + node_a] = dm[node_a * n + j];
                }
        }
        return tree[node_a];
}

// label_internal (post-order labels) + create_tasks, then sorted by c == post-order of internals
void emit_tasks(const TreeBuilder& B, int root, int N, std::vector<int>& abc)
{
        // iterative post-order
        std::vector<int> label(B.nodes.size(), -1);
        std::vector<std::pair<int, int>> stack;
        int next = N;
        stack.emplace_back(root, 0);
        while (!stack.empty()) {
                auto& top = stack.back();
                const Node& nd = B.nodes[(size_t)top.first];
                if (nd.left < 0 && nd.right < 0) {
                        label[(size_t)top.first] = nd.id;
                        stack.pop_back();
                        continue;
                }
                if (top.second == 0) {
                        top.second = 1;
                        stack.emplace_back(nd.left, 0);
                } else if (top.second == 1) {
                        top.second = 2;
                        stack.emplace_back(nd.right, 0);
                } else {
                        const int me = top.first;
                        label[(size_t)me] = next++;
                        abc.push_back(label[(size_t)nd.left]);
                        abc.push_back(label[(size_t)nd.right]);
                        abc.push_back(label[(size_t)me]);
                        stack.pop_back();
                }
        }
}

// select_anchors, anchor_consistency.c:124-198
void select_anchors(const float* sd, int N, int K, std::vector<int>& ids)
{
        ids.assign((size_t)K, 0);
        std::vector<float> min_dist((size_t)N);
        float sum = 0.0f;
        for (int i = 0; i < N; i++) sum += sd[i];
        const float mean = sum / (float)N;
        float best_diff = FLT_MAX;
        int best_idx = 0;
        for (int i = 0; i < N; i++) {
                float diff = sd[i] - mean;
                if (diff < 0) diff = -diff;
                if (diff < best_diff) { best_diff = diff; best_idx = i; }
        }
        ids[0] = best_idx;
        for (int i = 0; i < N; i++) {
                float d = sd[i] - sd[ids[0]];
                if (d < 0) d = -d;
                min_dist[(size_t)i] = d;
        }
        for (int k = 1; k < K; k++) {
                float best_min = -1.0f;
                int bi = 0;
                for (int i = 0; i < N; i++) {
                        bool skip = false;
                        for (int j = 0; j < k; j++) {
                                if (ids[(size_t)j] == i) { skip = true; break; }
                        }
                        if (skip) continue;
                        if (min_dist[(size_t)i] > best_min) { best_min = min_dist[(size_t)i]; bi = i; }
                }
                ids[(size_t)k] = bi;
                for (int i = 0; i < N; i++) {
                        float d = sd[i] - sd[bi];
                        if (d < 0) d = -d;
                        if (d < min_dist[(size_t)i]) min_dist[(size_t)i] = d;
                }
        }
}

} // namespace

// Guide tree on device-resident tree-alphabet sequences: tasks (a,b,c) sorted by c and
// msa->seq_distances (build_tree_kmeans, bisectingKmeans.c:177-271).
// The guide tree in three stages, so that the host-only middle one can run beside GPU work:
//   prepare : anchors, N x 32 distance matrix (GPU), msa->seq_distances
//   bisect  : bisecting k-means over the distance rows (host, OpenMP tasks)
//   finish  : leaf-cluster distances (GPU), UPGMA per cluster, task list
struct KbTreeJob {
        TreeBuilder B;
        std::vector<float> dm;
        int num_anchor = 0;
        int stride = 0;
        int root = -1;
        uint64_t tree_seed = 0;          // build_tree_kmeans_noisy (bisectingKmeans.c:76): seed != 0 && noise > 0
        float tree_noise = 0.0f;
};

// The multiplicative noise build_tree_kmeans_noisy puts on the N x 32 anchor distances
// (bisectingKmeans.c:104-116): factor = max(0.1, gaussian(1, sigma)) drawn row by row from the reference's
// generator (lib/src/tlrng.c): xoshiro256** (public domain, Blackman & Vigna) seeded by four splitmix64
// steps (init_rng :218-271), uniform = next / 2^64 redrawn while 0 (tl_random_double :87), Box-Muller with
// the second variate kept for the next call (tl_random_gaussian :105-124).  Same libm, same operation
// order: the factors are bit-identical, and so is everything that follows.
struct KbNoise {
        uint64_t s[4];
        bool gen = false;
        double z1 = 0.0;
        static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
        explicit KbNoise(uint64_t seed)
        {
                bool ok = false;
                while (!ok) {
                        for (int i = 0; i < 4; i++) {
                                uint64_t z = (seed += 0x9e3779b97f4a7c15ULL);
                                z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
                                z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
                                s[i] = z ^ (z >> 31);
                                if (s[i]) ok = true;
                        }
                }
        }
        uint64_t next()
        {
                const uint64_t r = rotl(s[1] * 5, 7) * 9;
                const uint64_t t = s[1] << 17;
                s[2] ^= s[0];
                s[3] ^= s[1];
                s[1] ^= s[2];
                s[0] ^= s[3];
                s[2] ^= t;
                s[3] = rotl(s[3], 45);
                return r;
        }
        double uniform()
        {
                double y;
                do {
                        y = (double)next() / 18446744073709551616.0;
                } while (y == 0.0);
                return y;
        }
        double gaussian(double mu, double sigma)
        {
                gen = !gen;
                if (!gen) {
                        return z1 * sigma + mu;
                }
                double u1, u2;
                do {
                        u1 = uniform();
                        u2 = uniform();
                } while (u1 <= DBL_EPSILON);
                const double z0 = sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
                z1 = sqrt(-2.0 * log(u1)) * sin(2.0 * M_PI * u2);
                return z0 * sigma + mu;
        }
        float factor(float sigma)
        {
                double n = gaussian(1.0, (double)sigma);
                if (n < 0.1) n = 0.1;
                return (float)n;
        }
};

int kb_tree_prepare(kb200_ctx* ctx, KbSeqs& S, KbTreeJob& T, std::vector<float>& seq_distances)
{
        const int N = S.n;
        const double tt0 = kb_now();
        TreeBuilder& B = T.B;
        B.N = N;
        // pick_anchor / select_seqs (pick_anchor.c:17-72)
        const int num_anchor = std::min(32, N);
        std::vector<int> anchors((size_t)num_anchor);
        {
                std::vector<LenId> items((size_t)N);
                std::vector<LenId*> ptrs((size_t)N);
                for (int i = 0; i < N; i++) {
                        items[(size_t)i].id = i;
                        items[(size_t)i].len = S.h_lens[i];
                        ptrs[(size_t)i] = &items[(size_t)i];
                }
                qsort(ptrs.data(), (size_t)N, sizeof(LenId*), cmp_len_only);
                const int stride = N / num_anchor;
                for (int i = 0; i < num_anchor; i++) anchors[(size_t)i] = ptrs[(size_t)(i * stride)]->id;
        }
        // d_estimation(pair = 0): N x num_anchor, rows padded to a multiple of 8 with zeros
        const int stride = ((num_anchor + 7) / 8) * 8;
        T.dm.assign((size_t)N * stride, 0.0f);
        {
                std::vector<int> rows((size_t)N);
                for (int i = 0; i < N; i++) rows[(size_t)i] = i;
                std::vector<float> tmp((size_t)N * num_anchor);
                KB_RUN(kb_distances_dev(ctx, S, rows.data(), N, anchors.data(), num_anchor, 0, tmp.data()));
                for (int i = 0; i < N; i++) {
                        memcpy(T.dm.data() + (size_t)i * stride, tmp.data() + (size_t)i * num_anchor, sizeof(float) * (size_t)num_anchor);
                }
        }
        if (T.tree_seed != 0 && T.tree_noise > 0.0f) {
                KbNoise rng(T.tree_seed);
                for (int i = 0; i < N; i++) {
                        for (int j = 0; j < num_anchor; j++) {
                                T.dm[(size_t)i * stride + j] *= rng.factor(T.tree_noise);
                        }
                }
        }
        T.num_anchor = num_anchor;
        T.stride = stride;
        B.dm = T.dm.data();
        B.stride = stride;
        B.num_anchors = num_anchor;
        B.nodes.reserve((size_t)2 * N + 64);
        // msa->seq_distances (bisectingKmeans.c:247-256)
        seq_distances.resize((size_t)N);
        for (int i = 0; i < N; i++) {
                float sum = 0.0f;
                for (int j = 0; j < num_anchor; j++) sum += T.dm[(size_t)i * stride + j];
                const float mean_dist = sum / (float)num_anchor;
                const float seq_len = (float)S.h_lens[i];
                seq_distances[(size_t)i] = (seq_len > 0.0f) ? mean_dist / seq_len : 0.0f;
        }
        if (kb_trace_on()) {
                fprintf(stderr, "[kb200 trace] guide tree: anchor distances %.1f ms\n", 1e3 * (kb_now() - tt0));
        }
        return KB200_OK;
}

// bisecting k-means on the device (kb_kmeans.cu); KB200_HOST_KMEANS=1 keeps the host restatement
// (kb_kmeans.h) for A/B checks
void kb_tree_bisect(KbTreeJob& T, int n_threads);

int kb_tree_bisect_any(kb200_ctx* ctx, KbTreeJob& T, int n_threads)
{
        if (getenv("KB200_HOST_KMEANS") != nullptr || T.num_anchor != 32 || T.stride != 32) {
                kb_tree_bisect(T, n_threads);
                return KB200_OK;
        }
        const double tt1 = kb_now();
        KbKmeansTree K;
        KB_RUN(kb_kmeans_bisect_dev(ctx, T.dm.data(), T.B.N, K));
        T.B.nodes.assign(K.left.size(), Node());
        for (size_t i = 0; i < K.left.size(); i++) {
                T.B.nodes[i].left = K.left[i];
                T.B.nodes[i].right = K.right[i];
        }
        T.B.clusters.clear();
        for (size_t l = 0; l < K.leaf_node.size(); l++) {
                Cluster c;
                c.samples.assign(K.order.begin() + K.leaf_begin[l], K.order.begin() + K.leaf_end[l]);
                c.placeholder = K.leaf_node[l];
                T.B.clusters.push_back(std::move(c));
        }
        T.root = K.root;
        if (kb_trace_on()) {
                fprintf(stderr, "[kb200 trace] guide tree: bisecting k-means on the device %.1f ms\n", 1e3 * (kb_now() - tt1));
        }
        return KB200_OK;
}

void kb_tree_bisect(KbTreeJob& T, int n_threads)
{
        const double tt1 = kb_now();
        const int N = T.B.N;
        std::vector<int> samples((size_t)N);
        for (int i = 0; i < N; i++) samples[(size_t)i] = i;
        int root = -1;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
#pragma omp single
#endif
        root = bisect(T.B, samples);
        T.root = root;
        if (kb_trace_on()) {
                fprintf(stderr, "[kb200 trace] guide tree: bisecting k-means %.1f ms (%d threads)\n", 1e3 * (kb_now() - tt1), n_threads);
        }
}

int kb_tree_finish(kb200_ctx* ctx, KbSeqs& S, KbTreeJob& T, int n_threads, std::vector<int>& abc)
{
        const double tt2 = kb_now();
        TreeBuilder& B = T.B;
        const int N = B.N;
        // leaf clusters: one batched launch for every i<j pair of every cluster.
        // d_estimation(pair=1) leaves dm[i][j] (i<j) = calc_distance(seq_j, seq_i)
        // (sequence_distance.c:53-81: the later (j,i) iteration overwrites both entries).
        {
                size_t npairs = 0;
                for (const Cluster& c : B.clusters) {
                        const size_t n = c.samples.size();
                        npairs += n * (n - 1) / 2;
                }
                std::vector<int> pa(npairs), pb(npairs);
                size_t e = 0;
                for (const Cluster& c : B.clusters) {
                        const int n = (int)c.samples.size();
                        for (int i = 0; i < n; i++) {
                                for (int j = i + 1; j < n; j++) {
                                        pa[e] = c.samples[(size_t)j];
                                        pb[e] = c.samples[(size_t)i];
                                        e++;
                                }
                        }
                }
                const bool host_upgma = getenv("KB200_HOST_KMEANS") != nullptr;
                std::vector<float> pd(host_upgma ? npairs : 0);
                if (npairs) {
                        KB_RUN(kb_distances_dev(ctx, S, pa.data(), (int)npairs, pb.data(), 0, 1, host_upgma ? pd.data() : nullptr));
                }
                // clusters were recorded in completion order; the result is order independent.
                // Every cluster gets its node range and its slice of the pair list up front.
                const size_t ncl = B.clusters.size();
                std::vector<int> cluster_root(ncl);
                std::vector<size_t> pair0(ncl);
                std::vector<int> node0(ncl);
                {
                        size_t pe = 0;
                        int nb = (int)B.nodes.size();
                        for (size_t ci = 0; ci < ncl; ci++) {
                                const size_t n = B.clusters[ci].samples.size();
                                pair0[ci] = pe;
                                node0[ci] = nb;
                                pe += n * (n - 1) / 2;
                                nb += (n > 0) ? (int)(2 * n - 1) : 0;
                        }
                        B.nodes.resize((size_t)nb);
                }
                if (!host_upgma) {
                        // UPGMA on the device: one warp per cluster, the pair distances never leave the GPU
                        // (they are still in the distance launch's output buffer); only the merge lists
                        // (2 ints per internal node) come back, the node bookkeeping of upgma() is replayed here
                        std::vector<int> csize(ncl);
                        std::vector<long long> p0(ncl), m0(ncl);
                        long long nm = 0;
                        for (size_t ci = 0; ci < ncl; ci++) {
                                csize[ci] = (int)B.clusters[ci].samples.size();
                                p0[ci] = (long long)pair0[ci];
                                m0[ci] = nm;
                                nm += std::max(0, csize[ci] - 1);
                        }
                        std::vector<int> merges((size_t)2 * (size_t)std::max<long long>(nm, 1));
                        KB_RUN(kb_upgma_dev(ctx, ctx->d_stage5.as<float>(), csize, p0, m0, nm, merges.data()));
                        for (size_t ci = 0; ci < ncl; ci++) {
                                const Cluster& c = B.clusters[ci];
                                const int n = csize[ci];
                                int next_node = node0[ci];
                                std::vector<int> tree((size_t)n);
                                for (int i = 0; i < n; i++) {
                                        Node nd;
                                        nd.id = c.samples[(size_t)i];
                                        tree[(size_t)i] = next_node;
                                        B.nodes[(size_t)next_node++] = nd;
                                }
                                int last_a = 0;
                                for (int sidx = 0; sidx < n - 1; sidx++) {
                                        const int a = merges[(size_t)2 * (size_t)(m0[ci] + sidx)];
                                        const int b = merges[(size_t)2 * (size_t)(m0[ci] + sidx) + 1];
                                        Node nd;
                                        nd.left = tree[(size_t)a];
                                        nd.right = tree[(size_t)b];
                                        tree[(size_t)a] = next_node;
                                        B.nodes[(size_t)next_node++] = nd;
                                        tree[(size_t)b] = -1;
                                        last_a = a;
                                }
                                cluster_root[ci] = tree[(size_t)last_a];
                        }
                } else {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads > 0 ? n_threads : 1)
#endif
                for (long long cl = 0; cl < (long long)ncl; cl++) {
                        const size_t ci = (size_t)cl;
                        const Cluster& c = B.clusters[ci];
                        const int n = (int)c.samples.size();
                        size_t e2 = pair0[ci];
                        std::vector<float> cdm((size_t)n * n, 0.0f);
                        for (int i = 0; i < n; i++) {
                                for (int j = i + 1; j < n; j++) {
                                        cdm[(size_t)i * n + j] = pd[e2];
                                        cdm[(size_t)j * n + i] = pd[e2];
                                        e2++;
                                }
                        }
                        cluster_root[ci] = upgma(B, cdm.data(), c.samples, node0[ci]);
                }
                }
                // splice the UPGMA sub-trees in place of the placeholders
                for (size_t ci = 0; ci < B.clusters.size(); ci++) {
                        B.nodes[(size_t)B.clusters[ci].placeholder] = B.nodes[(size_t)cluster_root[ci]];
                }
        }
        abc.clear();
        abc.reserve((size_t)3 * (N - 1));
        emit_tasks(B, T.root, N, abc);
        if (kb_trace_on()) {
                fprintf(stderr, "[kb200 trace] guide tree: leaf distances + upgma %.1f ms\n", 1e3 * (kb_now() - tt2));
        }
        if ((int)abc.size() != 3 * (N - 1)) {
                fprintf(stderr, "[kalign_b200] guide tree: %zu tasks for %d sequences\n", abc.size() / 3, N);
                return KB200_FAIL;
        }
        return KB200_OK;
}

// multi-GPU: the guide tree is built ONCE, by rank 0 with every host thread of the node (k-means and
// UPGMA are host code; N ranks each running them with 1/N of the CPU quota only made the public call
// slower), and the task list is broadcast over NCCL.  Collective: every rank must call it.
int kb_tree_broadcast(kb200_ctx* ctx, int N, std::vector<int>& abc)
{
        if (ctx->world <= 1) {
                return KB200_OK;
        }
        const size_t bytes = sizeof(int) * 3 * (size_t)(N - 1);
        KB_RUN(ctx->d_stage4.ensure(bytes + 16));
        if (ctx->rank == 0) {
                if (abc.size() != (size_t)3 * (size_t)(N - 1)) return KB200_FAIL;
                KB_CUDA(cudaMemcpyAsync(ctx->d_stage4.p, abc.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
        }
        std::vector<size_t> seg((size_t)ctx->world + 1, bytes);       // rank 0 owns the whole buffer
        seg[0] = 0;
        KB_RUN(kb_allgatherv(ctx, ctx->d_stage4.p, seg.data()));
        abc.resize((size_t)3 * (size_t)(N - 1));
        KB_CUDA(cudaMemcpyAsync(abc.data(), ctx->d_stage4.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        KB_CUDA(cudaStreamSynchronize(ctx->stream));
        return KB200_OK;
}

// threads for the host-only tree stages: rank 0 of a multi-GPU run works alone and takes the whole quota
static int tree_threads(const kb200_ctx* ctx, int n_threads)
{
        return (ctx->world > 1) ? kb_default_threads() : n_threads;
}

// build_tree_kmeans (bisectingKmeans.c:177-271) in one go
int kb_build_tree(kb200_ctx* ctx, KbSeqs& S, int n_threads, std::vector<int>& abc, std::vector<float>& seq_distances,
                  uint64_t tree_seed = 0, float tree_noise = 0.0f)
{
        KbTreeJob T;
        T.tree_seed = tree_seed;
        T.tree_noise = tree_noise;
        KB_RUN(kb_tree_prepare(ctx, S, T, seq_distances));      // N x 32 distances: rows sharded across the ranks
        int rc = KB200_OK;
        if (ctx->rank == 0) {
                rc = kb_tree_bisect_any(ctx, T, tree_threads(ctx, n_threads));
                if (rc == KB200_OK) rc = kb_tree_finish(ctx, S, T, tree_threads(ctx, n_threads), abc);
        }
        if (ctx->world > 1) {
                // a failure on rank 0 must not leave the other ranks waiting in the collective
                if (rc != KB200_OK) abc.assign((size_t)3 * (size_t)(S.n - 1), -1);
                KB_RUN(kb_tree_broadcast(ctx, S.n, abc));
                if (!abc.empty() && abc[0] < 0) rc = KB200_FAIL;
        }
        return rc;
}


// ---------------------------------------------------------------------------------------------
// encode / finalise on the device: the raw characters are uploaded ONCE (sorted order,
// concatenated); convert_msa_to_internal (lib/src/msa_op.c:344-375) is a table look-up kernel -- run twice
// for proteins (13-letter tree alphabet, then the 23-letter alignment alphabet) without touching the host --
// and finalise_alignment (lib/src/msa_op.c:546-598) a scatter: residue p of a sequence lands in column
// colof[p] = p + sum_{q<=p} gaps[q] of its row, everything else is '-'.
namespace {

struct KbLut { int8_t t[128]; };

__global__ void kb_encode_kernel(const uint8_t* __restrict__ raw, const long long total, const KbLut lut,
                                 uint8_t* __restrict__ codes, unsigned* __restrict__ unknown)
{
        const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        const long long nth = (long long)gridDim.x * blockDim.x;
        unsigned bad = 0;
        for (long long i = gid; i < total; i += nth) {
                const unsigned ch = raw[i];
                const int t = (ch < 128u) ? (int)lut.t[ch] : -1;
                if (t < 0) bad = ch ? ch : 1u;
                codes[i] = (uint8_t)((t < 0) ? 0 : t);     // unknown characters are coded as 0 (msa_op.c:358-362)
        }
        if (bad) atomicMax(unknown, bad);
}

__global__ void kb_finalise_kernel(const uint8_t* __restrict__ raw, const int64_t* __restrict__ offs, const int* __restrict__ lens,
                                   const int* __restrict__ rank, const int nseq, const int* __restrict__ colof,
                                   const long long stride, char* __restrict__ rows)
{
        // one warp per sequence; the rows were filled with '-' (and NUL terminators) before
        const int lane = threadIdx.x & 31;
        const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        if (i >= nseq) return;
        const long long off = offs[i];
        char* __restrict__ row = rows + (long long)rank[i] * stride;
        for (int p = lane; p < lens[i]; p += 32) {
                row[colof[off + p]] = (char)raw[off + p];
        }
}

__global__ void kb_fill_rows_kernel(char* __restrict__ rows, const long long stride, const int nseq)
{
        const long long n = stride * (long long)nseq;
        const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        const long long nth = (long long)gridDim.x * blockDim.x;
        for (long long x = gid; x < n; x += nth) {
                rows[x] = ((x % stride) == stride - 1) ? (char)0 : '-';
        }
}

} // namespace

// ---------------------------------------------------------------------------------------------
// staged pipeline object: create (encode, sort, upload, distances, guide tree) -> align (the DP
// stages on device-resident sequences: anchor batch + progressive alignment) -> result.
struct kb200_msa {
        kb200_ctx* ctx = nullptr;
        int n_threads = 1;
        int biotype = 2;
        int N = 0;
        std::vector<Seq> store;
        std::vector<Seq*> order;
        std::vector<int64_t> offs;
        std::vector<int> lens;
        uint8_t* raw = nullptr;          // the input characters, sorted order, concatenated (page-locked)
        KbDevBuf d_raw, d_rank, d_rows, d_colof;
        int aln_len = 0;                 // msa->alnlen of the last kb200_msa_align (= plen of the root task)
        bool gaps_on_host = false;
        int64_t total = 0;
        KbSeqs S;
        // deferred guide tree (kb200_kalign): k-means runs on a host thread beside the anchor batch;
        // proteins keep their 13-letter tree codes on the device for the leaf-cluster distances
        KbTreeJob* tree_job = nullptr;
        KbSeqs S_tree;
        std::vector<int> abc;
        std::vector<float> seq_distances;
        std::vector<int> posmaps;
        int* gaps = nullptr;             // total + N ints, page-locked (kb_host_take)
        std::vector<int> anchor_ids;
        kb200_params prm;
        int K = 0;
        float weight = 2.0f;
        bool aligned = false;
        double t_create = 0, t_tree = 0;
};

// convert_msa_to_internal on the device: S gets the codes of alphabet `alpha`
static int encode_seqs_dev(kb200_msa* M, KbSeqs& S, int alpha)
{
        kb200_ctx* ctx = M->ctx;
        const Alphabet a = make_alphabet(alpha);
        KbLut lut;
        for (int i = 0; i < 128; i++) lut.t[i] = a.to_internal[i];
        if (!S.d_seqs.p) {
                KB_RUN(S.alloc(ctx, M->offs.data(), M->lens.data(), M->N));
        }
        KB_RUN(ctx->d_counters.ensure(sizeof(KbRound) * (KB_MAX_ROUNDS + 2) + 64));
        unsigned* d_flag = ctx->d_counters.as<unsigned>();
        KB_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(unsigned), ctx->stream));
        const int grid = (int)std::min<long long>((M->total + 255) / 256 + 1, (long long)ctx->sm_count * 16);
        kb_encode_kernel<<<grid, 256, 0, ctx->stream>>>(M->d_raw.as<uint8_t>(), (long long)M->total, lut, S.d_seqs.as<uint8_t>(), d_flag);
        KB_CUDA(cudaGetLastError());
        ctx->stats.n_launches++;
        unsigned bad = 0;
        KB_CUDA(cudaMemcpyAsync(&bad, d_flag, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
        KB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (bad) {
                fprintf(stderr, "[kalign_b200] warning: character '%c' does not match the alphabet (coded as 0)\n", (int)bad < 128 ? (char)bad : '?');
        }
        return KB200_OK;
}

extern "C" {

// host-only helper for a binding that replaces anchor_consistency_build: the reference's anchor
// choice is a static function (select_anchors, lib/src/anchor_consistency.c:124-198)
int kb200_select_anchors(const float* seq_distances, int nseq, int K, int* anchor_ids)
{
        if (!seq_distances || !anchor_ids || nseq < 1 || K < 1 || K > nseq) {
                fprintf(stderr, "[kalign_b200] kb200_select_anchors: bad arguments\n");
                return KB200_FAIL;
        }
        std::vector<int> ids;
        select_anchors(seq_distances, nseq, K, ids);
        for (int k = 0; k < K; k++) {
                anchor_ids[k] = ids[(size_t)k];
        }
        return KB200_OK;
}

void kb200_msa_free(kb200_msa* M)
{
        if (!M) return;
        if (M->ctx) cudaSetDevice(M->ctx->device);
        M->S.release();
        M->S_tree.release();
        if (M->gaps && M->ctx) kb_host_give(M->ctx, M->gaps);
        if (M->raw && M->ctx) kb_host_give(M->ctx, M->raw);
        if (M->ctx) {
                // device buffers go back to the context's pool like the sequence buffers
                KbDevBuf* bufs[4] = {&M->d_raw, &M->d_rank, &M->d_rows, &M->d_colof};
                for (KbDevBuf* b : bufs) {
                        kb_give_pooled(M->ctx, *b);
                }
        }
        delete M->tree_job;
        delete M;
}

// everything of kalign_run_seeded (aln_wrap.c:133-205) that precedes the DP stages
// detect_alphabet (msa_op.c:142-215): 0 protein, 1 nucleotide, 2 undecided
static int kb_detect_biotype(const int* letter_freq)
{
        int biotype = 2;
        double DNA[128], protein[128];
        const char* DNA_letters = "acgtunACGTUN";
        const char* protein_letters = "acdefghiklmnpqrstvwyACDEFGHIKLMNPQRSTVWY";
        for (int i = 0; i < 128; i++) {
                DNA[i] = log(0.0001 * 1.0 / 116.0);
                protein[i] = log(0.0001 * 1.0 / 88.0);
        }
        for (int i = 0; i < 12; i++) DNA[(int)DNA_letters[i]] = log(0.9999 * 1.0 / 12.0);
        for (int i = 0; i < 40; i++) protein[(int)protein_letters[i]] = log(0.9999 * 1.0 / 40.0);
        double dna_prob = 0.0, prot_prob = 0.0;
        for (int i = 0; i < 128; i++) {
                if (letter_freq[i]) {
                        dna_prob += DNA[i] * (double)letter_freq[i];
                        prot_prob += protein[i] * (double)letter_freq[i];
                }
        }
        if (dna_prob > prot_prob) biotype = 1;
        else if (prot_prob > dna_prob) biotype = 0;
        return biotype;
}

// what kalign_run_seeded takes beyond kalign() (aln_wrap.c:133-139, 169-199), plus the two things only file input has
struct KbRunOpts {
        // names / file_letter_freq: only for file input (kb200_kalign_file): the record names break ties of the
        // (length, name) sort as in the reference's FASTA path, and detect_alphabet sees the letter frequencies
        // read_fasta counted on the sequence lines (msa_io.c:457; gap characters and digits included)
        const char* const* names = nullptr;
        const int* file_letter_freq = nullptr;
        uint64_t tree_seed = 0;          // != 0 with tree_noise > 0: build_tree_kmeans_noisy
        float tree_noise = 0.0f;
        float dist_scale = 0.0f;         // ap->dist_scale (always assigned)
        float vsm_amax = -1.0f;          // >= 0 overrides the default
        float use_seq_weights = -1.0f;   // >= 0 overrides the default
};
static const KbRunOpts kb_default_opts;

static int msa_create_impl(kb200_ctx* ctx, char** seq, int* len, int numseq, int n_threads, int type,
                           float gpo, float gpe, float tgpe, int consistency_anchors, float consistency_weight,
                           const bool defer_tree, kb200_msa** out, const KbRunOpts& opts = kb_default_opts)
{
        const char* const* names = opts.names;
        const int* file_letter_freq = opts.file_letter_freq;
        if (!ctx || !seq || !len || !out) {
                return KB200_FAIL;
        }
        *out = nullptr;
        if (n_threads < 1) n_threads = 1;
        n_threads = std::min(n_threads, kb_default_threads());   // never more than the CPU quota
        KB_CUDA(cudaSetDevice(ctx->device));
        const double tc0 = kb_now();
        // ---- kalign_arr_to_msa: letter frequencies, alphabet detection (msa_op.c:440-520,142-215)
        int letter_freq[128];
        memset(letter_freq, 0, sizeof(letter_freq));
        if (file_letter_freq) {
                memcpy(letter_freq, file_letter_freq, sizeof(letter_freq));
        } else
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads)
#endif
        {
                int local[128];
                memset(local, 0, sizeof(local));
#ifdef _OPENMP
#pragma omp for schedule(static) nowait
#endif
                for (int i = 0; i < numseq; i++) {
                        for (int j = 0; j < len[i]; j++) {
                                const int ch = (int)(unsigned char)seq[i][j];
                                if (ch < 128) local[ch]++;
                        }
                }
#ifdef _OPENMP
#pragma omp critical
#endif
                for (int c = 0; c < 128; c++) letter_freq[c] += local[c];
        }
        const int biotype = kb_detect_biotype(letter_freq);
        if (biotype == 2) {
                fprintf(stderr, "[kalign_b200] Unable to determine what alphabet to use.\n");
                return KB200_FAIL;
        }
        // ---- kalign_essential_input_check: ranks, drop empty sequences (msa_check.c:66-139)
        if (numseq <= 1) {
                fprintf(stderr, "[kalign_b200] only %d sequences found.\n", numseq);
                return KB200_FAIL;
        }
        kb200_msa* M = new kb200_msa();
        M->ctx = ctx;
        M->n_threads = n_threads;
        M->biotype = biotype;
        M->weight = consistency_weight;
        M->store.resize((size_t)numseq);
        M->order.reserve((size_t)numseq);
        for (int i = 0; i < numseq; i++) {
                Seq& s = M->store[(size_t)i];
                s.seq = seq[i];
                s.len = len[i];
                s.rank = i;
                s.name = names ? std::string(names[i]) : "s" + std::to_string(i);
                if (len[i] > 0) M->order.push_back(&s);
        }
        const int N = (int)M->order.size();
        M->N = N;
        if (N <= 1) {
                fprintf(stderr, "[kalign_b200] only %d sequences found.\n", N);
                delete M;
                return KB200_FAIL;
        }
        // ---- msa_sort_len_name (same libc qsort, same comparator semantics)
        qsort(M->order.data(), (size_t)N, sizeof(Seq*), cmp_len_name);
        M->offs.resize((size_t)N);
        M->lens.resize((size_t)N);
        int64_t total = 0;
        for (int i = 0; i < N; i++) {
                M->offs[(size_t)i] = total;
                M->lens[(size_t)i] = M->order[(size_t)i]->len;
                total += M->order[(size_t)i]->len;
        }
        M->total = total;
        M->raw = (uint8_t*)kb_host_take(ctx, (size_t)total + 16);
        if (!M->raw) {
                kb200_msa_free(M);
                return KB200_FAIL;
        }
        // the input characters in sorted order, one contiguous page-locked block: uploaded once, encoded
        // (and later expanded into the aligned rows) on the device
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads)
#endif
        for (int i = 0; i < N; i++) {
                memcpy(M->raw + M->offs[(size_t)i], M->order[(size_t)i]->seq, (size_t)M->lens[(size_t)i]);
        }
        const double tc1 = kb_now();
        int rc = KB200_OK;
        double tc2 = tc1;
        {
                // pooled device buffer for the raw characters and the rank table
                kb_take_pooled(ctx, M->d_raw, (size_t)total + 16);
                kb_take_pooled(ctx, M->d_rank, sizeof(int) * (size_t)N + 16);
                rc = M->d_raw.ensure((size_t)total + 16);
                if (rc == KB200_OK) rc = M->d_rank.ensure(sizeof(int) * (size_t)N + 16);
                // output row of sorted sequence i: its position among the kept sequences in input order
                // (msa_sort_rank + kalign_msa_to_arr; empty input sequences were dropped)
                std::vector<int> by_rank((size_t)N), rank((size_t)N);
                for (int i = 0; i < N; i++) by_rank[(size_t)i] = i;
                std::sort(by_rank.begin(), by_rank.end(), [&](int x, int y) { return M->order[(size_t)x]->rank < M->order[(size_t)y]->rank; });
                for (int r = 0; r < N; r++) rank[(size_t)by_rank[(size_t)r]] = r;
                if (rc == KB200_OK) {
                        if (cudaMemcpyAsync(M->d_raw.p, M->raw, (size_t)total, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
                            cudaMemcpyAsync(M->d_rank.p, rank.data(), sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
                            cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
                                rc = KB200_FAIL;
                        }
                        ctx->stats.h2d_bytes += (double)total + 4.0 * N;
                }
        }
        if (rc == KB200_OK && !defer_tree) {
                rc = encode_seqs_dev(M, M->S, biotype == 1 ? ALPHA_DNA : ALPHA_RED);
                tc2 = kb_now();
                if (rc == KB200_OK) rc = kb_build_tree(ctx, M->S, n_threads, M->abc, M->seq_distances, opts.tree_seed, opts.tree_noise);
                if (rc == KB200_OK && biotype == 0) {
                        rc = encode_seqs_dev(M, M->S, ALPHA_AMB);
                }
        } else if (rc == KB200_OK) {
                // only the first stage of the tree here (distance matrix -> seq_distances -> anchors);
                // kb200_kalign finishes the tree around the anchor batch
                M->tree_job = new KbTreeJob();
                M->tree_job->tree_seed = opts.tree_seed;
                M->tree_job->tree_noise = opts.tree_noise;
                if (biotype == 0) {
                        rc = encode_seqs_dev(M, M->S_tree, ALPHA_RED);        // 13-letter tree alphabet
                        tc2 = kb_now();
                        if (rc == KB200_OK) rc = kb_tree_prepare(ctx, M->S_tree, *M->tree_job, M->seq_distances);
                        if (rc == KB200_OK) rc = encode_seqs_dev(M, M->S, ALPHA_AMB);
                } else {
                        rc = encode_seqs_dev(M, M->S, ALPHA_DNA);
                        tc2 = kb_now();
                        if (rc == KB200_OK) rc = kb_tree_prepare(ctx, M->S, *M->tree_job, M->seq_distances);
                }
        }
        if (kb_trace_on()) {
                fprintf(stderr, "[kb200 trace] msa_create: detect/sort/encode %.1f ms, upload %.1f ms, guide tree%s %.1f ms\n",
                        1e3 * (tc1 - tc0), 1e3 * (tc2 - tc1), defer_tree ? " (first stage)" : "", 1e3 * (kb_now() - tc2));
        }
        if (rc == KB200_OK) {
                // resolve_pfasum_auto, aln_wrap.c:31-68
                if (type == KB200_TYPE_PROTEIN_PFASUM_AUTO) {
                        if (biotype != 0) {
                                type = KB200_TYPE_PROTEIN_PFASUM43;
                        } else {
                                int mn = M->lens[0], mxl = M->lens[0];
                                for (int i = 1; i < N; i++) { mn = std::min(mn, M->lens[(size_t)i]); mxl = std::max(mxl, M->lens[(size_t)i]); }
                                const float ratio = (mn > 0) ? (float)mxl / (float)mn : 1.0f;
                                type = (ratio < 1.5f) ? KB200_TYPE_PROTEIN_PFASUM43 : KB200_TYPE_PROTEIN_PFASUM60;
                        }
                }
                rc = kb200_params_init(&M->prm, biotype, type, gpo, gpe, tgpe);
                // aln_wrap.c:193-199
                M->prm.dist_scale = opts.dist_scale;
                if (opts.vsm_amax >= 0.0f) M->prm.vsm_amax = opts.vsm_amax;
                if (opts.use_seq_weights >= 0.0f) M->prm.use_seq_weights = opts.use_seq_weights;
        }
        if (rc == KB200_OK && consistency_anchors > 0 && N >= 3) {
                M->K = std::min(consistency_anchors, N);
                select_anchors(M->seq_distances.data(), N, M->K, M->anchor_ids);
                // host copy of the position maps only in the host-bonus A/B mode; otherwise the vector is a tag
                M->posmaps.assign(kb_bonus_on_host(M->K) ? (size_t)total * (size_t)M->K : (size_t)8, -1);
        }
        if (rc != KB200_OK) {
                kb200_msa_free(M);
                return KB200_FAIL;
        }
        *out = M;
        return KB200_OK;
}

int kb200_msa_create(kb200_ctx* ctx, char** seq, int* len, int numseq, int n_threads, int type,
                     float gpo, float gpe, float tgpe, int consistency_anchors, float consistency_weight,
                     kb200_msa** out)
{
        return msa_create_impl(ctx, seq, len, numseq, n_threads, type, gpo, gpe, tgpe, consistency_anchors, consistency_weight,
                               false, out);
}

// the DP stages (anchor_consistency_build + create_msa_tree) on device-resident sequences;
// may be called repeatedly (bench), every call recomputes everything.
static int msa_align_anchor(kb200_msa* M)
{
        if (M->K > 0) {
                KB_RUN(kb_anchor_posmaps_sharded(M->ctx, &M->prm, M->S, M->anchor_ids.data(), M->K, M->posmaps.data(), kb_bonus_on_host(M->K)));
        }
        return KB200_OK;
}

static int msa_align_tree(kb200_msa* M)
{
        // the gaps stay on the device (kb200_msa_result expands the rows there); only the host restatement
        // of the weave (KB200_HOST_BONUS) hands them back
        kb200_ctx* ctx = M->ctx;
        const bool host_state = kb_bonus_on_host(M->K) != 0;
        if (host_state && !M->gaps) {
                M->gaps = (int*)kb_host_take(ctx, sizeof(int) * ((size_t)M->total + (size_t)M->N));
                if (!M->gaps) return KB200_FAIL;
                memset(M->gaps, 0, sizeof(int) * ((size_t)M->total + (size_t)M->N));
        }
        std::vector<int> plens((size_t)std::max(1, M->N - 1));
        int* plen = plens.data();
        KB_RUN(kb_align_tree_dev(M->ctx, &M->prm, M->S, M->abc.data(), M->N - 1, M->seq_distances.data(),
                                 M->K > 0 ? M->posmaps.data() : nullptr, M->K, M->weight, M->n_threads,
                                 host_state ? M->gaps : nullptr, 1, nullptr, plen));
        M->aln_len = plens[(size_t)(M->N - 2)];          // msa->alnlen = plen of the root task
        M->gaps_on_host = host_state;
        if (!host_state) {
                // the context's column maps belong to whichever alignment ran last: keep this one's
                kb_take_pooled(ctx, M->d_colof, sizeof(int) * ((size_t)M->total + 8));
                KB_RUN(M->d_colof.ensure(sizeof(int) * ((size_t)M->total + 8)));
                KB_CUDA(cudaMemcpyAsync(M->d_colof.p, ctx->t_colof.p, sizeof(int) * (size_t)M->total, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        M->aligned = true;
        return KB200_OK;
}

int kb200_msa_align(kb200_msa* M)
{
        if (!M || M->tree_job) return KB200_FAIL;       // a deferred tree is finished by kb200_kalign only
        kb200_ctx* ctx = M->ctx;
        KB_CUDA(cudaSetDevice(ctx->device));
        // timed span of the whole step on the engine's stream (ev2 / ev3 live in the context)
        KB_CUDA(cudaEventRecord(ctx->ev2, ctx->stream));
        KB_RUN(msa_align_anchor(M));
        KB_RUN(msa_align_tree(M));
        KB_CUDA(cudaEventRecord(ctx->ev3, ctx->stream));
        KB_CUDA(cudaEventSynchronize(ctx->ev3));
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3);
        ctx->stats.align_seconds += 1e-3 * (double)ms;
        return KB200_OK;
}

// finalise_alignment (msa_op.c:546-598) + msa_sort_rank + kalign_msa_to_arr (msa_op.c:377)
int kb200_msa_result(kb200_msa* M, char*** aligned, int* out_aln_len)
{
        if (!M || !M->aligned || !aligned || !out_aln_len) return KB200_FAIL;
        const double tr0 = kb_now();
        const int N = M->N;
        kb200_ctx* ctx = M->ctx;
        char** out = (char**)calloc((size_t)N, sizeof(char*));
        if (!out) return KB200_FAIL;
        bool oom = false;
        int aln_len = 0;
        if (!M->gaps_on_host) {
                // finalise_alignment on the device: rows of '-' with the residues scattered to their columns
                // (colof, maintained by the weave kernels), one block copied back, cut into the caller's rows
                KB_CUDA(cudaSetDevice(ctx->device));
                aln_len = M->aln_len;
                const long long stride = (long long)aln_len + 1;
                const size_t bytes = (size_t)stride * (size_t)N;
                kb_take_pooled(ctx, M->d_rows, bytes + 16);
                KB_RUN(M->d_rows.ensure(bytes + 16));
                char* hrows = (char*)kb_host_take(ctx, bytes + 16);
                if (!hrows) { free(out); return KB200_FAIL; }
                const int fgrid = (int)std::min<long long>((long long)(bytes / 256) + 1, (long long)ctx->sm_count * 32);
                kb_fill_rows_kernel<<<fgrid, 256, 0, ctx->stream>>>(M->d_rows.as<char>(), stride, N);
                kb_finalise_kernel<<<(N * 32 + 127) / 128, 128, 0, ctx->stream>>>(M->d_raw.as<uint8_t>(), M->S.d_offs.as<int64_t>(), M->S.d_lens.as<int>(),
                                                                              M->d_rank.as<int>(), N, M->d_colof.as<int>(), stride, M->d_rows.as<char>());
                KB_CUDA(cudaGetLastError());
                ctx->stats.n_launches += 2;
                KB_CUDA(cudaMemcpyAsync(hrows, M->d_rows.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
                KB_CUDA(cudaStreamSynchronize(ctx->stream));
                ctx->stats.d2h_bytes += (double)bytes;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(kb_default_threads())
#endif
                for (int r = 0; r < N; r++) {
                        char* row = (char*)malloc((size_t)stride);
                        if (!row) {
                                oom = true;
                                continue;
                        }
                        memcpy(row, hrows + (size_t)r * (size_t)stride, (size_t)stride);
                        out[r] = row;
                }
                kb_host_give(ctx, hrows);
        } else {
        aln_len = M->lens[0];
        {
                const int* g = M->gaps + M->offs[0] + 0;
                for (int j = 0; j <= M->lens[0]; j++) aln_len += g[j];
        }
        std::vector<int> by_rank((size_t)N);
        for (int i = 0; i < N; i++) by_rank[(size_t)i] = i;
        std::sort(by_rank.begin(), by_rank.end(), [&](int x, int y) { return M->order[(size_t)x]->rank < M->order[(size_t)y]->rank; });
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(kb_default_threads())
#endif
        for (int r = 0; r < N; r++) {
                const int i = by_rank[(size_t)r];
                const int* g = M->gaps + M->offs[(size_t)i] + i;
                char* row = (char*)malloc((size_t)aln_len + 1);
                if (!row) {
                        oom = true;
                        continue;
                }
                int f = 0;
                const char* s = M->order[(size_t)i]->seq;
                const int li = M->lens[(size_t)i];
                for (int j = 0; j < li; j++) {
                        for (int c = 0; c < g[j] && f < aln_len; c++) row[f++] = '-';
                        if (f < aln_len) row[f++] = s[j];
                }
                for (int c = 0; c < g[li] && f < aln_len; c++) row[f++] = '-';
                while (f < aln_len) row[f++] = '-';
                row[aln_len] = 0;
                out[r] = row;
        }
        }
        if (oom) {
                for (int r = 0; r < N; r++) free(out[r]);
                free(out);
                return KB200_FAIL;
        }
        *aligned = out;
        *out_aln_len = aln_len;
        if (kb_trace_on()) {
                fprintf(stderr, "[kb200 trace] msa_result: %d rows x %d columns in %.1f ms\n", N, aln_len, 1e3 * (kb_now() - tr0));
        }
        return KB200_OK;
}

int kb200_msa_info(kb200_msa* M, int* numseq, int* biotype, int* n_anchors)
{
        if (!M) return KB200_FAIL;
        if (numseq) *numseq = M->N;
        if (biotype) *biotype = M->biotype;
        if (n_anchors) *n_anchors = M->K;
        return KB200_OK;
}

int kb200_msa_tree(kb200_msa* M, int* tasks_abc, float* seq_distances)
{
        if (!M || M->tree_job || (int)M->abc.size() != 3 * (M->N - 1)) return KB200_FAIL;
        if (tasks_abc) memcpy(tasks_abc, M->abc.data(), sizeof(int) * M->abc.size());
        if (seq_distances) memcpy(seq_distances, M->seq_distances.data(), sizeof(float) * (size_t)M->N);
        return KB200_OK;
}

static int kalign_named(kb200_ctx* ctx, char** seq, int* len, int numseq, int n_threads, int type,
                        float gpo, float gpe, float tgpe, int consistency_anchors, float consistency_weight,
                        char*** aligned, int* out_aln_len, const KbRunOpts& opts)
{
        if (!aligned || !out_aln_len) {
                return KB200_FAIL;
        }
        // One-shot call: the host-only stage of the guide tree (bisecting k-means) runs on a host
        // thread while the GPU works through the anchor batch, which needs only seq_distances.
        kb200_msa* M = nullptr;
        KB_RUN(msa_create_impl(ctx, seq, len, numseq, n_threads, type, gpo, gpe, tgpe,
                               consistency_anchors, consistency_weight, true, &M, opts));
        const double ta0 = kb_now();
        int rc = KB200_OK;
        {
                KbTreeJob* T = M->tree_job;
                const int nt = tree_threads(ctx, M->n_threads);
                const bool build = (ctx->rank == 0);          // multi-GPU: rank 0 builds the tree, the others receive it
                int krc = KB200_OK;
                // the device k-means runs BEFORE the anchor batch (its CTAs would starve behind the
                // persistent sweep grid); only the host restatement (KB200_HOST_KMEANS=1) runs beside it
                const bool host_km = getenv("KB200_HOST_KMEANS") != nullptr;
                std::thread kmeans;
                if (build && !host_km) {
                        krc = kb_tree_bisect_any(ctx, *T, nt);
                } else if (build) {
                        kmeans = std::thread([ctx, T, nt, &krc]() { krc = kb_tree_bisect_any(ctx, *T, nt); });
                }
                rc = msa_align_anchor(M);
                if (kmeans.joinable()) kmeans.join();
                int trc = krc;
                if (build && trc == KB200_OK) {
                        trc = kb_tree_finish(ctx, M->biotype == 0 ? M->S_tree : M->S, *T, nt, M->abc);
                }
                if (ctx->world > 1) {
                        if (trc != KB200_OK) M->abc.assign((size_t)3 * (size_t)(M->N - 1), -1);
                        if (kb_tree_broadcast(ctx, M->N, M->abc) != KB200_OK) rc = KB200_FAIL;
                        if (!M->abc.empty() && M->abc[0] < 0) trc = KB200_FAIL;
                }
                if (rc == KB200_OK) rc = trc;
                delete M->tree_job;
                M->tree_job = nullptr;
                M->S_tree.release();
        }
        const double ta1 = kb_now();
        if (rc == KB200_OK) rc = msa_align_tree(M);
        const double ta2 = kb_now();
        if (rc == KB200_OK) rc = kb200_msa_result(M, aligned, out_aln_len);
        kb200_msa_free(M);
        if (kb_trace_on()) {
                fprintf(stderr, "[kb200 trace] kalign: anchor batch || k-means, tree finish %.1f ms, tree levels %.1f ms, result + free %.1f ms\n",
                        1e3 * (ta1 - ta0), 1e3 * (ta2 - ta1), 1e3 * (kb_now() - ta2));
        }
        return rc;
}

int kb200_kalign(kb200_ctx* ctx, char** seq, int* len, int numseq, int n_threads, int type,
                 float gpo, float gpe, float tgpe, int consistency_anchors, float consistency_weight,
                 char*** aligned, int* out_aln_len)
{
        return kalign_named(ctx, seq, len, numseq, n_threads, type, gpo, gpe, tgpe, consistency_anchors, consistency_weight,
                            aligned, out_aln_len, kb_default_opts);
}

// kalign_run_seeded (lib/include/kalign/kalign.h:51, aln_wrap.c:133) on plain arrays; refine = none
int kb200_kalign_seeded(kb200_ctx* ctx, char** seq, int* len, int numseq, int n_threads, int type,
                        float gpo, float gpe, float tgpe, unsigned long long tree_seed, float tree_noise,
                        float dist_scale, float vsm_amax, float use_seq_weights,
                        int consistency_anchors, float consistency_weight, char*** aligned, int* out_aln_len)
{
        KbRunOpts o;
        o.tree_seed = (uint64_t)tree_seed;
        o.tree_noise = tree_noise;
        o.dist_scale = dist_scale;
        o.vsm_amax = vsm_amax;
        o.use_seq_weights = use_seq_weights;
        return kalign_named(ctx, seq, len, numseq, n_threads, type, gpo, gpe, tgpe, consistency_anchors, consistency_weight,
                            aligned, out_aln_len, o);
}

// the code tables of convert_msa_to_internal (create_alphabet, lib/src/alphabet.c:140-302): letters = 5
// (nucleotide), 13 (reduced protein, guide tree), 23 (protein, alignment); -1 = not in the alphabet.  Host code.
int kb200_alphabet(int letters, signed char* to_internal, int* L)
{
        if (!to_internal || (letters != ALPHA_DNA && letters != ALPHA_RED && letters != ALPHA_AMB)) {
                return KB200_FAIL;
        }
        const Alphabet a = make_alphabet(letters);
        for (int i = 0; i < 128; i++) to_internal[i] = (signed char)a.to_internal[i];
        if (L) *L = a.L;
        return KB200_OK;
}

// create_tasks (bisectingKmeans.c:1084-1114) fills the task list in pre-order -- a node, its left subtree, its
// right subtree -- and only create_msa_tree sorts it by c (sort_tasks, task.c:114).  Host code.
int kb200_tasks_creation_order(const int* tasks_sorted, int ntasks, int nseq, int* tasks_out)
{
        if (!tasks_sorted || !tasks_out || ntasks < 0 || nseq < 1) {
                return KB200_FAIL;
        }
        for (int t = 0; t < ntasks; t++) {
                if (tasks_sorted[3 * t + 2] != nseq + t) {
                        return KB200_FAIL;       // node ids are nseq + task index in the sorted list
                }
        }
        if (ntasks == 0) {
                return KB200_OK;
        }
        std::vector<int> stack;
        stack.push_back(nseq + ntasks - 1);
        int k = 0;
        while (!stack.empty()) {
                const int node = stack.back();
                stack.pop_back();
                if (node < nseq) {
                        continue;
                }
                const int t = node - nseq;
                if (t < 0 || t >= ntasks || k >= ntasks) {
                        return KB200_FAIL;
                }
                tasks_out[3 * k] = tasks_sorted[3 * t];
                tasks_out[3 * k + 1] = tasks_sorted[3 * t + 1];
                tasks_out[3 * k + 2] = node;
                k++;
                stack.push_back(tasks_sorted[3 * t + 1]);     // right is visited after the whole left subtree
                stack.push_back(tasks_sorted[3 * t]);
        }
        return k == ntasks ? KB200_OK : KB200_FAIL;
}

// build_tree_kmeans / build_tree_kmeans_noisy (bisectingKmeans.c:177,76) on plain arrays: the whole guide tree on
// the GPU -- anchor distances (bpm), bisecting k-means, UPGMA of the leaf clusters -- for callers that hold the
// sequences in the tree alphabet (the drop-in's seam).  tasks_abc: (nseq - 1) x 3 in the reference's creation order.
int kb200_guide_tree(kb200_ctx* ctx, const uint8_t* seqs, const int64_t* offs, const int* lens, int nseq, int n_threads,
                     unsigned long long tree_seed, float tree_noise, int* tasks_abc, float* seq_distances)
{
        if (!ctx || !seqs || !offs || !lens || nseq < 2 || !tasks_abc) {
                return KB200_FAIL;
        }
        KB_CUDA(cudaSetDevice(ctx->device));
        if (n_threads < 1) n_threads = 1;
        n_threads = std::min(n_threads, kb_default_threads());
        KbSeqs S;
        int rc = S.upload(ctx, seqs, offs, lens, nseq);
        std::vector<int> abc;
        std::vector<float> sd;
        if (rc == KB200_OK) {
                rc = kb_build_tree(ctx, S, n_threads, abc, sd, (uint64_t)tree_seed, tree_noise);
        }
        S.release();
        if (rc != KB200_OK || (int)abc.size() != 3 * (nseq - 1)) {
                return KB200_FAIL;
        }
        if (seq_distances) {
                memcpy(seq_distances, sd.data(), sizeof(float) * (size_t)nseq);
        }
        return kb200_tasks_creation_order(abc.data(), nseq - 1, nseq, tasks_abc);
}

// multiplicative factors of build_tree_kmeans_noisy, in the order they are applied (row by row)
int kb200_tree_noise(unsigned long long seed, float sigma, long long n, float* out)
{
        if (!out || n < 0 || seed == 0) {
                return KB200_FAIL;
        }
        KbNoise rng((uint64_t)seed);
        for (long long i = 0; i < n; i++) {
                out[i] = rng.factor(sigma);
        }
        return KB200_OK;
}

// resolve_run_params (ensemble.c:55-76): run 0 = the base penalties, no noise; run k > 0 = base penalties scaled by
// entry k % 12 of the reference's table (ensemble.c:33-46) and a noisy guide tree seeded seed + k
int kb200_ensemble_run_params(float base_gpo, float base_gpe, float base_tgpe, int run, unsigned long long seed,
                              float* gpo, float* gpe, float* tgpe, unsigned long long* tree_seed, float* tree_noise)
{
        static const float tbl[12][4] = {
                {1.0f, 1.0f, 1.0f, 0.0f},  {0.5f, 1.5f, 0.8f, 0.20f}, {1.5f, 0.5f, 1.2f, 0.20f}, {0.7f, 0.7f, 0.5f, 0.25f},
                {1.4f, 1.4f, 1.5f, 0.25f}, {0.8f, 1.2f, 1.0f, 0.30f}, {1.3f, 0.8f, 0.7f, 0.30f}, {0.6f, 1.0f, 1.3f, 0.15f},
                {1.0f, 0.6f, 0.6f, 0.15f}, {1.8f, 1.0f, 1.0f, 0.35f}, {1.0f, 1.8f, 1.8f, 0.35f}, {0.4f, 0.4f, 0.3f, 0.20f}};
        if (run < 0 || !gpo || !gpe || !tgpe || !tree_seed || !tree_noise) {
                return KB200_FAIL;
        }
        if (run == 0) {
                *gpo = base_gpo;
                *gpe = base_gpe;
                *tgpe = base_tgpe;
                *tree_seed = 0;
                *tree_noise = 0.0f;
        } else {
                const float* e = tbl[run % 12];
                *gpo = base_gpo * e[0];
                *gpe = base_gpe * e[1];
                *tgpe = base_tgpe * e[2];
                *tree_seed = seed + (unsigned long long)run;
                *tree_noise = e[3];
        }
        return KB200_OK;
}

// one of the independent alignment runs of kalign_ensemble (ensemble.c:286-340).  The runs share nothing:
// a caller with several GPUs (one process per GPU) gives run k to rank k % world -- no collective on the data
// path -- and collects the rows for the reference's POAR consensus / scoring, which stays host code.
int kb200_ensemble_run(kb200_ctx* ctx, char** seq, int* len, int numseq, int n_threads, int type,
                       float gpo, float gpe, float tgpe, int run, unsigned long long seed,
                       float dist_scale, float vsm_amax, float use_seq_weights,
                       int consistency_anchors, float consistency_weight, char*** aligned, int* out_aln_len)
{
        if (!ctx || !seq || !len || numseq < 2 || run < 0) {
                return KB200_FAIL;
        }
        // base penalties: aln_param_init's defaults for the detected alphabet where the caller passes < 0
        // (ensemble.c:268-273).  The alphabet is detected from the letters as kb200_kalign does.
        int freq[128];
        memset(freq, 0, sizeof(freq));
        for (int i = 0; i < numseq; i++) {
                for (int j = 0; j < len[i]; j++) {
                        const int ch = (int)(unsigned char)seq[i][j];
                        if (ch < 128) freq[ch]++;
                }
        }
        const int biotype = kb_detect_biotype(freq);
        if (biotype == 2) {
                fprintf(stderr, "[kalign_b200] Unable to determine what alphabet to use.\n");
                return KB200_FAIL;
        }
        kb200_params base;
        KB_RUN(kb200_params_init(&base, biotype, type == KB200_TYPE_PROTEIN_PFASUM_AUTO ? KB200_TYPE_PROTEIN_PFASUM43 : type, gpo, gpe, tgpe));
        float rg, re, rt, noise;
        unsigned long long ts;
        KB_RUN(kb200_ensemble_run_params(base.gpo, base.gpe, base.tgpe, run, seed, &rg, &re, &rt, &ts, &noise));
        if (use_seq_weights < 0.0f) use_seq_weights = 0.0f;      // ensemble.c:248-250
        return kb200_kalign_seeded(ctx, seq, len, numseq, n_threads, type, rg, re, rt, ts, noise, dist_scale, vsm_amax, use_seq_weights,
                                   consistency_anchors, consistency_weight, aligned, out_aln_len);
}

// the CLI's main path for one FASTA file (src/run_kalign.c:395-470: kalign_read_input -> kalign_run_seeded ->
// kalign_write_msa, default output format), file to file
int kb200_kalign_file(kb200_ctx* ctx, const char* infile, const char* outfile, int n_threads, int type,
                      float gpo, float gpe, float tgpe, int consistency_anchors, float consistency_weight)
{
        if (!ctx || !infile || !outfile) {
                return KB200_FAIL;
        }
        kb200_fasta* f = nullptr;
        KB_RUN(kb200_fasta_read(infile, n_threads, &f));
        const int n = kb200_fasta_numseq(f);
        if (n < 2) {
                // check_for_sequences, msa_io.c:176-191
                fprintf(stderr, "[kalign_b200] %s\n", n == 1 ? "Only 1 sequence was found in the input files or standard input"
                                                             : "No sequences were found in the input files or standard input.");
                kb200_fasta_free(f);
                return KB200_FAIL;
        }
        std::vector<char*> seqs((size_t)n);
        std::vector<int> lens((size_t)n);
        std::vector<const char*> names((size_t)n), kept;
        for (int i = 0; i < n; i++) {
                const char* sq = nullptr;
                kb200_fasta_get(f, i, &names[(size_t)i], &sq, &lens[(size_t)i], nullptr);
                seqs[(size_t)i] = (char*)sq;
                if (lens[(size_t)i] > 0) kept.push_back(names[(size_t)i]);      // empty records are dropped (msa_check.c:66-139)
        }
        char** rows = nullptr;
        int alnlen = 0;
        KbRunOpts fo;
        fo.names = names.data();
        fo.file_letter_freq = kb200_fasta_letter_freq(f);
        int rc = kalign_named(ctx, seqs.data(), lens.data(), n, n_threads, type, gpo, gpe, tgpe, consistency_anchors,
                              consistency_weight, &rows, &alnlen, fo);
        if (rc == KB200_OK) {
                rc = kb200_fasta_write(outfile, kept.data(), rows, (int)kept.size(), alnlen, n_threads);
        }
        if (rows) {
                for (size_t i = 0; i < kept.size(); i++) free(rows[i]);
                free(rows);
        }
        kb200_fasta_free(f);
        return rc;
}

} // extern "C"
