// kmeans_driver.cpp -- CPU check of kalign_b200/csrc/kb_kmeans.h (test infrastructure).
// Runs the bisecting k-means of the product header on a synthetic anchor-distance matrix and prints a
// hash of the resulting tree in canonical form (leaf clusters as sorted sample lists, left/right kept).
// tests/test_kmeans_host.py compiles this at -O2 and at the product's -O3 -mavx2 and runs it with
// several thread counts: the hash must not change, and it must equal the tree of the literal
// restatement of bisecting_kmeans kept below (seeds strictly in batches of 4).
#include "kb_kmeans.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <string>

namespace {

std::string canon(const TreeBuilder& B, int node)
{
        for (const Cluster& c : B.clusters) {
                if (c.placeholder == node) {
                        std::vector<int> v = c.samples;
                        std::sort(v.begin(), v.end());
                        std::string s = "[";
                        for (int x : v) { s += std::to_string(x); s += ' '; }
                        return s + "]";
                }
        }
        const Node& n = B.nodes[(size_t)node];
        return "(" + canon(B, n.left) + "," + canon(B, n.right) + ")";
}

// ground truth: the literal restatement of bisecting_kmeans (seeds in batches of 4, as the reference)
// bisecting_kmeans, bisectingKmeans.c:273-406.  Returns the node index of the sub-tree; leaf
// clusters (< 50 samples) are recorded and resolved by UPGMA after one batched distance launch.
int bisect_truth(TreeBuilder& B, std::vector<int>& samples)
{
        const int ns = (int)samples.size();
        if (ns < 50) {
                int slot;
#ifdef _OPENMP
#pragma omp critical(kb_tree_nodes)
#endif
                {
                        slot = (int)B.nodes.size();
                        B.nodes.push_back(Node());
                        Cluster c;
                        c.samples.swap(samples);
                        c.placeholder = slot;
                        B.clusters.push_back(std::move(c));
                }
                return slot;
        }
        const int tries = std::min(40, ns);
        const int step = ns / tries;
        Split best;
        bool have_best = false;
        for (int i = 0; i < tries; i += 4) {
                Split res[4];
#ifdef _OPENMP
#pragma omp taskloop if (ns > 2000) default(shared) grainsize(1)
#endif
                for (int j = 0; j < 4; j++) {
                        split2(B, samples, (i + j) * step, res[j]);
                }
                int change = 0;
                for (int j = 0; j < 4; j++) {
                        if (!have_best) {
                                best = std::move(res[j]);
                                have_best = true;
                                change++;
                        } else if (best.score > res[j].score) {
                                std::swap(best, res[j]);
                                change++;
                        }
                }
                if (!change) break;
        }
        std::vector<int>().swap(samples);
        int l = -1, r = -1;
#ifdef _OPENMP
#pragma omp task shared(B, best, l) if (ns > 2000)
#endif
        l = bisect_truth(B, best.sl);
#ifdef _OPENMP
#pragma omp task shared(B, best, r) if (ns > 2000)
#endif
        r = bisect_truth(B, best.sr);
#ifdef _OPENMP
#pragma omp taskwait
#endif
        int slot;
#ifdef _OPENMP
#pragma omp critical(kb_tree_nodes)
#endif
        {
                slot = (int)B.nodes.size();
                Node n;
                n.left = l; n.right = r;
                B.nodes.push_back(n);
        }
        return slot;
}


unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

} // namespace

int main(int argc, char** argv)
{
        const int N = argc > 1 ? atoi(argv[1]) : 5000;
        unsigned seed = argc > 2 ? (unsigned)atoi(argv[2]) : 1u;
        const int na = 32, stride = 32;
        // distance-like rows: a few "families" + noise, runs of exact duplicates (the tie rule of
        // split2), integer edit distances plus a fractional length term like d_estimation produces
        std::vector<float> dm((size_t)N * stride);
        const int fam = 7;
        std::vector<float> centers((size_t)fam * na);
        for (float& c : centers) c = (float)(lcg(seed) % 400);
        for (int i = 0; i < N; i++) {
                const int f = (int)(lcg(seed) % fam);
                for (int j = 0; j < na; j++) {
                        dm[(size_t)i * stride + j] = centers[(size_t)f * na + j] + (float)(lcg(seed) % 60) + 0.0301f * (float)(lcg(seed) % 7);
                }
                if (i > 0 && (lcg(seed) % 9) == 0) {
                        memcpy(&dm[(size_t)i * stride], &dm[(size_t)(i - 1) * stride], sizeof(float) * stride);
                }
        }
        TreeBuilder B;
        B.dm = dm.data(); B.stride = stride; B.num_anchors = na; B.N = N;
        B.nodes.reserve((size_t)2 * N + 64);
        std::vector<int> samples((size_t)N);
        for (int i = 0; i < N; i++) samples[(size_t)i] = i;
        int root = -1;
        const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel
#pragma omp single
        root = bisect(B, samples);
        const double ms = 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const std::string c = canon(B, root);
        {
                TreeBuilder T;
                T.dm = dm.data(); T.stride = stride; T.num_anchors = na; T.N = N;
                T.nodes.reserve((size_t)2 * N + 64);
                std::vector<int> s2((size_t)N);
                for (int i = 0; i < N; i++) s2[(size_t)i] = i;
                int rt = -1;
                const auto u0 = std::chrono::steady_clock::now();
#pragma omp parallel
#pragma omp single
                rt = bisect_truth(T, s2);
                const double ms2 = 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - u0).count();
                if (canon(T, rt) != c) {
                        fprintf(stderr, "tree differs from the literal restatement (N=%d)\n", N);
                        return 2;
                }
                printf("literal restatement: identical tree, %.1f ms\n", ms2);
        }
        unsigned long long h = 1469598103934665603ull;
        for (unsigned char ch : c) { h ^= ch; h *= 1099511628211ull; }
        size_t covered = 0;
        for (const Cluster& cl : B.clusters) covered += cl.samples.size();
        printf("N=%d clusters=%zu covered=%zu tree=%016llx ms=%.1f\n", N, B.clusters.size(), covered, h, ms);
        return covered == (size_t)N ? 0 : 1;
}
