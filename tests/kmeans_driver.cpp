// kmeans_driver.cpp -- CPU check of kalign_b200/csrc/kb_kmeans.h (test infrastructure).
// Runs the bisecting k-means of the product header on a synthetic anchor-distance matrix and prints a
// hash of the resulting tree in canonical form (leaf clusters as sorted sample lists, left/right kept).
// tests/test_kmeans_host.py compiles this at -O2 and at the product's -O3 -mavx2 and runs it with
// several thread counts: the hash must not change.
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
        unsigned long long h = 1469598103934665603ull;
        for (unsigned char ch : c) { h ^= ch; h *= 1099511628211ull; }
        size_t covered = 0;
        for (const Cluster& cl : B.clusters) covered += cl.samples.size();
        printf("N=%d clusters=%zu covered=%zu tree=%016llx ms=%.1f\n", N, B.clusters.size(), covered, h, ms);
        return covered == (size_t)N ? 0 : 1;
}
