// kb_kmeans.h -- host side of the guide tree: bisecting k-means over the anchor-distance rows
// (build_tree_kmeans / bisecting_kmeans / split2, lib/src/bisectingKmeans.c:177-971; edist_256's
// summation order, lib/src/euclidean_dist.c:161-206).  Plain C++ (no CUDA), so that
// tests/test_kmeans_host.py can compile it with g++ and check that the bisection tree does not depend
// on the optimisation level or the thread count: this file is built -O3 -mavx2 (the 8-lane distance
// loop vectorises like the reference's intrinsics; 1.7x faster than -O2 on the 100 000-sequence shape).
// Every float operation keeps the reference's order; compile with -ffp-contract=off, never -ffast-math.
#pragma once
#include <float.h>
#include <math.h>

#include <algorithm>
#include <utility>
#include <vector>

namespace {

// ---- guide tree ------------------------------------------------------------------------------
struct Node {
        int left = -1, right = -1;
        int id = -1;
};

struct Cluster {
        std::vector<int> samples;
        int root = -1;          // node index of the UPGMA sub-tree (filled later)
        int placeholder = -1;   // node slot that stands for this cluster in the k-means tree
};

struct TreeBuilder {
        std::vector<Node> nodes;
        std::vector<Cluster> clusters;
        const float* dm = nullptr;      // N x stride
        int stride = 0;
        int num_anchors = 0;
        int N = 0;
};

inline int cmp_floats(float a, float b)               // bisectingKmeans.c:63-73
{
        const float epsilon = 1e-6;
        if (fabsf(a - b) < epsilon) return 0;
        return (a > b) ? 1 : -1;
}

// edist_256: 8 lanes accumulate (a-b)^2 over chunks of 8, then the AVX horizontal sum order
// ((l0+l4)+(l1+l5)) + ((l2+l6)+(l3+l7)), then sqrtf  (euclidean_dist.c:161-206)
inline float edist8(const float* a, const float* b, int len)
{
        float r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < len; i += 8) {
                for (int l = 0; l < 8; l++) {
                        const float t = a[i + l] - b[i + l];
                        const float t2 = t * t;
                        r[l] = r[l] + t2;
                }
        }
        const float s0 = r[0] + r[4], s1 = r[1] + r[5], s2 = r[2] + r[6], s3 = r[3] + r[7];
        const float d = (s0 + s1) + (s2 + s3);
        return sqrtf(d);
}

struct Split {
        std::vector<int> sl, sr;
        float score = FLT_MAX;
};

// split2, bisectingKmeans.c:766-971
void split2(const TreeBuilder& B, const std::vector<int>& samples, int seed_pick, Split& res)
{
        const int na = B.num_anchors;
        const int num_var = ((na + 7) / 8) * 8;
        const int ns = (int)samples.size();
        std::vector<float> w(num_var, 0.0f), wl(num_var, 0.0f), wr(num_var, 0.0f), cl(num_var, 0.0f), cr(num_var, 0.0f);
        res.sl.resize(ns);
        res.sr.resize(ns);
        for (int i = 0; i < ns; i++) {
                const float* row = B.dm + (size_t)samples[i] * B.stride;
                for (int j = 0; j < na; j++) w[j] += row[j];
        }
        for (int j = 0; j < na; j++) w[j] /= (float)ns;
        {
                const float* row = B.dm + (size_t)samples[seed_pick] * B.stride;
                for (int j = 0; j < na; j++) cl[j] = row[j];
        }
        for (int j = 0; j < na; j++) cr[j] = w[j] - (cl[j] - w[j]);
        float* pcl = cl.data(); float* pcr = cr.data(); float* pwl = wl.data(); float* pwr = wr.data();
        int num_l = 0, num_r = 0;
        float score = 0.0f;
        for (int stop = 0; stop < 500; stop++) {
                num_l = 0; num_r = 0;
                for (int i = 0; i < na; i++) { pwr[i] = 0.0f; pwl[i] = 0.0f; }
                score = 0.0f;
                for (int i = 0; i < ns; i++) {
                        const int s = samples[i];
                        const float* row = B.dm + (size_t)s * B.stride;
                        const float dl = edist8(row, pcl, na);
                        const float dr = edist8(row, pcr, na);
                        score += (dl < dr) ? dl : dr;
                        const int r = cmp_floats(dr, dl);
                        float* wsel;
                        if (r == -1) { wsel = pwr; res.sr[num_r++] = s; }
                        else if (r == 1) { wsel = pwl; res.sl[num_l++] = s; }
                        else if (i & 1) { wsel = pwr; res.sr[num_r++] = s; }
                        else { wsel = pwl; res.sl[num_l++] = s; }
                        for (int j = 0; j < na; j++) wsel[j] += row[j];
                }
                if (num_l == 0 || num_r == 0) {
                        score = 0.0f;
                        num_l = 0; num_r = 0;
                        for (int i = 0; i < ns / 2; i++) res.sl[num_l++] = samples[i];
                        for (int i = ns / 2; i < ns; i++) res.sr[num_r++] = samples[i];
                        break;
                }
                for (int j = 0; j < na; j++) {
                        pwl[j] /= (float)num_l;
                        pwr[j] /= (float)num_r;
                }
                int changed = 0;
                for (int j = 0; j < na; j++) {
                        if (cmp_floats(pwl[j], pcl[j]) != 0) { changed = 1; break; }
                        if (cmp_floats(pwr[j], pcr[j]) != 0) { changed = 1; break; }
                }
                if (!changed) break;
                std::swap(pcl, pwl);
                std::swap(pcr, pwr);
        }
        res.sl.resize(num_l);
        res.sr.resize(num_r);
        res.score = score;
}

// bisecting_kmeans, bisectingKmeans.c:273-406.  Returns the node index of the sub-tree; leaf
// clusters (< 50 samples) are recorded and resolved by UPGMA after one batched distance launch.
int bisect(TreeBuilder& B, std::vector<int>& samples)
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
        l = bisect(B, best.sl);
#ifdef _OPENMP
#pragma omp task shared(B, best, r) if (ns > 2000)
#endif
        r = bisect(B, best.sr);
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


} // namespace
