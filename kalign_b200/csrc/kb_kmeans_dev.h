// kb_kmeans_dev.h -- interface of the device bisecting k-means (kb_kmeans.cu), plain types only
#pragma once
#include <vector>

struct kb200_ctx;

// the k-means part of the guide tree: internal nodes with two children, leaves = clusters of fewer
// than 50 samples (resolved by UPGMA afterwards); samples of leaf l are order[leaf_begin[l] .. leaf_end[l])
struct KbKmeansTree {
        std::vector<int> left, right;            // per node, -1 for a leaf cluster
        std::vector<int> leaf_node, leaf_begin, leaf_end;
        std::vector<int> order;                  // the N samples, grouped by leaf cluster, sample order kept
        int root = -1;
};

int kb_kmeans_bisect_dev(kb200_ctx* ctx, const float* dm_host, int N, KbKmeansTree& out);

// UPGMA of every leaf cluster on the device (kb_upgma_kernel); d_pd = pair distances on the device
int kb_upgma_dev(kb200_ctx* ctx, const float* d_pd, const std::vector<int>& csize, const std::vector<long long>& pair0,
                 const std::vector<long long>& merge0, long long nmerges, int* merges_out);
