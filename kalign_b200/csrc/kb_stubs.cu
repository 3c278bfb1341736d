// temporary: entry points not implemented yet fail loudly (no CPU fallback)
#include "kb_common.cuh"
extern "C" {
int kb200_anchor_posmaps(kb200_ctx*, const kb200_params*, const uint8_t*, const int64_t*, const int*, int,
                         const int*, int, long long, long long, int*)
{
        fprintf(stderr, "[kalign_b200] kb200_anchor_posmaps: not implemented\n");
        return KB200_FAIL;
}
int kb200_align_tree(kb200_ctx*, const kb200_params*, const uint8_t*, const int64_t*, const int*, int,
                     const int*, int, const float*, const int*, int, float, int*)
{
        fprintf(stderr, "[kalign_b200] kb200_align_tree: not implemented\n");
        return KB200_FAIL;
}
int kb200_kalign(kb200_ctx*, char**, int*, int, int, int, float, float, float, int, float, char***, int*)
{
        fprintf(stderr, "[kalign_b200] kb200_kalign: not implemented\n");
        return KB200_FAIL;
}
int kb200_distances(kb200_ctx*, const uint8_t*, const int64_t*, const int*, int, const int*, int, const int*, int, float*)
{
        fprintf(stderr, "[kalign_b200] kb200_distances: not implemented\n");
        return KB200_FAIL;
}
}
