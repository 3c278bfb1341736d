// kb_sweep_sparse.cu -- kb_sweep_kernel<BONUS_SPARSE>: sparse per-row bonus lists (tree levels in default mode): + 32 KB of staged lists per CTA, 4 x 47 KB = the whole carve-out.
#include "kb_sweep.cuh"

cudaError_t kb_sweep_launch_sparse(int grid, int block, cudaStream_t st, const KbJob* jobs, const KbBox* boxes, const void* units,
                                   KbRound* rnd, unsigned tag_base, const float* tbl, int tstride)
{
        // 4 resident CTAs per SM
        static bool carveout_set = false;
        if (!carveout_set) {
                cudaFuncSetAttribute(kb_sweep_kernel<BONUS_SPARSE>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
                carveout_set = true;
        }
        kb_sweep_kernel<BONUS_SPARSE><<<grid, block, 0, st>>>(jobs, boxes, static_cast<const KbUnit*>(units), rnd, tag_base, tbl, tstride);
        return cudaGetLastError();
}
