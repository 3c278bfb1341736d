// kb_sweep_dense.cu -- kb_sweep_kernel<BONUS_DENSE>: caller-supplied dense bonus matrix (kb200_pair_align_batch).
#include "kb_sweep.cuh"

cudaError_t kb_sweep_launch_dense(int grid, int block, cudaStream_t st, const KbJob* jobs, const KbBox* boxes, const void* units,
                                  KbRound* rnd, unsigned tag_base, const float* tbl, int tstride)
{
        // 4 resident CTAs per SM
        static bool carveout_set = false;
        if (!carveout_set) {
                cudaFuncSetAttribute(kb_sweep_kernel<BONUS_DENSE>, cudaFuncAttributePreferredSharedMemoryCarveout, 40);
                carveout_set = true;
        }
        kb_sweep_kernel<BONUS_DENSE><<<grid, block, 0, st>>>(jobs, boxes, static_cast<const KbUnit*>(units), rnd, tag_base, tbl, tstride);
        return cudaGetLastError();
}
