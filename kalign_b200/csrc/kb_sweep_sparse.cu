// kb_sweep_sparse.cu -- kb_sweep_kernel<BONUS_SPARSE>: sparse per-row bonus lists (tree levels in default mode): per-lane event queues in dynamic shared memory (38 KB per CTA; 4 CTAs x 52 KB per SM).
#include "kb_sweep.cuh"

cudaError_t kb_sweep_launch_sparse(int grid, int block, cudaStream_t st, const KbJob* jobs, const KbBox* boxes, const void* units,
                                   KbRound* rnd, unsigned tag_base, const float* tbl, int tstride)
{
        // 4 resident CTAs per SM
        static bool carveout_set = false;
        if (!carveout_set) {
                cudaFuncSetAttribute(kb_sweep_kernel<BONUS_SPARSE>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
                cudaFuncSetAttribute(kb_sweep_kernel<BONUS_SPARSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, BON_DYN_SMEM);
                carveout_set = true;
        }
        kb_sweep_kernel<BONUS_SPARSE><<<grid, block, BON_DYN_SMEM, st>>>(jobs, boxes, static_cast<const KbUnit*>(units), rnd, tag_base, tbl, tstride);
        return cudaGetLastError();
}
