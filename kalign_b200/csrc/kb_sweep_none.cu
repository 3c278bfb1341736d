// kb_sweep_none.cu -- kb_sweep_kernel<BONUS_NONE>: no consistency bonus (anchor batch, --fast): 24 KB of rings, tables and staged score vectors per CTA.
#include "kb_sweep.cuh"

cudaError_t kb_sweep_launch_none(int grid, int block, cudaStream_t st, const KbJob* jobs, const KbBox* boxes, const void* units,
                                 KbRound* rnd, unsigned tag_base, const float* tbl, int tstride)
{
        // 4 resident CTAs per SM
        static bool carveout_set = false;
        if (!carveout_set) {
                cudaFuncSetAttribute(kb_sweep_kernel<BONUS_NONE>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
                carveout_set = true;
        }
        kb_sweep_kernel<BONUS_NONE><<<grid, block, 0, st>>>(jobs, boxes, static_cast<const KbUnit*>(units), rnd, tag_base, tbl, tstride);
        return cudaGetLastError();
}
