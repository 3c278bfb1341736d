// kb_comm.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
//
// The reference has no distributed backend; its unit of parallelism is the
// OpenMP task per guide-tree node (lib/src/aln_run.c:95-109).  Here the tasks of one tree level
// (and the N x K anchor pairs) are sharded across ranks; the only data-path collectives are
// all-gathers of results that every rank needs next: anchor position maps once, and per level the
// coded paths and the merged sub-profiles.
#include "kb_common.cuh"

#include <nccl.h>
#include <string.h>

void kb_partition(const double* cost, int n, int world, int* bounds)
{
        double total = 0.0;
        for (int i = 0; i < n; i++) total += cost[i];
        bounds[0] = 0;
        double acc = 0.0;
        int i = 0;
        for (int r = 1; r < world; r++) {
                const double target = total * (double)r / (double)world;
                while (i < n && acc + 0.5 * cost[i] <= target) {
                        acc += cost[i];
                        i++;
                }
                bounds[r] = i;
        }
        bounds[world] = n;
        for (int r = 1; r <= world; r++) {
                if (bounds[r] < bounds[r - 1]) bounds[r] = bounds[r - 1];
        }
}

int kb_allgatherv(kb200_ctx* ctx, void* dbuf, const size_t* seg)
{
        if (ctx->world <= 1) {
                return KB200_OK;
        }
        if (!ctx->comm) {
                fprintf(stderr, "[kalign_b200] multi-GPU context without communicator\n");
                return KB200_FAIL;
        }
        ncclComm_t comm = (ncclComm_t)ctx->comm;
        ncclResult_t r = ncclGroupStart();
        for (int root = 0; root < ctx->world && r == ncclSuccess; root++) {
                const size_t nbytes = seg[root + 1] - seg[root];
                if (nbytes == 0) continue;
                char* p = (char*)dbuf + seg[root];
                r = ncclBroadcast(p, p, nbytes, ncclChar, root, comm, ctx->stream);
        }
        if (r == ncclSuccess) r = ncclGroupEnd(); else ncclGroupEnd();
        if (r != ncclSuccess) {
                fprintf(stderr, "[kalign_b200] NCCL error: %s\n", ncclGetErrorString(r));
                return KB200_FAIL;
        }
        ctx->stats.n_collectives += 1;
        ctx->stats.collective_bytes += (double)(seg[ctx->world] - seg[0]);
        return KB200_OK;
}

extern "C" {

int kb200_comm_unique_id(void* id_out, int nbytes)
{
        if (!id_out || nbytes < (int)sizeof(ncclUniqueId)) {
                return KB200_FAIL;
        }
        ncclUniqueId id;
        if (ncclGetUniqueId(&id) != ncclSuccess) {
                return KB200_FAIL;
        }
        memset(id_out, 0, (size_t)nbytes);
        memcpy(id_out, &id, sizeof(id));
        return KB200_OK;
}

int kb200_ctx_comm_init(kb200_ctx* ctx, int rank, int world, const void* id, int nbytes)
{
        if (!ctx || !id || world < 1 || rank < 0 || rank >= world || nbytes < (int)sizeof(ncclUniqueId)) {
                return KB200_FAIL;
        }
        KB_CUDA(cudaSetDevice(ctx->device));
        if (world == 1) {
                ctx->rank = 0; ctx->world = 1;
                return KB200_OK;
        }
        ncclUniqueId uid;
        memcpy(&uid, id, sizeof(uid));
        ncclComm_t comm;
        ncclResult_t r = ncclCommInitRank(&comm, world, uid, rank);
        if (r != ncclSuccess) {
                fprintf(stderr, "[kalign_b200] ncclCommInitRank failed: %s\n", ncclGetErrorString(r));
                return KB200_FAIL;
        }
        ctx->comm = comm;
        ctx->rank = rank;
        ctx->world = world;
        return KB200_OK;
}

void kb200_ctx_comm_destroy(kb200_ctx* ctx)
{
        if (ctx && ctx->comm) {
                ncclCommDestroy((ncclComm_t)ctx->comm);
                ctx->comm = nullptr;
                ctx->rank = 0;
                ctx->world = 1;
        }
}

// host-only helper (also used by the CPU tests of the sharding logic)
int kb200_partition(const double* cost, int n, int world, int* bounds)
{
        if (!cost || !bounds || n < 0 || world < 1) {
                return KB200_FAIL;
        }
        kb_partition(cost, n, world, bounds);
        return KB200_OK;
}

} // extern "C"
