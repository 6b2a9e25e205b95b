// libacn_b200_comm.so: NCCL-backed exchange entries for non-PyTorch hosts (include/acn_b200_comm.h).
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include "../../../include/acn_b200_comm.h"

struct acn_comm {
    ncclComm_t nccl;
    int device, rank, world;
};

namespace {
thread_local char g_err[512] = "";
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define NCCL_TRY(call)                                                                              \
    do {                                                                                            \
        ncclResult_t r__ = (call);                                                                  \
        if (r__ != ncclSuccess) return fail(ACN_ECUDA, "%s: %s failed: %s", __func__, #call, ncclGetErrorString(r__)); \
    } while (0)
#define CUDA_TRY(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) return fail(ACN_ECUDA, "%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__)); \
    } while (0)
int check_comm(const acn_comm* c, const char* fn) {
    if (!c) return fail(ACN_EINVAL, "%s: null communicator", fn);
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != c->device)
        return fail(ACN_EINVAL, "%s: the communicator belongs to device %d but device %d is current", fn, c->device, cur);
    return ACN_OK;
}
}  // namespace

extern "C" {

__attribute__((visibility("default"))) const char* acn_comm_last_error(void) { return g_err; }

__attribute__((visibility("default"))) int acn_comm_unique_id(void* id_out_128) {
    if (!id_out_128) return fail(ACN_EINVAL, "acn_comm_unique_id: null output");
    static_assert(sizeof(ncclUniqueId) == ACN_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCCL_TRY(ncclGetUniqueId(&id));
    memcpy(id_out_128, &id, sizeof id);
    return ACN_OK;
}

__attribute__((visibility("default"))) int acn_comm_init(int device, const void* unique_id_128, int rank, int world, acn_comm** out) {
    if (!unique_id_128 || !out) return fail(ACN_EINVAL, "acn_comm_init: null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(ACN_EINVAL, "acn_comm_init: bad rank %d / world %d", rank, world);
    CUDA_TRY(cudaSetDevice(device));
    ncclUniqueId id;
    memcpy(&id, unique_id_128, sizeof id);
    ncclComm_t c;
    NCCL_TRY(ncclCommInitRank(&c, world, id, rank));
    *out = new acn_comm{ c, device, rank, world };
    return ACN_OK;
}

__attribute__((visibility("default"))) int acn_comm_destroy(acn_comm* c) {
    if (!c) return ACN_OK;
    ncclCommDestroy(c->nccl);
    delete c;
    return ACN_OK;
}

__attribute__((visibility("default"))) int acn_comm_rank(const acn_comm* c, int* rank, int* world) {
    if (!c) return fail(ACN_EINVAL, "acn_comm_rank: null communicator");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    return ACN_OK;
}

__attribute__((visibility("default"))) int acn_allreduce(acn_comm* c, void* buf, int64_t n, int dtype, int op, acn_stream stream) {
    int rc = check_comm(c, "acn_allreduce");
    if (rc) return rc;
    if (n < 0 || (n > 0 && !buf)) return fail(ACN_EINVAL, "acn_allreduce: bad buffer / count");
    if (dtype != ACN_F32 && dtype != ACN_F16) return fail(ACN_EINVAL, "acn_allreduce: dtype must be ACN_F32 or ACN_F16");
    if (op != ACN_OP_SUM && op != ACN_OP_MAX) return fail(ACN_EINVAL, "acn_allreduce: op must be ACN_OP_SUM or ACN_OP_MAX");
    if (n == 0) return ACN_OK;
    NCCL_TRY(ncclAllReduce(buf, buf, (size_t)n, dtype == ACN_F32 ? ncclFloat32 : ncclFloat16, op == ACN_OP_SUM ? ncclSum : ncclMax,
                           c->nccl, (cudaStream_t)stream));
    return ACN_OK;
}

__attribute__((visibility("default"))) int acn_allgather(acn_comm* c, const void* send, void* recv, int64_t bytes, acn_stream stream) {
    int rc = check_comm(c, "acn_allgather");
    if (rc) return rc;
    if (bytes < 0 || (bytes > 0 && (!send || !recv))) return fail(ACN_EINVAL, "acn_allgather: bad buffer / size");
    if (bytes == 0) return ACN_OK;
    NCCL_TRY(ncclAllGather(send, recv, (size_t)bytes, ncclChar, c->nccl, (cudaStream_t)stream));
    return ACN_OK;
}

__attribute__((visibility("default"))) int acn_alltoall_samples(acn_comm* c, const void* send, const int64_t* send_counts, void* recv,
                                                               const int64_t* recv_counts, int row_bytes, acn_stream stream) {
    int rc = check_comm(c, "acn_alltoall_samples");
    if (rc) return rc;
    if (!send_counts || !recv_counts || row_bytes < 1) return fail(ACN_EINVAL, "acn_alltoall_samples: bad counts / row size");
    int64_t ts = 0, tr = 0;
    for (int r = 0; r < c->world; ++r) {
        if (send_counts[r] < 0 || recv_counts[r] < 0) return fail(ACN_EINVAL, "acn_alltoall_samples: negative count for rank %d", r);
        ts += send_counts[r]; tr += recv_counts[r];
    }
    if ((ts > 0 && !send) || (tr > 0 && !recv)) return fail(ACN_EINVAL, "acn_alltoall_samples: null buffer");
    const char* s = (const char*)send;
    char* d = (char*)recv;
    NCCL_TRY(ncclGroupStart());
    for (int r = 0; r < c->world; ++r) {
        const size_t sb = (size_t)send_counts[r] * row_bytes, rb = (size_t)recv_counts[r] * row_bytes;
        if (sb) NCCL_TRY(ncclSend(s, sb, ncclChar, r, c->nccl, (cudaStream_t)stream));
        if (rb) NCCL_TRY(ncclRecv(d, rb, ncclChar, r, c->nccl, (cudaStream_t)stream));
        s += sb; d += rb;
    }
    NCCL_TRY(ncclGroupEnd());
    return ACN_OK;
}

}  // extern "C"
