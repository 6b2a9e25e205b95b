/* libacn_b200_comm.so -- C ABI of the multi-GPU exchange of the adaptive-city-nerf hot path for hosts that are NOT
 * PyTorch (acn_comm_init / acn_alltoall_samples / acn_allreduce).  One process per GPU; NCCL over NVLink 5 /
 * NVSwitch underneath.  The Python package itself uses torch.distributed + symmetric memory for the same exchanges
 * (adaptive_city_nerf_b200/distributed.py); this library is the equivalent entry for a C / C++ host and links libnccl,
 * which is why it is separate from libacn_b200.so (no NCCL dependency in the product library).
 *
 * What it replaces in the reference: nothing -- the reference is single-GPU; the exchanges are the ones BASELINE's
 * north star names: routed sample rows to the GPU that owns their expert and back
 * (models/inr/meta_container.py:306-337 split across ranks), and the gradient all-reduce of single-expert data parallel
 * training (pipelines/offline_stage/meta_core.py:123-141 before the optimizer step).
 *
 * Every call returns 0 or a negative ACN_E* code (include/acn_b200.h); acn_comm_last_error() is thread-local.
 * Collectives are asynchronous on the caller's stream. */
#ifndef ACN_B200_COMM_H
#define ACN_B200_COMM_H
#include <stdint.h>
#include "acn_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct acn_comm acn_comm;
#define ACN_COMM_ID_BYTES 128
#define ACN_OP_SUM 0
#define ACN_OP_MAX 1

const char* acn_comm_last_error(void);

/* Rank 0 makes the 128-byte rendezvous id (ncclGetUniqueId) and ships it to the other ranks over any host channel. */
int acn_comm_unique_id(void* id_out_128);
/* Collective over all `world` ranks; `device` is this process's GPU. */
int acn_comm_init(int device, const void* unique_id_128, int rank, int world, acn_comm** out);
int acn_comm_destroy(acn_comm*);
int acn_comm_rank(const acn_comm*, int* rank, int* world);

/* In-place all-reduce of n elements (dtype ACN_F32 | ACN_F16; op ACN_OP_SUM | ACN_OP_MAX): hash-table and MLP gradients
 * of data-parallel training, the squared gradient norm of a sharded clip. */
int acn_allreduce(acn_comm*, void* buf, int64_t n, int dtype, int op, acn_stream);

/* Every rank contributes `bytes` bytes; recv holds world * bytes in rank order (the per-expert row counts). */
int acn_allgather(acn_comm*, const void* send, void* recv, int64_t bytes, acn_stream);

/* Variable-size all-to-all of routed sample rows (row_bytes each: 24 B [xyz, dir] out, 16 B [rgb, sigma] back).
 * send holds the rows for rank 0, then rank 1, ... (send_counts[r] rows each, HOST array of length world); recv
 * receives recv_counts[r] rows from rank r in rank order.  One grouped ncclSend / ncclRecv per peer. */
int acn_alltoall_samples(acn_comm*, const void* send, const int64_t* send_counts, void* recv, const int64_t* recv_counts,
                         int row_bytes, acn_stream);

#ifdef __cplusplus
}
#endif
#endif
