/* vsf_nccl.h — C ABI of libvsf_nccl.so: the only multi-GPU exchange of the path.
 *
 * The matching path shards by independent units: every rank owns a contiguous
 * pose range of the sequence and runs it on its own GPU through libvsf_cuda.so with no
 * data-path collective.  What remains is collecting results: the per-rank match lists
 * (variable length, device-resident) and, for a sharded slam::Frontend run, the pieces of the
 * SLAMProblem message (src/slam_frontend_main.cc:369-374 writes ONE message).  Both are NCCL
 * collectives over NVLink / NVSwitch, kept in a library of their own so that single-GPU users
 * of libvsf_cuda.so do not need NCCL at all.
 */
#ifndef VSF_NCCL_H_
#define VSF_NCCL_H_

#include <stddef.h>
#include <stdint.h>

#include "vsf.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vsf_nccl_comm vsf_nccl_comm;

#define VSF_NCCL_UNIQUE_ID_BYTES 128

/* Rank 0 makes an id (ncclGetUniqueId) and hands its 128 bytes to the other ranks by whatever
 * the launcher offers (a file, MPI, torch.distributed ...); every rank then creates the
 * communicator for the CUDA device it drives. */
int vsf_nccl_unique_id(char id[VSF_NCCL_UNIQUE_ID_BYTES]);
int vsf_nccl_comm_create(const char id[VSF_NCCL_UNIQUE_ID_BYTES], int world, int rank, int device,
                         vsf_nccl_comm** out);
/* Wrap a communicator the application already has (ncclComm_t passed as void*; not destroyed by
 * vsf_nccl_comm_destroy). */
int vsf_nccl_comm_adopt(void* nccl_comm, int world, int rank, int device, vsf_nccl_comm** out);
void vsf_nccl_comm_destroy(vsf_nccl_comm* comm);
const char* vsf_nccl_last_error(const vsf_nccl_comm* comm);

/* All-gather of the match lists of ctx's most recent window launch (vsf_window_match_device /
 * vsf_window_match_block_device / any host-API window call), straight from the device regions
 * the compaction kernel wrote - no host staging, no repack: one ncclAllGather of the n_frames
 * list lengths and one of the n_frames list regions, on the ctx's stream.  A region holds
 * `stride` records (vsf_device_match_lists reports it), of which the first counts[] are valid.
 * d_counts: device int32 [world][n_frames]; d_lists: device vsf_dmatch [world][n_frames][stride].
 * Asynchronous (vsf_synchronize(ctx) to wait). */
int vsf_gather_matches(vsf_ctx* ctx, vsf_nccl_comm* comm, int n_frames, int* d_counts, vsf_dmatch* d_lists);

/* Variable-length byte blobs to rank `root` (host memory in, host memory out; staged through
 * device buffers because NCCL moves device memory): sizes[r] = bytes rank r contributed, *out =
 * malloc'ed concatenation in rank order on root (caller frees), NULL elsewhere.  Used by the
 * sharded-sequence driver to assemble one SLAMProblem from the ranks' pieces. */
int vsf_nccl_gather_bytes(vsf_nccl_comm* comm, const void* data, size_t n, int root, void** out,
                          size_t* sizes);

#ifdef __cplusplus
}
#endif
#endif  /* VSF_NCCL_H_ */
