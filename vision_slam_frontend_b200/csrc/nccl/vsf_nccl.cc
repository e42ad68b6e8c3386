// libvsf_nccl.so (include/vsf_nccl.h): result gathers of the sharded path over NCCL.
#include "vsf_nccl.h"

#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct vsf_nccl_comm {
  ncclComm_t comm = nullptr;
  bool owned = false;
  int world = 1, rank = 0, device = 0;
  cudaStream_t stream = nullptr;   // for vsf_nccl_gather_bytes
  std::string err;
};

static_assert(VSF_NCCL_UNIQUE_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");

#define VN_NCCL(c, expr)                                                        \
  do {                                                                          \
    ncclResult_t r__ = (expr);                                                  \
    if (r__ != ncclSuccess) {                                                   \
      (c)->err = std::string(#expr) + ": " + ncclGetErrorString(r__);           \
      return VSF_ERR_CUDA;                                                      \
    }                                                                           \
  } while (0)
#define VN_CUDA(c, expr)                                                        \
  do {                                                                          \
    cudaError_t e__ = (expr);                                                   \
    if (e__ != cudaSuccess) {                                                   \
      (c)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);           \
      return VSF_ERR_CUDA;                                                      \
    }                                                                           \
  } while (0)

extern "C" int vsf_nccl_unique_id(char id[VSF_NCCL_UNIQUE_ID_BYTES]) {
  if (!id) return VSF_ERR_BAD_ARG;
  ncclUniqueId u;
  if (ncclGetUniqueId(&u) != ncclSuccess) return VSF_ERR_CUDA;
  std::memcpy(id, u.internal, NCCL_UNIQUE_ID_BYTES);
  return VSF_OK;
}

extern "C" int vsf_nccl_comm_create(const char id[VSF_NCCL_UNIQUE_ID_BYTES], int world, int rank, int device,
                                    vsf_nccl_comm** out) {
  if (!id || !out || world < 1 || rank < 0 || rank >= world) return VSF_ERR_BAD_ARG;
  *out = nullptr;
  if (cudaSetDevice(device) != cudaSuccess) return VSF_ERR_CUDA;
  vsf_nccl_comm* c = new vsf_nccl_comm();
  c->world = world;
  c->rank = rank;
  c->device = device;
  c->owned = true;
  ncclUniqueId u;
  std::memcpy(u.internal, id, NCCL_UNIQUE_ID_BYTES);
  if (ncclCommInitRank(&c->comm, world, u, rank) != ncclSuccess ||
      cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return VSF_ERR_CUDA;
  }
  *out = c;
  return VSF_OK;
}

extern "C" int vsf_nccl_comm_adopt(void* nccl_comm, int world, int rank, int device, vsf_nccl_comm** out) {
  if (!nccl_comm || !out || world < 1 || rank < 0 || rank >= world) return VSF_ERR_BAD_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return VSF_ERR_CUDA;
  vsf_nccl_comm* c = new vsf_nccl_comm();
  c->comm = static_cast<ncclComm_t>(nccl_comm);
  c->world = world;
  c->rank = rank;
  c->device = device;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return VSF_ERR_CUDA;
  }
  *out = c;
  return VSF_OK;
}

extern "C" void vsf_nccl_comm_destroy(vsf_nccl_comm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) {
    cudaStreamSynchronize(c->stream);
    cudaStreamDestroy(c->stream);
  }
  if (c->owned && c->comm) ncclCommDestroy(c->comm);
  delete c;
}

extern "C" const char* vsf_nccl_last_error(const vsf_nccl_comm* c) { return c ? c->err.c_str() : "null comm"; }

extern "C" int vsf_gather_matches(vsf_ctx* ctx, vsf_nccl_comm* c, int n_frames, int* d_counts, vsf_dmatch* d_lists) {
  if (!ctx || !c || n_frames < 1 || !d_counts || !d_lists) return VSF_ERR_BAD_ARG;
  const vsf_dmatch* src = nullptr;
  const int* src_counts = nullptr;
  int stride = 0, regions = 0;
  int rc = vsf_device_match_lists(ctx, &src, &src_counts, &stride, &regions);
  if (rc) return rc;
  if (n_frames > regions) return VSF_ERR_BAD_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(vsf_stream(ctx));
  VN_CUDA(c, cudaSetDevice(c->device));
  VN_NCCL(c, ncclGroupStart());
  VN_NCCL(c, ncclAllGather(src_counts, d_counts, size_t(n_frames), ncclInt32, c->comm, stream));
  VN_NCCL(c, ncclAllGather(src, d_lists, size_t(n_frames) * stride * sizeof(vsf_dmatch), ncclUint8, c->comm, stream));
  VN_NCCL(c, ncclGroupEnd());
  return VSF_OK;
}

extern "C" int vsf_nccl_gather_bytes(vsf_nccl_comm* c, const void* data, size_t n, int root, void** out, size_t* sizes) {
  if (!c || (n > 0 && !data) || root < 0 || root >= c->world || !out || !sizes) return VSF_ERR_BAD_ARG;
  *out = nullptr;
  VN_CUDA(c, cudaSetDevice(c->device));
  // sizes first (all ranks learn them), then one send / world receives
  unsigned long long* d_sz = nullptr;
  VN_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&d_sz), (c->world + 1) * sizeof(unsigned long long)));
  const unsigned long long mine = n;
  VN_CUDA(c, cudaMemcpyAsync(d_sz + c->world, &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));
  VN_NCCL(c, ncclAllGather(d_sz + c->world, d_sz, 1, ncclUint64, c->comm, c->stream));
  std::vector<unsigned long long> all(c->world);
  VN_CUDA(c, cudaMemcpyAsync(all.data(), d_sz, c->world * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  VN_CUDA(c, cudaStreamSynchronize(c->stream));
  size_t total = 0;
  for (int r = 0; r < c->world; ++r) {
    sizes[r] = size_t(all[r]);
    total += sizes[r];
  }
  uint8_t *d_send = nullptr, *d_recv = nullptr;
  if (n > 0) {
    VN_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&d_send), n));
    VN_CUDA(c, cudaMemcpyAsync(d_send, data, n, cudaMemcpyHostToDevice, c->stream));
  }
  if (c->rank == root && total > 0) VN_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&d_recv), total));
  VN_NCCL(c, ncclGroupStart());
  if (n > 0) VN_NCCL(c, ncclSend(d_send, n, ncclUint8, root, c->comm, c->stream));
  if (c->rank == root) {
    size_t off = 0;
    for (int r = 0; r < c->world; ++r) {
      if (sizes[r] > 0) VN_NCCL(c, ncclRecv(d_recv + off, sizes[r], ncclUint8, r, c->comm, c->stream));
      off += sizes[r];
    }
  }
  VN_NCCL(c, ncclGroupEnd());
  if (c->rank == root) {
    void* host = std::malloc(total > 0 ? total : 1);
    if (!host) return VSF_ERR_CUDA;
    if (total > 0) VN_CUDA(c, cudaMemcpyAsync(host, d_recv, total, cudaMemcpyDeviceToHost, c->stream));
    *out = host;
  }
  VN_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaFree(d_sz);
  if (d_send) cudaFree(d_send);
  if (d_recv) cudaFree(d_recv);
  return VSF_OK;
}
