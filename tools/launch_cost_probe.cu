// Host cost of one kernel launch as a function of the size of its parameter block and of the
// programmatic-launch attribute (tools only; nvcc -O2 -arch=sm_100a -o launch_cost_probe ...).
// For every size: (a) host time per launch, back to back (64 launches, then a sync);
// (b) launch -> host sees the kernel's store into mapped memory, one launch at a time (idle GPU).
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <vector>

template <int BYTES>
struct Blob { unsigned char b[BYTES]; };

template <int BYTES>
__global__ void probe_kernel(const __grid_constant__ Blob<BYTES> blob, volatile int* flag, int value) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *flag = value + blob.b[BYTES - 1];
}

static double now_us() {
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <int BYTES>
static void run(cudaStream_t st, volatile int* h_flag, int* d_flag, bool pdl) {
  Blob<BYTES> blob;
  std::memset(&blob, 0, sizeof(blob));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(64);
  cfg.blockDim = dim3(128);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  int value = 1;
  for (int i = 0; i < 200; ++i) cudaLaunchKernelEx(&cfg, probe_kernel<BYTES>, blob, (volatile int*)d_flag, value++);
  cudaStreamSynchronize(st);
  std::vector<double> per;
  for (int rep = 0; rep < 50; ++rep) {
    const double t0 = now_us();
    for (int i = 0; i < 64; ++i) cudaLaunchKernelEx(&cfg, probe_kernel<BYTES>, blob, (volatile int*)d_flag, value++);
    const double t1 = now_us();
    cudaStreamSynchronize(st);
    per.push_back((t1 - t0) / 64.0);
  }
  std::sort(per.begin(), per.end());
  std::vector<double> lat, host;
  for (int rep = 0; rep < 200; ++rep) {
    *h_flag = 0;
    const double t0 = now_us();
    cudaLaunchKernelEx(&cfg, probe_kernel<BYTES>, blob, (volatile int*)d_flag, 7);
    const double t1 = now_us();
    while (*h_flag != 7) {}
    const double t2 = now_us();
    cudaStreamSynchronize(st);
    host.push_back(t1 - t0);
    lat.push_back(t2 - t0);
  }
  std::sort(lat.begin(), lat.end());
  std::sort(host.begin(), host.end());
  std::printf("{\"param_bytes\": %d, \"pdl\": %d, \"back_to_back_us_per_launch_median\": %.2f, \"idle_launch_host_us_median\": %.2f, "
              "\"idle_launch_to_visible_us_median\": %.2f, \"idle_launch_to_visible_us_best\": %.2f}\n",
              BYTES + 16, pdl ? 1 : 0, per[per.size() / 2], host[host.size() / 2], lat[lat.size() / 2], lat[0]);
}

int main() {
  cudaSetDevice(0);
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  int* h_flag = nullptr;
  int* d_flag = nullptr;
  cudaHostAlloc(reinterpret_cast<void**>(&h_flag), 64, cudaHostAllocMapped);
  cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_flag), h_flag, 0);
  for (int pdl = 0; pdl < 2; ++pdl) {
    run<48>(st, h_flag, d_flag, pdl);
    run<496>(st, h_flag, d_flag, pdl);
    run<1008>(st, h_flag, d_flag, pdl);
    run<2032>(st, h_flag, d_flag, pdl);
    run<3312>(st, h_flag, d_flag, pdl);
    run<4064>(st, h_flag, d_flag, pdl);
  }
  return 0;
}
