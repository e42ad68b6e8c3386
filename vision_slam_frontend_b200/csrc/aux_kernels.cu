// Synthetic sequence generator (bench / test frame source) and the integer-pipe
// probe that gives the roofline its measured denominator.
#include "vsf_device.cuh"
#include "stereo_args.cuh"

namespace vsf {

// splitmix64 finaliser: counter-based, so any (pose, feature, word) can be
// generated independently on any rank.  tests/synth_twin.py restates it in numpy.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__host__ __device__ __forceinline__ uint32_t gcd_u32(uint32_t a, uint32_t b) {
  while (b) {
    const uint32_t t = a % b;
    a = b;
    b = t;
  }
  return a;
}

// Pose p observes landmarks [stride*p, stride*p + n); feature i of pose p is
// landmark stride*p + (a_p*i + b_p) mod n (affine permutation, a_p coprime to n).
// Landmark L has the code word w = low32(mix64(seed ^ (L*words + w))), words = 8 (256-bit rows)
// or 16 (64-byte rows; bytes at and beyond desc_bytes are zero, the device layout of 61-byte
// AKAZE rows).  Every bit is flipped with probability 1/32 (AND of five uniform words).
// One thread per (pose, feature, word).
__global__ void synth_sequence_kernel(uint32_t* out, int n, int first_pose, int n_poses,
                                      int stride, uint64_t seed, int words, int desc_bytes) {
  const size_t gid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = size_t(n_poses) * n * words;
  if (gid >= total) return;
  const int w = int(gid % unsigned(words));
  const size_t fi = gid / unsigned(words);
  const int i = int(fi % n);
  const int p = first_pose + int(fi / n);
  const uint64_t hp = mix64(seed ^ (0xA5A5A5A5ull + uint64_t(p) * 0x100000001B3ull));
  uint32_t a = uint32_t(hp % uint32_t(n)) | 1u;
  while (gcd_u32(a, uint32_t(n)) != 1u) a += 2u;
  const uint32_t b = uint32_t((hp >> 32) % uint32_t(n));
  const uint64_t perm = (uint64_t(a) * uint64_t(i) + b) % uint64_t(n);
  const uint64_t L = uint64_t(stride) * uint64_t(p) + perm;
  const uint32_t code = uint32_t(mix64(seed ^ (L * uint64_t(words) + w)));
  const uint64_t c = (uint64_t(p) * uint64_t(n) + uint64_t(i)) * uint64_t(words) + w;
  const uint64_t r0 = mix64(~seed ^ (c * 3 + 0));
  const uint64_t r1 = mix64(~seed ^ (c * 3 + 1));
  const uint64_t r2 = mix64(~seed ^ (c * 3 + 2));
  const uint32_t flip = uint32_t(r0) & uint32_t(r0 >> 32) & uint32_t(r1) & uint32_t(r1 >> 32) &
                        uint32_t(r2);
  const int valid = desc_bytes - 4 * w;            // bytes of this word inside the descriptor
  const uint32_t mask = valid >= 4 ? 0xFFFFFFFFu : (valid <= 0 ? 0u : ((1u << (8 * valid)) - 1u));
  out[gid] = (code ^ flip) & mask;
}

cudaError_t launch_synth(uint32_t* out, int n, int first_pose, int n_poses, int stride,
                         uint64_t seed, int words, int desc_bytes, cudaStream_t stream) {
  const size_t total = size_t(n_poses) * n * words;
  if (total == 0) return cudaSuccess;
  const int threads = 256;
  const size_t blocks = (total + threads - 1) / threads;
  synth_sequence_kernel<<<unsigned(blocks), threads, 0, stream>>>(out, n, first_pose, n_poses,
                                                                   stride, seed, words, desc_bytes);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Pipe probe: every thread runs 8 independent dependent-chains of one instruction
// kind.  Lane-ops/s = grid threads * 8 * iters * (ops per chain step) / time.
template <int KIND>
__global__ void __launch_bounds__(1024) probe_kernel(uint32_t* sink, int iters, uint32_t seed) {
  uint32_t x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u;
  const uint32_t a = seed | 1u, b = seed ^ 0x55555555u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0) {
        asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
      } else if (KIND == 1) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
      } else if (KIND == 2) {
        asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
      } else if (KIND == 3) {
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
      } else if (KIND == 4) {
        asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a + it));
      } else if (KIND == 5) {
        asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(x[i]) : "r"(b), "r"(a));
      } else if (KIND == 6) {
        asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
        asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= x[i];
  if (acc == 0x12345679u) sink[0] = acc;  // keep the chains alive
}

// ops issued per chain step for each kind
int probe_ops_per_step(int kind) {
  switch (kind) {
    case 2: return 2;
    case 5: return 3;
    case 6: return 2;
    default: return 1;
  }
}

cudaError_t launch_probe(int kind, uint32_t* sink, int iters, int blocks, int threads,
                         cudaStream_t stream) {
  switch (kind) {
    case 0: probe_kernel<0><<<blocks, threads, 0, stream>>>(sink, iters, 12345u); break;
    case 1: probe_kernel<1><<<blocks, threads, 0, stream>>>(sink, iters, 12345u); break;
    case 2: probe_kernel<2><<<blocks, threads, 0, stream>>>(sink, iters, 12345u); break;
    case 3: probe_kernel<3><<<blocks, threads, 0, stream>>>(sink, iters, 12345u); break;
    case 4: probe_kernel<4><<<blocks, threads, 0, stream>>>(sink, iters, 12345u); break;
    case 5: probe_kernel<5><<<blocks, threads, 0, stream>>>(sink, iters, 12345u); break;
    case 6: probe_kernel<6><<<blocks, threads, 0, stream>>>(sink, iters, 12345u); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// N1: cv::undistortPoints (stereo_args.cuh: undistort_one).  One thread per point; O(n) and
// trivially parallel, here so that the pixels of a node's features can stay on the device (the
// fused frame path applies the same function inside its triangulation launch).
__global__ void undistort_points_kernel(const float2* __restrict__ in, int n, const UndistortArgs a,
                                        float2* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = undistort_one(in[i], a);
}

cudaError_t launch_undistort_points(const float2* in, int n, const float* K9, const float* dist5, float2* out,
                                    cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  undistort_points_kernel<<<(n + 255) / 256, 256, 0, stream>>>(in, n, make_undistort_args(K9, dist5), out);
  return cudaGetLastError();
}

}  // namespace vsf
