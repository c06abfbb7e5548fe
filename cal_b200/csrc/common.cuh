// common.cuh -- shared device helpers for the cal_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cal_b200.h"

namespace cal {

constexpr int kWarp = 32;
constexpr int kRowThreads = 256;               // row-per-warp kernels: 8 warps per CTA
constexpr int kRowWarps = kRowThreads / kWarp;
constexpr int kMaxH = 256;
constexpr int kSMs = 148;                      // B200: 148 SMs (2 dies x 74)
constexpr int kMaxStatBlocks = 2 * kSMs;       // cap on CTAs that emit BN partials

// status bits (CAL_WS_STATUS[0])
constexpr int kStBadNode = 1;                  // edge endpoint outside [0, N)
constexpr int kStBadBatch = 2;                 // batch not sorted / outside [0, B)
constexpr int kStCapacity = 4;                 // N/E/B exceeds the workspace capacity

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// A lane's contiguous slice of one [*, H] row: H = 32 * VEC, lane owns channels
// [lane*VEC, lane*VEC + VEC).  VEC in {1, 2, 4, 8}; loads/stores are 4/8/16-byte vectors.
template <int VEC>
struct RowVec {
  float v[VEC];
  __device__ __forceinline__ void load(const float* __restrict__ row, int lane) {
    const float* p = row + lane * VEC;
    if constexpr (VEC == 1) {
      v[0] = __ldg(p);
    } else if constexpr (VEC == 2) {
      float2 t = __ldg(reinterpret_cast<const float2*>(p));
      v[0] = t.x; v[1] = t.y;
    } else {
#pragma unroll
      for (int i = 0; i < VEC / 4; ++i) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      }
    }
  }
  // plain (coherent) load: for buffers written earlier in the same kernel chain is fine with
  // __ldg too, but buffers written by THIS kernel must use this one.
  __device__ __forceinline__ void load_coherent(const float* row, int lane) {
    const float* p = row + lane * VEC;
    if constexpr (VEC == 1) {
      v[0] = *p;
    } else if constexpr (VEC == 2) {
      float2 t = *reinterpret_cast<const float2*>(p);
      v[0] = t.x; v[1] = t.y;
    } else {
#pragma unroll
      for (int i = 0; i < VEC / 4; ++i) {
        float4 t = *(reinterpret_cast<const float4*>(p) + i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      }
    }
  }
  __device__ __forceinline__ void store(float* __restrict__ row, int lane) const {
    float* p = row + lane * VEC;
    if constexpr (VEC == 1) {
      *p = v[0];
    } else if constexpr (VEC == 2) {
      *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    } else {
#pragma unroll
      for (int i = 0; i < VEC / 4; ++i)
        *(reinterpret_cast<float4*>(p) + i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
  }
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = 0.f;
  }
};

template <int VEC>
__device__ __forceinline__ float dot_lane(const RowVec<VEC>& a, const RowVec<VEC>& b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) s = fmaf(a.v[i], b.v[i], s);
  return s;
}

// Grid-wide "am I the last CTA" test on a self-resetting counter.  All threads must call.
// `expected` = number of CTAs that arrive on this counter.
__device__ __forceinline__ bool grid_last_block(unsigned int* counter, unsigned int expected) {
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {             // one thread fences on both sides of the atomic (cooperative-groups pattern)
    __threadfence();
    unsigned int t = atomicAdd(counter, 1u);
    s_last = (t == expected - 1u);
    if (s_last) {
      *counter = 0u;                  // ready for the next launch (graph replay safe)
      __threadfence();
    }
  }
  __syncthreads();
  return s_last != 0;
}

// Row-per-warp kernels: reduce NV per-lane double vectors (lane owns VEC channels) over the
// CTA's warps and write them to partial[(part * NVT + v0 + v) * K + koff + k], k < H.  `sbuf` holds
// kRowWarps * H doubles.  Summation order is fixed (warp 0..7) => deterministic.
template <int VEC, int NV>
__device__ __forceinline__ void block_partial_store_ex(double (&acc)[NV][VEC], double* sbuf, double* partial,
                                                       int H, int part, int NVT, int v0, int K, int koff) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VEC; ++i) sbuf[warp * H + lane * VEC + i] = acc[v][i];
    __syncthreads();
    for (int k = threadIdx.x; k < H; k += blockDim.x) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kRowWarps; ++w) s += sbuf[w * H + k];
      partial[((size_t)part * NVT + v0 + v) * K + koff + k] = s;
    }
  }
}
template <int VEC, int NV>
__device__ __forceinline__ void block_partial_store(double (&acc)[NV][VEC], double* sbuf, double* partial,
                                                    int H) {
  block_partial_store_ex<VEC, NV>(acc, sbuf, partial, H, blockIdx.x, NV, 0, H, 0);
}

// Sum over the G CTAs' partials of vector v, channel k (fixed order).
__device__ __forceinline__ double partial_total(const double* partial, int G, int NV, int v, int H, int k) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int g = 0;
  for (; g + 4 <= G; g += 4) {
    s0 += partial[((size_t)(g + 0) * NV + v) * H + k];
    s1 += partial[((size_t)(g + 1) * NV + v) * H + k];
    s2 += partial[((size_t)(g + 2) * NV + v) * H + k];
    s3 += partial[((size_t)(g + 3) * NV + v) * H + k];
  }
  for (; g < G; ++g) s0 += partial[((size_t)g * NV + v) * H + k];
  return (s0 + s1) + (s2 + s3);
}

// ---- programmatic dependent launch (PDL) ----
// Every kernel is launched with programmatic stream serialization: it may be scheduled as soon as
// its predecessor has passed its own dependency wait, so launch latency and the independent part
// of its prologue (weight staging) overlap the predecessor's tail.  Protocol used by all kernels:
//   [prologue that reads nothing written by the immediate predecessor]
//   pdl_wait();      // predecessor grid complete and flushed (hence, transitively, all earlier ones)
//   pdl_trigger();   // successor may be scheduled now
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
  pdl_wait();
  pdl_trigger();
}

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface in cudaGetLastError()
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int row_grid(int max_rows, int rows_per_cta = 4 * kRowWarps) {
  int g = ceil_div(max_rows > 0 ? max_rows : 1, rows_per_cta);
  if (g > kMaxStatBlocks) g = kMaxStatBlocks;
  if (g < 1) g = 1;
  return g;
}

#define CAL_CUDA_CHECK_LAUNCH()                      \
  do {                                               \
    cudaError_t e__ = cudaGetLastError();            \
    if (e__ != cudaSuccess) return (int)e__;         \
  } while (0)

#define CAL_DISPATCH_VEC(H, ...)                              \
  switch ((H) / 32) {                                         \
    case 1: { constexpr int VEC = 1; __VA_ARGS__; } break;    \
    case 2: { constexpr int VEC = 2; __VA_ARGS__; } break;    \
    case 4: { constexpr int VEC = 4; __VA_ARGS__; } break;    \
    default: return CAL_EINVAL;                               \
  }

}  // namespace cal
