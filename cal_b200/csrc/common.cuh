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
// status[0]: the bits of the batch prepared last (cal_prep rewrites it);  status[kStSticky]: every bit raised since the
// last cal_read_status (which clears it) -- a batch prepared ahead of its step must not hide the bits of the step before
constexpr int kStSticky = 4;

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

// -DCAL_TIMELINE builds: block 0 / thread 0 of the step's kernels stamp %globaltimer (ns, low 30 bits) into
// status[96 + id] -- the live schedule of one graph-replayed step (tools/timeline.py); compiled out otherwise.
#ifdef CAL_TIMELINE
#define CAL_TL(statusp, id)                                                          \
  do {                                                                               \
    if (threadIdx.x == 0) {                                                          \
      unsigned long long tl_;                                                        \
      asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(tl_));                      \
      (statusp)[96 + (id)] = (int)(tl_ & 0x3fffffffull);                             \
    }                                                                                \
  } while (0)
// per-CTA stamps of the two fused kernels: int[(kern * 160 + block) * 8 + slot] in the (idle) CAL_WS_D region
#define CAL_TLC(c, kern, slot)                                                       \
  do {                                                                               \
    if (threadIdx.x == 0) {                                                          \
      unsigned long long tl_;                                                        \
      asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(tl_));                      \
      reinterpret_cast<int*>((c).D)[((kern) * 160 + blockIdx.x) * 8 + (slot)] = (int)(tl_ & 0x3fffffffull); \
    }                                                                                \
  } while (0)
#else
#define CAL_TL(statusp, id) do { } while (0)
#define CAL_TLC(c, kern, slot) do { } while (0)
#endif

__device__ __forceinline__ void raise_status(int* status, int bits) {
  atomicOr(status, bits);
  atomicOr(status + kStSticky, bits);
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

// ... without the programmatic edge: the kernel starts after everything before it in the stream has completed (its
// own griddepcontrol.wait is then a no-op)
template <typename... KArgs, typename... Args>
inline void launch_k_plain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
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
