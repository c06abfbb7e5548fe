// comm.cu -- the data-parallel gradient exchange fused with the optimizer step, over NVLink peer
// memory (one sum of the flat gradient buffer per step, nothing else crosses
// NVLink).  Replaces the pair  ncclAllReduce(grads) ; Adam(params, grads / world)  by ONE kernel:
//
//   every rank owns an exchange region (cudaMalloc'ed here, exported with cudaIpcGetMemHandle and
//   mapped by its peers):   header | flags u32[2][world][kDpCtas] | slots f32[2][world][n_pad]
//
//   CTA b of rank r, exchange number s (parity = s & 1):
//     1. PUSH   its chunk b of the local gradient into slot[parity][r] of every peer (posted NVLink
//               stores, no read round trip), fence, then flag[parity][r][b] = s on every peer;
//     2. WAIT   until its own flag[parity][q][b] == s for every peer q (only chunk b is needed, so
//               there is no grid-wide barrier and no collective launch);
//     3. UPDATE sum the `world` copies of the chunk in RANK ORDER (bit-identical replicas on every
//               rank, unlike a ring), scale by 1/world and apply Adam to chunk b.
//
//   Double buffering by parity makes a second barrier unnecessary: a peer can only push exchange
//   s + 2 (same parity) after it has finished exchange s + 1, which needed this rank's pushes of
//   s + 1, which this rank issues after its own exchange s kernel has completed.
//   The wait is bounded (CAL_DP_TIMEOUT_S, default 600 s, 0 = wait for ever like NCCL): a peer that never
//   delivers makes the kernel record the error word and TRAP, so the rank fails loudly (every later CUDA
//   call returns the launch failure) instead of updating its parameters from a stale slot and silently
//   diverging from its peers.
#include <stdlib.h>
#include <string.h>

#include "internal.cuh"
#include "fsg.cuh"

namespace cal {

namespace {

constexpr int kDpCtas = kSMs;                       // one chunk per SM
constexpr size_t kDpHeaderBytes = 256;

struct DpHeader {
  unsigned int seq;        // exchanges completed by this rank
  unsigned int done;       // arrival counter of the running exchange
  unsigned int error;      // 1 = a peer's flag did not arrive within the timeout
  unsigned int world, n_pad_lo, n_pad_hi;
};

__host__ __device__ inline long long dp_pad(long long n) {
  const long long q = 4ll * kDpCtas;
  return (n + q - 1) / q * q;
}
__host__ __device__ inline size_t dp_flags_bytes(int world) {
  size_t b = (size_t)2 * world * kDpCtas * sizeof(unsigned int);
  return (b + 255) / 256 * 256;
}
__device__ __forceinline__ unsigned int* dp_flag(void* region, int world, int par, int src, int cta) {
  return reinterpret_cast<unsigned int*>(static_cast<char*>(region) + kDpHeaderBytes) +
         ((size_t)par * world + src) * kDpCtas + cta;
}
__device__ __forceinline__ float* dp_slot(void* region, int world, long long n_pad, int par, int src) {
  return reinterpret_cast<float*>(static_cast<char*>(region) + kDpHeaderBytes + dp_flags_bytes(world)) +
         ((size_t)par * world + src) * (size_t)n_pad;
}

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
  return t;
}

struct DpPeers {
  void* region[CAL_MAX_WORLD];
};

__global__ void __launch_bounds__(256) k_dp_adam(const DpPeers peers, const int world, const int rank,
                                                 float* __restrict__ p, const float* g, float* __restrict__ m,
                                                 float* __restrict__ v, const long long n, int* __restrict__ step,
                                                 float lr, const float* __restrict__ lr_dev, const float b1,
                                                 const float b2, const float eps, const float wd,
                                                 const unsigned long long timeout_ns, const cal_image_sink sink) {
  __shared__ float s_c[2];
  __shared__ int s_t;
  __shared__ unsigned int s_seq;
  DpHeader* hdr = static_cast<DpHeader*>(peers.region[rank]);
  const long long n_pad = dp_pad(n);
  const long long chunk = n_pad / kDpCtas;                         // floats, multiple of 4
  const long long c0 = (long long)blockIdx.x * chunk;
  const long long c1 = c0 + chunk < n ? c0 + chunk : n;            // n % 4 == 0 is checked on the host
  pdl_sync();                                                      // the gradients are complete
#ifdef CAL_TIMELINE
  int* dp_stamps = reinterpret_cast<int*>(hdr) + 16;                     // (timeline builds: CTA 0's stamps in the header's spare words)
#define DP_TL(id) do { if (blockIdx.x == 0) CAL_TL(dp_stamps - 96, id); } while (0)
#else
#define DP_TL(id) do { } while (0)
#endif
  DP_TL(0);
  // (hdr->seq is read AFTER the dependency wait: a directly preceding cal_dp_adam_step has then published it)
  if (threadIdx.x == 0) s_seq = *reinterpret_cast<volatile unsigned int*>(&hdr->seq) + 1u;
  if (threadIdx.x == 0) {       // bias corrections in fp64 like the Python scalars of torch.optim.Adam
    const int t = *reinterpret_cast<volatile int*>(step) + 1;
    s_t = t;
    if (lr_dev != nullptr) lr = *lr_dev;
    const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
    s_c[0] = (float)((double)lr / bc1);
    s_c[1] = (float)sqrt(bc2);
  }
  // This thread's piece of the chunk, fetched before anything waits: the gradient goes to the peers from registers,
  // the parameter and its moments are only needed after the peers' flags have arrived.  (A chunk is n_pad / 148 floats:
  // at most one float4 per thread up to 151 552 parameters; larger models take the strided loops.)
  const bool one = chunk <= 4ll * blockDim.x;                       // uniform
  const long long i0 = c0 + 4ll * threadIdx.x;
  const bool have = one && i0 < c1;
  float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f), p4 = g4, m4 = g4, v4 = g4;
  if (have) {
    g4 = *reinterpret_cast<const float4*>(g + i0);
    p4 = *reinterpret_cast<const float4*>(p + i0);
    m4 = *reinterpret_cast<const float4*>(m + i0);
    v4 = *reinterpret_cast<const float4*>(v + i0);
  }
  __syncthreads();
  const unsigned int seq = s_seq;
  const int par = (int)(seq & 1u);
  // ---- 1. push this CTA's chunk to every peer ----
  for (int d = 1; d < world; ++d) {
    const int q = (rank + d) % world;                              // staggered: ranks do not all hit peer 0 first
    float* dst = dp_slot(peers.region[q], world, n_pad, par, rank);
    if (one) {
      if (have) *reinterpret_cast<float4*>(dst + i0) = g4;
    } else {
      for (long long i = i0; i < c1; i += 4ll * blockDim.x)
        *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(g + i);
    }
  }
  DP_TL(1);
  // One release per flag: the barrier orders every thread's pushes before the flag writers' st.release.sys, whose
  // fence is cumulative over them (the pattern of a grid barrier: bar.sync, then one thread fences and signals).  An
  // extra __threadfence_system() by all 256 threads in front of the barrier cost 6.6 us here (timeline build).
  __syncthreads();
  DP_TL(2);
  if ((int)threadIdx.x < world && (int)threadIdx.x != rank)
    st_release_sys(dp_flag(peers.region[threadIdx.x], world, par, rank, blockIdx.x), seq);
  DP_TL(3);
  // ---- 2. wait for every peer's chunk ----
  if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
    const unsigned int* f = dp_flag(peers.region[rank], world, par, threadIdx.x, blockIdx.x);
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    while ((int)(ld_acquire_sys(f) - seq) < 0) {
      if ((++spins & 1023u) == 0u) {
        const unsigned long long now = global_ns();
        if (t0 == 0) t0 = now;
        else if (timeout_ns != 0ull && now - t0 > timeout_ns) {
          *reinterpret_cast<volatile unsigned int*>(&hdr->error) = 1u;
          __threadfence_system();
          __trap();                                                  // never apply an update from a stale slot
        }
      }
    }
  }
  __syncthreads();
  DP_TL(4);
  // ---- 3. rank-ordered sum + Adam on the chunk ----
  const float step_size = s_c[0], bc2s = s_c[1];
  const float gscale = 1.0f / (float)world;
  const float* mine = dp_slot(peers.region[rank], world, n_pad, par, 0);
  for (long long i = i0; i < c1; i += 4ll * blockDim.x) {
    if (!one) {
      g4 = *reinterpret_cast<const float4*>(g + i);
      p4 = *reinterpret_cast<const float4*>(p + i);
      m4 = *reinterpret_cast<const float4*>(m + i);
      v4 = *reinterpret_cast<const float4*>(v + i);
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < world; ++q) {
      const float4 t = q == rank ? g4 : __ldcg(reinterpret_cast<const float4*>(mine + (size_t)q * n_pad + i));
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    float gi[4] = {acc.x * gscale, acc.y * gscale, acc.z * gscale, acc.w * gscale};
    float pp[4] = {p4.x, p4.y, p4.z, p4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gk = gi[k];
      if (wd != 0.f) gk = fmaf(wd, pp[k], gk);
      mm[k] = mm[k] + (gk - mm[k]) * (1.f - b1);                   // torch: exp_avg.lerp_(grad, 1 - beta1)
      vv[k] = vv[k] * b2 + (1.f - b2) * gk * gk;
      const float denom = sqrtf(vv[k]) / bc2s + eps;
      pp[k] = pp[k] - step_size * (mm[k] / denom);
    }
    *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    *reinterpret_cast<float4*>(p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
    DP_TL(6);
    if (sink.count > 0) fsg_sink_emit4(sink, i, make_float4(pp[0], pp[1], pp[2], pp[3]));   // operand images of the fused small-graph path
  }
  DP_TL(5);
  // every CTA has read hdr->seq and *step before it arrives here, so the last one may advance them
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&hdr->done, 1u) == gridDim.x - 1u) {
      hdr->done = 0u;
      hdr->seq = seq;
      *step = s_t;
    }
  }
}

}  // namespace
}  // namespace cal

extern "C" size_t cal_dp_region_bytes(int32_t world, int64_t n) {
  if (world < 1 || world > CAL_MAX_WORLD || n <= 0) return 0;
  return cal::kDpHeaderBytes + cal::dp_flags_bytes(world) + (size_t)2 * world * (size_t)cal::dp_pad(n) * sizeof(float);
}

extern "C" int cal_dp_alloc(int32_t device, size_t bytes, void** region) {
  if (!region) return CAL_ENULL;
  if (bytes < cal::kDpHeaderBytes) return CAL_EINVAL;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return (int)e;
  void* p = nullptr;
  e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(p);
    return (int)e;
  }
  *region = p;
  return 0;
}

extern "C" int cal_dp_free(void* region) {
  if (!region) return CAL_ENULL;
  return (int)cudaFree(region);
}

extern "C" int cal_dp_export(void* region, unsigned char handle[CAL_DP_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == CAL_DP_HANDLE_BYTES, "handle size");
  if (!region || !handle) return CAL_ENULL;
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, region);
  if (e != cudaSuccess) return (int)e;
  memcpy(handle, &h, sizeof(h));
  return 0;
}

extern "C" int cal_dp_import(int32_t device, const unsigned char handle[CAL_DP_HANDLE_BYTES], void** mapped) {
  if (!handle || !mapped) return CAL_ENULL;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return (int)e;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return (int)e;
  *mapped = p;
  return 0;
}

extern "C" int cal_dp_unmap(void* mapped) {
  if (!mapped) return CAL_ENULL;
  return (int)cudaIpcCloseMemHandle(mapped);
}

extern "C" int cal_dp_read_error(const cal_dp_comm* comm, void* stream) {
  if (!comm) return CAL_ENULL;
  if (comm->world < 1 || comm->world > CAL_MAX_WORLD || comm->rank < 0 || comm->rank >= comm->world) return CAL_EINVAL;
  cal::DpHeader h;
  cudaError_t e = cudaMemcpyAsync(&h, comm->region[comm->rank], sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  return h.error ? CAL_ETIMEOUT : 0;
}

extern "C" int cal_dp_adam_step_images(const cal_dp_comm* comm, float* params, const float* grads, float* exp_avg,
                                       float* exp_avg_sq, int64_t n, int32_t* step, float lr, const float* lr_device,
                                       float beta1, float beta2, float eps, float weight_decay,
                                       const cal_image_sink* sink, void* stream) {
  cal_image_sink sk = {};
  if (sink != nullptr) sk = *sink;
  if (sk.count < 0 || sk.count > 16) return CAL_EINVAL;
  if (!comm || !params || !grads || !exp_avg || !exp_avg_sq || !step) return CAL_ENULL;
  if (comm->world < 1 || comm->world > CAL_MAX_WORLD || comm->rank < 0 || comm->rank >= comm->world) return CAL_EINVAL;
  if (n <= 0 || (n & 3) != 0) return CAL_EINVAL;
  if (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) return CAL_EALIGN;
  cal::DpPeers peers;
  for (int q = 0; q < CAL_MAX_WORLD; ++q) {
    peers.region[q] = q < comm->world ? comm->region[q] : nullptr;
    if (q < comm->world && peers.region[q] == nullptr) return CAL_ENULL;
  }
  static const unsigned long long timeout_ns = [] {
    const char* e = getenv("CAL_DP_TIMEOUT_S");
    const double sec = e != nullptr ? atof(e) : 600.0;
    return sec > 0.0 ? (unsigned long long)(sec * 1e9) : 0ull;
  }();
  cal::launch_k(cal::k_dp_adam, dim3(cal::kDpCtas), dim3(256), 0, (cudaStream_t)stream, peers, (int)comm->world,
                (int)comm->rank, params, grads, exp_avg, exp_avg_sq, (long long)n, step, lr, lr_device, beta1, beta2,
                eps, weight_decay, timeout_ns, sk);
  cal::note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int cal_dp_adam_step(const cal_dp_comm* comm, float* params, const float* grads, float* exp_avg,
                                float* exp_avg_sq, int64_t n, int32_t* step, float lr, const float* lr_device,
                                float beta1, float beta2, float eps, float weight_decay, void* stream) {
  return cal_dp_adam_step_images(comm, params, grads, exp_avg, exp_avg_sq, n, step, lr, lr_device, beta1, beta2, eps,
                                 weight_decay, nullptr, stream);
}
