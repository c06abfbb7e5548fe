// optim.cu -- fixed-order reduction of the per-CTA partial parameter gradients into the flat
// gradient buffer, and the fused Adam step on flat buffers (train_causal.py:21,192).
#include "internal.cuh"
#include "fsg.cuh"

namespace cal {

namespace {

constexpr int kMaxRed = 48;

// nparts rule: how many CTAs of the producing kernel actually wrote a partial
enum { PARTS_NODE_TILES = 0, PARTS_NODE_ROWS, PARTS_HEAD_TILES, PARTS_HEAD_ROWS };

struct RedEntry {
  long long dst;      // offset into grads
  size_t src;         // offset into gpart of part 0
  int n;              // elements
  int stride;         // floats between consecutive parts
  int rule, gmax;
  int split;          // lanes that share one element (1, 8 or 32): each sums every split-th partial
  int items;          // n * split rounded up to a multiple of 32 (work items of this entry)
};
struct RedTable {
  int count;
  long long total;    // sum of n
  RedEntry e[kMaxRed];
  // feat special
  size_t feat_src;
  int feat_stride, feat_gmax;
};

__device__ __forceinline__ int nparts_of(const Ctx& c, int rule, int gmax) {
  const int N = imin(imax(c.dims[0], 0), c.Nm), B = imin(imax(c.dims[2], 0), c.Bm);
  int n;
  switch (rule) {
    case PARTS_NODE_TILES: n = ceil_div(N, kTileRows); break;
    case PARTS_NODE_ROWS: n = ceil_div(N, kRowWarps); break;
    case PARTS_HEAD_TILES: n = ceil_div(B, kTileRows); break;
    default: n = ceil_div(B, kHeadRowsPerCta); break;
  }
  return imin(n, gmax);
}

__global__ void __launch_bounds__(256) k_grad_reduce(const Ctx c, const RedTable t, const int n_generic_blocks) {
  pdl_sync();
  if ((int)blockIdx.x < n_generic_blocks) {
    // generic entries: `split` adjacent lanes per output element (1 for the big conv matrices whose
    // partials are read fully coalesced; 32 for the small attention vectors that have hundreds of
    // partials), combined with a fixed xor-shuffle tree => deterministic
    for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < t.total; i0 += (long long)n_generic_blocks * blockDim.x) {
      const long long i = i0 + threadIdx.x;
      long long r = i;
      int ei = 0;
      while (ei < t.count - 1 && r >= t.e[ei].items) {
        r -= t.e[ei].items;
        ++ei;
      }
      const RedEntry& e = t.e[ei];
      const int sp = e.split;
      const long long el = r / sp;
      const int sub = (int)(r % sp);
      const bool live = i < t.total && el < e.n;
      const int np = nparts_of(c, e.rule, e.gmax);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      if (live) {
        const float* p = c.gpart + e.src + el;
        int g = sub;
#pragma unroll 4                                          // 16 independent loads in flight per thread (latency-bound otherwise)
        for (; g + 3 * sp < np; g += 4 * sp) {
          s0 += p[(size_t)g * e.stride];
          s1 += p[(size_t)(g + sp) * e.stride];
          s2 += p[(size_t)(g + 2 * sp) * e.stride];
          s3 += p[(size_t)(g + 3 * sp) * e.stride];
        }
        for (; g < np; g += sp) s0 += p[(size_t)g * e.stride];
      }
      float sum = (s0 + s1) + (s2 + s3);
      for (int o = sp >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);   // warp-uniform sp
      if (live && sub == 0) c.grads[e.dst + el] = sum;
    }
    return;
  }
  // input transform: one CTA per feature row f; thread = (column j, slice of the partials)
  //   d W_feat[f][j] = gamma0[f] M[f][j] + beta0[f] cs[j];  d gamma0[f] = sum_j W[f][j] M[f][j];
  //   d beta0[f] = sum_j W[f][j] cs[j];  conv_feat.bias never receives a gradient (gfn=True).
  __shared__ float s_m[256], s_c[256], s_g[8], s_b[8];
  const int H = c.H, F = c.F;
  const int tid = threadIdx.x;
  const int np = nparts_of(c, PARTS_NODE_TILES, t.feat_gmax);
  const float* W = c.params + c.po.conv_feat_w;
  const int nsl = 256 / H, j = tid % H, sl = tid / H;
  for (int f = (int)blockIdx.x - n_generic_blocks; f < F; f += (int)gridDim.x - n_generic_blocks) {
    float m0 = 0.f, m1 = 0.f, c0 = 0.f, c1 = 0.f;
    const float* pm = c.gpart + t.feat_src + (size_t)f * H + j;
    const float* pc = c.gpart + t.feat_src + (size_t)F * H + j;
    int g = sl;
    for (; g + 3 * nsl < np; g += 4 * nsl) {           // 8 independent loads in flight
      const float a0 = pm[(size_t)g * t.feat_stride], a1 = pm[(size_t)(g + nsl) * t.feat_stride];
      const float a2 = pm[(size_t)(g + 2 * nsl) * t.feat_stride], a3 = pm[(size_t)(g + 3 * nsl) * t.feat_stride];
      const float b0 = pc[(size_t)g * t.feat_stride], b1 = pc[(size_t)(g + nsl) * t.feat_stride];
      const float b2 = pc[(size_t)(g + 2 * nsl) * t.feat_stride], b3 = pc[(size_t)(g + 3 * nsl) * t.feat_stride];
      m0 += a0; m1 += a1; m0 += a2; m1 += a3;
      c0 += b0; c1 += b1; c0 += b2; c1 += b3;
    }
    for (; g < np; g += nsl) {
      m0 += pm[(size_t)g * t.feat_stride];
      c0 += pc[(size_t)g * t.feat_stride];
    }
    __syncthreads();
    s_m[tid] = m0 + m1;
    s_c[tid] = c0 + c1;
    __syncthreads();
    float dg = 0.f, db = 0.f;
    if (sl == 0) {
      float m = 0.f, cs = 0.f;
      for (int q = 0; q < nsl; ++q) {
        m += s_m[q * H + j];
        cs += s_c[q * H + j];
      }
      const float g0 = c.params[c.po.bn_feat_w + f], b0 = c.params[c.po.bn_feat_b + f];
      c.grads[c.po.conv_feat_w + (size_t)f * H + j] = g0 * m + b0 * cs;
      const float w = W[(size_t)f * H + j];
      dg = w * m;
      db = w * cs;
      if (f == 0 && c.po.conv_feat_b >= 0) c.grads[c.po.conv_feat_b + j] = 0.f;
    }
    // block sums of dg / db over the H columns (threads with sl != 0 contribute zeros; fixed order)
    dg = warp_sum(dg);
    db = warp_sum(db);
    if ((tid & 31) == 0) {
      s_g[tid >> 5] = dg;
      s_b[tid >> 5] = db;
    }
    __syncthreads();
    if (tid == 0) {
      float a = 0.f, b = 0.f;
      for (int w = 0; w < 8; ++w) {
        a += s_g[w];
        b += s_b[w];
      }
      c.grads[c.po.bn_feat_w + f] = a;
      c.grads[c.po.bn_feat_b + f] = b;
    }
  }
}

__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, long long n, int* __restrict__ step, unsigned int* __restrict__ done,
                       float lr, const float* __restrict__ lr_dev, float b1, float b2, float eps, float wd,
                       float gscale, const cal_image_sink sink) {
  pdl_sync();
  __shared__ float s_c[2];
  __shared__ int s_t;
  if (threadIdx.x == 0) {       // bias corrections in fp64 like the Python scalars of torch.optim.Adam
    const int t = *reinterpret_cast<volatile int*>(step) + (done != nullptr ? 1 : 0);   // 1-based step of this update
    s_t = t;
    if (lr_dev != nullptr) lr = *lr_dev;
    const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
    s_c[0] = (float)((double)lr / bc1);
    s_c[1] = (float)sqrt(bc2);
  }
  __syncthreads();
  const float step_size = s_c[0], bc2s = s_c[1];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    float mi = m[i] + (gi - m[i]) * (1.f - b1);          // torch: exp_avg.lerp_(grad, 1 - beta1)
    float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2s + eps;
    const float pn = pi - step_size * (mi / denom);
    p[i] = pn;
    if (sink.count > 0) fsg_sink_emit(sink, i, pn);       // operand images of the fused small-graph path
  }
  // fused tick: every CTA has read *step before it arrives here, so the last one may advance it
  if (done != nullptr) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(done, 1u) == gridDim.x - 1u) {
        *done = 0u;
        *step = s_t;
      }
    }
  }
}

__global__ void k_tick(int* step) {
  pdl_sync(); *step += 1; }

}  // namespace

int launch_grad_reduce(const Ctx& c, cudaStream_t s) {
  RedTable t;
  t.count = 0;
  t.total = 0;
  const int H = c.H, C = c.C;
  auto add = [&](long long dst, size_t src, int n, int stride, int rule, int gmax, int split = 1) {
    if (dst < 0 || n <= 0 || t.count >= kMaxRed) return;
    RedEntry& e = t.e[t.count++];
    e.dst = dst; e.src = src; e.n = n; e.stride = stride; e.rule = rule; e.gmax = gmax;
    e.split = split;
    e.items = (n * split + 31) / 32 * 32;      // a warp never straddles two entries
    t.total += e.items;
  };
  const int cs = H * H + H;
  if (c.model == CAL_MODEL_GCN)
    for (int l = 0; l < c.L; ++l) {
      add(c.po.convs_w[l], c.gp_conv[l], H * H, cs, PARTS_NODE_TILES, c.g_tile);
      add(c.po.convs_b[l], c.gp_conv[l] + H * H, H, cs, PARTS_NODE_TILES, c.g_tile);
    }
  else if (c.model == CAL_MODEL_GIN)
    for (int l = 0; l < c.L; ++l) {
      add(c.po.convs_w[l], c.gp_conv[l], H * H, cs, PARTS_NODE_TILES, c.g_tile);
      add(c.po.convs_b[l], c.gp_conv[l] + H * H, H, cs, PARTS_NODE_TILES, c.g_tile);
      add(c.po.gin_w2[l], c.gp_gin2[l], H * H, cs, PARTS_NODE_TILES, c.g_tile);
      add(c.po.gin_b2[l], c.gp_gin2[l] + H * H, H, cs, PARTS_NODE_TILES, c.g_tile);
    }
  else
    for (int l = 0; l < c.L; ++l) {
      add(c.po.convs_w[l], c.gp_conv[l], H * H, cs, PARTS_NODE_TILES, c.g_tile);
      add(c.po.convs_b[l], c.gp_conv[l] + H * H, H, cs, PARTS_NODE_TILES, c.g_tile);
      add(c.po.convs_att[l], c.gp_gat[l], 2 * H, 2 * H, PARTS_NODE_TILES, c.g_tile, 8);
    }
  add(c.po.context_w, c.gp_conv[c.L], H * H, cs, PARTS_NODE_TILES, c.g_tile);
  add(c.po.context_b, c.gp_conv[c.L] + H * H, H, cs, PARTS_NODE_TILES, c.g_tile);
  add(c.po.objects_w, c.gp_conv[c.L + 1], H * H, cs, PARTS_NODE_TILES, c.g_tile);
  add(c.po.objects_b, c.gp_conv[c.L + 1] + H * H, H, cs, PARTS_NODE_TILES, c.g_tile);
  // the attention vectors: ~ceil(N / 8) partials (one per k_att_bwd CTA) of 6H + 4 floats -> 32 lanes per element
  const int as = 8 * H + 4;
  add(c.po.node_att_w, c.gp_att, 2 * H, as, PARTS_NODE_ROWS, c.g_row, 32);
  add(c.po.edge_att_w, c.gp_att + 2 * H, 4 * H, as, PARTS_NODE_ROWS, c.g_row, 32);
  add(c.po.node_att_b, c.gp_att + 6 * H, 2, as, PARTS_NODE_ROWS, c.g_row, 32);
  add(c.po.edge_att_b, c.gp_att + 6 * H + 2, 2, as, PARTS_NODE_ROWS, c.g_row, 32);
  // (the readout parameters get their gradients straight from k_readout_bwd)
  (void)C;
  t.feat_src = c.gp_feat;
  t.feat_stride = c.F * H + H;
  t.feat_gmax = c.g_tile;
  const int nb = imax(1, imin((int)((t.total + 255) / 256), 4 * kSMs));
  const int nf = imax(1, imin(c.F, 2 * kSMs));          // one CTA per feature row
  launch_k(k_grad_reduce, dim3(nb + nf), dim3(256), 0, s, c, t, nb);
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace cal

extern "C" int cal_adam_step_images(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                    int32_t* step, float lr, const float* lr_device, float beta1, float beta2,
                                    float eps, float weight_decay, float grad_scale, const cal_image_sink* sink,
                                    void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !step) return CAL_ENULL;
  if (n <= 0) return CAL_EINVAL;
  cal_image_sink sk = {};
  if (sink != nullptr) sk = *sink;
  if (sk.count < 0 || sk.count > 16) return CAL_EINVAL;
  int g = (int)((n + 255) / 256);
  if (g > 4 * cal::kSMs) g = 4 * cal::kSMs;
  // step[0] = number of updates applied so far (advanced by this call); step[1] = arrival counter
  cal::launch_k(cal::k_adam, dim3(g), dim3(256), 0, (cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq, (long long)n, step,
                                                  reinterpret_cast<unsigned int*>(step + 1), lr, lr_device, beta1,
                                                  beta2, eps, weight_decay, grad_scale, sk);
  cal::note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int cal_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                             int32_t* step, float lr, const float* lr_device, float beta1, float beta2,
                             float eps, float weight_decay, float grad_scale, void* stream) {
  return cal_adam_step_images(params, grads, exp_avg, exp_avg_sq, n, step, lr, lr_device, beta1, beta2, eps,
                              weight_decay, grad_scale, nullptr, stream);
}

extern "C" int cal_adam_tick(int32_t* step, void* stream) {
  if (!step) return CAL_ENULL;
  cal::launch_k(cal::k_tick, dim3(1), dim3(1), 0, (cudaStream_t)stream, step);
  cal::note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}
