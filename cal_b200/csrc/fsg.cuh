// fsg.cuh -- shared definitions of the fused small-graph path (fsg.cu, fsg_bwd.cu, head_ro.cu; the image sink of optim.cu / comm.cu).
//
// "Small graphs": every graph of the batch has at most kFsgRows nodes and kFsgEntries CSR entries (edges
// after self-loop surgery), and the batch has at most kSMs row blocks.  Then a CTA owns a BLOCK of whole
// graphs (<= kFsgRows rows) for the entire forward (backward) pass: message passing is local to the CTA's
// shared memory, node transforms run on the tensor cores with the node rows on the MMA's N dimension,
// and only the BatchNorm statistics cross CTAs (one in-kernel all-reduce per BatchNorm boundary).
#pragma once
#include "internal.cuh"

namespace cal {

constexpr int kFsgRows = 40;            // rows (nodes) per block: 5 groups of 8
constexpr int kFsgEntries = 320;        // CSR entries per block
constexpr int kFsgPhases = 24;          // all-reduce sites per launch
constexpr int kFsgVec = 512;            // doubles per all-reduce vector slot
constexpr int kFsgCntStride = 40;       // u32 words between the arrival counters of two sites (own 128-byte line each)
constexpr int kFsgImgPart = 16384;      // floats of one operand part (128 x 128)
constexpr int kFsgImg = 2 * kFsgImgPart;   // hi | lo

// the CAL_WS_FSG region (byte offsets)
struct FsgLayout {
  size_t info, cnt, acc, img, part, total;
};
__host__ __device__ inline size_t fsg_up(size_t x) { return (x + 255) & ~(size_t)255; }

// Per-block partial parameter gradients of the fused backward (floats per block, H = 128):
//   conv matrix j in [0, L+2): [H*H] d W | [H] d b;  attention: node_att_w [2][H] | edge_att_w as p0 | q0 | p1 | q1 |
//   node_att_b [2] | edge_att_b [2] (padded to 8H + 4);  input transform: M [F][H] | column sums [H] (k_feat_bwd).
constexpr int kFsgH = 128;
__host__ __device__ inline size_t fsg_part_conv(int j) { return (size_t)j * (kFsgH * kFsgH + kFsgH); }
__host__ __device__ inline size_t fsg_part_att(int L) { return fsg_part_conv(L + 2); }
__host__ __device__ inline size_t fsg_part_feat(int L) { return fsg_part_att(L) + 8 * kFsgH + 4; }
__host__ __device__ inline size_t fsg_part_floats(int L, int F) { return fsg_part_feat(L) + (size_t)F * kFsgH + kFsgH; }

__host__ __device__ inline FsgLayout fsg_layout(int Bm, int L, int F) {
  FsgLayout f;
  size_t o = 0;
  f.info = o;  o = fsg_up(o + (size_t)(Bm > 0 ? Bm : 1) * 32);            // i32[8] per block, written by the forward kernel: graphs [g0, g1), nodes [n0, n1), in-CSR entries [e0, e1), first out-CSR entry, fits-the-limits flag
  f.cnt = o;   o = fsg_up(o + (size_t)(2 * kFsgPhases * kFsgCntStride + 8) * 4);     // site counters of both sets + the epoch words
  f.acc = o;   o = fsg_up(o + (size_t)2 * kFsgPhases * 8 * 2 * kFsgVec * 8);       // fixed-point all-reduce accumulators, two sets (fsg_dev.cuh)
  f.img = o;   o = fsg_up(o + (size_t)((L + 2) * 2 + 1 + 6) * kFsgImg * 4);   // forward / backward image per conv matrix + feat + fc1 of the 3 readouts
  f.part = o;                                                             // partial gradients, one slot per block
  if (Bm <= kSMs && F <= 128) o = fsg_up(o + (size_t)(Bm > 0 ? Bm : 1) * fsg_part_floats(L, F) * 4);
  f.total = o;
  return f;
}

struct FsgWs {
  int* info;
  unsigned int* cnt;
  long long* acc;
  float* img;
  float* part;
};
__host__ __device__ inline FsgWs fsg_ws(const Ctx& c) {
  const FsgLayout f = fsg_layout(c.Bm, c.L, c.F);
  FsgWs w;
  w.info = reinterpret_cast<int*>(c.fsg + f.info);
  w.cnt = reinterpret_cast<unsigned int*>(c.fsg + f.cnt);
  w.acc = reinterpret_cast<long long*>(c.fsg + f.acc);
  w.img = reinterpret_cast<float*>(c.fsg + f.img);
  w.part = reinterpret_cast<float*>(c.fsg + f.part);
  return w;
}
// weight-operand images: matrix j in [0, L+2) = convs[j] / context_convs / objects_convs
__host__ __device__ inline float* fsg_img_fwd(const FsgWs& w, int j) { return w.img + (size_t)(2 * j) * kFsgImg; }
__host__ __device__ inline float* fsg_img_bwd(const FsgWs& w, int j) { return w.img + (size_t)(2 * j + 1) * kFsgImg; }
__host__ __device__ inline float* fsg_img_feat(const FsgWs& w, int L) { return w.img + (size_t)(2 * (L + 2)) * kFsgImg; }
// readout fc1 (torch Linear [out, in], "add": in = H) of head h: forward image A[m = out][k = in], backward image A[m = in][k = out]
__host__ __device__ inline float* fsg_img_fc1_fwd(const FsgWs& w, int L, int h) { return w.img + (size_t)(2 * (L + 2) + 1 + 2 * h) * kFsgImg; }
__host__ __device__ inline float* fsg_img_fc1_bwd(const FsgWs& w, int L, int h) { return w.img + (size_t)(2 * (L + 2) + 2 + 2 * h) * kFsgImg; }


// The image words of parameter i (value v) -- what k_fsg_prep derives from the whole matrix, written by the optimizer
// kernel that just produced v (cal_image_sink, include/cal_b200.h).  Image element (m, k) lives at float offset
// (k / 4) * 512 + m * 4 + k % 4 of the hi part, the lo part kFsgImgPart floats behind it.
__device__ __forceinline__ void fsg_sink_emit(const cal_image_sink& sk, long long i, float v) {
  for (int q = 0; q < sk.count; ++q) {
    const long long d = i - sk.entry[q].offset;
    if (d < 0 || d >= (long long)sk.entry[q].rows * kFsgH) continue;
    const int r = (int)(d >> 7), cc = (int)(d & 127);
    const unsigned int u = __float_as_uint(v);
    const float hi = __uint_as_float((u + 0x1000u) & 0xffffe000u), lo = v - hi;      // umma::split_tf32
    if (sk.entry[q].dst_t != nullptr) {
      float* p = sk.entry[q].dst_t + (r >> 2) * 512 + cc * 4 + (r & 3);
      p[0] = hi;
      p[kFsgImgPart] = lo;
    }
    if (sk.entry[q].dst_n != nullptr) {
      float* p = sk.entry[q].dst_n + (cc >> 2) * 512 + r * 4 + (cc & 3);
      p[0] = hi;
      p[kFsgImgPart] = lo;
    }
    return;
  }
}

// ... of four consecutive parameters i .. i + 3 (i % 4 == 0: they share a matrix row, every sink offset is a multiple
// of 4).  The four words of the "n" image are contiguous (one 16-byte store per part), those of the "t" image are
// 16 bytes apart.
__device__ __forceinline__ void fsg_sink_emit4(const cal_image_sink& sk, long long i, float4 v) {
  for (int q = 0; q < sk.count; ++q) {
    const long long d = i - sk.entry[q].offset;
    if (d < 0 || d >= (long long)sk.entry[q].rows * kFsgH) continue;
    const int r = (int)(d >> 7), cc = (int)(d & 127);
    const float x[4] = {v.x, v.y, v.z, v.w};
    float hi[4], lo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      hi[k] = __uint_as_float((__float_as_uint(x[k]) + 0x1000u) & 0xffffe000u);      // umma::split_tf32
      lo[k] = x[k] - hi[k];
    }
    if (sk.entry[q].dst_n != nullptr) {
      float* p = sk.entry[q].dst_n + (cc >> 2) * 512 + r * 4;
      *reinterpret_cast<float4*>(p) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(p + kFsgImgPart) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
    if (sk.entry[q].dst_t != nullptr) {
      float* p = sk.entry[q].dst_t + (r >> 2) * 512 + cc * 4 + (r & 3);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        p[4 * k] = hi[k];
        p[4 * k + kFsgImgPart] = lo[k];
      }
    }
    return;
  }
}

}  // namespace cal
