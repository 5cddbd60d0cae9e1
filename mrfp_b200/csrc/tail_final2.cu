// HRFP+ tail fused through the classifier — replaces, on the bf16 path, the op sequence of
// /root/reference/deepv3.py:356-361:  dec1 = Upsample(dec1, (h/2, w/2));  dec1 = OCout_dec + dec1;  dec2 = final2(dec1)
// with final2 = Conv2d(256, num_classes, 1, bias=True) (deepv3.py:219-220), and its autograd backward.
//
// A 1x1 convolution commutes with the (linear) bilinear interpolation, so
//     dec2 = b2 + Upsample(W2 . dec1) + W2 . OCout_dec
// The first product is a (N, K, h/4, w/4) tensor the host computes with a plain GEMM at LOW resolution; this file
// evaluates the rest per 128-pixel tile of an output row: the span of the stored conv output Y_3 the tile gathers from
// arrives in shared memory as bf16 NHWC (16-byte cp.async, XOR-swizzled), `ldmatrix` hands it to the warps as
// m16n8k16 A fragments whose per-lane row addresses perform the nearest-neighbour gather, BatchNorm + ReLU are applied to
// the fragments in registers, and `mma.sync` contracts the 256 channels against W2 (bf16, K <= 24 classes).  Neither
// OCout_dec nor the up-sampled dec1 nor their sum (1.2 GB fp32 at batch 8) ever exists; the forward writes 90 MB.
//   Tensor-core choice: the A operand needs a register-side transform (BN/ReLU of gathered rows) and the contraction is
//   11.5 GFLOP per launch against ~0.6 GB of HBM traffic — 5 % of the tensor peak keeps up with the memory system, so
//   warp-level mma.sync on register fragments (no shared-memory round trip for the transformed operand) is the better
//   fit than a tcgen05 pipeline here.
// Backward (one pass over g = dL/d dec2, (N, K, h/2, w/2) fp32):
//     dA3[p][c]   = sum_k g[k][p] W2[k][c]        -> bf16 NHWC, joins the chain as the gradient of OCout_dec (rank K)
//     gW2[k][c]  += sum_p g[k][p] OCout_dec[c][p]  (OCout_dec regenerated from Y_3 in registers; accumulators stay in
//                                                   registers across the tiles of a persistent CTA)
//     gb2[k]     += sum_p g[k][p]
// The low-resolution half (gradient of W2 . dec1 through the Upsample) is the gather kernel of bilinear.cu plus the
// host's GEMM autograd.
#include "hrfp.cuh"
#include "tma.cuh"
#include <mutex>

namespace mrfp {
namespace {

using namespace tma;

constexpr int kPx = 128;             // output pixels per tile (a segment of one output row)
constexpr int kSpan = 120;           // source pixels a tile may gather from (128 * 332/384 + 2 = 113 in the reference geometry)
constexpr int kC = 256;              // channels of OCout_dec (widths[3])
constexpr int kCls = 24;             // classes, padded to three n-blocks of 8 (forward) / 32 (backward k-blocks)
constexpr int kSeg = 76;             // staged low-resolution row pitch (floats)
constexpr int kThreads = 256;
constexpr int kOutPitch = 132;       // floats; 132 mod 32 = 4: the accumulator scatter is bank-conflict free

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// D += A (16x16, row) * B (16x8, col), bf16 operands, fp32 accumulators
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// two bf16 values of CONSECUTIVE CHANNELS (c, c+1): BN + ReLU with (scale_c, shift_c, scale_c1, shift_c1)
__device__ __forceinline__ uint32_t bn_relu_pair_ch(uint32_t v, const float4 s) {
  const float y0 = fmaxf(fmaf(s.x, __uint_as_float(v << 16), s.y), 0.f);
  const float y1 = fmaxf(fmaf(s.z, __uint_as_float(v & 0xffff0000u), s.w), 0.f);
  return pack_bf16(y0, y1);
}
// two bf16 values of ONE channel (two pixels): BN + ReLU with (scale, shift)
__device__ __forceinline__ uint32_t bn_relu_pair_px(uint32_t v, const float2 s) {
  const float y0 = fmaxf(fmaf(s.x, __uint_as_float(v << 16), s.y), 0.f);
  const float y1 = fmaxf(fmaf(s.x, __uint_as_float(v & 0xffff0000u), s.y), 0.f);
  return pack_bf16(y0, y1);
}

struct TailArgs {
  const __nv_bfloat16* y;          // Y_3 (N, IH, IW, 256) bf16 NHWC
  const int* idx_h; const int* idx_w;
  const float* scale; const float* shift;
  int N, IH, IW, OH, OW, K;
  const float* w2;                 // (K, 256) fp32
};

// the tile's source span of Y_3 -> shared memory [pixel][256 ch] (512-byte rows, 16-byte chunks XOR-swizzled by pixel & 7)
__device__ __forceinline__ void gather_span(unsigned char* ysm, const TailArgs& a, int n, int oh, int w0, int w_last, int* s0_out) {
  const int s0 = a.idx_w[w0], nsp = a.idx_w[w_last] - s0 + 1;
  const __nv_bfloat16* yrow = a.y + (((size_t)n * a.IH + a.idx_h[oh]) * a.IW + s0) * kC;
  for (int e = threadIdx.x; e < nsp * 32; e += kThreads) {
    const int p = e >> 5, ch = e & 31;
    const uint32_t dst = smem_u32(ysm + p * 512 + ((ch ^ (p & 7)) << 4));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(yrow + (size_t)p * kC + ch * 8) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  *s0_out = s0;
}

// ======================================================================================================
// forward: out (N, K, OH, OW) = b2 + bilinear(T, align_corners) + W2 . ReLU(BN(gather(Y_3)))
// ======================================================================================================
constexpr size_t kFwdSmem = (size_t)kSpan * 512 + (size_t)kCls * 512 + kC * sizeof(float2) + (size_t)2 * kCls * kSeg * 4 +
                            (size_t)kCls * kOutPitch * 4 + kPx * sizeof(float4) + kPx * sizeof(int);

__global__ void __launch_bounds__(kThreads, 2)
tail_final2_fwd_kernel(const TailArgs a, const float* __restrict__ tlo, int LH, int LW, const float* __restrict__ b2,
                       float* __restrict__ out, int wtiles, int num_tiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* ysm = smem_raw;                                                        // [span px][256 ch bf16]
  unsigned char* w2sm = ysm + (size_t)kSpan * 512;                                      // [24 classes][256 ch bf16]
  float2* ss = reinterpret_cast<float2*>(w2sm + (size_t)kCls * 512);                    // (scale, shift) per channel
  float* trow = reinterpret_cast<float*>(ss + kC);                                      // [row 0/1][class][kSeg]
  float* outst = trow + 2 * kCls * kSeg;                                                // [class][kOutPitch]
  float4* pix = reinterpret_cast<float4*>(outst + kCls * kOutPitch);                    // per pixel: o0, o1, wl0, wl1
  int* spx = reinterpret_cast<int*>(pix + kPx);                                         // per pixel: source pixel - s0
  __shared__ __align__(8) uint64_t bar;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int K = a.K;
  if (t == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  pdl_sync();
  // classifier weights -> bf16 [class][channel], rows swizzled like the activations; BN table
  for (int e = t; e < kCls * kC; e += kThreads) {
    const int k = e >> 8, c = e & 255;
    const float v = k < K ? a.w2[k * kC + c] : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(w2sm + k * 512 + (((c >> 3) ^ (k & 7)) << 4) + (c & 7) * 2) = __float2bfloat16_rn(v);
  }
  for (int c = t; c < kC; c += kThreads) ss[c] = make_float2(a.scale[c], a.shift[c]);
  __syncthreads();
  const float rh = a.OH > 1 ? (float)(LH - 1) / (float)(a.OH - 1) : 0.f;               // ATen upsample_bilinear2d(align_corners=True)
  const float rw = a.OW > 1 ? (float)(LW - 1) / (float)(a.OW - 1) : 0.f;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int wt = tile % wtiles, orow = tile / wtiles;
    const int n = orow / a.OH, oh = orow - n * a.OH, w0 = wt * kPx;
    const int w_last = min(w0 + kPx - 1, a.OW - 1);
    const float h1r = rh * (float)oh;
    const int h1 = (int)h1r, h1p = h1 < LH - 1 ? 1 : 0;
    const float hl1 = h1r - (float)h1, hl0 = 1.f - hl1;
    const int ws = (int)(rw * (float)w0) & ~3;
    const int w_hi = min(LW - 1, (int)(rw * (float)w_last) + 1);
    const int cnt = min(kSeg, (w_hi - ws + 4) & ~3);
    if (t == 0) mbar_expect_tx(&bar, (uint32_t)(2 * K * cnt) * 4u);
    __syncthreads();                                   // previous tile's readers are done; expect_tx precedes the copies
    if (t < 2 * K) {                                   // the two low-resolution rows of every class
      const int k = t >> 1, r = t & 1;
      const float* src = tlo + (((size_t)n * K + k) * LH + h1 + (r ? h1p : 0)) * LW + ws;
      bulk_load(trow + (size_t)(r * kCls + k) * kSeg, src, (uint32_t)cnt * 4u, &bar);
    }
    int s0;
    gather_span(ysm, a, n, oh, w0, w_last, &s0);
    if (t < kPx) {
      const int ow = min(w0 + t, a.OW - 1);
      const float w1r = rw * (float)ow;
      const int w1 = (int)w1r;
      const float wl1 = w1r - (float)w1;
      pix[t] = make_float4(__int_as_float(w1 - ws), __int_as_float(w1 - ws + (w1 < LW - 1 ? 1 : 0)), 1.f - wl1, wl1);
      spx[t] = a.idx_w[ow] - s0;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    mbar_wait(&bar, phase);
    phase ^= 1u;
    __syncthreads();
    // vertical blend of the staged rows, in place over row 0
    for (int e = t; e < K * (cnt >> 2); e += kThreads) {
      const int k = e / (cnt >> 2), q = e - k * (cnt >> 2);
      float4* p0 = reinterpret_cast<float4*>(trow + (size_t)k * kSeg) + q;
      const float4 u = *p0, v = *(reinterpret_cast<const float4*>(trow + (size_t)(kCls + k) * kSeg) + q);
      *p0 = make_float4(fmaf(hl1, v.x, hl0 * u.x), fmaf(hl1, v.y, hl0 * u.y), fmaf(hl1, v.z, hl0 * u.z), fmaf(hl1, v.w, hl0 * u.w));
    }
    // ---- contraction over the 256 channels: warp = 16 pixels, 3 n-blocks of 8 classes ----
    {
      const int g = lane >> 2, tq = lane & 3;
      const int amat = lane >> 3, arow = (amat & 1) * 8 + (lane & 7), akc = amat >> 1;
      const int sidx = spx[16 * warp + arow];
      const uint32_t abase = smem_u32(ysm) + (uint32_t)sidx * 512u;
      const int asw = sidx & 7;
      const int bn01 = (amat >> 1) * 8 + (lane & 7), bkc = amat & 1;
      const uint32_t bbase01 = smem_u32(w2sm) + (uint32_t)bn01 * 512u;
      const int bn2 = 16 + (lane & 7), bkc2 = (lane >> 3) & 1;
      const uint32_t bbase2 = smem_u32(w2sm) + (uint32_t)bn2 * 512u;
      float acc[3][4];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
      for (int kb = 0; kb < kC / 16; ++kb) {
        uint32_t af[4], b01[4], b2r[2];
        ldsm_x4(abase + (uint32_t)(((kb * 2 + akc) ^ asw) << 4), af);
        ldsm_x4(bbase01 + (uint32_t)(((kb * 2 + bkc) ^ (bn01 & 7)) << 4), b01);
        ldsm_x2(bbase2 + (uint32_t)(((kb * 2 + bkc2) ^ (bn2 & 7)) << 4), b2r);
        const float4 s0v = *reinterpret_cast<const float4*>(ss + kb * 16 + 2 * tq);
        const float4 s1v = *reinterpret_cast<const float4*>(ss + kb * 16 + 8 + 2 * tq);
        af[0] = bn_relu_pair_ch(af[0], s0v); af[1] = bn_relu_pair_ch(af[1], s0v);
        af[2] = bn_relu_pair_ch(af[2], s1v); af[3] = bn_relu_pair_ch(af[3], s1v);
        mma_bf16(acc[0], af, b01[0], b01[1]);
        mma_bf16(acc[1], af, b01[2], b01[3]);
        mma_bf16(acc[2], af, b2r[0], b2r[1]);
      }
#pragma unroll
      for (int nb = 0; nb < 3; ++nb) {                 // c0,c1: (pixel g, class 2tq, 2tq+1); c2,c3: pixel g + 8
        float* o = outst + (nb * 8 + 2 * tq) * kOutPitch + 16 * warp + g;
        o[0] = acc[nb][0]; o[kOutPitch] = acc[nb][1]; o[8] = acc[nb][2]; o[kOutPitch + 8] = acc[nb][3];
      }
    }
    __syncthreads();
    // ---- + bias + horizontal taps of the blended low-resolution row, fp32 NCHW rows of 128 pixels ----
    const bool vec = (a.OW & 3) == 0;
    for (int e = t; e < K * (kPx / 4); e += kThreads) {
      const int k = e >> 5, q = e & 31, px = 4 * q, ow = w0 + px;
      if (ow >= a.OW) continue;
      const float* vb = trow + (size_t)k * kSeg;
      const float bias = b2 ? b2[k] : 0.f;
      float r[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 tp = pix[px + i];
        r[i] = outst[k * kOutPitch + px + i] + bias + fmaf(tp.z, vb[__float_as_int(tp.x)], tp.w * vb[__float_as_int(tp.y)]);
      }
      float* op = out + (((size_t)n * K + k) * a.OH + oh) * a.OW + ow;
      if (vec && ow + 3 < a.OW) *reinterpret_cast<float4*>(op) = make_float4(r[0], r[1], r[2], r[3]);
      else
        for (int i = 0; i < 4 && ow + i < a.OW; ++i) op[i] = r[i];
    }
  }
}

// ======================================================================================================
// backward: dA3 (N, OH, OW, 256) bf16 = W2^T g;  gW2 (K, 256) += g . OCout_dec^T;  gb2 (K) += sum g
// ======================================================================================================
constexpr size_t kBwdSmem = (size_t)kSpan * 512 + (size_t)kPx * 64 + (size_t)32 * 256 + (size_t)kC * 64 + (size_t)kPx * 128 +
                            kC * sizeof(float2) + kPx * sizeof(int) + 32 * sizeof(float);

__global__ void __launch_bounds__(kThreads, 2)
tail_final2_bwd_kernel(const TailArgs a, const float* __restrict__ g, __nv_bfloat16* __restrict__ dA, float* __restrict__ gW2,
                       float* __restrict__ gb2, int wtiles, int num_tiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* ysm = smem_raw;                                   // [span px][256 ch bf16]
  unsigned char* gT = ysm + (size_t)kSpan * 512;                   // [128 px][32 classes bf16]   (64-byte rows)
  unsigned char* gC = gT + (size_t)kPx * 64;                       // [32 classes][128 px bf16]   (256-byte rows)
  unsigned char* w2t = gC + (size_t)32 * 256;                      // [256 ch][32 classes bf16]   (64-byte rows)
  unsigned char* stg = w2t + (size_t)kC * 64;                      // [128 px][64 ch bf16] store staging (128-byte rows)
  float2* ss = reinterpret_cast<float2*>(stg + (size_t)kPx * 128);
  int* spx = reinterpret_cast<int*>(ss + kC);
  float* s_gb = reinterpret_cast<float*>(spx + kPx);               // [32] class sums of g
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int K = a.K;
  const int g8 = lane >> 2, tq = lane & 3;
  pdl_sync();
  // W2^T -> bf16 [channel][class], 16-byte chunks swizzled by (channel >> 1) & 3
  for (int e = t; e < kC * 32; e += kThreads) {
    const int c = e >> 5, k = e & 31;
    const float v = k < K ? a.w2[k * kC + c] : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(w2t + c * 64 + (((k >> 3) ^ ((c >> 1) & 3)) << 4) + (k & 7) * 2) = __float2bfloat16_rn(v);
  }
  if (t < 32) s_gb[t] = 0.f;
  for (int c = t; c < kC; c += kThreads) ss[c] = make_float2(a.scale[c], a.shift[c]);
  // weight-gradient accumulators of this CTA: warp = 32 channels (4 n-blocks) x 2 m-tiles of 16 classes
  float wacc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) wacc[i][j][q] = 0.f;
  __syncthreads();

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int wt = tile % wtiles, orow = tile / wtiles;
    const int n = orow / a.OH, oh = orow - n * a.OH, w0 = wt * kPx;
    const int w_last = min(w0 + kPx - 1, a.OW - 1);
    __syncthreads();                                   // previous tile's readers are done
    int s0;
    gather_span(ysm, a, n, oh, w0, w_last, &s0);
    if (t < kPx) spx[t] = a.idx_w[min(w0 + t, a.OW - 1)] - s0;
    // g tile (32 classes x 128 pixels; zero beyond K and beyond the row end) -> bf16 in both orientations
    for (int e = t; e < 32 * (kPx / 2); e += kThreads) {
      const int k = e >> 6, pp = (e & 63) * 2, ow = w0 + pp;
      float v0 = 0.f, v1 = 0.f;
      if (k < K) {
        const float* gp = g + (((size_t)n * K + k) * a.OH + oh) * a.OW + ow;
        if (ow < a.OW) v0 = __ldg(gp);
        if (ow + 1 < a.OW) v1 = __ldg(gp + 1);
      }
      // gC[k][pp..pp+1]: 256-byte rows, chunk (pp >> 3) swizzled by k & 7
      *reinterpret_cast<uint32_t*>(gC + k * 256 + (((pp >> 3) ^ (k & 7)) << 4) + (pp & 7) * 2) = pack_bf16(v0, v1);
      // gT[pp][k], gT[pp+1][k]: 64-byte rows, chunk (k >> 3) swizzled by (px >> 1) & 3
      *reinterpret_cast<__nv_bfloat16*>(gT + pp * 64 + (((k >> 3) ^ ((pp >> 1) & 3)) << 4) + (k & 7) * 2) = __float2bfloat16_rn(v0);
      *reinterpret_cast<__nv_bfloat16*>(gT + (pp + 1) * 64 + (((k >> 3) ^ (((pp + 1) >> 1) & 3)) << 4) + (k & 7) * 2) = __float2bfloat16_rn(v1);
      if (k < K) {                                     // bias gradient: the 64 lanes of a class row reduce through shuffles
        float sg = v0 + v1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sg += __shfl_xor_sync(0xffffffffu, sg, o);
        if (lane == 0) atomicAdd(&s_gb[k], sg);
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ---- (1) dA3[px][ch] = sum_k g[k][px] W2[k][ch]: warp = 16 pixels, four passes of 64 channels ----
    {
      const int amat = lane >> 3, arow = (amat & 1) * 8 + (lane & 7), akc = amat >> 1;
      const int prow = 16 * warp + arow;
      uint32_t af[2][4];
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
        ldsm_x4(smem_u32(gT + prow * 64 + (((kb * 2 + akc) ^ ((prow >> 1) & 3)) << 4)), af[kb]);
      const int bn = (amat >> 1) * 8 + (lane & 7), bkc = amat & 1;     // rows of the B matrices: channels
#pragma unroll 1
      for (int pass = 0; pass < 4; ++pass) {
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
        for (int np = 0; np < 4; ++np) {               // pairs of n-blocks: 16 channels
          const int ch = pass * 64 + np * 16 + bn;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            uint32_t bf[4];
            ldsm_x4(smem_u32(w2t + ch * 64 + (((kb * 2 + bkc) ^ ((ch >> 1) & 3)) << 4)), bf);
            mma_bf16(acc[2 * np], af[kb], bf[0], bf[1]);
            mma_bf16(acc[2 * np + 1], af[kb], bf[2], bf[3]);
          }
        }
        __syncthreads();                               // staging tile free (previous pass stored)
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {               // c0,c1: (pixel g8, channels nb*8 + 2tq, +1); c2,c3: pixel g8 + 8
          const int p0 = 16 * warp + g8, p1 = p0 + 8;
          *reinterpret_cast<uint32_t*>(stg + p0 * 128 + ((nb ^ (p0 & 7)) << 4) + tq * 4) = pack_bf16(acc[nb][0], acc[nb][1]);
          *reinterpret_cast<uint32_t*>(stg + p1 * 128 + ((nb ^ (p1 & 7)) << 4) + tq * 4) = pack_bf16(acc[nb][2], acc[nb][3]);
        }
        __syncthreads();
        __nv_bfloat16* drow = dA + (((size_t)n * a.OH + oh) * a.OW + w0) * kC + pass * 64;
#pragma unroll
        for (int q = 0; q < kPx * 8 / kThreads; ++q) {
          const int e = t + q * kThreads, px = e >> 3, chunk = e & 7;
          if (w0 + px < a.OW)
            *reinterpret_cast<uint4*>(drow + (size_t)px * kC + chunk * 8) =
                *reinterpret_cast<const uint4*>(stg + px * 128 + ((chunk ^ (px & 7)) << 4));
        }
      }
    }

    // ---- (2) gW2[k][ch] += sum_px g[k][px] ReLU(BN(Y_3[src(px)][ch])): warp = channels [32 warp, +32) ----
    {
      const int amat = lane >> 3, arow = (amat & 1) * 8 + (lane & 7), akc = amat >> 1;   // A = gC: rows = classes, k = pixels
      const int tm = lane >> 3;                        // B (.trans): matrices (k 0-7 | k 8-15) x (n-block j | j + 1)
      const int brow = (tm & 1) * 8 + (lane & 7);      // pixel within the k-block
      const int bnb = tm >> 1;                         // n-block within the pair
      float2 sc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) sc[j] = ss[32 * warp + j * 8 + g8];
#pragma unroll 2
      for (int kb = 0; kb < kPx / 16; ++kb) {
        uint32_t a0[4], a1[4];
        ldsm_x4(smem_u32(gC + arow * 256 + (((kb * 2 + akc) ^ (arow & 7)) << 4)), a0);
        ldsm_x4(smem_u32(gC + (16 + arow) * 256 + (((kb * 2 + akc) ^ ((16 + arow) & 7)) << 4)), a1);
        const int sidx = spx[kb * 16 + brow];
#pragma unroll
        for (int jp = 0; jp < 2; ++jp) {               // n-block pairs: channels 32 warp + 16 jp + {0..7, 8..15}
          const int chunk = 4 * warp + 2 * jp + bnb;   // 16-byte chunk (8 channels) of the 512-byte pixel row
          uint32_t bf[4];
          ldsm_x4_t(smem_u32(ysm) + (uint32_t)sidx * 512u + (uint32_t)((chunk ^ (sidx & 7)) << 4), bf);
          bf[0] = bn_relu_pair_px(bf[0], sc[2 * jp]); bf[1] = bn_relu_pair_px(bf[1], sc[2 * jp]);
          bf[2] = bn_relu_pair_px(bf[2], sc[2 * jp + 1]); bf[3] = bn_relu_pair_px(bf[3], sc[2 * jp + 1]);
          mma_bf16(wacc[0][2 * jp], a0, bf[0], bf[1]);
          mma_bf16(wacc[0][2 * jp + 1], a0, bf[2], bf[3]);
          mma_bf16(wacc[1][2 * jp], a1, bf[0], bf[1]);
          mma_bf16(wacc[1][2 * jp + 1], a1, bf[2], bf[3]);
        }
      }
    }
  }
  // ---- publish the CTA's weight- and bias-gradient partials ----
  __syncthreads();
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ch = 32 * warp + j * 8 + 2 * tq;
      const int k0 = mt * 16 + g8, k1 = k0 + 8;
      if (k0 < K) { atomicAdd(gW2 + k0 * kC + ch, wacc[mt][j][0]); atomicAdd(gW2 + k0 * kC + ch + 1, wacc[mt][j][1]); }
      if (k1 < K) { atomicAdd(gW2 + k1 * kC + ch, wacc[mt][j][2]); atomicAdd(gW2 + k1 * kC + ch + 1, wacc[mt][j][3]); }
    }
  if (t < K) atomicAdd(gb2 + t, s_gb[t]);
}

// ======================================================================================================
// Second form of both kernels: 64-pixel tiles whose inputs are DOUBLE-BUFFERED TMA copies issued by one thread (the next
// tile's span of Y_3 — four SWIZZLE_128B boxes of 64 channels — and its fp32 side rows are in flight while the current
// tile is contracted), and every operand that does not change from tile to tile — the classifier fragments and the BN
// constants — held in registers of the persistent CTA instead of being re-read from shared memory for every k-block.
// What was measured on the way (tools/trace_tail.py, profiles/README.md): per-thread cp.async copies and per-row bulk
// copies cost the issuing warps 2400-4300 cycles per tile; a lane-indexed register pick compiles to divergent branches.
// ======================================================================================================
__device__ unsigned long long g_tail_dbg[16];
#define TR_DECL unsigned long long tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tr_t = clock64();
#define TR(k) do { if (threadIdx.x == 0) { const long long _n = clock64(); tr_acc[k] += (unsigned long long)(_n - tr_t); tr_t = _n; } } while (0)
#define TR_END(off) do { if (threadIdx.x == 0) { for (int _k = 0; _k < 8; ++_k) atomicAdd(&g_tail_dbg[(off) + _k], tr_acc[_k]); } } while (0)
constexpr int kPx2 = 64;             // output pixels per tile
constexpr int kSpan2 = 64;           // source pixels a tile's box holds (64 * 332/384 + 2 = 58 needed in the reference geometry)
constexpr int kSeg2 = 44;            // staged low-resolution row pitch (floats): an Upsample by >= 2 needs <= 40
constexpr int kOutPitch2 = 68;       // floats; 68 mod 32 = 4: the accumulator scatter is bank-conflict free
constexpr uint32_t kSpanBytes2 = kSpan2 * 512;

__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// a span buffer is four boxes [64 px][64 ch] (128-byte rows, 16-byte chunks XOR-swizzled by pixel & 7): address of the
// 16-byte chunk `chunk` (8 channels, 0..31) of source pixel `p`
__device__ __forceinline__ uint32_t span_addr(uint32_t base, int p, int chunk) {
  return base + (uint32_t)(chunk >> 3) * (uint32_t)(kSpan2 * 128) + (uint32_t)p * 128u + (uint32_t)(((chunk & 7) ^ (p & 7)) << 4);
}

struct TileXY { int n, oh, w0, w_last; };
// tile -> (image, output row, segment) for the tiles blockIdx.x, + gridDim.x, ...: divisions once, then carries
struct TileWalk {
  int wt, oh, n, dwt, drow, wtiles, OH, OW;
  __device__ __forceinline__ TileWalk(const TailArgs& a, int tile, int step, int wtiles_) {
    wtiles = wtiles_; OH = a.OH; OW = a.OW;
    const int orow = tile / wtiles;
    wt = tile - orow * wtiles; n = orow / OH; oh = orow - n * OH;
    drow = step / wtiles; dwt = step - drow * wtiles;
  }
  __device__ __forceinline__ void advance() {
    wt += dwt;
    const int carry = wt >= wtiles ? 1 : 0;
    wt -= carry * wtiles;
    oh += drow + carry;
    while (oh >= OH) { oh -= OH; ++n; }
  }
  __device__ __forceinline__ TileXY xy() const {
    TileXY c;
    c.n = n; c.oh = oh; c.w0 = wt * kPx2; c.w_last = min(c.w0 + kPx2 - 1, OW - 1);
    return c;
  }
};

// BatchNorm + ReLU of two bf16 values in one register: one packed fp32 FMA, one rounding to bf16x2, ReLU on the pair.
// max(round(x), 0) == round(max(x, 0)): same results as the scalar form above, 5 instructions instead of 7.
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint32_t bn_relu_x2(uint32_t v, uint64_t scale2, uint64_t shift2) {
  uint64_t x, y;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "r"(v << 16), "r"(v & 0xffff0000u));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(y) : "l"(x), "l"(scale2), "l"(shift2));
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(y));
  uint32_t p;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(hi), "f"(lo));
  asm("max.bf16x2 %0, %1, %2;" : "=r"(p) : "r"(p), "r"(0u));
  return p;
}

// dynamic shared memory, rounded up to the 1024 bytes the swizzled boxes need (by pointer arithmetic, so that the
// compiler keeps the shared address space)
__device__ __forceinline__ unsigned char* smem_1k(unsigned char* raw) {
  const uint32_t a = smem_u32(raw);
  return raw + (((a + 1023u) & ~1023u) - a);
}

constexpr size_t kFwd2Smem = 1024 + (size_t)2 * kSpanBytes2 + (size_t)4 * kCls * kOutPitch2 * 4 + (size_t)2 * 2 * kCls * kSeg2 * 4 +
                             kPx2 * sizeof(float4) + kPx2 * sizeof(int) + 16;

// forward: warp = (channel quarter q, pixel-group pair u): 2 x 16 pixels x 64 channels x 24 classes; the four quarter
// sums of a pixel meet in the epilogue
__global__ void __launch_bounds__(kThreads, 2)
tail_final2_fwd2_kernel(const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_lo, const TailArgs a,
                        int LH, int LW, const float* __restrict__ b2, float* __restrict__ out, int wtiles, int num_tiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* ysm = smem_1k(smem_raw);                                               // [2][4 boxes][64 px][64 ch bf16]
  float* outst = reinterpret_cast<float*>(ysm + (size_t)2 * kSpanBytes2);               // [quarter][class][kOutPitch2]
  float* trow = outst + 4 * kCls * kOutPitch2;                                          // [buffer][row 0/1][class][kSeg2]
  float4* pix = reinterpret_cast<float4*>(trow + 2 * 2 * kCls * kSeg2);                 // per pixel: o0, o1, wl0, wl1
  int* spx = reinterpret_cast<int*>(pix + kPx2);                                        // per pixel: source pixel - s0
  uint64_t* bar = reinterpret_cast<uint64_t*>(spx + kPx2);                              // [2] full barriers
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int K = a.K;
  const int g = lane >> 2, tq = lane & 3;
  const int amat = lane >> 3, arow = (amat & 1) * 8 + (lane & 7), akc = amat >> 1;
  const int q = warp & 3, u = warp >> 2;
  if (t == 0) {
    mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_y)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_lo)) : "memory");
  }
  pdl_sync();
  // classifier weights -> bf16 [class][channel] (swizzled rows) in the span buffer, from there into B fragments
  for (int e = t; e < kCls * kC; e += kThreads) {
    const int k = e >> 8, c = e & 255;
    const float v = k < K ? a.w2[k * kC + c] : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(ysm + k * 512 + (((c >> 3) ^ (k & 7)) << 4) + (c & 7) * 2) = __float2bfloat16_rn(v);
  }
  __syncthreads();
  uint32_t bq[4][6];
  uint64_t sq[4][4];                                    // per k-block: (scale, scale), (shift, shift) of channels 2tq, +1 | 8 + 2tq, +1
  {
    const int bn01 = (amat >> 1) * 8 + (lane & 7), bkc = amat & 1;
    const int bn2 = 16 + (lane & 7), bkc2 = (lane >> 3) & 1;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int kb = 4 * q + kk;
      uint32_t r4[4], r2[2];
      ldsm_x4(smem_u32(ysm) + (uint32_t)bn01 * 512u + (uint32_t)(((kb * 2 + bkc) ^ (bn01 & 7)) << 4), r4);
      ldsm_x2(smem_u32(ysm) + (uint32_t)bn2 * 512u + (uint32_t)(((kb * 2 + bkc2) ^ (bn2 & 7)) << 4), r2);
      bq[kk][0] = r4[0]; bq[kk][1] = r4[1]; bq[kk][2] = r4[2]; bq[kk][3] = r4[3]; bq[kk][4] = r2[0]; bq[kk][5] = r2[1];
      const int c0 = kb * 16 + 2 * tq;
      sq[kk][0] = pack_f32x2(a.scale[c0], a.scale[c0 + 1]); sq[kk][1] = pack_f32x2(a.shift[c0], a.shift[c0 + 1]);
      sq[kk][2] = pack_f32x2(a.scale[c0 + 8], a.scale[c0 + 9]); sq[kk][3] = pack_f32x2(a.shift[c0 + 8], a.shift[c0 + 9]);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the staging writes precede the TMA writes of buffer 0
  __syncthreads();                                     // the staging region becomes span buffer 0
  const float rh = a.OH > 1 ? (float)(LH - 1) / (float)(a.OH - 1) : 0.f;               // ATen upsample_bilinear2d(align_corners=True)
  const float rw = a.OW > 1 ? (float)(LW - 1) / (float)(a.OW - 1) : 0.f;
  // one thread: the tile's span of Y_3 (4 boxes) and the two low-resolution rows of every class (2 boxes of [K][kSeg2];
  // columns beyond the row end arrive as zeros)
  auto issue = [&](const TileXY& c, int b) {
    const int h1 = (int)(rh * (float)c.oh), h1p = h1 < LH - 1 ? 1 : 0;
    const int ws = (int)(rw * (float)c.w0) & ~3;
    const int pix0 = (c.n * a.IH + a.idx_h[c.oh]) * a.IW + a.idx_w[c.w0];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&bar[b], 4u * (uint32_t)(kSpan2 * 128) + 2u * (uint32_t)K * (uint32_t)(kSeg2 * 4));
    const uint32_t yb = smem_u32(ysm) + (uint32_t)b * kSpanBytes2;
#pragma unroll
    for (int j = 0; j < 4; ++j) tma_2d(yb + (uint32_t)j * (uint32_t)(kSpan2 * 128), &tm_y, &bar[b], 64 * j, pix0);
    tma_3d(smem_u32(trow + (size_t)(b * 2 + 0) * kCls * kSeg2), &tm_lo, &bar[b], ws, h1, c.n * K);
    tma_3d(smem_u32(trow + (size_t)(b * 2 + 1) * kCls * kSeg2), &tm_lo, &bar[b], ws, h1 + h1p, c.n * K);
  };
  TileWalk walk(a, blockIdx.x, gridDim.x, wtiles);
  if (t == 0) issue(walk.xy(), 0);
  TR_DECL
  int it = 0;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
    const int b = it & 1;
    const TileXY c = walk.xy();
    walk.advance();
    __syncthreads();                                   // (A) the previous tile is consumed: the other buffers are free
    TR(0);
    if (t == 0 && tile + (int)gridDim.x < num_tiles) issue(walk.xy(), b ^ 1);
    const float h1r = rh * (float)c.oh;
    const float hl1 = h1r - (float)(int)h1r, hl0 = 1.f - hl1;
    const int ws = (int)(rw * (float)c.w0) & ~3;
    if (t < kPx2) {
      const int ow = min(c.w0 + t, a.OW - 1);
      const float w1r = rw * (float)ow;
      const int w1 = (int)w1r;
      const float wl1 = w1r - (float)w1;
      pix[t] = make_float4(__int_as_float(w1 - ws), __int_as_float(w1 - ws + (w1 < LW - 1 ? 1 : 0)), 1.f - wl1, wl1);
      spx[t] = a.idx_w[ow] - a.idx_w[c.w0];
    }
    TR(1);
    mbar_wait(&bar[b], (uint32_t)(it >> 1) & 1u);
    TR(2);
    __syncthreads();                                   // (B) pix / spx are visible (the copies are, through the barrier)
    TR(4);
    float* tr = trow + (size_t)b * 2 * kCls * kSeg2;
    for (int e = t; e < K * 16; e += kThreads) {       // vertical blend of the staged rows, in place over row 0
      const int k = e >> 4, qd = e & 15;
      if (qd >= kSeg2 / 4) continue;
      float4* p0 = reinterpret_cast<float4*>(tr + (size_t)k * kSeg2) + qd;
      const float4 x0 = *p0, x1 = *(reinterpret_cast<const float4*>(tr + (size_t)(kCls + k) * kSeg2) + qd);
      *p0 = make_float4(fmaf(hl1, x1.x, hl0 * x0.x), fmaf(hl1, x1.y, hl0 * x0.y), fmaf(hl1, x1.z, hl0 * x0.z), fmaf(hl1, x1.w, hl0 * x0.w));
    }
    {
      const int sidx0 = spx[32 * u + arow], sidx1 = spx[32 * u + 16 + arow];
      const uint32_t ybase = smem_u32(ysm) + (uint32_t)b * kSpanBytes2;
      float acc[2][3][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int kb = 4 * q + kk;
        uint32_t a0[4], a1[4];
        ldsm_x4(span_addr(ybase, sidx0, kb * 2 + akc), a0);
        ldsm_x4(span_addr(ybase, sidx1, kb * 2 + akc), a1);
        a0[0] = bn_relu_x2(a0[0], sq[kk][0], sq[kk][1]); a0[1] = bn_relu_x2(a0[1], sq[kk][0], sq[kk][1]);
        a0[2] = bn_relu_x2(a0[2], sq[kk][2], sq[kk][3]); a0[3] = bn_relu_x2(a0[3], sq[kk][2], sq[kk][3]);
        a1[0] = bn_relu_x2(a1[0], sq[kk][0], sq[kk][1]); a1[1] = bn_relu_x2(a1[1], sq[kk][0], sq[kk][1]);
        a1[2] = bn_relu_x2(a1[2], sq[kk][2], sq[kk][3]); a1[3] = bn_relu_x2(a1[3], sq[kk][2], sq[kk][3]);
        mma_bf16(acc[0][0], a0, bq[kk][0], bq[kk][1]);
        mma_bf16(acc[0][1], a0, bq[kk][2], bq[kk][3]);
        mma_bf16(acc[0][2], a0, bq[kk][4], bq[kk][5]);
        mma_bf16(acc[1][0], a1, bq[kk][0], bq[kk][1]);
        mma_bf16(acc[1][1], a1, bq[kk][2], bq[kk][3]);
        mma_bf16(acc[1][2], a1, bq[kk][4], bq[kk][5]);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int nb = 0; nb < 3; ++nb) {               // c0,c1: (pixel g, class 2tq, 2tq+1); c2,c3: pixel g + 8
          float* o = outst + (size_t)q * kCls * kOutPitch2 + (nb * 8 + 2 * tq) * kOutPitch2 + 32 * u + 16 * i + g;
          o[0] = acc[i][nb][0]; o[kOutPitch2] = acc[i][nb][1]; o[8] = acc[i][nb][2]; o[kOutPitch2 + 8] = acc[i][nb][3];
        }
    }
    TR(5);
    __syncthreads();                                   // (C)
    TR(6);
    // ---- quarter sums + bias + horizontal taps of the blended low-resolution row, fp32 NCHW rows of 64 pixels ----
    const bool vec = (a.OW & 3) == 0;
    for (int e = t; e < K * (kPx2 / 4); e += kThreads) {
      const int k = e >> 4, qd = e & 15, px = 4 * qd, ow = c.w0 + px;
      if (ow >= a.OW) continue;
      const float* vb = tr + (size_t)k * kSeg2;
      const float bias = b2 ? b2[k] : 0.f;
      const float* op0 = outst + k * kOutPitch2 + px;
      const float4 p0 = *reinterpret_cast<const float4*>(op0), p1 = *reinterpret_cast<const float4*>(op0 + kCls * kOutPitch2);
      const float4 p2 = *reinterpret_cast<const float4*>(op0 + 2 * kCls * kOutPitch2), p3 = *reinterpret_cast<const float4*>(op0 + 3 * kCls * kOutPitch2);
      float r[4] = {(p0.x + p1.x) + (p2.x + p3.x), (p0.y + p1.y) + (p2.y + p3.y), (p0.z + p1.z) + (p2.z + p3.z), (p0.w + p1.w) + (p2.w + p3.w)};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 tp = pix[px + i];
        r[i] += bias + fmaf(tp.z, vb[__float_as_int(tp.x)], tp.w * vb[__float_as_int(tp.y)]);
      }
      float* op = out + (((size_t)c.n * K + k) * a.OH + c.oh) * a.OW + ow;
      if (vec && ow + 3 < a.OW) *reinterpret_cast<float4*>(op) = make_float4(r[0], r[1], r[2], r[3]);
      else
        for (int i = 0; i < 4 && ow + i < a.OW; ++i) op[i] = r[i];
    }
    TR(7);
  }
  TR_END(0);
}

// 4 x 4 transpose among the four lanes of a quad (two butterfly stages, selects only): lane tq holds r[j], j = 0..3;
// afterwards lane tq holds the r[tq] of lanes 0..3 of its quad, in lane order
__device__ __forceinline__ uint4 quad_transpose(const uint32_t (&r)[4], int tq) {
  const bool up = (tq & 2) != 0, odd = (tq & 1) != 0;
  const uint32_t t0 = __shfl_xor_sync(0xffffffffu, up ? r[0] : r[2], 2);
  const uint32_t t1 = __shfl_xor_sync(0xffffffffu, up ? r[1] : r[3], 2);
  const uint32_t a0 = up ? t0 : r[0], a1 = up ? t1 : r[1], a2 = up ? r[2] : t0, a3 = up ? r[3] : t1;
  const uint32_t u0 = __shfl_xor_sync(0xffffffffu, odd ? a0 : a1, 1);
  const uint32_t u1 = __shfl_xor_sync(0xffffffffu, odd ? a2 : a3, 1);
  uint4 o;
  o.x = odd ? u0 : a0; o.y = odd ? a1 : u0; o.z = odd ? u1 : a2; o.w = odd ? a3 : u1;
  return o;
}

constexpr size_t kBwd2Smem = 1024 + (size_t)2 * kSpanBytes2 + (size_t)32 * kPx2 * 2 + (size_t)2 * 32 * kPx2 * 4 + kPx2 * sizeof(int) + 16;

// backward: warp = channels [32 warp, +32) in both products.  g_tma: g arrives as one [K][64] box per tile; otherwise
// (rows of g not 16-byte aligned) through plain loads
__global__ void __launch_bounds__(kThreads, 2)
tail_final2_bwd2_kernel(const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_g, const TailArgs a,
                        const float* __restrict__ gsrc, int g_tma, __nv_bfloat16* __restrict__ dA, float* __restrict__ gW2,
                        float* __restrict__ gb2, int wtiles, int num_tiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* ysm = smem_1k(smem_raw);                          // [2][4 boxes][64 px][64 ch bf16]
  unsigned char* gC = ysm + (size_t)2 * kSpanBytes2;               // [32 classes][64 px bf16]   (128-byte rows, swizzled)
  float* gF = reinterpret_cast<float*>(gC + (size_t)32 * kPx2 * 2);   // [2][32 classes][64 px fp32]: g as it arrives
  int* spx = reinterpret_cast<int*>(gF + 2 * 32 * kPx2);
  uint64_t* bar = reinterpret_cast<uint64_t*>(spx + kPx2);         // [2] full barriers
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int K = a.K;
  const int g8 = lane >> 2, tq = lane & 3;
  const int amat = lane >> 3, arow = (amat & 1) * 8 + (lane & 7), akc = amat >> 1;
  if (t == 0) {
    mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_y)) : "memory");
    if (g_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_g)) : "memory");
  }
  pdl_sync();
  // W2^T -> bf16 [channel][class] (16-byte chunks swizzled by (channel >> 1) & 3) in the span buffer, then into B fragments
  for (int e = t; e < kC * 32; e += kThreads) {
    const int c = e >> 5, k = e & 31;
    const float v = k < K ? a.w2[k * kC + c] : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(ysm + c * 64 + (((k >> 3) ^ ((c >> 1) & 3)) << 4) + (k & 7) * 2) = __float2bfloat16_rn(v);
  }
  for (int e = t; e < 2 * 32 * kPx2; e += kThreads) gF[e] = 0.f;    // rows of classes >= K stay zero from here on
  __syncthreads();
  uint32_t w1b[2][2][4];                                // [16-channel pair np][class k-block][n-block 2np: b0 b1 | 2np + 1: b0 b1]
  {
    const int bn = (amat >> 1) * 8 + (lane & 7), bkc = amat & 1;
#pragma unroll
    for (int np = 0; np < 2; ++np)
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const int ch = 32 * warp + np * 16 + bn;
        ldsm_x4(smem_u32(ysm + ch * 64 + (((kb * 2 + bkc) ^ ((ch >> 1) & 3)) << 4)), w1b[np][kb]);
      }
  }
  uint64_t sc[4][2];                                    // (scale, scale), (shift, shift) of channel 32 warp + 8 j + g8
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float s_ = a.scale[32 * warp + j * 8 + g8], b_ = a.shift[32 * warp + j * 8 + g8];
    sc[j][0] = pack_f32x2(s_, s_); sc[j][1] = pack_f32x2(b_, b_);
  }
  float wacc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) wacc[i][j][r] = 0.f;
  float sg[4] = {0.f, 0.f, 0.f, 0.f};                   // bias-gradient partials of classes warp + 8 i (this lane's two pixels)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staging writes and the zero fill precede the TMA writes
  __syncthreads();                                      // the staging region becomes span buffer 0

  // one thread: the tile's span of Y_3 (4 boxes) and its [K][64] box of g (pixels beyond the row end arrive as zeros)
  auto issue = [&](const TileXY& c, int b) {
    const int pix0 = (c.n * a.IH + a.idx_h[c.oh]) * a.IW + a.idx_w[c.w0];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&bar[b], 4u * (uint32_t)(kSpan2 * 128) + (g_tma ? (uint32_t)K * (uint32_t)(kPx2 * 4) : 0u));
    const uint32_t yb = smem_u32(ysm) + (uint32_t)b * kSpanBytes2;
#pragma unroll
    for (int j = 0; j < 4; ++j) tma_2d(yb + (uint32_t)j * (uint32_t)(kSpan2 * 128), &tm_y, &bar[b], 64 * j, pix0);
    if (g_tma) tma_3d(smem_u32(gF + (size_t)b * 32 * kPx2), &tm_g, &bar[b], c.w0, c.oh, c.n * K);
  };
  auto load_g_slow = [&](const TileXY& c, int b) {       // every thread: generic loads into the staging tile
    float* dstb = gF + (size_t)b * 32 * kPx2;
    for (int e = t; e < K * kPx2; e += kThreads) {
      const int k = e >> 6, px = e & 63, ow = c.w0 + px;
      dstb[k * kPx2 + px] = ow < a.OW ? __ldg(gsrc + (((size_t)c.n * K + k) * a.OH + c.oh) * a.OW + ow) : 0.f;
    }
  };
  TileWalk walk(a, blockIdx.x, gridDim.x, wtiles);
  if (t == 0) issue(walk.xy(), 0);
  if (!g_tma) load_g_slow(walk.xy(), 0);
  TR_DECL
  int it = 0;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
    const int b = it & 1;
    const TileXY c = walk.xy();
    walk.advance();
    const bool more = tile + (int)gridDim.x < num_tiles;
    __syncthreads();                                   // (A) the previous tile is consumed
    TR(0);
    if (more) {
      if (t == 0) issue(walk.xy(), b ^ 1);
      if (!g_tma) load_g_slow(walk.xy(), b ^ 1);
    }
    TR(1);
    if (t < kPx2) spx[t] = a.idx_w[min(c.w0 + t, a.OW - 1)] - a.idx_w[c.w0];
    TR(2);
    mbar_wait(&bar[b], (uint32_t)(it >> 1) & 1u);
    TR(3);
    {
      const float* gb = gF + (size_t)b * 32 * kPx2;
#pragma unroll
      for (int i = 0; i < 4; ++i) {                    // g tile -> bf16 [class][pixel]: a warp writes one whole 128-byte row
        const int k = warp + 8 * i;
        const float2 v = *reinterpret_cast<const float2*>(gb + k * kPx2 + 2 * lane);
        *reinterpret_cast<uint32_t*>(gC + k * 128 + (((lane >> 2) ^ (k & 7)) << 4) + (lane & 3) * 4) = pack_bf16(v.x, v.y);
        sg[i] += v.x + v.y;
      }
    }
    __syncthreads();                                   // (B) gC and spx are visible
    TR(4);
    const uint32_t ybase = smem_u32(ysm) + (uint32_t)b * kSpanBytes2;

    // ---- (1) dA3[px][ch] = sum_k g[k][px] W2[k][ch]: A fragments by ldmatrix.trans from gC, results leave from registers ----
#pragma unroll 2
    for (int mt = 0; mt < kPx2 / 16; ++mt) {
      uint32_t af[2][4];
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {                 // matrices: (px 0-7 | px 8-15) x (classes 16 kb + 0-7 | + 8-15)
        const int cls = 16 * kb + (amat >> 1) * 8 + (lane & 7), chunk = 2 * mt + (amat & 1);
        ldsm_x4_t(smem_u32(gC + cls * 128 + ((chunk ^ (cls & 7)) << 4)), af[kb]);
      }
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
      for (int np = 0; np < 2; ++np)
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          mma_bf16(acc[2 * np], af[kb], w1b[np][kb][0], w1b[np][kb][1]);
          mma_bf16(acc[2 * np + 1], af[kb], w1b[np][kb][2], w1b[np][kb][3]);
        }
      uint32_t r0[4], r1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { r0[j] = pack_bf16(acc[j][0], acc[j][1]); r1[j] = pack_bf16(acc[j][2], acc[j][3]); }
      const uint4 o0 = quad_transpose(r0, tq), o1 = quad_transpose(r1, tq);     // channels 32 warp + 8 tq + 0..7
      const int p0 = 16 * mt + g8, p1 = p0 + 8;
      __nv_bfloat16* drow = dA + (((size_t)c.n * a.OH + c.oh) * a.OW + c.w0) * kC + 32 * warp + 8 * tq;
      if (c.w0 + p0 < a.OW) *reinterpret_cast<uint4*>(drow + (size_t)p0 * kC) = o0;
      if (c.w0 + p1 < a.OW) *reinterpret_cast<uint4*>(drow + (size_t)p1 * kC) = o1;
    }

    TR(5);
    // ---- (2) gW2[k][ch] += sum_px g[k][px] ReLU(BN(Y_3[src(px)][ch])) ----
    {
      const int tm = lane >> 3;                        // B (.trans): matrices (k 0-7 | k 8-15) x (n-block j | j + 1)
      const int brow = (tm & 1) * 8 + (lane & 7);      // pixel within the k-block
      const int bnb = tm >> 1;                         // n-block within the pair
#pragma unroll 2
      for (int kb = 0; kb < kPx2 / 16; ++kb) {
        uint32_t a0[4], a1[4];
        ldsm_x4(smem_u32(gC + arow * 128 + (((kb * 2 + akc) ^ (arow & 7)) << 4)), a0);
        ldsm_x4(smem_u32(gC + (16 + arow) * 128 + (((kb * 2 + akc) ^ ((16 + arow) & 7)) << 4)), a1);
        const int sidx = spx[kb * 16 + brow];
#pragma unroll
        for (int jp = 0; jp < 2; ++jp) {               // n-block pairs: channels 32 warp + 16 jp + {0..7, 8..15}
          const int chunk = 4 * warp + 2 * jp + bnb;   // 16-byte chunk (8 channels) of the pixel's 256 channels
          uint32_t bf[4];
          ldsm_x4_t(span_addr(ybase, sidx, chunk), bf);
          bf[0] = bn_relu_x2(bf[0], sc[2 * jp][0], sc[2 * jp][1]); bf[1] = bn_relu_x2(bf[1], sc[2 * jp][0], sc[2 * jp][1]);
          bf[2] = bn_relu_x2(bf[2], sc[2 * jp + 1][0], sc[2 * jp + 1][1]); bf[3] = bn_relu_x2(bf[3], sc[2 * jp + 1][0], sc[2 * jp + 1][1]);
          mma_bf16(wacc[0][2 * jp], a0, bf[0], bf[1]);
          mma_bf16(wacc[0][2 * jp + 1], a0, bf[2], bf[3]);
          mma_bf16(wacc[1][2 * jp], a1, bf[0], bf[1]);
          mma_bf16(wacc[1][2 * jp + 1], a1, bf[2], bf[3]);
        }
      }
    }
    TR(6);
  }
  TR_END(8);
  // ---- publish the CTA's weight- and bias-gradient partials ----
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ch = 32 * warp + j * 8 + 2 * tq;
      const int k0 = mt * 16 + g8, k1 = k0 + 8;
      if (k0 < K) { atomicAdd(gW2 + k0 * kC + ch, wacc[mt][j][0]); atomicAdd(gW2 + k0 * kC + ch + 1, wacc[mt][j][1]); }
      if (k1 < K) { atomicAdd(gW2 + k1 * kC + ch, wacc[mt][j][2]); atomicAdd(gW2 + k1 * kC + ch + 1, wacc[mt][j][3]); }
    }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float s = warp_sum(sg[i]);
    if (lane == 0 && warp + 8 * i < K) atomicAdd(gb2 + warp + 8 * i, s);
  }
}

int tail_args(const mrfp_hrfp_plan* P, const void* saved, const void* lut, const float* w2, int K, TailArgs* a, int px, int span) {
  if (P->mode != MRFP_MATH_BF16) return MRFP_ERR_UNSUPPORTED;
  const HrfpStage& st = P->st[3];
  if (st.cout != kC || K <= 0 || K > kCls) return MRFP_ERR_UNSUPPORTED;
  const int* hidx = P->lut.data() + st.idx_w;           // every tile's source span must fit the buffer
  for (int w0 = 0; w0 < st.ow; w0 += px) {
    const int w1 = (w0 + px < st.ow ? w0 + px : st.ow) - 1;
    if (hidx[w1] - hidx[w0] + 1 > span) return MRFP_ERR_UNSUPPORTED;
  }
  const float* stats = reinterpret_cast<const float*>((const char*)saved + P->stats_off) + (size_t)3 * 4 * kMaxC;
  a->y = reinterpret_cast<const __nv_bfloat16*>((const char*)saved + st.y_off);
  a->idx_h = (const int*)lut + st.idx_h; a->idx_w = (const int*)lut + st.idx_w;
  a->scale = stats + 2 * kMaxC; a->shift = stats + 3 * kMaxC;
  a->N = P->N; a->IH = st.ch; a->IW = st.cw; a->OH = st.oh; a->OW = st.ow; a->K = K;
  a->w2 = w2;
  return MRFP_OK;
}

// the launch's two descriptors, cached in the plan per direction and re-encoded when an address or a shape changes
int tail_maps(const mrfp_hrfp_plan* P, int dir, const TailArgs& a, const float* aux, int d0, int d1, int box_w, bool want_aux,
              TailMaps* out) {
  std::lock_guard<std::mutex> lock(P->mu);
  TailMaps* m = &P->maps_tail[dir];
  if (!m->valid || m->key[0] != a.y || m->key[1] != aux || m->k != a.K || m->d0 != d0 || m->d1 != d1) {
    m->valid = 0;
    {
      const cuuint64_t dims[2] = {(cuuint64_t)kC, (cuuint64_t)a.N * a.IH * a.IW};
      const cuuint64_t strides[1] = {(cuuint64_t)kC * 2};
      const cuuint32_t box[2] = {64, kSpan2};
      int rc = conv_make_map(&m->y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a.y, 2, dims, strides, box, true);
      if (rc) return rc;
    }
    if (want_aux) {                                      // fp32 (N * K, d0, d1): one box = box_w columns of one row of K planes
      const cuuint64_t dims[3] = {(cuuint64_t)d1, (cuuint64_t)d0, (cuuint64_t)a.N * a.K};
      const cuuint64_t strides[2] = {(cuuint64_t)d1 * 4, (cuuint64_t)d0 * d1 * 4};
      const cuuint32_t box[3] = {(cuuint32_t)box_w, 1, (cuuint32_t)a.K};
      int rc = conv_make_map(&m->aux, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, aux, 3, dims, strides, box, false);
      if (rc) return rc;
    } else {
      m->aux = m->y;
    }
    m->key[0] = a.y; m->key[1] = aux; m->k = a.K; m->d0 = d0; m->d1 = d1;
    m->valid = 1;
  }
  *out = *m;
  return MRFP_OK;
}

}  // namespace
}  // namespace mrfp

using namespace mrfp;

extern "C" int mrfp_hrfp_tail_final2_fwd(const mrfp_hrfp_plan_t* P, const void* saved, const void* lut, const float* t_lo,
                                         int lh, int lw, const float* w2, const float* b2, int K, float* out, void* stream) {
  if (!P || P->magic != kPlanMagic) return MRFP_ERR_BAD_PLAN;
  if (!saved || !lut || !t_lo || !w2 || !out) return MRFP_ERR_NULL_POINTER;
  if (lh <= 0 || lw <= 0) return MRFP_ERR_BAD_SHAPE;
  static const int v2 = getenv("MRFP_TAIL_V") ? atoi(getenv("MRFP_TAIL_V")) : 2;
  TailArgs a;
  int rc = tail_args(P, saved, lut, w2, K, &a, v2 == 2 ? kPx2 : kPx, v2 == 2 ? kSpan2 : kSpan);
  if (rc) return rc;
  // the staged low-resolution span: an Upsample by >= 2 (as the reference's), rows 16-byte aligned for the bulk copies
  if (lw > a.OW || lh > a.OH || !(a.OW > 1 && 2 * (lw - 1) <= a.OW - 1) || (lw & 3) || ((uintptr_t)t_lo & 15)) return MRFP_ERR_UNSUPPORTED;
  DeviceInfo di;
  rc = get_device_info(&di);
  if (rc) return rc;
  const int px = v2 == 2 ? kPx2 : kPx;
  const int wtiles = (a.OW + px - 1) / px;
  const long long tiles = (long long)a.N * a.OH * wtiles;
  if (tiles > 0x7fffffffLL) return MRFP_ERR_BAD_SHAPE;
  const int grid = (int)(tiles < 2LL * di.sm_count ? tiles : 2LL * di.sm_count);
  if (v2 == 2) {
    if ((long long)a.N * a.IH * a.IW > 0x7fffffffLL) return MRFP_ERR_BAD_SHAPE;
    TailMaps tm;
    rc = tail_maps(P, 0, a, t_lo, lh, lw, kSeg2, true, &tm);
    if (rc) return rc;
    MRFP_SMEM_OPT_IN(tail_final2_fwd2_kernel, kFwd2Smem, di.device);
    launch_k(tail_final2_fwd2_kernel, dim3(grid), dim3(kThreads), kFwd2Smem, (cudaStream_t)stream, tm.y, tm.aux, a, lh, lw, b2, out,
             wtiles, (int)tiles);
  } else {
    MRFP_SMEM_OPT_IN(tail_final2_fwd_kernel, kFwdSmem, di.device);
    launch_k(tail_final2_fwd_kernel, dim3(grid), dim3(kThreads), kFwdSmem, (cudaStream_t)stream, a, t_lo, lh, lw, b2, out, wtiles,
             (int)tiles);
  }
  MRFP_CUDA_TRY(cudaGetLastError());
  return MRFP_OK;
}

extern "C" int mrfp_hrfp_tail_final2_bwd(const mrfp_hrfp_plan_t* P, const void* saved, const void* lut, const float* g,
                                         const float* w2, int K, void* g_dec_nhwc, float* g_w2, float* g_b2, void* stream) {
  if (!P || P->magic != kPlanMagic) return MRFP_ERR_BAD_PLAN;
  if (!saved || !lut || !g || !w2 || !g_dec_nhwc || !g_w2 || !g_b2) return MRFP_ERR_NULL_POINTER;
  if ((uintptr_t)g_dec_nhwc & 15) return MRFP_ERR_WORKSPACE;
  static const int v2 = getenv("MRFP_TAIL_V") ? atoi(getenv("MRFP_TAIL_V")) : 2;
  TailArgs a;
  int rc = tail_args(P, saved, lut, w2, K, &a, v2 == 2 ? kPx2 : kPx, v2 == 2 ? kSpan2 : kSpan);
  if (rc) return rc;
  DeviceInfo di;
  rc = get_device_info(&di);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  MRFP_CUDA_TRY(cudaMemsetAsync(g_w2, 0, (size_t)K * kC * sizeof(float), s));
  MRFP_CUDA_TRY(cudaMemsetAsync(g_b2, 0, (size_t)K * sizeof(float), s));
  const int px = v2 == 2 ? kPx2 : kPx;
  const int wtiles = (a.OW + px - 1) / px;
  const long long tiles = (long long)a.N * a.OH * wtiles;
  if (tiles > 0x7fffffffLL) return MRFP_ERR_BAD_SHAPE;
  const int grid = (int)(tiles < 2LL * di.sm_count ? tiles : 2LL * di.sm_count);
  if (v2 == 2) {
    if ((long long)a.N * a.IH * a.IW > 0x7fffffffLL) return MRFP_ERR_BAD_SHAPE;
    const int g_tma = (a.OW & 3) == 0 && ((uintptr_t)g & 15) == 0;     // rows of g 16-byte aligned: one box per tile
    TailMaps tm;
    rc = tail_maps(P, 1, a, g, a.OH, a.OW, kPx2, g_tma != 0, &tm);
    if (rc) return rc;
    MRFP_SMEM_OPT_IN(tail_final2_bwd2_kernel, kBwd2Smem, di.device);
    launch_k(tail_final2_bwd2_kernel, dim3(grid), dim3(kThreads), kBwd2Smem, s, tm.y, tm.aux, a, g, g_tma,
             reinterpret_cast<__nv_bfloat16*>(g_dec_nhwc), g_w2, g_b2, wtiles, (int)tiles);
  } else {
    MRFP_SMEM_OPT_IN(tail_final2_bwd_kernel, kBwdSmem, di.device);
    launch_k(tail_final2_bwd_kernel, dim3(grid), dim3(kThreads), kBwdSmem, s, a, g, reinterpret_cast<__nv_bfloat16*>(g_dec_nhwc), g_w2,
             g_b2, wtiles, (int)tiles);
  }
  MRFP_CUDA_TRY(cudaGetLastError());
  return MRFP_OK;
}

extern "C" int mrfp_debug_tail_trace(unsigned long long* host16, int reset) {
  if (host16) MRFP_CUDA_TRY(cudaMemcpyFromSymbol(host16, mrfp::g_tail_dbg, sizeof(unsigned long long) * 16));
  if (reset) { unsigned long long z[16] = {0}; MRFP_CUDA_TRY(cudaMemcpyToSymbol(mrfp::g_tail_dbg, z, sizeof(z))); }
  return 0;
}
