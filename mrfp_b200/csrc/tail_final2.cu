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

int tail_args(const mrfp_hrfp_plan* P, const void* saved, const void* lut, const float* w2, int K, TailArgs* a) {
  if (P->mode != MRFP_MATH_BF16) return MRFP_ERR_UNSUPPORTED;
  const HrfpStage& st = P->st[3];
  if (st.cout != kC || K <= 0 || K > kCls) return MRFP_ERR_UNSUPPORTED;
  const int* hidx = P->lut.data() + st.idx_w;           // every tile's source span must fit the buffer
  for (int w0 = 0; w0 < st.ow; w0 += kPx) {
    const int w1 = (w0 + kPx < st.ow ? w0 + kPx : st.ow) - 1;
    if (hidx[w1] - hidx[w0] + 1 > kSpan) return MRFP_ERR_UNSUPPORTED;
  }
  const float* stats = reinterpret_cast<const float*>((const char*)saved + P->stats_off) + (size_t)3 * 4 * kMaxC;
  a->y = reinterpret_cast<const __nv_bfloat16*>((const char*)saved + st.y_off);
  a->idx_h = (const int*)lut + st.idx_h; a->idx_w = (const int*)lut + st.idx_w;
  a->scale = stats + 2 * kMaxC; a->shift = stats + 3 * kMaxC;
  a->N = P->N; a->IH = st.ch; a->IW = st.cw; a->OH = st.oh; a->OW = st.ow; a->K = K;
  a->w2 = w2;
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
  TailArgs a;
  int rc = tail_args(P, saved, lut, w2, K, &a);
  if (rc) return rc;
  // the staged low-resolution span: an Upsample by >= 2 (as the reference's), rows 16-byte aligned for the bulk copies
  if (lw > a.OW || lh > a.OH || !(a.OW > 1 && 2 * (lw - 1) <= a.OW - 1) || (lw & 3) || ((uintptr_t)t_lo & 15)) return MRFP_ERR_UNSUPPORTED;
  DeviceInfo di;
  rc = get_device_info(&di);
  if (rc) return rc;
  const int wtiles = (a.OW + kPx - 1) / kPx;
  const long long tiles = (long long)a.N * a.OH * wtiles;
  if (tiles > 0x7fffffffLL) return MRFP_ERR_BAD_SHAPE;
  MRFP_SMEM_OPT_IN(tail_final2_fwd_kernel, kFwdSmem, di.device);
  const int grid = (int)(tiles < 2LL * di.sm_count ? tiles : 2LL * di.sm_count);
  launch_k(tail_final2_fwd_kernel, dim3(grid), dim3(kThreads), kFwdSmem, (cudaStream_t)stream, a, t_lo, lh, lw, b2, out, wtiles,
           (int)tiles);
  MRFP_CUDA_TRY(cudaGetLastError());
  return MRFP_OK;
}

extern "C" int mrfp_hrfp_tail_final2_bwd(const mrfp_hrfp_plan_t* P, const void* saved, const void* lut, const float* g,
                                         const float* w2, int K, void* g_dec_nhwc, float* g_w2, float* g_b2, void* stream) {
  if (!P || P->magic != kPlanMagic) return MRFP_ERR_BAD_PLAN;
  if (!saved || !lut || !g || !w2 || !g_dec_nhwc || !g_w2 || !g_b2) return MRFP_ERR_NULL_POINTER;
  if ((uintptr_t)g_dec_nhwc & 15) return MRFP_ERR_WORKSPACE;
  TailArgs a;
  int rc = tail_args(P, saved, lut, w2, K, &a);
  if (rc) return rc;
  DeviceInfo di;
  rc = get_device_info(&di);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  MRFP_CUDA_TRY(cudaMemsetAsync(g_w2, 0, (size_t)K * kC * sizeof(float), s));
  MRFP_CUDA_TRY(cudaMemsetAsync(g_b2, 0, (size_t)K * sizeof(float), s));
  const int wtiles = (a.OW + kPx - 1) / kPx;
  const long long tiles = (long long)a.N * a.OH * wtiles;
  if (tiles > 0x7fffffffLL) return MRFP_ERR_BAD_SHAPE;
  MRFP_SMEM_OPT_IN(tail_final2_bwd_kernel, kBwdSmem, di.device);
  const int grid = (int)(tiles < 2LL * di.sm_count ? tiles : 2LL * di.sm_count);
  launch_k(tail_final2_bwd_kernel, dim3(grid), dim3(kThreads), kBwdSmem, s, a, g, reinterpret_cast<__nv_bfloat16*>(g_dec_nhwc), g_w2,
           g_b2, wtiles, (int)tiles);
  MRFP_CUDA_TRY(cudaGetLastError());
  return MRFP_OK;
}
