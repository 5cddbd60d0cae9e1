// Backward (adjoint) of the reference's `Upsample` — F.interpolate(mode='bilinear', align_corners=True),
// /root/reference/network/mynn.py:114-119 — as a GATHER, for the up-sampling uses of the path: deepv3.py:356 (dec1 to
// (h/2, w/2) in front of the HRFP+ add) and deepv3.py:362 (logits to the image size).
//
// ATen's upsample_bilinear2d_backward scatters every gradient element into four low-resolution pixels with atomicAdd
// (15 ms per training step at batch 16 for the 1.2 GB gradient of deepv3.py:356).  The transpose of an interpolation is a
// gather with the same weights: a low-resolution row y receives (1 - l) * g[oy] from every output row oy whose first tap
// is y and l * g[oy] from every oy whose first tap is y - 1 — a short run of consecutive output rows (about 2 * scale + 1).
// The host builds, per axis, the table {first output index, weights of the run} with ATen's own float arithmetic
// (src = oy * float((L-1)/(O-1)), h1 = (int)src, l = src - h1), so the result is the exact adjoint of ATen's forward.
// One CTA: 8 low-resolution rows x 64 columns of one plane; vertical pass from global memory (coalesced along the row)
// into shared memory, horizontal pass from shared memory; every gradient element is read once from HBM.
#include "common.cuh"
#include <math.h>
#include <string.h>
#include <vector>

namespace mrfp {
namespace {

constexpr int kRows = 8, kCols = 64, kThreads = 256;   // kRows == warps per CTA
constexpr int kMaxRun = 6;                               // run lengths kept in registers (an Upsample by 2 has runs of 3-4)

// table blob (4-byte words): [0] = T (run length), [1] = span_max (host use), [2 .. 2+L) = start[L], then w[L][T] floats
struct AxisTable { int T; const int* start; const float* w; };
__device__ __forceinline__ AxisTable axis(const int* blob, int L) {
  return AxisTable{blob[0], blob + 2, reinterpret_cast<const float*>(blob + 2 + L)};
}

__global__ void __launch_bounds__(kThreads)
bilinear_up_bwd_kernel(const float* __restrict__ g, float* __restrict__ gl, int LH, int LW, int OH, int OW,
                       const int* __restrict__ tab_h, const int* __restrict__ tab_w, int span_max) {
  extern __shared__ float t[];                       // [kRows][span_max]
  pdl_sync();
  const AxisTable th = axis(tab_h, LH), tw = axis(tab_w, LW);
  const int x0 = blockIdx.x * kCols, y0 = blockIdx.y * kRows;
  const size_t plane = blockIdx.z;
  const int x1 = min(x0 + kCols, LW) - 1;
  const int ox0 = tw.start[x0];
  const int span = min(tw.start[x1] + tw.T, OW) - ox0;                 // output columns this tile gathers from
  const float* gp = g + plane * (size_t)OH * OW;
  // vertical pass: warp = one low-resolution row (kRows == warps); its run of weights and (clamped) source rows sit in
  // registers, so a lane's loads of one column are independent of each other and of the next column's
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (th.T <= kMaxRun) {
    const int y = y0 + warp;
    float* trow = t + warp * span_max;
    if (y < LH) {
      const int s = th.start[y];
      float wr[kMaxRun];
      size_t off[kMaxRun];
#pragma unroll
      for (int a = 0; a < kMaxRun; ++a) {              // weights beyond the run are zero in the table; rows beyond OH are clamped
        wr[a] = a < th.T ? th.w[(size_t)y * th.T + a] : 0.f;
        off[a] = (size_t)min(s + a, OH - 1) * OW;
      }
#pragma unroll 2
      for (int j = lane; j < span; j += 32) {
        const float* col = gp + ox0 + j;
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < kMaxRun; ++a)
          if (a < th.T) acc = fmaf(wr[a], __ldg(col + off[a]), acc);
        trow[j] = acc;
      }
    } else {
      for (int j = lane; j < span; j += 32) trow[j] = 0.f;
    }
  } else {
    for (int idx = threadIdx.x; idx < kRows * span; idx += kThreads) {
      const int r = idx / span, j = idx - r * span, y = y0 + r;
      float acc = 0.f;
      if (y < LH) {
        const int s = th.start[y];
        const float* wv = th.w + (size_t)y * th.T;
        for (int a = 0; a < th.T; ++a) {
          const int oy = s + a;
          if (oy < OH) acc = fmaf(wv[a], __ldg(gp + (size_t)oy * OW + ox0 + j), acc);
        }
      }
      t[r * span_max + j] = acc;
    }
  }
  __syncthreads();
  float* op = gl + plane * (size_t)LH * LW;
  for (int idx = threadIdx.x; idx < kRows * kCols; idx += kThreads) {
    const int r = idx / kCols, c = idx - r * kCols, y = y0 + r, x = x0 + c;
    if (y < LH && x < LW) {
      const int s = tw.start[x] - ox0;
      const float* wv = tw.w + (size_t)x * tw.T;
      float acc = 0.f;
      for (int b = 0; b < tw.T; ++b)
        if (s + b < span) acc = fmaf(wv[b], t[r * span_max + s + b], acc);
      op[(size_t)y * LW + x] = acc;
    }
  }
}

// ATen: area_pixel_compute_scale<float>(in, out, align_corners = true) and the source index of output pixel o
float aten_scale(int L, int O) { return O > 1 ? (float)(L - 1) / (float)(O - 1) : 0.f; }

// taps[y] = list of (o, weight) with o ascending; returns the common run length
int build_axis(int L, int O, std::vector<int>* start, std::vector<float>* w, int* span_max) {
  const float r = aten_scale(L, O);
  std::vector<std::vector<std::pair<int, float>>> taps(L);
  for (int o = 0; o < O; ++o) {
    const float src = r * (float)o;
    const int h1 = (int)src;
    const int h1p = h1 < L - 1 ? 1 : 0;
    const float l1 = src - (float)h1, l0 = 1.f - l1;
    taps[h1].push_back({o, l0});
    if (h1p) taps[h1 + 1].push_back({o, l1});
    else taps[h1].push_back({o, l1});                 // both taps on the last row (l1 == 0 there)
  }
  int T = 1;
  start->assign(L, 0);
  std::vector<std::vector<float>> runs(L);
  for (int y = 0; y < L; ++y) {
    if (taps[y].empty()) continue;                    // cannot happen for O >= L; kept for safety
    int lo = taps[y][0].first, hi = lo;
    for (auto& p : taps[y]) { lo = p.first < lo ? p.first : lo; hi = p.first > hi ? p.first : hi; }
    (*start)[y] = lo;
    runs[y].assign(hi - lo + 1, 0.f);
    for (auto& p : taps[y]) runs[y][p.first - lo] += p.second;
    if (hi - lo + 1 > T) T = hi - lo + 1;
  }
  w->assign((size_t)L * T, 0.f);
  for (int y = 0; y < L; ++y)
    for (size_t a = 0; a < runs[y].size(); ++a) (*w)[(size_t)y * T + a] = runs[y][a];
  // widest output span of a 64-column tile
  int sm = 1;
  for (int x0 = 0; x0 < L; x0 += kCols) {
    const int x1 = (x0 + kCols < L ? x0 + kCols : L) - 1;
    int e = (*start)[x1] + T;
    if (e > O) e = O;
    if (e - (*start)[x0] > sm) sm = e - (*start)[x0];
  }
  *span_max = sm;
  return T;
}

}  // namespace
}  // namespace mrfp

using namespace mrfp;

// bytes of the gather table of one axis (L low-resolution, O output positions, O >= L)
extern "C" size_t mrfp_bilinear_bwd_table_bytes(int L, int O) {
  if (L <= 0 || O < L) return 0;
  std::vector<int> start; std::vector<float> w; int sm;
  const int T = build_axis(L, O, &start, &w, &sm);
  return (size_t)(2 + L + (size_t)L * T) * 4;
}

extern "C" int mrfp_bilinear_bwd_write_table(int L, int O, void* host_dst, size_t bytes) {
  if (!host_dst) return MRFP_ERR_NULL_POINTER;
  if (L <= 0 || O < L) return MRFP_ERR_BAD_SHAPE;
  std::vector<int> start; std::vector<float> w; int sm;
  const int T = build_axis(L, O, &start, &w, &sm);
  const size_t need = (size_t)(2 + L + (size_t)L * T) * 4;
  if (bytes < need) return MRFP_ERR_WORKSPACE;
  int* p = reinterpret_cast<int*>(host_dst);
  p[0] = T; p[1] = sm;
  memcpy(p + 2, start.data(), (size_t)L * 4);
  memcpy(p + 2 + L, w.data(), (size_t)L * T * 4);
  return MRFP_OK;
}

// gl (planes, LH, LW) = adjoint of bilinear(align_corners=True) up-sampling applied to g (planes, OH, OW).
// tab_h / tab_w: DEVICE copies of the tables of (LH, OH) / (LW, OW); span_w: word [1] of the host copy of tab_w.
extern "C" int mrfp_bilinear_up_bwd_f32(const float* g, float* gl, long long planes, int LH, int LW, int OH, int OW,
                                        const void* tab_h, const void* tab_w, int span_w, void* stream) {
  if (!g || !gl || !tab_h || !tab_w) return MRFP_ERR_NULL_POINTER;
  if (planes <= 0 || LH <= 0 || LW <= 0 || OH < LH || OW < LW || span_w <= 0 || planes > 0x7fffffffLL) return MRFP_ERR_BAD_SHAPE;
  const size_t smem = (size_t)kRows * span_w * sizeof(float);
  if (smem > (48u << 10)) return MRFP_ERR_UNSUPPORTED;           // scale factors beyond ~x20
  if (planes > 65535) {                                          // gridDim.z limit: fold into chunks
    for (long long p0 = 0; p0 < planes; p0 += 65535) {
      const long long n = planes - p0 < 65535 ? planes - p0 : 65535;
      int rc = mrfp_bilinear_up_bwd_f32(g + (size_t)p0 * OH * OW, gl + (size_t)p0 * LH * LW, n, LH, LW, OH, OW, tab_h, tab_w, span_w, stream);
      if (rc) return rc;
    }
    return MRFP_OK;
  }
  dim3 grid((LW + kCols - 1) / kCols, (LH + kRows - 1) / kRows, (unsigned)planes);
  launch_k(bilinear_up_bwd_kernel, grid, dim3(kThreads), smem, (cudaStream_t)stream, g, gl, LH, LW, OH, OW,
           reinterpret_cast<const int*>(tab_h), reinterpret_cast<const int*>(tab_w), span_w);
  MRFP_CUDA_TRY(cudaGetLastError());
  return MRFP_OK;
}
