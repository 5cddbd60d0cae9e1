// NP+ forward / backward for sm_100a — replaces MRFPPlus.Normalization_Perturbation_Plus
// (/root/reference/deepv3.py:268-277) and its autograd backward (closed form).
//
// Both directions are the same memory-bound shape:  y = a[n,c] * x + b[n,c]  where (a, b) depend on the
// plane sums of x for ALL planes of the local batch (batch std of the plane means + global channel max).
// One cooperative persistent kernel, one CTA per SM:
//   phase A  each CTA streams its contiguous range of "units" (a unit = <= 32 KiB slice of one plane = four
//            128-bit loads per thread), fp32 per-thread partials -> warp-shuffle tree in double -> one partial
//            per unit; the loads of unit u+1 are issued before the reduction barrier of unit u, and the LAST
//            units of the range are kept in shared memory (up to 7 x 32 KiB per SM);
//   barrier  grid-wide (cooperative groups);
//   stats    every CTA derives d[c], max_c d (and the backward's extra reductions) from the unit partials
//            (<= N*C*K doubles, L2 resident) and the (a, b) pair of each plane it owns;
//   phase B  the CTA walks its range BACKWARDS: first the units still in shared memory (no re-read at all),
//            then the rest, most-recently-read first so the re-read is served from L2 while it lasts;
//            streaming (evict-first) stores.
// HBM traffic therefore sits between 1R+1W (everything cached on chip) and 2R+1W.
#include "common.cuh"
#include <cooperative_groups.h>
#include <math.h>

namespace cg = cooperative_groups;

namespace mrfp {
namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kUnitBytes = 32 * 1024;  // slice of a plane handled as one unit: kBatch vectors per thread
constexpr int kBatch = 4;

struct NpGeom {
  int N, C, HW;
  int P;         // planes = N*C
  int K;         // units per plane
  int Q;         // elements of vector type per unit (last unit of a plane may be shorter)
  int HWV;       // HW / VEC
  long long U;   // total units = P*K
  int cache_units;   // units kept in shared memory per CTA
  int max_local_planes;
};

template <int VEC> struct VecT;
template <> struct VecT<4> {
  using type = float4;
  static __device__ __forceinline__ float4 ld(const float4* p) { return ld_stream_f4(p); }
  static __device__ __forceinline__ void st(float4* p, const float4& v) { st_stream_f4(p, v); }
  static __device__ __forceinline__ float sum(const float4& v) { return (v.x + v.y) + (v.z + v.w); }
  static __device__ __forceinline__ float4 fma(const float4& v, float a, float b) {
    return make_float4(fmaf(a, v.x, b), fmaf(a, v.y, b), fmaf(a, v.z, b), fmaf(a, v.w, b));
  }
};
template <> struct VecT<1> {
  using type = float;
  static __device__ __forceinline__ float ld(const float* p) { return ld_stream_f1(p); }
  static __device__ __forceinline__ void st(float* p, const float& v) { st_stream_f1(p, v); }
  static __device__ __forceinline__ float sum(const float& v) { return v; }
  static __device__ __forceinline__ float fma(const float& v, float a, float b) { return fmaf(a, v, b); }
};

__device__ __forceinline__ double block_sum(double v, double* red /* kWarps */) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
#pragma unroll
  for (int i = 0; i < kWarps; ++i) t += red[i];
  return t;
}

// unit partials are written by other CTAs earlier in this launch: read through L2 (ld.global.cg)
__device__ __forceinline__ double unit_total(const double* ps, int plane, int K) {
  double s = 0;
  for (int k = 0; k < K; ++k) s += __ldcg(ps + (long long)plane * K + k);
  return s;
}

// per-channel statistics of the plane means over the local batch (deepv3.py:272): mbar, d = unbiased std
struct ChanStat {
  double mbar, d, dLds;
};

template <bool BWD>
__device__ __forceinline__ ChanStat channel_stat(const double* ps, const float* __restrict__ mean_in,
                                                 const float* __restrict__ eps, const NpGeom& g, int c) {
  ChanStat r;
  const double inv_hw = 1.0 / (double)g.HW;
  double s = 0;
  for (int n = 0; n < g.N; ++n) {
    const int p = n * g.C + c;
    s += BWD ? (double)mean_in[p] : unit_total(ps, p, g.K) * inv_hw;
  }
  r.mbar = s / (double)g.N;
  double q = 0, l = 0;
  for (int n = 0; n < g.N; ++n) {
    const int p = n * g.C + c;
    const double m = BWD ? (double)mean_in[p] : unit_total(ps, p, g.K) * inv_hw;
    q += (m - r.mbar) * (m - r.mbar);
    if (BWD) l += (double)eps[p] * m * unit_total(ps, p, g.K);   // dL/ds[c] = sum_n eps*m*G
  }
  r.d = sqrt(q / (double)(g.N - 1));   // N == 1 -> 0/0 -> NaN, as torch.std
  r.dLds = l;
  return r;
}

template <int VEC, bool BWD>
__global__ void __launch_bounds__(kThreads, 1)
npplus_kernel(const float* __restrict__ x, const float* __restrict__ alpha, const float* __restrict__ eps,
              const float* __restrict__ mean_in, float* __restrict__ out, float* __restrict__ mean_out,
              float* __restrict__ beta_out, double* __restrict__ ps, const NpGeom g) {
  using V = typename VecT<VEC>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V* cache = reinterpret_cast<V*>(smem_raw);
  float2* coef = reinterpret_cast<float2*>(smem_raw + (size_t)g.cache_units * g.Q * sizeof(V));
  __shared__ double red[kWarps];
  __shared__ double red2[2][kWarps];
  __shared__ double s_dmax, s_T;
  __shared__ int s_cstar;

  const V* xv = reinterpret_cast<const V*>(x);
  V* ov = reinterpret_cast<V*>(out);
  const int tid = threadIdx.x;
  const long long u0 = g.U * (long long)blockIdx.x / gridDim.x;
  const long long u1 = g.U * (long long)(blockIdx.x + 1) / gridDim.x;
  const long long cache_from = u1 - g.cache_units;   // units >= cache_from live in shared memory

  // ---------------- phase A: unit partial sums ----------------
  auto unit_len = [&](long long u) { return min(g.Q, g.HWV - (int)(u % g.K) * g.Q); };
  auto unit_off = [&](long long u) { return (u / g.K) * (long long)g.HWV + (u % g.K) * (long long)g.Q; };
  auto issue = [&](long long u, V (&r)[kBatch]) {
    const int len = unit_len(u);
    const V* src = xv + unit_off(u);
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      const int i = tid + j * kThreads;
      if (i < len) r[j] = VecT<VEC>::ld(src + i);
    }
  };
  {
    V cur[kBatch], nxt[kBatch];
    if (u0 < u1) issue(u0, cur);
    for (long long u = u0; u < u1; ++u) {
      if (u + 1 < u1) issue(u + 1, nxt);          // in flight across the reduction barrier below
      const int len = unit_len(u);
      V* keep = (u >= cache_from) ? cache + (size_t)(u - cache_from) * g.Q : nullptr;
      float acc[kBatch];
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        const int i = tid + j * kThreads;
        acc[j] = 0.f;
        if (i < len) {
          acc[j] = VecT<VEC>::sum(cur[j]);
          if (keep) keep[i] = cur[j];
        }
      }
      const double w = warp_sum((double)acc[0] + (double)acc[1] + (double)acc[2] + (double)acc[3]);
      double* rb = red2[u & 1];
      if ((tid & 31) == 0) rb[tid >> 5] = w;
      __syncthreads();
      if (tid == 0) {
        double t = 0;
#pragma unroll
        for (int i = 0; i < kWarps; ++i) t += rb[i];
        ps[u] = t;
      }
#pragma unroll
      for (int j = 0; j < kBatch; ++j) cur[j] = nxt[j];
    }
  }
  __threadfence();
  cg::this_grid().sync();

  // ---------------- statistics (every CTA, redundantly; (N,C)-sized) ----------------
  {
    double best = -1.0;
    int bestc = 0x7fffffff;
    bool anynan = false;
    for (int c = tid; c < g.C; c += kThreads) {
      // forward derives the plane means from the unit partials; backward receives the saved means
      const double d = channel_stat<BWD>(ps, mean_in, eps, g, c).d;
      if (d != d) anynan = true;
      if (d > best) { best = d; bestc = c; }
    }
    // block arg-max (first index wins on ties), NaN-propagating like Tensor.max (deepv3.py:273)
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oc = __shfl_xor_sync(0xffffffffu, bestc, o);
      const bool on = __shfl_xor_sync(0xffffffffu, (int)anynan, o);
      if (ob > best || (ob == best && oc < bestc)) { best = ob; bestc = oc; }
      anynan = anynan || on;
    }
    __shared__ double w_best[kWarps];
    __shared__ int w_c[kWarps];
    __shared__ int w_nan[kWarps];
    if ((tid & 31) == 0) { w_best[tid >> 5] = best; w_c[tid >> 5] = bestc; w_nan[tid >> 5] = anynan; }
    __syncthreads();
    if (tid == 0) {
      double b = w_best[0]; int bc = w_c[0]; bool nn = w_nan[0];
      for (int i = 1; i < kWarps; ++i) {
        if (w_best[i] > b || (w_best[i] == b && w_c[i] < bc)) { b = w_best[i]; bc = w_c[i]; }
        nn = nn || w_nan[i];
      }
      s_dmax = nn ? (double)NAN : b;
      s_cstar = bc;
    }
    __syncthreads();
    if (BWD) {
      double t = 0;
      const double dmax = s_dmax;
      for (int c = tid; c < g.C; c += kThreads) {
        const ChanStat cs = channel_stat<true>(ps, mean_in, eps, g, c);
        t += cs.dLds * 1.5 * cs.d / (dmax * dmax);
      }
      const double T = block_sum(t, red);
      if (tid == 0) s_T = T;
      __syncthreads();
    }
  }
  // (a, b) of every plane this CTA touches: one warp per plane, lanes over the batch dim
  const int pl0 = (int)(u0 / g.K);
  const int pl1 = (u1 > u0) ? (int)((u1 - 1) / g.K) : pl0 - 1;
  {
    const double dmax = s_dmax;
    const double inv_hw = 1.0 / (double)g.HW;
    const int lane = tid & 31, warp = tid >> 5;
    for (int j = warp; j <= pl1 - pl0; j += kWarps) {
      const int p = pl0 + j, n = p / g.C, c = p % g.C;
      double s = 0;
      for (int nn = lane; nn < g.N; nn += 32)
        s += BWD ? (double)mean_in[nn * g.C + c] : unit_total(ps, nn * g.C + c, g.K) * inv_hw;
      const double mbar = warp_sum(s) / (double)g.N;
      double q = 0, l = 0;
      for (int nn = lane; nn < g.N; nn += 32) {
        const int pp = nn * g.C + c;
        const double m = BWD ? (double)mean_in[pp] : unit_total(ps, pp, g.K) * inv_hw;
        q += (m - mbar) * (m - mbar);
        if (BWD) l += (double)eps[pp] * m * unit_total(ps, pp, g.K);
      }
      const double d = sqrt(warp_sum(q) / (double)(g.N - 1));
      const double dLds = BWD ? warp_sum(l) : 0.0;
      if (lane == 0) {
        const double a = (double)alpha[p];
        const double beta = 1.0 + (double)eps[p] * (d / dmax * 1.5);            // deepv3.py:273,275
        double b;
        if (!BWD) {
          const double m = unit_total(ps, p, g.K) * inv_hw;
          b = (beta - a) * m;                                                   // out = a*x + (beta-a)*m  (:276)
          if ((long long)p * g.K >= u0) {   // the CTA owning unit 0 of the plane publishes the side outputs
            mean_out[p] = (float)m;
            if (beta_out) beta_out[p] = (float)beta;
          }
        } else {
          const double m = (double)mean_in[p];
          const double G = unit_total(ps, p, g.K);
          double dLdd = 1.5 / dmax * dLds;
          if (c == s_cstar) dLdd -= s_T;
          // torch's std_backward zero-fills where std == 0
          const double dd_dm = (d == 0.0) ? 0.0 : (m - mbar) / ((double)(g.N - 1) * d);
          const double dLdm = (beta - a) * G + dLdd * dd_dm;
          b = dLdm * inv_hw;
        }
        coef[j] = make_float2((float)a, (float)b);
      }
    }
  }
  __syncthreads();

  // ---------------- phase B: rewrite, newest units first ----------------
  {
    V cur[kBatch], nxt[kBatch];
    long long u = u1 - 1;
    for (; u >= u0 && u >= cache_from; --u) {     // still on chip: no re-read
      const int len = unit_len(u);
      const float2 ab = coef[(int)(u / g.K) - pl0];
      const V* keep = cache + (size_t)(u - cache_from) * g.Q;
      V* dst = ov + unit_off(u);
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        const int i = tid + j * kThreads;
        if (i < len) VecT<VEC>::st(dst + i, VecT<VEC>::fma(keep[i], ab.x, ab.y));
      }
    }
    if (u >= u0) issue(u, cur);
    for (; u >= u0; --u) {
      if (u - 1 >= u0) issue(u - 1, nxt);
      const int len = unit_len(u);
      const float2 ab = coef[(int)(u / g.K) - pl0];
      V* dst = ov + unit_off(u);
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        const int i = tid + j * kThreads;
        if (i < len) VecT<VEC>::st(dst + i, VecT<VEC>::fma(cur[j], ab.x, ab.y));
      }
#pragma unroll
      for (int j = 0; j < kBatch; ++j) cur[j] = nxt[j];
    }
  }
}

struct NpLaunch {
  NpGeom g;
  int grid;
  size_t smem;
};

int plan_launch(int N, int C, int HW, int vec, const DeviceInfo& di, NpLaunch* L) {
  NpGeom& g = L->g;
  g.N = N; g.C = C; g.HW = HW; g.P = N * C;
  g.HWV = HW / vec;
  const int unit_elems = kThreads * kBatch;                      // vectors per unit (32 KiB for float4)
  g.Q = g.HWV < unit_elems ? g.HWV : unit_elems;
  g.K = (g.HWV + g.Q - 1) / g.Q;                                 // last unit of a plane may be shorter
  g.U = (long long)g.P * g.K;
  L->grid = (int)((g.U < di.sm_count) ? g.U : di.sm_count);
  const long long upc = (g.U + L->grid - 1) / L->grid;           // units per CTA (max)
  g.max_local_planes = (int)(upc / g.K + 2);
  const size_t coef_bytes = align_up((size_t)g.max_local_planes * sizeof(float2), 16);
  const size_t budget = (size_t)di.max_smem_optin - 1024 /* static smem */ - coef_bytes;
  const size_t unit_bytes = align_up((size_t)g.Q * 4 * vec, 16);
  long long cu = (long long)(budget / unit_bytes);
  if (cu > upc) cu = upc;
  // every CTA must have at least cache_units units, or the cached window would start before u0
  const long long min_upc = g.U / L->grid;
  if (cu > min_upc) cu = min_upc;
  g.cache_units = (int)cu;
  L->smem = (size_t)g.cache_units * g.Q * 4 * vec + coef_bytes;
  return MRFP_OK;
}

template <int VEC, bool BWD>
int launch(const float* x, const float* alpha, const float* eps, const float* mean_in, float* out,
           float* mean_out, float* beta_out, double* ps, const NpLaunch& L, cudaStream_t s) {
  auto kern = npplus_kernel<VEC, BWD>;
  MRFP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
  NpGeom g = L.g;
  void* args[] = {(void*)&x, (void*)&alpha, (void*)&eps, (void*)&mean_in, (void*)&out,
                  (void*)&mean_out, (void*)&beta_out, (void*)&ps, (void*)&g};
  MRFP_CUDA_TRY(cudaLaunchCooperativeKernel((void*)kern, dim3(L.grid), dim3(kThreads), args, L.smem, s));
  return MRFP_OK;
}

template <bool BWD>
int run(const float* x, const float* alpha, const float* eps, const float* mean_in, float* out, float* mean_out,
        float* beta_out, void* ws, size_t ws_bytes, int N, int C, int HW, void* stream) {
  if (!x || !alpha || !eps || !out || !ws || (BWD && !mean_in) || (!BWD && !mean_out)) return MRFP_ERR_NULL_POINTER;
  if (N <= 0 || C <= 0 || HW <= 0 || (long long)N * C > (1 << 24)) return MRFP_ERR_BAD_SHAPE;
  if (ws_bytes < mrfp_npplus_ws_bytes(N, C, HW) || ((uintptr_t)ws & 7)) return MRFP_ERR_WORKSPACE;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  const bool vec4 = (HW % 4 == 0) && (((uintptr_t)x | (uintptr_t)out) & 15) == 0;
  NpLaunch L;
  plan_launch(N, C, HW, vec4 ? 4 : 1, di, &L);
  cudaStream_t s = (cudaStream_t)stream;
  return vec4 ? launch<4, BWD>(x, alpha, eps, mean_in, out, mean_out, beta_out, (double*)ws, L, s)
              : launch<1, BWD>(x, alpha, eps, mean_in, out, mean_out, beta_out, (double*)ws, L, s);
}

}  // namespace
}  // namespace mrfp

extern "C" size_t mrfp_npplus_ws_bytes(int N, int C, int HW) {
  if (N <= 0 || C <= 0 || HW <= 0) return 0;
  // one double per unit; the scalar path has the most units per plane
  const long long per_plane = ((long long)HW + mrfp::kThreads * mrfp::kBatch - 1) / (mrfp::kThreads * mrfp::kBatch) + 1;
  return (size_t)((long long)N * C * per_plane * 8);
}

extern "C" int mrfp_npplus_fwd_f32(const float* x, const float* alpha, const float* eps, float* out, float* mean,
                                   float* beta, void* ws, size_t ws_bytes, int N, int C, int HW, void* stream) {
  return mrfp::run<false>(x, alpha, eps, nullptr, out, mean, beta, ws, ws_bytes, N, C, HW, stream);
}

extern "C" int mrfp_npplus_bwd_f32(const float* gout, const float* alpha, const float* eps, const float* mean,
                                   float* gin, void* ws, size_t ws_bytes, int N, int C, int HW, void* stream) {
  return mrfp::run<true>(gout, alpha, eps, mean, gin, nullptr, nullptr, ws, ws_bytes, N, C, HW, stream);
}
