// NP+ with its statistics taken by the PRODUCER of the feature, sm_100a.
//
// deepv3.py:332-335 applies NP+ to the output of layer1, whose last operation is a ReLU (Resnet.py:218-225).  The
// standalone NP+ kernel (npplus.cu) has to read its input twice (plane sums first, 2R+1W); here the ReLU pass — which
// touches every element anyway — also leaves the plane sums of its OUTPUT, so NP+ itself is one streaming pass:
//   mrfp_relu_psum_f32            y = max(x, 0), psum[n,c] = sum_hw y            (1R+1W, replaces the trunk's ReLU)
//   mrfp_npplus_fwd_presummed_f32 out = a[n,c]*y + b[n,c] from psum              (coefficient block + 1R+1W stream)
// The backward of NP+ still needs the plane sums of the incoming gradient and stays on the ring kernel
// (mrfp_npplus_bwd_f32).  Same formulas, in double, as npplus.cu (deepv3.py:268-277).
#include "hrfp.cuh"
#include <math.h>

namespace mrfp {
namespace {

constexpr int kChunk = 4096;            // floats of one plane per block iteration (256 threads x 4 float4)

// work item = (plane, chunk); a block sums its chunk and adds it to the plane total (one double atomic per item)
__global__ void __launch_bounds__(256)
relu_psum_kernel(const float* __restrict__ x, float* __restrict__ y, double* __restrict__ psum, int HW, int nchunk,
                 long long items) {
  pdl_sync();
  __shared__ float s_part[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool vec = (HW & 3) == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
  for (long long it = blockIdx.x; it < items; it += gridDim.x) {
    const long long plane = it / nchunk;
    const int ck = (int)(it - plane * nchunk);
    const int base = ck * kChunk, len = min(kChunk, HW - base);
    const float* src = x + plane * (long long)HW + base;
    float* dst = y + plane * (long long)HW + base;
    float acc = 0.f;
    if (vec) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = (tid + u * 256) * 4;
        v[u] = i < len ? ld_stream_f4(reinterpret_cast<const float4*>(src + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = (tid + u * 256) * 4;
        v[u].x = fmaxf(v[u].x, 0.f); v[u].y = fmaxf(v[u].y, 0.f); v[u].z = fmaxf(v[u].z, 0.f); v[u].w = fmaxf(v[u].w, 0.f);
        acc += (v[u].x + v[u].y) + (v[u].z + v[u].w);
        if (i < len) *reinterpret_cast<float4*>(dst + i) = v[u];      // (plain store: the next pass re-reads it from L2 while it lasts)
      }
    } else {
      for (int i = tid; i < len; i += 256) {
        const float v = fmaxf(src[i], 0.f);
        acc += v;
        dst[i] = v;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      double t = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += (double)s_part[i];
      atomicAdd(psum + plane, t);
    }
    __syncthreads();
  }
}

// read-only sibling: psum[plane] += sum of the chunk (the plane totals of an incoming gradient, NP+ backward)
__global__ void __launch_bounds__(256)
plane_psum_kernel(const float* __restrict__ x, double* __restrict__ psum, int HW, int nchunk, long long items) {
  pdl_sync();
  __shared__ float s_part[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool vec = (HW & 3) == 0 && ((uintptr_t)x & 15) == 0;
  for (long long it = blockIdx.x; it < items; it += gridDim.x) {
    const long long plane = it / nchunk;
    const int ck = (int)(it - plane * nchunk);
    const int base = ck * kChunk, len = min(kChunk, HW - base);
    const float* src = x + plane * (long long)HW + base;
    float acc = 0.f;
    if (vec) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = (tid + u * 256) * 4;
        v[u] = i < len ? __ldg(reinterpret_cast<const float4*>(src + i)) : make_float4(0.f, 0.f, 0.f, 0.f);   // (kept in L2: re-read next)
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += (v[u].x + v[u].y) + (v[u].z + v[u].w);
    } else {
      for (int i = tid; i < len; i += 256) acc += src[i];
    }
    acc = warp_sum(acc);
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      double t = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += (double)s_part[i];
      atomicAdd(psum + plane, t);
    }
    __syncthreads();
  }
}

// one block: channel statistics of the plane means, global max, per-plane (a, b); dynamic smem = 3*C doubles
__global__ void __launch_bounds__(256)
np_coef_kernel(const double* __restrict__ psum, const float* __restrict__ alpha, const float* __restrict__ eps,
               float2* __restrict__ coef, float* __restrict__ mean_out, float* __restrict__ beta_out, int N, int C, int HW) {
  pdl_sync();
  extern __shared__ double s_d[];                        // [C] batch std of the plane means
  __shared__ double w_best[8];
  __shared__ int w_nan[8];
  __shared__ double s_dmax;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double inv_hw = 1.0 / (double)HW;
  double best = -1.0;
  bool anynan = false;
  for (int c = tid; c < C; c += 256) {
    double sm = 0;
    for (int n = 0; n < N; ++n) sm += psum[(size_t)n * C + c] * inv_hw;
    const double mbar = sm / (double)N;
    double q = 0;
    for (int n = 0; n < N; ++n) {
      const double m = psum[(size_t)n * C + c] * inv_hw;
      q += (m - mbar) * (m - mbar);
    }
    const double d = sqrt(q / (double)(N - 1));          // N == 1 -> NaN, as torch.std (deepv3.py:272)
    s_d[c] = d;
    if (d != d) anynan = true;
    if (d > best) best = d;
  }
  for (int o = 16; o > 0; o >>= 1) {
    best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
    anynan = anynan || (__shfl_xor_sync(0xffffffffu, (int)anynan, o) != 0);
  }
  if (lane == 0) { w_best[warp] = best; w_nan[warp] = anynan; }
  __syncthreads();
  if (tid == 0) {
    double b = w_best[0];
    int nn = w_nan[0];
    for (int i = 1; i < 8; ++i) { b = fmax(b, w_best[i]); nn |= w_nan[i]; }
    s_dmax = nn ? (double)NAN : b;                       // NaN-propagating like Tensor.max (deepv3.py:273)
  }
  __syncthreads();
  const double dmax = s_dmax;
  for (int p = tid; p < N * C; p += 256) {
    const int c = p % C;
    const double a = (double)alpha[p], m = psum[p] * inv_hw;
    const double beta = 1.0 + (double)eps[p] * (s_d[c] / dmax * 1.5);   // deepv3.py:273,275
    coef[p] = make_float2((float)a, (float)((beta - a) * m));            // out = a*x + (beta-a)*m  (:276)
    mean_out[p] = (float)m;
    if (beta_out) beta_out[p] = (float)beta;
  }
}

__global__ void __launch_bounds__(256)
np_apply_kernel(const float* __restrict__ x, float* __restrict__ out, const float2* __restrict__ coef, int HW, int nchunk,
                long long items) {
  pdl_sync();
  const int tid = threadIdx.x;
  const bool vec = (HW & 3) == 0 && (((uintptr_t)x | (uintptr_t)out) & 15) == 0;
  for (long long it = blockIdx.x; it < items; it += gridDim.x) {
    const long long plane = it / nchunk;
    const int ck = (int)(it - plane * nchunk);
    const int base = ck * kChunk, len = min(kChunk, HW - base);
    const float2 ab = coef[plane];
    const float* src = x + plane * (long long)HW + base;
    float* dst = out + plane * (long long)HW + base;
    if (vec) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = (tid + u * 256) * 4;
        if (i < len) v[u] = ld_stream_f4(reinterpret_cast<const float4*>(src + i));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = (tid + u * 256) * 4;
        if (i < len)
          st_stream_f4(reinterpret_cast<float4*>(dst + i),
                       make_float4(fmaf(ab.x, v[u].x, ab.y), fmaf(ab.x, v[u].y, ab.y), fmaf(ab.x, v[u].z, ab.y), fmaf(ab.x, v[u].w, ab.y)));
      }
    } else {
      for (int i = tid; i < len; i += 256) dst[i] = fmaf(ab.x, src[i], ab.y);
    }
  }
}

}  // namespace

int plane_sums(const float* g, double* psum, long long planes, int HW, cudaStream_t stream) {
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  MRFP_CUDA_TRY(cudaMemsetAsync(psum, 0, (size_t)planes * sizeof(double), stream));
  const int nchunk = (HW + kChunk - 1) / kChunk;
  const long long items = planes * nchunk, cap = (long long)di.sm_count * 16;
  launch_k(plane_psum_kernel, dim3((unsigned)(items < cap ? items : cap)), dim3(256), 0, stream, g, psum, HW, nchunk, items);
  MRFP_CUDA_TRY(cudaGetLastError());
  return MRFP_OK;
}
}  // namespace mrfp

using namespace mrfp;

extern "C" int mrfp_relu_psum_f32(const float* x, float* y, double* psum, int NC, int HW, void* stream) {
  if (!x || !y || !psum) return MRFP_ERR_NULL_POINTER;
  if (NC <= 0 || HW <= 0) return MRFP_ERR_BAD_SHAPE;
  if ((uintptr_t)psum & 7) return MRFP_ERR_WORKSPACE;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  MRFP_CUDA_TRY(cudaMemsetAsync(psum, 0, (size_t)NC * sizeof(double), s));
  const int nchunk = (HW + kChunk - 1) / kChunk;
  const long long items = (long long)NC * nchunk;
  const long long cap = (long long)di.sm_count * 16;
  launch_k(relu_psum_kernel, dim3((unsigned)(items < cap ? items : cap)), dim3(256), 0, s, x, y, psum, HW, nchunk, items);
  MRFP_CUDA_TRY(cudaGetLastError());
  return MRFP_OK;
}

extern "C" size_t mrfp_npplus_presummed_ws_bytes(int N, int C) {
  if (N <= 0 || C <= 0) return 0;
  return (size_t)N * C * sizeof(float2);
}

extern "C" int mrfp_npplus_fwd_presummed_f32(const float* x, const double* psum, const float* alpha, const float* eps,
                                             float* out, float* mean, float* beta, void* ws, size_t ws_bytes, int N, int C,
                                             int HW, void* stream) {
  if (!x || !psum || !alpha || !eps || !out || !mean || !ws) return MRFP_ERR_NULL_POINTER;
  if (N <= 0 || C <= 0 || HW <= 0 || (long long)N * C > (1 << 24)) return MRFP_ERR_BAD_SHAPE;
  if (ws_bytes < mrfp_npplus_presummed_ws_bytes(N, C) || ((uintptr_t)ws & 15)) return MRFP_ERR_WORKSPACE;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  const size_t smem = (size_t)C * sizeof(double);
  if (smem > (size_t)di.max_smem_optin - 1024) return MRFP_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  float2* coef = reinterpret_cast<float2*>(ws);
  if (smem > (48u << 10))      // beyond the default limit only for C > 6144
    MRFP_CUDA_TRY(cudaFuncSetAttribute(np_coef_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  launch_k(np_coef_kernel, dim3(1), dim3(256), smem, s, psum, alpha, eps, coef, mean, beta, N, C, HW);
  const int nchunk = (HW + kChunk - 1) / kChunk;
  const long long items = (long long)N * C * nchunk;
  const long long cap = (long long)di.sm_count * 16;
  launch_k(np_apply_kernel, dim3((unsigned)(items < cap ? items : cap)), dim3(256), 0, s, x, out, (const float2*)coef, HW, nchunk,
           items);
  MRFP_CUDA_TRY(cudaGetLastError());
  return MRFP_OK;
}
