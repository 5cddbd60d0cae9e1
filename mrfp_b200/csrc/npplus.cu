// NP+ forward / backward for sm_100a — replaces MRFPPlus.Normalization_Perturbation_Plus
// (/root/reference/deepv3.py:268-277) and its autograd backward (closed form).
//
// Both directions are the same memory-bound shape:  y = a[n,c] * x + b[n,c]  where (a, b) depend on the
// plane sums of x for ALL planes of the local batch (batch std of the plane means + global channel max).
// One cooperative persistent kernel, one CTA per SM, each CTA owning a contiguous range of "units"
// (a unit = <= 32 KiB slice of one plane):
//   phase A  a producer thread streams the units into a shared-memory ring with 1-D TMA bulk copies
//            (cp.async.bulk + mbarrier complete_tx); 16 consumer warps reduce each unit from shared memory
//            (128-bit LDS, fp32 lane partials, double warp-shuffle tree); the producer folds the 16 warp
//            partials of a unit when it recycles the slot.  The LAST units of the range stay resident in the
//            ring (up to 7 x 32 KiB per SM);
//   barrier  grid-wide (cooperative groups);
//   stats    every CTA derives d[c], max_c d, arg-max and the backward's extra reductions from the unit
//            partials (<= N*C*K doubles, L2 resident) and the (a, b) pair of each plane it owns;
//   phase B  the CTA walks its range BACKWARDS: first the units still resident in shared memory (no re-read
//            at all), then the rest, most-recently-read first so the re-read is served from L2 while it
//            lasts; 128-bit streaming (evict-first) stores.
// HBM traffic therefore sits between 1R+1W (everything cached on chip) and 2R+1W.
// A register-staged variant of the same algorithm (no TMA; scalar loads) handles planes whose size or base
// address is not a multiple of 16 bytes.
#include "common.cuh"
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <atomic>

namespace cg = cooperative_groups;

namespace mrfp {
namespace {

constexpr int kConsumers = 512;                 // consumer threads (16 warps)
constexpr int kWarps = kConsumers / 32;
constexpr int kBatch = 4;                       // vectors per consumer thread per unit
constexpr int kUnitVecs = kConsumers * kBatch;  // 2048 float4 = 32 KiB
constexpr int kMaxSlots = 7;

struct NpGeom {
  int N, C, HW;
  int P;         // planes = N*C
  int K;         // units per plane
  int Q;         // vectors per unit (last unit of a plane may be shorter)
  int HWV;       // HW / VEC
  long long U;   // total units = P*K
  int slots;     // ring slots / resident units per CTA
  int max_local_planes;   // scalar path: planes a CTA can touch
  int grid;      // CTAs
  int keep_units;  // ring path: units (grid-wide) loaded with an L2 evict_last hint because phase B re-reads them
  long long scratch_off;   // scalar path, doubles: ps[0..scratch_off) partials, then pm[P], then chan[4*C]
  int grab;      // ring path: units taken from the grid-wide queue per atomic
  int max_jobs;  // ring path: capacity of a CTA's job list
};

// ------------------------------------------------------------------------------------------------------
// statistics shared by both kernels
// ------------------------------------------------------------------------------------------------------
// unit partials are written by other CTAs earlier in this launch: read through L2 (ld.global.cg)
__device__ __forceinline__ double plane_total(const double* ps, int plane, const NpGeom& g) {
  double s = 0;
#pragma unroll 4
  for (int k = 0; k < g.K; ++k) s += __ldcg(ps + (long long)plane * g.K + k);
  return s;
}

__device__ __forceinline__ double group_sum(double v, int width) {   // xor-butterfly inside aligned lane groups
  for (int o = width >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct NpShared {
  double w_best[kWarps + 1], w_tsum[kWarps + 1];
  int w_c[kWarps + 1], w_nan[kWarps + 1];
  double dmax, T;
  int cstar;
};

// Stage 1 (grid-parallel, between two grid barriers): per-channel statistics of the plane means over the
// local batch (deepv3.py:272) and the plane totals, written once to global scratch:
//   pm[p]      = plane total (sum over HW) of plane p
//   chan[4c..] = { mbar, d = unbiased std of the plane means, dL/ds (backward only), - }
template <bool BWD>
__device__ void np_stage1(const double* ps, const float* __restrict__ mean_in, const float* __restrict__ eps,
                          const NpGeom& g, double* pm, double* chan, int nthreads) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = nthreads >> 5;
  const double inv_hw = 1.0 / (double)g.HW;
  // lanes = (channel sub-group, batch index): group width = pow2 >= N (capped at 32)
  int gw = 1;
  while (gw < g.N && gw < 32) gw <<= 1;
  const int cpw = 32 / gw, sub = lane % gw, cg_i = lane / gw;
  const int gwarp = blockIdx.x * nwarps + warp, gwarps = gridDim.x * nwarps;
  for (int cbase = gwarp * cpw; cbase < g.C; cbase += gwarps * cpw) {     // warp-uniform trip count
    const int c = cbase + cg_i;
    const bool valid = c < g.C;
    double s = 0, m0 = 0, G0 = 0;
    if (valid) {
      for (int n = sub; n < g.N; n += gw) {
        const int p = n * g.C + c;
        const double tot = plane_total(ps, p, g);
        pm[p] = tot;
        const double m = BWD ? (double)mean_in[p] : tot * inv_hw;
        if (n == sub) { m0 = m; G0 = tot; }
        s += m;
      }
    }
    const double mbar = group_sum(s, gw) / (double)g.N;
    double q = 0, l = 0;
    if (valid) {
      for (int n = sub; n < g.N; n += gw) {
        const int p = n * g.C + c;
        double m, G;
        if (n == sub) { m = m0; G = G0; }
        else { G = plane_total(ps, p, g); m = BWD ? (double)mean_in[p] : G * inv_hw; }
        q += (m - mbar) * (m - mbar);
        if (BWD) l += (double)eps[p] * m * G;                            // dL/ds[c] = sum_n eps*m*G
      }
    }
    const double d = sqrt(group_sum(q, gw) / (double)(g.N - 1));         // N == 1 -> 0/0 -> NaN, as torch.std
    const double dLds = BWD ? group_sum(l, gw) : 0.0;
    if (valid && sub == 0) {
      chan[4 * c + 0] = mbar;
      chan[4 * c + 1] = d;
      chan[4 * c + 2] = dLds;
    }
  }
}

// Stage 2a (every CTA): global max / arg-max of d and the backward's cross-channel sum -> sh.dmax, sh.cstar, sh.T
template <bool BWD>
__device__ void np_global_reduce(const double* chan, const NpGeom& g, NpShared& sh, int nthreads) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
  double best = -1.0, tsum = 0.0;
  int bestc = 0x7fffffff;
  bool anynan = false;
  for (int c = tid; c < g.C; c += nthreads) {
    const double d = __ldcg(chan + 4 * c + 1);
    if (d != d) anynan = true;
    if (d > best) { best = d; bestc = c; }
    if (BWD) tsum += __ldcg(chan + 4 * c + 2) * d;
  }
  // block arg-max (first index wins on ties), NaN-propagating like Tensor.max (deepv3.py:273)
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oc = __shfl_xor_sync(0xffffffffu, bestc, o);
    const int on = __shfl_xor_sync(0xffffffffu, (int)anynan, o);
    tsum += __shfl_xor_sync(0xffffffffu, tsum, o);
    if (ob > best || (ob == best && oc < bestc)) { best = ob; bestc = oc; }
    anynan = anynan || (on != 0);
  }
  if (lane == 0) { sh.w_best[warp] = best; sh.w_c[warp] = bestc; sh.w_nan[warp] = anynan; sh.w_tsum[warp] = tsum; }
  __syncthreads();
  if (tid == 0) {
    double b = sh.w_best[0], t = sh.w_tsum[0];
    int bc = sh.w_c[0], nn = sh.w_nan[0];
    for (int i = 1; i < nwarps; ++i) {
      if (sh.w_best[i] > b || (sh.w_best[i] == b && sh.w_c[i] < bc)) { b = sh.w_best[i]; bc = sh.w_c[i]; }
      nn |= sh.w_nan[i];
      t += sh.w_tsum[i];
    }
    sh.dmax = nn ? (double)NAN : b;
    sh.cstar = bc;
    sh.T = 1.5 * t / (sh.dmax * sh.dmax);     // sum_c dL/ds[c] * 1.5 * d[c] / dmax^2
  }
  __syncthreads();
}

// Stage 2b: (a, b) of plane p, y = a*x + b.  `publish`: this thread also writes the (N,C) side outputs of the plane.
template <bool BWD>
__device__ __forceinline__ float2 np_plane_coef(int p, bool publish, const double* pm, const double* chan,
                                                const float* __restrict__ mean_in, const float* __restrict__ alpha,
                                                const float* __restrict__ eps, float* __restrict__ mean_out,
                                                float* __restrict__ beta_out, const NpGeom& g, const NpShared& sh) {
  const double inv_hw = 1.0 / (double)g.HW, dmax = sh.dmax;
  const int c = p % g.C;
  const double tot = __ldcg(pm + p);
  const double mbar = __ldcg(chan + 4 * c + 0), d = __ldcg(chan + 4 * c + 1);
  const double a = (double)alpha[p];
  const double beta = 1.0 + (double)eps[p] * (d / dmax * 1.5);            // deepv3.py:273,275
  double b;
  if (!BWD) {
    const double m = tot * inv_hw;
    b = (beta - a) * m;                                                   // out = a*x + (beta-a)*m  (:276)
    if (publish) {
      mean_out[p] = (float)m;
      if (beta_out) beta_out[p] = (float)beta;
    }
  } else {
    const double m = (double)mean_in[p];
    double dLdd = 1.5 / dmax * __ldcg(chan + 4 * c + 2);
    if (c == sh.cstar) dLdd -= sh.T;
    // torch's std_backward zero-fills where std == 0
    const double dd_dm = (d == 0.0) ? 0.0 : (m - mbar) / ((double)(g.N - 1) * d);
    const double dLdm = (beta - a) * tot + dLdd * dd_dm;
    b = dLdm * inv_hw;
  }
  return make_float2((float)a, (float)b);
}

// ------------------------------------------------------------------------------------------------------
// mbarrier / bulk-copy primitives
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                          uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

// ------------------------------------------------------------------------------------------------------
// TMA-ring kernel with a grid-wide dynamic unit queue (planes and base addresses 16-byte aligned)
// ------------------------------------------------------------------------------------------------------
// SMs do not stream from HBM at the same rate (the slowest took 40 % longer than the fastest on a static
// split), so units are handed out from an atomic counter in batches of g.grab; every CTA remembers its jobs in
// order and replays them backwards in phase B.  The counter is (re)initialised by CTA 0 of each launch and
// published with a per-launch nonce, so the workspace needs no host-side clearing.
struct NpCtrl {
  unsigned long long nonce;
  unsigned int counter;
  unsigned int pad[13];
};

template <bool BWD>
__global__ void __launch_bounds__(kConsumers + 32, 1)
npplus_ring_kernel(const float* __restrict__ x, const float* __restrict__ alpha, const float* __restrict__ eps,
                   const float* __restrict__ mean_in, float* __restrict__ out, float* __restrict__ mean_out,
                   float* __restrict__ beta_out, unsigned char* ws, const NpGeom g, unsigned long long nonce,
                   unsigned long long* trace) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int S = g.slots;
  auto stamp = [&](int i) {
    if (trace && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      trace[blockIdx.x * 8 + i] = t;
    }
  };
  stamp(0);
  float4* ring = reinterpret_cast<float4*>(smem_raw);                                   // S x 32 KiB
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)S * kUnitVecs * 16);  // [S]
  uint64_t* empty = full + S;                                                           // [S]
  float2* coefj = reinterpret_cast<float2*>(empty + S);                                 // [max_jobs]
  int* joblist = reinterpret_cast<int*>(coefj + g.max_jobs);                            // [max_jobs]
  __shared__ NpShared sh;

  NpCtrl* ctrl = reinterpret_cast<NpCtrl*>(ws);
  double* ps = reinterpret_cast<double*>(ws + sizeof(NpCtrl));      // [U] unit partials
  double* pm = ps + g.U;                                            // [P] plane totals
  double* chan = pm + g.P;                                          // [4C] channel statistics
  double* wpg = chan + 4 * g.C + (size_t)blockIdx.x * g.max_jobs * kWarps;   // [max_jobs][kWarps] warp partials of this CTA

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_producer = warp == kWarps;
  const int U = (int)g.U;
  const int last_len = g.HWV - (g.K - 1) * g.Q;          // the last unit of a plane may be shorter
  const float4* xv = reinterpret_cast<const float4*>(x);
  float4* ov = reinterpret_cast<float4*>(out);

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (blockIdx.x == 0) {                               // open this launch's queue: the first batches are static
      ctrl->counter = gridDim.x * g.grab;
      __threadfence();
      *reinterpret_cast<volatile unsigned long long*>(&ctrl->nonce) = nonce;
    }
  }
  __syncthreads();

  // ---------------- phase A ----------------
  int nA = 0;                                            // jobs taken by this CTA (known after the loop)
  if (is_producer) {
    if (lane == 0) {
      uint64_t pol_keep, pol_stream;
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
      // the globally last units are the ones re-read in phase B: the newest S per CTA stay in shared memory, the
      // g.keep_units before them are asked to stay in L2
      const int keep_hi = U - (int)gridDim.x * S, keep_lo = keep_hi - g.keep_units;
      int j = 0, s = 0, k = 0;
      unsigned base = blockIdx.x * g.grab, nbase = 0;
      bool open = blockIdx.x == 0, have_next = false;
      while (base < (unsigned)U && j < g.max_jobs - 1) {
        if (!have_next && j + 2 * g.grab < g.max_jobs - 1) {          // fetch the next batch while this one streams
          if (!open) {
            while (*reinterpret_cast<volatile unsigned long long*>(&ctrl->nonce) != nonce) {}
            __threadfence();
            open = true;
          }
          nbase = atomicAdd(&ctrl->counter, (unsigned)g.grab);
          have_next = true;
        }
        for (int q = 0; q < g.grab && base + q < (unsigned)U && j < g.max_jobs - 1; ++q) {
          const int u = (int)base + q;
          if (k > 0) mbar_wait(&empty[s], (k - 1) & 1);  // all 16 warps released fill k-1 of this slot
          const int plane = u / g.K, part = u - plane * g.K;
          const uint32_t bytes = (uint32_t)(part == g.K - 1 ? last_len : g.Q) * 16u;
          joblist[j] = u;
          mbar_expect_tx(&full[s], bytes);
          bulk_load(ring + (size_t)s * kUnitVecs, xv + (long long)plane * g.HWV + part * g.Q, bytes, &full[s],
                    (u >= keep_lo && u < keep_hi) ? pol_keep : pol_stream);
          ++j;
          if (++s == S) { s = 0; ++k; }
        }
        if (!have_next) break;
        base = nbase; have_next = false;
      }
      if (k > 0) mbar_wait(&empty[s], (k - 1) & 1);
      joblist[j] = -1;                                   // end marker: completes one phase of full[s] without data
      mbar_arrive(&full[s]);
    }
  } else {
    int s = 0, k = 0;
    for (;; ++nA) {
      mbar_wait(&full[s], k & 1);
      const int u = joblist[nA];
      if (u < 0) break;
      const int plane = u / g.K, part = u - plane * g.K;
      const int len = part == g.K - 1 ? last_len : g.Q;
      const float4* src = ring + (size_t)s * kUnitVecs;
      float acc = 0.f;
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        const int i = tid + b * kConsumers;
        if (i < len) { const float4 v = src[i]; acc += (v.x + v.y) + (v.z + v.w); }
      }
      const double w = warp_sum((double)acc);
      if (lane == 0) {
        wpg[nA * kWarps + warp] = w;
        mbar_arrive(&empty[s]);                          // always released; the last S fills simply stay in the ring
      }
      if (++s == S) { s = 0; ++k; }
    }
  }
  if (tid == 0) sh.cstar = nA;                           // publish the job count to the producer warp
  __syncthreads();
  nA = sh.cstar;
  __syncthreads();
  stamp(1);
  for (int j = tid; j < nA; j += kConsumers + 32) {      // fold the 16 warp partials of each unit, fixed order
    double t = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += __ldcg(wpg + j * kWarps + w);
    ps[joblist[j]] = t;
  }
  __threadfence();
  cg::this_grid().sync();
  stamp(2);

  // ---------------- statistics ----------------
  np_stage1<BWD>(ps, mean_in, eps, g, pm, chan, kConsumers + 32);
  __threadfence();
  cg::this_grid().sync();
  np_global_reduce<BWD>(chan, g, sh, kConsumers + 32);
  for (int j = tid; j < nA; j += kConsumers + 32) {
    const int u = joblist[j], plane = u / g.K;
    coefj[j] = np_plane_coef<BWD>(plane, u - plane * g.K == 0, pm, chan, mean_in, alpha, eps, mean_out, beta_out, g, sh);
  }
  __syncthreads();
  stamp(3);

  // ---------------- phase B: this CTA's jobs, newest first ----------------
  const int s_end = nA % S;                              // the slot whose full barrier took the end-marker arrive
  auto fills_a = [&](int s) { return (s < nA ? (nA - s + S - 1) / S : 0); };   // phase-A data fills of slot s
  if (is_producer) {
    if (lane == 0 && nA > S) {
      uint64_t pol_stream;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
      int s = (nA - 1 - S) % S;
      for (int i = S; i < nA; ++i) {                     // refill with job nA-1-i once job nA-1-(i-S) left the slot
        const int u = joblist[nA - 1 - i];
        mbar_wait(&empty[s], (fills_a(s) + i / S - 1) & 1);
        const int plane = u / g.K, part = u - plane * g.K;
        const uint32_t bytes = (uint32_t)(part == g.K - 1 ? last_len : g.Q) * 16u;
        mbar_expect_tx(&full[s], bytes);
        bulk_load(ring + (size_t)s * kUnitVecs, xv + (long long)plane * g.HWV + part * g.Q, bytes, &full[s], pol_stream);
        if (--s < 0) s = S - 1;
      }
    }
  } else {
    int s = nA > 0 ? (nA - 1) % S : 0;
    for (int i = 0; i < nA; ++i) {
      const int j = nA - 1 - i, u = joblist[j];
      const int plane = u / g.K, part = u - plane * g.K;
      const int len = part == g.K - 1 ? last_len : g.Q;
      if (i >= S) mbar_wait(&full[s], (fills_a(s) + (s == s_end ? 1 : 0) + i / S - 1) & 1);
      const float2 ab = coefj[j];
      const float4* src = ring + (size_t)s * kUnitVecs;
      float4 v[kBatch];
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        const int idx = tid + b * kConsumers;
        if (idx < len) {
          const float4 t = src[idx];
          v[b] = make_float4(fmaf(ab.x, t.x, ab.y), fmaf(ab.x, t.y, ab.y), fmaf(ab.x, t.z, ab.y), fmaf(ab.x, t.w, ab.y));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      float4* dst = ov + (long long)plane * g.HWV + part * g.Q;
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        const int idx = tid + b * kConsumers;
        if (idx < len) st_stream_f4(dst + idx, v[b]);
      }
      if (--s < 0) s = S - 1;
    }
  }
  if (trace) { __syncthreads(); stamp(4); }
}

// ------------------------------------------------------------------------------------------------------
// register-staged scalar kernel (any HW, any alignment)
// ------------------------------------------------------------------------------------------------------
template <bool BWD>
__global__ void __launch_bounds__(kConsumers, 1)
npplus_scalar_kernel(const float* __restrict__ x, const float* __restrict__ alpha, const float* __restrict__ eps,
                     const float* __restrict__ mean_in, float* __restrict__ out, float* __restrict__ mean_out,
                     float* __restrict__ beta_out, double* ps, const NpGeom g, unsigned long long* /*trace*/) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* cache = reinterpret_cast<float*>(smem_raw);                                        // slots x Q floats
  float2* coef = reinterpret_cast<float2*>(smem_raw + align_up((size_t)g.slots * g.Q * 4, 16));
  __shared__ NpShared sh;
  __shared__ double red2[2][kWarps];

  const int tid = threadIdx.x;
  const long long u0 = g.U * (long long)blockIdx.x / gridDim.x;
  const long long u1 = g.U * (long long)(blockIdx.x + 1) / gridDim.x;
  const long long cache_from = u1 - g.slots;   // units >= cache_from live in shared memory
  auto unit_len = [&](long long u) { return min(g.Q, g.HWV - (int)(u % g.K) * g.Q); };
  auto unit_off = [&](long long u) { return (u / g.K) * (long long)g.HWV + (u % g.K) * (long long)g.Q; };

  for (long long u = u0; u < u1; ++u) {
    const int len = unit_len(u);
    const float* src = x + unit_off(u);
    float* keep = (u >= cache_from) ? cache + (size_t)(u - cache_from) * g.Q : nullptr;
    float acc = 0.f;
    for (int i = tid; i < len; i += kConsumers) {
      const float v = ld_stream_f1(src + i);
      acc += v;
      if (keep) keep[i] = v;
    }
    const double w = warp_sum((double)acc);
    double* rb = red2[u & 1];
    if ((tid & 31) == 0) rb[tid >> 5] = w;
    __syncthreads();
    if (tid == 0) {
      double t = 0;
#pragma unroll
      for (int i = 0; i < kWarps; ++i) t += rb[i];
      ps[u] = t;
    }
  }
  __threadfence();
  cg::this_grid().sync();
  double* pm = ps + g.scratch_off;
  double* chan = pm + g.P;
  np_stage1<BWD>(ps, mean_in, eps, g, pm, chan, kConsumers);
  __threadfence();
  cg::this_grid().sync();
  np_global_reduce<BWD>(chan, g, sh, kConsumers);
  const int pl0 = (int)(u0 / g.K);
  {
    const int pl1 = (u1 > u0) ? (int)((u1 - 1) / g.K) : pl0 - 1;
    for (int j = tid; j <= pl1 - pl0; j += kConsumers)   // the CTA owning unit 0 of a plane publishes its side outputs
      coef[j] = np_plane_coef<BWD>(pl0 + j, (long long)(pl0 + j) * g.K >= u0, pm, chan, mean_in, alpha, eps, mean_out,
                                   beta_out, g, sh);
    __syncthreads();
  }
  for (long long u = u1 - 1; u >= u0; --u) {
    const int len = unit_len(u);
    const float2 ab = coef[(int)(u / g.K) - pl0];
    float* dst = out + unit_off(u);
    if (u >= cache_from) {
      const float* keep = cache + (size_t)(u - cache_from) * g.Q;
      for (int i = tid; i < len; i += kConsumers) st_stream_f1(dst + i, fmaf(ab.x, keep[i], ab.y));
    } else {
      const float* src = x + unit_off(u);
      for (int i = tid; i < len; i += kConsumers) st_stream_f1(dst + i, fmaf(ab.x, ld_stream_f1(src + i), ab.y));
    }
  }
}

struct NpLaunch {
  NpGeom g;
  int grid;
  size_t smem;
};

// workspace layout (bytes): [NpCtrl 64][ps: units][pm: P][chan: 4C][ring path: per-CTA warp partials]
size_t ws_layout_bytes(int N, int C, int HW) {
  const long long P = (long long)N * C;
  const long long u_scalar = P * (((long long)HW + kUnitVecs - 1) / kUnitVecs);                 // 2048 floats per unit
  const long long u_ring = P * (((long long)HW / 4 + kUnitVecs - 1) / kUnitVecs + 1);           // 2048 float4 per unit
  const long long units = u_scalar > u_ring ? u_scalar : u_ring;
  const long long jobs = 2 * u_ring + 32LL * 1024;       // sum over CTAs of max_jobs (<= 2*upc + 24 each, <= 1024 CTAs)
  return (size_t)(sizeof(NpCtrl) + 8 * (units + P + 4LL * C + jobs * kWarps));
}

void plan_launch(int N, int C, int HW, bool ring, const DeviceInfo& di, NpLaunch* L) {
  NpGeom& g = L->g;
  const int vec = ring ? 4 : 1;
  g.N = N; g.C = C; g.HW = HW; g.P = N * C;
  g.HWV = HW / vec;
  g.Q = g.HWV < kUnitVecs ? g.HWV : kUnitVecs;
  g.K = (g.HWV + g.Q - 1) / g.Q;                                 // last unit of a plane may be shorter
  g.U = (long long)g.P * g.K;
  L->grid = (int)((g.U < di.sm_count) ? g.U : di.sm_count);
  g.grid = L->grid;
  const long long upc = (g.U + L->grid - 1) / L->grid;           // units per CTA (static split: max)
  const long long min_upc = g.U / L->grid;                       // ... (min)
  g.max_local_planes = (int)(upc / g.K + 2);
  g.scratch_off = g.U;
  g.keep_units = 0; g.grab = 1; g.max_jobs = 0;
  const size_t slack = 1024;                                     // static smem + alignment
  if (ring) {
    long long grab = upc / 8;
    g.grab = (int)(grab < 1 ? 1 : (grab > 4 ? 4 : grab));
    g.max_jobs = (int)(2 * upc + 4 * g.grab + 8);
    const size_t tables = align_up((size_t)g.max_jobs * (sizeof(float2) + sizeof(int)), 16);
    const size_t per_slot = (size_t)kUnitVecs * 16 + 16;
    long long s = (long long)(((size_t)di.max_smem_optin - slack - tables) / per_slot);
    if (s > kMaxSlots) s = kMaxSlots;
    if (s > upc) s = upc;
    if (s < 1) s = 1;
    g.slots = (int)s;
    L->smem = (size_t)g.slots * per_slot + tables;
    // L2 share reserved for phase-B re-reads (MRFP_NPPLUS_KEEP_MB overrides; default 64 MiB)
    static const long long keep_mb = getenv("MRFP_NPPLUS_KEEP_MB") ? atoll(getenv("MRFP_NPPLUS_KEEP_MB")) : 64;
    g.keep_units = (int)((keep_mb << 20) / ((long long)kUnitVecs * 16));
  } else {
    const size_t coef_bytes = align_up((size_t)g.max_local_planes * sizeof(float2), 16);
    const size_t unit_bytes = (size_t)g.Q * 4;
    long long s = (long long)(((size_t)di.max_smem_optin - slack - coef_bytes) / unit_bytes);
    if (s > min_upc) s = min_upc;   // the resident window must not start before u0
    g.slots = (int)s;
    L->smem = align_up((size_t)g.slots * unit_bytes, 16) + coef_bytes;
  }
}

unsigned long long next_nonce() {
  static std::atomic<unsigned long long> counter{0x9E3779B97F4A7C15ull};
  return counter.fetch_add(0x9E3779B97F4A7C15ull) | 1ull;       // never 0, never repeats within a process
}

template <bool BWD>
int run(const float* x, const float* alpha, const float* eps, const float* mean_in, float* out, float* mean_out,
        float* beta_out, void* ws, size_t ws_bytes, int N, int C, int HW, void* stream) {
  if (!x || !alpha || !eps || !out || !ws || (BWD && !mean_in) || (!BWD && !mean_out)) return MRFP_ERR_NULL_POINTER;
  if (N <= 0 || C <= 0 || HW <= 0 || (long long)N * C > (1 << 24)) return MRFP_ERR_BAD_SHAPE;
  if (ws_bytes < mrfp_npplus_ws_bytes(N, C, HW) || ((uintptr_t)ws & 15)) return MRFP_ERR_WORKSPACE;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  const bool ring = (HW % 4 == 0) && (((uintptr_t)x | (uintptr_t)out) & 15) == 0;
  NpLaunch L;
  plan_launch(N, C, HW, ring, di, &L);
  if (ring && L.g.U >= (1LL << 31)) return MRFP_ERR_BAD_SHAPE;
  NpGeom g = L.g;
  // debug: MRFP_NPPLUS_TRACE=1 and a workspace with room for grid*8 extra u64 -> per-CTA phase timestamps
  static const bool want_trace = getenv("MRFP_NPPLUS_TRACE") != nullptr;
  const size_t base = align_up(mrfp_npplus_ws_bytes(N, C, HW), 8);
  unsigned long long* trace = nullptr;
  if (want_trace && ws_bytes >= base + (size_t)L.grid * 64) trace = (unsigned long long*)((char*)ws + base);
  if (ring) {
    unsigned char* wsb = (unsigned char*)ws;
    unsigned long long nonce = next_nonce();
    void* args[] = {(void*)&x, (void*)&alpha, (void*)&eps, (void*)&mean_in, (void*)&out, (void*)&mean_out,
                    (void*)&beta_out, (void*)&wsb, (void*)&g, (void*)&nonce, (void*)&trace};
    void* kern = (void*)npplus_ring_kernel<BWD>;
    MRFP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    MRFP_CUDA_TRY(cudaLaunchCooperativeKernel(kern, dim3(L.grid), dim3(kConsumers + 32), args, L.smem, (cudaStream_t)stream));
  } else {
    double* ps = (double*)((char*)ws + sizeof(NpCtrl));
    void* args[] = {(void*)&x, (void*)&alpha, (void*)&eps, (void*)&mean_in, (void*)&out,
                    (void*)&mean_out, (void*)&beta_out, (void*)&ps, (void*)&g, (void*)&trace};
    void* kern = (void*)npplus_scalar_kernel<BWD>;
    MRFP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    MRFP_CUDA_TRY(cudaLaunchCooperativeKernel(kern, dim3(L.grid), dim3(kConsumers), args, L.smem, (cudaStream_t)stream));
  }
  return MRFP_OK;
}

}  // namespace
}  // namespace mrfp

extern "C" size_t mrfp_npplus_ws_bytes(int N, int C, int HW) {
  if (N <= 0 || C <= 0 || HW <= 0) return 0;
  return mrfp::ws_layout_bytes(N, C, HW);
}

extern "C" int mrfp_npplus_fwd_f32(const float* x, const float* alpha, const float* eps, float* out, float* mean,
                                   float* beta, void* ws, size_t ws_bytes, int N, int C, int HW, void* stream) {
  return mrfp::run<false>(x, alpha, eps, nullptr, out, mean, beta, ws, ws_bytes, N, C, HW, stream);
}

extern "C" int mrfp_npplus_bwd_f32(const float* gout, const float* alpha, const float* eps, const float* mean,
                                   float* gin, void* ws, size_t ws_bytes, int N, int C, int HW, void* stream) {
  return mrfp::run<true>(gout, alpha, eps, mean, gin, nullptr, nullptr, ws, ws_bytes, N, C, HW, stream);
}
