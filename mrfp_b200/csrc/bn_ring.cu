// TMA-fed variants of the BN-backward passes of the HRFP chain (bf16 NHWC), sm_100a.
//
// The LDG-based row kernels in hrfp.cu plateau near 4.8 TB/s of HBM reads (ncu: warps on the long scoreboard, HBM
// channels 32-63 % busy), while the bulk-copy ring of the NP+ kernel streams at 6 TB/s.  Same structure here: one
// persistent CTA per SM, a producer warp that turns work items into 1-D bulk copies (cp.async.bulk + mbarrier
// complete_tx) into a shared-memory ring, 16 consumer warps that compute from shared memory.  An item is one
// 16 KiB segment of a gradient row plus the span of the saved conv-output row that its pixels gather from.
#include "hrfp.cuh"
#include "tma.cuh"

namespace mrfp {
namespace {

using namespace tma;

constexpr int kRConsumers = 512;
constexpr int kRWarps = kRConsumers / 32;
constexpr int kRSlots = 5;
constexpr int kDaBytes = 16384;                 // gradient segment
constexpr int kYBytes = 24576;                  // gathered span of y (up to 1.5x the segment: down-sampling stages)
constexpr int kSlotBytes = kDaBytes + kYBytes;

struct RingMeta { int ow0, npx, s0, pad; };

__device__ __forceinline__ void unpack8(const uint4 r, float (&v)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// U1[c] = sum mask*dA, U2[c] = sum mask*dA*y over all destination pixels; mask = [scale*y + shift > 0], y gathered.
__global__ void __launch_bounds__(kRConsumers + 32, 1)
bn_bwd_reduce_ring_kernel(const __nv_bfloat16* __restrict__ dA, const __nv_bfloat16* __restrict__ y,
                          const int* __restrict__ idx_h, const int* __restrict__ idx_w, const float* __restrict__ stats,
                          double* __restrict__ acc, int N, int C, int IH, int IW, int OH, int OW, int pseg, int nseg,
                          int items, int rev, float scale_w) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* ring = smem_raw;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kRSlots * kSlotBytes);
  uint64_t* empty = full + kRSlots;
  RingMeta* meta = reinterpret_cast<RingMeta*>(empty + kRSlots);
  float* s_acc = reinterpret_cast<float*>(meta + kRSlots);        // [2 * kMaxC]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int s = 0; s < kRSlots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kRWarps); }
    mbar_fence_init();
  }
  for (int i = tid; i < 2 * kMaxC; i += kRConsumers + 32) s_acc[i] = 0.f;
  __syncthreads();
  pdl_sync();

  if (warp == kRWarps) {
    // ---------------- producer warp: 32 item descriptors at a time, lane 0 issues ----------------
    int s = 0, k = 0;
    const uint64_t pol = policy_evict_first();
    for (int base = blockIdx.x; base < items; base += 32 * gridDim.x) {
      const int it = base + lane * gridDim.x;
      const bool valid = it < items;
      int ow0 = 0, npx = 0, s0 = 0, ybytes = 0;
      long long da_off = 0, y_off = 0;
      if (valid) {
        const int itr = rev ? items - 1 - it : it;
        const int row = itr / nseg, seg = itr - row * nseg;
        const int n = row / OH, oh = row - n * OH;
        ow0 = seg * pseg;
        npx = min(pseg, OW - ow0);
        s0 = idx_w[ow0];
        const int s1 = idx_w[ow0 + npx - 1];
        ybytes = (s1 - s0 + 1) * C * 2;
        da_off = ((long long)row * OW + ow0) * C;
        y_off = (((long long)n * IH + idx_h[oh]) * IW + s0) * C;
      }
      const int cnt = __popc(__ballot_sync(0xffffffffu, valid));   // valid lanes are a prefix
      for (int q = 0; q < cnt; ++q) {
        const int q_ow0 = __shfl_sync(0xffffffffu, ow0, q), q_npx = __shfl_sync(0xffffffffu, npx, q);
        const int q_s0 = __shfl_sync(0xffffffffu, s0, q), q_yb = __shfl_sync(0xffffffffu, ybytes, q);
        const long long q_da = __shfl_sync(0xffffffffu, da_off, q), q_y = __shfl_sync(0xffffffffu, y_off, q);
        if (lane == 0) {
          if (k > 0) mbar_wait(&empty[s], (k - 1) & 1);
          meta[s] = RingMeta{q_ow0, q_npx, q_s0, 0};
          const uint32_t dab = (uint32_t)(q_npx * C * 2);
          unsigned char* slot = ring + (size_t)s * kSlotBytes;
          mbar_expect_tx(&full[s], dab + (uint32_t)q_yb);
          bulk_load_hint(slot, dA + q_da, dab, &full[s], pol);
          bulk_load_hint(slot + kDaBytes, y + q_y, (uint32_t)q_yb, &full[s], pol);
        }
        if (++s == kRSlots) { s = 0; ++k; }
      }
    }
  } else {
    // ---------------- consumers: thread = fixed 8-channel group, pixels strided ----------------
    const int cg = C >> 3, cgs = 31 - __clz(cg);          // channel groups per pixel (power of two)
    const int c = (tid & (cg - 1)) << 3;
    float scale[8], shift[8], u1[8], u2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      scale[j] = stats[2 * kMaxC + c + j]; shift[j] = stats[3 * kMaxC + c + j];
      u1[j] = 0.f; u2[j] = 0.f;
    }
    int s = 0, k = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      mbar_wait(&full[s], k & 1);
      const RingMeta m = meta[s];
      const unsigned char* slot = ring + (size_t)s * kSlotBytes;
      const int nel = m.npx << cgs;
      for (int e = tid; e < nel; e += kRConsumers) {
        const int px = e >> cgs;
        // ATen's nearest rule, bit-identical to the idx_w table (single fp32 multiply, floor, clamp)
        const int j = min((int)floorf(__fmul_rn((float)(m.ow0 + px), scale_w)), IW - 1) - m.s0;
        float g[8], yv[8];
        unpack8(*reinterpret_cast<const uint4*>(slot + ((size_t)px * C + c) * 2), g);
        unpack8(*reinterpret_cast<const uint4*>(slot + kDaBytes + ((size_t)j * C + c) * 2), yv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float t = fmaf(scale[i], yv[i], shift[i]) > 0.f ? g[i] : 0.f;
          u1[i] += t;
          u2[i] = fmaf(t, yv[i], u2[i]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == kRSlots) { s = 0; ++k; }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&s_acc[c + j], u1[j]);
      atomicAdd(&s_acc[kMaxC + c + j], u2[j]);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kRConsumers) : "memory");
    for (int i = tid; i < C; i += kRConsumers) {
      atomicAdd(acc + i, (double)s_acc[i]);
      atomicAdd(acc + kMaxC + i, (double)s_acc[kMaxC + i]);
    }
  }
}

}  // namespace

// host_idx_w: the plan's host copy of idx_w (span check).  Returns MRFP_ERR_UNSUPPORTED when the geometry does not fit
// the ring (the caller then uses the LDG kernel).
int bn_bwd_reduce_ring(const __nv_bfloat16* dA, const __nv_bfloat16* y, const int* idx_h, const int* idx_w,
                       const int* host_idx_w, float scale_w, const float* stats, double* acc, int N, int C, int IH, int IW,
                       int OH, int OW, bool reverse, cudaStream_t stream) {
  if (C < 64 || C > kMaxC || (C & (C - 1))) return MRFP_ERR_UNSUPPORTED;
  const int pseg = kDaBytes / (2 * C);
  const int nseg = (OW + pseg - 1) / pseg;
  for (int sgi = 0; sgi < nseg; ++sgi) {
    const int ow0 = sgi * pseg, ow1 = (ow0 + pseg < OW ? ow0 + pseg : OW) - 1;
    if ((host_idx_w[ow1] - host_idx_w[ow0] + 1) * C * 2 > kYBytes) return MRFP_ERR_UNSUPPORTED;
  }
  const long long items = (long long)N * OH * nseg;
  if (items <= 0 || items > 0x3fffffff) return MRFP_ERR_UNSUPPORTED;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  const size_t smem = (size_t)kRSlots * kSlotBytes + 2 * kRSlots * 8 + kRSlots * sizeof(RingMeta) + 2 * kMaxC * 4;
  if (smem > (size_t)di.max_smem_optin) return MRFP_ERR_UNSUPPORTED;
  MRFP_CUDA_TRY(cudaFuncSetAttribute(bn_bwd_reduce_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = items < di.sm_count ? (int)items : di.sm_count;
  launch_k(bn_bwd_reduce_ring_kernel, dim3(grid), dim3(kRConsumers + 32), smem, stream, dA, y, idx_h, idx_w, stats, acc, N, C,
           IH, IW, OH, OW, pseg, nseg, (int)items, reverse ? 1 : 0, scale_w);
  MRFP_CUDA_TRY(cudaGetLastError());
  return MRFP_OK;
}

}  // namespace mrfp
