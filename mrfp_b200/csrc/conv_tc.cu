// tcgen05 implicit-GEMM 3x3 convolution for the HRFP chain (sm_100a), NHWC operands, fp32 accumulation in TMEM.
// Replaces the cuDNN fprop / dgrad calls behind nn.Conv2d at /root/reference/deepv3.py:320-327.
//
// Two element types share one kernel template:
//   __nv_bfloat16  tcgen05.mma kind::f16  (bf16 x bf16 -> fp32), 64 channels per k-step          (MRFP_MATH_BF16)
//   float          tcgen05.mma kind::tf32 (tf32 x tf32 -> fp32), 32 channels per k-step, fp32 in HBM  (MRFP_MATH_TF32:
//                  the arithmetic of the reference's own cuDNN convolutions under torch's TF32 default)
// In both a k-step is 128 bytes per pixel = one row of the 128-byte swizzle, so the tile geometry is identical.
//
// GEMM view per output tile:  D[128 pixels][COUT] = sum over (tap, channel chunk) A_tap[128][KB] * B_tap[COUT][KB]^T
//   * M tile = 8 rows x 16 cols of output pixels (UMMA M = 128, one TMEM lane per pixel)
//   * A operand: one 4-D TMA box {KB ch, 16, 8, 1} of the NHWC input at the tap's (dy,dx)*dilation offset —
//     TMA zero-fills the halo / out-of-image part, so padding costs nothing and no im2col buffer exists;
//     the box lands in shared memory as 128 rows x 128 B with the 128-byte swizzle = canonical K-major UMMA tile
//   * B operand: 2-D TMA box {KB, COUT} of the tap-major packed weights [9*COUT][CIN]
//   * accumulators: 2 TMEM stages x COUT fp32 columns (epilogue of tile i overlaps the MMAs of tile i+1)
// Warp roles (192 threads, persistent over tiles): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer
// (whole warp walks the loops, one elected lane issues), warps 2-5 = epilogue: tcgen05.ld -> pack into a swizzled
// staging tile -> TMA store (clips the ragged image edge) -> replication-count-weighted per-channel column sums of the
// stored tile (BN batch statistics of the resampled tensor); the last CTA finalises the statistics.
#include "hrfp.cuh"
#include <atomic>
#include <mutex>

namespace mrfp {
namespace {

constexpr int kThreads = 192;
constexpr int kATileBytes = 128 * 128;          // 128 pixels x 128 B (64 bf16 / 32 fp32 channels)
constexpr int kStageOutBytes = 128 * 128;       // 128 pixels x 128 B

template <typename T> struct Elem;
template <> struct Elem<__nv_bfloat16> {
  static constexpr int kBlockK = 64;            // channels per k-step = 128 bytes = one swizzle row
  static constexpr uint32_t kFmt = 1;           // UMMA instruction-descriptor operand format: BF16
  static constexpr CUtensorMapDataType kLoadType = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  static constexpr CUtensorMapDataType kStoreType = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
};
template <> struct Elem<float> {
  static constexpr int kBlockK = 32;
  static constexpr uint32_t kFmt = 2;           // TF32
  static constexpr CUtensorMapDataType kLoadType = CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;   // fp32 in HBM, tf32 on the way in
  static constexpr CUtensorMapDataType kStoreType = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
};

// MT = M sub-tiles (of 128 pixels, stacked vertically) per CTA tile.  With MT = 2 one B (weight) tile feeds two
// MMAs, halving the weight traffic per FLOP: the 64/128-wide layers are bound by the L2->SM operand stream
// (~100 B/clk/SM), not by the tensor pipe.  COUT = 256 already fills TMEM (2 x 256 columns) with MT = 1.
template <int COUT> struct Cfg {
  static constexpr int kMT = COUT == 256 ? 1 : 2;
  static constexpr int kAStageBytes = kMT * kATileBytes;
  static constexpr int kBTileBytes = COUT * 128;
  static constexpr int kStages = 4;
  static constexpr int kOutBufs = 2;                                // staging tiles: chunk i+1 is packed while chunk i drains
  static constexpr int kTmemCols = 2 * kMT * COUT;                  // 256 / 512 / 512: powers of two
  static constexpr int kSmemBytes = kStages * (kAStageBytes + kBTileBytes) + kOutBufs * kStageOutBytes +
                                    512 /* row weights */ + 256 /* barriers */ + 1024 /* alignment slack */;
};

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
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T with fp32 accumulation, issued by one thread for the CTA
template <typename T>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if constexpr (sizeof(T) == 2) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
  }
}
// K-major, 128-byte swizzle, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ int g_conv_dbg = 0;   // TEMP ablation flags
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int COUT, typename T>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_w,
                  const __grid_constant__ CUtensorMap tmap_out, int CIN, int dil, int tiles_h, int tiles_w,
                  int num_tiles, const int* __restrict__ cnt_h, const int* __restrict__ cnt_w,
                  double* __restrict__ stat_acc, int rev, const ConvBnFinalize fin, const T* __restrict__ add_src, int H, int W) {
  using C = Cfg<COUT>;
  using E = Elem<T>;
  constexpr int kBlockK = E::kBlockK;
  constexpr int kChunkC = 128 / (int)sizeof(T);                // channels of one 128-byte output chunk
  constexpr int kChunks = COUT / kChunkC;                      // chunks per M sub-tile
  constexpr bool kBf16 = sizeof(T) == 2;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;
  unsigned char* sB = sA + C::kStages * C::kAStageBytes;
  unsigned char* sOut = sB + C::kStages * C::kBTileBytes;
  float* s_wgt = reinterpret_cast<float*>(sOut + C::kOutBufs * kStageOutBytes);   // [128] replication count of each tile row's pixel
  uint64_t* full = reinterpret_cast<uint64_t*>(s_wgt + 128);
  uint64_t* empty = full + C::kStages;
  uint64_t* tmem_full = empty + C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = 9 * (CIN / kBlockK);       // k-steps per tile

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_in)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_out)) : "memory");
    for (int i = 0; i < C::kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(C::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;
  pdl_sync();                                 // set-up above overlaps the previous kernel's tail
  const int dbg = g_conv_dbg;

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks, one elected lane issues) =====================
    int stage = 0; uint32_t phase = 0;
    for (int t0 = blockIdx.x; t0 < num_tiles; t0 += gridDim.x) {
      const int t = rev ? num_tiles - 1 - t0 : t0;     // rev: walk the image from its end (where the producer of `in` finished)
      const int tw = t % tiles_w, th = (t / tiles_w) % tiles_h, n = t / (tiles_w * tiles_h);
      const int h0 = th * kTileH * C::kMT, w0 = tw * kTileW;
      for (int tap = 0; tap < 9; ++tap) {
        const int dy = (tap / 3 - 1) * dil, dx = (tap % 3 - 1) * dil;
        for (int kc = 0; kc < CIN / kBlockK; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&full[stage], ((dbg & 8) ? 0 : C::kAStageBytes) + ((dbg & 16) ? 0 : C::kBTileBytes));
            if (!(dbg & 8)) tma_load_4d(sA + stage * C::kAStageBytes, &tmap_in, &full[stage], kc * kBlockK, w0 + dx, h0 + dy, n);
            if (!(dbg & 16)) tma_load_2d(sB + stage * C::kBTileBytes, &tmap_w, &full[stage], kc * kBlockK, tap * COUT);
          }
          __syncwarp();
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    // all loads of this CTA are in flight: let the next kernel of the stream start its set-up (PDL)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loops with warp-uniform values (descriptors stay in uniform registers); one elected
    // lane issues the tcgen05 instructions.
    // instruction descriptor: D=f32, A=B=bf16 / tf32, both K-major, N=COUT, M=128
    constexpr uint32_t idesc = (1u << 4) | (E::kFmt << 7) | (E::kFmt << 10) | ((uint32_t)(COUT >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_tmem = tmem_u + (uint32_t)(acc * C::kMT * COUT);
      for (int ks = 0; ks < nk; ++ks) {
        mbar_wait(&full[stage], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint64_t da = make_desc_sw128(sA_u + stage * C::kAStageBytes);
          const uint64_t db = make_desc_sw128(sB_u + stage * C::kBTileBytes);
          if (!(dbg & 4))
#pragma unroll
          for (int k = 0; k < 4; ++k)                  // one UMMA_K = 32 bytes of the swizzle row (16 bf16 / 8 tf32)
#pragma unroll
            for (int mt = 0; mt < C::kMT; ++mt)        // the second M sub-tile is the next 128 rows (16 KiB) of the A box
              umma<T>(d_tmem + (uint32_t)(mt * COUT), da + (uint64_t)(mt * (kATileBytes >> 4) + k * 2),
                      db + (uint64_t)(k * 2), idesc, (ks | k) != 0);
          umma_commit(&empty[stage]);                  // frees the smem slot when the MMAs have read it
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tmem_full[acc]);   // accumulator complete
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;                // accumulator row = pixel inside the tile
    const int hl = r / kTileW, wl = r % kTileW;
    const int et = threadIdx.x - 64;            // 0..127
    const bool leader = et == 0;                // first epilogue thread issues the TMA stores
    // statistics: thread (word cp of a 128-byte row, row group pg) sums 32 rows of the staging tile (the values that are
    // stored and later normalised), weighted by the replication count of each row's pixel; a word is a channel pair
    // (bf16) or one channel (fp32)
    const int cp = et & 31, pg = et >> 5;
    float a1x[kChunks], a1y[kChunks], a2x[kChunks], a2y[kChunks];      // per 128-byte chunk of the output row (static indices)
#pragma unroll
    for (int j = 0; j < kChunks; ++j) a1x[j] = a1y[j] = a2x[j] = a2y[j] = 0.f;
    // add_src: this thread's 128 bytes of its pixel, fetched ONE CHUNK AHEAD so the loads overlap the previous chunk
    uint4 ad_nxt[8], ad_cur[8];
    bool ad_nxt_ok = false, ad_ok = false;
    auto add_fetch = [&](int t0_, int jj_) {
      ad_nxt_ok = false;
      if (add_src == nullptr || t0_ >= num_tiles) return;
      const int t_ = rev ? num_tiles - 1 - t0_ : t0_;
      const int tw_ = t_ % tiles_w, th_ = (t_ / tiles_w) % tiles_h, n_ = t_ / (tiles_w * tiles_h);
      const int mt_ = jj_ / kChunks, j_ = jj_ % kChunks;
      const int hh = th_ * kTileH * C::kMT + mt_ * kTileH + hl, ww = tw_ * kTileW + wl;
      if (hh < H && ww < W) {
        const uint4* ap = reinterpret_cast<const uint4*>(add_src + (((size_t)n_ * H + hh) * W + ww) * COUT + j_ * kChunkC);
#pragma unroll
        for (int c = 0; c < 8; ++c) ad_nxt[c] = __ldg(ap + c);
        ad_nxt_ok = true;
      }
    };
    add_fetch(blockIdx.x, 0);
    int it = 0;
    for (int t0 = blockIdx.x; t0 < num_tiles; t0 += gridDim.x, ++it) {
      const int t = rev ? num_tiles - 1 - t0 : t0;
      const int tw = t % tiles_w, th = (t / tiles_w) % tiles_h, n = t / (tiles_w * tiles_h);
      const int h0 = th * kTileH * C::kMT, w0 = tw * kTileW;
      const int acc = it & 1;
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (dbg & 2) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        continue;
      }
#pragma unroll
      for (int jj = 0; jj < C::kMT * kChunks; ++jj) {
        const int mt = jj / kChunks, j = jj % kChunks;
        if (add_src != nullptr) {                // rotate the prefetch: this chunk's data, then start the next chunk's loads
#pragma unroll
          for (int c = 0; c < 8; ++c) ad_cur[c] = ad_nxt[c];
          ad_ok = ad_nxt_ok;
          if (jj + 1 < C::kMT * kChunks) add_fetch(t0, jj + 1); else add_fetch(t0 + gridDim.x, 0);
        }
        unsigned char* ob = sOut + ((C::kMT * kChunks) % 2 == 0 ? (jj & 1) : ((it * C::kMT * kChunks + jj) & 1)) * kStageOutBytes;
        // the TMA store that last read this staging buffer (two chunks ago) must have drained
        if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        epi_bar_sync();
        if (stat_acc) s_wgt[r] = (float)(cnt_h[h0 + mt * kTileH + hl] * cnt_w[w0 + wl]);   // 0 outside the image (zero-padded tables)
        const uint32_t tcol = (uint32_t)((acc * C::kMT + mt) * COUT + j * kChunkC);
        if constexpr (kBf16) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + tcol + (uint32_t)(half * 32), v);
            if (ad_ok) {                           // summed in fp32, rounded once
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const uint4 a4 = ad_cur[half * 4 + c];
                const uint32_t w4[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  v[c * 8 + 2 * i] = __float_as_uint(__uint_as_float(v[c * 8 + 2 * i]) + __uint_as_float(w4[i] << 16));
                  v[c * 8 + 2 * i + 1] = __float_as_uint(__uint_as_float(v[c * 8 + 2 * i + 1]) + __uint_as_float(w4[i] & 0xffff0000u));
                }
              }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {          // four 16-byte chunks (8 channels each) per half
              uint32_t p[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[c * 8 + 2 * i]), __uint_as_float(v[c * 8 + 2 * i + 1]));
                p[i] = *reinterpret_cast<const uint32_t*>(&h2);
              }
              const int chunk = half * 4 + c;
              *reinterpret_cast<uint4*>(ob + r * 128 + ((chunk ^ (r & 7)) << 4)) = make_uint4(p[0], p[1], p[2], p[3]);
            }
          }
        } else {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + tcol, v);
#pragma unroll
          for (int c = 0; c < 8; ++c) {            // eight 16-byte chunks (4 channels each)
            uint4 o = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            if (ad_ok) {
              const uint4 a4 = ad_cur[c];
              o.x = __float_as_uint(__uint_as_float(o.x) + __uint_as_float(a4.x));
              o.y = __float_as_uint(__uint_as_float(o.y) + __uint_as_float(a4.y));
              o.z = __float_as_uint(__uint_as_float(o.z) + __uint_as_float(a4.z));
              o.w = __float_as_uint(__uint_as_float(o.w) + __uint_as_float(a4.w));
            }
            *reinterpret_cast<uint4*>(ob + r * 128 + ((c ^ (r & 7)) << 4)) = o;
          }
        }
        if (jj == C::kMT * kChunks - 1) {       // all TMEM reads of this accumulator are done
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        epi_bar_sync();
        if (leader) {
          tma_store_4d(&tmap_out, ob, j * kChunkC, w0, h0 + mt * kTileH, n);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (stat_acc) {
          // word cp lives in 16-byte chunk cp/4 of a row, position cp%4; 32 lanes read one whole (swizzled) row
          const unsigned char* col = ob + (cp & 3) * 4;
          const int ch = cp >> 2;
          float s1x = 0.f, s1y = 0.f, s2x = 0.f, s2y = 0.f;
#pragma unroll 8
          for (int i = 0; i < 32; ++i) {
            const int row = pg * 32 + i;
            const uint32_t w2 = *reinterpret_cast<const uint32_t*>(col + row * 128 + ((ch ^ (row & 7)) << 4));
            const float wg = s_wgt[row];
            if constexpr (kBf16) {
              const float y0 = __uint_as_float(w2 << 16), y1 = __uint_as_float(w2 & 0xffff0000u);
              const float t0 = wg * y0, t1 = wg * y1;
              s1x += t0; s1y += t1;
              s2x = fmaf(t0, y0, s2x); s2y = fmaf(t1, y1, s2y);
            } else {
              const float y0 = __uint_as_float(w2);
              const float t0 = wg * y0;
              s1x += t0;
              s2x = fmaf(t0, y0, s2x);
            }
          }
          a1x[j] += s1x; a2x[j] += s2x;
          if constexpr (kBf16) { a1y[j] += s1y; a2y[j] += s2y; }
        }
      }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (stat_acc) {
      // the four row groups meet once per kernel, in the first pipeline slot: every load of this CTA has been consumed and
      // every MMA has retired (the last tmem_full), so the operand ring is idle
      float* part = reinterpret_cast<float*>(sA) + pg * 2 * COUT;
#pragma unroll
      for (int j = 0; j < kChunks; ++j) {
        if constexpr (kBf16) {
          part[j * 64 + 2 * cp] = a1x[j]; part[j * 64 + 2 * cp + 1] = a1y[j];
          part[COUT + j * 64 + 2 * cp] = a2x[j]; part[COUT + j * 64 + 2 * cp + 1] = a2y[j];
        } else {
          part[j * 32 + cp] = a1x[j];
          part[COUT + j * 32 + cp] = a2x[j];
        }
      }
      epi_bar_sync();
      const float* all = reinterpret_cast<const float*>(sA);
      for (int c = et; c < 2 * COUT; c += 128) {
        const float sum = (all[c] + all[2 * COUT + c]) + (all[4 * COUT + c] + all[6 * COUT + c]);
        atomicAdd(stat_acc + (c < COUT ? c : kMaxC + c - COUT), (double)sum);
      }
    }
    if (stat_acc && fin.stats) {
      // BN finalisation by the last CTA to arrive (its adds and everybody else's are visible behind the fences)
      __threadfence();
      epi_bar_sync();
      if (leader) s_wgt[0] = (atomicAdd(fin.counter, 1u) == gridDim.x - 1) ? 1.f : 0.f;
      epi_bar_sync();
      if (s_wgt[0] != 0.f) {
        __threadfence();
        for (int c = et; c < COUT; c += 128) {
          const double mean = __ldcg(stat_acc + c) / fin.count;
          double var = __ldcg(stat_acc + kMaxC + c) / fin.count - mean * mean;
          if (var < 0) var = 0;
          const double invstd = 1.0 / sqrt(var + (double)fin.eps);
          const bool live = c < fin.cout_real;     // padded output channels (zero weights) are pinned to zero
          const float sc = live ? (float)((double)fin.gamma[c] * invstd) : 0.f;
          const float b = (live && fin.beta) ? fin.beta[c] : 0.f;
          fin.stats[0 * kMaxC + c] = (float)mean;
          fin.stats[1 * kMaxC + c] = (float)invstd;
          fin.stats[2 * kMaxC + c] = sc;
          fin.stats[3 * kMaxC + c] = (float)((double)b - mean * (double)sc);
          if (live && fin.running_mean) fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * (float)mean;
          if (live && fin.running_var)
            fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] +
                                 fin.momentum * (float)(var * (fin.count / (fin.count - 1.0)));
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------
// host side: tensor maps (driver entry point fetched through the runtime; no link against libcuda)
// ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_map(CUtensorMap* m, CUtensorMapDataType dt, const void* base, int rank, const cuuint64_t* dims,
             const cuuint64_t* strides, const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return MRFP_ERR_DRIVER;
  const cuuint32_t ones[4] = {1, 1, 1, 1};
  CUresult r = fn(m, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MRFP_OK : MRFP_ERR_DRIVER;
}

template <int COUT, typename T>
int launch(const T* in, const T* wpack, T* out, int N, int H, int W, int cin, int dil, const int* cnt_h, const int* cnt_w,
           double* stat_acc, int rev, const ConvBnFinalize& fin, const T* add_src, ConvMaps* cache, cudaStream_t stream) {
  using E = Elem<T>;
  constexpr int es = (int)sizeof(T);
  constexpr int kChunkC = 128 / es;
  // the three tensor maps depend on the buffer addresses and the (plan-constant) geometry only: a plan keeps them per
  // (stage, direction) and re-encodes when an address changes (the caching allocator hands the same blocks back)
  ConvMaps local;
  local.valid = 0;
  ConvMaps* m = cache ? cache : &local;
  if (!m->valid || m->key[0] != in || m->key[1] != wpack || m->key[2] != out) {
    m->valid = 0;
    {
      const cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
      const cuuint64_t strides[3] = {(cuuint64_t)cin * es, (cuuint64_t)W * cin * es, (cuuint64_t)H * W * cin * es};
      const cuuint32_t box[4] = {(cuuint32_t)E::kBlockK, kTileW, (cuuint32_t)(kTileH * Cfg<COUT>::kMT), 1};
      int rc = make_map(&m->in, E::kLoadType, in, 4, dims, strides, box);
      if (rc) return rc;
    }
    {
      const cuuint64_t dims[2] = {(cuuint64_t)cin, (cuuint64_t)9 * COUT};
      const cuuint64_t strides[1] = {(cuuint64_t)cin * es};
      const cuuint32_t box[2] = {(cuuint32_t)E::kBlockK, COUT};
      int rc = make_map(&m->w, E::kLoadType, wpack, 2, dims, strides, box);
      if (rc) return rc;
    }
    {
      const cuuint64_t dims[4] = {(cuuint64_t)COUT, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
      const cuuint64_t strides[3] = {(cuuint64_t)COUT * es, (cuuint64_t)W * COUT * es, (cuuint64_t)H * W * COUT * es};
      const cuuint32_t box[4] = {(cuuint32_t)kChunkC, kTileW, kTileH, 1};
      int rc = make_map(&m->out, E::kStoreType, out, 4, dims, strides, box);
      if (rc) return rc;
    }
    m->key[0] = in; m->key[1] = wpack; m->key[2] = out;
    m->valid = 1;
  }
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  const int th_px = kTileH * Cfg<COUT>::kMT;
  const int tiles_h = (H + th_px - 1) / th_px, tiles_w = (W + kTileW - 1) / kTileW;
  const int num_tiles = N * tiles_h * tiles_w;
  const int grid = num_tiles < di.sm_count ? num_tiles : di.sm_count;
  auto kern = conv3x3_tc_kernel<COUT, T>;
  static std::atomic<unsigned long long> attr_done{0};           // per device, once: the opt-in shared-memory size
  const unsigned long long bit = 1ull << (di.device & 63);
  if (!(attr_done.load(std::memory_order_acquire) & bit)) {
    MRFP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<COUT>::kSmemBytes));
    attr_done.fetch_or(bit, std::memory_order_release);
  }
  launch_k(kern, dim3(grid), dim3(kThreads), Cfg<COUT>::kSmemBytes, stream, m->in, m->w, m->out, cin, dil, tiles_h, tiles_w,
           num_tiles, cnt_h, cnt_w, stat_acc, rev, fin, add_src, H, W);
  MRFP_CUDA_TRY(cudaGetLastError());
  return MRFP_OK;
}

template <typename T>
int dispatch(const void* in, const void* wpack, void* out, int N, int H, int W, int cin, int cout, int dil, const int* cnt_h,
             const int* cnt_w, double* stat_acc, int rev, const ConvBnFinalize& fin, const void* add_src, ConvMaps* cache,
             cudaStream_t stream) {
  const T* i = static_cast<const T*>(in); const T* w = static_cast<const T*>(wpack); const T* a = static_cast<const T*>(add_src);
  T* o = static_cast<T*>(out);
  switch (cout) {
    case 64: return launch<64, T>(i, w, o, N, H, W, cin, dil, cnt_h, cnt_w, stat_acc, rev, fin, a, cache, stream);
    case 128: return launch<128, T>(i, w, o, N, H, W, cin, dil, cnt_h, cnt_w, stat_acc, rev, fin, a, cache, stream);
    case 256: return launch<256, T>(i, w, o, N, H, W, cin, dil, cnt_h, cnt_w, stat_acc, rev, fin, a, cache, stream);
  }
  return MRFP_ERR_UNSUPPORTED;
}

}  // namespace

bool conv3x3_tc_supported(int cin, int cout, int esize) {
  if (esize != 2 && esize != 4) return false;
  const int kb = 128 / esize;
  return cin % kb == 0 && cin <= kMaxC && (cout == 64 || cout == 128 || cout == 256);
}

int conv3x3_tc(const void* in, const void* wpack, void* out, int esize, int N, int H, int W, int cin, int cout, int dil,
               const int* cnt_h, const int* cnt_w, double* stat_acc, cudaStream_t stream, bool reverse_tiles,
               const ConvBnFinalize* finalize, const void* add_src, ConvMaps* cache) {
  ConvBnFinalize fin = {};
  if (finalize) {
    if (!stat_acc || !finalize->gamma || !finalize->stats || !finalize->counter) return MRFP_ERR_NULL_POINTER;
    fin = *finalize;
    if (fin.cout_real <= 0 || fin.cout_real > cout) fin.cout_real = cout;
  }
  const int rev = reverse_tiles ? 1 : 0;
  if (!conv3x3_tc_supported(cin, cout, esize)) return MRFP_ERR_UNSUPPORTED;
  if (((uintptr_t)in | (uintptr_t)wpack | (uintptr_t)out | (uintptr_t)add_src) & 15) return MRFP_ERR_WORKSPACE;
  if (esize == 2)
    return dispatch<__nv_bfloat16>(in, wpack, out, N, H, W, cin, cout, dil, cnt_h, cnt_w, stat_acc, rev, fin, add_src, cache, stream);
  return dispatch<float>(in, wpack, out, N, H, W, cin, cout, dil, cnt_h, cnt_w, stat_acc, rev, fin, add_src, cache, stream);
}

}  // namespace mrfp

// test / bench hooks (not part of the public header): one tcgen05 convolution on caller-provided NHWC buffers —
// the same kernel the chain launches.  bf16: in / wpack / out are bf16; tf32: fp32.
extern "C" int mrfp_debug_conv3x3_bf16(const void* in, const void* wpack, void* out, int N, int H, int W, int cin,
                                       int cout, int dil, const int* cnt_h, const int* cnt_w, double* stat_acc,
                                       void* stream) {
  return mrfp::conv3x3_tc(in, wpack, out, 2, N, H, W, cin, cout, dil, cnt_h, cnt_w, stat_acc, (cudaStream_t)stream, false,
                          nullptr, nullptr, nullptr);
}
extern "C" int mrfp_debug_conv_set(int v) { return (int)cudaMemcpyToSymbol(mrfp::g_conv_dbg, &v, sizeof(int)); }
extern "C" int mrfp_debug_conv3x3_tf32(const void* in, const void* wpack, void* out, int N, int H, int W, int cin,
                                       int cout, int dil, const int* cnt_h, const int* cnt_w, double* stat_acc,
                                       void* stream) {
  return mrfp::conv3x3_tc(in, wpack, out, 4, N, H, W, cin, cout, dil, cnt_h, cnt_w, stat_acc, (cudaStream_t)stream, false,
                          nullptr, nullptr, nullptr);
}
