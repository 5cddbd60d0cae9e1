// Internal declarations shared by hrfp.cu (plan, element-wise kernels, orchestration), bn_ring.cu (bulk-copy row
// passes) and conv_tc.cu (tcgen05 implicit-GEMM convolution).
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <mutex>
#include <vector>

namespace mrfp {

constexpr int kHrfpStages = 8;
constexpr int kMaxC = 256;
constexpr int kTileH = 8;    // tcgen05 conv output tile: 8 rows x 16 cols = 128 pixels = UMMA M
constexpr int kTileW = 16;

struct HrfpStage {
  int cin, cout, dil;       // channel counts as stored (padded to the tensor-core granule for a narrow stem)
  int cin_real, cout_real;  // channel counts of the module's conv (== cin, cout unless padded)
  int ch, cw;      // conv resolution (input and output of the 3x3 conv)
  int oh, ow;      // resolution after the nearest resample
  int max_rep;     // largest replication count of a source row / column
  float scale_h, scale_w;   // ATen's nearest rule: src = min(floorf(dst * scale), in - 1)  (== the idx tables)
  // offsets (in ints) into the LUT blob
  int idx_h, idx_w;       // dst -> src index, [oh], [ow]
  int cnt_h, cnt_w;       // replication count of each src row/col, zero-padded to a tile multiple (+1 tile)
  int start_h, start_w;   // first dst index of each src row/col, [ch], [cw]
  int lo_h, lo_w;         // first dst index whose source is >= s (monotone), [ch + 1], [cw + 1]: the replicas of source s
                          // are dst [lo[s], lo[s + 1]); a source segment [a, b) gathers from the contiguous dst span [lo[a], lo[b])
  size_t y_off;           // conv output Y_k in `saved` (bytes)
  size_t wf_off;          // packed forward weights in ws (bytes)
  size_t wb_off;          // packed dgrad weights in `saved` (bytes)
};

// the three TMA descriptors of one convolution launch; a plan keeps one per (stage, direction) and re-encodes only
// when a buffer address changes
struct ConvMaps {
  CUtensorMap in, w, out, aux;   // aux: the second operand of a rank-K term (conv_gather.cu)
  const void* key[3];
  const void* key_aux;
  int valid;
};
// the two descriptors of one classifier-tail launch (tail_final2.cu): Y_3 as [pixel][256 ch], and the fp32 side input
// (forward: the low-resolution product, backward: the incoming gradient) as [plane][row][col]
struct TailMaps {
  CUtensorMap y, aux;
  const void* key[2];
  int k, d0, d1;
  int valid;
};

}  // namespace mrfp

struct mrfp_hrfp_plan {
  uint32_t magic;
  int N, cin, xh, xw, h, w, mode, esize;
  int cin_pad;                     // stem channels as stored inside the chain
  mrfp::HrfpStage st[mrfp::kHrfpStages];
  std::vector<int> lut;
  size_t ws_bytes, saved_bytes;
  // ws layout
  size_t acc_fwd_off, acc_bwd_off, bufs_off, buf_a_bytes, buf_g_bytes, buf_dy_bytes;
  // saved layout
  size_t stats_off;   // 8 x 4 x kMaxC floats: mean, invstd, scale, shift
  // launch-side cache (not part of the geometry): tensor maps per (direction, stage)
  mutable std::mutex mu;
  mutable mrfp::ConvMaps maps[2][mrfp::kHrfpStages];
  mutable mrfp::ConvMaps maps_g[2][mrfp::kHrfpStages];   // the halo-tile variants (conv_gather.cu)
  mutable mrfp::TailMaps maps_tail[2];                   // classifier tail, forward / backward
  int fuse;                        // bf16 mode: bit 0 forward convs build their operand from Y_{k-1} on chip, bit 1 dgrads of
                                   // non-replicating stages build dY_k on chip
};

namespace mrfp {
constexpr uint32_t kPlanMagic = 0x4d524650u;   // 'MRFP'

// tcgen05 implicit-GEMM 3x3 convolution, NHWC in/out, fp32 accumulation in TMEM (conv_tc.cu).
//   esize 2: bf16 operands (kind::f16);  esize 4: fp32 storage, tf32 operands (kind::tf32)
//   in  [N][H][W][cin], wpack [9][cout][cin] (tap-major, K contiguous), out [N][H][W][cout]
//   cnt_h / cnt_w: zero-padded replication counts (device) -> per-channel weighted sum / sum of squares of the
//   stored conv output are added to stat_acc[0..cout) / stat_acc[kMaxC..kMaxC+cout); pass nullptr to skip.
//   add_src (dgrad use): a tensor of the output's shape that is added to the accumulators in fp32 before the single
//   rounding to the storage type (the gradient of OCout_dec joining dA_3).
//   finalize (forward use, with stat_acc): the LAST CTA to add its partial statistics turns them into the BN
//   mean / invstd / scale / shift table and updates the running statistics — no separate finalisation launch.
struct ConvBnFinalize {
  const float* gamma;         // [cout_real]
  const float* beta;          // [cout_real] or null
  float* running_mean;        // [cout_real] or null
  float* running_var;
  float* stats;               // [4][kMaxC]: mean, invstd, scale, shift
  unsigned int* counter;      // zero before the launch
  double count;               // elements per channel of the resampled tensor
  float momentum, eps;
  int cout_real;              // channels of the module's conv; stored channels beyond it get scale = shift = 0
};
int conv3x3_tc(const void* in, const void* wpack, void* out, int esize, int N, int H, int W, int cin, int cout, int dil,
               const int* cnt_h, const int* cnt_w, double* stat_acc, cudaStream_t stream, bool reverse_tiles = false,
               const ConvBnFinalize* finalize = nullptr, const void* add_src = nullptr, ConvMaps* cache = nullptr);
bool conv3x3_tc_supported(int cin, int cout, int esize);
int conv_make_map(CUtensorMap* m, CUtensorMapDataType dt, const void* base, int rank, const cuuint64_t* dims,
                  const cuuint64_t* strides, const cuuint32_t* box, bool swizzle128 = true);

// The same convolution on a halo-tile pipeline whose A operand is built on chip (conv_gather.cu, bf16 only):
//   fwd:   in[h][w][c] = ReLU(scale[c] * y_prev[idx_h[h]][idx_w[w]][c] + shift[c]) — BatchNorm (stats_prev: [4][kMaxC]) + ReLU +
//          nearest resample of Y_{k-1}, never written to HBM; everything else as conv3x3_tc
//   bwd:   in = dY_k = BN-backward apply of (dA_{k+1}, Y_k) with the nearest adjoint (lo_h / lo_w: first replica of each source
//          row / column, [H + 1] / [W + 1], on the device and — for max_rep = 2, an up-sampling stage with up to 2 x 2 replicas
//          per pixel — also on the host, where the capacity of the extras stage is derived); acc = the sums of bn_bwd_reduce;
//          out = dA_k; add_src as conv3x3_tc (non-replicating stages only)
// mode: 0 fwd, 1 bwd without replicas, 2 bwd with replicas.  MRFP_ERR_UNSUPPORTED -> the caller runs the separate element-wise pass and conv3x3_tc.
bool conv3x3_gather_supported(int mode, int N, int H, int W, int SH, int SW, int cin, int cout, int dil,
                              const int* host_lo_h = nullptr, const int* host_lo_w = nullptr);
int conv3x3_gather_fwd(const void* y_prev, int SH, int SW, const int* idx_h, const int* idx_w, const float* stats_prev,
                       const void* wpack, void* out, int N, int H, int W, int cin, int cout, int dil, const int* cnt_h,
                       const int* cnt_w, double* stat_acc, cudaStream_t stream, bool reverse_tiles,
                       const ConvBnFinalize* finalize, ConvMaps* cache);
int conv3x3_gather_bwd(const void* y, const void* dA, int OH, int OW, const int* lo_h, const int* lo_w, const int* host_lo_h,
                       const int* host_lo_w, int max_rep, const float* stats, const float* gamma, const double* acc, double count,
                       int c_real, const void* wpack, void* out, int N, int H, int W, int cin, int cout, int dil,
                       cudaStream_t stream, bool reverse_tiles, const void* add_src, ConvMaps* cache,
                       const void* rk_w2t = nullptr);

// NP+ per-plane coefficients from plane totals (hrfp.cu; one block, C <= kMaxC).  forward: psum = sum_hw x -> coef = (a, b)
// with out = a*x + b, mean_out / beta_out side arrays;  backward: psum = sum_hw g, mean_in = the forward's plane means ->
// coef = (a', b') with gin = a'*g + b' (deepv3.py:268-277)
int np_coef_launch(bool backward, const double* psum, const float* alpha, const float* eps, const float* mean_in, float2* coef,
                   float* mean_out, float* beta_out, int N, int C, int HW, cudaStream_t stream);
// psum[plane] = sum_hw g[plane][:] (np_fused.cu; zeroes psum first)
int plane_sums(const float* g, double* psum, long long planes, int HW, cudaStream_t stream);

// bulk-copy forward element-wise pass (bn_ring.cu): A_next = ReLU(scale * gather(Y) + shift); MRFP_ERR_UNSUPPORTED -> LDG kernel
int bn_relu_resample_bulk(const __nv_bfloat16* y, __nv_bfloat16* a, const int* idx_h, const int* idx_w, const int* host_idx_w,
                          const float* scale, const float* shift, int N, int C, int IH, int IW, int OH, int OW, bool reverse,
                          cudaStream_t stream);
// bulk-copy BN-backward apply pass (bn_ring.cu): dY = P*mask*sum(replicas of dA) - cnt*(Q + R*y); MRFP_ERR_UNSUPPORTED -> LDG kernel
int bn_bwd_apply_bulk(const __nv_bfloat16* dA, const __nv_bfloat16* y, __nv_bfloat16* dY, const int* lo_h, const int* lo_w,
                      const int* host_lo_h, const int* host_lo_w, const float* stats, const float* gamma, const double* acc,
                      int N, int C, int IH, int IW, int OH, int OW, double count, bool reverse, cudaStream_t stream, int c_real = 0);
// bulk-copy BN-backward reduction (bn_ring.cu): U1 = sum mask*dA, U2 = sum mask*dA*y; MRFP_ERR_UNSUPPORTED -> LDG kernel
int bn_bwd_reduce_bulk(const __nv_bfloat16* dA, const __nv_bfloat16* y, const int* idx_h, const int* idx_w,
                       const int* host_idx_w, const float* stats, double* acc, int N, int C, int IH, int IW, int OH, int OW,
                       bool reverse, cudaStream_t stream);
}  // namespace mrfp
