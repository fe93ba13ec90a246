// Split-bf16 ("bf16x3") implicit-GEMM on tcgen05: parameter block shared by host and device.
//
//   D[m, n] = sum_seg sum_tap sum_c A_seg[pixel(m) shifted by tap, c] * W_seg[tap, n, c]
//
// A (activations) and W (weights) are each stored as two bf16 tensors (hi, lo) with
// x ~= hi + lo; the kernel accumulates hi*hi + lo*hi + hi*lo into one fp32 TMEM accumulator,
// which reproduces fp32 products to ~2^-16 relative ("bf16x3").
//
// A is a 4-D NHWC bf16 tensor {C, W, H, N} read by TMA boxes {64, box_w, box_h, 1}
// (box_w*box_h = 128 output pixels); out-of-bounds coordinates are zero-filled by TMA, which
// implements the conv padding.  W is a 2-D K-major tensor {K, rows}, read by boxes {64, BN}.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace pf {

constexpr int GEMM_BM = 128;      // output rows (pixels / tokens) per CTA
constexpr int GEMM_BK = 64;       // bf16 elements per k-block = one 128-byte swizzle row
constexpr int GEMM_THREADS = 320; // warp0 TMA, warp1 MMA, warps 2..9 epilogue
constexpr int GEMM_MAX_TAPS = 12;
constexpr int GEMM_HALO_ROWS = 130;            // halo stage: one image row of 128 pixels + 1 on each side
constexpr int GEMM_HALO_A_SLOT = 17 * 1024;    // 130 x 128 B rounded up to the 1024-byte swizzle atom
constexpr int GEMM_RAW_THREADS = 512;          // kernels with raw segments: warpgroup 3 = operand-conversion warps

enum GemmOutMode : int {
  OUT_F32 = 0,        // fp32 row-major [m, ldc] (+ addvec[img, n] + resid[m, n])
  OUT_SPLIT = 1,      // bf16 hi/lo row-major [m, ldc] of (acc + addvec[img, n] + resid[m, n])
  OUT_SPLIT_T = 2,    // bf16 hi/lo transposed per image: [img][n][token]
  OUT_GEGLU = 3,      // tile = [BN/2 value cols | BN/2 gate cols]: split-bf16 of (x+b)*gelu(g+b)
  OUT_SPLIT8 = 4,     // f16f8 activation operand of (acc + addvec + resid): out_hi = fp16 [m, ldc],
                      // out_lo = fp8 rows [m][ldc / 64][h8 x 64 | l8 x 64] (common.cuh)
  OUT_GEGLU8 = 5,     // OUT_GEGLU with the result stored as an f16f8 activation operand
  OUT_QKV = 6,        // fused q|k|v projection: N tiles below qkv_split -> OUT_SPLIT into out_hi/lo (q|k rows),
                      // tiles from qkv_split on -> OUT_SPLIT_T into out2_hi/lo (V^T, the P V operand)
};

struct alignas(64) GemmSeg {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  int ntaps;         // 1 (linear / 1x1) or 9 (3x3)
  int kb_per_tap;    // channels / 64
  int img_mul;       // TMA image coordinate = img * img_mul + tap_dq[tap]
  int b_tap_stride;  // W row advance per tap (= total Cout of the packed weight)
  int a_col0, b_row0, b_col0, pad0;
  int a_img_zb, a_img_zh, a_col_zb, a_col_zh;  // per-batch (z) coordinate offsets
  int b_row_zb, b_row_zh, b_col_zb, b_col_zh;
  signed char tap_dx[GEMM_MAX_TAPS], tap_dy[GEMM_MAX_TAPS], tap_dq[GEMM_MAX_TAPS];
  // halo kernels only: taps [g*gtaps, (g+1)*gtaps) share one A box of a_rows pixels (tap j starts j
  // rows in); other kernels ignore these (every tap is its own stage)
  int ngroups, gtaps, a_rows;
  // RAW segment (1x1 taps only, kernels with conversion warps -- gemm_tc.cu): the A operand is formed INSIDE
  // the GEMM from fp32 NHWC tensors.  a_hi / a_lo are then fp32 maps {C, W, H, N} (box 64 x box_w x box_h, no
  // swizzle) of the first / second concatenated source; k-blocks below raw_c0 channels come from the first.
  // v = ((x - mean_r) * rstd_r) * scale[img or 0][c] + shift[..][c], optional SiLU; mean_r / rstd_r from the
  // per-row sums raw_rowstats[row][2] = (sum x, sum x^2) over raw_rowlen channels (LayerNorm) or 0 / 1 when
  // raw_rowstats is null; scale / shift null = identity.  The stage's 32 KB A area receives the fp32 tile by
  // TMA and is converted in place (registers in between) into the K-major SWIZZLE_128B operand tiles.
  int raw, raw_c0, raw_silu, raw_rowlen;
  const float* raw_scale;
  const float* raw_shift;
  const float* raw_rowstats;
  long long raw_ld;   // scale / shift row stride per image (0: one vector for all images)
  float raw_eps;
  int pad1;
};

struct alignas(64) GemmParams {
  GemmSeg seg[2];
  int nseg;
  int nstages;
  int tiles_per_img;  // M tiles per image
  int tiles_x;        // tiles along W inside an image
  int box_w, box_h;
  int zdiv;           // batch index z -> (zb = z / zdiv, zh = z % zdiv)
  int mode;
  int n_tiles, m_tiles, z_count;  // tile grid: n fastest, then m, then z
  float* out;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  long long ldc, out_zb, out_zh, out_img;
  __nv_bfloat16* out2_hi;  // OUT_QKV: transposed part [img][n - qkv_split][token], row stride ldc2, image stride out_img2
  __nv_bfloat16* out2_lo;
  long long ldc2, out_img2;
  int qkv_split, pad2;
  float* qknorm;        // OUT_QKV (optional): [img][2 (q, k)][heads][2 halves] fp32, zero-initialised; receives
                        // atomicMax of the 32-column partial squared norms of the q / k rows (attn_tc.cuh)
  const float* addvec;  // [img, addvec_ld] or null
  const float* resid;   // [m, ldr] or null
  long long addvec_ld, ldr;
  double* stats;        // OUT_F32 only: per-(image, column) sum / sum-of-squares [img][stats_ld][2]
  long long stats_ld;
  float* rowstats;      // OUT_F32 only (optional): per-row (sum, sum of squares) of the stored values, fp32
                        // atomics into zero-initialised [m][2] (LayerNorm statistics for a RAW consumer)
  int geglu_f;          // OUT_GEGLU: number of output features F (bias layout [x: F | gate: F])
  int two_cta;          // 1: cta_group::2 kernel (tile pairs; B maps have boxes of BN/2 rows)
  int stack;            // 1 (with two_cta, BN <= 128): stacked [B_hi ; B_lo] operand, 2 MMAs per K step
  int halo;             // 1 (with two_cta, BN = 64, box 128 x 1, OUT_F32): halo stages, see gemm_tc.cu
  int f8;               // 1: f16f8 operands (a_hi/b_hi = fp16 maps, a_lo/b_lo = byte maps of the fp8 rows)
  int fast;             // 1: NON-PARITY single-pass mode (PF_FAST=1, reported separately): only the hi x hi product
                        // is issued (fp16 x fp16 for f16f8 operands, bf16 x bf16 for split-bf16 ones)
  int raw;              // 1: at least one segment is RAW (512-thread kernel with conversion warps; two_cta only)
  // nearest-2x-upsample + conv3x3 evaluated as four 2x2 parity convolutions at LOW resolution:
  // OUT_F32 rows are scattered to pixel (2y + up_py, 2x + up_px) of the [img][2H][2W] output
  int up_mode, up_py, up_px;
  double flops_override;  // reference-algorithm FLOPs of this launch for reporting (0 = 2*M*N*K)
};

// smem bytes for a given BN / stage count (incl. 1 KB alignment slack)
inline int gemm_stage_bytes(int bn) { return 2 * GEMM_BM * 128 + 2 * bn * 128; }
inline int gemm_smem_bytes(int bn, int nstages) { return nstages * gemm_stage_bytes(bn) + 1024; }
// cta_group::2: each CTA stages its 128 A rows and half of the B rows
inline int gemm_stage_bytes2(int bn) { return 2 * GEMM_BM * 128 + bn * 128; }
// halo stages: 2 A slots of 130 rows + the half-B tile pairs of three taps
inline int gemm_stage_bytes2_halo(int bn) { return 2 * GEMM_HALO_A_SLOT + 3 * bn * 128; }
// epilogue staging: 8 warps x (32 rows x 32 fp32)
inline int gemm_epilogue_smem_bytes(int /*bn*/) { return 8 * 32 * 32 * 4; }

// persistent launch: min(#tiles, num_ctas) CTAs
cudaError_t launch_gemm(const GemmParams& p, int bn, int num_ctas, cudaStream_t stream);
cudaError_t gemm_init_attrs();
// whether a kernel exists for this (tile width, operand scheme, output mode, raw) combination
bool gemm_kernel_available(const GemmParams& p, int bn);

}  // namespace pf
