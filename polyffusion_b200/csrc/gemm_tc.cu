// tcgen05 implicit-GEMM kernel (bf16x3 split precision), see gemm_tc.cuh.
//
// Roles (320 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA
// issuer (one elected lane), warps 2..9 = epilogue (TMEM -> registers -> smem staging -> global).
// Pipelines: smem ring full[]/empty[] (TMA <-> MMA) and tmem_full[]/tmem_empty[] (MMA <-> epilogue).
// Persistent: grid = min(#tiles, #SMs); each CTA walks tiles (n fastest, then m, then z) with a
// static stride.  The TMEM accumulator is double buffered (2 x BN columns) so the epilogue of tile i
// overlaps the main loop of tile i+1.
#include "common.cuh"
#include "gemm_tc.cuh"

namespace pf {

constexpr int MAX_STAGES = 8;
// Register budget: __launch_bounds__(GEMM_LB_THREADS, 1).  320 lets ptxas use up to 204 registers (it
// takes 162-168).  Building with -DPF_GEMM_LB_THREADS=512 caps the kernels at 128 registers so that
// operand-transform blocks of the other half-batch lane (unet.cu, PF_LANE_MIN_HW) can be co-resident;
// measured on B200 (profiles/r3a_lanes_ab.txt): the cap costs 0.9 ms per step and the overlap buys
// nothing because the step is limited by board power, so it is off by default.
#ifndef PF_GEMM_LB_THREADS
#define PF_GEMM_LB_THREADS 320
#endif
constexpr int GEMM_LB_THREADS = PF_GEMM_LB_THREADS;
constexpr int STG_FLOATS = 32 * 32;  // per-warp staging tile: 32 rows x 32 fp32, 16-byte groups XOR-swizzled by row

struct TileCoord {
  int n0, tile, zb, zh, img, trem, x0, y0;
};

// pair_rank < 0: t indexes single tiles; otherwise t indexes tile pairs and the CTA takes m-tile
// 2 * pair + pair_rank
__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, long long t, int bn,
                                                 int pair_rank = -1) {
  TileCoord c;
  const int nt = static_cast<int>(t % p.n_tiles);
  const long long r = t / p.n_tiles;
  const int mt = pair_rank < 0 ? p.m_tiles : p.m_tiles / 2;
  c.tile = static_cast<int>(r % mt);
  const int z = static_cast<int>(r / mt);
  if (pair_rank >= 0) c.tile = c.tile * 2 + pair_rank;
  c.n0 = nt * bn;
  c.zb = z / p.zdiv;
  c.zh = z % p.zdiv;
  c.img = c.tile / p.tiles_per_img;
  c.trem = c.tile % p.tiles_per_img;
  c.x0 = (c.trem % p.tiles_x) * p.box_w;
  c.y0 = (c.trem / p.tiles_x) * p.box_h;
  return c;
}

// TWO = cta_group::2: a cluster of two CTAs computes a 256 x BN tile pair; each CTA stages its own
// 128 A rows and HALF of the B rows (the MMA reads the other half from the peer's smem), so a stage
// is 32 KB + BN*128 B instead of 32 KB + BN*256 B: more k-blocks in flight per SM and 1/3 fewer
// operand bytes through L2.  Only the leader CTA issues MMAs; both run the epilogue on their rows.
// MODE / STATS are compile-time so that every kernel carries only ONE epilogue variant: the fully
// general kernel was ~100 KB of SASS and the short (K = 256) launches stalled on instruction fetch
// (ncu: 35 % of samples `stall_no_inst`).
//
// STACK (cta_group::2, BN <= 128): the three products are issued as TWO instructions per K = 16 step,
//   A_hi x [B_hi ; B_lo]   (N = 2 BN: each CTA's B rows are its B_hi half followed by its B_lo half,
//                           which already sit back to back in the stage)
//   A_lo x  B_hi           (N = BN, accumulated BN/2 columns further right)
// so A_hi and B_hi are read from shared memory once instead of twice (15 -> 11 KB per step per CTA
// at BN = 64, 18 -> 14 KB at BN = 128; TMA fills and UMMA operand reads share the same 128 B/clk).
// The accumulator of a tile is then 2 BN columns wide,
//   [0, h) hi*hi (n < h) | [h, BN) hi*lo + lo*hi (n < h) | [BN, BN+h) hi*hi + lo*hi (n >= h) | [BN+h, 2BN) hi*lo (n >= h)
// with h = BN / 2, and the epilogue adds the two pieces of every output column.
//
// HALO (cta_group::2, tiles that are one image row of 128 pixels): the three dx taps of a 3x3
// convolution row are served from ONE 130-pixel halo row in shared memory.  The UMMA descriptor of
// tap dx simply starts dx rows (128 B each) further in: SWIZZLE_128B is applied to absolute
// shared-memory address bits, so a start address that is not a multiple of the 1024-byte atom reads
// the rows TMA wrote there (verified on B200 with tools/probe_umma_shift.cu; the descriptor's base
// offset field stays 0).  A stage then holds one (dy, k-block): A 2 x 130 rows + the B tiles of three
// taps, and the A bytes pulled through L2 drop 2.95x.  These N = 64 convolutions at 128 x 128 were
// bound by L2 -> SM bandwidth (~11 TB/s at 40 KB per 3.1 MFLOP issued), not by the tensor pipe.
//
// F8 (GemmParams::f8, BN <= 128): fp16 + fp8 operands (common.cuh "f16f8").  The stage layout is
// unchanged -- the "hi" tiles hold fp16 h16 values, the "lo" tiles hold the [h8 | l8] fp8 rows (weights:
// [l8 | h8]) of the same k-block, 128 bytes per row either way -- but a K = 64 block costs 4 fp16 MMAs
// into accumulator 0 plus 4 e4m3 MMAs (K = 32 each, over the 128-byte fp8 row) into accumulator 1
// instead of 12 bf16 MMAs; the epilogue returns acc0 + 2^-17 acc1.  CPU emulation of the scheme on the
// whole UNet (tools/experiments/precision_emul.py): max |err| 5.9e-5, rms 9.3e-6 against the fp32
// reference (bf16x3: 2.0e-5 / 3.9e-6; tolerance 1e-4 + 1e-3 |ref|).
// NACC = accumulator stages in TMEM.  2 everywhere (the epilogue of tile i overlaps the main loop of tile
// i + 1) except the BN = 256 f16f8 kernels: their two accumulators of 256 columns fill the 512 TMEM columns,
// so the epilogue is exposed (a few % of a K >= 2304 tile) in exchange for reading every A tile from shared
// memory once per 256 output columns -- the f16f8 kernels at BN = 128 are shared-memory-bandwidth bound
// (fill 94 B/clk + UMMA operand reads 96 B/clk against 128 B/clk).
//
// RAW (cta_group::2, 512 threads): segments flagged GemmSeg::raw take their A operand straight from fp32
// NHWC tensors.  TMA drops the fp32 tile (128 pixels x 64 channels = 32 KB, exactly the size of the stage's
// hi + lo A tiles) into the stage; four conversion warps (12..15) pull it into registers, apply the optional
// per-row (LayerNorm) and per-(image, channel) (GroupNorm / LayerNorm affine) maps and SiLU, and write the
// K-major SWIZZLE_128B operand tiles back IN PLACE (split-bf16 or f16f8), fence them into the async proxy and
// arrive on the leader's conv_full barrier.  For a 1x1 consumer
// every element is converted once per N tile, so the plain split of the ResBlock skip-conv input, the
// GroupNorm'd operand of proj_in and the LayerNorm'd operands of the q/k/v and GeGLU projections are never
// written to or read from HBM.
// STATS: 0 none, 1 per-(image, column) GroupNorm sums (fp64 atomics), 2 per-row sums (fp32 atomics, LayerNorm
// statistics for a RAW consumer).
template <int BN, bool TWO, int MODE, int STATS, bool STACK, bool HALO, bool F8 = false, int NACC = 2,
          bool RAW = false>
__device__ __forceinline__ void gemm_body(const GemmParams& p) {
  static_assert(!RAW || TWO, "raw segments are implemented for cta_group::2 only");
  static_assert(!F8 || (!STACK && NACC * 2 * BN <= 512), "f16f8 needs two accumulators of BN columns per stage");
  static_assert(!STACK || (TWO && BN <= 128), "stacked B operand needs cta_group::2 and 4*BN <= 512 TMEM columns");
  static_assert(!HALO || TWO, "halo stages are implemented for cta_group::2 only");
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ __align__(8) uint64_t raw_full[RAW ? MAX_STAGES : 1];   // fp32 tile landed (this CTA)
  __shared__ __align__(8) uint64_t conv_full[RAW ? MAX_STAGES : 1];  // operand tiles written (both CTAs -> leader)
  __shared__ uint32_t tmem_base_s;

  constexpr int A_BYTES = GEMM_BM * 128;
  constexpr int B_BYTES = (TWO ? BN / 2 : BN) * 128;  // B rows staged by this CTA
  constexpr int A_SLOT = HALO ? GEMM_HALO_A_SLOT : A_BYTES;  // bytes reserved per A half (hi / lo)
  constexpr int GT = HALO ? 3 : 1;                            // taps per stage (B tile pairs reserved)
  constexpr int STAGE_BYTES = 2 * A_SLOT + GT * 2 * B_BYTES;
  constexpr uint32_t IDESC = F8 ? umma_idesc_fmt0(BN, TWO ? 256 : 128)
                                : TWO ? umma_idesc_bf16_m256(BN) : umma_idesc_bf16(BN);
  constexpr uint32_t IDESC_2N = umma_idesc_bf16_m256(STACK ? 2 * BN : BN);
  constexpr int ACC_COLS = (STACK || F8) ? 2 * BN : BN;  // TMEM columns of one accumulator stage
  constexpr int LOMUL = F8 ? 2 : 1;  // the fp8 tensors are byte maps: 2 bytes per (channel) element
  constexpr int TMEM_COLS = NACC * ACC_COLS;     // 128 / 256 / 512: power of two >= 32
  const uint32_t rank = TWO ? cluster_ctarank() : 0u;
  const bool leader = (rank == 0);
  // tile walk: 1-CTA: tile t of this CTA; 2-CTA: pair-tile t of this cluster, my m-tile = 2*pair + rank
  const long long wid = TWO ? (blockIdx.x >> 1) : blockIdx.x;
  const long long wstride = TWO ? (gridDim.x >> 1) : gridDim.x;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nstages = p.nstages;
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const long long total_tiles =
      static_cast<long long>(p.n_tiles) * (TWO ? p.m_tiles / 2 : p.m_tiles) * p.z_count;

  // a stage = one (tap group, k-block); without HALO every group is a single tap

  if (threadIdx.x == 0) {
    for (int i = 0; i < nstages; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
      if (RAW) {
        mbar_init(smem_u32(&raw_full[i]), 1);
        mbar_init(smem_u32(&conv_full[i]), 8);  // one arrive per conversion warp of both CTAs
      }
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full_bar[i]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[i]), TWO ? 16 : 8);  // one arrive per epilogue warp (both CTAs)
    }
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) {
      tma_prefetch_desc(&p.seg[s].a_hi);
      tma_prefetch_desc(&p.seg[s].a_lo);
      tma_prefetch_desc(&p.seg[s].b_hi);
      tma_prefetch_desc(&p.seg[s].b_lo);
    }
  }
  if (warp == 1) {
    if (TWO) {
      tmem_alloc2(smem_u32(&tmem_base_s), TMEM_COLS);
      tmem_relinquish2();
    } else {
      tmem_alloc(smem_u32(&tmem_base_s), TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (TWO) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();     // operands / residual / statistics of earlier kernels are complete from here on
  pdl_trigger();  // let the next kernel's CTAs start their prologue during our tail

  // RAW kernels run 512 threads = 4 warpgroups (WG0: TMA, MMA, two idle warps; WG1-2: epilogue; WG3: conversion)
  // at 128 registers each.  Re-dividing the register file with setmaxnreg (56 / 168 / 120) made things worse:
  // ptxas 12.9 compiled the WHOLE kernel under the smallest of the three limits (4.4 KB of spills against
  // 0.2 KB at a flat 128).
  constexpr int EPI_WARP0 = RAW ? 4 : 2;  // first epilogue warp (8 of them: warp & 3 = TMEM lane quarter)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int it = 0;
      for (long long t = wid; t < total_tiles; t += wstride) {
        const TileCoord tc = decode_tile(p, t, BN, TWO ? static_cast<int>(rank) : -1);
        for (int s = 0; s < p.nseg; ++s) {
          const GemmSeg& sg = p.seg[s];
          const int acol = sg.a_col0 + tc.zb * sg.a_col_zb + tc.zh * sg.a_col_zh;
          const int aimg = (tc.img + tc.zb * sg.a_img_zb + tc.zh * sg.a_img_zh) * sg.img_mul;
          const int brow = sg.b_row0 + tc.zb * sg.b_row_zb + tc.zh * sg.b_row_zh + tc.n0 +
                           (TWO ? static_cast<int>(rank) * (BN / 2) : 0);
          const int bcol = sg.b_col0 + tc.zb * sg.b_col_zb + tc.zh * sg.b_col_zh;
          const int gtaps = HALO ? sg.gtaps : 1;
          const int ngroups = HALO ? sg.ngroups : sg.ntaps;
          const uint32_t tx = 2u * static_cast<uint32_t>(HALO ? sg.a_rows * 128 : A_BYTES) +
                              static_cast<uint32_t>(gtaps) * 2u * B_BYTES;  // bytes per CTA per stage
          for (int g = 0; g < ngroups; ++g) {
            const int tp = g * gtaps;
            const int ax = tc.x0 + sg.tap_dx[tp], ay = tc.y0 + sg.tap_dy[tp];
            const int ai = aimg + sg.tap_dq[tp];
            const int br = brow + tp * sg.b_tap_stride;
            for (int kb = 0; kb < sg.kb_per_tap; ++kb, ++it) {
              const int stage = it % nstages;
              const uint32_t ph = static_cast<uint32_t>(it / nstages) & 1u;
              mbar_wait(smem_u32(&empty_bar[stage]), ph ^ 1u);
              const uint32_t fb = smem_u32(&full_bar[stage]);
              const uint32_t sa = ring + stage * STAGE_BYTES;
              if (RAW && sg.raw) {
                // fp32 tile -> this CTA's A area (own barrier, converted in place by warps 12..15); the B
                // tiles complete on the leader's barrier as usual
                const uint32_t rf = smem_u32(&raw_full[stage]);
                const int ch = acol + kb * GEMM_BK;
                const bool second = ch >= sg.raw_c0;
                mbar_expect_tx(rf, 2u * A_BYTES);
                tma_load_4d(sa, second ? &sg.a_lo : &sg.a_hi, rf, second ? ch - sg.raw_c0 : ch, ax, ay, ai);
                if (leader) mbar_expect_tx(fb, 2u * 2u * B_BYTES);
                const uint32_t sb = sa + 2 * A_SLOT;
                tma2_load_2d(sb, &sg.b_hi, fb, bcol + kb * GEMM_BK, br);
                tma2_load_2d(sb + B_BYTES, &sg.b_lo, fb, (bcol + kb * GEMM_BK) * LOMUL, br);
              } else if (TWO) {
                // both CTAs' loads complete on the LEADER's barrier, which expects both halves
                if (leader) mbar_expect_tx(fb, 2 * tx);
                tma2_load_4d(sa, &sg.a_hi, fb, acol + kb * GEMM_BK, ax, ay, ai);
                tma2_load_4d(sa + A_SLOT, &sg.a_lo, fb, (acol + kb * GEMM_BK) * LOMUL, ax, ay, ai);
#pragma unroll
                for (int j = 0; j < GT; ++j) {
                  if (j < gtaps) {
                    const uint32_t sb = sa + 2 * A_SLOT + j * 2 * B_BYTES;
                    tma2_load_2d(sb, &sg.b_hi, fb, bcol + kb * GEMM_BK, br + j * sg.b_tap_stride);
                    tma2_load_2d(sb + B_BYTES, &sg.b_lo, fb, (bcol + kb * GEMM_BK) * LOMUL, br + j * sg.b_tap_stride);
                  }
                }
              } else {
                mbar_expect_tx(fb, tx);
                tma_load_4d(sa, &sg.a_hi, fb, acol + kb * GEMM_BK, ax, ay, ai);
                tma_load_4d(sa + A_BYTES, &sg.a_lo, fb, (acol + kb * GEMM_BK) * LOMUL, ax, ay, ai);
                tma_load_2d(sa + 2 * A_BYTES, &sg.b_hi, fb, bcol + kb * GEMM_BK, br);
                tma_load_2d(sa + 2 * A_BYTES + B_BYTES, &sg.b_lo, fb, (bcol + kb * GEMM_BK) * LOMUL, br);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && elect_one()) {
      int it = 0;
      int lt = 0;  // local tile counter
      uint32_t conv_bits = 0;  // phase parity of conv_full[] per stage
      for (long long t = wid; t < total_tiles; t += wstride, ++lt) {
        const int as = lt % NACC;
        const uint32_t aph = static_cast<uint32_t>(lt / NACC) & 1u;
        mbar_wait(smem_u32(&tmem_empty_bar[as]), aph ^ 1u);
        tc_fence_after();
        const uint32_t acc = tmem_base + static_cast<uint32_t>(as * ACC_COLS);
        int kbi = 0;  // stages consumed for this tile
        for (int s = 0; s < p.nseg; ++s) {
          const GemmSeg& sg = p.seg[s];
          const int gtaps = HALO ? sg.gtaps : 1;
          const int nst = (HALO ? sg.ngroups : sg.ntaps) * sg.kb_per_tap;
          for (int si = 0; si < nst; ++si, ++kbi, ++it) {
            const int stage = it % nstages;
            const uint32_t ph = static_cast<uint32_t>(it / nstages) & 1u;
            mbar_wait(smem_u32(&full_bar[stage]), ph);
            if (RAW && sg.raw) {
              mbar_wait(smem_u32(&conv_full[stage]), (conv_bits >> stage) & 1u);
              conv_bits ^= 1u << stage;
            }
            tc_fence_after();
            const uint32_t sa = ring + stage * STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < GT; ++j) {
              if (j < gtaps) {
                // tap j of the group: A starts j halo rows (128 B each) further in
                const uint64_t da_hi = umma_desc_sw128(sa + j * 128);
                const uint64_t da_lo = umma_desc_sw128(sa + A_SLOT + j * 128);
                const uint64_t db_hi = umma_desc_sw128(sa + 2 * A_SLOT + j * 2 * B_BYTES);
                const uint64_t db_lo = umma_desc_sw128(sa + 2 * A_SLOT + j * 2 * B_BYTES + B_BYTES);
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; ++k) {
                  const uint64_t ko = static_cast<uint64_t>(k * 2);  // 32 B per K=16 step (16 B units)
                  const uint32_t accum = (kbi | j | k) != 0;
                  if (p.fast) {
                    // single-pass mode: hi x hi only, plain accumulator layout (N = BN)
                    if (TWO) umma2_bf16(acc, da_hi + ko, db_hi + ko, IDESC, accum);
                    else umma_bf16(acc, da_hi + ko, db_hi + ko, IDESC, accum);
                  } else if (F8) {
                    // fp16 x fp16 (K = 16) and e4m3 x e4m3 (K = 32): 32 bytes of the row each
                    if (TWO) {
                      umma2_bf16(acc, da_hi + ko, db_hi + ko, IDESC, accum);
                      umma2_f8(acc + BN, da_lo + ko, db_lo + ko, IDESC, accum);
                    } else {
                      umma_bf16(acc, da_hi + ko, db_hi + ko, IDESC, accum);
                      umma_f8(acc + BN, da_lo + ko, db_lo + ko, IDESC, accum);
                    }
                  } else if (STACK) {
                    umma2_bf16(acc, da_hi + ko, db_hi + ko, IDESC_2N, accum);
                    umma2_bf16(acc + BN / 2, da_lo + ko, db_hi + ko, IDESC, 1u);
                  } else if (TWO) {
                    umma2_bf16(acc, da_lo + ko, db_hi + ko, IDESC, accum);
                    umma2_bf16(acc, da_hi + ko, db_lo + ko, IDESC, 1u);
                    umma2_bf16(acc, da_hi + ko, db_hi + ko, IDESC, 1u);
                  } else {
                    umma_bf16(acc, da_lo + ko, db_hi + ko, IDESC, accum);
                    umma_bf16(acc, da_hi + ko, db_lo + ko, IDESC, 1u);
                    umma_bf16(acc, da_hi + ko, db_hi + ko, IDESC, 1u);
                  }
                }
              }
            }
            if (TWO) umma2_commit_mc(smem_u32(&empty_bar[stage]));
            else umma_commit(smem_u32(&empty_bar[stage]));
          }
        }
        if (TWO) umma2_commit_mc(smem_u32(&tmem_full_bar[as]));
        else umma_commit(smem_u32(&tmem_full_bar[as]));
      }
    }
  } else if (RAW && warp >= 12) {
    // ------------------------------------------------------------------ operand conversion (4 warps)
    // thread = channels [4j, 4j + 4) and [32 + 4j, 32 + 4j + 4) of rows r0, r0 + 16, ... of every raw k-block:
    // a quarter-warp reads 128 contiguous bytes per request (no bank conflicts) and the two 8-byte halves of a
    // 16-byte operand chunk come from neighbouring lanes.  The code is specialised per transform kind and kept
    // small (the conversion warps were instruction-fetch bound with one general 36 KB loop body).
    const int tt = threadIdx.x - 384;  // 0..127
    const int j = tt & 7;
    // rows r0 + 16 i.  A warp takes rows w, w + 4, w + 8, w + 12: their swizzle phases (row & 7) pair up as
    // {w, w + 4}, so the four 64-byte row pieces of one STS.64 cover all 32 banks twice (2 wavefronts, the
    // minimum); with four consecutive rows per warp they fell on the same 16 banks (2-way conflicts)
    const int r0 = ((tt >> 3) & 3) * 4 + (tt >> 5);
    uint8_t* gsm = smem_raw + (ring - smem_u32(smem_raw));
    uint32_t raw_bits = 0;
    int it = 0;
    for (long long t = wid; t < total_tiles; t += wstride) {
      const TileCoord tc = decode_tile(p, t, BN, static_cast<int>(rank));
      for (int s = 0; s < p.nseg; ++s) {
        const GemmSeg& sg = p.seg[s];
        const int nst = (HALO ? sg.ngroups : sg.ntaps) * sg.kb_per_tap;
        if (!sg.raw) {
          it += nst;
          continue;
        }
        const bool rown = sg.raw_rowstats != nullptr;
        const bool aff = sg.raw_scale != nullptr;
        float rmul[8], radd[8];
        if (rown) {
          const float inv = 1.0f / static_cast<float>(sg.raw_rowlen);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const long long m = static_cast<long long>(tc.tile) * GEMM_BM + i * 16 + r0;
            const float2 st = __ldcg(reinterpret_cast<const float2*>(sg.raw_rowstats) + m);
            const float mean = st.x * inv;
            const float var = fmaxf(fmaf(-mean, mean, st.y * inv), 0.f);
            const float rstd = rsqrtf(var + sg.raw_eps);
            rmul[i] = rstd;
            radd[i] = -mean * rstd;
          }
        }
        // per-channel scale / shift of the first k-block (the next one is fetched while this one is converted)
        const float* scp = aff ? sg.raw_scale + static_cast<long long>(tc.img) * sg.raw_ld + sg.a_col0 + j * 4 : nullptr;
        const float* shp = aff ? sg.raw_shift + static_cast<long long>(tc.img) * sg.raw_ld + sg.a_col0 + j * 4 : nullptr;
        float4 sca, scb, sha, shb;
        if (aff) {
          sca = __ldg(reinterpret_cast<const float4*>(scp));
          scb = __ldg(reinterpret_cast<const float4*>(scp + 32));
          sha = __ldg(reinterpret_cast<const float4*>(shp));
          shb = __ldg(reinterpret_cast<const float4*>(shp + 32));
        }
        for (int kb = 0; kb < nst; ++kb, ++it) {
          const int stage = it % nstages;
          mbar_wait(smem_u32(&raw_full[stage]), (raw_bits >> stage) & 1u);
          raw_bits ^= 1u << stage;
          uint8_t* a0 = gsm + stage * STAGE_BYTES;
          float4 va[8], vb[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4* rp = reinterpret_cast<const float4*>(a0 + (i * 16 + r0) * 256 + j * 16);
            va[i] = rp[0];
            vb[i] = rp[8];
          }
          // every thread has its fp32 values in registers before anyone overwrites the tile
          asm volatile("bar.sync 2, 128;" ::: "memory");
          float4 nsa, nsb, nha, nhb;
          if (aff && kb + 1 < nst) {
            nsa = __ldg(reinterpret_cast<const float4*>(scp + (kb + 1) * GEMM_BK));
            nsb = __ldg(reinterpret_cast<const float4*>(scp + (kb + 1) * GEMM_BK + 32));
            nha = __ldg(reinterpret_cast<const float4*>(shp + (kb + 1) * GEMM_BK));
            nhb = __ldg(reinterpret_cast<const float4*>(shp + (kb + 1) * GEMM_BK + 32));
          }
          if (rown) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              va[i].x = fmaf(va[i].x, rmul[i], radd[i]); va[i].y = fmaf(va[i].y, rmul[i], radd[i]);
              va[i].z = fmaf(va[i].z, rmul[i], radd[i]); va[i].w = fmaf(va[i].w, rmul[i], radd[i]);
              vb[i].x = fmaf(vb[i].x, rmul[i], radd[i]); vb[i].y = fmaf(vb[i].y, rmul[i], radd[i]);
              vb[i].z = fmaf(vb[i].z, rmul[i], radd[i]); vb[i].w = fmaf(vb[i].w, rmul[i], radd[i]);
            }
          }
          if (aff) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              va[i].x = fmaf(va[i].x, sca.x, sha.x); va[i].y = fmaf(va[i].y, sca.y, sha.y);
              va[i].z = fmaf(va[i].z, sca.z, sha.z); va[i].w = fmaf(va[i].w, sca.w, sha.w);
              vb[i].x = fmaf(vb[i].x, scb.x, shb.x); vb[i].y = fmaf(vb[i].y, scb.y, shb.y);
              vb[i].z = fmaf(vb[i].z, scb.z, shb.z); vb[i].w = fmaf(vb[i].w, scb.w, shb.w);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int prow = i * 16 + r0;
            const int x7 = prow & 7;
            uint8_t* hrow = a0 + prow * 128 + (j & 1) * 8;
            if constexpr (F8) {
              uint2 ha, hb;
              uint32_t h8a, l8a, h8b, l8b;
              split_f8x4(va[i].x, va[i].y, va[i].z, va[i].w, 1.f, F8_ACT_LO_SCALE, ha, h8a, l8a);
              split_f8x4(vb[i].x, vb[i].y, vb[i].z, vb[i].w, 1.f, F8_ACT_LO_SCALE, hb, h8b, l8b);
              *reinterpret_cast<uint2*>(hrow + (((j >> 1) ^ x7) << 4)) = ha;
              *reinterpret_cast<uint2*>(hrow + (((4 + (j >> 1)) ^ x7) << 4)) = hb;
              uint8_t* lrow = a0 + A_SLOT + prow * 128 + (j & 3) * 4;
              *reinterpret_cast<uint32_t*>(lrow + (((j >> 2) ^ x7) << 4)) = h8a;
              *reinterpret_cast<uint32_t*>(lrow + (((2 + (j >> 2)) ^ x7) << 4)) = h8b;
              *reinterpret_cast<uint32_t*>(lrow + (((4 + (j >> 2)) ^ x7) << 4)) = l8a;
              *reinterpret_cast<uint32_t*>(lrow + (((6 + (j >> 2)) ^ x7) << 4)) = l8b;
            } else {
              uint2 ha, la, hb, lb;
              split2(va[i].x, va[i].y, ha.x, la.x);
              split2(va[i].z, va[i].w, ha.y, la.y);
              split2(vb[i].x, vb[i].y, hb.x, lb.x);
              split2(vb[i].z, vb[i].w, hb.y, lb.y);
              const int oa = ((j >> 1) ^ x7) << 4, ob = ((4 + (j >> 1)) ^ x7) << 4;
              *reinterpret_cast<uint2*>(hrow + oa) = ha;
              *reinterpret_cast<uint2*>(hrow + ob) = hb;
              *reinterpret_cast<uint2*>(hrow + A_SLOT + oa) = la;
              *reinterpret_cast<uint2*>(hrow + A_SLOT + ob) = lb;
            }
          }
          // generic-proxy writes -> visible to the tensor core (async proxy) of THIS SM, which is the one that
          // reads this CTA's A tile (the leader only issues the instruction).  The .shared::cta form is a
          // FENCE.VIEW.ASYNC; the unqualified fence and a cluster-scope release arrive both lower to
          // MEMBAR.ALL.GPU, which cost ~0.7 us per k-block here.
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(smem_u32(&conv_full[stage]));
          if (aff) {
            sca = nsa; scb = nsb; sha = nha; shb = nhb;
          }
        }
      }
    }
  } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + 8) {
    // ------------------------------------------------------------------ epilogue (8 warps)
    // TMEM -> registers (thread = output row) -> per-warp smem staging -> coalesced global stores:
    // in the store phase a quarter-warp (8 lanes x float4) covers one 128-byte row segment, so
    // every store / residual load touches whole 128 B lines.  Two warps share each TMEM lane
    // quarter and take alternate 32-column chunks.  Residual values are prefetched into registers
    // before the TMEM load so their DRAM latency overlaps the TMEM/smem round trip.  Optional
    // per-channel sum / sum-of-squares of the stored values (GroupNorm statistics for the consumer)
    // are reduced in registers -> shuffles -> smem -> one fp64 atomicAdd per (tile, column).
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int ewarp = warp - EPI_WARP0;  // 0..7
    const int chalf = ewarp >> 2;      // which alternate chunks this warp takes
    uint8_t* epi_base = smem_raw + (ring - smem_u32(smem_raw)) + nstages * STAGE_BYTES;
    float* stg_all = reinterpret_cast<float*>(epi_base);
    float* stg = stg_all + ewarp * STG_FLOATS;
    const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..255 among epilogue threads
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;
    int lt = 0;
    for (long long t = wid; t < total_tiles; t += wstride, ++lt) {
      const TileCoord tc = decode_tile(p, t, BN, TWO ? static_cast<int>(rank) : -1);
      const int as = lt % NACC;
      const uint32_t aph = static_cast<uint32_t>(lt / NACC) & 1u;
      const long long m0 = static_cast<long long>(tc.tile) * GEMM_BM + q * 32;  // first row of this warp
      const long long zoff = tc.zb * p.out_zb + tc.zh * p.out_zh;
      if ((MODE == OUT_F32 || MODE == OUT_SPLIT || MODE == OUT_SPLIT8) && p.resid != nullptr) {
        // Pull this warp's share of the residual into L2 one tile AHEAD (the residual was written
        // several kernels ago and has usually left L2; a DRAM round trip per chunk would otherwise
        // dominate the epilogue of the small-K GEMMs): one 128-byte line per (row, 32-col chunk).
        // The first tile of the CTA is prefetched here too, while its main loop is still running.
        const long long tn = t + wstride;
        if (tn < total_tiles) {
          const TileCoord nc = decode_tile(p, tn, BN, TWO ? static_cast<int>(rank) : -1);
#pragma unroll
          for (int ci = 0; ci < BN / 64; ++ci) {
            const float* ptr = p.resid + (static_cast<long long>(nc.tile) * GEMM_BM + q * 32 + lane) * p.ldr +
                               nc.n0 + chalf * 32 + ci * 64;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
          }
        }
        if (lt == 0) {
#pragma unroll
          for (int ci = 0; ci < BN / 64; ++ci) {
            const float* ptr = p.resid + (m0 + lane) * p.ldr + tc.n0 + chalf * 32 + ci * 64;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
          }
        }
      }
      mbar_wait(smem_u32(&tmem_full_bar[as]), aph);
      tc_fence_after();
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * ACC_COLS);
      // logical accumulator columns [c, c + 32) of this warp's 32 rows (wait included)
      auto ld_acc = [&](int c, uint32_t (&v)[32]) {
        if (p.fast) {
          tmem_ld32(taddr + c, v);
          tmem_ld_wait();
          return;
        }
        if constexpr (F8) {
          uint32_t w[32];
          tmem_ld32(taddr + c, v);
          tmem_ld32(taddr + BN + c, w);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = __float_as_uint(fmaf(__uint_as_float(w[j]), F8_CROSS_SCALE, __uint_as_float(v[j])));
        } else if constexpr (STACK) {
          constexpr int h = BN / 2;
          const int ca = (c / h) * BN + (c % h);
          uint32_t w[32];
          tmem_ld32(taddr + ca, v);
          tmem_ld32(taddr + ca + h, w);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
        } else {
          tmem_ld32(taddr + c, v);
          tmem_ld_wait();
        }
      };

      // OUT_QKV: the tile's column range decides between the row-major and the transposed store
      const bool transposed = MODE == OUT_SPLIT_T || (MODE == OUT_QKV && tc.n0 >= p.qkv_split);
      if (transposed) {
        if constexpr (MODE == OUT_SPLIT_T || MODE == OUT_QKV) {
          // [img][n][token]; consecutive lanes -> consecutive tokens (already coalesced)
          __nv_bfloat16* th = MODE == OUT_QKV ? p.out2_hi : p.out_hi;
          __nv_bfloat16* tl = MODE == OUT_QKV ? p.out2_lo : p.out_lo;
          const long long tld = MODE == OUT_QKV ? p.ldc2 : p.ldc;
          const long long timg = MODE == OUT_QKV ? p.out_img2 : p.out_img;
          const int tn0 = MODE == OUT_QKV ? tc.n0 - p.qkv_split : tc.n0;
          const long long tok = static_cast<long long>(tc.trem) * GEMM_BM + q * 32 + lane;
          const long long base = zoff + static_cast<long long>(tc.img) * timg + tok;
#pragma unroll 1
          for (int c = chalf * 32; c < BN; c += 64) {
            uint32_t v[32];
            ld_acc(c, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              __nv_bfloat16 h, l;
              split_bf16(__uint_as_float(v[j]), h, l);
              const long long o = base + static_cast<long long>(tn0 + c + j) * tld;
              th[o] = h;
              tl[o] = l;
            }
          }
        }
      } else if constexpr (MODE != OUT_SPLIT_T) {
        constexpr bool geglu = (MODE == OUT_GEGLU || MODE == OUT_GEGLU8);
        const int ncols = geglu ? BN / 2 : BN;          // output columns produced by this tile
        const int ocol0 = geglu ? (tc.n0 / 2) : tc.n0;  // first output column
        const float* av = p.addvec ? p.addvec + static_cast<long long>(tc.img) * p.addvec_ld : nullptr;
        const bool has_res = (MODE == OUT_F32 || MODE == OUT_SPLIT || MODE == OUT_SPLIT8) && p.resid != nullptr;
        float4 csum[STATS == 1 ? BN / 64 : 1], csq[STATS == 1 ? BN / 64 : 1];  // per-chunk column partial sums
        float rsum[STATS == 2 ? 8 : 1], rsq[STATS == 2 ? 8 : 1];               // per-row partial sums
        if constexpr (STATS == 2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) rsum[i] = rsq[i] = 0.f;
        }
#pragma unroll(STATS == 1 ? BN / 64 : 1)
        for (int ci = 0; ci < BN / 64; ++ci) {
          const int c = chalf * 32 + ci * 64;
          if (c >= ncols) break;
          float4 rres[8];
          if (has_res) {
#pragma unroll
            for (int it = 0; it < 8; ++it)
              rres[it] = __ldg(reinterpret_cast<const float4*>(
                  p.resid + (m0 + it * 4 + rsub) * p.ldr + tc.n0 + c + c4));
          }
          {
            uint32_t v[32];
            ld_acc(c, v);
            if constexpr (MODE == OUT_QKV) {
              // q / k tiles: bound of the largest squared row norm per (image, head, 32-column half) for the
              // attention kernel's single-pass stabiliser
              if (p.qknorm != nullptr) {
                float ss = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) ss = fmaf(__uint_as_float(v[j]), __uint_as_float(v[j]), ss);
                const unsigned mxu = __reduce_max_sync(0xffffffffu, __float_as_uint(ss));
                if (lane == 0) {
                  const int cdim = p.qkv_split >> 1, col = tc.n0 + c;
                  const int which = col / cdim, hc = col % cdim;
                  unsigned* d = reinterpret_cast<unsigned*>(p.qknorm) +
                                ((static_cast<long long>(tc.img) * 2 + which) * (cdim >> 6) + (hc >> 6)) * 2 + ((hc >> 5) & 1);
                  atomicMax(d, mxu);
                }
              }
            }
            if constexpr (geglu) {
              uint32_t g[32];
              ld_acc(BN / 2 + c, g);
              // value = cols [ocol0 + c, +32) of the first half, gate = same cols of the second
              // half of the (un-interleaved) projection: out = (x + b_x) * gelu(g + b_g)
              const float4* bx = reinterpret_cast<const float4*>(av + ocol0 + c);
              const float4* bg = reinterpret_cast<const float4*>(av + p.geglu_f + ocol0 + c);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b1 = __ldg(bx + j), b2 = __ldg(bg + j);
                const float bxv[4] = {b1.x, b1.y, b1.z, b1.w}, bgv[4] = {b2.x, b2.y, b2.z, b2.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float xv = __uint_as_float(v[4 * j + k]) + bxv[k];
                  const float gv = __uint_as_float(g[4 * j + k]) + bgv[k];
                  v[4 * j + k] = __float_as_uint(xv * gelu_erf_fast(gv));
                }
              }
            } else {
              if (av) {
                const float4* b4 = reinterpret_cast<const float4*>(av + tc.n0 + c);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 b = __ldg(b4 + j);
                  v[4 * j + 0] = __float_as_uint(__uint_as_float(v[4 * j + 0]) + b.x);
                  v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + b.y);
                  v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + b.z);
                  v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + b.w);
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                  make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          __syncwarp();
          float4 ssum = make_float4(0.f, 0.f, 0.f, 0.f), ssq = ssum;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int row = it * 4 + rsub;
            float4 o = *reinterpret_cast<const float4*>(stg + row * 32 + (((lane & 7) ^ (row & 7)) << 2));
            long long m = m0 + row;
            if (p.up_mode) {
              // low-res pixel (y, x) of image img -> high-res pixel (2y + py, 2x + px)
              const int r = q * 32 + row;
              const int y = tc.y0 + r / p.box_w, x = tc.x0 + r % p.box_w;
              const int Wl = p.tiles_x * p.box_w, Hl = (p.tiles_per_img / p.tiles_x) * p.box_h;
              m = (static_cast<long long>(tc.img) * (2 * Hl) + 2 * y + p.up_py) * (2 * Wl) + 2 * x + p.up_px;
            }
            if constexpr (MODE == OUT_F32) {
              if (has_res) {
                o.x += rres[it].x; o.y += rres[it].y; o.z += rres[it].z; o.w += rres[it].w;
              }
              // (st.global.cs / .cg hints measured: no change, profiles/r4m_*; without these stores the GEMMs of
              // a step take 1.75 ms less, about the HBM time of the 7.4 GB they write, profiles/r4l_*)
              *reinterpret_cast<float4*>(p.out + zoff + m * p.ldc + tc.n0 + c + c4) = o;
              if constexpr (STATS == 1) {
                ssum.x += o.x; ssum.y += o.y; ssum.z += o.z; ssum.w += o.w;
                ssq.x = fmaf(o.x, o.x, ssq.x); ssq.y = fmaf(o.y, o.y, ssq.y);
                ssq.z = fmaf(o.z, o.z, ssq.z); ssq.w = fmaf(o.w, o.w, ssq.w);
              }
              if constexpr (STATS == 2) {
                rsum[it] += (o.x + o.y) + (o.z + o.w);
                rsq[it] = fmaf(o.x, o.x, fmaf(o.y, o.y, fmaf(o.z, o.z, fmaf(o.w, o.w, rsq[it]))));
              }
            } else if constexpr (MODE == OUT_SPLIT8 || MODE == OUT_GEGLU8) {
              // f16f8 activation operand: h16 [m, ldc] fp16 + fp8 rows [m][ldc / 64][h8 x 64 | l8 x 64]
              if (MODE == OUT_SPLIT8 && has_res) {
                o.x += rres[it].x; o.y += rres[it].y; o.z += rres[it].z; o.w += rres[it].w;
              }
              uint2 h16;
              uint32_t h8, l8;
              split_f8x4(o.x, o.y, o.z, o.w, 1.f, F8_ACT_LO_SCALE, h16, h8, l8);
              const int col = ocol0 + c + c4;
              const long long row = zoff + m * p.ldc;
              *reinterpret_cast<uint2*>(p.out_hi + row + col) = h16;
              uint8_t* b8 = reinterpret_cast<uint8_t*>(p.out_lo) + row * 2 + (col >> 6) * 128 + (col & 63);
              *reinterpret_cast<uint32_t*>(b8) = h8;
              *reinterpret_cast<uint32_t*>(b8 + 64) = l8;
            } else {  // OUT_SPLIT (+ residual) / OUT_GEGLU: split-bf16 row-major
              if constexpr (MODE == OUT_SPLIT) {
                if (has_res) {
                  o.x += rres[it].x; o.y += rres[it].y; o.z += rres[it].z; o.w += rres[it].w;
                }
              }
              uint2 h, l;
              split2(o.x, o.y, h.x, l.x);
              split2(o.z, o.w, h.y, l.y);
              const long long off = zoff + m * p.ldc + ocol0 + c + c4;
              *reinterpret_cast<uint2*>(p.out_hi + off) = h;
              *reinterpret_cast<uint2*>(p.out_lo + off) = l;
            }
          }
          if constexpr (STATS == 1) {
            // reduce over the 4 row-groups of the warp (lanes with equal lane&7)
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
              ssum.x += __shfl_xor_sync(0xffffffffu, ssum.x, o);
              ssum.y += __shfl_xor_sync(0xffffffffu, ssum.y, o);
              ssum.z += __shfl_xor_sync(0xffffffffu, ssum.z, o);
              ssum.w += __shfl_xor_sync(0xffffffffu, ssum.w, o);
              ssq.x += __shfl_xor_sync(0xffffffffu, ssq.x, o);
              ssq.y += __shfl_xor_sync(0xffffffffu, ssq.y, o);
              ssq.z += __shfl_xor_sync(0xffffffffu, ssq.z, o);
              ssq.w += __shfl_xor_sync(0xffffffffu, ssq.w, o);
            }
            csum[ci] = ssum;
            csq[ci] = ssq;
          }
          __syncwarp();
        }
        if constexpr (STATS == 2) {
          // a row's 32-column pieces sit in the 8 lanes of a quarter-warp: reduce, then one atomic pair per
          // (row, warp, tile); the other column chunks / N tiles of the row add theirs
#pragma unroll
          for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int o = 1; o <= 4; o <<= 1) {
              rsum[i] += __shfl_xor_sync(0xffffffffu, rsum[i], o);
              rsq[i] += __shfl_xor_sync(0xffffffffu, rsq[i], o);
            }
          }
          if ((lane & 7) == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float* d = p.rowstats + (m0 + i * 4 + rsub) * 2;
              atomicAdd(d, rsum[i]);
              atomicAdd(d + 1, rsq[i]);
            }
          }
        }
        if (STATS == 1 && lane < 8) {
          // park the partials in this warp's (now idle) staging tile: [chunk][sum | sq][32 cols]
#pragma unroll
          for (int ci = 0; ci < BN / 64; ++ci) {
            *reinterpret_cast<float4*>(stg + (ci * 2 + 0) * 32 + c4) = csum[ci];
            *reinterpret_cast<float4*>(stg + (ci * 2 + 1) * 32 + c4) = csq[ci];
          }
        }
      }
      // all of this warp's TMEM reads for the tile are complete (tmem_ld_wait above)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (TWO) mbar_arrive_leader(smem_u32(&tmem_empty_bar[as]));
        else mbar_arrive(smem_u32(&tmem_empty_bar[as]));
      }
      if constexpr (STATS == 1) {
        // flush this tile's column sums: [img][stats_ld][2] fp64
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et < BN) {
          // column et lives in chunk et/32: warps (chunk&1)*4 + quarter, local chunk index chunk>>1
          const int chunk = et >> 5, cl = et & 31;
          const float* src = stg_all + ((chunk & 1) * 4) * STG_FLOATS + ((chunk >> 1) * 2) * 32 + cl;
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            s1 += src[qq * STG_FLOATS];
            s2 += src[qq * STG_FLOATS + 32];
          }
          double* d = p.stats + (static_cast<long long>(tc.img) * p.stats_ld + tc.n0 + et) * 2;
          atomicAdd(d, static_cast<double>(s1));
          atomicAdd(d + 1, static_cast<double>(s2));
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
  }

  tc_fence_before();
  if (TWO) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (TWO) tmem_dealloc2(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BN, int MODE, int STATS>
__global__ void __launch_bounds__(GEMM_LB_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  gemm_body<BN, false, MODE, STATS, false, false>(p);
}

template <int BN, int MODE, int STATS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_LB_THREADS, 1)
    gemm_tc2_kernel(const __grid_constant__ GemmParams p) {
  gemm_body<BN, true, MODE, STATS, false, false>(p);
}

// cta_group::2 with the stacked [B_hi ; B_lo] operand (BN <= 128)
template <int BN, int MODE, int STATS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_LB_THREADS, 1)
    gemm_tc2s_kernel(const __grid_constant__ GemmParams p) {
  gemm_body<BN, true, MODE, STATS, true, false>(p);
}

// stacked B + halo stages (one 130-pixel A row serves the three dx taps); fp32 output only
template <int BN, int STATS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_LB_THREADS, 1)
    gemm_tc2h_kernel(const __grid_constant__ GemmParams p) {
  gemm_body<BN, true, OUT_F32, STATS, true, true>(p);
}

// f16f8 operands (fp16 main product + e4m3 cross terms): cta_group::2, halo and single-CTA forms
template <int BN, int MODE, int STATS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_LB_THREADS, 1)
    gemm_tc2f_kernel(const __grid_constant__ GemmParams p) {
  gemm_body<BN, true, MODE, STATS, false, false, true>(p);
}
template <int BN, int STATS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_LB_THREADS, 1)
    gemm_tc2fh_kernel(const __grid_constant__ GemmParams p) {
  gemm_body<BN, true, OUT_F32, STATS, false, true, true>(p);
}
template <int BN, int MODE, int STATS>
__global__ void __launch_bounds__(GEMM_LB_THREADS, 1) gemm_tcf_kernel(const __grid_constant__ GemmParams p) {
  gemm_body<BN, false, MODE, STATS, false, false, true>(p);
}
// BN = 256, one accumulator stage (see NACC above)
template <int MODE, int STATS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_LB_THREADS, 1)
    gemm_tc2f256_kernel(const __grid_constant__ GemmParams p) {
  gemm_body<256, true, MODE, STATS, false, false, true, 1>(p);
}

// RAW-segment kernels (512 threads: the conversion warps are warpgroup 3), only the shapes the UNet plan uses
template <int BN, int MODE, int STATS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_RAW_THREADS, 1)
    gemm_tc2f_raw_kernel(const __grid_constant__ GemmParams p) {
  gemm_body<BN, true, MODE, STATS, false, false, true, 2, true>(p);
}
template <int BN, int STATS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_RAW_THREADS, 1)
    gemm_tc2fh_raw_kernel(const __grid_constant__ GemmParams p) {
  gemm_body<BN, true, OUT_F32, STATS, false, true, true, 2, true>(p);
}
template <int BN, int MODE, int STATS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_RAW_THREADS, 1)
    gemm_tc2s_raw_kernel(const __grid_constant__ GemmParams p) {
  gemm_body<BN, true, MODE, STATS, true, false, false, 2, true>(p);
}
template <int BN, int MODE, int STATS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_RAW_THREADS, 1)
    gemm_tc2_raw_kernel(const __grid_constant__ GemmParams p) {
  gemm_body<BN, true, MODE, STATS, false, false, false, 2, true>(p);
}

#ifndef PF_GEMM_NO_DISPATCH  // (register-allocation experiments instantiate single kernels: tools/)
typedef void (*GemmKernel)(GemmParams);

// variant index: 0 = fp32, 1 = fp32 + GroupNorm statistics, 2 = split, 3 = split transposed, 4 = GeGLU
// kind: 0 = one CTA per tile, 1 = cta_group::2, 2 = cta_group::2 with stacked B (BN <= 128)
template <int BN, int KIND>
static GemmKernel pick_variant(int v) {
  if constexpr (KIND == 2) {
    switch (v) {
      case 0: return gemm_tc2s_kernel<BN, OUT_F32, false>;
      case 1: return gemm_tc2s_kernel<BN, OUT_F32, true>;
      case 2: return gemm_tc2s_kernel<BN, OUT_SPLIT, false>;
      case 3: return gemm_tc2s_kernel<BN, OUT_SPLIT_T, false>;
      case 5: return gemm_tc2s_kernel<BN, OUT_SPLIT8, false>;
      case 6: return gemm_tc2s_kernel<BN, OUT_F32, 2>;
      case 8: return gemm_tc2s_kernel<BN, OUT_QKV, false>;
      case 4: return gemm_tc2s_kernel<BN, OUT_GEGLU, false>;
      default: return nullptr;
    }
  } else if constexpr (KIND == 1) {
    switch (v) {
      case 0: return gemm_tc2_kernel<BN, OUT_F32, false>;
      case 1: return gemm_tc2_kernel<BN, OUT_F32, true>;
      case 2: return gemm_tc2_kernel<BN, OUT_SPLIT, false>;
      case 3: return gemm_tc2_kernel<BN, OUT_SPLIT_T, false>;
      case 5: return gemm_tc2_kernel<BN, OUT_SPLIT8, false>;
      case 6: return gemm_tc2_kernel<BN, OUT_F32, 2>;
      case 4: return gemm_tc2_kernel<BN, OUT_GEGLU, false>;
      default: return nullptr;
    }
  } else {
    switch (v) {
      case 0: return gemm_tc_kernel<BN, OUT_F32, false>;
      case 1: return gemm_tc_kernel<BN, OUT_F32, true>;
      case 2: return gemm_tc_kernel<BN, OUT_SPLIT, false>;
      case 3: return gemm_tc_kernel<BN, OUT_SPLIT_T, false>;
      case 5: return gemm_tc_kernel<BN, OUT_SPLIT8, false>;
      case 6: return gemm_tc_kernel<BN, OUT_F32, 2>;
      case 4: return gemm_tc_kernel<BN, OUT_GEGLU, false>;
      default: return nullptr;
    }
  }
}

// f16f8 kernels: kind 4 = cta_group::2, 5 = cta_group::2 + halo stages, 6 = one CTA per tile;
// variants 0 = fp32, 1 = fp32 + statistics, 2 = split-bf16 output, 5 = f16f8 split output
template <int BN>
static GemmKernel pick_f8(int kind, int v) {
  if (kind == 5) {
    if constexpr (BN == 64) return v == 1 ? gemm_tc2fh_kernel<64, true> : v == 0 ? gemm_tc2fh_kernel<64, false> : nullptr;
    else return nullptr;
  }
  if (kind == 4) {
    switch (v) {
      case 0: return gemm_tc2f_kernel<BN, OUT_F32, false>;
      case 1: return gemm_tc2f_kernel<BN, OUT_F32, true>;
      case 2: return gemm_tc2f_kernel<BN, OUT_SPLIT, false>;
      case 5: return gemm_tc2f_kernel<BN, OUT_SPLIT8, false>;
      case 7: return gemm_tc2f_kernel<BN, OUT_GEGLU8, false>;
      default: return nullptr;
    }
  }
  switch (v) {
    case 0: return gemm_tcf_kernel<BN, OUT_F32, false>;
    case 1: return gemm_tcf_kernel<BN, OUT_F32, true>;
    case 2: return gemm_tcf_kernel<BN, OUT_SPLIT, false>;
    case 5: return gemm_tcf_kernel<BN, OUT_SPLIT8, false>;
    default: return nullptr;
  }
}

static GemmKernel pick_f8_256(int kind, int v) {
  if (kind != 4) return nullptr;
  switch (v) {
    case 0: return gemm_tc2f256_kernel<OUT_F32, false>;
    case 1: return gemm_tc2f256_kernel<OUT_F32, true>;
    case 2: return gemm_tc2f256_kernel<OUT_SPLIT, false>;
    case 5: return gemm_tc2f256_kernel<OUT_SPLIT8, false>;
    default: return nullptr;
  }
}

// kernels with RAW segments: kind 2 (stacked split-bf16, BN = 128: transformer linears), kind 1 (BN = 256:
// GeGLU projection), kind 4 (f16f8, BN = 64 / 128) and kind 5 (f16f8 halo, BN = 64): ResBlock second conv + raw
// 1x1 skip segment
static GemmKernel pick_raw(int bn, int kind, int v) {
  if (kind == 2 && bn == 128) {
    switch (v) {
      case 0: return gemm_tc2s_raw_kernel<128, OUT_F32, 0>;
      case 6: return gemm_tc2s_raw_kernel<128, OUT_F32, 2>;
      case 2: return gemm_tc2s_raw_kernel<128, OUT_SPLIT, 0>;
      case 3: return gemm_tc2s_raw_kernel<128, OUT_SPLIT_T, 0>;
      default: return nullptr;
    }
  }
  if (kind == 1 && bn == 256)
    return v == 4 ? gemm_tc2_raw_kernel<256, OUT_GEGLU, 0> : v == 0 ? gemm_tc2_raw_kernel<256, OUT_F32, 0> : nullptr;
  if (kind == 5 && bn == 64) return v == 1 ? gemm_tc2fh_raw_kernel<64, 1> : v == 0 ? gemm_tc2fh_raw_kernel<64, 0> : nullptr;
  if (kind == 4 && bn == 64) return v == 1 ? gemm_tc2f_raw_kernel<64, OUT_F32, 1> : nullptr;
  if (kind == 4 && bn == 128) {
    switch (v) {
      case 1: return gemm_tc2f_raw_kernel<128, OUT_F32, 1>;
      case 5: return gemm_tc2f_raw_kernel<128, OUT_SPLIT8, 0>;
      default: return nullptr;
    }
  }
  return nullptr;
}

static GemmKernel pick_kernel(int bn, int kind, int v, bool raw = false) {
  if (raw) return pick_raw(bn, kind, v);
  if (kind >= 4)
    return bn == 64 ? pick_f8<64>(kind, v) : bn == 128 ? pick_f8<128>(kind, v) : bn == 256 ? pick_f8_256(kind, v) : nullptr;
  if (kind == 3) {  // halo stages: BN = 64, fp32 output (with / without statistics)
    if (bn != 64 || v > 1) return nullptr;
    return v == 1 ? gemm_tc2h_kernel<64, true> : gemm_tc2h_kernel<64, false>;
  }
  switch (bn) {
    case 64: return kind == 2 ? pick_variant<64, 2>(v) : kind == 1 ? pick_variant<64, 1>(v) : pick_variant<64, 0>(v);
    case 128: return kind == 2 ? pick_variant<128, 2>(v) : kind == 1 ? pick_variant<128, 1>(v) : pick_variant<128, 0>(v);
    case 256: return kind == 2 ? nullptr : kind == 1 ? pick_variant<256, 1>(v) : pick_variant<256, 0>(v);
    default: return nullptr;
  }
}

cudaError_t gemm_init_attrs() {
  for (int bn : {64, 128, 256})
    for (int kind = 0; kind < 7; ++kind)
      for (int v = 0; v < 9; ++v)
        for (int raw = 0; raw < 2; ++raw) {
          GemmKernel k = pick_kernel(bn, kind, v, raw != 0);
          if (!k) continue;
          cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(k),
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
          if (e != cudaSuccess) return e;
        }
  return cudaSuccess;
}

static int gemm_kind(const GemmParams& p, int bn) {
  return p.f8 ? (p.two_cta ? (p.halo ? 5 : 4) : 6)
              : p.two_cta ? (p.halo ? 3 : (p.stack && bn <= 128) ? 2 : 1) : 0;
}
static int gemm_variant(const GemmParams& p) {
  switch (p.mode) {
    case OUT_F32: return p.stats ? 1 : p.rowstats ? 6 : 0;
    case OUT_SPLIT: return 2;
    case OUT_SPLIT_T: return 3;
    case OUT_SPLIT8: return 5;
    case OUT_GEGLU8: return 7;
    case OUT_QKV: return 8;
    default: return 4;
  }
}
bool gemm_kernel_available(const GemmParams& p, int bn) {
  return pick_kernel(bn, gemm_kind(p, bn), gemm_variant(p), p.raw != 0) != nullptr;
}

cudaError_t launch_gemm(const GemmParams& p, int bn, int num_ctas, cudaStream_t stream) {
  const int v = gemm_variant(p);
  const int kind = gemm_kind(p, bn);
  GemmKernel k = pick_kernel(bn, kind, v, p.raw != 0);
  if (!k) return cudaErrorInvalidValue;
  if (p.raw && !p.two_cta) return cudaErrorInvalidValue;
  const bool raw_kernel = p.raw != 0;
  const int nthreads = raw_kernel ? GEMM_RAW_THREADS : GEMM_THREADS;
  if (p.two_cta) {
    const int smem = p.nstages * (p.halo ? gemm_stage_bytes2_halo(bn) : gemm_stage_bytes2(bn)) + 1024 +
                     gemm_epilogue_smem_bytes(bn);
    const long long pairs = static_cast<long long>(p.n_tiles) * (p.m_tiles / 2) * p.z_count;
    const long long max_clusters = num_ctas / 2;
    const unsigned grid = 2u * static_cast<unsigned>(pairs < max_clusters ? pairs : max_clusters);
    return launch_pdl(k, dim3(grid), dim3(nthreads), smem, stream, p);
  }
  const int smem = gemm_smem_bytes(bn, p.nstages) + gemm_epilogue_smem_bytes(bn);
  const long long total = static_cast<long long>(p.n_tiles) * p.m_tiles * p.z_count;
  const unsigned grid = static_cast<unsigned>(total < num_ctas ? total : num_ctas);
  return launch_pdl(k, dim3(grid), dim3(GEMM_THREADS), smem, stream, p);
}

#endif  // PF_GEMM_NO_DISPATCH

}  // namespace pf
