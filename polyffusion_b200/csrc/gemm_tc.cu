// tcgen05 implicit-GEMM kernel (bf16x3 split precision), see gemm_tc.cuh.
//
// Roles (192 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA
// issuer (one elected lane), warps 2..5 = epilogue (TMEM -> registers -> global).
// Pipelines: smem ring full[]/empty[] (TMA <-> MMA) and tmem_full[]/tmem_empty[] (MMA <-> epilogue).
// Persistent: grid = min(#tiles, #SMs); each CTA walks tiles (n fastest, then m, then z) with a
// static stride.  The TMEM accumulator is double buffered (2 x BN columns) so the epilogue of tile i
// overlaps the main loop of tile i+1.
#include "common.cuh"
#include "gemm_tc.cuh"

namespace pf {

constexpr int MAX_STAGES = 8;

struct TileCoord {
  int n0, tile, zb, zh, img, trem, x0, y0;
};

__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, long long t, int bn) {
  TileCoord c;
  const int nt = static_cast<int>(t % p.n_tiles);
  const long long r = t / p.n_tiles;
  c.tile = static_cast<int>(r % p.m_tiles);
  const int z = static_cast<int>(r / p.m_tiles);
  c.n0 = nt * bn;
  c.zb = z / p.zdiv;
  c.zh = z % p.zdiv;
  c.img = c.tile / p.tiles_per_img;
  c.trem = c.tile % p.tiles_per_img;
  c.x0 = (c.trem % p.tiles_x) * p.box_w;
  c.y0 = (c.trem / p.tiles_x) * p.box_h;
  return c;
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_s;

  constexpr int A_BYTES = GEMM_BM * 128;
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  constexpr uint32_t IDESC = umma_idesc_bf16(BN);
  constexpr int TMEM_COLS = 2 * BN;  // 128 / 256 / 512: power of two >= 32

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nstages = p.nstages;
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const long long total_tiles = static_cast<long long>(p.n_tiles) * p.m_tiles * p.z_count;

  int nkb = 0;
  for (int s = 0; s < p.nseg; ++s) nkb += p.seg[s].ntaps * p.seg[s].kb_per_tap;

  if (threadIdx.x == 0) {
    for (int i = 0; i < nstages; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full_bar[i]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[i]), 4);  // one arrive per epilogue warp
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
    tmem_alloc(smem_u32(&tmem_base_s), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int it = 0;
      for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t, BN);
        for (int s = 0; s < p.nseg; ++s) {
          const GemmSeg& sg = p.seg[s];
          const int acol = sg.a_col0 + tc.zb * sg.a_col_zb + tc.zh * sg.a_col_zh;
          const int aimg = (tc.img + tc.zb * sg.a_img_zb + tc.zh * sg.a_img_zh) * sg.img_mul;
          const int brow = sg.b_row0 + tc.zb * sg.b_row_zb + tc.zh * sg.b_row_zh + tc.n0;
          const int bcol = sg.b_col0 + tc.zb * sg.b_col_zb + tc.zh * sg.b_col_zh;
          for (int tp = 0; tp < sg.ntaps; ++tp) {
            const int ax = tc.x0 + sg.tap_dx[tp], ay = tc.y0 + sg.tap_dy[tp];
            const int ai = aimg + sg.tap_dq[tp];
            const int br = brow + tp * sg.b_tap_stride;
            for (int kb = 0; kb < sg.kb_per_tap; ++kb, ++it) {
              const int stage = it % nstages;
              const uint32_t ph = static_cast<uint32_t>(it / nstages) & 1u;
              mbar_wait(smem_u32(&empty_bar[stage]), ph ^ 1u);
              const uint32_t fb = smem_u32(&full_bar[stage]);
              mbar_expect_tx(fb, STAGE_BYTES);
              const uint32_t sa = ring + stage * STAGE_BYTES;
              tma_load_4d(sa, &sg.a_hi, fb, acol + kb * GEMM_BK, ax, ay, ai);
              tma_load_4d(sa + A_BYTES, &sg.a_lo, fb, acol + kb * GEMM_BK, ax, ay, ai);
              tma_load_2d(sa + 2 * A_BYTES, &sg.b_hi, fb, bcol + kb * GEMM_BK, br);
              tma_load_2d(sa + 2 * A_BYTES + B_BYTES, &sg.b_lo, fb, bcol + kb * GEMM_BK, br);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      int it = 0;
      int lt = 0;  // local tile counter
      for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
        const int as = lt & 1;
        const uint32_t aph = static_cast<uint32_t>(lt >> 1) & 1u;
        mbar_wait(smem_u32(&tmem_empty_bar[as]), aph ^ 1u);
        tc_fence_after();
        const uint32_t acc = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kbi = 0; kbi < nkb; ++kbi, ++it) {
          const int stage = it % nstages;
          const uint32_t ph = static_cast<uint32_t>(it / nstages) & 1u;
          mbar_wait(smem_u32(&full_bar[stage]), ph);
          tc_fence_after();
          const uint32_t sa = ring + stage * STAGE_BYTES;
          const uint64_t da_hi = umma_desc_sw128(sa);
          const uint64_t da_lo = umma_desc_sw128(sa + A_BYTES);
          const uint64_t db_hi = umma_desc_sw128(sa + 2 * A_BYTES);
          const uint64_t db_lo = umma_desc_sw128(sa + 2 * A_BYTES + B_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t ko = static_cast<uint64_t>(k * 2);  // 32 B per K=16 step (16 B units)
            umma_bf16(acc, da_lo + ko, db_hi + ko, IDESC, (kbi | k) != 0);
            umma_bf16(acc, da_hi + ko, db_lo + ko, IDESC, 1u);
            umma_bf16(acc, da_hi + ko, db_hi + ko, IDESC, 1u);
          }
          umma_commit(smem_u32(&empty_bar[stage]));
        }
        umma_commit(smem_u32(&tmem_full_bar[as]));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;       // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;  // row inside the tile
    int lt = 0;
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
      const TileCoord tc = decode_tile(p, t, BN);
      const int as = lt & 1;
      const uint32_t aph = static_cast<uint32_t>(lt >> 1) & 1u;
      mbar_wait(smem_u32(&tmem_full_bar[as]), aph);
      tc_fence_after();
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN);
      const long long m = static_cast<long long>(tc.tile) * GEMM_BM + r;
      const long long zoff = tc.zb * p.out_zb + tc.zh * p.out_zh;

      if (p.mode == OUT_F32) {
        float* orow = p.out + zoff + m * p.ldc + tc.n0;
        const float* av =
            p.addvec ? p.addvec + static_cast<long long>(tc.img) * p.addvec_ld + tc.n0 : nullptr;
        const float* rr = p.resid ? p.resid + m * p.ldr + tc.n0 : nullptr;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o;
            o.x = __uint_as_float(v[4 * j + 0]);
            o.y = __uint_as_float(v[4 * j + 1]);
            o.z = __uint_as_float(v[4 * j + 2]);
            o.w = __uint_as_float(v[4 * j + 3]);
            if (av) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(av + c) + j);
              o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
            }
            if (rr) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(rr + c) + j);
              o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
            }
            reinterpret_cast<float4*>(orow + c)[j] = o;
          }
        }
      } else if (p.mode == OUT_SPLIT) {
        __nv_bfloat16* oh = p.out_hi + zoff + m * p.ldc + tc.n0;
        __nv_bfloat16* ol = p.out_lo + zoff + m * p.ldc + tc.n0;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 h, l;
            split2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]), h.x, l.x);
            split2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]), h.y, l.y);
            split2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]), h.z, l.z);
            split2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]), h.w, l.w);
            reinterpret_cast<uint4*>(oh + c)[j] = h;
            reinterpret_cast<uint4*>(ol + c)[j] = l;
          }
        }
      } else {  // OUT_SPLIT_T: [img][n][token]; consecutive lanes -> consecutive tokens
        const long long tok = static_cast<long long>(tc.trem) * GEMM_BM + r;
        const long long base = zoff + static_cast<long long>(tc.img) * p.out_img + tok;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            __nv_bfloat16 h, l;
            split_bf16(__uint_as_float(v[j]), h, l);
            const long long o = base + static_cast<long long>(tc.n0 + c + j) * p.ldc;
            p.out_hi[o] = h;
            p.out_lo[o] = l;
          }
        }
      }
      // all of this warp's TMEM reads for the tile are complete (tmem_ld_wait above)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[as]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

cudaError_t gemm_init_attrs() {
  cudaError_t e;
  e = cudaFuncSetAttribute(gemm_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  return e;
}

cudaError_t launch_gemm(const GemmParams& p, int bn, int num_ctas, cudaStream_t stream) {
  const int smem = gemm_smem_bytes(bn, p.nstages);
  const long long total = static_cast<long long>(p.n_tiles) * p.m_tiles * p.z_count;
  const unsigned grid = static_cast<unsigned>(total < num_ctas ? total : num_ctas);
  switch (bn) {
    case 64: gemm_tc_kernel<64><<<grid, GEMM_THREADS, smem, stream>>>(p); break;
    case 128: gemm_tc_kernel<128><<<grid, GEMM_THREADS, smem, stream>>>(p); break;
    case 256: gemm_tc_kernel<256><<<grid, GEMM_THREADS, smem, stream>>>(p); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace pf
