// Fused attention core on tcgen05:  O = softmax(Q K^T * scale) V  per (sample, head), d_head = 64,
// split-bf16 ("bf16x3") operands, scores and probabilities never leave the SM.
// Replaces the reference's normal_attention (stable_diffusion/model/unet_attention.py:261-293):
// einsum -> scale -> softmax -> einsum, which materialises a [B, heads, N, Nk] fp32 tensor.
//
// One CTA handles 128 query rows of one (sample, head) at a time (persistent over work items) and
// walks the keys in blocks of 64 TWICE:
//   pass 1:  S~_j = Q_hi K_hi_j^T     -> running row maximum m~           (no rescaling later)
//   pass 2:  S_j in full precision, P_j = exp2((S_j - m~) * scale*log2e), l += rowsum(P_j), O += P_j V_j
// and finally O / l.  softmax is invariant to the shift, so pass 1 only needs a stabiliser close to
// the true maximum: ONE bf16 product (hi*hi, K_lo is not even loaded) instead of three.  Pass 2 and
// P V use the stacked-operand form of the split product (see gemm_tc.cu):
//   Q_hi x [K_hi ; K_lo] (N = 128)  +  Q_lo x K_hi (N = 64, accumulated onto columns [0, 64))
// so an S / O accumulator is 128 columns wide and the consumer adds its two halves.  Recomputing S
// removes the accumulator-rescaling path of online softmax entirely.
//
// Roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocator), warps 2..9 =
// softmax / epilogue (two warps per TMEM lane quarter, 32 key columns each).
// TMEM: S double buffer (2 x 128 cols) + O (128 cols).  smem: Q, K ring, V^T ring, P double buffer
// (written by the softmax warps in the UMMA K-major 128B-swizzled layout, consumed as the A operand).
#include "common.cuh"
#include "attn_tc.cuh"

namespace pf {

constexpr int AT_KS = 4;  // K ring stages
constexpr int AT_VS = 3;  // V ring stages
constexpr int AT_Q_BYTES = 128 * 128;       // 128 rows x 64 bf16
constexpr int AT_KV_BYTES = 64 * 128;       // 64 rows x 64 bf16
constexpr int AT_P_BYTES = 128 * 128;       // 128 rows x 64 keys bf16
constexpr int AT_OFF_Q = 0;                                   // hi, lo
constexpr int AT_OFF_K = AT_OFF_Q + 2 * AT_Q_BYTES;           // [KS][hi, lo]
constexpr int AT_OFF_V = AT_OFF_K + AT_KS * 2 * AT_KV_BYTES;  // [VS][hi, lo]
constexpr int AT_OFF_P = AT_OFF_V + AT_VS * 2 * AT_KV_BYTES;  // [2][hi, lo]
constexpr int AT_OFF_X = AT_OFF_P + 2 * 2 * AT_P_BYTES;       // float xch[2][128][2]
constexpr int AT_SMEM = AT_OFF_X + 2 * 128 * 2 * 4;

int attn_smem_bytes() { return AT_SMEM + 1024; }

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__global__ void __launch_bounds__(ATTN_THREADS, 1) attn_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, q_empty, o_full, o_empty;
  __shared__ __align__(8) uint64_t k_full[AT_KS], k_empty[AT_KS], v_full[AT_VS], v_empty[AT_VS];
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2], p_full[2], p_empty[2];
  __shared__ uint32_t tmem_base_s;

  constexpr uint32_t IDESC = umma_idesc_bf16(64);
  constexpr uint32_t IDESC2 = umma_idesc_bf16(128);
  constexpr int TMEM_COLS = 512;  // S0 [0,128) S1 [128,256) O [256,384)
  constexpr uint32_t O_COL = 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const int nqt = p.N / 128, nb = p.Nk / 64;
  const long long items = static_cast<long long>(p.B) * p.heads * nqt;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&q_full), 1);
    mbar_init(smem_u32(&q_empty), 1);
    mbar_init(smem_u32(&o_full), 1);
    mbar_init(smem_u32(&o_empty), 8);
    for (int i = 0; i < AT_KS; ++i) { mbar_init(smem_u32(&k_full[i]), 1); mbar_init(smem_u32(&k_empty[i]), 1); }
    for (int i = 0; i < AT_VS; ++i) { mbar_init(smem_u32(&v_full[i]), 1); mbar_init(smem_u32(&v_empty[i]), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&s_empty[i]), 8);
      mbar_init(smem_u32(&p_full[i]), 8);
      mbar_init(smem_u32(&p_empty[i]), 1);
    }
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.q_hi); tma_prefetch_desc(&p.q_lo);
    tma_prefetch_desc(&p.k_hi); tma_prefetch_desc(&p.k_lo);
    tma_prefetch_desc(&p.v_hi); tma_prefetch_desc(&p.v_lo);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_s), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t kc = 0, vc = 0, it = 0;
      for (long long w = blockIdx.x; w < items; w += gridDim.x, ++it) {
        const int qt = static_cast<int>(w % nqt);
        const int bh = static_cast<int>(w / nqt);
        const int b = bh / p.heads, h = bh % p.heads;
        mbar_wait(smem_u32(&q_empty), (it & 1u) ^ 1u);
        mbar_expect_tx(smem_u32(&q_full), 2 * AT_Q_BYTES);
        tma_load_2d(base + AT_OFF_Q, &p.q_hi, smem_u32(&q_full), p.qcol0 + h * 64, b * p.N + qt * 128);
        tma_load_2d(base + AT_OFF_Q + AT_Q_BYTES, &p.q_lo, smem_u32(&q_full), p.qcol0 + h * 64,
                    b * p.N + qt * 128);
        for (int pass = 0; pass < 2; ++pass) {
          for (int j = 0; j < nb; ++j) {
            {
              const uint32_t st = kc % AT_KS, ph = (kc / AT_KS) & 1u;
              mbar_wait(smem_u32(&k_empty[st]), ph ^ 1u);
              mbar_expect_tx(smem_u32(&k_full[st]), (pass == 1 ? 2 : 1) * AT_KV_BYTES);
              const uint32_t dst = base + AT_OFF_K + st * 2 * AT_KV_BYTES;
              tma_load_2d(dst, &p.k_hi, smem_u32(&k_full[st]), p.kcol0 + h * 64, b * p.Nk + j * 64);
              if (pass == 1)
                tma_load_2d(dst + AT_KV_BYTES, &p.k_lo, smem_u32(&k_full[st]), p.kcol0 + h * 64,
                            b * p.Nk + j * 64);
              ++kc;
            }
            if (pass == 1) {
              const uint32_t st = vc % AT_VS, ph = (vc / AT_VS) & 1u;
              mbar_wait(smem_u32(&v_empty[st]), ph ^ 1u);
              mbar_expect_tx(smem_u32(&v_full[st]), 2 * AT_KV_BYTES);
              const uint32_t dst = base + AT_OFF_V + st * 2 * AT_KV_BYTES;
              tma_load_2d(dst, &p.v_hi, smem_u32(&v_full[st]), j * 64, bh * 64);
              tma_load_2d(dst + AT_KV_BYTES, &p.v_lo, smem_u32(&v_full[st]), j * 64, bh * 64);
              ++vc;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      uint32_t kc = 0, vc = 0, sc = 0, pc = 0, it = 0;
      const uint64_t dq_hi = umma_desc_sw128(base + AT_OFF_Q);
      const uint64_t dq_lo = umma_desc_sw128(base + AT_OFF_Q + AT_Q_BYTES);
      auto issue_s = [&](bool full) {
        const uint32_t st = kc % AT_KS, kph = (kc / AT_KS) & 1u;
        const uint32_t sb = sc & 1u, sph = (sc >> 1) & 1u;
        mbar_wait(smem_u32(&k_full[st]), kph);
        mbar_wait(smem_u32(&s_empty[sb]), sph ^ 1u);
        tc_fence_after();
        // K_hi (64 rows) is followed by K_lo (64 rows) in the stage: one 128-row B operand
        const uint64_t dk_hi = umma_desc_sw128(base + AT_OFF_K + st * 2 * AT_KV_BYTES);
        const uint32_t acc = tmem_base + sb * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ko = static_cast<uint64_t>(k * 2);
          if (full) {
            umma_bf16(acc, dq_hi + ko, dk_hi + ko, IDESC2, k != 0);
            umma_bf16(acc, dq_lo + ko, dk_hi + ko, IDESC, 1u);
          } else {
            umma_bf16(acc, dq_hi + ko, dk_hi + ko, IDESC, k != 0);
          }
        }
        umma_commit(smem_u32(&k_empty[st]));
        umma_commit(smem_u32(&s_full[sb]));
        ++kc;
        ++sc;
      };
      for (long long w = blockIdx.x; w < items; w += gridDim.x, ++it) {
        mbar_wait(smem_u32(&q_full), it & 1u);
        tc_fence_after();
        for (int j = 0; j < nb; ++j) issue_s(false);  // pass 1
        issue_s(true);                                // pass 2, block 0
        for (int j = 0; j < nb; ++j) {
          if (j + 1 < nb) issue_s(true);
          const uint32_t pb = pc & 1u, pph = (pc >> 1) & 1u;
          const uint32_t st = vc % AT_VS, vph = (vc / AT_VS) & 1u;
          mbar_wait(smem_u32(&p_full[pb]), pph);
          mbar_wait(smem_u32(&v_full[st]), vph);
          if (j == 0) mbar_wait(smem_u32(&o_empty), (it & 1u) ^ 1u);
          tc_fence_after();
          const uint64_t dp_hi = umma_desc_sw128(base + AT_OFF_P + pb * 2 * AT_P_BYTES);
          const uint64_t dp_lo = umma_desc_sw128(base + AT_OFF_P + pb * 2 * AT_P_BYTES + AT_P_BYTES);
          // V^T_hi (64 rows = d) followed by V^T_lo: one 128-row B operand
          const uint64_t dv_hi = umma_desc_sw128(base + AT_OFF_V + st * 2 * AT_KV_BYTES);
          const uint32_t acc = tmem_base + O_COL;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ko = static_cast<uint64_t>(k * 2);
            umma_bf16(acc, dp_hi + ko, dv_hi + ko, IDESC2, (j | k) != 0);
            umma_bf16(acc, dp_lo + ko, dv_hi + ko, IDESC, 1u);
          }
          umma_commit(smem_u32(&p_empty[pb]));
          umma_commit(smem_u32(&v_empty[st]));
          ++pc;
          ++vc;
        }
        umma_commit(smem_u32(&o_full));
        umma_commit(smem_u32(&q_empty));
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue warps
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    float* xch = reinterpret_cast<float*>(gbase + AT_OFF_X);
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float c = p.scale_log2e;
    uint32_t sc = 0, pc = 0, it = 0;
    for (long long w = blockIdx.x; w < items; w += gridDim.x, ++it) {
      const int qt = static_cast<int>(w % nqt);
      const int bh = static_cast<int>(w / nqt);
      const int b = bh / p.heads, h = bh % p.heads;
      // ---- pass 1: row maximum
      float mx = -INFINITY;
      for (int j = 0; j < nb; ++j, ++sc) {
        const uint32_t sb = sc & 1u, sph = (sc >> 1) & 1u;
        mbar_wait(smem_u32(&s_full[sb]), sph);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(lane_addr + sb * 128 + half * 32, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_empty[sb]));
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      xch[row * 2 + half] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float m = fmaxf(xch[row * 2], xch[row * 2 + 1]);
      const float mc = m * c;
      // ---- pass 2: probabilities -> P (smem, UMMA layout), row sums
      float sum = 0.f;
      for (int j = 0; j < nb; ++j, ++sc, ++pc) {
        const uint32_t sb = sc & 1u, sph = (sc >> 1) & 1u;
        const uint32_t pb = pc & 1u, pph = (pc >> 1) & 1u;
        mbar_wait(smem_u32(&s_full[sb]), sph);
        tc_fence_after();
        uint32_t v[32], v2[32];
        tmem_ld32(lane_addr + sb * 128 + half * 32, v);
        tmem_ld32(lane_addr + sb * 128 + 64 + half * 32, v2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_empty[sb]));
        uint32_t ph[16], pl[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float s0 = __uint_as_float(v[2 * i]) + __uint_as_float(v2[2 * i]);
          const float s1 = __uint_as_float(v[2 * i + 1]) + __uint_as_float(v2[2 * i + 1]);
          const float p0 = fast_ex2(fmaf(s0, c, -mc));
          const float p1 = fast_ex2(fmaf(s1, c, -mc));
          sum += p0 + p1;
          split2(p0, p1, ph[i], pl[i]);
        }
        mbar_wait(smem_u32(&p_empty[pb]), pph ^ 1u);
        uint8_t* prow = gbase + AT_OFF_P + pb * 2 * AT_P_BYTES + row * 128;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int chunk = ((half * 4 + i) ^ (row & 7)) << 4;
          *reinterpret_cast<uint4*>(prow + chunk) = make_uint4(ph[4 * i], ph[4 * i + 1], ph[4 * i + 2], ph[4 * i + 3]);
          *reinterpret_cast<uint4*>(prow + AT_P_BYTES + chunk) =
              make_uint4(pl[4 * i], pl[4 * i + 1], pl[4 * i + 2], pl[4 * i + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&p_full[pb]));
      }
      xch[256 + row * 2 + half] = sum;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float inv = 1.0f / (xch[256 + row * 2] + xch[256 + row * 2 + 1]);
      // ---- epilogue: O / l -> split-bf16 [B*N, ldo] at columns h*64 + half*32
      mbar_wait(smem_u32(&o_full), it & 1u);
      tc_fence_after();
      uint32_t o[32];
      {
        uint32_t o2[32];
        tmem_ld32(lane_addr + O_COL + half * 32, o);
        tmem_ld32(lane_addr + O_COL + 64 + half * 32, o2);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) + __uint_as_float(o2[i]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&o_empty));
      const long long off =
          (static_cast<long long>(b) * p.N + qt * 128 + row) * p.ldo + p.ocol0 + h * 64 + half * 32;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 hh, ll;
        split2(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv, hh.x, ll.x);
        split2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv, hh.y, ll.y);
        split2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv, hh.z, ll.z);
        split2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv, hh.w, ll.w);
        *reinterpret_cast<uint4*>(p.o_hi + off + 8 * i) = hh;
        *reinterpret_cast<uint4*>(p.o_lo + off + 8 * i) = ll;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

cudaError_t attn_init_attrs() {
  return cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes());
}

cudaError_t launch_attn(const AttnParams& p, int num_ctas, cudaStream_t stream) {
  const long long items = static_cast<long long>(p.B) * p.heads * (p.N / 128);
  const unsigned grid = static_cast<unsigned>(items < num_ctas ? items : num_ctas);
  return launch_pdl(attn_tc_kernel, dim3(grid), dim3(ATTN_THREADS), attn_smem_bytes(), stream, p);
}

}  // namespace pf
