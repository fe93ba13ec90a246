// Device-side primitives for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM).
// Thin inline-PTX wrappers; no CUTLASS dependency.  Everything here is cta_group::1.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include <cstdlib>
#include <utility>

namespace pf {

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args&&... args) {
  // measured on B200: no gain on this step (22.24 ms with vs 22.00 ms without), so opt-in only
  static const bool enabled = std::getenv("PF_PDL") != nullptr;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- programmatic dependent launch
// With PF_PDL=1 every hot-path kernel is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: its
// CTAs may be scheduled while the tail of the previous kernel is still running, do their private
// prologue (barrier init, TMEM alloc, descriptor prefetch), and block in pdl_wait() until the previous
// kernel has completed and flushed.  Rule: nothing produced by an earlier kernel is read, and no
// global memory is written, before pdl_wait().
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* desc, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* desc, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32.  Issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (lane i of the warp = TMEM lane
// lane_base + i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- cta_group::2 (CTA pair) variants
// In a 2-CTA cluster the shared::cta address of the odd CTA has bit 24 set; clearing it names the
// same object in the even ("leader") CTA as a shared::cluster address.
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar_local) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_local & PEER_MASK) : "memory");
}
// (cluster-scope release arrive / acquire wait variants were tried for the RAW conversion warps and dropped: both
// lower to MEMBAR.ALL.GPU; the data hand-off uses fence.proxy.async.shared::cta + a plain remote arrive, gemm_tc.cu)
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const void* desc, uint32_t bar_local, int c0,
                                             int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_local & PEER_MASK), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const void* desc, uint32_t bar_local, int c0,
                                             int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_local & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// M = 256 (128 rows per CTA) x N, issued by one thread of the leader CTA
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once) on the barrier at this offset in BOTH CTAs of the pair when prior MMAs complete
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar_local) {
  const unsigned short mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      :
      : "r"(bar_local), "h"(mask)
      : "memory");
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_m256(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(256 >> 4) << 24);
}

// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B, rows of 64 bf16 (128 B), 8-row
// groups 1024 B apart.  Bit layout follows the sm_100 UMMA descriptor: start>>4 [0,14),
// LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout_type=2 (SW128) [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;            // LBO (ignored for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO
  d |= static_cast<uint64_t>(1) << 46;            // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
  return d;
}
// Instruction descriptor: bf16 A/B (K-major both), fp32 accumulate, M=128, N=n.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(128 >> 4) << 24);
}

// ---------------------------------------------------------------- fp16 + fp8 ("f16f8") operands
// Second precision scheme of the GEMM kernels (gemm_tc.cuh): x = h16 + lo with h16 = rn_fp16(x); the
// product a*w is accumulated as
//     h16(a) * h16(w)                                   one kind::f16 MMA (K = 16) into accumulator 0
//   + [h8(a) | l8(a)] . [l8(w) | h8(w)] * 2^-17         one kind::f8f6f4 MMA (K = 32) into accumulator 1
// where h8 = e4m3(h16), l8 = e4m3(lo * 2^11) for activations and h8 = e4m3(h16 * 2^6),
// l8 = e4m3(lo * 2^17) for weights: the 128-byte shared-memory row of the fp8 tile holds the 64 h8
// bytes of a k-block followed by its 64 l8 bytes (weights: l8 then h8), so ONE pass over that row
// evaluates both cross terms lo*hi + hi*lo.  fp8 MMAs run at twice the bf16 rate: 2 tensor-time units
// per product instead of 3.  Instruction-descriptor format fields are 0 for both F16 and E4M3.
constexpr float F8_ACT_LO_SCALE = 2048.f;        // 2^11
constexpr float F8_W_HI_SCALE = 64.f;            // 2^6
constexpr float F8_W_LO_SCALE = 131072.f;        // 2^17
constexpr float F8_CROSS_SCALE = 1.f / 131072.f; // 2^-17 = 1 / (act_lo * w_hi) = 1 / (act_hi * w_lo)
__host__ __device__ constexpr uint32_t umma_idesc_fmt0(int n, int m) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 4 floats -> 4 packed e4m3 bytes (round to nearest even, saturate to +-448)
__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {
  unsigned short lo, hi;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %2, %1;" : "=h"(lo) : "f"(a), "f"(b));
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %2, %1;" : "=h"(hi) : "f"(c), "f"(d));
  return static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
}
// 2 floats -> packed fp16 pair (low half = first) and the two residuals x - h16(x)
__device__ __forceinline__ uint32_t split_h16x2(float a, float b, float& ra, float& rb) {
  const __half2 h = __floats2half2_rn(a, b);
  ra = a - __low2float(h);
  rb = b - __high2float(h);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// 4 consecutive channels -> h16 (2 words), h8 (1 word), l8 (1 word); lo_scale / hi_scale as above
__device__ __forceinline__ void split_f8x4(float a, float b, float c, float d, float hi_scale, float lo_scale,
                                           uint2& h16, uint32_t& h8, uint32_t& l8) {
  float ra, rb, rc, rd;
  h16.x = split_h16x2(a, b, ra, rb);
  h16.y = split_h16x2(c, d, rc, rd);
  h8 = pack_e4m3x4((a - ra) * hi_scale, (b - rb) * hi_scale, (c - rc) * hi_scale, (d - rd) * hi_scale);
  l8 = pack_e4m3x4(ra * lo_scale, rb * lo_scale, rc * lo_scale, rd * lo_scale);
}

// Branch-free MUFU wrappers (ex2.approx / rcp.approx: ~2^-22 relative error, no slow paths)
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// GELU(erf) with the Abramowitz-Stegun 7.1.26 rational approximation of erf (|abs err| <= 1.5e-7,
// i.e. at fp32 resolution) instead of erff: ~4x fewer instructions in the GeGLU epilogue.
__device__ __forceinline__ float gelu_erf_fast(float g) {
  const float z = fabsf(g) * 0.70710678118654752440f;
  const float t = fast_rcp(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float erf_abs = 1.0f - poly * fast_ex2(-1.4426950408889634f * z * z);
  const float erf_v = copysignf(erf_abs, g);
  return 0.5f * g * (1.0f + erf_v);
}

// ---------------------------------------------------------------- bf16 hi/lo split
// x ~= hi + lo with hi = rn_bf16(x), lo = rn_bf16(x - hi): 16 mantissa bits kept.
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(x, 0.f);
  const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
  const __nv_bfloat162 l2 = __floats2bfloat162_rn(x - __uint_as_float(hb << 16), 0.f);
  const uint32_t lb = *reinterpret_cast<const uint32_t*>(&l2);
  hi = __ushort_as_bfloat16(static_cast<unsigned short>(hb & 0xffffu));
  lo = __ushort_as_bfloat16(static_cast<unsigned short>(lb & 0xffffu));
}
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return static_cast<uint32_t>(__bfloat16_as_ushort(a)) |
         (static_cast<uint32_t>(__bfloat16_as_ushort(b)) << 16);
}
// split 2 floats -> packed hi pair and packed lo pair (low 16 bits = first element).
// Uses the packed cvt.rn.bf16x2.f32 (F2FP, full-rate ALU op) rather than two scalar F2F conversions.
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h2);
  const float ah = __uint_as_float(hi << 16);
  const float bh = __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - ah, b - bh);
  lo = *reinterpret_cast<const uint32_t*>(&l2);
}

}  // namespace pf
