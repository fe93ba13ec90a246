// Fused tcgen05 attention core (see attn_tc.cu): parameter block shared by host and device.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace pf {

constexpr int ATTN_THREADS = 576;  // warp0 TMA, warp1 MMA, warps 2..17 softmax / epilogue

struct alignas(64) AttnParams {
  const __nv_bfloat16* q_hi_ptr;  // split-bf16 queries [B*N, ldq], head h at columns qcol0 + h*64 (staged into TMEM)
  const __nv_bfloat16* q_lo_ptr;
  long long ldq;
  CUtensorMap k_hi, k_lo;  // split-bf16 keys,    2-D {ldk, B*Nk}, box {64, 64}
  CUtensorMap v_hi, v_lo;  // split-bf16 V^T,     2-D {Nk, B*heads*64}, box {64, 64}
  int B, heads, N, Nk;     // N % 128 == 0, Nk % 64 == 0, d_head == 64
  int qcol0, kcol0, ocol0; // column of head 0 inside the q / k / o row
  float scale_log2e;       // d_head^-0.5 * log2(e)
  __nv_bfloat16* o_hi;     // split-bf16 output [B*N, ldo], head h at columns ocol0 + h*64
  __nv_bfloat16* o_lo;
  long long ldo;
  // optional [B][2][heads][2] fp32: upper bounds of max_i |q_i|^2 and max_j |k_j|^2 per (sample, head), as two
  // 32-column halves each (written with atomicMax by the q|k|v projection's epilogue, gemm_tc.cu OUT_QKV).  With
  // it the kernel takes Q_max * K_max (Cauchy-Schwarz) as the softmax stabiliser and skips pass 1 whenever that
  // bound is small enough for fp32 (see attn_tc.cu); null = always two passes
  const float* qknorm;
};

int attn_smem_bytes();
cudaError_t attn_init_attrs();
cudaError_t launch_attn(const AttnParams& p, int num_ctas, cudaStream_t stream);

}  // namespace pf
