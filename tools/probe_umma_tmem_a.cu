// Probe: tcgen05.mma with the A operand in TENSOR MEMORY (M = 128, kind::f16, bf16), written by the
// threads themselves with tcgen05.st.32x32b (lane = row, one 32-bit column = two consecutive K
// elements).  This is what lets the attention kernel keep Q and the probabilities P out of shared
// memory.  Variants: packing order of the two bf16 in a column (low half = even k, or odd k).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/probe_umma_tmem_a tools/probe_umma_tmem_a.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../polyffusion_b200/csrc/common.cuh"

using namespace pf;

__device__ __forceinline__ void umma_bf16_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

__global__ void __launch_bounds__(128) probe_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b, float* d,
                                                    int swap_pack, int a_col) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sb = smem_raw + (base - smem_u32(smem_raw));
  for (int i = threadIdx.x; i < 64 * 8; i += 128) {
    const int r = i >> 3, j = i & 7;
    *reinterpret_cast<uint4*>(sb + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(b + r * 64 + j * 8);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_fence_init();
  }
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_base_s), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  {
    // thread = row: 64 bf16 of A -> 32 packed columns at a_col
    uint32_t v[32];
    const int row = threadIdx.x;
    for (int c = 0; c < 32; ++c) {
      const uint32_t e0 = __bfloat16_as_ushort(a[row * 64 + 2 * c]);
      const uint32_t e1 = __bfloat16_as_ushort(a[row * 64 + 2 * c + 1]);
      v[c] = swap_pack ? (e1 | (e0 << 16)) : (e0 | (e1 << 16));
    }
    tmem_st32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + a_col, v);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0 && elect_one()) {
    const uint64_t db = umma_desc_sw128(base);
    for (int k = 0; k < 4; ++k)  // K = 16 per instruction = 8 TMEM columns of A, 32 bytes of B
      umma_bf16_ta(tmem_base, tmem_base + a_col + 8 * k, db + 2 * k, umma_idesc_bf16(64), k != 0);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  uint32_t v[32];
  for (int c = 0; c < 64; c += 32) {
    tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) d[threadIdx.x * 64 + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

int main() {
  std::vector<float> ha(128 * 64), hb(64 * 64);
  std::vector<__nv_bfloat16> ba(128 * 64), bb(64 * 64);
  for (int r = 0; r < 128; ++r)
    for (int k = 0; k < 64; ++k) {
      ha[r * 64 + k] = static_cast<float>((r * 7 + k * 3 + (r / 8) * 5) % 13 - 6);
      ba[r * 64 + k] = __float2bfloat16(ha[r * 64 + k]);
    }
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 64; ++k) {
      hb[n * 64 + k] = static_cast<float>((n * 5 + k * k) % 11 - 5);
      bb[n * 64 + k] = __float2bfloat16(hb[n * 64 + k]);
    }
  __nv_bfloat16 *da, *db;
  float* dd;
  cudaMalloc(&da, ba.size() * 2);
  cudaMalloc(&db, bb.size() * 2);
  cudaMalloc(&dd, 128 * 64 * 4);
  cudaMemcpy(da, ba.data(), ba.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, bb.data(), bb.size() * 2, cudaMemcpyHostToDevice);
  const int smem = 64 * 128 + 1024;
  std::vector<float> hd(128 * 64);
  for (int swap_pack = 0; swap_pack < 2; ++swap_pack)
    for (int a_col : {64, 128, 200}) {
      cudaMemset(dd, 0, 128 * 64 * 4);
      probe_kernel<<<1, 128, smem>>>(da, db, dd, swap_pack, a_col);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("swap %d a_col %d: CUDA error %s\n", swap_pack, a_col, cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          float ref = 0;
          for (int k = 0; k < 64; ++k) ref += ha[m * 64 + k] * hb[n * 64 + k];
          const double err = fabs(ref - hd[m * 64 + n]);
          if (err > 1e-3) ++bad;
          if (err > maxerr) maxerr = err;
        }
      printf("A in TMEM at column %3d, pack %s: max err %.1f, bad %d / 8192  %s\n", a_col,
             swap_pack ? "(odd k low)" : "(even k low)", maxerr, bad, bad == 0 ? "OK" : "MISMATCH");
    }
  return 0;
}
