// Probe: does a tcgen05 K-major SWIZZLE_128B shared-memory descriptor accept a start address that is
// shifted by a whole number of 128-byte rows (NOT a multiple of the 1024-byte swizzle atom)?
// This is what a 3x3 convolution needs to serve the dx = -1, 0, +1 taps from ONE (W + 2)-pixel halo
// row held in shared memory.  The A tile is written with the swizzle pattern of its ABSOLUTE row
// index (what TMA does for a 1024-byte-aligned destination); the descriptor then starts s rows in,
// with the "matrix base offset" field (bits 49-51) either 0 or (start >> 7) & 7.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/probe tools/probe_umma_shift.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../polyffusion_b200/csrc/common.cuh"

using namespace pf;

constexpr int ROWS_A = 144;

__global__ void __launch_bounds__(128) probe_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b,
                                                    float* d, int shift, int bo_mode) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sa = g;                    // ROWS_A x 128 B
  uint8_t* sb = g + ROWS_A * 128;     // 64 x 128 B   (ROWS_A * 128 = 18432 = 18 * 1024: aligned)
  for (int i = threadIdx.x; i < ROWS_A * 8; i += 128) {
    const int r = i >> 3, j = i & 7;
    *reinterpret_cast<uint4*>(sa + r * 128 + ((j ^ (r & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(a + r * 64 + j * 8);
  }
  for (int i = threadIdx.x; i < 64 * 8; i += 128) {
    const int r = i >> 3, j = i & 7;
    *reinterpret_cast<uint4*>(sb + r * 128 + ((j ^ (r & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(b + r * 64 + j * 8);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_fence_init();
  }
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_base_s), 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 0 && elect_one()) {
    const uint32_t start = base + shift * 128;
    uint64_t da = umma_desc_sw128(start);
    if (bo_mode == 1) da |= static_cast<uint64_t>((start >> 7) & 7u) << 49;
    const uint64_t db = umma_desc_sw128(base + ROWS_A * 128);
    for (int k = 0; k < 4; ++k)
      umma_bf16(tmem_base, da + 2 * k, db + 2 * k, umma_idesc_bf16(64), k != 0);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  uint32_t v[32];
  for (int c = 0; c < 64; c += 32) {
    tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) d[(threadIdx.x) * 64 + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

int main() {
  std::vector<float> ha(ROWS_A * 64), hb(64 * 64);
  std::vector<__nv_bfloat16> ba(ROWS_A * 64), bb(64 * 64);
  for (int r = 0; r < ROWS_A; ++r)
    for (int k = 0; k < 64; ++k) {
      ha[r * 64 + k] = static_cast<float>((r * 7 + k * 3 + (r / 8) * 5) % 13 - 6);
      ba[r * 64 + k] = __float2bfloat16(ha[r * 64 + k]);
    }
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 64; ++k) {
      hb[n * 64 + k] = static_cast<float>((n * 5 + k) % 11 - 5);
      bb[n * 64 + k] = __float2bfloat16(hb[n * 64 + k]);
    }
  __nv_bfloat16 *da, *db;
  float* dd;
  cudaMalloc(&da, ba.size() * 2);
  cudaMalloc(&db, bb.size() * 2);
  cudaMalloc(&dd, 128 * 64 * 4);
  cudaMemcpy(da, ba.data(), ba.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, bb.data(), bb.size() * 2, cudaMemcpyHostToDevice);
  const int smem = ROWS_A * 128 + 64 * 128 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> hd(128 * 64);
  for (int mode = 0; mode < 2; ++mode)
    for (int s = 0; s <= 10; ++s) {
      cudaMemset(dd, 0, 128 * 64 * 4);
      probe_kernel<<<1, 128, smem>>>(da, db, dd, s, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("mode %d shift %d: CUDA error %s\n", mode, s, cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      int bad = 0, first_bad = -1;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          float ref = 0;
          for (int k = 0; k < 64; ++k) ref += ha[(m + s) * 64 + k] * hb[n * 64 + k];
          const double err = fabs(ref - hd[m * 64 + n]);
          if (err > 1e-3) {
            if (first_bad < 0) first_bad = m;
            ++bad;
          }
          if (err > maxerr) maxerr = err;
        }
      printf("base_offset_mode %d shift %2d rows: max err %.1f, bad %d / 8192, first bad row %d  %s\n", mode,
             s, maxerr, bad, first_bad, bad == 0 ? "OK" : "MISMATCH");
    }
  return 0;
}
