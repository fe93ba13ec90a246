#include "host_util.h"

#include <cstdlib>
#include <mutex>

namespace pf {

// cuTensorMapEncodeTiled is resolved through the runtime so the library has no link-time
// dependency on libcuda (it must load on a CPU-only box for the symbol-export tests).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  if (!fn) fail("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  return fn;
}

CUtensorMap make_map_4d(const void* ptr, int C, int W, int H, int N, int box_w, int box_h) {
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims,
                            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    fail("cuTensorMapEncodeTiled(4d C=%d W=%d H=%d N=%d box=%dx%d) failed: %d", C, W, H, N, box_w,
         box_h, (int)r);
  return m;
}

CUtensorMap make_map_4d_f32(const void* ptr, int C, int W, int H, int N, int box_w, int box_h) {
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr), dims, strides,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    fail("cuTensorMapEncodeTiled(4d f32 C=%d W=%d H=%d N=%d box=%dx%d) failed: %d", C, W, H, N, box_w,
         box_h, (int)r);
  return m;
}

CUtensorMap make_map_2d(const void* ptr, long long K, long long rows, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims,
                            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    fail("cuTensorMapEncodeTiled(2d K=%lld rows=%lld box=%d) failed: %d", K, rows, box_rows, (int)r);
  return m;
}

CUtensorMap make_map_4d_u8(const void* ptr, int C, int W, int H, int N, int box_w, int box_h) {
  CUtensorMap m;
  const cuuint64_t rb = (cuuint64_t)C * 2;  // bytes per pixel
  cuuint64_t dims[4] = {rb, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {rb, (cuuint64_t)W * rb, (cuuint64_t)H * W * rb};
  cuuint32_t box[4] = {128, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(ptr), dims, strides,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    fail("cuTensorMapEncodeTiled(4d u8 C=%d W=%d H=%d N=%d box=%dx%d) failed: %d", C, W, H, N, box_w,
         box_h, (int)r);
  return m;
}

CUtensorMap make_map_2d_u8(const void* ptr, long long K, long long rows, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)K * 2, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {128, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    fail("cuTensorMapEncodeTiled(2d u8 K=%lld rows=%lld box=%d) failed: %d", K, rows, box_rows, (int)r);
  return m;
}

// ------------------------------------------------------------------ arena
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void* Arena::alloc(size_t bytes) {
  bytes = align_up(bytes ? bytes : 1, 1024);
  for (auto it = free_.begin(); it != free_.end(); ++it) {
    if (it->second >= bytes) {
      const size_t off = it->first, sz = it->second;
      free_.erase(it);
      if (sz > bytes) free_[off + bytes] = sz - bytes;
      live_[off] = bytes;
      return base_ + off;
    }
  }
  const size_t off = top_;
  top_ += bytes;
  if (top_ > peak_) peak_ = top_;
  live_[off] = bytes;
  return base_ + off;
}

void Arena::free(void* p) {
  if (!p) return;
  const size_t off = static_cast<size_t>(static_cast<char*>(p) - base_);
  auto it = live_.find(off);
  if (it == live_.end()) fail("arena: free of unknown block");
  size_t sz = it->second;
  live_.erase(it);
  size_t start = off;
  // coalesce with neighbours
  auto nx = free_.find(off + sz);
  if (nx != free_.end()) {
    sz += nx->second;
    free_.erase(nx);
  }
  auto pv = free_.lower_bound(off);
  if (pv != free_.begin()) {
    --pv;
    if (pv->first + pv->second == off) {
      start = pv->first;
      sz += pv->second;
      free_.erase(pv);
    }
  }
  if (start + sz == top_) {
    top_ = start;
  } else {
    free_[start] = sz;
  }
}

// ------------------------------------------------------------------ tap tables
void fill_taps_3x3(GemmSeg& sg) {
  sg.ntaps = 9;
  sg.img_mul = 1;
  for (int t = 0; t < 9; ++t) {
    sg.tap_dy[t] = static_cast<signed char>(t / 3 - 1);
    sg.tap_dx[t] = static_cast<signed char>(t % 3 - 1);
    sg.tap_dq[t] = 0;
  }
}

void fill_taps_3x3_s2d(GemmSeg& sg) {
  // output (oy, ox) reads input (2*oy + ky - 1, 2*ox + kx - 1).  Row 2*oy-1 is odd-parity plane row
  // oy-1, row 2*oy is even-parity plane row oy, row 2*oy+1 is odd-parity plane row oy.
  sg.ntaps = 9;
  sg.img_mul = 4;
  for (int t = 0; t < 9; ++t) {
    const int ky = t / 3, kx = t % 3;
    const int py = (ky != 1), px = (kx != 1);
    sg.tap_dy[t] = static_cast<signed char>(ky == 0 ? -1 : 0);
    sg.tap_dx[t] = static_cast<signed char>(kx == 0 ? -1 : 0);
    sg.tap_dq[t] = static_cast<signed char>(py * 2 + px);
  }
}

void fill_taps_1x1(GemmSeg& sg) {
  sg.ntaps = 1;
  sg.img_mul = 1;
  sg.tap_dx[0] = sg.tap_dy[0] = sg.tap_dq[0] = 0;
}

static int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v && *v ? std::atoi(v) : dflt;
}

int gemm_default_stages(int bn) {
  int st = (200 * 1024 - 1024) / gemm_stage_bytes(bn);
  const int cap = env_int("PF_GEMM_MAX_STAGES", 6);
  if (st > cap) st = cap;
  if (st < 1) st = 1;
  return st;
}

int gemm_default_stages2(int bn) {
  // cta_group::2 stage = 32 KB (A) + bn * 128 B (half of B); 32 KB epilogue staging + 1 KB slack
  int st = (226 * 1024 - 1024 - 32768) / gemm_stage_bytes2(bn);
  const int cap = env_int("PF_GEMM_MAX_STAGES", 6);
  if (st > cap) st = cap;
  if (st < 1) st = 1;
  return st;
}

int choose_bn(int n) {
  // measured on B200 (profiles/): with cta_group::2, 256x128 tile pairs beat 256x256 on every shape
  // of this network except the GeGLU projection (which asks for 256 explicitly)
  static const int max_bn = env_int("PF_GEMM_MAX_BN", 128);
  if (n % 256 == 0 && max_bn >= 256) return 256;
  if (n % 128 == 0 && max_bn >= 128) return 128;
  if (n % 64 == 0) return 64;
  fail("GEMM N=%d is not a multiple of 64", n);
}

}  // namespace pf
