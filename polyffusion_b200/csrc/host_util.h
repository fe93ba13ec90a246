// Host-side helpers: error plumbing, TMA tensor-map encoding, plan-time arena, GEMM parameter
// builders.  Internal to libpf_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace pf {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

[[noreturn]] inline void fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Error(buf);
}

#define PF_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess)                                                             \
      ::pf::fail("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, \
                 cudaGetErrorString(e__));                                              \
  } while (0)

#define PF_CHECK(cond, ...)            \
  do {                                 \
    if (!(cond)) ::pf::fail(__VA_ARGS__); \
  } while (0)

// ------------------------------------------------------------------ tensor maps
// A operand: bf16 NHWC tensor {C, W, H, N}, box {64, box_w, box_h, 1}, 128B swizzle, zero OOB fill.
CUtensorMap make_map_4d(const void* ptr, int C, int W, int H, int N, int box_w, int box_h);
// RAW segment source (gemm_tc.cuh): fp32 NHWC tensor {C, W, H, N}, box {64, box_w, box_h, 1}, NO swizzle (the
// tile lands as 128 pixel rows of 256 bytes and is converted in shared memory)
CUtensorMap make_map_4d_f32(const void* ptr, int C, int W, int H, int N, int box_w, int box_h);
// B operand: bf16 K-major matrix {K, rows}, box {64, box_rows}, 128B swizzle.
CUtensorMap make_map_2d(const void* ptr, long long K, long long rows, int box_rows);
// f16f8 fp8 rows (common.cuh): byte tensors with 2 bytes per channel element, {2C, W, H, N} /
// {2K, rows}, boxes of 128 bytes (= one 64-channel k-block: 64 x h8 + 64 x l8), 128B swizzle
CUtensorMap make_map_4d_u8(const void* ptr, int C, int W, int H, int N, int box_w, int box_h);
CUtensorMap make_map_2d_u8(const void* ptr, long long K, long long rows, int box_rows);

// ------------------------------------------------------------------ plan-time arena
// First-fit free-list allocator over a caller-provided workspace.  Used only while a plan is being
// built (the op order is static), never on the launch path.
class Arena {
 public:
  explicit Arena(char* base) : base_(base) {}
  void* alloc(size_t bytes);
  void free(void* p);
  size_t peak() const { return peak_; }
  // true if p is the start of a live block of this arena
  bool owns(const void* p) const {
    const char* c = static_cast<const char*>(p);
    return c >= base_ && live_.count(static_cast<size_t>(c - base_)) != 0;
  }

 private:
  char* base_;
  size_t top_ = 0, peak_ = 0;
  std::map<size_t, size_t> free_;  // offset -> size
  std::map<size_t, size_t> live_;  // offset -> size
};

struct Split {  // split-bf16 operand tensor
  bf16* hi = nullptr;
  bf16* lo = nullptr;
};

inline int choose_box_w(int W) { return W < 128 ? W : 128; }

// 3x3 / 1x1 tap tables
void fill_taps_3x3(GemmSeg& sg);
void fill_taps_3x3_s2d(GemmSeg& sg);  // stride-2 conv over the 4 parity planes written by XF_S2D
void fill_taps_1x1(GemmSeg& sg);

int gemm_default_stages(int bn);
int gemm_default_stages2(int bn);  // cta_group::2 variant
int choose_bn(int n);

}  // namespace pf
