// Non-GEMM kernels of the UNet hot path.  See kernels.cuh for the contract of each launcher.
// All fp32 arithmetic here avoids fast-math so results track the reference's ATen fp32 ops.
#include "common.cuh"
#include "kernels.cuh"
#include <math.h>

namespace pf {

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }
// SiLU on the operand-transform path: ex2.approx + rcp (relative error ~1e-6, an order of magnitude
// below the 2^-16 operand rounding that follows)
__device__ __forceinline__ float silu_fast(float v) {
  return v * fast_rcp(1.0f + fast_ex2(-1.4426950408889634f * v));
}

// ------------------------------------------------------------------------------------------------
// conv_in: NCHW fp32 (tiny Cin) -> NHWC fp32, 3x3 pad 1.
// block = 256 threads = 64 pixels of one row x 4 channel quarters.
// ------------------------------------------------------------------------------------------------
constexpr int CI_ROWS = 4;  // image rows per block (amortises the weight / halo fill)
__global__ void __launch_bounds__(256) conv_in_kernel(const float* __restrict__ x,
                                                      const float* __restrict__ w,
                                                      const float* __restrict__ bias,
                                                      float* __restrict__ out, int Cin, int H, int W,
                                                      int Cout) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];
  const int K = Cin * 9;
  float* sw = sm;                    // [K][Cout]
  float* sx = sm + K * Cout;         // [Cin][CI_ROWS + 2][66]
  const int b = blockIdx.z, y0 = blockIdx.y * CI_ROWS, xb = blockIdx.x * 64;
  // destination-ordered fill (consecutive threads -> consecutive smem words: no bank conflicts);
  // the strided global reads of the 4.6 KB weight tensor are L2 hits
  for (int i = threadIdx.x; i < K * Cout; i += 256) {
    const int k = i / Cout, co = i % Cout;  // w is [Cout][Cin][3][3] -> k = ci*9 + ky*3 + kx
    sw[i] = __ldg(w + co * K + k);
  }
  const int HR = CI_ROWS + 2;
  for (int i = threadIdx.x; i < Cin * HR * 66; i += 256) {
    const int ci = i / (HR * 66), r = (i / 66) % HR, xx = i % 66;
    const int gy = y0 + r - 1, gx = xb + xx - 1;
    float v = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W)
      v = __ldg(x + ((static_cast<long long>(b) * Cin + ci) * H + gy) * W + gx);
    sx[i] = v;
  }
  __syncthreads();
  const int p = threadIdx.x >> 2, q = threadIdx.x & 3;
  if (xb + p >= W) return;
  // Cout / 4 == 16 (checked by the host): 16 accumulators per thread, weights broadcast from smem
  float4 bv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bv[j] = *reinterpret_cast<const float4*>(bias + q * 16 + 4 * j);
  for (int rr = 0; rr < CI_ROWS; ++rr) {
    const int y = y0 + rr;
    if (y >= H) break;
    float4 acc[4] = {bv[0], bv[1], bv[2], bv[3]};
    for (int ci = 0; ci < Cin; ++ci)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float v = sx[(ci * HR + rr + r) * 66 + p + kx];
          const float* wr = sw + (ci * 9 + r * 3 + kx) * Cout + q * 16;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 ww = *reinterpret_cast<const float4*>(wr + 4 * j);
            acc[j].x = fmaf(v, ww.x, acc[j].x);
            acc[j].y = fmaf(v, ww.y, acc[j].y);
            acc[j].z = fmaf(v, ww.z, acc[j].z);
            acc[j].w = fmaf(v, ww.w, acc[j].w);
          }
        }
    float* o = out + ((static_cast<long long>(b) * H + y) * W + xb + p) * Cout + q * 16;
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(o + 4 * j) = acc[j];
  }
}

// Fast path for the prmat2c geometry (Cin = 2, Cout = 64, W <= 128): a thread owns 4 output channels
// (its 72 weights live in registers) and walks pixels; a warp = 2 adjacent pixels x 16 channel quads,
// so every store instruction writes 512 contiguous bytes and the only shared-memory traffic is the
// 18 broadcast halo reads per pixel.  The GroupNorm statistics of the output (needed by the first
// ResBlock and by the last skip connection) are reduced here as well instead of a separate pass.
// Accumulation order (ci, ky, kx) is the same as in conv_in_kernel.
constexpr int CI2_ROWS = 4;
__global__ void __launch_bounds__(256) conv_in2_kernel(const float* __restrict__ x,
                                                       const float* __restrict__ w,
                                                       const float* __restrict__ bias,
                                                       float* __restrict__ out,
                                                       double* __restrict__ stats, int H, int W) {
  pdl_wait();
  pdl_trigger();
  constexpr int HR = CI2_ROWS + 2, PITCH = 132;
  __shared__ float sx[2 * HR * PITCH];
  __shared__ float s_red[2][16][64];
  const int b = blockIdx.y, y0 = blockIdx.x * CI2_ROWS;
  const int cq = threadIdx.x & 15, ps = threadIdx.x >> 4;
  for (int i = threadIdx.x; i < 2 * HR * PITCH; i += 256) {
    const int ci = i / (HR * PITCH), r = (i / PITCH) % HR, xx = i % PITCH;
    const int gy = y0 + r - 1, gx = xx - 1;
    float v = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W)
      v = __ldg(x + ((static_cast<long long>(b) * 2 + ci) * H + gy) * W + gx);
    sx[i] = v;
  }
  float4 wr[18];
#pragma unroll
  for (int k = 0; k < 18; ++k) {  // k = ci * 9 + ky * 3 + kx; w is [64][2][3][3]
    wr[k].x = __ldg(w + (4 * cq + 0) * 18 + k);
    wr[k].y = __ldg(w + (4 * cq + 1) * 18 + k);
    wr[k].z = __ldg(w + (4 * cq + 2) * 18 + k);
    wr[k].w = __ldg(w + (4 * cq + 3) * 18 + k);
  }
  const float4 bv = *reinterpret_cast<const float4*>(bias + 4 * cq);
  __syncthreads();
  float4 ssum = make_float4(0.f, 0.f, 0.f, 0.f), ssq = ssum;
  for (int rr = 0; rr < CI2_ROWS; ++rr) {
    const int y = y0 + rr;
    if (y >= H) break;
    for (int px = ps; px < W; px += 16) {
      float4 acc = bv;
#pragma unroll
      for (int ci = 0; ci < 2; ++ci)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float v = sx[(ci * HR + rr + r) * PITCH + px + kx];
            const float4 ww = wr[ci * 9 + r * 3 + kx];
            acc.x = fmaf(v, ww.x, acc.x);
            acc.y = fmaf(v, ww.y, acc.y);
            acc.z = fmaf(v, ww.z, acc.z);
            acc.w = fmaf(v, ww.w, acc.w);
          }
      *reinterpret_cast<float4*>(out + ((static_cast<long long>(b) * H + y) * W + px) * 64 + 4 * cq) = acc;
      ssum.x += acc.x; ssum.y += acc.y; ssum.z += acc.z; ssum.w += acc.w;
      ssq.x = fmaf(acc.x, acc.x, ssq.x); ssq.y = fmaf(acc.y, acc.y, ssq.y);
      ssq.z = fmaf(acc.z, acc.z, ssq.z); ssq.w = fmaf(acc.w, acc.w, ssq.w);
    }
  }
  if (stats) {
    *reinterpret_cast<float4*>(&s_red[0][ps][4 * cq]) = ssum;
    *reinterpret_cast<float4*>(&s_red[1][ps][4 * cq]) = ssq;
    __syncthreads();
    if (threadIdx.x < 128) {
      const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
      double t = 0.0;
#pragma unroll
      for (int r = 0; r < 16; ++r) t += static_cast<double>(s_red[which][r][c]);
      atomicAdd(stats + (static_cast<long long>(b) * 64 + c) * 2 + which, t);
    }
  }
}

void launch_gn_stats(const float* src, double* acc, int B, int HW, int Cs, int Ctot, int coff,
                     cudaStream_t s);

void launch_conv_in(const float* x, const float* w, const float* bias, float* out, double* stats, int B,
                    int Cin, int H, int W, int Cout, cudaStream_t s) {
  static const bool slow = std::getenv("PF_CONV_SLOW") != nullptr;  // A/B switch: generic kernels
  if (!slow && Cin == 2 && Cout == 64 && W <= 128) {
    dim3 grid((H + CI2_ROWS - 1) / CI2_ROWS, B);
    launch_pdl(conv_in2_kernel, grid, dim3(256), 0, s, x, w, bias, out, stats, H, W);
    return;
  }
  dim3 grid((W + 63) / 64, (H + CI_ROWS - 1) / CI_ROWS, B);
  const size_t smem = (static_cast<size_t>(Cin) * 9 * Cout + Cin * (CI_ROWS + 2) * 66) * sizeof(float);
  launch_pdl(conv_in_kernel, grid, dim3(256), smem, s, x, w, bias, out, Cin, H, W, Cout);
  if (stats) launch_gn_stats(out, stats, B, H * W, Cout, Cout, 0, s);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ src,
                                                       double* __restrict__ acc, int HW, int Cs,
                                                       int Ctot, int coff, int pix_per_block) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_sum[256 * 4];
  __shared__ float s_sq[256 * 4];
  const int b = blockIdx.y;
  const int nvec = Cs >> 2;            // float4 per pixel
  const int rows = 256 / nvec;         // pixel rows processed in parallel
  const int v = threadIdx.x % nvec, pr = threadIdx.x / nvec;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f), sq = sum;
  const float4* base = reinterpret_cast<const float4*>(src + static_cast<long long>(b) * HW * Cs);
  if (pr < rows) {
    for (int p = p0 + pr; p < p1; p += rows) {
      const float4 a = __ldg(base + static_cast<long long>(p) * nvec + v);
      sum.x += a.x; sum.y += a.y; sum.z += a.z; sum.w += a.w;
      sq.x = fmaf(a.x, a.x, sq.x); sq.y = fmaf(a.y, a.y, sq.y);
      sq.z = fmaf(a.z, a.z, sq.z); sq.w = fmaf(a.w, a.w, sq.w);
    }
  }
  reinterpret_cast<float4*>(s_sum)[threadIdx.x] = sum;
  reinterpret_cast<float4*>(s_sq)[threadIdx.x] = sq;
  __syncthreads();
  if (threadIdx.x < Cs) {
    const int c = threadIdx.x;
    double ts = 0.0, tq = 0.0;
    for (int r = 0; r < rows; ++r) {
      ts += static_cast<double>(s_sum[(r * nvec) * 4 + c]);
      tq += static_cast<double>(s_sq[(r * nvec) * 4 + c]);
    }
    double* a = acc + (static_cast<long long>(b) * Ctot + coff + c) * 2;
    atomicAdd(a, ts);
    atomicAdd(a + 1, tq);
  }
}

void launch_gn_stats(const float* src, double* acc, int B, int HW, int Cs, int Ctot, int coff,
                     cudaStream_t s) {
  // aim for >= ~4 waves of blocks; each block reduces pix_per_block pixels
  int chunks = max(1, min(HW / 64, (148 * 8 + B - 1) / B));
  int ppb = (HW + chunks - 1) / chunks;
  chunks = (HW + ppb - 1) / ppb;
  dim3 grid(chunks, B);
  launch_pdl(gn_stats_kernel, grid, dim3(256), 0, s, src, acc, HW, Cs, Ctot, coff, ppb);
}

// GroupNorm finalize for a RAW GEMM segment (gemm_tc.cuh): per-(sample, channel) scale = gamma * rstd and
// shift = beta - mean * scale from the producers' fp64 sums; one block per sample, one thread per channel.
__global__ void __launch_bounds__(512) gn_finalize_kernel(const double* __restrict__ stats,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float eps, int groups,
                                                          int HW, int C, float* __restrict__ scale,
                                                          float* __restrict__ shift) {
  pdl_wait();
  pdl_trigger();
  // same structure as the operand transform's prologue: every thread fetches its own channel's sums (one L2 round
  // trip), one thread per group reduces from shared memory in channel order, one thread per channel writes
  __shared__ double2 s_st[512];
  __shared__ float s_gmean[32], s_grstd[32];
  const int b = blockIdx.x;
  const int cpg = C / groups;
  const int c = threadIdx.x;  // blockDim.x == C <= 512
  float gam = 0.f, bet = 0.f;
  if (c < C) {
    gam = __ldg(gamma + c);
    bet = __ldg(beta + c);
    s_st[c] = __ldcg(reinterpret_cast<const double2*>(stats + (static_cast<long long>(b) * C + c) * 2));
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    double ts = 0.0, tq = 0.0;
    for (int i = 0; i < cpg; ++i) {
      const double2 st = s_st[g * cpg + i];
      ts += st.x;
      tq += st.y;
    }
    const double inv_n = 1.0 / (static_cast<double>(HW) * cpg);
    const double mean = ts * inv_n;
    double var = fma(tq, inv_n, -mean * mean);
    if (var < 0.0) var = 0.0;
    s_gmean[g] = static_cast<float>(mean);
    s_grstd[g] = rsqrtf(static_cast<float>(var) + eps);
  }
  __syncthreads();
  if (c < C) {
    const int g = c / cpg;
    const float sc = gam * s_grstd[g];
    scale[static_cast<long long>(b) * C + c] = sc;
    shift[static_cast<long long>(b) * C + c] = bet - s_gmean[g] * sc;
  }
}

void launch_gn_finalize(const double* stats, const float* gamma, const float* beta, float eps, int groups,
                        int B, int HW, int C, float* scale, float* shift, cudaStream_t s) {
  launch_pdl(gn_finalize_kernel, dim3(B), dim3(C), 0, s, stats, gamma, beta, eps, groups, HW, C, scale, shift);
}

// ------------------------------------------------------------------------------------------------
// act_split: fp32 NHWC (one or two concatenated sources) -> split bf16 operand tensor(s).
// GroupNorm is finalised in the block prologue from the per-(sample, channel) fp64 sums the
// producing GEMM epilogues left behind (no separate finalize launch): scale/shift for the block's
// sample go to smem.  One thread-item = 16 channels of one pixel (4 x 16-byte loads in flight).
// Optional second output = plain split of the same input (operand of the ResBlock 1x1 skip conv).
// ------------------------------------------------------------------------------------------------
constexpr int AS_IPT = 4;  // items per thread

__device__ __forceinline__ void store_split16(const float* v, bf16* oh, bf16* ol) {
  uint4 h0, l0, h1, l1;
  split2(v[0], v[1], h0.x, l0.x);   split2(v[2], v[3], h0.y, l0.y);
  split2(v[4], v[5], h0.z, l0.z);   split2(v[6], v[7], h0.w, l0.w);
  split2(v[8], v[9], h1.x, l1.x);   split2(v[10], v[11], h1.y, l1.y);
  split2(v[12], v[13], h1.z, l1.z); split2(v[14], v[15], h1.w, l1.w);
  reinterpret_cast<uint4*>(oh)[0] = h0;
  reinterpret_cast<uint4*>(oh)[1] = h1;
  reinterpret_cast<uint4*>(ol)[0] = l0;
  reinterpret_cast<uint4*>(ol)[1] = l1;
}

// f16f8 operand (common.cuh): 16 channels starting at channel c of the pixel whose row starts at
// element offset `row` (= pixel * C): h16 [pixel][C] fp16, fp8 rows [pixel][C / 64][h8 x 64 | l8 x 64]
__device__ __forceinline__ void store_f8_16(const float* v, bf16* o16, bf16* o8, long long row, int c) {
  uint2 h[4];
  uint32_t h8[4], l8[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    split_f8x4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3], 1.f, F8_ACT_LO_SCALE, h[i], h8[i], l8[i]);
  uint4* p16 = reinterpret_cast<uint4*>(o16 + row + c);
  p16[0] = make_uint4(h[0].x, h[0].y, h[1].x, h[1].y);
  p16[1] = make_uint4(h[2].x, h[2].y, h[3].x, h[3].y);
  uint8_t* p8 = reinterpret_cast<uint8_t*>(o8) + row * 2 + (c >> 6) * 128 + (c & 63);
  *reinterpret_cast<uint4*>(p8) = make_uint4(h8[0], h8[1], h8[2], h8[3]);
  *reinterpret_cast<uint4*>(p8 + 64) = make_uint4(l8[0], l8[1], l8[2], l8[3]);
}

__global__ void __launch_bounds__(256) act_split_kernel(ActSplitArgs a) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_scale[512], s_shift[512];
  __shared__ double2 s_st[512];
  __shared__ float s_gmean[32], s_grstd[32];
  const int C = a.C0 + a.C1;
  const int b = blockIdx.y;
  const bool norm = a.stats0 != nullptr;
  if (norm) {
    // GroupNorm finalize.  Every thread fetches the (sum, sum of squares) pair and the affine pair of its own
    // channel(s) -- ONE round trip to L2 for the whole block (the per-thread loop over the channels of a group
    // this replaces issued its loads one after the other: 19.9 -> 17.6 us per launch on the 16 x 16 maps, where a
    // block only streams 128 KB).  One thread per group then reduces from shared memory (same summation order as
    // before) and takes the fp64 square root; one thread per channel forms scale / shift.
    const int cpg = C / a.groups;
    float gam[4] = {0.f, 0.f, 0.f, 0.f}, bet[4] = {0.f, 0.f, 0.f, 0.f};  // C <= 512, blockDim >= 128
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = threadIdx.x + k * blockDim.x;
      if (c < C) {
        gam[k] = __ldg(a.gamma + c);
        bet[k] = __ldg(a.beta + c);
        s_st[c] = c < a.C0 ? __ldcg(reinterpret_cast<const double2*>(a.stats0 + (static_cast<long long>(b) * a.C0 + c) * 2))
                           : __ldcg(reinterpret_cast<const double2*>(a.stats1 + (static_cast<long long>(b) * a.C1 + (c - a.C0)) * 2));
      }
    }
    __syncthreads();
    if (threadIdx.x < a.groups) {
      const int g = threadIdx.x;
      double ts = 0.0, tq = 0.0;
      for (int i = 0; i < cpg; ++i) {
        const double2 st = s_st[g * cpg + i];
        ts += st.x;
        tq += st.y;
      }
      // mean and E[x^2] - mean^2 in fp64 (the subtraction cancels), the reciprocal square root in fp32: an fp64
      // divide + sqrt + divide chain on 32 threads was ~2 us of pure latency per launch
      const double inv_n = 1.0 / (static_cast<double>(a.H) * a.W * cpg);
      const double mean = ts * inv_n;
      double var = fma(tq, inv_n, -mean * mean);
      if (var < 0.0) var = 0.0;
      s_gmean[g] = static_cast<float>(mean);
      s_grstd[g] = rsqrtf(static_cast<float>(var) + a.eps);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = threadIdx.x + k * blockDim.x;
      if (c < C) {
        const int g = c / cpg;
        const float sc = gam[k] * s_grstd[g];
        s_scale[c] = sc;
        s_shift[c] = bet[k] - s_gmean[g] * sc;
      }
    }
    __syncthreads();
  }
  const int c16n = C >> 4;
  const int HW = a.H * a.W;
  const int items = HW * c16n;
  const long long pix_base = static_cast<long long>(b) * HW;
  // grid-stride over the sample's items: gridDim.x blocks per sample, one GroupNorm prologue per block
#pragma unroll 1
  for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < items; item += gridDim.x * blockDim.x) {
    const int c = (item % c16n) * 16;
    const int pl = item / c16n;  // pixel inside the sample
    const long long pix = pix_base + pl;
    float v[16];
    {
      const float* sp = (c < a.C0) ? a.src0 + pix * a.C0 + c : a.src1 + pix * a.C1 + (c - a.C0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(sp) + i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      }
    }
    if (a.out2_hi) {
      if (a.fmt8) store_f8_16(v, a.out2_hi, a.out2_lo, pix * C, c);
      else store_split16(v, a.out2_hi + pix * C + c, a.out2_lo + pix * C + c);
    }
    if (norm) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], s_scale[c + i], s_shift[c + i]);
    }
    if (a.silu) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = silu_fast(v[i]);
    }
    if (a.layout == XF_SAME) {
      if (a.fmt8) store_f8_16(v, a.out_hi, a.out_lo, pix * C, c);
      else store_split16(v, a.out_hi + pix * C + c, a.out_lo + pix * C + c);
    } else {
      const int y = pl / a.W, x = pl % a.W;
      if (a.layout == XF_UP2) {
        const int H2 = 2 * a.H, W2 = 2 * a.W;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const long long o = ((static_cast<long long>(b) * H2 + 2 * y + dy) * W2 + 2 * x + dx) * C;
            if (a.fmt8) store_f8_16(v, a.out_hi, a.out_lo, o, c);
            else store_split16(v, a.out_hi + o + c, a.out_lo + o + c);
          }
      } else {  // XF_S2D: [b*4 + (y&1)*2 + (x&1)][H/2][W/2][C]
        const int Hh = a.H >> 1, Wh = a.W >> 1;
        const long long o =
            (((static_cast<long long>(b) * 4 + (y & 1) * 2 + (x & 1)) * Hh + (y >> 1)) * Wh + (x >> 1)) * C;
        if (a.fmt8) store_f8_16(v, a.out_hi, a.out_lo, o, c);
        else store_split16(v, a.out_hi + o + c, a.out_lo + o + c);
      }
    }
  }
}

void launch_act_split(const ActSplitArgs& a, cudaStream_t s) {
  const int C = a.C0 + a.C1;
  const long long items = static_cast<long long>(a.H) * a.W * (C >> 4);
  // PF_ACT_THREADS=128: blocks small enough (64 registers per thread) to sit beside a register-capped
  // GEMM CTA in the half-batch lane experiment (unet.cu); 256 is the measured best otherwise
  static const int threads = std::getenv("PF_ACT_THREADS") ? std::atoi(std::getenv("PF_ACT_THREADS")) : 256;
  // Measured alternatives on B200, batch 64 (profiles/r4i_*, r4j_*; sum over the 47 transforms of one step):
  // 1 / 2 / 8 items per thread 4.76 / 4.27 / 4.17 ms against 3.98 ms with 4; issuing the loads of item k + 1
  // before item k is processed (77 registers, 3 blocks per SM, or capped at 64 registers for 4) 4.31 / 4.11 ms.
  // Normalising transforms run in the persistent form: ONE resident wave of blocks in total (4 blocks of 256
  // threads per SM), each block striding over its sample's items -- the GroupNorm prologue is paid once per block
  // and there is no partial last wave (measured on B200, batch 64, profiles/r4o_*: 103 -> 98.5 us on the 128 x 128
  // maps, 74 -> 70.5 us at 64 x 64, 44 -> 40 us at 32 x 32; two / four waves gain less).  The plain re-layout
  // transforms (no prologue) are slower that way (86 -> 98 us for the stride-2 planes) and keep one block per
  // 4 items per thread.  PF_ACT_WAVES overrides (0 = never persistent).
  static const int waves = std::getenv("PF_ACT_WAVES") ? std::atoi(std::getenv("PF_ACT_WAVES")) : 1;
  // PF_ACT_SMALL_IPT=<n>: items per thread on maps of at most 1024 pixels (more, smaller blocks)
  static const int small_ipt = std::getenv("PF_ACT_SMALL_IPT") ? std::atoi(std::getenv("PF_ACT_SMALL_IPT")) : AS_IPT;
  const int ipt = (a.H * a.W <= 1024 && small_ipt > 0) ? small_ipt : AS_IPT;
  long long bx = (items + threads * ipt - 1) / (threads * ipt);
  if (waves > 0 && a.stats0 != nullptr) {
    const long long slots = 148ll * (1024 / threads) * waves;
    const long long per_sample = std::max(1ll, slots / a.B);
    bx = std::min(bx, per_sample);
  }
  dim3 grid(static_cast<unsigned>(bx), a.B);
  launch_pdl(act_split_kernel, grid, dim3(threads), 0, s, a);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm + split (warp per row, values held in registers)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void __launch_bounds__(256) ln_split_kernel(const float* __restrict__ src,
                                                       const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float eps,
                                                       bf16* __restrict__ out_hi,
                                                       bf16* __restrict__ out_lo, long long rows,
                                                       int C, int fmt8) {
  pdl_wait();
  pdl_trigger();
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nch = C >> 7;  // float4 chunks per lane (C / 128), <= 4
  float4 v[4];
  const float4* sp = reinterpret_cast<const float4*>(src + row * C);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < nch) {
      v[i] = __ldg(sp + i * 32 + lane);
      sum += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  const float mean = warp_sum(sum) / static_cast<float>(C);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < nch) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += a * a + b * b + c * c + d * d;
    }
  const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(C) + eps);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < nch) {
      const int c0 = (i * 32 + lane) * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c0));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + c0));
      const float o0 = (v[i].x - mean) * rstd * g.x + bb.x;
      const float o1 = (v[i].y - mean) * rstd * g.y + bb.y;
      const float o2 = (v[i].z - mean) * rstd * g.z + bb.z;
      const float o3 = (v[i].w - mean) * rstd * g.w + bb.w;
      if (fmt8) {  // f16f8 operand (common.cuh): fp16 [row][C] + fp8 rows [row][C / 64][h8 x 64 | l8 x 64]
        uint2 h16;
        uint32_t h8, l8;
        split_f8x4(o0, o1, o2, o3, 1.f, F8_ACT_LO_SCALE, h16, h8, l8);
        *reinterpret_cast<uint2*>(out_hi + row * C + c0) = h16;
        uint8_t* p8 = reinterpret_cast<uint8_t*>(out_lo) + row * C * 2 + (c0 >> 6) * 128 + (c0 & 63);
        *reinterpret_cast<uint32_t*>(p8) = h8;
        *reinterpret_cast<uint32_t*>(p8 + 64) = l8;
      } else {
        uint2 h, l;
        split2(o0, o1, h.x, l.x);
        split2(o2, o3, h.y, l.y);
        *reinterpret_cast<uint2*>(out_hi + row * C + c0) = h;
        *reinterpret_cast<uint2*>(out_lo + row * C + c0) = l;
      }
    }
}

void launch_ln_split(const float* src, const float* gamma, const float* beta, float eps,
                     bf16* out_hi, bf16* out_lo, long long rows, int C, cudaStream_t s, int fmt8) {
  const unsigned blocks = static_cast<unsigned>((rows + 7) / 8);
  launch_pdl(ln_split_kernel, dim3(blocks), dim3(256), 0, s, src, gamma, beta, eps, out_hi, out_lo, rows, C, fmt8);
}

// ------------------------------------------------------------------------------------------------
// GeGLU + split: one thread = 8 outputs
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float g) {
  return 0.5f * g * (1.0f + erff(g * 0.70710678118654752440f));
}

__global__ void __launch_bounds__(256) geglu_split_kernel(const float* __restrict__ src,
                                                          bf16* __restrict__ out_hi,
                                                          bf16* __restrict__ out_lo, long long rows,
                                                          int F) {
  const int f8n = F >> 3;
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= rows * f8n) return;
  const long long row = idx / f8n;
  const int f = static_cast<int>(idx % f8n) * 8;
  const float* xp = src + row * 2 * F + f;
  const float* gp = xp + F;
  float xv[8], gv[8];
  *reinterpret_cast<float4*>(xv) = __ldg(reinterpret_cast<const float4*>(xp));
  *reinterpret_cast<float4*>(xv + 4) = __ldg(reinterpret_cast<const float4*>(xp) + 1);
  *reinterpret_cast<float4*>(gv) = __ldg(reinterpret_cast<const float4*>(gp));
  *reinterpret_cast<float4*>(gv + 4) = __ldg(reinterpret_cast<const float4*>(gp) + 1);
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = xv[i] * gelu_erf(gv[i]);
  uint4 h, l;
  split2(o[0], o[1], h.x, l.x);
  split2(o[2], o[3], h.y, l.y);
  split2(o[4], o[5], h.z, l.z);
  split2(o[6], o[7], h.w, l.w);
  *reinterpret_cast<uint4*>(out_hi + row * F + f) = h;
  *reinterpret_cast<uint4*>(out_lo + row * F + f) = l;
}

void launch_geglu_split(const float* src, bf16* out_hi, bf16* out_lo, long long rows, int F,
                        cudaStream_t s) {
  const long long total = rows * (F >> 3);
  geglu_split_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(src, out_hi, out_lo,
                                                                                  rows, F);
}

// ------------------------------------------------------------------------------------------------
// softmax + split: warp per row, Nk <= 1024 (8 float4 per lane)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_split_kernel(const float* __restrict__ S, float scale,
                                                            bf16* __restrict__ out_hi,
                                                            bf16* __restrict__ out_lo,
                                                            long long rows, int Nk) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nch = Nk >> 7;
  float4 v[8];
  const float4* sp = reinterpret_cast<const float4*>(S + row * Nk);
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nch) {
      float4 a = __ldg(sp + i * 32 + lane);
      a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
      v[i] = a;
      mx = fmaxf(mx, fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)));
    }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nch) {
      v[i].x = expf(v[i].x - mx); v[i].y = expf(v[i].y - mx);
      v[i].z = expf(v[i].z - mx); v[i].w = expf(v[i].w - mx);
      sum += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  const float inv = 1.0f / warp_sum(sum);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nch) {
      const int c0 = (i * 32 + lane) * 4;
      uint2 h, l;
      split2(v[i].x * inv, v[i].y * inv, h.x, l.x);
      split2(v[i].z * inv, v[i].w * inv, h.y, l.y);
      *reinterpret_cast<uint2*>(out_hi + row * Nk + c0) = h;
      *reinterpret_cast<uint2*>(out_lo + row * Nk + c0) = l;
    }
}

void launch_softmax_split(const float* S, float scale, bf16* out_hi, bf16* out_lo, long long rows,
                          int Nk, cudaStream_t s) {
  softmax_split_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, s>>>(S, scale, out_hi,
                                                                               out_lo, rows, Nk);
}

// ------------------------------------------------------------------------------------------------
// timestep embedding + small linears
// ------------------------------------------------------------------------------------------------
__global__ void time_sinusoid_kernel(const long long* __restrict__ t,
                                     const float* __restrict__ freqs, float* __restrict__ out, int B,
                                     int half) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, j = i % half;
  // args = t.float() * f_j ; emb = cat([cos(args), sin(args)])  (unet.py:165-169).  The frequency
  // table f_j = exp(-ln(10000) * j / half) is evaluated by the host exactly as the reference does.
  const float a = __fmul_rn(static_cast<float>(t[b]), freqs[j]);
  out[static_cast<long long>(b) * 2 * half + j] = cosf(a);
  out[static_cast<long long>(b) * 2 * half + half + j] = sinf(a);
}
void launch_time_sinusoid(const long long* t, const float* freqs, float* out, int B, int half,
                          cudaStream_t s) {
  launch_pdl(time_sinusoid_kernel, dim3((B * half + 127) / 128), dim3(128), 0, s, t, freqs, out, B, half);
}

// test helpers: fp32 = hi + lo, and [rows, C] -> transposed split [C, rows] per image
// ------------------------------------------------------------------------------------------------
// Generic building blocks of the legacy ddpm.unet.UNet evaluation (polyffusion_b200/ddpm/unet.py):
// any channel count / group count GroupNorm (+ Swish), row softmax, sin-then-cos timestep embedding.
// Not on the sdf sampling path (those shapes use the fused kernels above).
// ------------------------------------------------------------------------------------------------
// one block per (sample, group): two passes over the group's [HW, cpg] elements (fp64 accumulation)
__global__ void __launch_bounds__(256) groupnorm_generic_kernel(const float* __restrict__ x,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float eps,
                                                                int silu, float* __restrict__ out, int HW,
                                                                int C, int groups) {
  const int b = blockIdx.x / groups, g = blockIdx.x % groups;
  const int cpg = C / groups;
  const float* xb = x + static_cast<long long>(b) * HW * C + g * cpg;
  float* ob = out + static_cast<long long>(b) * HW * C + g * cpg;
  const int n = HW * cpg;
  double s1 = 0.0, s2 = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float v = xb[static_cast<long long>(i / cpg) * C + i % cpg];
    s1 += v;
    s2 += static_cast<double>(v) * v;
  }
  __shared__ double r1[256], r2[256];
  r1[threadIdx.x] = s1;
  r2[threadIdx.x] = s2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      r1[threadIdx.x] += r1[threadIdx.x + o];
      r2[threadIdx.x] += r2[threadIdx.x + o];
    }
    __syncthreads();
  }
  const double mean = r1[0] / n;
  double var = r2[0] / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float mu = static_cast<float>(mean);
  for (int i = threadIdx.x; i < n; i += 256) {
    const int c = i % cpg;
    const long long o = static_cast<long long>(i / cpg) * C + c;
    float v = (xb[o] - mu) * rstd * gamma[g * cpg + c] + beta[g * cpg + c];
    if (silu) v = v / (1.0f + expf(-v));
    ob[o] = v;
  }
}
void launch_groupnorm_generic(const float* x, const float* gamma, const float* beta, float eps, int silu,
                              float* out, int B, int HW, int C, int groups, cudaStream_t s) {
  groupnorm_generic_kernel<<<B * groups, 256, 0, s>>>(x, gamma, beta, eps, silu, out, HW, C, groups);
}

// out[r, :] = softmax(scale * S[r, :]); one warp per row, any n
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ S, float scale,
                                                           float* __restrict__ out, long long rows, int n) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* sp = S + row * n;
  float mx = -INFINITY;
  for (int i = lane; i < n; i += 32) mx = fmaxf(mx, sp[i] * scale);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int i = lane; i < n; i += 32) sum += expf(sp[i] * scale - mx);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int i = lane; i < n; i += 32) out[row * n + i] = expf(sp[i] * scale - mx) * inv;
}
void launch_softmax_rows(const float* S, float scale, float* out, long long rows, int n, cudaStream_t s) {
  softmax_rows_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, s>>>(S, scale, out, rows, n);
}

// direct fp32 3x3 convolution (pad 1) for the two edge layers whose channel count is not GEMM-shaped
// (image_proj 2 -> 64 from NCHW, final 64 -> 2 to NCHW; ddpm/unet.py:345-347, 405-407): one thread per
// output element, taps in the order ky, kx, ci of a [Cout][Cin][3][3] weight
__global__ void __launch_bounds__(256) conv3x3_direct_kernel(const float* __restrict__ x,
                                                             const float* __restrict__ w,
                                                             const float* __restrict__ bias,
                                                             float* __restrict__ out, int B, int Cin, int H,
                                                             int W, int Cout, int in_nchw, int out_nchw) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * H * W * Cout;
  if (i >= total) return;
  int b, co, y, xx;
  if (out_nchw) {
    xx = static_cast<int>(i % W); y = static_cast<int>((i / W) % H);
    co = static_cast<int>((i / (static_cast<long long>(W) * H)) % Cout);
    b = static_cast<int>(i / (static_cast<long long>(W) * H * Cout));
  } else {
    co = static_cast<int>(i % Cout); xx = static_cast<int>((i / Cout) % W);
    y = static_cast<int>((i / (static_cast<long long>(Cout) * W)) % H);
    b = static_cast<int>(i / (static_cast<long long>(Cout) * W * H));
  }
  float acc = bias ? bias[co] : 0.f;
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = y + ky - 1;
    if (iy < 0 || iy >= H) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = xx + kx - 1;
      if (ix < 0 || ix >= W) continue;
      for (int ci = 0; ci < Cin; ++ci) {
        const float v = in_nchw ? x[((static_cast<long long>(b) * Cin + ci) * H + iy) * W + ix]
                                : x[((static_cast<long long>(b) * H + iy) * W + ix) * Cin + ci];
        acc = fmaf(v, w[((static_cast<long long>(co) * Cin + ci) * 3 + ky) * 3 + kx], acc);
      }
    }
  }
  out[i] = acc;
}
void launch_conv3x3_direct(const float* x, const float* w, const float* bias, float* out, int B, int Cin,
                           int H, int W, int Cout, int in_nchw, int out_nchw, cudaStream_t s) {
  const long long total = static_cast<long long>(B) * H * W * Cout;
  conv3x3_direct_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(x, w, bias, out, B, Cin, H, W,
                                                                                    Cout, in_nchw, out_nchw);
}

// legacy TimeEmbedding (ddpm/unet.py:62-72): [sin(t f_i), cos(t f_i)], t converted to fp32 first
__global__ void time_sincos_kernel(const long long* __restrict__ t, const float* __restrict__ freqs,
                                   float* __restrict__ out, int B, int half) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, k = i % half;
  const float a = static_cast<float>(t[b]) * freqs[k];
  out[b * 2 * half + k] = sinf(a);
  out[b * 2 * half + half + k] = cosf(a);
}
void launch_time_sincos(const long long* t, const float* freqs, float* out, int B, int half, cudaStream_t s) {
  time_sincos_kernel<<<(B * half + 127) / 128, 128, 0, s>>>(t, freqs, out, B, half);
}

__global__ void merge_split_kernel(const bf16* __restrict__ hi, const bf16* __restrict__ lo,
                                   float* __restrict__ out, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(hi[i]) + __bfloat162float(lo[i]);
}
void launch_merge_split(const bf16* hi, const bf16* lo, float* out, long long n, cudaStream_t s) {
  merge_split_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(hi, lo, out, n);
}
__global__ void transpose_split_kernel(const float* __restrict__ src, bf16* __restrict__ hi,
                                       bf16* __restrict__ lo, int rows, int C) {
  // src [img][rows][C] -> out [img][C][rows]
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  const long long per = static_cast<long long>(rows) * C;
  const long long img = blockIdx.y;
  if (i >= per) return;
  const int r = static_cast<int>(i % rows), c = static_cast<int>(i / rows);
  bf16 h, l;
  split_bf16(src[img * per + static_cast<long long>(r) * C + c], h, l);
  hi[img * per + i] = h;
  lo[img * per + i] = l;
}
void launch_transpose_split(const float* src, bf16* hi, bf16* lo, int imgs, int rows, int C,
                            cudaStream_t s) {
  const long long per = static_cast<long long>(rows) * C;
  dim3 grid(static_cast<unsigned>((per + 255) / 256), imgs);
  transpose_split_kernel<<<grid, 256, 0, s>>>(src, hi, lo, rows, C);
}

// tiled fp32 SGEMM for the per-sample vectors: block = 16 samples x 16 output columns, K in chunks
// of 32 through smem (both operands are K-contiguous, so the tile loads are coalesced)
__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ in,
                                                           long long ld_in,
                                                           const float* __restrict__ W,
                                                           const float* __restrict__ bias,
                                                           float* __restrict__ out, long long ld_out,
                                                           int B, int N, int K, int out_act) {
  pdl_wait();
  pdl_trigger();
  // blockIdx.z = group g: independent problems laid side by side (in += g*K, W += g*N*K,
  // bias += g*N, out += g*N); a plain call has one group
  in += static_cast<long long>(blockIdx.z) * K;
  W += static_cast<long long>(blockIdx.z) * N * K;
  if (bias) bias += static_cast<long long>(blockIdx.z) * N;
  out += static_cast<long long>(blockIdx.z) * N;
  __shared__ float s_in[16][33];
  __shared__ float s_w[16][33];
  const int tn = threadIdx.x & 15, tb = threadIdx.x >> 4;
  const int n0 = blockIdx.x * 16, b0 = blockIdx.y * 16;
  const int lr = threadIdx.x >> 5, lk = threadIdx.x & 31;  // loader: rows lr and lr + 8, column lk
  float acc = 0.f;
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lr + 8 * h;
      const int k = k0 + lk;
      s_in[r][lk] = (b0 + r < B && k < K) ? __ldg(in + (b0 + r) * ld_in + k) : 0.f;
      s_w[r][lk] = (n0 + r < N && k < K) ? __ldg(W + static_cast<long long>(n0 + r) * K + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) acc = fmaf(s_in[tb][k], s_w[tn][k], acc);
    __syncthreads();
  }
  const int b = b0 + tb, n = n0 + tn;
  if (b < B && n < N) {
    acc += bias ? bias[n] : 0.f;
    if (out_act == 1) acc = silu_f(acc);
    else if (out_act == 2) acc = expf(acc);  // linear_var(.).exp_() of the condition encoders
    out[b * ld_out + n] = acc;
  }
}
void launch_small_linear(const float* in, long long ld_in, const float* W, const float* bias,
                         float* out, long long ld_out, int B, int N, int K, int out_act,
                         cudaStream_t s, int groups) {
  dim3 grid((N + 15) / 16, (B + 15) / 16, groups);
  launch_pdl(small_linear_kernel, grid, dim3(256), 0, s, in, ld_in, W, bias, out, ld_out, B, N, K, out_act);
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) + Box-Muller: counter = (element, global sample, index, stream),
// key = seed.  One normal per call; the integer path is bit-exact against oracle/philox_oracle.py.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float philox_normal(unsigned long long seed, long long sample, uint32_t elem, int index,
                                               int which) {
  uint32_t r[4];
  philox4x32_10(elem, static_cast<uint32_t>(sample), static_cast<uint32_t>(index), static_cast<uint32_t>(which),
                static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
  const float u1 = (static_cast<float>(r[0] >> 8) + 0.5f) * 5.9604644775390625e-08f;  // (0, 1), 2^-24 steps
  const float u2 = (static_cast<float>(r[1] >> 8) + 0.5f) * 5.9604644775390625e-08f;
  return sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
}

// one element of the fused reverse step: e = eps at flat NCHW position i of sample b (i inside the sample)
__device__ __forceinline__ void fused_step_apply(const FusedStep& fs, int b, long long per_sample, long long i,
                                                 float e) {
  const int idx = fs.index[0];
  // Philox key of this run: seed + nonce * 2^64/phi; the nonce sits next to the index in device memory so
  // that a captured graph draws fresh noise for every sampling run (polyffusion_b200/_loop.py)
  const unsigned long long seed = fs.seed + static_cast<unsigned long long>(static_cast<unsigned>(fs.index[1])) *
                                                0x9E3779B97F4A7C15ull;
  const float* cf = fs.coef + static_cast<long long>(idx) * 8;
  const long long g = static_cast<long long>(b) * per_sample + i;
  const float x = fs.x[g];
  const bool ddpm = fs.kind == 1;
  float nz = 0.f;
  if (!(ddpm && idx == 0))
    nz = fs.noise ? fs.noise[g] : philox_normal(seed, fs.sample0 + b, static_cast<uint32_t>(i), idx, 0);
  nz = __fmul_rn(nz, fs.temperature);
  float xp;
  if (ddpm) {
    const float x0 = __fsub_rn(__fmul_rn(cf[0], x), __fmul_rn(cf[1], e));
    const float mean = __fadd_rn(__fmul_rn(cf[2], x0), __fmul_rn(cf[3], x));
    xp = __fadd_rn(mean, __fmul_rn(cf[4], nz));
  } else {
    const float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(cf[0], e)), cf[1]);
    const float dir = __fmul_rn(cf[3], e);
    xp = __fadd_rn(__fadd_rn(__fmul_rn(cf[2], x0), dir), __fmul_rn(cf[4], nz));
  }
  if (fs.orig) {
    float nk = 0.f;
    if (!(ddpm && idx == 0))
      nk = fs.noise_kn ? fs.noise_kn[g] : philox_normal(seed, fs.sample0 + b, static_cast<uint32_t>(i), idx, 1);
    const float xk = __fadd_rn(__fmul_rn(cf[5], fs.orig[g]), __fmul_rn(cf[6], nk));
    const float m = fs.mask[g];
    xp = __fadd_rn(__fmul_rn(xk, m), __fmul_rn(xp, __fsub_rn(1.0f, m)));
  }
  fs.x[g] = xp;
  if (fs.eps_out) fs.eps_out[g] = e;
}

__global__ void __launch_bounds__(256) step_from_eps_kernel(const FusedStep fs, const float* __restrict__ eps,
                                                            int B, long long per_sample) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= B * per_sample) return;
  fused_step_apply(fs, static_cast<int>(i / per_sample), per_sample, i % per_sample, eps[i]);
}
void launch_step_from_eps(const FusedStep& fs, const float* eps, int B, long long per_sample, cudaStream_t s) {
  const long long n = B * per_sample;
  step_from_eps_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(fs, eps, B, per_sample);
}
__global__ void step_advance_kernel(int* index, long long* t, const long long* __restrict__ t_table, int B) {
  const int idx = max(*index - 1, 0);
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) t[b] = t_table[idx];
  if (threadIdx.x == 0) *index = idx;
}
void launch_step_advance(int* index, long long* t, const long long* t_table, int B, cudaStream_t s) {
  step_advance_kernel<<<1, 256, 0, s>>>(index, t, t_table, B);
}
__global__ void fill_normal_kernel(float* __restrict__ out, long long n_samples, long long per_sample,
                                   unsigned long long seed, long long sample0, int index, int which) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n_samples * per_sample) return;
  out[i] = philox_normal(seed, sample0 + i / per_sample, static_cast<uint32_t>(i % per_sample), index, which);
}
void launch_fill_normal(float* out, long long n_samples, long long per_sample, unsigned long long seed,
                        long long sample0, int index, int which, cudaStream_t s) {
  const long long n = n_samples * per_sample;
  fill_normal_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(out, n_samples, per_sample, seed,
                                                                            sample0, index, which);
}
__global__ void gather_rows_kernel(const long long* __restrict__ t, const float* __restrict__ table,
                                   float* __restrict__ dst, int n_rows, int width) {
  const int b = blockIdx.y;
  long long r = t[b];
  r = r < 0 ? 0 : (r >= n_rows ? n_rows - 1 : r);
  for (int i = blockIdx.x * 256 + threadIdx.x; i < width; i += gridDim.x * 256)
    dst[static_cast<long long>(b) * width + i] = table[r * width + i];
}
void launch_gather_rows(const long long* t, const float* table, float* dst, int B, int n_rows, int width,
                        cudaStream_t s) {
  dim3 grid((width + 255) / 256, B);
  gather_rows_kernel<<<grid, 256, 0, s>>>(t, table, dst, n_rows, width);
}

// ------------------------------------------------------------------------------------------------
// conv_out: GN-apply + SiLU + conv3x3 (C -> Cout<=4) -> NCHW.  16x16 pixel tile + halo in smem.
// ------------------------------------------------------------------------------------------------
constexpr int CO_T = 16;
__global__ void __launch_bounds__(256) conv_out_kernel(const float* __restrict__ h,
                                                       const double* __restrict__ stats,
                                                       const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float eps,
                                                       const float* __restrict__ w,
                                                       const float* __restrict__ bias,
                                                       float* __restrict__ out, int H, int W, int C,
                                                       int Cout) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];
  const int pitch = C + 4;
  float* st = sm;                                        // [(T+2)*(T+2)][pitch]
  float* sw = sm + (CO_T + 2) * (CO_T + 2) * pitch;      // [Cout][9][C]
  float* sc = sw + Cout * 9 * C;                         // [C] GroupNorm scale
  float* sh = sc + C;                                    // [C] GroupNorm shift
  const int b = blockIdx.z, ty = blockIdx.y * CO_T, tx = blockIdx.x * CO_T;
  const int c4n = C >> 2;
  {
    // GroupNorm(32) finalised from the producer's per-(sample, channel) fp64 sums
    const int cpg = C / 32;
    for (int c = threadIdx.x; c < C; c += 256) {
      const int g = c / cpg;
      double ts = 0.0, tq = 0.0;
      for (int i = 0; i < cpg; ++i) {
        const double* a = stats + (static_cast<long long>(b) * C + g * cpg + i) * 2;
        ts += a[0];
        tq += a[1];
      }
      const double n = static_cast<double>(H) * W * cpg;
      const double mean = ts / n;
      double var = tq / n - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
      const float s1 = gamma[c] * rstd;
      sc[c] = s1;
      sh[c] = beta[c] - static_cast<float>(mean) * s1;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < Cout * 9 * C; i += 256) {
    // w [Cout][C][3][3] -> sw[co][tap][c]
    const int co = i / (9 * C), r = i % (9 * C);
    const int tap = r / C, c = r % C;
    sw[i] = w[(static_cast<long long>(co) * C + c) * 9 + tap];
  }
  const int nfill = (CO_T + 2) * (CO_T + 2) * c4n;
  for (int i0 = threadIdx.x; i0 < nfill; i0 += 256 * 4) {
    float4 a[4];
    int ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {  // issue the four global loads first (latency overlap)
      const int i = i0 + u * 256;
      ok[u] = 0;
      a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < nfill) {
        const int pix = i / c4n, c = (i % c4n) * 4;
        const int gy = ty + pix / (CO_T + 2) - 1, gx = tx + pix % (CO_T + 2) - 1;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
          ok[u] = 1;
          a[u] = __ldg(reinterpret_cast<const float4*>(
              h + ((static_cast<long long>(b) * H + gy) * W + gx) * C + c));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * 256;
      if (i >= nfill) break;
      const int pix = i / c4n, c = (i % c4n) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok[u]) {
        v.x = silu_fast(fmaf(a[u].x, sc[c + 0], sh[c + 0]));
        v.y = silu_fast(fmaf(a[u].y, sc[c + 1], sh[c + 1]));
        v.z = silu_fast(fmaf(a[u].z, sc[c + 2], sh[c + 2]));
        v.w = silu_fast(fmaf(a[u].w, sc[c + 3], sh[c + 3]));
      }
      *reinterpret_cast<float4*>(st + pix * pitch + c) = v;
    }
  }
  __syncthreads();
  const int py = threadIdx.x / CO_T, px = threadIdx.x % CO_T;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int tap = 0; tap < 9; ++tap) {
    const float* ip = st + ((py + tap / 3) * (CO_T + 2) + px + tap % 3) * pitch;
    for (int c = 0; c < C; c += 4) {
      const float4 v = *reinterpret_cast<const float4*>(ip + c);
#pragma unroll
      for (int co = 0; co < 4; ++co)
        if (co < Cout) {
          const float4 ww = *reinterpret_cast<const float4*>(sw + (co * 9 + tap) * C + c);
          acc[co] = fmaf(v.x, ww.x, acc[co]);
          acc[co] = fmaf(v.y, ww.y, acc[co]);
          acc[co] = fmaf(v.z, ww.z, acc[co]);
          acc[co] = fmaf(v.w, ww.w, acc[co]);
        }
    }
  }
  const int gy = ty + py, gx = tx + px;
  if (gy < H && gx < W)
    for (int co = 0; co < Cout; ++co)
      out[((static_cast<long long>(b) * Cout + co) * H + gy) * W + gx] = acc[co] + bias[co];
}

// Fast path (C = 64, Cout = 2, W <= 128): scatter form.  The halo-tile kernel above reads every
// activated input value 9 times from shared memory (one LDS.128 per 8 FMAs: shared-memory bound,
// 370 us at batch 64).  Here a block walks the input rows of a 16-row strip once; each activated
// pixel is read ONCE and turned into its 9 x 2 tap partials  part[tap][co] = sum_c v[c] w[co][c][tap],
// kept for three rows in a ring; output row y then gathers 9 partials per (x, co):
//   out[y][x][co] = bias[co] + sum_{ky,kx} part(row y+ky-1)[ky*3+kx][co][x+kx-1].
// thread = (x, co); the weights are warp-uniform (broadcast LDS.128), pixel reads are conflict-free
// (pitch 68 floats).
constexpr int CO2_ROWS = 16, CO2_PITCH = 68, CO2_PW = 130;
constexpr int CO2_SMEM = (128 * CO2_PITCH + 3 * 9 * 2 * CO2_PW + 2 * 9 * 64 + 128) * 4;
template <bool FUSE>
__global__ void __launch_bounds__(256) conv_out2_kernel(const float* __restrict__ h,
                                                        const double* __restrict__ stats,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps,
                                                        const float* __restrict__ w,
                                                        const float* __restrict__ bias,
                                                        float* __restrict__ out, int H, int W,
                                                        const FusedStep fs) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];
  float* st = sm;                          // [128][68] activated input row
  float* part = st + 128 * CO2_PITCH;      // [3][9][2][130]
  float* sw = part + 3 * 9 * 2 * CO2_PW;   // [2][9][64]
  float* sc = sw + 2 * 9 * 64;             // [64] scale, [64] shift
  float* sh = sc + 64;
  const int b = blockIdx.y, y0 = blockIdx.x * CO2_ROWS;
  const int y1 = min(H, y0 + CO2_ROWS);
  if (threadIdx.x < 64) {
    // GroupNorm(32, 64): 2 channels per group, from the producer's per-(sample, channel) fp64 sums
    const int c = threadIdx.x, g = c >> 1;
    const double* a = stats + (static_cast<long long>(b) * 64 + g * 2) * 2;
    const double ts = a[0] + a[2], tq = a[1] + a[3];
    const double n = static_cast<double>(H) * W * 2;
    const double mean = ts / n;
    double var = tq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float s1 = gamma[c] * rstd;
    sc[c] = s1;
    sh[c] = beta[c] - static_cast<float>(mean) * s1;
  }
  for (int i = threadIdx.x; i < 2 * 9 * 64; i += 256) {
    // w [2][64][3][3] -> sw[co][tap][c]
    const int co = i / 576, r = i % 576, tap = r >> 6, c = r & 63;
    sw[i] = __ldg(w + (co * 64 + c) * 9 + tap);
  }
  for (int i = threadIdx.x; i < 3 * 9 * 2 * CO2_PW; i += 256) part[i] = 0.f;  // incl. the x = -1 / W columns
  __syncthreads();
  const int x = threadIdx.x & 127, co = threadIdx.x >> 7;
  const float bco = __ldg(bias + co);
  // the raw values of input row r + 1 are fetched into registers while row r is being reduced
  float4 a[8];
  auto fetch = [&](int r) {
    const float4* src = reinterpret_cast<const float4*>(h + (static_cast<long long>(b) * H + r) * W * 64);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = threadIdx.x + 256 * u;
      a[u] = (idx < W * 16) ? __ldg(src + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  if (y0 - 1 >= 0) fetch(y0 - 1);
  for (int r = y0 - 1; r <= y1; ++r) {
    const int slot = (r + 3) % 3;
    const bool inside = (r >= 0 && r < H);
    if (inside) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = threadIdx.x + 256 * u;
        const int pix = idx >> 4, c = (idx & 15) * 4;
        float4 v;
        v.x = silu_fast(fmaf(a[u].x, sc[c + 0], sh[c + 0]));
        v.y = silu_fast(fmaf(a[u].y, sc[c + 1], sh[c + 1]));
        v.z = silu_fast(fmaf(a[u].z, sc[c + 2], sh[c + 2]));
        v.w = silu_fast(fmaf(a[u].w, sc[c + 3], sh[c + 3]));
        *reinterpret_cast<float4*>(st + pix * CO2_PITCH + c) = v;
      }
    }
    if (r + 1 <= y1 && r + 1 >= 0 && r + 1 < H) fetch(r + 1);
    __syncthreads();
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    if (inside && x < W) {
      const float* ip = st + x * CO2_PITCH;
      const float* wp = sw + co * 576;
#pragma unroll 4
      for (int c = 0; c < 64; c += 4) {
        const float4 v = *reinterpret_cast<const float4*>(ip + c);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 ww = *reinterpret_cast<const float4*>(wp + t * 64 + c);
          acc[t] = fmaf(v.x, ww.x, acc[t]);
          acc[t] = fmaf(v.y, ww.y, acc[t]);
          acc[t] = fmaf(v.z, ww.z, acc[t]);
          acc[t] = fmaf(v.w, ww.w, acc[t]);
        }
      }
    }
    if (x < W) {
#pragma unroll
      for (int t = 0; t < 9; ++t) part[((slot * 9 + t) * 2 + co) * CO2_PW + x + 1] = acc[t];
    }
    __syncthreads();
    const int y = r - 1;
    if (y >= y0 && y < y1 && x < W) {
      float sum = bco;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int sl = (y + ky - 1 + 3) % 3;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
          sum += part[((sl * 9 + ky * 3 + kx) * 2 + co) * CO2_PW + x + kx];
      }
      if constexpr (FUSE) {
        // eps never leaves the SM: x_t -> x_{t-1} right here (north_star: "fused into the last kernel")
        fused_step_apply(fs, b, 2ll * H * W, (static_cast<long long>(co) * H + y) * W + x, sum);
      } else {
        out[((static_cast<long long>(b) * 2 + co) * H + y) * W + x] = sum;
      }
    }
  }
}

void launch_conv_out(const float* h, const double* stats, const float* gamma, const float* beta,
                     float eps, const float* w, const float* bias, float* out, int B, int H, int W,
                     int C, int Cout, cudaStream_t s, const FusedStep* fs) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(conv_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_out2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CO2_SMEM);
    cudaFuncSetAttribute(conv_out2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CO2_SMEM);
    attr_set = true;
  }
  static const bool slow = std::getenv("PF_CONV_SLOW") != nullptr;
  const bool fuse = fs && fs->kind != 0;
  if (!slow && C == 64 && Cout == 2 && W <= 128) {
    dim3 grid((H + CO2_ROWS - 1) / CO2_ROWS, B);
    FusedStep f{};
    if (fuse) f = *fs;
    if (fuse)
      launch_pdl(conv_out2_kernel<true>, grid, dim3(256), CO2_SMEM, s, h, stats, gamma, beta, eps, w, bias, out, H, W, f);
    else
      launch_pdl(conv_out2_kernel<false>, grid, dim3(256), CO2_SMEM, s, h, stats, gamma, beta, eps, w, bias, out, H, W, f);
    return;
  }
  const size_t smem = (static_cast<size_t>(CO_T + 2) * (CO_T + 2) * (C + 4) +
                       static_cast<size_t>(Cout) * 9 * C + 2 * C) * sizeof(float);
  dim3 grid((W + CO_T - 1) / CO_T, (H + CO_T - 1) / CO_T, B);
  launch_pdl(conv_out_kernel, grid, dim3(256), smem, s, h, stats, gamma, beta, eps, w, bias, out, H, W, C, Cout);
}

// ------------------------------------------------------------------------------------------------
// Condition encoders (reference dl_modules/chord_enc.py:5-22 RnnEncoder, dl_modules/txt_enc.py:5-35
// TextureEncoder): the step before the sampling loop.  fp32 throughout, precise expf / tanhf.
// ------------------------------------------------------------------------------------------------
// One GRU step (torch.nn.GRU gate order r, z, n):  gi = W_ih x_t + b_ih (row t of a [B, T, 3H]
// tensor), gh = W_hh h + b_hh;  r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r * gh_n),
// h' = (1 - z) * n + z * h.
__global__ void gru_cell_kernel(const float* __restrict__ gi, long long gi_ld, const float* __restrict__ gh,
                                const float* __restrict__ h, float* __restrict__ h_out, long long out_ld,
                                int B, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, j = i % H;
  const float* gib = gi + b * gi_ld;
  const float* ghb = gh + static_cast<long long>(b) * 3 * H;
  const float r = 1.0f / (1.0f + expf(-(gib[j] + ghb[j])));
  const float z = 1.0f / (1.0f + expf(-(gib[H + j] + ghb[H + j])));
  const float n = tanhf(gib[2 * H + j] + r * ghb[2 * H + j]);
  h_out[b * out_ld + j] = (1.0f - z) * n + z * h[static_cast<long long>(b) * H + j];
}
void launch_gru_cell(const float* gi, long long gi_ld, const float* gh, const float* h, float* h_out,
                     long long out_ld, int B, int H, cudaStream_t s) {
  gru_cell_kernel<<<(B * H + 255) / 256, 256, 0, s>>>(gi, gi_ld, gh, h, h_out, out_ld, B, H);
}

// TextureEncoder.cnn: Conv2d(1, C, (4, 12), stride (4, 1)) -> ReLU -> MaxPool2d((1, 4), (1, 4)) on a
// [B, 32, 128] piano roll -> [B, C, 8, 29] (conv width 117, pooled 29).  One thread per output.
__global__ void txt_cnn_kernel(const float* __restrict__ pr, const float* __restrict__ w,
                               const float* __restrict__ bias, float* __restrict__ out, int B, int C, int T,
                               int P) {
  const int Ho = T / 4, Wc = P - 12 + 1, Wo = Wc / 4;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(B) * C * Ho * Wo) return;
  const int j = static_cast<int>(idx % Wo), i = static_cast<int>((idx / Wo) % Ho);
  const int c = static_cast<int>((idx / (static_cast<long long>(Wo) * Ho)) % C);
  const int b = static_cast<int>(idx / (static_cast<long long>(Wo) * Ho * C));
  const float* in = pr + (static_cast<long long>(b) * T + 4 * i) * P;
  const float* wc = w + c * 48;
  float best = -INFINITY;
  for (int jj = 0; jj < 4; ++jj) {
    const int x = 4 * j + jj;
    float acc = 0.f;
    for (int ky = 0; ky < 4; ++ky)
      for (int kx = 0; kx < 12; ++kx) acc = fmaf(in[ky * P + x + kx], wc[ky * 12 + kx], acc);
    best = fmaxf(best, acc + bias[c]);
  }
  out[idx] = fmaxf(best, 0.f);
}
void launch_txt_cnn(const float* pr, const float* w, const float* bias, float* out, int B, int C, int T,
                    int P, cudaStream_t s) {
  const long long n = static_cast<long long>(B) * C * (T / 4) * ((P - 11) / 4);
  txt_cnn_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(pr, w, bias, out, B, C, T, P);
}

// ------------------------------------------------------------------------------------------------
// prmat2c -> prmat / note list (reference utils.py:240-269 prmat2c_to_prmat and the note loop of
// utils.py:446-470 prmat2c_to_midi_file).  Integer work, bit-exact:
//   on(s, key)  = int(round(onset[s, key])) > 0      (Python round = half to even = rintf)
//   dur(s, key) = 1 + number of consecutive steps s+1, s+2, .. < T with int(round(sustain)) > 0
// One thread per (segment, key) column walks the T steps backwards carrying the sustain run length,
// so consecutive threads touch consecutive keys (coalesced) and every cell is read once.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) prmat2c_dur_kernel(const float* __restrict__ x,
                                                          long long* __restrict__ out, int N, int C, int T,
                                                          int P) {
  const long long idx = static_cast<long long>(blockIdx.x) * 128 + threadIdx.x;
  if (idx >= static_cast<long long>(N) * P) return;
  const int n = static_cast<int>(idx / P), key = static_cast<int>(idx % P);
  const float* on = x + (static_cast<long long>(n) * C + 0) * T * P + key;
  const float* su = x + (static_cast<long long>(n) * C + 1) * T * P + key;
  long long* o = out + static_cast<long long>(n) * T * P + key;
  int run = 0;  // consecutive sustained steps starting at s + 1
  for (int s = T - 1; s >= 0; --s) {
    o[static_cast<long long>(s) * P] = (rintf(__ldg(on + static_cast<long long>(s) * P)) > 0.f) ? 1 + run : 0;
    run = (rintf(__ldg(su + static_cast<long long>(s) * P)) > 0.f) ? run + 1 : 0;
  }
}
void launch_prmat2c_dur(const float* x, long long* out, int N, int C, int T, int P, cudaStream_t s) {
  const long long cols = static_cast<long long>(N) * P;
  prmat2c_dur_kernel<<<static_cast<unsigned>((cols + 127) / 128), 128, 0, s>>>(x, out, N, C, T, P);
}

// note compaction in the reference's loop order (segment, step, key): per-row counts (warp per row,
// ballots), exclusive scan of the row counts, ordered write of (row, key, dur) triples.
__global__ void prmat_row_count_kernel(const long long* __restrict__ dur, int* __restrict__ counts,
                                       long long rows, int P) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  int c = 0;
  for (int k0 = 0; k0 < P; k0 += 32) {
    const int k = k0 + lane;
    const bool on = k < P && dur[row * P + k] > 0;
    c += __popc(__ballot_sync(0xffffffffu, on));
  }
  if (lane == 0) counts[row] = c;
}
// single block, in place: counts[0 .. rows) -> exclusive offsets, counts[rows] = total
__global__ void __launch_bounds__(1024) prmat_scan_kernel(int* __restrict__ counts, long long rows) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (long long base = 0; base < rows; base += 1024) {
    const long long i = base + threadIdx.x;
    const int v = i < rows ? counts[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
      int ws = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, ws, o);
        if (lane >= o) ws += t;
      }
      s_warp[lane] = ws;  // inclusive warp totals
    }
    __syncthreads();
    const int carry = s_carry;
    const int excl = carry + (w ? s_warp[w - 1] : 0) + incl - v;
    if (i < rows) counts[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_warp[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[rows] = s_carry;
}
__global__ void prmat_write_notes_kernel(const long long* __restrict__ dur, const int* __restrict__ offsets,
                                         int* __restrict__ notes, long long rows, int P, long long cap) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  long long pos = offsets[row];
  for (int k0 = 0; k0 < P; k0 += 32) {
    const int k = k0 + lane;
    const long long d = k < P ? dur[row * P + k] : 0;
    const unsigned m = __ballot_sync(0xffffffffu, d > 0);
    if (d > 0) {
      const long long at = pos + __popc(m & ((1u << lane) - 1u));
      if (at < cap) {
        notes[at * 3 + 0] = static_cast<int>(row);
        notes[at * 3 + 1] = k;
        notes[at * 3 + 2] = static_cast<int>(d);
      }
    }
    pos += __popc(m);
  }
}
void launch_prmat_notes(const long long* dur, int* offsets, int* notes, long long rows, int P, long long cap,
                        cudaStream_t s) {
  const unsigned blocks = static_cast<unsigned>((rows * 32 + 255) / 256);
  prmat_row_count_kernel<<<blocks, 256, 0, s>>>(dur, offsets, rows, P);
  prmat_scan_kernel<<<1, 1024, 0, s>>>(offsets, rows);
  if (notes && cap > 0) prmat_write_notes_kernel<<<blocks, 256, 0, s>>>(dur, offsets, notes, rows, P, cap);
}

// ------------------------------------------------------------------------------------------------
// weight packing / misc
// ------------------------------------------------------------------------------------------------
// f16f8 weight element (common.cuh): h16 at element o of the fp16 tensor; fp8 row bytes
// [row][Cin / 64][l8 x 64 | h8 x 64] (the ACTIVATION rows are [h8 | l8], so one K pass pairs
// act.h8 with w.l8 and act.l8 with w.h8)
__device__ __forceinline__ void store_weight_f8(float v, bf16* o16, bf16* o8, long long rowoff, int ci) {
  const __half h = __float2half_rn(v);
  const float hf = __half2float(h);
  reinterpret_cast<__half*>(o16)[rowoff + ci] = h;
  const uint32_t q = pack_e4m3x4(hf * F8_W_HI_SCALE, (v - hf) * F8_W_LO_SCALE, 0.f, 0.f);
  uint8_t* p8 = reinterpret_cast<uint8_t*>(o8) + rowoff * 2 + (ci >> 6) * 128 + (ci & 63);
  p8[0] = static_cast<uint8_t>((q >> 8) & 0xffu);   // l8
  p8[64] = static_cast<uint8_t>(q & 0xffu);         // h8
}

__global__ void pack_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out_hi,
                                   bf16* __restrict__ out_lo, int Cout, int Cin, int taps,
                                   int cout_total, int row0, int geglu_gran, int fmt8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(Cout) * Cin * taps;
  if (i >= total) return;
  // destination-ordered index: (tap, co, ci)
  const int ci = static_cast<int>(i % Cin);
  const int co = static_cast<int>((i / Cin) % Cout);
  const int tap = static_cast<int>(i / (static_cast<long long>(Cin) * Cout));
  const float v = w[(static_cast<long long>(co) * Cin + ci) * taps + tap];
  bf16 h, l;
  split_bf16(v, h, l);
  int dst = row0 + co;
  if (geglu_gran > 0) {
    // GeGLU projection [x: F rows | gate: F rows] -> tiles of [gran x rows | gran gate rows]
    const int F = Cout / 2;
    const int half = co / F, r = co % F;
    dst = row0 + (r / geglu_gran) * 2 * geglu_gran + half * geglu_gran + r % geglu_gran;
  }
  const long long o = (static_cast<long long>(tap) * cout_total + dst) * Cin + ci;
  if (fmt8) {
    store_weight_f8(v, out_hi, out_lo, o - ci, ci);
    return;
  }
  out_hi[o] = h;
  out_lo[o] = l;
}
void launch_pack_weight(const float* w, bf16* out_hi, bf16* out_lo, int Cout, int Cin, int taps,
                        int cout_total, int row0, int geglu_gran, cudaStream_t s, int fmt8) {
  const long long total = static_cast<long long>(Cout) * Cin * taps;
  pack_weight_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(
      w, out_hi, out_lo, Cout, Cin, taps, cout_total, row0, geglu_gran, fmt8);
}

// UpSample conv (nearest 2x then 3x3, unet.py:231-238) as four 2x2 parity kernels over the
// low-resolution input: for output parity (py, px) and tap (a, b) the effective weight is the sum of
// the 3x3 taps that read the same source pixel.  out: [parity 4][tap 4][Cout][Cin] split bf16.
__global__ void pack_weight_up_kernel(const float* __restrict__ w, bf16* __restrict__ out_hi,
                                      bf16* __restrict__ out_lo, int Cout, int Cin, int fmt8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long per = static_cast<long long>(Cout) * Cin;
  if (i >= 16 * per) return;
  const int ci = static_cast<int>(i % Cin);
  const int co = static_cast<int>((i / Cin) % Cout);
  const int tap = static_cast<int>((i / per) % 4);
  const int par = static_cast<int>(i / (4 * per));
  const int py = par >> 1, px = par & 1, a = tap >> 1, b = tap & 1;
  // rows of the 3x3 kernel folded into tap a: py=0: {0} | {1,2};  py=1: {0,1} | {2}
  const int ky0 = py == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2);
  const int ky1 = py == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
  const int kx0 = px == 0 ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2);
  const int kx1 = px == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
  const float* wp = w + (static_cast<long long>(co) * Cin + ci) * 9;
  float v = 0.f;
  for (int ky = ky0; ky <= ky1; ++ky)
    for (int kx = kx0; kx <= kx1; ++kx) v += wp[ky * 3 + kx];
  if (fmt8) {
    store_weight_f8(v, out_hi, out_lo, i - ci, ci);
    return;
  }
  bf16 h, l;
  split_bf16(v, h, l);
  out_hi[i] = h;
  out_lo[i] = l;
}
void launch_pack_weight_up(const float* w, bf16* out_hi, bf16* out_lo, int Cout, int Cin,
                           cudaStream_t s, int fmt8) {
  const long long total = 16LL * Cout * Cin;
  pack_weight_up_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(w, out_hi, out_lo,
                                                                                   Cout, Cin, fmt8);
}

// ------------------------------------------------------------------------------------------------
// get_mask("below" / "above") of inference_sdf.py:132-180 on the GPU, batched over songs.
// orig [n_seg, C, T, P] (channel 0 = onsets); each song = seg_per_song consecutive segments whose
// T-step rows form one sequence.  Per row: below -> first-maximum index of the onset row (0 when the
// row is empty), above -> P-1 - first-maximum index of the flipped row (P-1 when empty); rows whose
// value equals the "empty" value inherit the previous row's value (leading ones take the first
// non-empty value); mask[row, pitch >= v] = 1 (below) or mask[row, pitch <= v] = 1 (above).
// ------------------------------------------------------------------------------------------------
__global__ void mask_row_extreme_kernel(const float* __restrict__ orig, int* __restrict__ rowval,
                                        int C, int T, int P, int above, long long rows) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const long long seg = row / T, t = row % T;
  const float* rp = orig + ((seg * C + 0) * T + t) * P;
  // first index of the maximum, scanning pitches upward (below) or downward (above)
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int k = lane; k < P; k += 32) {
    const int pidx = above ? (P - 1 - k) : k;
    const float v = rp[pidx];
    if (v > best || (v == best && k < bi)) { best = v; bi = k; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) rowval[row] = above ? (P - 1 - bi) : bi;
}

__global__ void mask_fill_kernel(int* __restrict__ rowval, int rows_per_song, int n_songs, int empty,
                                 int* __restrict__ err) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_songs) return;
  int* v = rowval + static_cast<long long>(s) * rows_per_song;
  // first row whose value is non-zero (min_pitch.nonzero()[0] / max_pitch.nonzero()[0]); the
  // reference raises IndexError when there is none
  int first = -1;
  for (int i = 0; i < rows_per_song; ++i)
    if (v[i] != 0) { first = i; break; }
  if (first < 0) { atomicExch(err, 1); return; }
  const int fv = v[first];
  for (int i = 0; i < first; ++i) v[i] = fv;
  // rows with the "empty" value (0 below, P-1 above) inherit the previous row; index -1 wraps to the
  // last row exactly like the reference's Python indexing
  for (int i = 0; i < rows_per_song; ++i)
    if (v[i] == empty) v[i] = v[i > 0 ? i - 1 : rows_per_song - 1];
}

__global__ void mask_write_kernel(const int* __restrict__ rowval, float* __restrict__ mask, int C, int T,
                                  int P, int above, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= total) return;
  const int pch = static_cast<int>(i % P);
  const long long r = i / P;
  const int t = static_cast<int>(r % T);
  const long long sc = r / T;  // seg * C + c
  const long long seg = sc / C;
  const int v = rowval[seg * T + t];
  mask[i] = above ? (pch <= v ? 1.f : 0.f) : (pch >= v ? 1.f : 0.f);
}

int launch_get_mask(const float* orig, float* mask, int* rowval, int* err, int n_seg, int seg_per_song,
                    int C, int T, int P, int above, cudaStream_t s) {
  const long long rows = static_cast<long long>(n_seg) * T;
  const int n_songs = n_seg / seg_per_song;
  mask_row_extreme_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, s>>>(orig, rowval, C, T, P,
                                                                                above, rows);
  mask_fill_kernel<<<(n_songs + 63) / 64, 64, 0, s>>>(rowval, seg_per_song * T, n_songs,
                                                      above ? P - 1 : 0, err);
  const long long total = static_cast<long long>(n_seg) * C * T * P;
  mask_write_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(rowval, mask, C, T, P,
                                                                               above, total);
  return 0;
}

__global__ void vec_add_kernel(const float* a, const float* b, float* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + (b ? b[i] : 0.f);
}
void launch_vec_add(const float* a, const float* b, float* out, int n, cudaStream_t s) {
  vec_add_kernel<<<(n + 255) / 256, 256, 0, s>>>(a, b, out, n);
}

// ------------------------------------------------------------------------------------------------
// sampler step epilogues.  Every fp32 op is individually rounded (__fmul_rn/__fadd_rn, no FMA
// contraction) in the order the reference's ATen elementwise ops run, so that given identical
// inputs the result is the same fp32 value the reference computes.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float guided_eps(const StepArgs& a, long long i) {
  const float ec = a.e_cond[i];
  if (!a.e_uncond) return ec;
  const float eu = a.e_uncond[i];
  // e_u + s * (e_c - e_u)   (sampler/__init__.py:77)
  return __fadd_rn(eu, __fmul_rn(a.uncond_scale, __fsub_rn(ec, eu)));
}
__device__ __forceinline__ float step_noise(const StepArgs& a, long long i) {
  if (!a.noise) return 0.f;
  const float nz = a.noise_bcast > 0 ? a.noise[i % a.noise_bcast] : a.noise[i];
  return __fmul_rn(nz, a.temperature);
}
__device__ __forceinline__ float repaint_blend(const StepArgs& a, long long i, float x_unkn) {
  if (!a.orig) return x_unkn;
  // x_kn = kn_a*orig + kn_b*noise ; x = x_kn*mask + x_unkn*(1-mask)  (sampler_sdf.py:192,336)
  const float nk = a.noise_kn ? a.noise_kn[i] : 0.f;
  const float xk = __fadd_rn(__fmul_rn(a.kn_a, a.orig[i]), __fmul_rn(a.kn_b, nk));
  const float m = a.mask[i];
  return __fadd_rn(__fmul_rn(xk, m), __fmul_rn(x_unkn, __fsub_rn(1.0f, m)));
}

__global__ void __launch_bounds__(256) step_ddpm_kernel(const StepArgs a) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= a.n) return;
  const float x = a.x[i];
  const float e = guided_eps(a, i);
  // c0 = sqrt_recip_alpha_bar, c1 = sqrt_recip_m1_alpha_bar, c2 = mean_x0_coef, c3 = mean_xt_coef,
  // c4 = exp(0.5*log_var)
  const float x0 = __fsub_rn(__fmul_rn(a.c0, x), __fmul_rn(a.c1, e));
  const float mean = __fadd_rn(__fmul_rn(a.c2, x0), __fmul_rn(a.c3, x));
  const float xp = __fadd_rn(mean, __fmul_rn(a.c4, step_noise(a, i)));
  a.x_prev[i] = repaint_blend(a, i, xp);
  if (a.x0) a.x0[i] = x0;
  if (a.e_t) a.e_t[i] = e;
}
void launch_step_ddpm(const StepArgs& a, cudaStream_t s) {
  step_ddpm_kernel<<<static_cast<unsigned>((a.n + 255) / 256), 256, 0, s>>>(a);
}

__global__ void __launch_bounds__(256) step_ddim_kernel(const StepArgs a) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= a.n) return;
  const float x = a.x[i];
  const float e = guided_eps(a, i);
  // c0 = sqrt(1-alpha), c1 = alpha**0.5, c2 = alpha_prev**0.5, c3 = sqrt(1-alpha_prev-sigma^2),
  // c4 = sigma
  const float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(a.c0, e)), a.c1);
  const float dir = __fmul_rn(a.c3, e);
  float xp = __fadd_rn(__fmul_rn(a.c2, x0), dir);
  xp = __fadd_rn(xp, __fmul_rn(a.c4, step_noise(a, i)));
  a.x_prev[i] = repaint_blend(a, i, xp);
  if (a.x0) a.x0[i] = x0;
  if (a.e_t) a.e_t[i] = e;
}
void launch_step_ddim(const StepArgs& a, cudaStream_t s) {
  step_ddim_kernel<<<static_cast<unsigned>((a.n + 255) / 256), 256, 0, s>>>(a);
}

__global__ void __launch_bounds__(256) step_ddpm_legacy_kernel(const StepArgs a) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= a.n) return;
  // c0 = (1-alpha)/sqrt(1-alpha_bar), c1 = 1/sqrt(alpha), c2 = sqrt(sigma2)  (ddpm/__init__.py:79-88)
  const float mean = __fmul_rn(a.c1, __fsub_rn(a.x[i], __fmul_rn(a.c0, a.e_cond[i])));
  a.x_prev[i] = __fadd_rn(mean, __fmul_rn(a.c2, a.noise ? a.noise[i] : 0.f));
}
void launch_step_ddpm_legacy(const StepArgs& a, cudaStream_t s) {
  step_ddpm_legacy_kernel<<<static_cast<unsigned>((a.n + 255) / 256), 256, 0, s>>>(a);
}

__global__ void __launch_bounds__(256) q_sample_kernel(const float* __restrict__ x0,
                                                       const float* __restrict__ noise,
                                                       float* __restrict__ out, long long n, float a,
                                                       float b) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  out[i] = __fadd_rn(__fmul_rn(a, x0[i]), __fmul_rn(b, noise[i]));
}
void launch_q_sample(const float* x0, const float* noise, float* out, long long n, float a, float b,
                     cudaStream_t s) {
  q_sample_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(x0, noise, out, n, a, b);
}

}  // namespace pf
