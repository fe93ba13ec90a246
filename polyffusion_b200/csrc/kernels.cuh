// Non-GEMM kernels of the UNet hot path (all HBM-bound or tiny): GroupNorm statistics, the
// "operand transform" passes that turn fp32 NHWC activations into split-bf16 GEMM operands
// (fusing GN-apply / SiLU / LayerNorm / GeGLU / softmax / concat / 2x-upsample / stride-2
// re-layout), the fp32 edge convolutions (Cin=2 in, Cout=2 out), small per-sample linears, and the
// sampler step epilogues.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace pf {

typedef __nv_bfloat16 bf16;

// x NCHW [B,Cin,H,W] -> out NHWC [B,H,W,Cout], 3x3 pad 1, fp32 (unet.py:79 first conv); stats
// (optional, zero-initialised [B][Cout][2] fp64) receives the per-(sample, channel) sum / sum of
// squares of the output (GroupNorm statistics for its consumers)
void launch_conv_in(const float* x, const float* w, const float* bias, float* out, double* stats, int B,
                    int Cin, int H, int W, int Cout, cudaStream_t s);

// per-(sample, channel) sum / sum-of-squares (fp64 atomics) of an NHWC tensor [B,HW,Cs] written at
// channel offset coff of acc [B, Ctot, 2]  (only used for the first conv's output; every other
// feature map gets its statistics from the producing GEMM's epilogue)
void launch_gn_stats(const float* src, double* acc, int B, int HW, int Cs, int Ctot, int coff,
                     cudaStream_t s);
// scale[b][c] = gamma[c] * rstd(b, group(c)), shift[b][c] = beta[c] - mean * scale from the fp64 sums
// [B][C][2] (the same arithmetic as the operand transform's prologue); operand of a RAW GEMM segment
void launch_gn_finalize(const double* stats, const float* gamma, const float* beta, float eps, int groups,
                        int B, int HW, int C, float* scale, float* shift, cudaStream_t s);
enum XformLayout : int { XF_SAME = 0, XF_UP2 = 1, XF_S2D = 2 };
// Operand transform: out_hi/lo[b, y', x', c] = split(act(GN(cat(src0, src1)))); act = SiLU if silu.
// GroupNorm(groups, eps, gamma, beta) is applied when stats0 != null, using the per-(sample, channel)
// fp64 sum / sum-of-squares buffers of each source ([B][C0][2], [B][C1][2]).  out2 (optional) receives
// the plain split of the un-normalised input.  (C0 + C1) % 16 == 0, C0 % 16 == 0, C0 + C1 <= 512.
struct ActSplitArgs {
  const float* src0; int C0;
  const float* src1; int C1;
  const double* stats0; const double* stats1;
  const float* gamma; const float* beta;
  float eps; int groups;
  int silu, layout;
  bf16* out_hi; bf16* out_lo;
  bf16* out2_hi; bf16* out2_lo;
  int B, H, W;
  int fmt8;  // 1: outputs are f16f8 operands (out_hi = fp16 [pixel][C], out_lo = fp8 rows), see common.cuh
};
void launch_act_split(const ActSplitArgs& a, cudaStream_t s);

// LayerNorm over C (eps) + affine -> split bf16 [rows, C]; C multiple of 128, <= 512
// fmt8 = 1: the output is an f16f8 operand (out_hi = fp16, out_lo = fp8 rows; common.cuh)
void launch_ln_split(const float* src, const float* gamma, const float* beta, float eps,
                     bf16* out_hi, bf16* out_lo, long long rows, int C, cudaStream_t s, int fmt8 = 0);
// GeGLU: in [rows, 2F] -> split(in[:, :F] * gelu_erf(in[:, F:])) [rows, F]
void launch_geglu_split(const float* src, bf16* out_hi, bf16* out_lo, long long rows, int F,
                        cudaStream_t s);
// softmax(scale * S) over last dim Nk (multiple of 128, <= 1024) -> split bf16
void launch_softmax_split(const float* S, float scale, bf16* out_hi, bf16* out_lo, long long rows,
                          int Nk, cudaStream_t s);

// sinusoidal timestep embedding [B, 2*half] = [cos(t f_i), sin(t f_i)] (unet.py:151-169)
void launch_time_sinusoid(const long long* t, const float* freqs, float* out, int B, int half,
                          cudaStream_t s);
// generic blocks of the legacy ddpm.unet.UNet evaluation (kernels.cu)
void launch_groupnorm_generic(const float* x, const float* gamma, const float* beta, float eps, int silu,
                              float* out, int B, int HW, int C, int groups, cudaStream_t s);
void launch_softmax_rows(const float* S, float scale, float* out, long long rows, int n, cudaStream_t s);
void launch_conv3x3_direct(const float* x, const float* w, const float* bias, float* out, int B, int Cin,
                           int H, int W, int Cout, int in_nchw, int out_nchw, cudaStream_t s);
void launch_time_sincos(const long long* t, const float* freqs, float* out, int B, int half, cudaStream_t s);
// test helpers
void launch_merge_split(const bf16* hi, const bf16* lo, float* out, long long n, cudaStream_t s);
void launch_transpose_split(const float* src, bf16* hi, bf16* lo, int imgs, int rows, int C,
                            cudaStream_t s);
// out[b, n] = act(W[n,:] . in[b,:] + bias[n]); out_act: 0 none, 1 SiLU.  fp32.
void launch_small_linear(const float* in, long long ld_in, const float* W, const float* bias,
                         float* out, long long ld_out, int B, int N, int K, int out_act,
                         cudaStream_t s, int groups = 1);

// Reverse-diffusion step applied to eps where it is produced (the last kernel of the UNet evaluation),
// driven entirely by device-side state so that a whole sampling step replays as one CUDA graph:
//   index : device int32[2]: [0] row of `coef` for this replay (DDPM: the step; DDIM: the schedule index),
//           [1] run nonce folded into the Philox key (seed + nonce * 0x9E3779B97F4A7C15)
//   coef  : [n_index][8] = c0..c4 (meaning as in StepArgs), kn_a, kn_b, unused
//   x     : x_t in, x_{t-1} out (in place; NCHW like eps)
//   noise / noise_kn : injected N(0,1) tensors, or null -> Philox4x32-10 normals keyed by
//           (seed, global sample index = sample0 + b, element, index, stream) -- independent of how the
//           batch is sharded over ranks
//   orig / mask : RePaint known region (null -> plain step)
// DDPM adds no noise and re-noises the known region with zero noise at index 0 (sampler_sdf.py:152-153,
// 322-324).  Arithmetic and rounding order are those of step_ddpm_kernel / step_ddim_kernel.
struct FusedStep {
  int kind;  // 0 none, 1 DDPM (sampler_sdf.py:121-171, 322-336), 2 DDIM (sampler_ddim.py:233-272, 355-359)
  const int* index;
  const float* coef;
  float* x;
  float* eps_out;  // optional copy of eps
  const float* noise;
  const float* noise_kn;
  const float* orig;
  const float* mask;
  float temperature;
  unsigned long long seed;
  long long sample0;
};

// GroupNorm-apply + SiLU + conv3x3 (C -> Cout small) -> NCHW out (unet.py:145-149); with fs (kind != 0)
// the reverse step is applied in the same kernel and `out` is not written
void launch_conv_out(const float* h, const double* stats, const float* gamma, const float* beta,
                     float eps, const float* w, const float* bias, float* out, int B, int H, int W,
                     int C, int Cout, cudaStream_t s, const FusedStep* fs = nullptr);
// the same step from an eps tensor [B, per_sample] (fallback for shapes without the fused conv kernel)
void launch_step_from_eps(const FusedStep& fs, const float* eps, int B, long long per_sample, cudaStream_t s);
// index = max(index - 1, 0); t[b] = t_table[index]  (end of a fused step)
void launch_step_advance(int* index, long long* t, const long long* t_table, int B, cudaStream_t s);
// out[b, i] = Philox normal (same generator as FusedStep): N(0,1) keyed by (seed, sample0 + b, i, index, which)
void launch_fill_normal(float* out, long long n_samples, long long per_sample, unsigned long long seed,
                        long long sample0, int index, int which, cudaStream_t s);
// dst[b, :] = table[clamp(t[b], 0, n_rows - 1), :]   (time-embedding LUT gather)
void launch_gather_rows(const long long* t, const float* table, float* dst, int B, int n_rows, int width,
                        cudaStream_t s);

// condition encoders: one GRU step, and TextureEncoder.cnn (conv (4,12)/(4,1) + ReLU + maxpool (1,4))
void launch_gru_cell(const float* gi, long long gi_ld, const float* gh, const float* h, float* h_out,
                     long long out_ld, int B, int H, cudaStream_t s);
void launch_txt_cnn(const float* pr, const float* w, const float* bias, float* out, int B, int C, int T,
                    int P, cudaStream_t s);

// prmat2c [N, C>=2, T, P] fp32 (channel 0 onset, 1 sustain) -> note durations [N*T, P] int64
// (reference utils.py:240-269); then (row, key, dur) triples in (segment, step, key) order:
// offsets [rows + 1] receives the exclusive row offsets and the total, notes [cap][3] may be null
void launch_prmat2c_dur(const float* x, long long* out, int N, int C, int T, int P, cudaStream_t s);
void launch_prmat_notes(const long long* dur, int* offsets, int* notes, long long rows, int P, long long cap,
                        cudaStream_t s);

// weight packing: w [Cout, Cin, kh, kw] fp32 -> split bf16 [kh*kw][Cout_total][Cin] rows at row0
// geglu_gran > 0: interleave the [x | gate] halves of a GeGLU projection in blocks of geglu_gran rows
void launch_pack_weight(const float* w, bf16* out_hi, bf16* out_lo, int Cout, int Cin, int taps,
                        int cout_total, int row0, int geglu_gran, cudaStream_t s, int fmt8 = 0);
// UpSample conv weights w [Cout, Cin, 3, 3] -> four 2x2 parity kernels [parity 4][tap 4][Cout][Cin]
void launch_pack_weight_up(const float* w, bf16* out_hi, bf16* out_lo, int Cout, int Cin,
                           cudaStream_t s, int fmt8 = 0);
// get_mask("below"/"above") of inference_sdf.py:132-180, batched over songs (see kernels.cu).
// rowval: scratch [n_seg * T] ints; err: device int set to 1 if a song has no onset at all.
int launch_get_mask(const float* orig, float* mask, int* rowval, int* err, int n_seg, int seg_per_song,
                    int C, int T, int P, int above, cudaStream_t s);
// out[i] = a[i] + b[i] (bias pre-combination); b may be null
void launch_vec_add(const float* a, const float* b, float* out, int n, cudaStream_t s);

// ------------------------------------------------------------------ sampler step epilogues
struct StepArgs {
  // all tensors [n] fp32 contiguous (NCHW flattened); optional ones may be null
  const float* x;        // x_t
  const float* e_cond;   // eps (conditional half, or the only eps)
  const float* e_uncond; // eps for the unconditional half (CFG), or null
  const float* noise;    // N(0,1) for the reverse step, or null (=> 0)
  const float* orig;     // RePaint known image, or null
  const float* mask;     // RePaint mask
  const float* noise_kn; // noise used to diffuse `orig`
  float* x_prev;         // output
  float* x0;             // optional output
  float* e_t;            // optional output (guided eps)
  long long n;
  long long noise_bcast; // if >0: noise has this many elements, broadcast over batch (repeat_noise)
  float uncond_scale;
  // DDPM (sampler_sdf.py:121-171): x0 = c_rab*x - c_rm1*e; mean = c_x0*x0 + c_xt*x;
  //   x_prev = mean + exp(0.5*log_var)*noise*temperature
  // DDIM (sampler_ddim.py:233-272): x0 = (x - c_s1m*e)/sqrt(alpha); x_prev = sqrt(alpha_prev)*x0
  //   + sqrt(1-alpha_prev-sigma^2)*e + sigma*noise*temperature
  float c0, c1, c2, c3, c4;
  float temperature;
  float kn_a, kn_b;      // q_sample(orig): kn_a*orig + kn_b*noise_kn
};
void launch_step_ddpm(const StepArgs& a, cudaStream_t s);
void launch_step_ddim(const StepArgs& a, cudaStream_t s);
// legacy DDPM (ddpm/__init__.py:66-88): mean = (x - c0*e)*c1 ; x_prev = mean + c2*noise
void launch_step_ddpm_legacy(const StepArgs& a, cudaStream_t s);
// out = a*x0 + b*noise
void launch_q_sample(const float* x0, const float* noise, float* out, long long n, float a, float b,
                     cudaStream_t s);

}  // namespace pf
