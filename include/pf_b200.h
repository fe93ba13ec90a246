/*
 * pf_b200.h -- C ABI of libpf_b200.so: the B200-native (sm_100a) implementation of Polyffusion's
 * DDPM/DDIM sampling hot path.
 *
 * The reference (aik2mlj/polyffusion) is pure Python/PyTorch and has no FFI of its own; the
 * boundary it exposes is the Python module API.  This header is the thin
 * C ABI the Python drop-in classes in polyffusion_b200/ bind with ctypes.  Each entry point cites
 * the reference interface it replaces (paths relative to /root/reference/polyffusion/).
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a non-zero
 * code on failure (pf_last_error() gives the message); no C++ exceptions cross the ABI.  All
 * tensor pointers are DEVICE pointers unless the name ends in _host.  The caller owns every tensor
 * and the CUDA stream; the library owns packed weights and launch plans.  After the first call for
 * a given (batch, n_cond, workspace) triple, pf_unet_forward performs no allocation and no
 * host/device synchronisation, so a whole sampling step can be captured in a CUDA graph.
 */
#ifndef PF_B200_H
#define PF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pf_unet pf_unet;
typedef void* pf_stream; /* cudaStream_t */

/* Geometry of stable_diffusion/model/unet.py:35-47 UNetModel.__init__ (keys of params/sdf_*.yaml). */
typedef struct pf_unet_cfg {
  int32_t in_channels;
  int32_t out_channels;
  int32_t channels;
  int32_t n_res_blocks;
  int32_t n_levels;                 /* len(channel_multipliers) */
  int32_t channel_multipliers[8];
  int32_t attention_levels[8];      /* 1 if level i has SpatialTransformers */
  int32_t n_heads;
  int32_t tf_layers;
  int32_t d_cond;
} pf_unet_cfg;

const char* pf_last_error(void);
const char* pf_version(void);

/* UNetModel(**params)  -- stable_diffusion/model/unet.py:35 */
int pf_unet_create(const pf_unet_cfg* cfg, pf_unet** out);
void pf_unet_destroy(pf_unet* h);

/* load_state_dict: one call per tensor, `name` is the reference state_dict key
 * (e.g. "input_blocks.7.1.transformer_blocks.0.attn1.to_q.weight"); data is fp32, contiguous, on
 * the device; it is read during pf_unet_finalize and not referenced afterwards.
 * The special name "__time_freqs" carries the fp32 sinusoid frequency table
 * (unet.py:160-164, `channels // 2` entries). */
int pf_unet_set_weight(pf_unet* h, const char* name, const float* data, const int64_t* shape,
                       int32_t ndim);
/* Packs weights into split-bf16 tap-major GEMM operands, pre-combines biases, builds TMA maps. */
int pf_unet_finalize(pf_unet* h, pf_stream stream);

/* Bytes of scratch pf_unet_forward needs for this batch (images of height x width). */
size_t pf_unet_workspace_bytes(pf_unet* h, int32_t batch, int32_t n_cond, int32_t height,
                               int32_t width);

/* UNetModel.forward(x, time_steps, cond) -- stable_diffusion/model/unet.py:171-196.
 * x [B, in_channels, H, W] fp32 NCHW; time_steps [B] int64; cond [B, n_cond, d_cond] fp32;
 * out [B, out_channels, H, W] fp32 NCHW. */
int pf_unet_forward(pf_unet* h, const float* x, const int64_t* time_steps, const float* cond,
                    int32_t batch, int32_t n_cond, int32_t height, int32_t width, float* out,
                    void* workspace, size_t workspace_bytes, pf_stream stream);
/* One WHOLE reverse-diffusion step without host involvement (sampler_sdf.py:292-341, sampler_ddim.py:
 * 104-166 / 301-362 loop bodies): the UNet evaluation with the step arithmetic applied to eps inside its
 * last kernel (x0, posterior mean / DDIM direction, noise, RePaint blend with the re-noised known region),
 * then index -= 1 and time_steps[b] = t_table[index] on the device.  Everything that changes from step to
 * step lives in device memory, so a sampling loop is the replay of ONE captured CUDA graph:
 *   index    device int32[2]: [0] row of `coef` / `t_table` of this step (DDPM: the step; DDIM: schedule
 *            index), decremented by the call; [1] run nonce: the Philox key is seed + nonce * 0x9E3779B97F4A7C15
 *   coef     device [n_index][8] fp32: c0..c4 of pf_step_args for every index, then kn_a, kn_b, 0
 *   t_table  device [n_index] int64: timestep fed to the UNet at every index
 *   x        x_t in, x_{t-1} out, NCHW [B, out_channels, H, W] (x_in is the UNet input: x itself, or x
 *            concatenated with extra channels in a separate buffer)
 *   noise / noise_kn  injected N(0,1) tensors (the reference's torch.randn order), or NULL: Philox4x32-10
 *            normals keyed by (seed, sample0 + b, element, index) -- identical for any sharding of the batch
 *            over ranks; orig / mask: RePaint known region or NULL.
 * flags bit 0: skip the cond-only prologue (cross-attention vectors) -- pf_unet_prepare_cond was run for this
 * cond tensor.  DDPM draws no noise at index 0 (sampler_sdf.py:152-153, 322-324). */
typedef struct pf_fused_step {
  int32_t kind;  /* 1 = DDPM (pf_sample_step_ddpm arithmetic), 2 = DDIM (pf_sample_step_ddim) */
  int32_t flags;
  int32_t* index;
  const float* coef;
  const int64_t* t_table;
  float* x;
  float* eps_out; /* optional */
  const float* noise;
  const float* noise_kn;
  const float* orig;
  const float* mask;
  float temperature;
  uint64_t seed;
  int64_t sample0;
} pf_fused_step;
int pf_unet_forward_step(pf_unet* h, const float* x_in, int64_t* time_steps, const float* cond, int32_t batch,
                         int32_t n_cond, int32_t height, int32_t width, const pf_fused_step* step,
                         void* workspace, size_t workspace_bytes, pf_stream stream);
/* Runs only the part of the plan that depends on `cond` alone (the n_cond == 1 cross-attention vectors
 * to_out(to_v(cond)), unet_attention.py:186-212): once per sampling loop instead of once per step. */
int pf_unet_prepare_cond(pf_unet* h, const float* cond, int32_t batch, int32_t n_cond, int32_t height,
                         int32_t width, void* workspace, size_t workspace_bytes, pf_stream stream);
/* Tabulates time_embed + every ResBlock emb_layers projection (unet.py:64-68, 151-169, 286-289) for the
 * integer timesteps 0 .. n_steps-1 with the plan's own kernels (rows bit-identical to a forward at that t);
 * later forwards gather one row per sample (timesteps are clamped to the table).  Reset by pf_unet_finalize. */
int pf_unet_enable_time_lut(pf_unet* h, int32_t n_steps, pf_stream stream);
/* out[b, i] = the Philox normal pf_unet_forward_step would draw for (seed, sample0 + b, element i, index,
 * which: 0 = step noise, 1 = known-region noise); out [n_samples, per_sample] fp32. */
int pf_fill_normal(float* out, int64_t n_samples, int64_t per_sample, uint64_t seed, int64_t sample0,
                   int32_t index, int32_t which, pf_stream stream);

/* Same as pf_unet_forward but brackets every kernel of the plan with CUDA events on `stream`,
 * synchronises, and reports per-launch milliseconds, algorithmic FLOPs (2*M*N*K for the tcgen05
 * GEMM launches, 0 otherwise) and the op kind (0 = tcgen05 GEMM, 1 = first conv, 2/3 = GroupNorm
 * statistics / finalize, 4 = operand transform, 5 = LayerNorm, 6 = GeGLU, 7 = softmax, 8 = timestep
 * sinusoid, 9 = small linear, 10 = final conv, 11 = memset) into host arrays.  Measurement aid for
 * bench.py's roofline; not used on the sampling path. */
int pf_unet_forward_profiled(pf_unet* h, const float* x, const int64_t* time_steps,
                             const float* cond, int32_t batch, int32_t n_cond, int32_t height,
                             int32_t width, float* out, void* workspace, size_t workspace_bytes,
                             pf_stream stream, float* op_ms_host, double* op_flops_host,
                             int32_t* op_kind_host, int32_t max_ops, int32_t* n_ops);
/* human-readable description of launch i of the last-used plan (measurement aid) */
int pf_unet_op_desc(pf_unet* h, int32_t i, char* buf, int32_t len);
/* number of kernel launches one pf_unet_forward issues for the last-used plan */
int32_t pf_unet_launch_count(pf_unet* h);

/* Sampler step epilogues (elementwise over n = B*C*H*W floats).  Optional pointers may be NULL.
 * e_uncond != NULL selects classifier-free guidance e = e_u + s*(e_c - e_u)
 * (stable_diffusion/sampler/__init__.py:69-77); orig != NULL selects the RePaint blend
 * x = (kn_a*orig + kn_b*noise_kn)*mask + x_prev*(1-mask) (sampler_sdf.py:322-336,
 * sampler_ddim.py:355-359). */
typedef struct pf_step_args {
  const float* x;
  const float* e_cond;
  const float* e_uncond;
  const float* noise;
  const float* orig;
  const float* mask;
  const float* noise_kn;
  float* x_prev;
  float* x0;
  float* e_t;
  int64_t n;
  int64_t noise_bcast; /* >0: noise holds this many elements and repeats over the batch */
  float uncond_scale;
  float c0, c1, c2, c3, c4; /* per-step coefficients, see below */
  float temperature;
  float kn_a, kn_b;
} pf_step_args;

/* SDFSampler.p_sample -- sampler_sdf.py:121-171.  c0 = sqrt_recip_alpha_bar[step],
 * c1 = sqrt_recip_m1_alpha_bar[step], c2 = mean_x0_coef[step], c3 = mean_xt_coef[step],
 * c4 = exp(0.5*log_var[step]). */
int pf_sample_step_ddpm(const pf_step_args* a, pf_stream stream);
/* DDIMSampler.get_x_prev_and_pred_x0 -- sampler_ddim.py:233-272.  c0 = ddim_sqrt_one_minus_alpha[i],
 * c1 = ddim_alpha[i]**0.5, c2 = ddim_alpha_prev[i]**0.5, c3 = sqrt(1-alpha_prev-sigma^2),
 * c4 = ddim_sigma[i]. */
int pf_sample_step_ddim(const pf_step_args* a, pf_stream stream);
/* DenoiseDiffusion.p_sample -- ddpm/__init__.py:66-88.  c0 = (1-alpha_t)/sqrt(1-alpha_bar_t),
 * c1 = 1/sqrt(alpha_t), c2 = sqrt(sigma2_t). */
int pf_sample_step_ddpm_legacy(const pf_step_args* a, pf_stream stream);
/* q_sample: out = a*x0 + b*noise -- sampler_sdf.py:192, sampler_ddim.py:296-299 */
int pf_q_sample(const float* x0, const float* noise, float* out, int64_t n, float a, float b,
                pf_stream stream);

/* get_mask(orig, "below" | "above") -- inference_sdf.py:132-180, batched over songs.
 * orig / mask [n_seg, channels, steps, pitches] fp32 (channel 0 = onsets); every seg_per_song
 * consecutive segments form one song whose rows are scanned as one sequence (the reference scans the
 * whole batch as one sequence: seg_per_song = n_seg).  above = 0: keep pitch >= lowest onset
 * (inpaint accompaniment below a melody); above = 1: keep pitch <= highest onset.  Synchronises;
 * fails if a song has no onset (the reference raises IndexError). */
int pf_get_mask(const float* orig, float* mask, int32_t n_seg, int32_t seg_per_song, int32_t channels,
                int32_t steps, int32_t pitches, int32_t above, pf_stream stream);

/* Condition encoders (the step before the sampling loop).  fp32,
 * device pointers, asynchronous on `stream`, no allocation.
 * pf_linear: out[r, :n_out] = act(in[r, :n_in] . weight[n_out, n_in]^T + bias)   (torch.nn.Linear;
 *   act 0 = none, 1 = SiLU, 2 = exp as in `linear_var(x).exp_()`, dl_modules/chord_enc.py:20).
 * pf_gru_bidir_last: final hidden states of a 1-layer bidirectional batch_first torch.nn.GRU
 *   (`self.gru(x)[-1]` transposed to [B, 2H], forward half first: chord_enc.py:15-17,
 *   txt_enc.py:29-31).  x [B, T, n_in]; w_ih[d] [3H, n_in], w_hh[d] [3H, H], b_ih[d], b_hh[d] [3H]
 *   for d = 0 (forward), 1 (reverse), gate order r, z, n; h_last [B, 2H]. */
int pf_linear(const float* in, int64_t ld_in, const float* weight, const float* bias, float* out,
              int64_t ld_out, int32_t rows, int32_t n_out, int32_t n_in, int32_t act, pf_stream stream);
size_t pf_gru_workspace_bytes(int32_t batch, int32_t steps, int32_t hidden);
int pf_gru_bidir_last(const float* x, int32_t batch, int32_t steps, int32_t n_in, int32_t hidden,
                      const float* const* w_ih, const float* const* w_hh, const float* const* b_ih,
                      const float* const* b_hh, float* h_last, void* workspace, size_t workspace_bytes,
                      pf_stream stream);
/* TextureEncoder.cnn (txt_enc.py:10-14): Conv2d(1, channels, (4,12), stride (4,1)) -> ReLU ->
 * MaxPool2d((1,4)) of pr [B, steps, pitches] -> out [B, channels, steps/4, (pitches-11)/4]. */
int pf_txt_cnn(const float* pr, const float* weight, const float* bias, float* out, int32_t batch,
               int32_t channels, int32_t steps, int32_t pitches, pf_stream stream);

/* Piano-roll decode (replaces the Python loops of utils.py:240-269
 * prmat2c_to_prmat and the note loop of utils.py:446-470 prmat2c_to_midi_file).
 * prmat2c [n_seg, channels >= 2, steps, pitches] fp32 (channel 0 onset, channel 1 sustain), device.
 * prmat [n_seg * steps, pitches] int64, device: duration of the note starting at (step, pitch), 0
 * where int(round(onset)) <= 0 -- the same bytes as the reference's (n_seg*ratio, n_step, 128)
 * array.  Asynchronous on `stream`. */
int pf_prmat2c_to_prmat(const float* prmat2c, int32_t n_seg, int32_t channels, int32_t steps,
                        int32_t pitches, int64_t* prmat, pf_stream stream);
/* Notes of a duration matrix in the reference's loop order (segment, step, pitch):
 * row_offsets [rows + 1] int32 device scratch/out (exclusive offsets, total in the last slot),
 * notes [cap][3] int32 device (row = seg * steps + step, pitch, duration; may be NULL with cap 0 to
 * only count), *n_notes (host) receives the total.  Synchronises the stream. */
int pf_prmat_notes(const int64_t* prmat, int64_t rows, int32_t pitches, int32_t* row_offsets,
                   int32_t* notes, int64_t cap, int64_t* n_notes, pf_stream stream);

/* Building-block ops (used by the parity tests; they allocate their own scratch and synchronise).
 * conv: x NHWC fp32 [B,H,W,Cin] (Cin%64==0), w [Cout,Cin,k,k] (k in {1,3}), stride in {1,2},
 * upsample in {0,1} (nearest 2x before the conv), bias/resid optional; out NHWC [B,Ho,Wo,Cout]. */
int pf_op_conv2d_nhwc(const float* x, int32_t B, int32_t H, int32_t W, int32_t Cin, const float* w,
                      int32_t Cout, int32_t ksize, int32_t stride, int32_t upsample,
                      const float* bias, const float* resid, float* out, int32_t force_bn,
                      pf_stream stream);
/* Same, with a per-sample epilogue vector: bias [B, bias_ld] when bias_ld > 0 (e.g. conv bias +
 * time-embedding projection, ddpm/unet.py:140-141), a shared bias [Cout] when bias_ld == 0. */
int pf_op_conv2d_nhwc_ex(const float* x, int32_t B, int32_t H, int32_t W, int32_t Cin, const float* w,
                         int32_t Cout, int32_t ksize, int32_t stride, int32_t upsample,
                         const float* bias, int64_t bias_ld, const float* resid, float* out,
                         int32_t force_bn, pf_stream stream);
/* Generic blocks of the legacy unconditional UNet (ddpm/unet.py:410-444, SURVEY.md section 8a row D2;
 * composed by polyffusion_b200/ddpm/unet.py).  All asynchronous on `stream`.
 * groupnorm: GroupNorm(groups, eps) [+ Swish] of an NHWC fp32 tensor, any C % groups == 0
 *            (ResidualBlock norms with 32 groups up to C = 2048, final GroupNorm(8, 64), ddpm/unet.py:404);
 * softmax_rows: out[r, :] = softmax(scale * s[r, :]) (AttentionBlock, ddpm/unet.py:199-201);
 * time_sincos: [sin(t f_i) | cos(t f_i)] with t cast to fp32 (TimeEmbedding, ddpm/unet.py:62-72). */
int pf_op_groupnorm_generic(const float* x, int32_t B, int32_t HW, int32_t C, int32_t groups,
                            const float* gamma, const float* beta, float eps, int32_t silu, float* out,
                            pf_stream stream);
int pf_op_softmax_rows(const float* s, float scale, float* out, int64_t rows, int32_t n, pf_stream stream);
int pf_op_time_sincos(const int64_t* t, const float* freqs, float* out, int32_t B, int32_t half,
                      pf_stream stream);
/* direct fp32 3x3 convolution, pad 1, w [Cout,Cin,3,3]: the legacy UNet's edge layers (image_proj from
 * NCHW, final to NCHW; ddpm/unet.py:345-347, 405-407); x / out are NCHW or NHWC per the two flags. */
int pf_op_conv3x3_direct(const float* x, const float* w, const float* bias, float* out, int32_t B,
                         int32_t Cin, int32_t H, int32_t W, int32_t Cout, int32_t in_nchw, int32_t out_nchw,
                         pf_stream stream);
/* softmax(q k^T / sqrt(d_head)) v per head: q [B,N,heads*64], k/v [B,Nk,heads*64] fp32 -> out
 * [B,N,heads*64] fp32 (unet_attention.py:261-293 normal_attention, before to_out). */
int pf_op_attention(const float* q, const float* k, const float* v, int32_t B, int32_t N,
                    int32_t Nk, int32_t heads, float* out, pf_stream stream);
/* GroupNorm(32 groups, eps) [+ SiLU] of an NHWC tensor -> fp32 NHWC (hi+lo of the operand). */
int pf_op_groupnorm_nhwc(const float* x, int32_t B, int32_t HW, int32_t C, const float* gamma,
                         const float* beta, float eps, int32_t silu, float* out, pf_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* PF_B200_H */
