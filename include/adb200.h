/*
 * adb200 — C ABI of the B200-native EDM sampling / denoising hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no framework types. Every pointer named
 * `*_dev` / documented "device" is a CUDA device pointer on the current device; `stream` is a
 * `cudaStream_t` passed as `void*` (NULL = default stream). All functions enqueue asynchronously on
 * `stream`, never synchronise unless stated, and return 0 on success or a non-zero code, in which
 * case `adb_last_error()` holds a message. There is no CPU fallback: on a device that is not
 * sm_100 every compute entry point fails.
 *
 * Each entry point names the reference code (AgentCooper2002/AudioDiffuser, paths relative to the
 * reference root) whose arithmetic it replaces. INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 */
#ifndef ADB200_H
#define ADB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADB_OK 0
#define ADB_ERR_INVALID 1      /* bad argument (shape, alignment, NULL) */
#define ADB_ERR_CUDA 2         /* CUDA runtime / driver error */
#define ADB_ERR_UNSUPPORTED 3  /* option outside the fused path (caller must not silently fall back) */
#define ADB_ERR_PIPELINE 4     /* an in-kernel barrier wait timed out (programming error, reported by adb_check_async) */

#define ADB_PRECISION_FP32 0   /* CUDA-core fp32 path: <= 1e-5 relative to the reference */
#define ADB_PRECISION_BF16 1   /* tcgen05 path: bf16 operands, fp32 accumulate: <= 2e-2 relative */

const char* adb_last_error(void);
int adb_version(void);
/* 0 if `device` is sm_100 (B200) and usable. */
int adb_device_check(int device);
/* Host view of the z-stash residual-block kernel's job order (types 0 = G1a, 1 = G1b, 2 = G2r; wavenet_tc3.cuh): fills up to `cap`
 * (type, group) pairs and returns the number of jobs. No GPU involved; used by the CPU tests of the software-pipelined schedule. */
int adb_debug_zs_job_order(int n_groups, int write_h, int pipelined, int* types, int* groups, int cap);
/* Host view of the opt-in multi-layer wavefront launch (ADB_ZS_ML = S): walks every (sub-pass, layer, tile group) item of a chunk of
 * `bc` samples in launch order and returns how many awaited tiles do NOT belong to an earlier item of the same sub-pass (0 = the
 * round-robin dealing cannot deadlock); *min_distance = smallest index distance item -> dependency, *pipelined_ok = whether the
 * launch would use the software-pipelined job order. Test infrastructure (tests/test_host_logic.py), no GPU needed. */
int adb_debug_ml_order(int bc, int L, int layers, int cycle, int S, int pairs, long long* min_distance, int* n_items,
                       int* pipelined_ok);

/* Synchronise the device and report any asynchronous kernel / pipeline error (test & bench use). */
int adb_check_async(void);
/* Kernel launches issued so far by the op-level entry points (adb_edm_*, adb_cl_*, training step) on this process;
 * reset != 0 zeroes the counter. The DiffWave handle's own launches are reported by adb_wavenet_timers. */
long long adb_launch_count(int reset);

/* ------------------------------------------------------------------------------------------------
 * EDM preconditioning — src/models/components/diffusion.py:232-241 (get_scale_weights) and :46-63
 * (denoise_fn). State tensors are fp32, contiguous, `n_per` elements per sample, `B` samples.
 * `sigmas_dev` holds 1 value (sigma_stride = 0, the sampler's scalar sigma, utils.py:41-52) or B
 * values (sigma_stride = 1, training).
 * ---------------------------------------------------------------------------------------------- */
/* c_in[b] = (sigma_b^2 + sigma_data^2)^-1/2 and c_noise[b] = 0.25 ln(sigma_b) alone (diffusion.py:235, :240), for callers that fold
 * the input scale into their own first kernel (adb_cl_wavenc_prep) */
int adb_edm_precond_coef(const float* sigmas_dev, int sigma_stride, float sigma_data, float* c_in_dev, float* c_noise_dev, int B,
                         void* stream);
/* net_in = c_in(sigma) * x ; c_noise[b] = 0.25 ln(sigma_b)   (diffusion.py:50, :235) */
int adb_edm_precond_in(const float* x_dev, const float* sigmas_dev, int sigma_stride, float sigma_data,
                       float* net_in_dev, float* c_noise_dev, int B, int64_t n_per, void* stream);
/* out = clamp(c_skip x + c_out F, -1, 1)   (diffusion.py:60-63; dynamic_threshold == 0 only).
 * If f_null_dev != NULL: F = f_null + (f - f_null) * cond_scale first (diffusion.py:52-54). */
int adb_edm_precond_out(const float* x_dev, const float* f_dev, const float* f_null_dev, float cond_scale,
                        const float* sigmas_dev, int sigma_stride, float sigma_data, float* out_dev, int B,
                        int64_t n_per, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Sampler updates on a denoised estimate D (generic path: any `fn`) —
 * src/models/components/sampler_edm.py:343-367 (EDMSampler.step), :259-280 (EDMAlphaSampler.step).
 * ---------------------------------------------------------------------------------------------- */
/* out = a * x            (sampler_edm.py:380  x = sigmas[0] * noise) */
int adb_edm_scale(const float* x_dev, float a, float* out_dev, int64_t n, void* stream);
/* out = x + a * e        (churn: sampler_edm.py:347 with a = sqrt(sigma_hat^2 - sigma^2) * s_noise; also :273, :280) */
int adb_edm_axpy(const float* x_dev, const float* e_dev, float a, float* out_dev, int64_t n, void* stream);
/* out[b][i] = x[b][i] + a * (s_noise * eps), eps ~ N(0,1) drawn in the kernel: `epsilon = randn_like(x)` + the churn update of
 * sampler_edm.py:346-347 in one pass. Philox4x32-10, key = seed, counter = (i / 4, sample0 + b, step, high word of the sample
 * index), Box-Muller on 24-bit uniforms (oracle/philox.py restates it). Deterministic in (seed, global sample index, step,
 * element): independent of batch composition, rank and world size. In place (out == x) is allowed. */
int adb_edm_churn_rng(const float* x_dev, float* out_dev, float a, float s_noise, uint64_t seed, int step, int64_t sample0, int B,
                      int64_t n_per, void* stream);
/* d = (x - D) / sigma ; x_next = x + h d          (sampler_edm.py:354-357) */
int adb_edm_euler(const float* x_dev, const float* denoised_dev, float sigma, float h, float* d_dev,
                  float* x_next_dev, int64_t n, void* stream);
/* d2 = (x1 - D1) / sigma1 ; out = x + h (w0 d + w1 d2)   (Heun: w0 = w1 = 0.5, sampler_edm.py:366-367;
 * general RK2: w0 = 1 - 1/(2 alpha), w1 = 1/(2 alpha), sampler_edm.py:277-278) */
int adb_edm_rk2(const float* x_dev, const float* d_dev, const float* x1_dev, const float* denoised1_dev,
                float sigma1, float h, float w0, float w1, float* out_dev, int64_t n, void* stream);

/* out = a x - e d                       (d_old_dev == NULL: DPM-Solver++(2M) first-order step, sampler_edm.py:1097-1098)
 * out = a x - e (c0 d - c1 d_old)       (second-order multistep, sampler_edm.py:1100-1108) */
int adb_edm_lincomb(const float* x_dev, const float* d_dev, const float* d_old_dev, float a, float e, float c0, float c1,
                    float* out_dev, int64_t n, void* stream);
/* out = a x + sum_{i<k} coefs[i] terms[i]   (k <= 4; clamped to [-1, 1] when `clamp` != 0). `terms` is a HOST array of k
 * device pointers, `coefs` a host array of k scalars. One launch per update of the reference's DPMSampler (single-step
 * DPM-Solver-1/2/3 and multistep order 1-3, sampler_edm.py:562-704) and UniPCSampler (:870-987): every such update is a
 * combination of the state and the stored network outputs with scalars that depend only on the time grid. */
int adb_edm_lincomb_n(const float* x_dev, float a, const float* const* terms, const float* coefs, int k, int clamp,
                      float* out_dev, int64_t n, void* stream);
/* out = clamp(x, -1, 1)                 (the final x.clamp of DPM2MSampler.forward, sampler_edm.py:1131) */
int adb_edm_clamp(const float* x_dev, float* out_dev, int64_t n, void* stream);
/* ema = torch.lerp(ema, params, weight) on flat vectors, in place — PowerFunctionEMA.update / TraditionalEMA.update
 * (src/models/phema.py:104-108, :145-151), one launch per EMA instead of one per parameter tensor */
int adb_ema_lerp(float* ema_dev, const float* params_dev, float weight, int64_t n, void* stream);
/* pcm[i] = saturate_int16(round_half_even(x[i] * 32768)) — the float -> 16-bit PCM conversion behind
 * torchaudio.save(..., bits_per_sample=16) of the generated test samples (src/models/diffunet_complex_module.py:263-266),
 * done on the device so only 2 bytes per sample cross PCIe. NaN -> 0. 6 bytes of HBM traffic per sample. */
int adb_pcm16_encode(const float* x_dev, int16_t* pcm_dev, int64_t n, void* stream);

/* Fused Heun step around RAW network outputs F (what the fused trajectory launches between network
 * evaluations; sampler_edm.py:350-367 with diffusion.py:60-63 inlined):
 *   mid : D1 = clamp(c_skip(s) x + c_out(s) F1) ; d = (x - D1)/s ; x1 = x + h d        reads 8 B, writes 8 B / element
 *   post: x1 = x + h d ; D2 = clamp(c_skip(s1) x1 + c_out(s1) F2) ; d2 = (x1 - D2)/s1 ;
 *         x_next = x + h/2 (d + d2)                                                     reads 12 B, writes 4 B / element */
int adb_edm_heun_mid(const float* x_dev, const float* f1_dev, float sigma, float sigma_data, float h, float* d_dev,
                     float* x1_dev, int64_t n, void* stream);
int adb_edm_heun_post(const float* x_dev, const float* d_dev, const float* f2_dev, float sigma1, float sigma_data, float h,
                      float* x_next_dev, int64_t n, void* stream);
/* Final / Euler-only step on the RAW network output: D = clamp(c_skip x + c_out F) ; x_next = x + h (x - D) / sigma
 * (sampler_edm.py:354-357 with diffusion.py:46-63 folded in; 12 B per element; x_next may alias x). */
int adb_edm_euler_raw(const float* x_dev, const float* f_dev, float sigma, float sigma_data, float h, float* x_next_dev, int64_t n,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training twin — src/models/components/diffusion.py:65-97 (Diffusion.forward).
 * ---------------------------------------------------------------------------------------------- */
/* x_noisy = x + sigma_b * noise ; net_in = c_in(sigma_b) * x_noisy ; c_noise[b]   (diffusion.py:79, :57) */
int adb_edm_noise_in(const float* x_dev, const float* noise_dev, const float* sigmas_dev, float sigma_data,
                     float* x_noisy_dev, float* net_in_dev, float* c_noise_dev, int B, int64_t n_per, void* stream);
/* loss[b] = lambda(sigma_b) * mean_i (clamp(c_skip x_noisy + c_out F) - x)^2   (diffusion.py:60-63, :92-95) */
int adb_edm_dsm_loss(const float* x_dev, const float* x_noisy_dev, const float* f_dev, const float* sigmas_dev,
                     float sigma_data, float* loss_dev, int B, int64_t n_per, void* stream);
/* The same with the reference's `x_mask` (diffusion.py:80-83): mask_dev holds one byte per element of x (already broadcast to
 * x's shape); elements whose byte is 0 count with weight 0.01. mask_dev == NULL: no mask. */
int adb_edm_dsm_loss_masked(const float* x_dev, const float* x_noisy_dev, const float* f_dev, const float* sigmas_dev,
                            float sigma_data, const unsigned char* mask_dev, float* loss_dev, int B, int64_t n_per, void* stream);
/* d_f[b][i] = upstream[b] * d loss[b] / d F[b][i] of the loss above (torch.clamp passes the gradient on the closed interval): lets
 * PyTorch autograd carry the DSM loss into ANY differentiable `net` (Diffusion.forward trains whatever backbone it is given,
 * diffusion.py:65-97); the fused DiffWave training step has its own backward (adb_wavenet_dsm_backward). */
int adb_edm_dsm_loss_grad(const float* x_dev, const float* x_noisy_dev, const float* f_dev, const float* sigmas_dev,
                          float sigma_data, const unsigned char* mask_dev, const float* upstream_dev, float* d_f_dev, int B,
                          int64_t n_per, void* stream);

/* ------------------------------------------------------------------------------------------------
 * DiffWave backbone — src/models/backbones/wavenet.py:153-180 (WaveNetNoise), :117-151, :94-115.
 * ---------------------------------------------------------------------------------------------- */
typedef struct adb_wavenet adb_wavenet;

/* Number of fp32 values of the flat parameter vector for a given configuration: the concatenation
 * of WaveNetNoise(...).state_dict() values in state_dict order (wavenet.py:153-168). */
int64_t adb_wavenet_param_count(int residual_channels, int residual_layers);
/* Create from the flat parameter vector (device or host pointer; `on_device` says which). Folds the
 * weight norm (wavenet.py:44-51), packs weights for both precisions. Synchronises. */
int adb_wavenet_create(adb_wavenet** out, int residual_channels, int residual_layers, int dilation_cycle,
                       const float* params, int64_t n_params, int on_device);
/* Replace the parameters of an existing handle (same configuration) and rebuild the packed weights without
 * re-allocating: what a training loop calls after each optimizer step. Enqueues on the legacy default stream. */
int adb_wavenet_load_params(adb_wavenet* net, const float* params, int64_t n_params, int on_device);
void adb_wavenet_destroy(adb_wavenet* net);
/* Bytes of scratch the calls below need for (B, L, precision); the caller owns the buffer. */
int64_t adb_wavenet_workspace_bytes(const adb_wavenet* net, int B, int L, int precision);

/* F = WaveNetNoise(in_scale_b * x, c_noise)   (wavenet.py:170-180): x [B][L] -> out [B][L] (the
 * reference's [B,1,L]). `in_scale_dev` may be NULL (scale 1); stride 0 = one value for the batch. */
int adb_wavenet_forward(adb_wavenet* net, const float* x_dev, const float* c_noise_dev, const float* in_scale_dev,
                        int in_scale_stride, float* out_dev, int B, int L, int precision, void* workspace_dev,
                        int64_t workspace_bytes, void* stream);
/* Same, additionally copying h and the running skip sum after each of the first `dump_layers`
 * residual blocks into fp32 [dump_layers][B][L][C] buffers (debug / unit tests). */
int adb_wavenet_forward_debug(adb_wavenet* net, const float* x_dev, const float* c_noise_dev,
                              const float* in_scale_dev, int in_scale_stride, float* out_dev, int B, int L,
                              int precision, void* workspace_dev, int64_t workspace_bytes, float* dump_h_dev,
                              float* dump_skip_dev, int dump_layers, void* stream);
/* x0_hat = EluDiffusion.denoise_fn(x, net, sigma | sigmas)  fused with the backbone
 * (diffusion.py:32-63 + wavenet.py:170-180). */
int adb_wavenet_denoise(adb_wavenet* net, const float* x_dev, const float* sigmas_dev, int sigma_stride,
                        float sigma_data, float* out_dev, int B, int L, int precision, void* workspace_dev,
                        int64_t workspace_bytes, void* stream);

/* Whole EDM sampling trajectory on the device — EDMSampler.forward (sampler_edm.py:371-397) when
 * alpha < 0, EDMAlphaSampler.forward (sampler_edm.py:284-300) when alpha > 0.
 *   sigmas_host [n_sigmas]  the schedule (fp32, host), e.g. KarrasSchedule (scheduler.py:17-22)
 *   eps_dev                 churn noise, [num_steps][B][L] N(0,1) fp32, or NULL when s_churn == 0
 *                           (EDMSampler draws it every step, sampler_edm.py:346; only steps with
 *                            gamma > 0 read it)
 * Writes x [B][L] and, if nfe_out != NULL, the number of network evaluations. */
int adb_wavenet_sample_edm(adb_wavenet* net, const float* noise_dev, const float* sigmas_host, int n_sigmas,
                           int num_steps, float sigma_data, float s_tmin, float s_tmax, float s_churn, float s_noise,
                           int use_heun, float alpha, const float* eps_dev, float* x_out_dev, int B, int L,
                           int precision, void* workspace_dev, int64_t workspace_bytes, int* nfe_out, void* stream);

/* The same with the churn noise drawn in the kernel when eps_dev == NULL: Philox4x32-10 keyed by `churn_seed`, counter =
 * (element group, sample0 + b, step) — `sample0` is the global index of row 0 (the rank's shard offset), so a waveform does
 * not depend on the batch or world size it is sampled in. Replaces `torch.randn_like(x)` per step (sampler_edm.py:346)
 * without a [num_steps][B][L] tensor. adb_wavenet_sample_edm == this with churn_seed = 0, sample0 = 0. */
int adb_wavenet_sample_edm_seeded(adb_wavenet* net, const float* noise_dev, const float* sigmas_host, int n_sigmas,
                                  int num_steps, float sigma_data, float s_tmin, float s_tmax, float s_churn, float s_noise,
                                  int use_heun, float alpha, const float* eps_dev, uint64_t churn_seed, int64_t sample0,
                                  float* x_out_dev, int B, int L, int precision, void* workspace_dev,
                                  int64_t workspace_bytes, int* nfe_out, void* stream);

/* Per-kernel-class device time of the last *_timed call below (microseconds, CUDA events). */
#define ADB_TIMER_CONV 0   /* residual-block kernels (the dominant kernel) */
#define ADB_TIMER_STEP 1   /* fused sampler-step kernels */
#define ADB_TIMER_AUX 2    /* embedding MLP / E table / input projection */
#define ADB_TIMER_TAIL 3   /* skip-projection + output-projection kernel(s) */
#define ADB_TIMER_SKIP 4   /* skip GEMM over the stashed gated activations (z-stash path) */
#define ADB_TIMER_COUNT 5
/* Enable (1) / disable (0) event timing around kernel classes inside adb_wavenet_sample_edm and
 * adb_wavenet_forward; adb_wavenet_timers() synchronises and returns accumulated ms and kernel-launch counts per class
 * (launch counts are maintained whether or not timing is enabled). */
int adb_wavenet_set_timing(adb_wavenet* net, int enabled);
int adb_wavenet_timers(adb_wavenet* net, double* ms_out /*[ADB_TIMER_COUNT]*/, int64_t* launches_out /*[ADB_TIMER_COUNT]*/);

/* ------------------------------------------------------------------------------------------------
 * 1-D U-Net building blocks — src/models/backbones/unet1d.py, src/models/backbones/attention_utils.py.
 * Activations are channels-last [B][L][C] in `dtype` (ADB_DTYPE_F32: CUDA-core fp32 path, <= 1e-5 relative;
 * ADB_DTYPE_BF16: tcgen05 path with bf16 activations / weights, fp32 accumulate, statistics and softmax in fp32).
 * The host side (audiodiffuser_b200/backbones/unet1d.py) sequences these exactly like UNet1d.forward
 * (unet1d.py:769-816).
 * ---------------------------------------------------------------------------------------------- */
#define ADB_DTYPE_F32 0
#define ADB_DTYPE_BF16 1
#define ADB_ACT_NONE 0
#define ADB_ACT_RELU 1
#define ADB_ACT_SILU 2
#define ADB_ACT_GELU 3   /* erf form, nn.GELU() default (unet1d.py:55) */

/* Generic channels-last GEMM-convolution (nn.Conv1d / nn.ConvTranspose1d / bias-free nn.Linear call sites:
 * unet1d.py:154-158, :186-193, :214-255, :291-295, :49-61; attention_utils.py:95-110):
 *   Y[b][t][n] = act(bias[n] + sum_{j<taps} sum_{ci} X[b][t + off0 + j*dil][ci] * W[j][ci][n]) (+ res[b][t][n]),
 *   t in [0, rows), rows of X outside [0, L_in) read as zero.
 * ups == 0: out is [B][rows][N]. ups = f > 0 (ConvTranspose k = 2f, s = f): N = f*Cout and (t, n) is stored to
 * output row t*f + n/Cout - shift (kept if in [0, L_out)), channel n % Cout of out [B][L_out][Cout].
 * w: ADB_DTYPE_F32 -> fp32 [taps][Cin][N]; ADB_DTYPE_BF16 -> blocks written by adb_cl_pack_conv_weights. */
int adb_cl_conv(const void* in_dev, const void* w_dev, const float* bias_dev, const void* res_dev, void* out_dev, int B,
                int L_in, int rows, int Cin, int N, int taps, int off0, int dil, int act, int ups, int shift, int L_out,
                int dtype, void* stream);
int64_t adb_cl_conv_packed_elems(int Cin, int N, int taps);
/* adb_cl_conv (plain store) that runs only the first `last_tap_blocks` 64-channel K-blocks of the LAST tap: for packed weights
 * whose remaining rows of that tap are all zero — Downsample1d on the [L/f][f*C] view (unet1d.py:214-225), whose third coarse tap
 * holds a single fine tap (k = 2f + 1), so 3 of its 4 (f = 4) phase blocks would multiply zeros. fp32 ignores the hint. */
int adb_cl_conv_ktrim(const void* in_dev, const void* w_dev, const float* bias_dev, const void* res_dev, void* out_dev, int B, int L_in,
                      int rows, int Cin, int N, int taps, int off0, int dil, int act, int last_tap_blocks, int dtype, void* stream);
int adb_cl_pack_conv_weights(const float* w_f32_dev, void* packed_bf16_dev, int Cin, int N, int taps, void* stream);
/* the same blocks in fp16: the weight operand of adb_cl_gn_conv3, whose activation operand is produced in fp16 by the fused
 * GroupNorm / SiLU transform (11 significand bits instead of bf16's 8, and a SiLU that costs one packed MUFU per two elements) */
int adb_cl_pack_conv_weights_f16(const float* w_f32_dev, void* packed_f16_dev, int Cin, int N, int taps, void* stream);
/* Weight gradient of the same convolution (what autograd computes for nn.Conv1d weights): out[tap][i][j] (fp32
 * [taps][Ca][Cg], accumulated into) += scale * sum_{b,t} A[b][t + (tap - taps/2)*dil][i] * G[b][t][j], rows of A
 * outside [0, L) read as zero. ADB_DTYPE_BF16: tcgen05 with MN-major operands straight from the channels-last tensors. */
int adb_cl_wgrad(const void* a_dev, const void* g_dev, float* out_dev, int B, int L, int Ca, int Cg, int taps, int dil,
                 float scale, int dtype, void* stream);
/* out[b][n] = act(bias[n] + sum_k W[n][k] * f(in[b][k])), f = SiLU if silu_in; fp32 (time MLP unet1d.py:678-684,
 * to_cond_embedding :271-276, :304-308) */
int adb_cl_linear(const float* in_dev, const float* w_dev, const float* bias_dev, float* out_dev, int B, int K, int N,
                  int silu_in, int act, void* stream);
/* LabelEmbedder lookup with classifier-free-guidance dropout (conditioner.py:94-106): out[b] = drop[b] ? null_row :
 * table[labels[b]] (labels int64 [B]; drop int32 [B] or NULL = keep all); fp32 [B][C] */
/* to_bf16 != 0: out_bf16[i] = bf16(silu ? SiLU(in_f32[i]) : in_f32[i]); to_bf16 == 0: out_f32[i] = float(in_bf16[i]). The
 * operand / result conversions around the tensor-core form of the per-block conditioning projections
 * (`to_cond_embedding` = SiLU -> Linear, unet1d.py:279-283, :304-310), which in bf16 mode run as ONE adb_cl_conv over
 * [1][B][T] instead of the fp32 adb_cl_linear. */
int adb_cl_cast(const void* in_dev, void* out_dev, int64_t n, int to_bf16, int silu, void* stream);
int adb_cl_label_embed(const float* table_dev, const float* null_row_dev, const long long* labels_dev, const int* drop_dev,
                       float* out_dev, int B, int C, int num_classes, void* stream);
/* [t, sin(2 pi t w), cos(2 pi t w)] -> out [B][2*half+1]   (LearnedPositionalEmbedding, unet1d.py:128-142) */
int adb_cl_time_features(const float* t_dev, const float* w_dev, float* out_dev, int B, int half, void* stream);
/* nn.GroupNorm(G, C) + optional x*(scale+1)+shift (scale_shift_dev [B][ss_ld]: scale at [0,C), shift at [C,2C)) +
 * activation   (ConvBlock1d.forward, unet1d.py:195-207). sums_ws_dev: [B][G][2] doubles of scratch. */
int adb_cl_groupnorm(const void* in_dev, const float* gamma_dev, const float* beta_dev, const float* scale_shift_dev,
                     int64_t ss_ld, void* out_dev, double* sums_ws_dev, int B, int L, int C, int G, float eps, int act,
                     int dtype, void* stream);
/* GroupNorm statistics only: fp64 (sum, sum of squares) of the G groups of in_dev [B][L][C] into
 * sums_dev[b][g_off + g][2] of a [B][g_total][2] array (g_off / g_total place the groups of one half of a channel concatenation
 * next to the other half's, unet1d.py:552-556). zero_first != 0 clears the whole array first. */
int adb_cl_gn_stats(const void* in_dev, double* sums_dev, int B, int L, int C, int G, int g_total, int g_off, int zero_first,
                    int dtype, void* stream);
/* First half of the fused ConvBlock1d (bf16 only): GroupNorm statistics of in_dev [B][L][C] (G groups -> sums_dev[b][g_off + g] of a
 * [B][g_total][2] fp64 array) and, by the block that finishes a sample last, the per-(sample, channel) affine coefficients of
 * GroupNorm * gamma + beta followed by x * (scale + 1) + shift (unet1d.py:160-161, :195-207) into coef_dev [2][B][Cin_total]
 * (slopes, then offsets; channel c of this input is channel c_off + c of the possibly concatenated tensor, unet1d.py:552-556;
 * `scale` = a constant factor on this input, e.g. the skip scale 2^-1/2). sums_dev and tickets_dev ([B] int32) must be zero
 * before the first call and are left zero. scale_shift_dev as in adb_cl_groupnorm (indexed over the concatenated channels). */
int adb_cl_gn_coef(const void* in_dev, double* sums_dev, int* tickets_dev, float* coef_dev, int B, int L, int C, int G, int g_total,
                   int g_off, int c_off, int Cin_total, const float* gamma_dev, const float* beta_dev, const float* scale_shift_dev,
                   int64_t ss_ld, float eps, float scale, void* stream);
/* Second half: out = bias + res + Conv1d_k3_same(SiLU(slope * x + offset)) with the activation applied inside the convolution's
 * operand path, so the normalised tensor (and the concatenation) never exists in HBM (unet1d.py:186-193, the residual add of
 * ResnetBlock1d :315). x = in1_dev [B][L][C1], or the channel concatenation [in1 | in2] (in2_dev [B][L][C2], NULL / 0 for one
 * input); coef_dev from adb_cl_gn_coef over C1 + C2 channels; w_packed_dev from adb_cl_pack_conv_weights_f16(C1 + C2, N, 3). */
int adb_cl_gn_conv3(const void* in1_dev, int C1, const void* in2_dev, int C2, const float* coef_dev, const void* w_packed_dev,
                    const float* bias_dev, const void* res_dev, void* out_dev, int B, int L, int N, void* stream);
/* adb_cl_conv (bf16, plain store) over the channel concatenation of two inputs that is never materialised: K-blocks of the first
 * C1 channels come from in1_dev [B][L][C1], the rest from in2_dev [B][L][C2]; a constant factor on the second input is folded
 * into the packed weights by the caller (ResnetBlock1d.to_out on cat(x, skip * 2^-1/2), unet1d.py:291-295, :552-556). */
int adb_cl_conv_cat(const void* in1_dev, int C1, const void* in2_dev, int C2, const void* w_packed_dev, const float* bias_dev,
                    const void* res_dev, void* out_dev, int B, int L, int N, int taps, int off0, int dil, int act, void* stream);
/* row-wise LayerNorm over C (nn.LayerNorm unet1d.py:79 with b_dev; LayerNorm1d unet1d.py:32-45 with b_dev = NULL) */
int adb_cl_layernorm(const void* in_dev, const float* g_dev, const float* b_dev, void* out_dev, int64_t rows, int C, float eps,
                     int dtype, void* stream);
/* softmax(q k^T / sqrt(d)) v per (batch, head); q [B][L][C], kv [B][L][2C] (k | v)   (attention_utils.py:163-184) */
int adb_cl_attention(const void* q_dev, const void* kv_dev, void* out_dev, int B, int L, int C, int heads, int dtype,
                     void* stream);
/* out = [a | scale_b * b] along channels   (UpsampleBlock1d.add_skip, unet1d.py:536-537) */
int adb_cl_concat(const void* a_dev, const void* b_dev, float scale_b, void* out_dev, int64_t rows, int Ca, int Cb, int dtype,
                  void* stream);
/* WAVenc1d on the tensor cores (W == 2 S, L % W == 0, (W * Cin) % 64 == 0): re-lay the fp32 channels-first input as bf16
 * channels-last rows [B][L/W + 1][W * Cin] shifted by pad = W/2 - S/2 samples; adb_cl_conv with taps = 2 over these rows and
 * host-arranged weights [2][W * Cin][2 F] then yields frames (2m, 2m+1) in row m of [B][L/W][2 F] == [B][L/S][F]
 * (unet1d.py:572-594). */
int adb_cl_wavenc_prep(const float* x_dev, void* rows_bf16_dev, int B, int Cin, int L, int W, int S, const float* scale_dev,
                       void* stream);     /* scale_dev: per-sample factor on x (the EDM input scale c_in, diffusion.py:46-48) or NULL */
/* WAVenc1d (unet1d.py:572-594): x [B][Cin][L] fp32 channels-first -> [B][L/S][F] channels-last in `dtype`;
 * WAVdec1d (unet1d.py:596-622): [B][Lc][F] -> y [B][Cout][Lc*S] fp32 channels-first. w in torch layout. */
int adb_cl_wavenc(const float* x_dev, const float* w_dev, void* out_dev, int B, int Cin, int L, int F, int W, int S, int dtype,
                  void* stream);
int adb_cl_wavdec(const void* h_dev, const float* w_dev, float* y_dev, int B, int Lc, int F, int Cout, int W, int S, int dtype,
                  void* stream);
/* WAVdec1d on the tensor cores for bf16 activations (F % 64 == 0, W == 2 S, S * Cout <= 64): the filter bank is packed once
 * (adb_cl_wavdec_packed_elems(F) bf16 elements) and the transposed conv runs as a 2-tap GEMM-convolution with a
 * channels-first fp32 store. */
int64_t adb_cl_wavdec_packed_elems(int F);
int adb_cl_wavdec_pack(const float* w_dev, void* packed_bf16_dev, int F, int Cout, int W, int S, void* stream);
int adb_cl_wavdec_tc(const void* h_bf16_dev, const void* packed_bf16_dev, float* y_dev, int B, int Lc, int F, int Cout, int W,
                     int S, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training step — Diffusion.forward (src/models/components/diffusion.py:65-97) through WaveNetNoise with the
 * backward pass PyTorch autograd would build, AdamW (configs/model/diffunet_complex.yaml:7-12).
 * ADB_PRECISION_FP32: CUDA-core kernels, gradients <= 2e-4 relative to the fp64 evaluation of the same graph.
 * ADB_PRECISION_BF16: tcgen05 forward / data-gradient / weight-gradient GEMMs with bf16 operands and fp32 accumulation
 * (the precision of Lightning's bf16-mixed), block inputs saved, pre-gate activations recomputed.
 * ---------------------------------------------------------------------------------------------- */
int64_t adb_wavenet_train_workspace_bytes(const adb_wavenet* net, int B, int L, int precision);
/* loss[b] = lambda(sigma_b) mean (D(x + sigma_b noise) - x)^2, keeping the activations the backward needs in `workspace_dev` */
int adb_wavenet_dsm_forward_train(adb_wavenet* net, const float* x_dev, const float* noise_dev, const float* sigmas_dev,
                                  float sigma_data, float* loss_dev, int B, int L, int precision, void* workspace_dev,
                                  int64_t workspace_bytes, void* stream);
/* grad_flat[i] = d(sum_b upstream[b] loss[b]) / d(param_i), parameters in the flat order of adb_wavenet_param_count.
 * Must follow adb_wavenet_dsm_forward_train on the same workspace (same x, sigmas). */
int adb_wavenet_dsm_backward(adb_wavenet* net, const float* x_dev, const float* sigmas_dev, float sigma_data,
                             const float* upstream_dev, float* grad_flat_dev, int B, int L, int precision, void* workspace_dev,
                             int64_t workspace_bytes, void* stream);
/* torch.optim.AdamW update on flat fp32 vectors; gradients are multiplied by grad_scale first (1 / world for DDP averaging) */
int adb_adamw_step(float* params_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADB200_H */
