/*
 * ddb200.h -- C-ABI of the B200-native dDDPM hot path (libddb200.so).
 *
 * The reference (simonamtoft/downsampled-diffusion) has no FFI: every operation below is an
 * ATen call made from Python.  Each entry point cites the reference call site it replaces
 * (paths relative to the reference repo root).  A maintainer binds these with ctypes
 * (see INTEGRATION.md); downsampled_diffusion_b200/_lib.py is that binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - activations are NHWC ("channels-last") contiguous; `dtype` is DD_F32 or DD_BF16;
 *     tensors crossing the Python API (x_t, eps_hat, noise, images) are NCHW fp32 like the reference;
 *   - `stream` is a cudaStream_t passed as void*; every call is stream-ordered and re-entrant,
 *     the library keeps no global mutable state besides the thread-local error string;
 *   - return 0 on success, a negative DD_ERR_* otherwise (message: dd_last_error());
 *   - the library never allocates or frees device memory and never retains a pointer
 *     past the call (TMA descriptors are built by value into the launch parameters).
 */
#ifndef DDB200_H_
#define DDB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DD_F32  0
#define DD_BF16 1

#define DD_OK              0
#define DD_ERR_ARG        -1   /* unsupported shape / bad argument (Python raises ValueError/RuntimeError) */
#define DD_ERR_CUDA       -2   /* CUDA runtime / driver error */
#define DD_ERR_NO_DEVICE  -3   /* no sm_100 device */

const char* dd_last_error(void);
int dd_version(void);
/* 1 when the current device is compute capability 10.x (tcgen05/TMA kernels usable). */
int dd_device_ok(void);

/* ------------------------------------------------------------------------------------------
 * Diffusion arithmetic (models/diffusion/ddpm.py)
 * ---------------------------------------------------------------------------------------- */

/* q_sample: ddpm.py:256-273.  out = sqrt_ac[t_b]*x + sqrt_1mac[t_b]*eps, t per sample.
 * x/eps/out: (B, chw) fp32.  t: int64 (B,).  Separate mul/mul/add roundings like ATen. */
int dd_q_sample(const float* x, const float* eps, const int64_t* t,
                const float* sqrt_ac, const float* sqrt_1mac,
                float* out, int B, int64_t chw, void* stream);

/* One ancestral update given eps_hat: ddpm.py:149-158 (predict_x_from_eps, clamp), :160-185
 * (q_posterior), :217-227 (noise injection, masked at t==0).
 * coef: (T,5) fp32 rows {sqrt_recip_ac, sqrt_recipm1_ac, post_mean_coef1, post_mean_coef2,
 * exp(0.5*post_logvar_clipped)}.  t_idx: int32 device array, sample b uses t_idx[b*t_stride]
 * (t_stride 0 = one shared step, the sampling loop).  noise_step_stride != 0 selects
 * noise + ((T-1-t) mod noise_period)*noise_step_stride (no mod when noise_period == 0), i.e. the pre-drawn
 * chain noise of ddpm.py:223 for this step, held in a ring of noise_period steps.
 * clip != 0 applies clamp(-1,1) to x0 (ddpm.py:156-157). All NCHW fp32, (B, chw). */
int dd_posterior_step(const float* x_t, const float* eps_hat, const float* noise,
                      const float* coef, const int32_t* t_idx, int t_stride,
                      int64_t noise_step_stride, int T, int noise_period, int clip,
                      float* x_out, int B, int64_t chw, void* stream);

/* predict_x_from_eps alone (ddpm.py:149-158), per-sample int64 t (used by reconstruct / non-AE loss). */
int dd_predict_x0(const float* x_t, const float* eps, const int64_t* t,
                  const float* sqrt_recip_ac, const float* sqrt_recipm1_ac, int clip,
                  float* out, int B, int64_t chw, void* stream);

/* t_idx[i] -= 1 for i < n : advances the device-side step counter between graph replays
 * (replaces the host-side torch.full of ddpm.py:247). */
int dd_tick(int32_t* t_idx, int n, void* stream);

/* loss_ddpm's inner part: ddpm.py:279-283 + utils/utils.py:27-40.
 * out[b] = sum_chw (a-b)^2  (scale=1) or mean (scale=1/chw). */
int dd_mse_rowsum(const float* a, const float* b, float* out, int B, int64_t chw, float scale, void* stream);
/* d/d b of mean_b(out[b]*w[b]) :  grad_b = 2*(b-a)*scale*gscale*w[b] (w may be NULL). */
int dd_mse_rowsum_bwd(const float* a, const float* b, const float* w, float* grad_b,
                      int B, int64_t chw, float gscale, void* stream);

/* EMA.update: trainers/ema.py:36-44.  For every tensor i: shadow_i = shadow_i*decay + one_minus*param_i.
 * table: device array of n_tensors*3 uint64 {shadow_ptr, param_ptr, numel}; chunk: device array of
 * n_chunks*2 int32 {tensor index, chunk index}, each chunk = chunk_elems elements. */
int dd_ema_update(const uint64_t* table, const int32_t* chunks, int n_chunks, int chunk_elems,
                  float decay, float one_minus_decay, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer step of the training loop: trainer_ddpm.py:128-135 / :243-250 = clip_grad_norm_(1.0),
 * Adam.step (trainers/trainer.py:69), zero_grad, EMA.update.  table: device array of n_tensors*6 uint64
 * {param, grad, exp_avg, exp_avg_sq, shadow (0 = none), numel}, all fp32; chunks as in dd_ema_update.
 * ---------------------------------------------------------------------------------------- */

/* Total gradient L2 norm over every tensor and the clip coefficient, left on the device:
 * norm_out[0] = ||g||_2, norm_out[1] = min(max_norm / (||g||_2 + 1e-6), 1) (1 when max_norm <= 0).
 * partial: n_chunks floats of scratch.  torch.nn.utils.clip_grad_norm_ without its host synchronisation. */
int dd_grad_norm(const uint64_t* table, const int32_t* chunks, int n_chunks, int chunk_elems, float max_norm, float* partial,
                 float* norm_out, void* stream);

/* One pass over all parameters: g *= norm_out[1] (skipped when norm_out == NULL); torch.optim.Adam's update
 * (exp_avg.lerp_(g, 1-beta1); exp_avg_sq = beta2*exp_avg_sq + (1-beta2)*g*g; p += neg_step_size * exp_avg /
 * (sqrt(exp_avg_sq)/bias_correction2_sqrt + eps), neg_step_size = -lr/(1-beta1^step)); then the EMA of trainers/ema.py:36-44
 * on the fresh parameter (ema_mode 1), a plain copy (2: EMA.reset during warm-up, trainer_ddpm.py:107-109) or nothing (0);
 * zero_grad != 0 clears the gradient in place. */
int dd_adam_ema_step(const uint64_t* table, const int32_t* chunks, int n_chunks, int chunk_elems, const float* norm_out,
                     float one_minus_beta1, float beta2, float one_minus_beta2, float bias_correction2_sqrt, float eps,
                     float neg_step_size, int ema_mode, float decay, float one_minus_decay, int zero_grad, void* stream);

/* The 'deterministic' resampler (convblocks.py:8-26; wrapper.py:22-24, 49-53): F.interpolate(size=(Hout, Wout),
 * mode='bicubic', align_corners=True) on `planes` = B*C contiguous fp32 planes, cubic convolution A = -0.75, taps clamped to
 * the image; and its input gradient (gx is overwritten). */
int dd_bicubic2d(const float* x, float* y, int planes, int Hin, int Win, int Hout, int Wout, void* stream);
int dd_bicubic2d_bwd(const float* gy, float* gx, int planes, int Hin, int Win, int Hout, int Wout, void* stream);

/* Multi-tensor re-layout (weight packing after an optimizer step, packed weight gradients back to parameter layout):
 * for every element i of every segment, dst[i] = code ? ((const float*)src_table[(code >> 32) - 1])[code & 0xffffffff] : 0.
 * segs: n_segs*3 uint64 {dst pointer, index of the segment's first code, element count}; blocks: n_blocks*2 int32
 * {segment, chunk of 4096 elements}; codes: uint64 per destination element.  dst0 != NULL replaces segment 0's destination
 * (a buffer allocated per call). */
int dd_gather_f32(const uint64_t* segs, const int32_t* blocks, int n_blocks, const uint64_t* codes, const uint64_t* src_table,
                  float* dst0, void* stream);

/* ------------------------------------------------------------------------------------------
 * Evaluation-side chain and sampler output formatting
 * tab: (T, 8) fp32 rows {sqrt_ac, sqrt_1mac, sqrt_recip_ac, sqrt_recipm1_ac, post_mean_coef1, post_mean_coef2,
 * post_logvar_clipped, 0} -- the scalars ddpm.py gathers with extract() (helpers.py:31-34).
 * ---------------------------------------------------------------------------------------- */

/* q_sample (ddpm.py:256-273) for ONE step shared by the batch, read from the device-side step counter t_idx[0]
 * (test_losses_, ddpm.py:409-411); eps = noise + ((T-1-t) mod noise_period) * noise_step_stride (no mod when 0). */
int dd_q_sample_step(const float* x, const float* noise, int64_t noise_step_stride, int noise_period, const float* tab,
                     const int32_t* t_idx, int T, float* out, int B, int64_t chw, void* stream);

/* DDPM.vlb_terms (ddpm.py:317-365) given eps_hat = UNet(x_t, t): predict_x_from_eps(clip) -> q_posterior of both
 * means -> normal_kl (losses.py:17-52) for t > 0, minus discretized_gaussian_log_likelihood (losses.py:66-109) for
 * t == 0 -> flat_bits (utils/utils.py:43-48).  Sample b uses t_idx[b*t_stride].  vlb[b*out_stride + col], col = T-1-t
 * when col_from_t (the stacking order of ddpm.py:423) else 0.  When sq != NULL also sq[...] = sum_chw (eps - eps_hat)^2
 * (the L_simple term, ddpm.py:418-419), eps addressed like dd_q_sample_step's noise (eps_step_stride 0 = plain (B, chw)). */
int dd_vlb_terms(const float* x, const float* x_t, const float* eps_hat, const float* eps, int64_t eps_step_stride, int eps_period,
                 const float* tab, const int32_t* t_idx, int t_stride, int T, float* vlb, float* sq, int64_t out_stride,
                 int col_from_t, int B, int64_t chw, void* stream);

/* DDPM.calc_prior (ddpm.py:367-389): out[b] = flat_bits(normal_kl(sqrt_ac[T-1] x, log(1-ac[T-1]), 0, 0)). */
int dd_prior_kl(const float* x, float sqrt_ac_last, float log_1mac_last, float* out, int B, int64_t chw, void* stream);

/* fix_samples (utils/eval_helpers.py:37-41; min_max_norm_image, utils/utils.py:16-24): per-image
 * (x - min) / (max - min) * 255, NCHW fp32 -> NHWC fp32 (the np.moveaxis of the reference). */
int dd_fix_samples(const float* x, float* y, int B, int C, int H, int W, void* stream);

/* ------------------------------------------------------------------------------------------
 * U-Net pieces (models/unet/unet.py, models/unet/blocks.py)
 * ---------------------------------------------------------------------------------------- */

/* NCHW fp32 -> NHWC dtype (input staging for unet.py:74). */
int dd_nchw_to_nhwc(const float* x, void* y, int dtype, int B, int C, int H, int W, void* stream);
/* the same into bf16 NHWC with the channel dimension zero-padded to Cp (a multiple of 8, >= C): the U-Net input (unet.py:81) as a
 * regular 64-channel activation for the tensor-core path (pack the first convolution's weights with zeros for the pad channels) */
int dd_nchw_to_nhwc_pad(const float* x, void* y_bf16, int B, int C, int H, int W, int Cp, void* stream);
/* NHWC dtype -> NCHW fp32. */
int dd_nhwc_to_nchw(const void* x, int dtype, float* y, int B, int C, int H, int W, void* stream);
/* NCHW fp32 -> bf16 im2col rows (B*H*W, kpad): column tap*C+c holds x[b,c,h+dy,w+dx] (3x3, zero pad 1),
 * zero for columns >= 9*C.  Lets the first 3x3 conv (tiny C_in) and its res_conv run as K=kpad GEMMs. */
int dd_im2col3x3_nchw(const float* x, void* y_bf16, int B, int C, int H, int W, int kpad, void* stream);

/* SinusoidalPosEmb + time_mlp + every ResnetBlock's mlp (blocks.py:22-29, unet.py:30-35, blocks.py:92-95,108):
 * out[r, j] = Wcat[j,:] . mish(temb(t_r)) + bcat[j], temb = W2 . mish(W1 . sincos(t_r) + b1) + b2.
 * t: float32 (R,) time values (the reference multiplies the int64 t by fp32 frequencies); freq: (dim/2,)
 * fp32 = exp(arange(dim/2) * -log(10000)/(dim/2-1)) computed by the host exactly like blocks.py:25-26. */
int dd_time_bias(const float* t, int R, int dim, const float* freq,
                 const float* W1, const float* b1, const float* W2, const float* b2,
                 const float* Wcat, const float* bcat, int J, float* out, void* stream);

/* GroupNorm statistics, blocks.py:79 (biased variance, eps inside sqrt).  x: NHWC (B,HW,C).
 * stats: (B, G, 2) fp32 = {mean, rstd}. */
int dd_gn_stats(const void* x, int dtype, int B, int HW, int C, int G, float eps, float* stats, void* stream);

/* GroupNorm apply + Mish (+ time-embedding bias) (+ residual): blocks.py:79-84, 108-109, 115.
 * stats_mode 0: stats = {mean, rstd}; 1: stats = {sum, sumsq} accumulated by the conv epilogue.
 * tbias: row table (rows, tb_stride) already offset to this block's first column, or NULL;
 * sample b uses row trow[b*trow_stride] (trow NULL: row b).  residual: NHWC same dtype, or NULL. */
int dd_gn_mish(const void* x, void* y, int dtype, int B, int HW, int C, int G,
               const float* stats, int stats_mode, float eps,
               const float* gamma, const float* beta,
               const float* tbias, int tb_stride, const int32_t* trow, int trow_stride,
               const void* residual, void* stream);

/* dd_gn_mish for a split-K convolution: x = sum_s part[s] + bias (part: (S, B, HW, C) fp32 from dd_conv_tc with
 * DD_TC_SPLITK), GroupNorm statistics computed in the same launch (one CTA per image, HW*C <= 16384), then
 * y = mish(gn(x)*gamma+beta) [+ tbias] [+ residual] as above.  y / residual: bf16 NHWC. */
int dd_gn_mish_sum(const float* part, int S, const float* bias, void* y_bf16, int B, int HW, int C, int G, float eps,
                   const float* gamma, const float* beta,
                   const float* tbias, int tb_stride, const int32_t* trow, int trow_stride,
                   const void* residual_bf16, float* ln_part, void* stream);
/* CTAs per image dd_gn_mish_sum uses for (C, G) = the `parts` dimension of its ln_part output: (B*HW, parts, 2) fp32 per-pixel
 * {sum, sum of squares} over each CTA's channels of the bf16 values it wrote (channel-LayerNorm statistics for dd_conv_tc_ln). */
int dd_gn_mish_sum_parts(int C, int G);

/* Channel LayerNorm of blocks.py:50-60: (x-mean_c)/(sqrt(var_c)+eps)*g+b, per pixel. x,y NHWC (P, C). */
int dd_layernorm_c(const void* x, void* y, int dtype, int64_t P, int C, const float* g, const float* b,
                   float eps, void* stream);

/* LinearAttention core, blocks.py:128-133: qkv NHWC (B, n, 3*heads*dh), channel = (qkv, head, c);
 * k softmax over n, ctx = k v^T, out = ctx^T q.  out NHWC (B, n, heads*dh).  dh must be 32.
 * Two launches (partial contexts per n-split, then merge + output); ws: fp32 scratch of at least
 * dd_linattn_ws_floats(B, n, heads) floats, owned by the caller. */
int64_t dd_linattn_ws_floats(int B, int n, int heads);
int dd_linattn_core(const void* qkv, void* out, int dtype, int B, int n, int heads, int dh,
                    float* ws, int64_t ws_floats, void* stream);

/* Fused form for the tensor-core path (bf16 only): context on warp-level MMA with an online softmax, then
 *   Mb[b][c][h*dh+d] = sum_e Wout[c][h*dh+e] * ctx_{b,h}[d][e]   (bf16, (B, C, heads*dh), K-major)
 * so that to_out(attention)[n][c] = sum_k q[n][k] * Mb[b][c][k] + bias[c] (blocks.py:132-134) is a single
 * dd_conv_tc 1x1 launch with DD_TC_W_PER_SAMPLE reading q straight out of the qkv tensor.  Wout: (C, heads*dh) bf16.
 * One launch: the last split CTA of each (b, head) merges and folds.  ws: dd_linattn_mix_ws_floats(B, n, heads) floats
 * that must be ZERO before the first call (the trailing B*heads arrival tickets reset themselves). */
int64_t dd_linattn_mix_ws_floats(int B, int n, int heads);
int dd_linattn_mix(const void* qkv, int dtype, int B, int n, int heads, int dh, float* ws, int64_t ws_floats,
                   const void* Wout_bf16, int C, void* Mb_bf16, void* stream);

/* Generic direct convolution on CUDA cores, fp32 accumulate (validation mode, odd shapes, and the
 * down/up-sampling nets).  Replaces F.conv2d / F.conv_transpose2d call sites of blocks.py:35,44,78,103,
 * 123-124, unet.py:71, convblocks.py:29-67.
 *   x (B,H,W,C1) [+ x2 (B,H,W,C2) concatenated on channels: unet.py:97], w (taps, C1+C2, Cout) fp32,
 *   mode 0: conv k x k, stride, pad;  mode 1: transposed conv, oh = ih*stride - pad + ky, output size stride*H
 *           (ConvTranspose2d(4,2,1) of blocks.py:35, and the input gradient of a strided conv); w tap = ky*k+kx.
 *   flags: DD_CONV_PRE_MISH applies Mish to the input on load (convblocks.py:118-121),
 *          DD_CONV_TANH applies tanh to the output (dddpm.py:99-100,110-111),
 *          DD_CONV_OUT_NCHW writes fp32 NCHW, DD_CONV_IN_NCHW reads fp32 NCHW (x2 must be NULL). */
#define DD_CONV_PRE_MISH  1
#define DD_CONV_TANH      2
#define DD_CONV_OUT_NCHW  4
#define DD_CONV_IN_NCHW   8
int dd_conv_direct(const void* x, const void* x2, int C1, int C2, int in_dtype,
                   const float* w, const float* bias, const void* residual,
                   void* y, int out_dtype, int B, int H, int W, int Cout,
                   int ksize, int stride, int pad, int mode, int flags, void* stream);

/* avg_pool2d(2,2) (convblocks.py:129) and nearest x2 (convblocks.py:127), NHWC. */
int dd_avgpool2(const void* x, void* y, int dtype, int B, int H, int W, int C, void* stream);
int dd_upsample_nearest2(const void* x, void* y, int dtype, int B, int H, int W, int C, void* stream);

/* Space-to-depth: NHWC bf16 (B,H,W,C) -> 4 parity planes (4,B,H/2,W/2,C), plane = (h&1)*2+(w&1).
 * Feeds the stride-2 Downsample conv (blocks.py:44) as 9 unit-stride taps. */
int dd_space_to_depth2(const void* x, void* y, int B, int H, int W, int C, void* stream);

/* tcgen05 / TMEM / TMA implicit-GEMM convolution, bf16 operands, fp32 accumulate.
 * D[pixel, co] = sum_{tap, c} X_tap[pixel, c] * Wp[co, tap*(C1+C2) + c]
 *   kind DD_TC_CONV3x3 : 3x3 stride 1 pad 1 (blocks.py:78)            x (B,H,W,C)
 *   kind DD_TC_CONV1x1 : pointwise (blocks.py:103,123,124, unet.py:71) x (B,H,W,C)
 *   kind DD_TC_DOWN    : 3x3 stride 2 pad 1 (blocks.py:44)  x = space-to-depth planes (4,B,H,W,C), H,W = OUTPUT size
 *   kind DD_TC_UPT     : ConvTranspose2d(4,2,1) (blocks.py:35) as 4 sub-pixel phases; Wp rows = phase*Cout_pad+co,
 *                        K = 4 taps * C; output (B,2H,2W,Cout)
 * x_pitch: channels per pixel of the tensor x points into (0 = C1); lets a conv read the first C1 channels of a
 * wider tensor (q out of qkv).  x2 (optional, same geometry, C2 channels) is the skip tensor of unet.py:97 (concat-free).
 * C1, C2 multiples of 64; Wp (rows, K) bf16 K-major, rows padded to bn.  H, W: any size -- maps that are not powers of two
 * (28 -> 14 -> 7) are tiled as if padded to the next power of two (TMA zero fill over the edge, rows outside the image skipped
 * by the epilogue) and take this plain entry point only: dd_conv_tc_gn_cluster / dd_conv_tc_gn_ws_floats return 0 and
 * dd_conv_tc_splits returns 1 for them, DD_TC_SPLITK / DD_TC_PAIR are rejected.
 * Epilogue: + bias, GroupNorm {sum,sumsq} atomics into gn_stats (B, G, 2) when non-NULL (stats of the
 * fp32 accumulator + bias), + residual (bf16 NHWC, output geometry), store bf16 NHWC or fp32 NCHW
 * (out_nchw_f32, only the first cout_valid channels).
 * splitk_ws / splitk_cnt (optional): caller-owned fp32 scratch and int32 counters, ALL ZERO on entry and left all
 * zero on exit; when given, layers with fewer output tiles than SMs split their K loop over several CTAs that
 * reduce through the scratch (red.add), the last-arriving CTA running the epilogue. */
#define DD_TC_CONV3x3 0
#define DD_TC_CONV1x1 1
#define DD_TC_DOWN    2
#define DD_TC_UPT     3
/* flags: DD_TC_W_PER_SAMPLE (1x1 only): wp is (B, w_rows, K), image b multiplies its own matrix -- the fused
 * LinearAttention output GEMM  out = q . (ctx_b . W_out^T)  (blocks.py:132-134). */
#define DD_TC_W_PER_SAMPLE 1
/* DD_TC_SPLITK (3x3 only, where dd_conv_tc_splits() > 1): the K range is split over S = dd_conv_tc_splits(...)
 * CTAs per tile; every split writes its raw fp32 partial sums (no bias) to
 *   splitk_ws[split][b][h][w][Cout]            (S * B*H*W*Cout floats),
 * y, gn_stats and residual are ignored, and dd_gn_mish_sum consumes the partials (sum + bias + GroupNorm + Mish). */
#define DD_TC_SPLITK 2
/* DD_TC_PAIR (3x3 on maps >= 16x8, Cout a multiple of 128, an even number of pixel tiles): run the halo kernel as thread-block
 * clusters of two CTAs with tcgen05 cta_group::2 (M = 256 per MMA, each CTA loads half of the weight rows).  Opt-in. */
#define DD_TC_PAIR 4
/* DD_TC_STRIDED_IN (DD_TC_DOWN only): x is the plain NHWC input (B, 2H, 2W, C) instead of its four space-to-depth parity
 * planes; the kernel reads every other pixel through a stride-2 tensor map (TMA elementStrides), no copy. */
#define DD_TC_STRIDED_IN 8
/* Stream ordering of the PARAMETER operands (wp, bias, gamma / beta, wsum): the launches are chained by programmatic dependent
 * launch, and the weight producer of dd_conv_tc / dd_conv_tc_gn / dd_conv_tc_ln requests its first tiles BEFORE the dependency on
 * the preceding launch in the stream resolves (they never depend on it inside a step).  A caller that writes these buffers must
 * therefore not make the writing kernel the IMMEDIATELY preceding launch in the same stream (any launch in between, an event or a
 * synchronisation restores the order; the Python layer packs weights before the step's graph is launched).  Activations, residuals,
 * statistics and DD_TC_W_PER_SAMPLE weight matrices are ordered as usual.  DD_NO_PDL=1 in the environment turns the programmatic
 * launches (and with them this relaxation) off. */
int dd_conv_tc_splits(int kind, int B, int H, int W, int Cin, int Cout);
int dd_conv_tc(int kind, const void* x, int x_pitch, const void* x2, int C1, int C2,
               const void* wp, int w_rows, const float* bias, const void* residual,
               void* y, int out_nchw_f32, int cout_valid,
               float* gn_stats, int G,
               int B, int H, int W, int Cout, int flags,
               float* splitk_ws, int64_t splitk_ws_floats, int32_t* splitk_cnt, int splitk_cnt_n, void* stream);

/* dd_conv_tc with GroupNorm + Mish (+ time-embedding bias) (+ residual) fused into its epilogue -- the `Block` of
 * models/unet/blocks.py:73-84 as ONE launch, and the two halves of ResnetBlock.forward (:105-115):
 *     y = mish(GroupNorm_G(conv(x|x2) + bias) * gamma + beta)  [+ tbias[row(n), c]]  [+ residual]         bf16 NHWC
 * The statistics of a (sample, group) cover the whole image: the pixel tiles of one image are launched as one thread-block
 * cluster of dd_conv_tc_gn_cluster(...) CTAs which exchange their partial sums through distributed shared memory (1: a tile
 * holds whole images).  The normalisation reads the fp32 accumulator (no bf16 round trip, no statistics atomics).
 *   tbias  : optional fp32 rows (row stride tb_stride floats), pointer already offset to this layer's first column;
 *            row(n) = trow[n * trow_stride] (trow_stride 0: one device-side step counter for the whole batch) or n when trow == NULL
 *   residual: optional bf16 NHWC tensor of the output geometry, added AFTER the activation (blocks.py:115)
 *   ln_part : optional (B*H*W, Cout/bn, 2) fp32, bn = 128 (Cout >= 128, more than half a wave of tiles) else 64: per-pixel
 *             {sum, sum of squares} over the written channels, the channel-LayerNorm statistics a following PreNorm needs
 * dd_conv_tc_gn_cluster returns 0 for layers that cannot take this epilogue (transposed convs, images of more than 8 tiles,
 * fewer than 8 channels per group): use dd_conv_tc with gn_stats + dd_gn_mish there.  kinds / flags as dd_conv_tc (no
 * DD_TC_SPLITK, DD_TC_W_PER_SAMPLE, DD_TC_PAIR). */
int dd_conv_tc_gn_cluster(int kind, int B, int H, int W, int Cout, int G);
/* output channels per CTA tile that dd_conv_tc / dd_conv_tc_gn use for this layer (ln_part has Cout / tile_n parts) */
int dd_conv_tc_tile_n(int kind, int B, int H, int W, int Cout, int flags);
/* 1x1 convolution of the channel-LayerNorm'd input (PreNorm + to_qkv, blocks.py:57-69, 123) without materialising the norm:
 *     y[p, o] = inv_p * (sum_c Wg[o,c] x[p,c] - mean_p * wsum[o]) + bias[o],    inv_p = 1 / (std_p + eps)
 * wp = bf16(W * g) (gain folded into the weights), wsum[o] = sum_c wp[o,c], bias[o] = sum_c W[o,c] b[c]; mean_p / std_p come
 * from ln_in (B*H*W, ln_in_parts, 2) fp32 {sum, sum of squares} over the C input channels, as written by the producing launch
 * (dd_conv_tc_gn's ln_part, dd_gn_mish_sum's ln_part).  x, y bf16 NHWC. */
int dd_conv_tc_ln(const void* x, int C, const void* wp, int w_rows, const float* bias, const float* wsum,
                  const float* ln_in, int ln_in_parts, float ln_eps, void* y, int B, int H, int W, int Cout, void* stream);
/* Persistent form (3x3 stride 1 on maps of at least 16x8 pixels, Cout % 128 == 0): when dd_conv_tc_gn_ws_floats(...) > 0 and the
 * caller passes `ws`, a ZEROED fp32 workspace of that many floats, the layer runs as one CTA per SM walking its tiles with two
 * accumulators in TMEM -- the epilogue (statistics exchange through `ws`, normalisation, Mish, stores) of one tile under the
 * MMAs of the next (csrc/conv_tc_persist.cu).  ws == NULL selects the cluster form above.
 * dd_conv_tc_gn_ln_parts: the `parts` dimension of ln_part for the persistent (2 * Cout / 128) or the cluster form. */
int64_t dd_conv_tc_gn_ws_floats(int kind, int B, int H, int W, int Cout, int G);
int dd_conv_tc_gn_ln_parts(int kind, int B, int H, int W, int Cout, int G, int persistent);
int dd_conv_tc_gn(int kind, const void* x, int x_pitch, const void* x2, int C1, int C2,
                  const void* wp, int w_rows, const float* bias, void* y,
                  int B, int H, int W, int Cout, int flags,
                  int G, float eps, const float* gamma, const float* beta,
                  const float* tbias, int tb_stride, const int32_t* trow, int trow_stride,
                  const void* residual, float* ln_part, float* ws, void* stream);

/* fp32 training form of the tensor-core convolution (forward AND input gradient of the 3x3 stride-1 / 1x1 convolutions
 * of blocks.py:78,103,123-124 and convblocks.py:29-67): x, x2, y, addend fp32 NHWC; wp fp32 (w_rows, taps*(C1+C2)) K-major;
 * TF32 operands (tcgen05.mma.kind::tf32), fp32 accumulate, persistent CTAs with two TMEM accumulators.
 *   y = (conv(x|x2) + bias) [* mish'(mish_grad_of)] [+ addend];   y_mish = mish(y) when y_mish != NULL
 * addend may alias y (gradient accumulation); mish_grad_of (same shape as y) makes the launch the input gradient of a
 * conv that applies Mish first (convblocks.py:112-118); y_mish is the activated copy the next such conv consumes.
 * Channel counts multiples of 32, power-of-two maps; anything else stays on dd_conv_direct. */
int dd_conv_tc32(int kind, const float* x, const float* x2, int C1, int C2, const float* wp, int w_rows, const float* bias,
                 const float* addend, float* y, float* y_mish, const float* mish_grad_of,
                 int B, int H, int W, int Cout, void* stream);

/* 1x1 convolutions with a narrow side (<= 8 channels) -- the 3 -> 64 input layer and the 64 -> 3 (+tanh) output layer of the
 * resampling nets (convblocks.py:143-156, dddpm.py:92-112) and their gradients: HBM streaming kernels, the narrow operand in the
 * NCHW fp32 layout of the Python API, the wide one NHWC fp32 (a power-of-two multiple of 4 channels, <= 128).
 *   thin_in :  y[b][p][c] (+)= sum_s x[b][s][p] * w[s][c] + bias[c]            w: (Cs, Cout)   (forward 3->64; input gradient of 64->3)
 *   thin_out:  y[b][s][p] = act(sum_c x[b][p][c] * w[s][c] + bias[s])          w: (Cs, Cin)    (forward 64->3, optional tanh)
 *   thin_wgrad: dw[s][c] += sum_{b,p} narrow[b][s][p] * wide[b][p][c]  ((Cs, Cw) when narrow_major, else (Cw, Cs));
 *               dbias_wide[c] += sum wide (optional). */
int dd_conv1x1_thin_in(const float* x_nchw, const float* w, const float* bias, float* y_nhwc, int B, int HW, int Cs, int Cout,
                       int accumulate, void* stream);
int dd_conv1x1_thin_out(const float* x_nhwc, const float* w, const float* bias, float* y_nchw, int B, int HW, int Cin, int Cs,
                        int do_tanh, void* stream);
int dd_conv1x1_thin_wgrad(const float* narrow_nchw, const float* wide_nhwc, float* dw, int narrow_major, float* dbias_wide,
                          int B, int HW, int Cs, int Cw, void* stream);

/* fp32 NHWC space-to-depth (to_packed != 0): src (B, 2h, 2w, C) -> dst (B, h, w, 4C), packed channel (py*2+px)*C + c =
 * src[2i+py][2j+px][c]; depth-to-space (to_packed == 0): src packed -> dst full.  With it the stride-2 conv and the
 * transposed conv of the U-Net (blocks.py:35,44) are dense 3x3 stride-1 convolutions for dd_conv_tc32 (zero weight blocks
 * where a (tap, plane) pair does not occur). */
int dd_s2d_f32(const float* src, float* dst, int B, int h, int w, int C, int to_packed, void* stream);

/* Channel-major padded copies for the weight-gradient GEMM:  y[s][b][c][h + hpad][w] = x[b][h][w + s - nshift/2][c],
 * fp32, rows Wp (multiple of 32, >= W) wide, zeros outside the map; nshift = 3 bakes the column shifts of a 3x3 filter
 * into three copies (a TMA box origin must be 16-byte aligned), nshift = 1 is the plain copy.
 * colsum (optional, C floats): += sum over all pixels of x[..][c] -- the bias gradient when x is an output gradient. */
int dd_nhwc_to_chw_pad(const float* x, float* y, int B, int C, int H, int W, int Wp, int hpad, int nshift, float* colsum,
                       void* stream);

/* Weight gradient of the same convolutions on the tensor cores (TF32, pixels as the GEMM K dimension):
 *   dw[tap][ci][co] += sum_pixels x[pixel+tap][ci] * dy[pixel][co].
 * kind::tf32 reads K-major operands only, so the operands are channel-major copies made by dd_nhwc_to_chw_pad:
 *   x_cm (, x2_cm): (nshift, B, C, H+2, Wp), hpad = 1, nshift = 3 for 3x3 and 1 for 1x1 -- the conv's input
 *                   (the activated copy when the block applies Mish first);
 *   dy_cm: (B, Cout, H, Wp), hpad = 0, nshift = 1.
 * dw: fp32 (taps, C1+C2, Cout), ACCUMULATED with red.global.add (zero it first). */
int dd_conv_wgrad_tc32(int kind, const float* x_cm, const float* x2_cm, int C1, int C2, const float* dy_cm, float* dw,
                       int B, int H, int W, int Wp, int Cout, void* stream);

/* ------------------------------------------------------------------------------------------
 * Backward of the training denoising step (fp32 programs): autograd of ddpm.py:290-315 / dddpm.py:122-177.
 * Input gradients of convolutions are dd_conv_direct launches with transposed / flipped weights.
 * ---------------------------------------------------------------------------------------- */
/* dW[tap][ci][co] += sum_pixels X_tap[pixel][ci] * dY[pixel][co]; geometry and flags as dd_conv_direct
 * (DD_CONV_OUT_NCHW: dy is fp32 NCHW; DD_CONV_PRE_MISH / DD_CONV_IN_NCHW apply to x).  dw fp32 (taps, C1+C2, Cout). */
int dd_conv_wgrad(const void* x, const void* x2, int C1, int C2, int in_dtype, const float* dy, float* dw,
                  int B, int H, int W, int Cout, int ksize, int stride, int pad, int mode, int flags, void* stream);
/* out[c] += sum_m x[m][c]  (x: (M, C) fp32, or NCHW with M = B*HW when nchw != 0): bias gradients. */
int dd_colsum(const float* x, float* out, int64_t M, int C, int nchw, int64_t HW, void* stream);
/* y = x * mask / (1-p), mask from a counter-based hash of (seed, index): nn.Dropout of blocks.py:111 (the same
 * call with the same seed masks the gradient in the backward pass).  seed = hash(salt) ^ *seed_dev: the per-step seed
 * is read from device memory (seed_dev may be NULL) so that the launch can sit in a replayed CUDA graph; salt separates
 * the dropout layers of one network. */
int dd_dropout(const float* x, float* y, int64_t n, uint32_t salt, const uint32_t* seed_dev, float p, void* stream);
/* GroupNorm+Mish backward (blocks.py:79-84): x pre-norm conv output, stats {mean,rstd}; writes per-(b,c) sums
 * s_dhxh, s_dh, s_dy ((B, C) fp32: reduce over b for dgamma, dbeta; s_dy is the time-bias gradient) and dx. */
int dd_gn_mish_bwd(const float* x, const float* dy, const float* stats, const float* gamma, const float* beta,
                   int B, int HW, int C, int G, float* s_dhxh, float* s_dh, float* s_dy, float* dx, int accumulate,
                   void* stream);
/* channel LayerNorm backward (blocks.py:57-60); dg, db accumulate (+=). */
int dd_layernorm_c_bwd(const float* x, const float* dy, const float* g, float eps, int64_t P, int C, float* dx,
                       int accumulate, float* dg, float* db, void* stream);
/* merged attention state {max_d, sum_d, ctx[d][e]} per (b, head) from the partials dd_linattn_core left in ws. */
int dd_linattn_save(const float* ws, int B, int n, int heads, float* saved, void* stream);
/* LinearAttention core backward (blocks.py:128-133): dqkv from dout; dctx: (B*heads*1024) fp32 scratch. */
int dd_linattn_bwd(const float* qkv, const float* dout, const float* saved, float* dctx, float* dqkv,
                   int B, int n, int heads, int dh, void* stream);
/* elementwise: mode 0 y=mish(x); 1 y=g*mish'(x); 2 y=g*(1-x^2) (tanh backward, x = tanh output); 3 y=alpha*x;
 * 4 y=tanh(x); 5 y=(x-g)^2 (l2_loss(..., reduction='none'), models/utils/losses.py:12-14); accumulate != 0 adds into y. */
int dd_ew(int mode, const float* x, const float* g, float* y, int64_t n, float alpha, int accumulate, void* stream);
/* SinusoidalPosEmb (blocks.py:22-29): out (R, dim) = [sin(t*freq), cos(t*freq)]. */
int dd_sincos_emb(const float* t, const float* freq, float* out, int R, int dim, void* stream);
/* y[b,h/2,w/2,c] = scale * (2x2 block sum of x)  and  y[b,2h,2w,c] = scale * x[b,h,w,c]  (fp32 NHWC): the
 * forward/backward pair of avg_pool2d(2,2) and nearest x2 (convblocks.py:127-129). */
int dd_pool2_sum(const float* x, float* y, int B, int H, int W, int C, float scale, void* stream);
int dd_unpool2(const float* x, float* y, int B, int H, int W, int C, float scale, void* stream);

/* Debug only: when buf != NULL every dd_conv_tc CTA writes 8 clock64() stamps (start, prologue done, dependency
 * wait done, first operands landed, last MMA issued, accumulator visible, epilogue done) to buf[cta*8 + i]. */
int dd_debug_set_timeline(long long* buf);
/* Tuning aid: resident thread-block clusters of `cluster` halo-convolution CTAs (cudaOccupancyMaxActiveClusters), -1 on error. */
int dd_debug_max_clusters(int cluster);
/* Same for dd_linattn_mix (6 stamps per CTA: start, dependency resolved, context done, warps merged, context normalised, end);
 * only instrumented builds (-DDD_ATTN_TIMELINE=1) write them. */
int dd_debug_set_attn_timeline(long long* buf);

/* cudaMemsetAsync(ptr, 0, bytes): clears the GroupNorm {sum,sumsq} arena once per U-Net step. */
int dd_zero(void* ptr, int64_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* DDB200_H_ */
