/*
 * ddif_b200 — C ABI of the B200-native DDIF (Dif-PAN) denoising hot path.
 *
 * The reference (294coder/Dif-PAN) is 100 % Python/PyTorch and has no FFI layer; the entry points below are what
 * a binding for this path replaces.  Each one cites the reference code whose arithmetic it executes.  All
 * functions: plain pointers and sizes, no torch types; device pointers are BORROWED for the call; nothing is
 * allocated on the device by the library (workspace comes from the caller); work is enqueued on the given
 * cudaStream_t (pass torch.cuda.current_stream().cuda_stream) and the device is never synchronised.
 * Return value: 0 = ok, >0 = cudaError_t, <0 = DDIF_ERR_* (argument / shape / driver / state).
 *
 * Activations inside the UNet are NHWC bf16 ("pixels x channels", channel pitch = `*_ld` elements);
 * sampler state, conditioning and UNet inputs/outputs at the boundary are NCHW fp32 like the reference.
 *
 * Every kernel is described by a plain parameter struct whose fields are all 8 bytes wide (pointers, int64_t,
 * double) so that ctypes / cgo / JNI mirrors cannot get the padding wrong.  A struct can be
 *   - launched immediately:            ddif_launch(kind, &params, stream)
 *   - recorded into a plan (op list):  ddif_plan_add(plan, kind, &params)
 * A plan is the per-step schedule of one UNet forward (~280 ops); ddif_plan_run replays it with one C call and
 * ddif_plan_graph_launch with one CUDA-graph launch.
 */
#ifndef DDIF_B200_H
#define DDIF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ddif_stream_t; /* cudaStream_t */
typedef struct ddif_plan ddif_plan_t;

enum ddif_op_kind {
  DDIF_OP_GEMM = 1,          /* tcgen05 implicit-GEMM conv (3x3 / 1x1, stride 1|2, <=2 K-segments, fused epilogue) */
  DDIF_OP_IN_CONVERT = 2,    /* NCHW fp32 x (+self_cond) -> NHWC bf16 concat                       sr3_dwt.py:172-174 */
  DDIF_OP_TIME_EMBED = 3,    /* PositionalEncoding + noise_level_mlp + all FiLM Linears           sr3_dwt.py:57-64,223-258 */
  DDIF_OP_GN_APPLY = 4,      /* GroupNorm(1 group) affine (+Swish) (+depthwise 3x3), 1 or 2 sources sr3_dwt.py:292-294,507-512 */
  DDIF_OP_SOFTMAX_H = 5,     /* q.softmax(dim=-2) * scale                                          sr3_dwt.py:545,561 */
  DDIF_OP_ATTN = 6,          /* fused multi-head self-attention core                               sr3_dwt.py:347-357 */
  DDIF_OP_UPSAMPLE2X = 7,    /* nearest x2                                                         sr3_dwt.py:269 */
  DDIF_OP_CONV_DIRECT = 8,   /* CUDA-core direct conv (checker for the tcgen05 path; tiny layers)  */
  DDIF_OP_STATS = 9,         /* per-sample sum / sum of squares of an NHWC bf16 tensor             */
  DDIF_OP_MEMSET = 10,
  DDIF_OP_RESIZE = 11,       /* F.interpolate(cond, size, 'bilinear')                              sr3_dwt.py:661-663 */
  DDIF_OP_FWM_CONTEXT = 12,  /* kv = 1x1(dw3x3(c)); k.softmax(W); context = k v^T                  sr3_dwt.py:541,546,563 */
  DDIF_OP_FWM_WEFF = 13,     /* W_eff[b] = scale * attn_out.weight @ blockdiag(context[b]^T)       sr3_dwt.py:561-573 */
  DDIF_OP_DDPM_STEP = 14,    /* p_mean_variance + p_sample                                         diffusion_ddpm_pan.py:346-442 */
  DDIF_OP_DDIM_STEP = 15,    /* ddim_sample                                                        diffusion_ddpm_pan.py:594-621 */
  DDIF_OP_DPMPP_STEP = 16,   /* model_wrapper + data_prediction_fn + multistep update (orders 1-3)  dpm_solver.py:286-300,441-450,555-588,804-912 */
  DDIF_OP_Q_SAMPLE = 17,     /* q_sample                                                           diffusion_ddpm_pan.py:668-681 */
  DDIF_OP_HAAR_DWT2 = 18,    /* pywt.wavedec2(x,'db1',level=1)                                     dataset/pan_dataset.py:75-80 */
  DDIF_OP_HAAR_IDWT2 = 19,   /* inverse of the above (no reference call site)                      */
  DDIF_OP_COND_ASSEMBLE = 20,/* cond = cat[lms, pan, bilinear(wavelets)]                           diffusion_engine.py:221-228 */
  DDIF_OP_RANDN = 21,        /* Philox4x32-10 + Box-Muller standard normal (replaces torch.randn)  diffusion_ddpm_pan.py:86,484 */
  DDIF_OP_AXPBY_CLIP = 22,   /* sr = clip(sample + lms, 0, 1)                                      diffusion_engine.py:446-447 */
  DDIF_OP_DPM_SINGLE = 23,   /* one stage of the singlestep DPM-Solver / DPM-Solver++ updates      dpm_solver.py:555-600,602-802 */
  DDIF_OP_LOSS = 24,         /* sum_b w[b] * sum |a-b| or (a-b)^2 (training objective, forward)    diffusion_ddpm_pan.py:725-762 */
  DDIF_OP_AXPBY = 25,        /* out = ca[b]*x + cb[b]*y per sample (predict_start_from_noise / _v)  diffusion_ddpm_pan.py:284-312 */
  DDIF_OP_METRICS = 26,      /* per-image partial sums of SAM / ERGAS / PSNR / CC                  utils/_metric_legacy.py:299-346 */
  DDIF_OP_TILE = 27,         /* scene -> patch batch (gather) and patch batch -> scene (overlap-averaged stitch) */
  DDIF_OP_ATTN_BLOCK = 30,   /* whole SelfAttention block at 64 tokens: GN + qkv + attention + out + residual   sr3_dwt.py:330-360 */
  DDIF_OP_MULTI_TENSOR = 31, /* one launch over a LIST of tensors: EMA, copy, sum of squares, scale, clamp, AdamW   utils/optim_utils.py:24-58, utils/misc.py:25-36 */
  DDIF_OP_WGRAD = 32,        /* convolution weight gradient: dW[tap][o][i] += sum_p dY[p][o] X[p + tap][i]  (autograd of F.conv2d in p_losses().backward(),
                                diffusion_ddpm_pan.py:692-766, diffusion_engine.py:233) */
  DDIF_OP_COLSUM = 33,       /* per-channel sum of a gradient tensor over all pixels (bias gradient) or per sample (FiLM gradient, sr3_dwt.py:241-258) */
  DDIF_OP_FWM_FRONT = 34,    /* FWM front at 8 x 8 in one kernel: prenorm_x + DW3x3 + q 1x1 + softmax over H + attn_out (W_eff) + attn_res + bias   sr3_dwt.py:507-573 */
  DDIF_OP_DPM_ERR = 29,      /* adaptive DPM-Solver error estimate per sample                      dpm_solver.py:1003-1006 */
  DDIF_OP_WAVELET_COND = 28  /* raw lms, pan -> cond in one pass: Haar DWT, /division, channel order, bilinear up, concat
                                dataset/pan_dataset.py:73-142, dataset/hisr.py:48-59, diffusion_engine.py:221-228 */
};

/* ---- DDIF_OP_GEMM ------------------------------------------------------------------------------------------
 * out[b,y,x,n] = epilogue( sum_seg sum_tap sum_c  a_seg[b, y*stride+dy-pad, x*stride+dx-pad, c] * w_seg[z, n, c] )
 * z = tap (shared weights) or b (per-sample weights, 1x1 only).  Zero padding comes from TMA out-of-bounds fill.
 * Three kernels implement it: conv3x3_halo_tc_kernel (3x3, stride 1: ONE TMA halo tile feeds all nine taps, optional
 * fused GroupNorm(+Swish) prologue, 1-2 concatenated sources, N split over CTAs), cs_gemm_tc_kernel (1x1 with a_softmax_h: the
 * softmax over the image height is computed on the loaded tile) and conv_igemm_tc_kernel (everything else: 1x1, stride 2,
 * per-sample weights).
 * epilogue: v = acc + bias[n] + film[b*film_ld + n];  v = v*(1+mod[..,n]) + mod[..,n_valid+n];  v += residual;
 *           v = silu(v) if act;  stats[b] += (sum v, sum v^2);  store bf16 NHWC and/or fp32 NCHW.              */
typedef struct {
  const void* a[2];     int64_t a_ld[2];  int64_t a_c[2];  int64_t a_h[2];  int64_t a_w[2];
  const void* w[2];     int64_t w_s[2];   int64_t w_k[2];  int64_t taps[2]; int64_t w_per_sample[2];
  int64_t nseg, stride, batch, out_h, out_w, n_pad, n_valid;
  const float* bias;
  const float* film;    int64_t film_ld;
  const void* mod;
  const void* residual; int64_t res_ld;
  int64_t act;
  void* out;            int64_t out_ld;
  float* out_nchw;
  double* stats;
  /* Optional fused prologue on segment 0 (3x3 stride-1 convs only): a = act(GroupNorm_1group(a)) with the per-sample
   * (sum, sumsq) statistics of the input tensor; a_up is reserved and must be 0 (nearest x2, sr3_dwt.py:269, is
   * DDIF_OP_UPSAMPLE2X in front of the conv).  force_tma = 1 selects the generic TMA kernel even where the halo kernel applies. */
  const double* gn_stats; const float* gn_gamma; const float* gn_beta; double gn_eps;
  int64_t gn_act, a_up, force_tma;
  /* 3x3 stride-1 convs may take TWO segments = channel concat of two tensors (torch.cat((x, skip), 1), sr3_dwt.py:212);
   * the fused GroupNorm then runs over the concatenation: gn_stats2 = (sum, sumsq) of segment 1, gamma/beta in concat
   * order.  Used for FWM q = Conv1x1(DW3x3(GN(cat))) with the two convs composed into one dense 3x3 (sr3_dwt.py:507-541). */
  const double* gn_stats2;
  /* Depthwise mode of the fused 3x3 kernel (FWM q path, sr3_dwt.py:509-517,541,573): dw_w != NULL -> the (normalised) input goes
   * through a depthwise 3x3 with weights dw_w[9][sum a_c] (fp32, tap = ky*3+kx, zero padding of the normalised tensor) INSIDE the
   * kernel and `w` holds 1x1 weights ([1][n_pad][K], w_s = 1): output channels [0, dw_n) = w[0:dw_n] . dw3x3(a), channels
   * [dw_n, n_valid) = w[dw_n:] . a (the un-convolved normalised input: attn_res(x_hat)).  Needs gn_stats. */
  const float* dw_w; int64_t dw_n;
  /* a_softmax_h != 0 (one segment, 1x1, stride 1; shared or per-sample weights; epilogue = bias + optional residual): segment 0 is read
   * through a softmax over the image HEIGHT, a'[b,y,x,c] = exp(a[b,y,x,c]) / sum_y' exp(a[b,y',x,c]) -- FWM `q.softmax(dim=-2)` followed by
   * attn_out (sr3_dwt.py:541-573) in ONE kernel (cs_gemm_tc_kernel: a tile = 128/H image columns x all H lines).  Needs out_h in
   * {16, 32, 64}, out_w % (128 / out_h) == 0, a_c % 32 == 0, n_pad <= 128; DDIF_ERR_SHAPE otherwise (run DDIF_OP_SOFTMAX_H + a plain GEMM). */
  int64_t a_softmax_h;
} ddif_gemm_t;

typedef struct { const float* x; const float* self_cond; void* out; int64_t batch, c, h, w, c_pad; } ddif_in_convert_t;

typedef struct {
  const float* time; const float* w1; const float* b1; const float* w2; const float* b2;
  const float* wf; const float* bf; float* film; int64_t batch, inner, nfilm;
} ddif_time_embed_t;

/* y = act( (x - mean_b) * rstd_b * gamma[c] + beta[c] ), x = concat(src1[c1], src2[c2]) along channels.
 * stats1/stats2: per-sample (sum, sumsq) of src1/src2 over count1/count2 elements.
 * If dw_w != NULL also out_dw = depthwise3x3(y) with dw_w laid out [9][c1+c2] fp32 (zero padding of y). */
typedef struct {
  const void* src1; int64_t c1; const void* src2; int64_t c2;
  const double* stats1; const double* stats2;
  const float* gamma; const float* beta;
  void* out; const float* dw_w; void* out_dw;
  int64_t batch, h, w, act; double eps;
} ddif_gn_apply_t;

/* in: [B, h, w, in_ld] (first c channels of every pixel; in_ld = 0 means c), out: [B, h, w, c] */
typedef struct { const void* in; void* out; int64_t batch, h, w, c; double scale; int64_t in_ld; } ddif_softmax_h_t;
/* qkv: [B, ntok, 3*C] with per-head channel blocks [q(hd) k(hd) v(hd)]; out: [B, ntok, C] */
typedef struct { const void* qkv; void* out; int64_t batch, ntok, c, heads; double scale; } ddif_attn_t;
/* Fused SelfAttention block (sr3_dwt.py:330-360) for ntok = 64, c = 128, heads = 8 (what the UNet attends at): x [B, 64, c] bf16 is the block
 * input AND the residual; stats_in = its per-sample (sum, sumsq) in fp64; wqkv = the packed qkv 1x1 weight [3c][c] bf16; wout = the out 1x1
 * weight [c][c] bf16 with its K axis permuted for 16-byte fragment chunks (position 32(ks>>1) + 8t + 4(ks&1) + 2h + e holds input channel
 * 16ks + 8h + 2t + e, t < 4, ks < 8, h < 2, e < 2; csrc/attn_block.cu); out [B, 64, c] bf16; stats_out (optional) += (sum, sumsq) of out per sample. */
typedef struct {
  const void* x; const double* stats_in; const float* gamma; const float* beta; const void* wqkv; const void* wout; const float* bout;
  void* out; double* stats_out; int64_t batch, ntok, c, heads; double scale, eps;
} ddif_attn_block_t;
/* Fused front of FastAttnCondInjection for 8 x 8 images (csrc/fwm_front.cu; one CTA per sample): x [B, 8, 8, c1] and skip [B, 8, 8, c2] bf16 NHWC
 * (their channel concatenation is the block input, c1 + c2 = dim in {192, 256}), stats1 / stats2 = their per-sample (sum, sumsq) in fp64, gamma / beta
 * [dim] of prenorm_x; dw_w [9][dim] fp32 (q.0, tap = ky*3+kx); w1 [dim][w1_ld] bf16 + b1 [dim] (q.1); weff [B][weff_rows][weff_ld] bf16 = the
 * per-sample W_eff of the cond cache (DDIF_OP_FWM_WEFF); wres [o][wres_ld] bf16 (attn_res); bias [o] = attn_out.bias + attn_res.bias;
 * out [B, 8, 8, out_ld] bf16 (first o = 128 channels):  out = W_eff[b] . softmax_H(w1 . dw3x3(x_hat) + b1) + wres . x_hat + bias. */
typedef struct {
  const void* x; const void* skip; int64_t c1, c2;
  const double* stats1; const double* stats2; const float* gamma; const float* beta; double eps;
  const float* dw_w; const void* w1; int64_t w1_ld; const float* b1;
  const void* weff; int64_t weff_ld, weff_rows; const void* wres; int64_t wres_ld; const float* bias;
  void* out; int64_t out_ld, batch, h, w, o;
} ddif_fwm_front_t;
typedef struct { const void* in; void* out; int64_t batch, h, w, c; } ddif_upsample2x_t;
typedef struct {
  const void* in; int64_t in_ld, cin, in_h, in_w; const void* w; int64_t w_k, taps, stride;
  const float* bias; void* out; int64_t out_ld, batch, out_h, out_w, n_valid, n_pad; int64_t act;
} ddif_conv_direct_t;
typedef struct { const void* in; double* stats; int64_t batch, hw, c; } ddif_stats_t;
typedef struct { void* ptr; int64_t bytes; } ddif_memset_t;
/* src: fp32 NCHW [B, c_total, h, w]; channels [c0, c0+c) resized to out_h x out_w into a bf16 NHWC tensor with
 * c_pad channels (zero padded) and/or an fp32 NCHW tensor. */
typedef struct {
  const float* src; int64_t batch, c_total, c0, c, h, w, out_h, out_w;
  void* dst_nhwc; int64_t c_pad; float* dst_nchw;
} ddif_resize_t;
typedef struct {
  const float* c_dec; const float* kv0_w; const float* kv1_w; const float* kv1_b; float* ctx;
  int64_t batch, h, w, cd, dim, heads;
} ddif_fwm_context_t;
typedef struct {
  const float* ctx; const float* w_out; void* weff; int64_t batch, o, dim, heads, o_pad, k_pad; double scale;
} ddif_fwm_weff_t;

/* x0 = model_out (x_start) | sra*x - srm1*out (noise) | sa*x - s1ma*out (pred_v); clamp(x0+lms)-lms; posterior.
 * coef: fp32 device table [T][8] = {c1, c2, logvar, sqrt_recip_ac, sqrt_recipm1_ac, sqrt_ac, sqrt_1m_ac, 0}.
 * t: timestep index (same for the whole batch, like the reference loop).  noise: injected tensor or NULL
 * (then Philox(seed, offset)).  lms is read from `cond` (first c channels, cond_c channel count).
 * time_out (optional): float[batch] receives t-1 for the next UNet call. */
typedef struct {
  float* x; const float* model_out; const float* cond; const float* noise; const float* coef; float* time_out;
  int64_t batch, c, hw, cond_c, t, pred_mode, clip; double clamp_lo, clamp_hi; int64_t seed, offset;
} ddif_ddpm_step_t;
/* coef: [T][8] = {alphas_cumprod, alphas_cumprod_prev, sqrt_recip_ac, sqrt_recipm1_ac, sqrt_ac, sqrt_1m_ac, 0, 0} */
typedef struct {
  float* x; const float* model_out; const float* cond; const float* noise; const float* coef; float* time_out;
  int64_t batch, c, hw, cond_c, t, pred_mode, clip; double clamp_lo, clamp_hi, eta; int64_t seed, offset;
} ddif_ddim_step_t;
/* m0 = data prediction from the UNet output through the reference's x_start->noise->x_start round trip
 * (noise = (x - alpha_t*out)/sigma_t; m0 = (x - sigma_t*noise)/alpha_t, alpha/sigma of the CURRENT time), stored
 * to m_cur.  Then the multistep update to the NEXT time, with host-computed fp32 scalars and the reference's
 * operation order:
 *   order 0: no update (last evaluation only stores m_cur)
 *   order 1: x = cx*x - ca*m0                                                         (dpm_solver.py:581-584)
 *   order 2: x = cx*x - ca*m0 - cb*(inv_r0*(m0-m1))              cb = 0.5*ca           (dpm_solver.py:831-839)
 *   order 3: D10 = inv_r0*(m0-m1); D11 = inv_r1*(m1-m2); D1 = D10 + k1*(D10-D11); D2 = k2*(D10-D11);
 *            x = cx*x - ca*m0 + cb*D1 - cc*D2                                          (dpm_solver.py:888-901)
 * time_out (optional): float[batch] receives t_next_in (the UNet time label of the next evaluation). */
typedef struct {
  float* x; const float* model_out; float* m_cur; const float* m_prev1; const float* m_prev2; float* time_out;
  int64_t n, batch, order, model_type; /* model_type: 0 x_start, 1 noise, 2 v (dpm_solver.py:296-303) */
  double alpha_t, sigma_t, cx, ca, cb, cc, inv_r0, inv_r1, k1, k2, t_next_in;
  int64_t predict; /* 0: m = data prediction (algorithm_type dpmsolver++), 1: m = noise prediction (dpmsolver; the host passes
                      that algorithm's scalars in cx..cc with the signs of the ++ expression: dpm_solver.py:589-599,840-845,902-912) */
} ddif_dpmpp_step_t;
typedef struct { const float* x0; const float* noise; float* out; const float* sa; const float* s1ma; const int64_t* t; int64_t batch, chw; } ddif_q_sample_t;
/* One stage of a singlestep solver (DPM-Solver-1/2/3, both algorithm types).  m_cur = prediction at the evaluation point
 * (x_eval, alpha_e, sigma_e) from the denoiser output, through the same model_wrapper round trip as DDIF_OP_DPMPP_STEP:
 * predict 0 = data prediction (dpmsolver++), 1 = noise prediction (dpmsolver).  Then, with host-computed fp32 scalars,
 *   mode 0: x_out = c0*x_base - c1*m_cur                          (first update; stage x_s1)       dpm_solver.py:581-599,639-642
 *   mode 1: x_out = c0*x_base - c1*m_a + c2*(m_cur - m_a)         (2nd/3rd-order final; stage x_s2) dpm_solver.py:645-649,745-757
 *   mode 2: only m_cur is stored
 * x_out may alias x_base or x_eval.  time_out (optional): float[batch] receives t_next_in. */
typedef struct {
  const float* x_base; const float* x_eval; const float* model_out; float* m_cur; const float* m_a; float* x_out; float* time_out;
  int64_t n, batch, model_type, predict, mode;
  double alpha_e, sigma_e, c0, c1, c2, t_next_in;
} ddif_dpm_single_t;
/* Multi-tensor apply (the training loop's EMA / grad-clip / optimizer update touch ~350 small parameter tensors; the
 * reference issues 3-10 ATen kernels per tensor).  ptrs: device int64[ntensors][4] (addresses of up to four fp32 tensors per entry), sizes:
 * device int64[ntensors], chunks: device int64[nchunks][2] = (tensor index, first element); every chunk covers <= `chunk` elements.
 *   op 0 EMA    p0 = s0*p0 + s1*p1                       EmaUpdater.update, utils/optim_utils.py:44-52 (s0 = decay, s1 = 1 - decay)
 *   op 1 COPY   p0 = p1                                   :53-58
 *   op 2 SUMSQ  out[0] += sum p0^2                        clip_grad_norm_, utils/misc.py:33-34
 *   op 3 SCALE  p0 *= s0
 *   op 4 CLAMP  p0 = min(max(p0, -s0), s0)                clip_grad_value_, utils/misc.py:35-36
 *   op 5 ADAMW  torch.optim.AdamW step (p0 param, p1 grad, p2 exp_avg, p3 exp_avg_sq; s0 lr, s1 beta1, s2 beta2, s3 eps, s4 weight_decay,
 *               s5 = 1 - beta1^step, s6 = 1 - beta2^step), diffusion_engine.py:202-241 */
typedef struct { const int64_t* ptrs; const int64_t* sizes; const int64_t* chunks; double* out; int64_t nchunks, chunk, op; double s0, s1, s2, s3, s4, s5, s6; } ddif_multi_tensor_t;
/* Adaptive step-size control (dpm_solver.py:1003-1006): out[b] = sum_i ((x_higher - x_lower) / max(atol, rtol*max(|x_lower|, |x_prev|)))^2 for
 * sample b (every entry written); the host takes E = max_b sqrt(out[b] / chw). */
typedef struct { const float* x_higher; const float* x_lower; const float* x_prev; double* out; int64_t batch, chw; double atol, rtol; } ddif_dpm_err_t;
/* out[0] += sum_b weight[b] * sum_i l(a[b,i], b[b,i]); squared 0: |a-b| (l1), 1: (a-b)^2 (l2).  weight may be NULL (= 1).
 * `out` is a device double the caller zeroes; the mean is out[0] / (batch*chw). */
typedef struct { const float* a; const float* b; const float* weight; double* out; int64_t batch, chw, squared; } ddif_loss_t;
/* out[b,i] = ca[b]*x[b,i] + cb[b]*y[b,i]  (ca, cb: float[batch] on the device) */
typedef struct { const float* x; const float* y; const float* ca; const float* cb; float* out; int64_t batch, chw; } ddif_axpby_t;
/* gt, out: [batch, c, h, w] fp32.  Only pixels y < h-1, x < w-1 enter (the reference's bounds cut `[0:-1]`, :300-302).
 * sums: double[batch][2 + 6*c] (every entry is written; no pre-zeroing needed): {sum of spectral angles, pixels with |a||b| > 0,
 * per band: sum (a-b)^2, sum a, sum b, sum a^2, sum b^2, sum a*b}  with a = gt, b = out. */
typedef struct { const float* gt; const float* out; double* sums; int64_t batch, c, h, w; } ddif_metrics_t;
/* dir 0: tiles[(b*ny + iy)*nx + ix, c, :, :] = scene[b, c, iy*sy : iy*sy+ph, ix*sx : ix*sx+pw]
 * dir 1: scene[b, c, y, x] = mean over the tiles covering (y, x) (uniform average of overlaps; sy <= ph, sx <= pw) */
typedef struct { float* scene; float* tiles; int64_t batch, c, h, w, ph, pw, sy, sx, ny, nx, dir; } ddif_tile_t;
/* cond[b] = cat(lms/div, pan/div, up2(LL(lms)/div), up2(pan sub-bands / div))  in ONE pass from the raw arrays.
 * lms [batch, c, h, w], pan [batch, p, h, w] raw values; order 0 = Pan datasets (cH, cD, cV; pan_dataset.py:139-141),
 * 1 = HISR (cH, cV, cD; hisr.py:57-59); up2 = bilinear x2, align_corners=False (diffusion_engine.py:224-226).
 * wav (optional): [batch, c + 3p, h/2, w/2] also receives the wavelet stack the reference datasets return.  h even, w % 4 == 0. */
typedef struct { const float* lms; const float* pan; float* cond; float* wav; int64_t batch, c, p, h, w, order; double divisor; } ddif_wavelet_cond_t;
/* x: [planes, h, w] fp32 (output for IDWT); 4 sub-bands each [planes, h/2, w/2] (LL, cH, cV, cD).
 * DWT: every coefficient is divided by `divisor` (the dataset "division", pan_dataset.py:127-134); IDWT ignores it. */
typedef struct { float* x; float* ll; float* ch; float* cv; float* cd; int64_t planes, h, w; double divisor; } ddif_haar_t;
/* cond[b] = cat(lms[b] (c), pan[b] (p), bilinear_up(wavelets[b] (cw), h x w)); wavelets at h/2 x w/2 (any size) */
typedef struct { const float* lms; const float* pan; const float* wav; float* cond; int64_t batch, c, p, cw, h, w, wh, ww; } ddif_cond_assemble_t;
typedef struct { float* out; int64_t n, seed, offset; } ddif_randn_t;
/* Weight gradient of a convolution (3x3 pad 1 | 1x1; stride 1 | 2) from NHWC bf16 tensors: x [batch, in_h, in_w, x_ld] (first cin channels),
 * dy [batch, out_h, out_w, dy_ld] (first cout channels); dw: fp32 [taps][cout][cin], ACCUMULATED (the caller zeroes it), tap = ky*3+kx.
 * per_sample = 1 (1x1 only): dw is [batch][cout][cin], one gradient per sample (the FWM per-sample attn_out weights W_eff). */
typedef struct {
  const void* x; int64_t x_ld, cin; const void* dy; int64_t dy_ld, cout; float* dw;
  int64_t batch, in_h, in_w, out_h, out_w, taps, stride, per_sample;
} ddif_wgrad_t;
/* out[g][c] += sum over the pixels of group g of dy[p][c]; dy NHWC bf16 [batch, hw, ld]; per_sample = 0: one group (out[c]); 1: out[batch][c]. */
typedef struct { const void* dy; float* out; int64_t batch, hw, c, ld, per_sample; } ddif_colsum_t;
typedef struct { const float* x; const float* cond; float* out; int64_t batch, c, hw, cond_c; double lo, hi; } ddif_axpby_clip_t;

/* ---- entry points ---------------------------------------------------------------------------------------- */
int ddif_version(void);
const char* ddif_error_string(int code);
/* Launch one op now on `stream`. */
int ddif_launch(int kind, const void* params, ddif_stream_t stream);

ddif_plan_t* ddif_plan_create(void);
void ddif_plan_destroy(ddif_plan_t* plan);
/* Record an op (parameters are copied; GEMM tensor maps are encoded once here). Returns op index or <0. */
int ddif_plan_add(ddif_plan_t* plan, int kind, const void* params);
int ddif_plan_size(const ddif_plan_t* plan);
/* Which kernel a recorded DDIF_OP_GEMM resolved to: 0 conv_igemm_tc_kernel, 2 conv3x3_halo_tc_kernel, 3 cs_gemm_tc_kernel; -1 for every other op kind
 * (used by bench.py to attribute time per kernel). */
int ddif_plan_op_variant(const ddif_plan_t* plan, int index);
/* Enqueue ops [first, last) on `stream` (last<0 = all). */
int ddif_plan_run(ddif_plan_t* plan, int first, int last, ddif_stream_t stream);
/* Capture the whole plan into a CUDA graph once, then replay it. */
int ddif_plan_graph_build(ddif_plan_t* plan, ddif_stream_t stream);
/* Before ddif_plan_graph_build: ops [first, last) form a SIDE BRANCH of the captured graph -- they depend on everything before `first`, run
 * concurrently with the ops after them, and op `join_before` (> last - 1) is the first that waits for them.  The caller guarantees that no op in
 * [last, join_before) reads or writes what the branch writes.  Used for the time embedding + FiLM vectors (sr3_dwt.py:57-64,223-257), a
 * latency-bound launch whose first consumer is the FiLM add of the first ResBlock.  first < 0 clears.  ddif_plan_run ignores it (one stream). */
int ddif_plan_set_side_branch(ddif_plan_t* plan, int first, int last, int join_before);
int ddif_plan_graph_launch(ddif_plan_t* plan, ddif_stream_t stream);
/* Profiling: run the plan with a CUDA event pair around every op; ms[i] = duration of op i, kinds[i] = its kind. */
int ddif_plan_profile(ddif_plan_t* plan, ddif_stream_t stream, float* ms, int* kinds, int capacity);
/* Number of kernel launches the plan issues per run. */
int ddif_plan_launches(const ddif_plan_t* plan);

/* Debug: device buffer of 4*64*4 int64 receiving clock64() at the pipeline hand-offs of CTA 0 of the 3x3 conv kernels
 * (role-major: loader / TMA producer, MMA issuer, epilogue, GN transform; 64 tiles; 4 stamps).  NULL disables.
 * Not for production use. */
int ddif_debug_set_timestamps(void* device_ptr);

/* Named convenience wrappers (same structs), the symbols a reference-side binding would call directly. */
int ddif_haar_dwt2_f32(const ddif_haar_t* p, ddif_stream_t s);
int ddif_haar_idwt2_f32(const ddif_haar_t* p, ddif_stream_t s);
int ddif_cond_assemble_f32(const ddif_cond_assemble_t* p, ddif_stream_t s);
int ddif_ddpm_step_f32(const ddif_ddpm_step_t* p, ddif_stream_t s);
int ddif_ddim_step_f32(const ddif_ddim_step_t* p, ddif_stream_t s);
int ddif_dpmpp_step_f32(const ddif_dpmpp_step_t* p, ddif_stream_t s);
int ddif_q_sample_f32(const ddif_q_sample_t* p, ddif_stream_t s);
int ddif_conv_igemm_bf16(const ddif_gemm_t* p, ddif_stream_t s);
int ddif_dpm_single_f32(const ddif_dpm_single_t* p, ddif_stream_t s);
int ddif_dpm_err_f32(const ddif_dpm_err_t* p, ddif_stream_t s);
int ddif_multi_tensor_f32(const ddif_multi_tensor_t* p, ddif_stream_t s);
int ddif_loss_f32(const ddif_loss_t* p, ddif_stream_t s);
int ddif_metrics_f32(const ddif_metrics_t* p, ddif_stream_t s);
int ddif_tile_f32(const ddif_tile_t* p, ddif_stream_t s);
int ddif_wavelet_cond_f32(const ddif_wavelet_cond_t* p, ddif_stream_t s);
int ddif_wgrad_bf16(const ddif_wgrad_t* p, ddif_stream_t s);
int ddif_colsum_bf16(const ddif_colsum_t* p, ddif_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* DDIF_B200_H */
