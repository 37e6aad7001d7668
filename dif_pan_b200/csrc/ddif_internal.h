// Internal declarations shared by the translation units of libddif_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/ddif_b200.h"

namespace ddif {

// A GEMM op resolved to kernel parameters (tensor maps encoded) + launch geometry.
struct alignas(64) GemmLaunch {
  unsigned char kparams[1536];
  int grid_x, grid_y, smem_bytes;
  int flags;    // variant 2: compile-time epilogue specialisation (kEpi* bits)
  int variant;  // 3: cs_gemm_tc_kernel (column softmax fused into a 1x1 GEMM), 0: conv_igemm_tc_kernel (TMA tap boxes), 2: conv3x3_halo_tc_kernel (one TMA halo tile feeds all nine taps; in-place
                // GN+Swish).  (1 was an LDG-fed 3x3 kernel with the nearest-x2 folded into its loader: slower at every level, removed.)
};

int gemm_prepare(const ddif_gemm_t& g, GemmLaunch& L);
int gemm_launch(const GemmLaunch& L, cudaStream_t stream);
bool conv3_halo_applicable(const ddif_gemm_t& g);
int conv3_halo_prepare(const ddif_gemm_t& g, GemmLaunch& L);
int conv3_halo_launch(const GemmLaunch& L, cudaStream_t stream);
int conv3_halo_set_debug_ts(long long* ptr);
bool cs_gemm_applicable(const ddif_gemm_t& g);           // conv3x3_halo.cu: column-softmax GEMM (FWM softmax over H fused into attn_out)
int cs_gemm_prepare(const ddif_gemm_t& g, GemmLaunch& L);
int cs_gemm_launch(const GemmLaunch& L, cudaStream_t stream);

// elementwise.cu
int launch_in_convert(const ddif_in_convert_t& p, cudaStream_t s);
int launch_time_embed(const ddif_time_embed_t& p, cudaStream_t s);
int launch_gn_apply(const ddif_gn_apply_t& p, cudaStream_t s);
int launch_softmax_h(const ddif_softmax_h_t& p, cudaStream_t s);
int launch_attn(const ddif_attn_t& p, cudaStream_t s);
int launch_upsample2x(const ddif_upsample2x_t& p, cudaStream_t s);
int launch_conv_direct(const ddif_conv_direct_t& p, cudaStream_t s);
int launch_stats(const ddif_stats_t& p, cudaStream_t s);
int launch_resize(const ddif_resize_t& p, cudaStream_t s);
int launch_fwm_context(const ddif_fwm_context_t& p, cudaStream_t s);
int launch_fwm_weff(const ddif_fwm_weff_t& p, cudaStream_t s);

// sampler.cu
int launch_ddpm_step(const ddif_ddpm_step_t& p, cudaStream_t s);
int launch_ddim_step(const ddif_ddim_step_t& p, cudaStream_t s);
int launch_dpmpp_step(const ddif_dpmpp_step_t& p, cudaStream_t s);
int launch_q_sample(const ddif_q_sample_t& p, cudaStream_t s);
int launch_haar_dwt2(const ddif_haar_t& p, cudaStream_t s);
int launch_haar_idwt2(const ddif_haar_t& p, cudaStream_t s);
int launch_cond_assemble(const ddif_cond_assemble_t& p, cudaStream_t s);
int launch_randn(const ddif_randn_t& p, cudaStream_t s);
int launch_axpby_clip(const ddif_axpby_clip_t& p, cudaStream_t s);
int launch_dpm_single(const ddif_dpm_single_t& p, cudaStream_t s);
int launch_loss(const ddif_loss_t& p, cudaStream_t s);
int launch_dpm_err(const ddif_dpm_err_t& p, cudaStream_t s);
int launch_attn_block(const ddif_attn_block_t& p, cudaStream_t s);
int launch_fwm_front(const ddif_fwm_front_t& p, cudaStream_t s);  // fwm_front.cu
bool attn_tc_applicable(const ddif_attn_t& p);            // attn_tc.cu: head_dim 16, ntok % 128 == 0
int launch_attn_tc(const ddif_attn_t& p, cudaStream_t s);
int launch_multi_tensor(const ddif_multi_tensor_t& p, cudaStream_t s);
int launch_axpby(const ddif_axpby_t& p, cudaStream_t s);
int launch_metrics(const ddif_metrics_t& p, cudaStream_t s);
int launch_tile(const ddif_tile_t& p, cudaStream_t s);
int launch_wavelet_cond(const ddif_wavelet_cond_t& p, cudaStream_t s);
int launch_wgrad(const ddif_wgrad_t& p, cudaStream_t s);    // backward.cu
int launch_colsum(const ddif_colsum_t& p, cudaStream_t s);

}  // namespace ddif
