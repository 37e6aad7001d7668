// extern "C" surface of libddif_b200.so: immediate launches, the recorded op list ("plan") that replays one
// UNet forward per call, CUDA-graph capture of a plan, and per-op event profiling.  See include/ddif_b200.h.
#include <vector>

#include "common.cuh"
#include "ddif_internal.h"

namespace ddif {

bool ddif_pdl_enabled() {
#ifdef DDIF_VAR_NO_PDL  // tuning build (tools/): launches without the programmatic-stream-serialization attribute
  return false;
#else
  return true;
#endif
}

union OpParams {
  ddif_gemm_t gemm;
  ddif_in_convert_t in_convert;
  ddif_time_embed_t time_embed;
  ddif_gn_apply_t gn_apply;
  ddif_softmax_h_t softmax_h;
  ddif_attn_t attn;
  ddif_upsample2x_t upsample2x;
  ddif_conv_direct_t conv_direct;
  ddif_stats_t stats;
  ddif_memset_t memset_;
  ddif_resize_t resize;
  ddif_fwm_context_t fwm_context;
  ddif_fwm_weff_t fwm_weff;
  ddif_ddpm_step_t ddpm;
  ddif_ddim_step_t ddim;
  ddif_dpmpp_step_t dpmpp;
  ddif_q_sample_t q_sample;
  ddif_haar_t haar;
  ddif_cond_assemble_t cond_assemble;
  ddif_randn_t randn;
  ddif_axpby_clip_t axpby_clip;
  ddif_dpm_single_t dpm_single;
  ddif_loss_t loss;
  ddif_dpm_err_t dpm_err;
  ddif_attn_block_t attn_block;
  ddif_multi_tensor_t multi_tensor;
  ddif_axpby_t axpby;
  ddif_metrics_t metrics;
  ddif_tile_t tile;
  ddif_wavelet_cond_t wavelet_cond;
  ddif_wgrad_t wgrad;
  ddif_colsum_t colsum;
  ddif_fwm_front_t fwm_front;
};

static size_t params_size(int kind) {
  switch (kind) {
    case DDIF_OP_GEMM: return sizeof(ddif_gemm_t);
    case DDIF_OP_IN_CONVERT: return sizeof(ddif_in_convert_t);
    case DDIF_OP_TIME_EMBED: return sizeof(ddif_time_embed_t);
    case DDIF_OP_GN_APPLY: return sizeof(ddif_gn_apply_t);
    case DDIF_OP_SOFTMAX_H: return sizeof(ddif_softmax_h_t);
    case DDIF_OP_FWM_FRONT: return sizeof(ddif_fwm_front_t);
    case DDIF_OP_ATTN: return sizeof(ddif_attn_t);
    case DDIF_OP_UPSAMPLE2X: return sizeof(ddif_upsample2x_t);
    case DDIF_OP_CONV_DIRECT: return sizeof(ddif_conv_direct_t);
    case DDIF_OP_STATS: return sizeof(ddif_stats_t);
    case DDIF_OP_MEMSET: return sizeof(ddif_memset_t);
    case DDIF_OP_RESIZE: return sizeof(ddif_resize_t);
    case DDIF_OP_FWM_CONTEXT: return sizeof(ddif_fwm_context_t);
    case DDIF_OP_FWM_WEFF: return sizeof(ddif_fwm_weff_t);
    case DDIF_OP_DDPM_STEP: return sizeof(ddif_ddpm_step_t);
    case DDIF_OP_DDIM_STEP: return sizeof(ddif_ddim_step_t);
    case DDIF_OP_DPMPP_STEP: return sizeof(ddif_dpmpp_step_t);
    case DDIF_OP_Q_SAMPLE: return sizeof(ddif_q_sample_t);
    case DDIF_OP_HAAR_DWT2:
    case DDIF_OP_HAAR_IDWT2: return sizeof(ddif_haar_t);
    case DDIF_OP_COND_ASSEMBLE: return sizeof(ddif_cond_assemble_t);
    case DDIF_OP_RANDN: return sizeof(ddif_randn_t);
    case DDIF_OP_AXPBY_CLIP: return sizeof(ddif_axpby_clip_t);
    case DDIF_OP_DPM_SINGLE: return sizeof(ddif_dpm_single_t);
    case DDIF_OP_LOSS: return sizeof(ddif_loss_t);
    case DDIF_OP_DPM_ERR: return sizeof(ddif_dpm_err_t);
    case DDIF_OP_ATTN_BLOCK: return sizeof(ddif_attn_block_t);
    case DDIF_OP_MULTI_TENSOR: return sizeof(ddif_multi_tensor_t);
    case DDIF_OP_AXPBY: return sizeof(ddif_axpby_t);
    case DDIF_OP_METRICS: return sizeof(ddif_metrics_t);
    case DDIF_OP_TILE: return sizeof(ddif_tile_t);
    case DDIF_OP_WAVELET_COND: return sizeof(ddif_wavelet_cond_t);
    case DDIF_OP_WGRAD: return sizeof(ddif_wgrad_t);
    case DDIF_OP_COLSUM: return sizeof(ddif_colsum_t);
    default: return 0;
  }
}

struct Op {
  int kind;
  OpParams p;
  GemmLaunch* gemm;  // owned, only for DDIF_OP_GEMM
};

static int dispatch(const Op& op, cudaStream_t s) {
  switch (op.kind) {
    case DDIF_OP_GEMM: return gemm_launch(*op.gemm, s);
    case DDIF_OP_IN_CONVERT: return launch_in_convert(op.p.in_convert, s);
    case DDIF_OP_TIME_EMBED: return launch_time_embed(op.p.time_embed, s);
    case DDIF_OP_GN_APPLY: return launch_gn_apply(op.p.gn_apply, s);
    case DDIF_OP_SOFTMAX_H: return launch_softmax_h(op.p.softmax_h, s);
    case DDIF_OP_FWM_FRONT: return launch_fwm_front(op.p.fwm_front, s);
    case DDIF_OP_ATTN: return launch_attn(op.p.attn, s);
    case DDIF_OP_UPSAMPLE2X: return launch_upsample2x(op.p.upsample2x, s);
    case DDIF_OP_CONV_DIRECT: return launch_conv_direct(op.p.conv_direct, s);
    case DDIF_OP_STATS: return launch_stats(op.p.stats, s);
    case DDIF_OP_MEMSET: {
      cudaError_t e = cudaMemsetAsync(op.p.memset_.ptr, 0, (size_t)op.p.memset_.bytes, s);
      return e == cudaSuccess ? DDIF_OK : (int)e;
    }
    case DDIF_OP_RESIZE: return launch_resize(op.p.resize, s);
    case DDIF_OP_FWM_CONTEXT: return launch_fwm_context(op.p.fwm_context, s);
    case DDIF_OP_FWM_WEFF: return launch_fwm_weff(op.p.fwm_weff, s);
    case DDIF_OP_DDPM_STEP: return launch_ddpm_step(op.p.ddpm, s);
    case DDIF_OP_DDIM_STEP: return launch_ddim_step(op.p.ddim, s);
    case DDIF_OP_DPMPP_STEP: return launch_dpmpp_step(op.p.dpmpp, s);
    case DDIF_OP_Q_SAMPLE: return launch_q_sample(op.p.q_sample, s);
    case DDIF_OP_HAAR_DWT2: return launch_haar_dwt2(op.p.haar, s);
    case DDIF_OP_HAAR_IDWT2: return launch_haar_idwt2(op.p.haar, s);
    case DDIF_OP_COND_ASSEMBLE: return launch_cond_assemble(op.p.cond_assemble, s);
    case DDIF_OP_RANDN: return launch_randn(op.p.randn, s);
    case DDIF_OP_AXPBY_CLIP: return launch_axpby_clip(op.p.axpby_clip, s);
    case DDIF_OP_DPM_SINGLE: return launch_dpm_single(op.p.dpm_single, s);
    case DDIF_OP_LOSS: return launch_loss(op.p.loss, s);
    case DDIF_OP_DPM_ERR: return launch_dpm_err(op.p.dpm_err, s);
    case DDIF_OP_ATTN_BLOCK: return launch_attn_block(op.p.attn_block, s);
    case DDIF_OP_MULTI_TENSOR: return launch_multi_tensor(op.p.multi_tensor, s);
    case DDIF_OP_AXPBY: return launch_axpby(op.p.axpby, s);
    case DDIF_OP_METRICS: return launch_metrics(op.p.metrics, s);
    case DDIF_OP_TILE: return launch_tile(op.p.tile, s);
    case DDIF_OP_WAVELET_COND: return launch_wavelet_cond(op.p.wavelet_cond, s);
    case DDIF_OP_WGRAD: return launch_wgrad(op.p.wgrad, s);
    case DDIF_OP_COLSUM: return launch_colsum(op.p.colsum, s);
    default: return DDIF_ERR_ARG;
  }
}

static int make_op(int kind, const void* params, Op& op) {
  const size_t sz = params_size(kind);
  if (!sz || !params) return DDIF_ERR_ARG;
  op.kind = kind;
  op.gemm = nullptr;
  memset(&op.p, 0, sizeof(op.p));
  memcpy(&op.p, params, sz);
  if (kind == DDIF_OP_GEMM) {
    op.gemm = new GemmLaunch();
    int rc = gemm_prepare(op.p.gemm, *op.gemm);
    if (rc != DDIF_OK) {
      delete op.gemm;
      op.gemm = nullptr;
      return rc;
    }
  }
  return DDIF_OK;
}

}  // namespace ddif

struct ddif_plan {
  std::vector<ddif::Op> ops;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  // side branch of the captured graph (ddif_plan_set_side_branch): ops [side_first, side_last) run concurrently with the ops that follow them,
  // op join_before is the first one that needs their results
  int side_first = -1, side_last = -1, join_before = -1;
};

extern "C" {

int ddif_version(void) { return 100; }

const char* ddif_error_string(int code) {
  if (code == 0) return "ok";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  switch (code) {
    case DDIF_ERR_ARG: return "ddif: bad argument";
    case DDIF_ERR_SHAPE: return "ddif: unsupported shape / alignment";
    case DDIF_ERR_DRIVER: return "ddif: CUDA driver call failed (cuTensorMapEncodeTiled unavailable or rejected the map)";
    case DDIF_ERR_STATE: return "ddif: bad plan state";
    default: return "ddif: unknown error";
  }
}

int ddif_launch(int kind, const void* params, ddif_stream_t stream) {
  ddif::Op op;
  int rc = ddif::make_op(kind, params, op);
  if (rc != DDIF_OK) return rc;
  rc = ddif::dispatch(op, (cudaStream_t)stream);
  delete op.gemm;
  return rc;
}

ddif_plan_t* ddif_plan_create(void) { return new ddif_plan(); }

void ddif_plan_destroy(ddif_plan_t* plan) {
  if (!plan) return;
  if (plan->exec) cudaGraphExecDestroy(plan->exec);
  if (plan->graph) cudaGraphDestroy(plan->graph);
  for (auto& op : plan->ops) delete op.gemm;
  delete plan;
}

int ddif_plan_add(ddif_plan_t* plan, int kind, const void* params) {
  if (!plan) return DDIF_ERR_ARG;
  if (plan->exec) return DDIF_ERR_STATE;
  ddif::Op op;
  int rc = ddif::make_op(kind, params, op);
  if (rc != DDIF_OK) return rc;
  plan->ops.push_back(op);
  return (int)plan->ops.size() - 1;
}

int ddif_plan_op_variant(const ddif_plan_t* plan, int index) {
  if (!plan || index < 0 || index >= (int)plan->ops.size()) return DDIF_ERR_ARG;
  const ddif::Op& op = plan->ops[index];
  return op.kind == DDIF_OP_GEMM && op.gemm ? op.gemm->variant : -1;
}

int ddif_plan_size(const ddif_plan_t* plan) { return plan ? (int)plan->ops.size() : DDIF_ERR_ARG; }
int ddif_plan_launches(const ddif_plan_t* plan) { return plan ? (int)plan->ops.size() : DDIF_ERR_ARG; }

int ddif_plan_run(ddif_plan_t* plan, int first, int last, ddif_stream_t stream) {
  if (!plan) return DDIF_ERR_ARG;
  const int n = (int)plan->ops.size();
  if (last < 0 || last > n) last = n;
  if (first < 0) first = 0;
  for (int i = first; i < last; ++i) {
    int rc = ddif::dispatch(plan->ops[i], (cudaStream_t)stream);
    if (rc != DDIF_OK) return rc;
  }
  return DDIF_OK;
}

int ddif_plan_graph_build(ddif_plan_t* plan, ddif_stream_t stream) {
  if (!plan) return DDIF_ERR_ARG;
  if (plan->exec) return DDIF_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int n = (int)plan->ops.size();
  const bool fork = plan->side_first >= 0 && plan->side_first < plan->side_last && plan->side_last <= plan->join_before && plan->join_before < n;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  if (fork) {  // created before the capture starts; destroyed after it ends (the graph keeps the dependency edges, not the objects)
    DDIF_CUDA_CHECK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    DDIF_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    DDIF_CUDA_CHECK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
  }
  DDIF_CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  int rc = DDIF_OK;
  if (!fork) {
    rc = ddif_plan_run(plan, 0, -1, stream);
  } else {
    cudaError_t ce = cudaSuccess;
    for (int i = 0; i < n && rc == DDIF_OK && ce == cudaSuccess; ++i) {
      if (i == plan->side_first) {
        ce = cudaEventRecord(ev_fork, s);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(side, ev_fork, 0);
      }
      if (i == plan->join_before && ce == cudaSuccess) ce = cudaStreamWaitEvent(s, ev_join, 0);
      if (ce != cudaSuccess) break;
      const bool on_side = i >= plan->side_first && i < plan->side_last;
      rc = ddif::dispatch(plan->ops[i], on_side ? side : s);
      if (on_side && i == plan->side_last - 1 && rc == DDIF_OK) ce = cudaEventRecord(ev_join, side);
    }
    if (rc == DDIF_OK && ce != cudaSuccess) rc = (int)ce;
  }
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(s, &g);
  if (fork) {
    cudaEventDestroy(ev_fork);
    cudaEventDestroy(ev_join);
    cudaStreamDestroy(side);
  }
  if (rc != DDIF_OK) {
    if (g) cudaGraphDestroy(g);
    return rc;
  }
  if (e != cudaSuccess) return (int)e;
  plan->graph = g;
  DDIF_CUDA_CHECK(cudaGraphInstantiate(&plan->exec, g, 0));
  return DDIF_OK;
}

int ddif_plan_set_side_branch(ddif_plan_t* plan, int first, int last, int join_before) {
  if (!plan) return DDIF_ERR_ARG;
  if (plan->exec) return DDIF_ERR_STATE;
  const int n = (int)plan->ops.size();
  if (first < 0) {  // clear
    plan->side_first = plan->side_last = plan->join_before = -1;
    return DDIF_OK;
  }
  if (first >= last || last > join_before || join_before >= n) return DDIF_ERR_ARG;
  plan->side_first = first; plan->side_last = last; plan->join_before = join_before;
  return DDIF_OK;
}

int ddif_plan_graph_launch(ddif_plan_t* plan, ddif_stream_t stream) {
  if (!plan || !plan->exec) return DDIF_ERR_STATE;
  DDIF_CUDA_CHECK(cudaGraphLaunch(plan->exec, (cudaStream_t)stream));
  return DDIF_OK;
}

int ddif_plan_profile(ddif_plan_t* plan, ddif_stream_t stream, float* ms, int* kinds, int capacity) {
  if (!plan || !ms || !kinds) return DDIF_ERR_ARG;
  const int n = (int)plan->ops.size();
  if (capacity < n) return DDIF_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) DDIF_CUDA_CHECK(cudaEventCreate(&e));
  DDIF_CUDA_CHECK(cudaEventRecord(ev[0], s));
  int rc = DDIF_OK;
  for (int i = 0; i < n && rc == DDIF_OK; ++i) {
    rc = ddif::dispatch(plan->ops[i], s);
    cudaEventRecord(ev[i + 1], s);
  }
  cudaError_t e = cudaStreamSynchronize(s);
  if (rc == DDIF_OK && e != cudaSuccess) rc = (int)e;
  if (rc == DDIF_OK)
    for (int i = 0; i < n; ++i) {
      cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
      kinds[i] = plan->ops[i].kind;
    }
  for (auto& x : ev) cudaEventDestroy(x);
  return rc;
}

int ddif_debug_set_timestamps(void* device_ptr) {
  return ddif::conv3_halo_set_debug_ts((long long*)device_ptr);
}

int ddif_haar_dwt2_f32(const ddif_haar_t* p, ddif_stream_t s) { return ddif_launch(DDIF_OP_HAAR_DWT2, p, s); }
int ddif_haar_idwt2_f32(const ddif_haar_t* p, ddif_stream_t s) { return ddif_launch(DDIF_OP_HAAR_IDWT2, p, s); }
int ddif_cond_assemble_f32(const ddif_cond_assemble_t* p, ddif_stream_t s) { return ddif_launch(DDIF_OP_COND_ASSEMBLE, p, s); }
int ddif_ddpm_step_f32(const ddif_ddpm_step_t* p, ddif_stream_t s) { return ddif_launch(DDIF_OP_DDPM_STEP, p, s); }
int ddif_ddim_step_f32(const ddif_ddim_step_t* p, ddif_stream_t s) { return ddif_launch(DDIF_OP_DDIM_STEP, p, s); }
int ddif_dpmpp_step_f32(const ddif_dpmpp_step_t* p, ddif_stream_t s) { return ddif_launch(DDIF_OP_DPMPP_STEP, p, s); }
int ddif_dpm_single_f32(const ddif_dpm_single_t* p, ddif_stream_t s) { return ddif_launch(DDIF_OP_DPM_SINGLE, p, s); }
int ddif_loss_f32(const ddif_loss_t* p, ddif_stream_t s) { return ddif_launch(DDIF_OP_LOSS, p, s); }
int ddif_dpm_err_f32(const ddif_dpm_err_t* p, ddif_stream_t s) { return ddif_launch(DDIF_OP_DPM_ERR, p, s); }
int ddif_multi_tensor_f32(const ddif_multi_tensor_t* p, ddif_stream_t s) { return ddif_launch(DDIF_OP_MULTI_TENSOR, p, s); }
int ddif_metrics_f32(const ddif_metrics_t* p, ddif_stream_t s) { return ddif_launch(DDIF_OP_METRICS, p, s); }
int ddif_tile_f32(const ddif_tile_t* p, ddif_stream_t s) { return ddif_launch(DDIF_OP_TILE, p, s); }
int ddif_wavelet_cond_f32(const ddif_wavelet_cond_t* p, ddif_stream_t s) { return ddif_launch(DDIF_OP_WAVELET_COND, p, s); }
int ddif_wgrad_bf16(const ddif_wgrad_t* p, ddif_stream_t s) { return ddif_launch(DDIF_OP_WGRAD, p, s); }
int ddif_colsum_bf16(const ddif_colsum_t* p, ddif_stream_t s) { return ddif_launch(DDIF_OP_COLSUM, p, s); }
int ddif_q_sample_f32(const ddif_q_sample_t* p, ddif_stream_t s) { return ddif_launch(DDIF_OP_Q_SAMPLE, p, s); }
int ddif_conv_igemm_bf16(const ddif_gemm_t* p, ddif_stream_t s) { return ddif_launch(DDIF_OP_GEMM, p, s); }

}  // extern "C"
