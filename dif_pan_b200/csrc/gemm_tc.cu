// tcgen05 implicit-GEMM convolution for the SR3-DWT UNet (3x3 / 1x1, stride 1|2, NHWC bf16, fp32 accumulate).
//
// Replaces every dense F.conv2d of /root/reference/models/sr3_dwt.py (Block :296, ResnetBlock :318-320,
// CondInjection :380-385, FastAttnCondInjection :512-532, SelfAttention :338-339, Down/Upsample :270,279).
//
// GEMM view: M = output pixels (128 per CTA: a TW x TH x TN pixel box), N = Cout (<=256 per CTA), K = taps*Cin.
// One K-iteration = one (tap, 16|32|64-channel chunk): TMA loads the tap-shifted activation box (zero padding =
// TMA out-of-bounds fill; stride 2 = TMA element strides) and the matching weight slab into a SWIZZLE_32/64/128B
// K-major stage; one elected thread issues tcgen05.mma (M128 x N x K16) into a TMEM accumulator; tcgen05.commit
// frees the stage.  Warp roles: warp0 = TMA producer, warp1 = TMEM alloc + MMA issue, warps2-5 = epilogue
// (tcgen05.ld -> bias / FiLM / CSM modulation / residual / SiLU / GroupNorm statistics -> bf16 NHWC store).
#include "common.cuh"
#include "ddif_internal.h"

namespace ddif {

struct alignas(64) GemmKParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB[2];
  int nseg;
  int taps[2];
  int kchunks[2];
  int b_per_sample[2];
  int pad[2];
  int stride;
  int batch, out_h, out_w;
  int tw, th, tn;          // pixel box
  int tiles_x, tiles_y;    // tiles per image
  int a_rows;              // rows TMA fills per stage (tw*th*tn, 64 or 128)
  int bn;                  // N per CTA
  int n_valid;
  int bk;                  // K elements per stage (16/32/64)
  int span;                // bytes per smem row = bk*2
  int stages;
  uint32_t idesc;
  uint32_t layout_type;
  uint32_t tmem_cols;
  const float* bias;
  const float* film;
  int film_ld;
  const bf16* mod;
  const bf16* residual;
  int res_ld;
  int act;
  bf16* out;
  int out_ld;
  float* out_nchw;
  double* stats;
};

static constexpr int kGemmThreads = 192;

__global__ void __launch_bounds__(kGemmThreads, 1) conv_igemm_tc_kernel(const __grid_constant__ GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  // carve: [barriers | pad to 1024 | A stages | B stages]
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int a_stage_bytes = 128 * p.span;
  const int b_stage_bytes = p.bn * p.span;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + (size_t)p.stages * a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.stages * b_stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + p.stages;
  uint64_t* tmem_full_bar = bars + 2 * p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int m_tile = blockIdx.x;
  const int n_tile = blockIdx.y;
  const int img_grp = m_tile / tiles_per_img;
  const int t_in = m_tile - img_grp * tiles_per_img;
  const int ty0 = (t_in / p.tiles_x) * p.th;
  const int tx0 = (t_in % p.tiles_x) * p.tw;
  const int n0 = img_grp * p.tn;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) {
      tma_prefetch_desc(&p.tmA[s]);
      tma_prefetch_desc(&p.tmB[s]);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < p.stages; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 1);
      }
      mbar_init(tmem_full_bar, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int total_iters = 0;
  for (int s = 0; s < p.nseg; ++s) total_iters += p.taps[s] * p.kchunks[s];

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t tx_bytes = (uint32_t)(p.a_rows * p.span + b_stage_bytes);
      int it = 0;
      for (int s = 0; s < p.nseg; ++s) {
        for (int tap = 0; tap < p.taps[s]; ++tap) {
          const int dy = (p.taps[s] == 9) ? tap / 3 : 0;
          const int dx = (p.taps[s] == 9) ? tap % 3 : 0;
          const int cw = tx0 * p.stride + dx - p.pad[s];
          const int ch = ty0 * p.stride + dy - p.pad[s];
          const int bz = p.b_per_sample[s] ? n0 : tap;
          for (int kc = 0; kc < p.kchunks[s]; ++kc, ++it) {
            const int stage = it % p.stages;
            const uint32_t phase = (uint32_t)(it / p.stages) & 1u;
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            mbar_expect_tx(&full_bar[stage], tx_bytes);
            tma_load_4d(&p.tmA[s], &full_bar[stage], smem_a + (size_t)stage * a_stage_bytes, kc * p.bk, cw, ch, n0);
            tma_load_3d(&p.tmB[s], &full_bar[stage], smem_b + (size_t)stage * b_stage_bytes, kc * p.bk, n_tile * p.bn, bz);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t sbo = 8u * (uint32_t)p.span;
      const int ksteps = p.bk / 16;
      for (int it = 0; it < total_iters; ++it) {
        const int stage = it % p.stages;
        const uint32_t phase = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t da = make_smem_desc(smem_u32(smem_a + (size_t)stage * a_stage_bytes), sbo, p.layout_type);
        const uint64_t db = make_smem_desc(smem_u32(smem_b + (size_t)stage * b_stage_bytes), sbo, p.layout_type);
        for (int k = 0; k < ksteps; ++k) {
          // advance 16 K-elements = 32 bytes inside the swizzle span (start address field is in 16-byte units)
          umma_bf16_ss(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int px_per_img = p.tw * p.th;
    const int tn_i = row / px_per_img;
    const int r_in = row - tn_i * px_per_img;
    const int y = ty0 + r_in / p.tw;
    const int x = tx0 + r_in % p.tw;
    const int b = n0 + tn_i;
    const bool row_ok = (row < p.a_rows) && (b < p.batch) && (y < p.out_h) && (x < p.out_w);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    float s1 = 0.f, s2 = 0.f;
    const size_t pix = ((size_t)b * p.out_h + y) * p.out_w + x;
    if (q * 32 < p.a_rows) {
      for (int c0 = 0; c0 < p.bn; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        const int ng = n_tile * p.bn + c0;  // global output channel of r[0]
        if (row_ok && ng < p.n_valid) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          const int nrem = p.n_valid - ng;  // >= 1
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < nrem) v[j] += __ldg(p.bias + ng + j);
          }
          if (p.film) {
            const float* f = p.film + (size_t)b * p.film_ld + ng;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < nrem) v[j] += __ldg(f + j);
          }
          if (p.mod) {
            const bf16* m = p.mod + pix * (size_t)(2 * p.n_valid) + ng;
            if (nrem >= 16) {
              float sc[16], sh[16];
              unpack8(*reinterpret_cast<const bf16x8*>(m), sc);
              unpack8(*reinterpret_cast<const bf16x8*>(m + 8), sc + 8);
              unpack8(*reinterpret_cast<const bf16x8*>(m + p.n_valid), sh);
              unpack8(*reinterpret_cast<const bf16x8*>(m + p.n_valid + 8), sh + 8);
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = v[j] * (1.0f + sc[j]) + sh[j];
            } else {
              for (int j = 0; j < nrem; ++j)
                v[j] = v[j] * (1.0f + __bfloat162float(m[j])) + __bfloat162float(m[p.n_valid + j]);
            }
          }
          if (p.residual) {
            const bf16* rs = p.residual + pix * (size_t)p.res_ld + ng;
            if (nrem >= 16) {
              float rr[16];
              unpack8(*reinterpret_cast<const bf16x8*>(rs), rr);
              unpack8(*reinterpret_cast<const bf16x8*>(rs + 8), rr + 8);
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] += rr[j];
            } else {
              for (int j = 0; j < nrem; ++j) v[j] += __bfloat162float(rs[j]);
            }
          }
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = silu_f(v[j]);
          }
          if (p.stats) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < nrem) {
                s1 += v[j];
                s2 += v[j] * v[j];
              }
          }
          if (p.out) {
            bf16* o = p.out + pix * (size_t)p.out_ld + ng;
            if (nrem >= 16) {
              *reinterpret_cast<bf16x8*>(o) = pack8(v);
              *reinterpret_cast<bf16x8*>(o + 8) = pack8(v + 8);
            } else {
              for (int j = 0; j < nrem; ++j) o[j] = __float2bfloat16(v[j]);
            }
          }
          if (p.out_nchw) {
            const size_t hw = (size_t)p.out_h * p.out_w;
            float* o = p.out_nchw + ((size_t)b * p.n_valid + ng) * hw + (size_t)y * p.out_w + x;
            for (int j = 0; j < 16 && j < nrem; ++j) o[(size_t)j * hw] = v[j];
          }
        }
      }
      if (p.stats) {
        // all rows of one warp belong to one sample (px_per_img is 64 or 128)
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        const int bw = n0 + (q * 32) / px_per_img;
        if (lane == 0 && bw < p.batch) {
          atomicAdd(p.stats + 2 * (size_t)bw, (double)s1);
          atomicAdd(p.stats + 2 * (size_t)bw + 1, (double)s2);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

static CUtensorMapSwizzle swizzle_for_span(int span) {
  return span == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : span == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

int gemm_prepare(const ddif_gemm_t& g, GemmLaunch& L) {
  GemmKParams& p = *reinterpret_cast<GemmKParams*>(L.kparams);
  static_assert(sizeof(GemmKParams) <= sizeof(L.kparams), "kparams buffer too small");
  memset(&p, 0, sizeof(p));
  PFN_encodeTiled enc = get_encode();
  if (!enc) return DDIF_ERR_DRIVER;
  static bool smem_attr_set = false;  // outside any stream capture: gemm_prepare runs at plan-build time
  if (!smem_attr_set) {
    DDIF_CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    smem_attr_set = true;
  }
  if (g.nseg < 1 || g.nseg > 2) return DDIF_ERR_ARG;
  if (g.stride != 1 && g.stride != 2) return DDIF_ERR_ARG;
  if (g.n_pad % 16 != 0 || g.n_valid > g.n_pad || g.n_valid < 1) return DDIF_ERR_SHAPE;
  const int W = (int)g.out_w, H = (int)g.out_h, B = (int)g.batch;
  // pixel box: tw x th x tn = 128 rows (64 when the weights are per-sample and an image has only 64 pixels)
  int tw = W >= 16 ? 16 : 8;
  if (W < 8) return DDIF_ERR_SHAPE;
  int th = 128 / tw;
  if (th > H) th = H >= 8 ? 8 : H;
  if (tw * th != 128 && tw * th != 64) return DDIF_ERR_SHAPE;
  bool per_sample = false;
  for (int s = 0; s < g.nseg; ++s) per_sample |= g.w_per_sample[s] != 0;
  int tn = 128 / (tw * th);
  if (per_sample) tn = 1;
  p.tw = tw; p.th = th; p.tn = tn;
  p.a_rows = tw * th * tn;
  p.tiles_x = (int)ceil_div(W, tw);
  p.tiles_y = (int)ceil_div(H, th);
  p.batch = B; p.out_h = H; p.out_w = W;
  p.stride = (int)g.stride;
  p.nseg = (int)g.nseg;
  // K chunk shared by all segments
  int bk = 64;
  for (int s = 0; s < g.nseg; ++s) {
    if (g.a_c[s] % 16 != 0 || g.a_ld[s] % 8 != 0 || g.w_k[s] % 8 != 0 || g.w_k[s] < g.a_c[s]) return DDIF_ERR_SHAPE;
    while (g.a_c[s] % bk != 0) bk >>= 1;
    if (g.taps[s] != 1 && g.taps[s] != 9) return DDIF_ERR_ARG;
    if (g.w_per_sample[s] && g.taps[s] != 1) return DDIF_ERR_ARG;
  }
  p.bk = bk;
  p.span = bk * 2;
  p.layout_type = p.span == 128 ? 2u : p.span == 64 ? 4u : 6u;
  int bn = (int)g.n_pad;
  if (bn > 256) {
    bn = (g.n_pad % 256 == 0) ? 256 : 128;
    if (g.n_pad % bn != 0) return DDIF_ERR_SHAPE;
  }
  p.bn = bn;
  p.n_valid = (int)g.n_valid;
  uint32_t cols = 32;
  while ((int)cols < bn) cols <<= 1;
  p.tmem_cols = cols;
  // instruction descriptor: D=f32, A=B=bf16, K-major both, N>>3 at bit 17, M>>4 at bit 24
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const int stage_bytes = (128 + bn) * p.span;
  int stages = (96 * 1024) / stage_bytes;
  if (stages < 3) stages = (200 * 1024) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return DDIF_ERR_SHAPE;
  p.stages = stages;
  L.smem_bytes = stages * stage_bytes + (2 * stages + 2) * 8 + 1024;
  L.grid_x = p.tiles_x * p.tiles_y * (int)ceil_div(B, tn);
  L.grid_y = (int)(g.n_pad / bn);

  const CUtensorMapSwizzle sw = swizzle_for_span(p.span);
  for (int s = 0; s < g.nseg; ++s) {
    p.taps[s] = (int)g.taps[s];
    p.kchunks[s] = (int)(g.a_c[s] / bk);
    p.b_per_sample[s] = g.w_per_sample[s] ? 1 : 0;
    p.pad[s] = g.taps[s] == 9 ? 1 : 0;
    {
      cuuint64_t dims[4] = {(cuuint64_t)g.a_c[s], (cuuint64_t)g.a_w[s], (cuuint64_t)g.a_h[s], (cuuint64_t)B};
      cuuint64_t strides[3] = {(cuuint64_t)g.a_ld[s] * 2, (cuuint64_t)g.a_w[s] * g.a_ld[s] * 2,
                               (cuuint64_t)g.a_h[s] * g.a_w[s] * g.a_ld[s] * 2};
      cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)(tw * g.stride), (cuuint32_t)(th * g.stride), (cuuint32_t)tn};
      cuuint32_t es[4] = {1, (cuuint32_t)g.stride, (cuuint32_t)g.stride, 1};
      CUresult r = enc(&p.tmA[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(g.a[s]), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return DDIF_ERR_DRIVER;
    }
    {
      cuuint64_t dims[3] = {(cuuint64_t)g.w_k[s], (cuuint64_t)g.n_pad, (cuuint64_t)g.w_s[s]};
      cuuint64_t strides[2] = {(cuuint64_t)g.w_k[s] * 2, (cuuint64_t)g.n_pad * g.w_k[s] * 2};
      cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)bn, 1};
      cuuint32_t es[3] = {1, 1, 1};
      CUresult r = enc(&p.tmB[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(g.w[s]), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return DDIF_ERR_DRIVER;
    }
  }
  p.bias = g.bias;
  p.film = g.film;
  p.film_ld = (int)g.film_ld;
  p.mod = (const bf16*)g.mod;
  p.residual = (const bf16*)g.residual;
  p.res_ld = (int)g.res_ld;
  p.act = (int)g.act;
  p.out = (bf16*)g.out;
  p.out_ld = (int)g.out_ld;
  p.out_nchw = g.out_nchw;
  p.stats = g.stats;
  if (p.out && (p.out_ld % 8 != 0)) return DDIF_ERR_SHAPE;
  if (p.mod && (p.n_valid % 8 != 0)) return DDIF_ERR_SHAPE;
  if (p.residual && (p.res_ld % 8 != 0)) return DDIF_ERR_SHAPE;
  return DDIF_OK;
}

int gemm_launch(const GemmLaunch& L, cudaStream_t stream) {
  const GemmKParams& p = *reinterpret_cast<const GemmKParams*>(L.kparams);
  conv_igemm_tc_kernel<<<dim3(L.grid_x, L.grid_y), kGemmThreads, L.smem_bytes, stream>>>(p);
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

}  // namespace ddif
