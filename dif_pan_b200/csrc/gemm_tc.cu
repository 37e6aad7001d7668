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
  int num_m_tiles;         // tiles_x * tiles_y * ceil(batch / tn)
  int a_rows;              // rows TMA fills per stage (tw*th*tn, 64 or 128)
  int bn;                  // N per CTA
  int n_valid;
  int bk;                  // K elements per stage (16/32/64)
  int span;                // bytes per smem row = bk*2
  int stages;              // A (and streamed-B) ring depth
  int resident_b;          // 1: all weight slabs of this CTA's N tile stay in smem for the CTA's lifetime
  int b_slots;             // resident_b ? total K iterations : stages
  uint32_t idesc;
  uint32_t layout_type;
  uint32_t tmem_cols;      // 2 accumulators of bn columns, rounded to a power of two >= 32
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

static constexpr int kEpiWarps = 16;                      // 4 TMEM lane quarters x 4 column groups
static constexpr int kGemmThreads = 64 + 32 * kEpiWarps;  // warp0 TMA, warp1 MMA, warps 2..17 epilogue

__device__ __forceinline__ void ldg256(const void* p, uint32_t* r) {
  asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]),
               "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void unpack16(const uint32_t* r, float* f) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r[i]));
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Persistent, warp-specialised: each CTA owns one N tile (blockIdx.y) and walks M tiles blockIdx.x, +gridDim.x, ...
// Three pipelines: smem full/empty ring (TMA <-> MMA) running ACROSS tiles, TMEM full/empty (MMA <-> epilogue, two
// accumulators so tile i's epilogue overlaps tile i+1's MMAs), and the static tile walk.
__global__ void __launch_bounds__(kGemmThreads, 1) conv_igemm_tc_kernel(const __grid_constant__ GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int a_stage_bytes = 128 * p.span;
  const int b_slot_bytes = p.bn * p.span;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + (size_t)p.stages * a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.b_slots * b_slot_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + p.stages;
  uint64_t* tmem_full_bar = bars + 2 * p.stages;       // [2]
  uint64_t* tmem_empty_bar = bars + 2 * p.stages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.y;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

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
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tmem_full_bar[i], 1);
        mbar_init(&tmem_empty_bar[i], kEpiWarps);  // one arrive per epilogue warp
      }
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
      const uint32_t a_bytes = (uint32_t)(p.a_rows * p.span);
      uint32_t it = 0;
      bool first = true;
      for (int m_tile = blockIdx.x; m_tile < p.num_m_tiles; m_tile += gridDim.x, first = false) {
        const int img_grp = m_tile / tiles_per_img;
        const int t_in = m_tile - img_grp * tiles_per_img;
        const int ty0 = (t_in / p.tiles_x) * p.th;
        const int tx0 = (t_in % p.tiles_x) * p.tw;
        const int n0 = img_grp * p.tn;
        const bool load_b = !p.resident_b || first;
        int j = 0;
        for (int s = 0; s < p.nseg; ++s) {
          for (int tap = 0; tap < p.taps[s]; ++tap) {
            const int dy = (p.taps[s] == 9) ? tap / 3 : 0;
            const int dx = (p.taps[s] == 9) ? tap % 3 : 0;
            const int cw = tx0 * p.stride + dx - p.pad[s];
            const int ch = ty0 * p.stride + dy - p.pad[s];
            const int bz = p.b_per_sample[s] ? n0 : tap;
            for (int kc = 0; kc < p.kchunks[s]; ++kc, ++it, ++j) {
              const uint32_t stage = it % (uint32_t)p.stages;
              const uint32_t phase = (it / (uint32_t)p.stages) & 1u;
              mbar_wait(&empty_bar[stage], phase ^ 1u);
              mbar_expect_tx(&full_bar[stage], a_bytes + (load_b ? (uint32_t)b_slot_bytes : 0u));
              tma_load_4d(&p.tmA[s], &full_bar[stage], smem_a + (size_t)stage * a_stage_bytes, kc * p.bk, cw, ch, n0);
              if (load_b) {
                const int slot = p.resident_b ? j : (int)stage;
                tma_load_3d(&p.tmB[s], &full_bar[stage], smem_b + (size_t)slot * b_slot_bytes, kc * p.bk, n_tile * p.bn, bz);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t sbo = 8u * (uint32_t)p.span;
      const int ksteps = p.bk / 16;
      uint32_t it = 0, tcount = 0;
      for (int m_tile = blockIdx.x; m_tile < p.num_m_tiles; m_tile += gridDim.x, ++tcount) {
        const uint32_t acc = tcount & 1u, acc_phase = (tcount >> 1) & 1u;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * (uint32_t)p.bn;
        for (int j = 0; j < total_iters; ++j, ++it) {
          const uint32_t stage = it % (uint32_t)p.stages;
          const uint32_t phase = (it / (uint32_t)p.stages) & 1u;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const int slot = p.resident_b ? j : (int)stage;
          const uint64_t da = make_smem_desc(smem_u32(smem_a + (size_t)stage * a_stage_bytes), sbo, p.layout_type);
          const uint64_t db = make_smem_desc(smem_u32(smem_b + (size_t)slot * b_slot_bytes), sbo, p.layout_type);
          for (int k = 0; k < ksteps; ++k)
            umma_bf16_ss(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc, (j | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..17) =====================
    // warp -> (TMEM lane quarter q = warp % 4, column group cg): thread = one output pixel x (bn / ngroups) channels.
    // Global operands of the first 16-column chunk (residual / CSM scale+shift) are fetched BEFORE waiting for the
    // accumulator, so their latency hides behind the MMAs of the tile.
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int nchunks = p.bn >> 4;  // 16-column chunks of the accumulator; group cg owns chunks cg, cg+4, cg+8, ...
    const int row = q * 32 + lane;
    const int px_per_img = p.tw * p.th;
    const int tn_i = row / px_per_img;
    const int r_in = row - tn_i * px_per_img;
    const int ry = r_in / p.tw, rx = r_in % p.tw;
    const bool wide_io = (p.out_ld % 16 == 0) && (p.res_ld % 16 == 0) && (p.n_valid % 16 == 0);
    const bool active = (cg < nchunks) && (q * 32 < p.a_rows);
    uint32_t tcount = 0;
    for (int m_tile = blockIdx.x; m_tile < p.num_m_tiles; m_tile += gridDim.x, ++tcount) {
      const int img_grp = m_tile / tiles_per_img;
      const int t_in = m_tile - img_grp * tiles_per_img;
      const int y = (t_in / p.tiles_x) * p.th + ry;
      const int x = (t_in % p.tiles_x) * p.tw + rx;
      const int n0 = img_grp * p.tn;
      const int b = n0 + tn_i;
      const bool row_ok = active && (row < p.a_rows) && (b < p.batch) && (y < p.out_h) && (x < p.out_w);
      const size_t pix = ((size_t)b * p.out_h + y) * p.out_w + x;
      const int c_first = cg * 16;
      uint32_t pre_res[8], pre_sc[8], pre_sh[8];
      {
        const int ng = n_tile * p.bn + c_first;
        const bool full16 = wide_io && row_ok && (p.n_valid - ng >= 16);
        if (full16 && p.residual) ldg256(p.residual + pix * (size_t)p.res_ld + ng, pre_res);
        if (full16 && p.mod) {
          const bf16* m = p.mod + pix * (size_t)(2 * p.n_valid) + ng;
          ldg256(m, pre_sc);
          ldg256(m + p.n_valid, pre_sh);
        }
      }
      const uint32_t acc = tcount & 1u, acc_phase = (tcount >> 1) & 1u;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      float s1 = 0.f, s2 = 0.f;
      for (int cc = cg; cc < nchunks || cc == cg; cc += 4) {
        const int c0 = cc * 16;
        uint32_t r[16];
        if (active) {
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)p.bn + (uint32_t)c0, r);
          tmem_ld_wait();
        }
        if (cc + 4 >= nchunks) {
          // this warp's TMEM reads of the accumulator are done: hand it back to the MMA warp before the math/stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
        const int ng = n_tile * p.bn + c0;  // global output channel of r[0]
        if (row_ok && ng < p.n_valid) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          const int nrem = p.n_valid - ng;  // >= 1
          const bool full16 = wide_io && nrem >= 16;
          if (p.bias) {
            if (nrem >= 16) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + ng + j));
                v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
              }
            } else {
              for (int j = 0; j < nrem; ++j) v[j] += __ldg(p.bias + ng + j);
            }
          }
          if (p.film) {
            const float* f = p.film + (size_t)b * p.film_ld + ng;
            if (nrem >= 16 && (p.film_ld % 4 == 0)) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(f + j));
                v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
              }
            } else {
              for (int j = 0; j < 16 && j < nrem; ++j) v[j] += __ldg(f + j);
            }
          }
          if (p.mod) {
            const bf16* m = p.mod + pix * (size_t)(2 * p.n_valid) + ng;
            if (full16) {
              if (cc != cg) {
                ldg256(m, pre_sc);
                ldg256(m + p.n_valid, pre_sh);
              }
              float sc[16], sh[16];
              unpack16(pre_sc, sc);
              unpack16(pre_sh, sh);
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = v[j] * (1.0f + sc[j]) + sh[j];
            } else {
              for (int j = 0; j < 16 && j < nrem; ++j)
                v[j] = v[j] * (1.0f + __bfloat162float(m[j])) + __bfloat162float(m[p.n_valid + j]);
            }
          }
          if (p.residual) {
            const bf16* rs = p.residual + pix * (size_t)p.res_ld + ng;
            if (full16) {
              if (cc != cg) ldg256(rs, pre_res);
              float rr[16];
              unpack16(pre_res, rr);
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] += rr[j];
            } else {
              for (int j = 0; j < 16 && j < nrem; ++j) v[j] += __bfloat162float(rs[j]);
            }
          }
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __fdividef(v[j], 1.0f + __expf(-v[j]));
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
            if (full16) {
              uint32_t w[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                w[j] = *reinterpret_cast<const uint32_t*>(&t);
              }
              stg256(o, w);
            } else if (nrem >= 16) {
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
      if (p.stats && active) {
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
  while ((int)cols < 2 * bn) cols <<= 1;  // two accumulators
  p.tmem_cols = cols;
  // instruction descriptor: D=f32, A=B=bf16, K-major both, N>>3 at bit 17, M>>4 at bit 24
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  int total_iters = 0;
  for (int s = 0; s < g.nseg; ++s) total_iters += (int)(g.taps[s] * (g.a_c[s] / bk));
  const int a_stage = 128 * p.span, b_slot = bn * p.span;
  const int budget = 200 * 1024;
  // weights of this N tile stay resident in smem when they fit next to >= 4 A stages (never for per-sample weights)
  p.resident_b = (!per_sample && (int64_t)total_iters * b_slot + 4 * a_stage <= budget) ? 1 : 0;
  int stages;
  if (p.resident_b) {
    p.b_slots = total_iters;
    stages = (budget - total_iters * b_slot) / a_stage;
  } else {
    stages = budget / (a_stage + b_slot);
    p.b_slots = stages;
  }
  if (stages > 8) stages = 8;
  if (stages < 2) return DDIF_ERR_SHAPE;
  if (!p.resident_b) p.b_slots = stages;
  p.stages = stages;
  p.num_m_tiles = p.tiles_x * p.tiles_y * (int)ceil_div(B, tn);
  L.smem_bytes = stages * a_stage + p.b_slots * b_slot + (2 * stages + 6) * 8 + 1024;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      sms = n;
    else
      sms = 148;
  }
  L.grid_y = (int)(g.n_pad / bn);
  const int gx = (sms + L.grid_y - 1) / L.grid_y;  // persistent: ~one CTA per SM in total
  L.grid_x = p.num_m_tiles < gx ? p.num_m_tiles : gx;

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
