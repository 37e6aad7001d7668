// Fused 3x3 (stride 1, pad 1) convolution: [GroupNorm(1 group) + Swish] -> conv -> epilogue, tcgen05 / TMEM.
//
// Replaces `Block` = GN -> Swish -> Conv3x3 (/root/reference/models/sr3_dwt.py:288-300), the FWM ffn 3x3 convs
// (:529-531), `Upsample` = nearest x2 -> Conv3x3 (:266-273, the x2 is folded into the loader) and downs[0] (:86).
//
// Why not TMA for the activations: ncu on the TMA version (profiles/r01_*) shows the nine tap-shifted box loads are
// limited by the TMA row rate (~8-10 cycles per 64-128 B pixel row), i.e. ~9x the L2->SMEM traffic at ~13 B/clk/SM.
// Here 8 loader warps read every pixel of the (8+2)x(16+2) halo tile ONCE with 16-byte LDGs (register-prefetched
// several tiles ahead), apply the GroupNorm affine + Swish in registers (so the normalised tensor never exists in
// HBM), and store it into THREE dx-shifted K-major swizzled copies in shared memory.  Tap (dy, dx) of the implicit
// GEMM is then the aligned sub-tile `copy[dx] + dy*16 rows`: every UMMA descriptor start stays a multiple of the
// swizzle atom, no undocumented descriptor arithmetic.  Weights come by TMA (resident in smem when they fit, else a
// streamed ring); one thread issues tcgen05.mma into two TMEM accumulators; 4 epilogue warps run epilogue.cuh.
#include "common.cuh"
#include "ddif_internal.h"
#include "epilogue.cuh"

namespace ddif {

static constexpr int kLoadThreads = 256;
#ifndef DDIF_C3_EPI_WARPS
#define DDIF_C3_EPI_WARPS 8
#endif
#ifndef DDIF_C3_PF64
#define DDIF_C3_PF64 1   // halo tiles kept in flight in registers per loader thread, 64-channel slabs (6 x 16 B each)
#endif
#ifndef DDIF_C3_PF32
#define DDIF_C3_PF32 2   // ... 32-channel slabs (3 x 16 B each)
#endif
static constexpr int kC3EpiWarps = DDIF_C3_EPI_WARPS;   // 8 loader + 2 + {4,8} epilogue warps = 448 / 576 threads
static constexpr int kC3Threads = kLoadThreads + 64 + 32 * kC3EpiWarps;  // 8 loader warps, MMA, weight-TMA, epilogue warps
static constexpr int kHaloW = 18, kHaloH = 10, kHaloPx = kHaloW * kHaloH;
static constexpr int kCopyRows = kHaloH * 16;  // 10 lines x 16 pixels
static constexpr int kMaxBSlots = 32;
static constexpr int kMaxStatSamples = 1024;  // per-CTA (mean, rstd) table in smem; larger batches recompute per tile

struct alignas(64) Conv3KParams {
  CUtensorMap tmB;
  const bf16* src;
  int src_ld, src_h, src_w, up;
  int cin, kslab, nslab, span;
  int batch, out_h, out_w;
  int tiles_x, tiles_y, num_m_tiles;
  int bn;
  int resident_b, b_slots;
  uint32_t idesc, layout_type, tmem_cols;
  const double* gn_stats;
  const float* gn_gamma;
  const float* gn_beta;
  float gn_eps;
  int gn_act;
  double gn_count;
  EpiParams epi;
};

// Debug hook: when set (ddif_debug_set_timestamps), CTA 0 of the fused kernel records clock64() at the pipeline hand-offs
// of its first 64 tiles: ts[(role*64 + tile)*4 + k], role 0 = loader thread 0, 1 = MMA issuer, 2 = epilogue warp 10 lane 0.
__device__ long long* g_debug_ts = nullptr;
__device__ __forceinline__ void dbg_ts(int role, int tile, int k) {
  if (g_debug_ts && blockIdx.x == 0 && blockIdx.y == 0 && tile < 64) g_debug_ts[(role * 64 + tile) * 4 + k] = clock64();
}

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

struct TileCoord {
  int b, y0, x0;
};
__device__ __forceinline__ TileCoord tile_coord(const Conv3KParams& p, int m_tile) {
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  TileCoord t;
  t.b = m_tile / tiles_per_img;
  const int r = m_tile - t.b * tiles_per_img;
  t.y0 = (r / p.tiles_x) * 8;
  t.x0 = (r % p.tiles_x) * 16;
  return t;
}

// Walks the M tiles of one persistent CTA (blockIdx.x, +gridDim.x, ...) and their K slabs WITHOUT divisions in the
// loop: (image, tile row, tile column) advance by precomputed steps with carries.
struct UnitIter {
  int b, ty, tx, slab, remaining;  // remaining = units left including the current one
  int sb, sy, sx, tiles_x, tiles_y, nslab;
  __device__ __forceinline__ void init(const Conv3KParams& p) {
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    tiles_x = p.tiles_x; tiles_y = p.tiles_y; nslab = p.nslab;
    const int m0 = (int)blockIdx.x, g = (int)gridDim.x;
    b = m0 / tiles_per_img;
    int r = m0 - b * tiles_per_img;
    ty = r / tiles_x; tx = r - ty * tiles_x;
    sb = g / tiles_per_img;
    r = g - sb * tiles_per_img;
    sy = r / tiles_x; sx = r - sy * tiles_x;
    slab = 0;
    const int my_tiles = (p.num_m_tiles - m0 + g - 1) / g;
    remaining = my_tiles * nslab;
  }
  __device__ __forceinline__ void next() {
    --remaining;
    if (++slab < nslab) return;
    slab = 0;
    tx += sx; ty += sy; b += sb;
    if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
    if (ty >= tiles_y) { ty -= tiles_y; ++b; }
  }
};

// ---- loader warps ------------------------------------------------------------------------------------------------
// NCK = 16-byte chunks per pixel per K slab (kslab / 8); PF = units (tile x slab) kept in flight in registers.
// Everything that does not depend on the tile (halo pixel of each of the thread's loads, its three swizzled smem
// destinations) is computed once; per unit a thread only adds the tile origin, bounds-checks, and moves data.
template <int NCK, int PF>
__device__ __forceinline__ void loader_loop(const Conv3KParams& p, uint32_t a_base, uint64_t* a_full, uint64_t* a_empty, const float* s_gamma,
                                            const float* s_beta, const float2* s_stat, int lt) {
  constexpr int PXP = kLoadThreads / NCK;                 // pixels covered per pass
  constexpr int LPT = (kHaloPx + PXP - 1) / PXP;          // loads per thread per unit
  const int c = lt % NCK;                                 // this thread's chunk inside the slab (fixed)
  const int px0 = lt / NCK;
  const uint32_t copy_bytes = (uint32_t)(kCopyRows * p.span);
  const uint32_t stage_bytes = 3u * copy_bytes;
  const bool gn = p.gn_stats != nullptr;

  int hl[LPT], hj[LPT];          // halo line / column of load k (hl < 0: no such pixel)
  int koff[LPT];                 // element offset of halo pixel k relative to the tile origin pixel (up == 0 only)
  uint32_t soff[LPT][3];         // smem byte offset inside a stage for copy dx (0xffffffff: not stored)
#pragma unroll
  for (int k = 0; k < LPT; ++k) {
    const int pxi = px0 + k * PXP;
    const int l = pxi / kHaloW, j = pxi - l * kHaloW;
    hl[k] = pxi < kHaloPx ? l : -100000;
    hj[k] = j;
    koff[k] = ((l - 1) * p.src_w + (j - 1)) * p.src_ld;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int xx = j - dx;
      uint32_t o = 0xffffffffu;
      if (pxi < kHaloPx && xx >= 0 && xx <= 15) {
        const uint32_t row = (uint32_t)(l * 16 + xx);
        const uint32_t sw = NCK == 8 ? (row & 7u) : (NCK == 4 ? ((row >> 1) & 3u) : ((row >> 2) & 1u));
        o = (uint32_t)dx * copy_bytes + row * (uint32_t)p.span + (((uint32_t)c ^ sw) << 4);
      }
      soff[k][dx] = o;
    }
  }

  uint4 buf[PF][LPT];
  const size_t img_stride = (size_t)p.src_h * p.src_w * p.src_ld;

  auto issue = [&](const UnitIter& it, uint4 (&dst)[LPT]) {
    const int y0 = it.ty * 8 - 1, x0 = it.tx * 16 - 1;
    const bf16* base = p.src + (size_t)it.b * img_stride + it.slab * p.kslab + c * 8;
    if (p.up == 0) {
      const bf16* tbase = base + (it.ty * 8 * p.src_w + it.tx * 16) * p.src_ld;  // tile origin pixel
#pragma unroll
      for (int k = 0; k < LPT; ++k) {
        const int gy = y0 + hl[k], gx = x0 + hj[k];
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if ((unsigned)gy < (unsigned)p.out_h && (unsigned)gx < (unsigned)p.out_w) v = __ldg(reinterpret_cast<const uint4*>(tbase + koff[k]));
        dst[k] = v;
      }
    } else {
#pragma unroll
      for (int k = 0; k < LPT; ++k) {
        const int gy = y0 + hl[k], gx = x0 + hj[k];
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if ((unsigned)gy < (unsigned)p.out_h && (unsigned)gx < (unsigned)p.out_w)
          v = __ldg(reinterpret_cast<const uint4*>(base + ((gy >> 1) * p.src_w + (gx >> 1)) * p.src_ld));
        dst[k] = v;
      }
    }
  };

  auto consume = [&](const UnitIter& it, uint32_t u, const uint4 (&srcv)[LPT]) {
    const int y0 = it.ty * 8 - 1, x0 = it.tx * 16 - 1;
    const uint32_t stage = u & 1u, phase = (u >> 1) & 1u;
    float a[8], d[8];
    if (gn) {
      float mean, rstd;
      if (it.b < kMaxStatSamples) {
        const float2 mr = s_stat[it.b];
        mean = mr.x;
        rstd = mr.y;
      } else {
        const double s = p.gn_stats[2 * it.b], ss = p.gn_stats[2 * it.b + 1];
        const double m = s / p.gn_count;
        double var = ss / p.gn_count - m * m;
        if (var < 0) var = 0;
        mean = (float)m;
        rstd = rsqrtf((float)var + p.gn_eps);
      }
      // with Swish the affine is pre-halved: swish(t) = h*tanh(h) + h, h = t/2
      const float hs = p.gn_act ? 0.5f : 1.0f;
      const int ch0 = it.slab * p.kslab + c * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[j] = hs * rstd * s_gamma[ch0 + j];
        d[j] = hs * s_beta[ch0 + j] - mean * a[j];
      }
    }
    if (lt == 0) dbg_ts(0, (int)u, 0);
    mbar_wait(&a_empty[stage], phase ^ 1u);
    if (lt == 0) dbg_ts(0, (int)u, 1);
    const uint32_t sbase = a_base + stage * stage_bytes;
#pragma unroll
    for (int k = 0; k < LPT; ++k) {
      uint4 v = srcv[k];
      const int gy = y0 + hl[k], gx = x0 + hj[k];
      if (gn && (unsigned)gy < (unsigned)p.out_h && (unsigned)gx < (unsigned)p.out_w) {
        float f[8];
        unpack8(*reinterpret_cast<const bf16x8*>(&v), f);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float t = fmaf(f[q], a[q], d[q]);
          f[q] = p.gn_act ? swish_half(t) : t;
        }
        const bf16x8 pk = pack8(f);
        v = *reinterpret_cast<const uint4*>(&pk);
      }
#pragma unroll
      for (int dx = 0; dx < 3; ++dx)
        if (soff[k][dx] != 0xffffffffu) sts128(sbase + soff[k][dx], v);
    }
    if (lt == 0) dbg_ts(0, (int)u, 2);
    fence_proxy_async();
    mbar_arrive(&a_full[stage]);
    if (lt == 0) dbg_ts(0, (int)u, 3);
  };

  UnitIter it_c, it_i;  // consume / issue positions
  it_c.init(p);
  it_i = it_c;
#pragma unroll
  for (int s = 0; s < PF; ++s)
    if (it_i.remaining > 0) {
      issue(it_i, buf[s]);
      it_i.next();
    }
  uint32_t u = 0;
  while (it_c.remaining > 0) {
#pragma unroll
    for (int s = 0; s < PF; ++s) {
      if (it_c.remaining > 0) {
        consume(it_c, u, buf[s]);
        it_c.next();
        ++u;
        if (it_i.remaining > 0) {
          issue(it_i, buf[s]);
          it_i.next();
        }
      }
    }
  }
}

// ---- MMA warp ----------------------------------------------------------------------------------------------------
template <int KSTEPS>
__device__ __forceinline__ void mma_loop(const Conv3KParams& p, uint32_t a_base, uint32_t b_base, uint64_t* a_full, uint64_t* a_empty,
                                         uint64_t* b_full, uint64_t* b_empty, uint64_t* tmem_full, uint64_t* tmem_empty, uint32_t tmem_base,
                                         int my_tiles) {
  const uint32_t copy_bytes = (uint32_t)(kCopyRows * p.span);
  const uint32_t sbo = 8u * (uint32_t)p.span;
  const uint64_t desc_a0 = make_smem_desc(a_base, sbo, p.layout_type);
  const uint64_t desc_b0 = make_smem_desc(b_base, sbo, p.layout_type);
  const uint32_t a_stage16 = (3u * copy_bytes) >> 4, copy16 = copy_bytes >> 4, line16 = (16u * (uint32_t)p.span) >> 4;
  const uint32_t b16 = (uint32_t)(p.bn * p.span) >> 4;
  uint32_t tap_off[9];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) tap_off[tap] = (uint32_t)(tap % 3) * copy16 + (uint32_t)(tap / 3) * line16;
  const uint32_t nslots = (uint32_t)p.b_slots;
  const bool resident = p.resident_b != 0;
  uint32_t bslot = 0, bphase = 0, u = 0;
  for (int t = 0; t < my_tiles; ++t) {
    const uint32_t acc = (uint32_t)t & 1u;
    if ((threadIdx.x & 31) == 0) dbg_ts(1, t, 0);
    mbar_wait(&tmem_empty[acc], (((uint32_t)t >> 1) & 1u) ^ 1u);
    tc_fence_after();
    if ((threadIdx.x & 31) == 0) dbg_ts(1, t, 1);
    const uint32_t tmem_d = tmem_base + acc * (uint32_t)p.bn;
    for (int slab = 0; slab < p.nslab; ++slab, ++u) {
      const uint32_t stage = u & 1u;
      mbar_wait(&a_full[stage], (u >> 1) & 1u);
      tc_fence_after();
      if ((threadIdx.x & 31) == 0) dbg_ts(1, t, 2);
      const uint64_t da_stage = desc_a0 + (uint64_t)(stage * a_stage16);
      if (resident) {
        uint64_t db = desc_b0 + (uint64_t)((uint32_t)(slab * 9) * b16);
        if (t == 0) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_full[slab * 9 + tap], 0u);
            tc_fence_after();
            umma_bf16_ss_steps<KSTEPS>(tmem_d, da_stage + tap_off[tap], db, p.idesc, (slab | tap) != 0 ? 1u : 0u);
            db += b16;
          }
        } else {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            umma_bf16_ss_steps<KSTEPS>(tmem_d, da_stage + tap_off[tap], db, p.idesc, (slab | tap) != 0 ? 1u : 0u);
            db += b16;
          }
        }
      } else {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&b_full[bslot], bphase);
          tc_fence_after();
          umma_bf16_ss_steps<KSTEPS>(tmem_d, da_stage + tap_off[tap], desc_b0 + (uint64_t)(bslot * b16), p.idesc, (slab | tap) != 0 ? 1u : 0u);
          umma_commit_elect(&b_empty[bslot]);
          if (++bslot == nslots) {
            bslot = 0;
            bphase ^= 1u;
          }
        }
      }
      umma_commit_elect(&a_empty[stage]);
    }
    umma_commit_elect(&tmem_full[acc]);
    if ((threadIdx.x & 31) == 0) dbg_ts(1, t, 3);
  }
}

__global__ void __launch_bounds__(kC3Threads, 1) conv3x3_fused_tc_kernel(const __grid_constant__ Conv3KParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t copy_bytes = (uint32_t)(kCopyRows * p.span);
  const uint32_t a_stage_bytes = 3u * copy_bytes;
  const uint32_t b_slot_bytes = (uint32_t)(p.bn * p.span);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + 2 * a_stage_bytes;
  float* s_gamma = reinterpret_cast<float*>(smem_b + (size_t)p.b_slots * b_slot_bytes);
  float* s_beta = s_gamma + p.cin;
  float2* s_stat = reinterpret_cast<float2*>(s_beta + p.cin);
  const int n_stat = p.gn_stats ? (p.batch < kMaxStatSamples ? p.batch : kMaxStatSamples) : 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stat + n_stat);
  uint64_t* a_full = bars;            // [2]  count 256 (loader threads)
  uint64_t* a_empty = bars + 2;       // [2]  tcgen05.commit
  uint64_t* tmem_full = bars + 4;     // [2]
  uint64_t* tmem_empty = bars + 6;    // [2]  count kC3EpiWarps
  uint64_t* b_full = bars + 8;        // [b_slots]
  uint64_t* b_empty = b_full + kMaxBSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_empty + kMaxBSlots);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.y;
  const int my_tiles = (p.num_m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (p.gn_stats) {
    for (int i = threadIdx.x; i < p.cin; i += blockDim.x) {
      s_gamma[i] = p.gn_gamma[i];
      s_beta[i] = p.gn_beta[i];
    }
    for (int i = threadIdx.x; i < n_stat; i += blockDim.x) {  // per-sample (mean, rstd), fp64 once per CTA
      const double s = p.gn_stats[2 * i], ss = p.gn_stats[2 * i + 1];
      const double m = s / p.gn_count;
      double var = ss / p.gn_count - m * m;
      if (var < 0) var = 0;
      s_stat[i] = make_float2((float)m, rsqrtf((float)var + p.gn_eps));
    }
  }
  if (warp == 9 && lane == 0) tma_prefetch_desc(&p.tmB);
  if (warp == 8) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&a_full[i], kLoadThreads);
        mbar_init(&a_empty[i], 1);
        mbar_init(&tmem_full[i], 1);
        mbar_init(&tmem_empty[i], kC3EpiWarps);
      }
      for (int i = 0; i < p.b_slots; ++i) {
        mbar_init(&b_full[i], 1);
        mbar_init(&b_empty[i], 1);
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

  if (warp < 8) {
    // ===================== activation loaders (LDG -> GN+Swish -> 3 dx-shifted swizzled copies) =====================
    const uint32_t a_base = smem_u32(smem_a);
    if (p.kslab == 64) loader_loop<8, DDIF_C3_PF64>(p, a_base, a_full, a_empty, s_gamma, s_beta, s_stat, threadIdx.x);
    else if (p.kslab == 32) loader_loop<4, DDIF_C3_PF32>(p, a_base, a_full, a_empty, s_gamma, s_beta, s_stat, threadIdx.x);
    else loader_loop<2, 4>(p, a_base, a_full, a_empty, s_gamma, s_beta, s_stat, threadIdx.x);
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    // The single-thread instruction count per MMA bounds the kernel (a lone thread retires one dependent instruction
    // every ~5 cycles), so this warp stays converged, MMAs are issued by an elected lane inside one asm block per tap
    // (K steps unrolled at compile time), and every descriptor is `base + small precomputed offset`.
    if (p.kslab == 64) mma_loop<4>(p, smem_u32(smem_a), smem_u32(smem_b), a_full, a_empty, b_full, b_empty, tmem_full, tmem_empty, tmem_base, my_tiles);
    else if (p.kslab == 32) mma_loop<2>(p, smem_u32(smem_a), smem_u32(smem_b), a_full, a_empty, b_full, b_empty, tmem_full, tmem_empty, tmem_base, my_tiles);
    else mma_loop<1>(p, smem_u32(smem_a), smem_u32(smem_b), a_full, a_empty, b_full, b_empty, tmem_full, tmem_empty, tmem_base, my_tiles);
  } else if (warp == 9) {
    // ===================== weight producer (TMA) =====================
    if (lane == 0) {
      if (p.resident_b) {
        for (int slab = 0; slab < p.nslab; ++slab)
          for (int tap = 0; tap < 9; ++tap) {
            const int slot = slab * 9 + tap;
            mbar_expect_tx(&b_full[slot], b_slot_bytes);
            tma_load_3d(&p.tmB, &b_full[slot], smem_b + (size_t)slot * b_slot_bytes, slab * p.kslab, n_tile * p.bn, tap);
          }
      } else {
        uint32_t slot = 0, phase = 0;
        const uint32_t nslots = (uint32_t)p.b_slots;
        for (int t = 0; t < my_tiles; ++t)
          for (int slab = 0; slab < p.nslab; ++slab)
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&b_empty[slot], phase ^ 1u);
              mbar_expect_tx(&b_full[slot], b_slot_bytes);
              tma_load_3d(&p.tmB, &b_full[slot], smem_b + (size_t)slot * b_slot_bytes, slab * p.kslab, n_tile * p.bn, tap);
              if (++slot == nslots) {
                slot = 0;
                phase ^= 1u;
              }
            }
      }
    }
  } else {
    // ===================== epilogue (warps 10..13) =====================
    const int q = warp & 3;
    const int cg = (warp - 10) >> 2;
    const int row = q * 32 + lane;
    const int ry = row >> 4, rx = row & 15;
    const bool active = cg < (p.bn >> 4);
    for (int t = 0; t < my_tiles; ++t) {
      const TileCoord tc = tile_coord(p, (int)blockIdx.x + t * (int)gridDim.x);
      const int y = tc.y0 + ry, x = tc.x0 + rx, b = tc.b;
      const bool row_ok = active && (y < p.out_h) && (x < p.out_w);
      const size_t pix = ((size_t)b * p.out_h + y) * p.out_w + x;
      EpiPrefetch pf;
      epilogue_prefetch<kC3EpiWarps / 4>(p.epi, pf, n_tile, p.bn, cg, row_ok, pix);
      const uint32_t acc = (uint32_t)t & 1u, acc_phase = ((uint32_t)t >> 1) & 1u;
      if (warp == 10 && lane == 0) dbg_ts(2, t, 0);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (warp == 10 && lane == 0) dbg_ts(2, t, 1);
      epilogue_tile<kC3EpiWarps / 4>(p.epi, pf, tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)p.bn, &tmem_empty[acc], p.bn, n_tile, cg, lane,
                       active, row_ok, b, y, x, pix, b);
      if (warp == 10 && lane == 0) dbg_ts(2, t, 2);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled ddif_get_encode();
int ddif_sm_count();

bool conv3_applicable(const ddif_gemm_t& g) {
  if (g.nseg != 1 || g.taps[0] != 9 || g.stride != 1 || g.w_per_sample[0]) return false;
  if (g.out_w < 16 || g.out_h < 8) return false;
  if (g.a_c[0] % 16 != 0 || g.a_c[0] > 256) return false;
  if (g.a_up != 0 && g.a_up != 1) return false;
  return true;
}

int conv3_prepare(const ddif_gemm_t& g, GemmLaunch& L) {
  Conv3KParams& p = *reinterpret_cast<Conv3KParams*>(L.kparams);
  static_assert(sizeof(Conv3KParams) <= sizeof(L.kparams), "kparams buffer too small");
  memset(&p, 0, sizeof(p));
  PFN_encodeTiled enc = ddif_get_encode();
  if (!enc) return DDIF_ERR_DRIVER;
  static bool attr_set = false;
  if (!attr_set) {
    DDIF_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_fused_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  if (!conv3_applicable(g)) return DDIF_ERR_SHAPE;
  if (g.n_pad % 16 != 0 || g.n_valid > g.n_pad || g.n_valid < 1) return DDIF_ERR_SHAPE;
  if (g.a_ld[0] % 8 != 0 || g.w_k[0] % 8 != 0 || g.w_k[0] < g.a_c[0]) return DDIF_ERR_SHAPE;
  const int cin = (int)g.a_c[0];
  p.src = (const bf16*)g.a[0];
  p.src_ld = (int)g.a_ld[0];
  p.up = (int)g.a_up;
  p.src_h = (int)g.a_h[0];
  p.src_w = (int)g.a_w[0];
  p.batch = (int)g.batch; p.out_h = (int)g.out_h; p.out_w = (int)g.out_w;
  if ((p.src_h << p.up) != p.out_h || (p.src_w << p.up) != p.out_w) return DDIF_ERR_SHAPE;
  p.cin = cin;
  p.kslab = cin % 64 == 0 ? 64 : (cin % 32 == 0 ? 32 : 16);
  p.nslab = cin / p.kslab;
  p.span = p.kslab * 2;
  p.layout_type = p.span == 128 ? 2u : p.span == 64 ? 4u : 6u;
  p.tiles_x = (int)ceil_div(p.out_w, 16);
  p.tiles_y = (int)ceil_div(p.out_h, 8);
  p.num_m_tiles = p.tiles_x * p.tiles_y * p.batch;
  int bn = (int)g.n_pad;
  if (bn > 256) {
    bn = (g.n_pad % 256 == 0) ? 256 : 128;
    if (g.n_pad % bn != 0) return DDIF_ERR_SHAPE;
  }
  p.bn = bn;
  uint32_t cols = 32;
  while ((int)cols < 2 * bn) cols <<= 1;
  p.tmem_cols = cols;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const int a_bytes = 2 * 3 * kCopyRows * p.span;
  const int b_slot = bn * p.span;
  const int n_stat = g.gn_stats ? (int)(g.batch < kMaxStatSamples ? g.batch : kMaxStatSamples) : 0;
  const int misc = 2 * cin * 4 + n_stat * 8 + (8 + 2 * kMaxBSlots) * 8 + 64 + 1024;
  const int budget = 225 * 1024 - a_bytes - misc;
  const int nb_res = 9 * p.nslab;
  if (nb_res <= kMaxBSlots && (int64_t)nb_res * b_slot <= budget) {
    p.resident_b = 1;
    p.b_slots = nb_res;
  } else {
    p.resident_b = 0;
    int s = budget / b_slot;
    if (s > 8) s = 8;
    if (s < 2) return DDIF_ERR_SHAPE;
    p.b_slots = s;
  }
  L.smem_bytes = a_bytes + p.b_slots * b_slot + misc;
  L.grid_y = (int)(g.n_pad / bn);
  const int sms = ddif_sm_count();
  const int gx = (sms + L.grid_y - 1) / L.grid_y;
  L.grid_x = p.num_m_tiles < gx ? p.num_m_tiles : gx;
  L.variant = 1;
  {
    cuuint64_t dims[3] = {(cuuint64_t)g.w_k[0], (cuuint64_t)g.n_pad, (cuuint64_t)g.w_s[0]};
    cuuint64_t strides[2] = {(cuuint64_t)g.w_k[0] * 2, (cuuint64_t)g.n_pad * g.w_k[0] * 2};
    cuuint32_t box[3] = {(cuuint32_t)p.kslab, (cuuint32_t)bn, 1};
    cuuint32_t es[3] = {1, 1, 1};
    const CUtensorMapSwizzle sw = p.span == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : p.span == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = enc(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(g.w[0]), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return DDIF_ERR_DRIVER;
  }
  p.gn_stats = g.gn_stats;
  p.gn_gamma = g.gn_gamma;
  p.gn_beta = g.gn_beta;
  p.gn_eps = (float)g.gn_eps;
  p.gn_act = (int)g.gn_act;
  p.gn_count = (double)cin * p.src_h * p.src_w;
  if (p.gn_stats && (!p.gn_gamma || !p.gn_beta)) return DDIF_ERR_ARG;
  EpiParams& e = p.epi;
  e.bias = g.bias; e.film = g.film; e.film_ld = (int)g.film_ld; e.mod = (const bf16*)g.mod; e.residual = (const bf16*)g.residual;
  e.res_ld = (int)g.res_ld; e.act = (int)g.act; e.out = (bf16*)g.out; e.out_ld = (int)g.out_ld; e.out_nchw = g.out_nchw; e.stats = g.stats;
  e.n_valid = (int)g.n_valid; e.batch = p.batch; e.out_h = p.out_h; e.out_w = p.out_w;
  if (e.out && (e.out_ld % 8 != 0)) return DDIF_ERR_SHAPE;
  if (e.mod && (e.n_valid % 8 != 0)) return DDIF_ERR_SHAPE;
  if (e.residual && (e.res_ld % 8 != 0)) return DDIF_ERR_SHAPE;
  return DDIF_OK;
}

int conv3_set_debug_ts(long long* ptr) {
  DDIF_CUDA_CHECK(cudaMemcpyToSymbol(g_debug_ts, &ptr, sizeof(ptr)));
  return DDIF_OK;
}

int conv3_launch(const GemmLaunch& L, cudaStream_t stream) {
  const Conv3KParams& p = *reinterpret_cast<const Conv3KParams*>(L.kparams);
  conv3x3_fused_tc_kernel<<<dim3(L.grid_x, L.grid_y), kC3Threads, L.smem_bytes, stream>>>(p);
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

}  // namespace ddif
