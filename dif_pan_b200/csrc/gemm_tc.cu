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
#include "epilogue.cuh"

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
  uint32_t tmem_cols;      // `groups` accumulators of bn columns, rounded to a power of two >= 32
  int groups;              // epilogue warp groups working on different tiles (generic: 2 or 4; lean: 1)
  int naccs;               // TMEM accumulators (generic: = groups; lean: 2)
  int lean;                // kLean* flags of the compile-time epilogue, 0 = generic
  // kLeanOpTma: the epilogue operands of a tile (CSM modulation maps scale | shift, or the residual) are brought into a shared-memory ring by
  // the TMA producer instead of per-thread global loads (see the enum below)
  CUtensorMap tmOp;        // mod: [B, H, W, 2N] box 64 ch x tw x th x tn (SWIZZLE_128B); residual: [B, H, W, res_ld] box N ch x tw x th x tn
  int op_stages;           // ring depth (multiple of `groups`), 0 = operands via global loads
  int op_slabs;            // TMA boxes per tile: mod 2N / 64, residual 1
  int op_span;             // bytes per row of one slab: 128 (SWIZZLE_128B), 64 (SWIZZLE_64B) or 32 (SWIZZLE_32B)
  int op_chan0;            // first channel of the residual inside its tensor (n_tile * bn is added)
  uint32_t op_bytes;       // bytes of one ring slot = op_slabs * 128 rows * op_span
  uint32_t op_off, bar_off;  // shared-memory offsets of the operand ring and of the barriers
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

// CONTIGUOUS M-tile range of this CTA: [base, base + count) (sequential memory walk; see conv3x3_halo.cu halo_range)
__device__ __forceinline__ void gemm_range(int num_m_tiles, int& base, int& count) {
  const int G = (int)gridDim.x, c = (int)blockIdx.x;
  const int q = num_m_tiles / G, r = num_m_tiles - q * G;
  base = c * q + (c < r ? c : r);
  count = q + (c < r ? 1 : 0);
}

// Persistent, warp-specialised: each CTA owns one N tile (blockIdx.y) and walks its contiguous range of M tiles.
// Three pipelines: smem full/empty ring (TMA <-> MMA) running ACROSS tiles, TMEM full/empty (MMA <-> epilogue, two
// accumulators so tile i's epilogue overlaps tile i+1's MMAs), and the static tile walk.
// F = 0: generic run-time epilogue (epilogue.cuh), tile groups.  F & 1: lean epilogue for N <= 64 with 32-byte aligned rows --
// compile-time operand set (kLeanMod / kLeanRes / kLeanStats), one 16-channel chunk per thread, all 16 warps on one tile, and
// the NEXT tile's residual / modulation vectors requested before the current tile is processed (the generic path exposed
// one L2/DRAM round trip per tile and chunk: 80 us instead of 30 for the 64x64-level 1x1 convs,
// profiles/r01s4_gemm1x1_ablation.txt).
// F & 16 (kLeanOpTma, round 2): with kLeanMod / kLeanRes the operands of a tile are NOT fetched by the epilogue threads (32 bytes per thread at a
// 64-256 byte pixel pitch, one tile ahead: ncu showed the epilogue warps stalled on exactly those loads -- long_scoreboard 4.3 per issue,
// DRAM at 38 % -- profiles/r02_ncu_prof_igemm_xconv_*) but by the TMA producer, as swizzled boxes into a 4-deep shared-memory ring
// (full / empty mbarriers per slot, one slot per M tile); the epilogue reads them with conflict-free LDS.128.
enum : int { kLean = 1, kLeanMod = 2, kLeanRes = 4, kLeanStats = 8, kLeanOpTma = 16 };
// shared-memory address of 16-byte chunk `chunk` of row `row` inside a TMA box whose rows are `span` bytes (hardware swizzle of that span)
__device__ __forceinline__ uint32_t op_addr(uint32_t base, int row, int chunk, int span) {
  const int sw = span == 128 ? (row & 7) : (span == 64 ? ((row >> 1) & 3) : ((row >> 2) & 1));
  return base + (uint32_t)(row * span) + ((uint32_t)(chunk ^ sw) << 4);
}
__device__ __forceinline__ void lds256(uint32_t a0, uint32_t a1, uint32_t* r) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a0) : "memory");
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a1) : "memory");
}

template <int F>
struct LeanOperands {
  uint32_t res[(F & kLeanRes) ? 8 : 1];
  uint32_t sc[(F & kLeanMod) ? 8 : 1];
  uint32_t sh[(F & kLeanMod) ? 8 : 1];
};

template <int F>
__global__ void __launch_bounds__(kGemmThreads, 1) conv_igemm_tc_kernel(const __grid_constant__ GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int a_stage_bytes = 128 * p.span;
  const int b_slot_bytes = p.bn * p.span;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + (size_t)p.stages * a_stage_bytes;
  uint8_t* smem_op = smem + p.op_off;                  // [op_stages][op_bytes] epilogue operand ring (kLeanOpTma), 1024-byte aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + p.stages;
  uint64_t* tmem_full_bar = bars + 2 * p.stages;       // [4]
  uint64_t* tmem_empty_bar = bars + 2 * p.stages + 4;  // [4]
  uint64_t* op_full = bars + 2 * p.stages + 8;         // [4] operand slot landed
  uint64_t* op_empty = bars + 2 * p.stages + 12;       // [4] operand slot read by the tile group's warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.y;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  int tile_base, tile_count;
  gemm_range(p.num_m_tiles, tile_base, tile_count);
  const int tile_end = tile_base + tile_count;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) {
      tma_prefetch_desc(&p.tmA[s]);
      tma_prefetch_desc(&p.tmB[s]);
    }
    if (p.op_stages) tma_prefetch_desc(&p.tmOp);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < p.stages; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < p.naccs; ++i) {
        mbar_init(&tmem_full_bar[i], 1);
        mbar_init(&tmem_empty_bar[i], kEpiWarps / p.groups);  // one arrive per epilogue warp of the group that drains it
      }
      for (int i = 0; i < p.op_stages; ++i) {
        mbar_init(&op_full[i], 1);
        mbar_init(&op_empty[i], kEpiWarps / p.groups);
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
  pdl_wait();  // prologue above overlaps the previous kernel's tail; everything below reads what earlier kernels wrote

  int total_iters = 0;
  for (int s = 0; s < p.nseg; ++s) total_iters += p.taps[s] * p.kchunks[s];

  if (warp == 0) {
    // ===================== TMA producer =====================
    // NOTE: this is ONE thread; its instruction count per K-iteration bounds the whole kernel, so the loop keeps
    // running (stage, phase) counters instead of dividing, and walks dy/dx incrementally.
    if (lane == 0) {
      const uint32_t a_bytes = (uint32_t)(p.a_rows * p.span);
      const uint32_t tx_ab = a_bytes + (uint32_t)b_slot_bytes;
      const uint32_t nstages = (uint32_t)p.stages;
      uint32_t stage = 0, phase = 0;
      uint32_t ostage = 0, ophase = 0;
      uint8_t* a_dst = smem_a;
      bool first = true;
      int img_grp = tile_base / tiles_per_img;
      int t_in = tile_base - img_grp * tiles_per_img;
      const int step_grp = 0, step_in = 1;
      for (int m_tile = tile_base; m_tile < tile_end; ++m_tile, first = false) {
        const int trow = t_in / p.tiles_x;
        const int ty0 = trow * p.th;
        const int tx0 = (t_in - trow * p.tiles_x) * p.tw;
        const int n0 = img_grp * p.tn;
        const bool load_b = !p.resident_b || first;
        const uint32_t tx_bytes = load_b ? tx_ab : a_bytes;
        if (p.op_stages) {  // this tile's epilogue operands (stride 1: output pixel box == the box of the operand tensor)
          mbar_wait(&op_empty[ostage], ophase ^ 1u);
          mbar_expect_tx(&op_full[ostage], (uint32_t)(p.op_slabs * p.a_rows * p.op_span));
          uint8_t* dst = smem_op + (size_t)ostage * p.op_bytes;
          const uint32_t slab_bytes = 128u * (uint32_t)p.op_span;
          for (int sl = 0; sl < p.op_slabs; ++sl)
            tma_load_4d(&p.tmOp, &op_full[ostage], dst + (size_t)sl * slab_bytes, p.op_chan0 + n_tile * p.bn + sl * 64, tx0, ty0, n0);
          if (++ostage == (uint32_t)p.op_stages) { ostage = 0; ophase ^= 1u; }
        }
        int j = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const int ntaps = p.taps[s], nkc = p.kchunks[s];
          const int cw0 = tx0 * p.stride - p.pad[s], ch0 = ty0 * p.stride - p.pad[s];
          int dy = 0, dx = 0;
          for (int tap = 0; tap < ntaps; ++tap) {
            const int bz = p.b_per_sample[s] ? n0 : tap;
            for (int kc = 0; kc < nkc; ++kc, ++j) {
              mbar_wait(&empty_bar[stage], phase ^ 1u);
#ifdef DDIF_VAR_G_NO_TMA
              mbar_arrive(&full_bar[stage]);
#else
              mbar_expect_tx(&full_bar[stage], tx_bytes);
              tma_load_4d(&p.tmA[s], &full_bar[stage], a_dst, kc * p.bk, cw0 + dx, ch0 + dy, n0);
              if (load_b) {
                const int slot = p.resident_b ? j : (int)stage;
                tma_load_3d(&p.tmB[s], &full_bar[stage], smem_b + (size_t)slot * b_slot_bytes, kc * p.bk, n_tile * p.bn, bz);
              }
#endif
              a_dst += a_stage_bytes;
              if (++stage == nstages) {
                stage = 0;
                phase ^= 1u;
                a_dst = smem_a;
              }
            }
            if (++dx == 3) {
              dx = 0;
              ++dy;
            }
          }
        }
        t_in += step_in;
        img_grp += step_grp;
        if (t_in >= tiles_per_img) {
          t_in -= tiles_per_img;
          ++img_grp;
        }
      }
      pdl_trigger();  // all loads of this CTA are in flight: the next kernel's CTAs may take over SMs as ours exit
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, elected lane issues; see common.cuh) =====================
    {
      const uint32_t sbo = 8u * (uint32_t)p.span;
      const int ksteps = p.bk / 16;
      const uint32_t nstages = (uint32_t)p.stages;
      const uint64_t desc_a0 = make_smem_desc(smem_u32(smem_a), sbo, p.layout_type);
      const uint64_t desc_b0 = make_smem_desc(smem_u32(smem_b), sbo, p.layout_type);
      const uint32_t a_step = (uint32_t)a_stage_bytes >> 4, b_step = (uint32_t)b_slot_bytes >> 4;
      const bool resident = p.resident_b != 0;
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      const uint32_t ngroups = (uint32_t)p.naccs;
      for (int m_tile = tile_base; m_tile < tile_end; ++m_tile) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * (uint32_t)p.bn;
        uint32_t b_res = 0;
        for (int j = 0; j < total_iters; ++j) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = desc_a0 + (uint64_t)(stage * a_step);
          const uint64_t db = desc_b0 + (uint64_t)(resident ? b_res : stage * b_step);
          const uint32_t accum = j != 0 ? 1u : 0u;
#ifndef DDIF_VAR_G_NO_MMA
          if (ksteps == 4) umma_bf16_ss_steps<4>(tmem_d, da, db, p.idesc, accum);
          else if (ksteps == 2) umma_bf16_ss_steps<2>(tmem_d, da, db, p.idesc, accum);
          else umma_bf16_ss_steps<1>(tmem_d, da, db, p.idesc, accum);
#endif
          umma_commit_elect(&empty_bar[stage]);
          b_res += b_step;
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_elect(&tmem_full_bar[acc]);
        if (++acc == ngroups) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..17) =====================
    // warp -> (TMEM lane quarter q = warp % 4, tile group grp, column group cg); see epilogue.cuh.  Group g owns TMEM
    // accumulator g and the CTA's tiles g, g + groups, ...: while one group waits for its residual / modulation loads and
    // stores a tile, the others work on the next tiles (a single group left the 1x1 convs latency-bound: ~4000 cycles per
    // 128 x 32 tile, profiles/r01_step_profile_B256_s4.txt x_conv rows).
    const int q = warp & 3;
    const int sub = (warp - 2) >> 2;
    if constexpr ((F & kLean) != 0) {
      constexpr bool kMod = (F & kLeanMod) != 0, kRes = (F & kLeanRes) != 0, kStats = (F & kLeanStats) != 0;
      constexpr bool kOpTma = (F & kLeanOpTma) != 0 && (kMod || kRes);
      // N <= 32 needs only 1-2 of the 4 column groups: the others form further TILE groups (group g drains accumulator g for the
      // CTA's tiles g, g + groups, ...), so two (four) tiles' wait -> tcgen05.ld -> math -> store chains overlap instead of running
      // back to back (a single group left the 64x64-level 1x1 convs latency-bound at ~3400 cycles per 128 x 32 tile).
      const int ncg = 4 / p.groups;            // column groups per tile group
      const int cg = sub % ncg;                // this thread's 16-channel chunk
      const int grp = sub / ncg;               // this thread's tile group
      const int row = q * 32 + lane;
      const int px_per_img = p.tw * p.th;
      const int tn_i = row / px_per_img;
      const int r_in = row - tn_i * px_per_img;
      const int ry = r_in / p.tw, rx = r_in % p.tw;
      const bool active = (cg < (p.bn >> 4)) && (q * 32 < p.a_rows);
      const int ng = cg * 16;
      float bv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) bv[j] = (active && p.bias) ? __ldg(p.bias + ng + j) : 0.f;
      const int step = p.groups;
      LeanOperands<F> nxt;
      bool nxt_ok = false;
      size_t nxt_pix = 0;
      int nxt_b = 0;
      auto locate = [&](int m_tile, size_t& pix, int& b) -> bool {
        const int img_grp = m_tile / tiles_per_img;
        const int t_in = m_tile - img_grp * tiles_per_img;
        const int y = (t_in / p.tiles_x) * p.th + ry;
        const int x = (t_in % p.tiles_x) * p.tw + rx;
        b = img_grp * p.tn + tn_i;
        pix = ((size_t)b * p.out_h + y) * p.out_w + x;
        return active && (row < p.a_rows) && (b < p.batch) && (y < p.out_h) && (x < p.out_w);
      };
      auto fetch = [&](LeanOperands<F>& o, size_t pix, bool ok) {
        if (kRes && ok) ldg256(p.residual + pix * (size_t)p.res_ld + ng, o.res);
        if (kMod && ok) {
          const bf16* m = p.mod + pix * (size_t)(2 * p.n_valid) + ng;
          ldg256(m, o.sc);
          ldg256(m + p.n_valid, o.sh);
        }
      };
      int m_tile = tile_base + grp;
      if (!kOpTma && m_tile < tile_end) {
        nxt_ok = locate(m_tile, nxt_pix, nxt_b);
        fetch(nxt, nxt_pix, nxt_ok);
      }
      const uint32_t op_base = smem_u32(smem_op), op_mask = (uint32_t)p.op_stages - 1u, op_shift = p.op_stages == 4 ? 2u : 1u;
      const uint32_t op_slab_bytes = 128u * (uint32_t)p.op_span;
      for (uint32_t tcount = 0; m_tile < tile_end; m_tile += step, ++tcount) {
        LeanOperands<F> cur;
        bool row_ok;
        size_t pix;
        int b;
        if constexpr (kOpTma) {
          row_ok = locate(m_tile, pix, b);
        } else {
          cur = nxt;
          row_ok = nxt_ok;
          pix = nxt_pix;
          b = nxt_b;
          if (m_tile + step < tile_end) {
            nxt_ok = locate(m_tile + step, nxt_pix, nxt_b);
            fetch(nxt, nxt_pix, nxt_ok);  // next tile's operands travel while this tile is processed
          }
        }
        float fv[16];
        if (p.film) {
          const float* f = p.film + (size_t)b * p.film_ld + ng;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 t = row_ok ? __ldg(reinterpret_cast<const float4*>(f + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            fv[j] = t.x; fv[j + 1] = t.y; fv[j + 2] = t.z; fv[j + 3] = t.w;
          }
        }
        // one group: two accumulators alternate per tile; several groups: group g owns accumulator g
        const uint32_t acc = p.groups == 1 ? (tcount & 1u) : (uint32_t)grp;
        mbar_wait(&tmem_full_bar[acc], (p.groups == 1 ? (tcount >> 1) : tcount) & 1u);
        tc_fence_after();
        uint32_t r[16];
        if (active) {
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)p.bn + (uint32_t)ng, r);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        uint32_t op_slot = 0;
        if constexpr (kOpTma) {  // this tile's operands: slot (tile index in the CTA's range) % op_stages of the ring the producer fills
          const uint32_t jt = (uint32_t)(m_tile - tile_base);
          op_slot = jt & op_mask;
          mbar_wait(&op_full[op_slot], (jt >> op_shift) & 1u);
          if (row_ok) {
            const uint32_t ob = op_base + op_slot * p.op_bytes;
            if (kMod) {  // channels [ng, ng + 16) of scale and [N + ng, N + ng + 16) of shift inside the 64-channel slabs
              const int g1 = p.n_valid + ng;
              lds256(op_addr(ob + (uint32_t)(ng >> 6) * op_slab_bytes, row, (ng & 63) >> 3, 128), op_addr(ob + (uint32_t)(ng >> 6) * op_slab_bytes, row, ((ng & 63) >> 3) + 1, 128), cur.sc);
              lds256(op_addr(ob + (uint32_t)(g1 >> 6) * op_slab_bytes, row, (g1 & 63) >> 3, 128), op_addr(ob + (uint32_t)(g1 >> 6) * op_slab_bytes, row, ((g1 & 63) >> 3) + 1, 128), cur.sh);
            }
            if (kRes) lds256(op_addr(ob, row, ng >> 3, p.op_span), op_addr(ob, row, (ng >> 3) + 1, p.op_span), cur.res);
          }
        }
        float s1 = 0.f, s2 = 0.f;
        if (row_ok) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) + bv[j];
          if (p.film) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += fv[j];
          }
          if (kMod) {
            float sc[16], sh[16];
            unpack16(cur.sc, sc);
            unpack16(cur.sh, sh);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = v[j] * (1.0f + sc[j]) + sh[j];
          }
          if (kRes) {
            float rr[16];
            unpack16(cur.res, rr);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += rr[j];
          }
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = swish_half(0.5f * v[j]);
          }
          if (kStats) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              s1 += v[j];
              s2 = fmaf(v[j], v[j], s2);
            }
          }
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            w[j] = *reinterpret_cast<const uint32_t*>(&t);
          }
#ifndef DDIF_VAR_G_NO_STG
          stg256(p.out + pix * (size_t)p.out_ld + ng, w);
#endif
        }
        if constexpr (kOpTma) {  // the operand values have been consumed: hand the slot back to the producer
          __syncwarp();
          if (lane == 0) mbar_arrive(&op_empty[op_slot]);
        }
#ifndef DDIF_VAR_G_NO_STATRED
        if (kStats && active) {  // all rows of one warp belong to one sample
          s1 = warp_sum(s1);
          s2 = warp_sum(s2);
          const int img_grp = m_tile / tiles_per_img;
          const int stat_sample = img_grp * p.tn + (q * 32) / px_per_img;
          if (lane == 0 && stat_sample < p.batch) {
            atomicAdd(p.stats + 2 * (size_t)stat_sample, (double)s1);
            atomicAdd(p.stats + 2 * (size_t)stat_sample + 1, (double)s2);
          }
        }
#else
        if (kStats && s1 + s2 == 1.2345f) atomicAdd(p.stats, 1.0);
#endif
      }
      tc_fence_before();
    } else {
    const int grp = p.groups == 4 ? sub : (sub & 1);
    const int cg = p.groups == 4 ? 0 : (sub >> 1);
    const int row = q * 32 + lane;
    const int px_per_img = p.tw * p.th;
    const int tn_i = row / px_per_img;
    const int r_in = row - tn_i * px_per_img;
    const int ry = r_in / p.tw, rx = r_in % p.tw;
    const bool active = (cg < (p.bn >> 4)) && (q * 32 < p.a_rows);
    EpiParams e{p.bias, p.film, p.film_ld, p.mod, p.residual, p.res_ld, p.act, p.out, p.out_ld, p.out_nchw, p.stats,
                p.n_valid, p.batch, p.out_h, p.out_w};
    uint32_t tcount = 0;
    for (int m_tile = tile_base + grp; m_tile < tile_end; m_tile += p.groups, ++tcount) {
      const int img_grp = m_tile / tiles_per_img;
      const int t_in = m_tile - img_grp * tiles_per_img;
      const int y = (t_in / p.tiles_x) * p.th + ry;
      const int x = (t_in % p.tiles_x) * p.tw + rx;
      const int n0 = img_grp * p.tn;
      const int b = n0 + tn_i;
      const bool row_ok = active && (row < p.a_rows) && (b < p.batch) && (y < p.out_h) && (x < p.out_w);
      const size_t pix = ((size_t)b * p.out_h + y) * p.out_w + x;
      EpiPrefetch pf;
      epilogue_prefetch<4>(e, pf, n_tile, p.bn, cg, row_ok, pix);
      mbar_wait(&tmem_full_bar[grp], tcount & 1u);
      tc_fence_after();
      const uint32_t tm = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)grp * (uint32_t)p.bn;
      if (p.groups == 4)
        epilogue_tile<1>(e, pf, tm, &tmem_empty_bar[grp], p.bn, n_tile, cg, lane, active, row_ok, b, y, x, pix, n0 + (q * 32) / px_per_img);
      else
        epilogue_tile<2>(e, pf, tm, &tmem_empty_bar[grp], p.bn, n_tile, cg, lane, active, row_ok, b, y, x, pix, n0 + (q * 32) / px_per_img);
    }
    tc_fence_before();
    }
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

PFN_encodeTiled ddif_get_encode();
int ddif_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      sms = n;
    else
      sms = 148;
  }
  return sms;
}
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

PFN_encodeTiled ddif_get_encode() { return get_encode(); }

static CUtensorMapSwizzle swizzle_for_span(int span) {
  return span == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : span == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

typedef void (*IgemmKernel)(const GemmKParams);
static IgemmKernel igemm_kernel(int f) {
  switch (f) {
    case 0: return conv_igemm_tc_kernel<0>;
    case kLean: return conv_igemm_tc_kernel<kLean>;
    case kLean | kLeanMod: return conv_igemm_tc_kernel<kLean | kLeanMod>;
    case kLean | kLeanRes: return conv_igemm_tc_kernel<kLean | kLeanRes>;
    case kLean | kLeanStats: return conv_igemm_tc_kernel<kLean | kLeanStats>;
    case kLean | kLeanMod | kLeanStats: return conv_igemm_tc_kernel<kLean | kLeanMod | kLeanStats>;
    case kLean | kLeanRes | kLeanStats: return conv_igemm_tc_kernel<kLean | kLeanRes | kLeanStats>;
    case kLean | kLeanOpTma | kLeanMod: return conv_igemm_tc_kernel<kLean | kLeanOpTma | kLeanMod>;
    case kLean | kLeanOpTma | kLeanRes: return conv_igemm_tc_kernel<kLean | kLeanOpTma | kLeanRes>;
    case kLean | kLeanOpTma | kLeanMod | kLeanStats: return conv_igemm_tc_kernel<kLean | kLeanOpTma | kLeanMod | kLeanStats>;
    case kLean | kLeanOpTma | kLeanRes | kLeanStats: return conv_igemm_tc_kernel<kLean | kLeanOpTma | kLeanRes | kLeanStats>;
    default: return nullptr;
  }
}

int gemm_prepare(const ddif_gemm_t& g, GemmLaunch& L) {
  if (g.a_up != 0) return DDIF_ERR_SHAPE;  // reserved: nearest x2 runs as its own kernel (DDIF_OP_UPSAMPLE2X) in front of the conv
  if (g.a_softmax_h) return cs_gemm_prepare(g, L);  // softmax over the image height fused into the loader (DDIF_ERR_SHAPE if unsupported)
  if (!g.force_tma && conv3_halo_applicable(g)) {
#ifdef DDIF_VAR_NO_HALO  // tuning build (tools/): plain 3x3 convs through the generic TMA kernel
    if (g.nseg == 2 || g.gn_stats || g.dw_w)
#endif
      return conv3_halo_prepare(g, L);
  }
  if (g.gn_stats != nullptr || g.dw_w != nullptr) return DDIF_ERR_SHAPE;  // the fused prologues exist only in the halo kernel
  L.variant = 0;
  GemmKParams& p = *reinterpret_cast<GemmKParams*>(L.kparams);
  static_assert(sizeof(GemmKParams) <= sizeof(L.kparams), "kparams buffer too small");
  memset(&p, 0, sizeof(p));
  PFN_encodeTiled enc = get_encode();
  if (!enc) return DDIF_ERR_DRIVER;
  static bool smem_attr_set = false;  // outside any stream capture: gemm_prepare runs at plan-build time
  if (!smem_attr_set) {
    for (int f = 0; f < 32; ++f) {
      IgemmKernel k = igemm_kernel(f);
      if (k) DDIF_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
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
  const bool lean = bn <= 64 && g.n_pad == bn && g.n_valid % 16 == 0 && g.out && !g.out_nchw && g.out_ld % 16 == 0 &&
                    (!g.residual || g.res_ld % 16 == 0) && (!g.film || g.film_ld % 4 == 0) && !(g.mod && g.residual);
  p.lean = lean ? (kLean | (g.mod ? kLeanMod : 0) | (g.residual ? kLeanRes : 0) | (g.stats ? kLeanStats : 0)) : 0;
  // epilogue operands through a TMA-fed shared-memory ring (kLeanOpTma): stride-1 convs whose operand rows are whole swizzle spans
  int op_slot_bytes = 0;
#ifndef DDIF_VAR_G_NO_OPTMA  // A/B build: per-thread global loads one tile ahead (round 1)
  if (lean && g.stride == 1 && g.mod && g.n_valid % 32 == 0 && (reinterpret_cast<uintptr_t>(g.mod) & 15u) == 0) {
    p.op_slabs = (int)(2 * g.n_valid / 64);
    p.op_span = 128;
  }
#ifdef DDIF_VAR_G_RES_OPTMA  // the residual through the same ring: implemented and correct, but a wash in a same-box A/B (64 -> 32 @64^2 61.4 -> 55.7 us,
  // 32 -> 32 @64^2 44.6 -> 53.5 us, the UNet's attn_out GEMMs 0.545 -> 0.538 ms per step), so the product keeps the one-tile-ahead global loads
  else if (lean && g.stride == 1 && g.residual && !g.mod && (bn == 16 || bn == 32 || bn == 64) && g.n_valid == bn &&
           (reinterpret_cast<uintptr_t>(g.residual) & 15u) == 0) {
    p.op_slabs = 1;
    p.op_span = bn * 2;
  }
#endif
  if (p.op_slabs) {
    op_slot_bytes = p.op_slabs * 128 * p.op_span;
    p.lean |= kLeanOpTma;
  }
#endif
  // lean: N = 64 (48) -> all 16 warps on one tile, two alternating accumulators; N = 32 -> 2 tile groups; N = 16 -> 4 tile groups
  p.groups = lean ? (bn <= 16 ? 4 : (bn <= 32 ? 2 : 1)) : (4 * bn <= 512 ? 4 : 2);
  p.naccs = lean ? (p.groups == 1 ? 2 : p.groups) : p.groups;
  uint32_t cols = 32;
  while ((int)cols < p.naccs * bn) cols <<= 1;
  p.tmem_cols = cols;
  // instruction descriptor: D=f32, A=B=bf16, K-major both, N>>3 at bit 17, M>>4 at bit 24
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  int total_iters = 0;
  for (int s = 0; s < g.nseg; ++s) total_iters += (int)(g.taps[s] * (g.a_c[s] / bk));
  const int a_stage = 128 * p.span, b_slot = bn * p.span;
  // operand ring: 4 slots when they leave >= 4 A stages, else 2 (the slot count must be a multiple of the tile groups sharing the ring)
  if (op_slot_bytes) {
    p.op_stages = (200 * 1024 - 4 * op_slot_bytes >= 4 * (a_stage + b_slot)) ? 4 : 2;
    if (p.op_stages < p.groups) p.op_stages = p.groups;
    if (p.op_stages != 2 && p.op_stages != 4) return DDIF_ERR_SHAPE;
  }
  const int budget = 200 * 1024 - p.op_stages * op_slot_bytes;
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
#ifndef DDIF_VAR_GEMM_MAX_STAGES  // tuning builds (tools/) pass -DDDIF_VAR_GEMM_MAX_STAGES=n
#define DDIF_VAR_GEMM_MAX_STAGES 8
#endif
  if (stages > DDIF_VAR_GEMM_MAX_STAGES) stages = DDIF_VAR_GEMM_MAX_STAGES;
  if (stages < 2) return DDIF_ERR_SHAPE;
  if (!p.resident_b) p.b_slots = stages;
  p.stages = stages;
  p.num_m_tiles = p.tiles_x * p.tiles_y * (int)ceil_div(B, tn);
  p.op_off = (uint32_t)((stages * a_stage + p.b_slots * b_slot + 1023) & ~1023);
  p.op_bytes = (uint32_t)op_slot_bytes;
  p.bar_off = p.op_off + (uint32_t)(p.op_stages * op_slot_bytes);
  L.smem_bytes = (int)p.bar_off + (2 * stages + 18) * 8 + 1024;
  if (p.op_stages) {  // tensor map of the epilogue operand: pixel box of one M tile, rows = tw * th * tn in the TMEM lane order
    const bool is_mod = g.mod != nullptr;
    const cuuint64_t ch = is_mod ? (cuuint64_t)(2 * g.n_valid) : (cuuint64_t)g.n_valid;
    const cuuint64_t ld = is_mod ? ch : (cuuint64_t)g.res_ld;
    cuuint64_t dims[4] = {ch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)(p.op_span / 2), (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&p.tmOp, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(is_mod ? g.mod : g.residual), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_span(p.op_span), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return DDIF_ERR_DRIVER;
    p.op_chan0 = 0;
  }
  const int sms = ddif_sm_count();
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
  if (L.variant == 2) return conv3_halo_launch(L, stream);
  if (L.variant == 3) return cs_gemm_launch(L, stream);
  const GemmKParams& p = *reinterpret_cast<const GemmKParams*>(L.kparams);
  IgemmKernel k = igemm_kernel(p.lean);
  if (!k) return DDIF_ERR_STATE;
  DDIF_CUDA_CHECK(launch_pdl(k, dim3(L.grid_x, L.grid_y), dim3(kGemmThreads), (size_t)L.smem_bytes, stream, p));
  return DDIF_OK;
}

}  // namespace ddif
