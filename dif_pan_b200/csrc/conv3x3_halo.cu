// 3x3 (stride 1, pad 1) convolution with an optional fused GroupNorm(1 group)+Swish prologue: ONE halo tile in shared
// memory feeds all nine taps.  tcgen05 / TMEM / TMA, persistent, warp-specialised.
//
// Replaces `Block` = GN -> Swish -> Conv3x3 (/root/reference/models/sr3_dwt.py:288-300), the FWM ffn 3x3 convs
// (:529-531) and downs[0] (:86).
//
// Output tile = 8 (x) x 16 (y) pixels = the 128 rows of one UMMA; its halo = 10 x 18 pixels.  One TMA 4D box
// (C_slab x 10 x 18 x 1, hardware swizzle, out-of-bounds = zero padding) brings the halo into a stage of a 3-4 deep
// ring: halo pixel (hy, hx) is row r = hy*10 + hx of `span` = C_slab*2 bytes.  Tap (dy, dx) of the implicit GEMM is the
// SAME stage read through a descriptor that starts (dy*10 + dx) rows further and strides SBO = 10 rows between the
// 8-row groups (8 pixels of one tile line): the UMMA swizzle is a function of the absolute shared-memory address
// (tools/microbench/umma_offset.cu, profiles/r01_microbench_umma_offset.txt: every start row / SBO reads back exactly),
// so no shifted copies are needed and each activation byte crosses L2->SMEM once (1.4x halo overhead instead of 9x).
//
// With the GN prologue, 8 transform warps normalise the landed stage IN PLACE (LDS.128 -> fma + tanh.approx -> STS.128;
// padding pixels stay zero, as the reference pads the normalised tensor) and hand it to the MMA warp through a second
// mbarrier; without it the TMA barrier feeds the MMA warp directly.  Weights of all taps stay resident in shared
// memory; one elected thread issues the MMAs into two TMEM accumulators; 8 epilogue warps run epilogue.cuh.
//
// Measured floors (profiles/r01_microbench_*.txt): one M128xNxK16 MMA = max(44.8, N/2) cycles, so a 32->32 tile
// (18 MMAs) needs >= 810 cycles; the TMA halo feed needs 420 (C=32) / 750 (C=64) cycles per tile.
#include "common.cuh"
#include "ddif_internal.h"
#include "epilogue.cuh"

namespace ddif {

static constexpr int kHxfThreads = 256;                    // transform warps 0..7: two groups of 128, alternating stages
static constexpr int kHxfGroup = kHxfThreads / 2;
static constexpr int kHEpiWarps = 8;                       // warps 10..17: two groups of 4, one TMEM accumulator each
#ifdef DDIF_VAR_HALO_2PROD  // tuning build: one TMA producer warp per half-pipeline ring (warps 18, 19)
static constexpr int kHProducers = 2;
#else
static constexpr int kHProducers = 1;
#endif
static constexpr int kHThreads = kHxfThreads + 64 + 32 * kHEpiWarps + 32 * kHProducers;  // + MMA warps 8, 9 + TMA producer warp(s) 18(, 19)
static constexpr int kHW = 10, kHH = 18, kHPx = kHW * kHH;  // halo of an 8 x 16 tile
#ifndef DDIF_VAR_HALO_MAX_STAGES  // tuning builds (tools/) pass -DDDIF_VAR_HALO_MAX_STAGES=n
#define DDIF_VAR_HALO_MAX_STAGES 12
#endif
static constexpr int kHMaxStages = DDIF_VAR_HALO_MAX_STAGES;
static constexpr int kHMaxStat = 1024;
// Negative results of rounds 1-2, removed from the source (numbers in DESIGN.md section 4.2): a TMA-store epilogue (tile staged in shared
// memory, cp.async.bulk.tensor store: slower for N >= 64, +4 % for N = 32 -- the staging traffic competes with the tcgen05 operand fetch);
// fp64 running statistics per thread in shared memory instead of the per-tile warp reduction (same-box A/B: 64 -> 64 @32^2 26.5 -> 29.2 us);
// letting the MMA warp wait on the TMA barrier of a residual slab itself instead of the transform group's relay (32 -> 32 @64^2 60.9 -> 64.0 us);
// a TMA-fed shared-memory ring for the epilogue-side residual of the N >= 64 layers (one slot per epilogue group, loaded one tile ahead by the
// group's elected thread): it costs two of the six halo stages of the 64 -> 64 layers and loses (34.9 -> 36.5 us; profiles/r02_halo_*_negative.txt);
// an L2 prefetch (cp.async.bulk.prefetch.tensor) of the halo box 4 / 8 tiles ahead of the loads: 32 -> 32 @64^2 53.2 -> 59.5 us, 64 -> 32 @64^2
// 66.5 -> 69.7 / 85.3 us (profiles/r02s2_halo_sweep.txt); K slabs of 32 for the 64-channel layers (more, smaller stages): 26.0 -> 28.6 us.

struct alignas(64) HaloKParams {
  CUtensorMap tmA[2];  // activations of K segment 0 / 1 (virtual channel concat: torch.cat((x, skip), 1), sr3_dwt.py:212)
  CUtensorMap tmB[2];  // weights of segment 0 / 1
  CUtensorMap tmR;     // residual tile box (L2 prefetch only)
  CUtensorMap tmRA;    // residual-as-operand mode: halo box of the residual tensor (kslab_r x 10 x 18 x 1), see `nslab_r`
  const float* dw_w;   // depthwise mode (see ddif_gemm_t.dw_w): [9][cin] fp32
  int dw_n;            // leading output channels fed by the depthwise result; the rest read the normalised input itself
  int ntap_w;          // weight taps resident per K slab: 9, or 1 in depthwise mode
  uint32_t idesc_q, idesc_r, dw_bytes;  // depthwise mode: instruction descriptors (N = dw_n / bn - dw_n), bytes of one A buffer
  int has_res_map;
  int cin, kslab, nslab, span;
  int nslab0;          // slabs that come from segment 0
  // Residual-as-operand mode (nslab_r > 0): the residual add `conv(x) + r` (sr3_dwt.py:326) runs on the tensor pipe as one more
  // K segment with IDENTITY weights and only the centre tap, accumulated into the same TMEM tile.  The epilogue then has no
  // residual loads (whose L2/HBM latency it could not hide: 25 % of the 64 -> 64 layers) and no residual math.  bf16 x 1.0 in fp32 is exact.
  int nslab_r;         // residual slabs per tile (bn / kslab_r), appended after the `nslab` conv slabs; nslab_t = nslab + nslab_r
  int nslab_t, kslab_r, span_r, ksteps_r;
  uint32_t layout_r, r_slot_bytes;
  int batch, out_h, out_w;
  int tiles_x, tiles_y, num_tiles;
  int bn;
  int stages;
  uint32_t stage_bytes, b_slot_bytes;
  uint32_t idesc, layout_type, tmem_cols;
  int nacc;            // TMEM accumulators: 2 (one per half-pipeline) or 4 (two per half-pipeline: the MMA warp runs one tile ahead of its epilogue group)
  int tile_w_log2, tile_h;  // output tile = (1 << tile_w_log2) x tile_h pixels, row = y * tile_w + x: 8 x 16 for the halo conv, (128 / H) x H for
                            // the column-softmax GEMM below
  uint32_t a_bytes;         // column-softmax GEMM: bytes of the A slab of a stage (the sample's weight slab follows it)
  int w_per_sample;         // column-softmax GEMM: weights indexed by the sample (FWM W_eff) instead of shared
  const double* gn_stats;
  const double* gn_stats2;  // statistics of segment 1 (the GroupNorm runs over the concatenated tensor)
  const float* gn_gamma;
  const float* gn_beta;
  float gn_eps;
  int gn_act;
  double gn_count;
  EpiParams epi;
};

// Debug hook (ddif_debug_set_timestamps): CTA 0 records clock64() at the pipeline hand-offs of its first 64 tiles:
// ts[(role*64 + tile)*4 + k], role 0 = TMA producer, 1 = MMA issuer, 2 = epilogue warp 10 lane 0, 3 = transform thread 0.
__device__ long long* g_halo_ts = nullptr;
// `ts` is read from g_halo_ts ONCE per thread (a probe is then a clock read + one fire-and-forget store).
__device__ __forceinline__ void h_ts(long long* ts, int role, int tile, int k) {
  if (ts && tile < 64) ts[(role * 64 + tile) * 4 + k] = clock64();
}

__device__ __forceinline__ void h_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint4 h_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void h_sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// (mean, rstd) of sample b over the (concatenated) input, from the fp64 (sum, sumsq) pairs of its one or two sources.
__device__ __forceinline__ float2 halo_mean_rstd(const HaloKParams& p, int b) {
  double s = p.gn_stats[2 * b], ss = p.gn_stats[2 * b + 1];
  if (p.gn_stats2) {
    s += p.gn_stats2[2 * b];
    ss += p.gn_stats2[2 * b + 1];
  }
  const double m = s / p.gn_count;
  double var = ss / p.gn_count - m * m;
  if (var < 0) var = 0;
  return make_float2((float)m, rsqrtf((float)var + p.gn_eps));
}

// CONTIGUOUS tile ranges: CTA c owns tiles [base, base + count): a CTA's halo loads, residual reads, stores and statistics atomics
// walk memory sequentially (consecutive tiles of one sample) instead of striding over the whole batch (m = c, c + grid, ...):
// 32 -> 32 @64^2 65.8 -> 60.8 us, 6.22 -> 6.14 ms per denoise step.
__device__ __forceinline__ void halo_range(const HaloKParams& p, int& base, int& count) {
  const int G = (int)gridDim.x, c = (int)blockIdx.x;
  const int q = p.num_tiles / G, r = p.num_tiles - q * G;
  base = c * q + (c < r ? c : r);
  count = q + (c < r ? 1 : 0);
}

// Divide-free walk over tiles m = start, start + stride, ... (< end) with their K slabs.
// PIPELINE OWNERSHIP RULE: the CTA runs two half-pipelines r = 0, 1 (tiles of its walk with even / odd position): stage
// ring r, transform group r, MMA warp r, TMEM accumulator r and epilogue group r.  Every mbarrier is therefore waited on
// by the same thread(s) for each of its phases, in order -- a parity wait issued two phases ahead of the barrier would
// return immediately on the stale phase (the bug a shared ring with alternating consumers has).
struct HaloIter {
  int b, ty, tx, slab, remaining;
  int sb, sy, sx, tiles_x, tiles_y, nslab;
  __device__ __forceinline__ void init(const HaloKParams& p, int start, int stride, int end) {
    const int tpi = p.tiles_x * p.tiles_y;
    tiles_x = p.tiles_x; tiles_y = p.tiles_y; nslab = p.nslab_t;
    b = start / tpi;
    int r = start - b * tpi;
    ty = r / tiles_x; tx = r - ty * tiles_x;
    sb = stride / tpi;
    r = stride - sb * tpi;
    sy = r / tiles_x; sx = r - sy * tiles_x;
    slab = 0;
    remaining = start < end ? ((end - start + stride - 1) / stride) * nslab : 0;
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

// ---- transform warps: GroupNorm affine (+Swish) of a landed stage, in place -------------------------------------------
// Group g (warps 4g..4g+3) serves half-pipeline g: its wait -> LDS -> math -> STS -> fence -> arrive latency chain on one
// stage overlaps the other group's.
// SHORT (images of <= 8 lines, i.e. the 8 x 8 level: the lower half of every 8 x 16 tile is padding): chunks are batched in pairs and the
// batches whose halo rows all lie below the image are skipped (a group-uniform test), a third of the transform at that level.
template <int NCK, bool SHORT = false>
__device__ __forceinline__ void halo_transform_loop(const HaloKParams& p, uint32_t a_base, uint64_t* a_tma, uint64_t* a_ready,
                                                    const float* s_gamma, const float* s_beta, const float2* s_stat, int tid, long long* dts) {
  constexpr int RPP = kHxfGroup / NCK;                 // halo rows per pass of one group
  constexpr int NPASS = (kHPx + RPP - 1) / RPP;        // 12 / 6 / 3 chunks per thread
  constexpr int BATCH = SHORT ? 2 : (NPASS % 6 == 0 ? 6 : 3);
  static_assert(NPASS % BATCH == 0, "pass batching");
  const int grp = tid / kHxfGroup, lt = tid % kHxfGroup;
  const int c = lt % NCK;
  const int r0 = lt / NCK;
  const float hs = p.gn_act ? 0.5f : 1.0f;  // swish(t) = h*tanh(h) + h with h = t/2: the affine is pre-halved
  const uint32_t nst = (uint32_t)p.stages >> 1;  // stages of this group's ring
  // per-thread tables (tile independent): swizzled smem offset and halo (line, column) of each of the thread's chunks
  // one packed word per chunk (soff < 2^15: a stage is < 32 KB): bits 0..15 smem offset, 16..23 hx, 24..31 hy; rows past the halo
  // get hy = 255 (never in bounds).  Two separate tables spilled to local memory at NPASS = 12 under the 96-register cap.
  uint32_t tab[NPASS];
#pragma unroll
  for (int k = 0; k < NPASS; ++k) {
    const int r = r0 + k * RPP;
    const int hy = r / kHW, hx = r - hy * kHW;
    const uint32_t sw = NCK == 8 ? ((uint32_t)r & 7u) : (NCK == 4 ? (((uint32_t)r >> 1) & 3u) : (((uint32_t)r >> 2) & 1u));
    const uint32_t so = (uint32_t)r * (uint32_t)(NCK * 16) + (((uint32_t)c ^ sw) << 4);
    tab[k] = so | (r < kHPx ? (((uint32_t)hy << 24) | ((uint32_t)hx << 16)) : (255u << 24));
  }
  const bool act = p.gn_act != 0;
  const unsigned H = (unsigned)p.out_h, W = (unsigned)p.out_w;
  HaloIter it;
  {
    int base, count;
    halo_range(p, base, count);
    it.init(p, base + grp, 2, base + count);
  }
  a_tma += grp * nst; a_ready += grp * nst;
  a_base += (uint32_t)grp * nst * p.stage_bytes;
  uint32_t stage = 0, phase = 0;
  int cur_b = -1, cur_slab = -1, tcount = 0;
  f32x2 a2[4], d2[4];
  for (; it.remaining > 0; it.next()) {
    if (it.slab >= p.nslab) {  // residual slab (identity-weight K segment): TMA -> MMA untouched, this group only relays the barrier
      mbar_wait(&a_tma[stage], phase);
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&a_ready[stage]);
      if (++stage == nst) { stage = 0; phase ^= 1u; }
      continue;
    }
#ifdef DDIF_VAR_TS2  // probe build: stamp 0 = top of the conv-slab iteration (after a residual relay), 1 = tables ready, 2 = landed, 3 = arrived
    if (tid == 0) h_ts(dts, 3, tcount, 0);
#endif
    if (it.b != cur_b || it.slab != cur_slab) {
      cur_b = it.b; cur_slab = it.slab;
      float mean, rstd;
      if (it.b < kHMaxStat) {
        const float2 mr = s_stat[it.b];
        mean = mr.x; rstd = mr.y;
      } else {
        const float2 mr = halo_mean_rstd(p, it.b);
        mean = mr.x; rstd = mr.y;
      }
      const int ch0 = it.slab * p.kslab + c * 8;
      const float sc = hs * rstd;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a0 = sc * s_gamma[ch0 + 2 * j], a1 = sc * s_gamma[ch0 + 2 * j + 1];
        a2[j] = pk2(a0, a1);
        d2[j] = pk2(fmaf(-mean, a0, hs * s_beta[ch0 + 2 * j]), fmaf(-mean, a1, hs * s_beta[ch0 + 2 * j + 1]));
      }
    }
    const int y0 = it.ty * 16 - 1, x0 = it.tx * 8 - 1;
    const uint32_t sbase = a_base + stage * p.stage_bytes;
#ifdef DDIF_VAR_TS2
    if (tid == 0) h_ts(dts, 3, tcount, 1);
    mbar_wait(&a_tma[stage], phase);
    if (tid == 0) h_ts(dts, 3, tcount, 2);
#else
    if (tid == 0) h_ts(dts, 3, tcount, 0);
    mbar_wait(&a_tma[stage], phase);
    if (tid == 0) h_ts(dts, 3, tcount, 1);
#endif
#ifndef DDIF_VAR_NO_XFORM
    // Branch-free: every chunk is loaded and normalised unconditionally (rows past the halo / padding pixels compute on
    // whatever is there) and only the STORE is predicated, so the BATCH chunks of a thread form independent dependency
    // chains the scheduler can interleave (MUFU.TANH issues at 8 cycles per warp instruction per SM sub-partition:
    // profiles/r01_microbench_sfu_rate.txt); per-chunk divergent regions serialised them.
    const int rlim = ((int)H - y0 < kHH ? (int)H - y0 : kHH) * kHW;  // halo rows >= rlim lie below the image
#pragma unroll
    for (int k0 = 0; k0 < NPASS; k0 += BATCH) {
      if (SHORT && k0 * RPP >= rlim) break;
      uint4 v[BATCH];
      bool ok[BATCH];
#pragma unroll
      for (int k = 0; k < BATCH; ++k) {
        ok[k] = (unsigned)(y0 + (int)(tab[k0 + k] >> 24)) < H && (unsigned)(x0 + (int)((tab[k0 + k] >> 16) & 255u)) < W;
        v[k] = h_lds128(sbase + (tab[k0 + k] & 0xffffu));
      }
#pragma unroll
      for (int k = 0; k < BATCH; ++k) {
        const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          f32x2 t = fma2(bf2_to_f2(w[q]), a2[q], d2[q]);
#ifndef DDIF_VAR_NO_TANH
          if (act) t = swish_half2(t);
#endif
          o[q] = f2_to_bf2(t);
        }
        v[k] = make_uint4(o[0], o[1], o[2], o[3]);
      }
#pragma unroll
      for (int k = 0; k < BATCH; ++k)
        if (ok[k]) h_sts128(sbase + (tab[k0 + k] & 0xffffu), v[k]);
    }
#endif
#ifndef DDIF_VAR_TS2
    if (tid == 0) h_ts(dts, 3, tcount, 2);
#endif
    h_fence_proxy_async();
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&a_ready[stage]);  // one arrive per warp (count kHxfGroup / 32)
    if (tid == 0) h_ts(dts, 3, tcount, 3);
    ++tcount;
    if (++stage == nst) { stage = 0; phase ^= 1u; }
  }
}

// ---- MMA warps -------------------------------------------------------------------------------------------------------
// Two issuing warps: warp w takes tiles t = w, w+2, ... into TMEM accumulator w.  A tcgen05.mma issue blocks for its
// ~45-cycle slot (profiles/r01_microbench_umma_rate.txt), so a single issuer adds its barrier waits (~600 cycles per
// tile) to the tensor pipe's critical path; with two, one warp waits for its next stage / accumulator while the other's
// MMAs execute.
template <int KSTEPS>
__device__ __forceinline__ void halo_mma_loop(const HaloKParams& p, uint32_t a_base, uint32_t b_base, uint64_t* a_full, uint64_t* a_empty,
                                              uint64_t* b_full, uint64_t* tmem_full, uint64_t* tmem_empty, uint32_t tmem_base, int my_tiles,
                                              int w, long long* dts) {
  const uint32_t span = (uint32_t)p.span;
  const uint64_t desc_b0 = make_smem_desc(b_base, 8u * span, p.layout_type);
  const uint32_t stage16 = p.stage_bytes >> 4, b16 = p.b_slot_bytes >> 4;
  uint32_t tap_off[9];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) tap_off[tap] = ((uint32_t)((tap / 3) * kHW + (tap % 3)) * span) >> 4;
  const uint32_t nst = (uint32_t)p.stages >> 1;  // stages of this warp's ring
  a_full += (uint32_t)w * nst; a_empty += (uint32_t)w * nst;
  a_base += (uint32_t)w * nst * p.stage_bytes;
  const uint64_t desc_a0 = make_smem_desc(a_base, (uint32_t)kHW * span, p.layout_type);  // SBO = one halo line (10 rows)
  // residual-as-operand: centre tap (one line + one pixel into the halo) of a stage holding rows of span_r bytes; identity weights
  // sit behind the conv weights
  const uint32_t span_r = (uint32_t)p.span_r;
  const uint64_t desc_r0 = make_smem_desc(a_base + (uint32_t)(kHW + 1) * span_r, (uint32_t)kHW * span_r, p.layout_r);
  const uint64_t desc_br0 = make_smem_desc(b_base + (uint32_t)(9 * p.nslab) * p.b_slot_bytes, 8u * span_r, p.layout_r);
  uint32_t stage = 0, phase = 0;
  mbar_wait(b_full, 0u);
  tc_fence_after();
  const bool four = p.nacc == 4;
  uint32_t itn = 0;
  for (int t = w; t < my_tiles; t += 2, ++itn) {
    if ((threadIdx.x & 31) == 0) h_ts(dts, 1, t, 0);
    const uint32_t acc = (uint32_t)w + (four ? 2u * (itn & 1u) : 0u);
    const uint32_t tmem_d = tmem_base + acc * (uint32_t)p.bn;
    mbar_wait(&tmem_empty[acc], ((four ? itn >> 1 : itn) & 1u) ^ 1u);
    tc_fence_after();
    if ((threadIdx.x & 31) == 0) h_ts(dts, 1, t, 1);
    uint64_t db = desc_b0;
    for (int slab = 0; slab < p.nslab; ++slab) {
      mbar_wait(&a_full[stage], phase);
      tc_fence_after();
      if ((threadIdx.x & 31) == 0) h_ts(dts, 1, t, 2);
      const uint64_t da = desc_a0 + (uint64_t)(stage * stage16);
#ifndef DDIF_VAR_NO_MMA
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        umma_bf16_ss_steps<KSTEPS>(tmem_d, da + tap_off[tap], db, p.idesc, (slab | tap) != 0 ? 1u : 0u);
        db += b16;
      }
#endif
      umma_commit_elect(&a_empty[stage]);
      if (++stage == nst) { stage = 0; phase ^= 1u; }
    }
    for (int rs = 0; rs < p.nslab_r; ++rs) {  // + residual: centre tap of the residual's halo stage x identity weights
      mbar_wait(&a_full[stage], phase);
      tc_fence_after();
      const uint64_t da = desc_r0 + (uint64_t)(stage * stage16);
      const uint64_t dbr = desc_br0 + (uint64_t)((uint32_t)rs * (p.r_slot_bytes >> 4));
      for (int k = 0; k < p.ksteps_r; ++k) umma_bf16_ss_steps<1>(tmem_d, da + (uint64_t)(2 * k), dbr + (uint64_t)(2 * k), p.idesc, 1u);
      umma_commit_elect(&a_empty[stage]);
      if (++stage == nst) { stage = 0; phase ^= 1u; }
    }
    umma_commit_elect(&tmem_full[acc]);
    if ((threadIdx.x & 31) == 0) h_ts(dts, 1, t, 3);
  }
}

// ---- depthwise mode (FWM q path) ------------------------------------------------------------------------------------------
// q = Conv1x1(DW3x3(x_hat)) and r = attn_res(x_hat) (sr3_dwt.py:509-517,541,573) without the 9x dense-conv FLOPs of the composed
// formulation: the transform group normalises the landed halo stage in place (as above), synchronises, and then computes the
// depthwise 3x3 of the 8 x 16 centre pixels from the normalised halo (fp32, sliding 3x3 window in registers: 3 new LDS.32 per
// output word) into a K-major swizzled A buffer.  The MMA warp then issues ONE tap: q columns from that buffer, r columns from
// the centre-tap view of the stage itself.  Two A buffers per half-pipeline, released by tcgen05.commit.
template <int NCK>
__device__ __forceinline__ uint32_t halo_sw(uint32_t r) {
  return NCK == 8 ? (r & 7u) : (NCK == 4 ? ((r >> 1) & 3u) : ((r >> 2) & 1u));
}

template <int NCK>
__device__ __forceinline__ void halo_transform_dw_loop(const HaloKParams& p, uint32_t a_base, uint32_t dw_base, uint64_t* a_tma, uint64_t* a_ready,
                                                       uint64_t* dw_empty, const float* s_gamma, const float* s_beta, const float2* s_stat,
                                                       const float* s_dw, int tid) {
  constexpr int RPP = kHxfGroup / NCK;
  constexpr int NPASS = (kHPx + RPP - 1) / RPP;
  constexpr int BATCH = NPASS % 6 == 0 ? 6 : 3;
  constexpr int WPR = NCK * 4;            // bf16x2 words per row of the slab: 8 / 16 / 32
  constexpr int NCG = kHxfGroup / WPR;    // column groups: 16 / 8 / 4
  constexpr uint32_t SPAN = NCK * 16;
  const int grp = tid / kHxfGroup, lt = tid % kHxfGroup;
  const int c = lt % NCK;
  const int r0 = lt / NCK;
  const float hs = p.gn_act ? 0.5f : 1.0f;
  const uint32_t nst = (uint32_t)p.stages >> 1;
  uint32_t soff[NPASS];
  int hyx[NPASS];
#pragma unroll
  for (int k = 0; k < NPASS; ++k) {
    const int r = r0 + k * RPP;
    const int hy = r / kHW, hx = r - hy * kHW;
    soff[k] = (uint32_t)r * SPAN + (((uint32_t)c ^ halo_sw<NCK>((uint32_t)r)) << 4);
    hyx[k] = r < kHPx ? ((hy << 8) | hx) : (255 << 8);
  }
  // depthwise work split: word w of the slab row, column group cgp
  const int w = lt % WPR, cgp = lt / WPR;
  const uint32_t wchunk = (uint32_t)w >> 2, wsub = ((uint32_t)w & 3u) << 2;
  constexpr int UNITS = NCG == 4 ? 2 : 1;           // columns per thread
  constexpr int NROWS = NCG == 16 ? 8 : 16;         // output rows per unit
  const int col0 = NCG == 16 ? (cgp & 7) : (NCG == 8 ? cgp : cgp * 2);
  const int row0 = NCG == 16 ? (cgp >> 3) * 8 : 0;
  const bool act = p.gn_act != 0;
  const unsigned H = (unsigned)p.out_h, W = (unsigned)p.out_w;
  HaloIter it;
  {
    int base, count;
    halo_range(p, base, count);
    it.init(p, base + grp, 2, base + count);
  }
  a_tma += grp * nst; a_ready += grp * nst;
  a_base += (uint32_t)grp * nst * p.stage_bytes;
  dw_base += (uint32_t)(grp * 2) * p.dw_bytes;
  dw_empty += grp * 2;
  uint32_t stage = 0, phase = 0, dbuf = 0, dphase = 0;
  int cur_b = -1, cur_slab = -1;
  f32x2 a2[4], d2[4], wt[9];
  for (; it.remaining > 0; it.next()) {
    if (it.b != cur_b || it.slab != cur_slab) {
      if (it.slab != cur_slab) {
        const int ch = it.slab * p.kslab + 2 * w;
#pragma unroll
        for (int t = 0; t < 9; ++t) wt[t] = pk2(s_dw[t * p.cin + ch], s_dw[t * p.cin + ch + 1]);
      }
      cur_b = it.b; cur_slab = it.slab;
      const float2 mr = it.b < kHMaxStat ? s_stat[it.b] : halo_mean_rstd(p, it.b);
      const int ch0 = it.slab * p.kslab + c * 8;
      const float sc = hs * mr.y;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a0 = sc * s_gamma[ch0 + 2 * j], a1 = sc * s_gamma[ch0 + 2 * j + 1];
        a2[j] = pk2(a0, a1);
        d2[j] = pk2(fmaf(-mr.x, a0, hs * s_beta[ch0 + 2 * j]), fmaf(-mr.x, a1, hs * s_beta[ch0 + 2 * j + 1]));
      }
    }
    const int y0 = it.ty * 16 - 1, x0 = it.tx * 8 - 1;
    const uint32_t sbase = a_base + stage * p.stage_bytes;
    mbar_wait(&a_tma[stage], phase);
    // phase 1: GroupNorm affine (+Swish) in place, branch-free (see halo_transform_loop)
#pragma unroll
    for (int k0 = 0; k0 < NPASS; k0 += BATCH) {
      uint4 v[BATCH];
      bool ok[BATCH];
#pragma unroll
      for (int k = 0; k < BATCH; ++k) {
        ok[k] = (unsigned)(y0 + (hyx[k0 + k] >> 8)) < H && (unsigned)(x0 + (hyx[k0 + k] & 255)) < W;
        v[k] = h_lds128(sbase + soff[k0 + k]);
      }
#pragma unroll
      for (int k = 0; k < BATCH; ++k) {
        const uint32_t ww[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          f32x2 t = fma2(bf2_to_f2(ww[q]), a2[q], d2[q]);
          if (act) t = swish_half2(t);
          o[q] = f2_to_bf2(t);
        }
        v[k] = make_uint4(o[0], o[1], o[2], o[3]);
      }
#pragma unroll
      for (int k = 0; k < BATCH; ++k)
        if (ok[k]) h_sts128(sbase + soff[k0 + k], v[k]);
    }
    named_bar_sync(3 + grp, kHxfGroup);  // the whole normalised halo is visible to the group
    // phase 2: depthwise 3x3 of the centre pixels -> A buffer `dbuf`
    mbar_wait(&dw_empty[dbuf], dphase ^ 1u);
    const uint32_t dbase = dw_base + dbuf * p.dw_bytes;
    auto ldw = [&](int hy, int hx) -> f32x2 {
      const uint32_t r = (uint32_t)(hy * kHW + hx);
      uint32_t word;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(word) : "r"(sbase + r * SPAN + ((wchunk ^ halo_sw<NCK>(r)) << 4) + wsub) : "memory");
      return bf2_to_f2(word);
    };
#pragma unroll
    for (int u = 0; u < UNITS; ++u) {
      const int col = col0 + u;  // output column ox; halo columns col, col+1, col+2
      f32x2 ra[3], rb[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        ra[j] = ldw(row0, col + j);
        rb[j] = ldw(row0 + 1, col + j);
      }
#pragma unroll
      for (int oy = 0; oy < NROWS; ++oy) {
        f32x2 rc[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) rc[j] = ldw(row0 + oy + 2, col + j);
        // three independent chains (one per window row) instead of nine dependent FMAs
        f32x2 acc0 = fma2(ra[2], wt[2], fma2(ra[1], wt[1], mul2(ra[0], wt[0])));
        f32x2 acc1 = fma2(rb[2], wt[5], fma2(rb[1], wt[4], mul2(rb[0], wt[3])));
        f32x2 acc2 = fma2(rc[2], wt[8], fma2(rc[1], wt[7], mul2(rc[0], wt[6])));
        const f32x2 acc = add2(add2(acc0, acc1), acc2);
        const uint32_t ro = (uint32_t)((row0 + oy) * 8 + col);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(dbase + ro * SPAN + ((wchunk ^ halo_sw<NCK>(ro)) << 4) + wsub), "r"(f2_to_bf2(acc)) : "memory");
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          ra[j] = rb[j];
          rb[j] = rc[j];
        }
      }
    }
    h_fence_proxy_async();
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&a_ready[stage]);
    if (++stage == nst) { stage = 0; phase ^= 1u; }
    if (++dbuf == 2u) { dbuf = 0; dphase ^= 1u; }
  }
}

template <int KSTEPS>
__device__ __forceinline__ void halo_mma_dw_loop(const HaloKParams& p, uint32_t a_base, uint32_t dw_base, uint32_t b_base, uint64_t* a_full,
                                                 uint64_t* a_empty, uint64_t* dw_empty, uint64_t* b_full, uint64_t* tmem_full, uint64_t* tmem_empty,
                                                 uint32_t tmem_base, int my_tiles, int w) {
  const uint32_t span = (uint32_t)p.span;
  const uint32_t stage16 = p.stage_bytes >> 4, b16 = p.b_slot_bytes >> 4, dw16 = p.dw_bytes >> 4;
  const uint32_t centre = ((uint32_t)(kHW + 1) * span) >> 4;  // centre tap of the halo: one line + one pixel
  const uint32_t nst = (uint32_t)p.stages >> 1;
  a_full += (uint32_t)w * nst; a_empty += (uint32_t)w * nst;
  dw_empty += w * 2;
  a_base += (uint32_t)w * nst * p.stage_bytes;
  dw_base += (uint32_t)(w * 2) * p.dw_bytes;
  const uint64_t desc_s0 = make_smem_desc(a_base, (uint32_t)kHW * span, p.layout_type);  // stage view: SBO = one halo line
  const uint64_t desc_d0 = make_smem_desc(dw_base, 8u * span, p.layout_type);             // depthwise buffer: dense 8-row groups
  const uint64_t desc_bq = make_smem_desc(b_base, 8u * span, p.layout_type);
  const uint64_t desc_br = desc_bq + (uint64_t)(((uint32_t)p.dw_n * span) >> 4);
  const bool has_r = p.dw_n < p.bn;
  uint32_t stage = 0, phase = 0, dbuf = 0;
  mbar_wait(b_full, 0u);
  tc_fence_after();
  const bool four = p.nacc == 4;
  uint32_t itn = 0;
  for (int t = w; t < my_tiles; t += 2, ++itn) {
    const uint32_t acc = (uint32_t)w + (four ? 2u * (itn & 1u) : 0u);
    const uint32_t tmem_d = tmem_base + acc * (uint32_t)p.bn;
    mbar_wait(&tmem_empty[acc], ((four ? itn >> 1 : itn) & 1u) ^ 1u);
    tc_fence_after();
    for (int slab = 0; slab < p.nslab; ++slab) {
      mbar_wait(&a_full[stage], phase);
      tc_fence_after();
      const uint64_t db = (uint64_t)((uint32_t)slab * b16);
      umma_bf16_ss_steps<KSTEPS>(tmem_d, desc_d0 + (uint64_t)(dbuf * dw16), desc_bq + db, p.idesc_q, slab != 0 ? 1u : 0u);
      if (has_r)
        umma_bf16_ss_steps<KSTEPS>(tmem_d + (uint32_t)p.dw_n, desc_s0 + (uint64_t)(stage * stage16 + centre), desc_br + db, p.idesc_r, slab != 0 ? 1u : 0u);
      umma_commit_elect(&a_empty[stage]);
      umma_commit_elect(&dw_empty[dbuf]);
      if (++stage == nst) { stage = 0; phase ^= 1u; }
      dbuf ^= 1u;
    }
    umma_commit_elect(&tmem_full[acc]);
  }
}

// ---- epilogue --------------------------------------------------------------------------------------------------------
// Two groups of 4 warps; group g owns TMEM accumulator g (tiles t = g, g+2, ...), so a warp has two tile periods for one
// tile and the per-tile fixed cost (coordinates, barrier, statistics reduction) is paid once per 32 rows x ALL columns.
// Compile-time flags keep the instruction stream short: the generic run-time epilogue (epilogue.cuh) spent ~2000
// cycles per 128 x 32 tile on constant loads and branches (profiles/r01_halo_v1_*).
enum : int { kEpiRes = 1, kEpiAct = 2, kEpiStats = 4, kEpiNchw = 8 };

template <int F>
__device__ __forceinline__ void halo_epilogue_loop(const HaloKParams& p, uint32_t tmem_base, uint64_t* tmem_full, uint64_t* tmem_empty, int warp,
                                                   int lane, int my_tiles, long long* dts, float* s_add, int grp, int ng, int slot, int tile_first, int tile_step) {
  // tile_first / tile_step: the CTA's tiles are tile_first + k * tile_step, k < my_tiles (contiguous range for the halo conv, grid-strided for the
  // column-softmax GEMM)
  // grp / ng: this warp's tile group and the number of groups (2: warps 10..17; 4: + warps 0..7, which have no GroupNorm transform to do --
  // then group g drains accumulator g alone); slot: this warp's 256-float vector in s_add
  constexpr bool kRes = (F & kEpiRes) != 0, kAct = (F & kEpiAct) != 0, kStats = (F & kEpiStats) != 0, kNchw = (F & kEpiNchw) != 0;
  const int q = warp & 3;
  const int row = q * 32 + lane;
  const int twl = p.tile_w_log2, tile_h = p.tile_h;
  const int ry = row >> twl, rx = row & ((1 << twl) - 1);
  const int nch = p.bn >> 4;
  const int n0 = (int)blockIdx.y * p.bn;           // first output channel of this CTA's N tile
  const int n_valid = p.epi.n_valid - n0;          // valid channels of the tile (may exceed bn: clipped by nch)
  const int n_total = p.epi.n_valid;
  const float* bias = p.epi.bias ? p.epi.bias + n0 : nullptr;
  const float* film = p.epi.film ? p.epi.film + n0 : nullptr;
  const int film_ld = p.epi.film_ld;
  const bf16* resid = p.epi.residual ? p.epi.residual + n0 : nullptr;
  const int res_ld = p.epi.res_ld, out_ld = p.epi.out_ld;
  bf16* outp = p.epi.out ? p.epi.out + n0 : nullptr;
  float* out_nchw = p.epi.out_nchw;
  double* stats = p.epi.stats;
  const int out_h = p.out_h, out_w = p.out_w, tiles_x = p.tiles_x, tiles_y = p.tiles_y;
  const size_t hw = (size_t)out_h * out_w;
  const uint32_t tm_lane0 = tmem_base + ((uint32_t)(q * 32) << 16);
  const bool four = p.nacc == 4;
  // per-warp additive vector (bias, or FiLM row of the tile's sample [+ bias]) in shared memory: with ~200 KB of dynamic
  // smem the L1 is a few KB, so per-tile __ldg of these vectors paid an L2 round trip per 16-channel chunk
  float* addv = s_add + slot * 256;
  const bool has_add = bias != nullptr || film != nullptr;
  const int nadd = (p.bn + 31) >> 5;  // values per lane
  if (!film && bias) {
    for (int k = 0; k < nadd; ++k) {
      const int ch = lane + 32 * k;
      addv[ch] = ch < n_valid ? bias[ch] : 0.f;
    }
    __syncwarp();
  }
  // divide-free walk over this group's tiles of the CTA's contiguous range: start base + grp, stride 2
  int b, ty, tx, sb, sy, sx;
  {
    const int tpi = tiles_x * tiles_y;
    const int m0 = tile_first + grp * tile_step, g2 = ng * tile_step;
    b = m0 / tpi;
    int r = m0 - b * tpi;
    ty = r / tiles_x; tx = r - ty * tiles_x;
    sb = g2 / tpi;
    r = g2 - sb * tpi;
    sy = r / tiles_x; sx = r - sy * tiles_x;
  }
  // Statistics are reduced per TILE (two warp reductions + two fp64 atomics): the per-tile partials are the same values whatever the
  // tile -> CTA assignment, so results do not depend on the batch size.  Carrying per-thread sums across a sample's tiles was tried:
  // plain fp32 running sums are 10 % faster on the 32 -> 32 layers but make the statistics batch-size dependent at the 1e-5 level
  // (= the bf16 noise level of the network output); compensated or fp64 running sums keep the numerics but their four extra live
  // registers spill in this 96-register epilogue and cost more than the reduction they save.
  f32x2 s1 = pk2(0.f, 0.f), s2 = pk2(0.f, 0.f);
  int stat_b = -1;
  auto tile_stats = [&]() {
    float l1, h1, l2, h2;
    upk2(s1, l1, h1);
    upk2(s2, l2, h2);
    const float t1 = warp_sum(l1 + h1), t2 = warp_sum(l2 + h2);
    if (lane == 0) {
      atomicAdd(stats + 2 * (size_t)stat_b, (double)t1);
      atomicAdd(stats + 2 * (size_t)stat_b + 1, (double)t2);
    }
    s1 = pk2(0.f, 0.f);
    s2 = pk2(0.f, 0.f);
  };
  uint32_t it = 0;
  const bool solo = ng == 4;  // four groups: accumulator grp belongs to this group alone
  for (int t = grp; t < my_tiles; t += ng, ++it) {
    const int y = ty * tile_h + ry, x = (tx << twl) + rx;
    const bool row_ok = (y < out_h) && (x < out_w);
    const size_t pix = ((size_t)b * out_h + y) * out_w + x;
    const bf16* res_px = kRes ? resid + pix * (size_t)res_ld : nullptr;
    float fa[8];
    if (film) {  // this tile's FiLM row travels while the MMAs finish
      const float* film_b = film + (size_t)b * film_ld;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int ch = lane + 32 * k;
        if (k < nadd) fa[k] = ch < n_valid ? __ldg(film_b + ch) + (bias ? __ldg(bias + ch) : 0.f) : 0.f;
      }
    }
    uint32_t rs0[8], rs1[8];
#ifdef DDIF_VAR_NO_RESLD
#pragma unroll
    for (int j = 0; j < 8; ++j) rs0[j] = rs1[j] = 0u;
#endif
    if (kRes) {  // first two 16-channel chunks of the residual travel while the MMAs finish
#ifndef DDIF_VAR_NO_RESLD
      if (row_ok) {
        ldg256(res_px, rs0);
        if (nch > 1) ldg256(res_px + 16, rs1);
      }
#endif
    }
    const uint32_t acc = solo ? (uint32_t)grp : (uint32_t)grp + (four ? 2u * (it & 1u) : 0u);
    const uint32_t tm_lane = tm_lane0 + acc * (uint32_t)p.bn;
    uint64_t* empty = &tmem_empty[acc];
    mbar_wait(&tmem_full[acc], (solo ? it : (four ? it >> 1 : it)) & 1u);
    tc_fence_after();
    if (warp == 10 && lane == 0) h_ts(dts, 2, t, 0);
    if (film) {
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < nadd) addv[lane + 32 * k] = fa[k];
      __syncwarp();
    }
    stat_b = b;
    auto process = [&](const uint32_t (&acc)[16], const uint32_t (&rs)[8], int cc) {
      const int ng = cc * 16;
      const int nrem = n_valid - ng;
      if (!row_ok || nrem <= 0) return;
      if (nrem >= 16) {
        f32x2 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = pk2(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
        if (has_add) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 tq = *reinterpret_cast<const float4*>(addv + ng + 4 * j);
            v[2 * j] = add2(v[2 * j], pk2(tq.x, tq.y));
            v[2 * j + 1] = add2(v[2 * j + 1], pk2(tq.z, tq.w));
          }
        }
        if (kRes) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = add2(v[j], bf2_to_f2(rs[j]));
        }
        if (kAct) {
          const f32x2 half2 = pk2(0.5f, 0.5f);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = swish_half2(mul2(v[j], half2));
        }
        if (kStats) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            s1 = add2(s1, v[j]);
            s2 = fma2(v[j], v[j], s2);
          }
        }
        if (kNchw) {
          float* o = out_nchw + ((size_t)b * n_total + n0 + ng) * hw + (size_t)y * out_w + x;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float lo, hi;
            upk2(v[j], lo, hi);
            o[(size_t)(2 * j) * hw] = lo;
            o[(size_t)(2 * j + 1) * hw] = hi;
          }
        } else {
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w[j] = f2_to_bf2(v[j]);
#ifndef DDIF_VAR_NO_STG
          stg256(outp + pix * (size_t)out_ld + ng, w);
#endif
        }
      } else {  // ragged last chunk (n_valid % 16 != 0): scalar path
        float ss1 = 0.f, ss2 = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (j < nrem) {
            float t2 = __uint_as_float(acc[j]);
            if (has_add) t2 += addv[ng + j];
            if (kRes) t2 += __bfloat162float(res_px[ng + j]);
            if (kAct) t2 = swish_half(0.5f * t2);
            if (kStats) { ss1 += t2; ss2 = fmaf(t2, t2, ss2); }
            if (kNchw) out_nchw[((size_t)b * n_total + n0 + ng + j) * hw + (size_t)y * out_w + x] = t2;
            else outp[pix * (size_t)out_ld + ng + j] = __float2bfloat16(t2);
          }
        }
        if (kStats) { s1 = add2(s1, pk2(ss1, 0.f)); s2 = add2(s2, pk2(ss2, 0.f)); }
      }
    };
    for (int cc = 0; cc < nch; cc += 2) {
      uint32_t a0[16], a1[16];
      const bool two = cc + 1 < nch;
#ifndef DDIF_VAR_NO_RESLD
      if (kRes && cc > 0 && row_ok) {
        ldg256(res_px + cc * 16, rs0);
        if (two) ldg256(res_px + cc * 16 + 16, rs1);
      }
#endif
      tmem_ld16(tm_lane + (uint32_t)(cc * 16), a0);
      if (two) tmem_ld16(tm_lane + (uint32_t)(cc * 16 + 16), a1);
      tmem_ld_wait();
      if (warp == 10 && lane == 0 && cc == 0) h_ts(dts, 2, t, 1);
      if (cc + 2 >= nch) {  // last TMEM read of this tile: hand the accumulator back before the math and the stores
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(empty);
      }
#ifndef DDIF_VAR_NO_EPI
      process(a0, rs0, cc);
      if (two) process(a1, rs1, cc + 1);
#endif
    }
    if (warp == 10 && lane == 0) h_ts(dts, 2, t, 2);
    if (kStats) tile_stats();
    if (warp == 10 && lane == 0) h_ts(dts, 2, t, 3);
    tx += sx; ty += sy; b += sb;
    if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
    if (ty >= tiles_y) { ty -= tiles_y; ++b; }
  }
}

template <int F, bool DW>
__global__ void __launch_bounds__(kHThreads, 1) conv3x3_halo_tc_kernel(const __grid_constant__ HaloKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint8_t* smem_a = smem;
  uint8_t* smem_dw = smem + (size_t)p.stages * p.stage_bytes;  // [4][dw_bytes] depthwise A buffers (2 per half-pipeline), only in depthwise mode
  uint8_t* smem_b = smem_dw + (DW ? 4u * p.dw_bytes : 0u);
  uint8_t* smem_id = smem_b + (size_t)(p.ntap_w * p.nslab) * p.b_slot_bytes;  // [nslab_r][bn rows x span_r] identity weights (residual-as-operand)
  float* s_gamma = reinterpret_cast<float*>(smem_id + (size_t)p.nslab_r * p.r_slot_bytes);
  float* s_beta = s_gamma + p.cin;
  float2* s_stat = reinterpret_cast<float2*>(s_beta + p.cin);
  const int n_stat = p.gn_stats ? (p.batch < kHMaxStat ? p.batch : kHMaxStat) : 0;
  float* s_add = reinterpret_cast<float*>(s_stat + ((n_stat + 1) & ~1));  // [8 epilogue warps (16 without the GN prologue)][256], 16-byte aligned
  float* s_dw = s_add + (p.gn_stats ? 1 : 2) * kHEpiWarps * 256;          // [9][cin] depthwise weights (depthwise mode)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_dw + (DW ? ((9 * p.cin + 3) & ~3) : 0));
  uint64_t* a_tma = bars;                       // [stages] TMA landed
  uint64_t* a_ready = bars + kHMaxStages;       // [stages] transformed (count 256)
  uint64_t* a_empty = bars + 2 * kHMaxStages;   // [stages] MMAs done
  uint64_t* tmem_full = bars + 3 * kHMaxStages; // [4]
  uint64_t* tmem_empty = tmem_full + 4;         // [4]
  uint64_t* b_full = tmem_empty + 4;            // [1]
  uint64_t* dw_empty = b_full + 1;              // [4] depthwise A buffer consumed by the MMAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dw_empty + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int tile_base, my_tiles;
  halo_range(p, tile_base, my_tiles);
  const bool gn = p.gn_stats != nullptr;
  long long* const dts = (blockIdx.x == 0 && blockIdx.y == 0) ? g_halo_ts : nullptr;

  if (warp == 18 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB[0]);
    if (p.nslab_r) tma_prefetch_desc(&p.tmRA);
    if (p.nslab0 < p.nslab) {
      tma_prefetch_desc(&p.tmA[1]);
      tma_prefetch_desc(&p.tmB[1]);
    }
  }
  if (warp == 8) {
    if (lane == 0) {
      for (int i = 0; i < p.stages; ++i) {
        mbar_init(&a_tma[i], 1);
        mbar_init(&a_ready[i], kHxfGroup / 32);
        mbar_init(&a_empty[i], 1);
      }
      for (int i = 0; i < p.nacc; ++i) {
        mbar_init(&tmem_full[i], 1);
        mbar_init(&tmem_empty[i], kHEpiWarps / 2);
      }
      mbar_init(b_full, 1);
      for (int i = 0; i < 4; ++i) mbar_init(&dw_empty[i], 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  if (p.nslab_r) {
    // Identity weights, K-major rows in the UMMA swizzle of span_r (address bits [4, 4+b) ^= bits [7, 7+b)): row n (output channel
    // n0 + n of this CTA) has its single 1.0 at residual channel n, i.e. slab n / kslab_r, column n % kslab_r.
    const uint32_t idb = smem_u32(smem_id), words = ((uint32_t)p.nslab_r * p.r_slot_bytes) >> 4;
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) h_sts128(idb + (i << 4), make_uint4(0u, 0u, 0u, 0u));
    __syncthreads();
    const uint32_t msk = p.span_r == 128 ? 7u : (p.span_r == 64 ? 3u : 1u);
    for (int n = threadIdx.x; n < p.bn; n += blockDim.x) {
      const uint32_t rs = (uint32_t)n / (uint32_t)p.kslab_r, k = (uint32_t)n % (uint32_t)p.kslab_r;
      uint32_t a = idb + rs * p.r_slot_bytes + (uint32_t)n * (uint32_t)p.span_r + k * 2u;
      a ^= ((a >> 7) & msk) << 4;
      asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"((unsigned short)0x3F80) : "memory");  // bf16 1.0
    }
    h_fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Resident weights: static data, requested BEFORE the dependency wait so that the load overlaps the previous kernel's tail.
  if (warp == 18 && lane == 0) {
    mbar_expect_tx(b_full, (uint32_t)(p.ntap_w * p.nslab) * p.b_slot_bytes);
    for (int slab = 0; slab < p.nslab; ++slab)
      for (int tap = 0; tap < p.ntap_w; ++tap)
        tma_load_3d(&p.tmB[slab >= p.nslab0], b_full, smem_b + (size_t)(slab * p.ntap_w + tap) * p.b_slot_bytes,
                    (slab >= p.nslab0 ? slab - p.nslab0 : slab) * p.kslab, (int)blockIdx.y * p.bn, tap);
  }
  pdl_wait();  // everything below reads what earlier kernels of the step wrote

  // Without the GroupNorm prologue warps 0..7 have nothing to transform: they can form two MORE epilogue groups (four in all, one per TMEM
  // accumulator).
  // Same-box A/B (profiles/r02s2_extra_epi_cs4pipe_ab.txt): a gain only where the epilogue is heavy AND has no global operand -- Swish without residual
  // (FWM ffn0: 32 -> 64 @64^2 55.0 -> 51.4 us, 64 -> 128 @32^2 32.2 -> 30.7 us); with an epilogue-side residual four groups LOSE (128 -> 64 @32^2
  // 40.0 -> 44.5 us), so those keep two.
  const bool extra_epi = !gn && !DW && p.nacc == 4 && (F & kEpiAct) != 0 && (F & kEpiRes) == 0;
  if (warp < 8 && !extra_epi) {
    if (gn) {
      // GroupNorm tables are private to the transform warps: the TMA producer, the MMA issuers and the epilogue warps start
      // their loops right after the dependency wait instead of behind this block's global loads + fp64 math.
      for (int i = threadIdx.x; i < p.cin; i += kHxfThreads) {
        s_gamma[i] = p.gn_gamma[i];
        s_beta[i] = p.gn_beta[i];
      }
      for (int i = threadIdx.x; i < n_stat; i += kHxfThreads) s_stat[i] = halo_mean_rstd(p, i);  // fp64 once per CTA
      if (DW)
        for (int i = threadIdx.x; i < 9 * p.cin; i += kHxfThreads) s_dw[i] = p.dw_w[i];
      named_bar_sync(5, kHxfThreads);
    }
    // ===================== transform warps (only with the GroupNorm prologue) =====================
    if constexpr (DW) {
      const uint32_t a_base = smem_u32(smem_a), dwb = smem_u32(smem_dw);
      if (p.kslab == 64) halo_transform_dw_loop<8>(p, a_base, dwb, a_tma, a_ready, dw_empty, s_gamma, s_beta, s_stat, s_dw, threadIdx.x);
      else if (p.kslab == 32) halo_transform_dw_loop<4>(p, a_base, dwb, a_tma, a_ready, dw_empty, s_gamma, s_beta, s_stat, s_dw, threadIdx.x);
      else halo_transform_dw_loop<2>(p, a_base, dwb, a_tma, a_ready, dw_empty, s_gamma, s_beta, s_stat, s_dw, threadIdx.x);
    } else if (gn) {
      const uint32_t a_base = smem_u32(smem_a);
      if (p.kslab == 64) halo_transform_loop<8>(p, a_base, a_tma, a_ready, s_gamma, s_beta, s_stat, threadIdx.x, dts);
      else if (p.kslab == 32 && p.out_h <= 8) halo_transform_loop<4, true>(p, a_base, a_tma, a_ready, s_gamma, s_beta, s_stat, threadIdx.x, dts);
      else if (p.kslab == 32) halo_transform_loop<4>(p, a_base, a_tma, a_ready, s_gamma, s_beta, s_stat, threadIdx.x, dts);
      else halo_transform_loop<2>(p, a_base, a_tma, a_ready, s_gamma, s_beta, s_stat, threadIdx.x, dts);
    }
  } else if (warp == 8 || warp == 9) {
    // ===================== MMA issuers (converged warps, elected lane issues) =====================
    uint64_t* a_full = gn ? a_ready : a_tma;
    const int w = warp - 8;
    if constexpr (DW) {
      const uint32_t sa = smem_u32(smem_a), sd = smem_u32(smem_dw), sb = smem_u32(smem_b);
      if (p.kslab == 64) halo_mma_dw_loop<4>(p, sa, sd, sb, a_full, a_empty, dw_empty, b_full, tmem_full, tmem_empty, tmem_base, my_tiles, w);
      else if (p.kslab == 32) halo_mma_dw_loop<2>(p, sa, sd, sb, a_full, a_empty, dw_empty, b_full, tmem_full, tmem_empty, tmem_base, my_tiles, w);
      else halo_mma_dw_loop<1>(p, sa, sd, sb, a_full, a_empty, dw_empty, b_full, tmem_full, tmem_empty, tmem_base, my_tiles, w);
    } else
    if (p.kslab == 64) halo_mma_loop<4>(p, smem_u32(smem_a), smem_u32(smem_b), a_full, a_empty, b_full, tmem_full, tmem_empty, tmem_base, my_tiles, w, dts);
    else if (p.kslab == 32) halo_mma_loop<2>(p, smem_u32(smem_a), smem_u32(smem_b), a_full, a_empty, b_full, tmem_full, tmem_empty, tmem_base, my_tiles, w, dts);
    else halo_mma_loop<1>(p, smem_u32(smem_a), smem_u32(smem_b), a_full, a_empty, b_full, tmem_full, tmem_empty, tmem_base, my_tiles, w, dts);
  } else if (warp >= 18) {
    // ===================== TMA producer: the halo ring (the resident weights were requested above) =====================
    if (lane == 0) {
      const uint32_t tx = (uint32_t)(kHPx * p.span);
      const uint32_t nst = (uint32_t)p.stages >> 1;  // per ring
      const bool own_ring = kHProducers == 2;        // this producer feeds ring (warp - 18) only: tiles base + ring, base + ring + 2, ...
      HaloIter it;
      {
        int base, count;
        halo_range(p, base, count);
        if (own_ring) it.init(p, base + (warp - 18), 2, base + count);
        else it.init(p, base, 1, base + count);
      }
      long long* const pts = warp == 18 ? dts : nullptr;
      // (stage, phase) of ring 0 / 1 as scalars: a run-time indexed rs[ring] lives in LOCAL memory (LDL/STL on the producer's
      // dependent chain every tile)
      uint32_t st0 = 0u, ph0 = 0u, st1 = 0u, ph1 = 0u;
      uint32_t ring = own_ring ? (uint32_t)(warp - 18) : 0u;
      int u = 0;
      for (; it.remaining > 0; it.next(), ++u) {
        const uint32_t rs = ring ? st1 : st0, rp = ring ? ph1 : ph0;
        const uint32_t stage = ring * nst + rs;
        h_ts(pts, 0, u, 0);
        mbar_wait(&a_empty[stage], rp ^ 1u);
        h_ts(pts, 0, u, 1);
#ifdef DDIF_VAR_NO_TMA
        mbar_arrive(&a_tma[stage]);
#else
        if (it.slab >= p.nslab) {  // residual slab of this CTA's N range
          mbar_expect_tx(&a_tma[stage], (uint32_t)(kHPx * p.span_r));
          tma_load_4d(&p.tmRA, &a_tma[stage], smem_a + (size_t)stage * p.stage_bytes, (int)blockIdx.y * p.bn + (it.slab - p.nslab) * p.kslab_r,
                      it.tx * 8 - 1, it.ty * 16 - 1, it.b);
        } else {
          mbar_expect_tx(&a_tma[stage], tx);
          const int seg = it.slab >= p.nslab0;
          tma_load_4d(&p.tmA[seg], &a_tma[stage], smem_a + (size_t)stage * p.stage_bytes, (seg ? it.slab - p.nslab0 : it.slab) * p.kslab, it.tx * 8 - 1,
                      it.ty * 16 - 1, it.b);
        }
#endif
#ifndef DDIF_VAR_NO_RESPF
        if (p.has_res_map && it.slab == 0) tma_prefetch_l2_4d(&p.tmR, (int)blockIdx.y * p.bn, it.tx * 8, it.ty * 16, it.b);  // residual tile -> L2, `stages` tiles ahead
#endif
        {
          uint32_t ns = rs + 1u, np = rp;
          if (ns == nst) { ns = 0u; np ^= 1u; }
          if (ring) { st1 = ns; ph1 = np; } else { st0 = ns; ph0 = np; }
        }
        if (!own_ring && it.slab == p.nslab_t - 1) ring ^= 1u;  // next tile -> other half-pipeline
      }
      pdl_trigger();  // all loads of this CTA are in flight: the next kernel's CTAs may take over SMs as ours exit
    }
  } else if ((warp >= 10 && warp < 18) || warp < 8) {
    // ===================== epilogue (warps 10..17 [+ 0..7]): groups of 4 warps =====================
    const int grp = warp >= 10 ? (warp - 10) >> 2 : 2 + (warp >> 2);
    halo_epilogue_loop<F>(p, tmem_base, tmem_full, tmem_empty, warp, lane, my_tiles, dts, s_add, grp, extra_epi ? 4 : 2, warp >= 10 ? warp - 10 : 8 + warp,
                          tile_base, 1);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// =====================================================================================================================
// Column-softmax GEMM (round 2): out = W[b] . softmax_over_H(a) + bias (+ residual) in ONE kernel.
//
// Replaces `q.softmax(dim=-2)` followed by the per-sample `attn_out` 1x1 product of FastAttnCondInjection
// (/root/reference/models/sr3_dwt.py:541-573): the stand-alone softmax kernel wrote the normalised q tensor to HBM and the 1x1 GEMM read it
// back (4 + 2 bytes per element: 0.99 ms of the 5.84 ms denoise step at B = 256 for the two launches of the 12 FWM blocks at >= 16 lines).
// Here a tile of the GEMM is (128 / H) image columns x ALL H lines (H = 16, 32, 64; row = y * tile_w + x), so the softmax over the height is
// local to the tile: the TMA producer brings the raw q tile (and the sample's W_eff slab: per-sample weights cannot stay resident) into a
// ring stage, the transform warps replace it IN PLACE by its column softmax, the MMA warp multiplies (one tap), the epilogue of the halo
// conv adds bias and the residual.  Same warp roles, barriers and two half-pipelines as conv3x3_halo_tc_kernel.
//
// Transform of one stage (128 rows x NCK 16-byte chunks) by one group of 128 threads: thread (j, rq, c8) owns chunk c8 of the rows
// j + 8 (rq * NR + i): eight consecutive lanes read eight consecutive rows of one chunk column (conflict-free under the TMA swizzle), and all
// rows of a thread belong to image column x = j mod tile_w.  (1) column maximum with packed max.bf16x2, finished over the lanes that share
// (x, c8) by xor-shuffles (lane masks tile_w .. 4 RQ); (2) e = 2^((v - max) log2 e) in fp32, summed in fp32, parked as f16x2 (values in
// (0, 1]: 11 bits against the 8 of the bf16 result, where fp32 would need 64 live registers per thread); (3) e / sum -> bf16 -> STS.128.
__device__ __forceinline__ float cs_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t cs_max_bf2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t cs_pack_h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ f32x2 cs_unpack_h2(uint32_t w) {
  float lo, hi;
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}\n" : "=f"(lo), "=f"(hi) : "r"(w));
  return pk2(lo, hi);
}

// Warp roles of the column-softmax GEMM: kCsPipes transform groups of 4 warps (one ring each; tile t of the CTA's walk goes through pipeline
// t % kCsPipes), 2 MMA warps (even / odd tiles), 2 epilogue groups of 4 warps (even / odd tiles), 1 TMA producer warp.  Same-box A/B of 2 pipelines
// (608 threads, 96 registers) against 4 (864 threads, 72 registers, 76-212 bytes of spills), profiles/r02s2_cs_walk_ab.txt: dim 64 @64^2 73.2 vs
// 73.0 us, dim 128 @64^2 124.4 vs 101.1 us, dim 128 @32^2 45.6 vs 48.9 us, whole step 5.422 vs 5.450 ms -> 2 (the transform is not what paces the
// common dim-64 case).
#ifndef DDIF_VAR_CS_PIPES  // tuning builds: 2 (608 threads, 96 registers) or 4 (864 threads, 72 registers)
#define DDIF_VAR_CS_PIPES 2
#endif
static constexpr int kCsPipes = DDIF_VAR_CS_PIPES;
// Tile walk of a CTA: grid-strided (tile c, c + G, ...: at any time neighbouring CTAs read neighbouring column groups of the same image lines, so
// the 256-512 byte line segments of a tile's TMA box combine into full DRAM pages) unless DDIF_VAR_CS_CONTIG (contiguous ranges like the halo conv).
__device__ __forceinline__ void cs_walk(const HaloKParams& p, int& first, int& step, int& count) {
#ifdef DDIF_VAR_CS_CONTIG
  halo_range(p, first, count);
  step = 1;
#else
  first = (int)blockIdx.x;
  step = (int)gridDim.x;
  count = first < p.num_tiles ? (p.num_tiles - first + step - 1) / step : 0;
#endif
}
static constexpr int kCsMmaWarp = 4 * kCsPipes, kCsEpiWarp = kCsMmaWarp + 2, kCsProdWarp = kCsEpiWarp + kHEpiWarps;
static constexpr int kCsThreads = 32 * (kCsProdWarp + 1);

template <int NCK>
__device__ __forceinline__ void cs_transform_loop(const HaloKParams& p, uint32_t a_base, uint64_t* a_tma, uint64_t* a_ready, int tid) {
  constexpr int RQ = 128 / (8 * NCK);  // row blocks of 8 * NR rows: 2 (64-channel slabs) or 4 (32-channel slabs)
  constexpr int NR = 16 / RQ;          // rows per thread: 8 or 4
  constexpr uint32_t SPAN = NCK * 16;
  const int grp = tid / kHxfGroup, lt = tid % kHxfGroup;
  const int j = lt & 7, rq = (lt >> 3) % RQ, c8 = lt / (8 * RQ);
  const uint32_t sw = NCK == 8 ? (uint32_t)j : (((uint32_t)j >> 1) & 3u);  // rows j + 8 m all share the swizzle phase of row j
  const uint32_t off0 = (uint32_t)(j + 8 * rq * NR) * SPAN + (((uint32_t)c8 ^ sw) << 4);
  const int tw = 1 << p.tile_w_log2;
  const uint32_t nst = (uint32_t)p.stages / kCsPipes;
  int remaining;
  {
    int first, step, count;
    cs_walk(p, first, step, count);
    remaining = count > grp ? ((count - grp + kCsPipes - 1) / kCsPipes) * p.nslab : 0;  // this group's tiles (positions grp, grp + P, ...) x K slabs
  }
  a_tma += grp * nst; a_ready += grp * nst;
  a_base += (uint32_t)grp * nst * p.stage_bytes;
  uint32_t stage = 0, phase = 0;
  const f32x2 l2e = pk2(1.4426950408889634f, 1.4426950408889634f);
  for (; remaining > 0; --remaining) {
    const uint32_t sbase = a_base + stage * p.stage_bytes + off0;
    mbar_wait(&a_tma[stage], phase);
    uint4 v[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) v[i] = h_lds128(sbase + (uint32_t)i * 8u * SPAN);
    uint32_t mx[4] = {v[0].x, v[0].y, v[0].z, v[0].w};
#pragma unroll
    for (int i = 1; i < NR; ++i) {
      mx[0] = cs_max_bf2(mx[0], v[i].x); mx[1] = cs_max_bf2(mx[1], v[i].y);
      mx[2] = cs_max_bf2(mx[2], v[i].z); mx[3] = cs_max_bf2(mx[3], v[i].w);
    }
    for (int m = tw; m <= 4 * RQ; m <<= 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) mx[q] = cs_max_bf2(mx[q], __shfl_xor_sync(0xffffffffu, mx[q], m));
    }
    f32x2 nm[4], sum[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      nm[q] = mul2(bf2_to_f2(mx[q]), pk2(-1.4426950408889634f, -1.4426950408889634f));
      sum[q] = pk2(0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float lo, hi;
        upk2(fma2(bf2_to_f2(w[q]), l2e, nm[q]), lo, hi);
        const float e0 = cs_ex2(lo), e1 = cs_ex2(hi);
        sum[q] = add2(sum[q], pk2(e0, e1));
        w[q] = cs_pack_h2(e0, e1);
      }
      v[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    for (int m = tw; m <= 4 * RQ; m <<= 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float lo, hi;
        upk2(sum[q], lo, hi);
        lo += __shfl_xor_sync(0xffffffffu, lo, m);
        hi += __shfl_xor_sync(0xffffffffu, hi, m);
        sum[q] = pk2(lo, hi);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float lo, hi;
      upk2(sum[q], lo, hi);
      sum[q] = pk2(__frcp_rn(lo), __frcp_rn(hi));
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      uint32_t o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q] = f2_to_bf2(mul2(cs_unpack_h2(w[q]), sum[q]));
      h_sts128(sbase + (uint32_t)i * 8u * SPAN, make_uint4(o[0], o[1], o[2], o[3]));
    }
    h_fence_proxy_async();
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&a_ready[stage]);
    if (++stage == nst) { stage = 0; phase ^= 1u; }
  }
}

// MMA warp w takes tiles t = w, w + 2, ... into TMEM accumulator w + 2 * (itn & 1); tile t runs in pipeline t % kCsPipes (4 pipelines: = that
// accumulator index, the warp alternates between the rings w and w + 2; 2 pipelines: ring w).
template <int KSTEPS>
__device__ __forceinline__ void cs_mma_loop(const HaloKParams& p, uint32_t a_base, uint64_t* a_full, uint64_t* a_empty, uint64_t* tmem_full,
                                            uint64_t* tmem_empty, uint32_t tmem_base, int my_tiles, int w) {
  static_assert(kCsPipes == 4 || kCsPipes == 2, "pipelines");
  const uint32_t span = (uint32_t)p.span, nst = (uint32_t)p.stages / kCsPipes, stage16 = p.stage_bytes >> 4;
  const uint64_t desc_a0 = make_smem_desc(a_base, 8u * span, p.layout_type);              // dense rows: SBO = 8 rows
  const uint64_t desc_b0 = make_smem_desc(a_base + p.a_bytes, 8u * span, p.layout_type);  // the stage's weight slab
  uint32_t st0 = 0, ph0 = 0, st1 = 0, ph1 = 0, itn = 0;  // ring state of pipelines w and (4 pipelines) w + 2
  for (int t = w; t < my_tiles; t += 2, ++itn) {
    const uint32_t acc = (uint32_t)w + 2u * (itn & 1u);
    const uint32_t hi = kCsPipes == 4 ? (itn & 1u) : 0u, pl = (uint32_t)w + 2u * hi;
    const uint32_t tmem_d = tmem_base + acc * (uint32_t)p.bn;
    mbar_wait(&tmem_empty[acc], ((itn >> 1) & 1u) ^ 1u);
    tc_fence_after();
    uint32_t stage = hi ? st1 : st0, phase = hi ? ph1 : ph0;
    const uint32_t ring0 = pl * nst;
    for (int slab = 0; slab < p.nslab; ++slab) {
      mbar_wait(&a_full[ring0 + stage], phase);
      tc_fence_after();
      const uint64_t so = (uint64_t)((ring0 + stage) * stage16);
      umma_bf16_ss_steps<KSTEPS>(tmem_d, desc_a0 + so, desc_b0 + so, p.idesc, slab != 0 ? 1u : 0u);
      umma_commit_elect(&a_empty[ring0 + stage]);
      if (++stage == nst) { stage = 0; phase ^= 1u; }
    }
    if (hi) { st1 = stage; ph1 = phase; } else { st0 = stage; ph0 = phase; }
    umma_commit_elect(&tmem_full[acc]);
  }
}

template <int F>
__global__ void __launch_bounds__(kCsThreads, 1) cs_gemm_tc_kernel(const __grid_constant__ HaloKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint8_t* smem_a = smem;                                                       // [stages][A slab | weight slab]
  float* s_add = reinterpret_cast<float*>(smem + (size_t)p.stages * p.stage_bytes);  // [8 epilogue warps][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_add + kHEpiWarps * 256);
  uint64_t* a_tma = bars;
  uint64_t* a_ready = bars + kHMaxStages;
  uint64_t* a_empty = bars + 2 * kHMaxStages;
  uint64_t* tmem_full = bars + 3 * kHMaxStages;
  uint64_t* tmem_empty = tmem_full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 4);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int tile_first, tile_step, my_tiles;
  cs_walk(p, tile_first, tile_step, my_tiles);
  if (warp == kCsProdWarp && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB[0]);
  }
  if (warp == kCsMmaWarp) {
    if (lane == 0) {
      for (int i = 0; i < p.stages; ++i) {
        mbar_init(&a_tma[i], 1);
        mbar_init(&a_ready[i], kHxfGroup / 32);
        mbar_init(&a_empty[i], 1);
      }
      for (int i = 0; i < p.nacc; ++i) {
        mbar_init(&tmem_full[i], 1);
        mbar_init(&tmem_empty[i], kHEpiWarps / 2);
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
  pdl_wait();  // everything below reads what earlier kernels of the step wrote

  if (warp < kCsMmaWarp) {
    if (p.kslab == 64) cs_transform_loop<8>(p, smem_u32(smem_a), a_tma, a_ready, threadIdx.x);
    else cs_transform_loop<4>(p, smem_u32(smem_a), a_tma, a_ready, threadIdx.x);
  } else if (warp < kCsEpiWarp) {
    if (p.kslab == 64) cs_mma_loop<4>(p, smem_u32(smem_a), a_ready, a_empty, tmem_full, tmem_empty, tmem_base, my_tiles, warp - kCsMmaWarp);
    else cs_mma_loop<2>(p, smem_u32(smem_a), a_ready, a_empty, tmem_full, tmem_empty, tmem_base, my_tiles, warp - kCsMmaWarp);
  } else if (warp == kCsProdWarp) {
    if (lane == 0) {
      const uint32_t tx = p.a_bytes + p.b_slot_bytes;
      const uint32_t nst = (uint32_t)p.stages / kCsPipes;
      const int twl = p.tile_w_log2;
      HaloIter it;
      it.init(p, tile_first, tile_step, tile_first + my_tiles * tile_step);
      // (stage, phase) of the kCsPipes rings packed into one word (8 bits per ring: stage | phase << 7): a run-time indexed array would live
      // in local memory on this thread's dependent chain
      uint32_t rstate = 0u, ring = 0u;
      for (; it.remaining > 0; it.next()) {
        const uint32_t sh = ring * 8u;
        const uint32_t rs = (rstate >> sh) & 0x7fu, rp = (rstate >> (sh + 7u)) & 1u;
        const uint32_t stage = ring * nst + rs;
        uint8_t* dst = smem_a + (size_t)stage * p.stage_bytes;
        mbar_wait(&a_empty[stage], rp ^ 1u);
        mbar_expect_tx(&a_tma[stage], tx);
        tma_load_4d(&p.tmA[0], &a_tma[stage], dst, it.slab * p.kslab, it.tx << twl, 0, it.b);
        tma_load_3d(&p.tmB[0], &a_tma[stage], dst + p.a_bytes, it.slab * p.kslab, 0, p.w_per_sample ? it.b : 0);
        if (p.has_res_map && it.slab == 0) tma_prefetch_l2_4d(&p.tmR, 0, it.tx << twl, 0, it.b);  // residual tile -> L2, a ring ahead of its use
        {
          uint32_t ns = rs + 1u, np = rp;
          if (ns == nst) { ns = 0u; np ^= 1u; }
          rstate = (rstate & ~(0xffu << sh)) | ((ns | (np << 7)) << sh);
        }
        if (it.slab == p.nslab - 1) ring = (ring + 1u) & (uint32_t)(kCsPipes - 1);
      }
      pdl_trigger();
    }
  } else if (warp >= kCsEpiWarp && warp < kCsProdWarp) {
    halo_epilogue_loop<F>(p, tmem_base, tmem_full, tmem_empty, warp, lane, my_tiles, nullptr, s_add, (warp - kCsEpiWarp) >> 2, 2, warp - kCsEpiWarp,
                          tile_first, tile_step);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == kCsMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

int conv3_halo_set_debug_ts(long long* ptr) {
  DDIF_CUDA_CHECK(cudaMemcpyToSymbol(g_halo_ts, &ptr, sizeof(ptr)));
  return DDIF_OK;
}

// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled ddif_get_encode();
int ddif_sm_count();

typedef void (*HaloKernel)(const HaloKParams);
// Residual-as-operand (HaloKParams::nslab_r): needs a TMA-addressable residual (16-byte aligned base and pixel pitch).
// DDIF_NO_RESMMA=1 keeps the epilogue-side residual add (A/B switch for profiling).
static bool halo_res_mma(const ddif_gemm_t& g) {
#ifdef DDIF_VAR_NO_RESMMA  // tuning build (tools/): keep the epilogue-side residual add everywhere
  const bool off = true;
#else
  const bool off = false;
#endif
  // Measured on B200 (layer_bench, B = 256): 32 -> 32 @64^2 66.9 -> 65.8 us, 64 -> 32 @64^2 71.4 -> 67.3 us, but 64 -> 64 @32^2 35.4 -> 38.7 us and
  // 128 -> 64 @32^2 40.9 -> 58.8 us: with N >= 64 the resident weights leave only 2-3 halo stages per ring and the extra residual
  // stage per tile starves the conv slabs.  Enabled for N <= 32 only.
  return !off && g.residual && !g.dw_w && g.n_pad <= 32 && g.res_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(g.residual) & 15u) == 0;
}
static int halo_flags(const ddif_gemm_t& g) {
  return ((g.residual && !halo_res_mma(g)) ? kEpiRes : 0) | (g.act ? kEpiAct : 0) | (g.stats ? kEpiStats : 0) | (g.out_nchw ? kEpiNchw : 0);
}
static HaloKernel halo_kernel(int f, bool dw = false) {
  if (dw) return f == 0 ? conv3x3_halo_tc_kernel<0, true> : nullptr;  // the q path: bias only, bf16 NHWC output
  switch (f) {
    case 0: return conv3x3_halo_tc_kernel<0, false>;
    case 1: return conv3x3_halo_tc_kernel<1, false>;
    case 2: return conv3x3_halo_tc_kernel<2, false>;
    case 3: return conv3x3_halo_tc_kernel<3, false>;
    case 4: return conv3x3_halo_tc_kernel<4, false>;
    case 5: return conv3x3_halo_tc_kernel<5, false>;
    case 6: return conv3x3_halo_tc_kernel<6, false>;
    case 7: return conv3x3_halo_tc_kernel<7, false>;
    case 8: return conv3x3_halo_tc_kernel<8, false>;
    default: return nullptr;
  }
}
static cudaError_t halo_set_attrs() {
  static bool done = false;
  if (done) return cudaSuccess;
  for (int f = 0; f <= 9; ++f) {
    cudaError_t e = cudaFuncSetAttribute(f == 9 ? halo_kernel(0, true) : halo_kernel(f), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
  }
  done = true;
  return cudaSuccess;
}

struct HaloGeom {
  int cin, kslab, nslab, nslab0, span, stages, stage_bytes, b_slot, n_stat, misc, smem, split, bn, ntap_w, dw_bytes, kslab_r, r_total;
};

static bool halo_geometry(const ddif_gemm_t& g, HaloGeom& h) {
  if (g.nseg < 1 || g.nseg > 2 || g.stride != 1 || g.a_up != 0) return false;
  if (g.out_w < 8 || g.out_h < 8) return false;
  if (g.n_pad % 16 != 0 || g.n_pad > 256 || g.n_pad < 16) return false;
  int cin = 0, gcd = 64;
  for (int s = 0; s < g.nseg; ++s) {
    if (g.taps[s] != 9 || g.w_per_sample[s]) return false;
    if (g.a_c[s] % 16 != 0 || g.a_c[s] <= 0) return false;
    if (g.a_ld[s] % 8 != 0 || g.w_k[s] % 8 != 0 || g.w_k[s] < g.a_c[s]) return false;
    if (g.a_h[s] != g.out_h || g.a_w[s] != g.out_w) return false;
    while (g.a_c[s] % gcd != 0) gcd >>= 1;
    cin += (int)g.a_c[s];
  }
  if (cin > 512) return false;
  const bool dw = g.dw_w != nullptr;
  const bool res_mma = halo_res_mma(g);
  if (dw && (!g.gn_stats || g.dw_n < 16 || g.dw_n % 16 != 0 || g.dw_n > g.n_pad || (g.n_pad - g.dw_n) % 16 != 0 || g.n_valid != g.n_pad)) return false;
  if (g.nseg == 2 && g.gn_stats && !g.gn_stats2) return false;
  if (g.mod || halo_kernel(halo_flags(g), g.dw_w != nullptr) == nullptr) return false;  // CSM modulation / fp32 store + residual: generic kernel
  if (g.out && g.out_nchw) return false;
  if (g.out && g.out_ld % 16 != 0) return false;                       // 32-byte stores
  if (g.residual && g.res_ld % 16 != 0) return false;
  if (g.film && g.film_ld % 4 != 0) return false;
  if (!g.out && !g.out_nchw) return false;
  h.cin = cin;
  h.n_stat = g.gn_stats ? (int)(g.batch < kHMaxStat ? g.batch : kHMaxStat) : 0;
  h.misc = 2 * ((cin + 3) & ~3) * 4 + ((h.n_stat + 1) & ~1) * 8 + (g.gn_stats ? 1 : 2) * kHEpiWarps * 256 * 4 + (dw ? ((9 * cin + 3) & ~3) * 4 : 0) + (3 * kHMaxStages + 16) * 8 + 64 + 1024;
  h.ntap_w = dw ? 1 : 9;
  // Resident weights of one CTA (9 taps x all K slabs x bn rows) must leave room for two rings of >= 2 halo stages:
  // split N over blockIdx.y (1, 2, 4 CTAs per tile) and, before splitting further, halve the K slab (smaller stages).
#ifndef DDIF_VAR_HALO_KSLAB_MAX  // tuning builds (tools/) pass -DDDIF_VAR_HALO_KSLAB_MAX=n
#define DDIF_VAR_HALO_KSLAB_MAX 64
#endif
  // Ring depth (same-box sweep, profiles/r02s2_halo_sweep.txt): with the GroupNorm prologue a stage is held by the transform group and then
  // by the MMAs, so only `stages - 4` halo tiles are really in flight from L2 / HBM; 12 stages instead of 8 take the 32 -> 32 @64^2 layers from
  // 61.0 to 52.6 us (residual) / 56.9 to 45.3 us (no residual) and 16 are no better.  Without the prologue (TMA -> MMA directly) deeper rings
  // LOSE (32 -> 64 + Swish epilogue: 54.3 -> 57.3 us at 12, 59.1 at 16), so those layers keep 8.  (Layers with >= 64 input channels never have
  // room for more than 6-8 stages next to their resident weights.)
  int max_stages = (g.gn_stats && !dw) ? kHMaxStages : (kHMaxStages < 8 ? kHMaxStages : 8), kslab_max = DDIF_VAR_HALO_KSLAB_MAX;
#ifdef DDIF_VAR_HALO_TUNE_ENV  // tuning build only (tools/): sweep the ring depth / K slab from the environment
  if (const char* e = getenv("DDIF_HALO_STAGES")) max_stages = atoi(e) < kHMaxStages ? atoi(e) : kHMaxStages;
  if (const char* e = getenv("DDIF_HALO_KSLAB")) kslab_max = atoi(e);
#endif
  for (int pass = 1; pass < 3; ++pass) {  // pass 1: >= 4 stages; pass 2: accept 2
    for (int split = 1; split <= 4; split *= 2) {
      if (g.n_pad % (16 * split) != 0) break;
      if (split > 1 && (g.out_nchw || dw)) break;
      const int bn = (int)g.n_pad / split;
      for (int kslab = gcd; kslab >= 16; kslab >>= 1) {
        if (kslab > kslab_max && kslab > 16) continue;
        const int span = kslab * 2;
        const int stage_bytes = (kHPx * span + 1023) & ~1023;
        const int b_total = h.ntap_w * cin * bn * 2;  // independent of the slab size
        const int dw_bytes = dw ? 128 * span : 0;
        int kslab_r = 0;
        if (res_mma) {
          kslab_r = 64;
          while (bn % kslab_r != 0 || kslab_r > kslab) kslab_r >>= 1;  // bn % 16 == 0, kslab >= 16: terminates at >= 16
        }
        const int r_total = res_mma ? bn * bn * 2 : 0;  // identity weights: (bn / kslab_r) slots of bn rows x kslab_r columns
        int st = ((227 * 1024 - h.misc - b_total - r_total - 4 * dw_bytes) / stage_bytes) & ~1;
        if (st > max_stages) st = max_stages;
        if (st >= (pass <= 1 ? 4 : 2)) {
          h.kslab = kslab; h.nslab = cin / kslab; h.nslab0 = (int)g.a_c[0] / kslab; h.span = span;
          h.stage_bytes = stage_bytes; h.split = split; h.bn = bn; h.b_slot = bn * span; h.stages = st;
          h.dw_bytes = dw_bytes;
          h.kslab_r = kslab_r; h.r_total = r_total;
          h.smem = st * stage_bytes + 4 * dw_bytes + b_total + r_total + h.misc;
          return true;
        }
      }
    }
  }
  return false;
}

bool conv3_halo_applicable(const ddif_gemm_t& g) {
  HaloGeom h;
  return halo_geometry(g, h);
}

int conv3_halo_prepare(const ddif_gemm_t& g, GemmLaunch& L) {
  HaloKParams& p = *reinterpret_cast<HaloKParams*>(L.kparams);
  static_assert(sizeof(HaloKParams) <= sizeof(L.kparams), "kparams buffer too small");
  memset(&p, 0, sizeof(p));
  PFN_encodeTiled enc = ddif_get_encode();
  if (!enc) return DDIF_ERR_DRIVER;
  DDIF_CUDA_CHECK(halo_set_attrs());
  HaloGeom h;
  if (!halo_geometry(g, h)) return DDIF_ERR_SHAPE;
  if (g.n_valid > g.n_pad || g.n_valid < 1) return DDIF_ERR_SHAPE;
  p.cin = h.cin; p.kslab = h.kslab; p.nslab = h.nslab; p.nslab0 = h.nslab0; p.span = h.span;
  p.batch = (int)g.batch; p.out_h = (int)g.out_h; p.out_w = (int)g.out_w;
  p.tiles_x = (int)ceil_div(p.out_w, 8);
  p.tiles_y = (int)ceil_div(p.out_h, 16);
  p.num_tiles = p.tiles_x * p.tiles_y * p.batch;
  p.bn = h.bn;
  p.stages = h.stages;
  p.stage_bytes = (uint32_t)h.stage_bytes;
  p.b_slot_bytes = (uint32_t)h.b_slot;
  p.layout_type = p.span == 128 ? 2u : p.span == 64 ? 4u : 6u;
  p.nacc = 4 * p.bn <= 512 ? 4 : 2;
  p.tile_w_log2 = 3; p.tile_h = 16;
  uint32_t cols = 32;
  while ((int)cols < p.nacc * p.bn) cols <<= 1;
  p.tmem_cols = cols;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  L.smem_bytes = h.smem;
  L.grid_y = h.split;
  const int sms = ddif_sm_count();
  const int gx = sms / h.split > 0 ? sms / h.split : 1;
  L.grid_x = p.num_tiles < gx ? p.num_tiles : gx;
  L.variant = 2;
  L.flags = halo_flags(g);
  const CUtensorMapSwizzle sw = p.span == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : p.span == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  for (int s = 0; s < g.nseg; ++s) {
    {
      cuuint64_t dims[4] = {(cuuint64_t)g.a_c[s], (cuuint64_t)g.a_w[s], (cuuint64_t)g.a_h[s], (cuuint64_t)g.batch};
      cuuint64_t strides[3] = {(cuuint64_t)g.a_ld[s] * 2, (cuuint64_t)g.a_w[s] * g.a_ld[s] * 2, (cuuint64_t)g.a_h[s] * g.a_w[s] * g.a_ld[s] * 2};
      cuuint32_t box[4] = {(cuuint32_t)p.kslab, (cuuint32_t)kHW, (cuuint32_t)kHH, 1};
      cuuint32_t es[4] = {1, 1, 1, 1};
      CUresult r = enc(&p.tmA[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(g.a[s]), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return DDIF_ERR_DRIVER;
    }
    {
      cuuint64_t dims[3] = {(cuuint64_t)g.w_k[s], (cuuint64_t)g.n_pad, (cuuint64_t)g.w_s[s]};
      cuuint64_t strides[2] = {(cuuint64_t)g.w_k[s] * 2, (cuuint64_t)g.n_pad * g.w_k[s] * 2};
      cuuint32_t box[3] = {(cuuint32_t)p.kslab, (cuuint32_t)p.bn, 1};
      cuuint32_t es[3] = {1, 1, 1};
      CUresult r = enc(&p.tmB[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(g.w[s]), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return DDIF_ERR_DRIVER;
    }
  }
  p.dw_w = g.dw_w;
  p.dw_n = (int)g.dw_n;
  p.ntap_w = h.ntap_w;
  p.dw_bytes = (uint32_t)h.dw_bytes;
  if (g.dw_w) {
    p.idesc_q = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.dw_n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    p.idesc_r = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((p.bn - p.dw_n) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  }
  p.nslab_t = p.nslab;
  p.span_r = p.span; p.layout_r = p.layout_type;  // harmless defaults (descriptors are built even when unused)
  if (h.kslab_r) {
    p.kslab_r = h.kslab_r; p.nslab_r = p.bn / h.kslab_r; p.nslab_t = p.nslab + p.nslab_r;
    p.span_r = h.kslab_r * 2; p.ksteps_r = h.kslab_r / 16;
    p.layout_r = p.span_r == 128 ? 2u : p.span_r == 64 ? 4u : 6u;
    p.r_slot_bytes = (uint32_t)(p.bn * p.span_r);
    cuuint64_t dims[4] = {(cuuint64_t)g.n_valid, (cuuint64_t)g.out_w, (cuuint64_t)g.out_h, (cuuint64_t)g.batch};
    cuuint64_t strides[3] = {(cuuint64_t)g.res_ld * 2, (cuuint64_t)g.out_w * g.res_ld * 2, (cuuint64_t)g.out_h * g.out_w * g.res_ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)h.kslab_r, (cuuint32_t)kHW, (cuuint32_t)kHH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    const CUtensorMapSwizzle swr = p.span_r == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : p.span_r == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    if (enc(&p.tmRA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(g.residual), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swr,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return DDIF_ERR_DRIVER;
  } else if (g.residual) {
    const int nb = (int)(g.n_valid < p.bn ? g.n_valid : p.bn) / 8 * 8;
    cuuint64_t dims[4] = {(cuuint64_t)g.n_valid, (cuuint64_t)g.out_w, (cuuint64_t)g.out_h, (cuuint64_t)g.batch};
    cuuint64_t strides[3] = {(cuuint64_t)g.res_ld * 2, (cuuint64_t)g.out_w * g.res_ld * 2, (cuuint64_t)g.out_h * g.out_w * g.res_ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)nb, 8, 16, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (nb >= 8 && nb <= 256 &&
        enc(&p.tmR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(g.residual), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
      p.has_res_map = 1;
  }
  p.gn_stats = g.gn_stats;
  p.gn_stats2 = g.nseg == 2 ? g.gn_stats2 : nullptr;
  p.gn_gamma = g.gn_gamma;
  p.gn_beta = g.gn_beta;
  p.gn_eps = (float)g.gn_eps;
  p.gn_act = (int)g.gn_act;
  p.gn_count = (double)h.cin * p.out_h * p.out_w;
  if (p.gn_stats && (!p.gn_gamma || !p.gn_beta)) return DDIF_ERR_ARG;
  EpiParams& e = p.epi;
  e.bias = g.bias; e.film = g.film; e.film_ld = (int)g.film_ld; e.mod = (const bf16*)g.mod; e.residual = (const bf16*)g.residual;
  e.res_ld = (int)g.res_ld; e.act = (int)g.act; e.out = (bf16*)g.out; e.out_ld = (int)g.out_ld; e.out_nchw = g.out_nchw; e.stats = g.stats;
  e.n_valid = (int)g.n_valid; e.batch = p.batch; e.out_h = p.out_h; e.out_w = p.out_w;
  return DDIF_OK;
}

int conv3_halo_launch(const GemmLaunch& L, cudaStream_t stream) {
  const HaloKParams& p = *reinterpret_cast<const HaloKParams*>(L.kparams);
  HaloKernel k = halo_kernel(L.flags, p.dw_w != nullptr);
  if (!k) return DDIF_ERR_STATE;
  DDIF_CUDA_CHECK(launch_pdl(k, dim3(L.grid_x, L.grid_y), dim3(kHThreads), (size_t)L.smem_bytes, stream, p));
  return DDIF_OK;
}

// ---- column-softmax GEMM: host side ---------------------------------------------------------------------------------------------------
typedef void (*CsKernel)(const HaloKParams);
static CsKernel cs_kernel(int f) { return f == 0 ? cs_gemm_tc_kernel<0> : (f == kEpiRes ? cs_gemm_tc_kernel<kEpiRes> : nullptr); }

bool cs_gemm_applicable(const ddif_gemm_t& g) {
  if (!g.a_softmax_h || g.nseg != 1 || g.taps[0] != 1 || g.stride != 1 || g.a_up != 0) return false;
  if (g.gn_stats || g.mod || g.film || g.dw_w || g.act || g.stats || g.out_nchw || !g.out) return false;
  const int64_t H = g.out_h, W = g.out_w;
  if (g.a_h[0] != H || g.a_w[0] != W || (H != 16 && H != 32 && H != 64) || W % (128 / H) != 0) return false;
  if (g.a_c[0] % 32 != 0 || g.a_c[0] <= 0 || g.a_c[0] > 512 || g.a_ld[0] % 8 != 0 || g.w_k[0] % 8 != 0 || g.w_k[0] < g.a_c[0]) return false;
  if (g.n_pad % 16 != 0 || g.n_pad < 16 || g.n_pad > 128 || g.n_valid > g.n_pad || g.n_valid < 1) return false;
  if (g.out_ld % 16 != 0 || (g.residual && g.res_ld % 16 != 0)) return false;
  if (g.w_per_sample[0] && g.w_s[0] < g.batch) return false;
  return true;
}

int cs_gemm_prepare(const ddif_gemm_t& g, GemmLaunch& L) {
  HaloKParams& p = *reinterpret_cast<HaloKParams*>(L.kparams);
  memset(&p, 0, sizeof(p));
  PFN_encodeTiled enc = ddif_get_encode();
  if (!enc) return DDIF_ERR_DRIVER;
  if (!cs_gemm_applicable(g)) return DDIF_ERR_SHAPE;
  static bool attrs = false;
  if (!attrs) {
    DDIF_CUDA_CHECK(cudaFuncSetAttribute(cs_kernel(0), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DDIF_CUDA_CHECK(cudaFuncSetAttribute(cs_kernel(kEpiRes), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attrs = true;
  }
  const int H = (int)g.out_h, W = (int)g.out_w, tw = 128 / H;
  p.cin = (int)g.a_c[0];
  p.kslab = p.cin % 64 == 0 ? 64 : 32;
  p.nslab = p.nslab0 = p.nslab_t = p.cin / p.kslab;
  p.span = p.kslab * 2;
  p.ntap_w = 1;
  p.batch = (int)g.batch; p.out_h = H; p.out_w = W;
  p.tile_h = H;
  p.tile_w_log2 = tw == 8 ? 3 : (tw == 4 ? 2 : 1);
  p.tiles_x = W / tw; p.tiles_y = 1;
  p.num_tiles = p.tiles_x * p.batch;
  p.bn = (int)g.n_pad;
  p.a_bytes = 128u * (uint32_t)p.span;
  p.b_slot_bytes = (((uint32_t)p.bn * (uint32_t)p.span) + 1023u) & ~1023u;
  p.stage_bytes = p.a_bytes + p.b_slot_bytes;
  p.w_per_sample = g.w_per_sample[0] ? 1 : 0;
  p.layout_type = p.span == 128 ? 2u : 4u;
  p.span_r = p.span; p.layout_r = p.layout_type;
  const int misc = kHEpiWarps * 256 * 4 + (3 * kHMaxStages + 16) * 8 + 64 + 1024;
  int st = (227 * 1024 - misc) / (int)p.stage_bytes;
  if (st > kHMaxStages) st = kHMaxStages;
  st = st / kCsPipes * kCsPipes;  // one ring per pipeline
  if (st < kCsPipes) return DDIF_ERR_SHAPE;
  p.stages = st;
  p.nacc = 4;
  uint32_t cols = 32;
  while ((int)cols < p.nacc * p.bn) cols <<= 1;
  p.tmem_cols = cols;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  L.smem_bytes = st * (int)p.stage_bytes + misc;
  L.grid_y = 1;
  const int sms = ddif_sm_count();
  L.grid_x = p.num_tiles < sms ? p.num_tiles : sms;
  L.variant = 3;
  L.flags = g.residual ? kEpiRes : 0;
  const CUtensorMapSwizzle sw = p.span == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.a_c[0], (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)g.batch};
    cuuint64_t strides[3] = {(cuuint64_t)g.a_ld[0] * 2, (cuuint64_t)W * g.a_ld[0] * 2, (cuuint64_t)H * W * g.a_ld[0] * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.kslab, (cuuint32_t)tw, (cuuint32_t)H, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&p.tmA[0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(g.a[0]), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return DDIF_ERR_DRIVER;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)g.w_k[0], (cuuint64_t)g.n_pad, (cuuint64_t)g.w_s[0]};
    cuuint64_t strides[2] = {(cuuint64_t)g.w_k[0] * 2, (cuuint64_t)g.n_pad * g.w_k[0] * 2};
    cuuint32_t box[3] = {(cuuint32_t)p.kslab, (cuuint32_t)p.bn, 1};
    cuuint32_t es[3] = {1, 1, 1};
    if (enc(&p.tmB[0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(g.w[0]), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return DDIF_ERR_DRIVER;
  }
  if (g.residual) {
    const int nb = (int)(g.n_valid < p.bn ? g.n_valid : p.bn) / 8 * 8;
    cuuint64_t dims[4] = {(cuuint64_t)g.n_valid, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)g.batch};
    cuuint64_t strides[3] = {(cuuint64_t)g.res_ld * 2, (cuuint64_t)W * g.res_ld * 2, (cuuint64_t)H * W * g.res_ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)nb, (cuuint32_t)tw, (cuuint32_t)H, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (nb >= 8 && nb <= 256 && (reinterpret_cast<uintptr_t>(g.residual) & 15u) == 0 &&
        enc(&p.tmR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(g.residual), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
      p.has_res_map = 1;
  }
  EpiParams& e = p.epi;
  e.bias = g.bias; e.film = nullptr; e.film_ld = 0; e.mod = nullptr; e.residual = (const bf16*)g.residual;
  e.res_ld = (int)g.res_ld; e.act = 0; e.out = (bf16*)g.out; e.out_ld = (int)g.out_ld; e.out_nchw = nullptr; e.stats = nullptr;
  e.n_valid = (int)g.n_valid; e.batch = p.batch; e.out_h = p.out_h; e.out_w = p.out_w;
  return DDIF_OK;
}

int cs_gemm_launch(const GemmLaunch& L, cudaStream_t stream) {
  const HaloKParams& p = *reinterpret_cast<const HaloKParams*>(L.kparams);
  CsKernel k = cs_kernel(L.flags);
  if (!k) return DDIF_ERR_STATE;
  DDIF_CUDA_CHECK(launch_pdl(k, dim3(L.grid_x, L.grid_y), dim3(kCsThreads), (size_t)L.smem_bytes, stream, p));
  return DDIF_OK;
}

}  // namespace ddif
