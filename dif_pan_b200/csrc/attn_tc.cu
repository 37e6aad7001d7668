// Flash-style multi-head self-attention core on tcgen05 / TMEM / TMA for LONG token counts (whole-scene mode: a 256x256 / 512x512
// scene is ONE sample whose attention level has 1 024 / 4 096 tokens, /root/reference/diffusion_engine.py:371-380,441-447).
//
// Replaces, for head_dim = 16 and ntok % 128 == 0, the CUDA-core `attn_kernel<HD>` behind
//     qkv = Conv1x1(GN(x));  attn = softmax(Q K^T * scale) V   per head            (/root/reference/models/sr3_dwt.py:347-357)
// (the 64-token patches keep the one-CTA-per-sample register-chained block kernel, attn_block.cu).
//
// One CTA = one (sample, head, 128-query tile); 6 warps:
//   warps 0-3  softmax: thread r owns query row r = TMEM lane r: tcgen05.ld of the 64 scores of a key block, online max / sum in the
//              exp2 domain, P -> bf16 -> shared memory in the UMMA SWIZZLE_128B K-major layout, V^T of the block built from the TMA
//              staging tile (V arrives [key][d]; the P.V MMA wants its B operand K-major = [d][key]), running output O in REGISTERS
//              (head_dim = 16 values per row: the P.V product of each key block goes to a fresh TMEM accumulator and is added after the
//              usual exp2(m_old - m_new) rescale, so there is no TMEM read-modify-write)
//   warp 4     MMA issuer: S_j = Q K_j^T (M128 x N64 x K16, one instruction) one key block AHEAD of the softmax warps, then
//              O_j = P_j V_j (M128 x N16 x K64, four instructions) as soon as P_j / V_j^T are in shared memory
//   warp 5     TMA producer: Q tile once, then a 3-stage ring of (K_j, V_j) tiles straight out of the [B, ntok, 3C] qkv tensor
//              (per-head channel blocks [q16 | k16 | v16], row pitch 3C)
// TMEM: two S accumulators (2 x 64 columns) + two O accumulators (2 x 16) = 256 columns -> two CTAs per SM; shared memory ~53 KB.
// The exponentials bound the kernel (MUFU: 128 x 64 ex2 per key block and CTA = 512 cycles at 16 / clk / SM against ~230 cycles of tensor work).
#include "common.cuh"
#include "ddif_internal.h"
#include "epilogue.cuh"

namespace ddif {

static constexpr int kAtQ = 128, kAtK = 64, kAtHd = 16, kAtStages = 3;
static constexpr int kAtThreads = 192;
static constexpr uint32_t kAtQBytes = kAtQ * 32, kAtKBytes = kAtK * 32, kAtVtBytes = 16 * 128, kAtPBytes = kAtQ * 128;
// shared-memory map (offsets from a 1024-byte aligned base): P[2] | VT[2] | Q | K[3] | Vst[3] | barriers
static constexpr uint32_t kAtOffP = 0, kAtOffVt = kAtOffP + 2 * kAtPBytes, kAtOffQ = kAtOffVt + 2 * kAtVtBytes, kAtOffK = kAtOffQ + kAtQBytes,
                          kAtOffV = kAtOffK + kAtStages * kAtKBytes, kAtOffBar = kAtOffV + kAtStages * kAtKBytes;
static constexpr uint32_t kAtSmem = kAtOffBar + 256 + 1024;

struct alignas(64) AttnTcParams {
  CUtensorMap tmK;  // qkv as (3C, ntok, B), box 16 x 64 x 1, SWIZZLE_32B: Q and K tiles (K-major operand rows of 32 bytes)
  CUtensorMap tmV;  // same tensor, box 16 x 64 x 1, no swizzle: V staging tile [key][d]
  bf16* out;
  int ntok, c, heads;
  float scale_log2e;
};

__device__ __forceinline__ float at_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void at_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(kAtThreads, 2) attn_tc_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAtOffBar);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* kv_full = bars + 1;            // [3] TMA landed (K and V of the stage)
  uint64_t* kv_empty = kv_full + kAtStages;  // [3] S MMA done with K (1) + softmax warps done with the V staging tile (4)
  uint64_t* s_full = kv_empty + kAtStages;   // [2] S accumulator written
  uint64_t* s_empty = s_full + 2;          // [2] S accumulator read by the 4 softmax warps
  uint64_t* p_full = s_empty + 2;          // [2] P and V^T of the block are in shared memory (4 warps)
  uint64_t* o_full = p_full + 2;           // [2] P.V accumulator written (also: P / V^T buffers free again)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = (int)blockIdx.x * kAtQ, head = (int)blockIdx.y, b = (int)blockIdx.z;
  const int nkb = p.ntok / kAtK;

  if (warp == 5 && lane == 0) {
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
  }
  if (warp == 4) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int i = 0; i < kAtStages; ++i) {
        mbar_init(&kv_full[i], 1);
        mbar_init(&kv_empty[i], 5);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&s_empty[i], 4);
        mbar_init(&p_full[i], 4);
        mbar_init(&o_full[i], 1);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // qkv is written by the previous kernel of the step

  if (warp == 5) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int cq = head * 3 * kAtHd;
      mbar_expect_tx(q_full, kAtQBytes);
      tma_load_3d(&p.tmK, q_full, smem + kAtOffQ, cq, q0, b);
      tma_load_3d(&p.tmK, q_full, smem + kAtOffQ + kAtKBytes, cq, q0 + kAtK, b);
      uint32_t stage = 0, phase = 0;
      for (int j = 0; j < nkb; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1u);
        mbar_expect_tx(&kv_full[stage], 2 * kAtKBytes);
        tma_load_3d(&p.tmK, &kv_full[stage], smem + kAtOffK + stage * kAtKBytes, cq + kAtHd, j * kAtK, b);
        tma_load_3d(&p.tmV, &kv_full[stage], smem + kAtOffV + stage * kAtKBytes, cq + 2 * kAtHd, j * kAtK, b);
        if (++stage == kAtStages) { stage = 0; phase ^= 1u; }
      }
      pdl_trigger();
    }
  } else if (warp == 4) {
    // ===================== MMA issuer (converged warp, one elected lane issues) =====================
    const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kAtK >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kAtHd >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t desc_q = make_smem_desc(base + kAtOffQ, 8u * 32u, 6u);        // SWIZZLE_32B, 8-row groups of 256 bytes
    const uint64_t desc_k0 = make_smem_desc(base + kAtOffK, 8u * 32u, 6u);
    const uint64_t desc_p0 = make_smem_desc(base + kAtOffP, 8u * 128u, 2u);      // SWIZZLE_128B, 8-row groups of 1024 bytes
    const uint64_t desc_v0 = make_smem_desc(base + kAtOffVt, 8u * 128u, 2u);
    mbar_wait(q_full, 0u);
    tc_fence_after();
    uint32_t stage = 0, phase = 0;
    auto issue_pv = [&](int j) {  // O_j = P_j V_j into O accumulator j & 1
      const uint32_t bsel = (uint32_t)j & 1u;
      mbar_wait(&p_full[bsel], ((uint32_t)j >> 1) & 1u);
      tc_fence_after();
      umma_bf16_ss_steps<4>(tmem_base + 128u + 32u * bsel, desc_p0 + (uint64_t)(bsel * (kAtPBytes >> 4)), desc_v0 + (uint64_t)(bsel * (kAtVtBytes >> 4)),
                            idesc_o, 0u);
      umma_commit_elect(&o_full[bsel]);
    };
    for (int j = 0; j < nkb; ++j) {
      const uint32_t bsel = (uint32_t)j & 1u;
      mbar_wait(&kv_full[stage], phase);
      mbar_wait(&s_empty[bsel], (((uint32_t)j >> 1) & 1u) ^ 1u);
      tc_fence_after();
      umma_bf16_ss_steps<1>(tmem_base + 64u * bsel, desc_q, desc_k0 + (uint64_t)(stage * (kAtKBytes >> 4)), idesc_s, 0u);
      umma_commit_elect(&s_full[bsel]);
      umma_commit_elect(&kv_empty[stage]);
      if (++stage == kAtStages) { stage = 0; phase ^= 1u; }
      if (j > 0) issue_pv(j - 1);
    }
    issue_pv(nkb - 1);
  } else {
    // ===================== softmax warps: thread = query row = TMEM lane =====================
    const int r = warp * 32 + lane;
    const uint32_t tm_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float c = p.scale_log2e;
    float m_run = -INFINITY, l_run = 0.f;
    float o[kAtHd];
#pragma unroll
    for (int d = 0; d < kAtHd; ++d) o[d] = 0.f;
    // V^T transposition work split: this thread moves 8 head-dim values of one key of the block
    const int vkey = threadIdx.x & 63, vhalf = threadIdx.x >> 6;
    uint32_t stage = 0, phase = 0;
    auto add_pv = [&](int j) {  // o += P_j V_j
      const uint32_t bsel = (uint32_t)j & 1u;
      mbar_wait(&o_full[bsel], ((uint32_t)j >> 1) & 1u);
      tc_fence_after();
      uint32_t pv[16];
      tmem_ld16(tm_lane + 128u + 32u * bsel, pv);
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < kAtHd; ++d) o[d] += __uint_as_float(pv[d]);
    };
    for (int j = 0; j < nkb; ++j) {
      const uint32_t bsel = (uint32_t)j & 1u;
      mbar_wait(&s_full[bsel], ((uint32_t)j >> 1) & 1u);
      tc_fence_after();
      uint32_t s[kAtK];
#pragma unroll
      for (int k = 0; k < kAtK / 16; ++k) tmem_ld16(tm_lane + 64u * bsel + 16u * k, s + 16 * k);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[bsel]);  // the S accumulator can take key block j + 2
      float m_blk = __uint_as_float(s[0]);
#pragma unroll
      for (int k = 1; k < kAtK; ++k) m_blk = fmaxf(m_blk, __uint_as_float(s[k]));
      const float m_new = fmaxf(m_run, m_blk);
      const float alpha = at_ex2((m_run - m_new) * c);   // first block: exp2(-inf) = 0
      const float mc = -m_new * c;
      m_run = m_new;
      if (j > 0) add_pv(j - 1);                          // P_{j-1} V_{j-1} joins before this block's rescale
      l_run *= alpha;
#pragma unroll
      for (int d = 0; d < kAtHd; ++d) o[d] *= alpha;
      // P_j -> bf16 -> shared memory, row r of the SWIZZLE_128B K-major tile: 16-byte chunk k holds keys 8k .. 8k+7
      const uint32_t prow = base + kAtOffP + bsel * kAtPBytes + (uint32_t)r * 128u;
      float lsum = 0.f;
#pragma unroll
      for (int k = 0; k < kAtK / 8; ++k) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float e0 = at_ex2(fmaf(__uint_as_float(s[8 * k + 2 * q]), c, mc));
          const float e1 = at_ex2(fmaf(__uint_as_float(s[8 * k + 2 * q + 1]), c, mc));
          lsum += e0 + e1;
          const __nv_bfloat162 t = __floats2bfloat162_rn(e0, e1);
          w[q] = *reinterpret_cast<const uint32_t*>(&t);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + ((uint32_t)(k ^ (r & 7)) << 4)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                     : "memory");
      }
      l_run += lsum;
      // V_j^T: staging tile [key][16 d] (32-byte rows, as TMA delivered it) -> [d][key] rows of 128 bytes in the SWIZZLE_128B layout
      mbar_wait(&kv_full[stage], phase);
      {
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(base + kAtOffV + stage * kAtKBytes + (uint32_t)vkey * 32u + (uint32_t)vhalf * 16u) : "memory");
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        const uint32_t vt = base + kAtOffVt + bsel * kAtVtBytes + (uint32_t)(vkey & 7) * 2u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t d = (uint32_t)(vhalf * 8 + i);
          const unsigned short h = (unsigned short)((i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xffffu));
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(vt + d * 128u + ((((uint32_t)vkey >> 3) ^ (d & 7u)) << 4)), "h"(h) : "memory");
        }
      }
      at_fence_proxy_async();  // generic-proxy writes of P and V^T become visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&p_full[bsel]);
        mbar_arrive(&kv_empty[stage]);  // this warp is done with the V staging tile
      }
      if (++stage == kAtStages) { stage = 0; phase ^= 1u; }
    }
    add_pv(nkb - 1);
    tc_fence_before();
    const float inv = 1.0f / l_run;
    uint32_t w[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      const __nv_bfloat162 t = __floats2bfloat162_rn(o[2 * d] * inv, o[2 * d + 1] * inv);
      w[d] = *reinterpret_cast<const uint32_t*>(&t);
    }
    stg256(p.out + ((size_t)b * p.ntok + q0 + r) * (size_t)p.c + head * kAtHd, w);
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled ddif_get_encode();

bool attn_tc_applicable(const ddif_attn_t& p) {
  return p.heads > 0 && p.c == p.heads * kAtHd && p.ntok >= kAtQ && p.ntok % kAtQ == 0 && p.c % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(p.qkv) & 15u) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 31u) == 0 && p.batch > 0 && p.batch < 65536;
}

int launch_attn_tc(const ddif_attn_t& p, cudaStream_t s) {
  if (!attn_tc_applicable(p)) return DDIF_ERR_SHAPE;
  PFN_encodeTiled enc = ddif_get_encode();
  if (!enc) return DDIF_ERR_DRIVER;
  static bool attr_set = false;
  if (!attr_set) {
    DDIF_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAtSmem));
    attr_set = true;
  }
  AttnTcParams k;
  memset(&k, 0, sizeof(k));
  const cuuint64_t c3 = (cuuint64_t)(3 * p.c);
  cuuint64_t dims[3] = {c3, (cuuint64_t)p.ntok, (cuuint64_t)p.batch};
  cuuint64_t strides[2] = {c3 * 2, (cuuint64_t)p.ntok * c3 * 2};
  cuuint32_t box[3] = {(cuuint32_t)kAtHd, (cuuint32_t)kAtK, 1};
  cuuint32_t es[3] = {1, 1, 1};
  if (enc(&k.tmK, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(p.qkv), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return DDIF_ERR_DRIVER;
  if (enc(&k.tmV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(p.qkv), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return DDIF_ERR_DRIVER;
  k.out = (bf16*)p.out;
  k.ntok = (int)p.ntok;
  k.c = (int)p.c;
  k.heads = (int)p.heads;
  k.scale_log2e = (float)(p.scale * 1.4426950408889634);
  const dim3 grid((unsigned)(p.ntok / kAtQ), (unsigned)p.heads, (unsigned)p.batch);
  DDIF_CUDA_CHECK(launch_pdl(attn_tc_kernel, grid, dim3(kAtThreads), (size_t)kAtSmem, s, k));
  return DDIF_OK;
}

}  // namespace ddif
