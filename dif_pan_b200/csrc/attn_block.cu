// Fused 64-token self-attention block: GroupNorm -> qkv 1x1 -> 8-head attention -> out 1x1 -> + bias + residual (+ output statistics)
// in ONE kernel, one CTA per sample.
//
// Replaces `SelfAttention.forward` (/root/reference/models/sr3_dwt.py:330-360) at the resolution the UNet attends at (8x8 = 64
// tokens, C = 128, 8 heads of 16): previously gn_apply + qkv GEMM + attn64 + out GEMM = 4 launches and ~93 us per block at B = 256
// for 0.1 GFLOP per sample-block -- every one of them launch/latency-bound (M = B*64 rows = 128 tiles < 148 SMs), 8 blocks per step.
//
// Why warp-level mma.sync (m16n8k16, bf16 -> fp32) and not tcgen05 here: a sample is 64 x 128 and every GEMM of the block is
// M = 64; the whole block is 10 MFLOP-sample sized and lives in registers.  The FlashAttention-2 register chaining lets each stage
// feed the next WITHOUT shared memory or TMEM round trips: the qkv accumulator fragments of a warp's 16 token rows ARE the A
// fragments of Q K^T, the exponentiated score fragments ARE the A fragments of P V, and the per-head outputs ARE the A fragments
// (one k-step per head) of the out projection.  Only K and V^T cross warps, through 35 KB of shared memory and one __syncthreads.
// A tcgen05 version needs four TMEM->register->shared-memory->descriptor hand-offs per sample for the same arithmetic.
//
// K ordering trick: an MMA sums over its k slots, so A and B may enumerate the channels in ANY common order.  Thread t of a quad
// takes the four 16-byte chunks c = 4j + t (j < 4) of every 256-byte operand row for all 8 k-steps (slot (ks, half, e) <-> channel
// 32(ks>>1) + 8t + 4(ks&1) + 2half + e): every operand row is fetched with four 16-byte accesses per thread instead of sixteen 4-byte
// ones, and with the chunk position XOR-ed by 4*(row & 1) the shared-memory reads of a quarter warp (two rows x four t) hit eight
// distinct 16-byte bank groups.  The out projection's A operand comes from registers in the natural order, so its weight is stored
// with that permutation applied on the host (unet.py).
//
// Weights go through shared memory: the 12 KB qkv slice of head h+1 and the whole 32 KB out weight are fetched with cp.async while
// head h computes (ncu on the first version, which read weight fragments straight from global memory: every HMMA stalled on
// long_scoreboard, 25 us per CTA; profiles/r01s5_ncu_prof_attn_block_v1_global_weights_stalls.txt).
#include "common.cuh"
#include "ddif_internal.h"

namespace ddif {

static constexpr int kAbTok = 64, kAbC = 128, kAbHeads = 8, kAbHd = 16;
static constexpr int kAbKsLd = kAbC + 8;     // bf16 elements per K row (+16 B: conflict-free fragment reads)
static constexpr int kAbVtLd = kAbTok + 8;   // bf16 elements per V^T row

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  // not volatile: a pure function of its operands, so the compiler may interleave independent accumulator chains
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf2(uint32_t w) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}
// word index (0..15) inside a thread's 64-byte operand slice of slot (ks, half)
__device__ __forceinline__ constexpr int ab_word(int ks, int half) { return (ks >> 1) * 4 + (ks & 1) * 2 + half; }

static constexpr int kAbW1Rows = 48, kAbRowB = kAbC * 2;  // weight rows of one head's [q|k|v] slice; bytes per weight row
static constexpr int kAbSmemK = 0, kAbSmemVt = kAbSmemK + kAbTok * kAbKsLd * 2, kAbSmemW1 = kAbSmemVt + kAbC * kAbVtLd * 2,
                     kAbSmemW3 = kAbSmemW1 + 2 * kAbW1Rows * kAbRowB, kAbSmemRed = kAbSmemW3 + kAbC * kAbRowB, kAbSmemBytes = kAbSmemRed + 32;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// `rows` weight rows of 256 bytes -> shared memory, 16-byte chunk c of row r stored at chunk position c ^ (4 * (r & 1))
__device__ __forceinline__ void ab_fill(uint32_t dst, const bf16* src, int rows, int tid) {
  for (int q = tid; q < rows * 16; q += 128) {
    const int r = q >> 4, c = q & 15;
    cp_async16(dst + (uint32_t)(r * kAbRowB + ((c ^ ((r & 1) << 2)) << 4)), src + (size_t)r * kAbC + 8 * c);
  }
}
// the 16 operand words of thread t for weight row r (chunks 4j + t, j < 4)
__device__ __forceinline__ void ab_row(uint32_t base, int r, int t, uint32_t (&w)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t a = base + (uint32_t)(r * kAbRowB + (((4 * j + t) ^ ((r & 1) << 2)) << 4));
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[4 * j]), "=r"(w[4 * j + 1]), "=r"(w[4 * j + 2]), "=r"(w[4 * j + 3]) : "r"(a));
  }
}

__global__ void __launch_bounds__(128) attn_block64_kernel(ddif_attn_block_t p) {
  extern __shared__ __align__(16) uint8_t ab_smem[];
  bf16* s_k = reinterpret_cast<bf16*>(ab_smem + kAbSmemK);
  bf16* s_vt = reinterpret_cast<bf16*>(ab_smem + kAbSmemVt);
  float* s_red = reinterpret_cast<float*>(ab_smem + kAbSmemRed);
  const uint32_t s_w1 = smem_u32(ab_smem + kAbSmemW1), s_w3 = smem_u32(ab_smem + kAbSmemW3);
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int r0 = warp * 16 + g, r1 = r0 + 8;
  const bf16* x = reinterpret_cast<const bf16*>(p.x) + (size_t)b * kAbTok * kAbC;
  const bf16* wqkv = reinterpret_cast<const bf16*>(p.wqkv);
  const bf16* wout = reinterpret_cast<const bf16*>(p.wout);
  // static weights: requested before the dependency wait (they overlap the previous kernel's tail)
  ab_fill(s_w3, wout, kAbC, threadIdx.x);
  ab_fill(s_w1, wqkv, kAbW1Rows, threadIdx.x);
  cp_async_commit();
  pdl_wait();

  // ---- GroupNorm(1 group) statistics of the sample (fp64 sums from the producer's epilogue) ----
  const double cnt = (double)kAbTok * kAbC;
  const double mean_d = p.stats_in[2 * b] / cnt;
  double var_d = p.stats_in[2 * b + 1] / cnt - mean_d * mean_d;
  if (var_d < 0) var_d = 0;
  const float mean = (float)mean_d, rstd = rsqrtf((float)var_d + (float)p.eps);

  // ---- A fragments of the normalised rows r0, r1: chunks 4j + t, i.e. channels 32j + 8t .. + 7 ----
  uint32_t af[8][4];
  {
    uint32_t xr[2][16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 v0 = *reinterpret_cast<const uint4*>(x + (size_t)r0 * kAbC + 32 * j + 8 * t);
      const uint4 v1 = *reinterpret_cast<const uint4*>(x + (size_t)r1 * kAbC + 32 * j + 8 * t);
      xr[0][4 * j] = v0.x; xr[0][4 * j + 1] = v0.y; xr[0][4 * j + 2] = v0.z; xr[0][4 * j + 3] = v0.w;
      xr[1][4 * j] = v1.x; xr[1][4 * j + 1] = v1.y; xr[1][4 * j + 2] = v1.z; xr[1][4 * j + 3] = v1.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + 32 * j + 8 * t)), g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + 32 * j + 8 * t + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + 32 * j + 8 * t)), b1 = __ldg(reinterpret_cast<const float4*>(p.beta + 32 * j + 8 * t + 4));
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float a0 = rstd * gm[2 * w], a1 = rstd * gm[2 * w + 1];
        const float d0 = bt[2 * w] - mean * a0, d1 = bt[2 * w + 1] - mean * a1;
        const float2 u0 = unpack_bf2(xr[0][4 * j + w]), u1 = unpack_bf2(xr[1][4 * j + w]);
        xr[0][4 * j + w] = pack_bf2(fmaf(u0.x, a0, d0), fmaf(u0.y, a1, d1));
        xr[1][4 * j + w] = pack_bf2(fmaf(u1.x, a0, d0), fmaf(u1.y, a1, d1));
      }
    }
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      af[ks][0] = xr[0][ab_word(ks, 0)];
      af[ks][1] = xr[1][ab_word(ks, 0)];
      af[ks][2] = xr[0][ab_word(ks, 1)];
      af[ks][3] = xr[1][ab_word(ks, 1)];
    }
  }

  // ---- qkv = n Wqkv^T, head by head: q stays in registers (A fragments of Q K^T), k -> s_k[token][ch], v -> s_vt[ch][token] ----
  uint32_t qf[kAbHeads][4];
#pragma unroll
  for (int h = 0; h < kAbHeads; ++h) {
    cp_async_wait_all();   // this head's weight slice (and, the first time, the out weight) has landed for this thread's copies ...
    __syncthreads();       // ... and for everybody's; all warps are also done reading the buffer the next fill overwrites
    if (h + 1 < kAbHeads) {
      ab_fill(s_w1 + (uint32_t)(((h + 1) & 1) * kAbW1Rows * kAbRowB), wqkv + (size_t)(h + 1) * kAbW1Rows * kAbC, kAbW1Rows, threadIdx.x);
      cp_async_commit();
    }
    const uint32_t wb = s_w1 + (uint32_t)((h & 1) * kAbW1Rows * kAbRowB);
    float acc[6][4];
#pragma unroll
    for (int nt = 0; nt < 6; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
#pragma unroll
    for (int half6 = 0; half6 < 2; ++half6) {  // three n-tiles at a time: 48 operand registers, three independent accumulator chains
      uint32_t bw[3][16];
#pragma unroll
      for (int u = 0; u < 3; ++u) ab_row(wb, 8 * (3 * half6 + u) + g, t, bw[u]);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
#pragma unroll
        for (int u = 0; u < 3; ++u) mma16816(acc[3 * half6 + u], af[ks], bw[u][ab_word(ks, 0)], bw[u][ab_word(ks, 1)]);
    }
    qf[h][0] = pack_bf2(acc[0][0], acc[0][1]);
    qf[h][1] = pack_bf2(acc[0][2], acc[0][3]);
    qf[h][2] = pack_bf2(acc[1][0], acc[1][1]);
    qf[h][3] = pack_bf2(acc[1][2], acc[1][3]);
#pragma unroll
    for (int nt = 2; nt < 4; ++nt) {
      const int c = 16 * h + 8 * (nt - 2) + 2 * t;
      *reinterpret_cast<uint32_t*>(s_k + r0 * kAbKsLd + c) = pack_bf2(acc[nt][0], acc[nt][1]);
      *reinterpret_cast<uint32_t*>(s_k + r1 * kAbKsLd + c) = pack_bf2(acc[nt][2], acc[nt][3]);
    }
#pragma unroll
    for (int nt = 4; nt < 6; ++nt) {
      const int c = 16 * h + 8 * (nt - 4) + 2 * t;
      s_vt[c * kAbVtLd + r0] = __float2bfloat16(acc[nt][0]);
      s_vt[(c + 1) * kAbVtLd + r0] = __float2bfloat16(acc[nt][1]);
      s_vt[c * kAbVtLd + r1] = __float2bfloat16(acc[nt][2]);
      s_vt[(c + 1) * kAbVtLd + r1] = __float2bfloat16(acc[nt][3]);
    }
  }
  __syncthreads();

  // ---- per head: S = Q K^T (one k-step), softmax over the 64 keys in registers, O = P V; O becomes k-step h of the out projection ----
  const float sl2 = (float)(p.scale * 1.4426950408889634);
  uint32_t of[kAbHeads][4];
#pragma unroll
  for (int h = 0; h < kAbHeads; ++h) {
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s[j][i] = 0.f;
      const bf16* kr = s_k + (8 * j + g) * kAbKsLd + 16 * h + 2 * t;
      mma16816(s[j], qf[h], *reinterpret_cast<const uint32_t*>(kr), *reinterpret_cast<const uint32_t*>(kr + 8));
    }
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      m0 = fmaxf(m0, fmaxf(s[j][0], s[j][1]));
      m1 = fmaxf(m1, fmaxf(s[j][2], s[j][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = exp2f((s[j][0] - m0) * sl2);
      s[j][1] = exp2f((s[j][1] - m0) * sl2);
      s[j][2] = exp2f((s[j][2] - m1) * sl2);
      s[j][3] = exp2f((s[j][3] - m1) * sl2);
      l0 += s[j][0] + s[j][1];
      l1 += s[j][2] + s[j][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    float o[2][4];
#pragma unroll
    for (int dt = 0; dt < 2; ++dt)
#pragma unroll
      for (int i = 0; i < 4; ++i) o[dt][i] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_bf2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dt = 0; dt < 2; ++dt) {
        const bf16* vr = s_vt + (16 * h + 8 * dt + g) * kAbVtLd + 16 * kk + 2 * t;
        mma16816(o[dt], pa, *reinterpret_cast<const uint32_t*>(vr), *reinterpret_cast<const uint32_t*>(vr + 8));
      }
    }
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    of[h][0] = pack_bf2(o[0][0] * i0, o[0][1] * i0);
    of[h][1] = pack_bf2(o[0][2] * i1, o[0][3] * i1);
    of[h][2] = pack_bf2(o[1][0] * i0, o[1][1] * i0);
    of[h][3] = pack_bf2(o[1][2] * i1, o[1][3] * i1);
  }

  // ---- y = O Wout^T + bias + x, four n-tiles at a time (K-permuted weight rows from shared memory) ----
  bf16* out = reinterpret_cast<bf16*>(p.out) + (size_t)b * kAbTok * kAbC;
  float s1 = 0.f, s2 = 0.f;
  for (int ng = 0; ng < 4; ++ng) {
    uint32_t bw[4][16];
#pragma unroll
    for (int u = 0; u < 4; ++u) ab_row(s_w3, 8 * (4 * ng + u) + g, t, bw[u]);
    float yy[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 4; ++i) yy[u][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
#pragma unroll
      for (int u = 0; u < 4; ++u) mma16816(yy[u], of[ks], bw[u][ab_word(ks, 0)], bw[u][ab_word(ks, 1)]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int nt = 4 * ng + u;
      const float (&y)[4] = yy[u];
      const int c = 8 * nt + 2 * t;
      const float2 bo = p.bout ? __ldg(reinterpret_cast<const float2*>(p.bout + c)) : make_float2(0.f, 0.f);
      const float2 x0 = unpack_bf2(*reinterpret_cast<const uint32_t*>(x + (size_t)r0 * kAbC + c));
      const float2 x1 = unpack_bf2(*reinterpret_cast<const uint32_t*>(x + (size_t)r1 * kAbC + c));
      const float v00 = y[0] + bo.x + x0.x, v01 = y[1] + bo.y + x0.y;
      const float v10 = y[2] + bo.x + x1.x, v11 = y[3] + bo.y + x1.y;
      s1 += (v00 + v01) + (v10 + v11);
      s2 = fmaf(v00, v00, fmaf(v01, v01, fmaf(v10, v10, fmaf(v11, v11, s2))));
      *reinterpret_cast<uint32_t*>(out + (size_t)r0 * kAbC + c) = pack_bf2(v00, v01);
      *reinterpret_cast<uint32_t*>(out + (size_t)r1 * kAbC + c) = pack_bf2(v10, v11);
    }
  }
  if (p.stats_out) {
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      s_red[warp] = s1;
      s_red[4 + warp] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      atomicAdd(p.stats_out + 2 * (size_t)b, (double)((s_red[0] + s_red[1]) + (s_red[2] + s_red[3])));
      atomicAdd(p.stats_out + 2 * (size_t)b + 1, (double)((s_red[4] + s_red[5]) + (s_red[6] + s_red[7])));
    }
  }
}

bool attn_block_applicable(const ddif_attn_block_t& p) { return p.ntok == kAbTok && p.c == kAbC && p.heads == kAbHeads; }

int launch_attn_block(const ddif_attn_block_t& p, cudaStream_t s) {
  if (!attn_block_applicable(p)) return DDIF_ERR_SHAPE;
  if (!p.x || !p.stats_in || !p.gamma || !p.beta || !p.wqkv || !p.wout || !p.out || p.batch < 1) return DDIF_ERR_ARG;
  static bool attr_done = false;
  if (!attr_done) {
    DDIF_CUDA_CHECK(cudaFuncSetAttribute(attn_block64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAbSmemBytes));
    attr_done = true;
  }
  DDIF_CUDA_CHECK(launch_pdl(attn_block64_kernel, dim3((unsigned)p.batch), dim3(128), (size_t)kAbSmemBytes, s, p));
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

}  // namespace ddif
