// Training backward, first slice (BASELINE configs[4]; /root/reference/diffusion/diffusion_ddpm_pan.py:692-766 calls loss.backward() and the
// reference gets every gradient from autograd): the two reductions a convolution's backward needs besides the data gradient -- which IS a
// convolution (flipped / transposed weights) and runs on the forward tcgen05 kernels --
//
//   wgrad    dW[tap][o][i] += sum over output pixels p of dY[p][o] * X[p shifted by the tap][i]          (F.conv2d weight gradient)
//   colsum   out[g][c]     += sum over the pixels of group g of dY[p][c]      (bias gradient: one group; FiLM gradient: one group per sample)
//
// wgrad is a GEMM whose reduction dimension is the PIXEL index (K = B*H*W, up to 10^5 .. 10^6) and whose output is tiny (taps x Cout x Cin),
// so it is split over K: one CTA = one (tap, 64 x 64 block of (o, i), pixel range), fp32 partial results added with atomics.  Both operands
// are pixel-major in memory (NHWC: channels contiguous = "MN-major"), so the fragments are built with ldmatrix.trans from a double-buffered
// cp.async ring (zero fill = padding) and multiplied with warp-level mma.sync.m16n8k16 (bf16 -> fp32).  This is the round's correctness-first
// version; a tcgen05 formulation with MN-major shared-memory descriptors is the next step.
#include "common.cuh"
#include "ddif_internal.h"

namespace ddif {

static constexpr int kWgBM = 64, kWgBN = 64, kWgBK = 64, kWgThreads = 256;

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct WgK {
  const bf16* x;
  const bf16* dy;
  float* dw;
  int x_ld, cin, dy_ld, cout;
  int batch, in_h, in_w, out_h, out_w, taps, stride, pad;
  int groups;        // 1, or batch (per-sample weights)
  int splits;        // pixel ranges per group
  int m_blocks, n_blocks;
  int px_per_group;  // output pixels per group
};

// shared tile: [64 pixel rows][64 channels] bf16 = 128 bytes per row, 16-byte chunk c of row r stored at chunk (c ^ (r & 7))
__device__ __forceinline__ uint32_t wg_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

__global__ void __launch_bounds__(kWgThreads) wgrad_kernel(const WgK p) {
  __shared__ __align__(128) uint8_t sA[2][kWgBK * 128];  // dY tile: rows = pixels, columns = 64 output channels of this block
  __shared__ __align__(128) uint8_t sB[2][kWgBK * 128];  // X tile (shifted by the tap): rows = pixels, columns = 64 input channels
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // blockIdx.y -> (tap, m block, n block); blockIdx.x -> (group, split)
  int t = (int)blockIdx.y;
  const int nb = t % p.n_blocks; t /= p.n_blocks;
  const int mb = t % p.m_blocks; t /= p.m_blocks;
  const int tap = t;
  const int g = (int)blockIdx.x / p.splits, sp = (int)blockIdx.x % p.splits;
  const int ky = p.taps == 9 ? tap / 3 : 0, kx = p.taps == 9 ? tap % 3 : 0;
  const int chunks_total = (p.px_per_group + kWgBK - 1) / kWgBK;
  const int c_begin = (int)((int64_t)chunks_total * sp / p.splits), c_end = (int)((int64_t)chunks_total * (sp + 1) / p.splits);
  const int o0 = mb * kWgBM, i0 = nb * kWgBN;
  const int ohw = p.out_h * p.out_w;

  // loader mapping: thread -> (row = tid / 4 [+ 0], chunks (tid % 4) * 2, +1) of both tiles: 64 rows x 8 chunks = 512 chunks per tile
  const int lrow = tid >> 2, lch = (tid & 3) * 2;
  auto load_stage = [&](int stage, int chunk) {
    const int pix = chunk * kWgBK + lrow;            // output pixel inside the group
    const bool pv = pix < p.px_per_group;
    const int64_t gp = (int64_t)g * p.px_per_group + pix;  // global output pixel (b, y, x)
    const int b = (int)(gp / ohw);
    const int r = (int)(gp - (int64_t)b * ohw);
    const int y = r / p.out_w, x = r - y * p.out_w;
    const int iy = y * p.stride + ky - p.pad, ix = x * p.stride + kx - p.pad;
    const bool xv = pv && iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w;
    const bf16* dyp = p.dy + (pv ? gp : 0) * (int64_t)p.dy_ld + o0;
    const bf16* xp = p.x + (xv ? ((int64_t)b * p.in_h + iy) * p.in_w + ix : 0) * (int64_t)p.x_ld + i0;
    const uint32_t a_base = smem_u32(sA[stage]), b_base = smem_u32(sB[stage]);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int ch = lch + k;
      cp_async16_zfill(a_base + wg_off(lrow, ch), dyp + ch * 8, pv && (o0 + ch * 8 < p.cout));
      cp_async16_zfill(b_base + wg_off(lrow, ch), xp + ch * 8, xv && (i0 + ch * 8 < p.cin));
    }
  };

  // warp tile: 16 (o) x 32 (i): warp -> (wm = warp % 4, wn = warp / 4)
  const int wm = warp & 3, wn = warp >> 2;
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;

  if (c_begin < c_end) {
    load_stage(0, c_begin);
    cp_async_commit();
    for (int c = c_begin; c < c_end; ++c) {
      const int st = (c - c_begin) & 1;
      if (c + 1 < c_end) load_stage(st ^ 1, c + 1);
      cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();
      const uint32_t a_base = smem_u32(sA[st]), b_base = smem_u32(sB[st]);
      const int q = lane >> 3, rr = lane & 7;
#pragma unroll
      for (int k0 = 0; k0 < kWgBK; k0 += 16) {
        // A fragment (m16 x k16) = transposed 8x8 blocks of the pixel-major tile: matrices (k lo, m lo), (k lo, m hi), (k hi, m lo), (k hi, m hi)
        uint32_t a[4];
        {
          const int prow = k0 + rr + ((q >> 1) << 3);
          const int ch = wm * 2 + (q & 1);
          ldsm_x4_trans(a_base + wg_off(prow, ch), a[0], a[1], a[2], a[3]);
        }
        // B fragments for four n8 blocks: two ldmatrix.x4.trans, each (k lo, n), (k hi, n), (k lo, n + 8), (k hi, n + 8)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t b0, b1, b2, b3;
          const int prow = k0 + rr + ((q & 1) << 3);
          const int ch = wn * 4 + h * 2 + (q >> 1);
          ldsm_x4_trans(b_base + wg_off(prow, ch), b0, b1, b2, b3);
          mma_bf16_16816(acc[2 * h], a, b0, b1);
          mma_bf16_16816(acc[2 * h + 1], a, b2, b3);
        }
      }
      __syncthreads();
    }
    cp_async_wait<0>();
  }
  // accumulator fragment: rows (lane / 4, + 8) of the m16 block, columns (lane % 4) * 2 + {0, 1} of each n8 block
  float* dst = p.dw + ((size_t)(p.groups > 1 ? g : 0) * p.taps + tap) * (size_t)p.cout * p.cin;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int o = o0 + wm * 16 + (lane >> 2) + ((e >> 1) << 3);
      const int i = i0 + wn * 32 + j * 8 + (lane & 3) * 2 + (e & 1);
      if (o < p.cout && i < p.cin && acc[j][e] != 0.f) atomicAdd(dst + (size_t)o * p.cin + i, acc[j][e]);
    }
  }
}

// A fragment layout note: mma.m16n8k16 wants a0 = (m lo, k lo), a1 = (m hi, k lo), a2 = (m lo, k hi), a3 = (m hi, k hi); the x4 load above
// orders its matrices by lane octet q = 0..3 as (k lo, m lo), (k lo, m hi), (k hi, m lo), (k hi, m hi) = exactly a0..a3 after the transpose.

int launch_wgrad(const ddif_wgrad_t& g, cudaStream_t s) {
  if (!g.x || !g.dy || !g.dw) return DDIF_ERR_ARG;
  if (g.taps != 1 && g.taps != 9) return DDIF_ERR_ARG;
  if (g.stride != 1 && g.stride != 2) return DDIF_ERR_ARG;
  if (g.cin % 8 != 0 || g.cout % 8 != 0 || g.x_ld % 8 != 0 || g.dy_ld % 8 != 0 || g.cin > g.x_ld || g.cout > g.dy_ld) return DDIF_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(g.x) & 15u) || (reinterpret_cast<uintptr_t>(g.dy) & 15u)) return DDIF_ERR_SHAPE;
  if (g.per_sample && g.taps != 1) return DDIF_ERR_ARG;
  WgK k;
  k.x = (const bf16*)g.x; k.dy = (const bf16*)g.dy; k.dw = g.dw;
  k.x_ld = (int)g.x_ld; k.cin = (int)g.cin; k.dy_ld = (int)g.dy_ld; k.cout = (int)g.cout;
  k.batch = (int)g.batch; k.in_h = (int)g.in_h; k.in_w = (int)g.in_w; k.out_h = (int)g.out_h; k.out_w = (int)g.out_w;
  k.taps = (int)g.taps; k.stride = (int)g.stride; k.pad = g.taps == 9 ? 1 : 0;
  k.groups = g.per_sample ? (int)g.batch : 1;
  const int64_t px = g.batch * g.out_h * g.out_w;
  k.px_per_group = (int)(px / k.groups);
  k.m_blocks = (int)ceil_div(g.cout, kWgBM);
  k.n_blocks = (int)ceil_div(g.cin, kWgBN);
  const int tiles = k.taps * k.m_blocks * k.n_blocks * k.groups;
  const int chunks = (k.px_per_group + kWgBK - 1) / kWgBK;
  int splits = (int)ceil_div(4 * 148, tiles);
  if (splits > chunks) splits = chunks;
  if (splits < 1) splits = 1;
  k.splits = splits;
  const dim3 grid((unsigned)(k.groups * splits), (unsigned)(k.taps * k.m_blocks * k.n_blocks));
  wgrad_kernel<<<grid, kWgThreads, 0, s>>>(k);
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

// out[g][c] += sum over the pixels of group g of dy[p][c];  groups = 1 (bias gradient) or batch (FiLM gradient: per sample)
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ dy, float* __restrict__ out, int ld, int c, int px_per_group, int splits) {
  const int g = (int)blockIdx.x / splits, sp = (int)blockIdx.x % splits;
  const int cw = c >> 1;                    // channel pairs
  const int lanes = 256 / cw > 0 ? 256 / cw : 1;  // pixel lanes per CTA
  const int cp = threadIdx.x % cw, pl = threadIdx.x / cw;
  if (pl >= lanes) return;
  const int p_begin = (int)((int64_t)px_per_group * sp / splits), p_end = (int)((int64_t)px_per_group * (sp + 1) / splits);
  const uint32_t* src = reinterpret_cast<const uint32_t*>(dy + (size_t)g * px_per_group * ld) + cp;
  float a0 = 0.f, a1 = 0.f;
  for (int p = p_begin + pl; p < p_end; p += lanes) {
    const uint32_t w = __ldg(src + (size_t)p * (ld >> 1));
    a0 += __uint_as_float(w << 16);
    a1 += __uint_as_float(w & 0xffff0000u);
  }
  atomicAdd(out + (size_t)g * c + 2 * cp, a0);
  atomicAdd(out + (size_t)g * c + 2 * cp + 1, a1);
}

int launch_colsum(const ddif_colsum_t& p, cudaStream_t s) {
  if (!p.dy || !p.out) return DDIF_ERR_ARG;
  if (p.c % 2 != 0 || p.ld % 2 != 0 || p.c > p.ld || p.c > 512 || p.c < 2) return DDIF_ERR_SHAPE;
  const int groups = p.per_sample ? (int)p.batch : 1;
  const int64_t px = p.batch * p.hw / groups;
  int splits = (int)ceil_div(2 * 148, groups);
  const int64_t max_splits = ceil_div(px, 64);
  if (splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  colsum_kernel<<<groups * splits, 256, 0, s>>>((const bf16*)p.dy, p.out, (int)p.ld, (int)p.c, (int)px, splits);
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

}  // namespace ddif
