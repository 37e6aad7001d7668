// Fused front of FastAttnCondInjection at the 8 x 8 level (one CTA per sample, everything between x and y stays in shared memory / registers):
//   x_hat = GroupNorm_1group(cat(x, skip))                      prenorm_x                      /root/reference/models/sr3_dwt.py:507,541
//   q     = Conv1x1(DW3x3(x_hat)) + b                           self.q                         :509-512,543
//   qs    = softmax over the image HEIGHT of q                  q.softmax(dim=-2)              :547
//   y     = W_eff[b] . qs + attn_res(x_hat) + bias              (q @ context) -> attn_out, + attn_res   :549-573 (context folded into W_eff per
//                                                               sample by the cond cache, unet.py)
// Before: gn_dw_tile_kernel + 1x1 GEMM + softmax_h + two-segment 1x1 GEMM = 4 launches and 78 us per block at B = 256 (26.7 + 20.7 + 10.3 + 20.5)
// for 17 MFLOP per sample -- each of them bound by launch / fill latency (16 384 pixels = 128 GEMM tiles on 148 SMs), 4 blocks per step.
//
// Like the 64-token attention block (attn_block.cu) this uses warp-level mma.sync.m16n8k16 (bf16 -> fp32): a sample is 64 x 256, every GEMM has
// M = 64 and the chain needs two block-wide exchanges (depthwise neighbours, softmax over the lines) that tcgen05 would pay with
// TMEM -> register -> shared-memory -> descriptor round trips.  Eight warps: warp w owns the token rows 16 (w & 3) .. + 15 (image lines 2 (w & 3),
// + 1) and half w >> 2 of the 32 output channels of every weight slice (four warps with all 32: 42 instead of 39 us per launch).  Measured and
// rejected: a ring of four 16-row slices with three in flight and ldmatrix A loads (47 us: twice the barriers, one HMMA chain per warp).
//   P0  x, skip -> GroupNorm affine -> bufA (bf16, x_hat)                      one 16-byte chunk per thread and step
//   P1  depthwise 3x3 of x_hat -> bufB                                         thread = one bf16x2 channel pair for all 64 pixels, weights in
//                                                                              registers, a sliding 3 x 3 window along x (3 new LDS per pixel)
//   P2  q = dw . W1^T + b1 -> bufB (in place: the warps hold their A fragments in registers before they overwrite their own rows)
//   P3  softmax over the 8 lines of every (column, channel pair) of bufB
//   P4  y = qs . W_eff[b]^T + x_hat . W_res^T + bias -> global
// The weights (W1 128 KB, W_eff[b] and W_res 64 KB each at dim 256) stream through a double-buffered 32-row slice in shared memory with cp.async
// while the previous slice multiplies.  101-110 KB of shared memory -> two CTAs per SM, so one CTA's exchanges overlap the other's MMAs.
#include "common.cuh"
#include "ddif_internal.h"

namespace ddif {

static constexpr int kFfTok = 64, kFfO = 128, kFfSlice = 32, kFfThreads = 256;

template <int DIM>
struct FfCfg {
  static constexpr int LD = DIM + 8;                  // bf16 elements per shared-memory row (+16 B: conflict-free 32-bit fragment reads)
  static constexpr int KS = DIM / 16;                 // k-steps of one GEMM pass
  static constexpr int NCH = DIM / 8;                 // 16-byte chunks per row
  static constexpr int kBuf = kFfTok * LD * 2, kW = kFfSlice * LD * 2;
  static constexpr int offA = 0, offB = kBuf, offW = 2 * kBuf, offTab = offW + 2 * kW, bytes = offTab + 2 * DIM * 4;
  static constexpr int nP2 = DIM / kFfSlice, nP4 = 2 * (kFfO / kFfSlice), nSlices = nP2 + nP4;
};

__device__ __forceinline__ void ff_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t ff_pack(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}
__device__ __forceinline__ float2 ff_unpack(uint32_t w) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w)); }
__device__ __forceinline__ uint32_t ff_lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void ff_sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void ff_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ff_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ff_ldsm4(uint32_t a, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a) : "memory");
}
__device__ __forceinline__ void ff_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int DIM>
__global__ void __launch_bounds__(kFfThreads, 2) fwm_front64_kernel(ddif_fwm_front_t p) {
  using C = FfCfg<DIM>;
  extern __shared__ __align__(16) uint8_t ff_smem[];
  const uint32_t sA = smem_u32(ff_smem + C::offA), sB = smem_u32(ff_smem + C::offB), sW = smem_u32(ff_smem + C::offW);
  float* s_a = reinterpret_cast<float*>(ff_smem + C::offTab);
  float* s_d = s_a + DIM;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nh = warp >> 2;                        // which 16 of the 32 output channels of a weight slice
  const int r0 = (warp & 3) * 16 + g, r1 = r0 + 8;
  const bf16* w1 = reinterpret_cast<const bf16*>(p.w1);
  const bf16* weff = reinterpret_cast<const bf16*>(p.weff) + (size_t)b * p.weff_rows * p.weff_ld;
  const bf16* wres = reinterpret_cast<const bf16*>(p.wres);

  // slice i of the weight stream: 32 rows of K = DIM elements -> buffer i & 1
  auto issue = [&](int i) {
    const bf16* src;
    int ld;
    if (i < C::nP2) { src = w1 + (size_t)(i * kFfSlice) * p.w1_ld; ld = (int)p.w1_ld; }
    else {
      const int j = i - C::nP2, ns = j >> 1;
      if ((j & 1) == 0) { src = weff + (size_t)(ns * kFfSlice) * p.weff_ld; ld = (int)p.weff_ld; }
      else { src = wres + (size_t)(ns * kFfSlice) * p.wres_ld; ld = (int)p.wres_ld; }
    }
    const uint32_t dst = sW + (uint32_t)((i & 1) * C::kW);
    for (int q = tid; q < kFfSlice * C::NCH; q += kFfThreads) {
      const int r = q / C::NCH, c = q - r * C::NCH;
      ff_cp16(dst + (uint32_t)(r * C::LD * 2 + c * 16), src + (size_t)r * ld + 8 * c);
    }
    ff_commit();
  };
  issue(0);     // static weights: requested before the dependency wait (they overlap the previous kernel's tail)
  // ... as are the GroupNorm scale / shift vectors and this thread's depthwise weights (one CTA wave covers the whole batch, so the kernel's
  // duration IS one CTA's chain of dependent global round trips: everything static is fetched here, everything else in batches)
  const int c1 = (int)p.c1;
  float gam = 0.f, bet = 0.f;
  if (tid < DIM) { gam = __ldg(p.gamma + tid); bet = __ldg(p.beta + tid); }
  const int wc = tid % (DIM / 2), yh = tid / (DIM / 2);  // P1: channel pair, upper / lower four lines
  float2 w[9];
  if (tid < DIM) {
#pragma unroll
    for (int k = 0; k < 9; ++k) w[k] = make_float2(__ldg(p.dw_w + (size_t)k * DIM + 2 * wc), __ldg(p.dw_w + (size_t)k * DIM + 2 * wc + 1));
  }
  pdl_wait();

  // ---- P0: x, skip rows of this sample (all loads in flight at once), GroupNorm(1 group) affine -> bufA (x_hat, bf16) ----
  constexpr int kChunks = kFfTok * C::NCH / kFfThreads;  // 16-byte chunks per thread: 8 (dim 256) or 6 (dim 192)
  static_assert(kFfTok * C::NCH % kFfThreads == 0, "chunks per thread");
  uint4 xin[kChunks];
  {
    const bf16* x = reinterpret_cast<const bf16*>(p.x) + (size_t)b * kFfTok * c1;
    const bf16* sk = reinterpret_cast<const bf16*>(p.skip) + (size_t)b * kFfTok * (DIM - c1);
#pragma unroll
    for (int k = 0; k < kChunks; ++k) {
      const int idx = tid + k * kFfThreads, px = idx / C::NCH, ch = (idx - px * C::NCH) * 8;
      xin[k] = ch < c1 ? *reinterpret_cast<const uint4*>(x + (size_t)px * c1 + ch)
                       : *reinterpret_cast<const uint4*>(sk + (size_t)px * (DIM - c1) + (ch - c1));
    }
  }
  {
    double s = p.stats1[2 * b], ss = p.stats1[2 * b + 1];
    if (p.stats2) { s += p.stats2[2 * b]; ss += p.stats2[2 * b + 1]; }
    const double cnt = (double)DIM * kFfTok;
    const double mean_d = s / cnt;
    double var_d = ss / cnt - mean_d * mean_d;
    if (var_d < 0) var_d = 0;
    const float mean = (float)mean_d, rstd = rsqrtf((float)var_d + (float)p.eps);
    if (tid < DIM) {
      const float a = rstd * gam;
      s_a[tid] = a;
      s_d[tid] = bet - mean * a;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kChunks; ++k) {
    const int idx = tid + k * kFfThreads, px = idx / C::NCH, ch = (idx - px * C::NCH) * 8;
    const uint32_t wv[4] = {xin[k].x, xin[k].y, xin[k].z, xin[k].w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 u = ff_unpack(wv[j]);
      o[j] = ff_pack(fmaf(u.x, s_a[ch + 2 * j], s_d[ch + 2 * j]), fmaf(u.y, s_a[ch + 2 * j + 1], s_d[ch + 2 * j + 1]));
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sA + (uint32_t)(px * C::LD * 2 + ch * 2)), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
  }
  __syncthreads();
  // ---- P1: depthwise 3x3 (zero padding of x_hat) -> bufB; thread = channel pair wc, all 64 pixels ----
  if (tid < DIM) {
    const uint32_t colA = sA + (uint32_t)(wc * 4), colB = sB + (uint32_t)(wc * 4);
    for (int y = 4 * yh; y < 4 * yh + 4; ++y) {
      const bool up = y > 0, dn = y < 7;
      // column triple (rows y-1, y, y+1) of image column xx; zero outside the image
      auto col = [&](int xx, float2 (&c)[3]) {
        const uint32_t a = colA + (uint32_t)((y * 8 + xx) * C::LD * 2);
        c[0] = up ? ff_unpack(ff_lds32(a - 8 * C::LD * 2)) : make_float2(0.f, 0.f);
        c[1] = ff_unpack(ff_lds32(a));
        c[2] = dn ? ff_unpack(ff_lds32(a + 8 * C::LD * 2)) : make_float2(0.f, 0.f);
      };
      float2 cl[3] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, cm[3], cr[3];
      col(0, cm);
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        if (x < 7) col(x + 1, cr);
        else cr[0] = cr[1] = cr[2] = make_float2(0.f, 0.f);
        float ax = 0.f, ay = 0.f;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          ax = fmaf(w[dy * 3 + 0].x, cl[dy].x, ax); ay = fmaf(w[dy * 3 + 0].y, cl[dy].y, ay);
          ax = fmaf(w[dy * 3 + 1].x, cm[dy].x, ax); ay = fmaf(w[dy * 3 + 1].y, cm[dy].y, ay);
          ax = fmaf(w[dy * 3 + 2].x, cr[dy].x, ax); ay = fmaf(w[dy * 3 + 2].y, cr[dy].y, ay);
        }
        ff_sts32(colB + (uint32_t)((y * 8 + x) * C::LD * 2), ff_pack(ax, ay));
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) { cl[dy] = cm[dy]; cm[dy] = cr[dy]; }
      }
    }
  }

  // ---- P2 / P4: two GEMM passes over a stream of 32-row weight slices ----
  uint32_t af[C::KS][4];
  // Fragments come from shared memory with ldmatrix.x4 (ncu of the LDS.32 version: 44 % issue utilisation, mio_throttle the first stall reason --
  // 3 100 of a warp's 8 600 instructions were 4-byte fragment loads).  A: one instruction per k-step (matrices rows 0-7 / 8-15 x k 0-7, rows 0-7 /
  // 8-15 x k 8-15 = a0..a3 of mma.m16n8k16); B: one per k-step for both n-tiles of the warp (n-tile 0 x k lo / hi = b0, b1; n-tile 1 x k lo / hi).
  // Lane l addresses row (l & 7) of matrix l >> 3.
  const uint32_t a_lane = (uint32_t)((((warp & 3) * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * C::LD + (lane >> 4) * 8) * 2);
  const uint32_t b_lane = (uint32_t)(((16 * nh + (lane >> 4) * 8 + (lane & 7)) * C::LD + ((lane >> 3) & 1) * 8) * 2);
  auto load_a = [&](uint32_t base) {
#pragma unroll
    for (int ks = 0; ks < C::KS; ++ks) ff_ldsm4(base + a_lane + (uint32_t)(32 * ks), af[ks]);
  };
  float acc[2][4];
  bf16* out = reinterpret_cast<bf16*>(p.out) + (size_t)b * kFfTok * p.out_ld;
  for (int i = 0; i < C::nSlices; ++i) {
    ff_wait_all();
    __syncthreads();  // slice i has landed for everybody; everybody is done with the buffer slice i + 1 goes to (and, at i = 0, with P1)
    if (i + 1 < C::nSlices) issue(i + 1);
    if (i == 0) {
      load_a(sB);       // A = depthwise output, rows of this warp
      __syncthreads();  // both warps of a row block hold their fragments before either overwrites those rows with q
    }
    if (i == C::nP2) {
      // ---- P3: softmax over the 8 lines for every (image column, channel pair) of q (all warps passed the barrier above: q is complete) ----
      const int x = tid & 7, wcb = tid >> 3;
      for (int wc = wcb; wc < DIM / 2; wc += kFfThreads / 8) {
        const uint32_t a = sB + (uint32_t)(x * C::LD * 2 + wc * 4);
        float2 v[8];
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int y = 0; y < 8; ++y) {
          v[y] = ff_unpack(ff_lds32(a + (uint32_t)(y * 8 * C::LD * 2)));
          m0 = fmaxf(m0, v[y].x);
          m1 = fmaxf(m1, v[y].y);
        }
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) {
          v[y].x = exp2f((v[y].x - m0) * 1.4426950408889634f);
          v[y].y = exp2f((v[y].y - m1) * 1.4426950408889634f);
          l0 += v[y].x;
          l1 += v[y].y;
        }
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
        for (int y = 0; y < 8; ++y) ff_sts32(a + (uint32_t)(y * 8 * C::LD * 2), ff_pack(v[y].x * i0, v[y].y * i1));
      }
      __syncthreads();
    }
    const uint32_t wb = sW + (uint32_t)((i & 1) * C::kW);
    const bool p2 = i < C::nP2;
    const int j = i - C::nP2;           // P4: n-slice j >> 1, part j & 1 (0: W_eff x qs, 1: W_res x x_hat)
    if (!p2) load_a((j & 1) ? sA : sB);
    if (p2 || (j & 1) == 0) {
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[u][e] = 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < C::KS; ++ks) {
      uint32_t bf[4];
      ff_ldsm4(wb + b_lane + (uint32_t)(32 * ks), bf);
      ff_mma(acc[0], af[ks], bf[0], bf[1]);
      ff_mma(acc[1], af[ks], bf[2], bf[3]);
    }
    if (p2) {  // q = acc + b1 -> bufB (bf16, like the q tensor of the unfused path), channels 32 i + 8 u + 2 t
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = kFfSlice * i + 16 * nh + 8 * u + 2 * t;
        const float b0 = p.b1 ? __ldg(p.b1 + c) : 0.f, b1v = p.b1 ? __ldg(p.b1 + c + 1) : 0.f;
        ff_sts32(sB + (uint32_t)(r0 * C::LD * 2 + c * 2), ff_pack(acc[u][0] + b0, acc[u][1] + b1v));
        ff_sts32(sB + (uint32_t)(r1 * C::LD * 2 + c * 2), ff_pack(acc[u][2] + b0, acc[u][3] + b1v));
      }
    } else if (j & 1) {  // y = acc + bias -> global
      const int ns = j >> 1;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = kFfSlice * ns + 16 * nh + 8 * u + 2 * t;
        if (c < (int)p.o) {
          const float b0 = p.bias ? __ldg(p.bias + c) : 0.f, b1v = p.bias ? __ldg(p.bias + c + 1) : 0.f;
          *reinterpret_cast<uint32_t*>(out + (size_t)r0 * p.out_ld + c) = ff_pack(acc[u][0] + b0, acc[u][1] + b1v);
          *reinterpret_cast<uint32_t*>(out + (size_t)r1 * p.out_ld + c) = ff_pack(acc[u][2] + b0, acc[u][3] + b1v);
        }
      }
    }
  }
}

static bool ff_applicable(const ddif_fwm_front_t& p) {
  const int64_t dim = p.c1 + p.c2;
  return p.h == 8 && p.w == 8 && (dim == 256 || dim == 192) && p.c1 % 8 == 0 && p.c2 % 8 == 0 && p.c1 > 0 && p.c2 >= 0 && p.o == kFfO && p.o % 2 == 0 &&
         p.w1_ld >= dim && p.weff_ld >= dim && p.wres_ld >= dim && p.w1_ld % 8 == 0 && p.weff_ld % 8 == 0 && p.wres_ld % 8 == 0 &&
         p.weff_rows >= kFfO && p.out_ld >= p.o && p.out_ld % 2 == 0 && p.batch >= 1 && p.batch <= 65535;
}

int launch_fwm_front(const ddif_fwm_front_t& p, cudaStream_t s) {
  if (!ff_applicable(p)) return DDIF_ERR_SHAPE;
  if (!p.x || !p.stats1 || !p.gamma || !p.beta || !p.dw_w || !p.w1 || !p.weff || !p.wres || !p.out || (p.c2 > 0 && (!p.skip || !p.stats2))) return DDIF_ERR_ARG;
  static bool attr_done = false;
  if (!attr_done) {
    DDIF_CUDA_CHECK(cudaFuncSetAttribute(fwm_front64_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfCfg<256>::bytes));
    DDIF_CUDA_CHECK(cudaFuncSetAttribute(fwm_front64_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfCfg<192>::bytes));
    attr_done = true;
  }
  if (p.c1 + p.c2 == 256)
    DDIF_CUDA_CHECK(launch_pdl(fwm_front64_kernel<256>, dim3((unsigned)p.batch), dim3(kFfThreads), (size_t)FfCfg<256>::bytes, s, p));
  else
    DDIF_CUDA_CHECK(launch_pdl(fwm_front64_kernel<192>, dim3((unsigned)p.batch), dim3(kFfThreads), (size_t)FfCfg<192>::bytes, s, p));
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

}  // namespace ddif
