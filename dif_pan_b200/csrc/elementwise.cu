// Memory-bound kernels of the SR3-DWT UNet forward: layout conversion, GroupNorm(1 group) apply (+Swish, +depthwise
// 3x3), FWM softmaxes / context, self-attention core, time embedding, bilinear cond resize, nearest upsample,
// per-sample statistics, and a CUDA-core direct convolution used as on-GPU checker for the tcgen05 path.
// All activations are NHWC bf16; every thread moves 16-byte vectors (8 channels) with channel-contiguous,
// warp-coalesced accesses.  Reference: /root/reference/models/sr3_dwt.py (line numbers per kernel).
#include "common.cuh"
#include "ddif_internal.h"

namespace ddif {

static inline int grid_for(int64_t items, int threads, int cap = 148 * 16) {
  int64_t b = ceil_div(items, threads);
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ---- NCHW fp32 (x, self_cond) -> NHWC bf16 [self_cond | x | 0-pad]   (sr3_dwt.py:172-174) -----------------
__global__ void in_convert_kernel(const float* __restrict__ x, const float* __restrict__ sc, bf16* __restrict__ out,
                                  int B, int C, int HW, int c_pad) {
  pdl_wait();
  const int64_t total = (int64_t)B * HW;
  const int nsrc = sc ? 2 : 1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const int pix = (int)(i - (int64_t)b * HW);
    bf16* o = out + i * c_pad;
    for (int c0 = 0; c0 < c_pad; c0 += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        float t = 0.f;
        if (c < nsrc * C) {
          const float* src = (nsrc == 2 && c < C) ? sc : x;
          const int cc = (nsrc == 2 && c >= C) ? c - C : c;
          t = __ldg(src + ((size_t)b * C + cc) * HW + pix);
        }
        v[j] = t;
      }
      *reinterpret_cast<bf16x8*>(o + c0) = pack8(v);
    }
  }
}
int launch_in_convert(const ddif_in_convert_t& p, cudaStream_t s) {
  if (p.c_pad % 8 != 0 || p.c_pad < (p.self_cond ? 2 : 1) * p.c) return DDIF_ERR_SHAPE;
  const int64_t hw = p.h * p.w;
  DDIF_CUDA_CHECK(launch_pdl(in_convert_kernel, dim3(grid_for(p.batch * hw, 256)), dim3(256), (size_t)(0), s, p.x, p.self_cond, (bf16*)p.out, (int)p.batch, (int)p.c, (int)hw, (int)p.c_pad));
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

// ---- time embedding + all FiLM vectors (sr3_dwt.py:57-64, 223-238, 245-257) ------------------------------
__global__ void time_embed_kernel(ddif_time_embed_t p) {
  pdl_wait();
  extern __shared__ float sm[];
  const int inner = (int)p.inner, hid = 4 * inner, count = inner / 2;
  float* enc = sm;
  float* h = sm + inner;
  float* te = h + hid;
  const int b = blockIdx.x;
  const float t = p.time[b];
  for (int i = threadIdx.x; i < inner; i += blockDim.x) {
    const int k = i < count ? i : i - count;
    const float step = (float)k / (float)count;
    const float e = t * expf(-9.210340371976184f * step);
    enc[i] = i < count ? sinf(e) : cosf(e);
  }
  __syncthreads();
  // warp-cooperative mat-vecs: a warp reads one weight row with coalesced loads (lanes split K) and reduces by shuffles;
  // thread-per-row reads touched 32 cache lines per load instruction (97 us for 73 k MACs per sample).
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int j = warp; j < hid; j += nwarps) {
    float a = 0.f;
    for (int k = lane; k < inner; k += 32) a = fmaf(p.w1[j * inner + k], enc[k], a);
    a = warp_sum(a) + p.b1[j];
    if (lane == 0) h[j] = a / (1.0f + expf(-a));
  }
  __syncthreads();
  for (int j = warp; j < inner; j += nwarps) {
    float a = 0.f;
    for (int k = lane; k < hid; k += 32) a = fmaf(p.w2[j * hid + k], h[k], a);
    a = warp_sum(a) + p.b2[j];
    if (lane == 0) te[j] = a;
  }
  __syncthreads();
  if (inner == 32) {
    const float tk = te[lane];
    float* dst = p.film + (size_t)b * p.nfilm;
    constexpr int R = 8;  // independent rows per warp iteration: 8 L2 loads in flight per lane (the loop is latency-bound)
    for (int j0 = warp * R; j0 < (int)p.nfilm; j0 += nwarps * R) {
      float a[R];
#pragma unroll
      for (int r = 0; r < R; ++r) a[r] = (j0 + r < (int)p.nfilm) ? __ldg(p.wf + (size_t)(j0 + r) * 32 + lane) * tk : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] += __shfl_xor_sync(0xffffffffu, a[r], o);
      float mine = a[0];
#pragma unroll
      for (int r = 1; r < R; ++r) mine = lane == r ? a[r] : mine;
      if (lane < R && j0 + lane < (int)p.nfilm) dst[j0 + lane] = mine + __ldg(p.bf + j0 + lane);
    }
  } else {
    for (int j = warp; j < (int)p.nfilm; j += nwarps) {
      float a = 0.f;
      for (int k = lane; k < inner; k += 32) a = fmaf(p.wf[(size_t)j * inner + k], te[k], a);
      a = warp_sum(a) + p.bf[j];
      if (lane == 0) p.film[(size_t)b * p.nfilm + j] = a;
    }
  }
}
int launch_time_embed(const ddif_time_embed_t& p, cudaStream_t s) {
  if (p.inner % 2 != 0 || p.inner > 256) return DDIF_ERR_SHAPE;
  DDIF_CUDA_CHECK(launch_pdl(time_embed_kernel, dim3((int)p.batch), dim3(1024), (size_t)((size_t)(6 * p.inner) * sizeof(float)), s, p));
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

// ---- GroupNorm(1 group) apply (+Swish) (+depthwise 3x3)   (sr3_dwt.py:292-294, 336, 381-382, 507-512) ------
struct GnK {
  const bf16* src1; const bf16* src2; int c1, c2;
  const double* st1; const double* st2;
  const float* gamma; const float* beta;
  bf16* out; const float* dw_w; bf16* out_dw;
  int H, W, act; float eps;
};

__device__ __forceinline__ void gn_norm8(const bf16* src, const float* a, const float* d, int act, float* y) {
  float v[8];
  unpack8(*reinterpret_cast<const bf16x8*>(src), v);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float t = fmaf(v[j], a[j], d[j]);  // a, d are pre-halved when act (see gn_apply_kernel)
    y[j] = act ? swish_half(t) : t;
  }
}

// One thread owns one 8-channel chunk for its whole lifetime (the grid-stride is a multiple of the chunk count), so
// the per-channel affine (rstd*gamma, beta - mean*rstd*gamma) lives in registers and the loop has no division.
static constexpr int kGnThreads = 192;  // divisible by every chunk count the UNet uses (4,8,12,16,24,32,64)

template <bool DW>
__global__ void __launch_bounds__(kGnThreads) gn_apply_kernel(GnK p) {
  pdl_wait();
  const int b = blockIdx.y;
  const int C = p.c1 + p.c2;
  const int nchunk = C >> 3;
  const int HW = p.H * p.W;
  double s = p.st1[2 * b], ss = p.st1[2 * b + 1];
  double n = (double)p.c1 * HW;
  if (p.c2) {
    s += p.st2[2 * b];
    ss += p.st2[2 * b + 1];
    n += (double)p.c2 * HW;
  }
  const double mean_d = s / n;
  double var_d = ss / n - mean_d * mean_d;
  if (var_d < 0) var_d = 0;
  const float mean = (float)mean_d;
  const float rstd = rsqrtf((float)var_d + p.eps);
  const uint32_t stride_items = gridDim.x * kGnThreads;  // multiple of nchunk (checked on the host)
  const uint32_t i0 = blockIdx.x * kGnThreads + threadIdx.x;
  const int ch = (int)(i0 % (uint32_t)nchunk) << 3;
  uint32_t pix = i0 / (uint32_t)nchunk;
  const uint32_t dpix = stride_items / (uint32_t)nchunk;
  float a[8], d[8];
  const float hs = p.act ? 0.5f : 1.0f;  // swish(t) = h*tanh(h) + h with h = t/2
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = hs * rstd * __ldg(p.gamma + ch + j);
    d[j] = hs * __ldg(p.beta + ch + j) - mean * a[j];
  }
  const bool first = ch < p.c1;
  const int ld = first ? p.c1 : p.c2;
  const bf16* src = (first ? p.src1 : p.src2) + (size_t)b * HW * ld + (first ? ch : ch - p.c1);
  bf16* dst = p.out + (size_t)b * HW * C + ch;
  bf16* dst_dw = p.out_dw ? p.out_dw + (size_t)b * HW * C + ch : nullptr;
  float w[DW ? 9 : 1][8];
  if (DW) {
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.dw_w + (size_t)tap * C + ch));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.dw_w + (size_t)tap * C + ch + 4));
      w[tap][0] = w0.x; w[tap][1] = w0.y; w[tap][2] = w0.z; w[tap][3] = w0.w;
      w[tap][4] = w1.x; w[tap][5] = w1.y; w[tap][6] = w1.z; w[tap][7] = w1.w;
    }
  }
  if (!DW) {
    // 4 independent 16-byte loads in flight per thread before any use
    for (; pix + 3 * dpix < (uint32_t)HW; pix += 4 * dpix) {
      bf16x8 in[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) in[u] = *reinterpret_cast<const bf16x8*>(src + (size_t)(pix + u * dpix) * ld);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float v[8], y[8];
        unpack8(in[u], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = fmaf(v[j], a[j], d[j]);
          y[j] = p.act ? swish_half(t) : t;
        }
        *reinterpret_cast<bf16x8*>(dst + (size_t)(pix + u * dpix) * C) = pack8(y);
      }
    }
  }
  for (; pix < (uint32_t)HW; pix += dpix) {
    float y[8];
    gn_norm8(src + (size_t)pix * ld, a, d, p.act, y);
    *reinterpret_cast<bf16x8*>(dst + (size_t)pix * C) = pack8(y);
    if (DW) {
      const int py = (int)(pix / (uint32_t)p.W), px = (int)pix - py * p.W;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
        if (yy < 0 || yy >= p.H || xx < 0 || xx >= p.W) continue;
        float z[8];
        if (tap == 4) {
#pragma unroll
          for (int j = 0; j < 8; ++j) z[j] = y[j];
        } else {
          gn_norm8(src + ((size_t)yy * p.W + xx) * ld, a, d, p.act, z);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += w[tap][j] * z[j];
      }
      *reinterpret_cast<bf16x8*>(dst_dw + (size_t)pix * C) = pack8(acc);
    }
  }
}
// ---- GroupNorm(1 group) + depthwise 3x3, shared-memory tiled   (FWM prenorm_x + q[0], sr3_dwt.py:507-512,541) -----
// One CTA = one (16 x 8)-pixel tile x one 32-channel group: the (18 x 10) halo is read ONCE (16-byte loads), normalised
// in fp32 into shared memory (zero outside the image: the depthwise conv pads the NORMALISED tensor), the centre pixels
// are written out as x_hat, then every thread produces two (pixel, 8-channel) outputs of the depthwise conv from
// shared memory.  HBM traffic = 2 B read + 4 B written per element instead of nine re-normalised neighbour reads
// (profiles/r01_step_B256_v1: 840 GB/s for the per-pixel version).  Row pitch 144 B keeps every LDS.128 / STS.128
// phase on distinct banks.
static constexpr int kDwTW = 16, kDwTH = 8, kDwHW = kDwTW + 2, kDwHH = kDwTH + 2, kDwHalo = kDwHW * kDwHH;
static constexpr int kDwCh = 32, kDwPitch = 144, kDwThreads = 256;

__global__ void __launch_bounds__(kDwThreads) gn_dw_tile_kernel(GnK p) {
  pdl_wait();
  __shared__ __align__(16) uint8_t s_tile[kDwHalo * kDwPitch];
  __shared__ __align__(16) float s_w[9][kDwCh];
  __shared__ __align__(16) float s_a[kDwCh], s_d[kDwCh];
  const int C = p.c1 + p.c2;
  const int ngrp = C / kDwCh;
  const int cg = blockIdx.x % ngrp;
  const int tile = blockIdx.x / ngrp;
  const int tiles_x = (p.W + kDwTW - 1) / kDwTW;
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int b = blockIdx.y;
  const int ch0 = cg * kDwCh;
  const int HW = p.H * p.W;
  const int t = threadIdx.x;
  if (t < kDwCh) {
    double s = p.st1[2 * b], ss = p.st1[2 * b + 1];
    double n = (double)p.c1 * HW;
    if (p.c2) {
      s += p.st2[2 * b];
      ss += p.st2[2 * b + 1];
      n += (double)p.c2 * HW;
    }
    const double mean_d = s / n;
    double var_d = ss / n - mean_d * mean_d;
    if (var_d < 0) var_d = 0;
    const float mean = (float)mean_d;
    const float rstd = rsqrtf((float)var_d + p.eps);
    const float a = rstd * __ldg(p.gamma + ch0 + t);
    s_a[t] = a;
    s_d[t] = __ldg(p.beta + ch0 + t) - mean * a;
  }
  for (int i = t; i < 9 * kDwCh; i += kDwThreads) s_w[i / kDwCh][i % kDwCh] = __ldg(p.dw_w + (size_t)(i / kDwCh) * C + ch0 + (i % kDwCh));
  __syncthreads();
  // phase 1: halo -> normalise -> smem (fp32) and x_hat (bf16, centre pixels)
  const int y0 = ty * kDwTH - 1, x0 = tx * kDwTW - 1;
  const int c = t & 3;            // 8-channel chunk inside the 32-channel group
  const int chc = ch0 + c * 8;    // its first channel in the concatenated tensor
  const bool first = chc < p.c1;
  const int ld = first ? p.c1 : p.c2;
  const bf16* src = (first ? p.src1 + chc : p.src2 + (chc - p.c1)) + (size_t)b * HW * ld;
  float a[8], d[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = s_a[c * 8 + j];
    d[j] = s_d[c * 8 + j];
  }
  bf16x8 in[3];
  bool ok[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int px = (t >> 2) + k * (kDwThreads / 4);
    const int hy = px / kDwHW, hx = px - hy * kDwHW;
    const int gy = y0 + hy, gx = x0 + hx;
    ok[k] = px < kDwHalo && (unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W;
    if (ok[k]) in[k] = *reinterpret_cast<const bf16x8*>(src + (size_t)(gy * p.W + gx) * ld);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int px = (t >> 2) + k * (kDwThreads / 4);
    if (px < kDwHalo) {
      float y[8];
      if (ok[k]) {
        float v[8];
        unpack8(in[k], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = fmaf(v[j], a[j], d[j]);
        const int hy = px / kDwHW, hx = px - hy * kDwHW;
        if (hy >= 1 && hy <= kDwTH && hx >= 1 && hx <= kDwTW)
          *reinterpret_cast<bf16x8*>(p.out + ((size_t)b * HW + (size_t)(y0 + hy) * p.W + (x0 + hx)) * C + chc) = pack8(y);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = 0.f;
      }
      float4* dst = reinterpret_cast<float4*>(s_tile + px * kDwPitch + c * 32);
      dst[0] = make_float4(y[0], y[1], y[2], y[3]);
      dst[1] = make_float4(y[4], y[5], y[6], y[7]);
    }
  }
  __syncthreads();
  // phase 2: depthwise 3x3 from smem; thread -> pixels (t>>2) and (t>>2)+64 of the 128-pixel tile, chunk c
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int pt = (t >> 2) + k * 64;
    const int py = pt / kDwTW, pxx = pt - py * kDwTW;
    const int gy = ty * kDwTH + py, gx = tx * kDwTW + pxx;
    if (gy >= p.H || gx >= p.W) continue;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float4* z = reinterpret_cast<const float4*>(s_tile + ((py + tap / 3) * kDwHW + pxx + tap % 3) * kDwPitch + c * 32);
      const float4* w = reinterpret_cast<const float4*>(&s_w[tap][c * 8]);
      const float4 z0 = z[0], z1 = z[1], w0 = w[0], w1 = w[1];
      acc[0] += w0.x * z0.x; acc[1] += w0.y * z0.y; acc[2] += w0.z * z0.z; acc[3] += w0.w * z0.w;
      acc[4] += w1.x * z1.x; acc[5] += w1.y * z1.y; acc[6] += w1.z * z1.z; acc[7] += w1.w * z1.w;
    }
    *reinterpret_cast<bf16x8*>(p.out_dw + ((size_t)b * HW + (size_t)gy * p.W + gx) * C + chc) = pack8(acc);
  }
}

int launch_gn_apply(const ddif_gn_apply_t& p, cudaStream_t s) {
  if (p.c1 % 8 != 0 || p.c2 % 8 != 0 || p.c1 <= 0) return DDIF_ERR_SHAPE;
  if (p.c2 && (!p.src2 || !p.stats2)) return DDIF_ERR_ARG;
  if (p.dw_w && !p.out_dw) return DDIF_ERR_ARG;
  const int nchunk = (int)((p.c1 + p.c2) / 8);
  if (p.dw_w && (p.c1 + p.c2) % kDwCh == 0 && !p.act && p.batch <= 65535) {
    GnK k{(const bf16*)p.src1, (const bf16*)p.src2, (int)p.c1, (int)p.c2, p.stats1, p.stats2, p.gamma, p.beta,
          (bf16*)p.out, p.dw_w, (bf16*)p.out_dw, (int)p.h, (int)p.w, (int)p.act, (float)p.eps};
    const int64_t tiles = ceil_div(p.w, kDwTW) * ceil_div(p.h, kDwTH);
    DDIF_CUDA_CHECK(launch_pdl(gn_dw_tile_kernel, dim3(dim3((unsigned)(tiles * ((p.c1 + p.c2) / kDwCh)), (unsigned)p.batch)), dim3(kDwThreads), (size_t)(0), s, k));
    DDIF_LAUNCH_CHECK();
    return DDIF_OK;
  }
  if (kGnThreads % nchunk != 0) return DDIF_ERR_SHAPE;  // channel counts are multiples of 8 dividing 1536
  GnK k{(const bf16*)p.src1, (const bf16*)p.src2, (int)p.c1, (int)p.c2, p.stats1, p.stats2, p.gamma, p.beta,
        (bf16*)p.out, p.dw_w, (bf16*)p.out_dw, (int)p.h, (int)p.w, (int)p.act, (float)p.eps};
  const int64_t items = p.h * p.w * nchunk;
  // ~8 items per thread, at most ~32 CTAs per SM over the whole launch
  int64_t gx = ceil_div(items, (int64_t)kGnThreads * 8);
  const int64_t cap = ceil_div((int64_t)148 * 32, p.batch);
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  if (p.dw_w)
    DDIF_CUDA_CHECK(launch_pdl(gn_apply_kernel<true>, dim3(dim3((unsigned)gx, (unsigned)p.batch)), dim3(kGnThreads), (size_t)(0), s, k));
  else
    DDIF_CUDA_CHECK(launch_pdl(gn_apply_kernel<false>, dim3(dim3((unsigned)gx, (unsigned)p.batch)), dim3(kGnThreads), (size_t)(0), s, k));
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

// ---- q.softmax(dim=-2) * scale   (sr3_dwt.py:545, 561): softmax over H for every (b, x, channel) ---------
__global__ void softmax_h_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int B, int H, int W, int C, int ld, float scale) {
  pdl_wait();
  const int nchunk = C >> 3;
  const int64_t items = (int64_t)B * W * nchunk;
  const size_t row = (size_t)W * C, row_in = (size_t)W * ld;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < items; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / ((int64_t)W * nchunk));
    const int r = (int)(i - (int64_t)b * W * nchunk);  // (x, chunk) linear == offset/8 inside a dense row
    const int x = r / nchunk, ch = r - x * nchunk;
    const bf16* src = in + (size_t)b * H * row_in + (size_t)x * ld + (size_t)ch * 8;
    bf16* dst = out + (size_t)b * H * row + (size_t)r * 8;
    float m[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { m[j] = -INFINITY; l[j] = 0.f; }
    for (int y = 0; y < H; ++y) {
      float v[8];
      unpack8(*reinterpret_cast<const bf16x8*>(src + (size_t)y * row_in), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float mn = fmaxf(m[j], v[j]);
        l[j] = l[j] * __expf(m[j] - mn) + __expf(v[j] - mn);
        m[j] = mn;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) l[j] = scale / l[j];
    for (int y = 0; y < H; ++y) {
      float v[8];
      unpack8(*reinterpret_cast<const bf16x8*>(src + (size_t)y * row_in), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __expf(v[j] - m[j]) * l[j];
      *reinterpret_cast<bf16x8*>(dst + (size_t)y * row) = pack8(v);
    }
  }
}
// Register-resident variant for H <= 64: one thread owns one bf16x2 word (2 channels) of one image column, keeps its H
// values packed in registers (H independent 4-byte loads in flight; a warp covers 128 contiguous bytes per row) and
// makes ONE pass over HBM: 4 B/element instead of 6, one exp per element instead of three.
template <int H>
__global__ void __launch_bounds__(128) softmax_h_reg_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int64_t items,
                                                            int W, int cw, int ldw, float scale) {
  pdl_wait();
  constexpr float kLog2e = 1.4426950408889634f;
  const int row_out = W * cw, row_in = W * ldw;  // words per image row (cw = c/2 output words, ldw = in_ld/2 input words per pixel)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < items; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / row_out;
    const int r = (int)(i - b * row_out);
    const int x = r / cw, c = r - x * cw;
    const uint32_t* src = in + (size_t)b * H * row_in + (size_t)x * ldw + c;
    uint32_t* dst = out + (size_t)b * H * row_out + r;
    uint32_t w[H];
#pragma unroll
    for (int y = 0; y < H; ++y) w[y] = __ldg(src + (size_t)y * row_in);
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int y = 0; y < H; ++y) {
      m0 = fmaxf(m0, __uint_as_float(w[y] << 16));
      m1 = fmaxf(m1, __uint_as_float(w[y] & 0xffff0000u));
    }
    const float c0 = -m0 * kLog2e, c1 = -m1 * kLog2e;
    float l0 = 0.f, l1 = 0.f;
    float e0[H], e1[H];
#pragma unroll
    for (int y = 0; y < H; ++y) {
      e0[y] = exp2f(fmaf(__uint_as_float(w[y] << 16), kLog2e, c0));
      e1[y] = exp2f(fmaf(__uint_as_float(w[y] & 0xffff0000u), kLog2e, c1));
      l0 += e0[y];
      l1 += e1[y];
    }
    const float r0 = scale / l0, r1 = scale / l1;
#pragma unroll
    for (int y = 0; y < H; ++y) {
      const __nv_bfloat162 t = __floats2bfloat162_rn(e0[y] * r0, e1[y] * r1);
      dst[(size_t)y * row_out] = *reinterpret_cast<const uint32_t*>(&t);
    }
  }
}
// Tall images (whole-scene mode: H = 128 .. 512): the column of one (b, x, channel pair) is split over TY = H / 32 threads of a CTA that
// owns one (b, x): each thread keeps its 32 rows packed in registers, the column maximum and the sum of exponentials are combined through
// shared memory, and HBM is read once and written once.  The round-1 fall-back walked a whole column per thread (H serial strided loads,
// two passes, B*W*C/8 threads in total): 175 GB/s at B = 1, 23 % of a whole-scene denoise step (profiles/r02_whole_scene_profile_before.txt).
static constexpr int kSmRows = 32;
__global__ void __launch_bounds__(768) softmax_h_col_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int H, int W, int cw, int ldw,
                                                             float scale) {
  pdl_wait();
  __shared__ float2 red[16 * 128];
  constexpr float kLog2e = 1.4426950408889634f;
  const int c = threadIdx.x, ty = threadIdx.y, TY = blockDim.y;
  const int x = blockIdx.x;
  const size_t b = blockIdx.y;
  const size_t row_in = (size_t)W * ldw, row_out = (size_t)W * cw;
  const uint32_t* src = in + (b * H + ty) * row_in + (size_t)x * ldw + c;
  uint32_t* dst = out + (b * H + ty) * row_out + (size_t)x * cw + c;
  uint32_t w[kSmRows];
#pragma unroll
  for (int i = 0; i < kSmRows; ++i) w[i] = __ldg(src + (size_t)i * TY * row_in);  // rows ty, ty + TY, ...
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < kSmRows; ++i) {
    m0 = fmaxf(m0, __uint_as_float(w[i] << 16));
    m1 = fmaxf(m1, __uint_as_float(w[i] & 0xffff0000u));
  }
  red[ty * cw + c] = make_float2(m0, m1);
  __syncthreads();
  for (int t = 0; t < TY; ++t) {
    const float2 v = red[t * cw + c];
    m0 = fmaxf(m0, v.x);
    m1 = fmaxf(m1, v.y);
  }
  __syncthreads();
  const float c0 = -m0 * kLog2e, c1 = -m1 * kLog2e;
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int i = 0; i < kSmRows; ++i) {
    l0 += exp2f(fmaf(__uint_as_float(w[i] << 16), kLog2e, c0));
    l1 += exp2f(fmaf(__uint_as_float(w[i] & 0xffff0000u), kLog2e, c1));
  }
  red[ty * cw + c] = make_float2(l0, l1);
  __syncthreads();
  l0 = 0.f;
  l1 = 0.f;
  for (int t = 0; t < TY; ++t) {  // same order in every thread of the column: identical totals
    const float2 v = red[t * cw + c];
    l0 += v.x;
    l1 += v.y;
  }
  const float r0 = scale / l0, r1 = scale / l1;
#pragma unroll
  for (int i = 0; i < kSmRows; ++i) {  // the exponentials are recomputed (MUFU is idle, registers are not: up to 768 threads per CTA)
    const __nv_bfloat162 t = __floats2bfloat162_rn(exp2f(fmaf(__uint_as_float(w[i] << 16), kLog2e, c0)) * r0,
                                                   exp2f(fmaf(__uint_as_float(w[i] & 0xffff0000u), kLog2e, c1)) * r1);
    dst[(size_t)i * TY * row_out] = *reinterpret_cast<const uint32_t*>(&t);
  }
}

template <int H>
static int launch_softmax_h_reg(const ddif_softmax_h_t& p, cudaStream_t s) {
  const int ld = (int)(p.in_ld ? p.in_ld : p.c);
  const int64_t items = p.batch * p.w * (p.c / 2);
  DDIF_CUDA_CHECK(launch_pdl(softmax_h_reg_kernel<H>, dim3(grid_for(items, 128, 148 * 24)), dim3(128), (size_t)0, s, (const uint32_t*)p.in,
                             (uint32_t*)p.out, items, (int)p.w, (int)(p.c / 2), ld / 2, (float)p.scale));
  return DDIF_OK;
}
int launch_softmax_h(const ddif_softmax_h_t& p, cudaStream_t s) {
  if (p.c % 8 != 0 || p.in_ld % 8 != 0 || (p.in_ld && p.in_ld < p.c)) return DDIF_ERR_SHAPE;
  if (p.h == 64) return launch_softmax_h_reg<64>(p, s);
  if (p.h == 32) return launch_softmax_h_reg<32>(p, s);
  if (p.h == 16) return launch_softmax_h_reg<16>(p, s);
  if (p.h == 8) return launch_softmax_h_reg<8>(p, s);
  if (p.h > 64 && p.h % kSmRows == 0 && p.h <= 16 * kSmRows && (p.c / 2) <= 128 && (p.c / 2) * (p.h / kSmRows) <= 768 && p.c % 2 == 0 && p.w <= 65535 &&
      p.batch <= 65535) {
    const int ld = (int)(p.in_ld ? p.in_ld : p.c);
    DDIF_CUDA_CHECK(launch_pdl(softmax_h_col_kernel, dim3((unsigned)p.w, (unsigned)p.batch), dim3((unsigned)(p.c / 2), (unsigned)(p.h / kSmRows)), (size_t)0, s,
                               (const uint32_t*)p.in, (uint32_t*)p.out, (int)p.h, (int)p.w, (int)(p.c / 2), ld / 2, (float)p.scale));
    return DDIF_OK;
  }
  DDIF_CUDA_CHECK(launch_pdl(softmax_h_kernel, dim3(grid_for(p.batch * p.w * (p.c / 8), 128)), dim3(128), (size_t)0, s, (const bf16*)p.in, (bf16*)p.out,
                             (int)p.batch, (int)p.h, (int)p.w, (int)p.c, (int)(p.in_ld ? p.in_ld : p.c), (float)p.scale));
  return DDIF_OK;
}

// ---- self-attention core (sr3_dwt.py:347-357): flash-style, one block per (b, head, 64-query tile) ---------
template <int HD>
__global__ void attn_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int ntok, int C, int heads, float scale) {
  pdl_wait();
  __shared__ float sk[64][HD];
  __shared__ float sv[64][HD];
  const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * 64;
  const int tid = threadIdx.x;  // 64 threads
  const int C3 = 3 * C;
  const bf16* base = qkv + (size_t)b * ntok * C3 + head * 3 * HD;
  const int qi = q0 + tid;
  float q[HD], acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    q[d] = (qi < ntok) ? __bfloat162float(base[(size_t)qi * C3 + d]) * scale : 0.f;
    acc[d] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < ntok; k0 += 64) {
    __syncthreads();
    for (int i = tid; i < 64 * HD; i += 64) {
      const int j = i / HD, d = i - j * HD;
      const int kj = k0 + j;
      sk[j][d] = (kj < ntok) ? __bfloat162float(base[(size_t)kj * C3 + HD + d]) : 0.f;
      sv[j][d] = (kj < ntok) ? __bfloat162float(base[(size_t)kj * C3 + 2 * HD + d]) : 0.f;
    }
    __syncthreads();
    const int kn = min(64, ntok - k0);
    for (int j = 0; j < kn; ++j) {
      float sdot = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) sdot += q[d] * sk[j][d];
      const float mn = fmaxf(m, sdot);
      const float corr = __expf(m - mn);
      const float pj = __expf(sdot - mn);
      l = l * corr + pj;
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[d] = acc[d] * corr + pj * sv[j][d];
      m = mn;
    }
  }
  if (qi < ntok) {
    const float inv = 1.0f / l;
    bf16* o = out + ((size_t)b * ntok + qi) * C + head * HD;
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = __float2bfloat16(acc[d] * inv);
  }
}
// 64-token specialisation (the UNet attends only at the 8x8 level): one block per (b, head), one thread per query.  The
// token's [q | k | v] head slice is 3*HD contiguous bf16 -> 16-byte loads; K/V live in shared memory as fp32 and are
// read as broadcast float4; the 64 scores stay in registers (independent dot products, one exp2 per score instead of
// the online-softmax's two exps and a serial rescale chain per key).
template <int HD>
__global__ void __launch_bounds__(64) attn64_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int C, float scale_log2e) {
  pdl_wait();
  __shared__ __align__(16) float sk[64][HD];
  __shared__ __align__(16) float sv[64][HD];
  const int b = blockIdx.y, head = blockIdx.x, tid = threadIdx.x;
  const bf16* row = qkv + ((size_t)b * 64 + tid) * (size_t)(3 * C) + head * 3 * HD;
  float q[HD];
#pragma unroll
  for (int c = 0; c < HD / 8; ++c) {
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(row + c * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) q[c * 8 + j] = f[j] * scale_log2e;
    unpack8(*reinterpret_cast<const bf16x8*>(row + HD + c * 8), f);
    *reinterpret_cast<float4*>(&sk[tid][c * 8]) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(&sk[tid][c * 8 + 4]) = make_float4(f[4], f[5], f[6], f[7]);
    unpack8(*reinterpret_cast<const bf16x8*>(row + 2 * HD + c * 8), f);
    *reinterpret_cast<float4*>(&sv[tid][c * 8]) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(&sv[tid][c * 8 + 4]) = make_float4(f[4], f[5], f[6], f[7]);
  }
  __syncthreads();
  float sc[64];
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      const float4 kk = *reinterpret_cast<const float4*>(&sk[j][d]);
      a0 = fmaf(q[d], kk.x, a0);
      a1 = fmaf(q[d + 1], kk.y, a1);
      a2 = fmaf(q[d + 2], kk.z, a2);
      a3 = fmaf(q[d + 3], kk.w, a3);
    }
    sc[j] = (a0 + a1) + (a2 + a3);
    m = fmaxf(m, sc[j]);
  }
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    sc[j] = exp2f(sc[j] - m);
    l += sc[j];
  }
  float acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) acc[d] = 0.f;
#pragma unroll
  for (int j = 0; j < 64; ++j) {
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      const float4 vv = *reinterpret_cast<const float4*>(&sv[j][d]);
      acc[d] = fmaf(sc[j], vv.x, acc[d]);
      acc[d + 1] = fmaf(sc[j], vv.y, acc[d + 1]);
      acc[d + 2] = fmaf(sc[j], vv.z, acc[d + 2]);
      acc[d + 3] = fmaf(sc[j], vv.w, acc[d + 3]);
    }
  }
  const float inv = 1.0f / l;
  bf16* o = out + ((size_t)b * 64 + tid) * C + head * HD;
#pragma unroll
  for (int c = 0; c < HD / 8; ++c) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = acc[c * 8 + j] * inv;
    *reinterpret_cast<bf16x8*>(o + c * 8) = pack8(f);
  }
}
int launch_attn(const ddif_attn_t& p, cudaStream_t s) {
  if (p.heads <= 0 || p.c % p.heads != 0) return DDIF_ERR_SHAPE;
  const int hd = (int)(p.c / p.heads);
  // long token counts (whole-scene mode, non-64x64 inputs): the tcgen05 / TMEM flash kernel (attn_tc.cu)
  if (attn_tc_applicable(p)) return launch_attn_tc(p, s);
  if (p.ntok == 64 && (hd == 16 || hd == 8) && p.c % 8 == 0) {
    const float sl2 = (float)(p.scale * 1.4426950408889634);
    const dim3 g64((unsigned)p.heads, (unsigned)p.batch);
    if (hd == 16)
      DDIF_CUDA_CHECK(launch_pdl(attn64_kernel<16>, g64, dim3(64), (size_t)0, s, (const bf16*)p.qkv, (bf16*)p.out, (int)p.c, sl2));
    else
      DDIF_CUDA_CHECK(launch_pdl(attn64_kernel<8>, g64, dim3(64), (size_t)0, s, (const bf16*)p.qkv, (bf16*)p.out, (int)p.c, sl2));
    return DDIF_OK;
  }
  dim3 grid((unsigned)ceil_div(p.ntok, 64), (unsigned)p.heads, (unsigned)p.batch);
  const bf16* in = (const bf16*)p.qkv;
  bf16* out = (bf16*)p.out;
  switch (hd) {
    case 8: DDIF_CUDA_CHECK(launch_pdl(attn_kernel<8>, dim3(grid), dim3(64), (size_t)(0), s, in, out, (int)p.ntok, (int)p.c, (int)p.heads, (float)p.scale)); break;
    case 16: DDIF_CUDA_CHECK(launch_pdl(attn_kernel<16>, dim3(grid), dim3(64), (size_t)(0), s, in, out, (int)p.ntok, (int)p.c, (int)p.heads, (float)p.scale)); break;
    case 32: DDIF_CUDA_CHECK(launch_pdl(attn_kernel<32>, dim3(grid), dim3(64), (size_t)(0), s, in, out, (int)p.ntok, (int)p.c, (int)p.heads, (float)p.scale)); break;
    case 64: DDIF_CUDA_CHECK(launch_pdl(attn_kernel<64>, dim3(grid), dim3(64), (size_t)(0), s, in, out, (int)p.ntok, (int)p.c, (int)p.heads, (float)p.scale)); break;
    default: return DDIF_ERR_SHAPE;
  }
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

// ---- nearest x2 (sr3_dwt.py:269) ---------------------------------------------------------------------------
// One thread per INPUT 16-byte chunk: one load, four stores (the 2x2 output pixels), 32-bit index math; grid (chunks of a sample, batch).
__global__ void __launch_bounds__(256) upsample2x_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int B, int H, int W, int C) {
  pdl_wait();
  const int nchunk = C >> 3;
  const int per_sample = H * W * nchunk;
  const int b = blockIdx.y;
  const uint4* src = reinterpret_cast<const uint4*>(in) + (size_t)b * per_sample;
  uint4* dst = reinterpret_cast<uint4*>(out) + (size_t)b * per_sample * 4;
  const int OW = 2 * W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_sample; i += gridDim.x * blockDim.x) {
    const int ch = i % nchunk;
    const int px = i / nchunk;
    const int y = px / W, x = px - y * W;
    const uint4 v = src[i];
    const int o = ((2 * y) * OW + 2 * x) * nchunk + ch;
    dst[o] = v;
    dst[o + nchunk] = v;
    dst[o + OW * nchunk] = v;
    dst[o + OW * nchunk + nchunk] = v;
  }
}
int launch_upsample2x(const ddif_upsample2x_t& p, cudaStream_t s) {
  if (p.c % 8 != 0) return DDIF_ERR_SHAPE;
  if (p.batch < 1 || p.batch > 65535 || p.h * p.w * p.c > (1 << 28)) return DDIF_ERR_SHAPE;
  const int per_sample = (int)(p.h * p.w * (p.c / 8));
  int gx = (per_sample + 255) / 256;
  if (gx > 64) gx = 64;
  DDIF_CUDA_CHECK(launch_pdl(upsample2x_kernel, dim3((unsigned)gx, (unsigned)p.batch), dim3(256), (size_t)0, s, (const bf16*)p.in, (bf16*)p.out,
                             (int)p.batch, (int)p.h, (int)p.w, (int)p.c));
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

// ---- CUDA-core direct convolution: checker for the tcgen05 path -------------------------------------------
__global__ void conv_direct_kernel(ddif_conv_direct_t p) {
  const bf16* in = (const bf16*)p.in;
  const bf16* w = (const bf16*)p.w;
  bf16* out = (bf16*)p.out;
  const int N = (int)p.n_valid;
  const int64_t items = p.batch * p.out_h * p.out_w * N;
  const int pad = p.taps == 9 ? 1 : 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < items; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    int64_t r = i / N;
    const int ox = (int)(r % p.out_w); r /= p.out_w;
    const int oy = (int)(r % p.out_h);
    const int b = (int)(r / p.out_h);
    float acc = p.bias ? p.bias[n] : 0.f;
    for (int tap = 0; tap < (int)p.taps; ++tap) {
      const int dy = p.taps == 9 ? tap / 3 : 0, dx = p.taps == 9 ? tap % 3 : 0;
      const int iy = oy * (int)p.stride + dy - pad, ix = ox * (int)p.stride + dx - pad;
      if (iy < 0 || iy >= p.in_h || ix < 0 || ix >= p.in_w) continue;
      const bf16* a = in + (((size_t)b * p.in_h + iy) * p.in_w + ix) * p.in_ld;
      const bf16* ww = w + ((size_t)tap * p.n_pad + n) * p.w_k;
      for (int c = 0; c < (int)p.cin; ++c) acc += __bfloat162float(a[c]) * __bfloat162float(ww[c]);
    }
    if (p.act == 1) acc = acc / (1.0f + __expf(-acc));
    out[(((size_t)b * p.out_h + oy) * p.out_w + ox) * p.out_ld + n] = __float2bfloat16(acc);
  }
}
int launch_conv_direct(const ddif_conv_direct_t& p, cudaStream_t s) {
  if (p.taps != 1 && p.taps != 9) return DDIF_ERR_ARG;
  conv_direct_kernel<<<grid_for(p.batch * p.out_h * p.out_w * p.n_valid, 256, 148 * 32), 256, 0, s>>>(p);
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

// ---- per-sample sum / sum of squares (GroupNorm statistics for tensors not produced by a GEMM epilogue) ----
__global__ void stats_kernel(const bf16* __restrict__ in, double* __restrict__ stats, int64_t per_sample) {
  const int b = blockIdx.y;
  const bf16* src = in + (size_t)b * per_sample;
  float s1 = 0.f, s2 = 0.f;
  const int64_t nvec = per_sample >> 3;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8];
    unpack8(*reinterpret_cast<const bf16x8*>(src + i * 8), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1 += v[j]; s2 += v[j] * v[j]; }
  }
  __shared__ float r1[8], r2[8];
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s1; r2[threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, c = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += r1[i]; c += r2[i]; }
    atomicAdd(stats + 2 * b, a);
    atomicAdd(stats + 2 * b + 1, c);
  }
}
int launch_stats(const ddif_stats_t& p, cudaStream_t s) {
  const int64_t per = p.hw * p.c;
  if (per % 8 != 0) return DDIF_ERR_SHAPE;
  int gx = grid_for(per / 8, 256, 64);
  stats_kernel<<<dim3(gx, (unsigned)p.batch), 256, 0, s>>>((const bf16*)p.in, p.stats, per);
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

// ---- F.interpolate(cond[:, c0:c0+c], size, 'bilinear', align_corners=False) (sr3_dwt.py:661-663) ----------
__device__ __forceinline__ void bilinear_src(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.0f - l1;
}
__global__ void resize_kernel(ddif_resize_t p) {
  const int64_t items = p.batch * p.out_h * p.out_w;
  const float sh = (float)p.h / (float)p.out_h, sw = (float)p.w / (float)p.out_w;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < items; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % p.out_w);
    const int oy = (int)((i / p.out_w) % p.out_h);
    const int b = (int)(i / (p.out_w * p.out_h));
    int y0, y1, x0, x1;
    float hy0, hy1, wx0, wx1;
    bilinear_src(oy, sh, (int)p.h, y0, y1, hy0, hy1);
    bilinear_src(ox, sw, (int)p.w, x0, x1, wx0, wx1);
    bf16* dn = p.dst_nhwc ? (bf16*)p.dst_nhwc + (size_t)i * p.c_pad : nullptr;
    for (int c = 0; c < (int)p.c; ++c) {
      const float* pl = p.src + ((size_t)b * p.c_total + p.c0 + c) * p.h * p.w;
      float v;
      if (p.h == p.out_h && p.w == p.out_w) {
        v = pl[(size_t)oy * p.w + ox];
      } else {
        v = hy0 * (wx0 * pl[(size_t)y0 * p.w + x0] + wx1 * pl[(size_t)y0 * p.w + x1]) +
            hy1 * (wx0 * pl[(size_t)y1 * p.w + x0] + wx1 * pl[(size_t)y1 * p.w + x1]);
      }
      if (dn) dn[c] = __float2bfloat16(v);
      if (p.dst_nchw) p.dst_nchw[(((size_t)b * p.c + c) * p.out_h + oy) * p.out_w + ox] = v;
    }
    if (dn)
      for (int c = (int)p.c; c < (int)p.c_pad; ++c) dn[c] = __float2bfloat16(0.f);
  }
}
int launch_resize(const ddif_resize_t& p, cudaStream_t s) {
  if (p.dst_nhwc && p.c_pad < p.c) return DDIF_ERR_SHAPE;
  resize_kernel<<<grid_for(p.batch * p.out_h * p.out_w, 128), 128, 0, s>>>(p);
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

// ---- FWM cond-only context (sr3_dwt.py:541,546,563): one block per (b, head, row range) ---------------------
//  kv = Conv1x1(DW3x3(c));  k.softmax over W;  ctx[d][e] = sum_n k[d,n] v[e,n]
// rows_per_cta >= H: one CTA owns the whole image and stores its context (patches: deterministic, as in round 1); otherwise the rows are
// split over blockIdx.z and the partial contexts are added with fp32 atomics into a zeroed buffer (whole-scene mode: H = 128 .. 512; one CTA
// per (b, head) walked 512 rows serially on 8 SMs: 127 ms of cond-cache build for a 512x512 scene).
__global__ void fwm_context_kernel(ddif_fwm_context_t p, int rows_per_cta) {
  extern __shared__ float sm[];
  const int b = blockIdx.y, head = blockIdx.x;
  const int H = (int)p.h, W = (int)p.w, cd = (int)p.cd, dim = (int)p.dim;
  const int d = dim / (int)p.heads;
  float* dwb = sm;               // [cd][W]
  float* kb = dwb + cd * W;      // [d][W]
  float* vb = kb + d * W;        // [d][W]
  const int tid = threadIdx.x, nt = blockDim.x;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};  // pairs tid, tid+nt, ... of the d*d context (d <= 32, nt = 256)
  const float* cb = p.c_dec + (size_t)b * cd * H * W;
  const int y_begin = (int)blockIdx.z * rows_per_cta, y_end = min(H, y_begin + rows_per_cta);
  for (int y = y_begin; y < y_end; ++y) {
    for (int i = tid; i < cd * W; i += nt) {
      const int cc = i / W, x = i - cc * W;
      float a = 0.f;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        a += p.kv0_w[cc * 9 + tap] * cb[((size_t)cc * H + yy) * W + xx];
      }
      dwb[i] = a;
    }
    __syncthreads();
    for (int i = tid; i < 2 * d * W; i += nt) {
      const int r = i / W, x = i - r * W;
      const int row = r < d ? head * d + r : dim + head * d + (r - d);
      float a = p.kv1_b[row];
      for (int cc = 0; cc < cd; ++cc) a += p.kv1_w[row * cd + cc] * dwb[cc * W + x];
      if (r < d) kb[r * W + x] = a; else vb[(r - d) * W + x] = a;
    }
    __syncthreads();
    // softmax over W for each k row: one warp per row
    for (int r = tid >> 5; r < d; r += nt >> 5) {
      float mx = -INFINITY;
      for (int x = tid & 31; x < W; x += 32) mx = fmaxf(mx, kb[r * W + x]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int x = tid & 31; x < W; x += 32) {
        const float e = expf(kb[r * W + x] - mx);
        kb[r * W + x] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
      for (int x = tid & 31; x < W; x += 32) kb[r * W + x] *= inv;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int pr = tid + k * nt;
      if (pr < d * d) {
        const int dd = pr / d, e = pr - dd * d;
        float a = 0.f;
        for (int x = 0; x < W; ++x) a += kb[dd * W + x] * vb[e * W + x];
        acc[k] += a;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int pr = tid + k * nt;
    if (pr < d * d) {
      float* dst = p.ctx + ((size_t)b * p.heads + head) * d * d + pr;
      if (gridDim.z == 1) *dst = acc[k];
      else atomicAdd(dst, acc[k]);
    }
  }
}
int launch_fwm_context(const ddif_fwm_context_t& p, cudaStream_t s) {
  if (p.dim % p.heads != 0) return DDIF_ERR_SHAPE;
  const int d = (int)(p.dim / p.heads);
  if (d > 32) return DDIF_ERR_SHAPE;
  const size_t smem = (size_t)(p.cd + 2 * d) * p.w * sizeof(float);
  if (smem > 220 * 1024) return DDIF_ERR_SHAPE;
  if (smem > 48 * 1024) DDIF_CUDA_CHECK(cudaFuncSetAttribute(fwm_context_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int rows = p.h > 64 ? 16 : (int)p.h;  // patches (H <= 64): one CTA per (b, head), plain stores
  const int nz = (int)ceil_div(p.h, rows);
  if (nz > 1) DDIF_CUDA_CHECK(cudaMemsetAsync(p.ctx, 0, (size_t)p.batch * p.dim * d * sizeof(float), s));
  fwm_context_kernel<<<dim3((unsigned)p.heads, (unsigned)p.batch, (unsigned)nz), 256, smem, s>>>(p, rows);
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

// ---- W_eff[b][o][h*d+dd] = scale * sum_e W_out[o][h*d+e] * ctx[b][h][dd][e]   (sr3_dwt.py:561-573) ----------
__global__ void fwm_weff_kernel(ddif_fwm_weff_t p) {
  const int dim = (int)p.dim, d = dim / (int)p.heads;
  const int64_t items = p.batch * p.o * dim;
  bf16* out = (bf16*)p.weff;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < items; i += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % dim);
    const int o = (int)((i / dim) % p.o);
    const int b = (int)(i / ((int64_t)dim * p.o));
    const int h = col / d, dd = col - h * d;
    const float* cx = p.ctx + (((size_t)b * p.heads + h) * d + dd) * d;
    const float* wo = p.w_out + (size_t)o * dim + h * d;
    float a = 0.f;
    for (int e = 0; e < d; ++e) a += wo[e] * cx[e];
    out[((size_t)b * p.o_pad + o) * p.k_pad + col] = __float2bfloat16(a * (float)p.scale);
  }
}
int launch_fwm_weff(const ddif_fwm_weff_t& p, cudaStream_t s) {
  if (p.o_pad < p.o || p.k_pad < p.dim) return DDIF_ERR_SHAPE;
  fwm_weff_kernel<<<grid_for(p.batch * p.o * p.dim, 256), 256, 0, s>>>(p);
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

}  // namespace ddif
