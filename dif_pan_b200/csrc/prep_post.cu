// Kernels on either side of the sampling loop (SURVEY.md section 8(f) "next" rows N1, N2, N4 and the forward half of S9):
//   fused conditioning prep  raw lms, pan -> cond        /root/reference/dataset/pan_dataset.py:73-142, dataset/hisr.py:48-59,
//                                                        diffusion_engine.py:221-228
//   scene <-> patch batch    tiling / overlap-averaged stitching around diffusion_engine.py:351-505 (the reference feeds whole scenes)
//   training objective       sum_b w[t_b] * l1/l2        diffusion_ddpm_pan.py:725-762
//   per-sample axpby         predict_start_from_noise / predict_v / predict_start_from_v   diffusion_ddpm_pan.py:284-312
//   validation metrics       SAM / ERGAS / PSNR / CC partial sums   utils/_metric_legacy.py:299-346,365
// All memory-bound, fp32 in / fp64 accumulators, coalesced along the pixel axis.
#include "common.cuh"
#include "ddif_internal.h"

namespace ddif {

#define MUL(a, b) __fmul_rn((a), (b))
#define ADD(a, b) __fadd_rn((a), (b))
#define SUB(a, b) __fsub_rn((a), (b))
#define DIV(a, b) __fdiv_rn((a), (b))

static inline int pgrid(int64_t items, int threads = 256) {
  int64_t b = ceil_div(items, threads);
  if (b > 148 * 8) b = 148 * 8;
  return (int)(b < 1 ? 1 : b);
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = l < (int)(blockDim.x >> 5) ? sh[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;  // valid in thread 0
}

// ---- fused conditioning prep ----------------------------------------------------------------------------------------
// One CTA (256 threads) per 32 x 64 pixel tile of one INPUT plane (lms band or pan band); grid (tiles, c + p, batch), 32-bit indexing.  The CTA
// writes everything that derives from that plane: the scaled copy (cond channel) and the bilinear x2 of its Haar sub-bands (LL for an lms band; the
// three detail bands for a pan band) -- so every input element is read once (+ a one-coefficient halo) instead of once per derived channel.
// The coefficients the tile's outputs need -- (32/2 + 2) x (64/2 + 2) per sub-band, with index clamping at the image edges, exactly the clamped
// rows / columns the per-output expressions pick -- are computed ONCE into shared memory (two float2 loads give all four sub-bands of a 2x2
// block); every thread then forms its 4 consecutive outputs per sub-band from 8 shared-memory values.  Round 1 recomputed 8 coefficients per
// thread and channel (~500 instructions per 4 outputs) and ran at 0.21-0.25 of the HBM rate.
static constexpr int kWcTh = 32, kWcTw = 64, kWcCh = kWcTh / 2 + 2, kWcCw = kWcTw / 2 + 2;

__global__ void __launch_bounds__(256) wavelet_cond_kernel(ddif_wavelet_cond_t p) {
  __shared__ float cs[3][kWcCh][kWcCw + 1];
  const int c = (int)p.c, pp = (int)p.p, h = (int)p.h, w = (int)p.w, hh = h / 2, wh = w / 2;
  const int cw = c + 3 * pp, ct = c + pp + cw;
  const int plane = blockIdx.y, b = blockIdx.z;
  const bool is_lms = plane < c;
  const int nsub = is_lms ? 1 : 3;  // sub-bands derived from this plane
  const int tiles_x = (w + kWcTw - 1) / kWcTw;
  const int ty = (int)blockIdx.x / tiles_x, tx = (int)blockIdx.x - ty * tiles_x;
  const int oy0 = ty * kWcTh, ox0 = tx * kWcTw;
  const size_t hw = (size_t)h * w;
  const float dv = (float)(1.0 / p.divisor);  // reciprocal (<= 1.5 ulp from the IEEE division of haar_dwt2_kernel): the divisions made the first
                                              // version of this kernel instruction-bound (FCHK + slow path)
  const float* pl = is_lms ? p.lms + ((size_t)b * c + plane) * hw : p.pan + ((size_t)b * pp + (plane - c)) * hw;
  // coefficient tiles: slot (i, j) = block (clamp(ry0 + i), clamp(rx0 + j)); edge rows / columns repeat like the clamped bilinear taps
  const int ry0 = oy0 / 2 - 1, rx0 = ox0 / 2 - 1;
  for (int s = threadIdx.x; s < kWcCh * kWcCw; s += 256) {
    const int i = s / kWcCw, j = s - i * kWcCw;
    int by = ry0 + i, bx = rx0 + j;
    by = by < 0 ? 0 : (by > hh - 1 ? hh - 1 : by);
    bx = bx < 0 ? 0 : (bx > wh - 1 ? wh - 1 : bx);
    const float2 r0 = *reinterpret_cast<const float2*>(pl + (size_t)(2 * by) * w + 2 * bx);
    const float2 r1 = *reinterpret_cast<const float2*>(pl + (size_t)(2 * by + 1) * w + 2 * bx);
    const float ab = ADD(r0.x, r0.y), cd = ADD(r1.x, r1.y), amb = SUB(r0.x, r0.y), cmd = SUB(r1.x, r1.y);
    if (is_lms) {
      cs[0][i][j] = MUL(MUL(ADD(ab, cd), 0.5f), dv);                       // LL
    } else {
      const float vh = MUL(MUL(SUB(ab, cd), 0.5f), dv), vv = MUL(MUL(ADD(amb, cmd), 0.5f), dv), vd = MUL(MUL(SUB(amb, cmd), 0.5f), dv);
      cs[0][i][j] = vh;                                                    // group 0: cH
      cs[1][i][j] = p.order == 0 ? vd : vv;                                // Pan: h, d, v;  HISR: h, v, d
      cs[2][i][j] = p.order == 0 ? vv : vd;
    }
  }
  __syncthreads();
  for (int qd = threadIdx.x; qd < kWcTh * kWcTw / 4; qd += 256) {  // quads of 4 consecutive output pixels
    const int ly = qd / (kWcTw / 4), q4 = qd - ly * (kWcTw / 4);
    const int oy = oy0 + ly, ox = ox0 + 4 * q4;
    if (oy >= h || ox >= w) continue;                       // w % 4 == 0: a quad is entirely inside or outside
    const size_t px = (size_t)oy * w + ox;
    {  // scaled copy: cond channel `plane`
      const float4 v = *reinterpret_cast<const float4*>(pl + px);
      *reinterpret_cast<float4*>(p.cond + ((size_t)b * ct + plane) * hw + px) = make_float4(MUL(v.x, dv), MUL(v.y, dv), MUL(v.z, dv), MUL(v.w, dv));
    }
    float sy = 0.5f * ((float)oy + 0.5f) - 0.5f;
    if (sy < 0.f) sy = 0.f;
    int y0 = (int)sy;
    if (y0 > hh - 1) y0 = hh - 1;
    const int y1 = y0 + (y0 < hh - 1 ? 1 : 0);
    const float ly1 = sy - (float)y0, ly0 = 1.f - ly1;
    int xi0[4], xi1[4];
    float lx1v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float sx = 0.5f * ((float)(ox + e) + 0.5f) - 0.5f;
      if (sx < 0.f) sx = 0.f;
      int x0 = (int)sx;
      if (x0 > wh - 1) x0 = wh - 1;
      const int x1 = x0 + (x0 < wh - 1 ? 1 : 0);
      lx1v[e] = sx - (float)x0;
      xi0[e] = x0 - rx0;
      xi1[e] = x1 - rx0;
    }
    for (int g = 0; g < nsub; ++g) {
      // wavelet channel k: [LL(lms) x c | pan sub-band 0 x p | sub-band 1 x p | sub-band 2 x p]
      const int k = is_lms ? plane : c + g * pp + (plane - c);
      const float* r0 = cs[g][y0 - ry0];
      const float* r1 = cs[g][y1 - ry0];
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float lx1 = lx1v[e], lx0 = 1.f - lx1;
        const float v00 = r0[xi0[e]], v01 = r0[xi1[e]], v10 = r1[xi0[e]], v11 = r1[xi1[e]];
        o[e] = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);  // same expression as cond_assemble_kernel
      }
      *reinterpret_cast<float4*>(p.cond + ((size_t)b * ct + c + pp + k) * hw + px) = make_float4(o[0], o[1], o[2], o[3]);
      if (p.wav && !(oy & 1)) {  // raw coefficients of blocks (oy/2, ox/2) and (oy/2, ox/2 + 1)
        const float* rr = cs[g][(oy >> 1) - ry0];
        *reinterpret_cast<float2*>(p.wav + (((size_t)b * cw + k) * hh + (oy >> 1)) * wh + (ox >> 1)) = make_float2(rr[(ox >> 1) - rx0], rr[(ox >> 1) + 1 - rx0]);
      }
    }
  }
}
int launch_wavelet_cond(const ddif_wavelet_cond_t& p, cudaStream_t s) {
  if (p.h % 2 || p.w % 4 || p.c < 1 || p.p < 1 || p.order < 0 || p.order > 1 || p.batch > 65535 || p.h * p.w > (1 << 28)) return DDIF_ERR_SHAPE;
  if (p.batch < 1) return DDIF_OK;
  const int64_t tiles = ceil_div(p.h, (int64_t)kWcTh) * ceil_div(p.w, (int64_t)kWcTw);
  const dim3 grid((unsigned)tiles, (unsigned)(p.c + p.p), (unsigned)p.batch);
  wavelet_cond_kernel<<<grid, 256, 0, s>>>(p);
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

// ---- scene <-> patch batch --------------------------------------------------------------------------------------------
__global__ void tile_gather_kernel(ddif_tile_t p) {
  const int64_t items = p.batch * p.ny * p.nx * p.c * p.ph * p.pw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < items; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int64_t x = r % p.pw; r /= p.pw;
    const int64_t y = r % p.ph; r /= p.ph;
    const int64_t ch = r % p.c; r /= p.c;
    const int64_t ix = r % p.nx; r /= p.nx;
    const int64_t iy = r % p.ny;
    const int64_t b = r / p.ny;
    p.tiles[i] = p.scene[((b * p.c + ch) * p.h + iy * p.sy + y) * p.w + ix * p.sx + x];
  }
}
__global__ void tile_stitch_kernel(ddif_tile_t p) {
  const int64_t items = p.batch * p.c * p.h * p.w;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < items; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int64_t x = r % p.w; r /= p.w;
    const int64_t y = r % p.h; r /= p.h;
    const int64_t ch = r % p.c;
    const int64_t b = r / p.c;
    // tiles iy with iy*sy <= y < iy*sy + ph
    int64_t iy1 = y / p.sy; if (iy1 > p.ny - 1) iy1 = p.ny - 1;
    int64_t iy0 = y - p.ph + 1 <= 0 ? 0 : (y - p.ph + p.sy) / p.sy;
    int64_t ix1 = x / p.sx; if (ix1 > p.nx - 1) ix1 = p.nx - 1;
    int64_t ix0 = x - p.pw + 1 <= 0 ? 0 : (x - p.pw + p.sx) / p.sx;
    float acc = 0.f;
    int cnt = 0;
    for (int64_t iy = iy0; iy <= iy1; ++iy)
      for (int64_t ix = ix0; ix <= ix1; ++ix) {
        acc += p.tiles[((((b * p.ny + iy) * p.nx + ix) * p.c + ch) * p.ph + (y - iy * p.sy)) * p.pw + (x - ix * p.sx)];
        ++cnt;
      }
    p.scene[i] = cnt > 1 ? acc / (float)cnt : acc;
  }
}
int launch_tile(const ddif_tile_t& p, cudaStream_t s) {
  if (p.ph > p.h || p.pw > p.w || p.sy < 1 || p.sx < 1 || p.sy > p.ph || p.sx > p.pw) return DDIF_ERR_SHAPE;
  if ((p.ny - 1) * p.sy + p.ph != p.h || (p.nx - 1) * p.sx + p.pw != p.w) return DDIF_ERR_SHAPE;  // tiles cover the scene exactly
  if (p.dir == 0) tile_gather_kernel<<<pgrid(p.batch * p.ny * p.nx * p.c * p.ph * p.pw), 256, 0, s>>>(p);
  else tile_stitch_kernel<<<pgrid(p.batch * p.c * p.h * p.w), 256, 0, s>>>(p);
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

// ---- training objective (forward) -----------------------------------------------------------------------------------------
__global__ void loss_kernel(ddif_loss_t p) {
  __shared__ double sh[8];
  const int64_t n = p.batch * p.chw;
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = SUB(p.a[i], p.b[i]);
    float l = p.squared == 0 ? fabsf(d) : MUL(d, d);
    if (p.weight) l = MUL(l, p.weight[i / p.chw]);
    acc += (double)l;
  }
  const double t = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(p.out, t);
}
int launch_loss(const ddif_loss_t& p, cudaStream_t s) {
  if (p.squared < 0 || p.squared > 1 || !p.out) return DDIF_ERR_ARG;
  loss_kernel<<<pgrid(p.batch * p.chw), 256, 0, s>>>(p);
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

// ---- adaptive DPM-Solver: per-sample error norm (dpm_solver.py:1003-1006), one block per sample ---------------------------------
__global__ void __launch_bounds__(256) dpm_err_kernel(ddif_dpm_err_t p) {
  __shared__ double sh[8];
  const size_t base = (size_t)blockIdx.x * p.chw;
  const float atol = (float)p.atol, rtol = (float)p.rtol;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < p.chw; i += blockDim.x) {
    const float xl = p.x_lower[base + i];
    const float delta = fmaxf(atol, MUL(rtol, fmaxf(fabsf(xl), fabsf(p.x_prev[base + i]))));
    const float e = DIV(SUB(p.x_higher[base + i], xl), delta);
    acc += (double)MUL(e, e);
  }
  const double t = block_sum(acc, sh);
  if (threadIdx.x == 0) p.out[blockIdx.x] = t;
}
int launch_dpm_err(const ddif_dpm_err_t& p, cudaStream_t s) {
  if (p.batch < 1 || p.batch > 0x7fffffff || !p.out) return DDIF_ERR_ARG;
  dpm_err_kernel<<<(unsigned)p.batch, 256, 0, s>>>(p);
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

__global__ void axpby_kernel(ddif_axpby_t p) {
  const int64_t n = p.batch * p.chw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / p.chw;
    p.out[i] = ADD(MUL(p.ca[b], p.x[i]), MUL(p.cb[b], p.y[i]));
  }
}
int launch_axpby(const ddif_axpby_t& p, cudaStream_t s) {
  axpby_kernel<<<pgrid(p.batch * p.chw), 256, 0, s>>>(p);
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

// ---- validation metrics: partial sums per image ---------------------------------------------------------------------------
// grid (1 + c, batch): block (0, b) sums the spectral angles of image b, block (1 + k, b) the six sums of band k; each block owns its
// outputs (plain stores, no atomics), 32-bit pixel indexing, one shared-memory stage for all of a block's sums.
template <int N>
__device__ __forceinline__ void block_sums(double (&v)[N], double* sh /* [N][8] */, double* out) {
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0)
#pragma unroll
    for (int j = 0; j < N; ++j) sh[j * 8 + w] = v[j];
  __syncthreads();
  if (threadIdx.x < N) {
    double r = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += sh[threadIdx.x * 8 + i];
    out[threadIdx.x] = r;
  }
}

__global__ void __launch_bounds__(256) metrics_kernel(ddif_metrics_t p) {
  __shared__ double sh[6 * 8];
  const int b = blockIdx.y;
  const int c = (int)p.c, h = (int)p.h, w = (int)p.w;
  const int hc = h - 1, wc = w - 1;  // bounds cut [0:-1] on both axes
  const int npix = hc * wc;
  const size_t hw = (size_t)h * w;
  const float* ga = p.gt + (size_t)b * c * hw;
  const float* oa = p.out + (size_t)b * c * hw;
  double* out = p.sums + (size_t)b * (2 + 6 * c);
  if (blockIdx.x == 0) {
    double v[2] = {0.0, 0.0};
    for (int i = threadIdx.x; i < npix; i += blockDim.x) {
      const int y = i / wc, off = y * w + (i - y * wc);
      float s1 = 0.f, na = 0.f, nb = 0.f;  // fp32 like the reference's torch float32 reductions over the band axis
      for (int k = 0; k < c; ++k) {
        const float a = ga[k * hw + off], o = oa[k * hw + off];
        s1 += a * o; na += a * a; nb += o * o;
      }
      const float t = sqrtf(na * nb);
      if (t > 0.f) v[1] += 1.0;
      const float ang = acosf(s1 / t);
      if (!isnan(ang)) v[0] += (double)ang;
    }
    block_sums<2>(v, sh, out);
  } else {
    const int k = blockIdx.x - 1;
    const float* gk = ga + k * hw;
    const float* ok = oa + k * hw;
    double v[6] = {0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < npix; i += blockDim.x) {
      const int y = i / wc, off = y * w + (i - y * wc);
      const float a = gk[off], o = ok[off];
      const float d = a - o;
      v[0] += (double)(d * d); v[1] += a; v[2] += o; v[3] += (double)a * a; v[4] += (double)o * o; v[5] += (double)a * o;
    }
    block_sums<6>(v, sh, out + 2 + 6 * k);
  }
}
int launch_metrics(const ddif_metrics_t& p, cudaStream_t s) {
  if (p.h < 2 || p.w < 2 || p.c < 1 || p.c > 1024 || p.batch < 1 || p.batch > 65535 || p.h * p.w > (1 << 30)) return DDIF_ERR_SHAPE;
  metrics_kernel<<<dim3((unsigned)(1 + p.c), (unsigned)p.batch), 256, 0, s>>>(p);
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

// ---- multi-tensor apply (training-loop surround: EMA, grad clip, AdamW) ---------------------------------------------------------
// One block per chunk of one tensor; element-wise, fp32, explicit operation order (no FMA contraction) so that results match the
// eager torch expressions the reference executes tensor by tensor.
__global__ void __launch_bounds__(256) multi_tensor_kernel(ddif_multi_tensor_t p) {
  __shared__ double sh[8];
  const int64_t ti = p.chunks[2 * blockIdx.x], start = p.chunks[2 * blockIdx.x + 1];
  int64_t n = p.sizes[ti] - start;
  if (n > p.chunk) n = p.chunk;
  float* p0 = reinterpret_cast<float*>(p.ptrs[4 * ti]) + start;
  const float* p1 = reinterpret_cast<const float*>(p.ptrs[4 * ti + 1]) + start;
  float* p2 = reinterpret_cast<float*>(p.ptrs[4 * ti + 2]) + start;
  float* p3 = reinterpret_cast<float*>(p.ptrs[4 * ti + 3]) + start;
  const float s0 = (float)p.s0, s1 = (float)p.s1;
  if (p.op == 0) {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) p0[i] = ADD(MUL(p0[i], s0), MUL(p1[i], s1));
  } else if (p.op == 1) {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) p0[i] = p1[i];
  } else if (p.op == 2) {
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += (double)p0[i] * (double)p0[i];
    const double t = block_sum(acc, sh);
    if (threadIdx.x == 0) atomicAdd(p.out, t);
  } else if (p.op == 3) {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) p0[i] = MUL(p0[i], s0);
  } else if (p.op == 4) {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) p0[i] = fminf(fmaxf(p0[i], -s0), s0);
  } else {
    // torch.optim.AdamW, single-tensor expression order (torch/optim/adamw.py): decoupled decay, lerp first moment, mul/addcmul
    // second moment, denom = sqrt(v) / sqrt(bc2) + eps, param -= (lr / bc1) * m / denom
    const float lr = s0, b1 = s1, b2 = (float)p.s2, eps = (float)p.s3, wd = (float)p.s4;
    const float decay = (float)(1.0 - p.s0 * p.s4), w1 = (float)(1.0 - p.s1), w2 = (float)(1.0 - p.s2);
    const float step_size = (float)(p.s0 / p.s5), bc2s = (float)sqrt(p.s6);
    (void)lr; (void)b1; (void)wd;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      const float g = p1[i];
      const float w = MUL(p0[i], decay);
      const float m = ADD(p2[i], MUL(w1, SUB(g, p2[i])));
      const float v = ADD(MUL(p3[i], b2), MUL(MUL(g, g), w2));
      const float denom = ADD(DIV(sqrtf(v), bc2s), eps);
      p2[i] = m;
      p3[i] = v;
      p0[i] = SUB(w, MUL(step_size, DIV(m, denom)));
    }
  }
}
int launch_multi_tensor(const ddif_multi_tensor_t& p, cudaStream_t s) {
  if (p.op < 0 || p.op > 5 || p.chunk < 1 || !p.ptrs || !p.sizes || !p.chunks || (p.op == 2 && !p.out)) return DDIF_ERR_ARG;
  if (p.nchunks < 1) return DDIF_OK;
  multi_tensor_kernel<<<(unsigned)p.nchunks, 256, 0, s>>>(p);
  DDIF_LAUNCH_CHECK();
  return DDIF_OK;
}

}  // namespace ddif
