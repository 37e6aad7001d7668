// Shared accumulator epilogue of the tcgen05 conv kernels (gemm_tc.cu, conv3x3_tc.cu):
//   TMEM -> registers -> bias / FiLM / CSM modulation / residual / SiLU / GroupNorm statistics -> bf16 NHWC or fp32 NCHW.
// One thread = one output pixel (TMEM lane) x the 16-column chunks {cg, cg+NG, cg+2NG, ...} of the accumulator.
#pragma once
#include "common.cuh"

namespace ddif {

struct EpiParams {
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
  int n_valid;
  int batch, out_h, out_w;
};

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

// Operands of the thread's first chunk are fetched before the accumulator is waited for (latency hides behind MMAs).
struct EpiPrefetch {
  uint32_t res[8], sc[8], sh[8];
};

template <int NG>
__device__ __forceinline__ void epilogue_prefetch(const EpiParams& e, EpiPrefetch& pf, int n_tile, int bn, int cg, bool row_ok, size_t pix) {
  const bool wide_io = (e.out_ld % 16 == 0) && (e.res_ld % 16 == 0) && (e.n_valid % 16 == 0);
  const int ng = n_tile * bn + cg * 16;
  const bool full16 = wide_io && row_ok && (e.n_valid - ng >= 16);
  if (full16 && e.residual) ldg256(e.residual + pix * (size_t)e.res_ld + ng, pf.res);
  if (full16 && e.mod) {
    const bf16* m = e.mod + pix * (size_t)(2 * e.n_valid) + ng;
    ldg256(m, pf.sc);
    ldg256(m + e.n_valid, pf.sh);
  }
}

// tmem_acc: TMEM address of (lane quarter base, first column of this accumulator).  `empty_bar` receives one arrive per
// warp as soon as the warp's last TMEM read has completed.  active: warp-uniform.  Returns nothing; accumulates stats.
template <int NG>
__device__ __forceinline__ void epilogue_tile(const EpiParams& e, EpiPrefetch& pf, uint32_t tmem_acc, uint64_t* empty_bar, int bn, int n_tile,
                                              int cg, int lane, bool active, bool row_ok, int b, int y, int x, size_t pix, int stat_sample) {
  const int nchunks = bn >> 4;
  const bool wide_io = (e.out_ld % 16 == 0) && (e.res_ld % 16 == 0) && (e.n_valid % 16 == 0);
  float s1 = 0.f, s2 = 0.f;
  auto process = [&](const uint32_t (&r)[16], int cc) {
    const int c0 = cc * 16;
    const int ng = n_tile * bn + c0;  // global output channel of r[0]
    if (row_ok && ng < e.n_valid) {
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
      const int nrem = e.n_valid - ng;  // >= 1
      const bool full16 = wide_io && nrem >= 16;
      if (e.bias) {
        if (nrem >= 16) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(e.bias + ng + j));
            v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < nrem) v[j] += __ldg(e.bias + ng + j);
        }
      }
      if (e.film) {
        const float* f = e.film + (size_t)b * e.film_ld + ng;
        if (nrem >= 16 && (e.film_ld % 4 == 0)) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(f + j));
            v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < nrem) v[j] += __ldg(f + j);
        }
      }
      if (e.mod) {
        const bf16* m = e.mod + pix * (size_t)(2 * e.n_valid) + ng;
        if (full16) {
          if (cc != cg) {
            ldg256(m, pf.sc);
            ldg256(m + e.n_valid, pf.sh);
          }
          float sc[16], sh[16];
          unpack16(pf.sc, sc);
          unpack16(pf.sh, sh);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = v[j] * (1.0f + sc[j]) + sh[j];
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < nrem) v[j] = v[j] * (1.0f + __bfloat162float(m[j])) + __bfloat162float(m[e.n_valid + j]);
        }
      }
      if (e.residual) {
        const bf16* rs = e.residual + pix * (size_t)e.res_ld + ng;
        if (full16) {
          if (cc != cg) ldg256(rs, pf.res);
          float rr[16];
          unpack16(pf.res, rr);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += rr[j];
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < nrem) v[j] += __bfloat162float(rs[j]);
        }
      }
      if (e.act == 1) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = swish_half(0.5f * v[j]);
      }
      if (e.stats) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < nrem) {
            s1 += v[j];
            s2 += v[j] * v[j];
          }
      }
      if (e.out) {
        bf16* o = e.out + pix * (size_t)e.out_ld + ng;
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
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < nrem) o[j] = __float2bfloat16(v[j]);
        }
      }
      if (e.out_nchw) {
        const size_t hw = (size_t)e.out_h * e.out_w;
        float* o = e.out_nchw + ((size_t)b * e.n_valid + ng) * hw + (size_t)y * e.out_w + x;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < nrem) o[(size_t)j * hw] = v[j];
      }
    }
    };
  for (int cc = cg; cc < nchunks || cc == cg; cc += NG) {
    uint32_t r[16];
    if (active) {
      tmem_ld16(tmem_acc + (uint32_t)(cc * 16), r);
      tmem_ld_wait();
    }
    if (cc + NG >= nchunks) {  // this warp's last TMEM read is done: hand the accumulator back before the math/stores
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar);
    }
#ifndef DDIF_VAR_G_NO_EPI
    process(r, cc);
#endif
  }
  if (e.stats && active) {
    // all rows of one warp belong to one sample
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0 && stat_sample < e.batch) {
      atomicAdd(e.stats + 2 * (size_t)stat_sample, (double)s1);
      atomicAdd(e.stats + 2 * (size_t)stat_sample + 1, (double)s2);
    }
  }
}

}  // namespace ddif
