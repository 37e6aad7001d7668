// Microbenchmark: how fast can one SM be fed (10 x 18)-pixel halo tiles of an NHWC bf16 tensor [B, 64, 64, C]?
//   mode 0: one TMA 4D box per tile (C x 10 x 18 x 1, hardware swizzle, OOB zero fill), STAGES-deep mbarrier ring
//   mode 1: cp.async 16 B per thread (zfill for the padding), software swizzle, commit-group ring of STAGES tiles
//   mode 2: mode 1 + every thread re-reads its own chunks (LDS), applies a GroupNorm-affine + Swish, writes back (STS)
// 148 persistent CTAs walk the tiles (8 x 16 output pixels each) like the conv kernel would.  Nothing consumes the data.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../dif_pan_b200/csrc/common.cuh"
using namespace ddif;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static constexpr int kStages = 4;
static constexpr int kHaloW = 10, kHaloH = 18, kHaloPx = 180;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(256, 1) feed(const __grid_constant__ CUtensorMap tm, const bf16* src, int C, int B, int mode, int nthreads,
                                               long long* out, float* sink) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
  __shared__ uint64_t full[kStages];
  const int span = C * 2, nck = C / 8;
  const uint32_t stage_bytes = (uint32_t)(((kHaloPx * span) + 1023) & ~1023);
  const int tiles_x = 8, tiles_y = 4, tpi = 32, ntiles = B * tpi;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) mbar_init(&full[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  int count = 0;
  if (mode == 0) {
    if (threadIdx.x == 0) {
      int issued = 0, done = 0;
      for (int t = blockIdx.x; t < ntiles || done < issued; t += gridDim.x) {
        if (t < ntiles) {
          if (issued - done == kStages) {
            mbar_wait(&full[done % kStages], (done / kStages) & 1);
            ++done;
          }
          const int b = t / tpi, r = t % tpi, y0 = (r / tiles_x) * 16, x0 = (r % tiles_x) * 8;
          const int s = issued % kStages;
          mbar_expect_tx(&full[s], (uint32_t)(kHaloPx * span));
          tma_load_4d(&tm, &full[s], sm + s * stage_bytes, 0, x0 - 1, y0 - 1, b);
          ++issued;
        } else {
          mbar_wait(&full[done % kStages], (done / kStages) & 1);
          ++done;
        }
      }
      count = done;
    }
  } else {
    const int lt = threadIdx.x;
    if (lt < nthreads) {
      const int per = (kHaloPx * nck + nthreads - 1) / nthreads;
      float acc = 0.f;
      int inflight = 0;
      int pend_t[kStages];
      auto issue = [&](int t, int s) {
        const int b = t / tpi, r = t % tpi, y0 = (r / tiles_x) * 16 - 1, x0 = (r % tiles_x) * 8 - 1;
        for (int k = 0; k < per; ++k) {
          const int i = lt + k * nthreads;
          if (i < kHaloPx * nck) {
            const int px = i / nck, c = i % nck;
            const int hy = px / kHaloW, hx = px % kHaloW;
            const int gy = y0 + hy, gx = x0 + hx;
            const bool ok = (unsigned)gy < 64u && (unsigned)gx < 64u;
            const bf16* g = src + ((size_t)(b * 64 + (ok ? gy : 0)) * 64 + (ok ? gx : 0)) * C + c * 8;
            const int f = span == 128 ? (px & 7) : span == 64 ? ((px >> 1) & 3) : ((px >> 2) & 1);
            cp_async16(smem_u32(sm + s * stage_bytes + px * span + ((c ^ f) << 4)), g, ok ? 16u : 0u);
          }
        }
        cp_commit();
      };
      auto transform = [&](int s) {
        for (int k = 0; k < per; ++k) {
          const int i = lt + k * nthreads;
          if (i < kHaloPx * nck) {
            const int px = i / nck, c = i % nck;
            const int f = span == 128 ? (px & 7) : span == 64 ? ((px >> 1) & 3) : ((px >> 2) & 1);
            bf16x8* p = reinterpret_cast<bf16x8*>(sm + s * stage_bytes + px * span + ((c ^ f) << 4));
            float v[8];
            unpack8(*p, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = swish_half(fmaf(v[j], 0.37f, 0.11f));
            *p = pack8(v);
            acc += v[0];
          }
        }
      };
      int ti = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++ti) {
        issue(t, ti % kStages);
        ++inflight;
        if (inflight == kStages) {
          cp_wait<kStages - 1>();
          if (mode == 2) transform((ti - (kStages - 1)) % kStages);
          --inflight;
          ++count;
        }
      }
      cp_wait<0>();
      count += inflight;
      if (acc == 123.456f) sink[0] = acc;
      (void)pend_t;
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) {
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = count;
  }
}

int main() {
  const int B = 256;
  PFN_encodeTiled enc = nullptr;
  {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    enc = (PFN_encodeTiled)p;
  }
  long long* d;
  float* sink;
  cudaMalloc(&d, 148 * 16);
  cudaMalloc(&sink, 16);
  cudaFuncSetAttribute(feed, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("mode C threads | ms | GB/s (unique bytes) | cycles/tile/SM\n");
  for (int C : {32, 64}) {
    bf16* src;
    const size_t n = (size_t)B * 64 * 64 * C;
    cudaMalloc(&src, n * 2);
    cudaMemset(src, 0, n * 2);
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)C, 64, 64, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * 64, (cuuint64_t)C * 2 * 64 * 64};
    cuuint32_t box[4] = {(cuuint32_t)C, kHaloW, kHaloH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    for (int mode : {0, 1, 2})
      for (int nthreads : {128, 256}) {
        if (mode == 0 && nthreads != 128) continue;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        feed<<<148, 256, 200 * 1024>>>(tm, src, C, B, mode, nthreads, d, sink);
        cudaEventRecord(e0);
        feed<<<148, 256, 200 * 1024>>>(tm, src, C, B, mode, nthreads, d, sink);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> h(296);
        cudaMemcpy(h.data(), d, 296 * 8, cudaMemcpyDeviceToHost);
        double cyc = 0, tiles = 0;
        for (int i = 0; i < 148; ++i) { cyc += h[2 * i]; tiles += h[2 * i + 1]; }
        printf("%d %3d %3d | %.3f | %7.0f | %7.0f  (tiles %.0f)\n", mode, C, nthreads, ms, n * 2 / ms / 1e6, cyc / tiles, tiles);
      }
    cudaFree(src);
  }
  return 0;
}
