// Microbenchmark: tcgen05.mma (M128 x N x K16, bf16, SS) issue/completion rate vs N, number of independent accumulators,
// and swizzle span.  One CTA per SM, one issuing thread.  Operand contents are irrelevant (zeros).
#include <cstdio>
#include <cstdlib>
#include "../../dif_pan_b200/csrc/common.cuh"
using namespace ddif;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) k(int N, int nacc, int span, int iters, int same_desc, long long* out, int ts_mode, int M) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0;
  if (warp == 0) {
    if (lane == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncwarp();
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0 && lane == 0) {
    const uint32_t layout = span == 128 ? 2u : span == 64 ? 4u : 6u;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t da0 = make_smem_desc(smem_u32(sm), 8u * span, layout);
    const uint64_t db0 = make_smem_desc(smem_u32(sm) + 96 * 1024, 8u * span, layout);
    const int ksteps = span / 32;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    long long t0 = clock64();
    // lean issue loop: 8 MMAs per iteration, descriptors advanced by adds only
    const uint64_t da1 = da0 + 2, da2 = da0 + 4, da3 = da0 + 6;
    const uint64_t db1 = db0 + 2, db2 = db0 + 4, db3 = db0 + 6;
    const uint32_t d0 = tm, d1 = tm + (nacc > 1 ? (uint32_t)N : 0u);
    if (ts_mode) {
      for (int it = 0; it < iters / 8; ++it) {
        umma_bf16_ts(d0, tm + 256u, db0, idesc, 1u); umma_bf16_ts(d1, tm + 264u, db1, idesc, 1u);
        umma_bf16_ts(d0, tm + 272u, db2, idesc, 1u); umma_bf16_ts(d1, tm + 280u, db3, idesc, 1u);
        umma_bf16_ts(d0, tm + 288u, db0, idesc, 1u); umma_bf16_ts(d1, tm + 296u, db1, idesc, 1u);
        umma_bf16_ts(d0, tm + 304u, db2, idesc, 1u); umma_bf16_ts(d1, tm + 312u, db3, idesc, 1u);
      }
    } else {
      for (int it = 0; it < iters / 8; ++it) {
        umma_bf16_ss(d0, da0, db0, idesc, 1u); umma_bf16_ss(d1, da1, db1, idesc, 1u);
        umma_bf16_ss(d0, da2, db2, idesc, 1u); umma_bf16_ss(d1, da3, db3, idesc, 1u);
        umma_bf16_ss(d0, da0, db0, idesc, 1u); umma_bf16_ss(d1, da1, db1, idesc, 1u);
        umma_bf16_ss(d0, da2, db2, idesc, 1u); umma_bf16_ss(d1, da3, db3, idesc, 1u);
      }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2048;
  printf("N span nacc grid | cycles/MMA issue | cycles/MMA complete | MAC/clk/SM\n");
  for (int M : {64, 128})
  for (int ts : {0, 1})
  for (int grid : {148})
    for (int span : {64, 128})
      for (int N : {32, 64, 128, 256})
        for (int nacc : {1, 2}) {
          if (N * nacc > 512) continue;
          k<<<grid, 128, 200 * 1024>>>(N, nacc, span, iters, 0, d, ts, M);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          const double n = (double)iters;
          printf("M=%d ts=%d %3d %3d %d %3d | %8.1f | %8.1f | %8.1f\n", M, ts, N, span, nacc, grid, h[0] / n, h[1] / n, (double)M * N * 16 / (h[1] / n));
        }
  return 0;
}
