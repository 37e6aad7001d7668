// Microtest: how does tcgen05.mma address a swizzled K-major A operand whose descriptor does NOT start on a swizzle-atom
// boundary and whose 8-row groups are NOT 8 rows apart?  (Decides whether one halo tile in shared memory can feed all
// nine taps of a 3x3 convolution through nine descriptors: start = base + (dy*pitch + dx) rows, SBO = pitch rows.)
//
// Shared memory holds rows r = 0..R-1 of `span` bytes, 16-byte chunk c of row r stored at  r*span + ((c ^ f(r)) << 4)
// with f = the swizzle of the ABSOLUTE address (what TMA writes): SW128 f = r & 7, SW64 f = (r >> 1) & 3, SW32 f = (r >> 2) & 1.
// B is a 16 x 16 identity, so D[m, n] = A[row(m), ks*16 + n] and the expected row is  r0 + (m / 8) * sbo_rows + (m % 8).
#include <cstdio>
#include <cstdlib>
#include "../../dif_pan_b200/csrc/common.cuh"
using namespace ddif;

static constexpr int kRows = 448;

__device__ __host__ inline float aval(int r, int k) { return (float)((r * 7 + k * 3) % 251); }

__global__ void __launch_bounds__(128, 1) k(int span, int r0, int sbo_rows, int bo_mode, int ks, int* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
  uint8_t* smB = sm + 64 * 1024;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nck = span / 16;
  for (int i = threadIdx.x; i < kRows * nck; i += blockDim.x) {
    const int r = i / nck, c = i % nck;
    const int f = span == 128 ? (r & 7) : span == 64 ? ((r >> 1) & 3) : ((r >> 2) & 1);
    bf16* dst = reinterpret_cast<bf16*>(sm + r * span + ((c ^ f) << 4));
    for (int e = 0; e < 8; ++e) dst[e] = __float2bfloat16(aval(r, c * 8 + e));
  }
  for (int i = threadIdx.x; i < 16 * nck; i += blockDim.x) {
    const int r = i / nck, c = i % nck;
    const int f = span == 128 ? (r & 7) : span == 64 ? ((r >> 1) & 3) : ((r >> 2) & 1);
    bf16* dst = reinterpret_cast<bf16*>(smB + r * span + ((c ^ f) << 4));
    for (int e = 0; e < 8; ++e) dst[e] = __float2bfloat16((c * 8 + e) == r ? 1.0f : 0.0f);
  }
  if (warp == 0) {
    if (lane == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncwarp();
    tmem_alloc(&slot, 32);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0 && lane == 0) {
    const uint32_t layout = span == 128 ? 2u : span == 64 ? 4u : 6u;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_start = smem_u32(sm) + (uint32_t)(r0 * span + ks * 32);
    uint64_t da = make_smem_desc(a_start, (uint32_t)(sbo_rows * span), layout);
    if (bo_mode == 1) da |= (uint64_t)((a_start >> 7) & 7u) << 49;
    const uint64_t db = make_smem_desc(smem_u32(smB), 8u * span, layout);
    umma_bf16_ss(tm, da, db, idesc, 0u);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
  }
  __syncthreads();
  tc_fence_after();
  uint32_t r[16];
  tmem_ld16(tm + ((uint32_t)(warp * 32) << 16), r);
  tmem_ld_wait();
  const int m = warp * 32 + lane;
  const int row = r0 + (m / 8) * sbo_rows + (m % 8);
  int bad = 0;
  for (int n = 0; n < 16; ++n) bad += (__uint_as_float(r[n]) != aval(row, ks * 16 + n)) ? 1 : 0;
  atomicAdd(&out[0], bad);
  if (bad && atomicAdd(&out[1], 1) == 0) {  // first failing lane: which row did the hardware read?
    out[2] = m;
    out[3] = (int)__uint_as_float(r[0]);
    out[4] = (int)__uint_as_float(r[1]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 32); }
}

int main() {
  int* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  printf("span r0 sbo_rows bo_mode ks | mismatches(of 2048) [first bad m, d0, d1 -> candidate rows]\n");
  for (int span : {128, 64, 32})
    for (int sbo : {8, 10, 16, 18})
      for (int bo : {0, 1})
        for (int r0 : {0, 1, 2, 3, 7, 8, 9, 10, 11, 12, 20, 21, 22}) {
          int tot = 0, h[8] = {0};
          int hf[8] = {0};
          for (int ks = 0; ks < span / 32; ++ks) {
            cudaMemset(d, 0, 64);
            k<<<1, 128, 100 * 1024>>>(span, r0, sbo, bo, ks, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s (span %d r0 %d sbo %d bo %d ks %d)\n", cudaGetErrorString(e), span, r0, sbo, bo, ks); return 1; }
            cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
            if (h[0] && !tot) for (int i = 0; i < 8; ++i) hf[i] = h[i];
            tot += h[0];
          }
          printf("%3d %2d %2d %d | %5d", span, r0, sbo, bo, tot);
          if (tot) {
            printf("  first bad m=%d d0=%d d1=%d rows:", hf[2], hf[3], hf[4]);
            for (int r = 0; r < kRows; ++r)
              for (int kk = 0; kk < span / 2; kk += 8)
                if ((int)aval(r, kk) == hf[3] && (int)aval(r, kk + 1) == hf[4]) printf(" (r%d,k%d)", r, kk);
          }
          printf("\n");
        }
  return 0;
}
