// Microbenchmark: per-SM issue throughput of the instructions the GroupNorm+Swish transform and the conv epilogue are
// made of (MUFU.TANH / EX2 / RCP in f32 and bf16x2, FFMA2, F2FP pack, bf16 unpack, HFMA2.BF16), as a function of the
// number of resident warps.  Output: lane-operations per clock per SM (one "op" = one instruction result lane).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sfu_rate sfu_rate.cu && ./sfu_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int OP>
__device__ __forceinline__ void body(float (&v)[8], uint32_t (&w)[8], uint64_t (&d)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
    if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
    if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
    if (OP == 3) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(w[i]));
    if (OP == 4) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(w[i]));
    if (OP == 5) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
    if (OP == 6) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(d[i]));
    if (OP == 7) {  // f32 pair -> bf16x2; result fed back so that the chain stays live
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(w[i]) : "f"(v[i]));
      v[i] = __uint_as_float(w[i]);
    }
    if (OP == 8) asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(w[i]));
    if (OP == 9) {  // bf16x2 -> f32 halves (shift + and), fed back
      uint32_t lo, hi;
      asm volatile("shl.b32 %0, %1, 16;" : "=r"(lo) : "r"(w[i]));
      asm volatile("and.b32 %0, %1, 0xffff0000;" : "=r"(hi) : "r"(w[i]));
      asm volatile("xor.b32 %0, %1, %2;" : "=r"(w[i]) : "r"(lo), "r"(hi));
    }
    if (OP == 11) asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(d[i]));
    if (OP == 10) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(w[i]));
  }
}

template <int OP>
__global__ void k(float* out, int iters, long long* cyc) {
  float v[8];
  uint32_t w[8];
  uint64_t d[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = 0.001f * (threadIdx.x + i);
    w[i] = 0x3c003c00u + threadIdx.x + i;
    d[i] = ((uint64_t)__float_as_uint(v[i]) << 32) | __float_as_uint(0.5f * v[i]);
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) body<OP>(v, w, d);
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i] + __uint_as_float(w[i]) + __uint_as_float((uint32_t)d[i]) + __uint_as_float((uint32_t)(d[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(const char* name, int lanes_per_inst) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  printf("%-28s", name);
  for (int warps : {1, 4, 8, 16, 32}) {
    k<OP><<<148, warps * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    avg /= 148;
    const double inst = (double)iters * 8 * (OP == 9 ? 3 : 1) * warps;  // warp-instructions per SM
    printf("  w%-2d %6.2f cyc/winst %7.1f lane/clk", warps, avg / inst, inst * 32 * lanes_per_inst / avg);
  }
  printf("\n");
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("tanh.approx.f32", 1);
  run<1>("ex2.approx.ftz.f32", 1);
  run<2>("rcp.approx.ftz.f32", 1);
  run<11>("add.rn.f32x2", 2);
  run<3>("tanh.approx.bf16x2", 2);
  run<10>("tanh.approx.f16x2", 2);
  run<4>("ex2.approx.ftz.bf16x2", 2);
  run<5>("fma.rn.f32", 1);
  run<6>("fma.rn.f32x2", 2);
  run<7>("cvt.rn.bf16x2.f32", 2);
  run<8>("fma.rn.bf16x2", 2);
  run<9>("shl+and+xor", 1);
  return 0;
}
