// Microbenchmark: HBM throughput of 32-byte-per-thread accesses as a function of how a warp's 32 accesses are laid out.
//   mode 0: lane i touches bytes [32 i, 32 i + 32) of a 1 KB block (fully coalesced, what a TMA / smem-staged epilogue gives)
//   mode 1: lane i touches 32 B at pitch 64 B  (row = pixel with 32 channels bf16, thread = row, one 16-channel chunk)
//   mode 2: lane i touches 32 B at pitch 128 B (64-channel rows)            mode 3: pitch 256 B
// In modes 1-3 the OTHER 32-byte pieces of every row are touched by later iterations of the same warp (like the second
// chunk of the conv epilogue), so the total bytes are identical.  read+write copy of `n` bytes, CUDA events.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void k(const uint4* __restrict__ in, uint4* __restrict__ out, size_t rows, int pitch32, int do_read, int do_write) {
  // rows of pitch32*32 bytes; a warp handles 32 consecutive rows, looping over the pitch32 pieces
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const size_t nwarps = (gridDim.x * (size_t)blockDim.x) >> 5;
  for (size_t r0 = warp * 32; r0 < rows; r0 += nwarps * 32) {
    for (int c = 0; c < pitch32; ++c) {
      size_t idx;  // in uint4 (16 B) units: two per 32-byte piece
      if (pitch32 == 0) idx = 0;
      idx = ((r0 + lane) * pitch32 + c) * 2;
      uint4 a = make_uint4(1, 2, 3, 4), b = a;
      if (do_read) { a = in[idx]; b = in[idx + 1]; }
      if (do_write) { out[idx] = a; out[idx + 1] = b; }
      else if (a.x == 0x12345 && b.y == 0x777) out[0] = a;
    }
  }
}
__global__ void kc(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n16, int do_read, int do_write) {
  for (size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 2; i < n16; i += gridDim.x * (size_t)blockDim.x * 2) {
    uint4 a = make_uint4(1, 2, 3, 4), b = a;
    if (do_read) { a = in[i]; b = in[i + 1]; }
    if (do_write) { out[i] = a; out[i + 1] = b; }
    else if (a.x == 0x12345 && b.y == 0x777) out[0] = a;
  }
}
int main() {
  const size_t n = 512ull << 20;
  uint4 *in, *out;
  cudaMalloc(&in, n); cudaMalloc(&out, n);
  cudaMemset(in, 1, n); cudaMemset(out, 0, n);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[3] = {"read+write", "read only", "write only"};
  for (int rw = 0; rw < 3; ++rw) {
    const int rd = rw != 2, wr = rw != 1;
    for (int mode = 0; mode < 4; ++mode) {
      const int pitch32 = mode == 0 ? 0 : (1 << mode);  // 2, 4, 8 pieces of 32 B per row
      float best = 1e9;
      for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) kc<<<148 * 16, 256>>>(in, out, n / 16, rd, wr);
        else k<<<148 * 16, 256>>>(in, out, n / (pitch32 * 32), pitch32, rd, wr);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
      }
      printf("%-10s mode %d (pitch %4d B): %7.3f ms  %7.1f GB/s\n", names[rw], mode, mode == 0 ? 32 : pitch32 * 32, best, (rd + wr) * (double)n / best / 1e6);
    }
  }
  return 0;
}
