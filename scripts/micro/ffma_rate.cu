// FFMA vs FFMA2 issue rate per SM sub-partition on sm_100a: W warps per CTA (one CTA per SM), 8 independent
// accumulator chains per thread, register operands only.  Prints cycles per warp-instruction per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float2 a[8], u[8];
  for (int i = 0; i < 8; ++i) { a[i] = make_float2(threadIdx.x * 1e-3f + i, 0.5f * i); u[i] = make_float2(1.0001f + i * 1e-4f, 0.9999f - i * 1e-4f); }
  float2 h = make_float2(out[threadIdx.x & 31], out[(threadIdx.x + 1) & 31]);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) { a[i].x = fmaf(h.x, u[i].x, a[i].x); a[i].y = fmaf(h.y, u[i].y, a[i].y); }   // 2 scalar FFMA
        else a[i] = __ffma2_rn(h, u[i], a[i]);                                                       // 1 FFMA2
      }
  }
  long long t1 = clock64();
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMemset(out, 0, 1 << 22); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  for (int warps : {1, 2, 4, 7, 8, 13, 16}) {
    for (int mode = 0; mode < 2; ++mode) {
      long long c = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, cyc); else k<1><<<148, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double fma_instr_per_warp = (double)iters * 64 * (mode == 0 ? 2 : 1);
      const double warps_per_smsp = (warps + 3) / 4;
      printf("warps/CTA %2d %s: %lld cycles, %.2f cycles per warp-instruction per sub-partition (busiest), %.1f FMA lanes/clk/SM\n", warps,
             mode == 0 ? "FFMA " : "FFMA2", c, c / (fma_instr_per_warp * warps_per_smsp), (double)iters * 64 * 2 * 32 * warps / c);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
