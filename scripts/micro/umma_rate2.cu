// Micro-benchmark: tcgen05.mma (kind::f16, M=128, K=16, SS) issue/execute rate vs N, and the cost of
// tcgen05.commit -> mbarrier -> try_wait round trips, one CTA per SM (optionally all SMs).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory"); } while (!done);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t a) {
  uint64_t d = 0; d |= (uint64_t)((a >> 4) & 0x3FFF); d |= (uint64_t)1 << 16; d |= (uint64_t)64 << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
// mode 0: `per` MMAs then commit+wait, repeated `iters` times.   mode 1: `per` MMAs + commit (no wait; wait lags by 2)
__global__ void __launch_bounds__(160, 1) k(int N, int per, int iters, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[16];
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tptr;
  if (warp >= 1) {
    uint64_t* mybar = bar + (warp - 1) * 4;
    const uint32_t mytm = tm + (uint32_t)((warp - 1) * 128);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t dA = desc(smem_u32(smem)), dB = desc(smem_u32(smem + 32768));
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode == 1 && it >= 2) mbar_wait(&mybar[it & 1], ((it - 2) >> 1) & 1);
      if (elect_one()) {
        for (int j = 0; j < per; ++j) umma(mytm, dA + (uint64_t)((j & 3) * 2), dB + (uint64_t)((j & 3) * 2), idesc, j > 0);
        commit(&mybar[mode == 1 ? (it & 1) : 0]);
      }
      __syncwarp();
      if (mode == 0) mbar_wait(&mybar[0], it & 1);
    }
    if (mode == 1) { for (int it = iters - 2; it < iters; ++it) if (it >= 0) mbar_wait(&mybar[it & 1], (it >> 1) & 1); }
    long long t1 = clock64();
    if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;  // warp 1 (all issuing warps run the same loop)
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}
int main() {
  long long* out; cudaMalloc(&out, 256 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 200;
  for (int W : {1, 2, 4})
    for (int N : {64, 128})
      for (int per : {0, 4, 8, 16, 32}) {
        const int ctas = 1, mode = 1;
        k<<<ctas, 32 * (1 + W), 98 * 1024>>>(N, per, iters, mode, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
        long long h[256]; cudaMemcpy(h, out, ctas * 8, cudaMemcpyDeviceToHost);
        printf("issuing warps=%d N=%3d per=%2d: %7.1f cycles/iter  (%.1f per MMA over all warps)\n", W, N, per, (double)h[0] / iters, per ? (double)h[0] / iters / (per * W) : 0.0);
      }
  return 0;
}
