// Micro-benchmark: per-SM bulk-copy ingest rate from L2 (cp.async.bulk, 16 KB tiles) when CTAs of a
// group read the SAME 128/256 KB region (as the LSTM recurrence's h broadcast) vs distinct regions.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory"); } while (!done);
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// each CTA: `iters` rounds of loading `ntiles` tiles (tile bytes TB) through a ring of S stages
__global__ void ingest(const uint8_t* src, int group, size_t group_stride, int ntiles, int TB, int S, int iters, int rotate, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[16];
  if (threadIdx.x == 0) { for (int s = 0; s < S; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint8_t* base = src + (size_t)(blockIdx.x / group) * group_stride;
    const int rot = rotate ? (blockIdx.x % group) % ntiles : 0;
    long long t0 = clock64();
    int it = 0, issued = 0, total = iters * ntiles;
    // prime
    for (; issued < S && issued < total; ++issued) {
      mbar_expect_tx(&full[issued % S], TB);
      bulk_load(smem + (size_t)(issued % S) * TB, base + (size_t)((issued + rot) % ntiles) * TB, TB, &full[issued % S]);
    }
    for (; it < total; ++it) {
      mbar_wait(&full[it % S], (it / S) & 1);
      if (issued < total) {
        mbar_expect_tx(&full[issued % S], TB);
        bulk_load(smem + (size_t)(issued % S) * TB, base + (size_t)((issued + rot) % ntiles) * TB, TB, &full[issued % S]);
        ++issued;
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}
int main() {
  uint8_t* src; size_t bytes = 256u << 20; cudaMalloc(&src, bytes); cudaMemset(src, 1, bytes);
  long long* out; cudaMalloc(&out, 256 * 8);
  cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Cfg { int ctas, group, ntiles, TB, S, rotate; const char* name; };
  Cfg cfgs[] = {
    {1, 1, 64, 2048, 12, 0, "1 CTA 2KB S=12"},
    {1, 1, 64, 4096, 12, 0, "1 CTA 4KB S=12"},
    {1, 1, 32, 8192, 12, 0, "1 CTA 8KB S=12"},
    {1, 1, 16, 16384, 6, 0, "1 CTA 16KB S=6"},
    {1, 1, 8, 32768, 4, 0, "1 CTA 32KB S=4"},
    {1, 1, 4, 65536, 2, 0, "1 CTA 64KB S=2"},
    {1, 1, 4, 65536, 3, 0, "1 CTA 64KB S=3"},
    {1, 1, 4, 65536, 1, 0, "1 CTA 64KB S=1"},
    {1, 1, 16, 16384, 1, 0, "1 CTA 16KB S=1"},
    {128, 32, 4, 65536, 3, 0, "128 CTAs groups of 32, 64KB S=3"},
    {128, 32, 4, 65536, 2, 0, "128 CTAs groups of 32, 64KB S=2"},
    {128, 1, 4, 65536, 3, 0, "128 CTAs own, 64KB S=3"},
    {128, 32, 8, 32768, 4, 0, "128 CTAs groups of 32, 32KB S=4"},
    {148, 37, 8, 32768, 4, 0, "148 CTAs groups of 37, 32KB S=4"},
  };
  for (auto& c : cfgs) {
    int iters = 50;
    for (int rep = 0; rep < 2; ++rep)
      ingest<<<c.ctas, 32, (size_t)c.S * c.TB + 1024>>>(src, c.group, (size_t)1 << 20, c.ntiles, c.TB, c.S, iters, c.rotate, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    long long h[256]; cudaMemcpy(h, out, c.ctas * 8, cudaMemcpyDeviceToHost);
    long long mx = 0, mn = 1LL << 60; for (int i = 0; i < c.ctas; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
    double bytes_per_cta = (double)iters * c.ntiles * c.TB;
    printf("%-50s: %.1f B/clk/SM (slowest) %.1f (fastest); chip %.0f B/clk\n", c.name, bytes_per_cta / mx, bytes_per_cta / mn, bytes_per_cta * c.ctas / mx);
  }
  return 0;
}
