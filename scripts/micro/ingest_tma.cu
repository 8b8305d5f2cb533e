// Micro-benchmark: per-SM ingest rate of TENSOR TMA loads (cp.async.bulk.tensor 3D, SWIZZLE_128B) of a
// (rows=128, K) bf16 h tile, box = (64 cols, 128 rows, CH chunks), vs the 1D bulk numbers in ingest.cu.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory"); } while (!done);
}
__device__ __forceinline__ void tma3(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// tensor: (64, rows_total, nchunks) ; CTA group g reads rows [g*128, +128), all chunks, CH chunks per request
__global__ void ingest(const __grid_constant__ CUtensorMap tm, int group, int nch, int CH, int S, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[16];
  if (threadIdx.x == 0) { for (int s = 0; s < S; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int row0 = (blockIdx.x / group) * 128;
    const uint32_t TB = 16384u * CH;
    const int nreq = nch / CH;
    long long t0 = clock64();
    int issued = 0, total = iters * nreq;
    for (; issued < S && issued < total; ++issued) {
      mbar_expect_tx(&full[issued % S], TB);
      tma3(smem + (size_t)(issued % S) * TB, &tm, &full[issued % S], 0, row0, (issued % nreq) * CH);
    }
    for (int it = 0; it < total; ++it) {
      mbar_wait(&full[it % S], (it / S) & 1);
      if (issued < total) {
        mbar_expect_tx(&full[issued % S], TB);
        tma3(smem + (size_t)(issued % S) * TB, &tm, &full[issued % S], 0, row0, (issued % nreq) * CH);
        ++issued;
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)ptr;
  const int rows = 1024, K = 512, nch = K / 64;
  uint8_t* src; cudaMalloc(&src, (size_t)rows * K * 2); cudaMemset(src, 1, (size_t)rows * K * 2);
  long long* out; cudaMalloc(&out, 256 * 8);
  cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Cfg { int ctas, group, CH, S; const char* name; };
  Cfg cfgs[] = {
    {1, 1, 1, 6, "1 CTA, 16KB box, S=6"}, {1, 1, 1, 3, "1 CTA, 16KB box, S=3"}, {1, 1, 1, 1, "1 CTA, 16KB box, S=1"},
    {1, 1, 2, 3, "1 CTA, 32KB box, S=3"}, {1, 1, 4, 2, "1 CTA, 64KB box, S=2"}, {1, 1, 4, 3, "1 CTA, 64KB box, S=3"},
    {128, 32, 1, 6, "128 CTAs groups of 32, 16KB box, S=6"}, {128, 32, 1, 3, "128 CTAs groups of 32, 16KB box, S=3"},
    {128, 32, 2, 3, "128 CTAs groups of 32, 32KB box, S=3"}, {128, 32, 4, 2, "128 CTAs groups of 32, 64KB box, S=2"},
    {128, 16, 1, 6, "128 CTAs groups of 16, 16KB box, S=6"},
  };
  for (auto& c : cfgs) {
    CUtensorMap tm;
    cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)nch};
    cuuint64_t strides[2] = {(cuuint64_t)K * 2, 128};
    cuuint32_t box[3] = {64, 128, (cuuint32_t)c.CH};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d for %s\n", (int)r, c.name); continue; }
    int iters = 50;
    for (int rep = 0; rep < 2; ++rep)
      ingest<<<c.ctas, 32, (size_t)c.S * c.CH * 16384 + 2048>>>(tm, c.group, nch, c.CH, c.S, iters, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    long long h[256]; cudaMemcpy(h, out, c.ctas * 8, cudaMemcpyDeviceToHost);
    long long mx = 0, mn = 1LL << 60; for (int i = 0; i < c.ctas; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
    double bytes_per_cta = (double)iters * nch * 16384;
    printf("%-45s: %.1f B/clk/SM (slowest) %.1f (fastest); chip %.0f B/clk\n", c.name, bytes_per_cta / mx, bytes_per_cta / mn, bytes_per_cta * c.ctas / mx);
  }
  return 0;
}
