// Does cuTensorMapEncodeTiled accept non-monotonic strides (dims (u, b, t, g) over a (B,T,8H) fp32 tensor)
// and does the box land in smem as [g][b][u] with SWIZZLE_64B?
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, float* out, int j0, int dir4, int t, int b0) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(32768) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(smem)), "l"((uint64_t)&tm), "r"(smem_u32(&bar)), "r"(j0), "r"(b0), "r"(t), "r"(dir4) : "memory");
    uint32_t done;
    do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory"); } while (!done);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)ptr;
  const int B = 200, T = 7, H = 500;
  size_t n = (size_t)B * T * 8 * H;
  std::vector<float> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = (float)i;
  float* d; cudaMalloc(&d, n * 4); cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
  float* out; cudaMalloc(&out, 32768);
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)T, 8};
  cuuint64_t strides[3] = {(cuuint64_t)T * 8 * H * 4, (cuuint64_t)8 * H * 4, (cuuint64_t)H * 4};
  cuuint32_t box[4] = {16, 128, 1, 4};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode (u,b,t,g) permuted strides: %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 0;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  const int j0 = 496, dir4 = 4, t = 3, b0 = 128;
  k<<<1, 256, 34000>>>(tm, out, j0, dir4, t, b0);
  cudaError_t e = cudaDeviceSynchronize(); printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<float> o(8192); cudaMemcpy(o.data(), out, 32768, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int g = 0; g < 4; ++g) for (int b = 0; b < 128; ++b) for (int u = 0; u < 16; ++u) {
    int chunk = u / 4, sw = (b >> 1) & 3;   // SW64: 16B chunk index ^= address bits [7,9) = (row>>1)&3 for 64 B rows
    float got = o[g * 2048 + b * 16 + ((chunk ^ sw) * 4) + (u & 3)];
    float want = (b0 + b < B && j0 + u < H) ? (float)(((size_t)(b0 + b) * T + t) * 8 * H + (size_t)(dir4 + g) * H + j0 + u) : 0.f;
    if (got != want) { if (bad < 5) printf("mismatch g%d b%d u%d got %f want %f\n", g, b, u, got, want); ++bad; }
  }
  printf("mismatches: %d\n", bad);
  return 0;
}
