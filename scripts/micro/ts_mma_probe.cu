// Probe for the swapped-operand LSTM recurrence (lstm_tcu.cu): tcgen05.mma with the A operand in TMEM
// (the resident recurrent kernel U) and the B operand (h tile) in shared memory.
//   part 1: correctness of (a) the A-in-TMEM packing written with tcgen05.st.32x32b (lane = M row, 32-bit column c =
//           K elements 2c | 2c+1 << 16), (b) the K-major SWIZZLE_128B B tile with N rows, (c) the register layout of
//           tcgen05.ld.16x256b.x4 that the epilogue's gate gather relies on.
//   part 2: issue / execute rate of TS-mode MMAs for N = 16..256 (one issuing thread), against SS mode.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
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
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t a) {
  uint64_t d = 0; d |= (uint64_t)((a >> 4) & 0x3FFF); d |= (uint64_t)1 << 16; d |= (uint64_t)64 << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
__host__ __device__ inline int a_val(int r, int k) { return ((r * 7 + k * 3) % 13) - 6; }
__host__ __device__ inline int b_val(int n, int k) { return ((n * 5 + k * 11) % 9) - 4; }

constexpr int KT = 64, NT = 64;

__global__ void __launch_bounds__(160, 1) probe(float* out_a, float* out_b) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // B tile: NT rows x 64 k bf16, K-major, SWIZZLE_128B
  for (int e = threadIdx.x; e < NT * KT; e += blockDim.x) {
    const int n = e / KT, k = e % KT;
    const uint32_t off = n * 128 + ((((k * 2) >> 4) ^ (n & 7)) << 4) + ((k * 2) & 15);
    *reinterpret_cast<__nv_bfloat16*>(smem + off) = __float2bfloat16((float)b_val(n, k));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tptr;
  if (warp < 4) {
    const int r = warp * 32 + lane;
    for (int j = 0; j < KT / 16; ++j) {
      uint32_t w[8];
      for (int c = 0; c < 8; ++c) {
        const __nv_bfloat16 lo = __float2bfloat16((float)a_val(r, 16 * j + 2 * c)), hi = __float2bfloat16((float)a_val(r, 16 * j + 2 * c + 1));
        w[c] = (uint32_t)(*reinterpret_cast<const uint16_t*>(&lo)) | ((uint32_t)(*reinterpret_cast<const uint16_t*>(&hi)) << 16);
      }
      const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)(8 * j);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 4) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t dB = desc(smem_u32(smem));
    if (elect_one()) {
      for (int j = 0; j < KT / 16; ++j) umma_ts(tm + 256, tm + 8 * j, dB + (uint64_t)(2 * j), idesc, j > 0);
      commit(&bar);
    }
    __syncwarp();
  }
  if (warp < 4) {
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int r = warp * 32 + lane;
    for (int h = 0; h < 2; ++h) {
      uint32_t v[32];
      const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)(256 + 32 * h);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
          "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
            "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int c = 0; c < 32; ++c) out_a[r * NT + 32 * h + c] = __uint_as_float(v[c]);
    }
    for (int G = 0; G < 2; ++G)
      for (int h = 0; h < 2; ++h) {
        uint32_t v[16];
        const uint32_t taddr = tm + ((uint32_t)(warp * 32 + 16 * G) << 16) + (uint32_t)(256 + 32 * h);
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < 16; ++c) out_b[(((warp * 2 + G) * 2 + h) * 32 + lane) * 16 + c] = __uint_as_float(v[c]);
      }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

// part 2: `per` MMAs (K = 16 each) + commit + wait, `iters` times; ts = 1: A from TMEM
__global__ void __launch_bounds__(64, 1) rate(int N, int per, int iters, int ts, long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t dA = desc(smem_u32(smem)), dB = desc(smem_u32(smem + 32768));
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
        if (ts) for (int j = 0; j < per; ++j) umma_ts(tm + 256, tm + (uint32_t)((j & 31) * 8), dB + (uint64_t)((j & 3) * 2), idesc, j > 0);
        else for (int j = 0; j < per; ++j) umma_ss(tm + 256, dA + (uint64_t)((j & 3) * 2), dB + (uint64_t)((j & 3) * 2), idesc, j > 0);
        commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, it & 1);
    }
    long long t1 = clock64();
    if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
  float *oa, *ob;
  cudaMalloc(&oa, 128 * NT * 4); cudaMalloc(&ob, 4 * 2 * 2 * 32 * 16 * 4);
  cudaMemset(oa, 0xff, 128 * NT * 4); cudaMemset(ob, 0xff, 4 * 2 * 2 * 32 * 16 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  probe<<<1, 160, 32 * 1024>>>(oa, ob);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe: err %s\n", cudaGetErrorString(e)); return 1; }
  static float ha[128 * NT], hb[4 * 2 * 2 * 32 * 16], D[128][NT];
  cudaMemcpy(ha, oa, sizeof(ha), cudaMemcpyDeviceToHost); cudaMemcpy(hb, ob, sizeof(hb), cudaMemcpyDeviceToHost);
  for (int r = 0; r < 128; ++r) for (int n = 0; n < NT; ++n) { int s = 0; for (int k = 0; k < KT; ++k) s += a_val(r, k) * b_val(n, k); D[r][n] = (float)s; }
  int bad_a = 0, bad_b = 0;
  for (int r = 0; r < 128; ++r) for (int n = 0; n < NT; ++n) if (ha[r * NT + n] != D[r][n]) { if (bad_a < 8) printf("  32x32b mismatch r=%d n=%d got %g want %g\n", r, n, ha[r * NT + n], D[r][n]); ++bad_a; }
  for (int w = 0; w < 4; ++w) for (int G = 0; G < 2; ++G) for (int h = 0; h < 2; ++h) for (int t = 0; t < 32; ++t) for (int c = 0; c < 16; ++c) {
    const int j = c >> 2, p = (c >> 1) & 1, ee = c & 1;
    const float want = D[32 * w + 16 * G + (t >> 2) + 8 * p][32 * h + 8 * j + 2 * (t & 3) + ee];
    const float got = hb[(((w * 2 + G) * 2 + h) * 32 + t) * 16 + c];
    if (got != want) { if (bad_b < 8) printf("  16x256b mismatch w=%d G=%d h=%d t=%d c=%d got %g want %g\n", w, G, h, t, c, got, want); ++bad_b; }
  }
  printf("TS-mode MMA (A in TMEM via tcgen05.st 32x32b, B = %d-row SW128 tile): 32x32b readback %s (%d bad), 16x256b.x4 layout %s (%d bad)\n",
         NT, bad_a ? "FAIL" : "ok", bad_a, bad_b ? "FAIL" : "ok", bad_b);

  long long* out; cudaMalloc(&out, 256 * 8);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 100;
  for (int ctas : {1, 128})
    for (int ts : {1, 0})
      for (int N : {16, 32, 64, 128, 256})
        for (int per : {8, 64}) {
          rate<<<ctas, 64, 98 * 1024>>>(N, per, iters, ts, out);
          e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("rate: err %s\n", cudaGetErrorString(e)); return 1; }
          long long h[256]; cudaMemcpy(h, out, ctas * 8, cudaMemcpyDeviceToHost);
          long long mx = 0; for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
          printf("ctas=%3d %s N=%3d per=%2d: %8.1f cycles/iter  (%.1f per MMA)\n", ctas, ts ? "TS" : "SS", N, per, (double)mx / iters, (double)mx / iters / per);
        }
  return 0;
}
