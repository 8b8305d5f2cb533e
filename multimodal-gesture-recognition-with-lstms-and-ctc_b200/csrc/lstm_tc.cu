// Bidirectional Keras-LSTM forward recurrence on tcgen05 tensor cores (sm_100a).
// Replaces the tf.while_loop body behind `Bidirectional(LSTM(H, tanh, hard_sigmoid))`
// (/root/reference/audio_network/speech_lstm_ctc_words.py:56-77, skeletal_lstm_ctc.py:309-331,
// multimodal.py:159-168) for the wide layers, where the per-step h_{t-1} U product
// (B x H x 4H) is far beyond the CUDA cores.
//
// Persistent kernel, both directions concurrently.  CTA = (direction, 128-row batch tile,
// group of 16 hidden units).  The 64 columns of U that produce those units' i,f,c,o gates are
// split once into bf16 hi + lo and stay RESIDENT in shared memory (K-major, SWIZZLE_128B) for
// all T steps.  Per step:
//   warp 0   waits for the step barrier, then TMA-streams h_{t-1} (bf16 hi+lo, K-chunks of 64)
//   warp 1   issues tcgen05.mma  D[128 x 64] += h_hi U_hi + h_hi U_lo + h_lo U_hi  (fp32 in TMEM)
//   warps2-5 (one thread per batch row) prefetch the pre-activations P_t, tcgen05.ld the
//            accumulator, apply hard_sigmoid/tanh, update c (registers), write y_t and publish
//            h_t as bf16 hi/lo into the exchange buffer, then arrive on the step barrier.
// CTAs of one (direction, batch tile) exchange h through L2 and a monotonic counter.
#include <stdlib.h>
#include "tc_common.cuh"

namespace gr {

static constexpr int kTcThreads = 320;  // TMA warp, MMA warp, 8 epilogue warps
static constexpr int kEU = 8;           // units per epilogue thread
static constexpr int kUnits = 16;       // hidden units per CTA
static constexpr int kNcols = 64;       // 4 gates x 16 units
static constexpr int kStages = 3;

struct LstmTcParams {
  float* gates;       // (B, T, 8H): P in; post-activation gates out when save != 0
  float* y;           // (B, T, 2H)
  float* cell;        // (B, T, 2H) when save != 0
  __nv_bfloat16* hb_hi;  // (2 dir, 2 parity, Bpad, Kp64)
  __nv_bfloat16* hb_lo;
  unsigned* counters;    // (2 dir, NBT) x 32
  int B, T, H, Bpad, Kp64, UGn, NBT, nchunks, save;
  long long* trace;      // debug: per-step clock64 stamps of CTA 0 (GR_TC_TRACE), else null
};

#define TC_TRACE(slot, step) do { if (p.trace && blockIdx.x == 0 && (step) < 128) p.trace[(step) * 8 + (slot)] = clock64(); } while (0)

__device__ __forceinline__ float hsig(float v) { return fminf(fmaxf(0.2f * v + 0.5f, 0.f), 1.f); }
// tanh(x) = 1 - 2/(1 + e^{2x}) on MUFU ex2 + IEEE division: absolute error ~1e-7 (the libm tanhf
// costs ~30 instructions and sat on the per-step critical path of every recurrence step)
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = ex2_approx(x * 2.8853900817779268f);  // e^{2x}
  return 1.0f - __fdividef(2.0f, 1.0f + e);
}

__global__ void __launch_bounds__(kTcThreads, 1)
lstm_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmHh, const __grid_constant__ CUtensorMap tmHl,
                   const __grid_constant__ CUtensorMap tmUh, const __grid_constant__ CUtensorMap tmUl,
                   LstmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nch = p.nchunks;
  const uint32_t u_chunk_bytes = kNcols * kBK * 2;   // 8 KB
  const uint32_t a_bytes = 128 * kBK * 2;            // 16 KB
  uint8_t* Uh = smem;                                // nch * 8 KB
  uint8_t* Ul = Uh + (size_t)nch * u_chunk_bytes;
  uint8_t* stg = Ul + (size_t)nch * u_chunk_bytes;   // kStages * (hi 16 KB + lo 16 KB)
  uint64_t* full = reinterpret_cast<uint64_t*>(stg + (size_t)kStages * 2 * a_bytes);
  uint64_t* empty = full + kStages;
  uint64_t* ufull = empty + kStages;
  uint64_t* tmem_full = ufull + 1;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ug = blockIdx.x % p.UGn;
  const int bt = (blockIdx.x / p.UGn) % p.NBT;
  const int dir = blockIdx.x / (p.UGn * p.NBT);
  const int j0 = ug * kUnits;
  const int H = p.H, T = p.T;
  unsigned* ctr = p.counters + (dir * p.NBT + bt) * 32;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmHh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmHl)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmUh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmUl)) : "memory");
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(ufull, 1);
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    if (lane == 0) {
      // ---- one-time: this CTA's 64 columns of U (4 gates x 16 units), hi and lo, all K chunks
      mbar_expect_tx(ufull, (uint32_t)nch * u_chunk_bytes * 2);
      for (int c = 0; c < nch; ++c)
        for (int g = 0; g < 4; ++g) {
          const int row = dir * 4 * H + g * H + j0;
          tma_load_2d(Uh + (size_t)c * u_chunk_bytes + g * (kUnits * 128), &tmUh, ufull, c * kBK, row);
          tma_load_2d(Ul + (size_t)c * u_chunk_bytes + g * (kUnits * 128), &tmUl, ufull, c * kBK, row);
        }
      // ---- per step: stream h_{t-1} for this batch tile
      int it = 0;
      for (int s = 1; s < T; ++s) {
        const unsigned target = (unsigned)s * p.UGn;
        unsigned v;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < target);
        asm volatile("fence.proxy.async;" ::: "memory");
        TC_TRACE(0, s);
        const int par_prev = (s + 1) & 1;
        const int row = (dir * 2 + par_prev) * p.Bpad + bt * 128;
        for (int c = 0; c < nch; ++c, ++it) {
          const int st = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&empty[st], ph ^ 1);
          uint8_t* dst = stg + (size_t)st * 2 * a_bytes;
          mbar_expect_tx(&full[st], 2 * a_bytes);
          tma_load_2d(dst, &tmHh, &full[st], c * kBK, row);
          tma_load_2d(dst + a_bytes, &tmHl, &full[st], c * kBK, row);
        }
        TC_TRACE(1, s);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kNcols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      mbar_wait(ufull, 0);
      int it = 0;
      for (int s = 1; s < T; ++s) {
        for (int c = 0; c < nch; ++c, ++it) {
          const int st = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&full[st], ph);
          if (c == 0) TC_TRACE(2, s);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(stg + (size_t)st * 2 * a_bytes);
          const uint64_t dAh = make_sw128_desc(sa);
          const uint64_t dAl = make_sw128_desc(sa + a_bytes);
          const uint64_t dBh = make_sw128_desc(smem_u32(Uh + (size_t)c * u_chunk_bytes));
          const uint64_t dBl = make_sw128_desc(smem_u32(Ul + (size_t)c * u_chunk_bytes));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);
            umma_bf16(tmem_base, dAh + adv, dBh + adv, idesc, (c > 0 || k > 0) ? 1u : 0u);
            umma_bf16(tmem_base, dAh + adv, dBl + adv, idesc, 1u);
            umma_bf16(tmem_base, dAl + adv, dBh + adv, idesc, 1u);
          }
          umma_commit(&empty[st]);
        }
        umma_commit(tmem_full);
        TC_TRACE(3, s);
      }
    }
  } else {
    // ---- epilogue: 8 warps; thread <-> (batch row, half of the 16 units).  Warp w may only touch
    // TMEM lanes 32*(w%4)..+31, so warps w and w+4 share a row group and split the units.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int b = bt * 128 + q * 32 + lane;
    const bool bok = b < p.B;
    const size_t G8 = (size_t)8 * H, Y2 = (size_t)2 * H;
    const int ju = j0 + half * kEU;        // first unit of this thread
    int nu = H - ju;                        // valid units for this thread (multiple of 4)
    if (nu > kEU) nu = kEU;
    if (nu < 0) nu = 0;
    float c_state[kEU];
#pragma unroll
    for (int u = 0; u < kEU; ++u) c_state[u] = 0.f;
    for (int s = 0; s < T; ++s) {
      const int t = dir == 0 ? s : T - 1 - s;
      float pre[4][kEU];
      float* grow = p.gates + ((size_t)b * T + t) * G8 + (size_t)dir * 4 * H + ju;
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int u4 = 0; u4 < kEU; u4 += 4) {
          if (bok && u4 < nu) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(grow + (size_t)g * H + u4));
            pre[g][u4] = v.x; pre[g][u4 + 1] = v.y; pre[g][u4 + 2] = v.z; pre[g][u4 + 3] = v.w;
          } else {
            pre[g][u4] = pre[g][u4 + 1] = pre[g][u4 + 2] = pre[g][u4 + 3] = 0.f;
          }
        }
      if (s > 0) {
        mbar_wait(tmem_full, (uint32_t)((s - 1) & 1));
        if (threadIdx.x == 64) TC_TRACE(4, s);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t v[4][kEU];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * kUnits + half * kEU);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
              : "=r"(v[g][0]), "=r"(v[g][1]), "=r"(v[g][2]), "=r"(v[g][3]), "=r"(v[g][4]), "=r"(v[g][5]),
                "=r"(v[g][6]), "=r"(v[g][7])
              : "r"(taddr)
              : "memory");
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int u = 0; u < kEU; ++u) pre[g][u] += __uint_as_float(v[g][u]);
        if (threadIdx.x == 64) TC_TRACE(5, s);
      }
      float hv[kEU];
#pragma unroll
      for (int u = 0; u < kEU; ++u) {
        const float gi = hsig(pre[0][u]);
        const float gf = hsig(pre[1][u]);
        const float gg = tanh_fast(pre[2][u]);
        const float go = hsig(pre[3][u]);
        const float c = gf * c_state[u] + gi * gg;
        c_state[u] = c;
        hv[u] = go * tanh_fast(c);
        pre[0][u] = gi; pre[1][u] = gf; pre[2][u] = gg; pre[3][u] = go;
      }
      if (bok) {
        if (s + 1 < T) {
          // publish h_t first (it is on the critical path of every CTA of this group)
          uint32_t hi_w[kEU / 2], lo_w[kEU / 2];
#pragma unroll
          for (int u = 0; u < kEU; u += 2) {
            const float a0 = (u < nu) ? hv[u] : 0.f, a1 = (u + 1 < nu) ? hv[u + 1] : 0.f;
            const __nv_bfloat16 h0 = __float2bfloat16_rn(a0), h1 = __float2bfloat16_rn(a1);
            const __nv_bfloat16 l0 = __float2bfloat16_rn(a0 - __bfloat162float(h0));
            const __nv_bfloat16 l1 = __float2bfloat16_rn(a1 - __bfloat162float(h1));
            hi_w[u / 2] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            lo_w[u / 2] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
          }
          const size_t off = ((size_t)(dir * 2 + (s & 1)) * p.Bpad + b) * p.Kp64 + ju;
          *reinterpret_cast<uint4*>(p.hb_hi + off) = make_uint4(hi_w[0], hi_w[1], hi_w[2], hi_w[3]);
          *reinterpret_cast<uint4*>(p.hb_lo + off) = make_uint4(lo_w[0], lo_w[1], lo_w[2], lo_w[3]);
        }
      }
      if (threadIdx.x == 64) TC_TRACE(6, s);
      if (s + 1 < T) {
        // the release below is cumulative over everything ordered before it by bar.sync, so the
        // 256 publishing threads do not each need a gpu-scope fence
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 64) TC_TRACE(7, s);
        if (threadIdx.x == 64) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
      }
      if (bok) {
        // off the critical path: the layer output and (training) the saved gates / cell state
        float* yrow = p.y + ((size_t)b * T + t) * Y2 + (size_t)dir * H + ju;
#pragma unroll
        for (int u4 = 0; u4 < kEU; u4 += 4)
          if (u4 < nu) __stcs(reinterpret_cast<float4*>(yrow + u4), make_float4(hv[u4], hv[u4 + 1], hv[u4 + 2], hv[u4 + 3]));
        if (p.save) {
          float* crow = p.cell + ((size_t)b * T + t) * Y2 + (size_t)dir * H + ju;
#pragma unroll
          for (int u4 = 0; u4 < kEU; u4 += 4)
            if (u4 < nu) {
              *reinterpret_cast<float4*>(crow + u4) = make_float4(c_state[u4], c_state[u4 + 1], c_state[u4 + 2], c_state[u4 + 3]);
#pragma unroll
              for (int g = 0; g < 4; ++g)
                *reinterpret_cast<float4*>(grow + (size_t)g * H + u4) =
                    make_float4(pre[g][u4], pre[g][u4 + 1], pre[g][u4 + 2], pre[g][u4 + 3]);
            }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
  }
}

// transposing fp32 -> bf16 hi/lo split (gemm.cu)
int split_bf16_t_launch(const float* x, int R, int K, int ldx, __nv_bfloat16* hi, __nv_bfloat16* lo, int ld_out,
                        cudaStream_t s);

struct TcLayout {
  int Bpad, Kp64, Kp8, UGn, NBT, nch;
  size_t off_hb_hi, off_hb_lo, off_ut_hi, off_ut_lo, off_trace, total;
};
static TcLayout tc_layout(int B, int H) {
  TcLayout L;
  L.Bpad = (B + 127) / 128 * 128;
  L.Kp64 = (H + 63) / 64 * 64;
  L.Kp8 = (H + 7) / 8 * 8;
  L.UGn = (H + kUnits - 1) / kUnits;
  L.NBT = L.Bpad / 128;
  L.nch = L.Kp64 / 64;
  size_t o = 1024;
  const size_t hb = (size_t)4 * L.Bpad * L.Kp64 * 2;
  L.off_hb_hi = o; o += hb;
  L.off_hb_lo = o; o += hb;
  const size_t ut = (size_t)8 * H * L.Kp8 * 2;
  L.off_ut_hi = o; o += (ut + 255) & ~(size_t)255;
  L.off_ut_lo = o; o += (ut + 255) & ~(size_t)255;
  L.off_trace = o; o += 128 * 8 * 8;
  L.total = o + 256;
  return L;
}

size_t lstm_tc_workspace_bytes(int B, int H) { return tc_layout(B, H).total; }
size_t lstm_tc_trace_offset(int B, int H) { return tc_layout(B, H).off_trace; }

bool lstm_tc_supported(int B, int H) {
  if (H % 4 != 0 || H < 32) return false;
  TcLayout L = tc_layout(B, H);
  if (2 * L.NBT * L.UGn > num_sms()) return false;
  const size_t smem = 1024 + (size_t)L.nch * 8192 * 2 + (size_t)kStages * 32768 + 128;
  return smem <= 227 * 1024;
}

int lstm_fwd_tc_launch(float* gates, const float* U, int B, int T, int H, float* y, float* cell, void* workspace,
                       cudaStream_t s) {
  TcLayout L = tc_layout(B, H);
  char* w = static_cast<char*>(workspace);
  LstmTcParams p;
  p.gates = gates; p.y = y; p.cell = cell; p.save = cell != nullptr;
  p.hb_hi = reinterpret_cast<__nv_bfloat16*>(w + L.off_hb_hi);
  p.hb_lo = reinterpret_cast<__nv_bfloat16*>(w + L.off_hb_lo);
  p.counters = reinterpret_cast<unsigned*>(w);
  p.trace = getenv("GR_TC_TRACE") ? reinterpret_cast<long long*>(w + L.off_trace) : nullptr;
  p.B = B; p.T = T; p.H = H; p.Bpad = L.Bpad; p.Kp64 = L.Kp64; p.UGn = L.UGn; p.NBT = L.NBT; p.nchunks = L.nch;
  GR_CUDA(cudaMemsetAsync(w, 0, L.off_ut_hi, s));  // counters + both exchange buffers
  __nv_bfloat16* ut_hi = reinterpret_cast<__nv_bfloat16*>(w + L.off_ut_hi);
  __nv_bfloat16* ut_lo = reinterpret_cast<__nv_bfloat16*>(w + L.off_ut_lo);
  for (int d = 0; d < 2; ++d) {
    // U_d (H, 4H) -> U_d^T (4H, Kp8) bf16 hi/lo
    int rc0 = split_bf16_t_launch(U + (size_t)d * H * 4 * H, H, 4 * H, 4 * H, ut_hi + (size_t)d * 4 * H * L.Kp8,
                                  ut_lo + (size_t)d * 4 * H * L.Kp8, L.Kp8, s);
    if (rc0 != GR_OK) return rc0;
  }
  CUtensorMap tHh, tHl, tUh, tUl;
  int rc;
  if ((rc = make_map(&tHh, p.hb_hi, (uint64_t)4 * L.Bpad, L.Kp64, L.Kp64, 128)) != GR_OK) return rc;
  if ((rc = make_map(&tHl, p.hb_lo, (uint64_t)4 * L.Bpad, L.Kp64, L.Kp64, 128)) != GR_OK) return rc;
  if ((rc = make_map(&tUh, ut_hi, (uint64_t)8 * H, L.Kp8, L.Kp8, kUnits)) != GR_OK) return rc;
  if ((rc = make_map(&tUl, ut_lo, (uint64_t)8 * H, L.Kp8, L.Kp8, kUnits)) != GR_OK) return rc;
  const size_t smem = 1024 + (size_t)L.nch * 8192 * 2 + (size_t)kStages * 32768 + 128;
  GR_CUDA(cudaFuncSetAttribute(lstm_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {&tHh, &tHl, &tUh, &tUl, &p};
  GR_CUDA(cudaLaunchCooperativeKernel((void*)lstm_fwd_tc_kernel, dim3(2 * L.NBT * L.UGn), dim3(kTcThreads), args, smem, s));
  return GR_OK;
}

}  // namespace gr
