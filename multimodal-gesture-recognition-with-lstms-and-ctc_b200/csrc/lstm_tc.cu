// Bidirectional Keras-LSTM forward recurrence on tcgen05 tensor cores (sm_100a).
// Replaces the tf.while_loop body behind `Bidirectional(LSTM(H, tanh, hard_sigmoid))`
// (/root/reference/audio_network/speech_lstm_ctc_words.py:56-77, skeletal_lstm_ctc.py:309-331,
// multimodal.py:159-168) for the wide layers, where the per-step h_{t-1} U product
// (B x H x 4H) is far beyond the CUDA cores.
//
// Persistent kernel, both directions concurrently.  CTA = (direction, 128-row batch tile, group of
// 16 hidden units).  The 64 columns of U that produce those units' i,f,c,o gates are split once into
// bf16 hi + lo and stay RESIDENT in shared memory for all T steps (K-major, SWIZZLE_128B; per 64-wide
// K chunk one B tile of 128 rows = [64 hi columns ; 64 lo columns]).  Warp roles:
//   warp 0    step barrier poll, then TMA (3D tensor map, one 32 KB request = [h_hi | h_lo] K-chunk)
//             of h_{t-1} through a shared-memory ring
//   warp 1    tcgen05.mma, fp32 accumulator D[128 x 128] in TMEM:  one N=128 MMA gives
//             h_hi U_hi | h_hi U_lo, one N=64 MMA adds h_lo U_hi  (bf16x3, ~2^-17 relative)
//   warps 2-9 epilogue, thread = (batch row, 8 units): pre-activations P_t from the TMA-fed shared
//             tile, tcgen05.ld of the accumulator, hard_sigmoid / tanh, cell update (c in registers),
//             publish h_t (bf16 hi/lo) to the exchange buffer, arrive on the step barrier, then
//             stage y_t (+ gates, c when training) in shared memory
//   warp 10   all bulk global traffic as TMA tensor copies: P_{t+1} load, y / gates / c stores.
//   warp 11   second tcgen05.mma issuer (odd K chunks, second accumulator set).
// Nothing but the 16 KB h publish and the barrier counter goes through the LSU: per-thread-row
// global accesses (32 lines per warp instruction) used to occupy the LSU for ~5000 cycles per step
// and sat in front of the barrier's release fence and poll (profiles/r01_lstm_tc_trace.txt).
// CTAs of one (direction, batch tile) exchange h through L2 and a monotonic counter
// (red.release / ld.acquire at gpu scope).
//
// Measured TMA behaviour that shapes the ring (scripts/micro/ingest*.cu, B200): one bulk/tensor copy
// request costs ~385 cycles of a per-SM serial engine for any size up to 32 KB (530 cycles at
// 64 KB), so the stage is one 32 KB request, never two 16 KB ones.
#include <stdlib.h>
#include "tc_common.cuh"

namespace gr {

static constexpr int kTcThreads = 384;   // TMA warp, MMA warp A, 8 epilogue warps, IO warp, MMA warp B
static constexpr int kUnits = 16;        // hidden units per CTA
static constexpr int kEU = 8;            // units per epilogue thread
static constexpr int kEpiThreads = 256;   // (16 epilogue warps x 4 units were measured slower: MUFU-bound math, 8-byte publishes)
static constexpr int kMaxStages = 5;     // ring stages incl. the P / gates tile, which serves as the last one
static constexpr uint32_t kTile = 16384; // 128 rows x 128 B
static constexpr uint32_t kStageBytes = 2 * kTile;
static constexpr uint32_t kIoBytes = 32768;  // P / gates tile: [4 gates][128 rows][16 units] fp32

struct LstmTcParams {
  uint8_t* hx;         // exchange: [(K chunk, hi|lo)][(dir, parity, padded batch row)][64 bf16]
  unsigned* counters;  // (2 dir, NBT) x 32
  long long* trace;    // debug: per-step clock64 stamps (GR_TC_TRACE), else null
  int B, T, H, Bpad, UGn, NBT, nchunks, nstages, save;
  int dbg;             // GR_TC_DBG timing experiments: 1 = no h loads, 2 = no MMAs (results invalid)
};

#define TC_TRACE(slot, step) do { if (p.trace && (step) < 128) p.trace[((size_t)blockIdx.x * 128 + (step)) * 16 + (slot)] = clock64(); } while (0)

__device__ __forceinline__ float hsig(float v) { return fminf(fmaxf(0.2f * v + 0.5f, 0.f), 1.f); }
// tanh(x) = 1 - 2/(1 + e^{2x}) on MUFU ex2 + fast division: absolute error ~1e-7 (libm tanhf costs
// ~30 instructions and sat on the per-step critical path)
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = ex2_approx(x * 2.8853900817779268f);
  return 1.0f - __fdividef(2.0f, 1.0f + e);
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// byte offset of units [u0, u0+4) of row r inside a [128 rows][16 fp32] tile written / read by TMA
// with SWIZZLE_64B (64-byte rows; 16-byte chunk index ^= address bits [7,9) = (r >> 1) & 3)
__device__ __forceinline__ uint32_t sw64_off(uint32_t r, uint32_t chunk) { return r * 64u + ((chunk ^ ((r >> 1) & 3u)) << 4); }

__global__ void __launch_bounds__(kTcThreads, 1)
lstm_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmUh, const __grid_constant__ CUtensorMap tmUl,
                   const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmG,
                   const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmC, LstmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nch = p.nchunks, NST = p.nstages;
  uint8_t* Ub = smem;                                  // nch B tiles of 16 KB
  uint8_t* ring = Ub + (size_t)nch * kTile;            // NST stages of [h_hi tile | h_lo tile]
  uint8_t* io = ring + (size_t)NST * kStageBytes;      // P_t in, gates_t out
  uint8_t* ystage = ring + (size_t)(NST - 1) * kStageBytes;   // y_t / c_t staging aliases the last ring
  uint8_t* cstage = ystage + 8192;                            // stage (idle between two steps' MMAs)
  uint64_t* full = reinterpret_cast<uint64_t*>(io + kIoBytes);
  uint64_t* empty = full + kMaxStages;
  uint64_t* ufull = empty + kMaxStages;
  uint64_t* tmem_full = ufull + 1;
  uint64_t* p_full = tmem_full + 1;        // P_t tile landed
  uint64_t* stage_ready = p_full + 1;      // 256 epilogue threads staged step t's outputs
  uint64_t* out_done = stage_ready + 1;    // step t's TMA stores have read their staging
  uint64_t* p_consumed = out_done + 1;     // 256 epilogue threads hold P_t in registers: the tile may carry h
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(p_consumed + 1);
  // The P / gates tile is idle from the moment the epilogue has P_t in registers until the step's
  // MMAs are done, so during the MMA phase it is ring stage NST (the stages are contiguous).
  const int NSTT = NST + 1;
  // chunk c -> ring stage c % NSTT, its rank among the step's uses of that stage, and the stage's uses
  // per step: tabulated once (runtime divisions sat on the TMA / MMA issue paths)
  __shared__ int tab_st[16], tab_m[16], tab_u[16];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ug = blockIdx.x % p.UGn;
  const int bt = (blockIdx.x / p.UGn) % p.NBT;
  const int dir = blockIdx.x / (p.UGn * p.NBT);
  const int j0 = ug * kUnits;
  const int H = p.H, T = p.T;
  unsigned* ctr = p.counters + (dir * p.NBT + bt) * 32;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmUh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmUl)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmH)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmG)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(ufull, 1);
    mbar_init(tmem_full, 2);
    mbar_init(p_full, 1);
    mbar_init(stage_ready, kEpiThreads);
    mbar_init(out_done, 1);
    mbar_init(p_consumed, kEpiThreads);
    for (int c = 0; c < 16; ++c) {
      const int st = c % NSTT;
      tab_st[c] = st; tab_m[c] = c / NSTT; tab_u[c] = (nch - st + NSTT - 1) / NSTT;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // The whole warp runs the loops with warp-uniform operands and ONE elected lane issues: a
    // single-lane branch makes ptxas wrap every UTMALDG / UTCHMMA in an R2UR "waterfall" loop.
    if (elect_one()) {
      // ---- one-time: this CTA's columns of U, all K chunks.  B-tile row = part*64 + gate*16 + unit
      mbar_expect_tx(ufull, (uint32_t)nch * kTile);
      for (int c = 0; c < nch; ++c)
        for (int g = 0; g < 4; ++g) {
          const int row = dir * 4 * H + g * H + j0;
          uint8_t* dst = Ub + (size_t)c * kTile + g * (kUnits * 128);
          tma_load_2d(dst, &tmUh, ufull, c * kBK, row);
          tma_load_2d(dst + 64 * 128, &tmUl, ufull, c * kBK, row);
        }
    }
    __syncwarp();
    // ---- per step: stream h_{t-1} of this batch tile
    for (int s = 1; s < T; ++s) {
      // step barrier: one monotonic counter per (direction, batch tile).  (Per-CTA flags in one line
      // polled warp-wide were tried: 32 writers + 32x32 pollers on one L2 line took 2.4x longer.)
      const unsigned target = (unsigned)s * p.UGn;
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      } while (v < target);
      asm volatile("fence.proxy.async;" ::: "memory");
      if (lane == 0) TC_TRACE(0, s);
      mbar_wait(out_done, (uint32_t)((s - 1) & 1));   // ring stage NST-1 doubled as y/c staging
      const int row = (dir * 2 + ((s + 1) & 1)) * p.Bpad + bt * 128;
      for (int c = 0; c < nch; ++c) {
        // chunk c -> stage c % NSTT (restarting every step); its use index gives the phase
        const int st = tab_st[c];
        const uint32_t use = (uint32_t)(s - 1) * (uint32_t)tab_u[c] + (uint32_t)tab_m[c];
        mbar_wait(&empty[st], (use & 1) ^ 1);
        if (st == NST) mbar_wait(p_consumed, (uint32_t)(s & 1));
        if (elect_one()) {
          if (p.dbg & 1) {
            mbar_arrive(&full[st]);
          } else {
            mbar_expect_tx(&full[st], kStageBytes);
            tma_load_3d(ring + (size_t)st * kStageBytes, &tmH, &full[st], 0, row, 2 * c);
          }
        }
        __syncwarp();
      }
      if (lane == 0) TC_TRACE(1, s);
    }
  } else if (warp == 1 || warp == 11) {
    // Two issuing warps (even / odd K chunks, one accumulator set each): a single thread issues a
    // small-N UTCHMMA only every ~45 cycles and a tcgen05.commit costs it ~375 cycles, so one
    // issuer left the tensor pipe ~50% idle (scripts/micro/umma_rate*.cu).
    const int par = warp == 1 ? 0 : 1;
    const uint32_t acc = tmem_base + (uint32_t)(par * 128);
    // D = F32; A, B = BF16; M = 128; N = 128 (h_hi x [U_hi | U_lo]) and N = 64 (h_lo x U_hi)
    constexpr uint32_t idesc128 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idesc64 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    mbar_wait(ufull, 0);
    for (int s = 1; s < T; ++s) {
      for (int c = 0; c < nch; ++c) {
        const int st = tab_st[c];
        const uint32_t ph = ((uint32_t)(s - 1) * (uint32_t)tab_u[c] + (uint32_t)tab_m[c]) & 1;
        // both warps observe every phase of every stage in order (a parity wait may lag the barrier
        // by at most one phase), the other warp's chunks are then skipped
        mbar_wait(&full[st], ph);
        if ((c & 1) != par) continue;
        if (lane == 0) { if (c == 0) TC_TRACE(2, s); if (c < 8) TC_TRACE(9 + (c >> 1), s); }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(ring + (size_t)st * kStageBytes);
        const uint64_t dAh = make_sw128_desc(sa);
        const uint64_t dAl = make_sw128_desc(sa + kTile);
        const uint64_t dB = make_sw128_desc(smem_u32(Ub + (size_t)c * kTile));
        if (elect_one()) {
          if (!(p.dbg & 2)) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              const uint64_t adv = (uint64_t)(k * 2);
              umma_bf16(acc, dAh + adv, dB + adv, idesc128, (c > 1 || k > 0) ? 1u : 0u);
              umma_bf16(acc, dAl + adv, dB + adv, idesc64, 1u);
            }
          }
          umma_commit(&empty[st]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(tmem_full);   // this warp's chunks of step s are complete
      __syncwarp();
      if (lane == 0 && par == 0) TC_TRACE(3, s);
    }
  } else if (warp < 10) {
    // ---- epilogue: 8 warps; thread <-> (batch row, half of the 16 units).  Warp w may only touch
    // TMEM lanes 32*(w%4)..+31, so warps w and w+4 share a row group and split the units.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t rl = (uint32_t)(q * 32 + lane);    // row inside the batch tile
    const int ju = j0 + half * kEU;                   // first unit of this thread
    int nu = H - ju;                                  // valid units of this thread (multiple of 4)
    if (nu > kEU) nu = kEU;
    if (nu < 0) nu = 0;
    const uint32_t so0 = sw64_off(rl, (uint32_t)half * 2), so1 = sw64_off(rl, (uint32_t)half * 2 + 1);
    // exchange position of this thread's 8 units (16 B): slab (chunk, part), row, column
    const size_t R = (size_t)4 * p.Bpad;
    const size_t hx_col = (size_t)(ju & 63) * 2;
    const size_t hx_slab = (size_t)(ju >> 6) * 2;
    float c_state[kEU];
#pragma unroll
    for (int u = 0; u < kEU; ++u) c_state[u] = 0.f;
    for (int s = 0; s < T; ++s) {
      float pre[4][kEU];
      mbar_wait(p_full, (uint32_t)(s & 1));
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4 a = *reinterpret_cast<const float4*>(io + g * 8192 + so0);
        const float4 b = *reinterpret_cast<const float4*>(io + g * 8192 + so1);
        pre[g][0] = a.x; pre[g][1] = a.y; pre[g][2] = a.z; pre[g][3] = a.w;
        pre[g][4] = b.x; pre[g][5] = b.y; pre[g][6] = b.z; pre[g][7] = b.w;
      }
      {
        // The tile may be overwritten by the TMA (h chunk, async proxy) as soon as p_consumed completes,
        // so the eight LDS must have RETURNED before the arrive, not merely been issued: consume one
        // word of each.  (Without this some rows read h bytes instead of P: found by the run-to-run
        // determinism check of tests/test_gpu_lstm.py::test_full_size_recurrence_properties.)
        float chk = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) chk += pre[g][0] + pre[g][4];
        asm volatile("" ::"f"(chk) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads before the async-proxy overwrite (free: 0.04 us/step)
      mbar_arrive(p_consumed);
      if (s > 0) {
        mbar_wait(tmem_full, (uint32_t)((s - 1) & 1));
        if (threadIdx.x == 64) TC_TRACE(4, s);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // acc(g,u) = sum over the accumulator sets of columns [g*16+u] (h_hi U_hi + h_lo U_hi) and
        // [64 + g*16+u] (h_hi U_lo)
#pragma unroll
        for (int set = 0; set < 2; ++set) {
          if (set == 1 && nch == 1) break;
          uint32_t v[8][8];
#pragma unroll
          for (int part = 0; part < 2; ++part)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * 128 + part * 64 + g * kUnits + half * kEU);
              uint32_t* w = v[part * 4 + g];
              asm volatile(
                  "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                  : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                  : "r"(taddr)
                  : "memory");
            }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int u = 0; u < kEU; ++u) pre[g][u] += __uint_as_float(v[g][u]) + __uint_as_float(v[4 + g][u]);
        }
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
      if (s + 1 < T) {
        // publish h_t first (it is on the critical path of every CTA of this group)
        uint32_t w0[4], w1[4];
#pragma unroll
        for (int u = 0; u < kEU; u += 2) {
          const float a0 = (u < nu) ? hv[u] : 0.f, a1 = (u + 1 < nu) ? hv[u + 1] : 0.f;
          const __nv_bfloat162 hh = __floats2bfloat162_rn(a0, a1);
          w0[u / 2] = *reinterpret_cast<const uint32_t*>(&hh);
          const float f0 = __uint_as_float(w0[u / 2] << 16), f1 = __uint_as_float(w0[u / 2] & 0xffff0000u);
          const __nv_bfloat162 ll = __floats2bfloat162_rn(a0 - f0, a1 - f1);
          w1[u / 2] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        if (ju < nch * 64) {
          const size_t rowR = (size_t)(dir * 2 + (s & 1)) * p.Bpad + (size_t)bt * 128 + rl;
          uint8_t* dst = p.hx + (hx_slab * R + rowR) * 128 + hx_col;
          *reinterpret_cast<uint4*>(dst) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
          *reinterpret_cast<uint4*>(dst + R * 128) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
        }
        if (threadIdx.x == 64) TC_TRACE(6, s);
        // the release below is cumulative over everything ordered before it by bar.sync, so the
        // 256 publishing threads do not each need a gpu-scope fence (a per-thread fence.acq_rel.gpu
        // here costs 0.75 us/step and changes nothing in 160 determinism runs)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 64) {
          TC_TRACE(7, s);
          asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
          TC_TRACE(8, s);
        }
      }
      // off the critical path: stage the layer output (and, training, the gates / cell state) for
      // the IO warp's TMA stores
      *reinterpret_cast<float4*>(ystage + so0) = make_float4(hv[0], hv[1], hv[2], hv[3]);
      *reinterpret_cast<float4*>(ystage + so1) = make_float4(hv[4], hv[5], hv[6], hv[7]);
      if (p.save) {
        *reinterpret_cast<float4*>(cstage + so0) = make_float4(c_state[0], c_state[1], c_state[2], c_state[3]);
        *reinterpret_cast<float4*>(cstage + so1) = make_float4(c_state[4], c_state[5], c_state[6], c_state[7]);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          *reinterpret_cast<float4*>(io + g * 8192 + so0) = make_float4(pre[g][0], pre[g][1], pre[g][2], pre[g][3]);
          *reinterpret_cast<float4*>(io + g * 8192 + so1) = make_float4(pre[g][4], pre[g][5], pre[g][6], pre[g][7]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(stage_ready);
    }
  } else if (warp == 10) {
    // ---- IO warp: P_t loads and y_t / gates_t / c_t stores, all as 4D tensor copies (box = 16 units
    // x 128 rows x 1 step x 4 gates); out-of-range units / rows are clipped by the tensor maps
    if (lane == 0) {
      const int b0 = bt * 128;
      mbar_expect_tx(p_full, kIoBytes);
      tma_load_4d(io, &tmG, p_full, j0, b0, dir == 0 ? 0 : T - 1, dir * 4);
      for (int s = 0; s < T; ++s) {
        const int t = dir == 0 ? s : T - 1 - s;
        mbar_wait(stage_ready, (uint32_t)(s & 1));
        tma_store_4d(&tmY, ystage, j0, b0, t, dir);
        if (p.save) {
          tma_store_4d(&tmC, cstage, j0, b0, t, dir);
          tma_store_4d(&tmG, io, j0, b0, t, dir * 4);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(out_done);
        if (s + 1 < T) {
          mbar_expect_tx(p_full, kIoBytes);
          tma_load_4d(io, &tmG, p_full, j0, b0, dir == 0 ? s + 1 : T - 2 - s, dir * 4);
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// transposing fp32 -> bf16 hi/lo split of U (gemm.cu)
int split_bf16_t_launch(const float* x, int R, int K, int ldx, __nv_bfloat16* hi, __nv_bfloat16* lo, int ld_out,
                        cudaStream_t s);

struct TcLayout {
  int Bpad, Kp64, Kp8, UGn, NBT, nch, nstages;
  size_t off_hx, off_ut_hi, off_ut_lo, off_trace, total, smem;
};
static TcLayout tc_layout(int B, int H) {
  TcLayout L;
  L.Bpad = (B + 127) / 128 * 128;
  L.Kp64 = (H + 63) / 64 * 64;
  L.Kp8 = (H + 7) / 8 * 8;
  L.UGn = (H + kUnits - 1) / kUnits;
  L.NBT = L.Bpad / 128;
  L.nch = L.Kp64 / 64;
  const long budget = 227 * 1024 - 1024 - 256 - (long)L.nch * kTile - kIoBytes;
  L.nstages = (int)(budget / (long)kStageBytes);
  if (L.nstages > kMaxStages) L.nstages = kMaxStages;
  if (L.nstages > kMaxStages - 1) L.nstages = kMaxStages - 1;
  L.smem = 1024 + (size_t)L.nch * kTile + (size_t)(L.nstages > 0 ? L.nstages : 0) * kStageBytes + kIoBytes + 256;
  size_t o = 1024;
  L.off_hx = o; o += (size_t)2 * L.nch * 4 * L.Bpad * 128;   // (chunk, part) slabs of (dir, parity, row) x 128 B
  const size_t ut = (size_t)8 * H * L.Kp8 * 2;
  L.off_ut_hi = o; o += (ut + 255) & ~(size_t)255;
  L.off_ut_lo = o; o += (ut + 255) & ~(size_t)255;
  L.off_trace = o; o += (size_t)160 * 128 * 16 * 8;
  L.total = o + 256;
  return L;
}

size_t lstm_tc_workspace_bytes(int B, int H) { return tc_layout(B, H).total; }
size_t lstm_tc_trace_offset(int B, int H) { return tc_layout(B, H).off_trace; }

bool lstm_tc_supported(int B, int H) {
  if (H % 4 != 0 || H < 32) return false;
  TcLayout L = tc_layout(B, H);
  if (2 * L.NBT * L.UGn > num_sms()) return false;
  return L.nstages >= 1 && L.nch <= 16;
}

// (units, batch rows, time, variant) view of a (B, T, nvar*H) fp32 tensor; box = 16 x 128 x 1 x nbox
static int make_io_map(CUtensorMap* tm, const float* base, int B, int T, int H, int nvar, int nbox) {
  const cuuint64_t dims[4] = {(cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)T, (cuuint64_t)nvar};
  const cuuint64_t strides[3] = {(cuuint64_t)T * nvar * H * 4, (cuuint64_t)nvar * H * 4, (cuuint64_t)H * 4};
  const cuuint32_t box[4] = {(cuuint32_t)kUnits, 128, 1, (cuuint32_t)nbox};
  return make_map_nd(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
}

int lstm_fwd_tc_launch(float* gates, const float* U, int B, int T, int H, float* y, float* cell, void* workspace,
                       cudaStream_t s) {
  TcLayout L = tc_layout(B, H);
  char* w = static_cast<char*>(workspace);
  LstmTcParams p;
  p.save = cell != nullptr;
  p.hx = reinterpret_cast<uint8_t*>(w + L.off_hx);
  p.counters = reinterpret_cast<unsigned*>(w);
  p.dbg = getenv("GR_TC_DBG") ? atoi(getenv("GR_TC_DBG")) : 0;
  p.trace = getenv("GR_TC_TRACE") ? reinterpret_cast<long long*>(w + L.off_trace) : nullptr;
  p.B = B; p.T = T; p.H = H; p.Bpad = L.Bpad; p.UGn = L.UGn; p.NBT = L.NBT; p.nchunks = L.nch; p.nstages = L.nstages;
  GR_CUDA(cudaMemsetAsync(w, 0, L.off_ut_hi, s));  // counters + exchange buffer (its K padding stays zero)
  __nv_bfloat16* ut_hi = reinterpret_cast<__nv_bfloat16*>(w + L.off_ut_hi);
  __nv_bfloat16* ut_lo = reinterpret_cast<__nv_bfloat16*>(w + L.off_ut_lo);
  for (int d = 0; d < 2; ++d) {
    // U_d (H, 4H) -> U_d^T (4H, Kp8) bf16 hi/lo
    int rc0 = split_bf16_t_launch(U + (size_t)d * H * 4 * H, H, 4 * H, 4 * H, ut_hi + (size_t)d * 4 * H * L.Kp8,
                                  ut_lo + (size_t)d * 4 * H * L.Kp8, L.Kp8, s);
    if (rc0 != GR_OK) return rc0;
  }
  CUtensorMap tUh, tUl, tH, tG, tY, tC;
  int rc;
  if ((rc = make_map(&tUh, ut_hi, (uint64_t)8 * H, L.Kp8, L.Kp8, kUnits)) != GR_OK) return rc;
  if ((rc = make_map(&tUl, ut_lo, (uint64_t)8 * H, L.Kp8, L.Kp8, kUnits)) != GR_OK) return rc;
  {
    const cuuint64_t R = (cuuint64_t)4 * L.Bpad;
    const cuuint64_t dims[3] = {64, R, (cuuint64_t)2 * L.nch};
    const cuuint64_t strides[2] = {128, R * 128};
    const cuuint32_t box[3] = {64, 128, 2};
    if ((rc = make_map_nd(&tH, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, p.hx, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) != GR_OK)
      return rc;
  }
  if ((rc = make_io_map(&tG, gates, B, T, H, 8, 4)) != GR_OK) return rc;
  if ((rc = make_io_map(&tY, y, B, T, H, 2, 1)) != GR_OK) return rc;
  if ((rc = make_io_map(&tC, cell ? cell : y, B, T, H, 2, 1)) != GR_OK) return rc;
  void* args[] = {&tUh, &tUl, &tH, &tG, &tY, &tC, &p};
  GR_CUDA(cudaFuncSetAttribute(lstm_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
  GR_CUDA(cudaLaunchCooperativeKernel((void*)lstm_fwd_tc_kernel, dim3(2 * L.NBT * L.UGn), dim3(kTcThreads), args, L.smem, s));
  return GR_OK;
}

}  // namespace gr
