// Bidirectional Keras-LSTM forward recurrence, second tcgen05 design: the recurrent kernel U lives in TENSOR
// MEMORY and the operands are swapped.  Replaces the tf.while_loop body behind
// `Bidirectional(LSTM(H, tanh, hard_sigmoid))` (/root/reference/audio_network/speech_lstm_ctc_words.py:56-77,
// skeletal_network/skeletal_lstm_ctc.py:309-331) for the wide layers (H = 300 / 500).
//
// What lstm_tc.cu (round 1) ran into: with U resident in shared memory (128 KB of bf16 hi/lo per 16 units) the
// MMA phase sits at the shared-memory bandwidth floor (h tile written by TMA + read twice as the A operand +
// the U tile read as B), only two or three 32 KB ring stages fit beside U, and a CTA can hold a single batch tile,
// so the ~7000-cycle publish -> counter -> poll -> TMA chain between two steps is fully exposed.
//
// Here, per CTA = (direction, 16 hidden units, up to two batch tiles of NB rows):
//   * A operand = this CTA's 64 gate columns of U as 128 M-rows (bf16 hi rows and lo rows), K-major in TMEM
//     (<= 256 columns, written once with tcgen05.st); tcgen05.mma reads it from TMEM (TS form), so shared memory
//     carries only the streamed h tiles: B operand = [NB batch rows x 64 k] K-major SWIZZLE_128B, N = NB.
//     D[128 x NB] (fp32, TMEM) accumulates A x h_hi^T and A x h_lo^T:  rows 0..63 of a lane group hold
//     U_hi (h_hi + h_lo), rows +8 hold U_lo (h_hi + h_lo); their sum is the bf16x3 product plus the (harmless)
//     lo x lo term.
//   * The freed shared memory holds a 4-stage 32 KB h ring and double-buffered P / y / c / gates staging, and the
//     CTA interleaves TWO independent batch tiles: while tile 0's epilogue publishes h_t and the group's counter
//     round trip runs, the tensor pipe works on tile 1 (and vice versa).
//   * Epilogue: TMEM lane = gate row, column = batch row, so one unit's four gates (x hi/lo rows) sit in eight
//     lanes of one 32-lane quadrant.  tcgen05.ld.32x32b hands thread l its row for 8X columns; a three-stage
//     shuffle butterfly over the 8 threads that share a unit (lane bits 4,3,2) adds the hi and lo rows and turns
//     "one gate row x 8X cells" into "four gates x X cells" per thread.  (tcgen05.ld.16x256b, whose fragment
//     would save the first stage, was measured at ~300 cycles per 16-lane x 8-column atom: 5400 cycles per
//     128-row tile.)
//   * h_t is staged as fp32 in shared memory (the y tile that the IO warp stores with TMA anyway), converted to
//     bf16 hi/lo by one thread per (row, part) and written to the L2 exchange buffer as whole 32-byte sectors.
// Warp roles: 0 = h TMA (polls the step counter of the tile's group), 1 = MMA issue, 2..9 = epilogue,
// 10 = IO (P loads, y / c / gates stores as 4-D TMA tensor copies), 11 idle.
#include <stdlib.h>
#include "tc_common.cuh"

namespace gr {

static constexpr int kUThreads = 384;
static constexpr int kUEpi = 256;
static constexpr int kUUnits = 16;
static constexpr int kURing = 4;
static constexpr uint32_t kUStage = 32768;
static constexpr uint32_t kDBase = 256;   // first accumulator column; A occupies columns [0, 8*KS) <= 256

struct LstmTcuParams {
  uint8_t* hx;                 // exchange: [(K chunk, hi|lo) slab][(dir, parity, padded batch row)][64 bf16]
  unsigned* counters;          // (dir, global tile) x 32 words
  const __nv_bfloat16* ut_hi;  // (8H, Kp8): row = dir*4H + gate*H + unit, K-major
  const __nv_bfloat16* ut_lo;
  long long* trace;
  int B, T, H, Bpad, UGn, NTg, NSB, Kc, KS, Kp8, CP, NREQ, save;
  int store_y;                 // 0: y is not written (inference with an auxiliary destination only)
  int aux_mode;                // 0 none, 1: h tiles are ALSO stored to the auxiliary tensor, 2: reduce-added to it (TMA .add)
  int dbg;                     // GR_TCU_DBG experiments: 1 = MMAs of the next tile are NOT held back behind the epilogue's tcgen05.ld, 2 = no MMAs, 4 = no h loads
};

#define TCU_TRACE(slot, n) do { if (p.trace && (n) < 256) p.trace[((size_t)blockIdx.x * 256 + (n)) * 16 + (slot)] = clock64(); } while (0)

__device__ __forceinline__ float hsig_u(float v) { return fminf(fmaxf(0.2f * v + 0.5f, 0.f), 1.f); }
__device__ __forceinline__ float tanh_u(float x) {
  const float e = ex2_approx(x * 2.8853900817779268f);
  return 1.0f - __fdividef(2.0f, 1.0f + e);
}
__device__ __forceinline__ void tma_load_3d_u(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d_u(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d_u(const CUtensorMap* tm, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// shared -> global tile ADDED to the destination (fp32 add performed at the memory side)
__device__ __forceinline__ void tma_reduce_add_4d_u(const CUtensorMap* tm, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ bool elect_one_u() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive_u(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// A operand from tensor memory (TS form)
__device__ __forceinline__ void umma_ts_bf16(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ float fsel(bool c, float a, float b) {   // c ? a : b as one SEL
  float r;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tselp.f32 %0, %1, %2, p;\n\t}" : "=f"(r) : "f"(a), "f"(b), "r"((uint32_t)c));
  return r;
}
// byte offset of the 16-byte chunk `chunk` of row r inside a [rows][16 fp32] tile moved by TMA with SWIZZLE_64B
__device__ __forceinline__ uint32_t sw64u(uint32_t r, uint32_t chunk) { return r * 64u + ((chunk ^ ((r >> 1) & 3u)) << 4); }

// tcgen05.ld.32x32b.x8: register c <- (lane base + thread, column base + c)
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t* w) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "r"(taddr) : "memory");
}

__device__ __forceinline__ void tmem_st_zero_32x8(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z) : "memory");
}

// NB = batch rows per tile (16, 32, 64, 128); a CTA owns up to NT tiles (NT * accumulator width <= 256 TMEM columns).
template <int NB, int NT>
__global__ void __launch_bounds__(kUThreads, 1)
lstm_fwd_tcu_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmG,
                    const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmC,
                    const __grid_constant__ CUtensorMap tmA, LstmTcuParams p) {
  constexpr int X = NB / 16;                    // 8-column groups per epilogue warp (each warp: NB/2 columns)
  // Narrow tiles: ONE MMA per k-step with N = 2 NB over the stacked [h_hi rows ; h_lo rows] slabs (they are adjacent
  // in the stage), D columns [0, NB) = A h_hi^T and [NB, 2NB) = A h_lo^T, added in the epilogue.  At N = 16 / 32 the
  // issue rate (~23 cycles per MMA), not the tensor pipe, bounds the MMA phase: half the instructions, half the time.
  constexpr bool STK = NB <= 32;
  constexpr int DW = STK ? 2 * NB : NB;         // accumulator columns per tile
  static_assert(NT * DW <= 256 && NT <= 4, "accumulators must fit the 256 TMEM columns beside U");
  constexpr uint32_t kIo = NB * 256;            // P / gates tile: [4 gates][NB rows][16 units] fp32
  constexpr uint32_t kYs = NB * 64;             // y (and c) staging: [NB rows][16 units] fp32
  constexpr uint32_t kSlot = kIo + 2 * kYs;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;                                   // kURing stages of <= 32 KB
  uint8_t* slots = ring + (size_t)kURing * kUStage;       // 2 x { io, ystage, cstage }
  uint64_t* full = reinterpret_cast<uint64_t*>(slots + 2 * kSlot);
  uint64_t* empty = full + kURing;
  uint64_t* tmem_full = empty + kURing;    // [NT <= 4]  tile's MMAs of this step are complete
  uint64_t* p_full = tmem_full + 4;        // [2]  P tile of the slot landed (implies: the slot's staging is free)
  uint64_t* stage_ready = p_full + 2;      // [2]  256 epilogue threads staged the slot's outputs
  uint64_t* ld_done = stage_ready + 2;     // 256 epilogue threads hold the tile's accumulator in registers
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(ld_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ug = blockIdx.x % p.UGn;
  const int sb = (blockIdx.x / p.UGn) % p.NSB;
  const int dir = blockIdx.x / (p.UGn * p.NSB);
  const int j0 = ug * kUUnits;
  const int H = p.H, T = p.T;
  const int ntl = min(NT, p.NTg - NT * sb);             // tiles of this CTA
  const uint32_t stage_bytes = (uint32_t)p.CP * 2u * NB * 128u;
  // When one step's requests of all tiles fit the ring, a stage is only re-used by the SAME tile's next step, whose h
  // exists only after this CTA's epilogue saw the tile's MMAs complete (tmem_full): the stage-release commits (each
  // blocks the MMA warp for ~375 cycles) are implied by the data dependence and are skipped.
  const bool need_empty = ntl * p.NREQ > kURing;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmH)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmG)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    for (int s = 0; s < kURing; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 4; ++s) mbar_init(&tmem_full[s], 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&p_full[s], 1); mbar_init(&stage_ready[s], kUEpi); }
    mbar_init(ld_done, kUEpi);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_s;

  // ---- one-time: U slice -> TMEM.  Lane L = 32q + 16 part + 4 gate + ul holds unit 4q + ul.
  if (warp >= 2 && warp < 6) {
    const int q = warp & 3;
    const int part = lane >> 4, gate = (lane >> 2) & 3, ul = lane & 3;
    const int unit = j0 + 4 * q + ul;
    const __nv_bfloat16* src = (part ? p.ut_lo : p.ut_hi) + ((size_t)dir * 4 * H + (size_t)gate * H + unit) * p.Kp8;
    const bool live = unit < H;
    for (int ks = 0; ks < p.KS; ++ks) {
      uint32_t w[8];
#pragma unroll
      for (int hseg = 0; hseg < 2; ++hseg) {
        const int k0 = ks * 16 + hseg * 8;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (live && k0 < H) v = *reinterpret_cast<const uint4*>(src + k0);   // k0 + 8 <= Kp8 (H is a multiple of 4)
        uint32_t e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int k = k0 + 2 * c;
          if (k >= H) e[c] = 0u;
          else if (k + 1 >= H) e[c] &= 0xffffu;
          w[hseg * 4 + c] = e[c];
        }
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ks * 8);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                   ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (warp == 0) {
    // ---- h stream: for every (step, tile) poll the tile group's counter, then NREQ tensor requests of
    // [64 k][NB rows][2*CP slabs] (one elected lane issues; the loops stay warp-uniform)
    int st = 0;
    uint32_t ph = 0;
    for (int s = 1; s < T; ++s) {
      for (int tau = 0; tau < ntl; ++tau) {
        const int gt = NT * sb + tau;
        const unsigned* ctr = p.counters + (dir * p.NTg + gt) * 32;
        const unsigned target = (unsigned)s * p.UGn;
        unsigned v;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < target);
        asm volatile("fence.proxy.async;" ::: "memory");
        if (lane == 0) TCU_TRACE(0, s * ntl + tau);
        const int row = (dir * 2 + ((s + 1) & 1)) * p.Bpad + gt * NB;
        for (int r = 0; r < p.NREQ; ++r) {
          if (need_empty && (st & 1) == 0) mbar_wait(&empty[st >> 1], ph ^ 1);   // stages are released in pairs (see the MMA warp)
          if (elect_one_u()) {
            if (p.dbg & 4) {
              mbar_arrive_u(&full[st]);      // timing experiment: no h loads (results invalid)
            } else {
              mbar_expect_tx(&full[st], stage_bytes);
              tma_load_3d_u(ring + (size_t)st * kUStage, &tmH, &full[st], 0, row, 2 * p.CP * r);
            }
          }
          __syncwarp();
          if (++st == kURing) { st = 0; ph ^= 1; }
        }
        if (lane == 0) TCU_TRACE(1, s * ntl + tau);
      }
    }
  } else if (warp == 1) {
    // ---- MMA issue: D[tile] (128 x NB, fp32) = sum_k A[:, k] (TMEM) x (h_hi + h_lo)[tile rows, k]^T.
    // A tcgen05.commit blocks the issuing thread for ~375 cycles (scripts/micro/umma_rate.cu): with one commit per
    // 8-MMA stage the pipe ran at ~95 cycles per N=128 MMA instead of 69, so ring stages are released in PAIRS (one
    // commit per two stages).  (Two issuing warps on alternate stages of the same accumulator were tried: faster,
    // but MMAs of different threads into one accumulator are not interlocked -- 6 of 22 parity cases failed.)
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(DW >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int st = 0;
    uint32_t ph = 0;
    for (int s = 1; s < T; ++s) {
      for (int tau = 0; tau < ntl; ++tau) {
        const uint32_t dcol = tmem_base + kDBase + (uint32_t)(tau * DW);
        // A tcgen05.ld issued while MMAs are queued is served behind them, so this tile's MMAs are not issued before
        // the epilogue has the previous tile's accumulator in registers; the rest of that epilogue overlaps with them.
        const int m = (s - 1) * ntl + tau;
        if (m > 0 && !(p.dbg & 1)) mbar_wait(ld_done, (uint32_t)((m - 1) & 1));
        for (int r = 0; r < p.NREQ; ++r) {
          mbar_wait(&full[st], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (lane == 0 && r == 0) TCU_TRACE(2, s * ntl + tau);
          const uint32_t sa = smem_u32(ring + (size_t)st * kUStage);
          if (elect_one_u()) {
            // running descriptor / address increments only: a lone warp executes dependent scalar code at ~4-5 cycles
            // per instruction, and the per-chunk index arithmetic of the first version cost ~140 cycles per chunk
            // (more than the four MMAs of a narrow tile)
            constexpr uint64_t kSlab = (uint64_t)((NB * 128u) >> 4);     // descriptor units per [NB rows x 64 k] slab
            int c = r * p.CP;
            const int cend = (p.dbg & 2) ? c : min(p.Kc, c + p.CP);      // dbg 2: timing experiment without MMAs
            uint64_t dB = make_sw128_desc(sa);                           // hi slab of the stage's first chunk
            uint32_t a = tmem_base + (uint32_t)(c * 32);                 // 4 k-steps x 8 columns per chunk
            uint32_t acc = c > 0 ? 1u : 0u;
            for (; c < cend; ++c) {
              const int nk = p.KS - 4 * c;                               // < 4 only in the last chunk
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k < nk) {
                  umma_ts_bf16(dcol, a + (uint32_t)(8 * k), dB + (uint64_t)(2 * k), idesc, k > 0 ? 1u : acc);
                  if (!STK) umma_ts_bf16(dcol, a + (uint32_t)(8 * k), dB + kSlab + (uint64_t)(2 * k), idesc, 1u);
                }
              }
              acc = 1u;
              a += 32u;
              dB += 2 * kSlab;
            }
            if (need_empty && (st & 1)) umma_commit(&empty[st >> 1]);   // stages st-1 and st are free once these MMAs are done
          }
          __syncwarp();
          if (++st == kURing) { st = 0; ph ^= 1; }
        }
        if (lane == 0) TCU_TRACE(15, s * ntl + tau);
        if (elect_one_u()) umma_commit(&tmem_full[tau]);
        __syncwarp();
        if (lane == 0) TCU_TRACE(3, s * ntl + tau);
      }
    }
  } else if (warp < 10) {
    // ---- epilogue.  Warp w may touch TMEM lanes 32*(w%4)..+31 (units 4q..4q+3); warps w and w+4 split the
    // tile's columns.  Thread l: row part = l>>4, gate gam = (l>>2)&3, unit 4q + (l&3).
    const int q = warp & 3;
    const int hw = (warp - 2) >> 2;
    const int part = lane >> 4, gam = (lane >> 2) & 3, ul = lane & 3;
    const int et = threadIdx.x - 64;                 // 0..255
    // after the butterfly this thread owns the cells (unit 4q + ul, column hw*NB/2 + 8*mm + 4*part + gam), mm < X:
    // the 32 lanes of one access then touch 8 consecutive rows x 4 consecutive words = 32 different banks
    const uint32_t uword = (uint32_t)ul * 4u;
    uint32_t coff[X];     // byte offset of the cell inside a [NB rows][16 units] SWIZZLE_64B tile
#pragma unroll
    for (int mm = 0; mm < X; ++mm) {
      const uint32_t col = (uint32_t)(hw * (NB / 2) + 8 * mm + 4 * part + gam);
      coff[mm] = sw64u(col, (uint32_t)q) + uword;
    }
    // publisher role: thread et < 2*NB converts row et>>1, part et&1 of the staged y tile
    const int prow = et >> 1, ppart = et & 1;
    const size_t R = (size_t)4 * p.Bpad;
    float c_state[NT][X];
#pragma unroll
    for (int a = 0; a < NT; ++a)
#pragma unroll
      for (int mm = 0; mm < X; ++mm) c_state[a][mm] = 0.f;

    for (int s = 0; s < T; ++s) {
#pragma unroll
      for (int tau = 0; tau < NT; ++tau) {
        if (tau >= ntl) break;
        const int n = s * ntl + tau;
        const int slot = n & 1;
        const uint32_t sph = (uint32_t)(n >> 1) & 1u;
        uint8_t* io = slots + (size_t)slot * kSlot;
        uint8_t* ystage = io + kIo;
        uint8_t* cstage = ystage + kYs;
        float pre[X][4];
        if (et == 0) TCU_TRACE(11, n);
        mbar_wait(&p_full[slot], sph);
#pragma unroll
        for (int mm = 0; mm < X; ++mm)
#pragma unroll
          for (int g = 0; g < 4; ++g) pre[mm][g] = *reinterpret_cast<const float*>(io + g * kYs + coff[mm]);
        if (et == 0) TCU_TRACE(9, n);
        if (s > 0) {
          mbar_wait(&tmem_full[tau], (uint32_t)((s - 1) & 1));
          if (et == 0) TCU_TRACE(4, n);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          uint32_t v[8 * X];          // this thread's row, columns hw*NB/2 + [0, 8X)
          const uint32_t taddr = tmem_base + kDBase + (uint32_t)(tau * DW + hw * (NB / 2)) + ((uint32_t)(q * 32) << 16);
#pragma unroll
          for (int j = 0; j < X; ++j) tmem_ld_32x8(taddr + (uint32_t)(8 * j), v + 8 * j);
          if constexpr (STK) {
            uint32_t v2[8 * X];         // the h_lo half of the stacked accumulator
#pragma unroll
            for (int j = 0; j < X; ++j) tmem_ld_32x8(taddr + (uint32_t)(NB + 8 * j), v2 + 8 * j);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 8 * X; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (et == 0) TCU_TRACE(12, n);
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive_u(ld_done);
          if (et == 0) TCU_TRACE(13, n);
          // column 8*mm + k, k = 4 kp + 2 k1 + k0.  Stage 1 (lane ^ 16): keep kp == part, add the partner's row
          // (same gate column, other bf16 part).  Stage 2 (lane ^ 8, gate bit 1): keep k1.  Stage 3 (lane ^ 4): keep k0.
          const bool bp = part != 0, b1 = (gam & 2) != 0, b0 = (gam & 1) != 0;
          float w4[4 * X];              // [mm][k1 k0]
#pragma unroll
          for (int mm = 0; mm < X; ++mm)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float keep = __uint_as_float(bp ? v[8 * mm + 4 + k] : v[8 * mm + k]);
              const float send = __uint_as_float(bp ? v[8 * mm + k] : v[8 * mm + 4 + k]);
              w4[4 * mm + k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
          if (et == 0) { asm volatile("" ::"f"(w4[0]), "f"(w4[4 * X - 1]) : "memory"); TCU_TRACE(14, n); }
          float a0[2 * X], a1[2 * X];   // [mm][k0]: a0 = own gate gam, a1 = gate gam^2
#pragma unroll
          for (int mm = 0; mm < X; ++mm)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const float keep = b1 ? w4[4 * mm + 2 + k] : w4[4 * mm + k];
              const float send = b1 ? w4[4 * mm + k] : w4[4 * mm + 2 + k];
              a0[2 * mm + k] = keep;
              a1[2 * mm + k] = __shfl_xor_sync(0xffffffffu, send, 8);
            }
          float acc[X][4];              // slots: 0 = gate gam, 1 = gam^2, 2 = gam^1, 3 = gam^3
#pragma unroll
          for (int mm = 0; mm < X; ++mm) {
            const float k0 = b0 ? a0[2 * mm + 1] : a0[2 * mm], s0 = b0 ? a0[2 * mm] : a0[2 * mm + 1];
            const float k1 = b0 ? a1[2 * mm + 1] : a1[2 * mm], s1 = b0 ? a1[2 * mm] : a1[2 * mm + 1];
            acc[mm][0] = k0;
            acc[mm][1] = k1;
            acc[mm][2] = __shfl_xor_sync(0xffffffffu, s0, 4);
            acc[mm][3] = __shfl_xor_sync(0xffffffffu, s1, 4);
          }
          // gate g sits in slot (d>>1) | ((d&1)<<1), d = g ^ gam.  selp, not ?: chains: ptxas turned the 4-way
          // compare chain into divergent branches (32 of them: 5200 cycles per tile)
#pragma unroll
          for (int mm = 0; mm < X; ++mm)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const bool x0 = b0 != ((g & 1) != 0), x1 = b1 != ((g & 2) != 0);
              const float lo2 = fsel(x1, acc[mm][1], acc[mm][0]), hi2 = fsel(x1, acc[mm][3], acc[mm][2]);
              pre[mm][g] += fsel(x0, hi2, lo2);
            }
          if (et == 0) TCU_TRACE(5, n);
        }
#pragma unroll
        for (int mm = 0; mm < X; ++mm) {
          const float gi = hsig_u(pre[mm][0]);
          const float gf = hsig_u(pre[mm][1]);
          const float gg = tanh_u(pre[mm][2]);
          const float go = hsig_u(pre[mm][3]);
          const float c = gf * c_state[tau][mm] + gi * gg;
          c_state[tau][mm] = c;
          const float h = go * tanh_u(c);
          *reinterpret_cast<float*>(ystage + coff[mm]) = h;
          if (p.save) {
            *reinterpret_cast<float*>(cstage + coff[mm]) = c;
            *reinterpret_cast<float*>(io + 0 * kYs + coff[mm]) = gi;
            *reinterpret_cast<float*>(io + 1 * kYs + coff[mm]) = gf;
            *reinterpret_cast<float*>(io + 2 * kYs + coff[mm]) = gg;
            *reinterpret_cast<float*>(io + 3 * kYs + coff[mm]) = go;
          }
        }
        if (et == 0) TCU_TRACE(10, n);
        if (s + 1 < T) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          asm volatile("bar.sync 1, 256;" ::: "memory");    // y tile staged
          if (et < 2 * NB) {
            float hv[16];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              const float4 f = *reinterpret_cast<const float4*>(ystage + sw64u((uint32_t)prow, (uint32_t)ch));
              hv[4 * ch] = f.x; hv[4 * ch + 1] = f.y; hv[4 * ch + 2] = f.z; hv[4 * ch + 3] = f.w;
            }
            uint32_t w[8];
#pragma unroll
            for (int u = 0; u < 16; u += 2) {
              const __nv_bfloat162 hh = __floats2bfloat162_rn(hv[u], hv[u + 1]);
              uint32_t wv = *reinterpret_cast<const uint32_t*>(&hh);
              if (ppart) {
                const float f0 = __uint_as_float(wv << 16), f1 = __uint_as_float(wv & 0xffff0000u);
                const __nv_bfloat162 ll = __floats2bfloat162_rn(hv[u] - f0, hv[u + 1] - f1);
                wv = *reinterpret_cast<const uint32_t*>(&ll);
              }
              w[u >> 1] = wv;
            }
            const size_t rowR = (size_t)(dir * 2 + (s & 1)) * p.Bpad + (size_t)(NT * sb + tau) * NB + prow;
            uint8_t* dst = p.hx + (((size_t)(j0 >> 6) * 2 + ppart) * R + rowR) * 128 + (size_t)(j0 & 63) * 2;
            *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(dst + 16) = make_uint4(w[4], w[5], w[6], w[7]);
          }
          if (et == 0) TCU_TRACE(6, n);
          // the release below is cumulative over everything ordered before it by bar.sync (lstm_tc.cu).  (One release
          // per warp after bar.warp.sync instead -- 8 MEMBAR.GPU side by side, no second CTA barrier -- was measured
          // SLOWER: red done at +4855 instead of +4137 cycles after tmem_full.)
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (et == 0) {
            TCU_TRACE(7, n);
            unsigned* ctr = p.counters + (dir * p.NTg + NT * sb + tau) * 32;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
          }
          if (et == 0) TCU_TRACE(8, n);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_u(&stage_ready[slot]);
      }
    }
  } else if (warp == 10) {
    // ---- IO warp: P loads two (step, tile) slots ahead, y / c / gates stores; box = 16 units x NB rows x 1 step
    if (lane == 0) {
      const int total = T * ntl;
      for (int n = 0; n < min(2, total); ++n) {
        const int s = n / ntl, tau = n - s * ntl;
        mbar_expect_tx(&p_full[n & 1], kIo);
        tma_load_4d_u(slots + (size_t)(n & 1) * kSlot, &tmG, &p_full[n & 1], j0, (NT * sb + tau) * NB, dir == 0 ? s : T - 1 - s, dir * 4);
      }
      int s = 0, tau = 0;
      for (int n = 0; n < total; ++n) {
        const int slot = n & 1;
        uint8_t* io = slots + (size_t)slot * kSlot;
        const int t = dir == 0 ? s : T - 1 - s;
        const int b0 = (NT * sb + tau) * NB;
        mbar_wait(&stage_ready[slot], (uint32_t)(n >> 1) & 1u);
        if (p.store_y) tma_store_4d_u(&tmY, io + kIo, j0, b0, t, dir);
        // auxiliary destination (the fusion model's Merge(concat) buffer): layer 1 of a tower stores its h there as well,
        // layer 2 ADDS its h (residual `add`, speech_lstm_ctc_words.py:79) -- no separate add pass over (B,T,2H)
        if (p.aux_mode == 1) tma_store_4d_u(&tmA, io + kIo, j0, b0, t, dir);
        else if (p.aux_mode == 2) tma_reduce_add_4d_u(&tmA, io + kIo, j0, b0, t, dir);
        if (p.save) {
          tma_store_4d_u(&tmC, io + kIo + kYs, j0, b0, t, dir);
          tma_store_4d_u(&tmG, io, j0, b0, t, dir * 4);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        // slot free again: P of (step, tile) sequence number n + 2
        int s2 = s, tau2 = tau;
        for (int a = 0; a < 2; ++a) { if (++tau2 == ntl) { tau2 = 0; ++s2; } }
        if (s2 < T) {
          mbar_expect_tx(&p_full[slot], kIo);
          tma_load_4d_u(io, &tmG, &p_full[slot], j0, (NT * sb + tau2) * NB, dir == 0 ? s2 : T - 1 - s2, dir * 4);
        }
        if (++tau == ntl) { tau = 0; ++s; }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// =================================================================================================
// Backward through time on the same design.  dh_t = dy_t + dP_{t+1} U^T needs ALL 4H gate gradients of a batch row
// (4x the forward exchange), and U^T for 16 units is [16 x 4H]: as the A operand it becomes 128 M-rows by splitting
// K = 4H into its four gate blocks -- TMEM lane (unit u, bf16 part p, gate g) holds U[u, g*H + k], k < H (the same
// 256 columns as forward).  The B operand of gate block g is the exchanged dP_g tile [NB rows x H]; the MMA's
// disable-output-lane mask restricts each MMA to the 32 lanes of its gate block, so ONE accumulator D[128 x NB] ends
// up with the partial sums (u, p, g) and the epilogue's shuffle butterfly ADDS the eight lanes of a unit instead
// of transposing them.  Everything else is the forward kernel: L2 exchange + step counter, TMA ring, two interleaved
// batch tiles, IO warp (loads gates_t / c_{t-1} / dy_t, stores dP_t in place of the gates).
struct LstmTcuBwdParams {
  uint8_t* dx;                 // exchange: [slab = (gate*Kc + chunk)*2 + part][(dir, parity, padded batch row)][64 bf16]
  unsigned* counters;
  const float* U;              // (2, H, 4H)
  int B, T, H, Bpad, UGn, NTg, NSB, Kc, KS, CP, NREQ, NCH;
};

__device__ __forceinline__ void umma_ts_bf16_masked(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                                    uint32_t accum, uint32_t dis) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum), "r"(dis) : "memory");
}
__device__ __forceinline__ float dhsig_u(float s) { return (s > 0.f && s < 1.f) ? 0.2f : 0.f; }

template <int NB>
__global__ void __launch_bounds__(kUThreads, 1)
lstm_bwd_tcu_kernel(const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmG,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmY, LstmTcuBwdParams p) {
  constexpr int X = NB / 16;
  constexpr bool STK = NB <= 32;
  constexpr int DW = STK ? 2 * NB : NB;
  constexpr int RING = NB == 128 ? 3 : 4;       // 128-row tiles: the input staging leaves room for three stages
  constexpr uint32_t kIo = NB * 256;            // gates in / dP out: [4 gates][NB rows][16 units] fp32
  constexpr uint32_t kYs = NB * 64;             // c_{t-1}, dy_t, c_T tiles: [NB rows][16 units] fp32
  constexpr uint32_t kSlot = kIo + 2 * kYs;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* slots = ring + (size_t)RING * kUStage;         // 2 x { gates/dP, c_prev, dy }
  uint8_t* cinit = slots + 2 * kSlot;                     // 2 x c at the first processed time step
  uint64_t* full = reinterpret_cast<uint64_t*>(cinit + 2 * kYs);
  uint64_t* empty = full + 4;
  uint64_t* tmem_full = empty + 4;
  uint64_t* in_full = tmem_full + 2;       // [2]  the slot's gates / c_prev / dy tiles landed
  uint64_t* stage_ready = in_full + 2;     // [2]  256 epilogue threads staged the slot's dP tile
  uint64_t* cinit_full = stage_ready + 2;  // [2]
  uint64_t* ld_done = cinit_full + 2;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(ld_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ug = blockIdx.x % p.UGn;
  const int sb = (blockIdx.x / p.UGn) % p.NSB;
  const int dir = blockIdx.x / (p.UGn * p.NSB);
  const int j0 = ug * kUUnits;
  const int H = p.H, T = p.T;
  const int ntl = min(2, p.NTg - 2 * sb);
  const uint32_t stage_bytes = (uint32_t)p.CP * 2u * NB * 128u;
  const bool need_empty = ntl * p.NREQ > RING;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmD)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmG)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
    for (int s = 0; s < 4; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1); mbar_init(&in_full[s], 1); mbar_init(&stage_ready[s], kUEpi); mbar_init(&cinit_full[s], 1);
    }
    mbar_init(ld_done, kUEpi);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_s;

  // ---- one-time: rows of U -> TMEM.  Lane 32q + 16 part + 4 gate + ul holds U[unit 4q+ul][gate*H + k] (bf16 hi | lo)
  if (warp >= 2 && warp < 6) {
    const int q = warp & 3;
    const int part = lane >> 4, gate = (lane >> 2) & 3, ul = lane & 3;
    const int unit = j0 + 4 * q + ul;
    const float* src = p.U + ((size_t)dir * H + unit) * 4 * H + (size_t)gate * H;
    const bool live = unit < H;
    for (int ks = 0; ks < p.KS; ++ks) {
      uint32_t w[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k0 = ks * 16 + 4 * i;
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live && k0 < H) f = *reinterpret_cast<const float4*>(src + k0);   // H is a multiple of 4
        const float e[4] = {f.x, f.y, f.z, f.w};
        uint32_t pk[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const __nv_bfloat162 hh = __floats2bfloat162_rn(e[2 * c], e[2 * c + 1]);
          uint32_t wv = *reinterpret_cast<const uint32_t*>(&hh);
          if (part) {
            const float f0 = __uint_as_float(wv << 16), f1 = __uint_as_float(wv & 0xffff0000u);
            const __nv_bfloat162 ll = __floats2bfloat162_rn(e[2 * c] - f0, e[2 * c + 1] - f1);
            wv = *reinterpret_cast<const uint32_t*>(&ll);
          }
          pk[c] = wv;
        }
        w[2 * i] = pk[0]; w[2 * i + 1] = pk[1];
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ks * 8);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                   ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  // The accumulators are cleared by the epilogue warps (here and after every read) and EVERY MMA accumulates: with
  // lane masks, a non-accumulating first MMA per gate block would depend on what the hardware does to disabled lanes.
  if (warp >= 2 && warp < 10) {
    const int q = warp & 3, hw = (warp - 2) >> 2;
    for (int tau = 0; tau < 2; ++tau) {
      const uint32_t taddr = tmem_base + kDBase + (uint32_t)(tau * DW + hw * (NB / 2)) + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int j = 0; j < X; ++j) {
        tmem_st_zero_32x8(taddr + (uint32_t)(8 * j));
        if (STK) tmem_st_zero_32x8(taddr + (uint32_t)(NB + 8 * j));
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (warp == 0) {
    // ---- dP stream of the previous backward step: (gate, chunk, part) slabs in slab order
    int st = 0;
    uint32_t ph = 0;
    for (int sp = 1; sp < T; ++sp) {
      for (int tau = 0; tau < ntl; ++tau) {
        const int gt = 2 * sb + tau;
        const unsigned* ctr = p.counters + (dir * p.NTg + gt) * 32;
        const unsigned target = (unsigned)sp * p.UGn;
        unsigned v;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < target);
        asm volatile("fence.proxy.async;" ::: "memory");
        const int row = (dir * 2 + ((sp + 1) & 1)) * p.Bpad + gt * NB;
        for (int r = 0; r < p.NREQ; ++r) {
          if (need_empty) mbar_wait(&empty[st], ph ^ 1);
          if (elect_one_u()) {
            mbar_expect_tx(&full[st], stage_bytes);
            tma_load_3d_u(ring + (size_t)st * kUStage, &tmD, &full[st], 0, row, 2 * p.CP * r);
          }
          __syncwarp();
          if (++st == RING) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issue: D[(u, p, g) x NB] += A[(u, p, g), k] x dP_g[rows, k]^T, lanes of the other gate blocks disabled
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(DW >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint64_t kSlab = (uint64_t)((NB * 128u) >> 4);
    int st = 0;
    uint32_t ph = 0;
    for (int sp = 1; sp < T; ++sp) {
      for (int tau = 0; tau < ntl; ++tau) {
        const uint32_t dcol = tmem_base + kDBase + (uint32_t)(tau * DW);
        const int m = (sp - 1) * ntl + tau;
        if (m > 0) mbar_wait(ld_done, (uint32_t)((m - 1) & 1));
        for (int r = 0; r < p.NREQ; ++r) {
          mbar_wait(&full[st], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(ring + (size_t)st * kUStage);
          if (elect_one_u()) {
            int gc = r * p.CP;
            const int gend = min(p.NCH, gc + p.CP);
            int g = gc / p.Kc, c = gc - g * p.Kc;
            uint64_t dB = make_sw128_desc(sa);
            for (; gc < gend; ++gc) {
              const uint32_t dis = ~((0xFu << (4 * g)) | (0xFu << (16 + 4 * g)));
              const uint32_t a = tmem_base + (uint32_t)(c * 32);
              const int nk = p.KS - 4 * c;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k < nk) {
                  umma_ts_bf16_masked(dcol, a + (uint32_t)(8 * k), dB + (uint64_t)(2 * k), idesc, 1u, dis);
                  if (!STK) umma_ts_bf16_masked(dcol, a + (uint32_t)(8 * k), dB + kSlab + (uint64_t)(2 * k), idesc, 1u, dis);
                }
              }
              dB += 2 * kSlab;
              if (++c == p.Kc) { c = 0; ++g; }
            }
            if (need_empty) umma_commit(&empty[st]);
          }
          __syncwarp();
          if (++st == RING) { st = 0; ph ^= 1; }
        }
        if (elect_one_u()) umma_commit(&tmem_full[tau]);
        __syncwarp();
      }
    }
  } else if (warp < 10) {
    // ---- epilogue: thread l of quadrant q holds lane (part, gate, ul); after the reducing butterfly it owns the cells
    // (unit 4q + ul, column hw*NB/2 + 8*mm + 4*part + gam), mm < X -- the forward kernel's ownership
    const int q = warp & 3;
    const int hw = (warp - 2) >> 2;
    const int part = lane >> 4, gam = (lane >> 2) & 3, ul = lane & 3;
    const int et = threadIdx.x - 64;
    const uint32_t uword = (uint32_t)ul * 4u;
    uint32_t coff[X];
#pragma unroll
    for (int mm = 0; mm < X; ++mm) {
      const uint32_t col = (uint32_t)(hw * (NB / 2) + 8 * mm + 4 * part + gam);
      coff[mm] = sw64u(col, (uint32_t)q) + uword;
    }
    const size_t R = (size_t)4 * p.Bpad;
    float c_cur[2][X], dcn[2][X];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int mm = 0; mm < X; ++mm) { c_cur[a][mm] = 0.f; dcn[a][mm] = 0.f; }

    for (int sp = 0; sp < T; ++sp) {
#pragma unroll
      for (int tau = 0; tau < 2; ++tau) {
        if (tau >= ntl) break;
        const int n = sp * ntl + tau;
        const int slot = n & 1;
        uint8_t* io = slots + (size_t)slot * kSlot;
        const uint8_t* cpt = io + kIo;
        const uint8_t* dyt = cpt + kYs;
        mbar_wait(&in_full[slot], (uint32_t)(n >> 1) & 1u);
        if (sp == 0) {
          mbar_wait(&cinit_full[tau], 0u);
#pragma unroll
          for (int mm = 0; mm < X; ++mm) c_cur[tau][mm] = *reinterpret_cast<const float*>(cinit + tau * kYs + coff[mm]);
        }
        float gt4[X][4], cpv[X], dh[X];
#pragma unroll
        for (int mm = 0; mm < X; ++mm) {
#pragma unroll
          for (int g = 0; g < 4; ++g) gt4[mm][g] = *reinterpret_cast<const float*>(io + g * kYs + coff[mm]);
          cpv[mm] = *reinterpret_cast<const float*>(cpt + coff[mm]);
          dh[mm] = *reinterpret_cast<const float*>(dyt + coff[mm]);
        }
        if (sp > 0) {
          mbar_wait(&tmem_full[tau], (uint32_t)((sp - 1) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          uint32_t v[8 * X];
          const uint32_t taddr = tmem_base + kDBase + (uint32_t)(tau * DW + hw * (NB / 2)) + ((uint32_t)(q * 32) << 16);
#pragma unroll
          for (int j = 0; j < X; ++j) tmem_ld_32x8(taddr + (uint32_t)(8 * j), v + 8 * j);
          if constexpr (STK) {
            uint32_t v2[8 * X];
#pragma unroll
            for (int j = 0; j < X; ++j) tmem_ld_32x8(taddr + (uint32_t)(NB + 8 * j), v2 + 8 * j);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 8 * X; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < X; ++j) {          // clear the accumulator for the tile's next step
            tmem_st_zero_32x8(taddr + (uint32_t)(8 * j));
            if (STK) tmem_st_zero_32x8(taddr + (uint32_t)(NB + 8 * j));
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive_u(ld_done);
          // reducing butterfly over the unit's eight lanes (part, gate): each stage keeps half of the columns and adds
          // the partner's values for them
          const bool bp = part != 0, b1 = (gam & 2) != 0, b0 = (gam & 1) != 0;
          float w4[4 * X];
#pragma unroll
          for (int mm = 0; mm < X; ++mm)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float keep = __uint_as_float(bp ? v[8 * mm + 4 + k] : v[8 * mm + k]);
              const float send = __uint_as_float(bp ? v[8 * mm + k] : v[8 * mm + 4 + k]);
              w4[4 * mm + k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
          float w2[2 * X];
#pragma unroll
          for (int mm = 0; mm < X; ++mm)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const float keep = b1 ? w4[4 * mm + 2 + k] : w4[4 * mm + k];
              const float send = b1 ? w4[4 * mm + k] : w4[4 * mm + 2 + k];
              w2[2 * mm + k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
#pragma unroll
          for (int mm = 0; mm < X; ++mm) {
            const float keep = b0 ? w2[2 * mm + 1] : w2[2 * mm];
            const float send = b0 ? w2[2 * mm] : w2[2 * mm + 1];
            dh[mm] += keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
        }
#pragma unroll
        for (int mm = 0; mm < X; ++mm) {
          const float gi = gt4[mm][0], gf = gt4[mm][1], gg = gt4[mm][2], go = gt4[mm][3];
          const float tc = tanh_u(c_cur[tau][mm]);
          const float dc = dcn[tau][mm] + dh[mm] * go * (1.f - tc * tc);
          const float d_o = dh[mm] * tc * dhsig_u(go);
          const float d_i = dc * gg * dhsig_u(gi);
          const float d_g = dc * gi * (1.f - gg * gg);
          const float d_f = dc * cpv[mm] * dhsig_u(gf);
          dcn[tau][mm] = dc * gf;
          c_cur[tau][mm] = cpv[mm];
          *reinterpret_cast<float*>(io + 0 * kYs + coff[mm]) = d_i;
          *reinterpret_cast<float*>(io + 1 * kYs + coff[mm]) = d_f;
          *reinterpret_cast<float*>(io + 2 * kYs + coff[mm]) = d_g;
          *reinterpret_cast<float*>(io + 3 * kYs + coff[mm]) = d_o;
        }
        if (sp + 1 < T) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          asm volatile("bar.sync 1, 256;" ::: "memory");    // dP tile staged
          // publish dP_t as bf16 hi / lo: one 32-byte sector per (row, gate, part)
          for (int it = et; it < 8 * NB; it += kUEpi) {
            const int ppart = it & 1, pg = (it >> 1) & 3, prow = it >> 3;
            float hv[16];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              const float4 f = *reinterpret_cast<const float4*>(io + pg * kYs + sw64u((uint32_t)prow, (uint32_t)ch));
              hv[4 * ch] = f.x; hv[4 * ch + 1] = f.y; hv[4 * ch + 2] = f.z; hv[4 * ch + 3] = f.w;
            }
            uint32_t w[8];
#pragma unroll
            for (int u = 0; u < 16; u += 2) {
              const __nv_bfloat162 hh = __floats2bfloat162_rn(hv[u], hv[u + 1]);
              uint32_t wv = *reinterpret_cast<const uint32_t*>(&hh);
              if (ppart) {
                const float f0 = __uint_as_float(wv << 16), f1 = __uint_as_float(wv & 0xffff0000u);
                const __nv_bfloat162 ll = __floats2bfloat162_rn(hv[u] - f0, hv[u + 1] - f1);
                wv = *reinterpret_cast<const uint32_t*>(&ll);
              }
              w[u >> 1] = wv;
            }
            const size_t rowR = (size_t)(dir * 2 + (sp & 1)) * p.Bpad + (size_t)(2 * sb + tau) * NB + prow;
            const size_t slab = ((size_t)pg * p.Kc + (size_t)(j0 >> 6)) * 2 + ppart;
            uint8_t* dst = p.dx + (slab * R + rowR) * 128 + (size_t)(j0 & 63) * 2;
            *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(dst + 16) = make_uint4(w[4], w[5], w[6], w[7]);
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (et == 0) {
            unsigned* ctr = p.counters + (dir * p.NTg + 2 * sb + tau) * 32;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_u(&stage_ready[slot]);
      }
    }
  } else if (warp == 10) {
    // ---- IO warp.  Backward step sp differentiates forward step s = T-1-sp at time t (forward-previous time tp; a tp
    // outside [0, T) makes the whole c_prev box out of bounds: TMA fills it with zeros = the initial cell state)
    if (lane == 0) {
      const int total = T * ntl;
      for (int tau = 0; tau < ntl; ++tau) {
        mbar_expect_tx(&cinit_full[tau], kYs);
        tma_load_4d_u(cinit + tau * kYs, &tmC, &cinit_full[tau], j0, (2 * sb + tau) * NB, dir == 0 ? T - 1 : 0, dir);
      }
      auto load_inputs = [&](int slot, int sp, int tau) {
        uint8_t* io = slots + (size_t)slot * kSlot;
        const int t = dir == 0 ? T - 1 - sp : sp;
        const int tp = dir == 0 ? t - 1 : t + 1;
        const int b0 = (2 * sb + tau) * NB;
        mbar_expect_tx(&in_full[slot], kIo + 2 * kYs);
        tma_load_4d_u(io, &tmG, &in_full[slot], j0, b0, t, dir * 4);
        tma_load_4d_u(io + kIo, &tmC, &in_full[slot], j0, b0, tp, dir);
        tma_load_4d_u(io + kIo + kYs, &tmY, &in_full[slot], j0, b0, t, dir);
      };
      for (int n = 0; n < min(2, total); ++n) load_inputs(n & 1, n / ntl, n % ntl);
      int sp = 0, tau = 0;
      for (int n = 0; n < total; ++n) {
        const int slot = n & 1;
        uint8_t* io = slots + (size_t)slot * kSlot;
        const int t = dir == 0 ? T - 1 - sp : sp;
        mbar_wait(&stage_ready[slot], (uint32_t)(n >> 1) & 1u);
        tma_store_4d_u(&tmG, io, j0, (2 * sb + tau) * NB, t, dir * 4);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        int sp2 = sp, tau2 = tau;
        for (int a = 0; a < 2; ++a) { if (++tau2 == ntl) { tau2 = 0; ++sp2; } }
        if (sp2 < T) load_inputs(slot, sp2, tau2);
        if (++tau == ntl) { tau = 0; ++sp; }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int split_bf16_t_launch(const float* x, int R, int K, int ldx, __nv_bfloat16* hi, __nv_bfloat16* lo, int ld_out,
                        cudaStream_t s);

struct TcuLayout {
  int NB, NT, NTg, NSB, Bpad, UGn, Kc, KS, Kp8, CP, NREQ;
  size_t off_hx, off_ut_hi, off_ut_lo, off_trace, total, smem;
};
static TcuLayout tcu_layout(int B, int H) {
  TcuLayout L;
  // Two tiles per CTA.  (FOUR 64-row tiles per CTA for B > 128 -- GR_TCU_NT=4 -- were measured SLOWER: 9.7 vs 8.3 us per
  // step at B=256, H=500.  The single TMA warp and the 4-stage ring serialise the tiles: a 64-row tile needs the whole
  // ring, so the next tile's first load is issued only when stages drain and its latency is exposed at every tile switch
  // -- 4 x (3 670 MMA phase + ~700) = 17 500 cycles per step against 14 400 with two 128-row tiles.)
  int nb = 16;
  while (nb < 128 && 2 * nb < B) nb *= 2;
  int nt = 2;
  if (const char* e = getenv("GR_TCU_NB")) {   // experiments: force the tile width
    const int v = atoi(e);
    if (v == 16 || v == 32 || v == 64 || v == 128) nb = v;
  }
  if (const char* e = getenv("GR_TCU_NT")) {
    if (atoi(e) == 4 && B > 128) { nb = 64; nt = 4; }
  }
  L.NB = nb;
  L.NT = nt;
  L.NTg = (B + nb - 1) / nb;
  L.NSB = (L.NTg + nt - 1) / nt;
  L.Bpad = L.NTg * nb;
  L.UGn = (H + kUUnits - 1) / kUUnits;
  L.Kc = (H + 63) / 64;
  L.KS = (H + 15) / 16;
  L.Kp8 = (H + 7) / 8 * 8;
  L.CP = 128 / nb < L.Kc ? 128 / nb : L.Kc;     // K chunks per TMA request (<= 32 KB)
  L.NREQ = (L.Kc + L.CP - 1) / L.CP;
  L.smem = 1024 + (size_t)kURing * kUStage + (size_t)2 * nb * 384 + 256;
  size_t o = (size_t)2 * L.NTg * 128;
  o = (o + 1023) & ~(size_t)1023;
  L.off_hx = o; o += (size_t)2 * L.Kc * 4 * L.Bpad * 128;
  const size_t ut = (size_t)8 * H * L.Kp8 * 2;
  L.off_ut_hi = o; o += (ut + 255) & ~(size_t)255;
  L.off_ut_lo = o; o += (ut + 255) & ~(size_t)255;
  L.off_trace = o; o += (size_t)160 * 256 * 16 * 8;
  L.total = o + 256;
  return L;
}

size_t lstm_tcu_workspace_bytes(int B, int H) { return tcu_layout(B, H).total; }
int lstm_tcu_grid(int B, int H) { TcuLayout L = tcu_layout(B, H); return 2 * L.NSB * L.UGn; }
size_t lstm_tcu_trace_offset(int B, int H) { return tcu_layout(B, H).off_trace; }

bool lstm_tcu_supported(int B, int H) {
  if (H % 4 != 0 || H < 32 || H > 512) return false;
  TcuLayout L = tcu_layout(B, H);
  return 2 * L.NSB * L.UGn <= num_sms() && 2 * L.NSB * L.UGn <= 160;
}

static int make_io_map_u(CUtensorMap* tm, const float* base, int B, int T, int H, int nvar, int nb, int nbox) {
  const cuuint64_t dims[4] = {(cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)T, (cuuint64_t)nvar};
  const cuuint64_t strides[3] = {(cuuint64_t)T * nvar * H * 4, (cuuint64_t)nvar * H * 4, (cuuint64_t)H * 4};
  const cuuint32_t box[4] = {(cuuint32_t)kUUnits, (cuuint32_t)nb, 1, (cuuint32_t)nbox};
  return make_map_nd(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
}

int lstm_fwd_tcu_launch(float* gates, const float* U, int B, int T, int H, float* y, float* cell, void* workspace,
                        cudaStream_t s, float* aux, int ld_aux, int aux_mode) {
  TcuLayout L = tcu_layout(B, H);
  char* w = static_cast<char*>(workspace);
  LstmTcuParams p;
  p.save = cell != nullptr;
  p.store_y = y != nullptr;
  p.aux_mode = aux ? aux_mode : 0;
  p.hx = reinterpret_cast<uint8_t*>(w + L.off_hx);
  p.counters = reinterpret_cast<unsigned*>(w);
  p.dbg = getenv("GR_TCU_DBG") ? atoi(getenv("GR_TCU_DBG")) : 0;
  p.trace = getenv("GR_TC_TRACE") ? reinterpret_cast<long long*>(w + L.off_trace) : nullptr;
  p.B = B; p.T = T; p.H = H; p.Bpad = L.Bpad; p.UGn = L.UGn; p.NTg = L.NTg; p.NSB = L.NSB; p.Kc = L.Kc; p.KS = L.KS;
  p.Kp8 = L.Kp8; p.CP = L.CP; p.NREQ = L.NREQ;
  GR_CUDA(cudaMemsetAsync(w, 0, L.off_ut_hi, s));   // counters + exchange buffer (its K padding stays zero)
  __nv_bfloat16* ut_hi = reinterpret_cast<__nv_bfloat16*>(w + L.off_ut_hi);
  __nv_bfloat16* ut_lo = reinterpret_cast<__nv_bfloat16*>(w + L.off_ut_lo);
  p.ut_hi = ut_hi; p.ut_lo = ut_lo;
  for (int d = 0; d < 2; ++d) {
    int rc0 = split_bf16_t_launch(U + (size_t)d * H * 4 * H, H, 4 * H, 4 * H, ut_hi + (size_t)d * 4 * H * L.Kp8,
                                  ut_lo + (size_t)d * 4 * H * L.Kp8, L.Kp8, s);
    if (rc0 != GR_OK) return rc0;
  }
  CUtensorMap tH, tG, tY, tC;
  int rc;
  {
    const cuuint64_t R = (cuuint64_t)4 * L.Bpad;
    const cuuint64_t dims[3] = {64, R, (cuuint64_t)2 * L.Kc};
    const cuuint64_t strides[2] = {128, R * 128};
    const cuuint32_t box[3] = {64, (cuuint32_t)L.NB, (cuuint32_t)(2 * L.CP)};
    if ((rc = make_map_nd(&tH, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, p.hx, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) != GR_OK)
      return rc;
  }
  if ((rc = make_io_map_u(&tG, gates, B, T, H, 8, L.NB, 4)) != GR_OK) return rc;
  CUtensorMap tA;
  if (aux) {
    // (unit, batch row, time, direction) view of columns [0, 2H) of a (B*T, ld_aux) matrix starting at `aux`
    const cuuint64_t dims[4] = {(cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)T, 2};
    const cuuint64_t strides[3] = {(cuuint64_t)T * ld_aux * 4, (cuuint64_t)ld_aux * 4, (cuuint64_t)H * 4};
    const cuuint32_t box[4] = {(cuuint32_t)kUUnits, (cuuint32_t)L.NB, 1, 1};
    if ((rc = make_map_nd(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, aux, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) != GR_OK)
      return rc;
  }
  float* ymap = y ? y : (cell ? cell : gates);     // any valid base: the map is not used when y is not stored
  if ((rc = make_io_map_u(&tY, ymap, B, T, H, 2, L.NB, 1)) != GR_OK) return rc;
  if ((rc = make_io_map_u(&tC, cell ? cell : ymap, B, T, H, 2, L.NB, 1)) != GR_OK) return rc;
  if (!aux) tA = tY;
  void* args[] = {&tH, &tG, &tY, &tC, &tA, &p};
  const dim3 grid(2 * L.NSB * L.UGn), block(kUThreads);
  auto go = [&](auto kern) -> int {
    GR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    GR_CUDA(cudaLaunchCooperativeKernel((void*)kern, grid, block, args, L.smem, s));
    return GR_OK;
  };
  switch (L.NB) {
    case 16: return go(lstm_fwd_tcu_kernel<16, 2>);
    case 32: return go(lstm_fwd_tcu_kernel<32, 2>);
    case 64: return L.NT == 4 ? go(lstm_fwd_tcu_kernel<64, 4>) : go(lstm_fwd_tcu_kernel<64, 2>);
    default: return go(lstm_fwd_tcu_kernel<128, 2>);
  }
}


struct TcuBwdLayout {
  int NB, NTg, NSB, Bpad, UGn, Kc, KS, CP, NREQ, NCH;
  size_t off_dx, total, smem;
};
static TcuBwdLayout tcu_bwd_layout(int B, int H) {
  TcuBwdLayout L;
  int nb = 16;
  while (nb < 128 && 2 * nb < B) nb *= 2;
  if (const char* e = getenv("GR_TCU_BWD_NB")) {   // experiments: force the tile width
    const int v = atoi(e);
    if (v == 16 || v == 32 || v == 64 || v == 128) nb = v;
  }
  L.NB = nb;
  L.NTg = (B + nb - 1) / nb;
  L.NSB = (L.NTg + 1) / 2;
  L.Bpad = L.NTg * nb;
  L.UGn = (H + kUUnits - 1) / kUUnits;
  L.Kc = (H + 63) / 64;
  L.KS = (H + 15) / 16;
  L.NCH = 4 * L.Kc;
  L.CP = 128 / nb < L.NCH ? 128 / nb : L.NCH;
  L.NREQ = (L.NCH + L.CP - 1) / L.CP;
  const int ring = nb == 128 ? 3 : 4;
  L.smem = 1024 + (size_t)ring * kUStage + (size_t)2 * nb * 384 + (size_t)2 * nb * 64 + 256;
  size_t o = (size_t)2 * L.NTg * 128;
  o = (o + 1023) & ~(size_t)1023;
  L.off_dx = o; o += (size_t)2 * L.NCH * 4 * L.Bpad * 128;
  L.total = o + 256;
  return L;
}

size_t lstm_tcu_bwd_workspace_bytes(int B, int H) { return tcu_bwd_layout(B, H).total; }

bool lstm_tcu_bwd_supported(int B, int H) {
  if (H % 4 != 0 || H < 32 || H > 512) return false;
  TcuBwdLayout L = tcu_bwd_layout(B, H);
  return 2 * L.NSB * L.UGn <= num_sms();
}

int lstm_bwd_tcu_launch(float* gates, const float* cell, const float* dy, const float* U, int B, int T, int H,
                        void* workspace, cudaStream_t s) {
  TcuBwdLayout L = tcu_bwd_layout(B, H);
  char* w = static_cast<char*>(workspace);
  LstmTcuBwdParams p;
  p.dx = reinterpret_cast<uint8_t*>(w + L.off_dx);
  p.counters = reinterpret_cast<unsigned*>(w);
  p.U = U;
  p.B = B; p.T = T; p.H = H; p.Bpad = L.Bpad; p.UGn = L.UGn; p.NTg = L.NTg; p.NSB = L.NSB; p.Kc = L.Kc; p.KS = L.KS;
  p.CP = L.CP; p.NREQ = L.NREQ; p.NCH = L.NCH;
  GR_CUDA(cudaMemsetAsync(w, 0, L.total - 256, s));   // counters + exchange buffer (K padding and padded rows stay zero)
  CUtensorMap tD, tG, tC, tY;
  int rc;
  {
    const cuuint64_t R = (cuuint64_t)4 * L.Bpad;
    const cuuint64_t dims[3] = {64, R, (cuuint64_t)2 * L.NCH};
    const cuuint64_t strides[2] = {128, R * 128};
    const cuuint32_t box[3] = {64, (cuuint32_t)L.NB, (cuuint32_t)(2 * L.CP)};
    if ((rc = make_map_nd(&tD, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, p.dx, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) != GR_OK)
      return rc;
  }
  if ((rc = make_io_map_u(&tG, gates, B, T, H, 8, L.NB, 4)) != GR_OK) return rc;
  if ((rc = make_io_map_u(&tC, cell, B, T, H, 2, L.NB, 1)) != GR_OK) return rc;
  if ((rc = make_io_map_u(&tY, dy, B, T, H, 2, L.NB, 1)) != GR_OK) return rc;
  void* args[] = {&tD, &tG, &tC, &tY, &p};
  const dim3 grid(2 * L.NSB * L.UGn), block(kUThreads);
  auto go = [&](auto kern) -> int {
    GR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    GR_CUDA(cudaLaunchCooperativeKernel((void*)kern, grid, block, args, L.smem, s));
    return GR_OK;
  };
  switch (L.NB) {
    case 16: return go(lstm_bwd_tcu_kernel<16>);
    case 32: return go(lstm_bwd_tcu_kernel<32>);
    case 64: return go(lstm_bwd_tcu_kernel<64>);
    default: return go(lstm_bwd_tcu_kernel<128>);
  }
}

}  // namespace gr
