// Projection GEMM with a FUSED fp32 prologue, tcgen05 + TMA, sm_100a.
//
//   for v in [0, nvar):   C[:, v*Nv : (v+1)*Nv]  (+)=  op(A, mask_v) * B[v*Nv : (v+1)*Nv, :]^T  + bias
//
// A is the fp32 activation tensor itself (never materialised as bf16 in HBM): the four
// "epilogue" warps spend the main loop as A-PRODUCERS -- they load the fp32 tile with coalesced
// 128-bit loads, apply the per-sequence LSTM input-dropout mask of gate/direction variant v
// (/root/reference/audio_network/speech_lstm_ctc_words.py:61,73 `dropout=`; Keras applies one
// mask per gate to x before each gate's matmul), split every value into bf16 hi + lo and store
// both in the SWIZZLE_128B K-major layout that tcgen05.mma reads.  transA != 0 reads A as
// (K, M) -- the X^T / H_prev^T operands of the BPTT weight-gradient contractions -- with an
// optional time shift inside each sequence (h_{t-1} / h_{t+1} pairing for dU).
// B (weights, or dP^T) arrives pre-split through TMA.  Arithmetic: bf16x3, fp32 accumulate.
//
// PERSISTENT: grid = #SMs, every CTA walks output tiles (M-tile major, so the CTAs running at the
// same time share their fp32 A rows in L2) with one continuous smem ring and TWO TMEM accumulators:
//   warp 0      TMA of the B k-blocks            warp 1      tcgen05.mma issue (one elected lane,
//   warps 2-9   A producers                                  warp-uniform operands)
//   warps 10-13 epilogue of tile i (tcgen05.ld -> +bias -> swizzled smem -> TMA tensor store, or
//               fp32 atomics for split-K / accumulate) while the other warps are on tile i+1.
// The K = 39 / 20 first-layer projections are a pure 4 GB store stream: one CTA per tile spent
// ~50 k cycles per tile on launch, TMEM allocation and per-thread-row stores (32 lines per warp
// store instruction); see profiles/r01_gemm_a32_*.
#include <stdlib.h>
#include <string.h>
#include "tc_common.cuh"

namespace gr {

static constexpr int kA32Threads = 480;  // TMA warp, MMA warp, 8 A-producer warps, 4 epilogue warps, tile scheduler
static constexpr int kSchedDepth = 4;    // tile ids in flight between the scheduler and the 14 consumer warps
static constexpr uint32_t kEpiStage = 16384;   // 128 rows x 32 fp32, SWIZZLE_128B, two buffers

struct A32Params {
  const float* A;
  const float* mask;     // (nvar, nseq, Kdim) or null;  Kdim = K (!transA) or M (transA)
  const float* bias;     // (nvar*Nv) or null
  float* C;
  long long mask_var_stride;
  int lda, ldc, M, Nv, K, BN, nvar, ntile, rows_per_seq, row_shift, transA, aligned4;
  int kb_total, kb_per_split, stages, use_atomic, tmem_cols;
  long long* trace;     // debug (GR_A32_TRACE): clock64 stamps [cta][k-block < 64][8]
  int trace_off;        // first stage recorded (GR_A32_TRACE_OFF)
  int splits, tiles_n, ntiles_total;   // tile t = ((m tile * tiles_n) + (variant group, n tile)) * splits + k split
  int korder, mtiles;   // korder (split-K contractions over B*T): t = (k split * mtiles + m tile) * tiles_n + (variant group, n tile),
                        // so that the CTAs running together share one k range: its rows of A and of dP^T are read from DRAM
                        // once (k split fastest made every m tile re-read all of dP^T: 13 x 0.8 GB for the fusion dW)
  int epi_bufs;      // epilogue staging buffers (2, 4 or 6): TMA stores in flight per CTA
  int bias_vec;      // bias is 16-byte aligned and Nv % 4 == 0: full chunks read it with 128-bit loads
  int epi_stg;       // 1 (GR_A32_EPI=stg): the epilogue warps write the staged chunk themselves (128-byte rows,
                     // coalesced st.global) instead of a TMA tensor store; measured slower (0.83 vs 0.73 ms on the
                     // K = 40 store stream), kept as a cross-check of the TMA path
  unsigned* sched;   // dynamic tile counter (zeroed per launch): CTAs that get an SM late find less work
  float out_scale;   // applied to the accumulator before bias / store (binmask: the dropout scale 1/(1-p))
  int binmask;       // every mask element is 0 or mask_scale (dropout): the producers split each loaded fp32 tile ONCE and
                     // form the nvg variants by AND-ing the packed bf16 words with the keep bits; scale in the epilogue
  int nvg;   // variants per CTA: 4 when the variant is <= 128 columns wide (one loaded A tile, four masked
             // conversions, four 128-column accumulators), else 1 (two 256-column accumulators, double buffered)
};

#define A32_TRACE(slot, it) do { if (p.trace && (unsigned)((it) - p.trace_off) < 64u) p.trace[((size_t)blockIdx.x * 64 + ((it) - p.trace_off)) * 16 + (slot)] = clock64(); } while (0)
__device__ __forceinline__ bool a32_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void a32_tma_store_3d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ void a32_tmem_ld_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void a32_tmem_ld_x64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void split2(float v, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi = __bfloat16_as_ushort(h);
  lo = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
}

// pack two floats into bf16x2 (hi) and the bf16x2 of their residuals (lo); element a in the low half.
// hi is the TRUNCATED upper half (one PRMT for the pair, no conversion instruction), lo = a - hi is
// exact in fp32 and rounded to nearest: |a - (hi + lo)| <= 2^-16 |a|, unbiased.
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
  hi = __byte_perm(ua, ub, 0x7632);
  const float fa = __uint_as_float(ua & 0xffff0000u), fb = __uint_as_float(ub & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - fa, b - fb);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// MODE 0: generic producer (any alignment, any sequence length)
// MODE 1: fast row-major A   (lda, K multiples of 4; a 128-row tile spans <= 2 sequences)
// MODE 2: fast transposed A  (a 64-row k-block spans <= 2 sequences)
template <int MODE>
__global__ void __launch_bounds__(kA32Threads, 1)
gemm_a32_kernel(const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                const __grid_constant__ CUtensorMap tmC, A32Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int BN = p.BN;
  const uint32_t a_bytes = 128 * kBK * 2;
  const uint32_t b_bytes = (uint32_t)BN * kBK * 2;
  const uint32_t stage_bytes = 2 * (a_bytes + b_bytes);     // a multiple of 1024 (BN % 16 == 0)
  uint8_t* epi = smem + (size_t)p.stages * stage_bytes;     // epi_bufs x kEpiStage
  uint64_t* fullA = reinterpret_cast<uint64_t*>(epi + (size_t)p.epi_bufs * kEpiStage);
  uint64_t* fullB = fullA + p.stages;
  uint64_t* empty = fullB + p.stages;
  uint64_t* tmem_full = empty + p.stages;    // [4]: per accumulator buffer (nvg == 1: two) / per variant (nvg == 4)
  uint64_t* tmem_empty = tmem_full + 4;      // [4]
  uint64_t* sfull = tmem_empty + 4;          // [kSchedDepth] tile id published
  uint64_t* sempty = sfull + kSchedDepth;    // [kSchedDepth] tile id read by all 14 consumer warps
  int* tile_ring = reinterpret_cast<int*>(sempty + kSchedDepth);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tile_ring + kSchedDepth);
  // dropout-mask keep bits of the current / next k-block: [2 slots][4 variants x 2 sequences][4 words]
  uint32_t* kbits = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(tmem_ptr_s + 1) + 15) & ~uintptr_t(15));
  // store-stream epilogue with the drain on the producer warps (epi_stg == 3): two 64 KB staging tiles
  uint64_t* staged = reinterpret_cast<uint64_t*>(kbits + 64);   // [2] 128 epilogue threads filled the staging tile
  uint64_t* drained = staged + 2;                                 // [2] the 8 producer warps wrote it to global memory
  float* bias_s = reinterpret_cast<float*>(drained + 2);          // [256] the tile's bias slice (epi_stg == 3 only)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile decode (identical in every role): t -> (m0, var, n0, k-block range)
// every consumer warp walks the same tile sequence, published by the scheduler warp through a small ring
#define A32_NEXT_TILE(n, tile)                                                 \
  int tile;                                                                    \
  {                                                                            \
    const int sl_ = (n) % kSchedDepth;                                         \
    mbar_wait(&sfull[sl_], (uint32_t)(((n) / kSchedDepth) & 1));              \
    tile = tile_ring[sl_];                                                     \
    __syncwarp();                                                              \
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sempty[sl_])) : "memory"); \
  }                                                                            \
  if (tile < 0) break;
#define A32_TILE_DECODE(t)                                                    \
  const int z_ = p.korder ? (t) / (p.tiles_n * p.mtiles) : (t) % p.splits;    \
  const int nv_ = p.korder ? (t) % p.tiles_n : ((t) / p.splits) % p.tiles_n;  \
  const int m0 = (p.korder ? ((t) / p.tiles_n) % p.mtiles : (t) / p.splits / p.tiles_n) * 128; \
  const int var = (nv_ / p.ntile) * p.nvg, n0 = (nv_ % p.ntile) * BN;         \
  const int kb_begin = z_ * p.kb_per_split;                                   \
  const int nkb = min(kb_begin + p.kb_per_split, p.kb_total) - kb_begin;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBl)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
    for (int s = 0; s < p.stages; ++s) { mbar_init(&fullA[s], 8); mbar_init(&fullB[s], 1); mbar_init(&empty[s], 1); }   // fullA: one arrival per producer warp
    for (int b = 0; b < 4; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 128); }
    for (int b = 0; b < kSchedDepth; ++b) { mbar_init(&sfull[b], 1); mbar_init(&sempty[b], 14); }
    for (int b = 0; b < 2; ++b) { mbar_init(&staged[b], 128); mbar_init(&drained[b], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // ---- B operand: two TMA requests (hi, lo) per k-block; the whole warp runs the loop with
    // warp-uniform operands, one elected lane issues (no R2UR waterfall around UTMALDG)
    int it = 0;
    for (int tn = 0;; ++tn) {
      A32_NEXT_TILE(tn, tile)
      A32_TILE_DECODE(tile)
      (void)m0;
      for (int i = 0; i < nkb; ++i)
        for (int vi = 0; vi < p.nvg; ++vi, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          if (lane == 0) A32_TRACE(6, it);
          if (a32_elect_one()) {
            uint8_t* st = smem + (size_t)s * stage_bytes + 2 * a_bytes;
            mbar_expect_tx(&fullB[s], 2 * b_bytes);
            const int k0 = (kb_begin + i) * kBK;
            tma_load_2d(st, &tmBh, &fullB[s], k0, (var + vi) * p.Nv + n0);
            tma_load_2d(st + b_bytes, &tmBl, &fullB[s], k0, (var + vi) * p.Nv + n0);
          }
          __syncwarp();
        }
    }
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int it = 0, lt = 0;
    for (;; ++lt) {
      A32_NEXT_TILE(lt, tile)
      A32_TILE_DECODE(tile)
      (void)m0; (void)var; (void)n0;
      const int nbuf = p.nvg == 1 ? 2 : 1;
      const int buf = lt % nbuf;
      // nvg == 1: two 256-column accumulators alternate between tiles.  nvg == 4: the four variants' accumulators are
      // handed over ONE BY ONE (tmem_full / tmem_empty per variant): variant 0 of the next tile starts as soon as the
      // epilogue has drained accumulator 0, while it is still draining 1..3 (all four at once stalled the MMA warp
      // for the whole epilogue, 39 k of 210 k cycles per tile of the fusion projection)
      if (p.nvg == 1) {
        mbar_wait(&tmem_empty[buf], (uint32_t)(((lt / nbuf) & 1) ^ 1));   // epilogue drained the accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      for (int i = 0; i < nkb; ++i)
        for (int vi = 0; vi < p.nvg; ++vi, ++it) {
          const uint32_t acc = tmem_base + (uint32_t)(p.nvg == 1 ? buf * 256 : vi * 128);
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          if (p.nvg > 1 && i == 0) {
            if (lane == 0 && vi == 0) A32_TRACE(10, lt);
            mbar_wait(&tmem_empty[vi], (uint32_t)((lt & 1) ^ 1));
            if (lane == 0 && vi == 0) A32_TRACE(11, lt);
          }
          mbar_wait(&fullB[s], ph);
          if (lane == 0) A32_TRACE(3, it);
          mbar_wait(&fullA[s], ph);
          if (lane == 0) A32_TRACE(4, it);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
          const uint64_t dAh = make_sw128_desc(sa);
          const uint64_t dAl = make_sw128_desc(sa + a_bytes);
          const uint64_t dBh = make_sw128_desc(sa + 2 * a_bytes);
          const uint64_t dBl = make_sw128_desc(sa + 2 * a_bytes + b_bytes);
          if (a32_elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              const uint64_t adv = (uint64_t)(k * 2);
              umma_bf16(acc, dAh + adv, dBh + adv, idesc, (i > 0 || k > 0) ? 1u : 0u);
              umma_bf16(acc, dAh + adv, dBl + adv, idesc, 1u);
              umma_bf16(acc, dAl + adv, dBh + adv, idesc, 1u);
            }
            umma_commit(&empty[s]);
            if (i == nkb - 1) {
              if (p.nvg > 1) umma_commit(&tmem_full[vi]);
              else umma_commit(&tmem_full[buf]);
            }
          }
          __syncwarp();
          if (lane == 0) A32_TRACE(5, it);
        }
    }
  } else if (warp < 10) {
    // =============== A producers: 8 warps, 256 threads ===============
    const int t = threadIdx.x - 64;  // 0..255
    int it = 0;
    int gblk = 0;   // k-blocks handled so far (all tiles): slot parity of the keep-bit ring
    (void)gblk; (void)kbits;
    // epi_stg == 3 (store-stream tiles): the producers have next to nothing to do (one k-block of K <= 64), so THEY write
    // the staged output tiles to global memory -- warp w rows 16w .. 16w+15 of every 128-column chunk, 512 contiguous
    // bytes per row -- one tile behind their own A tile, while the four epilogue warps only move TMEM -> staging.
    int dcn = 0;                      // chunks drained so far (same sequence as the epilogue warps' counter)
    int pm0 = 0, pvar = 0, pn0 = 0;   // the tile whose chunks are drained next
    bool have_prev = false;
    auto drain_tile = [&]() {
      const int w = t >> 5;
      const int cbase = pvar * p.Nv + pn0;
      for (int c0 = 0; c0 < BN && pn0 + c0 < p.Nv; c0 += 128, ++dcn) {
        const int db = dcn & 1;
        mbar_wait(&staged[db], (uint32_t)((dcn >> 1) & 1));
        const uint8_t* stg = epi + (size_t)db * 65536;
        const bool cok = pn0 + c0 + 4 * lane < p.Nv && c0 + 4 * lane < BN;
        const uint32_t q4l = (uint32_t)lane >> 3, jl = (uint32_t)lane & 7u;
        const int nrow = min(16, p.M - pm0 - w * 16);
        float* cptr = p.C + (size_t)(pm0 + w * 16) * p.ldc + cbase + c0 + 4 * lane;
        const uint8_t* sbase = stg + (uint32_t)(w * 16) * 512u + q4l * 128u;
        if (cok) {
#pragma unroll 1
          for (int i0 = 0; i0 < 16; i0 += 8) {
            if (i0 >= nrow) break;
            float4 o[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) o[u] = *reinterpret_cast<const float4*>(sbase + (uint32_t)(i0 + u) * 512u + ((jl ^ (uint32_t)u) << 4));
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              if (i0 + u < nrow) *reinterpret_cast<float4*>(cptr) = o[u];
              cptr += p.ldc;
            }
          }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&drained[db])) : "memory");
      }
    };
    for (int tn = 0;; ++tn) {
    A32_NEXT_TILE(tn, tile)
    A32_TILE_DECODE(tile)
    (void)n0;
    const float* mv = p.mask ? p.mask + (long long)var * p.mask_var_stride : nullptr;
    if constexpr (MODE == 1) {
      // ---------------- fast row-major producer ----------------
      const int kq = t & 15, rbase = t >> 4;             // float4 column, first row; rows rbase + 16 j
      const int T_ = p.rows_per_seq;
      const int seqA = m0 / T_;
      const int rbound = (seqA + 1) * T_ - m0;           // tile rows >= rbound belong to sequence seqA+1
      const int nseq = (p.M + T_ - 1) / T_;
      const float* ap[8];
      bool rok[8], selB[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = rbase + 16 * j;
        rok[j] = m0 + r < p.M;
        selB[j] = r >= rbound;
        ap[j] = p.A + (size_t)(rok[j] ? m0 + r : 0) * p.lda + kq * 4;
      }
      const float* mpA = mv ? mv + (size_t)seqA * p.K + kq * 4 : nullptr;
      const float* mpB = mv ? mv + (size_t)min(seqA + 1, nseq - 1) * p.K + kq * 4 : nullptr;
      const uint32_t off0 = (uint32_t)rbase * 128u + ((((uint32_t)kq >> 1) ^ ((uint32_t)rbase & 7u)) << 4) + ((uint32_t)kq & 1u) * 8u;
      const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f), zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 xa[8], mAa = one4, mBa = one4;
      // loads of k-block i: rows j0..j0+3 of this thread (and, with the second half, the masks)
      auto ld = [&](int i, int j0, float4* x, float4& mA, float4& mB) {
        const int koff = (kb_begin + i) * kBK;
        const bool kv = koff + kq * 4 < p.K;
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j0 + j] = (kv && rok[j0 + j]) ? __ldg(reinterpret_cast<const float4*>(ap[j0 + j] + koff)) : zero4;
        if (mv && j0 == 4) {
          mA = kv ? __ldg(reinterpret_cast<const float4*>(mpA + koff)) : zero4;
          mB = kv ? __ldg(reinterpret_cast<const float4*>(mpB + koff)) : zero4;
        }
      };
      if (p.nvg > 1 && p.binmask) {
        // ---- four variants per CTA, masks in {0, s}: the fp32 tile of k-block i is loaded and split ONCE; variant vi
        //      stores the packed hi / lo words AND-ed with its keep bits (s is applied by the epilogue).  The fp32
        //      registers are free right after the split, so the loads of k-block i+1 have four stages to land.
        //      Keep bits: no per-stage global mask loads (their latency was exposed every stage: 2 200 cycles per
        //      stage with 600 of work) -- producer warp w owns row (variant w/2, sequence w%2) of the mask, loads the 64
        //      floats of the NEXT k-block one block ahead, ballots them into two words of a 2-slot shared ring; one
        //      named barrier per k-block, then every thread picks its 8 nibbles.
        auto ldx = [&](int i) {
          const int koff = (kb_begin + i) * kBK;
          const bool kv = koff + kq * 4 < p.K;
#pragma unroll
          for (int j = 0; j < 8; ++j) xa[j] = (kv && rok[j]) ? __ldg(reinterpret_cast<const float4*>(ap[j] + koff)) : zero4;
        };
        const int bw = t >> 5, bl = t & 31;
        const float* mrow = mv + (long long)(bw >> 1) * p.mask_var_stride + (size_t)((bw & 1) ? min(seqA + 1, nseq - 1) : seqA) * p.K;
        float f0 = 0.f, f1 = 0.f;
        auto bld = [&](int i) {
          const int k = (kb_begin + i) * kBK + bl;
          f0 = (i < nkb && k < p.K) ? __ldg(mrow + k) : 0.f;
          f1 = (i < nkb && k + 32 < p.K) ? __ldg(mrow + k + 32) : 0.f;
        };
        if (nkb > 0) { ldx(0); bld(0); }
        for (int i = 0; i < nkb; ++i, ++gblk) {
          // publish the keep bits of block i (loaded one block ago), prefetch those of block i+1
          {
            const uint32_t b0 = __ballot_sync(0xffffffffu, f0 != 0.f), b1 = __ballot_sync(0xffffffffu, f1 != 0.f);
            if (bl == 0) *reinterpret_cast<uint2*>(&kbits[((gblk & 1) * 8 + bw) * 4]) = make_uint2(b0, b1);
          }
          bld(i + 1);
          if (t == 0) A32_TRACE(7, it);
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (t == 0) A32_TRACE(8, it);
          uint32_t nib = 0;     // nibble (vi*2 + sel): keep bits of this thread's four k
#pragma unroll
          for (int r = 0; r < 8; ++r)
            nib |= ((kbits[((gblk & 1) * 8 + r) * 4 + (kq >> 3)] >> ((kq & 7) * 4)) & 0xfu) << (4 * r);
          uint2 hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            split_pair(xa[j].x, xa[j].y, hi[j].x, lo[j].x);
            split_pair(xa[j].z, xa[j].w, hi[j].y, lo[j].y);
          }
          if (t == 0 && p.trace) { asm volatile("" ::"r"(hi[7].y) : "memory"); A32_TRACE(9, it); }   // stamp after the split
          if (i + 1 < nkb) ldx(i + 1);
          if (t == 0) A32_TRACE(12, it);
#pragma unroll
          for (int vi = 0; vi < 4; ++vi, ++it) {
            const int s = it % p.stages;
            const uint32_t ph = (it / p.stages) & 1;
            uint8_t* Ah = smem + (size_t)s * stage_bytes + off0;
            uint8_t* Al = Ah + a_bytes;
            const uint32_t nA = nib >> (8 * vi), nB = nib >> (8 * vi + 4);
            const uint2 zA = make_uint2(((nA & 1u) ? 0x0000ffffu : 0u) | ((nA & 2u) ? 0xffff0000u : 0u),
                                        ((nA & 4u) ? 0x0000ffffu : 0u) | ((nA & 8u) ? 0xffff0000u : 0u));
            const uint2 zB = make_uint2(((nB & 1u) ? 0x0000ffffu : 0u) | ((nB & 2u) ? 0xffff0000u : 0u),
                                        ((nB & 4u) ? 0x0000ffffu : 0u) | ((nB & 8u) ? 0xffff0000u : 0u));
            if (t == 0) A32_TRACE(0, it);
            mbar_wait(&empty[s], ph ^ 1);
            if (t == 0) A32_TRACE(1, it);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint2 z = selB[j] ? zB : zA;
              *reinterpret_cast<uint2*>(Ah + j * 2048) = make_uint2(hi[j].x & z.x, hi[j].y & z.y);
              *reinterpret_cast<uint2*>(Al + j * 2048) = make_uint2(lo[j].x & z.x, lo[j].y & z.y);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&fullA[s])) : "memory");
            if (t == 0) A32_TRACE(2, it);
          }
        }
      } else if (p.nvg > 1) {
        // ---- several variants per CTA: the fp32 tile of k-block i is loaded ONCE and converted nvg
        //      times, each time under the dropout mask of one variant (masks prefetched one ahead)
        auto ldmask = [&](int q, float4& mA, float4& mB) {
          const int i = q / p.nvg, vi = q - i * p.nvg;
          const int koff = (kb_begin + i) * kBK;
          const bool kv = koff + kq * 4 < p.K;
          const long long vo = (long long)vi * p.mask_var_stride;
          mA = kv ? __ldg(reinterpret_cast<const float4*>(mpA + vo + koff)) : zero4;
          mB = kv ? __ldg(reinterpret_cast<const float4*>(mpB + vo + koff)) : zero4;
        };
        auto ldx = [&](int i) {
          const int koff = (kb_begin + i) * kBK;
          const bool kv = koff + kq * 4 < p.K;
#pragma unroll
          for (int j = 0; j < 8; ++j) xa[j] = (kv && rok[j]) ? __ldg(reinterpret_cast<const float4*>(ap[j] + koff)) : zero4;
        };
        const int Q = nkb * p.nvg;
        float4 mAn = one4, mBn = one4;
        if (nkb > 0) { ldx(0); ldmask(0, mAa, mBa); }
        int vi = 0, i = 0;
        for (int q = 0; q < Q; ++q, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          uint8_t* Ah = smem + (size_t)s * stage_bytes + off0;
          uint8_t* Al = Ah + a_bytes;
          if (q + 1 < Q) ldmask(q + 1, mAn, mBn);
          if (t == 0) A32_TRACE(0, it);
          mbar_wait(&empty[s], ph ^ 1);
          if (t == 0) A32_TRACE(1, it);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 x = xa[j];
            const float4 mm = selB[j] ? mBa : mAa;
            x.x *= mm.x; x.y *= mm.y; x.z *= mm.z; x.w *= mm.w;
            uint32_t h01, l01, h23, l23;
            split_pair(x.x, x.y, h01, l01);
            split_pair(x.z, x.w, h23, l23);
            *reinterpret_cast<uint2*>(Ah + j * 2048) = make_uint2(h01, h23);
            *reinterpret_cast<uint2*>(Al + j * 2048) = make_uint2(l01, l23);
          }
          if (++vi == p.nvg) { vi = 0; ++i; if (i < nkb) ldx(i); }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&fullA[s])) : "memory");
          if (t == 0) A32_TRACE(2, it);
          mAa = mAn; mBa = mBn;
        }
      } else {
      if (nkb > 0) { ld(0, 0, xa, mAa, mBa); ld(0, 4, xa, mAa, mBa); }
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        uint8_t* Ah = smem + (size_t)s * stage_bytes + off0;
        uint8_t* Al = Ah + a_bytes;
        if (t == 0) A32_TRACE(0, it);
        mbar_wait(&empty[s], ph ^ 1);
        if (t == 0) A32_TRACE(1, it);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
          for (int j = 4 * hf; j < 4 * hf + 4; ++j) {
            float4 x = xa[j];
            const float4 mm = selB[j] ? mBa : mAa;
            x.x *= mm.x; x.y *= mm.y; x.z *= mm.z; x.w *= mm.w;
            uint32_t h01, l01, h23, l23;
            split_pair(x.x, x.y, h01, l01);
            split_pair(x.z, x.w, h23, l23);
            *reinterpret_cast<uint2*>(Ah + j * 2048) = make_uint2(h01, h23);
            *reinterpret_cast<uint2*>(Al + j * 2048) = make_uint2(l01, l23);
          }
          if (t == 0) A32_TRACE(8 + 2 * hf, it);
          // the registers of this half are free: its loads of block i+1 overlap the other half
          if (i + 1 < nkb) ld(i + 1, 4 * hf, xa, mAa, mBa);
          if (t == 0) A32_TRACE(9 + 2 * hf, it);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (t == 0) A32_TRACE(12, it);
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&fullA[s])) : "memory");
        if (t == 0) A32_TRACE(2, it);
      }
      }
    } else if constexpr (MODE == 2) {
      // ---------------- fast transposed producer ----------------
      // warp w <-> tile rows (m) 16w..16w+15; lane = (mq, kc): k rows 8kc..8kc+7 of the k-block (one
      // 16-byte chunk of each output row) x the float4 of m = 16w + 4mq..+3.  Loads are 128-bit
      // (8 per thread per k-block instead of 32 scalar ones, which capped the SM's requests in
      // flight); stores are STS.128.
      const int w = t >> 5, ln = t & 31;
      // lane = mq*8 + kc: the 8 lanes of a quarter-warp (one STS.128 wavefront) share a tile row and cover its 8 swizzled
      // 16-byte chunks -- conflict-free (kc*4 + mq put two rows 8 apart on the same banks: 2-way conflicts, ncu:
      // 106 M wavefronts for 53 M ideal in the fusion dW); the global loads touch the same addresses as before
      const int kc = ln & 7, mq = ln >> 3;
      const int T_ = p.rows_per_seq;
      const int nseq = (p.K + T_ - 1) / T_;
      const int rloc = 16 * w + 4 * mq;              // first tile row of this thread
      const bool mok = m0 + rloc < p.M;              // M % 4 == 0: the float4 is all-in or all-out
      const float* ap = p.A + (mok ? m0 + rloc : 0);
      const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f), one4 = make_float4(1.f, 1.f, 1.f, 1.f);
      float4 xa[8], ma[2];
      int selA_mask = 0;   // bit kk set: row kk belongs to the second sequence of the block
      auto ld = [&](int i, float4* x, float4* mk, int& sel) {
        const int kbase = (kb_begin + i) * kBK + 8 * kc;
        const int seq0 = ((kb_begin + i) * kBK) / T_;
        sel = 0;
        const int sqb = kbase / T_;
        const int ttb = kbase - sqb * T_;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const int kg = kbase + kk;
          const bool wrap = ttb + kk >= T_;            // T_ >= 64 > 8: at most one wrap inside the chunk
          const int sq = sqb + (wrap ? 1 : 0);
          const int tt = ttb + kk - (wrap ? T_ : 0) + p.row_shift;
          const bool ok = kg < p.K && tt >= 0 && tt < T_;
          if (sq != seq0) sel |= 1 << kk;
          const size_t roff = (size_t)(ok ? kg + p.row_shift : 0) * p.lda;
          x[kk] = (ok && mok) ? __ldg(reinterpret_cast<const float4*>(ap + roff)) : zero4;
        }
        if (mv) {
#pragma unroll
          for (int sI = 0; sI < 2; ++sI)
            mk[sI] = mok ? __ldg(reinterpret_cast<const float4*>(mv + (size_t)min(seq0 + sI, nseq - 1) * p.M + m0 + rloc)) : zero4;
        }
      };
      ma[0] = ma[1] = one4;
      if (p.nvg > 1 && p.binmask) {
        // ---- four variants per CTA, masks in {0, s} (see MODE 1): one load and ONE split of the k-block; a variant's
        //      row g is the packed words or zeros (one AND per word; the per-k select only in the rare chunk that
        //      straddles two sequences).  Keep bits through the shared ring: warp w owns (variant w/2, sequence
        //      seq0 + w%2) and ballots the mask of the tile's 128 rows into four words.
        const int bw = t >> 5, bl = t & 31;
        float f[4] = {0.f, 0.f, 0.f, 0.f};
        auto bld = [&](int i) {
          const int seq0 = ((kb_begin + min(i, nkb - 1)) * kBK) / T_;
          const float* mrow = mv + (long long)(bw >> 1) * p.mask_var_stride + (size_t)min(seq0 + (bw & 1), nseq - 1) * p.M + m0;
#pragma unroll
          for (int u = 0; u < 4; ++u) f[u] = (i < nkb && m0 + 32 * u + bl < p.M) ? __ldg(mrow + 32 * u + bl) : 0.f;
        };
        float4 dummy[2];
        if (nkb > 0) { ld(0, xa, dummy, selA_mask); bld(0); }
        for (int i = 0; i < nkb; ++i, ++gblk) {
          {
            uint32_t b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) b[u] = __ballot_sync(0xffffffffu, f[u] != 0.f);
            if (bl == 0) *reinterpret_cast<uint4*>(&kbits[((gblk & 1) * 8 + bw) * 4]) = make_uint4(b[0], b[1], b[2], b[3]);
          }
          bld(i + 1);
          asm volatile("bar.sync 1, 256;" ::: "memory");
          uint32_t nib = 0;     // nibble (vi*2 + sel): keep bits of this thread's four tile rows
#pragma unroll
          for (int r = 0; r < 8; ++r)
            nib |= ((kbits[((gblk & 1) * 8 + r) * 4 + (rloc >> 5)] >> (rloc & 31)) & 0xfu) << (4 * r);
          uint32_t hh[4][4], ll[4][4];
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int kk = 0; kk < 8; kk += 2) {
              const float4 v0 = xa[kk], v1 = xa[kk + 1];
              const float e0 = g == 0 ? v0.x : g == 1 ? v0.y : g == 2 ? v0.z : v0.w;
              const float e1 = g == 0 ? v1.x : g == 1 ? v1.y : g == 2 ? v1.z : v1.w;
              split_pair(e0, e1, hh[g][kk >> 1], ll[g][kk >> 1]);
            }
          const int sel = selA_mask;
          if (i + 1 < nkb) ld(i + 1, xa, dummy, selA_mask);
#pragma unroll
          for (int vi = 0; vi < 4; ++vi, ++it) {
            const int s = it % p.stages;
            const uint32_t ph = (it / p.stages) & 1;
            uint8_t* Ah = smem + (size_t)s * stage_bytes;
            uint8_t* Al = Ah + a_bytes;
            const uint32_t kb = nib >> (8 * vi);
            if (t == 0) A32_TRACE(0, it);
            mbar_wait(&empty[s], ph ^ 1);
            if (t == 0) A32_TRACE(1, it);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const bool k0 = (kb >> g) & 1u, k1 = (kb >> (4 + g)) & 1u;
              uint32_t z[4];
              if (sel == 0) {
                z[0] = z[1] = z[2] = z[3] = k0 ? 0xffffffffu : 0u;
              } else {
#pragma unroll
                for (int w = 0; w < 4; ++w)
                  z[w] = ((((sel >> (2 * w)) & 1) ? k1 : k0) ? 0x0000ffffu : 0u) |
                         ((((sel >> (2 * w + 1)) & 1) ? k1 : k0) ? 0xffff0000u : 0u);
              }
              const uint32_t r = (uint32_t)(rloc + g);
              const uint32_t off = r * 128u + (((uint32_t)kc ^ (r & 7u)) << 4);
              *reinterpret_cast<uint4*>(Ah + off) = make_uint4(hh[g][0] & z[0], hh[g][1] & z[1], hh[g][2] & z[2], hh[g][3] & z[3]);
              *reinterpret_cast<uint4*>(Al + off) = make_uint4(ll[g][0] & z[0], ll[g][1] & z[1], ll[g][2] & z[2], ll[g][3] & z[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&fullA[s])) : "memory");
            if (t == 0) A32_TRACE(2, it);
          }
        }
      } else if (p.nvg > 1) {
        // ---- several variants per CTA (see MODE 1): one load of the k-block, nvg masked conversions
        auto ldmask = [&](int q, float4* mk) {
          const int i = q / p.nvg, vi = q - i * p.nvg;
          const int seq0 = ((kb_begin + i) * kBK) / T_;
          const float* mvv = mv + (long long)vi * p.mask_var_stride;
#pragma unroll
          for (int sI = 0; sI < 2; ++sI)
            mk[sI] = mok ? __ldg(reinterpret_cast<const float4*>(mvv + (size_t)min(seq0 + sI, nseq - 1) * p.M + m0 + rloc)) : zero4;
        };
        float4 mn[2] = {one4, one4};
        float4 dummy[2];
        const int Q = nkb * p.nvg;
        if (nkb > 0) { ld(0, xa, dummy, selA_mask); ldmask(0, ma); }
        int vi = 0, i = 0;
        for (int q = 0; q < Q; ++q, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          uint8_t* Ah = smem + (size_t)s * stage_bytes;
          uint8_t* Al = Ah + a_bytes;
          if (q + 1 < Q) ldmask(q + 1, mn);
          if (t == 0) A32_TRACE(0, it);
          mbar_wait(&empty[s], ph ^ 1);
          if (t == 0) A32_TRACE(1, it);
          const float4 m_0 = ma[0], m_1 = ma[1];
          const int sel = selA_mask;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float mg0 = g == 0 ? m_0.x : g == 1 ? m_0.y : g == 2 ? m_0.z : m_0.w;
            const float mg1 = g == 0 ? m_1.x : g == 1 ? m_1.y : g == 2 ? m_1.z : m_1.w;
            uint32_t hh[4], ll[4];
#pragma unroll
            for (int kk = 0; kk < 8; kk += 2) {
              const float4 v0 = xa[kk], v1 = xa[kk + 1];
              const float e0 = g == 0 ? v0.x : g == 1 ? v0.y : g == 2 ? v0.z : v0.w;
              const float e1 = g == 0 ? v1.x : g == 1 ? v1.y : g == 2 ? v1.z : v1.w;
              const float a0 = e0 * (((sel >> kk) & 1) ? mg1 : mg0);
              const float a1 = e1 * (((sel >> (kk + 1)) & 1) ? mg1 : mg0);
              split_pair(a0, a1, hh[kk >> 1], ll[kk >> 1]);
            }
            const uint32_t r = (uint32_t)(rloc + g);
            const uint32_t off = r * 128u + (((uint32_t)kc ^ (r & 7u)) << 4);
            *reinterpret_cast<uint4*>(Ah + off) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
            *reinterpret_cast<uint4*>(Al + off) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
          }
          if (++vi == p.nvg) { vi = 0; ++i; if (i < nkb) ld(i, xa, dummy, selA_mask); }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&fullA[s])) : "memory");
          if (t == 0) A32_TRACE(2, it);
          ma[0] = mn[0]; ma[1] = mn[1];
        }
      } else {
      if (nkb > 0) ld(0, xa, ma, selA_mask);
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        uint8_t* Ah = smem + (size_t)s * stage_bytes;
        uint8_t* Al = Ah + a_bytes;
        if (t == 0) A32_TRACE(0, it);
        mbar_wait(&empty[s], ph ^ 1);
        if (t == 0) A32_TRACE(1, it);
        const float4 m_0 = ma[0], m_1 = ma[1];
        const int sel = selA_mask;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float mg0 = g == 0 ? m_0.x : g == 1 ? m_0.y : g == 2 ? m_0.z : m_0.w;
          const float mg1 = g == 0 ? m_1.x : g == 1 ? m_1.y : g == 2 ? m_1.z : m_1.w;
          uint32_t hh[4], ll[4];
#pragma unroll
          for (int kk = 0; kk < 8; kk += 2) {
            const float4 v0 = xa[kk], v1 = xa[kk + 1];
            const float e0 = g == 0 ? v0.x : g == 1 ? v0.y : g == 2 ? v0.z : v0.w;
            const float e1 = g == 0 ? v1.x : g == 1 ? v1.y : g == 2 ? v1.z : v1.w;
            const float a0 = e0 * (((sel >> kk) & 1) ? mg1 : mg0);
            const float a1 = e1 * (((sel >> (kk + 1)) & 1) ? mg1 : mg0);
            split_pair(a0, a1, hh[kk >> 1], ll[kk >> 1]);
          }
          const uint32_t r = (uint32_t)(rloc + g);
          const uint32_t off = r * 128u + (((uint32_t)kc ^ (r & 7u)) << 4);
          *reinterpret_cast<uint4*>(Ah + off) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
          *reinterpret_cast<uint4*>(Al + off) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
        }
        if (i + 1 < nkb) ld(i + 1, xa, ma, selA_mask);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&fullA[s])) : "memory");
        if (t == 0) A32_TRACE(2, it);
      }
      }
    } else {
    constexpr int NV = 8;            // float4 per thread per k-block (128 x 64 floats / 256 threads / 4)
    // per-thread static coordinates
    int rr[NV], cc[NV];              // !transA: (row in tile, float4 index in row)   transA: (k row, float4 index over m)
    int seq[NV], tin[NV];            // sequence id and position-in-sequence of the row that carries the mask
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int f = j * 256 + t;
      if (!p.transA) { rr[j] = f >> 4; cc[j] = f & 15; const int m = m0 + rr[j]; seq[j] = m / p.rows_per_seq; tin[j] = 0; }
      else { rr[j] = f >> 5; cc[j] = f & 31; const int kg = kb_begin * kBK + rr[j]; seq[j] = kg / p.rows_per_seq; tin[j] = kg - seq[j] * p.rows_per_seq; }
    }
    auto load_tile = [&](int i, float4* v) {
      const int k0 = (kb_begin + i) * kBK;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!p.transA) {
          const int m = m0 + rr[j], k = k0 + cc[j] * 4;
          if (m < p.M && k < p.K) {
            const float* src = p.A + (size_t)m * p.lda + k;
            if (p.aligned4 && k + 3 < p.K) {
              x = __ldg(reinterpret_cast<const float4*>(src));
              if (mv) {
                const float4 mm = __ldg(reinterpret_cast<const float4*>(mv + (size_t)seq[j] * p.K + k));
                x.x *= mm.x; x.y *= mm.y; x.z *= mm.z; x.w *= mm.w;
              }
            } else {
              const float* ms = mv ? mv + (size_t)seq[j] * p.K + k : nullptr;
              x.x = src[0] * (ms ? ms[0] : 1.f);
              if (k + 1 < p.K) x.y = src[1] * (ms ? ms[1] : 1.f);
              if (k + 2 < p.K) x.z = src[2] * (ms ? ms[2] : 1.f);
              if (k + 3 < p.K) x.w = src[3] * (ms ? ms[3] : 1.f);
            }
          }
        } else {
          const int kg = k0 + rr[j], m = m0 + cc[j] * 4;
          const int tt = tin[j] + p.row_shift;
          if (kg < p.K && m < p.M && tt >= 0 && tt < p.rows_per_seq) {
            const float* src = p.A + (size_t)(kg + p.row_shift) * p.lda + m;
            if (p.aligned4 && m + 3 < p.M) {
              x = __ldg(reinterpret_cast<const float4*>(src));
              if (mv) {
                const float4 mm = __ldg(reinterpret_cast<const float4*>(mv + (size_t)seq[j] * p.M + m));
                x.x *= mm.x; x.y *= mm.y; x.z *= mm.z; x.w *= mm.w;
              }
            } else {
              const float* ms = mv ? mv + (size_t)seq[j] * p.M + m : nullptr;
              x.x = src[0] * (ms ? ms[0] : 1.f);
              if (m + 1 < p.M) x.y = src[1] * (ms ? ms[1] : 1.f);
              if (m + 2 < p.M) x.z = src[2] * (ms ? ms[2] : 1.f);
              if (m + 3 < p.M) x.w = src[3] * (ms ? ms[3] : 1.f);
            }
          }
          // advance this row's (sequence, position) to the next k-block
          tin[j] += kBK;
          while (tin[j] >= p.rows_per_seq) { tin[j] -= p.rows_per_seq; ++seq[j]; }
        }
        v[j] = x;
      }
    };
    float4 va[NV], vb[NV];
    if (nkb > 0) load_tile(0, va);
    for (int i = 0; i < nkb; ++i, ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      uint8_t* Ah = smem + (size_t)s * stage_bytes;
      uint8_t* Al = Ah + a_bytes;
      float4* cur = (i & 1) ? vb : va;
      float4* nxt = (i & 1) ? va : vb;
      if (i + 1 < nkb) load_tile(i + 1, nxt);   // in flight while k-block i is converted
      mbar_wait(&empty[s], ph ^ 1);
      if (!p.transA) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const uint32_t r = (uint32_t)rr[j], kq = (uint32_t)cc[j];
          uint32_t h0, l0, h1, l1, h2, l2, h3, l3;
          split2(cur[j].x, h0, l0); split2(cur[j].y, h1, l1); split2(cur[j].z, h2, l2); split2(cur[j].w, h3, l3);
          const uint32_t off = r * 128u + (((kq >> 1) ^ (r & 7u)) << 4) + (kq & 1u) * 8u;
          *reinterpret_cast<uint2*>(Ah + off) = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
          *reinterpret_cast<uint2*>(Al + off) = make_uint2(l0 | (l1 << 16), l2 | (l3 << 16));
        }
      } else {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const uint32_t kk = (uint32_t)rr[j], mq = (uint32_t)cc[j];
          const float xs[4] = {cur[j].x, cur[j].y, cur[j].z, cur[j].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t r = mq * 4 + e;
            uint32_t h, l;
            split2(xs[e], h, l);
            const uint32_t off = r * 128u + (((kk >> 3) ^ (r & 7u)) << 4) + (kk & 7u) * 2u;
            *reinterpret_cast<uint16_t*>(Ah + off) = (uint16_t)h;
            *reinterpret_cast<uint16_t*>(Al + off) = (uint16_t)l;
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&fullA[s])) : "memory");
    }
    }
    if constexpr (MODE == 1) {
      if (p.epi_stg == 3) {           // (end of the tile's producer work) drain the PREVIOUS tile, remember this one
        if (have_prev) drain_tile();
        pm0 = m0; pvar = var; pn0 = n0; have_prev = true;
      }
    }
    }  // tile loop (producers)
    if constexpr (MODE == 1) {
      if (p.epi_stg == 3 && have_prev) drain_tile();
    }
  } else if (warp < 14) {
    // =============== epilogue: 4 warps; warp q owns TMEM lanes (= tile rows) 32q..32q+31 ===============
    const int q = warp & 3;           // warps 10..13 -> 2, 3, 0, 1: any bijection onto the lane groups works
    const int et = (int)threadIdx.x - 320;   // 0..127
    const uint32_t r = (uint32_t)(q * 32 + lane);   // row inside the tile
    int lt = 0, cc = 0;
    for (;; ++lt) {
      A32_NEXT_TILE(lt, tile)
      A32_TILE_DECODE(tile)
      (void)nkb;
      const int nbuf = p.nvg == 1 ? 2 : 1;
      const int buf = lt % nbuf;
      if (et == 0) A32_TRACE(13, lt);
      if (p.nvg == 1) {
        mbar_wait(&tmem_full[buf], (uint32_t)((lt / nbuf) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      if (et == 0) A32_TRACE(14, lt);
      const int row = m0 + (int)r;
      const bool add_bias = p.bias != nullptr && z_ == 0;
      if (p.epi_stg == 3) {
        // ---- store-stream epilogue, drain on the producer warps: these four warps move TMEM -> (+bias) -> one of two
        // 64 KB swizzled staging tiles.  The tile's bias slice sits in shared memory (8 uniform global loads per 32
        // columns had their L2 latency in every iteration) and the tcgen05.ld of the next 32 columns is in flight while
        // the current ones are stored.
        const int cbase = var * p.Nv + n0;
        if (add_bias) {
          asm volatile("bar.sync 2, 128;" ::: "memory");          // the previous tile's chunks have read bias_s
          for (int c = et; c < BN; c += 128) bias_s[c] = (n0 + c < p.Nv) ? __ldg(p.bias + cbase + c) : 0.f;
          asm volatile("bar.sync 2, 128;" ::: "memory");
        }
        for (int c0 = 0; c0 < BN && n0 + c0 < p.Nv; c0 += 128, ++cc) {
          uint8_t* const stgw = epi + (size_t)(cc & 1) * 65536;
          const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + c0);
          // 64 columns per tcgen05.ld (a load + wait pair costs several hundred cycles whatever its width: x16 steps with
          // the next load in flight were SLOWER than x32 steps, 11.5 k against 9.7 k cycles per tile)
          const int ncols = min(128, min(BN - c0, p.Nv - n0 - c0));
          mbar_wait(&drained[cc & 1], (uint32_t)(((cc >> 1) & 1) ^ 1));
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            if (64 * h >= ncols) break;
            uint32_t v[64];
            a32_tmem_ld_x64(tbase + (uint32_t)(64 * h), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                     __uint_as_float(v[4 * j + 3]));
              if (add_bias) {
                const float4 bq = *reinterpret_cast<const float4*>(bias_s + c0 + 64 * h + 4 * j);
                o.x += bq.x; o.y += bq.y; o.z += bq.z; o.w += bq.w;
              }
              // 32-column group 2h + j/8, 16-byte chunk j % 8 of its 128-byte row piece
              *reinterpret_cast<float4*>(stgw + r * 512u + (uint32_t)(2 * h + (j >> 3)) * 128u +
                                         ((((uint32_t)(j & 7)) ^ (r & 7u)) << 4)) = o;
            }
          }
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&staged[cc & 1])) : "memory");
        }
      } else if (p.epi_stg >= 2) {
        // ---- wide coalesced epilogue (store-stream tiles, one k-block; GR_A32_EPI=tma restores the tensor stores): 128
        // columns at a time through a 64 KB swizzled staging tile, then every warp instruction writes 512 contiguous
        // bytes of ONE output row (the 32-column TMA tensor stores write 128-byte row pieces).
        const int cbase = var * p.Nv + n0;
        for (int c0 = 0; c0 < BN && n0 + c0 < p.Nv; c0 += 128) {
          uint8_t* const stgw = epi;
          asm volatile("bar.sync 2, 128;" ::: "memory");          // the previous drain has read the staging tile
#pragma unroll 1
          for (int q4 = 0; q4 < 4; ++q4) {
            if (c0 + 32 * q4 >= BN || n0 + c0 + 32 * q4 >= p.Nv) break;
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + c0 + 32 * q4);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int ncol = min(32, p.Nv - (n0 + c0 + 32 * q4));
            const float* brow = p.bias ? p.bias + cbase + c0 + 32 * q4 : nullptr;
            const bool bvec = add_bias && p.bias_vec && ncol == 32;
            float4 bb[8];
            if (bvec) {
#pragma unroll
              for (int j = 0; j < 8; ++j) bb[j] = __ldg(reinterpret_cast<const float4*>(brow) + j);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                     __uint_as_float(v[4 * j + 3]));
              if (bvec) {
                o.x += bb[j].x; o.y += bb[j].y; o.z += bb[j].z; o.w += bb[j].w;
              } else if (add_bias) {
                if (4 * j < ncol) o.x += brow[4 * j];
                if (4 * j + 1 < ncol) o.y += brow[4 * j + 1];
                if (4 * j + 2 < ncol) o.z += brow[4 * j + 2];
                if (4 * j + 3 < ncol) o.w += brow[4 * j + 3];
              }
              *reinterpret_cast<float4*>(stgw + r * 512u + (uint32_t)q4 * 128u + (((uint32_t)j ^ (r & 7u)) << 4)) = o;
            }
          }
          asm volatile("bar.sync 2, 128;" ::: "memory");
          // drain: warp q writes rows 32q .. 32q+31, lane l the 16 bytes of columns 4l .. 4l+3.  Eight rows per
          // iteration: eight independent LDS.128 (the swizzle term repeats every 8 rows), then eight STG.128 off one
          // running pointer -- a lone warp per sub-partition is latency-bound, so the loop is kept to ~3 instructions a row
          const bool cok = n0 + c0 + 4 * lane < p.Nv && c0 + 4 * lane < BN;
          const uint32_t q4l = (uint32_t)lane >> 3, jl = (uint32_t)lane & 7u;
          const int nrow = min(32, p.M - m0 - q * 32);                // rows of this warp inside the matrix
          float* cptr = p.C + (size_t)(m0 + q * 32) * p.ldc + cbase + c0 + 4 * lane;
          const uint8_t* sbase = epi + (uint32_t)(q * 32) * 512u + q4l * 128u;
          if (cok) {
#pragma unroll 1
            for (int i0 = 0; i0 < 32; i0 += 8) {
              if (i0 >= nrow) break;
              float4 o[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) o[u] = *reinterpret_cast<const float4*>(sbase + (uint32_t)(i0 + u) * 512u + ((jl ^ (uint32_t)u) << 4));
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                if (i0 + u < nrow) *reinterpret_cast<float4*>(cptr) = o[u];
                cptr += p.ldc;
              }
            }
          }
        }
      } else
      for (int vi = 0; vi < p.nvg; ++vi) {
      if (p.nvg > 1) {
        mbar_wait(&tmem_full[vi], (uint32_t)(lt & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      for (int c0 = 0; c0 < BN && n0 + c0 < p.Nv; c0 += 32) {
        const int cbase = (var + vi) * p.Nv + n0;     // first output column of this tile / variant
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((p.nvg == 1 ? buf * 256 : vi * 128) + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int ncol = min(32, p.Nv - (n0 + c0));
        const float* brow = p.bias ? p.bias + cbase + c0 : nullptr;
        // bias of a full 16-byte aligned chunk: eight uniform 128-bit loads (the per-element guarded loads of round 1
        // cost ~1 ms of the 2.3 ms first-layer projection)
        const bool bvec = add_bias && p.bias_vec && ncol == 32;
        float4 bb[8];
        if (bvec) {
#pragma unroll
          for (int j = 0; j < 8; ++j) bb[j] = __ldg(reinterpret_cast<const float4*>(brow) + j);
        }
        if (p.binmask) {
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(__uint_as_float(v[c]) * p.out_scale);
        }
        if (p.use_atomic) {
          if (row < p.M) {
            float* crow = p.C + (size_t)row * p.ldc + cbase + c0;
            for (int c = 0; c < ncol; ++c) {
              float x = __uint_as_float(v[c]);
              if (add_bias) x += brow[c];
              atomicAdd(crow + c, x);
            }
          }
        } else {
          // registers -> swizzled staging (conflict-free STS.128) -> one TMA tensor store per 32
          // columns; rows >= M and columns >= Nv are clipped by the tensor map
          uint8_t* stg = epi + (size_t)(cc % p.epi_bufs) * kEpiStage;
          if (et == 0 && !p.epi_stg) {   // the store that used this buffer epi_bufs chunks ago has read it
            if (p.epi_bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else if (p.epi_bufs == 4) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 5;" ::: "memory");
          }
          asm volatile("bar.sync 2, 128;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                   __uint_as_float(v[4 * j + 3]));
            if (bvec) {
              o.x += bb[j].x; o.y += bb[j].y; o.z += bb[j].z; o.w += bb[j].w;
            } else if (add_bias) {
              if (4 * j < ncol) o.x += brow[4 * j];
              if (4 * j + 1 < ncol) o.y += brow[4 * j + 1];
              if (4 * j + 2 < ncol) o.z += brow[4 * j + 2];
              if (4 * j + 3 < ncol) o.w += brow[4 * j + 3];
            }
            *reinterpret_cast<float4*>(stg + r * 128u + (((uint32_t)j ^ (r & 7u)) << 4)) = o;
          }
          if (p.epi_stg) {
            // 8 lanes per 128-byte row, 4 rows per store instruction; the next use of this buffer is
            // epi_bufs chunks away and ordered by the bar.sync at the top of that chunk
            asm volatile("bar.sync 2, 128;" ::: "memory");
            const uint32_t ch = (uint32_t)et & 7u;                    // 16-byte chunk = columns 4 ch .. 4 ch + 3
            const bool cok = n0 + c0 + 4 * (int)ch < p.Nv;
            float* cdst = p.C + (size_t)cbase + c0 + 4 * ch;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t rr = (uint32_t)(et >> 3) + 16u * i;      // tile row
              const float4 o = *reinterpret_cast<const float4*>(stg + rr * 128u + ((ch ^ (rr & 7u)) << 4));
              if (cok && m0 + (int)rr < p.M) *reinterpret_cast<float4*>(cdst + (size_t)(m0 + rr) * p.ldc) = o;
            }
          } else {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("bar.sync 2, 128;" ::: "memory");
          if (et == 0) {
            a32_tma_store_3d(&tmC, stg, n0 + c0, m0, var + vi);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          }
          ++cc;
        }
      }
      if (p.nvg > 1) {   // accumulator vi is free for the next tile
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[vi])) : "memory");
      }
      }
      if (et == 0) A32_TRACE(15, lt);
      if (p.nvg == 1) {
        // this accumulator may be overwritten by the MMA warp (tile lt + 2)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[buf])) : "memory");
      }
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    // =============== tile scheduler: first tile = blockIdx.x, then an atomic counter ===============
    if (lane == 0) {
      for (int n = 0;; ++n) {
        const int sl = n % kSchedDepth;
        mbar_wait(&sempty[sl], (uint32_t)(((n / kSchedDepth) & 1) ^ 1));
        int tile = n == 0 ? (int)blockIdx.x : (int)(gridDim.x + atomicAdd(p.sched, 1u));
        if (tile >= p.ntiles_total) tile = -1;
        tile_ring[sl] = tile;
        asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sfull[sl])) : "memory");
        if (tile < 0) break;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

}  // namespace gr

static long long* g_a32_trace = nullptr;
static unsigned* g_a32_sched = nullptr;     // ring of per-launch tile counters
static unsigned g_a32_sched_next = 0;
static constexpr unsigned kSchedSlots = 256;
extern "C" int gr_debug_a32_trace(long long* host_out, size_t n) {
  if (!g_a32_trace) return -1;
  return cudaMemcpy(host_out, g_a32_trace, n * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}

static int gemm_a32_launch(const float* A, int lda, int transA, int row_shift, const float* mask, float mask_scale,
                           int rows_per_seq, int nvar, const void* b_hi, const void* b_lo, int ldb,
                           const float* bias, float* C, int ldc, int M, int Nv, int K, int accumulate,
                           void* stream);

extern "C" int gr_gemm_a32_f32(const float* A, int lda, int transA, int row_shift, const float* mask,
                               int rows_per_seq, int nvar, const void* b_hi, const void* b_lo, int ldb,
                               const float* bias, float* C, int ldc, int M, int Nv, int K, int accumulate,
                               void* stream) {
  return gemm_a32_launch(A, lda, transA, row_shift, mask, 0.f, rows_per_seq, nvar, b_hi, b_lo, ldb, bias, C, ldc, M, Nv, K,
                         accumulate, stream);
}

extern "C" int gr_gemm_a32_dropout_f32(const float* A, int lda, int transA, int row_shift, const float* mask,
                                       float mask_scale, int rows_per_seq, int nvar, const void* b_hi, const void* b_lo,
                                       int ldb, const float* bias, float* C, int ldc, int M, int Nv, int K,
                                       int accumulate, void* stream) {
  if (mask && !(mask_scale > 0.f)) return gr::set_error(GR_EINVAL, "gemm_a32_dropout: mask_scale must be > 0");
  return gemm_a32_launch(A, lda, transA, row_shift, mask, mask ? mask_scale : 0.f, rows_per_seq, nvar, b_hi, b_lo, ldb, bias,
                         C, ldc, M, Nv, K, accumulate, stream);
}

static int gemm_a32_launch(const float* A, int lda, int transA, int row_shift, const float* mask, float mask_scale,
                           int rows_per_seq, int nvar, const void* b_hi, const void* b_lo, int ldb,
                           const float* bias, float* C, int ldc, int M, int Nv, int K, int accumulate,
                           void* stream) {
  using namespace gr;
  if (!A || !b_hi || !b_lo || !C) return set_error(GR_EINVAL, "gemm_a32: null pointer");
  if (M <= 0 || Nv <= 0 || K <= 0 || nvar <= 0 || rows_per_seq <= 0 || ldc < nvar * Nv)
    return set_error(GR_EINVAL, "gemm_a32: bad shape");
  if ((ldb % 8) || ldb < K) return set_error(GR_EINVAL, "gemm_a32: ldb must be a multiple of 8 and >= K");
  if (!transA && lda < K) return set_error(GR_EINVAL, "gemm_a32: lda < K");
  if (transA && lda < M) return set_error(GR_EINVAL, "gemm_a32: lda < M (transA)");
  if (!transA && row_shift != 0) return set_error(GR_EUNSUPPORTED, "gemm_a32: row_shift needs transA");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  A32Params p;
  p.A = A; p.mask = mask; p.bias = bias; p.C = C; p.lda = lda; p.ldc = ldc; p.M = M; p.Nv = Nv; p.K = K;
  p.nvar = nvar; p.rows_per_seq = rows_per_seq; p.row_shift = row_shift; p.transA = transA ? 1 : 0;
  p.bias_vec = (bias && (Nv % 4) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0) ? 1 : 0;
  const int seqs = ((transA ? K : M) + rows_per_seq - 1) / rows_per_seq;
  p.mask_var_stride = (long long)seqs * (transA ? M : K);
  const int inner = transA ? M : K;
  p.aligned4 = ((lda % 4) == 0 && (inner % 4) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
                (!mask || (reinterpret_cast<uintptr_t>(mask) & 15) == 0)) ? 1 : 0;
  // tile width: as few equal tiles as possible, each <= 256 and a multiple of 32 (the epilogue moves
  // 32-column chunks)
  const int nt = (Nv + 255) / 256;
  p.BN = (((Nv + nt - 1) / nt) + 31) / 32 * 32;
  p.ntile = (Nv + p.BN - 1) / p.BN;
  p.kb_total = (K + kBK - 1) / kBK;
  const int mt = (M + 127) / 128;
  int mode = 0;
  const char* force = getenv("GR_A32_MODE");
  if (!transA && p.aligned4 && (!mask || rows_per_seq >= 128)) mode = 1;
  if (transA && p.aligned4 && rows_per_seq >= 64) mode = 2;
  if (force && force[0] == '0') mode = 0;
  const char* nvg_env = getenv("GR_A32_NVG");
  // measured (scripts/trace_a32.py): sharing the loaded tile pays for the transposed operand (dW: 1.57 ->
  // 1.32 ms per 64K rows), not for the row-major one (2 086 vs 1 935 cycles per stage: its half-tile
  // prefetch already hides the loads, and both are near the shared-memory bandwidth bound); GR_A32_NVG=4
  // forces it there too, GR_A32_NVG=1 disables it
  const bool nvg_ok = mode != 0 && mask && (nvar % 4) == 0 && p.ntile == 1 && p.BN <= 128;
  // masks known to be {0, mask_scale} (gr_gemm_a32_dropout_f32): one split per loaded tile serves all four variants, which
  // makes the shared tile pay for the row-major operand too; GR_A32_BINMASK=0 falls back to the multiply path
  const char* bm_env = getenv("GR_A32_BINMASK");
  const bool bin_ok = nvg_ok && mask_scale > 0.f && !(bm_env && bm_env[0] == '0');
  p.nvg = nvg_ok && (((mode == 2 || bin_ok) && !(nvg_env && nvg_env[0] == '1')) || (nvg_env && nvg_env[0] == '4')) ? 4 : 1;
  p.binmask = (bin_ok && p.nvg == 4) ? 1 : 0;
  p.out_scale = p.binmask ? mask_scale : 1.f;
  const long long tiles = (long long)mt * p.ntile * (nvar / p.nvg);
  int splits = 1;
  const int sms = num_sms();
  if (tiles < sms && p.kb_total >= 8) {
    // split K so that the work units (tiles x splits, handed out dynamically) fill whole waves of the SMs: the old
    // rule (about two units per SM) gave the fusion dW contraction 26 x 12 = 312 units = 2.1 waves, i.e. a third wave
    // for 16 of them.  Score = wave efficiency x the share of a unit that is main loop (about 8 k-blocks of fill,
    // drain and atomic epilogue per unit).
    const int smax = (int)min((long long)(4 * sms + tiles - 1) / tiles, (long long)p.kb_total / 4);
    double best = -1.0;
    for (int sp = 1; sp <= smax; ++sp) {
      const long long units = tiles * sp;
      const long long waves = (units + sms - 1) / sms;
      const int kbps = (p.kb_total + sp - 1) / sp;
      const double score = (double)units / (double)(waves * sms) * ((double)kbps / (double)(kbps + 8));
      if (score > best + 1e-9) { best = score; splits = sp; }
    }
    if (const char* se = getenv("GR_A32_SPLITS")) { const int v = atoi(se); if (v >= 1 && v <= p.kb_total) splits = v; }
  }
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  // the TMA-store epilogue needs 16-byte aligned rows / variant blocks; otherwise fp32 atomics
  const bool tma_ok = (ldc % 4) == 0 && (Nv % 4) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
  p.use_atomic = (splits > 1 || accumulate || !tma_ok) ? 1 : 0;
  if (p.use_atomic && !accumulate) GR_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)nvar * Nv * 4, M, s));
  if (!g_a32_sched) GR_CUDA(cudaMalloc(&g_a32_sched, kSchedSlots * sizeof(unsigned)));
  p.sched = g_a32_sched + (__atomic_fetch_add(&g_a32_sched_next, 1u, __ATOMIC_RELAXED) % kSchedSlots);
  GR_CUDA(cudaMemsetAsync(p.sched, 0, sizeof(unsigned), s));
  p.trace = nullptr;
  p.trace_off = 0;
  if (const char* to = getenv("GR_A32_TRACE_OFF")) p.trace_off = atoi(to);
  if (getenv("GR_A32_TRACE")) {
    if (!g_a32_trace) GR_CUDA(cudaMalloc(&g_a32_trace, (size_t)160 * 64 * 16 * 8));
    GR_CUDA(cudaMemsetAsync(g_a32_trace, 0, (size_t)160 * 64 * 16 * 8, s));
    p.trace = g_a32_trace;
  }
  p.splits = splits;
  p.mtiles = mt;
  { const char* ko = getenv("GR_A32_KORDER"); p.korder = (splits > 1 && !(ko && ko[0] == '0')) ? 1 : 0; }
  p.tiles_n = p.ntile * (nvar / p.nvg);
  p.ntiles_total = (int)(tiles * splits);
  const size_t stage_bytes = 2 * ((size_t)128 * kBK * 2 + (size_t)p.BN * kBK * 2);
  // single-k-block problems (the K = 39 / 20 first-layer projections) are a pure store stream: one
  // operand stage is enough, the rest of shared memory keeps more TMA stores in flight
  p.epi_bufs = 2;
  { const char* es = getenv("GR_A32_EPI"); p.epi_stg = (es && es[0] == 's') ? 1 : 0; }
  int stages = (int)((227 * 1024 - 1024 - 2 * kEpiStage - 768) / stage_bytes);
  if (stages > 4) stages = 4;
  if (p.kb_total == 1 && !p.use_atomic) {
    const char* eb = getenv("GR_A32_EPI_BUFS");
    stages = 1;
    p.epi_bufs = eb ? atoi(eb) : 2;   // measured: 2, 4, 6 buffers within noise (the TMA store engine is the limit)
    if (p.epi_bufs != 2 && p.epi_bufs != 4 && p.epi_bufs != 6) p.epi_bufs = 2;
    // store-stream tiles: the wide coalesced epilogue (64 KB staging tile = four 16 KB buffers) unless GR_A32_EPI=tma / stg.
    // Measured (profiles/r02_store_probe.txt): speech first layer 1.52 ms against 1.73 with the TMA tensor stores, skeletal
    // 1.10 against 1.13 -- still 2.7 TB/s of the 7.4 TB/s a plain fill reaches: 512-byte pieces of 128 rows, 16 KB apart
    { const char* es = getenv("GR_A32_EPI");
      if (p.nvg == 1 && !(es && (es[0] == 't' || es[0] == 's'))) { p.epi_stg = 2; p.epi_bufs = 4; }
      // two staging tiles drained by the producer warps (row-major fast producer only); GR_A32_EPI=wide1: one tile,
      // drained by the epilogue warps themselves
      if (p.epi_stg == 2 && mode == 1 && !(es && strcmp(es, "wide1") == 0) &&
          1024 + stage_bytes + (size_t)8 * kEpiStage + 2048 <= 227 * 1024) { p.epi_stg = 3; p.epi_bufs = 8; } }
    while (p.epi_stg != 3 && p.epi_bufs > 2 && 1024 + stage_bytes + (size_t)p.epi_bufs * kEpiStage + 512 > 227 * 1024) p.epi_bufs -= 2;
    if (p.epi_stg == 2 && p.epi_bufs < 4) p.epi_stg = 0;   // no room for the 64 KB staging tile: TMA stores
    if (1024 + stage_bytes + (size_t)p.epi_bufs * kEpiStage + 512 > 227 * 1024) return set_error(GR_EUNSUPPORTED, "gemm_a32: tile does not fit shared memory");
  } else
  if (stages < 2) return set_error(GR_EUNSUPPORTED, "gemm_a32: tile does not fit shared memory");
  p.stages = stages;
  p.tmem_cols = 512;   // two accumulators of <= 256 columns
  const size_t smem = 1024 + stages * stage_bytes + (size_t)p.epi_bufs * kEpiStage + (3 * stages + 8 + 2 * kSchedDepth) * 8 + kSchedDepth * 4 + 16 + 16 + 256 + 32 + (p.epi_stg == 3 ? 1024 : 0);
  CUtensorMap tBh, tBl, tC;
  int rc;
  if ((rc = make_map(&tBh, b_hi, (uint64_t)nvar * Nv, ldb, ldb, p.BN)) != GR_OK) return rc;
  if ((rc = make_map(&tBl, b_lo, (uint64_t)nvar * Nv, ldb, ldb, p.BN)) != GR_OK) return rc;
  if (tma_ok) {
    // (column within variant, row, variant): columns >= Nv and rows >= M are clipped on store
    const cuuint64_t dims[3] = {(cuuint64_t)Nv, (cuuint64_t)M, (cuuint64_t)nvar};
    const cuuint64_t strides[2] = {(cuuint64_t)ldc * 4, (cuuint64_t)Nv * 4};
    const cuuint32_t box[3] = {32, 128, 1};
    if ((rc = make_map_nd(&tC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, C, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) != GR_OK)
      return rc;
  } else {
    tC = tBh;   // unused by the atomic epilogue
  }
  int grid_cap = sms;
  if (const char* gc = getenv("GR_A32_GRID")) grid_cap = max(1, min(sms, atoi(gc)));   // scaling experiments
  dim3 grid((unsigned)min((long long)grid_cap, tiles * splits));
  if (mode == 1) {
    GR_CUDA(cudaFuncSetAttribute(gemm_a32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_a32_kernel<1><<<grid, kA32Threads, smem, s>>>(tBh, tBl, tC, p);
  } else if (mode == 2) {
    GR_CUDA(cudaFuncSetAttribute(gemm_a32_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_a32_kernel<2><<<grid, kA32Threads, smem, s>>>(tBh, tBl, tC, p);
  } else {
    GR_CUDA(cudaFuncSetAttribute(gemm_a32_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_a32_kernel<0><<<grid, kA32Threads, smem, s>>>(tBh, tBl, tC, p);
  }
  GR_CHECK_LAUNCH("gemm_a32_kernel");
  return GR_OK;
}
