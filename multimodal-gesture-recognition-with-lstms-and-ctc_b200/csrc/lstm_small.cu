// Keras-LSTM recurrence for NARROW layers (H <= 104, e.g. the fusion BLSTM(100) of
// /root/reference/multimodal_fusion/multimodal.py:159-168), forward and BPTT, fp32.
//
// When U (H x 4H fp32 <= 170 KB) fits in ONE SM's register file the recurrence needs no inter-CTA
// exchange at all: CTA = (direction, BS <= 4 sequences), U stays in REGISTERS for all T steps, the
// recurrent vector (h_{t-1} / dG_{t+1}) lives in shared memory, ONE __syncthreads per step.
//
// Round-2 layout (the round-1 kernels gave every thread a whole column of U and therefore read the
// whole recurrent vector per thread: 416 threads x 104 floats of broadcast LDS.128 per sequence and
// step = 1 350 cycles of shared-memory return bandwidth, 1.35 us per step at BS = 1 and 3.3 us at
// BS = 4; the global operands were fetched one step ahead only):
//   * register tiles of U are (8 outputs x H/4 inputs), 7 warps: forward thread (pair of units, K-quarter
//     kq) owns U[kq*H/4 .. +H/4][{i,f,c,o} of both units]; backward thread (octet of units, slice ks of 16)
//     owns U[8jo .. 8jo+7][ks*H/4 .. +H/4] of the flattened gate axis.  A thread reads H/4 floats of the
//     recurrent vector per sequence (8x less shared-memory traffic), does its 8*H/4 FMAs as packed FFMA2,
//     and the partial sums meet in a TRANSPOSING shuffle reduction (6 shuffles over a quad / 8 over 16
//     lanes) that leaves every lane with the two gates of one unit that it activates, stores and shares;
//   * P_t (forward) / gates_t, c_t, dy_t (backward) arrive through a ring of D stages filled by 1-D
//     bulk copies (cp.async.bulk + mbarrier complete_tx), D steps ahead of their use, so no step waits
//     for DRAM; c_{t-1} of the backward pass is the c of the next ring stage.
#include "tc_common.cuh"

namespace gr {

__device__ __forceinline__ float hsig_s(float v) { return fminf(fmaxf(0.2f * v + 0.5f, 0.f), 1.f); }
__device__ __forceinline__ float dhsig_s(float s) { return (s > 0.f && s < 1.f) ? 0.2f : 0.f; }
// tanh(x) = 1 - 2/(1 + e^{2x}) on MUFU ex2 + fast division: absolute error ~1e-7 (as lstm_tc.cu)
__device__ __forceinline__ float tanh_s(float x) {
  const float e = ex2_approx(x * 2.8853900817779268f);
  return 1.0f - __fdividef(2.0f, 1.0f + e);
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct SmallParams {
  float* gates;        // (B,T,8H)  fwd: P in / gates out (if save);  bwd: gates in / dP out
  const float* U;      // (2,H,4H)
  float* y;            // fwd out (B,T,2H)
  float* cell;         // fwd: out if save;  bwd: in
  const float* dy;     // bwd in (B,T,2H)
  int B, T, H, BS, save;
};

static constexpr int kSmallDepth = 8;   // ring stages = steps of prefetch distance

// KQ = number of recurrent inputs per thread held in registers (>= H/4): 8 (H <= 32), 16 (H <= 64), 25 (H <= 100),
// 26 (H = 104); in shared memory a quarter / slice occupies KQP = KQ rounded up to a multiple of 4 floats (16-byte
// aligned vector loads).  Seven warps at most, so that the 8 x KQ register tile of U fits the 255-register budget
// (thirteen warps of 4 x KQ tiles are capped at 128 registers: ptxas then funnels every h load through one register
// quad and the matvec becomes a serial LDS -> FFMA2 chain: 0.86 us per sequence and step).
__host__ __device__ constexpr int small_threads(int KQ) { return ((8 * KQ + 31) / 32) * 32; }
__host__ __device__ constexpr int small_kqp(int KQ) { return (KQ + 3) & ~3; }

// acc[o] += v[0..KQ) . u[o][0..KQ) for the 8 outputs of the thread; v is 16-byte aligned shared memory
template <int KQ>
__device__ __forceinline__ void small_matvec(const float* v, const float2 (&u2)[8][KQ / 2], const float (&u1)[8], float (&z)[8]) {
  float2 acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = make_float2(0.f, 0.f);
#pragma unroll
  for (int m = 0; m < KQ / 4; ++m) {
    const float4 h = reinterpret_cast<const float4*>(v)[m];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      acc[o] = __ffma2_rn(make_float2(h.x, h.y), u2[o][2 * m], acc[o]);
      acc[o] = __ffma2_rn(make_float2(h.z, h.w), u2[o][2 * m + 1], acc[o]);
    }
  }
  if (KQ % 4 >= 2) {
    const float2 h = *reinterpret_cast<const float2*>(v + (KQ & ~3));
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = __ffma2_rn(h, u2[o][KQ / 2 - 1], acc[o]);
  }
  if (KQ % 2 == 1) {
    const float h = v[KQ - 1];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o].x = fmaf(h, u1[o], acc[o].x);
  }
#pragma unroll
  for (int o = 0; o < 8; ++o) z[o] = acc[o].x + acc[o].y;
}

template <int KQ, int BS>
__global__ void __launch_bounds__(small_threads(KQ), 1) lstm_small_fwd_kernel(SmallParams p) {
  constexpr int D = kSmallDepth, KQP = small_kqp(KQ);
  extern __shared__ __align__(128) unsigned char smraw[];
  const int H = p.H, T = p.T, HQ = H >> 2, H4 = 4 * H;
  float* ring = reinterpret_cast<float*>(smraw);            // D x BS x 4H    P rows of this direction
  float* hs = ring + D * BS * H4;                            // 2 x BS x 4KQP  h_{t-1}, quarter q at [q*KQP, q*KQP + HQ), zero padded
  uint64_t* full = reinterpret_cast<uint64_t*>(hs + 2 * BS * 4 * KQP);
  const int nbg = (p.B + BS - 1) / BS;
  const int dir = blockIdx.x / nbg, b0 = (blockIdx.x % nbg) * BS;
  const int nact = min(BS, p.B - b0);
  const int tid = threadIdx.x;
  const int pair = tid >> 2, kq = tid & 3;
  const bool act = 2 * pair < H;                             // H is even: both units of the pair exist or neither
  // U tile, output o = (unit 2*pair + (o >> 2), gate o & 3), inputs k0 + i, k0 = kq*HQ, zero beyond the quarter
  float2 u2[8][KQ / 2];
  float u1[8];
  {
    const float* Ud = p.U + (size_t)dir * H * H4;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const int col = (o & 3) * H + 2 * pair + (o >> 2);
#pragma unroll
      for (int m = 0; m < KQ / 2; ++m) {
        u2[o][m].x = (act && 2 * m < HQ) ? Ud[(size_t)(kq * HQ + 2 * m) * H4 + col] : 0.f;
        u2[o][m].y = (act && 2 * m + 1 < HQ) ? Ud[(size_t)(kq * HQ + 2 * m + 1) * H4 + col] : 0.f;
      }
      u1[o] = (act && (KQ & 1) && KQ - 1 < HQ) ? Ud[(size_t)(kq * HQ + KQ - 1) * H4 + col] : 0.f;
    }
  }
  for (int e = tid; e < 2 * BS * 4 * KQP; e += blockDim.x) hs[e] = 0.f;
  if (tid == 0) {
#pragma unroll
    for (int d = 0; d < D; ++d) mbar_init(&full[d], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t G8 = (size_t)8 * H, Y2 = (size_t)2 * H;
  auto issue = [&](int s) {   // P rows of step s -> stage s % D (one elected thread)
    const int st = s % D;
    const int t = dir == 0 ? s : T - 1 - s;
    mbar_expect_tx(&full[st], (uint32_t)nact * H4 * 4u);
    for (int b = 0; b < nact; ++b)
      bulk_g2s(ring + (size_t)(st * BS + b) * H4, p.gates + ((size_t)(b0 + b) * T + t) * G8 + (size_t)dir * H4, H4 * 4u, &full[st]);
  };
  if (tid == 0)
    for (int s = 0; s < D && s < T; ++s) issue(s);
  const bool hi2 = (kq & 2) != 0, hi1 = (kq & 1) != 0;
  // after the reduction this lane owns gates (2*hi1, 2*hi1 + 1) of unit ju
  const int ju = 2 * pair + (hi2 ? 1 : 0);
  const int qh = act ? ju / HQ : 0;
  const int hslot = qh * KQP + (ju - qh * HQ);
  const int pcol = act ? (hi1 ? 2 : 0) * H + ju : 0;         // column of the lane's first gate in a P / gates row of this direction
  float creg[BS];
#pragma unroll
  for (int b = 0; b < BS; ++b) creg[b] = 0.f;
  for (int s = 0; s < T; ++s) {
    const int st = s % D;
    const int t = dir == 0 ? s : T - 1 - s;
    const float* hcur = hs + (s & 1) * (BS * 4 * KQP);
    float* hnxt = hs + ((s + 1) & 1) * (BS * 4 * KQP);
    float zz[BS][2];
#pragma unroll
    for (int b = 0; b < BS; ++b) {
      float z[8];
      small_matvec<KQ>(hcur + b * 4 * KQP + kq * KQP, u2, u1, z);
      // transposing reduction over the quad: xor 2 leaves the four gates of unit ju, xor 1 the lane's two gates
      float k[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) k[g] = (hi2 ? z[4 + g] : z[g]) + __shfl_xor_sync(0xffffffffu, hi2 ? z[g] : z[4 + g], 2);
      zz[b][0] = (hi1 ? k[2] : k[0]) + __shfl_xor_sync(0xffffffffu, hi1 ? k[0] : k[2], 1);
      zz[b][1] = (hi1 ? k[3] : k[1]) + __shfl_xor_sync(0xffffffffu, hi1 ? k[1] : k[3], 1);
    }
    // gate math of all sequences together, branch-free (one dependent chain per step, not per sequence)
    mbar_wait(&full[st], (uint32_t)(s / D) & 1u);
    const float* Pst = ring + (size_t)st * BS * H4 + pcol;
    float a0[BS], a1[BS], hv[BS];
#pragma unroll
    for (int b = 0; b < BS; ++b) {
      const float s0 = zz[b][0] + Pst[b * H4], s1 = zz[b][1] + Pst[b * H4 + H];
      const float th = tanh_s(s0), sg = hsig_s(s0);
      a0[b] = hi1 ? th : sg;                                  // gate c (tanh) on the odd lanes, gate i on the even ones
      a1[b] = hsig_s(s1);                                     // gate f / gate o
    }
#pragma unroll
    for (int b = 0; b < BS; ++b) {
      const float o0 = __shfl_xor_sync(0xffffffffu, a0[b], 1), o1 = __shfl_xor_sync(0xffffffffu, a1[b], 1);
      const float gi = hi1 ? o0 : a0[b], gf = hi1 ? o1 : a1[b], gg = hi1 ? a0[b] : o0, go = hi1 ? a1[b] : o1;
      const float c = gf * creg[b] + gi * gg;                 // both lanes of the unit carry the (bit-identical) cell state
      creg[b] = c;
      hv[b] = go * tanh_s(c);
    }
    {
      const size_t row0 = (size_t)b0 * T + t;
      float* yp = p.y + row0 * Y2 + (size_t)dir * H + ju;
      float* cp = p.cell + row0 * Y2 + (size_t)dir * H + ju;
      float* gp = p.gates + row0 * G8 + (size_t)dir * H4 + pcol;
      const size_t ysq = (size_t)T * Y2, gsq = (size_t)T * G8;
      const bool sv = p.save != 0;
#pragma unroll
      for (int b = 0; b < BS; ++b) {
        const bool on = act && b < nact;
        if (on && !hi1) {
          hnxt[b * 4 * KQP + hslot] = hv[b];
          yp[b * ysq] = hv[b];
        }
        if (on && hi1 && sv) cp[b * ysq] = creg[b];
        if (on && sv) {
          gp[b * gsq] = a0[b];
          gp[b * gsq + H] = a1[b];
        }
      }
    }
    __syncthreads();
    if (tid == 0 && s + D < T) issue(s + D);
  }
}

template <int KQ, int BS>
__global__ void __launch_bounds__(small_threads(KQ), 1) lstm_small_bwd_kernel(SmallParams p) {
  constexpr int D = kSmallDepth, KQP = small_kqp(KQ);
  extern __shared__ __align__(128) unsigned char smraw[];
  const int H = p.H, T = p.T, HQ = H >> 2, H4 = 4 * H, H6 = 6 * H;
  float* ring = reinterpret_cast<float*>(smraw);            // D x BS x 6H    [gates 4H | c H | dy H] of this direction
  float* dgs = ring + D * BS * H6;                           // 2 x BS x 16KQP dG_{next}, slice k at [k*KQP, k*KQP + HQ), zero padded
  uint64_t* full = reinterpret_cast<uint64_t*>(dgs + 2 * BS * 16 * KQP);
  const int nbg = (p.B + BS - 1) / BS;
  const int dir = blockIdx.x / nbg, b0 = (blockIdx.x % nbg) * BS;
  const int nact = min(BS, p.B - b0);
  const int tid = threadIdx.x, lane = tid & 31;
  const int ks = lane & 15, jo = 2 * (tid >> 5) + (lane >> 4);
  // U tile, output r = unit 8*jo + r, inputs = slice ks of the flattened gate axis: columns ks*HQ + i
  float2 uu2[8][KQ / 2];
  float uu1[8];
  {
    const float* Ud = p.U + (size_t)dir * H * H4;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int unit = 8 * jo + r;
      const bool ok = unit < H;
#pragma unroll
      for (int m = 0; m < KQ / 2; ++m) {
        uu2[r][m].x = (ok && 2 * m < HQ) ? Ud[(size_t)unit * H4 + ks * HQ + 2 * m] : 0.f;
        uu2[r][m].y = (ok && 2 * m + 1 < HQ) ? Ud[(size_t)unit * H4 + ks * HQ + 2 * m + 1] : 0.f;
      }
      uu1[r] = (ok && (KQ & 1) && KQ - 1 < HQ) ? Ud[(size_t)unit * H4 + ks * HQ + KQ - 1] : 0.f;
    }
  }
  for (int e = tid; e < 2 * BS * 16 * KQP; e += blockDim.x) dgs[e] = 0.f;
  if (tid == 0) {
#pragma unroll
    for (int d = 0; d < D; ++d) mbar_init(&full[d], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t G8 = (size_t)8 * H, Y2 = (size_t)2 * H;
  auto issue = [&](int sp) {   // operands of backward step sp -> stage sp % D (one elected thread)
    const int st = sp % D;
    const int s = T - 1 - sp;
    const int t = dir == 0 ? s : T - 1 - s;
    mbar_expect_tx(&full[st], (uint32_t)nact * H6 * 4u);
    for (int b = 0; b < nact; ++b) {
      const size_t row = (size_t)(b0 + b) * T + t;
      float* dst = ring + (size_t)(st * BS + b) * H6;
      bulk_g2s(dst, p.gates + row * G8 + (size_t)dir * H4, H4 * 4u, &full[st]);
      bulk_g2s(dst + H4, p.cell + row * Y2 + (size_t)dir * H, H * 4u, &full[st]);
      bulk_g2s(dst + H4 + H, p.dy + row * Y2 + (size_t)dir * H, H * 4u, &full[st]);
    }
  };
  if (tid == 0)
    for (int sp = 0; sp < D && sp < T; ++sp) issue(sp);
  // element-wise role after the reduction: lane bits 3..1 pick the unit of the octet, bit 0 the gate pair
  const bool b3 = (lane & 8) != 0, b2 = (lane & 4) != 0, b1 = (lane & 2) != 0, gh = (lane & 1) != 0;
  const int uj = 8 * jo + (b3 ? 4 : 0) + (b2 ? 2 : 0) + (b1 ? 1 : 0);
  const bool eact = uj < H;
  const int ujs = eact ? uj : 0;                               // safe ring index for the idle lanes of the last octet
  const int c0 = (gh ? 2 : 0) * H + uj, c1 = c0 + H;          // flattened gate columns of the lane's two gates
  const int q0 = eact ? c0 / HQ : 0, q1 = eact ? c1 / HQ : 0;
  const int dslot0 = q0 * KQP + (c0 - q0 * HQ), dslot1 = q1 * KQP + (c1 - q1 * HQ);
  float dcreg[BS];
#pragma unroll
  for (int b = 0; b < BS; ++b) dcreg[b] = 0.f;
  for (int sp = 0; sp < T; ++sp) {
    const int st = sp % D;
    const int s = T - 1 - sp;
    const int t = dir == 0 ? s : T - 1 - s;
    const bool has_prev = sp + 1 < T;
    const float* dcur = dgs + (sp & 1) * (BS * 16 * KQP);
    float* dnxt = dgs + ((sp + 1) & 1) * (BS * 16 * KQP);
    float dhv[BS];
#pragma unroll
    for (int b = 0; b < BS; ++b) {
      float z[8];
      small_matvec<KQ>(dcur + b * 16 * KQP + ks * KQP, uu2, uu1, z);
      // transposing reduction over the 16 slices: every lane ends with the full sum of unit uj
      float k[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) k[r] = (b3 ? z[4 + r] : z[r]) + __shfl_xor_sync(0xffffffffu, b3 ? z[r] : z[4 + r], 8);
      const float q0v = (b2 ? k[2] : k[0]) + __shfl_xor_sync(0xffffffffu, b2 ? k[0] : k[2], 4);
      const float q1v = (b2 ? k[3] : k[1]) + __shfl_xor_sync(0xffffffffu, b2 ? k[1] : k[3], 4);
      float r_ = (b1 ? q1v : q0v) + __shfl_xor_sync(0xffffffffu, b1 ? q0v : q1v, 2);
      r_ += __shfl_xor_sync(0xffffffffu, r_, 1);
      dhv[b] = r_;
    }
    mbar_wait(&full[st], (uint32_t)(sp / D) & 1u);
    if (has_prev) mbar_wait(&full[(sp + 1) % D], (uint32_t)((sp + 1) / D) & 1u);   // c_{t-1} = the c of the next stage
    {
      const float* R0 = ring + (size_t)st * BS * H6 + ujs;
      const float* R1 = ring + (size_t)((sp + 1) % D) * BS * H6 + H4 + ujs;
      float* gr = p.gates + ((size_t)b0 * T + t) * G8 + (size_t)dir * H4;
      const size_t gsq = (size_t)T * G8;
      float m0[BS], m1[BS];
#pragma unroll
      for (int b = 0; b < BS; ++b) {
        const float* R = R0 + b * H6;
        const float gi = R[0], gf = R[H], gg = R[2 * H], go = R[3 * H];
        const float c = R[H4];
        const float dh = dhv[b] + R[H4 + H];
        const float cp = has_prev ? R1[b * H6] : 0.f;
        const float tc = tanh_s(c);
        const float dc = dcreg[b] + dh * go * (1.f - tc * tc);   // both lanes of the unit carry the (bit-identical) dc
        dcreg[b] = dc * gf;
        m0[b] = gh ? dc * gi * (1.f - gg * gg) : dc * gg * dhsig_s(gi);      // d_c  | d_i
        m1[b] = gh ? dh * tc * dhsig_s(go) : dc * cp * dhsig_s(gf);          // d_o  | d_f
      }
#pragma unroll
      for (int b = 0; b < BS; ++b) {
        if (eact && b < nact) {
          gr[b * gsq + c0] = m0[b];
          gr[b * gsq + c1] = m1[b];
          dnxt[b * 16 * KQP + dslot0] = m0[b];
          dnxt[b * 16 * KQP + dslot1] = m1[b];
        }
      }
    }
    __syncthreads();
    if (tid == 0 && sp + D < T) issue(sp + D);
  }
}

bool lstm_small_supported(int H) { return H % 4 == 0 && H >= 4 && H <= 104; }

static int small_bs(int B) {
  const int per_dir = max(1, num_sms() / 2);
  const int bs = (B + per_dir - 1) / per_dir;
  return bs <= 1 ? 1 : (bs <= 2 ? 2 : 4);
}

template <int KQ, int BS>
static int small_launch(bool bwd, SmallParams& p, cudaStream_t s) {
  constexpr int D = kSmallDepth, KQP = small_kqp(KQ);
  const int nbg = (p.B + BS - 1) / BS;
  // forward: 4 lanes per pair of units; backward: 16 lanes per octet of units
  const int threads = ((max(4 * ((p.H + 1) / 2), 16 * ((p.H + 7) / 8)) + 31) / 32) * 32;
  const size_t tail = sizeof(uint64_t) * D;
  if (!bwd) {
    const size_t smem = sizeof(float) * ((size_t)D * BS * 4 * p.H + (size_t)2 * BS * 4 * KQP) + tail;
    GR_CUDA(cudaFuncSetAttribute(lstm_small_fwd_kernel<KQ, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_small_fwd_kernel<KQ, BS><<<2 * nbg, threads, smem, s>>>(p);
  } else {
    const size_t smem = sizeof(float) * ((size_t)D * BS * 6 * p.H + (size_t)2 * BS * 16 * KQP) + tail;
    GR_CUDA(cudaFuncSetAttribute(lstm_small_bwd_kernel<KQ, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_small_bwd_kernel<KQ, BS><<<2 * nbg, threads, smem, s>>>(p);
  }
  GR_CHECK_LAUNCH("lstm_small_kernel");
  return GR_OK;
}

template <int KQ>
static int small_launch_bs(bool bwd, SmallParams& p, cudaStream_t s) {
  if (p.BS == 1) return small_launch<KQ, 1>(bwd, p, s);
  if (p.BS == 2) return small_launch<KQ, 2>(bwd, p, s);
  return small_launch<KQ, 4>(bwd, p, s);
}

// the bulk copies need 16-byte aligned rows: H % 4 == 0 (checked by lstm_small_supported) and aligned bases
bool lstm_small_aligned(const float* gates, const float* cell, const float* dy) {
  return ((reinterpret_cast<uintptr_t>(gates) | reinterpret_cast<uintptr_t>(cell) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0;
}

int lstm_small_run(bool bwd, float* gates, const float* U, int B, int T, int H, float* y, float* cell,
                   const float* dy, cudaStream_t s) {
  SmallParams p;
  p.gates = gates; p.U = U; p.y = y; p.cell = cell; p.dy = dy; p.B = B; p.T = T; p.H = H;
  p.BS = small_bs(B);
  if (const char* e = getenv("GR_SMALL_BS")) {   // experiments: sequences per CTA
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4) p.BS = v;
  }
  p.save = (!bwd && cell != nullptr) ? 1 : 0;
  // B > 4 * (SMs/2) sequences per direction are covered by more CTAs than SMs (several waves)
  if (H <= 32) return small_launch_bs<8>(bwd, p, s);
  if (H <= 64) return small_launch_bs<16>(bwd, p, s);
  if (H <= 100) return small_launch_bs<25>(bwd, p, s);
  return small_launch_bs<26>(bwd, p, s);
}

}  // namespace gr
