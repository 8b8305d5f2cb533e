// CTC loss + gradient for sm_100a.  Replaces ctc_lambda_func -> K.ctc_batch_cost -> tf.nn.ctc_loss
// (/root/reference/audio_network/losses.py:4-15; TF CTCLossCalculator semantics).
//
// Design: one CTA of two warps per sequence.  Warp 0 runs the alpha recursion
// forward in time; warp 1 runs the SAME recursion on the time-reversed, label-reversed problem,
// which is the beta recursion with the emission folded in (gamma = beta + log y).  They meet at
// t* = Tn/2: log p = logsumexp_u(alpha + gamma - log y)(t*).  Each warp then keeps going through
// the other half, where the other warp's stored lattice rows turn every step directly into
// posterior occupancies, so the softmax/CTC gradient is produced on the fly and only HALF of the
// (alpha, beta) lattice ever touches memory.  Log space (base 2, MUFU ex2/lg2), extended labels
// split into blank states b_k (k = 0..L) and label states l_k (k = 0..L-1), K consecutive k per
// lane, one shuffle per step.  Probabilities are streamed in chunks of TC frames through shared
// memory with coalesced loads; gradients leave the same way.
#include <algorithm>
#include <stdlib.h>
#include "common.cuh"

namespace gr {

static constexpr float kNeg = -1.0e30f;  // "log zero": finite so that no inf-inf NaN can arise

__device__ __forceinline__ float lse2(float a, float b) {
  // lo - hi == -|a - b| exactly, so the max and the difference do not depend on each other (one
  // instruction and one dependent latency less per call; the |.| and the sign fold into the MUFU operand)
  return fmaxf(a, b) + lg2_approx(1.0f + ex2_approx(-fabsf(a - b)));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  float hi = fmaxf(a, b), lo = fminf(a, b);
  float m = fmaxf(hi, c), x1 = fminf(hi, c);
  return m + lg2_approx(1.0f + ex2_approx(x1 - m) + ex2_approx(lo - m));
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}

__device__ __forceinline__ void sts_f32_if(uint32_t a, float v, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.f32 [%0], %1;\n\t}" ::"r"(a), "f"(v), "r"((uint32_t)pred) : "memory");
}

struct CtcParams {
  const float* x;
  int is_logits, B, T, C, drop;
  float eps;
  const int32_t* labels;
  int Lmax;
  const int32_t* label_len;
  const int32_t* input_len;
  const float* upstream;
  float* loss;
  float* grad;
  int32_t* status;
  float* ws;
  size_t ws_seq_floats;
  int RS;  // lattice row stride (floats)
};

__host__ __device__ inline int ctc_cp(int C) { return C | 1; }
__host__ __device__ inline int ctc_rsp(int RS) { return RS | 1; }

static int ctc_row_stride(int Lmax) { return (2 * Lmax + 3 + 3) & ~3; }  // states + per-row offset (hi, lo) in the last two slots; 16-byte rows

// ============================================================================================
// v4 (round 1; GR_CTC_IMPL=v4).  What it changed against its predecessor v2 (removed in round 2) is everything
// AROUND the recursion, which is where v2 spent more than half of its ~200 warp-instructions per step:
//  * chunks of 32 frames wherever shared memory allows (v2: 16 at the BASELINE config-4 shape, so the
//    lane-per-row passes ran half empty).  The room comes from a different phase-1 layout: the
//    probabilities are single-buffered there, the gradient is formed in place over them, and the
//    other warp's lattice rows arrive through a circular buffer of TC+PD rows filled row by row,
//    PD rows ahead of the consumer, with one 16-byte cp.async per lane (rows are 16-byte aligned,
//    RS % 4 == 0) instead of a second full chunk buffer filled with 4-byte copies;
//  * probabilities are staged with 8-byte cp.async when C is even; gradients leave as float2;
//  * the per-row constant (my offset + the other warp's offset - log p) is formed in the step from the
//    two offset words that travel with the lattice row, not in a pre-pass over the staged chunk;
//  * occupancy -> class reduction gathers through a class-sorted index list built once per sequence
//    (labels are fixed), instead of read-modify-write accumulation into a zeroed row;
//  * the first step of each phase is peeled, so the loop body carries no first-step predicates, and
//    invalid blank slots are kept at "log zero" by the arithmetic itself instead of by masks.
template <int K, int TC>
__global__ void __launch_bounds__(64) ctc_loss_grad_kernel_v4(CtcParams p) {
  constexpr int PD = 4;          // lattice rows in flight ahead of the consumer (phase 1)
  constexpr int EM = TC + PD;    // circular lattice buffer, rows
  extern __shared__ float smem[];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.C, T = p.T, Lmax = p.Lmax, RS = p.RS;
  const int blank = C - 1;
  // smem carve-up (floats): [labs Lmax4][cstart C+1 -> 4][clist Lmax4][bcast 4] then per warp
  //   phase 0: X double buffer (2 x TC x C);  phase 1: X (TC x C) | E circular (EM x RS)
  const int Lmax4 = (Lmax + 3) & ~3, C4 = (C + 1 + 3) & ~3;
  int* labs = reinterpret_cast<int*>(smem);
  int* cstart = labs + Lmax4;
  int* clist = cstart + C4;
  float* bcast = smem + 2 * Lmax4 + C4;
  const int xc = TC * C;                                   // multiple of 4 (TC is)
  const int per_warp = max(2 * xc, xc + EM * RS);
  float* Xs = bcast + 4 + warp * per_warp;
  float* Es = Xs + xc;

  const float* xb = p.x + (size_t)b * T * C;
  float* gb = p.grad ? p.grad + (size_t)b * T * C : nullptr;
  const int32_t* lab_g = p.labels + (size_t)b * Lmax;
  const int Lraw = p.label_len[b];
  const int Tn = p.input_len[b];

  // ---- validation (TF CTCLossOp order) + label rule: a label >= C-1 terminates the sequence
  int st = GR_CTC_OK;
  int L = 0;
  if (Tn < 1 || Tn > T - p.drop) st = GR_CTC_BAD_INPUT_LENGTH;
  else if (Lraw <= 0) st = GR_CTC_ZERO_LABELS;
  else {
    int first_null = Lraw, last_nonnull = -1, bad = 0;
    const int Lr = min(Lraw, Lmax);
    for (int k = lane; k < Lr; k += 32) {
      int v = lab_g[k];
      if (v >= blank) first_null = min(first_null, k);
      else { last_nonnull = max(last_nonnull, k); if (v < 0) bad = 1; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      first_null = min(first_null, __shfl_xor_sync(0xffffffffu, first_null, o));
      last_nonnull = max(last_nonnull, __shfl_xor_sync(0xffffffffu, last_nonnull, o));
      bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    L = min(first_null, Lr);
    if (last_nonnull > first_null || bad || Lraw > Lmax) st = GR_CTC_NONNULL_AFTER_NULL;
    else if (Lraw > Tn) st = GR_CTC_NOT_ENOUGH_TIME;
  }
  if (st != GR_CTC_OK) {  // block-uniform
    if (threadIdx.x == 0) {
      if (p.status) p.status[b] = st;
      p.loss[b] = __int_as_float(0x7fc00000);
    }
    if (gb) for (int e = threadIdx.x; e < T * C; e += 64) gb[e] = 0.f;
    return;
  }
  for (int k = threadIdx.x; k < L; k += 64) labs[k] = lab_g[k];
  if (gb) {  // gradient rows outside [drop, drop+Tn) are zero
    for (int e = threadIdx.x; e < p.drop * C; e += 64) gb[e] = 0.f;
    for (int e = (p.drop + Tn) * C + threadIdx.x; e < T * C; e += 64) gb[e] = 0.f;
  }
  __syncthreads();
  if (gb) {  // class-sorted list of label positions: clist[cstart[c] .. cstart[c+1]) = {k : labs[k] == c}, ascending
    for (int c = threadIdx.x; c < C; c += 64) {
      int n = 0;
      for (int k = 0; k < L; ++k) n += labs[k] == c;
      cstart[c + 1] = n;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      cstart[0] = 0;
      for (int c = 0; c < C; ++c) cstart[c + 1] += cstart[c];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 64) {
      int m = cstart[c];
      for (int k = 0; k < L; ++k)
        if (labs[k] == c) clist[m++] = k;
    }
    __syncthreads();
  }

  const int dir = warp;  // 0: alpha, natural order; 1: gamma, reversed time and labels
  const int tstar = Tn >> 1;
  float* wsb = p.ws + (size_t)b * p.ws_seq_floats;
  // alpha rows t<=t* live at row t; gamma rows t>=t* live at row t+1 (disjoint)
  float* my_rows = wsb + (dir == 0 ? 0 : RS);
  const float* other_rows = wsb + (dir == 0 ? RS : 0);

  // per-lane state description
  float sb[K], sl[K];
  int labr[K], posb[K], posl[K];
  bool vb[K], vl[K], skip[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int k = lane * K + j;
    vb[j] = k <= L;
    vl[j] = k < L;
    const int nk_l = dir == 0 ? k : L - 1 - k;
    labr[j] = vl[j] ? labs[nk_l] : 0;
    int prev = 0;
    if (vl[j] && k >= 1) prev = labs[dir == 0 ? k - 1 : L - k];
    skip[j] = vl[j] && k >= 1 && labr[j] != prev;
    posb[j] = dir == 0 ? k : L - k;
    posl[j] = Lmax + 1 + nk_l;
    sb[j] = kNeg;
    sl[j] = kNeg;
  }
  const float up_scale = p.upstream ? p.upstream[b] : 1.0f;
  const float eps = p.eps;
  const bool vec2 = ((C & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 7) == 0);
  const bool gvec2 = ((C & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.grad) & 7) == 0);

  double off = 0.0;       // states are kept relative to this running offset, re-centred once per chunk
  double logp2 = 0.0;
  bool novalid = false;
  float lpb_last = kNeg, lpl_last[K];
#pragma unroll
  for (int j = 0; j < K; ++j) lpl_last[j] = kNeg;

  auto stage_x = [&](float* dx, int tlo, int n) {
    const float* src = xb + (size_t)(p.drop + tlo) * C;
    const int tot = n * C;
    if (vec2) {
      for (int e = 2 * lane; e < tot; e += 64)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dx + e)), "l"(src + e) : "memory");
    } else {
      for (int e = lane; e < tot; e += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dx + e)), "l"(src + e) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // ---- per-row pre-pass: lp2 = log2 q, q = softmax(log(p + eps)) = (p+eps)/sum(p+eps); returns Z of the lane's row
  auto pre_pass = [&](float* Xc, int n) -> float {
    float Z = 1.f;
    if (lane < n) {
      float* row = Xc + lane * C;
      Z = 0.f;
      if (p.is_logits) {
        float m = row[0];
        for (int c = 1; c < C; ++c) m = fmaxf(m, row[c]);
        float s_ = 0.f;
        for (int c = 0; c < C; ++c) { float e_ = ex2_approx((row[c] - m) * kLog2e); row[c] = e_; s_ += e_; }
        const float inv = 1.0f / s_;
        for (int c = 0; c < C; ++c) { float pe = row[c] * inv + eps; row[c] = pe; Z += pe; }
      } else {
        for (int c = 0; c < C; ++c) { float pe = row[c] + eps; row[c] = pe; Z += pe; }
      }
      const float lz = lg2_approx(Z);
      for (int c = 0; c < C; ++c) row[c] = fmaxf(lg2_approx(row[c]) - lz, kNeg);
    }
    return Z;
  };
  auto recentre = [&]() {
    float mx = kNeg;
#pragma unroll
    for (int j = 0; j < K; ++j) mx = fmaxf(mx, fmaxf(sb[j], sl[j]));
    mx = warp_max(mx);
    if (mx > -1.0e29f) {
#pragma unroll
      for (int j = 0; j < K; ++j) { sb[j] -= mx; sl[j] -= mx; }
      off += (double)mx;
    }
  };
  // one step of the recursion; invalid label slots are pinned to "log zero" through lpl = kNeg, and the
  // invalid blank slots behind them stay there on their own (lse2(kNeg, kNeg) + lpb == kNeg in fp32)
  auto recur = [&](float lpb, const float (&lpl)[K]) {
    float upv = __shfl_up_sync(0xffffffffu, sl[K - 1], 1);
    if (lane == 0) upv = kNeg;
    float nb[K], nl[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const float prevl = (j == 0) ? upv : sl[j - 1];
      const float tb = lse2(sb[j], prevl);        // = the blank update; shared by the skip transition
      nb[j] = tb + lpb;
      nl[j] = lse2(sl[j], skip[j] ? tb : sb[j]) + lpl[j];
    }
#pragma unroll
    for (int j = 0; j < K; ++j) { sb[j] = nb[j]; sl[j] = nl[j]; }
  };

  // =========================== phase 0: own half, lattice rows to memory ===========================
  // invalid state slots write to the row's unused pad word (RS-3) instead of being predicated off
  const int padpos = RS - 3;
  int gposb[K], gposl[K];
  uint32_t xoffl[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    gposb[j] = vb[j] ? posb[j] : padpos;
    gposl[j] = vl[j] ? posl[j] : padpos;
    xoffl[j] = 4u * (uint32_t)labr[j];
  }
  const uint32_t xoffb = 4u * (uint32_t)blank;
  {
    const int i_end = dir == 0 ? tstar + 1 : Tn - tstar;
    const int nchunks = (i_end + TC - 1) / TC;
    auto chunk_range = [&](int ci, int& ic, int& ie, int& tlo) {
      ic = ci * TC;
      ie = min(ic + TC, i_end);
      tlo = dir == 0 ? ic : Tn - ie;
    };
    { int ic, ie, tlo; chunk_range(0, ic, ie, tlo); stage_x(Xs, tlo, ie - ic); }
    for (int ci = 0; ci < nchunks; ++ci) {
      int ic, ie, tlo;
      chunk_range(ci, ic, ie, tlo);
      const int n = ie - ic;
      float* Xc = Xs + (ci & 1) * xc;
      if (ci + 1 < nchunks) {
        int ic2, ie2, tlo2;
        chunk_range(ci + 1, ic2, ie2, tlo2);
        stage_x(Xs + ((ci + 1) & 1) * xc, tlo2, ie2 - ic2);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncwarp();
      if (ic > 0) recentre();
      const float2 offw = make_float2((float)off, (float)(off - (double)(float)off));
      pre_pass(Xc, n);
      __syncwarp();
      const int rstep = dir == 0 ? 4 * C : -4 * C;
      uint32_t xrow = (uint32_t)__cvta_generic_to_shared(Xc + (dir == 0 ? 0 : n - 1) * C);
      float* dst = my_rows + (size_t)(dir == 0 ? ic : Tn - 1 - ic) * RS;
      const int dstep = dir == 0 ? RS : -RS;
      float lpb = kNeg, lpl[K];
      auto load_lp = [&]() {
        lpb = lds_f32(xrow + xoffb);
#pragma unroll
        for (int j = 0; j < K; ++j) { const float v = lds_f32(xrow + xoffl[j]); lpl[j] = vl[j] ? v : kNeg; }
      };
      auto emit = [&]() {
#pragma unroll
        for (int j = 0; j < K; ++j) { dst[gposb[j]] = sb[j]; dst[gposl[j]] = sl[j]; }
        if (lane == 0) *reinterpret_cast<float2*>(dst + RS - 2) = offw;
        xrow += rstep;
        dst += dstep;
      };
      int i = ic;
      if (ic == 0) {   // initial state of the recursion
        load_lp();
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const int k = lane * K + j;
          sb[j] = (k == 0) ? lpb : kNeg;
          sl[j] = (k == 0) ? lpl[j] : kNeg;
        }
        emit();
        ++i;
      }
      for (; i < ie; ++i) {
        load_lp();
        recur(lpb, lpl);
        emit();
      }
      lpb_last = lpb;
#pragma unroll
      for (int j = 0; j < K; ++j) lpl_last[j] = lpl[j];
      __syncwarp();   // every lane is done with Xc before the next-but-one chunk is staged over it
    }
  }
  // =========================== meeting row: log p ===========================
  __syncthreads();  // both half-lattices are in memory
  if (dir == 0) {
    const float* orow = other_rows + (size_t)tstar * RS;
    float v[2 * K];
    float m = kNeg;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      v[2 * j] = vb[j] ? sb[j] + orow[posb[j]] - lpb_last : kNeg;
      v[2 * j + 1] = vl[j] ? sl[j] + orow[posl[j]] - lpl_last[j] : kNeg;
      m = fmaxf(m, fmaxf(v[2 * j], v[2 * j + 1]));
    }
    m = warp_max(m);
    float s_ = 0.f;
#pragma unroll
    for (int j = 0; j < 2 * K; ++j) s_ += ex2_approx(v[j] - m);
    s_ = warp_sum(s_);
    const float lp2 = m + lg2_approx(s_);
    if (lane == 0) {
      const bool nv = !(m > -1.0e29f) || !(lp2 > -1.0e29f);
      double* bd = reinterpret_cast<double*>(bcast);
      bd[0] = nv ? -1.0e300 : off + (double)orow[RS - 2] + (double)orow[RS - 1] + (double)lp2;
    }
  }
  __syncthreads();
  logp2 = reinterpret_cast<const double*>(bcast)[0];
  novalid = !(logp2 > -1.0e299);
  if (threadIdx.x == 0) {
    p.loss[b] = novalid ? __int_as_float(0x7f800000) : (float)(-logp2 * 0.6931471805599453);
    if (p.status) p.status[b] = novalid ? GR_CTC_NO_VALID_PATH : GR_CTC_OK;
  }
  if (p.grad == nullptr) return;

  // =========================== phase 1: the other half, occupancies -> gradient ===========================
  {
    // dir 0 re-uses its state at row t* for the first row (no recursion step); dir 1 starts one row further
    const int i_begin = dir == 0 ? tstar : Tn - tstar;
    const int i_end = Tn;
    const int nchunks = (i_end - i_begin + TC - 1) / TC;
    const uint32_t e_base = (uint32_t)__cvta_generic_to_shared(Es);
    const uint32_t e_bytes = (uint32_t)(EM * RS) * 4u, e_pitch = (uint32_t)RS * 4u;
    uint32_t eoffb[K], eoffl[K];
#pragma unroll
    for (int j = 0; j < K; ++j) { eoffb[j] = 4u * (uint32_t)gposb[j]; eoffl[j] = 4u * (uint32_t)gposl[j]; }
    // prefetch stream: row i + PD of the other warp's half-lattice -> circular slot, one 16-byte copy per lane
    const bool pf_lane = 4 * lane < RS;
    const bool wide_rows = RS > 128;
    const float* pf_src = other_rows + (size_t)(dir == 0 ? i_begin : Tn - 1 - i_begin) * RS + 4 * lane;
    const ptrdiff_t pf_step = dir == 0 ? RS : -RS;
    uint32_t pf_dst = e_base + 16u * (uint32_t)lane;
    const uint32_t pf_end = e_base + e_bytes + 16u * (uint32_t)lane;
    int pf_left = i_end - i_begin;
    auto fetch_next = [&]() {
      if (pf_lane && pf_left > 0) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(pf_dst), "l"(pf_src) : "memory");
        if (wide_rows)   // rows wider than 32 x 16 bytes (Lmax > 62)
          for (int o = 128 + 4 * lane; o < RS; o += 128)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(pf_dst + 4u * (uint32_t)o - 16u * (uint32_t)lane), "l"(pf_src + o - 4 * lane) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      pf_src += pf_step;
      pf_dst += e_pitch;
      if (pf_dst == pf_end) pf_dst -= e_bytes;
      --pf_left;
    };
#pragma unroll
    for (int d = 0; d < PD; ++d) fetch_next();
    uint32_t erow = e_base;
    int slot = 0;
    for (int ci = 0; ci < nchunks; ++ci) {
      const int ic = i_begin + ci * TC;
      const int ie = min(ic + TC, i_end);
      const int n = ie - ic;
      const int tlo = dir == 0 ? ic : Tn - ie;
      const int slot0 = slot;
      stage_x(Xs, tlo, n);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      recentre();
      const double D0 = novalid ? -1.0e30 : off - logp2;     // no valid path: every occupancy becomes ex2(-huge) = 0
      const float Zrow = pre_pass(Xs, n);
      __syncwarp();
      const int rstep = dir == 0 ? 4 * C : -4 * C;
      uint32_t xrow = (uint32_t)__cvta_generic_to_shared(Xs + (dir == 0 ? 0 : n - 1) * C);
      float lpb, lpl[K];
      auto load_lp = [&]() {
        lpb = lds_f32(xrow + xoffb);
#pragma unroll
        for (int j = 0; j < K; ++j) { const float v = lds_f32(xrow + xoffl[j]); lpl[j] = vl[j] ? v : kNeg; }
      };
      // the other warp's row is read BEFORE the recursion step so that its latency (and the offset arithmetic)
      // overlaps the MUFU chain of the step; the row constant is re-formed only when the offsets change
      float ev_b[K], ev_l[K];
      float2 oo_prev = make_float2(__int_as_float(0x7fc00000), 0.f);
      float cst = kNeg;
      auto emit_load = [&]() {
        asm volatile("cp.async.wait_group %0;" ::"n"(PD) : "memory");
        __syncwarp();
        const float2 oo = lds_f32x2(erow + e_pitch - 8u);
#pragma unroll
        for (int j = 0; j < K; ++j) { ev_b[j] = lds_f32(erow + eoffb[j]); ev_l[j] = lds_f32(erow + eoffl[j]); }
        if (oo.x != oo_prev.x || oo.y != oo_prev.y) {   // warp-uniform: once per chunk of the producer
          cst = (float)(D0 + (double)oo.x + (double)oo.y);
          oo_prev = oo;
        }
      };
      auto emit = [&]() {
        const float cb = cst - lpb;
#pragma unroll
        for (int j = 0; j < K; ++j) {
          sts_f32_if(erow + eoffb[j], ex2_approx((sb[j] + ev_b[j]) + cb), vb[j]);      // (predicated, not routed to the
          sts_f32_if(erow + eoffl[j], ex2_approx((sl[j] + ev_l[j]) + (cst - lpl[j])), vl[j]);  //  pad word: no smem race)
        }
        xrow += rstep;
        erow += e_pitch;
        ++slot;
        if (slot == EM) { slot = 0; erow = e_base; }
      };
      int i = ic;
      if (dir == 0 && ci == 0) {   // row t*: the state is already there
        fetch_next();
        emit_load();
        load_lp();
        emit();
        ++i;
      }
      for (; i < ie; ++i) {
        fetch_next();
        emit_load();
        load_lp();
        recur(lpb, lpl);
        emit();
      }
      __syncwarp();
      // ---- per-row post-pass: occupancies -> gradient, in place over X
      if (lane < n) {
        // lane = row of the chunk (natural time order), classes walked through the class-sorted index list
        float* xrow = Xs + lane * C;
        int sidx = slot0 + (dir == 0 ? lane : n - 1 - lane);
        if (sidx >= EM) sidx -= EM;
        const float* erow = Es + sidx * RS;
        const float* elab = erow + Lmax + 1;
        float occb = 0.f;
        for (int k = 0; k <= L; ++k) occb += erow[k];
        const float Z = Zrow;
        auto occ_of = [&](int c) -> float {
          if (c == blank) return occb;
          float o = 0.f;
          const int m1 = cstart[c + 1];
          for (int m = cstart[c]; m < m1; ++m) o += elab[clist[m]];
          return o;
        };
        if (p.is_logits) {
          float dot = 0.f;
          for (int c = 0; c < C; ++c) {
            const float q = ex2_approx(xrow[c]);
            const float gz = up_scale * (q - occ_of(c));
            const float pe = q * Z;                  // p + eps
            const float pr = fmaxf(pe - eps, 0.f);   // p
            dot += (pr / pe) * gz;
          }
          for (int c = 0; c < C; ++c) {
            const float q = ex2_approx(xrow[c]);
            const float gz = up_scale * (q - occ_of(c));
            const float pe = q * Z;
            const float pr = fmaxf(pe - eps, 0.f);
            xrow[c] = (pr / pe) * gz - pr * dot;
          }
        } else {
          for (int c = 0; c < C; ++c) {
            const float q = ex2_approx(xrow[c]);
            xrow[c] = up_scale * __fdividef(q - occ_of(c), q * Z);
          }
        }
      }
      __syncwarp();
      float* gdst = gb + (size_t)(p.drop + tlo) * C;
      const int tot = n * C;
      if (gvec2) {
        for (int e = 2 * lane; e < tot; e += 64) *reinterpret_cast<float2*>(gdst + e) = *reinterpret_cast<const float2*>(Xs + e);
      } else {
        for (int e = lane; e < tot; e += 32) gdst[e] = Xs[e];
      }
      __syncwarp();
    }
  }
}


__host__ __device__ inline int ctc5_lp(int Lmax) { return (Lmax + 1) | 1; }                       // OCC line: slot 0 + Lmax labels, odd pitch
__host__ __device__ inline int ctc5_nlp(int Lmax) { const int K = (Lmax + 1 + 31) / 32; return ((Lmax + K) / K) | 1; }   // BL line: one float per lane
__host__ __device__ inline int ctc5_per_warp(int C, int Lmax, int RS, int TC) {
  const int xc = TC * C;
  const int p1 = xc + 12 * RS + TC * ctc5_lp(Lmax) + TC * ctc5_nlp(Lmax);
  const int m = 2 * xc > p1 ? 2 * xc : p1;
  return (m + 3) & ~3;
}

// ============================================================================================
// v5 (default).  v4's algorithm and arithmetic with fewer instructions around the recursion (v4: 68 / ~110
// warp-instructions per step in the two phases against 32 of recursion arithmetic):
//  * lattice rows are stored in the PRODUCER's state order, interleaved [b_0, l_0, b_1, l_1, ...], so a lane's 2K states
//    are contiguous: ONE vector store and one 64-bit pointer increment per lane and step (v4: 2K scalar stores, each with
//    its own 64-bit address); the consumer's positions are 2(L-k) / 2(L-1-k)+1 in that order;
//  * the emissions of slots behind the last state are no longer masked every step (see the note at the state set-up);
//  * the occupancy kept per state is (occupancy x emission probability); the division happens once per class in the
//    post-pass, where 1/q = ex2(-log2 q) needs no division at all.
template <int K, int TC>
__global__ void __launch_bounds__(64) ctc_loss_grad_kernel_v5(CtcParams p) {
  constexpr int PG = 4;          // lattice rows per prefetch group (phase 1)
  constexpr int EMR = 3 * PG;    // circular lattice buffer: the group in use + two groups (8..11 rows) in flight
  extern __shared__ float smem[];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.C, T = p.T, Lmax = p.Lmax, RS = p.RS;
  const int blank = C - 1;
  // smem carve-up (floats): [labs Lmax4][cstart C+1 -> 4][clist Lmax4][bcast 4] then per warp
  //   phase 0: X double buffer (2 x TC x C)
  //   phase 1: X (TC x C) | E circular (EMR x RS) | OCC (TC x LP: label occupancies in class-sorted order, slot 0 = 0)
  //            | BL (TC x NLP: one blank partial sum per lane)
  const int Lmax4 = (Lmax + 3) & ~3, C4 = (C + 1 + 3) & ~3;
  int* labs = reinterpret_cast<int*>(smem);
  int* cstart = labs + Lmax4;
  int* clist = cstart + C4;              // after the set-up: rank[k] = position of label k in the class-sorted order
  float* segf = smem + 2 * Lmax4 + C4;   // segf[m] = 0 where sorted position m starts a class, else 1 (segmented running sum)
  float* bcast = segf + Lmax4;
  const int xc = TC * C;                                   // multiple of 4 (TC is)
  const int LP = ctc5_lp(Lmax), NLP = ctc5_nlp(Lmax);      // odd pitches: conflict-free lane-per-row passes
  const int per_warp = ctc5_per_warp(C, Lmax, RS, TC);
  float* Xs = bcast + 4 + warp * per_warp;
  float* Es = Xs + xc;
  float* Oc = Es + EMR * RS;
  float* Bl = Oc + TC * LP;

  const float* xb = p.x + (size_t)b * T * C;
  float* gb = p.grad ? p.grad + (size_t)b * T * C : nullptr;
  const int32_t* lab_g = p.labels + (size_t)b * Lmax;
  const int Lraw = p.label_len[b];
  const int Tn = p.input_len[b];

  // ---- validation (TF CTCLossOp order) + label rule: a label >= C-1 terminates the sequence
  int st = GR_CTC_OK;
  int L = 0;
  if (Tn < 1 || Tn > T - p.drop) st = GR_CTC_BAD_INPUT_LENGTH;
  else if (Lraw <= 0) st = GR_CTC_ZERO_LABELS;
  else {
    int first_null = Lraw, last_nonnull = -1, bad = 0;
    const int Lr = min(Lraw, Lmax);
    for (int k = lane; k < Lr; k += 32) {
      int v = lab_g[k];
      if (v >= blank) first_null = min(first_null, k);
      else { last_nonnull = max(last_nonnull, k); if (v < 0) bad = 1; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      first_null = min(first_null, __shfl_xor_sync(0xffffffffu, first_null, o));
      last_nonnull = max(last_nonnull, __shfl_xor_sync(0xffffffffu, last_nonnull, o));
      bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    L = min(first_null, Lr);
    if (last_nonnull > first_null || bad || Lraw > Lmax) st = GR_CTC_NONNULL_AFTER_NULL;
    else if (Lraw > Tn) st = GR_CTC_NOT_ENOUGH_TIME;
  }
  if (st != GR_CTC_OK) {  // block-uniform
    if (threadIdx.x == 0) {
      if (p.status) p.status[b] = st;
      p.loss[b] = __int_as_float(0x7fc00000);
    }
    if (gb) for (int e = threadIdx.x; e < T * C; e += 64) gb[e] = 0.f;
    return;
  }
  for (int k = threadIdx.x; k < L; k += 64) labs[k] = lab_g[k];
  if (gb) {  // gradient rows outside [drop, drop+Tn) are zero
    for (int e = threadIdx.x; e < p.drop * C; e += 64) gb[e] = 0.f;
    for (int e = (p.drop + Tn) * C + threadIdx.x; e < T * C; e += 64) gb[e] = 0.f;
  }
  __syncthreads();
  if (gb) {  // class-sorted list of label positions: clist[cstart[c] .. cstart[c+1]) = {k : labs[k] == c}, ascending
    for (int c = threadIdx.x; c < C; c += 64) {
      int n = 0;
      for (int k = 0; k < L; ++k) n += labs[k] == c;
      cstart[c + 1] = n;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      cstart[0] = 0;
      for (int c = 0; c < C; ++c) cstart[c + 1] += cstart[c];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 64) {   // rank[k] = position of label k when the labels are sorted by class
      int m = cstart[c];
      for (int k = 0; k < L; ++k)
        if (labs[k] == c) clist[k] = m++;
      for (int q = cstart[c]; q < cstart[c + 1]; ++q) segf[q] = q == cstart[c] ? 0.f : 1.f;
    }
    __syncthreads();
  }

  const int dir = warp;  // 0: alpha, natural order; 1: gamma, reversed time and labels
  const int tstar = Tn >> 1;
  float* wsb = p.ws + (size_t)b * p.ws_seq_floats;
  // alpha rows t<=t* live at row t; gamma rows t>=t* live at row t+1 (disjoint)
  float* my_rows = wsb + (dir == 0 ? 0 : RS);
  const float* other_rows = wsb + (dir == 0 ? RS : 0);
  // row format (v5): [b_0, l_0, b_1, l_1, ...] in the PRODUCER's own state order (lane j's 2K states are contiguous:
  // one vector store per lane and step); the last two floats of the LAST row of each producer chunk carry that chunk's
  // (offset hi, offset lo) -- the consumer walks a chunk from its last row down, so that is the row where it needs them
  const int OFFW = RS - 2;

  // per-lane state description
  float sb[K], sl[K];
  int labr[K], posb[K], posl[K];
  bool vb[K], vl[K], skip[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int k = lane * K + j;
    vb[j] = k <= L;
    vl[j] = k < L;
    const int nk_l = dir == 0 ? k : L - 1 - k;
    labr[j] = vl[j] ? labs[nk_l] : 0;
    int prev = 0;
    if (vl[j] && k >= 1) prev = labs[dir == 0 ? k - 1 : L - k];
    skip[j] = vl[j] && k >= 1 && labr[j] != prev;
    // the same natural position in the OTHER warp's row: its state index is L - k (blank) / L - 1 - k (label)
    posb[j] = 2 * (L - k);
    posl[j] = 2 * (L - 1 - k) + 1;
    sb[j] = kNeg;
    sl[j] = kNeg;
  }
  // Slots behind the last state (k > L blanks, k >= L labels) are NOT pinned to "log zero": they take mass from blank L
  // and from each other but nothing ever flows back into a valid state (every transition goes up in k), they are never
  // stored, and the re-centring and the meeting row leave them out.  That saves the per-step masking of the emissions.
  const float up_scale = p.upstream ? p.upstream[b] : 1.0f;
  const float eps = p.eps;
  const bool vec2 = ((C & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 7) == 0);
  const bool gvec2 = ((C & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.grad) & 7) == 0);

  double off = 0.0;       // states are kept relative to this running offset, re-centred once per chunk
  double logp2 = 0.0;
  bool novalid = false;
  float lpb_last = kNeg, lpl_last[K];
#pragma unroll
  for (int j = 0; j < K; ++j) lpl_last[j] = kNeg;

  auto stage_x = [&](float* dx, int tlo, int n) {
    const float* src = xb + (size_t)(p.drop + tlo) * C;
    const int tot = n * C;
    if (vec2) {
      for (int e = 2 * lane; e < tot; e += 64)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dx + e)), "l"(src + e) : "memory");
    } else {
      for (int e = lane; e < tot; e += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dx + e)), "l"(src + e) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // ---- per-row pre-pass: lp2 = log2 q, q = softmax(log(p + eps)) = (p+eps)/sum(p+eps); returns Z of the lane's row
  auto pre_pass = [&](float* Xc, int n) -> float {
    float Z = 1.f;
    if (lane < n) {
      float* row = Xc + lane * C;
      Z = 0.f;
      if (p.is_logits) {
        float m = row[0];
        for (int c = 1; c < C; ++c) m = fmaxf(m, row[c]);
        float s_ = 0.f;
        for (int c = 0; c < C; ++c) { float e_ = ex2_approx((row[c] - m) * kLog2e); row[c] = e_; s_ += e_; }
        const float inv = 1.0f / s_;
        for (int c = 0; c < C; ++c) { float pe = row[c] * inv + eps; row[c] = pe; Z += pe; }
      } else {
        for (int c = 0; c < C; ++c) { float pe = row[c] + eps; row[c] = pe; Z += pe; }
      }
      const float lz = lg2_approx(Z);
      for (int c = 0; c < C; ++c) row[c] = fmaxf(lg2_approx(row[c]) - lz, kNeg);
    }
    return Z;
  };
  auto recentre = [&]() {
    float mx = kNeg;
#pragma unroll
    for (int j = 0; j < K; ++j) mx = fmaxf(mx, fmaxf(vb[j] ? sb[j] : kNeg, vl[j] ? sl[j] : kNeg));
    mx = warp_max(mx);
    if (mx > -1.0e29f) {
#pragma unroll
      for (int j = 0; j < K; ++j) { sb[j] = fmaxf(sb[j] - mx, kNeg); sl[j] = fmaxf(sl[j] - mx, kNeg); }   // (junk slots stay finite)
      off += (double)mx;
    }
  };
  // one step of the recursion
  auto recur = [&](float lpb, const float (&lpl)[K]) {
    float upv = __shfl_up_sync(0xffffffffu, sl[K - 1], 1);
    if (lane == 0) upv = kNeg;
    float nb[K], nl[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const float prevl = (j == 0) ? upv : sl[j - 1];
      const float tb = lse2(sb[j], prevl);        // = the blank update; shared by the skip transition
      nb[j] = tb + lpb;
      nl[j] = lse2(sl[j], skip[j] ? tb : sb[j]) + lpl[j];
    }
#pragma unroll
    for (int j = 0; j < K; ++j) { sb[j] = nb[j]; sl[j] = nl[j]; }
  };

  // =========================== phase 0: own half, lattice rows to memory ===========================
  // invalid state slots read the other warp's row at an unused word
  const int padpos = OFFW;
  int gposb[K], gposl[K];
  uint32_t xoffl[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    gposb[j] = vb[j] ? posb[j] : padpos;
    gposl[j] = vl[j] ? posl[j] : padpos;
    xoffl[j] = 4u * (uint32_t)labr[j];
  }
  const bool st_lane = vb[0];                      // this lane holds at least one valid state: it stores its 2K floats
  const uint32_t xoffb = 4u * (uint32_t)blank;
  {
    const int i_end = dir == 0 ? tstar + 1 : Tn - tstar;
    const int nchunks = (i_end + TC - 1) / TC;
    auto chunk_range = [&](int ci, int& ic, int& ie, int& tlo) {
      ic = ci * TC;
      ie = min(ic + TC, i_end);
      tlo = dir == 0 ? ic : Tn - ie;
    };
    { int ic, ie, tlo; chunk_range(0, ic, ie, tlo); stage_x(Xs, tlo, ie - ic); }
    for (int ci = 0; ci < nchunks; ++ci) {
      int ic, ie, tlo;
      chunk_range(ci, ic, ie, tlo);
      const int n = ie - ic;
      float* Xc = Xs + (ci & 1) * xc;
      if (ci + 1 < nchunks) {
        int ic2, ie2, tlo2;
        chunk_range(ci + 1, ic2, ie2, tlo2);
        stage_x(Xs + ((ci + 1) & 1) * xc, tlo2, ie2 - ic2);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncwarp();
      if (ic > 0) recentre();
      const float2 offw = make_float2((float)off, (float)(off - (double)(float)off));
      pre_pass(Xc, n);
      __syncwarp();
      const int rstep = dir == 0 ? 4 * C : -4 * C;
      uint32_t xrow = (uint32_t)__cvta_generic_to_shared(Xc + (dir == 0 ? 0 : n - 1) * C);
      float* dst = my_rows + (size_t)(dir == 0 ? ic : Tn - 1 - ic) * RS + lane * 2 * K;
      const int dstep = dir == 0 ? RS : -RS;
      float lpb = kNeg, lpl[K];
      auto load_lp = [&]() {
        lpb = lds_f32(xrow + xoffb);
#pragma unroll
        for (int j = 0; j < K; ++j) lpl[j] = lds_f32(xrow + xoffl[j]);
      };
      auto emit = [&]() {
        if (st_lane) {
          if constexpr (K % 2 == 0) {
#pragma unroll
            for (int j = 0; j < K; j += 2) *reinterpret_cast<float4*>(dst + 2 * j) = make_float4(sb[j], sl[j], sb[j + 1], sl[j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < K; ++j) *reinterpret_cast<float2*>(dst + 2 * j) = make_float2(sb[j], sl[j]);
          }
        }
        xrow += rstep;
        dst += dstep;
      };
      int i = ic;
      if (ic == 0) {   // initial state of the recursion
        load_lp();
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const int k = lane * K + j;
          sb[j] = (k == 0) ? lpb : kNeg;
          sl[j] = (k == 0 && vl[j]) ? lpl[j] : kNeg;
        }
        emit();
        ++i;
      }
      for (; i < ie; ++i) {
        load_lp();
        recur(lpb, lpl);
        emit();
      }
      lpb_last = lpb;
#pragma unroll
      for (int j = 0; j < K; ++j) lpl_last[j] = lpl[j];
      __syncwarp();   // every lane is done with Xc before the next-but-one chunk is staged over it; and the vector stores
      //                 of the chunk's last row are ordered before the offset words that share its tail
      if (lane == 0) *reinterpret_cast<float2*>(dst - dstep + OFFW) = offw;
    }
  }
  // =========================== meeting row: log p ===========================
  __syncthreads();  // both half-lattices are in memory
  if (dir == 0) {
    const float* orow = other_rows + (size_t)tstar * RS;
    float v[2 * K];
    float m = kNeg;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      v[2 * j] = vb[j] ? sb[j] + orow[gposb[j]] - lpb_last : kNeg;
      v[2 * j + 1] = vl[j] ? sl[j] + orow[gposl[j]] - lpl_last[j] : kNeg;
      m = fmaxf(m, fmaxf(v[2 * j], v[2 * j + 1]));
    }
    m = warp_max(m);
    float s_ = 0.f;
#pragma unroll
    for (int j = 0; j < 2 * K; ++j) s_ += ex2_approx(v[j] - m);
    s_ = warp_sum(s_);
    const float lp2 = m + lg2_approx(s_);
    if (lane == 0) {
      const bool nv = !(m > -1.0e29f) || !(lp2 > -1.0e29f);
      double* bd = reinterpret_cast<double*>(bcast);
      bd[0] = nv ? -1.0e300 : off + (double)orow[OFFW] + (double)orow[OFFW + 1] + (double)lp2;
    }
  }
  __syncthreads();
  logp2 = reinterpret_cast<const double*>(bcast)[0];
  novalid = !(logp2 > -1.0e299);
  if (threadIdx.x == 0) {
    p.loss[b] = novalid ? __int_as_float(0x7f800000) : (float)(-logp2 * 0.6931471805599453);
    if (p.status) p.status[b] = novalid ? GR_CTC_NO_VALID_PATH : GR_CTC_OK;
  }
  if (p.grad == nullptr) return;

  // =========================== phase 1: the other half, occupancies -> gradient ===========================
  {
    // dir 0 re-uses its state at row t* for the first row (no recursion step); dir 1 starts one row further
    const int i_begin = dir == 0 ? tstar : Tn - tstar;
    const int i_end = Tn;
    const int nrows = i_end - i_begin;
    const int nchunks = (nrows + TC - 1) / TC;
    const uint32_t e_base = (uint32_t)__cvta_generic_to_shared(Es);
    const uint32_t e_pitch = (uint32_t)RS * 4u;
    uint32_t eoffb[K], eoffl[K], ooffl[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      eoffb[j] = 4u * (uint32_t)gposb[j];
      eoffl[j] = 4u * (uint32_t)gposl[j];
      const int k = lane * K + j;
      ooffl[j] = 4u * (uint32_t)(vl[j] ? 1 + clist[dir == 0 ? k : L - 1 - k] : 0);   // class-sorted slot of the label state
    }
    // prefetch stream: the other warp's half-lattice in groups of PG rows, one 16-byte copy per lane and row; group g
    // lives in ring rows PG*(g % 3) ..; two groups (8 to 11 rows) are in flight ahead of the consumer
    const bool pf_lane = 4 * lane < RS;
    const bool wide_rows = RS > 128;
    const float* pf_src = other_rows + (size_t)(dir == 0 ? i_begin : Tn - 1 - i_begin) * RS + 4 * lane;
    const ptrdiff_t pf_step = dir == 0 ? RS : -RS;
    int pf_row = 0, pf_slot = 0;
    auto fetch_group = [&]() {
      uint32_t d = e_base + (uint32_t)(pf_slot * PG) * e_pitch + 16u * (uint32_t)lane;
#pragma unroll
      for (int u = 0; u < PG; ++u) {
        if (pf_lane && pf_row + u < nrows) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(pf_src) : "memory");
          if (wide_rows) {   // rows wider than 32 x 16 bytes (Lmax > 62)
#pragma unroll 1
            for (int o = 128 + 4 * lane; o < RS; o += 128)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 4u * (uint32_t)o - 16u * (uint32_t)lane), "l"(pf_src + o - 4 * lane) : "memory");
          }
        }
        d += e_pitch;
        pf_src += pf_step;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      pf_row += PG;
      pf_slot = pf_slot == 2 ? 0 : pf_slot + 1;
    };
    fetch_group();
    fetch_group();
    // offsets of the producer chunk the first consumed row belongs to: in the last row of that chunk
    float2 oo;
    {
      const int ipf = Tn - 1 - i_begin;                                  // producer index of the first consumed row
      const int ipl = min(ipf | (TC - 1), dir == 0 ? Tn - tstar - 1 : tstar);
      oo = *reinterpret_cast<const float2*>(other_rows + (size_t)(dir == 0 ? Tn - 1 - ipl : ipl) * RS + OFFW);
    }
    int r = 0;                 // rows consumed so far
    uint32_t erow = e_base;
    int eslot = 0;             // ring group of the row in use
    const uint32_t oc_base = (uint32_t)__cvta_generic_to_shared(Oc), bl_base = (uint32_t)__cvta_generic_to_shared(Bl) + 4u * (uint32_t)lane;
    const uint32_t oc_pitch = 4u * (uint32_t)LP, bl_pitch = 4u * (uint32_t)NLP;
    const bool bl_lane = vb[0];
    for (int ci = 0; ci < nchunks; ++ci) {
      const int ic = i_begin + ci * TC;
      const int ie = min(ic + TC, i_end);
      const int n = ie - ic;
      const int tlo = dir == 0 ? ic : Tn - ie;
      stage_x(Xs, tlo, n);
      if (ci + 1 < nchunks) {   // the single X buffer of this phase cannot be staged ahead: pull the next chunk into L2 at least
        const int ic2 = ic + TC, ie2 = min(ic2 + TC, i_end);
        const float* nx = xb + (size_t)(p.drop + (dir == 0 ? ic2 : Tn - ie2)) * C;
        if (32 * lane < (ie2 - ic2) * C) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 32 * lane) : "memory");
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      recentre();
      const double D0 = novalid ? -1.0e30 : off - logp2;     // no valid path: every occupancy becomes ex2(-huge) = 0
      const float Zrow = pre_pass(Xs, n);
      __syncwarp();
      const int rstep = dir == 0 ? 4 * C : -4 * C;
      uint32_t xrow = (uint32_t)__cvta_generic_to_shared(Xs + (dir == 0 ? 0 : n - 1) * C);
      uint32_t ocrow = oc_base, blrow = bl_base;
      float lpb, lpl[K];
      auto load_lp = [&]() {
        lpb = lds_f32(xrow + xoffb);
#pragma unroll
        for (int j = 0; j < K; ++j) lpl[j] = lds_f32(xrow + xoffl[j]);
      };
      // the other warp's row is read BEFORE the recursion step so that its latency (and the offset arithmetic)
      // overlaps the MUFU chain of the step.  The row constant changes where the PRODUCER re-centred, i.e. at its
      // chunk boundaries: producer index Tn-1-i, so whenever (Tn - i) is a multiple of TC -- and at the start of
      // each of this warp's chunks (D0 moved)
      float ev_b[K], ev_l[K];
      float cst = kNeg;
      auto emit_load = [&](int i) {
        if ((r & (PG - 1)) == 0) {
          fetch_group();
          asm volatile("cp.async.wait_group 2;" ::: "memory");
          __syncwarp();
          erow = e_base + (uint32_t)(eslot * PG) * e_pitch;
          eslot = eslot == 2 ? 0 : eslot + 1;
        }
#pragma unroll
        for (int j = 0; j < K; ++j) { ev_b[j] = lds_f32(erow + eoffb[j]); ev_l[j] = lds_f32(erow + eoffl[j]); }
        const bool pb = ((Tn - i) & (TC - 1)) == 0;        // warp-uniform: first row of a producer chunk on this walk
        if (pb) oo = lds_f32x2(erow + 4u * (uint32_t)OFFW);
        if (pb || i == ic) cst = (float)(D0 + (double)oo.x + (double)oo.y);
      };
      // occupancy x emission probability: ex2(alpha + gamma - log p); the division by the emission probability is done
      // once per CLASS in the post-pass (all states of a class share it).  Labels go to their class-sorted slot of the
      // row's OCC line, the lane's blanks are added up first.
      auto emit = [&]() {
        float ob = 0.f;
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const float o = ex2_approx((sb[j] + ev_b[j]) + cst);
          ob += vb[j] ? o : 0.f;
          sts_f32_if(ocrow + ooffl[j], ex2_approx((sl[j] + ev_l[j]) + cst), vl[j]);
        }
        sts_f32_if(blrow, ob, bl_lane);
        xrow += rstep;
        erow += e_pitch;
        ocrow += oc_pitch;
        blrow += bl_pitch;
        ++r;
      };
      int i = ic;
      if (dir == 0 && ci == 0) {   // row t*: the state is already there
        emit_load(i);
        load_lp();
        emit();
        ++i;
      }
      for (; i < ie; ++i) {
        emit_load(i);
        load_lp();
        recur(lpb, lpl);
        emit();
      }
      __syncwarp();
      // ---- per-row post-pass: occupancies -> gradient, in place over X
      if (lane < n) {
        // lane = row of the chunk (natural time order)
        float* xr = Xs + lane * C;
        const int pidx = dir == 0 ? lane : n - 1 - lane;          // processing index of that row
        float* oc = Oc + pidx * LP;
        const float* bl = Bl + pidx * NLP;
        float occb = 0.f;
        const int nbl = L / K + 1;                               // lanes holding a valid blank state
        for (int l = 0; l < nbl; ++l) occb += bl[l];
        // running sums over the class-sorted label occupancies, restarted where a class starts (no differences of
        // prefix sums: the occupancies of one row span many orders of magnitude): the class total is the running sum at
        // the last label of the class, slot cstart[c+1]
        float run = 0.f;
        for (int m = 1; m <= L; ++m) { run = fmaf(run, segf[m - 1], oc[m]); oc[m] = run; }
        const float Z = Zrow;
        int cs0 = 0;
        if (p.is_logits) {
          // gz_c = up (q - occ_c); grad_c = w_c gz_c - p_c sum_c' w_c' gz_c' with w = p / (p + eps): two sweeps over the
          // classes, the class occupancies are cheap prefix differences both times
          const float rz = 1.0f / Z;
          float dot = 0.f;
          for (int c = 0; c < C; ++c) {
            float oc_c = occb;
            if (c != blank) { const int cs1 = cstart[c + 1]; oc_c = cs1 > cs0 ? oc[cs1] : 0.f; cs0 = cs1; }
            const float x = xr[c];
            const float q = ex2_approx(x), r_ = ex2_approx(-x);
            const float pr = fmaxf(q * Z - eps, 0.f);   // p
            dot += pr * (r_ * rz) * (up_scale * (q - oc_c * r_));
          }
          cs0 = 0;
          for (int c = 0; c < C; ++c) {
            float oc_c = occb;
            if (c != blank) { const int cs1 = cstart[c + 1]; oc_c = cs1 > cs0 ? oc[cs1] : 0.f; cs0 = cs1; }
            const float x = xr[c];
            const float q = ex2_approx(x), r_ = ex2_approx(-x);
            const float pr = fmaxf(q * Z - eps, 0.f);
            xr[c] = pr * (r_ * rz) * (up_scale * (q - oc_c * r_)) - pr * dot;
          }
        } else {
          const float uz = up_scale / Z;
          for (int c = 0; c < C; ++c) {
            float oc_c = occb;
            if (c != blank) { const int cs1 = cstart[c + 1]; oc_c = cs1 > cs0 ? oc[cs1] : 0.f; cs0 = cs1; }
            const float r_ = ex2_approx(-xr[c]);      // 1 / q
            xr[c] = uz * (1.0f - oc_c * (r_ * r_));
          }
        }
      }
      __syncwarp();
      float* gdst = gb + (size_t)(p.drop + tlo) * C;
      const int tot = n * C;
      if (gvec2) {
        for (int e = 2 * lane; e < tot; e += 64) *reinterpret_cast<float2*>(gdst + e) = *reinterpret_cast<const float2*>(Xs + e);
      } else {
        for (int e = lane; e < tot; e += 32) gdst[e] = Xs[e];
      }
      __syncwarp();
    }
  }
}


static size_t ctc_smem_bytes_v4(int C, int Lmax, int RS, int TC) {
  const int Lmax4 = (Lmax + 3) & ~3, C4 = (C + 1 + 3) & ~3;
  const size_t xc = (size_t)TC * C;
  const size_t per_warp = std::max(2 * xc, xc + (size_t)(TC + 4) * RS);
  return (2 * (size_t)Lmax4 + C4 + 4 + 2 * per_warp) * sizeof(float);
}

static int ctc5_row_stride(int Lmax) {   // 2 floats per state pair of every lane that can hold a valid state (+ room for the offset words)
  const int K = (Lmax + 1 + 31) / 32;
  const int nl = (Lmax + 1 + K - 1) / K;
  int rs = (nl * 2 * K + 3) & ~3;
  if (rs - 2 <= 2 * Lmax) rs += 4;      // the two offset words sit behind the last valid state (index 2 Lmax)
  return rs;
}

static size_t ctc_smem_bytes_v5(int C, int Lmax, int RS, int TC) {
  const int Lmax4 = (Lmax + 3) & ~3, C4 = (C + 1 + 3) & ~3;
  return (3 * (size_t)Lmax4 + C4 + 4 + 2 * (size_t)ctc5_per_warp(C, Lmax, RS, TC)) * sizeof(float);
}

template <int K>
static int launch_ctc(const CtcParams& p, cudaStream_t stream, bool v4) {
  // the largest chunk that still lets the whole batch be resident in one wave
  const int sms = num_sms();
  const int need_per_sm = (p.B + sms - 1) / sms;
  const size_t smem_budget = 227 * 1024;
  int TC = 32;
  while (TC > 8) {
    size_t per_cta = (v4 ? ctc_smem_bytes_v4(p.C, p.Lmax, p.RS, TC) : ctc_smem_bytes_v5(p.C, p.Lmax, p.RS, TC)) + 1024;
    if (per_cta * need_per_sm <= smem_budget && per_cta <= 200 * 1024) break;
    TC >>= 1;
  }
  if (const char* tcs = getenv("GR_CTC_TC")) {   // tests: force a smaller chunk than the batch size would pick
    const int v = atoi(tcs);
    if ((v == 8 || v == 16) && v < TC) TC = v;
  }
  const size_t smem = v4 ? ctc_smem_bytes_v4(p.C, p.Lmax, p.RS, TC) : ctc_smem_bytes_v5(p.C, p.Lmax, p.RS, TC);
  if (smem > 220 * 1024) return set_error(GR_EUNSUPPORTED, "ctc: C/Lmax too large for shared memory");
  auto go = [&](auto kern) -> int {
    GR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<p.B, 64, smem, stream>>>(p);
    GR_CHECK_LAUNCH("ctc_loss_grad_kernel");
    return GR_OK;
  };
  if (v4) {
    if (TC == 32) return go(ctc_loss_grad_kernel_v4<K, 32>);
    if (TC == 16) return go(ctc_loss_grad_kernel_v4<K, 16>);
    return go(ctc_loss_grad_kernel_v4<K, 8>);
  }
  if (TC == 32) return go(ctc_loss_grad_kernel_v5<K, 32>);
  if (TC == 16) return go(ctc_loss_grad_kernel_v5<K, 16>);
  return go(ctc_loss_grad_kernel_v5<K, 8>);
}

}  // namespace gr

extern "C" int gr_ctc_workspace_bytes(int B, int T, int C, int Lmax, size_t* bytes_out) {
  if (B <= 0 || T <= 0 || C < 2 || Lmax <= 0 || !bytes_out) return gr::set_error(GR_EINVAL, "ctc_workspace_bytes: bad argument");
  const size_t RS = std::max(gr::ctc_row_stride(Lmax), gr::ctc5_row_stride(Lmax));   // either kernel (GR_CTC_IMPL)
  *bytes_out = (size_t)B * (size_t)(T + 2) * RS * sizeof(float);
  return GR_OK;
}

extern "C" int gr_ctc_loss_grad_f32(const float* x, int input_is_logits, int B, int T, int C,
                                    int drop_frames, float eps, const int32_t* labels, int Lmax,
                                    const int32_t* label_len, const int32_t* input_len,
                                    const float* upstream, float* loss, float* grad_out,
                                    int32_t* status, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  using namespace gr;
  if (!x || !labels || !label_len || !input_len || !loss || !workspace)
    return set_error(GR_EINVAL, "ctc_loss_grad: null pointer");
  if (B <= 0 || T <= 0 || C < 2 || Lmax <= 0 || drop_frames < 0 || drop_frames >= T)
    return set_error(GR_EINVAL, "ctc_loss_grad: bad shape");
  if (Lmax > 255) return set_error(GR_EUNSUPPORTED, "ctc_loss_grad: Lmax > 255");
  if (reinterpret_cast<uintptr_t>(workspace) & 15) return set_error(GR_EINVAL, "ctc_loss_grad: workspace must be 16-byte aligned");
  size_t need = 0;
  gr_ctc_workspace_bytes(B, T, C, Lmax, &need);
  if (workspace_bytes < need) return set_error(GR_EWORKSPACE, "ctc_loss_grad: workspace too small");
  CtcParams p;
  p.x = x; p.is_logits = input_is_logits; p.B = B; p.T = T; p.C = C; p.drop = drop_frames;
  p.eps = eps; p.labels = labels; p.Lmax = Lmax; p.label_len = label_len; p.input_len = input_len;
  p.upstream = upstream; p.loss = loss; p.grad = grad_out; p.status = status;
  p.ws = static_cast<float*>(workspace);
  const char* impl = getenv("GR_CTC_IMPL");
  const bool v4 = impl && impl[0] == 'v' && impl[1] == '4';
  p.RS = v4 ? ctc_row_stride(Lmax) : ctc5_row_stride(Lmax);
  p.ws_seq_floats = (size_t)(T + 2) * p.RS;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int K = (Lmax + 1 + 31) / 32;
  // GR_CTC_IMPL=v4: the round-1 kernel (natural-order lattice rows), kept as a cross-check of v5
  switch (K) {
    case 1: return launch_ctc<1>(p, s, v4);
    case 2: return launch_ctc<2>(p, s, v4);
    case 3: return launch_ctc<3>(p, s, v4);
    case 4: return launch_ctc<4>(p, s, v4);
    case 5: return launch_ctc<5>(p, s, v4);
    case 6: return launch_ctc<6>(p, s, v4);
    default: return launch_ctc<8>(p, s, v4);
  }
}
