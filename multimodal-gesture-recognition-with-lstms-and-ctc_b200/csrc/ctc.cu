// CTC loss + gradient for sm_100a.  Replaces ctc_lambda_func -> K.ctc_batch_cost -> tf.nn.ctc_loss
// (/root/reference/audio_network/losses.py:4-15; TF CTCLossCalculator semantics).
//
// Design (DESIGN.md "K6"): one CTA of two warps per sequence.  Warp 0 runs the alpha recursion
// forward in time; warp 1 runs the SAME recursion on the time-reversed, label-reversed problem,
// which is the beta recursion with the emission folded in (gamma = beta + log y).  They meet at
// t* = Tn/2: log p = logsumexp_u(alpha + gamma - log y)(t*).  Each warp then keeps going through
// the other half, where the other warp's stored lattice rows turn every step directly into
// posterior occupancies, so the softmax/CTC gradient is produced on the fly and only HALF of the
// (alpha, beta) lattice ever touches memory.  Log space (base 2, MUFU ex2/lg2), extended labels
// split into blank states b_k (k = 0..L) and label states l_k (k = 0..L-1), K consecutive k per
// lane, one shuffle per step.  Probabilities are streamed in chunks of TC frames through shared
// memory with coalesced loads; gradients leave the same way.
#include "common.cuh"

namespace gr {

static constexpr float kNeg = -1.0e30f;  // "log zero": finite so that no inf-inf NaN can arise

__device__ __forceinline__ float lse2(float a, float b) {
  float hi = fmaxf(a, b), lo = fminf(a, b);
  return hi + lg2_approx(1.0f + ex2_approx(lo - hi));
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

template <int K, int TC>
__global__ void __launch_bounds__(64) ctc_loss_grad_kernel(CtcParams p) {
  extern __shared__ float smem[];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.C, T = p.T, Lmax = p.Lmax, RS = p.RS;
  const int blank = C - 1;
  // smem carve-up: [labs Lmax ints][logp broadcast 4 floats] then per warp: Xs, Os, Es
  int* labs = reinterpret_cast<int*>(smem);
  float* bcast = smem + ((Lmax + 3) & ~3);
  float* wbase = bcast + 4 + warp * (3 * TC * C + 2 * TC * RS);
  float* Xs = wbase;                 // 2 buffers x TC rows x C
  float* Os = Xs + 2 * TC * C;       // TC x C
  float* Es = Os + TC * C;           // 2 buffers x TC rows x RS (RS odd: conflict-free lane-per-row)

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
  // gradient rows outside [drop, drop+Tn) are zero
  if (gb) {
    for (int e = threadIdx.x; e < p.drop * C; e += 64) gb[e] = 0.f;
    for (int e = (p.drop + Tn) * C + threadIdx.x; e < T * C; e += 64) gb[e] = 0.f;
  }
  __syncthreads();

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

  // Renormalisation: states are kept relative to a running offset `off` (double), re-centred on
  // the warp maximum once per chunk, so fp32 never has to resolve 1e-7 at magnitude 1e3..1e4.
  double off = 0.0;
  double logp2 = 0.0;
  bool novalid = false;
  float lpb_last = kNeg, lpl_last[K];
#pragma unroll
  for (int j = 0; j < K; ++j) lpl_last[j] = kNeg;

  // shared-memory staging: two buffers per warp, filled by cp.async one chunk ahead
  auto stage_chunk = [&](int buf, int tlo, int n, bool with_lattice) {
    const float* src = xb + (size_t)(p.drop + tlo) * C;
    float* dx = Xs + buf * (TC * C);
    const int tot = n * C;
    for (int e = lane; e < tot; e += 32)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dx + e)), "l"(src + e) : "memory");
    if (with_lattice) {
      const float* lsrc = other_rows + (size_t)tlo * RS;
      float* de = Es + buf * (TC * RS);
      const int ltot = n * RS;
      for (int e = lane; e < ltot; e += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(de + e)), "l"(lsrc + e) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  for (int phase = 0; phase < 2; ++phase) {
    if (phase == 1 && p.grad == nullptr) break;
    // processing-order step range of this phase
    int i_begin, i_end;
    bool hold_first = false;  // first step of the phase re-uses the current state (row t*)
    if (phase == 0) { i_begin = 0; i_end = dir == 0 ? tstar + 1 : Tn - tstar; }
    else if (dir == 0) { i_begin = tstar; i_end = Tn; hold_first = true; }
    else { i_begin = Tn - tstar; i_end = Tn; }
    const bool lat = phase == 1;
    const int nchunks = (i_end - i_begin + TC - 1) / TC;
    auto chunk_range = [&](int ci, int& ic, int& ie, int& tlo) {
      ic = i_begin + ci * TC;
      ie = min(ic + TC, i_end);
      tlo = dir == 0 ? ic : Tn - ie;
    };
    if (nchunks > 0) { int ic, ie, tlo; chunk_range(0, ic, ie, tlo); stage_chunk(0, tlo, ie - ic, lat); }

    for (int ci = 0; ci < nchunks; ++ci) {
      int ic, ie, tlo;
      chunk_range(ci, ic, ie, tlo);
      const int n = ie - ic;
      const int buf = ci & 1;
      if (ci + 1 < nchunks) {
        int ic2, ie2, tlo2;
        chunk_range(ci + 1, ic2, ie2, tlo2);
        stage_chunk(buf ^ 1, tlo2, ie2 - ic2, lat);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncwarp();
      float* Xc = Xs + buf * (TC * C);
      float* Ec = Es + buf * (TC * RS);
      // ---- re-centre the states on their maximum (once per chunk; uniform across the warp)
      if (ic > 0) {
        float mx = kNeg;
#pragma unroll
        for (int j = 0; j < K; ++j) mx = fmaxf(mx, fmaxf(sb[j], sl[j]));
        mx = warp_max(mx);
        if (mx > -1.0e29f) {
#pragma unroll
          for (int j = 0; j < K; ++j) { sb[j] -= mx; sl[j] -= mx; }
          off += (double)mx;
        }
      }
      const float off_hi = (float)off, off_lo = (float)(off - (double)off_hi);
      // ---- per-row pre-pass: lp2 = log2 q, q = softmax(log(p + eps)) = (p+eps)/sum(p+eps);
      //      phase 1 also folds (my offset + other offset - log p) into the row's last slot
      float Zrow = 1.f;
      if (lane < n) {
        float* row = Xc + lane * C;
        float Z = 0.f;
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
        Zrow = Z;
        if (lat) {
          float* erow = Ec + lane * RS;
          erow[RS - 1] = novalid ? kNeg : (float)(off + (double)erow[RS - 2] + (double)erow[RS - 1] - logp2);
        }
      }
      __syncwarp();
      // ---- the serial part
      int i = ic;
      if (i == 0) {  // initial state of the recursion (phase 0 only)
        const float* row = Xc + (dir == 0 ? 0 : n - 1) * C;
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const int k = lane * K + j;
          sb[j] = (k == 0) ? row[blank] : kNeg;
          sl[j] = (k == 0 && vl[j]) ? row[labr[j]] : kNeg;
        }
      }
      const int rstep = dir == 0 ? 1 : -1;
      int r = dir == 0 ? 0 : n - 1;
      float* dst = my_rows + (size_t)(dir == 0 ? ic : Tn - 1 - ic) * RS;
      const int dstep = dir == 0 ? RS : -RS;
      for (; i < ie; ++i, r += rstep, dst += dstep) {
        const float* row = Xc + r * C;
        const float lpb = row[blank];
        float lpl[K];
#pragma unroll
        for (int j = 0; j < K; ++j) lpl[j] = vl[j] ? row[labr[j]] : kNeg;
        if (i != 0 && !(hold_first && i == i_begin)) {
          float upv = __shfl_up_sync(0xffffffffu, sl[K - 1], 1);
          if (lane == 0) upv = kNeg;
          float nb[K], nl[K];
#pragma unroll
          for (int j = 0; j < K; ++j) {
            const float prevl = (j == 0) ? upv : sl[j - 1];
            // lse(sl, sb, prevl) = lse(sl, lse(sb, prevl)): the inner term is the blank update anyway
            const float tb = lse2(sb[j], prevl);
            nb[j] = tb + (vb[j] ? lpb : kNeg);
            nl[j] = lse2(sl[j], skip[j] ? tb : sb[j]) + lpl[j];
          }
#pragma unroll
          for (int j = 0; j < K; ++j) { sb[j] = nb[j]; sl[j] = nl[j]; }
        }
        if (!lat) {
#pragma unroll
          for (int j = 0; j < K; ++j) {
            if (vb[j]) dst[posb[j]] = sb[j];
            if (vl[j]) dst[posl[j]] = sl[j];
          }
          if (lane == 0) { dst[RS - 2] = off_hi; dst[RS - 1] = off_lo; }
          if (i == ie - 1) {
            lpb_last = lpb;
#pragma unroll
            for (int j = 0; j < K; ++j) lpl_last[j] = lpl[j];
          }
        } else {
          float* erow = Ec + r * RS;
          const float cst = erow[RS - 1];
#pragma unroll
          for (int j = 0; j < K; ++j) {
            if (vb[j]) erow[posb[j]] = ex2_approx((sb[j] + erow[posb[j]] - lpb) + cst);
            if (vl[j]) erow[posl[j]] = ex2_approx((sl[j] + erow[posl[j]] - lpl[j]) + cst);
          }
        }
      }
      if (lat) {
        __syncwarp();
        // ---- per-row post-pass: occupancies -> gradient
        if (lane < n) {
          const float* row = Xc + lane * C;
          float* orow = Os + lane * C;
          const float* erow = Ec + lane * RS;
          for (int c = 0; c < C; ++c) orow[c] = 0.f;
          float occb = 0.f;
          for (int k = 0; k <= L; ++k) occb += erow[k];
          for (int k = 0; k < L; ++k) orow[labs[k]] += erow[Lmax + 1 + k];
          orow[blank] += occb;
          const float Z = Zrow;
          if (p.is_logits) {
            float dot = 0.f;
            for (int c = 0; c < C; ++c) {
              const float q = ex2_approx(row[c]);
              const float gz = up_scale * (q - orow[c]);
              const float pe = q * Z;                  // p + eps
              const float pr = fmaxf(pe - eps, 0.f);   // p
              const float w = pr / pe;
              orow[c] = w * gz;
              dot += w * gz;
            }
            for (int c = 0; c < C; ++c) {
              const float q = ex2_approx(row[c]);
              const float pr = fmaxf(q * Z - eps, 0.f);
              orow[c] = orow[c] - pr * dot;
            }
          } else {
            for (int c = 0; c < C; ++c) {
              const float q = ex2_approx(row[c]);
              orow[c] = up_scale * __fdividef(q - orow[c], q * Z);
            }
          }
        }
        __syncwarp();
        float* gdst = gb + (size_t)(p.drop + tlo) * C;
        const int tot = n * C;
        for (int e = lane; e < tot; e += 32) gdst[e] = Os[e];
        __syncwarp();
      }
    }
    if (phase == 0) {
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
    }
  }
}

static size_t ctc_smem_bytes(int C, int Lmax, int RS, int TC) {
  size_t fl = ((Lmax + 3) & ~3) + 4 + 2 * (size_t)(3 * TC * C + 2 * TC * RS);
  return fl * sizeof(float);
}
static int ctc_row_stride(int Lmax) { return (2 * Lmax + 3) | 1; }  // +2: per-row offset (hi, lo); odd stride

template <int K>
static int launch_ctc(const CtcParams& p, cudaStream_t stream) {
  // pick the largest chunk that still lets the whole batch be resident in one wave
  const int sms = num_sms();
  const int need_per_sm = (p.B + sms - 1) / sms;
  const size_t smem_budget = 220 * 1024;
  int TC = 32;
  while (TC > 8) {
    size_t per_cta = ctc_smem_bytes(p.C, p.Lmax, p.RS, TC) + 1024;
    if (per_cta * need_per_sm <= smem_budget && per_cta <= 200 * 1024) break;
    TC >>= 1;
  }
  const size_t smem = ctc_smem_bytes(p.C, p.Lmax, p.RS, TC);
  if (smem > 220 * 1024) return set_error(GR_EUNSUPPORTED, "ctc: C/Lmax too large for shared memory");
  auto go = [&](auto kern) -> int {
    GR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<p.B, 64, smem, stream>>>(p);
    GR_CHECK_LAUNCH("ctc_loss_grad_kernel");
    return GR_OK;
  };
  if (TC == 32) return go(ctc_loss_grad_kernel<K, 32>);
  if (TC == 16) return go(ctc_loss_grad_kernel<K, 16>);
  return go(ctc_loss_grad_kernel<K, 8>);
}

}  // namespace gr

extern "C" int gr_ctc_workspace_bytes(int B, int T, int C, int Lmax, size_t* bytes_out) {
  if (B <= 0 || T <= 0 || C < 2 || Lmax <= 0 || !bytes_out) return gr::set_error(GR_EINVAL, "ctc_workspace_bytes: bad argument");
  const size_t RS = gr::ctc_row_stride(Lmax);
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
  size_t need = 0;
  gr_ctc_workspace_bytes(B, T, C, Lmax, &need);
  if (workspace_bytes < need) return set_error(GR_EWORKSPACE, "ctc_loss_grad: workspace too small");
  CtcParams p;
  p.x = x; p.is_logits = input_is_logits; p.B = B; p.T = T; p.C = C; p.drop = drop_frames;
  p.eps = eps; p.labels = labels; p.Lmax = Lmax; p.label_len = label_len; p.input_len = input_len;
  p.upstream = upstream; p.loss = loss; p.grad = grad_out; p.status = status;
  p.ws = static_cast<float*>(workspace);
  p.RS = ctc_row_stride(Lmax);
  p.ws_seq_floats = (size_t)(T + 2) * p.RS;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int K = (Lmax + 1 + 31) / 32;
  switch (K) {
    case 1: return launch_ctc<1>(p, s);
    case 2: return launch_ctc<2>(p, s);
    case 3: return launch_ctc<3>(p, s);
    case 4: return launch_ctc<4>(p, s);
    case 5: return launch_ctc<5>(p, s);
    case 6: return launch_ctc<6>(p, s);
    default: return launch_ctc<8>(p, s);
  }
}
