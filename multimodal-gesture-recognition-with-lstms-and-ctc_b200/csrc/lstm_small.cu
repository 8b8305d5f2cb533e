// Keras-LSTM recurrence for NARROW layers (H <= 128, e.g. the fusion BLSTM(100) of
// /root/reference/multimodal_fusion/multimodal.py:159-168), forward and BPTT, fp32.
//
// When U (H x 4H fp32 <= 256 KB) fits in ONE SM's register file the recurrence needs no
// inter-CTA exchange at all: CTA = (direction, BS <= 4 sequences); thread n keeps column n of U
// (forward) / a quarter of row j of U (backward) in REGISTERS for all T steps, h_{t-1} / dG_{t+1}
// live in shared memory and are read with broadcast LDS.128.  Two __syncthreads per step, no
// grid barrier, no atomics; 2*ceil(B/BS) independent CTAs fill the SMs.
#include "common.cuh"

namespace gr {

__device__ __forceinline__ float hsig_s(float v) { return fminf(fmaxf(0.2f * v + 0.5f, 0.f), 1.f); }
__device__ __forceinline__ float dhsig_s(float s) { return (s > 0.f && s < 1.f) ? 0.2f : 0.f; }
// tanh(x) = 1 - 2/(1 + e^{2x}) on MUFU ex2 + fast division: absolute error ~1e-7 (as lstm_tc.cu)
__device__ __forceinline__ float tanh_s(float x) {
  const float e = ex2_approx(x * 2.8853900817779268f);
  return 1.0f - __fdividef(2.0f, 1.0f + e);
}

struct SmallParams {
  float* gates;        // (B,T,8H)  fwd: P in / gates out (if save);  bwd: gates in / dP out
  const float* U;      // (2,H,4H)
  float* y;            // fwd out (B,T,2H)
  float* cell;         // fwd: out if save;  bwd: in
  const float* dy;     // bwd in (B,T,2H)
  int B, T, H, BS, save;
};

// BS (sequences per CTA) is a template parameter: with a runtime bound the matvec loop kept a branch
// per (k, sequence) and every LDS -> 4 dependent FMAs chain ran serially (6.8 us / step at H = 100).
template <int HP, int BS>
__global__ void __launch_bounds__(4 * HP, 1) lstm_small_fwd_kernel(SmallParams p) {
  extern __shared__ __align__(16) float sm[];
  const int H = p.H, T = p.T, H4 = 4 * H;
  float* hs = sm;                 // BS * HP   (h_{t-1}, zero padded to HP)
  float* zs = hs + BS * HP;       // BS * H4
  const int nbg = (p.B + BS - 1) / BS;
  const int dir = blockIdx.x / nbg, bg = blockIdx.x % nbg;
  const int b0 = bg * BS;
  const int n = threadIdx.x;
  float ureg[HP];
  const float* Ud = p.U + (size_t)dir * H * H4;
#pragma unroll
  for (int k = 0; k < HP; ++k) ureg[k] = (k < H && n < H4) ? Ud[(size_t)k * H4 + n] : 0.f;
  for (int e = threadIdx.x; e < BS * HP; e += blockDim.x) hs[e] = 0.f;
  // element-wise role: thread e < BS*H  <->  (sequence b0+eb, unit ej)
  const int eb = n / H, ej = n - eb * H;
  const bool eact = n < BS * H && (b0 + eb) < p.B;
  float c_state = 0.f;
  const size_t G8 = (size_t)8 * H, Y2 = (size_t)2 * H;
  __syncthreads();
  // pre-activations of step 0; step s+1's are prefetched while step s computes
  float pre[4] = {0.f, 0.f, 0.f, 0.f};
  if (eact) {
    const float* g0 = p.gates + ((size_t)(b0 + eb) * T + (dir == 0 ? 0 : T - 1)) * G8 + (size_t)dir * H4 + ej;
#pragma unroll
    for (int g = 0; g < 4; ++g) pre[g] = g0[(size_t)g * H];
  }
  for (int s = 0; s < T; ++s) {
    const int t = dir == 0 ? s : T - 1 - s;
    float* grow = eact ? p.gates + ((size_t)(b0 + eb) * T + t) * G8 + (size_t)dir * H4 + ej : nullptr;
    float nxt[4] = {0.f, 0.f, 0.f, 0.f};
    if (eact && s + 1 < T) {
      const float* gn = grow + (dir == 0 ? (ptrdiff_t)G8 : -(ptrdiff_t)G8);
#pragma unroll
      for (int g = 0; g < 4; ++g) nxt[g] = gn[(size_t)g * H];
    }
    if (s > 0 && n < H4) {
      // packed fp32x2 FMAs (sm_100 FFMA2): half the FMA issue slots of the scalar loop
      float2 acc[BS][2];
#pragma unroll
      for (int b = 0; b < BS; ++b) acc[b][0] = acc[b][1] = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < HP; k += 4) {
        const float2 u01 = make_float2(ureg[k], ureg[k + 1]), u23 = make_float2(ureg[k + 2], ureg[k + 3]);
#pragma unroll
        for (int b = 0; b < BS; ++b) {
          const float4 h = *reinterpret_cast<const float4*>(hs + b * HP + k);
          acc[b][0] = __ffma2_rn(make_float2(h.x, h.y), u01, acc[b][0]);
          acc[b][1] = __ffma2_rn(make_float2(h.z, h.w), u23, acc[b][1]);
        }
      }
#pragma unroll
      for (int b = 0; b < BS; ++b) zs[b * H4 + n] = (acc[b][0].x + acc[b][0].y) + (acc[b][1].x + acc[b][1].y);
    }
    __syncthreads();
    if (eact) {
      float z[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) z[g] = pre[g] + (s > 0 ? zs[eb * H4 + g * H + ej] : 0.f);
      const float gi = hsig_s(z[0]), gf = hsig_s(z[1]), gg = tanh_s(z[2]), go = hsig_s(z[3]);
      const float c = gf * c_state + gi * gg;
      c_state = c;
      const float h = go * tanh_s(c);
      hs[eb * HP + ej] = h;
      p.y[((size_t)(b0 + eb) * T + t) * Y2 + (size_t)dir * H + ej] = h;
      if (p.save) {
        grow[0] = gi; grow[H] = gf; grow[2 * (size_t)H] = gg; grow[3 * (size_t)H] = go;
        p.cell[((size_t)(b0 + eb) * T + t) * Y2 + (size_t)dir * H + ej] = c;
      }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) pre[g] = nxt[g];
    __syncthreads();
  }
}

template <int HP, int BS>
__global__ void __launch_bounds__(4 * HP, 1) lstm_small_bwd_kernel(SmallParams p) {
  extern __shared__ __align__(16) float sm[];
  const int H = p.H, T = p.T, H4 = 4 * H;
  const int QS = 4 * HP;            // padded stride of one sequence's dG row: 4 quarters of HP
  float* dgs = sm;                  // BS * QS   dG_{next}, quarter q at [q*HP, q*HP+H), zero padded
  float* part = dgs + BS * QS;      // 4 * BS * H partial sums
  const int nbg = (p.B + BS - 1) / BS;
  const int dir = blockIdx.x / nbg, bg = blockIdx.x % nbg;
  const int b0 = bg * BS;
  const int tid = threadIdx.x;
  const int q = tid / H, j = tid - q * H;     // quarter of K (= gate block), output unit
  const bool mact = tid < H4;
  float ureg[HP];
  const float* Ud = p.U + (size_t)dir * H * H4;
#pragma unroll
  for (int i = 0; i < HP; ++i) ureg[i] = (i < H && mact) ? Ud[(size_t)j * H4 + q * H + i] : 0.f;
  for (int e = tid; e < BS * QS; e += blockDim.x) dgs[e] = 0.f;
  const int eb = tid / H, ej = tid - eb * H;
  const bool eact = tid < BS * H && (b0 + eb) < p.B;
  float dc_carry = 0.f;
  const size_t G8 = (size_t)8 * H, Y2 = (size_t)2 * H;
  __syncthreads();
  for (int sp = 0; sp < T; ++sp) {
    const int s = T - 1 - sp;
    const int t = dir == 0 ? s : T - 1 - s;
    const int tp = dir == 0 ? t - 1 : t + 1;
    float gi = 0.f, gf = 0.f, gg = 0.f, go = 0.f, c = 0.f, cp = 0.f, dyv = 0.f;
    float* grow = nullptr;
    if (eact) {
      const size_t row = (size_t)(b0 + eb) * T + t;
      grow = p.gates + row * G8 + (size_t)dir * H4 + ej;
      gi = grow[0]; gf = grow[H]; gg = grow[2 * (size_t)H]; go = grow[3 * (size_t)H];
      c = p.cell[row * Y2 + (size_t)dir * H + ej];
      cp = (s > 0) ? p.cell[((size_t)(b0 + eb) * T + tp) * Y2 + (size_t)dir * H + ej] : 0.f;
      dyv = p.dy[row * Y2 + (size_t)dir * H + ej];
    }
    if (sp > 0 && mact) {
      float2 acc[BS][2];
#pragma unroll
      for (int b = 0; b < BS; ++b) acc[b][0] = acc[b][1] = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < HP; i += 4) {
        const float2 u01 = make_float2(ureg[i], ureg[i + 1]), u23 = make_float2(ureg[i + 2], ureg[i + 3]);
#pragma unroll
        for (int b = 0; b < BS; ++b) {
          const float4 g4 = *reinterpret_cast<const float4*>(dgs + b * QS + q * HP + i);
          acc[b][0] = __ffma2_rn(make_float2(g4.x, g4.y), u01, acc[b][0]);
          acc[b][1] = __ffma2_rn(make_float2(g4.z, g4.w), u23, acc[b][1]);
        }
      }
#pragma unroll
      for (int b = 0; b < BS; ++b) part[(q * BS + b) * H + j] = (acc[b][0].x + acc[b][0].y) + (acc[b][1].x + acc[b][1].y);
    }
    __syncthreads();
    if (eact) {
      float dh = dyv;
      if (sp > 0) dh += (part[(0 * BS + eb) * H + ej] + part[(1 * BS + eb) * H + ej]) +
                        (part[(2 * BS + eb) * H + ej] + part[(3 * BS + eb) * H + ej]);
      const float tc = tanh_s(c);
      const float dc = dc_carry + dh * go * (1.f - tc * tc);
      const float d_o = dh * tc * dhsig_s(go);
      const float d_i = dc * gg * dhsig_s(gi);
      const float d_g = dc * gi * (1.f - gg * gg);
      const float d_f = dc * cp * dhsig_s(gf);
      dc_carry = dc * gf;
      grow[0] = d_i; grow[H] = d_f; grow[2 * (size_t)H] = d_g; grow[3 * (size_t)H] = d_o;
      float* dr = dgs + eb * QS + ej;
      dr[0] = d_i; dr[HP] = d_f; dr[2 * HP] = d_g; dr[3 * HP] = d_o;
    }
    __syncthreads();
  }
}

bool lstm_small_supported(int H) { return H % 4 == 0 && H >= 4 && H <= 104; }

static int small_bs(int B) {
  const int per_dir = max(1, num_sms() / 2);
  const int bs = (B + per_dir - 1) / per_dir;
  return bs <= 1 ? 1 : (bs <= 2 ? 2 : 4);
}

template <int HP, int BS>
static int small_launch(bool bwd, SmallParams& p, cudaStream_t s) {
  const int nbg = (p.B + BS - 1) / BS;
  const int threads = ((4 * p.H + 31) / 32) * 32;
  if (!bwd) {
    const size_t smem = sizeof(float) * ((size_t)BS * HP + (size_t)BS * 4 * p.H);
    GR_CUDA(cudaFuncSetAttribute(lstm_small_fwd_kernel<HP, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_small_fwd_kernel<HP, BS><<<2 * nbg, threads, smem, s>>>(p);
  } else {
    const size_t smem = sizeof(float) * ((size_t)BS * 4 * HP + (size_t)4 * BS * p.H);
    GR_CUDA(cudaFuncSetAttribute(lstm_small_bwd_kernel<HP, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_small_bwd_kernel<HP, BS><<<2 * nbg, threads, smem, s>>>(p);
  }
  GR_CHECK_LAUNCH("lstm_small_kernel");
  return GR_OK;
}

template <int HP>
static int small_launch_bs(bool bwd, SmallParams& p, cudaStream_t s) {
  if (p.BS == 1) return small_launch<HP, 1>(bwd, p, s);
  if (p.BS == 2) return small_launch<HP, 2>(bwd, p, s);
  return small_launch<HP, 4>(bwd, p, s);
}

int lstm_small_run(bool bwd, float* gates, const float* U, int B, int T, int H, float* y, float* cell,
                   const float* dy, cudaStream_t s) {
  SmallParams p;
  p.gates = gates; p.U = U; p.y = y; p.cell = cell; p.dy = dy; p.B = B; p.T = T; p.H = H;
  p.BS = small_bs(B);
  p.save = (!bwd && cell != nullptr) ? 1 : 0;
  // B > 4 * (SMs/2) sequences per direction are covered by more CTAs than SMs (several waves)
  if (H <= 32) return small_launch_bs<32>(bwd, p, s);
  if (H <= 64) return small_launch_bs<64>(bwd, p, s);
  return small_launch_bs<104>(bwd, p, s);
}

}  // namespace gr
