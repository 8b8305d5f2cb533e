// Bidirectional Keras-LSTM recurrence (forward and BPTT) for sm_100a -- generic fp32 version.
// Replaces the tf.while_loop body behind `Bidirectional(LSTM(H, tanh, hard_sigmoid))`
// (/root/reference/audio_network/speech_lstm_ctc_words.py:56-77, skeletal_network/
// skeletal_lstm_ctc.py:309-331, multimodal_fusion/multimodal.py:159-168; maths SURVEY.md 3.5/A.4).
//
// Persistent kernel, both directions concurrently.  CTA = (direction, group of HS hidden units):
// the 4*HS columns of the recurrent kernel U that produce those units' i,f,c,o gates stay
// resident in shared memory for all T steps; the gate non-linearities and the cell update are
// fused behind the h*U product (thread = 4 batch rows x 1 unit x 4 gates, so no exchange is
// needed for the cell update).  Per step the CTAs of one direction exchange h_t through a
// transposed (H, B) buffer in L2 and a monotonic counter barrier.  `gates` is used in place:
// pre-activations P in, post-activation gates out (forward), dP out (backward).
#include <stdlib.h>
#include "common.cuh"

namespace gr {

static constexpr int kLstmThreads = 256;
static constexpr int kKC = 32;  // K-chunk rows staged in shared memory

__device__ __forceinline__ float hard_sigmoid(float v) { return fminf(fmaxf(0.2f * v + 0.5f, 0.f), 1.f); }
__device__ __forceinline__ float dhard_sigmoid_from_out(float s) { return (s > 0.f && s < 1.f) ? 0.2f : 0.f; }

__device__ __forceinline__ void grid_wait(const unsigned* ctr, unsigned target) {
  if (threadIdx.x == 0) {
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target);
  }
  __syncthreads();
}
__device__ __forceinline__ void grid_arrive(unsigned* ctr) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
  }
}

struct LstmFwdParams {
  float* gates;       // (B, T, 8H)
  const float* U;     // (2, H, 4H)
  float* y;           // (B, T, 2H)
  float* cell;        // (B, T, 2H) or null
  float* hT;          // (2 dir, 2 parity, H, Bp)
  float* cstate;      // (2, B, H) used when cell == null
  unsigned* counters; // 2 x 32 uints
  int B, T, H, Bp, HS, UG, NBQ;
};

__global__ void __launch_bounds__(kLstmThreads, 1) lstm_fwd_kernel(LstmFwdParams p) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, B = p.B, T = p.T, HS = p.HS, Bp = p.Bp;
  const int dir = blockIdx.x / p.UG, ug = blockIdx.x % p.UG;
  const int j0 = ug * HS;
  const int NBQ = p.NBQ, BBT = 4 * NBQ;
  float* Us = smem;                       // H * HS * 4
  float* hs = Us + (size_t)H * HS * 4;    // kKC * BBT
  const float* Ud = p.U + (size_t)dir * H * 4 * H;
  for (int e = threadIdx.x; e < H * HS * 4; e += kLstmThreads) {
    const int g = e & 3, jj = (e >> 2) % HS, k = (e >> 2) / HS;
    const int j = j0 + jj;
    Us[e] = (j < H) ? Ud[(size_t)k * 4 * H + g * H + j] : 0.f;
  }
  __syncthreads();
  const int jj = threadIdx.x % HS, bq = threadIdx.x / HS;
  const int j = j0 + jj;
  const bool tactive = (bq < NBQ) && (j < H);
  unsigned* ctr = p.counters + dir * 32;
  const size_t G8 = (size_t)8 * H, Y2 = (size_t)2 * H;

  for (int s = 0; s < T; ++s) {
    const int t = dir == 0 ? s : T - 1 - s;
    const int tp = dir == 0 ? t - 1 : t + 1;
    const float* hprev = p.hT + ((size_t)(dir * 2 + ((s + 1) & 1)) * H) * Bp;
    float* hcur = p.hT + ((size_t)(dir * 2 + (s & 1)) * H) * Bp;
    if (s > 0) grid_wait(ctr, (unsigned)s * p.UG);
    for (int b0 = 0; b0 < B; b0 += BBT) {
      const int bb = b0 + bq * 4;
      float pre[4][4], cprev[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int b = bb + r;
        const bool ok = tactive && b < B;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          pre[r][g] = ok ? p.gates[((size_t)b * T + t) * G8 + dir * 4 * H + g * H + j] : 0.f;
        if (s == 0 || !ok) cprev[r] = 0.f;
        else if (p.cell) cprev[r] = p.cell[((size_t)b * T + tp) * Y2 + dir * H + j];
        else cprev[r] = p.cstate[((size_t)dir * B + b) * H + j];
      }
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int g = 0; g < 4; ++g) acc[r][g] = 0.f;
      if (s > 0) {
        const int nb = min(BBT, Bp - b0);  // multiple of 4
        for (int kc = 0; kc < H; kc += kKC) {
          const int nk = min(kKC, H - kc);
          __syncthreads();
          for (int e = threadIdx.x * 4; e < nk * BBT; e += kLstmThreads * 4) {
            const int kk = e / BBT, c = e - kk * BBT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < nb) v = __ldcg(reinterpret_cast<const float4*>(hprev + (size_t)(kc + kk) * Bp + b0 + c));
            *reinterpret_cast<float4*>(hs + kk * BBT + c) = v;
          }
          __syncthreads();
          if (tactive) {
            const float4* up = reinterpret_cast<const float4*>(Us) + (size_t)kc * HS + jj;
            const float4* hp = reinterpret_cast<const float4*>(hs) + bq;
#pragma unroll 4
            for (int kk = 0; kk < nk; ++kk) {
              const float4 u = up[(size_t)kk * HS];
              const float4 h = hp[kk * NBQ];
              acc[0][0] = fmaf(h.x, u.x, acc[0][0]); acc[0][1] = fmaf(h.x, u.y, acc[0][1]);
              acc[0][2] = fmaf(h.x, u.z, acc[0][2]); acc[0][3] = fmaf(h.x, u.w, acc[0][3]);
              acc[1][0] = fmaf(h.y, u.x, acc[1][0]); acc[1][1] = fmaf(h.y, u.y, acc[1][1]);
              acc[1][2] = fmaf(h.y, u.z, acc[1][2]); acc[1][3] = fmaf(h.y, u.w, acc[1][3]);
              acc[2][0] = fmaf(h.z, u.x, acc[2][0]); acc[2][1] = fmaf(h.z, u.y, acc[2][1]);
              acc[2][2] = fmaf(h.z, u.z, acc[2][2]); acc[2][3] = fmaf(h.z, u.w, acc[2][3]);
              acc[3][0] = fmaf(h.w, u.x, acc[3][0]); acc[3][1] = fmaf(h.w, u.y, acc[3][1]);
              acc[3][2] = fmaf(h.w, u.z, acc[3][2]); acc[3][3] = fmaf(h.w, u.w, acc[3][3]);
            }
          }
        }
      }
      if (tactive) {
        float hv[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int b = bb + r;
          const float gi = hard_sigmoid(pre[r][0] + acc[r][0]);
          const float gf = hard_sigmoid(pre[r][1] + acc[r][1]);
          const float gg = tanhf(pre[r][2] + acc[r][2]);
          const float go = hard_sigmoid(pre[r][3] + acc[r][3]);
          const float c = gf * cprev[r] + gi * gg;
          const float h = go * tanhf(c);
          hv[r] = (b < B) ? h : 0.f;
          if (b < B) {
            float* gp = p.gates + ((size_t)b * T + t) * G8 + dir * 4 * H + j;
            gp[0] = gi; gp[H] = gf; gp[2 * H] = gg; gp[3 * H] = go;
            if (p.cell) p.cell[((size_t)b * T + t) * Y2 + dir * H + j] = c;
            else p.cstate[((size_t)dir * B + b) * H + j] = c;
            p.y[((size_t)b * T + t) * Y2 + dir * H + j] = h;
          }
        }
        if (bb < Bp)
          *reinterpret_cast<float4*>(hcur + (size_t)j * Bp + bb) = make_float4(hv[0], hv[1], hv[2], hv[3]);
      }
    }
    if (s + 1 < T) grid_arrive(ctr);
  }
}

// ---------------------------------------------------------------------------------------------
struct LstmBwdParams {
  float* gates;        // (B, T, 8H) in: i,f,g,o   out: dP
  const float* cell;   // (B, T, 2H)
  const float* dy;     // (B, T, 2H)
  const float* U;      // (2, H, 4H)
  float* dGT;          // (2 dir, 2 parity, 4H, Bp)
  float* dcs;          // (2, B, H) carried dc
  unsigned* counters;
  int B, T, H, Bp, HS, UG, NBQ, KS;
};

template <int HST>
__global__ void __launch_bounds__(kLstmThreads, 1) lstm_bwd_kernel(LstmBwdParams p) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, B = p.B, T = p.T, HS = p.HS, Bp = p.Bp;
  const int dir = blockIdx.x / p.UG, ug = blockIdx.x % p.UG;
  const int j0 = ug * HS;
  const int NBQ = p.NBQ, BBT = 4 * NBQ, KS = p.KS;
  const int K4 = 4 * H;
  float* Us = smem;                          // K4 * HST   Us[n*HST + jj] = U[dir][j0+jj][n]
  float* gs = Us + (size_t)K4 * HST;         // kKC * BBT  staged dG^T chunk
  float* red = gs + (size_t)kKC * BBT;       // KS * NBQ * 4 * HST partial sums
  const float* Ud = p.U + (size_t)dir * H * K4;
  for (int e = threadIdx.x; e < K4 * HST; e += kLstmThreads) {
    const int jj = e % HST, n = e / HST;
    const int j = j0 + jj;
    Us[e] = (jj < HS && j < H) ? Ud[(size_t)j * K4 + n] : 0.f;
  }
  __syncthreads();
  const int bq = threadIdx.x % NBQ, kpart = threadIdx.x / NBQ;
  unsigned* ctr = p.counters + dir * 32;
  const size_t G8 = (size_t)8 * H, Y2 = (size_t)2 * H;

  for (int sp = 0; sp < T; ++sp) {
    const int s = T - 1 - sp;                  // forward step index being differentiated
    const int t = dir == 0 ? s : T - 1 - s;
    const int tp = dir == 0 ? t - 1 : t + 1;   // forward-previous time (c_{prev})
    const float* gnext = p.dGT + ((size_t)(dir * 2 + ((sp + 1) & 1)) * K4) * Bp;
    float* gcur = p.dGT + ((size_t)(dir * 2 + (sp & 1)) * K4) * Bp;
    if (sp > 0) grid_wait(ctr, (unsigned)sp * p.UG);
    for (int b0 = 0; b0 < B; b0 += BBT) {
      float acc[4][HST];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < HST; ++q) acc[r][q] = 0.f;
      if (sp > 0) {
        const int nb = min(BBT, Bp - b0);
        for (int kc = 0; kc < K4; kc += kKC) {
          const int nk = min(kKC, K4 - kc);
          __syncthreads();
          for (int e = threadIdx.x * 4; e < nk * BBT; e += kLstmThreads * 4) {
            const int kk = e / BBT, c = e - kk * BBT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < nb) v = __ldcg(reinterpret_cast<const float4*>(gnext + (size_t)(kc + kk) * Bp + b0 + c));
            *reinterpret_cast<float4*>(gs + kk * BBT + c) = v;
          }
          __syncthreads();
          if (kpart < KS) {
            for (int kk = kpart; kk < nk; kk += KS) {
              const float4 g4 = *reinterpret_cast<const float4*>(gs + kk * BBT + bq * 4);
              const float* u = Us + (size_t)(kc + kk) * HST;
#pragma unroll
              for (int q = 0; q < HST; ++q) {
                const float uq = u[q];
                acc[0][q] = fmaf(g4.x, uq, acc[0][q]);
                acc[1][q] = fmaf(g4.y, uq, acc[1][q]);
                acc[2][q] = fmaf(g4.z, uq, acc[2][q]);
                acc[3][q] = fmaf(g4.w, uq, acc[3][q]);
              }
            }
          }
        }
      }
      __syncthreads();
      if (kpart < KS) {
        float* rp = red + ((size_t)(kpart * NBQ + bq) * 4) * HST;
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int q = 0; q < HST; ++q) rp[r * HST + q] = acc[r][q];
      }
      __syncthreads();
      // element-wise part: one (b, unit) per thread-iteration; adjacent threads -> adjacent b
      for (int o = threadIdx.x; o < BBT * HS; o += kLstmThreads) {
        const int jj = o / BBT, bl = o - jj * BBT;
        const int b = b0 + bl, j = j0 + jj;
        if (b >= Bp) continue;
        float d4[4] = {0.f, 0.f, 0.f, 0.f};
        if (b < B && j < H) {
          float dh = p.dy[((size_t)b * T + t) * Y2 + dir * H + j];
          if (sp > 0) {
            const int q4 = bl >> 2, r = bl & 3;
            for (int kp = 0; kp < KS; ++kp) dh += red[((size_t)(kp * NBQ + q4) * 4 + r) * HST + jj];
          }
          float* gp = p.gates + ((size_t)b * T + t) * G8 + dir * 4 * H + j;
          const float gi = gp[0], gf = gp[H], gg = gp[2 * H], go = gp[3 * H];
          const float c = p.cell[((size_t)b * T + t) * Y2 + dir * H + j];
          const float cp = (s > 0) ? p.cell[((size_t)b * T + tp) * Y2 + dir * H + j] : 0.f;
          float* dcp = p.dcs + ((size_t)dir * B + b) * H + j;
          const float dcn = (sp > 0) ? *dcp : 0.f;
          const float tc = tanhf(c);
          const float dc = dcn + dh * go * (1.f - tc * tc);
          d4[3] = dh * tc * dhard_sigmoid_from_out(go);
          d4[0] = dc * gg * dhard_sigmoid_from_out(gi);
          d4[2] = dc * gi * (1.f - gg * gg);
          d4[1] = dc * cp * dhard_sigmoid_from_out(gf);
          *dcp = dc * gf;
          gp[0] = d4[0]; gp[H] = d4[1]; gp[2 * H] = d4[2]; gp[3 * H] = d4[3];
        }
        if (j < H) {
#pragma unroll
          for (int g = 0; g < 4; ++g) gcur[((size_t)(g * H + j)) * Bp + b] = d4[g];
        }
      }
    }
    if (sp + 1 < T) grid_arrive(ctr);
  }
}

// tensor-core forward path (lstm_tc.cu)
bool lstm_tc_supported(int B, int H);
size_t lstm_tc_workspace_bytes(int B, int H);
size_t lstm_tc_trace_offset(int B, int H);
int lstm_fwd_tc_launch(float* gates, const float* U, int B, int T, int H, float* y, float* cell, void* workspace,
                       cudaStream_t s);

// tensor-core forward path with U resident in tensor memory (lstm_tcu.cu)
bool lstm_tcu_supported(int B, int H);
size_t lstm_tcu_workspace_bytes(int B, int H);
size_t lstm_tcu_trace_offset(int B, int H);
int lstm_tcu_grid(int B, int H);
int lstm_fwd_tcu_launch(float* gates, const float* U, int B, int T, int H, float* y, float* cell, void* workspace,
                        cudaStream_t s, float* aux = nullptr, int ld_aux = 0, int aux_mode = 0);

bool lstm_tcu_bwd_supported(int B, int H);
size_t lstm_tcu_bwd_workspace_bytes(int B, int H);
int lstm_bwd_tcu_launch(float* gates, const float* cell, const float* dy, const float* U, int B, int T, int H,
                        void* workspace, cudaStream_t s);

// register-resident-U path for narrow layers (lstm_small.cu)
bool lstm_small_supported(int H);
bool lstm_small_aligned(const float* gates, const float* cell, const float* dy);   // 16-byte rows for the bulk copies
int lstm_small_run(bool bwd, float* gates, const float* U, int B, int T, int H, float* y, float* cell,
                   const float* dy, cudaStream_t s);

// GR_LSTM_IMPL = generic | tc | tcu | small forces one implementation (debugging / cross-checks)
static bool use_small_path(int H, const float* gates, const float* cell, const float* dy) {
  const char* e = getenv("GR_LSTM_IMPL");
  if (e && strcmp(e, "small") != 0) return false;
  return lstm_small_supported(H) && lstm_small_aligned(gates, cell, dy);
}
static bool use_tcu_path(int B, int H) {
  const char* e = getenv("GR_LSTM_IMPL");
  if (e) return strcmp(e, "tcu") == 0 && lstm_tcu_supported(B, H);
  const char* d = getenv("GR_LSTM_TCU");           // default forward kernel for the wide layers; GR_LSTM_TCU=0: lstm_tc.cu
  if (d && d[0] == '0') return false;
  return lstm_tcu_supported(B, H);
}
static bool use_tcu_bwd_path(int B, int H) {
  const char* e = getenv("GR_LSTM_IMPL");
  if (e) return strcmp(e, "tcu") == 0 && lstm_tcu_bwd_supported(B, H);
  const char* d = getenv("GR_LSTM_TCU");
  if (d && d[0] == '0') return false;
  return lstm_tcu_bwd_supported(B, H);
}
static bool use_tc_path(int B, int H) {
  const char* e = getenv("GR_LSTM_IMPL");
  if (e && strcmp(e, "tc") != 0) return false;
  return lstm_tc_supported(B, H);
}

static void lstm_config(int B, int H, int* HS, int* UG, int* Bp) {
  const int per_dir = max(1, num_sms() / 2);
  int hs = (H + per_dir - 1) / per_dir;
  if (hs > 8) hs = 8;  // more unit groups than SMs/2 is not allowed (co-residency); checked by caller
  *HS = hs;
  *UG = (H + hs - 1) / hs;
  *Bp = (B + 3) & ~3;
}

}  // namespace gr

extern "C" size_t gr_debug_lstm_tc_trace_offset(int B, int H) { return gr::lstm_tc_trace_offset(B, H); }
extern "C" size_t gr_debug_lstm_tcu_trace_offset(int B, int H) { return gr::lstm_tcu_trace_offset(B, H); }

extern "C" int gr_lstm_workspace_bytes(int B, int H, size_t* bytes_out) {
  if (B <= 0 || H <= 0 || !bytes_out) return gr::set_error(GR_EINVAL, "lstm_workspace_bytes: bad argument");
  const size_t Bp = (B + 3) & ~3;
  // counters (256 B) + dG^T exchange (2*2*4H*Bp) [the h^T exchange aliases it] + carried state (2*B*H)
  size_t generic = 256 + sizeof(float) * ((size_t)16 * H * Bp + (size_t)2 * B * H) + 256;
  size_t tc = gr::lstm_tc_supported(B, H) ? gr::lstm_tc_workspace_bytes(B, H) : 0;
  size_t tcu = gr::lstm_tcu_supported(B, H) ? gr::lstm_tcu_workspace_bytes(B, H) : 0;
  if (tcu > tc) tc = tcu;
  size_t tcub = gr::lstm_tcu_bwd_supported(B, H) ? gr::lstm_tcu_bwd_workspace_bytes(B, H) : 0;
  if (tcub > tc) tc = tcub;
  *bytes_out = generic > tc ? generic : tc;
  return GR_OK;
}

extern "C" int gr_lstm_recurrence_fwd_f32(float* gates, const float* U, int B, int T, int H,
                                          float* y, float* cell, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  using namespace gr;
  if (!gates || !U || !y || !workspace) return set_error(GR_EINVAL, "lstm_fwd: null pointer");
  if (B <= 0 || T <= 0 || H <= 0) return set_error(GR_EINVAL, "lstm_fwd: bad shape");
  size_t need = 0;
  gr_lstm_workspace_bytes(B, H, &need);
  if (workspace_bytes < need) return set_error(GR_EWORKSPACE, "lstm_fwd: workspace too small");
  // Narrow layers (H <= 104) at a large batch: the register-resident kernel needs one CTA per 4 sequences (128 SMs for
  // 2.2 ms at B = 256, H = 100 = 280 SM-ms), the tensor-memory kernel 2 x ceil(H/16) CTAs per 256 sequences (14 SMs for
  // 6.4 ms = 89 SM-ms).  The fusion step is bound by the sum of its kernels' SM-time, not by this layer's latency, so the
  // FORWARD pass of such a layer goes to the tensor-memory kernel from B = 256 on: 42.2 -> 40.5 ms per step on the same box
  // (profiles/r02_narrow_layer_on_tcu.txt; the BPTT kernel does not pay off: 11.8 ms on 14 SMs, 6.5 on 28).
  // GR_LSTM_NARROW_FWD=small|tcu forces the choice (tcu from B = 128 on); GR_LSTM_IMPL overrides everything.
  const char* nf = getenv("GR_LSTM_NARROW_FWD");
  const bool narrow_tcu = !(nf && nf[0] == 's') && B >= ((nf && nf[0] == 't') ? 128 : 256) && !getenv("GR_LSTM_IMPL") &&
                          lstm_small_supported(H) && lstm_tcu_supported(B, H);
  if (!narrow_tcu && use_small_path(H, gates, nullptr, nullptr))
    return lstm_small_run(false, gates, U, B, T, H, y, cell, nullptr, static_cast<cudaStream_t>(stream));
  if (narrow_tcu || use_tcu_path(B, H))
    return lstm_fwd_tcu_launch(gates, U, B, T, H, y, cell, workspace, static_cast<cudaStream_t>(stream));
  if (use_tc_path(B, H))
    return lstm_fwd_tc_launch(gates, U, B, T, H, y, cell, workspace, static_cast<cudaStream_t>(stream));
  LstmFwdParams p;
  lstm_config(B, H, &p.HS, &p.UG, &p.Bp);
  if (2 * p.UG > num_sms()) return set_error(GR_EUNSUPPORTED, "lstm_fwd: H too large for the resident-U kernel");
  p.gates = gates; p.U = U; p.y = y; p.cell = cell; p.B = B; p.T = T; p.H = H;
  char* w = static_cast<char*>(workspace);
  p.counters = reinterpret_cast<unsigned*>(w);
  p.hT = reinterpret_cast<float*>(w + 256);
  p.cstate = p.hT + (size_t)16 * H * p.Bp;
  p.NBQ = min(kLstmThreads / p.HS, min(64, p.Bp / 4));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GR_CUDA(cudaMemsetAsync(p.counters, 0, 256, s));
  const size_t smem = sizeof(float) * ((size_t)H * p.HS * 4 + (size_t)kKC * 4 * p.NBQ);
  if (smem > 220 * 1024) return set_error(GR_EUNSUPPORTED, "lstm_fwd: shared memory");
  GR_CUDA(cudaFuncSetAttribute(lstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {&p};
  GR_CUDA(cudaLaunchCooperativeKernel((void*)lstm_fwd_kernel, dim3(2 * p.UG), dim3(kLstmThreads), args, smem, s));
  return GR_OK;
}

extern "C" int gr_lstm_recurrence_grid(int B, int H) {
  if (B <= 0 || H <= 0 || gr::lstm_small_supported(H) || !gr::use_tcu_path(B, H)) return 0;
  return gr::lstm_tcu_grid(B, H);
}

extern "C" int gr_lstm_recurrence_aux_supported(int B, int H) {
  return (!gr::lstm_small_supported(H) && gr::use_tcu_path(B, H)) ? 1 : 0;
}

extern "C" int gr_lstm_recurrence_fwd_aux_f32(float* gates, const float* U, int B, int T, int H, float* y, float* cell,
                                              float* aux, int ld_aux, int accumulate, void* workspace,
                                              size_t workspace_bytes, void* stream) {
  using namespace gr;
  if (!gates || !U || !aux || !workspace) return set_error(GR_EINVAL, "lstm_fwd_aux: null pointer");
  if (B <= 0 || T <= 0 || H <= 0 || ld_aux < 2 * H) return set_error(GR_EINVAL, "lstm_fwd_aux: bad shape");
  if ((ld_aux % 4) || (H % 4) || (reinterpret_cast<uintptr_t>(aux) & 15))
    return set_error(GR_EUNSUPPORTED, "lstm_fwd_aux: aux, ld_aux and H must be 16-byte aligned");
  if (!gr_lstm_recurrence_aux_supported(B, H))
    return set_error(GR_EUNSUPPORTED, "lstm_fwd_aux: only the tensor-memory recurrence (wide layers) has the auxiliary output");
  size_t need = 0;
  gr_lstm_workspace_bytes(B, H, &need);
  if (workspace_bytes < need) return set_error(GR_EWORKSPACE, "lstm_fwd_aux: workspace too small");
  return lstm_fwd_tcu_launch(gates, U, B, T, H, y, cell, workspace, static_cast<cudaStream_t>(stream), aux, ld_aux,
                             accumulate ? 2 : 1);
}

extern "C" int gr_lstm_recurrence_bwd_f32(float* gates, const float* cell, const float* dy,
                                          const float* U, int B, int T, int H, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  using namespace gr;
  if (!gates || !cell || !dy || !U || !workspace) return set_error(GR_EINVAL, "lstm_bwd: null pointer");
  if (B <= 0 || T <= 0 || H <= 0) return set_error(GR_EINVAL, "lstm_bwd: bad shape");
  size_t need = 0;
  gr_lstm_workspace_bytes(B, H, &need);
  if (workspace_bytes < need) return set_error(GR_EWORKSPACE, "lstm_bwd: workspace too small");
  const char* nbw = getenv("GR_LSTM_NARROW_BWD");
  const bool narrow_tcu = nbw && nbw[0] == 't' && B >= 128 && !getenv("GR_LSTM_IMPL") && lstm_tcu_bwd_supported(B, H);
  if (!narrow_tcu && use_small_path(H, gates, cell, dy))
    return lstm_small_run(true, gates, U, B, T, H, nullptr, const_cast<float*>(cell), dy, static_cast<cudaStream_t>(stream));
  if (narrow_tcu || use_tcu_bwd_path(B, H))
    return lstm_bwd_tcu_launch(gates, cell, dy, U, B, T, H, workspace, static_cast<cudaStream_t>(stream));
  LstmBwdParams p;
  lstm_config(B, H, &p.HS, &p.UG, &p.Bp);
  if (2 * p.UG > num_sms()) return set_error(GR_EUNSUPPORTED, "lstm_bwd: H too large for the resident-U kernel");
  p.gates = gates; p.cell = cell; p.dy = dy; p.U = U; p.B = B; p.T = T; p.H = H;
  char* w = static_cast<char*>(workspace);
  p.counters = reinterpret_cast<unsigned*>(w);
  p.dGT = reinterpret_cast<float*>(w + 256);
  p.dcs = p.dGT + (size_t)16 * H * p.Bp;
  int nbq = 1;
  while (nbq * 2 <= min(64, p.Bp / 4)) nbq *= 2;
  if (nbq * 4 < p.Bp && nbq < 64) nbq *= 2;  // cover ragged small batches in one tile
  p.NBQ = nbq;
  p.KS = kLstmThreads / nbq;
  if (p.KS > kKC) p.KS = kKC;
  const int HST = p.HS <= 1 ? 1 : p.HS <= 2 ? 2 : p.HS <= 4 ? 4 : 8;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GR_CUDA(cudaMemsetAsync(p.counters, 0, 256, s));
  const size_t smem = sizeof(float) * ((size_t)4 * H * HST + (size_t)kKC * 4 * p.NBQ +
                                       (size_t)p.KS * p.NBQ * 4 * HST);
  if (smem > 220 * 1024) return set_error(GR_EUNSUPPORTED, "lstm_bwd: shared memory");
  void* args[] = {&p};
  auto go = [&](auto kern) -> int {
    GR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GR_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(2 * p.UG), dim3(kLstmThreads), args, smem, s));
    return GR_OK;
  };
  switch (HST) {
    case 1: return go(lstm_bwd_kernel<1>);
    case 2: return go(lstm_bwd_kernel<2>);
    case 4: return go(lstm_bwd_kernel<4>);
    default: return go(lstm_bwd_kernel<8>);
  }
}
