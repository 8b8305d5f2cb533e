// Best-path decoders for sm_100a.
//  * gr_ctc_bestpath_ref_f32: the reference's thresholded decoder, exact semantics of the
//    Python-2 loop in /root/reference/multimodal_fusion/sequence_decoding.py:41-50
//    (count-based filter: for every class s, the FIRST n_low[s] frames whose argmax is s are
//    deleted, n_low[s] = #frames with argmax s and max prob < threshold; then adjacent repeats
//    collapse; blank is kept).  SURVEY.md A.7.
//  * gr_ctc_greedy_f32: K.ctc_decode(greedy=True) -> TF ctc_greedy_decoder (merge_repeated=True).
// One CTA per sequence: (1) coalesced tile loads -> per-frame argmax/max (thread per frame),
// (2) warp 0 walks the frame labels 32 at a time (match_any ranks, ballot compaction).
#include "common.cuh"

namespace gr {

static constexpr int kDecThreads = 256;

template <bool kRefFilter>
__global__ void __launch_bounds__(kDecThreads)
bestpath_kernel(const float* __restrict__ probs, int T, int C, int drop, double threshold,
                const int32_t* __restrict__ seq_len, float eps, int32_t* __restrict__ out_ids,
                int32_t* __restrict__ out_len, float* __restrict__ out_score) {
  extern __shared__ __align__(16) float smem[];
  const int n = blockIdx.x;
  const int Cp = C | 1;
  float* tile = smem;                                              // kDecThreads * C (Cp reserved)
  int* best = reinterpret_cast<int*>(tile + kDecThreads * Cp);     // T entries: class | low<<16
  int* cnt = best + T;                                             // C running counts
  int* nlow = cnt + C;                                             // C low-confidence counts
  float* red = reinterpret_cast<float*>(nlow + C);                 // 8 partial scores
  const float* x = probs + (size_t)n * T * C;
  int Tn = T - drop;  // frames considered
  if (!kRefFilter && seq_len) Tn = max(0, min(seq_len[n], T));
  for (int c = threadIdx.x; c < C; c += kDecThreads) { cnt[c] = 0; nlow[c] = 0; }
  __syncthreads();
  float score = 0.f;
  for (int t0 = 0; t0 < Tn; t0 += kDecThreads) {
    const int rows = min(kDecThreads, Tn - t0);
    const float* src = x + (size_t)(drop + t0) * C;
    // flat copy of rows*C floats (128-bit when the tile start is 16-byte aligned: no per-element
    // division, 4x fewer load instructions); rows are then read at stride C
    const int tot = rows * C;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      const int tot4 = tot >> 2;
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* t4 = reinterpret_cast<float4*>(tile);
      for (int e = threadIdx.x; e < tot4; e += kDecThreads) t4[e] = __ldg(s4 + e);
      for (int e = (tot4 << 2) + threadIdx.x; e < tot; e += kDecThreads) tile[e] = __ldg(src + e);
    } else {
      for (int e = threadIdx.x; e < tot; e += kDecThreads) tile[e] = __ldg(src + e);
    }
    __syncthreads();
    if (threadIdx.x < rows) {
      const float* row = tile + threadIdx.x * C;
      float m = row[0];
      int am = 0;
      for (int c = 1; c < C; ++c) {
        const float v = row[c];
        if (v > m) { m = v; am = c; }
      }
      int low = 0;
      if (kRefFilter) {
        low = ((double)m < threshold) ? 1 : 0;
        if (low) atomicAdd(&nlow[am], 1);
      } else {
        score -= logf(m + eps);
      }
      best[t0 + threadIdx.x] = am | (low << 16);
    }
    __syncthreads();
  }
  if (!kRefFilter && out_score) {
    // deterministic block reduction of the per-thread partial scores
    for (int o = 16; o; o >>= 1) score += __shfl_xor_sync(0xffffffffu, score, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = score;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < kDecThreads / 32; ++w) s += red[w];
      out_score[n] = s;
    }
  }
  int32_t* out = out_ids + (size_t)n * T;
  __shared__ int s_len;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const unsigned lt = (1u << lane) - 1u;
    const int blank = C - 1;
    int prev_id = -1;  // last kept id (ref) / previous frame's id (greedy)
    int pos = 0;
    for (int t0 = 0; t0 < Tn; t0 += 32) {
      const int t = t0 + lane;
      const bool in = t < Tn;
      const int id = in ? (best[t] & 0xffff) : -1;
      bool keep = in;
      if (kRefFilter) {
        const unsigned same = __match_any_sync(0xffffffffu, id);
        if (in) {
          const int rank = cnt[id] + __popc(same & lt);
          keep = rank >= nlow[id];
        }
        __syncwarp();
        if (in && (same & lt) == 0) cnt[id] += __popc(same);
        __syncwarp();
      }
      const unsigned km = __ballot_sync(0xffffffffu, keep);
      // id of the previous element in the (kept) stream
      int pid;
      if (kRefFilter) {
        const unsigned below = km & lt;
        const int src = below ? 31 - __clz(below) : 0;
        pid = __shfl_sync(0xffffffffu, id, src);
        if (!below) pid = prev_id;
      } else {
        pid = __shfl_up_sync(0xffffffffu, id, 1);
        if (lane == 0) pid = prev_id;
      }
      bool emit = keep && id != pid;
      if (!kRefFilter) emit = emit && id != blank;
      const unsigned em = __ballot_sync(0xffffffffu, emit);
      if (emit) out[pos + __popc(em & lt)] = id;
      pos += __popc(em);
      // carry
      if (kRefFilter) {
        if (km) prev_id = __shfl_sync(0xffffffffu, id, 31 - __clz(km));
      } else {
        const int last = min(31, Tn - 1 - t0);
        prev_id = __shfl_sync(0xffffffffu, id, last);
      }
    }
    if (lane == 0) { s_len = pos; out_len[n] = pos; }
  }
  __syncthreads();
  for (int i = s_len + threadIdx.x; i < T; i += kDecThreads) out[i] = -1;
}

static size_t dec_smem(int T, int C) {
  return (size_t)kDecThreads * (C | 1) * 4 + (size_t)T * 4 + 2 * (size_t)C * 4 + 64;
}

}  // namespace gr

extern "C" int gr_ctc_bestpath_ref_f32(const float* probs, int N, int T, int C, int drop_frames,
                                       double threshold, int32_t* out_ids, int32_t* out_len,
                                       void* stream) {
  using namespace gr;
  if (!probs || !out_ids || !out_len) return set_error(GR_EINVAL, "bestpath_ref: null pointer");
  if (N <= 0 || T <= 0 || C < 1 || C > 32768 || drop_frames < 0 || drop_frames > T)
    return set_error(GR_EINVAL, "bestpath_ref: bad shape");
  const size_t smem = dec_smem(T, C);
  if (smem > 220 * 1024) return set_error(GR_EUNSUPPORTED, "bestpath_ref: T*C too large for shared memory");
  GR_CUDA(cudaFuncSetAttribute(bestpath_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  bestpath_kernel<true><<<N, kDecThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      probs, T, C, drop_frames, threshold, nullptr, 0.f, out_ids, out_len, nullptr);
  GR_CHECK_LAUNCH("bestpath_kernel<ref>");
  return GR_OK;
}

extern "C" int gr_ctc_greedy_f32(const float* probs, int N, int T, int C, const int32_t* seq_len,
                                 float eps, int32_t* out_ids, int32_t* out_len, float* out_score,
                                 void* stream) {
  using namespace gr;
  if (!probs || !out_ids || !out_len) return set_error(GR_EINVAL, "greedy: null pointer");
  if (N <= 0 || T <= 0 || C < 2 || C > 32768) return set_error(GR_EINVAL, "greedy: bad shape");
  const size_t smem = dec_smem(T, C);
  if (smem > 220 * 1024) return set_error(GR_EUNSUPPORTED, "greedy: T*C too large for shared memory");
  GR_CUDA(cudaFuncSetAttribute(bestpath_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  bestpath_kernel<false><<<N, kDecThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      probs, T, C, 0, 0.0, seq_len, eps, out_ids, out_len, out_score);
  GR_CHECK_LAUNCH("bestpath_kernel<greedy>");
  return GR_OK;
}
