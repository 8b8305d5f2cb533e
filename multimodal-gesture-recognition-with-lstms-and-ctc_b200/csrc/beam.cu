// CTC beam search for sm_100a: K.ctc_decode(greedy=False, beam_width=100, top_paths) ->
// TensorFlow CTCBeamSearchDecoder semantics (no call site in /root/reference; named by
// BASELINE.json config 5), in the batch form that oracle/decode_ref.py and
// oracle/beam_ref.c restate: per step every current leaf is updated, every (leaf, non-blank
// label) pair whose child is not already a leaf is a candidate, and the new leaf set is the top
// `beam_width` of (leaves ++ candidates) by total log-probability, ties broken by position.
//
// One CTA per sequence.  (1) a pre-pass kernel turns the probabilities into per-frame
// log-softmax(log(p+eps)) rows with the deterministic fp32 math of det_math.h; (2) the search
// kernel keeps the <= W leaves in shared memory, builds the W*C candidate keys
// (order-preserving float bits << 32 | ~index), bitonic-sorts them in shared memory and rebuilds
// the leaf set; the prefix tree lives in a (T+1) x W node pool in global memory.
// Bit-exact against oracle/beam_ref.c (same arithmetic, same tie-breaking).
#include "common.cuh"
#include "det_math.h"

namespace gr {

static constexpr int kBeamThreads = 512;

__global__ void beam_logsoftmax_kernel(const float* __restrict__ probs, float* __restrict__ lp, size_t rows, int C,
                                       float eps) {
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (size_t)gridDim.x * blockDim.x) {
    const float* row = probs + r * C;
    float* out = lp + r * C;
    float mx = DM_NEG_INF;
    for (int c = 0; c < C; ++c) {
      const float x = dm_logf(DM_ADD(row[c], eps));
      out[c] = x;
      if (x > mx) mx = x;
    }
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum = DM_ADD(sum, dm_expf(DM_ADD(out[c], -mx)));
    const float lse = DM_ADD(mx, dm_logf(sum));
    for (int c = 0; c < C; ++c) out[c] = DM_ADD(out[c], -lse);
  }
}

struct BeamParams {
  const float* lp;            // (N, T, C) log-softmax rows
  const int32_t* seq_len;     // (N) or null
  int32_t* pool_parent;       // (N, (T+1)*W)
  int32_t* pool_label;
  int32_t* out_ids;           // (N, top_paths, T)
  int32_t* out_len;           // (N, top_paths)
  float* out_logp;            // (N, top_paths)
  int N, T, C, W, NS, top_paths, merge_repeated;
};

struct LeafArrays {
  int* node; int* label; int* pslot; float* pb; float* pl; float* pt;
};

__global__ void __launch_bounds__(kBeamThreads) beam_search_kernel(BeamParams p) {
  extern __shared__ __align__(16) unsigned char bsm[];
  const int W = p.W, C = p.C, NC = p.C - 1, NS = p.NS, T = p.T;
  const int n = blockIdx.x, tid = threadIdx.x;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(bsm);           // NS
  int* ibase = reinterpret_cast<int*>(keys + NS);
  LeafArrays L[2];
  for (int k = 0; k < 2; ++k) {
    L[k].node = ibase; ibase += W;
    L[k].label = ibase; ibase += W;
    L[k].pslot = ibase; ibase += W;
    L[k].pb = reinterpret_cast<float*>(ibase); ibase += W;
    L[k].pl = reinterpret_cast<float*>(ibase); ibase += W;
    L[k].pt = reinterpret_cast<float*>(ibase); ibase += W;
  }
  float* npb = reinterpret_cast<float*>(ibase); ibase += W;
  float* npl = reinterpret_cast<float*>(ibase); ibase += W;
  float* npt = reinterpret_cast<float*>(ibase); ibase += W;
  int* map = ibase; ibase += W;
  float* lps = reinterpret_cast<float*>(ibase); ibase += C;
  int* s_nleaf = ibase; ibase += 4;
  unsigned char* has_child = reinterpret_cast<unsigned char*>(ibase);                // W * C
  int* hist = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(has_child + W * C) + 15) & ~uintptr_t(15));   // 256 digit counts
  int* s_sel = hist + 256;                                                           // [0] selected count, [1] need, [2] stop
  unsigned long long* s_thr = reinterpret_cast<unsigned long long*>(s_sel + 4);      // threshold prefix
  unsigned long long* sel = s_thr + 1;                                               // NSEL selected keys
  int NSEL = 32;
  while (NSEL < W) NSEL <<= 1;

  const int Tn = p.seq_len ? max(0, min(p.seq_len[n], T)) : T;
  int32_t* pparent = p.pool_parent + (size_t)n * (T + 1) * W;
  int32_t* plabel = p.pool_label + (size_t)n * (T + 1) * W;
  int cur = 0;
  if (tid == 0) {
    L[0].node[0] = 0; L[0].label[0] = -1; L[0].pslot[0] = -1;
    L[0].pb[0] = 0.f; L[0].pl[0] = DM_NEG_INF; L[0].pt[0] = 0.f;
    pparent[0] = -1; plabel[0] = -1;
    s_nleaf[0] = 1;
  }
  __syncthreads();
  int nleaf = 1;
  for (int t = 0; t < Tn; ++t) {
    const LeafArrays A = L[cur], Bn = L[cur ^ 1];
    const float* row = p.lp + ((size_t)n * T + t) * C;
    for (int c = tid; c < C; c += kBeamThreads) lps[c] = row[c];
    for (int e = tid; e < W * C; e += kBeamThreads) has_child[e] = 0;
    // only as many key slots as this step can populate: W leaf slots + nleaf * (C-1) candidates
    int ns = W + nleaf * NC;
    if (ns > NS) ns = NS;
    for (int e = tid; e < ns; e += kBeamThreads) keys[e] = 0ull;
    __syncthreads();
    // ---- update the current leaves
    if (tid < nleaf) {
      const int s = tid;
      float nl = DM_NEG_INF;
      if (A.node[s] != 0) {
        nl = A.pl[s];
        const int ps = A.pslot[s];
        if (ps >= 0) {
          const float prev = (A.label[s] == A.label[ps]) ? A.pb[ps] : A.pt[ps];
          nl = dm_lse(nl, prev);
          has_child[ps * C + A.label[s]] = 1;
        }
        nl = DM_ADD(nl, lps[A.label[s]]);
      }
      const float nb = DM_ADD(A.pt[s], lps[C - 1]);
      const float nt = dm_lse(nb, nl);
      npb[s] = nb; npl[s] = nl; npt[s] = nt;
      if (nt > DM_NEG_INF) keys[s] = ((unsigned long long)dm_ord(nt) << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)s);
    }
    __syncthreads();
    // ---- candidate children
    for (int e = tid; e < nleaf * NC; e += kBeamThreads) {
      const int s = e / NC, c = e - s * NC;
      if (has_child[s * C + c]) continue;
      if (!(A.pt[s] > DM_NEG_INF)) continue;
      const float prev = (c == A.label[s]) ? A.pb[s] : A.pt[s];
      if (!(prev > DM_NEG_INF)) continue;
      const float tot = DM_ADD(lps[c], prev);
      if (!(tot > DM_NEG_INF)) continue;
      const unsigned i = (unsigned)(W + s * NC + c);
      keys[i] = ((unsigned long long)dm_ord(tot) << 32) | (unsigned)(0xFFFFFFFFu - i);
    }
    __syncthreads();
    // ---- top-W selection.  Only the W largest keys matter and all keys are distinct, so the set is
    // independent of the algorithm: an MSB-first radix select (8-bit digits, shared histogram) finds
    // the W-th largest key, the <= W keys above it are compacted and only those are bitonic-sorted
    // (a full sort of the 4096 candidate slots took ~380 k cycles per step).
    if (tid == 0) { s_sel[0] = 0; s_sel[1] = W; s_sel[2] = 0; *s_thr = 0ull; }
    __syncthreads();
    for (int d = 7; d >= 0; --d) {
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
      const unsigned long long prefix = *s_thr;
      const int sh = 8 * d;
      for (int e = tid; e < ns; e += kBeamThreads) {
        const unsigned long long k = keys[e];
        if (k == 0ull) continue;
        if (d < 7 && ((k ^ prefix) >> (sh + 8)) != 0ull) continue;   // not in the current prefix class
        atomicAdd(&hist[(int)((k >> sh) & 255ull)], 1);
      }
      __syncthreads();
      if (tid < 32) {
        // digits 255 - 8*lane .. 248 - 8*lane belong to this lane; scan from the top
        int c[8], tot = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { c[i] = hist[255 - (tid * 8 + i)]; tot += c[i]; }
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (tid >= o) incl += v;
        }
        const int need = s_sel[1];
        const int above = incl - tot;                 // keys with a larger digit than this lane's
        const int all = __shfl_sync(0xffffffffu, incl, 31);
        if (all < need) {
          // fewer than `need` keys left in this class (only possible at d == 7: fewer valid keys than W):
          // everything valid is selected
          if (tid == 0) { *s_thr = 1ull; s_sel[2] = 1; }
        } else if (above < need && incl >= need) {
          int run = above;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (run < need && run + c[i] >= need) {
              const int g = 255 - (tid * 8 + i);
              *s_thr = prefix | ((unsigned long long)g << sh);
              s_sel[1] = need - run;
              if (run + c[i] == need) s_sel[2] = 1;    // the whole digit class is in: lower digits are free
            }
            run += c[i];
          }
        }
      }
      __syncthreads();
      if (s_sel[2]) break;
    }
    const unsigned long long thr = *s_thr;
    for (int e = tid; e < NSEL; e += kBeamThreads) sel[e] = 0ull;
    __syncthreads();
    for (int e = tid; e < ns; e += kBeamThreads) {
      const unsigned long long k = keys[e];
      if (k != 0ull && k >= thr) sel[atomicAdd(&s_sel[0], 1)] = k;
    }
    __syncthreads();
    // ---- bitonic sort of the selected keys, descending
    for (int k = 2; k <= NSEL; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < NSEL; i += kBeamThreads) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const unsigned long long a = sel[i], b = sel[ixj];
            const bool desc = (i & k) == 0;
            if (desc ? (a < b) : (a > b)) { sel[i] = b; sel[ixj] = a; }
          }
        }
        __syncthreads();
      }
    }
    if (tid < W) keys[tid] = sel[tid];
    __syncthreads();
    // ---- rebuild the leaf set
    if (tid < W) map[tid] = -1;
    __syncthreads();
    const bool valid = tid < W && keys[tid] != 0ull;
    unsigned idx = 0;
    if (valid) {
      idx = 0xFFFFFFFFu - (unsigned)(keys[tid] & 0xFFFFFFFFull);
      if (idx < (unsigned)W) map[idx] = tid;
    }
    const int cnt = __syncthreads_count(valid ? 1 : 0);
    if (valid) {
      const int r = tid;
      if (idx < (unsigned)W) {
        const int s = (int)idx;
        Bn.node[r] = A.node[s]; Bn.label[r] = A.label[s];
        Bn.pslot[r] = A.pslot[s] >= 0 ? map[A.pslot[s]] : -1;
        Bn.pb[r] = npb[s]; Bn.pl[r] = npl[s]; Bn.pt[r] = npt[s];
      } else {
        const int s = (int)(idx - W) / NC, c = (int)(idx - W) % NC;
        const float prev = (c == A.label[s]) ? A.pb[s] : A.pt[s];
        const float tot = DM_ADD(lps[c], prev);
        const int node = (t + 1) * W + r;
        pparent[node] = A.node[s]; plabel[node] = c;
        Bn.node[r] = node; Bn.label[r] = c; Bn.pslot[r] = map[s];
        Bn.pb[r] = DM_NEG_INF; Bn.pl[r] = tot; Bn.pt[r] = tot;
      }
    }
    nleaf = cnt;
    cur ^= 1;
    __syncthreads();
  }
  // ---- emit the best paths (leaf slots are already sorted by total)
  __threadfence_block();
  __syncthreads();
  const LeafArrays A = L[cur];
  if (tid < p.top_paths) {
    const int k = tid;
    int32_t* o = p.out_ids + ((size_t)n * p.top_paths + k) * T;
    int len = 0;
    if (k < nleaf) {
      // walk to the root writing labels backwards into the tail of the row, then compact forward
      int node = A.node[k], m = 0;
      while (node != 0) { o[T - 1 - m] = plabel[node]; node = pparent[node]; ++m; }
      int prev = -1;
      for (int i = 0; i < m; ++i) {
        const int v = o[T - m + i];
        if (p.merge_repeated && v == prev) continue;
        o[len++] = v; prev = v;
      }
      p.out_logp[(size_t)n * p.top_paths + k] = A.pt[k];
    } else {
      p.out_logp[(size_t)n * p.top_paths + k] = DM_NEG_INF;
    }
    for (int i = len; i < T; ++i) o[i] = -1;
    p.out_len[(size_t)n * p.top_paths + k] = len;
  }
}

static int beam_ns(int W, int C) {
  int ns = 2;
  while (ns < W * C) ns <<= 1;
  return ns;
}
static size_t beam_smem(int W, int C) {
  int nsel = 32;
  while (nsel < W) nsel <<= 1;
  return (size_t)beam_ns(W, C) * 8 + (size_t)W * 4 * (12 + 4) + (size_t)C * 4 + 16 + (size_t)((W * C + 15) & ~15) + 16 + 256 * 4 + 16 + 8 +
         (size_t)nsel * 8 + 64;
}

}  // namespace gr

extern "C" int gr_ctc_beam_workspace_bytes(int N, int T, int C, int beam_width, size_t* bytes_out) {
  if (!bytes_out || N <= 0 || T <= 0 || C < 2 || beam_width <= 0) return gr::set_error(GR_EINVAL, "beam_workspace_bytes: bad argument");
  const size_t lp = (size_t)N * T * C * sizeof(float);
  const size_t pool = (size_t)N * (T + 1) * beam_width * sizeof(int32_t);
  *bytes_out = ((lp + 255) & ~(size_t)255) + 2 * ((pool + 255) & ~(size_t)255) + 256;
  return GR_OK;
}

extern "C" int gr_ctc_beam_f32(const float* probs, int N, int T, int C, const int32_t* seq_len, float eps,
                               int beam_width, int top_paths, int merge_repeated, int32_t* out_ids,
                               int32_t* out_len, float* out_logp, void* workspace, size_t workspace_bytes,
                               void* stream) {
  using namespace gr;
  if (!probs || !out_ids || !out_len || !out_logp || !workspace) return set_error(GR_EINVAL, "beam: null pointer");
  if (N <= 0 || T <= 0 || C < 2 || beam_width <= 0 || top_paths <= 0 || top_paths > beam_width)
    return set_error(GR_EINVAL, "beam: bad shape");
  if (beam_width > kBeamThreads) return set_error(GR_EUNSUPPORTED, "beam: beam_width > 512");
  size_t need = 0;
  gr_ctc_beam_workspace_bytes(N, T, C, beam_width, &need);
  if (workspace_bytes < need) return set_error(GR_EWORKSPACE, "beam: workspace too small");
  const size_t smem = beam_smem(beam_width, C);
  if (smem > 220 * 1024) return set_error(GR_EUNSUPPORTED, "beam: beam_width*C too large for shared memory");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* w = static_cast<char*>(workspace);
  const size_t lp_bytes = (((size_t)N * T * C * sizeof(float)) + 255) & ~(size_t)255;
  const size_t pool_bytes = (((size_t)N * (T + 1) * beam_width * sizeof(int32_t)) + 255) & ~(size_t)255;
  BeamParams p;
  p.lp = reinterpret_cast<float*>(w);
  p.pool_parent = reinterpret_cast<int32_t*>(w + lp_bytes);
  p.pool_label = reinterpret_cast<int32_t*>(w + lp_bytes + pool_bytes);
  p.seq_len = seq_len; p.out_ids = out_ids; p.out_len = out_len; p.out_logp = out_logp;
  p.N = N; p.T = T; p.C = C; p.W = beam_width; p.NS = beam_ns(beam_width, C); p.top_paths = top_paths;
  p.merge_repeated = merge_repeated;
  const size_t rows = (size_t)N * T;
  const int blocks = (int)min((size_t)num_sms() * 8, (rows + 127) / 128);
  beam_logsoftmax_kernel<<<blocks, 128, 0, s>>>(probs, reinterpret_cast<float*>(w), rows, C, eps);
  GR_CHECK_LAUNCH("beam_logsoftmax_kernel");
  GR_CUDA(cudaFuncSetAttribute(beam_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  beam_search_kernel<<<N, kBeamThreads, smem, s>>>(p);
  GR_CHECK_LAUNCH("beam_search_kernel");
  return GR_OK;
}
