/* beam_ref.c -- CPU oracle for the CTC beam-search decoder (TEST INFRASTRUCTURE, see
 * oracle/__init__.py; parity unpinned by reference fixtures).
 *
 * Restates TensorFlow 1.12.1 CTCBeamSearchDecoder as reached through Keras 2.1.4
 * K.ctc_decode(greedy=False, beam_width=100, top_paths=1) (no call site in /root/reference;
 * named by BASELINE.json config 5) in its batch form -- identical to
 * oracle/decode_ref.py:ctc_beam_search, against which it is checked -- using the deterministic
 * fp32 math of csrc/det_math.h so that the CUDA kernel can be compared BIT-EXACTLY. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../multimodal-gesture-recognition-with-lstms-and-ctc_b200/csrc/det_math.h"

typedef struct { int node, label, pslot; float pb, pl, pt; } Leaf;

static int cmp_desc(const void* a, const void* b) {
  uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return x < y ? 1 : (x > y ? -1 : 0);
}

/* probs (T, C) softmax probabilities; out_ids (top_paths, T) padded -1. returns 0 on success */
int beam_ref(const float* probs, int T, int C, int seq_len, float eps, int W, int top_paths,
             int merge_repeated, int32_t* out_ids, int32_t* out_len, float* out_logp) {
  const int blank = C - 1, NC = C - 1;
  const int ncand_max = W + W * NC;
  Leaf* cur = (Leaf*)calloc(W, sizeof(Leaf));
  Leaf* nxt = (Leaf*)calloc(W, sizeof(Leaf));
  float* npb = (float*)malloc(sizeof(float) * W);
  float* npl = (float*)malloc(sizeof(float) * W);
  float* npt = (float*)malloc(sizeof(float) * W);
  uint8_t* has_child = (uint8_t*)malloc((size_t)W * C);
  uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * ncand_max);
  int* map = (int*)malloc(sizeof(int) * W);
  float* lp = (float*)malloc(sizeof(float) * C);
  int32_t* pool_parent = (int32_t*)malloc(sizeof(int32_t) * (size_t)(seq_len + 1) * W);
  int32_t* pool_label = (int32_t*)malloc(sizeof(int32_t) * (size_t)(seq_len + 1) * W);
  int nleaf = 1;
  cur[0].node = 0; cur[0].label = -1; cur[0].pslot = -1; cur[0].pb = 0.f; cur[0].pl = DM_NEG_INF; cur[0].pt = 0.f;
  pool_parent[0] = -1; pool_label[0] = -1;
  for (int t = 0; t < seq_len; ++t) {
    const float* row = probs + (size_t)t * C;
    float mx = DM_NEG_INF;
    for (int c = 0; c < C; ++c) { lp[c] = dm_logf(DM_ADD(row[c], eps)); if (lp[c] > mx) mx = lp[c]; }
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum = DM_ADD(sum, dm_expf(DM_ADD(lp[c], -mx)));
    const float lse = DM_ADD(mx, dm_logf(sum));
    for (int c = 0; c < C; ++c) lp[c] = DM_ADD(lp[c], -lse);
    memset(has_child, 0, (size_t)W * C);
    for (int s = 0; s < nleaf; ++s) {
      float nl = DM_NEG_INF;
      if (cur[s].node != 0) {
        nl = cur[s].pl;
        const int ps = cur[s].pslot;
        if (ps >= 0) {
          const float prev = (cur[s].label == cur[ps].label) ? cur[ps].pb : cur[ps].pt;
          nl = dm_lse(nl, prev);
          has_child[(size_t)ps * C + cur[s].label] = 1;
        }
        nl = DM_ADD(nl, lp[cur[s].label]);
      }
      const float nb = DM_ADD(cur[s].pt, lp[blank]);
      npb[s] = nb; npl[s] = nl; npt[s] = dm_lse(nb, nl);
    }
    int nk = 0;
    for (int s = 0; s < nleaf; ++s)
      if (npt[s] > DM_NEG_INF) keys[nk++] = ((uint64_t)dm_ord(npt[s]) << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)s);
    for (int s = 0; s < nleaf; ++s) {
      if (!(cur[s].pt > DM_NEG_INF)) continue;
      for (int c = 0; c < NC; ++c) {
        if (has_child[(size_t)s * C + c]) continue;
        const float prev = (c == cur[s].label) ? cur[s].pb : cur[s].pt;
        if (!(prev > DM_NEG_INF)) continue;
        const float tot = DM_ADD(lp[c], prev);
        if (!(tot > DM_NEG_INF)) continue;
        const uint32_t i = (uint32_t)(W + s * NC + c);
        keys[nk++] = ((uint64_t)dm_ord(tot) << 32) | (uint32_t)(0xFFFFFFFFu - i);
      }
    }
    qsort(keys, nk, sizeof(uint64_t), cmp_desc);
    const int nn = nk < W ? nk : W;
    for (int s = 0; s < W; ++s) map[s] = -1;
    for (int r = 0; r < nn; ++r) {
      const uint32_t i = 0xFFFFFFFFu - (uint32_t)(keys[r] & 0xFFFFFFFFu);
      if (i < (uint32_t)W) map[i] = r;
    }
    for (int r = 0; r < nn; ++r) {
      const uint32_t i = 0xFFFFFFFFu - (uint32_t)(keys[r] & 0xFFFFFFFFu);
      if (i < (uint32_t)W) {
        nxt[r] = cur[i];
        nxt[r].pb = npb[i]; nxt[r].pl = npl[i]; nxt[r].pt = npt[i];
        nxt[r].pslot = cur[i].pslot >= 0 ? map[cur[i].pslot] : -1;
      } else {
        const int s = (int)(i - W) / NC, c = (int)(i - W) % NC;
        const float prev = (c == cur[s].label) ? cur[s].pb : cur[s].pt;
        const float tot = DM_ADD(lp[c], prev);
        const int node = (t + 1) * W + r;
        pool_parent[node] = cur[s].node; pool_label[node] = c;
        nxt[r].node = node; nxt[r].label = c; nxt[r].pslot = map[s];
        nxt[r].pb = DM_NEG_INF; nxt[r].pl = tot; nxt[r].pt = tot;
      }
    }
    Leaf* tmp = cur; cur = nxt; nxt = tmp;
    nleaf = nn;
  }
  int32_t* rev = (int32_t*)malloc(sizeof(int32_t) * (size_t)(seq_len + 1));
  for (int k = 0; k < top_paths; ++k) {
    int32_t* o = out_ids + (size_t)k * T;
    for (int i = 0; i < T; ++i) o[i] = -1;
    if (k >= nleaf) { out_len[k] = 0; out_logp[k] = DM_NEG_INF; continue; }
    int n = 0, node = cur[k].node;
    while (node != 0) { rev[n++] = pool_label[node]; node = pool_parent[node]; }
    int len = 0, prev = -1;
    for (int i = n - 1; i >= 0; --i) {
      if (merge_repeated && rev[i] == prev) continue;
      o[len++] = rev[i]; prev = rev[i];
    }
    out_len[k] = len; out_logp[k] = cur[k].pt;
  }
  free(rev); free(cur); free(nxt); free(npb); free(npl); free(npt); free(has_child); free(keys); free(map); free(lp);
  free(pool_parent); free(pool_label);
  return 0;
}

/* exposed for the accuracy test of det_math.h */
float dm_test_expf(float x) { return dm_expf(x); }
float dm_test_logf(float x) { return dm_logf(x); }
float dm_test_log1pf(float x) { return dm_log1pf(x); }
float dm_test_lse(float a, float b) { return dm_lse(a, b); }
