/* det_math.h -- deterministic fp32 exp / log / log1p built ONLY from correctly-rounded IEEE
 * operations (+, *, /, fma, integer bit manipulation), so that the CUDA beam-search kernel and
 * its C oracle (oracle/beam_ref.c) produce bit-identical log-probabilities and therefore make
 * identical beam-pruning decisions.  (libm's expf/log1pf and CUDA's differ in the last ulp.)
 * Accuracy: a few ulp -- checked against libm in tests/test_oracle_beam.py.
 * Host build: compile with -ffp-contract=off.  Device build: the *_rn intrinsics are never
 * contracted by nvcc. */
#ifndef GR_DET_MATH_H_
#define GR_DET_MATH_H_

#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define DM_FN __host__ __device__ __forceinline__
#else
#include <math.h>
#define DM_FN static inline
#endif

#if defined(__CUDA_ARCH__)
#define DM_ADD(a, b) __fadd_rn((a), (b))
#define DM_MUL(a, b) __fmul_rn((a), (b))
#define DM_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define DM_DIV(a, b) __fdiv_rn((a), (b))
#else
#define DM_ADD(a, b) ((a) + (b))
#define DM_MUL(a, b) ((a) * (b))
#define DM_FMA(a, b, c) fmaf((a), (b), (c))
#define DM_DIV(a, b) ((a) / (b))
#endif

#define DM_NEG_INF (-__builtin_inff())

DM_FN uint32_t dm_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
DM_FN float dm_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* exp(x) for x <= 0 (returns 0 below ~-87.3, i.e. flushes instead of producing denormals) */
DM_FN float dm_expf(float x) {
  if (!(x > -87.0f)) return 0.0f;
  if (x > 0.0f) x = 0.0f;
  /* k = round(x * log2(e)) via the 1.5*2^23 trick */
  const float magic = 12582912.0f;
  float kf = DM_ADD(DM_MUL(x, 1.4426950408889634f), magic);
  int32_t k = (int32_t)dm_f2u(kf) - (int32_t)dm_f2u(magic);
  kf = DM_ADD(kf, -magic);
  /* r = x - k*ln2 (Cody-Waite, two constants) */
  float r = DM_FMA(kf, -0.693145751953125f, x);
  r = DM_FMA(kf, -1.42860682030941723212e-6f, r);
  /* e^r, |r| <= 0.3466: degree-6 Taylor/minimax in Horner form */
  float p = 1.3888889e-3f;
  p = DM_FMA(p, r, 8.3333333e-3f);
  p = DM_FMA(p, r, 4.1666668e-2f);
  p = DM_FMA(p, r, 1.6666667e-1f);
  p = DM_FMA(p, r, 0.5f);
  p = DM_FMA(p, r, 1.0f);
  p = DM_FMA(p, r, 1.0f);
  /* scale by 2^k, k in [-126, 0] */
  return dm_u2f(dm_f2u(p) + ((uint32_t)k << 23));
}

/* log(x) for finite x > 0 (normal range) */
DM_FN float dm_logf(float x) {
  uint32_t ix = dm_f2u(x);
  /* reduce x into [sqrt(2)/2, sqrt(2)) */
  ix += 0x3f800000u - 0x3f3504f3u;
  int32_t k = (int32_t)(ix >> 23) - 0x7f;
  ix = (ix & 0x007fffffu) + 0x3f3504f3u;
  float f = DM_ADD(dm_u2f(ix), -1.0f);
  float s = DM_DIV(f, DM_ADD(2.0f, f));
  float z = DM_MUL(s, s);
  float w = DM_MUL(z, z);
  float t1 = DM_MUL(w, DM_FMA(w, 0.24279078841f, 0.40000972152f));
  float t2 = DM_MUL(z, DM_FMA(w, 0.28498786688f, 0.66666662693f));
  float R = DM_ADD(t2, t1);
  float hfsq = DM_MUL(0.5f, DM_MUL(f, f));
  float dk = (float)k;
  /* log(x) = k*ln2_hi + (f - (hfsq - (s*(hfsq+R) + k*ln2_lo))) */
  float a = DM_FMA(dk, 9.0580006145e-06f, DM_MUL(s, DM_ADD(hfsq, R)));
  float b = DM_ADD(hfsq, -a);
  float c = DM_ADD(f, -b);
  return DM_FMA(dk, 6.9313812256e-01f, c);
}

/* log(1 + y) for 0 <= y <= 1 */
DM_FN float dm_log1pf(float y) {
  float u = DM_ADD(1.0f, y);
  float l = dm_logf(u);
  /* first-order correction for the rounding of 1+y:  + (y - (u-1)) / u */
  float d = DM_ADD(y, -DM_ADD(u, -1.0f));
  return DM_ADD(l, DM_DIV(d, u));
}

/* log(exp(a) + exp(b)) with -inf as log(0) */
DM_FN float dm_lse(float a, float b) {
  if (a == DM_NEG_INF) return b;
  if (b == DM_NEG_INF) return a;
  float m = a > b ? a : b;
  float lo = a > b ? b : a;
  return DM_ADD(m, dm_log1pf(dm_expf(DM_ADD(lo, -m))));
}

/* order-preserving map float -> uint32 (total order; -inf lowest) */
DM_FN uint32_t dm_ord(float f) {
  uint32_t u = dm_f2u(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

#endif /* GR_DET_MATH_H_ */
