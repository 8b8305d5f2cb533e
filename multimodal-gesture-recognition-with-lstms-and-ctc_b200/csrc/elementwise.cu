// Dense+softmax head, reductions, residual/concat, optimiser step and Philox regularisers.
// Reference call sites: Dense/softmax speech_lstm_ctc_words.py:86-90, multimodal.py:175-179;
// layers.add speech:79, multimodal:111,117; Merge(concat) multimodal:155-156; Dropout speech:82;
// GaussianNoise speech:53; Adam(clipvalue)/maxnorm multimodal.py:206-208,165.
#include <curand_kernel.h>
#include "common.cuh"

namespace gr {

static constexpr int kCMax = 64;  // classes held in registers by the head kernels

// ------------------------------------------------------------------ Dense + softmax forward
// warp per row; lanes stride over Fin; W^T staged in shared memory with odd row stride.
template <int CR>
__global__ void __launch_bounds__(256) dense_softmax_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mask,
                                                               const float* __restrict__ Wd, const float* __restrict__ bd,
                                                               int R, int Fin, int C, float* __restrict__ logits,
                                                               float* __restrict__ probs) {
  extern __shared__ float ws[];  // Fin * Cs
  const int Cs = C | 1;
  for (int e = threadIdx.x; e < Fin * C; e += blockDim.x) {
    const int k = e / C, c = e - k * C;
    ws[k * Cs + c] = Wd[e];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + warp; r < R; r += gridDim.x * wpb) {
    float acc[CR];
#pragma unroll
    for (int c = 0; c < CR; ++c) acc[c] = 0.f;
    const float* xr = x + (size_t)r * Fin;
    const float* mr = mask ? mask + (size_t)r * Fin : nullptr;
    for (int k = lane; k < Fin; k += 32) {
      float xv = xr[k];
      if (mr) xv *= mr[k];
      const float* w = ws + k * Cs;
#pragma unroll
      for (int c = 0; c < CR; ++c)
        if (c < C) acc[c] = fmaf(xv, w[c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < CR; ++c)
      if (c < C) {
#pragma unroll
        for (int o = 16; o; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
        acc[c] += bd[c];
      }
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < CR; ++c)
      if (c < C) m = fmaxf(m, acc[c]);
    float s = 0.f;
    float e_[CR];
#pragma unroll
    for (int c = 0; c < CR; ++c)
      if (c < C) { e_[c] = expf(acc[c] - m); s += e_[c]; }
    const float inv = 1.0f / s;
    // lane c writes column c (and c+32)
#pragma unroll
    for (int c = 0; c < CR; ++c)
      if (c < C && (c & 31) == lane) {
        if (logits) logits[(size_t)r * C + c] = acc[c];
        if (probs) probs[(size_t)r * C + c] = e_[c] * inv;
      }
  }
}

// Thread-per-row variant (Fin % 4 == 0, 16-byte aligned rows): W^T in shared memory with rows padded to
// CP floats and read with broadcast LDS.128, the row's dot products and its softmax stay in one thread
// (no shuffles): ~5x fewer instructions per row than the warp-per-row kernel (1.14 -> see DESIGN 4.5).
template <int CP>
__global__ void __launch_bounds__(128) dense_softmax_fwd_row_kernel(const float* __restrict__ x, const float* __restrict__ mask,
                                                                    const float* __restrict__ Wd, const float* __restrict__ bd,
                                                                    int R, int Fin, int C, float* __restrict__ logits,
                                                                    float* __restrict__ probs) {
  extern __shared__ __align__(16) float ws[];  // Fin * CP, zero padded; then CP biases
  for (int e = threadIdx.x; e < Fin * CP; e += blockDim.x) {
    const int k = e / CP, c = e - k * CP;
    ws[e] = c < C ? Wd[k * C + c] : 0.f;
  }
  float* bs = ws + Fin * CP;
  for (int c = threadIdx.x; c < CP; c += blockDim.x) bs[c] = c < C ? bd[c] : 0.f;
  __syncthreads();
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
    float acc[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) acc[c] = bs[c];
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)r * Fin);
    const float4* mr = mask ? reinterpret_cast<const float4*>(mask + (size_t)r * Fin) : nullptr;
    for (int k4 = 0; k4 < Fin / 4; ++k4) {
      float4 xv = __ldg(xr + k4);
      if (mr) { const float4 mv = __ldg(mr + k4); xv.x *= mv.x; xv.y *= mv.y; xv.z *= mv.z; xv.w *= mv.w; }
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4* w = reinterpret_cast<const float4*>(ws + (k4 * 4 + j) * CP);
#pragma unroll
        for (int c4 = 0; c4 < CP / 4; ++c4) {
          const float4 wv = w[c4];
          acc[4 * c4] = fmaf(xs[j], wv.x, acc[4 * c4]);
          acc[4 * c4 + 1] = fmaf(xs[j], wv.y, acc[4 * c4 + 1]);
          acc[4 * c4 + 2] = fmaf(xs[j], wv.z, acc[4 * c4 + 2]);
          acc[4 * c4 + 3] = fmaf(xs[j], wv.w, acc[4 * c4 + 3]);
        }
      }
    }
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < CP; ++c)
      if (c < C) m = fmaxf(m, acc[c]);
    float sum = 0.f;
    float e_[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      e_[c] = c < C ? expf(acc[c] - m) : 0.f;
      sum += e_[c];
    }
    const float inv = 1.0f / sum;
    float* lr = logits ? logits + (size_t)r * C : nullptr;
    float* pr = probs ? probs + (size_t)r * C : nullptr;
#pragma unroll
    for (int c = 0; c < CP; ++c)
      if (c < C) {
        if (lr) lr[c] = acc[c];
        if (pr) pr[c] = e_[c] * inv;
      }
  }
}

// ------------------------------------------------------------------ Dense backward
// thread <-> input feature k; C accumulators (dWd[k,:]) and W[k,:] in registers; rows tiled by 32.  The row loop is
// unrolled by four with all x / mask loads issued first (one load pair in flight per thread left the kernel latency-bound
// at 0.9 TB/s: 0.71 ms for the fusion head), and the staged gradients are zero-padded to CR columns and read as float4.
template <int CR>
__global__ void __launch_bounds__(256) dense_bwd_kernel(const float* __restrict__ x, const float* __restrict__ mask,
                                                       const float* __restrict__ Wd, const float* __restrict__ g,
                                                       int R, int Fin, int C, int rows_per_cta, float* __restrict__ dWd,
                                                       float* __restrict__ dbd, float* __restrict__ dx) {
  __shared__ __align__(16) float gs[32][CR];
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(R, r_begin + rows_per_cta);
  if (r_begin >= r_end) return;
  float accb = 0.f;
  for (int kbase = 0; kbase < Fin; kbase += blockDim.x) {
    const int k = kbase + threadIdx.x;
    const bool kok = k < Fin;
    float acc[CR], w[CR];
#pragma unroll
    for (int c = 0; c < CR; ++c) { acc[c] = 0.f; w[c] = (kok && c < C && dx) ? Wd[(size_t)k * C + c] : 0.f; }
    for (int r0 = r_begin; r0 < r_end; r0 += 32) {
      const int nr = min(32, r_end - r0);
      __syncthreads();
      for (int e = threadIdx.x; e < nr * CR; e += blockDim.x) {
        const int rr = e / CR, c = e - rr * CR;
        gs[rr][c] = c < C ? g[(size_t)(r0 + rr) * C + c] : 0.f;
      }
      __syncthreads();
      if (kbase == 0 && threadIdx.x < C)
        for (int r = 0; r < nr; ++r) accb += gs[r][threadIdx.x];
      if (kok) {
        for (int r = 0; r < nr; r += 4) {
          float xv[4], mv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool ok = r + u < nr;
            const size_t off = (size_t)(r0 + r + u) * Fin + k;
            mv[u] = ok ? (mask ? mask[off] : 1.f) : 0.f;
            xv[u] = ok ? x[off] : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (r + u < nr) {
              const float xm = xv[u] * mv[u];
              const float4* g4 = reinterpret_cast<const float4*>(gs[r + u]);
              float d = 0.f;
#pragma unroll
              for (int c4 = 0; c4 < CR / 4; ++c4) {
                const float4 gq = g4[c4];
                acc[4 * c4] = fmaf(xm, gq.x, acc[4 * c4]);         d = fmaf(gq.x, w[4 * c4], d);
                acc[4 * c4 + 1] = fmaf(xm, gq.y, acc[4 * c4 + 1]); d = fmaf(gq.y, w[4 * c4 + 1], d);
                acc[4 * c4 + 2] = fmaf(xm, gq.z, acc[4 * c4 + 2]); d = fmaf(gq.z, w[4 * c4 + 2], d);
                acc[4 * c4 + 3] = fmaf(xm, gq.w, acc[4 * c4 + 3]); d = fmaf(gq.w, w[4 * c4 + 3], d);
              }
              if (dx) dx[(size_t)(r0 + r + u) * Fin + k] = d * mv[u];
            }
          }
        }
      }
    }
    if (kok) {
#pragma unroll
      for (int c = 0; c < CR; ++c)
        if (c < C) atomicAdd(dWd + (size_t)k * C + c, acc[c]);
    }
  }
  if (threadIdx.x < C) atomicAdd(dbd + threadIdx.x, accb);
}

// ------------------------------------------------------------------ column sums
__global__ void colsum_kernel(const float* __restrict__ a, int R, int N, int lda, int rows_per_cta, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(R, r0 + rows_per_cta);
  if (n >= N) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int r = r0;
  for (; r + 3 < r1; r += 4) {
    s0 += a[(size_t)r * lda + n];
    s1 += a[(size_t)(r + 1) * lda + n];
    s2 += a[(size_t)(r + 2) * lda + n];
    s3 += a[(size_t)(r + 3) * lda + n];
  }
  for (; r < r1; ++r) s0 += a[(size_t)r * lda + n];
  atomicAdd(out + n, (s0 + s1) + (s2 + s3));
}

__global__ void add_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ o, size_t n4,
                           const float* at, const float* bt, float* ot, size_t tail0, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 u = a[i], v = b[i];
    o[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
  }
  if (blockIdx.x == 0)
    for (size_t i = tail0 + threadIdx.x; i < n; i += blockDim.x) ot[i] = at[i] + bt[i];
}

// out[r, col0 + c] = a[r, c] + b[r, c]: the residual add of a tower (layers.add, speech:79,
// multimodal.py:111,117) written straight into its column block of the Merge(concat) buffer
// (multimodal.py:155-156), so no separate concat pass reads and rewrites 3.3 GB
__global__ void add_into_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float* __restrict__ o,
                                size_t rows, int cols4, int ldo) {
  const size_t total = rows * (size_t)cols4;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t r = e / cols4;
    const int c = (int)(e - r * cols4);
    const float4 u = a[e], v = b[e];
    *reinterpret_cast<float4*>(o + r * ldo + 4 * c) = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
  }
}

__global__ void concat2_kernel(const float* __restrict__ a, int Fa, const float* __restrict__ b, int Fb,
                               float* __restrict__ o, size_t rows) {
  const int Fo = Fa + Fb;
  const size_t total = rows * Fo;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t r = e / Fo;
    const int c = (int)(e - r * Fo);
    o[e] = c < Fa ? a[r * Fa + c] : b[r * Fb + (c - Fa)];
  }
}

// out[r,k] (+)= tmp[r,k] * mask[r / rows_per_seq, k]   (input-dropout mask applied to dX)
__global__ void mask_mul_acc_kernel(float* __restrict__ out, const float* __restrict__ tmp, const float* __restrict__ mask,
                                    int rows_per_seq, int R, int K, int accumulate) {
  const size_t total = (size_t)R * K;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t r = e / K;
    const int k = (int)(e - r * K);
    const float v = tmp[e] * mask[(r / rows_per_seq) * K + k];
    out[e] = accumulate ? out[e] + v : v;
  }
}

// ------------------------------------------------------------------ Adam (+clipvalue) and maxnorm
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr_t, float beta1, float beta2, float eps,
                            float clip) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    if (clip > 0.f) gi = fminf(fmaxf(gi, -clip), clip);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
  }
}
// w (rows, cols): per column n = sqrt(sum_r w^2); w *= clip(n, 0, max_norm) / (1e-7 + n)
__global__ void maxnorm_kernel(float* __restrict__ w, int rows, int cols, float max_norm) {
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < cols)
    for (int r = threadIdx.y; r < rows; r += 8) { const float x = w[(size_t)r * cols + c]; s = fmaf(x, x, s); }
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < 8; ++i) tot += part[i][threadIdx.x];
  const float nrm = sqrtf(tot);
  const float scale = fminf(fmaxf(nrm, 0.f), max_norm) / (1e-7f + nrm);
  if (c < cols)
    for (int r = threadIdx.y; r < rows; r += 8) w[(size_t)r * cols + c] *= scale;
}

// ------------------------------------------------------------------ Philox regularisers
__global__ void dropout_mask_kernel(float* __restrict__ out, size_t n, float p, float scale, uint64_t seed, uint64_t offset) {
  const size_t n4 = (n + 3) / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    curandStatePhilox4_32_10_t st;
    curand_init(seed, i, offset, &st);
    const float4 u = curand_uniform4(&st);
    const float vals[4] = {u.x, u.y, u.z, u.w};
    for (int j = 0; j < 4; ++j) {
      const size_t e = i * 4 + j;
      if (e < n) out[e] = (vals[j] > p) ? scale : 0.f;  // keep with probability 1-p
    }
  }
}
__global__ void gaussian_noise_kernel(float* __restrict__ out, size_t n, float stddev, uint64_t seed, uint64_t offset) {
  const size_t n4 = (n + 3) / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    curandStatePhilox4_32_10_t st;
    curand_init(seed, i, offset, &st);
    const float4 z = curand_normal4(&st);
    const float vals[4] = {z.x, z.y, z.z, z.w};
    for (int j = 0; j < 4; ++j) {
      const size_t e = i * 4 + j;
      if (e < n) out[e] = vals[j] * stddev;
    }
  }
}

static int grid_for(size_t n, int block = 256) {
  size_t b = (n + block - 1) / block;
  size_t cap = (size_t)num_sms() * 16;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace gr

extern "C" int gr_dense_softmax_fwd_f32(const float* x, const float* drop_mask, const float* Wd, const float* bd,
                                        int R, int Fin, int C, float* logits, float* probs, void* stream) {
  using namespace gr;
  if (!x || !Wd || !bd || (!logits && !probs)) return set_error(GR_EINVAL, "dense_fwd: null pointer");
  if (R <= 0 || Fin <= 0 || C <= 0 || C > kCMax) return set_error(GR_EINVAL, "dense_fwd: bad shape (C <= 64)");
  const size_t smem = (size_t)Fin * (C | 1) * sizeof(float);
  if (smem > 220 * 1024) return set_error(GR_EUNSUPPORTED, "dense_fwd: Fin*C too large for shared memory");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool rowk = (Fin % 4) == 0 && C <= 48 && !getenv("GR_DENSE_WARP") &&
                    ((reinterpret_cast<uintptr_t>(x) | (drop_mask ? reinterpret_cast<uintptr_t>(drop_mask) : 0)) & 15) == 0;
  if (rowk) {
    const int CP = C <= 24 ? 24 : 48;
    const size_t sm2 = ((size_t)Fin * CP + CP) * sizeof(float);
    if (sm2 <= 200 * 1024) {
      const int g2 = min((R + 127) / 128, num_sms() * 8);
      if (CP == 24) {
        GR_CUDA(cudaFuncSetAttribute(dense_softmax_fwd_row_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
        dense_softmax_fwd_row_kernel<24><<<g2, 128, sm2, s>>>(x, drop_mask, Wd, bd, R, Fin, C, logits, probs);
      } else {
        GR_CUDA(cudaFuncSetAttribute(dense_softmax_fwd_row_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
        dense_softmax_fwd_row_kernel<48><<<g2, 128, sm2, s>>>(x, drop_mask, Wd, bd, R, Fin, C, logits, probs);
      }
      GR_CHECK_LAUNCH("dense_softmax_fwd_row_kernel");
      return GR_OK;
    }
  }
  const int grid = min((R + 7) / 8, num_sms() * (smem > 100 * 1024 ? 1 : 2));
  if (C <= 32) {
    GR_CUDA(cudaFuncSetAttribute(dense_softmax_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dense_softmax_fwd_kernel<32><<<grid, 256, smem, s>>>(x, drop_mask, Wd, bd, R, Fin, C, logits, probs);
  } else {
    GR_CUDA(cudaFuncSetAttribute(dense_softmax_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dense_softmax_fwd_kernel<64><<<grid, 256, smem, s>>>(x, drop_mask, Wd, bd, R, Fin, C, logits, probs);
  }
  GR_CHECK_LAUNCH("dense_softmax_fwd_kernel");
  return GR_OK;
}

extern "C" int gr_dense_bwd_f32(const float* x, const float* drop_mask, const float* Wd, const float* g_logits,
                                int R, int Fin, int C, float* dWd, float* dbd, float* dx, void* stream) {
  using namespace gr;
  if (!x || !Wd || !g_logits || !dWd || !dbd) return set_error(GR_EINVAL, "dense_bwd: null pointer");
  if (R <= 0 || Fin <= 0 || C <= 0 || C > kCMax) return set_error(GR_EINVAL, "dense_bwd: bad shape (C <= 64)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GR_CUDA(cudaMemsetAsync(dWd, 0, (size_t)Fin * C * sizeof(float), s));
  GR_CUDA(cudaMemsetAsync(dbd, 0, (size_t)C * sizeof(float), s));
  const int ctas = min((R + 31) / 32, num_sms() * 4);
  int rows_per_cta = (R + ctas - 1) / ctas;
  rows_per_cta = (rows_per_cta + 31) / 32 * 32;
  const int grid = (R + rows_per_cta - 1) / rows_per_cta;
  if (C <= 24) dense_bwd_kernel<24><<<grid, 256, 0, s>>>(x, drop_mask, Wd, g_logits, R, Fin, C, rows_per_cta, dWd, dbd, dx);
  else if (C <= 32) dense_bwd_kernel<32><<<grid, 256, 0, s>>>(x, drop_mask, Wd, g_logits, R, Fin, C, rows_per_cta, dWd, dbd, dx);
  else if (C <= 44) dense_bwd_kernel<44><<<grid, 256, 0, s>>>(x, drop_mask, Wd, g_logits, R, Fin, C, rows_per_cta, dWd, dbd, dx);
  else dense_bwd_kernel<64><<<grid, 256, 0, s>>>(x, drop_mask, Wd, g_logits, R, Fin, C, rows_per_cta, dWd, dbd, dx);
  GR_CHECK_LAUNCH("dense_bwd_kernel");
  return GR_OK;
}

extern "C" int gr_colsum_f32(const float* a, int R, int N, int lda, float* out, void* stream) {
  using namespace gr;
  if (!a || !out || R <= 0 || N <= 0 || lda < N) return set_error(GR_EINVAL, "colsum: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GR_CUDA(cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), s));
  const int gx = (N + 127) / 128;
  int gy = max(1, (num_sms() * 8) / gx);
  int rows_per_cta = (R + gy - 1) / gy;
  if (rows_per_cta < 16) rows_per_cta = 16;
  gy = (R + rows_per_cta - 1) / rows_per_cta;
  colsum_kernel<<<dim3(gx, gy), 128, 0, s>>>(a, R, N, lda, rows_per_cta, out);
  GR_CHECK_LAUNCH("colsum_kernel");
  return GR_OK;
}

extern "C" int gr_add_f32(const float* a, const float* b, float* out, size_t n, void* stream) {
  using namespace gr;
  if (!a || !b || !out || n == 0) return set_error(GR_EINVAL, "add: bad argument");
  const bool al = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  const size_t n4 = al ? n / 4 : 0;
  add_kernel<<<grid_for(n4 ? n4 : 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), reinterpret_cast<float4*>(out), n4, a, b, out,
      n4 * 4, n);
  GR_CHECK_LAUNCH("add_kernel");
  return GR_OK;
}

extern "C" int gr_add_into_f32(const float* a, const float* b, float* out, size_t rows, int cols, int ldo, void* stream) {
  using namespace gr;
  if (!a || !b || !out || rows == 0 || cols <= 0 || ldo < cols) return set_error(GR_EINVAL, "add_into: bad argument");
  if ((cols % 4) || (ldo % 4) || ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15))
    return set_error(GR_EUNSUPPORTED, "add_into: cols, ldo and the pointers must be 16-byte aligned");
  add_into_kernel<<<grid_for(rows * (size_t)(cols / 4)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), out, rows, cols / 4, ldo);
  GR_CHECK_LAUNCH("add_into_kernel");
  return GR_OK;
}

extern "C" int gr_mask_mul_acc_f32(float* out, const float* tmp, const float* mask, int rows_per_seq, int R, int K,
                                   int accumulate, void* stream) {
  using namespace gr;
  if (!out || !tmp || !mask || R <= 0 || K <= 0 || rows_per_seq <= 0) return set_error(GR_EINVAL, "mask_mul_acc: bad argument");
  mask_mul_acc_kernel<<<grid_for((size_t)R * K), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, tmp, mask, rows_per_seq, R, K, accumulate);
  GR_CHECK_LAUNCH("mask_mul_acc_kernel");
  return GR_OK;
}

extern "C" int gr_concat2_f32(const float* a, int Fa, const float* b, int Fb, float* out, size_t rows, void* stream) {
  using namespace gr;
  if (!a || !b || !out || Fa <= 0 || Fb <= 0 || rows == 0) return set_error(GR_EINVAL, "concat2: bad argument");
  concat2_kernel<<<grid_for(rows * (Fa + Fb)), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, Fa, b, Fb, out, rows);
  GR_CHECK_LAUNCH("concat2_kernel");
  return GR_OK;
}

extern "C" int gr_adam_step_f32(float* param, const float* grad, float* m, float* v, size_t n, int rows, int cols,
                                float lr, float beta1, float beta2, float eps, float decay, float clipvalue,
                                float max_norm, int64_t step, void* stream) {
  using namespace gr;
  if (!param || !grad || !m || !v || n == 0) return set_error(GR_EINVAL, "adam: bad argument");
  if (max_norm > 0.f && (size_t)rows * cols != n) return set_error(GR_EINVAL, "adam: rows*cols != n for maxnorm");
  // Keras 2.1.4 Adam.get_updates: lr *= 1/(1+decay*iterations); t = iterations+1;
  // lr_t = lr*sqrt(1-beta2^t)/(1-beta1^t)
  double lr_d = lr;
  if (decay > 0.f) lr_d *= 1.0 / (1.0 + (double)decay * (double)step);
  const double t = (double)step + 1.0;
  const float lr_t = (float)(lr_d * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t)));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  adam_kernel<<<grid_for(n), 256, 0, s>>>(param, grad, m, v, n, lr_t, beta1, beta2, eps, clipvalue);
  GR_CHECK_LAUNCH("adam_kernel");
  if (max_norm > 0.f) {
    maxnorm_kernel<<<(cols + 31) / 32, dim3(32, 8), 0, s>>>(param, rows, cols, max_norm);
    GR_CHECK_LAUNCH("maxnorm_kernel");
  }
  return GR_OK;
}

// ------------------------------------------------------------------ multi-tensor optimiser epilogue
// One launch over the flat gradient bucket (behind the all-reduce): clipvalue + Adam for every trainable tensor, the
// parameter pointers travelling in the kernel arguments (multimodal.py:206-208: five tensors in the fusion model).
namespace gr {
struct MtTable {
  float* ptr[GR_MT_MAX];
  unsigned long long off[GR_MT_MAX + 1];   // element offsets of the tensors inside the flat buffers
  int n;
};
__global__ void pack_kernel(MtTable tab, float* __restrict__ flat) {
  const size_t total = tab.off[tab.n];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int t = 0;
    while (t + 1 < tab.n && i >= tab.off[t + 1]) ++t;
    flat[i] = tab.ptr[t][i - tab.off[t]];
  }
}
__global__ void adam_flat_kernel(MtTable tab, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                 float lr_t, float beta1, float beta2, float eps, float clip) {
  const size_t total = tab.off[tab.n];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int t = 0;
    while (t + 1 < tab.n && i >= tab.off[t + 1]) ++t;
    float gi = g[i];
    if (clip > 0.f) gi = fminf(fmaxf(gi, -clip), clip);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float* p = tab.ptr[t] + (i - tab.off[t]);
    *p = *p - lr_t * mi / (sqrtf(vi) + eps);
  }
}
static int fill_table(MtTable* tab, float* const* ptrs, const size_t* sizes, int n) {
  if (!ptrs || !sizes || n <= 0 || n > GR_MT_MAX) return set_error(GR_EINVAL, "multi-tensor: 1..GR_MT_MAX tensors");
  tab->n = n;
  tab->off[0] = 0;
  for (int t = 0; t < n; ++t) {
    if (!ptrs[t] || sizes[t] == 0) return set_error(GR_EINVAL, "multi-tensor: null or empty tensor");
    tab->ptr[t] = ptrs[t];
    tab->off[t + 1] = tab->off[t] + sizes[t];
  }
  return GR_OK;
}
}  // namespace gr

extern "C" int gr_pack_f32(const float* const* srcs, const size_t* sizes, int n_tensors, float* flat, void* stream) {
  using namespace gr;
  MtTable tab;
  int rc = fill_table(&tab, const_cast<float* const*>(reinterpret_cast<const float* const*>(srcs)), sizes, n_tensors);
  if (rc != GR_OK) return rc;
  if (!flat) return set_error(GR_EINVAL, "pack: null pointer");
  pack_kernel<<<grid_for(tab.off[tab.n]), 256, 0, static_cast<cudaStream_t>(stream)>>>(tab, flat);
  GR_CHECK_LAUNCH("pack_kernel");
  return GR_OK;
}

extern "C" int gr_adam_flat_f32(float* const* params, const size_t* sizes, int n_tensors, const float* flat_grad,
                                float* flat_m, float* flat_v, float lr, float beta1, float beta2, float eps, float decay,
                                float clipvalue, int64_t step, void* stream) {
  using namespace gr;
  MtTable tab;
  int rc = fill_table(&tab, params, sizes, n_tensors);
  if (rc != GR_OK) return rc;
  if (!flat_grad || !flat_m || !flat_v) return set_error(GR_EINVAL, "adam_flat: null pointer");
  double lr_d = lr;                       // Keras 2.1.4 Adam.get_updates (see gr_adam_step_f32)
  if (decay > 0.f) lr_d *= 1.0 / (1.0 + (double)decay * (double)step);
  const double t = (double)step + 1.0;
  const float lr_t = (float)(lr_d * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t)));
  adam_flat_kernel<<<grid_for(tab.off[tab.n]), 256, 0, static_cast<cudaStream_t>(stream)>>>(tab, flat_grad, flat_m, flat_v, lr_t,
                                                                                            beta1, beta2, eps, clipvalue);
  GR_CHECK_LAUNCH("adam_flat_kernel");
  return GR_OK;
}

extern "C" int gr_maxnorm_f32(float* w, int rows, int cols, float max_norm, void* stream) {
  using namespace gr;
  if (!w || rows <= 0 || cols <= 0 || max_norm <= 0.f) return set_error(GR_EINVAL, "maxnorm: bad argument");
  maxnorm_kernel<<<(cols + 31) / 32, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(w, rows, cols, max_norm);
  GR_CHECK_LAUNCH("maxnorm_kernel");
  return GR_OK;
}

extern "C" int gr_dropout_mask_f32(float* out, size_t n, float p, uint64_t seed, uint64_t offset, void* stream) {
  using namespace gr;
  if (!out || n == 0 || p < 0.f || p >= 1.f) return set_error(GR_EINVAL, "dropout_mask: bad argument");
  dropout_mask_kernel<<<grid_for((n + 3) / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, n, p, 1.0f / (1.0f - p), seed, offset);
  GR_CHECK_LAUNCH("dropout_mask_kernel");
  return GR_OK;
}

extern "C" int gr_gaussian_noise_f32(float* out, size_t n, float stddev, uint64_t seed, uint64_t offset, void* stream) {
  using namespace gr;
  if (!out || n == 0) return set_error(GR_EINVAL, "gaussian_noise: bad argument");
  gaussian_noise_kernel<<<grid_for((n + 3) / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, n, stddev, seed, offset);
  GR_CHECK_LAUNCH("gaussian_noise_kernel");
  return GR_OK;
}
