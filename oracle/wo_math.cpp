// wo_math.cpp -- the galois_* op semantics the oracle assumes.  TEST INFRASTRUCTURE ONLY.
//
// `galois` 0.1.0 (Cargo.toml:13) is absent, so each op below restates the identically named
// ggml op of whisper.cpp v1.0.3; call sites cited per function.
#include <algorithm>

#include "wo_common.hpp"

namespace wo {

Luts::Luts() : gelu(65536), exp(65536) {
  for (uint32_t i = 0; i < 65536; ++i) {
    float f = f16_bits_to_f32((uint16_t)i);
    gelu[i] = f32_to_f16_bits(gelu_tanh_f32(f));
    exp[i] = f32_to_f16_bits(expf(f));
  }
}
const Luts& luts() {
  static Luts l;
  return l;
}

void round_f16_inplace(float* x, size_t n) {
  for (size_t i = 0; i < n; ++i) x[i] = f16_round(x[i]);
}

double abs_sum(const float* x, size_t n) {
  double s = 0.0;
  for (size_t i = 0; i < n; ++i) s += std::fabs((double)x[i]);
  return s;
}

// ---- dot-product GEMM ----------------------------------------------------------------------
typedef float v8f __attribute__((vector_size(32)));
static inline v8f ld8(const float* p) {
  v8f v;
  std::memcpy(&v, p, 32);
  return v;
}
static inline float hsum8(v8f v) {
  return ((v[0] + v[4]) + (v[2] + v[6])) + ((v[1] + v[5]) + (v[3] + v[7]));
}

// 2 rows of A x 4 rows of B, f32 lanes (ggml_vec_dot_f16 accumulates in f32 SIMD lanes too)
static inline void micro_2x4(const float* a0, const float* a1, const float* b0, const float* b1,
                             const float* b2, const float* b3, int K, float* c0, float* c1) {
  v8f acc[2][4] = {};
  int k = 0;
  for (; k + 8 <= K; k += 8) {
    v8f x0 = ld8(a0 + k), x1 = ld8(a1 + k);
    v8f w0 = ld8(b0 + k), w1 = ld8(b1 + k), w2 = ld8(b2 + k), w3 = ld8(b3 + k);
    acc[0][0] += x0 * w0; acc[0][1] += x0 * w1; acc[0][2] += x0 * w2; acc[0][3] += x0 * w3;
    acc[1][0] += x1 * w0; acc[1][1] += x1 * w1; acc[1][2] += x1 * w2; acc[1][3] += x1 * w3;
  }
  float r[2][4];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 4; ++j) r[i][j] = hsum8(acc[i][j]);
  const float* bs[4] = {b0, b1, b2, b3};
  for (; k < K; ++k) {
    for (int j = 0; j < 4; ++j) {
      r[0][j] += a0[k] * bs[j][k];
      r[1][j] += a1[k] * bs[j][k];
    }
  }
  for (int j = 0; j < 4; ++j) {
    c0[j] = r[0][j];
    c1[j] = r[1][j];
  }
}

static inline float dot1(const float* a, const float* b, int K) {
  v8f acc = {};
  int k = 0;
  for (; k + 8 <= K; k += 8) acc += ld8(a + k) * ld8(b + k);
  float r = hsum8(acc);
  for (; k < K; ++k) r += a[k] * b[k];
  return r;
}

void gemm_nt(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N,
             int K, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  const int M2 = M & ~1, N4 = N & ~3;
  // cache blocking only: blocks of IB rows of A (one per worker at a time) x JB rows of B, sized so that a block
  // pair (~(IB + JB) * K * 4 bytes) stays in a core's L2 while it is reused.  Every output element is still
  // one micro_2x4 / dot1 over the whole K in the same order, so the results do not depend on the blocking.
  const int IB = 32;
  int JB = (int)((size_t)(768 * 1024) / ((size_t)K * 4)) & ~3;
  if (JB < 16) JB = 16;
  if (JB > 256) JB = 256;
  const int n_ib = (M2 + IB - 1) / IB;
  parallel_for(n_ib, n_threads, [&](int ib) {
    const int i_lo = ib * IB, i_hi = std::min(M2, i_lo + IB);
    for (int j_lo = 0; j_lo < N4; j_lo += JB) {
      const int j_hi = std::min(N4, j_lo + JB);
      for (int i = i_lo; i < i_hi; i += 2) {
        const float* a0 = A + (size_t)i * lda;
        const float* a1 = a0 + lda;
        float* c0 = C + (size_t)i * ldc;
        float* c1 = c0 + ldc;
        for (int j = j_lo; j < j_hi; j += 4) {
          const float* b0 = B + (size_t)j * ldb;
          micro_2x4(a0, a1, b0, b0 + ldb, b0 + 2 * (size_t)ldb, b0 + 3 * (size_t)ldb, K, c0 + j, c1 + j);
        }
      }
    }
    for (int i = i_lo; i < i_hi; i += 2) {
      const float* a0 = A + (size_t)i * lda;
      const float* a1 = a0 + lda;
      float* c0 = C + (size_t)i * ldc;
      float* c1 = c0 + ldc;
      for (int j = N4; j < N; ++j) {
        c0[j] = dot1(a0, B + (size_t)j * ldb, K);
        c1[j] = dot1(a1, B + (size_t)j * ldb, K);
      }
    }
  });
  if (M2 < M) {
    const float* a0 = A + (size_t)M2 * lda;
    float* c0 = C + (size_t)M2 * ldc;
    const int NB = (N + 63) / 64;
    parallel_for(NB, n_threads, [&](int jb) {
      const int j1 = std::min(N, (jb + 1) * 64);
      for (int j = jb * 64; j < j1; ++j) c0[j] = dot1(a0, B + (size_t)j * ldb, K);
    });
  }
}

// galois_norm (src/main.rs:1781-1785) followed by repeat/mul/add with w and b (1882-1886,
// 1948-1952, 1980-1984).  ggml-sem: eps = 1e-5, mean and variance accumulated in f64.
void layer_norm(const float* x, int T, int d, const float* w, const float* b, float* y) {
  const float eps = 1e-5f;
  for (int t = 0; t < T; ++t) {
    const float* xr = x + (size_t)t * d;
    float* yr = y + (size_t)t * d;
    double mean = 0.0;
    for (int i = 0; i < d; ++i) mean += (double)xr[i];
    mean /= d;
    double sum2 = 0.0;
    for (int i = 0; i < d; ++i) {
      double v = (double)xr[i] - mean;
      yr[i] = (float)v;
      sum2 += v * v;
    }
    const float scale = (float)(1.0 / std::sqrt(sum2 / d + (double)eps));
    for (int i = 0; i < d; ++i) {
      float n = yr[i] * scale;   // norm
      n = w[i] * n;              // mul(repeat(w), cur)
      yr[i] = n + b[i];          // add(.., repeat(b))
    }
  }
}

// galois_matmul (src/main.rs:1752-1767) with a weight as `a` and activations as `b`, then the
// repeat/add bias idiom (e.g. 1891-1893).  ggml-sem: activations rounded to F16 before the
// dot, f32 accumulate.  x: [T][K] token-major; W: ne = [K, N] i.e. rows of K; y: [T][N].
void linear(const orc_ctx* ctx, const float* x, int T, int K, const Tensor& W, const Tensor* bias,
            float* y, int n_threads) {
  const int N = W.ne[1];
  const float* wf = W.f32_rows();
  std::vector<float> xr((size_t)T * K);
  std::memcpy(xr.data(), x, xr.size() * 4);
  if (ctx->opt.act_f16_round && W.f16) round_f16_inplace(xr.data(), xr.size());
  gemm_nt(xr.data(), K, wf, K, y, N, T, N, K, n_threads);
  if (bias) {
    const float* bp = bias->f32();
    for (int t = 0; t < T; ++t) {
      float* yr = y + (size_t)t * N;
      for (int n = 0; n < N; ++n) yr[n] = bp[n] + yr[n];   // add(repeat(b), cur)
    }
  }
}

// galois_gelu (src/main.rs:1775-1779)
void gelu_inplace(const orc_ctx* ctx, float* x, size_t n) {
  switch (ctx->opt.gelu_mode) {
    case 0:
      for (size_t i = 0; i < n; ++i) x[i] = gelu_lut(x[i]);
      break;
    case 1:
      for (size_t i = 0; i < n; ++i) x[i] = gelu_tanh_f32(x[i]);
      break;
    default:
      for (size_t i = 0; i < n; ++i) x[i] = 0.5f * x[i] * (1.0f + erff(x[i] * 0.70710678118654752440f));
      break;
  }
}

}  // namespace wo
