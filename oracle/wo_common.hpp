// wo_common.hpp -- shared pieces of the CPU oracle.  TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// F16 helpers, the two F16 lookup tables ggml v1.0.3 evaluates GELU and exp through, the
// dot-product GEMM every galois_* op with a contraction reduces to, and the model container
// filled by the loader restatement (wo_loader.cpp).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "oracle.h"

namespace wo {

// ---- F16 (galois::F16 / ggml_fp16_t): IEEE binary16, round-to-nearest-even ----------------
static inline uint16_t f32_to_f16_bits(float x) {
  _Float16 h = (_Float16)x;
  uint16_t b;
  std::memcpy(&b, &h, 2);
  return b;
}
static inline float f16_bits_to_f32(uint16_t b) {
  _Float16 h;
  std::memcpy(&h, &b, 2);
  return (float)h;
}
static inline float f16_round(float x) { return (float)(_Float16)x; }

// ---- ggml-sem lookup tables (galois_gelu 1777, galois_flash_attn 1795)
struct Luts {
  std::vector<uint16_t> gelu;  // table_gelu_f16[i] = F16(gelu_tanh(F32(i)))
  std::vector<uint16_t> exp;   // table_exp_f16[i]  = F16(expf(F32(i)))
  Luts();
};
const Luts& luts();

static inline float gelu_tanh_f32(float x) {
  const float GELU_COEF_A = 0.044715f;
  const float SQRT_2_OVER_PI = 0.79788456080286535587989211986876f;
  return 0.5f * x * (1.0f + tanhf(SQRT_2_OVER_PI * x * (1.0f + GELU_COEF_A * x * x)));
}
static inline float gelu_lut(float x) { return f16_bits_to_f32(luts().gelu[f32_to_f16_bits(x)]); }
static inline float exp_lut(float x) { return f16_bits_to_f32(luts().exp[f32_to_f16_bits(x)]); }

// ---- options (switchable rounding points) -------------------------------------------------
struct Options {
  int act_f16_round = 1;
  int gelu_mode = 0;
  int softmax_exp = 0;
  int prob_f16_round = 1;
};

// ---- model container ----------------------------------------------------------------------
struct Tensor {
  std::vector<int> ne;         // ggml order: ne[0] innermost
  bool f16 = false;
  std::vector<uint8_t> data;   // raw bytes as in the file
  bool loaded = false;
  size_t nelem() const {
    size_t n = 1;
    for (int v : ne) n *= (size_t)v;
    return n;
  }
  const float* f32() const { return reinterpret_cast<const float*>(data.data()); }
  const uint16_t* h() const { return reinterpret_cast<const uint16_t*>(data.data()); }
  // weight rows as f32 (F16 -> F32 is exact)
  void to_f32(std::vector<float>& out) const;
  // the same, converted once and kept (a decode step walks every decoder weight: converting them at every
  // call made the oracle's per-token time ~10x its arithmetic).  Called from the driving thread only.
  const float* f32_rows() const {
    if (!f16) return f32();
    if (f32_cache.empty()) to_f32(f32_cache);
    return f32_cache.data();
  }
  mutable std::vector<float> f32_cache;
};

struct HParams {  // src/main.rs:607-619
  int32_t n_vocab, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer;
  int32_t n_text_ctx, n_text_state, n_text_head, n_text_layer, n_mels, f16;
};

struct Vocab {  // src/main.rs:541-575 (ids only; token text is out of scope for parity)
  int32_t n_vocab = 51864;
  int32_t token_eot = 50256, token_sot = 50257, token_prev = 50360, token_solm = 50361;
  int32_t token_not = 50362, token_beg = 50363, token_translate = 50358, token_transcribe = 50359;
  std::vector<std::string> id_to_token;
};

struct Model {
  HParams hp{};
  int filt_n_mel = 0, filt_n_fft = 0;
  std::vector<float> filters;            // [n_mel][n_fft]
  Vocab vocab;
  std::map<std::string, Tensor> t;       // name -> tensor (table of src/main.rs:960-1334)
  const Tensor& get(const std::string& name) const;
};

int load_model(const char* path, Model& m, std::string& err);   // wo_loader.cpp

// ---- contraction kernel ---------------------------------------------------------------------
// C[i][j] = sum_k A[i][k] * B[j][k]   (both operands contiguous along k; f32 accumulate)
// This is galois_matmul's contract (`dst[m,n] = sum_k a[k,m]*b[k,n]`, src/main.rs:1752-1767)
// with A = activations (rows = tokens) and B = weight rows.
void gemm_nt(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N,
             int K, int n_threads);

void round_f16_inplace(float* x, size_t n);

// static-chunked parallel loop over [0, n) on std::thread workers (no OpenMP: the image's gcc
// wrapper cannot link libgomp)
template <class F>
static inline void parallel_for(int n, int n_threads, F&& body) {
  if (n_threads <= 1 || n <= 1) {
    for (int i = 0; i < n; ++i) body(i);
    return;
  }
  const int nt = n_threads < n ? n_threads : n;
  std::vector<std::thread> th;
  th.reserve(nt);
  for (int w = 0; w < nt; ++w) {
    th.emplace_back([&, w]() {
      const int lo = (int)((long long)n * w / nt), hi = (int)((long long)n * (w + 1) / nt);
      for (int i = lo; i < hi; ++i) body(i);
    });
  }
  for (auto& t : th) t.join();
}

}  // namespace wo

// ---- context ------------------------------------------------------------------------------
struct orc_ctx {
  wo::Model model;
  wo::Options opt;
  // mel (WhisperMel, src/main.rs:733-748)
  int mel_n_mel = 0, mel_n_len = 0;
  std::vector<float> mel;                 // [n_mel][n_len]
  // exp_n_audio_ctx (src/main.rs:362, read at 1803-1807): > 0 shortens the audio context of whisper_encode
  int exp_n_audio_ctx = 0;
  int enc_n_ctx = 0;                      // the n_ctx the last encode ran with (rows of enc_out / cross K,V)
  // encoder results
  std::vector<float> enc_out;             // ln_post output [n_ctx][d]
  std::vector<uint16_t> cross_k, cross_v; // [L_text][n_ctx][d] F16 (memory_cross_k/v, 1350-1354)
  std::vector<float> cross_kf, cross_vf;  // the same values widened to f32 once per encode (decoder's operand form)
  std::vector<uint16_t> mem_k, mem_v;     // [L_text][n_text_ctx][d] F16 (memory_k/v, 1346-1347)
  std::map<int, double> chk;              // stage*1000+layer -> sum|x|
  std::vector<float> logits;              // [n_vocab] of the last decoded position
};

namespace wo {
int pcm_to_mel(orc_ctx* ctx, const float* pcm, size_t n, int n_threads);   // wo_mel.cpp
int encode(orc_ctx* ctx, int n_threads, size_t mel_offset);                // wo_encoder.cpp
int decode(orc_ctx* ctx, const int32_t* tokens, int n_tokens, int n_past, int n_threads);  // wo_decoder.cpp
void fft(const std::vector<float>& in, std::vector<float>& out);
void dft(const std::vector<float>& in, std::vector<float>& out);
// shared by encoder + decoder
void layer_norm(const float* x, int T, int d, const float* w, const float* b, float* y);
void linear(const orc_ctx* ctx, const float* x, int T, int K, const Tensor& W, const Tensor* bias,
            float* y, int n_threads);
void gelu_inplace(const orc_ctx* ctx, float* x, size_t n);
double abs_sum(const float* x, size_t n);
}  // namespace wo
