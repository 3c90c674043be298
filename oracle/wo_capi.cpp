// wo_capi.cpp -- C API of the CPU oracle (see oracle.h).  TEST INFRASTRUCTURE ONLY.
#include <algorithm>

#include "wo_common.hpp"

static thread_local std::string g_err;

extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }

int orc_ctx_create(const char* model_path, orc_ctx** out) {
  if (!model_path || !out) return ORC_ERR_UNEXPECTED;
  orc_ctx* c = new orc_ctx();
  int rc = wo::load_model(model_path, c->model, g_err);
  if (rc != ORC_OK) {
    delete c;
    *out = nullptr;
    return rc;
  }
  *out = c;
  return ORC_OK;
}

void orc_ctx_free(orc_ctx* ctx) { delete ctx; }

int orc_get_hparams(const orc_ctx* ctx, int32_t out[11]) {
  std::memcpy(out, &ctx->model.hp, 44);
  return ORC_OK;
}

int orc_get_special_tokens(const orc_ctx* ctx, int32_t out[8]) {
  const wo::Vocab& v = ctx->model.vocab;
  out[0] = v.token_eot; out[1] = v.token_sot; out[2] = v.token_prev; out[3] = v.token_solm;
  out[4] = v.token_not; out[5] = v.token_beg; out[6] = v.token_translate; out[7] = v.token_transcribe;
  return ORC_OK;
}

int orc_set_option(orc_ctx* ctx, int opt, int value) {
  switch (opt) {
    case ORC_OPT_ACT_F16_ROUND: ctx->opt.act_f16_round = value; break;
    case ORC_OPT_GELU_MODE: ctx->opt.gelu_mode = value; break;
    case ORC_OPT_SOFTMAX_EXP: ctx->opt.softmax_exp = value; break;
    case ORC_OPT_PROB_F16_ROUND: ctx->opt.prob_f16_round = value; break;
    default: return ORC_ERR_UNEXPECTED;
  }
  return ORC_OK;
}

int orc_set_audio_ctx(orc_ctx* ctx, int n_ctx) {   // exp_n_audio_ctx (src/main.rs:362); 0 = the model's
  if (n_ctx < 0 || n_ctx > ctx->model.hp.n_audio_ctx) return ORC_ERR_UNEXPECTED;
  ctx->exp_n_audio_ctx = n_ctx;
  return ORC_OK;
}

int orc_pcm_to_mel(orc_ctx* ctx, const float* pcm, size_t n_samples, int n_threads) {
  return wo::pcm_to_mel(ctx, pcm, n_samples, n_threads);
}

int orc_mel_dims(const orc_ctx* ctx, int* n_mel, int* n_len) {
  *n_mel = ctx->mel_n_mel;
  *n_len = ctx->mel_n_len;
  return ORC_OK;
}

int orc_mel_read(const orc_ctx* ctx, float* out) {
  std::memcpy(out, ctx->mel.data(), ctx->mel.size() * 4);
  return ORC_OK;
}

int orc_mel_set(orc_ctx* ctx, const float* mel, int n_mel, int n_len) {
  ctx->mel_n_mel = n_mel;
  ctx->mel_n_len = n_len;
  ctx->mel.assign(mel, mel + (size_t)n_mel * n_len);
  return ORC_OK;
}

int orc_encode(orc_ctx* ctx, int n_threads, size_t mel_offset) { return wo::encode(ctx, n_threads, mel_offset); }

int orc_encoder_out_read(const orc_ctx* ctx, float* out) {
  if (ctx->enc_out.empty()) return ORC_ERR_UNEXPECTED;
  std::memcpy(out, ctx->enc_out.data(), ctx->enc_out.size() * 4);
  return ORC_OK;
}

int orc_cross_kv_read(const orc_ctx* ctx, int layer, uint16_t* k, uint16_t* v) {
  const auto& hp = ctx->model.hp;
  const size_t n = (size_t)(ctx->enc_n_ctx > 0 ? ctx->enc_n_ctx : hp.n_audio_ctx) * hp.n_text_state;
  if (layer < 0 || layer >= hp.n_text_layer || ctx->cross_k.size() < n * (layer + 1)) return ORC_ERR_UNEXPECTED;
  if (k) std::memcpy(k, ctx->cross_k.data() + n * layer, n * 2);
  if (v) std::memcpy(v, ctx->cross_v.data() + n * layer, n * 2);
  return ORC_OK;
}

int orc_checksum(const orc_ctx* ctx, int stage, int layer, double* abs_sum) {
  auto it = ctx->chk.find(stage * 1000 + layer);
  if (it == ctx->chk.end()) return ORC_ERR_UNEXPECTED;
  *abs_sum = it->second;
  return ORC_OK;
}

int orc_decode(orc_ctx* ctx, const int32_t* tokens, int n_tokens, int n_past, int n_threads) {
  return wo::decode(ctx, tokens, n_tokens, n_past, n_threads);
}

int orc_logits_read(const orc_ctx* ctx, float* out) {
  if (ctx->logits.empty()) return ORC_ERR_UNEXPECTED;
  std::memcpy(out, ctx->logits.data(), ctx->logits.size() * 4);
  return ORC_OK;
}

// D6: greedy = plain arg-max over all n_vocab logits (first index wins ties); prompt given by
// the caller; stops after `eot` or max_new tokens or when the text context is full.  The oracle
// DEFINES this rule -- the reference has no sampler.
int orc_decode_greedy(orc_ctx* ctx, const int32_t* prompt, int n_prompt, int max_new, int eot,
                      int n_threads, int32_t* out_tokens, float* out_margin, int* out_len) {
  const int n_text_ctx = ctx->model.hp.n_text_ctx;
  int n_past = 0, n_out = 0;
  std::vector<int32_t> feed(prompt, prompt + n_prompt);
  while (n_out < max_new) {
    if (n_past + (int)feed.size() > n_text_ctx) break;
    int rc = wo::decode(ctx, feed.data(), (int)feed.size(), n_past, n_threads);
    if (rc != ORC_OK) return rc;
    n_past += (int)feed.size();
    const std::vector<float>& lg = ctx->logits;
    int best = 0;
    float b1 = -INFINITY, b2 = -INFINITY;
    for (int i = 0; i < (int)lg.size(); ++i) {
      if (lg[i] > b1) { b2 = b1; b1 = lg[i]; best = i; }
      else if (lg[i] > b2) b2 = lg[i];
    }
    out_tokens[n_out] = best;
    if (out_margin) out_margin[n_out] = b1 - b2;
    ++n_out;
    if (best == eot) break;
    feed.assign(1, best);
  }
  *out_len = n_out;
  return ORC_OK;
}

void orc_fft(const float* in, int n, float* out) {
  std::vector<float> i(in, in + n), o;
  wo::fft(i, o);
  std::memcpy(out, o.data(), o.size() * 4);
}
void orc_dft(const float* in, int n, float* out) {
  std::vector<float> i(in, in + n), o;
  wo::dft(i, o);
  std::memcpy(out, o.data(), o.size() * 4);
}
float orc_f16_round(float x) { return wo::f16_round(x); }
float orc_gelu_lut(float x) { return wo::gelu_lut(x); }
float orc_exp_lut(float x) { return wo::exp_lut(x); }

}  // extern "C"
