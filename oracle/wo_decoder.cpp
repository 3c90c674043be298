// wo_decoder.cpp -- the decode step the reference leaves unimplemented.  TEST INFRASTRUCTURE ONLY.
//
// The reference declares the decoder's state and weights (WhisperLayerDecoder src/main.rs:694-731,
// d_pe/d_te/d_ln 788-793, memory_k/v F16 [L*n_text_ctx*d] 1343-1347, logits/probs 351-352) but has
// no whisper_decode.  This file restates upstream whisper.cpp v1.0.3's whisper_decode -- the code
// base main.rs transliterates (paths at 2066-2088) -- constrained by the state the reference's
// encoder leaves behind: cross-K already scaled by (d/H)^-1/4 (1994-1996), no key bias (704, 718),
// F16 KV (1346-1354).  SURVEY.md section 8a rows D1-D6.  PARITY UNPINNED: nothing in the reference
// can check this.
#include "wo_common.hpp"

namespace wo {

void attention(const orc_ctx* ctx, const float* q, int Tq, const float* k, const float* v, int Tk,
               int d, int H, float scale, int causal_past, float* out, int n_threads);

int decode(orc_ctx* ctx, const int32_t* tokens, int N, int n_past, int n_threads) {
  const Model& m = ctx->model;
  const HParams& hp = m.hp;
  const int d = hp.n_text_state, H = hp.n_text_head, L = hp.n_text_layer;
  const int n_ctx = hp.n_text_ctx, M = ctx->enc_n_ctx > 0 ? ctx->enc_n_ctx : hp.n_audio_ctx, n_vocab = hp.n_vocab;
  if (N < 1 || n_past < 0 || n_past + N > n_ctx) return ORC_ERR_NOT_ENOUGH_SPACE;
  if (ctx->cross_k.size() != (size_t)L * M * d) return ORC_ERR_UNEXPECTED;   // encode first
  if (ctx->mem_k.size() != (size_t)L * n_ctx * d) {
    ctx->mem_k.assign((size_t)L * n_ctx * d, 0);
    ctx->mem_v.assign((size_t)L * n_ctx * d, 0);
  }
  const float qk_scale = powf((float)d / (float)H, -0.25f);
  // a single-token step does too little work per op to pay for spawning workers (parallel_for starts fresh
  // threads at every call): the per-layer ops run on the calling thread, only the vocabulary projection is split
  const int nt_all = n_threads;
  if ((size_t)N * d * d < ((size_t)1 << 24)) n_threads = 1;

  // D1: x = d_te[:, tok] + d_pe[:, n_past + i]
  std::vector<float> inpL((size_t)N * d);
  {
    const Tensor& te = m.get("decoder.token_embedding.weight");   // ne = [d, n_vocab]
    const float* pe = m.get("decoder.positional_embedding").f32();
    for (int i = 0; i < N; ++i) {
      const int tok = tokens[i];
      if (tok < 0 || tok >= n_vocab) return ORC_ERR_UNEXPECTED;
      for (int c = 0; c < d; ++c) {
        float e = te.f16 ? f16_bits_to_f32(te.h()[(size_t)tok * d + c]) : te.f32()[(size_t)tok * d + c];
        inpL[(size_t)i * d + c] = e + pe[(size_t)(n_past + i) * d + c];
      }
    }
  }
  std::vector<float> cur((size_t)N * d), q((size_t)N * d), kk((size_t)N * d), vv((size_t)N * d);
  std::vector<float> att((size_t)N * d), inpCA((size_t)N * d), inpFF((size_t)N * d), hid((size_t)N * 4 * d);
  for (int il = 0; il < L; ++il) {
    const std::string p = "decoder.blocks." + std::to_string(il) + ".";
    // D2: self-attention
    layer_norm(inpL.data(), N, d, m.get(p + "attn_ln.weight").f32(), m.get(p + "attn_ln.bias").f32(), cur.data());
    linear(ctx, cur.data(), N, d, m.get(p + "attn.query.weight"), &m.get(p + "attn.query.bias"), q.data(), n_threads);
    for (auto& x : q) x *= qk_scale;
    linear(ctx, cur.data(), N, d, m.get(p + "attn.key.weight"), nullptr, kk.data(), n_threads);
    for (auto& x : kk) x *= qk_scale;
    linear(ctx, cur.data(), N, d, m.get(p + "attn.value.weight"), &m.get(p + "attn.value.bias"), vv.data(), n_threads);
    {   // append K,V as F16 at row il*n_ctx + n_past
      uint16_t* kd = ctx->mem_k.data() + ((size_t)il * n_ctx + n_past) * d;
      uint16_t* vd = ctx->mem_v.data() + ((size_t)il * n_ctx + n_past) * d;
      for (size_t i = 0; i < (size_t)N * d; ++i) {
        kd[i] = f32_to_f16_bits(kk[i]);
        vd[i] = f32_to_f16_bits(vv[i]);
      }
    }
    const int Tk = n_past + N;
    std::vector<float> Kf((size_t)Tk * d), Vf((size_t)Tk * d);
    {
      const uint16_t* ks = ctx->mem_k.data() + (size_t)il * n_ctx * d;
      const uint16_t* vs = ctx->mem_v.data() + (size_t)il * n_ctx * d;
      for (size_t i = 0; i < Kf.size(); ++i) {
        Kf[i] = f16_bits_to_f32(ks[i]);
        Vf[i] = f16_bits_to_f32(vs[i]);
      }
    }
    if (ctx->opt.act_f16_round) round_f16_inplace(q.data(), q.size());   // mul_mat(K f16, Q) rounds Q to F16
    attention(ctx, q.data(), N, Kf.data(), Vf.data(), Tk, d, H, 1.0f, n_past, att.data(), n_threads);
    linear(ctx, att.data(), N, d, m.get(p + "attn.out.weight"), &m.get(p + "attn.out.bias"), cur.data(), n_threads);
    for (size_t i = 0; i < inpCA.size(); ++i) inpCA[i] = cur[i] + inpL[i];
    // D3: cross-attention over the encoder memory
    layer_norm(inpCA.data(), N, d, m.get(p + "cross_attn_ln.weight").f32(), m.get(p + "cross_attn_ln.bias").f32(), cur.data());
    linear(ctx, cur.data(), N, d, m.get(p + "cross_attn.query.weight"), &m.get(p + "cross_attn.query.bias"), q.data(), n_threads);
    for (auto& x : q) x *= qk_scale;
    if (ctx->opt.act_f16_round) round_f16_inplace(q.data(), q.size());
    if (ctx->cross_kf.size() != ctx->cross_k.size()) {   // F16 -> f32 once per encode (encode clears the copies)
      ctx->cross_kf.resize(ctx->cross_k.size());
      ctx->cross_vf.resize(ctx->cross_v.size());
      for (size_t i = 0; i < ctx->cross_k.size(); ++i) {
        ctx->cross_kf[i] = f16_bits_to_f32(ctx->cross_k[i]);
        ctx->cross_vf[i] = f16_bits_to_f32(ctx->cross_v[i]);
      }
    }
    const float* Kc = ctx->cross_kf.data() + (size_t)il * M * d;
    const float* Vc = ctx->cross_vf.data() + (size_t)il * M * d;
    attention(ctx, q.data(), N, Kc, Vc, M, d, H, 1.0f, -1, att.data(), n_threads);
    linear(ctx, att.data(), N, d, m.get(p + "cross_attn.out.weight"), &m.get(p + "cross_attn.out.bias"), cur.data(), n_threads);
    for (size_t i = 0; i < inpFF.size(); ++i) inpFF[i] = cur[i] + inpCA[i];
    // D4: MLP
    layer_norm(inpFF.data(), N, d, m.get(p + "mlp_ln.weight").f32(), m.get(p + "mlp_ln.bias").f32(), cur.data());
    linear(ctx, cur.data(), N, d, m.get(p + "mlp.0.weight"), &m.get(p + "mlp.0.bias"), hid.data(), n_threads);
    gelu_inplace(ctx, hid.data(), hid.size());
    linear(ctx, hid.data(), N, 4 * d, m.get(p + "mlp.2.weight"), &m.get(p + "mlp.2.bias"), cur.data(), n_threads);
    for (size_t i = 0; i < inpL.size(); ++i) inpL[i] = cur[i] + inpFF[i];
  }
  // D5: logits of the last position = d_te^T . LN_{decoder.ln}(x)
  std::vector<float> last(d);
  layer_norm(inpL.data() + (size_t)(N - 1) * d, 1, d, m.get("decoder.ln.weight").f32(), m.get("decoder.ln.bias").f32(), last.data());
  ctx->logits.resize(n_vocab);
  linear(ctx, last.data(), 1, d, m.get("decoder.token_embedding.weight"), nullptr, ctx->logits.data(), nt_all);
  return ORC_OK;
}

}  // namespace wo
