// wo_encoder.cpp -- restatement of whisper_encode (src/main.rs:1799-2063).  TEST INFRASTRUCTURE ONLY.
//
// Dataflow and layout glue follow main.rs line by line (cited below); each galois_* kernel body
// follows the ggml v1.0.3 op of the same name (galois itself is absent).
// Activations are held token-major [t][channel] (the reference's ne0 = channel tensors);
// conv inputs/outputs are held as the reference holds them and noted where they differ.
#include <algorithm>

#include "wo_common.hpp"

namespace wo {

// galois_flash_attn (src/main.rs:1787-1797, call 1922).  q,k: [T][d] f16-valued, head h uses
// columns [h*Dh, (h+1)*Dh); v likewise.  Per (h, n): s_m = (k_m . q_n) * scale; softmax with
// exp through the F16 table on (s - max); sum in f64; p -> F16 before the P.V dot.
// `n_kv_of_q(n)` = number of keys query n may see (causal mask of the decoder; T for the encoder).
void attention(const orc_ctx* ctx, const float* q, int Tq, const float* k, const float* v, int Tk,
               int d, int H, float scale, int causal_past, float* out, int n_threads) {
  const int Dh = d / H;
  if (n_threads < 1) n_threads = 1;
  parallel_for(H, n_threads, [&](int h) {
    // gather head operands contiguous: Qh [Tq][Dh], Kh [Tk][Dh], Vt [Dh][Tk] (the reference
    // stores V transposed, time contiguous: 1914-1920)
    std::vector<float> Qh((size_t)Tq * Dh), Kh((size_t)Tk * Dh), Vt((size_t)Dh * Tk);
    for (int t = 0; t < Tq; ++t)
      for (int c = 0; c < Dh; ++c) Qh[(size_t)t * Dh + c] = q[(size_t)t * d + h * Dh + c];
    for (int t = 0; t < Tk; ++t)
      for (int c = 0; c < Dh; ++c) {
        Kh[(size_t)t * Dh + c] = k[(size_t)t * d + h * Dh + c];
        Vt[(size_t)c * Tk + t] = v[(size_t)t * d + h * Dh + c];
      }
    std::vector<float> S((size_t)Tq * Tk);
    gemm_nt(Qh.data(), Dh, Kh.data(), Dh, S.data(), Tk, Tq, Tk, Dh, 1);
    for (int n = 0; n < Tq; ++n) {
      float* s = S.data() + (size_t)n * Tk;
      const int M = causal_past >= 0 ? std::min(Tk, causal_past + n + 1) : Tk;
      for (int m = 0; m < M; ++m) s[m] *= scale;
      float mx = -INFINITY;
      for (int m = 0; m < M; ++m) mx = std::max(mx, s[m]);
      double sum = 0.0;
      for (int m = 0; m < M; ++m) {
        float val = ctx->opt.softmax_exp == 0 ? exp_lut(s[m] - mx) : expf(s[m] - mx);
        sum += (double)val;
        s[m] = val;
      }
      const float inv = (float)(1.0 / sum);
      for (int m = 0; m < M; ++m) {
        float p = s[m] * inv;
        s[m] = ctx->opt.prob_f16_round ? f16_round(p) : p;
      }
      for (int m = M; m < Tk; ++m) s[m] = 0.0f;   // masked keys (diag_mask_inf -> exp = 0)
    }
    std::vector<float> O((size_t)Tq * Dh);
    gemm_nt(S.data(), Tk, Vt.data(), Tk, O.data(), Dh, Tq, Dh, Tk, 1);
    for (int t = 0; t < Tq; ++t)   // merge heads: permute(0,2,1,3) + cpy -> [d, T] (1924-1929)
      for (int c = 0; c < Dh; ++c) out[(size_t)t * d + h * Dh + c] = O[(size_t)t * Dh + c];
  });
}

// galois_conv_1d_{1s,2s} (src/main.rs:1709-1721): kernel ne = [3, Cin, Cout], src [T][Cin]
// token-major here; dst[t][co] = sum_{k,ci} W[co][ci][k] * src[stride*t + k - 1][ci], zero pad.
// ggml-sem: src rounded to F16, f32 accumulate.
static void conv1d_k3(const orc_ctx* ctx, const float* src, int T, int Cin, const Tensor& W,
                      int stride, float* dst, int n_threads) {
  const int Cout = W.ne[2];
  const int To = T / stride;
  std::vector<float> A((size_t)To * Cin * 3);
  const bool rnd = ctx->opt.act_f16_round && W.f16;
  for (int t = 0; t < To; ++t) {
    float* a = A.data() + (size_t)t * Cin * 3;
    for (int ci = 0; ci < Cin; ++ci)
      for (int k = 0; k < 3; ++k) {
        int ts = stride * t + k - 1;
        float x = (ts >= 0 && ts < T) ? src[(size_t)ts * Cin + ci] : 0.0f;
        a[ci * 3 + k] = rnd ? f16_round(x) : x;
      }
  }
  std::vector<float> wf;
  W.to_f32(wf);   // [Cout][Cin][3] -> rows of Cin*3, same (ci,k) order as A
  gemm_nt(A.data(), Cin * 3, wf.data(), Cin * 3, dst, Cout, To, Cout, Cin * 3, n_threads);
}

static void add_bias_rows(float* y, int T, int N, const float* b) {
  for (int t = 0; t < T; ++t)
    for (int n = 0; n < N; ++n) y[(size_t)t * N + n] = b[n] + y[(size_t)t * N + n];
}

int encode(orc_ctx* ctx, int n_threads, size_t mel_offset) {
  const Model& m = ctx->model;
  const HParams& hp = m.hp;
  // 1803-1807: exp_n_audio_ctx when positive (the reference never sets it: 500), else the model's audio context
  const int n_ctx = ctx->exp_n_audio_ctx > 0 ? ctx->exp_n_audio_ctx : hp.n_audio_ctx;
  if (n_ctx > hp.n_audio_ctx) return ORC_ERR_NOT_ENOUGH_SPACE;   // the positional embedding has n_audio_ctx rows
  ctx->enc_n_ctx = n_ctx;
  const int d = hp.n_audio_state;
  const int H = hp.n_audio_head;
  const int L = hp.n_audio_layer;
  const int n_mels = hp.n_mels;
  if (ctx->mel_n_mel != n_mels) return ORC_ERR_UNEXPECTED;   // assert 1813
  const int Tm = 2 * n_ctx;
  ctx->chk.erase(ctx->chk.upper_bound(ORC_STAGE_MEL * 1000 + 999), ctx->chk.end());

  // E0: mel window (1816-1829).  Reference tensor is [n_mels][2*n_ctx] (time contiguous);
  // held here token-major [t][n_mels].
  std::vector<float> mel((size_t)Tm * n_mels, 0.0f);
  {
    const size_t n_len = (size_t)ctx->mel_n_len;
    const size_t i0 = std::min(mel_offset, n_len);
    const size_t i1 = std::min(mel_offset + (size_t)Tm, n_len);
    for (int j = 0; j < n_mels; ++j)
      for (size_t i = i0; i < i1; ++i) mel[(i - i0) * n_mels + j] = ctx->mel[(size_t)j * n_len + i];
  }
  // E1: conv1 + bias + GELU (1834-1855)
  std::vector<float> h1((size_t)Tm * d);
  conv1d_k3(ctx, mel.data(), Tm, n_mels, m.get("encoder.conv1.weight"), 1, h1.data(), n_threads);
  add_bias_rows(h1.data(), Tm, d, m.get("encoder.conv1.bias").f32());
  gelu_inplace(ctx, h1.data(), h1.size());
  ctx->chk[ORC_STAGE_CONV1 * 1000] = abs_sum(h1.data(), h1.size());
  // E2: conv2 (stride 2) + bias + GELU (1856-1860)
  std::vector<float> cur((size_t)n_ctx * d);
  conv1d_k3(ctx, h1.data(), Tm, d, m.get("encoder.conv2.weight"), 2, cur.data(), n_threads);
  add_bias_rows(cur.data(), n_ctx, d, m.get("encoder.conv2.bias").f32());
  gelu_inplace(ctx, cur.data(), cur.size());
  // E3: inpL = e_pe[:n_ctx] + cur^T (1862-1875); e_pe ne = [d, n_audio_ctx] -> rows of d
  std::vector<float> inpL((size_t)n_ctx * d);
  {
    const float* pe = m.get("encoder.positional_embedding").f32();
    for (size_t i = 0; i < inpL.size(); ++i) inpL[i] = pe[i] + cur[i];
  }
  ctx->chk[ORC_STAGE_CONV2_POS * 1000] = abs_sum(inpL.data(), inpL.size());

  std::vector<float> q((size_t)n_ctx * d), k((size_t)n_ctx * d), v((size_t)n_ctx * d);
  std::vector<float> att((size_t)n_ctx * d), inpFF((size_t)n_ctx * d), hid((size_t)n_ctx * 4 * d);
  const float att_scale = 1.0f / sqrtf((float)(d / H));   // flash_attn scales by 1/sqrt(D)
  for (int il = 0; il < L; ++il) {                          // 1877-1975
    const std::string p = "encoder.blocks." + std::to_string(il) + ".";
    // E4: norm, *w, +b (1882-1886)
    layer_norm(inpL.data(), n_ctx, d, m.get(p + "attn_ln.weight").f32(), m.get(p + "attn_ln.bias").f32(), cur.data());
    // E5: Q (+b), K (no bias), V (+b) (1891-1897)
    linear(ctx, cur.data(), n_ctx, d, m.get(p + "attn.query.weight"), &m.get(p + "attn.query.bias"), q.data(), n_threads);
    linear(ctx, cur.data(), n_ctx, d, m.get(p + "attn.key.weight"), nullptr, k.data(), n_threads);
    linear(ctx, cur.data(), n_ctx, d, m.get(p + "attn.value.weight"), &m.get(p + "attn.value.bias"), v.data(), n_threads);
    // E6: cpy into F16 tensors (1898-1920): round-to-nearest-even
    if (ctx->opt.act_f16_round) {
      round_f16_inplace(q.data(), q.size());
      round_f16_inplace(k.data(), k.size());
      round_f16_inplace(v.data(), v.size());
    }
    // E7: flash_attn + merge (1922-1929)
    attention(ctx, q.data(), n_ctx, k.data(), v.data(), n_ctx, d, H, att_scale, -1, att.data(), n_threads);
    // E8: out projection + bias (1936-1938), residual (1942)
    linear(ctx, att.data(), n_ctx, d, m.get(p + "attn.out.weight"), &m.get(p + "attn.out.bias"), cur.data(), n_threads);
    for (size_t i = 0; i < inpFF.size(); ++i) inpFF[i] = cur[i] + inpL[i];
    // E9: mlp_ln, fc1 + b, GELU, fc2 + b (1948-1965)
    layer_norm(inpFF.data(), n_ctx, d, m.get(p + "mlp_ln.weight").f32(), m.get(p + "mlp_ln.bias").f32(), cur.data());
    linear(ctx, cur.data(), n_ctx, d, m.get(p + "mlp.0.weight"), &m.get(p + "mlp.0.bias"), hid.data(), n_threads);
    gelu_inplace(ctx, hid.data(), hid.size());
    linear(ctx, hid.data(), n_ctx, 4 * d, m.get(p + "mlp.2.weight"), &m.get(p + "mlp.2.bias"), cur.data(), n_threads);
    // E10: inpO = cur + inpFF, carried into inpL (1968-1972)
    for (size_t i = 0; i < inpL.size(); ++i) inpL[i] = cur[i] + inpFF[i];
    ctx->chk[ORC_STAGE_LAYER * 1000 + il] = abs_sum(inpL.data(), inpL.size());
  }
  // E11: ln_post (1980-1984)
  ctx->enc_out.resize((size_t)n_ctx * d);
  layer_norm(inpL.data(), n_ctx, d, m.get("encoder.ln_post.weight").f32(), m.get("encoder.ln_post.bias").f32(), ctx->enc_out.data());
  ctx->chk[ORC_STAGE_LN_POST * 1000] = abs_sum(ctx->enc_out.data(), ctx->enc_out.size());

  // E12: cross-attention memory (1990-2030)
  const int Lt = hp.n_text_layer;
  const int dt = hp.n_text_state;
  ctx->cross_kf.clear();
  ctx->cross_vf.clear();
  ctx->cross_k.assign((size_t)Lt * n_ctx * dt, 0);
  ctx->cross_v.assign((size_t)Lt * n_ctx * dt, 0);
  const float kscale = powf((float)d / (float)H, -0.25f);   // 1994
  std::vector<float> kc((size_t)n_ctx * dt), vc((size_t)n_ctx * dt);
  for (int il = 0; il < Lt; ++il) {
    const std::string p = "decoder.blocks." + std::to_string(il) + ".";
    linear(ctx, ctx->enc_out.data(), n_ctx, d, m.get(p + "cross_attn.key.weight"), nullptr, kc.data(), n_threads);   // 1992
    for (auto& x : kc) x *= kscale;                                                                                  // 1996
    linear(ctx, ctx->enc_out.data(), n_ctx, d, m.get(p + "cross_attn.value.weight"), &m.get(p + "cross_attn.value.bias"), vc.data(), n_threads);   // 2013-2016
    uint16_t* kd = ctx->cross_k.data() + (size_t)il * n_ctx * dt;   // offset 2*d*il*n_ctx bytes (2018-2027)
    uint16_t* vd = ctx->cross_v.data() + (size_t)il * n_ctx * dt;
    double sk = 0.0, sv = 0.0;
    for (size_t i = 0; i < kc.size(); ++i) {                       // cpy F32 -> F16 (2029-2030)
      kd[i] = f32_to_f16_bits(kc[i]);
      vd[i] = f32_to_f16_bits(vc[i]);
      sk += std::fabs((double)f16_bits_to_f32(kd[i]));
      sv += std::fabs((double)f16_bits_to_f32(vd[i]));
    }
    ctx->chk[ORC_STAGE_CROSS_K * 1000 + il] = sk;
    ctx->chk[ORC_STAGE_CROSS_V * 1000 + il] = sv;
  }
  return ORC_OK;
}

}  // namespace wo
