// wb_decode.cu -- the decode step + device-side greedy loop behind wb_decode / wb_decode_greedy.
//
// The reference stops after whisper_encode (src/main.rs:2065-2075); it declares the decoder's
// weights (WhisperLayerDecoder 694-731, tensor table 1139-1333), the self-attention KV cache
// memory_k/v (F16 [n_text_layer * n_text_ctx * n_text_state], 1343-1347) and logits/probs
// (351-352) but no function that uses them.  This file implements SURVEY.md section 8a rows
// D1-D6 (upstream whisper.cpp v1.0.3 semantics) on that state.
//
// Every linear layer runs on the tcgen05 GEMM in swap-AB form: the weight [N_out][K] is the
// 128-row "A" operand streamed once from HBM by TMA, the few activation rows (sequences x
// tokens) are the narrow N tile, and the epilogue stores C^T so activations stay token-major.
// n_past and the greedy step counter live in device memory, so one captured CUDA graph replays
// for every position of the greedy loop (the step is launch-bound otherwise).
#include <math.h>
#include <string.h>

#include "wb_internal.hpp"

namespace wb {

static int bn_for_rows(int R) { return R <= 32 ? 32 : R <= 64 ? 64 : R <= 128 ? 128 : 256; }
static int bn_index(int bn) { return bn == 32 ? 0 : bn == 64 ? 1 : bn == 128 ? 2 : 3; }

static int make_act_maps(wb_ctx* ctx, ActMaps& am, const __half* base, int K, int rows_cap) {
  const int bns[4] = {32, 64, 128, 256};
  const char* err = "";
  for (int i = 0; i < 4; ++i) {
    if (!tmap_2d_rows(&am.m[i], base, (uint64_t)K, (uint64_t)rows_cap, (uint64_t)K, (uint32_t)bns[i], &err))
      return fail_msg(ctx, WB_ERR_TENSOR_OP, std::string("galois tensor:'") + err + "'");
  }
  return WB_OK;
}

// C^T = W * X^T : out[r][n_out] for r < R
// few rows (one token per sequence, <= 32 sequences): the HBM-bound skinny kernel, every 16 weight
// rows on their own CTA (decode_kernels.cu); more rows: the tcgen05 GEMM in swap-AB form
static int run_linear_rows(wb_ctx* ctx, const Linear& l, const __half* x, const ActMaps& act, int R, GemmEpilogue epi,
                           const char* family, float* top2 = nullptr, cudaStream_t st = nullptr);

static int run_gemm_swapped(wb_ctx* ctx, const Linear& l, const ActMaps& act, int R, GemmEpilogue epi,
                            const char* family) {
  GemmProblem g;
  g.a_map = l.map_a;
  const int bn = bn_for_rows(R);
  g.w_map = act.m[bn_index(bn)];
  g.M_rows = l.N;
  g.batch = 1;
  g.N = R;
  g.K = l.K;
  g.bn = bn;
  epi.transpose_out = 1;
  if (!epi.bias) epi.bias = l.bias;
  if (!epi.colscale) epi.colscale = l.colscale;
  g.epi = epi;
  LaunchTimer t(ctx, family);
  WB_CK(launch_gemm(g, ctx->num_sms, ctx->stream));
  return WB_OK;
}

static int run_linear_rows(wb_ctx* ctx, const Linear& l, const __half* x, const ActMaps& act, int R, GemmEpilogue epi,
                           const char* family, float* top2, cudaStream_t st) {
  if (!st) st = ctx->stream;
  if (R > 32 || l.K % 64 != 0) {
    if (st != ctx->stream) return fail_msg(ctx, WB_ERR_TENSOR_OP, "galois tensor:'sequence groups need the skinny linear kernel'");
    if (epi.ln_stats_in || epi.ln_stats_out) return fail_msg(ctx, WB_ERR_TENSOR_OP, "galois tensor:'folded LayerNorm needs the skinny linear kernel'");
    return run_gemm_swapped(ctx, l, act, R, epi, family);
  }
  DecodeLinear a;
  a.w = l.w;
  a.x = x;
  a.N = l.N;
  a.K = l.K;
  a.R = R;
  a.ldx = l.K;
  a.bias = epi.bias ? epi.bias : l.bias;
  a.colscale = epi.colscale ? epi.colscale : l.colscale;
  a.scale = epi.scale;
  a.gelu = epi.gelu;
  a.residual = epi.residual;
  a.res_ld = epi.res_ld;
  a.out = epi.out;
  a.out_f16 = epi.out_f16;
  a.out_ld = epi.out_ld;
  a.top2 = top2;
  a.ln_center = epi.ln_center;
  if (epi.ln_stats_in) {   // folded LayerNorm, consumer side: l carries gamma (upload_cat_ln)
    a.ln_in = epi.ln_stats_in;
    a.ln_c1 = l.ln_c1;
    a.ln_inv_d = 1.0f / (float)l.K;
    a.ln_eps = epi.ln_eps;
  }
  if (epi.ln_stats_out) {  // producer side
    a.ln_out = epi.ln_stats_out;
    a.x16_out = epi.x16_out;
    a.x16_ld = epi.x16_ld;
  }
  LaunchTimer t(ctx, family);
  WB_CK(launch_decode_linear(a, st));
  return WB_OK;
}

int decode_setup(wb_ctx* ctx, const ModelFileView& mv) {
  const ModelHParams& hp = ctx->hp;
  const int d = hp.n_text_state, Lt = hp.n_text_layer, S = ctx->cfg.max_segments;
  const float s = powf((float)d / (float)hp.n_text_head, -0.25f);   // Dh^-1/4 on Q and K (D2, D3)
  int rc;
  // token embedding doubles as the logits projection (D5): d_te [n_vocab][d]
  if ((rc = upload_linear(ctx, mv, "decoder.token_embedding.weight", "", ctx->logits_lin, true))) return rc;
  ctx->d_te = ctx->logits_lin.w;
  if ((rc = upload_f32(ctx, mv, "decoder.positional_embedding", &ctx->d_pe))) return rc;
  if ((rc = upload_f32(ctx, mv, "decoder.ln.weight", &ctx->d_ln_w))) return rc;
  if ((rc = upload_f32(ctx, mv, "decoder.ln.bias", &ctx->d_ln_b))) return rc;
  ctx->dec.resize(Lt);
  for (int i = 0; i < Lt; ++i) {
    const std::string p = "decoder.blocks." + std::to_string(i) + ".";
    DecLayer& l = ctx->dec[i];
    if ((rc = upload_f32(ctx, mv, p + "attn_ln.weight", &l.attn_ln_w))) return rc;
    if ((rc = upload_f32(ctx, mv, p + "attn_ln.bias", &l.attn_ln_b))) return rc;
    if ((rc = upload_f32(ctx, mv, p + "cross_attn_ln.weight", &l.cross_ln_w))) return rc;
    if ((rc = upload_f32(ctx, mv, p + "cross_attn_ln.bias", &l.cross_ln_b))) return rc;
    if ((rc = upload_f32(ctx, mv, p + "mlp_ln.weight", &l.mlp_ln_w))) return rc;
    if ((rc = upload_f32(ctx, mv, p + "mlp_ln.bias", &l.mlp_ln_b))) return rc;
    // Q = (Wq x + bq) * s ; K = (Wk x) * s ; V = Wv x + bv   (D2)
    // the three linears that read a LayerNorm carry its gamma / beta (and the Dh^-1/4 scale) in their weights:
    // the single-token step applies the row statistics in the linear's epilogue (no LayerNorm kernel), the
    // many-row prompt pass normalises with a plain LayerNorm (gamma = 1, beta = 0) in front of the same weights
    if ((rc = upload_cat_ln(ctx, mv,
                            {{p + "attn.query.weight", p + "attn.query.bias", s},
                             {p + "attn.key.weight", "", s},
                             {p + "attn.value.weight", p + "attn.value.bias", 1.0f}},
                            p + "attn_ln.weight", p + "attn_ln.bias", l.qkv, true)))
      return rc;
    if ((rc = upload_linear(ctx, mv, p + "attn.out.weight", p + "attn.out.bias", l.out, true))) return rc;
    if ((rc = upload_cat_ln(ctx, mv, {{p + "cross_attn.query.weight", p + "cross_attn.query.bias", s}},
                            p + "cross_attn_ln.weight", p + "cross_attn_ln.bias", l.cq, true)))
      return rc;
    if ((rc = upload_linear(ctx, mv, p + "cross_attn.out.weight", p + "cross_attn.out.bias", l.cout, true))) return rc;
    if ((rc = upload_cat_ln(ctx, mv, {{p + "mlp.0.weight", p + "mlp.0.bias", 1.0f}}, p + "mlp_ln.weight", p + "mlp_ln.bias",
                            l.fc1, true)))
      return rc;
    if ((rc = upload_linear(ctx, mv, p + "mlp.2.weight", p + "mlp.2.bias", l.fc2, true))) return rc;
  }
  // ---- state
  const size_t n_ctx = hp.n_text_ctx;
  size_t R = (size_t)S * n_ctx;
  if (R < 256) R = 256;
  ctx->dec_rows_cap = (int)R;
  if ((rc = dev_alloc(ctx, &ctx->self_k, (size_t)Lt * S * n_ctx * d))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->self_v, (size_t)Lt * S * n_ctx * d))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->dx, R * d))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_ln, R * d))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_qkv, R * 3 * d))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_att, R * d))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_q, R * d))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_hid, R * 4 * d))) return rc;
  const size_t Sf = S < 256 ? 256 : S;
  if ((rc = dev_alloc(ctx, &ctx->d_lnf, Sf * d))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_logits, (size_t)S * hp.n_vocab))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_top2, (size_t)32 * decode_linear_parts(hp.n_vocab) * 3))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_tokens, R))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_next, (size_t)S))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_out_tokens, (size_t)S * n_ctx))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_out_margin, (size_t)S * n_ctx))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_margin, (size_t)S))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_out_len, (size_t)S))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_done, (size_t)S))) return rc;
  {
    std::vector<float> ones((size_t)d, 1.0f), zeros((size_t)d, 0.0f);
    if ((rc = upload_f32_vec(ctx, ones, &ctx->d_ones))) return rc;
    if ((rc = upload_f32_vec(ctx, zeros, &ctx->d_zeros))) return rc;
  }
  if ((rc = dev_alloc(ctx, &ctx->dec_ln_stats, (size_t)DEC_LN_SLOT * (3 * Lt + 1)))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->dec_ln_center, (size_t)DEC_LN_ROWS))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_npast, 4))) return rc;
  ctx->d_step = ctx->d_npast + 1;
  // split-K cross-attention partials (only used for few rows: n_tok <= 8)
  const size_t prow = (size_t)S * 8 * hp.n_text_head * 8;
  if ((rc = dev_alloc(ctx, &ctx->d_part_o, prow * 64))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_part_ml, prow * 2))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_split_cnt, (size_t)S * 8 * hp.n_text_head))) return rc;   // zeroed; self-resetting
  if ((rc = make_act_maps(ctx, ctx->m_ln, ctx->d_ln, d, (int)R))) return rc;
  if ((rc = make_act_maps(ctx, ctx->m_att, ctx->d_att, d, (int)R))) return rc;
  if ((rc = make_act_maps(ctx, ctx->m_hid, ctx->d_hid, 4 * d, (int)R))) return rc;
  if ((rc = make_act_maps(ctx, ctx->m_lnf, ctx->d_lnf, d, (int)Sf))) return rc;
  return WB_OK;
}

// One decode pass over `n_tok` new tokens per sequence: embed -> L x (self-attn, cross-attn, MLP)
// -> final LN of the last position -> logits.  Launch-only (no host sync, no copies): capturable.
// WB_DEC_SKIP (timing experiments only, results become meaningless): comma list of kernel families to leave out of
// the step -- "cross", "self", "linear", "logits"
static bool dec_skip(const char* what) {
  const char* e = getenv("WB_DEC_SKIP");
  return e && strstr(e, what) != nullptr;
}

// `seq0` / `st`: the pass covers sequences [seq0, seq0 + n_seq) and is issued on stream `st` -- the single-token step
// of a large batch is split into sequence groups that run on parallel branches of the step graph (wb_decode_greedy):
// a group's chain of ~100 small latency-bound kernels overlaps the other groups' HBM-bound cross-attention.  Every
// activation buffer is row-indexed, so a group simply works on its rows (seq0 > 0 needs n_tok == 1).
static int decode_pass(wb_ctx* ctx, const int* tokens_dev, int n_seq, int n_tok, int seq0 = 0, cudaStream_t st = nullptr) {
  if (!st) st = ctx->stream;
  if (seq0 > 0 && n_tok != 1) return fail_msg(ctx, WB_ERR_UNEXPECTED, "Unexpected: sequence groups are single-token passes");
  const ModelHParams& hp = ctx->hp;
  const bool skip_cross = dec_skip("cross"), skip_self = dec_skip("self"), skip_lin = dec_skip("linear"),
             skip_logits = dec_skip("logits");
  const int d = hp.n_text_state, H = hp.n_text_head, Lt = hp.n_text_layer;
  const int T = ctx->enc_T > 0 ? ctx->enc_T : hp.n_audio_ctx;   // rows of the cross K/V the last wb_encode wrote
  const int n_ctx = hp.n_text_ctx;
  const int R = n_seq * n_tok;
  const long long ld_kv = d;   // cross K / V of one layer: dense [seg][T][d]
  int rc;
  // this group's rows of every buffer
  const size_t r0 = (size_t)seq0 * n_tok;
  tokens_dev += r0;
  float* const dx = ctx->dx + r0 * d;
  __half* const d_ln = ctx->d_ln + r0 * d;
  __half* const d_qkv = ctx->d_qkv + r0 * 3 * d;
  __half* const d_att = ctx->d_att + r0 * d;
  __half* const d_q = ctx->d_q + r0 * d;
  __half* const d_hid = ctx->d_hid + r0 * 4 * d;
  __half* const d_lnf = ctx->d_lnf + (size_t)seq0 * d;
  float* const d_logits = ctx->d_logits + (size_t)seq0 * hp.n_vocab;
  float* const d_top2 = ctx->d_top2 + (size_t)seq0 * decode_linear_parts(hp.n_vocab) * 3;
  // few rows (the single-token step): LayerNorm folded into the linears around it -- the producers of the
  // residual stream leave row statistics + an F16 copy (d_ln), the consumers apply them (slot 3 il + {0, 1, 2} =
  // attn_ln, cross_attn_ln, mlp_ln of layer il).  Many rows (prompt pass on the tcgen05 GEMM): a LayerNorm
  // kernel without affine part in front of the same folded weights.
  const bool fold = R <= DEC_LN_ROWS && d % 64 == 0;
  auto slot = [&](int i) { return ctx->dec_ln_stats + (size_t)i * DEC_LN_SLOT + r0; };
  {
    LaunchTimer t(ctx, "dec_embed");
    WB_CK(launch_embed(ctx->d_te, ctx->d_pe, tokens_dev, n_seq, n_tok, ctx->d_npast, d, dx, st,
                       fold ? slot(0) : nullptr, d_ln, fold ? 3 * Lt : 0, fold ? ctx->dec_ln_center + r0 : nullptr));
  }
  const int n_split = n_tok <= 8 ? decode_cross_splits(n_seq, H, T, ctx->num_sms) : 1;
  // L2 prefetch branch (captured single-token step only): the cross K / V of the first n_pf sequences of layer il
  const bool prefetch = ctx->dec_pf_active && n_tok == 1 && ctx->dec_pf_mb > 0 && !skip_cross;
  const size_t seq_kv_bytes = (size_t)T * d * sizeof(__half);
  int n_pf = prefetch ? (int)(((size_t)ctx->dec_pf_mb << 20) / (2 * seq_kv_bytes)) : 0;
  if (n_pf > n_seq) n_pf = n_seq;
  auto prefetch_layer = [&](int il) -> cudaError_t {
    cudaStream_t ps = ctx->dec_pf_stream;
    cudaError_t e = cudaEventRecord(ctx->dec_pf_fork, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ps, ctx->dec_pf_fork, 0);
    const __half* kx = ctx->cross + (size_t)(2 * il) * ctx->cross_slab + (size_t)seq0 * T * d;
    if (e == cudaSuccess) e = launch_l2_prefetch(kx, kx + ctx->cross_slab, (size_t)n_pf * seq_kv_bytes, ps);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->dec_pf_join, ps);
    return e;
  };
  if (n_pf > 0) WB_CK(prefetch_layer(0));
  for (int il = 0; il < Lt; ++il) {
    const DecLayer& l = ctx->dec[il];
    if (!fold) {   // D2: self-attention
      LaunchTimer t(ctx, "dec_layernorm");
      WB_CK(launch_layernorm(dx, ctx->d_ones, ctx->d_zeros, R, d, d_ln, nullptr, st, 0, true));
    }
    {
      GemmEpilogue e;
      e.out = d_qkv;
      e.out_f16 = 1;
      e.out_ld = 3 * d;
      if (fold) {
        e.ln_stats_in = slot(3 * il);
        e.ln_center = ctx->dec_ln_center + r0;
      }
      if (!skip_lin && (rc = run_linear_rows(ctx, l.qkv, d_ln, ctx->m_ln, R, e, "dec_gemm", nullptr, st))) return rc;
    }
    if (!skip_self) {
      LaunchTimer t(ctx, "dec_self_attn");
      __half* kc = ctx->self_k + ((size_t)il * ctx->cfg.max_segments + seq0) * n_ctx * d;
      __half* vc = ctx->self_v + ((size_t)il * ctx->cfg.max_segments + seq0) * n_ctx * d;
      WB_CK(launch_decode_self_attn(d_qkv, d, kc, vc, n_seq, n_tok, ctx->d_npast, n_ctx, H, d_att, st));
    }
    {
      GemmEpilogue e;
      e.residual = dx;
      e.res_ld = d;
      e.out = dx;
      e.out_f16 = 0;
      e.out_ld = d;
      if (fold) {
        e.ln_stats_out = slot(3 * il + 1);
        e.ln_center = ctx->dec_ln_center + r0;
        e.x16_out = d_ln;
        e.x16_ld = d;
      }
      if (!skip_lin && (rc = run_linear_rows(ctx, l.out, d_att, ctx->m_att, R, e, "dec_gemm", nullptr, st))) return rc;
    }
    if (!fold) {   // D3: cross-attention over memory_cross_k/v written by wb_encode
      LaunchTimer t(ctx, "dec_layernorm");
      WB_CK(launch_layernorm(dx, ctx->d_ones, ctx->d_zeros, R, d, d_ln, nullptr, st, 0, true));
    }
    {
      GemmEpilogue e;
      e.out = d_q;
      e.out_f16 = 1;
      e.out_ld = d;
      if (fold) {
        e.ln_stats_in = slot(3 * il + 1);
        e.ln_center = ctx->dec_ln_center + r0;
      }
      if (!skip_lin && (rc = run_linear_rows(ctx, l.cq, d_ln, ctx->m_ln, R, e, "dec_gemm", nullptr, st))) return rc;
    }
    if (n_pf > 0) WB_CK(cudaStreamWaitEvent(st, ctx->dec_pf_join, 0));   // join: the branch always ends before its consumer
    if (!skip_cross) {
      LaunchTimer t(ctx, "dec_cross_attn");
      const __half* kx = ctx->cross + (size_t)(2 * il) * ctx->cross_slab + (size_t)seq0 * T * d;   // segment seq0 onwards
      // (timing experiment WB_DEC_SKIP=headmajor: read the slab as if it were head-major -- garbage results)
      const bool hm = dec_skip("headmajor");
      WB_CK(launch_decode_cross_attn(d_q, d, kx, kx + ctx->cross_slab, hm ? 64 : ld_kv,
                                     hm ? (long long)ctx->cfg.max_segments * T * 64 : 64, n_seq, n_tok, T, H, d_att,
                                     ctx->d_part_o + (size_t)seq0 * H * 8 * 64, ctx->d_part_ml + (size_t)seq0 * H * 8 * 2, n_split,
                                     ctx->d_split_cnt + (size_t)seq0 * H, st));
    }
    if (n_pf > 0 && il + 1 < Lt) WB_CK(prefetch_layer(il + 1));   // behind this layer's cross-attention, under the linears that follow
    {
      GemmEpilogue e;
      e.residual = dx;
      e.res_ld = d;
      e.out = dx;
      e.out_f16 = 0;
      e.out_ld = d;
      if (fold) {
        e.ln_stats_out = slot(3 * il + 2);
        e.ln_center = ctx->dec_ln_center + r0;
        e.x16_out = d_ln;
        e.x16_ld = d;
      }
      if (!skip_lin && (rc = run_linear_rows(ctx, l.cout, d_att, ctx->m_att, R, e, "dec_gemm", nullptr, st))) return rc;
    }
    if (!fold) {   // D4: MLP
      LaunchTimer t(ctx, "dec_layernorm");
      WB_CK(launch_layernorm(dx, ctx->d_ones, ctx->d_zeros, R, d, d_ln, nullptr, st, 0, true));
    }
    {
      GemmEpilogue e;
      e.gelu = 1;
      e.out = d_hid;
      e.out_f16 = 1;
      e.out_ld = 4 * d;
      if (fold) {
        e.ln_stats_in = slot(3 * il + 2);
        e.ln_center = ctx->dec_ln_center + r0;
      }
      if (!skip_lin && (rc = run_linear_rows(ctx, l.fc1, d_ln, ctx->m_ln, R, e, "dec_gemm", nullptr, st))) return rc;
    }
    {
      GemmEpilogue e;
      e.residual = dx;
      e.res_ld = d;
      e.out = dx;
      e.out_f16 = 0;
      e.out_ld = d;
      if (fold && il + 1 < Lt) {   // (decoder.ln in front of the logits keeps its kernel: d_te is shared with the embedding)
        e.ln_stats_out = slot(3 * il + 3);
        e.ln_center = ctx->dec_ln_center + r0;
        e.x16_out = d_ln;
        e.x16_ld = d;
      }
      if (!skip_lin && (rc = run_linear_rows(ctx, l.fc2, d_hid, ctx->m_hid, R, e, "dec_gemm", nullptr, st))) return rc;
    }
  }
  {   // D5: logits of the last position of every sequence
    LaunchTimer t(ctx, "dec_layernorm");
    WB_CK(launch_layernorm(dx + (size_t)(n_tok - 1) * d, ctx->d_ln_w, ctx->d_ln_b, n_seq, d, d_lnf, nullptr,
                           st, (long long)n_tok * d, true));
  }
  {
    GemmEpilogue e;
    e.out = d_logits;
    e.out_f16 = 0;
    e.out_ld = hp.n_vocab;
    // <= 32 sequences: the skinny kernel also leaves per-CTA top-2 partials, so D6 never re-reads the logits
    ctx->logits_top2_valid = n_seq <= 32 && ctx->logits_lin.K % 64 == 0;
    if (!skip_logits && (rc = run_linear_rows(ctx, ctx->logits_lin, d_lnf, ctx->m_lnf, n_seq, e, "dec_gemm_logits",
                                              ctx->logits_top2_valid ? d_top2 : nullptr, st)))
      return rc;
  }
  return WB_OK;
}

// D6 bookkeeping after a decode pass: from the top-2 partials when the skinny logits kernel ran
// advance_by > 0: the kernel also moves n_past / step on once every sequence is done (no separate advance launch)
static cudaError_t run_argmax(wb_ctx* ctx, int n_seqs, int max_new, int eot, int seq0 = 0, cudaStream_t st = nullptr,
                              int advance_by = 0) {
  const ModelHParams& hp = ctx->hp;
  if (!st) st = ctx->stream;
  const size_t o = (size_t)seq0, om = (size_t)seq0 * max_new;   // this group's sequences
  if (ctx->logits_top2_valid)
    return launch_argmax_partials(ctx->d_top2 + o * decode_linear_parts(hp.n_vocab) * 3, decode_linear_parts(hp.n_vocab), n_seqs,
                                  ctx->d_next + o, ctx->d_margin + o, ctx->d_out_tokens + om, ctx->d_out_margin + om,
                                  ctx->d_out_len + o, ctx->d_done + o, max_new, ctx->d_step, eot, st,
                                  advance_by > 0 ? ctx->d_npast : nullptr, advance_by);
  return launch_argmax(ctx->d_logits + o * hp.n_vocab, n_seqs, hp.n_vocab, ctx->d_next + o, ctx->d_margin + o,
                       ctx->d_out_tokens + om, ctx->d_out_margin + om, ctx->d_out_len + o, ctx->d_done + o, max_new, ctx->d_step,
                       eot, st, advance_by > 0 ? ctx->d_npast : nullptr, advance_by);
}

static int check_decode_args(wb_ctx* ctx, int n_tok, int n_past, int n_seq) {
  if (!ctx->cfg.decode_capacity || !ctx->self_k)
    return fail_msg(ctx, WB_ERR_UNEXPECTED, "Unexpected: context created without decode_capacity");
  if (n_seq < 1 || n_seq > ctx->cfg.max_segments || n_tok < 1 || n_past < 0 ||
      n_past + n_tok > ctx->hp.n_text_ctx || (long long)n_seq * n_tok > ctx->dec_rows_cap)
    return fail_msg(ctx, WB_ERR_NOT_ENOUGH_SPACE, "not enough space in the context's memory pool\n");
  if (n_seq > ctx->enc_n_seg)
    return fail_msg(ctx, WB_ERR_UNEXPECTED, "Unexpected: decode needs the cross K/V of wb_encode for every sequence");
  return WB_OK;
}

}  // namespace wb

using namespace wb;

extern "C" {

int wb_decode(wb_ctx* ctx, const int32_t* tokens, int n_tokens, int n_past, int n_seqs) {
  if (!ctx || !tokens) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  int rc = check_decode_args(ctx, n_tokens, n_past, n_seqs);
  if (rc) return rc;
  const int R = n_seqs * n_tokens;
  for (int i = 0; i < R; ++i)
    if (tokens[i] < 0 || tokens[i] >= ctx->hp.n_vocab) return fail_msg(ctx, WB_ERR_UNEXPECTED, "Unexpected: token id out of range");
  cudaStream_t st = ctx->stream;
  cudaEventRecord(ctx->ev[2][0], st);
  WB_CK(cudaMemcpyAsync(ctx->d_tokens, tokens, sizeof(int) * R, cudaMemcpyHostToDevice, st));
  WB_CK(launch_fill_i32(ctx->d_npast, 1, n_past, st));
  rc = decode_pass(ctx, ctx->d_tokens, n_seqs, n_tokens);
  if (rc) return rc;
  cudaEventRecord(ctx->ev[2][1], st);
  ctx->ev_used[2] = true;
  ctx->dec_n_seq = n_seqs;
  ctx->tm.n_decode_calls += 1;
  WB_CK(cudaStreamSynchronize(st));   // `tokens` is caller memory: the H2D copy must have drained
  return WB_OK;
}

int wb_logits_read(wb_ctx* ctx, int seq, float* out) {
  if (!ctx || !out || seq < 0 || seq >= ctx->dec_n_seq) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  const size_t n = (size_t)ctx->hp.n_vocab;
  WB_CK(cudaMemcpyAsync(out, ctx->d_logits + (size_t)seq * n, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  WB_CK(cudaStreamSynchronize(ctx->stream));
  return WB_OK;
}

// D6: greedy = arg-max over all logits; prompt shared by all sequences; a sequence stops recording
// after `eot` or max_new tokens or when the text context is full.
int wb_decode_greedy(wb_ctx* ctx, const int32_t* prompt, int n_prompt, int max_new, int eot, int n_seqs,
                     int32_t* out_tokens, float* out_margin, int32_t* out_len) {
  if (!ctx || !prompt || !out_tokens || !out_len || max_new < 1) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  int rc = check_decode_args(ctx, n_prompt, 0, n_seqs);
  if (rc) return rc;
  const ModelHParams& hp = ctx->hp;
  const int n_ctx = hp.n_text_ctx;
  // the caller's arrays are [n_seqs][max_new]; the device-side ones are [n_seqs][min(max_new, n_text_ctx)] --
  // only the step count is clamped, the caller's row stride is kept (2-D copies at the end)
  const int out_stride = max_new;
  if (max_new > n_ctx) max_new = n_ctx;
  cudaStream_t st = ctx->stream;
  std::vector<int> toks((size_t)n_seqs * n_prompt);
  for (int s = 0; s < n_seqs; ++s)
    for (int i = 0; i < n_prompt; ++i) {
      if (prompt[i] < 0 || prompt[i] >= hp.n_vocab) return fail_msg(ctx, WB_ERR_UNEXPECTED, "Unexpected: token id out of range");
      toks[(size_t)s * n_prompt + i] = prompt[i];
    }
  cudaEventRecord(ctx->ev[2][0], st);
  WB_CK(cudaMemcpyAsync(ctx->d_tokens, toks.data(), sizeof(int) * toks.size(), cudaMemcpyHostToDevice, st));
  WB_CK(cudaMemsetAsync(ctx->d_done, 0, sizeof(int) * n_seqs, st));
  WB_CK(cudaMemsetAsync(ctx->d_out_len, 0, sizeof(int) * n_seqs, st));
  WB_CK(cudaMemsetAsync(ctx->d_npast, 0, sizeof(int) * 2, st));   // n_past = 0, step = 0
  // out buffers are indexed [seq][max_new]
  WB_CK(cudaMemsetAsync(ctx->d_out_tokens, 0, sizeof(int) * (size_t)n_seqs * max_new, st));
  WB_CK(cudaMemsetAsync(ctx->d_out_margin, 0, sizeof(float) * (size_t)n_seqs * max_new, st));
  // ---- prompt pass
  if ((rc = decode_pass(ctx, ctx->d_tokens, n_seqs, n_prompt))) return rc;
  {
    LaunchTimer t(ctx, "dec_argmax");
    WB_CK(run_argmax(ctx, n_seqs, max_new, eot, 0, nullptr, n_prompt));   // + n_past += n_prompt, step = 1
  }
  WB_CK(cudaStreamSynchronize(st));   // `toks` is a stack temporary
  // ---- single-token steps: captured once, replayed for every position
  int n_past = n_prompt;
  const int n_steps_max = max_new - 1;
  const bool use_graph = !ctx->time_kernels;
  // everything decode_pass bakes into its launches is part of the key: the audio context of the last encode (the
  // segment stride of the cross K/V and the key-range split of the cross-attention) as well as the batch shape
  const int enc_T = ctx->enc_T > 0 ? ctx->enc_T : hp.n_audio_ctx;
  // sequence groups: the step of a large batch runs as G parallel branches of the graph, each over its own block of
  // sequences (multiples of 8 = the skinny linear's row groups).  A branch is a chain of ~8 L small kernels that is
  // bound by launch / L2 latency except for its cross-attention, which streams 2 T d bytes per sequence and layer
  // from HBM: with several branches in flight one group's latency-bound kernels run under another group's stream.
  // The weights are read once per group (from L2 for all but the first).  Off by default (see the measurements below);
  // WB_DEC_GROUPS=2|4 selects it.
  int n_groups = 1;
  {
    static const int forced = [] { const char* e = getenv("WB_DEC_GROUPS"); return e ? atoi(e) : 0; }();
    const double w_bytes = (14.0 * hp.n_text_layer * hp.n_text_state * hp.n_text_state + (double)hp.n_vocab * hp.n_text_state) * 2.0;
    const double kv_bytes = (double)n_seqs * hp.n_text_layer * 2.0 * enc_T * hp.n_text_state * 2.0;
    (void)w_bytes;
    (void)kv_bytes;
    // measured on B200 (tools/dec_groups.py, profiles/r02_dec_groups.jsonl): small B = 32: 1 group 0.755 ms/step, 2 groups
    // 0.760, 4 groups 0.903; medium B = 32: 2.01 / 2.00 / 2.44; large-v3 B = 15: 2.10 / 2.48 -- the branches do not
    // overlap usefully (a cross-attention grid fills every SM's registers while it runs), so the default stays 1
    if (forced > 0) n_groups = forced;
    if (n_seqs > 32) n_groups = 1;   // more than 32 rows: the tcgen05 swap-AB path, one stream
    while (n_groups > 1 && (n_seqs + n_groups - 1) / n_groups < 8) n_groups >>= 1;
    if (n_groups > WB_MAX_DEC_GROUPS) n_groups = WB_MAX_DEC_GROUPS;
  }
  const int per_group = n_groups > 1 ? (((n_seqs + n_groups - 1) / n_groups) + 7) / 8 * 8 : n_seqs;
  const int n_split = decode_cross_splits(n_groups > 1 ? per_group : n_seqs, hp.n_text_head, enc_T, ctx->num_sms);
  if (use_graph && (ctx->step_graph == nullptr || ctx->step_graph_n_seq != n_seqs ||
                    ctx->step_graph_max_new != max_new || ctx->step_graph_eot != eot ||
                    ctx->step_graph_enc_T != enc_T || ctx->step_graph_n_split != n_split ||
                    ctx->step_graph_groups != n_groups)) {
    if (ctx->step_graph) {
      cudaGraphExecDestroy(ctx->step_graph);
      ctx->step_graph = nullptr;
    }
    cudaGraph_t graph = nullptr;
    if (n_groups > 1 && !ctx->dec_group_stream[0]) {
      for (int g = 0; g < WB_MAX_DEC_GROUPS; ++g) {
        WB_CK(cudaStreamCreateWithFlags(&ctx->dec_group_stream[g], cudaStreamNonBlocking));
        WB_CK(cudaEventCreateWithFlags(&ctx->dec_group_done[g], cudaEventDisableTiming));
      }
      WB_CK(cudaEventCreateWithFlags(&ctx->dec_fork, cudaEventDisableTiming));
    }
    {
      // B200, 224 greedy steps (tools/dec_groups.py): small B = 32 0.721 ms/step without, 0.703 / 0.701 / 0.700 / 0.713 /
      // 0.729 with 24 / 32 / 48 / 72 / 96 MB per layer; medium B = 32 1.728 without, 1.693 / 1.691 / 1.715 / 1.808 / 1.888;
      // large-v3 B = 15 2.122 without, 2.150 / 2.161 with 24 / 48 (its linears stream 50 MB of weights per layer and
      // want the bandwidth themselves).  Identical token ids throughout.
      static const int forced_mb = [] { const char* e = getenv("WB_DEC_PREFETCH_MB"); return e ? atoi(e) : -1; }();
      const double layer_w_mb = 16.0 * hp.n_text_state * hp.n_text_state * 2.0 / (1 << 20);
      const int pf_mb = forced_mb >= 0 ? forced_mb : (layer_w_mb <= 36.0 ? 32 : 0);
      ctx->dec_pf_mb = pf_mb;
      if (pf_mb > 0 && !ctx->dec_pf_stream) {
        WB_CK(cudaStreamCreateWithFlags(&ctx->dec_pf_stream, cudaStreamNonBlocking));
        WB_CK(cudaEventCreateWithFlags(&ctx->dec_pf_fork, cudaEventDisableTiming));
        WB_CK(cudaEventCreateWithFlags(&ctx->dec_pf_join, cudaEventDisableTiming));
      }
    }
    WB_CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    cudaError_t e1 = cudaSuccess, e2 = cudaSuccess;
    if (n_groups <= 1) {
      ctx->dec_pf_active = true;
      rc = decode_pass(ctx, ctx->d_next, n_seqs, 1);
      ctx->dec_pf_active = false;
      if (rc == WB_OK) e1 = run_argmax(ctx, n_seqs, max_new, eot, 0, nullptr, 1);   // + the step's bookkeeping
    } else {
      // fork: every branch starts behind the previous step's bookkeeping on `st`; join: `st` waits for every branch
      e1 = cudaEventRecord(ctx->dec_fork, st);
      rc = WB_OK;
      for (int g = 0, s0 = 0; s0 < n_seqs && rc == WB_OK && e1 == cudaSuccess; ++g, s0 += per_group) {
        const int ng = n_seqs - s0 < per_group ? n_seqs - s0 : per_group;
        cudaStream_t gs = ctx->dec_group_stream[g];
        e1 = cudaStreamWaitEvent(gs, ctx->dec_fork, 0);
        if (e1 == cudaSuccess) rc = decode_pass(ctx, ctx->d_next, ng, 1, s0, gs);
        if (rc == WB_OK && e1 == cudaSuccess) e1 = run_argmax(ctx, ng, max_new, eot, s0, gs);
        if (e1 == cudaSuccess) e1 = cudaEventRecord(ctx->dec_group_done[g], gs);
        if (e1 == cudaSuccess) e1 = cudaStreamWaitEvent(st, ctx->dec_group_done[g], 0);
      }
    }
    if (rc == WB_OK && e1 == cudaSuccess && n_groups > 1) e2 = launch_advance(ctx->d_npast, 1, ctx->d_step, st);
    cudaError_t e3 = cudaStreamEndCapture(st, &graph);
    if (rc != WB_OK) return rc;
    if (e1 != cudaSuccess) return fail(ctx, WB_ERR_TENSOR_OP, "argmax (capture)", e1);
    if (e2 != cudaSuccess) return fail(ctx, WB_ERR_TENSOR_OP, "advance (capture)", e2);
    if (e3 != cudaSuccess) return fail(ctx, WB_ERR_TENSOR_OP, "cudaStreamEndCapture", e3);
    cudaError_t e4 = cudaGraphInstantiate(&ctx->step_graph, graph, 0);
    cudaGraphDestroy(graph);
    if (e4 != cudaSuccess) return fail(ctx, WB_ERR_TENSOR_OP, "cudaGraphInstantiate", e4);
    ctx->step_graph_n_seq = n_seqs;
    ctx->step_graph_max_new = max_new;
    ctx->step_graph_eot = eot;
    ctx->step_graph_enc_T = enc_T;
    ctx->step_graph_n_split = n_split;
    ctx->step_graph_groups = n_groups;
  }
  std::vector<int> done_h(n_seqs, 0);
  for (int it = 0; it < n_steps_max; ++it) {
    if (n_past + 1 > n_ctx) break;   // text context full
    if (use_graph) {
      WB_CK(cudaGraphLaunch(ctx->step_graph, st));
      // per layer: 6 linear + self-attn + cross-attn (LayerNorm folded into the linears); + embed, final LN,
      // logits, arg-max, advance
      ctx->tm.n_kernel_launches += (8 * hp.n_text_layer + 4) * (n_groups > 1 ? (n_seqs + per_group - 1) / per_group : 1) + (n_groups > 1 ? 1 : 0) +
                                   (n_groups <= 1 && ctx->dec_pf_mb > 0 ? hp.n_text_layer : 0);   // + the L2 prefetch branch
    } else {
      if ((rc = decode_pass(ctx, ctx->d_next, n_seqs, 1))) return rc;
      LaunchTimer t(ctx, "dec_argmax");
      WB_CK(run_argmax(ctx, n_seqs, max_new, eot, 0, nullptr, 1));
    }
    n_past += 1;
    if ((it & 31) == 31) {   // every 32 steps: stop early once every sequence has emitted eot
      WB_CK(cudaMemcpyAsync(done_h.data(), ctx->d_done, sizeof(int) * n_seqs, cudaMemcpyDeviceToHost, st));
      WB_CK(cudaStreamSynchronize(st));
      bool all = true;
      for (int v : done_h) all = all && v != 0;
      if (all) break;
    }
  }
  cudaEventRecord(ctx->ev[2][1], st);
  ctx->ev_used[2] = true;
  ctx->dec_n_seq = n_seqs;
  ctx->tm.n_decode_calls += 1;
  WB_CK(cudaMemcpy2DAsync(out_tokens, sizeof(int) * (size_t)out_stride, ctx->d_out_tokens, sizeof(int) * (size_t)max_new,
                          sizeof(int) * (size_t)max_new, (size_t)n_seqs, cudaMemcpyDeviceToHost, st));
  if (out_margin)
    WB_CK(cudaMemcpy2DAsync(out_margin, sizeof(float) * (size_t)out_stride, ctx->d_out_margin, sizeof(float) * (size_t)max_new,
                            sizeof(float) * (size_t)max_new, (size_t)n_seqs, cudaMemcpyDeviceToHost, st));
  WB_CK(cudaMemcpyAsync(out_len, ctx->d_out_len, sizeof(int) * n_seqs, cudaMemcpyDeviceToHost, st));
  WB_CK(cudaStreamSynchronize(st));
  return WB_OK;
}

}  // extern "C"
