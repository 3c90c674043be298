// wb_loader.hpp -- ggml-v1 ("lmgg") model-file parser of the product library (host side).
//
// Parses the layout WhisperContext::new reads (src/main.rs:366-503): magic (46, 368-371),
// 11 x i32 hparams (622-633), mel filterbank (513-524), vocab (430-431, 578-589) and tensor
// records `n_dims, name_len, ftype, ne[], name, data` (1385-1437), applying the reference's
// checks (unknown name / element count / per-dim shape / byte size, 1401-1434) against the
// tensor table of 960-1334.  The file is read once into memory; tensors are views into it.
#pragma once

#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/whisper_b200.h"

namespace wb {

struct HostTensor {
  int n_dims = 0;
  int64_t ne[3] = {1, 1, 1};   // ne[0] innermost
  bool f16 = false;
  const uint8_t* data = nullptr;
  size_t bytes = 0;
  int64_t nelem() const { return ne[0] * ne[1] * ne[2]; }
};

struct ModelHParams {   // src/main.rs:607-619
  int32_t n_vocab, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer;
  int32_t n_text_ctx, n_text_state, n_text_head, n_text_layer, n_mels, f16;
};

struct ModelFileView {
  std::vector<uint8_t> blob;
  ModelHParams hp{};
  int filt_n_mel = 0, filt_n_fft = 0;
  const float* filters = nullptr;
  int32_t n_vocab_file = 0;
  std::vector<std::string> vocab;   // id_to_token (544): file entries + the reference's names for extra ids (442-467)
  int32_t special[8] = {50256, 50257, 50360, 50361, 50362, 50363, 50358, 50359};   // 557-575
  std::unordered_map<std::string, HostTensor> tensors;
};

namespace detail {
struct Expect {
  int n_dims;
  int64_t ne[3];
  bool weight;   // dtype follows hparams.f16 (817-821)
};

inline void expect_table(const ModelHParams& hp, std::unordered_map<std::string, Expect>& t) {
  const int64_t da = hp.n_audio_state, dt = hp.n_text_state;
  auto add = [&](const std::string& n, int nd, int64_t a, int64_t b, int64_t c, bool w) { t[n] = Expect{nd, {a, b, c}, w}; };
  add("encoder.positional_embedding", 2, da, hp.n_audio_ctx, 1, false);
  add("encoder.conv1.weight", 3, 3, hp.n_mels, da, true);
  add("encoder.conv1.bias", 2, 1, da, 1, false);
  add("encoder.conv2.weight", 3, 3, da, da, true);
  add("encoder.conv2.bias", 2, 1, da, 1, false);
  add("encoder.ln_post.weight", 1, da, 1, 1, false);
  add("encoder.ln_post.bias", 1, da, 1, 1, false);
  add("decoder.positional_embedding", 2, dt, hp.n_text_ctx, 1, false);
  add("decoder.token_embedding.weight", 2, dt, hp.n_vocab, 1, true);
  add("decoder.ln.weight", 1, dt, 1, 1, false);
  add("decoder.ln.bias", 1, dt, 1, 1, false);
  auto block = [&](const std::string& p, int64_t d, bool cross) {
    add(p + "mlp_ln.weight", 1, d, 1, 1, false);
    add(p + "mlp_ln.bias", 1, d, 1, 1, false);
    add(p + "mlp.0.weight", 2, d, 4 * d, 1, true);
    add(p + "mlp.0.bias", 1, 4 * d, 1, 1, false);
    add(p + "mlp.2.weight", 2, 4 * d, d, 1, true);
    add(p + "mlp.2.bias", 1, d, 1, 1, false);
    add(p + "attn_ln.weight", 1, d, 1, 1, false);
    add(p + "attn_ln.bias", 1, d, 1, 1, false);
    const char* pre[2] = {"attn", "cross_attn"};
    for (int c = 0; c < (cross ? 2 : 1); ++c) {
      const std::string q = p + pre[c] + ".";
      add(q + "query.weight", 2, d, d, 1, true);
      add(q + "query.bias", 1, d, 1, 1, false);
      add(q + "key.weight", 2, d, d, 1, true);
      add(q + "value.weight", 2, d, d, 1, true);
      add(q + "value.bias", 1, d, 1, 1, false);
      add(q + "out.weight", 2, d, d, 1, true);
      add(q + "out.bias", 1, d, 1, 1, false);
    }
    if (cross) {
      add(p + "cross_attn_ln.weight", 1, d, 1, 1, false);
      add(p + "cross_attn_ln.bias", 1, d, 1, 1, false);
    }
  };
  for (int i = 0; i < hp.n_audio_layer; ++i) block("encoder.blocks." + std::to_string(i) + ".", da, false);
  for (int i = 0; i < hp.n_text_layer; ++i) block("decoder.blocks." + std::to_string(i) + ".", dt, true);
}
}  // namespace detail

// returns WB_OK or a negative WsError code; `err` gets the WsError Display text (52-71)
inline int parse_model_file(const char* path, ModelFileView& mv, std::string& err) {
  FILE* f = fopen(path, "rb");
  if (!f) {
    err = std::string("Unexpected IO: cannot open '") + path + "'";
    return WB_ERR_IO;
  }
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (sz < 0) {
    fclose(f);
    err = "Unexpected IO: cannot size model file";
    return WB_ERR_IO;
  }
  mv.blob.resize((size_t)sz);
  const size_t got = fread(mv.blob.data(), 1, (size_t)sz, f);
  fclose(f);
  if (got != (size_t)sz) {
    err = "Unexpected IO: short read";
    return WB_ERR_IO;
  }
  const uint8_t* p = mv.blob.data();
  const uint8_t* end = p + mv.blob.size();
  auto need = [&](size_t n) { return (size_t)(end - p) >= n; };
  auto rd32 = [&](int32_t& v) {
    if (!need(4)) return false;
    memcpy(&v, p, 4);
    p += 4;
    return true;
  };
  const char* eof_msg = "Unexpected IO: failed to fill whole buffer";
  int32_t magic = 0;
  if (!rd32(magic)) { err = eof_msg; return WB_ERR_IO; }
  if ((uint32_t)magic != 0x67676d6cu) {
    err = std::string("invalid model file '") + path + "' (bad magic)\n";
    return WB_ERR_BAD_MAGIC;
  }
  int32_t h[11];
  for (int i = 0; i < 11; ++i)
    if (!rd32(h[i])) { err = eof_msg; return WB_ERR_IO; }
  memcpy(&mv.hp, h, sizeof(h));
  const ModelHParams& hp = mv.hp;
  if (hp.n_audio_state <= 0 || hp.n_audio_head <= 0 || hp.n_audio_state % hp.n_audio_head != 0 ||
      hp.n_text_state <= 0 || hp.n_text_head <= 0 || hp.n_text_state % hp.n_text_head != 0 || hp.n_vocab <= 0 ||
      hp.n_audio_ctx <= 0 || hp.n_text_ctx <= 0 || hp.n_mels <= 0 || hp.n_audio_layer < 0 || hp.n_text_layer < 0) {
    err = "Unexpected: implausible hparams";
    return WB_ERR_UNEXPECTED;
  }
  if (!rd32(mv.filt_n_mel) || !rd32(mv.filt_n_fft)) { err = eof_msg; return WB_ERR_IO; }
  if (mv.filt_n_mel <= 0 || mv.filt_n_fft <= 0 || !need((size_t)mv.filt_n_mel * mv.filt_n_fft * 4)) { err = eof_msg; return WB_ERR_IO; }
  mv.filters = reinterpret_cast<const float*>(p);
  p += (size_t)mv.filt_n_mel * mv.filt_n_fft * 4;
  if (!rd32(mv.n_vocab_file) || mv.n_vocab_file < 0) { err = eof_msg; return WB_ERR_IO; }
  mv.vocab.reserve((size_t)(hp.n_vocab > mv.n_vocab_file ? hp.n_vocab : mv.n_vocab_file));
  for (int i = 0; i < mv.n_vocab_file; ++i) {   // WhisperVocab::load (578-589): u32 length + bytes per token
    int32_t len = 0;
    if (!rd32(len) || len < 0 || !need((size_t)len)) { err = eof_msg; return WB_ERR_IO; }
    mv.vocab.emplace_back(reinterpret_cast<const char*>(p), (size_t)len);
    p += len;
  }
  // special-token fix-up (433-440).  The reference tests n_vocab == 51865 only (594-596) and shifts
  // eot, sot, prev, solm, not, beg by one.  A vocabulary with MORE language tokens (large-v3: 51866 = one
  // extra language) moves every id behind the language block by the number of extra languages as well --
  // upstream's `dt = num_languages - 98` rule: translate / transcribe and prev, solm, not, beg shift by
  // `extra`; eot and sot sit in front of the language block and keep the single shift.  With 51865 entries
  // this is exactly the reference's fix-up.
  if (hp.n_vocab >= 51865) {
    const int extra = hp.n_vocab - 51865;
    mv.special[0] += 1;                                   // eot
    mv.special[1] += 1;                                   // sot
    for (int i = 2; i < 6; ++i) mv.special[i] += 1 + extra;   // prev, solm, not, beg
    mv.special[6] += extra;                               // translate
    mv.special[7] += extra;                               // transcribe
  }
  // ids the file has no text for get the reference's placeholder names (442-467)
  for (int i = mv.n_vocab_file; i < hp.n_vocab; ++i) {
    const int eot = mv.special[0], sot = mv.special[1], prev = mv.special[2], tnot = mv.special[4], beg = mv.special[5];
    std::string w;
    if (i > beg) w = "[_TT_" + std::to_string(i - beg) + "]";
    else if (i == eot) w = "[_EOT_]";
    else if (i == sot) w = "[_SOT_]";
    else if (i == prev) w = "[_PREV_]";
    else if (i == tnot) w = "[_NOT_]";
    else if (i == beg) w = "[_BEG_]";
    else w = "[_extra_token_" + std::to_string(i) + "]";
    mv.vocab.push_back(w);
  }

  std::unordered_map<std::string, detail::Expect> table;
  detail::expect_table(hp, table);
  while (p < end) {
    int32_t n_dims = 0, name_len = 0, ftype = 0;
    if (!rd32(n_dims) || !rd32(name_len) || !rd32(ftype)) { err = eof_msg; return WB_ERR_IO; }
    if (n_dims < 1 || n_dims > 3 || name_len < 0 || !need((size_t)n_dims * 4 + (size_t)name_len)) {
      err = "Unexpected: malformed tensor record";
      return WB_ERR_UNEXPECTED;
    }
    HostTensor t;
    t.n_dims = n_dims;
    int64_t nelements = 1;
    for (int i = 0; i < n_dims; ++i) {
      int32_t v;
      rd32(v);
      t.ne[i] = v;
      nelements *= v;
    }
    const std::string name(reinterpret_cast<const char*>(p), (size_t)name_len);
    p += name_len;
    auto it = table.find(name);
    if (it == table.end()) {
      err = "unknown tensor '" + name + "' in model file\n";
      return WB_ERR_UNKNOWN_TENSOR;
    }
    const detail::Expect& ex = it->second;
    const int64_t expect_n = ex.ne[0] * ex.ne[1] * ex.ne[2];
    if (expect_n != nelements) {
      err = "tensor " + name + " has wrong size in model file, got:" + std::to_string(expect_n) +
            ", expected:" + std::to_string(nelements) + "\n";
      return WB_ERR_WRONG_SIZE_TENSOR;
    }
    for (int i = 0; i < ex.n_dims; ++i) {
      if (ex.ne[i] != t.ne[i]) {
        err = "tensor " + name + " has wrong shape in model file\n";
        return WB_ERR_WRONG_SHAPE_TENSOR;
      }
    }
    const bool expect_f16 = ex.weight && hp.f16 == 1;
    const size_t bpe = ftype == 0 ? 4 : 2;
    const size_t expect_bytes = (size_t)expect_n * (expect_f16 ? 2 : 4);
    if ((size_t)nelements * bpe != expect_bytes) {
      err = "tensor " + name + " has wrong bytes in model file, got:" + std::to_string(expect_bytes) +
            ", expected:" + std::to_string((size_t)nelements * bpe) + "\n";
      return WB_ERR_WRONG_BYTES_TENSOR;
    }
    if (!need(expect_bytes)) { err = eof_msg; return WB_ERR_IO; }
    t.f16 = expect_f16;
    t.data = p;
    t.bytes = expect_bytes;
    p += expect_bytes;
    mv.tensors[name] = t;
  }
  for (auto& kv : table) {
    if (!mv.tensors.count(kv.first)) {
      err = "invalid ref tensor '" + kv.first + "'\n";   // declared but never filled
      return WB_ERR_BAD_REF_TENSOR;
    }
  }
  return WB_OK;
}

}  // namespace wb
