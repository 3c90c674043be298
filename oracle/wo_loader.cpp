// wo_loader.cpp -- restatement of the reference's model-file loader.  TEST INFRASTRUCTURE ONLY.
//
// Follows WhisperContext::new (src/main.rs:366-503), WhisperHparams::load (622-658),
// WhisperFilters::load (513-535), WhisperVocab::load (578-592) and WhisperModel::load
// (809-1483): the tensor table is declared first (960-1334), then records are matched by
// name and checked for element count, per-dimension shape and byte size (1401-1434).
#include <cstdio>
#include <memory>

#include "wo_common.hpp"

namespace wo {

namespace {
struct Reader {
  FILE* f;
  bool rd(void* p, size_t n) { return fread(p, 1, n, f) == n; }
  bool i32(int32_t& v) { return rd(&v, 4); }
  bool u32(uint32_t& v) { return rd(&v, 4); }
};

void declare(Model& m, const std::string& name, std::vector<int> ne, bool f16) {
  Tensor t;
  t.ne = std::move(ne);
  t.f16 = f16;
  m.t.emplace(name, std::move(t));
}

// src/main.rs:945-1334
void declare_tensors(Model& m) {
  const HParams& hp = m.hp;
  const bool w16 = hp.f16 == 1;  // 817-821
  const int da = hp.n_audio_state, dt = hp.n_text_state;
  declare(m, "encoder.positional_embedding", {da, hp.n_audio_ctx}, false);   // 960
  declare(m, "encoder.conv1.weight", {3, hp.n_mels, da}, w16);               // 961
  declare(m, "encoder.conv1.bias", {1, da}, false);                          // 962
  declare(m, "encoder.conv2.weight", {3, da, da}, w16);                      // 964-965
  declare(m, "encoder.conv2.bias", {1, da}, false);                          // 966
  declare(m, "encoder.ln_post.weight", {da}, false);                         // 968
  declare(m, "encoder.ln_post.bias", {da}, false);                           // 969
  auto block = [&](const std::string& p, int d, bool cross) {
    declare(m, p + "mlp_ln.weight", {d}, false);
    declare(m, p + "mlp_ln.bias", {d}, false);
    declare(m, p + "mlp.0.weight", {d, 4 * d}, w16);
    declare(m, p + "mlp.0.bias", {4 * d}, false);
    declare(m, p + "mlp.2.weight", {4 * d, d}, w16);
    declare(m, p + "mlp.2.bias", {d}, false);
    declare(m, p + "attn_ln.weight", {d}, false);
    declare(m, p + "attn_ln.bias", {d}, false);
    declare(m, p + "attn.query.weight", {d, d}, w16);
    declare(m, p + "attn.query.bias", {d}, false);
    declare(m, p + "attn.key.weight", {d, d}, w16);   // no key bias (675, 704)
    declare(m, p + "attn.value.weight", {d, d}, w16);
    declare(m, p + "attn.value.bias", {d}, false);
    declare(m, p + "attn.out.weight", {d, d}, w16);
    declare(m, p + "attn.out.bias", {d}, false);
    if (cross) {
      declare(m, p + "cross_attn_ln.weight", {d}, false);
      declare(m, p + "cross_attn_ln.bias", {d}, false);
      declare(m, p + "cross_attn.query.weight", {d, d}, w16);
      declare(m, p + "cross_attn.query.bias", {d}, false);
      declare(m, p + "cross_attn.key.weight", {d, d}, w16);
      declare(m, p + "cross_attn.value.weight", {d, d}, w16);
      declare(m, p + "cross_attn.value.bias", {d}, false);
      declare(m, p + "cross_attn.out.weight", {d, d}, w16);
      declare(m, p + "cross_attn.out.bias", {d}, false);
    }
  };
  for (int i = 0; i < hp.n_audio_layer; ++i)                                  // 1006-1136
    block("encoder.blocks." + std::to_string(i) + ".", da, false);
  declare(m, "decoder.positional_embedding", {dt, hp.n_text_ctx}, false);     // 1139
  declare(m, "decoder.token_embedding.weight", {dt, hp.n_vocab}, w16);        // 1141
  declare(m, "decoder.ln.weight", {dt}, false);                               // 1142
  declare(m, "decoder.ln.bias", {dt}, false);                                 // 1143
  for (int i = 0; i < hp.n_text_layer; ++i)                                   // 1160-1333
    block("decoder.blocks." + std::to_string(i) + ".", dt, true);
}
}  // namespace

const Tensor& Model::get(const std::string& name) const { return t.at(name); }

void Tensor::to_f32(std::vector<float>& out) const {
  size_t n = nelem();
  out.resize(n);
  if (f16) {
    const uint16_t* p = h();
    for (size_t i = 0; i < n; ++i) out[i] = f16_bits_to_f32(p[i]);
  } else {
    std::memcpy(out.data(), data.data(), n * 4);
  }
}

int load_model(const char* path, Model& m, std::string& err) {
  FILE* fp = fopen(path, "rb");
  if (!fp) {
    err = std::string("Unexpected IO: cannot open ") + path;
    return ORC_ERR_IO;
  }
  std::unique_ptr<FILE, int (*)(FILE*)> guard(fp, fclose);
  Reader r{fp};
  uint32_t magic = 0;
  if (!r.u32(magic)) { err = "Unexpected IO: short read (magic)"; return ORC_ERR_IO; }
  if (magic != 0x67676d6cu) {                                                // 368-371
    err = std::string("invalid model file '") + path + "' (bad magic)";
    return ORC_ERR_BAD_MAGIC;
  }
  int32_t hp[11];
  for (int i = 0; i < 11; ++i)                                                // 622-633
    if (!r.i32(hp[i])) { err = "Unexpected IO: short read (hparams)"; return ORC_ERR_IO; }
  std::memcpy(&m.hp, hp, sizeof(hp));
  int32_t n_mel = 0, n_fft = 0;                                               // 513-524
  if (!r.i32(n_mel) || !r.i32(n_fft) || n_mel <= 0 || n_fft <= 0) {
    err = "Unexpected IO: short read (filters)";
    return ORC_ERR_IO;
  }
  m.filt_n_mel = n_mel;
  m.filt_n_fft = n_fft;
  m.filters.resize((size_t)n_mel * n_fft);
  if (!r.rd(m.filters.data(), m.filters.size() * 4)) { err = "Unexpected IO: short read (filters)"; return ORC_ERR_IO; }
  int32_t n_vocab_file = 0;                                                   // 430-431
  if (!r.i32(n_vocab_file) || n_vocab_file < 0) { err = "Unexpected IO: short read (vocab)"; return ORC_ERR_IO; }
  m.vocab = Vocab();
  m.vocab.id_to_token.resize(n_vocab_file);
  for (int i = 0; i < n_vocab_file; ++i) {                                    // 578-589
    uint32_t len = 0;
    if (!r.u32(len)) { err = "Unexpected IO: short read (vocab)"; return ORC_ERR_IO; }
    std::string w(len, '\0');
    if (len && !r.rd(&w[0], len)) { err = "Unexpected IO: short read (vocab)"; return ORC_ERR_IO; }
    m.vocab.id_to_token[i] = std::move(w);
  }
  m.vocab.n_vocab = m.hp.n_vocab;                                             // 432
  // is_multilingual (594-596) tests n_vocab == 51865 only; large-v3 carries 51866 and is
  // multilingual too, so the oracle generalises to >= 51865 -- an
  // oracle decision, since the reference would treat v3 as an English-only vocabulary.
  // Every extra language token beyond the 99 of a 51865-entry vocabulary moves the ids behind the
  // language block once more (upstream's dt = num_languages - 98 rule): large-v3's <|notimestamps|> is
  // 50364 and its first time stamp 50365.
  if (m.vocab.n_vocab >= 51865) {                                             // 433-440
    const int extra = m.vocab.n_vocab - 51865;
    m.vocab.token_eot += 1;
    m.vocab.token_sot += 1;
    m.vocab.token_prev += 1 + extra;
    m.vocab.token_solm += 1 + extra;
    m.vocab.token_not += 1 + extra;
    m.vocab.token_beg += 1 + extra;
    m.vocab.token_translate += extra;
    m.vocab.token_transcribe += extra;
  }
  declare_tensors(m);
  // records until EOF (1384-1475; true EOF rather than the reference's fill_buf() < 12 test)
  for (;;) {
    int32_t n_dims = 0, name_len = 0, ftype = 0;
    if (!r.i32(n_dims)) break;  // clean EOF
    if (!r.i32(name_len) || !r.i32(ftype)) { err = "Unexpected IO: short read (record header)"; return ORC_ERR_IO; }
    if (n_dims < 1 || n_dims > 3 || name_len < 0 || name_len > 4096) {
      err = "Unexpected: malformed tensor record";
      return ORC_ERR_UNEXPECTED;
    }
    size_t nelements = 1;
    int32_t ne[3] = {1, 1, 1};
    for (int i = 0; i < n_dims; ++i) {
      if (!r.i32(ne[i])) { err = "Unexpected IO: short read (ne)"; return ORC_ERR_IO; }
      nelements *= (size_t)ne[i];
    }
    std::string name(name_len, '\0');
    if (name_len && !r.rd(&name[0], name_len)) { err = "Unexpected IO: short read (name)"; return ORC_ERR_IO; }
    auto it = m.t.find(name);
    if (it == m.t.end()) {                                                    // 1401-1403
      err = "unknown tensor '" + name + "' in model file";
      return ORC_ERR_UNKNOWN_TENSOR;
    }
    Tensor& t = it->second;
    if (t.nelem() != nelements) {                                             // 1406-1412
      err = "tensor " + name + " has wrong size in model file";
      return ORC_ERR_WRONG_SIZE_TENSOR;
    }
    for (size_t i = 0; i < t.ne.size(); ++i) {                                // 1413-1422
      if (t.ne[i] != ne[i]) {
        err = "tensor " + name + " has wrong shape in model file";
        return ORC_ERR_WRONG_SHAPE_TENSOR;
      }
    }
    size_t bpe = ftype == 0 ? 4 : 2;                                          // 1423-1427
    size_t expect_bytes = t.nelem() * (t.f16 ? 2 : 4);
    if (nelements * bpe != expect_bytes) {                                    // 1428-1434
      err = "tensor " + name + " has wrong bytes in model file";
      return ORC_ERR_WRONG_BYTES_TENSOR;
    }
    t.data.resize(expect_bytes);
    if (!r.rd(t.data.data(), expect_bytes)) { err = "Unexpected IO: short read (tensor data)"; return ORC_ERR_IO; }  // 1437
    t.loaded = true;
  }
  for (auto& kv : m.t) {
    if (!kv.second.loaded) {
      err = "Unexpected: tensor '" + kv.first + "' missing from model file";
      return ORC_ERR_UNEXPECTED;
    }
  }
  return ORC_OK;
}

}  // namespace wo
