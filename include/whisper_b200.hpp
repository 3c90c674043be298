// whisper_b200.hpp -- header-only C++17 host side over the C-ABI of whisper_b200.h, carrying the reference's own
// names, argument meaning and error behaviour (szuwgh/whisper.rs, src/main.rs):
//
//     WhisperContext::new(fname) -> WsResult<WhisperContext>              src/main.rs:366
//     whisper_pcm_to_mel(ctx, samples) -> WsResult<()>                     src/main.rs:1681
//     whisper_encode(ctx, n_threads, mel_offset) -> WsResult<()>           src/main.rs:1799
//     whisper_decode(ctx, tokens, n_past, n_threads) -> WsResult<()>       (declared state only: 351-352, 694-731)
//
// The reference is compiled code (Rust) and this image has no Rust toolchain, so the compiled host side is this
// header (the Rust crates of the same shape ship as source under whisper.rs_b200/rust/).  Rust's `WsResult<T>` with
// `?` propagation becomes a thrown `WsError` whose `kind` is the reference's variant (src/main.rs:50-72) and whose
// what() is the reference's Display text; `&mut WhisperContext` exclusivity becomes a move-only class used from one
// thread at a time.  Results stay inside the context, as in the reference, and are read through accessors.
// There is no CPU path: without a B200 `WhisperContext::new_` throws WsError{WrongGTensor}.
#ifndef WHISPER_B200_HPP
#define WHISPER_B200_HPP

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "whisper_b200.h"

namespace whisper_b200 {

// WsError (src/main.rs:50-72)
enum class WsErrorKind {
  Unexpected, UnexpectIO, BadMagic, NotEnoughSpace, UnknownTensor, BadRefTensor, WrongSizeTensor, WrongShapeTensor,
  WrongBytesTensor, WrongGTensor
};

class WsError : public std::runtime_error {
 public:
  WsError(int code, const std::string& msg) : std::runtime_error(msg), code(code), kind(kind_of(code)) {}
  int code;            // the WB_ERR_* value the C-ABI returned
  WsErrorKind kind;
  const char* variant() const {
    static const char* names[] = {"Unexpected", "UnexpectIO", "BadMagic", "NotEnoughSpace", "UnknownTensor", "BadRefTensor",
                                  "WrongSizeTensor", "WrongShapeTensor", "WrongBytesTensor", "WrongGTensor"};
    return names[static_cast<int>(kind)];
  }
  static WsErrorKind kind_of(int code) {
    switch (code) {
      case WB_ERR_IO: return WsErrorKind::UnexpectIO;
      case WB_ERR_BAD_MAGIC: return WsErrorKind::BadMagic;
      case WB_ERR_NOT_ENOUGH_SPACE: return WsErrorKind::NotEnoughSpace;
      case WB_ERR_UNKNOWN_TENSOR: return WsErrorKind::UnknownTensor;
      case WB_ERR_BAD_REF_TENSOR: return WsErrorKind::BadRefTensor;
      case WB_ERR_WRONG_SIZE_TENSOR: return WsErrorKind::WrongSizeTensor;
      case WB_ERR_WRONG_SHAPE_TENSOR: return WsErrorKind::WrongShapeTensor;
      case WB_ERR_WRONG_BYTES_TENSOR: return WsErrorKind::WrongBytesTensor;
      case WB_ERR_TENSOR_OP: return WsErrorKind::WrongGTensor;
      default: return WsErrorKind::Unexpected;
    }
  }
};

// WhisperHparams (src/main.rs:607-619), as carried in the file header
struct WhisperHparams {
  int32_t n_vocab, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer;
  int32_t n_text_ctx, n_text_state, n_text_head, n_text_layer, n_mels, f16;
};

// special token ids (WhisperVocab, src/main.rs:557-575 after the fix-up of 433-440)
struct WhisperTokens {
  int32_t eot, sot, prev, solm, not_, beg, translate, transcribe;
};

// convert_integer_to_float_audio (src/main.rs:1673-1679)
inline std::vector<float> convert_integer_to_float_audio(const std::vector<int16_t>& samples) {
  std::vector<float> out(samples.size());
  for (size_t i = 0; i < samples.size(); ++i) out[i] = static_cast<float>(samples[i]) / 32768.0f;
  return out;
}

// WhisperContext (src/main.rs:333-363) resident on one B200
class WhisperContext {
 public:
  // WhisperContext::new (366): one 30 s window, one clip, decoder enabled -- the reference's shape.  (`new` is a
  // keyword in C++, hence the underscore.)
  static WhisperContext new_(const std::string& fname) { return with_capacity(fname, 0, 1, 1, 480000); }

  static WhisperContext with_capacity(const std::string& fname, int device, int max_segments, int max_clips,
                                      int64_t max_clip_samples, bool decode_capacity = true) {
    wb_config cfg;
    wb_config_default(&cfg);
    cfg.device = device;
    cfg.max_segments = max_segments;
    cfg.max_clips = max_clips;
    cfg.max_clip_samples = max_clip_samples;
    cfg.decode_capacity = decode_capacity ? 1 : 0;
    wb_ctx* h = nullptr;
    const int rc = wb_ctx_create(fname.c_str(), &cfg, &h);
    if (rc != WB_OK) throw WsError(rc, wb_last_error(nullptr));
    WhisperContext c(h);
    int32_t hp[11];
    wb_get_hparams(h, hp);
    c.hparams = WhisperHparams{hp[0], hp[1], hp[2], hp[3], hp[4], hp[5], hp[6], hp[7], hp[8], hp[9], hp[10]};
    int32_t st[8];
    wb_get_special_tokens(h, st);
    c.tokens = WhisperTokens{st[0], st[1], st[2], st[3], st[4], st[5], st[6], st[7]};
    return c;
  }

  WhisperContext(WhisperContext&& o) noexcept { *this = std::move(o); }
  WhisperContext& operator=(WhisperContext&& o) noexcept {
    if (this != &o) {
      release();
      h_ = o.h_;
      o.h_ = nullptr;
      hparams = o.hparams;
      tokens = o.tokens;
      logits = std::move(o.logits);
      samples_ = std::move(o.samples_);
    }
    return *this;
  }
  WhisperContext(const WhisperContext&) = delete;
  WhisperContext& operator=(const WhisperContext&) = delete;
  ~WhisperContext() { release(); }

  WhisperHparams hparams{};
  WhisperTokens tokens{};
  std::vector<float> logits;   // logits of the last decoded position (src/main.rs:351), filled by whisper_decode

  // ---- read-backs of state the reference keeps inside the context
  // mel.data (1633): [n_mel][n_len] f32
  std::vector<float> mel(int clip, int* n_mel_out = nullptr, int* n_len_out = nullptr) {
    int n_mel = 0, n_len = 0, n_clips = 0;
    wb_mel_dims(h_, &n_mel, &n_len, &n_clips);
    std::vector<float> v(static_cast<size_t>(n_mel) * n_len);
    check(wb_mel_read(h_, clip, v.data(), v.size()));
    if (n_mel_out) *n_mel_out = n_mel;
    if (n_len_out) *n_len_out = n_len;
    return v;
  }
  // `cur` after ln_post (1980-1984): [n_audio_ctx][n_audio_state] f32; the reference drops it with buf_compute
  std::vector<float> encoder_out(int seg = 0) {
    std::vector<float> v(static_cast<size_t>(hparams.n_audio_ctx) * hparams.n_audio_state);
    check(wb_encoder_out_read(h_, seg, v.data()));
    return v;
  }
  wb_timings timings() const {   // t_*_us (334-339)
    wb_timings t;
    wb_timings_get(h_, &t);
    return t;
  }
  std::string token_text(int32_t id) const {   // WhisperVocab::id_to_token (544)
    char buf[256];
    const int n = wb_token_text(h_, id, buf, sizeof(buf));
    if (n < 0) throw WsError(n, "token id out of range");
    return std::string(buf, static_cast<size_t>(n < 255 ? n : 255));
  }

  wb_ctx* handle() { return h_; }
  void check(int rc) const {
    if (rc != WB_OK) throw WsError(rc, wb_last_error(h_));
  }
  // the samples of the last whisper_pcm_to_mel stay alive until the upload has certainly finished (the reference
  // shares them with its mel threads through the same Arc, 1584)
  void hold(std::shared_ptr<const std::vector<float>> s) { samples_ = std::move(s); }

 private:
  explicit WhisperContext(wb_ctx* h) : h_(h) {}
  void release() {
    if (h_) wb_ctx_free(h_);
    h_ = nullptr;
  }
  wb_ctx* h_ = nullptr;
  std::shared_ptr<const std::vector<float>> samples_;
};

// whisper_pcm_to_mel (src/main.rs:1681): `samples` is the reference's Arc<Vec<f32>>; the mel stays inside the context
inline void whisper_pcm_to_mel(WhisperContext& ctx, std::shared_ptr<const std::vector<float>> samples) {
  ctx.check(wb_pcm_to_mel(ctx.handle(), samples->data(), samples->size(), 1));
  ctx.hold(std::move(samples));
}

// whisper_encode (src/main.rs:1799): `n_threads` is accepted and ignored, exactly as in the reference (1799, 2074)
inline void whisper_encode(WhisperContext& wctx, size_t /*n_threads*/, size_t mel_offset) {
  const int32_t clip = 0;
  wctx.check(wb_encode(wctx.handle(), &clip, &mel_offset, 1));
}

// whisper_decode: the step implied by the reference's state; fills ctx.logits (351) for the last token
inline void whisper_decode(WhisperContext& ctx, const std::vector<int32_t>& tokens, size_t n_past, size_t /*n_threads*/) {
  ctx.check(wb_decode(ctx.handle(), tokens.data(), static_cast<int>(tokens.size()), static_cast<int>(n_past), 1));
  ctx.logits.resize(static_cast<size_t>(ctx.hparams.n_vocab));
  ctx.check(wb_logits_read(ctx.handle(), 0, ctx.logits.data()));
}

// device-side greedy loop (arg-max over all logits, stop at `eot` or `max_new`)
inline std::vector<int32_t> whisper_decode_greedy(WhisperContext& ctx, const std::vector<int32_t>& prompt, size_t max_new,
                                                  int32_t eot) {
  std::vector<int32_t> toks(max_new);
  int32_t len = 0;
  ctx.check(wb_decode_greedy(ctx.handle(), prompt.data(), static_cast<int>(prompt.size()), static_cast<int>(max_new), eot, 1,
                             toks.data(), nullptr, &len));
  toks.resize(static_cast<size_t>(len));
  return toks;
}

}  // namespace whisper_b200

#endif
