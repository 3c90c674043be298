// reference_main.cpp -- the reference's `fn main` and `test_load_model` (src/main.rs:2065-2075, 2081-2091) written
// against include/whisper_b200.hpp, the C++ host side that carries the reference's names:
//
//     let s16: Vec<i16> = reader.samples::<i16>() ...;
//     let samples = convert_integer_to_float_audio(&s16);
//     let mut wctx = WhisperContext::new(model_path).unwrap();
//     whisper_pcm_to_mel(&mut wctx, Arc::new(samples)).unwrap();
//     whisper_encode(&mut wctx, 1, 0).unwrap();
//
//   reference_main <model.bin> <pcm_s16.raw> <out_prefix>      run it (dumps mel / encoder output / logits for the
//                                                              parity test, tests/test_gpu_cabi.py)
//   reference_main --errors <dir>                              loader error behaviour on the bad files in <dir>
// Exit code 0 = ok; WsError escaping main prints "<variant>: <Display text>" and exits with -code.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include "whisper_b200.hpp"

using namespace whisper_b200;

static void dump(const std::string& path, const std::vector<float>& v) {
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char*>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(float)));
}

static int run(const char* model_path, const char* pcm_path, const std::string& prefix) {
  std::ifstream in(pcm_path, std::ios::binary | std::ios::ate);
  if (!in) return 65;
  const std::streamsize bytes = in.tellg();
  in.seekg(0);
  std::vector<int16_t> s16(static_cast<size_t>(bytes) / 2);
  in.read(reinterpret_cast<char*>(s16.data()), bytes);
  std::printf("len:%zu\n", s16.size());                                         // 2069
  auto samples = std::make_shared<const std::vector<float>>(convert_integer_to_float_audio(s16));   // 2070
  // (the reference's context is sized for a 30 s clip; a shorter test clip needs no more)
  WhisperContext wctx = WhisperContext::with_capacity(model_path, 0, 1, 1, static_cast<int64_t>(samples->size()) < 400 ? 400 : static_cast<int64_t>(samples->size()));   // 2072
  whisper_pcm_to_mel(wctx, samples);                                            // 2073
  whisper_encode(wctx, 1, 0);                                                   // 2074
  int n_mel = 0, n_len = 0;
  dump(prefix + ".mel.f32", wctx.mel(0, &n_mel, &n_len));
  dump(prefix + ".enc.f32", wctx.encoder_out(0));
  const int32_t first = wctx.tokens.sot < wctx.hparams.n_vocab ? wctx.tokens.sot : 7;   // micro test vocabularies are tiny
  whisper_decode(wctx, {first}, 0, 1);
  dump(prefix + ".logits.f32", wctx.logits);
  const auto toks = whisper_decode_greedy(wctx, {first}, 4, -1);
  const wb_timings tm = wctx.timings();
  std::printf("reference_main ok: n_mel=%d n_len=%d d=%d tokens=%zu t_mel_us=%lld t_encode_us=%lld launches=%lld\n", n_mel, n_len,
              wctx.hparams.n_audio_state, toks.size(), static_cast<long long>(tm.t_mel_us), static_cast<long long>(tm.t_encode_us),
              static_cast<long long>(tm.n_kernel_launches));
  return 0;
}

// every loader failure surfaces as the reference's WsError variant with its Display text (src/main.rs:50-72)
static int errors(const std::string& dir) {
  struct Case { const char* file; WsErrorKind kind; const char* text; };
  const Case cases[] = {
      {"bad_magic.bin", WsErrorKind::BadMagic, "bad magic"},
      {"nope.bin", WsErrorKind::UnexpectIO, "cannot open"},
      {"unknown.bin", WsErrorKind::UnknownTensor, "unknown tensor"},
      {"size.bin", WsErrorKind::WrongSizeTensor, "wrong size"},
      {"shape.bin", WsErrorKind::WrongShapeTensor, "wrong shape"},
      {"bytes.bin", WsErrorKind::WrongBytesTensor, "wrong bytes"},
  };
  int bad = 0;
  for (const Case& c : cases) {
    try {
      WhisperContext ctx = WhisperContext::new_(dir + "/" + c.file);
      std::printf("%s: no error\n", c.file);
      ++bad;
    } catch (const WsError& e) {
      const bool ok = e.kind == c.kind && std::strstr(e.what(), c.text) != nullptr;
      std::printf("%s: %s (%d) %s\n", c.file, e.variant(), e.code, ok ? "ok" : "UNEXPECTED");
      bad += ok ? 0 : 1;
    }
  }
  return bad;
}

int main(int argc, char** argv) {
  try {
    if (argc == 3 && std::strcmp(argv[1], "--errors") == 0) return errors(argv[2]);
    if (argc == 4) return run(argv[1], argv[2], argv[3]);
    std::fprintf(stderr, "usage: %s model.bin pcm_s16.raw out_prefix | --errors dir\n", argv[0]);
    return 64;
  } catch (const WsError& e) {                                                  // the reference's .unwrap()
    std::fprintf(stderr, "%s: %s\n", e.variant(), e.what());
    return -e.code;
  }
}
