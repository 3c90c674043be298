/* main_replay.c -- the reference's `fn main` (src/main.rs:2065-2075) replayed in C against
 * include/whisper_b200.h and libwhisper_b200.so: WhisperContext::new -> whisper_pcm_to_mel ->
 * whisper_encode(ctx, 1, 0), then one whisper_decode step on the state the reference only declares.
 * Compiled with gcc (no CUDA headers, no C++): what a maintainer's FFI layer sees.
 *
 *   main_replay <model.bin> <pcm_f32.raw> <out_prefix>
 *
 * Writes <out_prefix>.mel.f32, <out_prefix>.enc.f32 and <out_prefix>.logits.f32 for the parity test
 * (tests/test_gpu_cabi.py compares them with the oracle), and prints one summary line.  Exit code
 * = -(the failing call's WsError code), 0 on success.  No CPU fallback: without a B200 wb_ctx_create
 * fails with WB_ERR_TENSOR_OP and this program exits 10. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "whisper_b200.h"

static int fail(const char* what, int rc, const wb_ctx* ctx) {
  fprintf(stderr, "%s failed (%d): %s\n", what, rc, wb_last_error(ctx));
  return -rc;
}

static int dump(const char* prefix, const char* suffix, const float* p, size_t n) {
  char path[4096];
  snprintf(path, sizeof(path), "%s.%s", prefix, suffix);
  FILE* f = fopen(path, "wb");
  if (!f) return 1;
  const size_t w = fwrite(p, sizeof(float), n, f);
  fclose(f);
  return w == n ? 0 : 1;
}

int main(int argc, char** argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: %s model.bin pcm_f32.raw out_prefix\n", argv[0]);
    return 64;
  }
  /* the clip (the reference reads a wav with hound, 2067-2070; here raw f32 samples) */
  FILE* f = fopen(argv[2], "rb");
  if (!f) return 65;
  fseek(f, 0, SEEK_END);
  const long bytes = ftell(f);
  fseek(f, 0, SEEK_SET);
  const size_t n_samples = (size_t)bytes / sizeof(float);
  float* pcm = (float*)malloc(n_samples * sizeof(float));
  if (!pcm || fread(pcm, sizeof(float), n_samples, f) != n_samples) return 66;
  fclose(f);

  wb_config cfg;
  wb_config_default(&cfg);
  cfg.max_segments = 1;
  cfg.max_clips = 1;
  cfg.max_clip_samples = (int64_t)n_samples;
  cfg.decode_capacity = 1;
  wb_ctx* ctx = NULL;
  int rc = wb_ctx_create(argv[1], &cfg, &ctx);                       /* 2066: WhisperContext::new */
  if (rc != WB_OK) return fail("wb_ctx_create", rc, NULL);
  int32_t hp[11], special[8];
  wb_get_hparams(ctx, hp);
  wb_get_special_tokens(ctx, special);
  const int n_vocab = hp[0], n_audio_ctx = hp[1], d = hp[2];

  rc = wb_pcm_to_mel(ctx, pcm, n_samples, 1);                         /* 2072: whisper_pcm_to_mel */
  if (rc != WB_OK) return fail("wb_pcm_to_mel", rc, ctx);
  int n_mel = 0, n_len = 0, n_clips = 0;
  wb_mel_dims(ctx, &n_mel, &n_len, &n_clips);
  float* mel = (float*)malloc((size_t)n_mel * n_len * sizeof(float));
  rc = wb_mel_read(ctx, 0, mel, (size_t)n_mel * n_len);
  if (rc != WB_OK) return fail("wb_mel_read", rc, ctx);

  const int32_t clip0 = 0;
  const size_t off0 = 0;
  rc = wb_encode(ctx, &clip0, &off0, 1);                              /* 2074: whisper_encode(ctx, 1, 0) */
  if (rc != WB_OK) return fail("wb_encode", rc, ctx);
  float* enc = (float*)malloc((size_t)n_audio_ctx * d * sizeof(float));
  rc = wb_encoder_out_read(ctx, 0, enc);
  if (rc != WB_OK) return fail("wb_encoder_out_read", rc, ctx);

  /* [sot]; the micro test architecture's vocabulary is smaller than the special-token ids: any valid id then */
  const int32_t prompt[1] = {special[1] < n_vocab ? special[1] : 7};
  rc = wb_decode(ctx, prompt, 1, 0, 1);
  if (rc != WB_OK) return fail("wb_decode", rc, ctx);
  float* logits = (float*)malloc((size_t)n_vocab * sizeof(float));
  rc = wb_logits_read(ctx, 0, logits);
  if (rc != WB_OK) return fail("wb_logits_read", rc, ctx);

  wb_timings tm;
  wb_timings_get(ctx, &tm);
  if (dump(argv[3], "mel.f32", mel, (size_t)n_mel * n_len) || dump(argv[3], "enc.f32", enc, (size_t)n_audio_ctx * d) ||
      dump(argv[3], "logits.f32", logits, (size_t)n_vocab))
    return 67;
  printf("main_replay ok: %s n_mel=%d n_len=%d n_ctx=%d d=%d n_vocab=%d t_mel_us=%lld t_encode_us=%lld t_decode_us=%lld launches=%lld\n",
         wb_version(), n_mel, n_len, n_audio_ctx, d, n_vocab, (long long)tm.t_mel_us, (long long)tm.t_encode_us,
         (long long)tm.t_decode_us, (long long)tm.n_kernel_launches);
  wb_ctx_free(ctx);
  free(pcm); free(mel); free(enc); free(logits);
  return 0;
}
