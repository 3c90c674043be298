/*
 * oracle.h -- C API of the CPU oracle.  TEST INFRASTRUCTURE ONLY.
 *
 * This library is a CPU restatement of the reference's algorithm for the Whisper hot path
 * (szuwgh/whisper.rs, src/main.rs): model-file loader (366-503, 513-535, 578-592, 622-658,
 * 809-1483), log-mel front end (1487-1707), encoder (1799-2063), plus the decode step the
 * reference leaves unimplemented (scaffolding at 694-731, 1336-1354; semantics of upstream
 * whisper.cpp v1.0.3, the code base main.rs transliterates -- see its paths at 2066-2088).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libwhisper_b200.so) never links, loads or calls it.
 *
 * PARITY STATUS
 *   mel (1487-1671): fully specified by the reference source; restated operation by
 *       operation (same recursion, same f32 evaluation order, no FMA contraction).  The
 *       reference cannot be built here (no rustc; path dependency `galois` absent), so even
 *       this part is checked only against an independent f64 evaluation -- parity unpinned.
 *   encoder / decoder arithmetic: lives in the un-vendored crate `galois` 0.1.0
 *       (Cargo.toml:13, no Cargo.lock).  Restated from the identically named ggml ops of
 *       whisper.cpp v1.0.3.  The reference holds no golden vectors,
 *       known-answer tests or fixtures for this path (its two tests assert nothing,
 *       2081-2117) -- PARITY UNPINNED.
 */
#ifndef WHISPER_ORACLE_H
#define WHISPER_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

/* error codes mirror WsError (src/main.rs:50-72) */
enum {
  ORC_OK = 0,
  ORC_ERR_UNEXPECTED = -1,
  ORC_ERR_IO = -2,
  ORC_ERR_BAD_MAGIC = -3,
  ORC_ERR_NOT_ENOUGH_SPACE = -4,
  ORC_ERR_UNKNOWN_TENSOR = -5,
  ORC_ERR_BAD_REF_TENSOR = -6,
  ORC_ERR_WRONG_SIZE_TENSOR = -7,
  ORC_ERR_WRONG_SHAPE_TENSOR = -8,
  ORC_ERR_WRONG_BYTES_TENSOR = -9,
  ORC_ERR_TENSOR_OP = -10
};

/* switchable rounding points (SURVEY.md section 7 "hard parts"); defaults = canonical */
enum {
  ORC_OPT_ACT_F16_ROUND = 0, /* 1 (default): matmul/conv round the activation operand to F16 */
  ORC_OPT_GELU_MODE = 1,     /* 0 (default): tanh via F16 LUT; 1: tanh f32; 2: erf f32 */
  ORC_OPT_SOFTMAX_EXP = 2,   /* 0 (default): exp via F16 LUT; 1: expf f32 */
  ORC_OPT_PROB_F16_ROUND = 3 /* 1 (default): probabilities rounded to F16 before P*V */
};

/* checkpoints mirroring the author's sum|x| probes (src/main.rs:1439-1454, 1836-1849, ...) */
enum {
  ORC_STAGE_MEL = 0,
  ORC_STAGE_CONV1 = 1,
  ORC_STAGE_CONV2_POS = 2,
  ORC_STAGE_LAYER = 3,   /* residual stream after encoder block `layer` */
  ORC_STAGE_LN_POST = 4,
  ORC_STAGE_CROSS_K = 5, /* per text layer */
  ORC_STAGE_CROSS_V = 6
};

int orc_ctx_create(const char* model_path, orc_ctx** out);       /* WhisperContext::new, 366 */
void orc_ctx_free(orc_ctx* ctx);
int orc_get_hparams(const orc_ctx* ctx, int32_t out[11]);         /* 607-619 order */
int orc_get_special_tokens(const orc_ctx* ctx, int32_t out[8]);   /* eot,sot,prev,solm,not,beg,translate,transcribe (557-575, 433-440) */
int orc_set_option(orc_ctx* ctx, int opt, int value);
int orc_set_audio_ctx(orc_ctx* ctx, int n_ctx);                   /* exp_n_audio_ctx, src/main.rs:362, 1803-1807 */

/* whisper_pcm_to_mel (1681): whole clip, n_threads frame-strided workers (reference uses 4) */
int orc_pcm_to_mel(orc_ctx* ctx, const float* pcm, size_t n_samples, int n_threads);
int orc_mel_dims(const orc_ctx* ctx, int* n_mel, int* n_len);
int orc_mel_read(const orc_ctx* ctx, float* out);                 /* [n_mel][n_len], layout of 1633 */
int orc_mel_set(orc_ctx* ctx, const float* mel, int n_mel, int n_len);

/* whisper_encode (1799): one 2*n_ctx-frame window at mel_offset */
int orc_encode(orc_ctx* ctx, int n_threads, size_t mel_offset);
int orc_encoder_out_read(const orc_ctx* ctx, float* out);         /* ln_post output [n_ctx][d] */
int orc_cross_kv_read(const orc_ctx* ctx, int layer, uint16_t* k, uint16_t* v); /* F16 bits [n_ctx][d], 2018-2030 */
int orc_checksum(const orc_ctx* ctx, int stage, int layer, double* abs_sum);

/* decode step (absent in the reference; upstream semantics) */
int orc_decode(orc_ctx* ctx, const int32_t* tokens, int n_tokens, int n_past, int n_threads);
int orc_logits_read(const orc_ctx* ctx, float* out);              /* [n_vocab], last position */
int orc_decode_greedy(orc_ctx* ctx, const int32_t* prompt, int n_prompt, int max_new, int eot,
                      int n_threads, int32_t* out_tokens, float* out_margin, int* out_len);

/* stand-alone pieces, for pinning the restatement against independent evaluations */
void orc_fft(const float* in, int n, float* out /* 2n interleaved re,im */);   /* 1505-1551 */
void orc_dft(const float* in, int n, float* out);                               /* 1487-1502 */
float orc_f16_round(float x);
float orc_gelu_lut(float x);
float orc_exp_lut(float x);

const char* orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
