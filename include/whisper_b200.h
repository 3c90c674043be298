/*
 * whisper_b200.h -- C-ABI of libwhisper_b200.so, the B200 (sm_100a) drop-in for the Whisper hot
 * path of szuwgh/whisper.rs.
 *
 * The reference exposes no FFI/plugin interface: its boundary is three crate-private Rust
 * functions sharing one context struct.  Each entry point below names the
 * reference item it replaces (file:line in /root/reference, all in src/main.rs).  A Rust `-sys`
 * crate binds exactly these symbols (whisper.rs_b200/rust/).
 *
 * Conventions: opaque handle; caller-owned host buffers (plain pointers + sizes); `int` return,
 * 0 = ok, negative = a WsError variant (src/main.rs:50-72); one handle <-> one CUDA device + one
 * stream; a handle is thread-compatible, not thread-safe (the reference's `&mut WhisperContext`).
 * There is NO CPU fallback: every compute call fails with WB_ERR_TENSOR_OP when no sm_100 device
 * is usable.
 */
#ifndef WHISPER_B200_H
#define WHISPER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wb_ctx wb_ctx;

/* WsError (src/main.rs:50-72) */
enum {
  WB_OK = 0,
  WB_ERR_UNEXPECTED = -1,         /* WsError::Unexpected */
  WB_ERR_IO = -2,                 /* WsError::UnexpectIO */
  WB_ERR_BAD_MAGIC = -3,          /* WsError::BadMagic */
  WB_ERR_NOT_ENOUGH_SPACE = -4,   /* WsError::NotEnoughSpace (batch / context capacity exceeded) */
  WB_ERR_UNKNOWN_TENSOR = -5,     /* WsError::UnknownTensor */
  WB_ERR_BAD_REF_TENSOR = -6,     /* WsError::BadRefTensor */
  WB_ERR_WRONG_SIZE_TENSOR = -7,  /* WsError::WrongSizeTensor */
  WB_ERR_WRONG_SHAPE_TENSOR = -8, /* WsError::WrongShapeTensor */
  WB_ERR_WRONG_BYTES_TENSOR = -9, /* WsError::WrongBytesTensor */
  WB_ERR_TENSOR_OP = -10          /* WsError::WrongGTensor: a device/kernel failure */
};

enum { WB_NORM_CLIP = 0, WB_NORM_SEGMENT = 1 };

typedef struct wb_config {
  int32_t device;          /* CUDA ordinal */
  int32_t max_segments;    /* encoder/decoder batch capacity (30 s windows per wb_encode call) */
  int32_t max_clips;       /* clips per wb_pcm_to_mel call */
  int64_t max_clip_samples;/* longest clip, in samples */
  int32_t norm_scope;      /* WB_NORM_CLIP = whole-clip max as clamp_and_normalize (1654-1671), the reference's
                              behaviour (a clip split across GPUs keeps the whole-clip maximum through
                              wb_pcm_to_logmel / wb_mel_normalize).  WB_NORM_SEGMENT = every encoder window is
                              clamped against its OWN maximum (what the reference computes for a clip of exactly
                              that window): windows become fully independent; the stored mel (wb_mel_read) then
                              holds the un-normalised log10 values */
  int32_t checkpoints;     /* 1: keep sum|x| probes per stage (the author's debug probes, 1836-1849) */
  void*   stream;          /* cudaStream_t to run on; NULL = the library creates its own */
  int32_t decode_capacity; /* 1: allocate self-attention KV for max_segments sequences */
  int32_t reserved[7];
} wb_config;

typedef struct wb_timings {  /* t_*_us of WhisperContext (src/main.rs:334-339), device-timed */
  int64_t t_load_us, t_mel_us, t_sample_us, t_encode_us, t_decode_us;
  int64_t n_mel_calls, n_encode_calls, n_decode_calls, n_kernel_launches;
} wb_timings;

/* stages of wb_checksum: the sum|x| probes the reference's author compared against whisper.cpp
 * (src/main.rs:1439-1454, 1836-1849, 1998-2010, 2032-2058) */
enum {
  WB_STAGE_MEL = 0, WB_STAGE_CONV1 = 1, WB_STAGE_CONV2_POS = 2, WB_STAGE_LAYER = 3,
  WB_STAGE_LN_POST = 4, WB_STAGE_CROSS_K = 5, WB_STAGE_CROSS_V = 6
};

void wb_config_default(wb_config* cfg);

/* WhisperContext::new (366-503): parse the ggml-v1 file (magic 368-371, hparams 622-658, filters
 * 513-535, vocab 578-592, tensor records 1381-1481 with the checks of 1401-1434), upload the
 * weights, size every activation buffer from hparams (the MEM_REQ_* tables 117-189 are not used). */
int wb_ctx_create(const char* model_path, const wb_config* cfg, wb_ctx** out);
void wb_ctx_free(wb_ctx* ctx);
int wb_get_hparams(const wb_ctx* ctx, int32_t out[11]);           /* WhisperHparams 607-619 */
int wb_get_special_tokens(const wb_ctx* ctx, int32_t out[8]);     /* WhisperVocab ids 557-575, 433-440:
                                                                      eot,sot,prev,solm,not,beg,translate,transcribe */

/* whisper_pcm_to_mel (1681-1707) -> log_mel_spectrogram (1554-1652) + clamp_and_normalize
 * (1654-1671), for n_clips clips of n_samples each.  The result stays on the device, like
 * ctx.mel (349).  `pcm` is a HOST pointer ([n_clips][n_samples] f32); the _device variant takes a
 * device pointer (inputs already resident in HBM).  The calls do not wait for the GPU: with PINNED host
 * memory the upload is asynchronous, so the buffer must stay unchanged until a later blocking call on the
 * handle (wb_sync, wb_wait, any read-back) has returned -- the reference holds its samples in an
 * Arc<Vec<f32>> for the same reason (1584, 1681); pageable memory is staged before the call returns. */
int wb_pcm_to_mel(wb_ctx* ctx, const float* pcm, size_t n_samples, int n_clips);
int wb_pcm_to_mel_device(wb_ctx* ctx, const float* pcm_dev, size_t n_samples, int n_clips);
/* same with i16 PCM, the reference's real input: convert_integer_to_float_audio (1673-1679) */
int wb_pcm16_to_mel(wb_ctx* ctx, const int16_t* pcm, size_t n_samples, int n_clips);
/* Start uploading the NEXT batch's host PCM (n_bytes of f32 or i16 samples, ideally pinned memory)
 * on a separate copy stream while the current batch is still being encoded.  A following
 * wb_pcm_to_mel / wb_pcm16_to_mel with the same pointer and byte count uses the uploaded copy
 * instead of copying again.  The caller keeps the host buffer unchanged until that call.  (The
 * reference holds its samples in an Arc<Vec<f32>> shared with its mel threads, 1584, 1681.) */
int wb_pcm_prefetch(wb_ctx* ctx, const void* pcm, size_t n_bytes);
/* whisper_pcm_to_mel in two phases, for ONE long clip whose samples are split across GPUs (each
 * GPU holds the span of its own 30 s windows plus the 240-sample halo frame i = [160i, 160i+400)
 * reaches into, 1594-1597).  The whole-clip maximum of clamp_and_normalize (1655-1662) is the only
 * coupling between the parts:
 *   wb_pcm_to_logmel   log10 mel (1554-1652) of exactly n_frames frames per clip (0 = n_samples/160;
 *                      samples past n_samples read as zero, 1596-1600) + the local per-clip maximum
 *   wb_mel_max_read    that maximum -> the caller takes the MAX over all parts (one 4-byte all-reduce)
 *   wb_mel_normalize   x = max(x, max - 8); x = (x + 4) / 4 (1664-1670) with the combined maxima
 *                      (clip_max NULL = the local ones, which makes the two calls equal wb_pcm_to_mel)
 * wb_encode refuses a mel that has not been normalised. */
int wb_pcm_to_logmel(wb_ctx* ctx, const float* pcm, size_t n_samples, int n_clips, int n_frames);
int wb_mel_max_read(wb_ctx* ctx, float* out, int n_clips);
int wb_mel_normalize(wb_ctx* ctx, const float* clip_max, int n_clips);
int wb_mel_dims(const wb_ctx* ctx, int* n_mel, int* n_len, int* n_clips);
int wb_mel_read(wb_ctx* ctx, int clip, float* out, size_t cap_floats);   /* [n_mel][n_len], layout of 1633 */
int wb_mel_write(wb_ctx* ctx, const float* mel, int n_mel, int n_len, int n_clips); /* set ctx.mel directly */

/* whisper_encode (1799-2063), batched: segment s encodes the 2*n_ctx-frame window starting at
 * mel_offsets[s] of clip clip_ids[s] (NULL = clip 0 / offset 0).  n_segments = 1, offset 0
 * reproduces the reference's call (2074).  Leaves ln_post output and the per-layer cross-attention
 * K/V (memory_cross_k/v, 1990-2030) on the device. */
int wb_encode(wb_ctx* ctx, const int32_t* clip_ids, const size_t* mel_offsets, int n_segments);
/* exp_n_audio_ctx (362, read at 1803-1807): n_ctx > 0 makes the following wb_encode calls run with an audio context
 * of n_ctx <= n_audio_ctx positions (a window of 2 * n_ctx mel frames, the first n_ctx rows of the positional
 * embedding); 0 = the model's.  Read-backs, digests and the decoder's cross-attention follow the last encode. */
int wb_set_audio_ctx(wb_ctx* ctx, int n_ctx);
int wb_encoder_out_read(wb_ctx* ctx, int seg, float* out);        /* `cur` after ln_post (1980-1984): [n_ctx][d] f32 */
int wb_cross_kv_read(wb_ctx* ctx, int seg, int layer, uint16_t* k, uint16_t* v); /* F16 bits [n_ctx][d], 2018-2030 */
int wb_checksum(wb_ctx* ctx, int stage, int layer, int seg, double* abs_sum);
/* sum|x| of every encoded segment's ln_post output in one call (the author's probe, 1836-1849, applied
 * to `cur` after 1984): out[n_segments] doubles.  The small host-visible result of an encode step. */
int wb_encoder_digest(wb_ctx* ctx, double* out, int cap);
/* Non-blocking form: the digest and its read-back into `out` (pinned host memory) are queued on the
 * handle's stream; returns a ticket (>= 0) for wb_wait.  Lets a caller submit batch i+1 (mel + encode)
 * before it waits for batch i's result, so one handle keeps its stream full.  Up to 8 tickets may be
 * outstanding. */
int wb_encoder_digest_async(wb_ctx* ctx, double* out, int cap);
int wb_wait(wb_ctx* ctx, int ticket);

/* The decode step the reference declares state for but never implements (694-731, 1336-1354,
 * logits/probs 351-352): tokens is HOST [n_seqs][n_tokens]; sequence i attends to the cross K/V
 * of encoder segment i.  Logits of the last position stay on the device. */
int wb_decode(wb_ctx* ctx, const int32_t* tokens, int n_tokens, int n_past, int n_seqs);
int wb_logits_read(wb_ctx* ctx, int seq, float* out);             /* [n_vocab] f32 */
/* greedy loop on the device: arg-max over all logits, stop per sequence at `eot` or max_new.
 * out_tokens [n_seqs][max_new], out_margin (top1 - top2 logit, may be NULL), out_len [n_seqs].
 * At most n_text_ctx tokens are generated per sequence whatever max_new says; the row stride of the
 * output arrays stays the caller's max_new (columns past the generated ones are left untouched). */
int wb_decode_greedy(wb_ctx* ctx, const int32_t* prompt, int n_prompt, int max_new, int eot,
                     int n_seqs, int32_t* out_tokens, float* out_margin, int32_t* out_len);

/* WhisperVocab::id_to_token (544) incl. the placeholder names of ids the file has no text for
 * (442-467).  Returns the token's byte length (the copy is truncated to cap-1 and NUL-terminated),
 * negative on a bad id.  wb_tokens_to_text concatenates the text tokens (ids below eot) of a
 * greedy result and skips the special / timestamp ids. */
int wb_token_text(const wb_ctx* ctx, int32_t id, char* out, size_t cap);
int wb_tokens_to_text(const wb_ctx* ctx, const int32_t* ids, int n, char* out, size_t cap);

int wb_sync(wb_ctx* ctx);                                         /* wait for the handle's stream */
int wb_timings_get(const wb_ctx* ctx, wb_timings* out);           /* t_*_us 334-339 */
const char* wb_last_error(const wb_ctx* ctx);                     /* WsError Display text (52-71); ctx may be NULL */
const char* wb_version(void);

/* ---- parity / measurement probes (used by tests and bench.py; not part of the reference surface)
 * Single-op entry points over HOST buffers so each kernel can be checked in isolation. */
int wb_dbg_gemm(wb_ctx* ctx, int M, int N, int K, const uint16_t* a_f16, const uint16_t* w_f16,
                const float* bias, const float* residual, int gelu, float scale, int out_f16,
                void* out);
int wb_dbg_attention(wb_ctx* ctx, int n_seg, int T, int H, const uint16_t* qkv_f16 /* [n_seg*T][3*H*64] */,
                     uint16_t* out_f16 /* [n_seg*T][H*64] */);
int wb_dbg_layernorm(wb_ctx* ctx, int rows, int d, const float* x, const float* w, const float* b,
                     uint16_t* out_f16);
/* With wb_config.reserved[1] = 1 (or WB_CANARY=1 in the environment) every device buffer of the handle is allocated
 * between two 256-byte guard zones; this returns how many guard zones have been written into since (0 = no kernel
 * wrote outside its buffers; -1 = the handle has no guards). */
int wb_dbg_canary_check(wb_ctx* ctx);
/* device-side timing of the last call of each kind, in microseconds, per kernel family */
int wb_kernel_time_us(const wb_ctx* ctx, const char* family, double* total_us, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif
