// wb_kernels.hpp -- host-side launch interface of the hand-written sm_100a kernels.
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

namespace wb {

// Launch with programmatic stream serialisation (see pdl_wait() in ptx.cuh): the kernel's prologue
// overlaps the tail of the previous kernel in the stream.
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}

// ---- TMA descriptors (cuTensorMapEncodeTiled resolved through cudaGetDriverEntryPoint so the
// library has no link-time dependency on libcuda) ------------------------------------------------
// f16 tensor, up to 4 dims; dims[0] innermost (contiguous); strides_bytes[i] = stride of dim i+1.
bool make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, const char** err);
// same for an f32 tensor (epilogue output / residual boxes of gemm2.cu)
bool make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, const char** err);

// ---- GEMM: C[b][m][n] = epi( sum_k A[b][m][k] * W[n][k] ), f16 x f16 -> f32 (TMEM) -------------
// Restates galois_matmul (src/main.rs:1752-1767) / galois_conv_1d_* (1709-1721) call sites with
// their bias / scale / GELU / residual / F16-repack followers fused into the epilogue.
// Row statistics of the decoder's folded LayerNorms: sum(x) * 2^30 and sum(x^2) * 2^20 as signed 64-bit integers
// (two's complement in unsigned words for atomicAdd).  |sum x| < 8e9 and sum x^2 < 8e12 fit; the quantisation (1e-9 /
// 1e-6 absolute) is far below the f32 rounding of the terms themselves.
struct DecLnStat {
  unsigned long long s1, s2;
};
constexpr float DEC_LN_S1_SCALE = 1073741824.0f;   // 2^30
constexpr float DEC_LN_S2_SCALE = 1048576.0f;      // 2^20

struct GemmEpilogue {
  const float* bias = nullptr;      // [N] added to the accumulator
  const float* colscale = nullptr;  // [N] multiplies (acc + bias)
  float scale = 1.0f;               // scalar multiplier after bias (cross-K (d/H)^-1/4, 1994-1996)
  int gelu = 0;                     // GELU (1775-1779), F16-rounded input
  const float* residual = nullptr;  // f32 [b*res_bstride + m*res_ld + n], added last
  long long res_bstride = 0;
  int res_ld = 0;
  void* out = nullptr;              // [b*out_bstride + m*out_ld + n]
  int out_f16 = 1;
  long long out_bstride = 0;
  int out_ld = 0;
  // column slabs: with out_slab_cols = C > 0, column n goes to slab n / C at column n % C:
  //   out[(n / C) * out_slab_stride + m*out_ld + n % C]   (batch must be 1)
  // -- the cross-attention K/V of every text layer leave one GEMM as per-layer [rows][d] matrices
  int out_slab_cols = 0;
  long long out_slab_stride = 0;
  // columns n >= vt_col0 are written transposed (time contiguous) as F16: with nn = n - vt_col0,
  //   vt_out[((seg*vt_heads + nn/64) * vt_head_rows + nn%64) * vt_ld + t],  seg = m / vt_T, t = m % vt_T
  // which is the reference's V layout `[T, Dh, H]` per segment (1914-1920); each head block holds
  // vt_head_rows >= 64 rows (the attention kernel keeps a row of ones after the 64 head rows).
  __half* vt_out = nullptr;
  int vt_col0 = 1 << 30;
  int vt_heads = 0;
  int vt_head_rows = 64;
  int vt_ld = 0;
  int vt_T = 1;
  // LayerNorm folded into the GEMMs around it (pair kernel only, gemm2.cu): a producer (f32 residual-stream
  // output) also writes an F16 copy of x - ln_center[row] to x16_out[row*x16_ld + n] and leaves the (sum, sum of
  // squares) of x - center over its columns in ln_part_out[row*ln_parts + slot] (ln_parts = 2 * N tiles: plain
  // stores, no atomics); a consumer adds the ln_parts slots of its A rows in slot order (ln_part_in, row width = K),
  // applies rstd * (acc - mu' * colscale[n]) + bias[n] and moves ln_center[row] on by mu'.
  float2* ln_part_out = nullptr;
  const float2* ln_part_in = nullptr;
  int ln_parts = 0;
  float* ln_center = nullptr;
  // the same fold in the decoder's single-token step (decode_kernels.cu): per-row (sum, sum of squares) of x as
  // 64-bit FIXED-POINT numbers (DecLnStat), so that the atomic accumulation over CTAs is exact integer addition and the
  // result does not depend on the order the CTAs arrive in
  DecLnStat* ln_stats_out = nullptr;
  __half* x16_out = nullptr;
  int x16_ld = 0;
  const DecLnStat* ln_stats_in = nullptr;
  float ln_eps = 1e-5f;
  // swap-AB mode for skinny activations (decoder): the GEMM computes C^T; element (m, n) is
  // stored at out[n*out_ld + m] and bias/colscale/residual are indexed by m instead of n.
  int transpose_out = 0;
};

struct GemmProblem {
  CUtensorMap a_map;   // dims {K, M_rows, batch}, box {64, 128, 1}, SWIZZLE_128B
  CUtensorMap w_map;   // dims {K, N},             box {64, BN},     SWIZZLE_128B
  int M_rows = 0, batch = 1, N = 0, K = 0;
  int bn = 256;        // 32 | 64 | 128 | 192 | 256  (must match w_map's box)
  GemmEpilogue epi;
  // ---- CTA-pair kernel (gemm2.cu) only: w_map box {64, bn/2}; the epilogue stores through out_map
  // (dims {N, M_rows, batch}, box {64 f16 | 32 f32, 32, 1}, SWIZZLE_128B) and reads the f32 residual
  // through res_map (same box shape; epi.residual is not used); res_bcast: residual has no batch dim
  const CUtensorMap* out_map = nullptr;
  const CUtensorMap* res_map = nullptr;
  int res_bcast = 0;
  long long* dbg = nullptr;   // optional clock64() trace of pair 0 (tools/gemm_trace.py); nullptr in production
};
cudaError_t launch_gemm(const GemmProblem& g, int num_sms, cudaStream_t st);
int gemm_pick_bn(int N);
bool gemm_setup_attributes(const char** err);
cudaError_t launch_gemm2(const GemmProblem& g, int num_sms, cudaStream_t st);
int gemm2_pick_bn(int N);   // 0: N not eligible for the pair kernel

// ---- fused softmax attention (galois_flash_attn src/main.rs:1787-1797, call 1922) -------------
constexpr int ATTN_VT_HEAD_ROWS = 80;   // 64 head rows + 1 row of ones + 15 zero rows (MMA N = 80)
struct AttnProblem {
  CUtensorMap qk_map;  // dims {64, 2H, T, B} over the [B*T][2d] Q|K buffer, box {64,1,128,1}
  CUtensorMap vt_map;  // dims {Tp, B*H*80} over V^T, box {64, 80}
  int B = 0, T = 0, H = 0;
  __half* out = nullptr;  // [B*T][H*64] merged heads (1924-1929)
  float scale = 0.125f;
  long long* dbg = nullptr;  // optional clock64() trace buffer (1024 entries) for tools/prof_attention.py
};
cudaError_t launch_attention(const AttnProblem& a, cudaStream_t st);
cudaError_t launch_vt_init(__half* vt, int n_heads_total, int Tp, cudaStream_t st);   // ones / zero rows
bool attention_setup_attributes(const char** err);

// ---- log-mel (src/main.rs:1554-1671) -----------------------------------------------------------
// every table the mel kernel needs, laid out exactly as it sits in shared memory: each CTA copies the blob
// verbatim with 16-byte loads that are all in flight together
struct MelTableBlob {
  float2 w200[200];           // W_200^j, j < 200
  float2 w400[202];           // W_400^k, k <= 200 (+ 1 of padding: every member starts on 16 bytes)
  float hann[400];            // periodic Hann window (1567-1569)
  float fw[1024];             // non-zero filterbank taps, mel after mel (bin order inside a mel)
  int2 frange[128];           // per mel: first bin, first tap index in fw
  int fcount[128];            // per mel: number of taps
};
static_assert(sizeof(MelTableBlob) % 16 == 0, "copied as float4");
struct MelTables {            // built once per context
  const MelTableBlob* blob;   // device
  int n_mel;
};
// frames -> log10 mel power, [clip][n_mel][n_len]; also per-clip running max (ordered-int encoding)
cudaError_t launch_mel_frames(const MelTables& t, const void* pcm, int pcm_is_i16, size_t n_samples,
                              int n_clips, int n_len, float* mel_out, int* clip_max_enc, cudaStream_t st);
// clamp_and_normalize (1654-1671): x = max(x, max - 8); x = (x + 4) / 4
cudaError_t launch_mel_normalize(float* mel, int n_clips, size_t per_clip, const int* clip_max_enc, cudaStream_t st);
cudaError_t launch_fill_i32(int* p, int n, int v, cudaStream_t st);

// ---- small fused kernels -------------------------------------------------------------------------
// E0 (1816-1829) + F16 rounding of the conv operand: [clip][n_mel][n_len] f32 window ->
// [seg][Tm + 2][n_mel] f16 token-major with one zero row before and after.
// norm_mode 0: the mel is already normalised; 1 / 2: it holds log10 values and clamp_and_normalize (1654-1671) is
// applied on the way through with the maximum max_enc[clip] / max_enc[seg] (ordered-int encoding, mel.cu)
cudaError_t launch_mel_window(const float* mel, int n_mel, int n_len, const int* clip_ids,
                              const long long* offsets, int n_seg, int Tm, __half* out, cudaStream_t st,
                              const int* max_enc = nullptr, int norm_mode = 0);
// per-segment maximum of the window [offset, offset + Tm) of its clip's log10 mel (WB_NORM_SEGMENT)
cudaError_t launch_mel_window_max(const float* mel, int n_mel, int n_len, const int* clip_ids, const long long* offsets,
                                  int n_seg, int Tm, int* seg_max_enc, cudaStream_t st);
// galois_norm + repeat/mul/add (1781-1785, 1882-1886): rows of d, f32 in; f16 and/or f32 out
cudaError_t launch_layernorm(const float* x, const float* w, const float* b, int rows, int d,
                             __half* out_f16, float* out_f32, cudaStream_t st, long long in_row_stride = 0,
                             bool pdl = false, float* center_out = nullptr);   // center_out[row] = the row's mean
// sum|x| probes (1836-1849): out[seg] = sum over that segment's elements, bit-reproducible (per-block partial
// slots added in slot order by the last block).  `scratch`: abs_sum_scratch_bytes(scratch_segs) zero-initialised bytes.
size_t abs_sum_scratch_bytes(int max_seg);
cudaError_t launch_abs_sum_f32(const float* x, long long per_seg, long long seg_stride, int n_seg, double* out,
                               void* scratch, int scratch_segs, cudaStream_t st);
cudaError_t launch_abs_sum_f16(const __half* x, int rows, int cols, long long row_stride, long long seg_stride,
                               int n_seg, double* out, void* scratch, int scratch_segs, cudaStream_t st);

// ---- decoder step kernels (absent in the reference) ------------------------
// n_past / step live in device memory so one captured CUDA graph serves every position.
// D1: x[s][i][:] = d_te[tok] + d_pe[n_past + i]
cudaError_t launch_embed(const __half* te, const float* pe, const int* tokens, int n_seq, int n_tok,
                         const int* n_past_dev, int d, float* x, cudaStream_t st, DecLnStat* stats = nullptr,
                         __half* x16 = nullptr, int n_clear_slots = 0, float* center = nullptr);
// D2: append this step's K/V to the F16 cache [seq][n_text_ctx][d], causal attention over it
cudaError_t launch_decode_self_attn(const __half* qkv, int d, __half* kc, __half* vc, int n_seq, int n_tok,
                                    const int* n_past_dev, int n_text_ctx, int H, __half* out, cudaStream_t st);
// D3: cross-attention over the encoder memory; rows at kv[(seq*T + t)*ld_kv + h*64 ..]
int decode_cross_splits(int n_seq, int H, int T, int num_sms);
cudaError_t launch_decode_cross_attn(const __half* q, int d, const __half* k, const __half* v, long long ld_kv,
                                     long long head_stride, int n_seq, int n_tok, int T, int H, __half* out,
                                     float* part_o, float* part_ml, int n_split, int* split_cnt, cudaStream_t st);
// D6 / K13: per-sequence arg-max + top-2 margin + greedy-loop bookkeeping
cudaError_t launch_argmax(const float* logits, int n_seq, int n_vocab, int* next_tok, float* margin, int* out_tokens,
                          float* out_margin, int* out_len, int* done, int max_new, int* step_dev, int eot,
                          cudaStream_t st, int* n_past_dev = nullptr, int advance_by = 0);   // n_past_dev: also advance n_past / step
// skinny linear layer of a decode step (R <= 32 activation rows): out[r][n] = epi(sum_k x[r][k] W[n][k])
struct DecodeLinear {
  const __half* w = nullptr;       // [N][K] f16, K contiguous
  const __half* x = nullptr;       // [32][ldx] f16 (rows >= R may hold anything finite; they are never stored)
  int N = 0, K = 0, R = 0, ldx = 0;
  const float* bias = nullptr;     // [N]
  const float* colscale = nullptr; // [N]
  float scale = 1.0f;
  int gelu = 0;
  const float* residual = nullptr; // f32 [R][res_ld], added last
  int res_ld = 0;
  void* out = nullptr;             // [R][out_ld]
  int out_f16 = 1, out_ld = 0;
  float* top2 = nullptr;           // optional [R][n_parts][3]: per-CTA (top value, second value, index bits)
  // LayerNorm folded into the single-token step (no LayerNorm kernels between the linears):
  const DecLnStat* ln_in = nullptr;   // consumer: per-row (sum, sum of squares) of x (fixed point); w carries gamma, bias = c2
  const float* ln_c1 = nullptr;    //           c1[n] = sum_k w[n][k]
  float ln_inv_d = 0.0f, ln_eps = 1e-5f;
  DecLnStat* ln_out = nullptr;     // producer: statistics of (result - centre) per row, accumulated with integer atomics
  float* ln_center = nullptr;      // [R] per-row centre: read by producers, advanced by the row mean by consumers (CTA 0)
  __half* x16_out = nullptr;       //           and their F16 copy [R][x16_ld] (the next linear's activations)
  int x16_ld = 0;
};
constexpr int DEC_LN_ROWS = 32;    // rows per statistics slot of the folded decode step
// A slot is DEC_LN_SUB sub-slots of DEC_LN_ROWS rows: a producer CTA adds into sub-slot (its index mod DEC_LN_SUB), the
// consumer adds the sub-slots up -- same-address atomics of the 48-320 CTAs of a producer serialise in L2 (measured:
// one 64-bit word per row cost 0.36 us per producer kernel), spreading them over 4 words removes most of that.
constexpr int DEC_LN_SUB = 4;
constexpr int DEC_LN_SLOT = DEC_LN_SUB * DEC_LN_ROWS;
cudaError_t launch_decode_linear(const DecodeLinear& a, cudaStream_t st);
int decode_linear_parts(int N);    // CTAs (= top-2 partials per sequence) for N output features
cudaError_t launch_argmax_partials(const float* part, int n_part, int n_seq, int* next_tok, float* margin,
                                   int* out_tokens, float* out_margin, int* out_len, int* done, int max_new,
                                   int* step_dev, int eot, cudaStream_t st, int* n_past_dev = nullptr, int advance_by = 0);
cudaError_t launch_advance(int* n_past_dev, int add, int* step_dev, cudaStream_t st);
cudaError_t launch_l2_prefetch(const void* p0, const void* p1, size_t bytes_each, cudaStream_t st);   // two ranges, 4 KB granules

}  // namespace wb
