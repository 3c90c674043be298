// wb_ctx.hpp -- the context behind the C-ABI handle (the reference's WhisperContext,
// src/main.rs:333-363, re-laid-out for one B200: weights, activations and KV caches live in HBM,
// sized from hparams instead of the MEM_REQ_* tables).
#pragma once

#include <string>
#include <unordered_map>
#include <vector>

#include "wb_kernels.hpp"
#include "wb_loader.hpp"

namespace wb {

struct Linear {            // one weight matrix [N][K] f16 (row-major, K contiguous) on the device
  __half* w = nullptr;
  int N = 0, K = 0;
  int bn = 0;
  CUtensorMap map_w;       // as the GEMM's W operand: dims {K, N}, box {64, bn}
  CUtensorMap map_a;       // as the A operand (swap-AB decode GEMMs): dims {K, N, 1}, box {64, 128, 1}
  bool has_map_a = false;
  int bn2 = 0;             // pair-kernel tile width (gemm2.cu); 0 = not eligible
  CUtensorMap map_w2;      // dims {K, N}, box {64, bn2 / 2}: each CTA of a pair loads half of the W tile
  const float* bias = nullptr;
  const float* colscale = nullptr;
  // LayerNorm-folded weights (upload_cat_ln): w = s * W diag(gamma), bias = c2 = s * (W beta + b),
  // ln_c1[n] = sum_k w[n][k]:  s * (LN(x) W^T + b) = rstd * (x w^T - mu * ln_c1) + bias
  const float* ln_c1 = nullptr;
};

struct EncLayer {
  const float *attn_ln_w, *attn_ln_b, *mlp_ln_w, *mlp_ln_b;
  Linear qkv, out, fc1, fc2;
};
struct DecLayer {
  const float *attn_ln_w, *attn_ln_b, *cross_ln_w, *cross_ln_b, *mlp_ln_w, *mlp_ln_b;
  Linear qkv, out, cq, cout, fc1, fc2;
};

struct ActMaps {          // one activation buffer as the skinny GEMM operand, per tile width
  CUtensorMap m[4];       // box {64, 32|64|128|256}
};

struct EncodeMaps {       // wb_encode's tensor maps for one (n_seg, audio context); rebuilt when either changes
  int n_seg = 0, T = 0;
  CUtensorMap m_conv1, m_conv2, m_ln, m_att, m_hid, m_enc;               // A operands
  CUtensorMap o_conv1, o_x3, o_pe, o_x, o_qk, o_hid, o_cross;            // epilogue output / residual boxes
  AttnProblem ap;
};

struct EncodeGraph {      // the launches of one wb_encode, captured for one shape
  cudaGraphExec_t exec = nullptr;
  int n_seg = 0, T = 0, mel_n_len = 0, norm_mode = -1;
  int launches = 0;       // kernels one replay runs (n_kernel_launches accounting)
};

struct PendingEvent {   // one bracketed launch whose events have not been read yet
  const char* fam;
  cudaEvent_t a, b;
};

struct KernelClock {   // device time per kernel family (CUDA events on the handle's stream)
  double total_us = 0;
  int64_t launches = 0;
};

}  // namespace wb

constexpr int WB_N_TICKETS = 8;
constexpr int WB_MAX_DEC_GROUPS = 4;
constexpr size_t WB_MAX_ENC_GRAPHS = 4;

struct wb_ctx {
  wb_config cfg{};
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  wb::ModelHParams hp{};
  int32_t special[8]{};
  std::vector<std::string> vocab;   // id_to_token (src/main.rs:544)
  std::vector<void*> allocs;
  bool canary = false;                // guard zones around every device buffer (wb_dbg_canary_check)
  struct Guarded { uint8_t* user; size_t bytes; };
  std::vector<Guarded> guarded;

  // ---- weights
  const float* e_pe = nullptr;
  wb::Linear conv1, conv2;            // rearranged [Cout][k][Cin]
  const float *ln_post_w = nullptr, *ln_post_b = nullptr;
  std::vector<wb::EncLayer> enc;
  wb::Linear cross_kv;                // all text layers: N = Lt * 2 * d
  const __half* d_te = nullptr;       // [n_vocab][d]
  wb::Linear logits_lin;              // d_te as a Linear (swap-AB A operand)
  const float *d_pe = nullptr, *d_ln_w = nullptr, *d_ln_b = nullptr;
  std::vector<wb::DecLayer> dec;

  // ---- mel
  wb::MelTables mel_tab{};
  void* d_pcm = nullptr;              // staging for host PCM (the buffer the next mel reads)
  size_t d_pcm_bytes = 0;
  // double-buffered upload (wb_pcm_prefetch): the next batch's H2D runs on copy_stream under the
  // current batch's encoder
  void* d_pcm_buf[2] = {nullptr, nullptr};
  int pcm_cur = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copy_done[2] = {nullptr, nullptr};   // recorded on copy_stream after the upload into buffer i
  cudaEvent_t ev_mel_read[2] = {nullptr, nullptr};    // recorded on stream after the mel kernel that read buffer i
  const void* pf_host[2] = {nullptr, nullptr};        // host pointer / size a prefetched buffer was filled from
  size_t pf_bytes[2] = {0, 0};
  float* d_mel = nullptr;             // [clip][n_mel][n_len]
  size_t d_mel_floats = 0;
  int* d_clip_max = nullptr;
  int mel_n_len = 0, mel_n_clips = 0;
  bool mel_normalized = false;        // false between wb_pcm_to_logmel and wb_mel_normalize
  bool mel_materialized = false;      // d_mel holds normalised values (after a read-back / wb_mel_write / checkpoints);
                                      // otherwise log10 values, normalised on the way into the encoder's window copy
  int* d_seg_max = nullptr;           // [max_segments] per-window maxima (WB_NORM_SEGMENT)

  // ---- encoder activations (capacity = cfg.max_segments)
  int* d_clip_ids = nullptr;
  long long* d_offsets = nullptr;
  int* h_clip_ids = nullptr;          // pinned ring [WB_N_TICKETS][max_segments] the segment table is uploaded from
  long long* h_offsets = nullptr;
  cudaEvent_t ev_seg_slot[WB_N_TICKETS] = {};
  int seg_slot_next = 0;
  __half* conv_in = nullptr;          // [seg][Tm+2][n_mel]
  __half* h1 = nullptr;               // [seg][Tm+2][d]
  float* x = nullptr;                 // residual stream [seg*T][d] f32
  __half* ln_out = nullptr;           // [seg*T][d]: LayerNorm output, or (ln_fold) the F16 copy of x
  bool ln_fold = false;               // attn_ln / mlp_ln folded into the GEMMs around them (gemm2.cu LN template)
  float2* ln_part[2] = {nullptr, nullptr};   // [max_segments*T][ln_parts] partial (sum, sum of squares) of x - centre, one
                                      // slot per (N tile, epilogue half) of the producing GEMM; two buffers in alternation
  float* ln_center = nullptr;         // [max_segments*T] per-row centre of the folded LayerNorms (gemm2.cu)
  int ln_parts = 0;                   // 2 * (d / pair tile width)
  const float *enc_ones = nullptr, *enc_zeros = nullptr;   // identity affine of layer 0's attn_ln (its gamma / beta live in the QKV weights)
  __half* qk = nullptr;               // [seg*T][2d]
  __half* vt = nullptr;               // [seg][H*64][Tp]
  __half* attn_out = nullptr;         // [seg*T][d]
  __half* hidden = nullptr;           // [seg*T][4d]
  float* enc_out = nullptr;           // [seg*T][d] f32 (ln_post)
  __half* enc_f16 = nullptr;
  __half* cross = nullptr;            // memory_cross_k/v: [2*Lt slabs][max_segments*T][d]; K of text layer il is slab
                                      // 2*il, V is slab 2*il+1 (each a dense [seg][T][d] matrix, as in src/main.rs:2018-2030)
  size_t cross_slab = 0;              // elements per slab = max_segments * T * d
  int Tp = 0;
  wb::EncodeMaps enc_maps;
  std::vector<wb::EncodeGraph> enc_graphs;   // captured encodes, one per shape seen twice (at most WB_MAX_ENC_GRAPHS)
  int enc_n_seg = 0;                  // segments of the last wb_encode
  int exp_n_audio_ctx = 0;            // exp_n_audio_ctx (src/main.rs:362): > 0 shortens the encoder's audio context
  int enc_T = 0;                      // audio context the last wb_encode ran with (rows per segment of its outputs)
  int pad_T = 0;                      // audio context the zero pad rows of conv_in / h1 are laid out for (0 = as allocated)
  double* d_chk = nullptr;            // [slot][max_segments]
  void* d_chk_scratch = nullptr;      // per-block partial sums + arrival counters of the sum|x| kernels (misc.cu)
  int n_chk_slots = 0;
  std::vector<char> chk_valid;

  // ---- decoder state
  int dec_rows_cap = 0;                          // rows (sequences x tokens) one wb_decode call may carry
  __half *self_k = nullptr, *self_v = nullptr;   // [layer][seq][n_text_ctx][d] F16  (memory_k/v, 1343-1347)
  float* dx = nullptr;                           // residual stream [rows][d] f32
  __half *d_ln = nullptr, *d_qkv = nullptr, *d_att = nullptr, *d_hid = nullptr, *d_q = nullptr, *d_lnf = nullptr;
  wb::ActMaps m_ln, m_att, m_hid, m_lnf;         // swap-AB activation operands
  float* d_logits = nullptr;                     // [seq][n_vocab] of the last position  (logits, 351)
  float* d_top2 = nullptr;            // [32][ceil(n_vocab/16)][3] per-CTA top-2 of the vocabulary projection
  bool logits_top2_valid = false;
  int *d_tokens = nullptr, *d_next = nullptr, *d_out_tokens = nullptr, *d_done = nullptr, *d_out_len = nullptr;
  int *d_npast = nullptr, *d_step = nullptr;
  float *d_margin = nullptr, *d_out_margin = nullptr;
  float *d_part_o = nullptr, *d_part_ml = nullptr;
  int* d_split_cnt = nullptr;                    // per (row, head) arrival counter of the split cross-attention
  int dec_n_seq = 0;
  const float *d_ones = nullptr, *d_zeros = nullptr;   // identity affine for the prompt pass's plain LayerNorm
  float* dec_ln_center = nullptr;                // [DEC_LN_ROWS] per-row centre of the folded decode step
  wb::DecLnStat* dec_ln_stats = nullptr;         // [3 Lt + 1][DEC_LN_SUB][DEC_LN_ROWS] row statistics of the folded single-token step
  cudaGraphExec_t step_graph = nullptr;          // one single-token greedy step, captured per n_seq
  int step_graph_n_seq = 0, step_graph_max_new = 0, step_graph_eot = -1;
  int step_graph_enc_T = 0, step_graph_n_split = 0;   // host parameters decode_pass bakes into the captured launches
  int step_graph_groups = 0;                     // sequence groups (parallel branches) of the captured step
  cudaStream_t dec_group_stream[WB_MAX_DEC_GROUPS] = {};   // capture-only streams of the branches
  cudaEvent_t dec_group_done[WB_MAX_DEC_GROUPS] = {};
  cudaEvent_t dec_fork = nullptr;
  // side branch of the step graph that prefetches the next layer's cross K / V into L2 (decode_pass)
  cudaStream_t dec_pf_stream = nullptr;
  cudaEvent_t dec_pf_fork = nullptr, dec_pf_join = nullptr;
  int dec_pf_mb = 0;           // MB per layer (0: off)
  bool dec_pf_active = false;  // set while the single-token step is being captured

  // ---- results read back without blocking (wb_encoder_digest_async / wb_wait)
  cudaEvent_t ev_ticket[WB_N_TICKETS] = {};
  int ticket_next = 0;

  // ---- timing
  cudaEvent_t ev[3][2] = {};          // [mel|encode|decode][start|stop] of the most recent call
  bool ev_used[3] = {false, false, false};
  wb_timings tm{};
  std::unordered_map<std::string, wb::KernelClock> clocks;
  std::vector<wb::PendingEvent> pending_events;
  std::vector<cudaEvent_t> free_events;
  bool time_kernels = false;
};
