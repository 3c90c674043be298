// wb_api.cu -- C-ABI entry points: context creation (loader + weight upload), log-mel, encoder.
// Mirrors the reference's call order main -> WhisperContext::new -> whisper_pcm_to_mel ->
// whisper_encode (src/main.rs:2065-2075) behind include/whisper_b200.h.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <thread>
#if defined(__F16C__)
#include <immintrin.h>
#endif

#include "wb_internal.hpp"

namespace wb {

static thread_local std::string g_err;
void set_global_error(const std::string& msg) { g_err = msg; }

int fail(wb_ctx* ctx, int code, const char* what, cudaError_t e) {
  std::string m = std::string("galois tensor:'") + what + ": " + cudaGetErrorString(e) + "'";
  if (ctx) ctx->err = m;
  g_err = m;
  return code;
}
int fail_msg(wb_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  g_err = msg;
  return code;
}

// ------------------------------------------------------------------------------------------------
// TMA descriptors
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static bool make_tmap(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, const char** err) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    *err = "cuTensorMapEncodeTiled not available (no CUDA driver?)";
    return false;
  }
  cuuint64_t gd[5];
  cuuint64_t gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gs[i] = strides_bytes[i];
  }
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    static thread_local char buf[160];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (CUresult %d, rank %d, dims %llu,%llu box %u,%u)", (int)r,
             rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
             rank > 1 ? box[1] : 0);
    *err = buf;
    return false;
  }
  return true;
}

bool make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const char** err) {
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, base, rank, dims, strides_bytes, box, err);
}
bool make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const char** err) {
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, err);
}

// epilogue output / residual tensor of the pair GEMM: [batch][rows][N] with row stride ld and
// batch stride bstride (elements); box = 128 bytes of columns x 32 rows
bool tmap_out(CUtensorMap* m, const void* base, bool f32, uint64_t N, uint64_t rows, uint64_t batch, uint64_t ld_elems,
              uint64_t bstride_elems, const char** err) {
  const uint64_t es = f32 ? 4 : 2;
  const uint64_t dims[3] = {N, rows, batch};
  const uint64_t st[2] = {ld_elems * es, (batch > 1 ? bstride_elems : rows * ld_elems) * es};
  const uint32_t box[3] = {f32 ? 32u : 64u, 32, 1};
  return f32 ? make_tmap_f32(m, base, 3, dims, st, box, err) : make_tmap_f16(m, base, 3, dims, st, box, err);
}

bool tmap_2d_rows(CUtensorMap* m, const void* base, uint64_t K, uint64_t rows, uint64_t ld_elems, uint32_t box_rows,
                  const char** err) {
  const uint64_t dims[2] = {K, rows};
  const uint64_t st[1] = {ld_elems * 2};
  const uint32_t box[2] = {64, box_rows};
  return make_tmap_f16(m, base, 2, dims, st, box, err);
}
bool tmap_3d_rows(CUtensorMap* m, const void* base, uint64_t K, uint64_t rows, uint64_t batch, uint64_t ld_elems,
                  uint64_t bstride_elems, const char** err) {
  const uint64_t dims[3] = {K, rows, batch};
  const uint64_t st[2] = {ld_elems * 2, bstride_elems * 2};
  const uint32_t box[3] = {64, 128, 1};
  return make_tmap_f16(m, base, 3, dims, st, box, err);
}

bool make_linear_maps(wb_ctx* ctx, Linear& l, bool want_a_map) {
  const char* err = "";
  l.bn = gemm_pick_bn(l.N);
  if (!tmap_2d_rows(&l.map_w, l.w, l.K, l.N, l.K, l.bn, &err)) {
    fail_msg(ctx, WB_ERR_TENSOR_OP, err);
    return false;
  }
  l.bn2 = want_a_map ? 0 : gemm2_pick_bn(l.N);   // encoder-side weights also get the pair kernel's half-tile map
  if (l.bn2 && !tmap_2d_rows(&l.map_w2, l.w, l.K, l.N, l.K, l.bn2 / 2, &err)) {
    fail_msg(ctx, WB_ERR_TENSOR_OP, err);
    return false;
  }
  if (want_a_map) {
    if (!tmap_3d_rows(&l.map_a, l.w, l.K, l.N, 1, l.K, (uint64_t)l.K * l.N, &err)) {
      fail_msg(ctx, WB_ERR_TENSOR_OP, err);
      return false;
    }
    l.has_map_a = true;
  }
  return true;
}

// ------------------------------------------------------------------------------------------------
// kernel-family timing
// (the event lists live in the handle: two handles driven from two threads share no mutable state)
static cudaEvent_t get_event(wb_ctx* c) {
  auto& pool = c->free_events;
  if (!pool.empty()) {
    cudaEvent_t e = pool.back();
    pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
LaunchTimer::LaunchTimer(wb_ctx* ctx, const char* family) : c(ctx), fam(family) {
  ++c->tm.n_kernel_launches;
  if (c->time_kernels) {
    a = get_event(c);
    b = get_event(c);
    cudaEventRecord(a, c->stream);
  }
}
LaunchTimer::~LaunchTimer() {
  if (a) {
    cudaEventRecord(b, c->stream);
    c->pending_events.push_back(PendingEvent{fam, a, b});
  }
}
void resolve_kernel_clocks(wb_ctx* ctx) {
  for (auto& pe : ctx->pending_events) {
    float ms = 0.0f;
    if (cudaEventSynchronize(pe.b) == cudaSuccess && cudaEventElapsedTime(&ms, pe.a, pe.b) == cudaSuccess) {
      KernelClock& k = ctx->clocks[pe.fam];
      k.total_us += (double)ms * 1000.0;
      k.launches += 1;
    }
    ctx->free_events.push_back(pe.a);
    ctx->free_events.push_back(pe.b);
  }
  ctx->pending_events.clear();
}

static cudaError_t run_attention(const AttnProblem& ap, cudaStream_t st) { return launch_attention(ap, st); }

static bool force_gemm1() {
  static const bool v = getenv("WB_GEMM1") != nullptr;   // developer aid: A/B the single-CTA kernel
  return v;
}

static long long* g_gemm_trace = nullptr;   // developer aid: set by wb_dbg_gemm when WB_GEMM_TRACE is given

int run_gemm(wb_ctx* ctx, const CUtensorMap& a_map, int M_rows, int batch, const Linear& l, GemmEpilogue epi,
             const char* family, const CUtensorMap* out_map, const CUtensorMap* res_map, int res_bcast) {
  GemmProblem g;
  g.dbg = g_gemm_trace;
  g.a_map = a_map;
  g.M_rows = M_rows;
  g.batch = batch;
  g.N = l.N;
  g.K = l.K;
  if (!epi.bias) epi.bias = l.bias;
  if (!epi.colscale) epi.colscale = epi.ln_part_in ? l.ln_c1 : l.colscale;   // LN == 2 reads c1 from the column-scale slot
  g.epi = epi;
  // pair kernel: needs the output (and residual) tensor maps; a residual needs whole tiles of f32 output
  const bool pair = out_map && l.bn2 && !epi.transpose_out && !force_gemm1() && (!epi.residual || res_map) &&
                    (!res_map || (!epi.out_f16 && l.N % l.bn2 == 0 && epi.vt_col0 >= l.N));
  if ((epi.ln_part_out || epi.ln_part_in) && !pair)
    return fail_msg(ctx, WB_ERR_TENSOR_OP, "galois tensor:'LayerNorm-folded GEMM needs the pair kernel'");
  LaunchTimer t(ctx, family);
  if (pair) {
    g.w_map = l.map_w2;
    g.bn = l.bn2;
    g.out_map = out_map;
    g.res_map = epi.residual ? res_map : nullptr;
    g.res_bcast = res_bcast;
    WB_CK(launch_gemm2(g, ctx->num_sms, ctx->stream));
  } else {
    g.w_map = l.map_w;
    g.bn = l.bn;
    WB_CK(launch_gemm(g, ctx->num_sms, ctx->stream));
  }
  return WB_OK;
}

// ------------------------------------------------------------------------------------------------
// weight upload
const HostTensor* find(const ModelFileView& mv, const std::string& n) {
  auto it = mv.tensors.find(n);
  return it == mv.tensors.end() ? nullptr : &it->second;
}

// host tensor -> f16 vector (weights stored F32 in the file are rounded to F16: the tensor cores
// take F16 operands; files with hparams.f16 == 1 pass through bit-exactly)
void to_f16_host(const HostTensor& t, std::vector<__half>& out) {
  const int64_t n = t.nelem();
  out.resize((size_t)n);
  if (t.f16) {
    memcpy(out.data(), t.data, (size_t)n * 2);
  } else {
    for (int64_t i = 0; i < n; ++i) {
      float f;
      memcpy(&f, t.data + 4 * i, 4);
      out[(size_t)i] = __float2half_rn(f);
    }
  }
}

int upload_f32(wb_ctx* ctx, const ModelFileView& mv, const std::string& name, const float** out) {
  const HostTensor* t = find(mv, name);
  if (!t || t->f16) return fail_msg(ctx, WB_ERR_BAD_REF_TENSOR, "invalid ref tensor '" + name + "'\n");
  float* d = nullptr;
  int rc = dev_alloc(ctx, &d, (size_t)t->nelem(), false);
  if (rc) return rc;
  WB_CK(cudaMemcpyAsync(d, t->data, t->bytes, cudaMemcpyHostToDevice, ctx->stream));
  *out = d;
  return WB_OK;
}

int upload_f16_vec(wb_ctx* ctx, const std::vector<__half>& h, __half** out) {
  __half* d = nullptr;
  int rc = dev_alloc(ctx, &d, h.size(), false);
  if (rc) return rc;
  WB_CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));   // h is a temporary: synchronous
  *out = d;
  return WB_OK;
}
int upload_f32_vec(wb_ctx* ctx, const std::vector<float>& h, const float** out) {
  float* d = nullptr;
  int rc = dev_alloc(ctx, &d, h.size(), false);
  if (rc) return rc;
  WB_CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  *out = d;
  return WB_OK;
}

int upload_linear(wb_ctx* ctx, const ModelFileView& mv, const std::string& wname, const std::string& bname,
                  Linear& l, bool want_a_map) {
  const HostTensor* t = find(mv, wname);
  if (!t) return fail_msg(ctx, WB_ERR_BAD_REF_TENSOR, "invalid ref tensor '" + wname + "'\n");
  std::vector<__half> h;
  to_f16_host(*t, h);
  l.K = (int)t->ne[0];
  l.N = (int)t->ne[1];
  int rc = upload_f16_vec(ctx, h, &l.w);
  if (rc) return rc;
  if (!bname.empty()) {
    rc = upload_f32(ctx, mv, bname, &l.bias);
    if (rc) return rc;
  }
  return make_linear_maps(ctx, l, want_a_map) ? WB_OK : WB_ERR_TENSOR_OP;
}

// conv weight file layout [Cout][Cin][3] (ne = [3, Cin, Cout], src/main.rs:961-965) ->
// [Cout][k][Cin] so that one GEMM row of the activation operand is the contiguous run of three
// consecutive token-major input rows.
int upload_conv(wb_ctx* ctx, const ModelFileView& mv, const std::string& wname, const std::string& bname,
                Linear& l) {
  const HostTensor* t = find(mv, wname);
  if (!t) return fail_msg(ctx, WB_ERR_BAD_REF_TENSOR, "invalid ref tensor '" + wname + "'\n");
  std::vector<__half> h, r;
  to_f16_host(*t, h);
  const int Cin = (int)t->ne[1], Cout = (int)t->ne[2];
  r.resize(h.size());
  for (int co = 0; co < Cout; ++co)
    for (int ci = 0; ci < Cin; ++ci)
      for (int k = 0; k < 3; ++k) r[((size_t)co * 3 + k) * Cin + ci] = h[((size_t)co * Cin + ci) * 3 + k];
  l.K = 3 * Cin;
  l.N = Cout;
  int rc = upload_f16_vec(ctx, r, &l.w);
  if (rc) return rc;
  rc = upload_f32(ctx, mv, bname, &l.bias);
  if (rc) return rc;
  return make_linear_maps(ctx, l, false) ? WB_OK : WB_ERR_TENSOR_OP;
}

// concatenate weight rows of several tensors into one [sum N][K] matrix (+ bias / column scale)
int upload_cat(wb_ctx* ctx, const ModelFileView& mv, const std::vector<CatPart>& parts, Linear& l, bool want_a_map) {
  std::vector<__half> h;
  std::vector<float> bias, cs;
  bool any_scale = false;
  int K = 0;
  for (const CatPart& p : parts) {
    const HostTensor* t = find(mv, p.w);
    if (!t) return fail_msg(ctx, WB_ERR_BAD_REF_TENSOR, "invalid ref tensor '" + p.w + "'\n");
    std::vector<__half> part;
    to_f16_host(*t, part);
    K = (int)t->ne[0];
    const int N = (int)t->ne[1];
    h.insert(h.end(), part.begin(), part.end());
    if (!p.b.empty()) {
      const HostTensor* bt = find(mv, p.b);
      if (!bt) return fail_msg(ctx, WB_ERR_BAD_REF_TENSOR, "invalid ref tensor '" + p.b + "'\n");
      const size_t o = bias.size();
      bias.resize(o + N);
      memcpy(bias.data() + o, bt->data, (size_t)N * 4);
    } else {
      bias.resize(bias.size() + N, 0.0f);
    }
    cs.resize(cs.size() + N, p.scale);
    if (p.scale != 1.0f) any_scale = true;
  }
  l.K = K;
  l.N = (int)(h.size() / (size_t)K);
  int rc = upload_f16_vec(ctx, h, &l.w);
  if (rc) return rc;
  rc = upload_f32_vec(ctx, bias, &l.bias);
  if (rc) return rc;
  if (any_scale) {
    rc = upload_f32_vec(ctx, cs, &l.colscale);
    if (rc) return rc;
  }
  return make_linear_maps(ctx, l, want_a_map) ? WB_OK : WB_ERR_TENSOR_OP;
}

// host F16 <-> f32 (round to nearest even): F16C when the host compiler has it, else cuda_fp16's software path
static inline float h2f_host(uint16_t h) {
#if defined(__F16C__)
  return _cvtsh_ss(h);
#else
  __half v;
  memcpy(&v, &h, 2);
  return __half2float(v);
#endif
}
static inline uint16_t f2h_host(float f) {
#if defined(__F16C__)
  return _cvtss_sh(f, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
#else
  const __half v = __float2half_rn(f);
  uint16_t h;
  memcpy(&h, &v, 2);
  return h;
#endif
}
template <class F>
static void parallel_rows(int n, F&& body) {
  unsigned nt = std::thread::hardware_concurrency();
  if (nt > 16) nt = 16;
  if (nt < 2 || n < 64) {
    body(0, n);
    return;
  }
  std::vector<std::thread> th;
  const int per = (n + (int)nt - 1) / (int)nt;
  for (unsigned t = 0; t < nt; ++t) {
    const int lo = (int)t * per, hi = lo + per < n ? lo + per : n;
    if (lo < hi) th.emplace_back([&body, lo, hi] { body(lo, hi); });
  }
  for (auto& x : th) x.join();
}

// A Linear whose input is LayerNorm(x) (gamma, beta), with the LayerNorm's affine part and the part's output
// scale s folded into it:  w = s * W diag(gamma) rounded to F16;  ln_c1[n] = sum_k w[n][k];
// bias = c2[n] = s * (sum_k W[n][k] beta[k] + b[n]).  Then
//   s * (LN(x) W^T + b) = rstd * (x w^T - mu * ln_c1) + c2      (statistics applied in the GEMM epilogue), or
//                       = xhat w^T + c2                          (xhat = (x - mu) * rstd from a plain LayerNorm kernel)
int upload_cat_ln(wb_ctx* ctx, const ModelFileView& mv, const std::vector<CatPart>& parts, const std::string& gamma,
                  const std::string& beta, Linear& l, bool want_a_map) {
  const HostTensor* gt = find(mv, gamma);
  const HostTensor* bt = find(mv, beta);
  if (!gt || !bt || gt->f16 || bt->f16) return fail_msg(ctx, WB_ERR_BAD_REF_TENSOR, "invalid ref tensor '" + gamma + "'\n");
  const float* g = reinterpret_cast<const float*>(gt->data);
  const float* be = reinterpret_cast<const float*>(bt->data);
  std::vector<__half> h;
  std::vector<float> c1, c2;
  int K = 0;
  for (const CatPart& p : parts) {
    const HostTensor* t = find(mv, p.w);
    if (!t) return fail_msg(ctx, WB_ERR_BAD_REF_TENSOR, "invalid ref tensor '" + p.w + "'\n");
    std::vector<__half> part;
    to_f16_host(*t, part);
    K = (int)t->ne[0];
    const int N = (int)t->ne[1];
    if ((int64_t)K != gt->nelem() || (int64_t)K != bt->nelem())
      return fail_msg(ctx, WB_ERR_WRONG_SHAPE_TENSOR, "tensor '" + gamma + "' has wrong shape in model file\n");
    const float* pb = nullptr;
    if (!p.b.empty()) {
      const HostTensor* pbt = find(mv, p.b);
      if (!pbt) return fail_msg(ctx, WB_ERR_BAD_REF_TENSOR, "invalid ref tensor '" + p.b + "'\n");
      pb = reinterpret_cast<const float*>(pbt->data);
    }
    const size_t c0 = c1.size();
    c1.resize(c0 + N);
    c2.resize(c0 + N);
    // rows are independent: host threads, F16C conversions (a 3 GB model folds ~8e8 weights at load)
    parallel_rows(N, [&](int n_lo, int n_hi) {
      for (int n = n_lo; n < n_hi; ++n) {
        double s1 = 0.0, s2 = 0.0;
        uint16_t* row = reinterpret_cast<uint16_t*>(part.data()) + (size_t)n * K;
        for (int k = 0; k < K; ++k) {
          const float w = h2f_host(row[k]);
          const uint16_t wf = f2h_host(p.scale * w * g[k]);
          row[k] = wf;
          s1 += (double)h2f_host(wf);
          s2 += (double)w * (double)be[k];
        }
        c1[c0 + n] = (float)s1;
        c2[c0 + n] = (float)((double)p.scale * (s2 + (pb ? (double)pb[n] : 0.0)));
      }
    });
    h.insert(h.end(), part.begin(), part.end());
  }
  l.K = K;
  l.N = (int)(h.size() / (size_t)K);
  int rc = upload_f16_vec(ctx, h, &l.w);
  if (rc) return rc;
  if ((rc = upload_f32_vec(ctx, c2, &l.bias))) return rc;
  if ((rc = upload_f32_vec(ctx, c1, &l.ln_c1))) return rc;
  return make_linear_maps(ctx, l, want_a_map) ? WB_OK : WB_ERR_TENSOR_OP;
}

int build_mel_tables(wb_ctx* ctx, const ModelFileView& mv) {
  const int n_mel = mv.filt_n_mel;
  if (mv.filt_n_fft != 201) return fail_msg(ctx, WB_ERR_UNEXPECTED, "Unexpected: mel filterbank must have 201 bins");
  std::vector<float> hann(400);
  std::vector<float2> w200(200), w400(201);
  const double PI = 3.14159265358979323846;
  for (int i = 0; i < 400; ++i) hann[i] = (float)(0.5 * (1.0 - cos(2.0 * PI * i / 400.0)));   // periodic (1567-1569)
  for (int i = 0; i < 200; ++i) w200[i] = make_float2((float)cos(2.0 * PI * i / 200.0), (float)-sin(2.0 * PI * i / 200.0));
  for (int i = 0; i < 201; ++i) w400[i] = make_float2((float)cos(2.0 * PI * i / 400.0), (float)-sin(2.0 * PI * i / 400.0));
  std::vector<float> filt((size_t)n_mel * 201);
  memcpy(filt.data(), mv.filters, filt.size() * 4);
  std::vector<int2> range(n_mel);
  for (int j = 0; j < n_mel; ++j) {   // span of non-zero taps; exact zeros add nothing to the f32 sum
    int lo = 201, hi = 0;
    for (int k = 0; k < 201; ++k)
      if (filt[(size_t)j * 201 + k] != 0.0f) {
        lo = k < lo ? k : lo;
        hi = k + 1;
      }
    if (lo >= hi) lo = hi = 0;
    range[j] = make_int2(lo, hi);
  }
  // compact taps: [lo, hi) of each mel back to back (zeros inside the span are kept: the sum order stays bin order)
  std::vector<float> nz;
  std::vector<int> start(n_mel);
  for (int j = 0; j < n_mel; ++j) {
    start[j] = (int)nz.size();
    for (int k = range[j].x; k < range[j].y; ++k) nz.push_back(filt[(size_t)j * 201 + k]);
  }
  if (nz.size() > 1024 || n_mel > 128)
    return fail_msg(ctx, WB_ERR_UNEXPECTED, "Unexpected: mel filterbank too dense for the front-end kernel");
  std::vector<MelTableBlob> blob(1);
  MelTableBlob& b = blob[0];
  memset(&b, 0, sizeof(b));
  memcpy(b.w200, w200.data(), sizeof(float2) * 200);
  memcpy(b.w400, w400.data(), sizeof(float2) * 201);
  memcpy(b.hann, hann.data(), sizeof(float) * 400);
  memcpy(b.fw, nz.data(), sizeof(float) * nz.size());
  for (int j = 0; j < n_mel; ++j) {
    b.frange[j] = make_int2(range[j].x, start[j]);
    b.fcount[j] = range[j].y - range[j].x;
  }
  MelTableBlob* d_blob = nullptr;
  int rc;
  if ((rc = dev_alloc(ctx, &d_blob, 1, false))) return rc;
  WB_CK(cudaMemcpy(d_blob, &b, sizeof(b), cudaMemcpyHostToDevice));
  ctx->mel_tab = MelTables{d_blob, n_mel};
  return WB_OK;
}

int alloc_activations(wb_ctx* ctx) {
  const ModelHParams& hp = ctx->hp;
  const size_t S = (size_t)ctx->cfg.max_segments;
  const size_t T = hp.n_audio_ctx, Tm = 2 * T, d = hp.n_audio_state, Lt = hp.n_text_layer;
  ctx->Tp = (int)((T + 127) / 128 * 128);
  int rc;
  if ((rc = dev_alloc(ctx, &ctx->d_clip_ids, S))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_offsets, S))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->conv_in, S * (Tm + 2) * hp.n_mels))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->h1, S * (Tm + 2) * d))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->x, S * T * d))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->ln_out, S * T * d))) return rc;
  if (ctx->ln_fold) {
    ctx->ln_parts = 2 * ((int)d / gemm2_pick_bn((int)d));
    for (int i = 0; i < 2; ++i)
      if ((rc = dev_alloc(ctx, &ctx->ln_part[i], S * T * (size_t)ctx->ln_parts))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->ln_center, S * T))) return rc;
    std::vector<float> ones(d, 1.0f), zeros(d, 0.0f);
    if ((rc = upload_f32_vec(ctx, ones, &ctx->enc_ones))) return rc;
    if ((rc = upload_f32_vec(ctx, zeros, &ctx->enc_zeros))) return rc;
  }
  if ((rc = dev_alloc(ctx, &ctx->qk, S * T * 2 * d))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->vt, S * (d / 64) * ATTN_VT_HEAD_ROWS * ctx->Tp))) return rc;
  WB_CK(launch_vt_init(ctx->vt, (int)(S * (d / 64)), ctx->Tp, ctx->stream));
  if ((rc = dev_alloc(ctx, &ctx->attn_out, S * T * d))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->hidden, S * T * 4 * d))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->enc_out, S * T * d))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->enc_f16, S * T * d))) return rc;
  ctx->cross_slab = S * T * d;
  if ((rc = dev_alloc(ctx, &ctx->cross, S * T * Lt * 2 * d))) return rc;
  ctx->n_chk_slots = 4 + hp.n_audio_layer + 2 * hp.n_text_layer;
  if ((rc = dev_alloc(ctx, &ctx->d_chk, (size_t)ctx->n_chk_slots * S))) return rc;
  ctx->chk_valid.assign(ctx->n_chk_slots, 0);
  {
    uint8_t* scratch = nullptr;
    if ((rc = dev_alloc(ctx, &scratch, abs_sum_scratch_bytes((int)S)))) return rc;   // zeroed: the arrival counters start at 0
    ctx->d_chk_scratch = scratch;
  }
  // mel buffers
  const size_t max_len = (size_t)(ctx->cfg.max_clip_samples / 160);
  ctx->d_mel_floats = (size_t)ctx->cfg.max_clips * ctx->mel_tab.n_mel * (max_len ? max_len : 1);
  if ((rc = dev_alloc(ctx, &ctx->d_mel, ctx->d_mel_floats))) return rc;
  ctx->d_pcm_bytes = (size_t)ctx->cfg.max_clips * (size_t)ctx->cfg.max_clip_samples * 4;
  for (int i = 0; i < 2; ++i) {
    uint8_t* pcm = nullptr;
    if ((rc = dev_alloc(ctx, &pcm, ctx->d_pcm_bytes, false))) return rc;
    ctx->d_pcm_buf[i] = pcm;
    WB_CK(cudaEventCreateWithFlags(&ctx->ev_copy_done[i], cudaEventDisableTiming));
    WB_CK(cudaEventCreateWithFlags(&ctx->ev_mel_read[i], cudaEventDisableTiming));
  }
  ctx->d_pcm = ctx->d_pcm_buf[0];
  for (int i = 0; i < WB_N_TICKETS; ++i) {
    WB_CK(cudaEventCreateWithFlags(&ctx->ev_ticket[i], cudaEventDisableTiming));
    WB_CK(cudaEventCreateWithFlags(&ctx->ev_seg_slot[i], cudaEventDisableTiming));
  }
  WB_CK(cudaHostAlloc((void**)&ctx->h_clip_ids, sizeof(int) * WB_N_TICKETS * ctx->cfg.max_segments, cudaHostAllocDefault));
  WB_CK(cudaHostAlloc((void**)&ctx->h_offsets, sizeof(long long) * WB_N_TICKETS * ctx->cfg.max_segments, cudaHostAllocDefault));
  WB_CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  if ((rc = dev_alloc(ctx, &ctx->d_clip_max, (size_t)ctx->cfg.max_clips))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_seg_max, S))) return rc;
  return WB_OK;
}

int chk_slot(const wb_ctx* ctx, int stage, int layer) {
  const int L = ctx->hp.n_audio_layer, Lt = ctx->hp.n_text_layer;
  switch (stage) {
    case WB_STAGE_MEL: return 0;
    case WB_STAGE_CONV1: return 1;
    case WB_STAGE_CONV2_POS: return 2;
    case WB_STAGE_LAYER: return (layer >= 0 && layer < L) ? 3 + layer : -1;
    case WB_STAGE_LN_POST: return 3 + L;
    case WB_STAGE_CROSS_K: return (layer >= 0 && layer < Lt) ? 4 + L + 2 * layer : -1;
    case WB_STAGE_CROSS_V: return (layer >= 0 && layer < Lt) ? 5 + L + 2 * layer : -1;
    default: return -1;
  }
}

}  // namespace wb

using namespace wb;

extern "C" {

const char* wb_version(void) { return "whisper_b200 0.1 (sm_100a)"; }

const char* wb_last_error(const wb_ctx* ctx) { return ctx ? ctx->err.c_str() : wb::g_err.c_str(); }

void wb_config_default(wb_config* cfg) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->device = 0;
  cfg->max_segments = 1;
  cfg->max_clips = 1;
  cfg->max_clip_samples = 480000;   // WHISPER_SAMPLE_RATE * WHISPER_CHUNK_SIZE (25, 29)
  cfg->norm_scope = WB_NORM_CLIP;
  cfg->checkpoints = 0;
  cfg->stream = nullptr;
  cfg->decode_capacity = 1;
}

int wb_ctx_create(const char* model_path, const wb_config* cfg_in, wb_ctx** out) {
  if (!model_path || !out) return fail_msg(nullptr, WB_ERR_UNEXPECTED, "Unexpected: null argument");
  *out = nullptr;
  const auto t_start = std::chrono::steady_clock::now();
  const bool load_trace = getenv("WB_LOAD_TRACE") != nullptr;   // developer aid: where context creation spends its time
  auto trace = [&](const char* what) {
    if (load_trace)
      fprintf(stderr, "[libwhisper_b200] load %-24s %8.3f s\n", what,
              std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
  };
  wb_config cfg;
  if (cfg_in) cfg = *cfg_in;
  else wb_config_default(&cfg);
  if (cfg.max_segments < 1 || cfg.max_clips < 1 || cfg.max_clip_samples < 400)
    return fail_msg(nullptr, WB_ERR_UNEXPECTED, "Unexpected: bad wb_config capacities");

  ModelFileView mv;
  std::string perr;
  int rc = parse_model_file(model_path, mv, perr);
  if (rc != WB_OK) return fail_msg(nullptr, rc, perr);
  trace("file parsed");
  const ModelHParams& hp = mv.hp;
  if (hp.n_audio_state / hp.n_audio_head != 64 || hp.n_text_state / hp.n_text_head != 64)
    return fail_msg(nullptr, WB_ERR_UNEXPECTED, "Unexpected: head dimension must be 64");
  if (hp.n_audio_state != hp.n_text_state)
    return fail_msg(nullptr, WB_ERR_UNEXPECTED, "Unexpected: n_audio_state != n_text_state");
  if (hp.n_audio_state % 64 != 0 || hp.n_audio_state > 1280 || hp.n_mels % 8 != 0 || mv.filt_n_mel != hp.n_mels)
    return fail_msg(nullptr, WB_ERR_UNEXPECTED, "Unexpected: unsupported model dimensions");
  if (cfg.norm_scope != WB_NORM_CLIP && cfg.norm_scope != WB_NORM_SEGMENT)
    return fail_msg(nullptr, WB_ERR_UNEXPECTED, "Unexpected: norm_scope must be WB_NORM_CLIP or WB_NORM_SEGMENT");

  // ---- device: no CPU fallback
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev <= 0)
    return fail_msg(nullptr, WB_ERR_TENSOR_OP, std::string("galois tensor:'no CUDA device: ") + cudaGetErrorString(e) +
                                                   "' (libwhisper_b200 has no CPU fallback)");
  if (cfg.device < 0 || cfg.device >= n_dev) return fail_msg(nullptr, WB_ERR_UNEXPECTED, "Unexpected: bad device ordinal");
  cudaDeviceProp prop;
  if ((e = cudaSetDevice(cfg.device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, cfg.device)) != cudaSuccess)
    return fail(nullptr, WB_ERR_TENSOR_OP, "cudaSetDevice", e);
  if (prop.major != 10)
    return fail_msg(nullptr, WB_ERR_TENSOR_OP, "galois tensor:'device is not sm_100 (Blackwell B200); kernels are sm_100a only'");

  wb_ctx* ctx = new wb_ctx();
  ctx->cfg = cfg;
  ctx->device = cfg.device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->hp = hp;
  memcpy(ctx->special, mv.special, sizeof(mv.special));
  ctx->vocab = mv.vocab;
  ctx->time_kernels = cfg.reserved[0] != 0;
  {
    const char* cv = getenv("WB_CANARY");
    ctx->canary = cfg.reserved[1] != 0 || (cv && cv[0] == '1');
  }
  auto bail = [&](int code) {
    std::string m = ctx->err;
    wb_ctx_free(ctx);
    wb::g_err = m;
    return code;
  };
  if (cfg.stream) {
    ctx->stream = reinterpret_cast<cudaStream_t>(cfg.stream);
  } else {
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
      fail(ctx, WB_ERR_TENSOR_OP, "cudaStreamCreate", e);
      return bail(WB_ERR_TENSOR_OP);
    }
    ctx->own_stream = true;
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 2; ++j) cudaEventCreate(&ctx->ev[i][j]);
  const char* aerr = "";
  if (!gemm_setup_attributes(&aerr) || !attention_setup_attributes(&aerr)) {
    fail_msg(ctx, WB_ERR_TENSOR_OP, std::string("galois tensor:'kernel attribute setup: ") + aerr + "'");
    return bail(WB_ERR_TENSOR_OP);
  }

  // ---- weights (tensor table of src/main.rs:960-1334)
#define TRY(x)            \
  do {                    \
    rc = (x);             \
    if (rc) return bail(rc); \
  } while (0)
  trace("device + stream ready");
  TRY(build_mel_tables(ctx, mv));
  {
    // LayerNorm folded into the GEMMs around it (default; WB_LN_FOLD=0 keeps the separate LayerNorm kernel):
    // needs the pair kernel for every GEMM involved
    const char* ev = getenv("WB_LN_FOLD");
    const int dd = hp.n_audio_state;
    ctx->ln_fold = !(ev && ev[0] == '0') && !force_gemm1() && gemm2_pick_bn(dd) && gemm2_pick_bn(3 * dd) && gemm2_pick_bn(4 * dd);
  }
  TRY(upload_f32(ctx, mv, "encoder.positional_embedding", &ctx->e_pe));
  TRY(upload_conv(ctx, mv, "encoder.conv1.weight", "encoder.conv1.bias", ctx->conv1));
  TRY(upload_conv(ctx, mv, "encoder.conv2.weight", "encoder.conv2.bias", ctx->conv2));
  TRY(upload_f32(ctx, mv, "encoder.ln_post.weight", &ctx->ln_post_w));
  TRY(upload_f32(ctx, mv, "encoder.ln_post.bias", &ctx->ln_post_b));
  ctx->enc.resize(hp.n_audio_layer);
  for (int i = 0; i < hp.n_audio_layer; ++i) {
    const std::string p = "encoder.blocks." + std::to_string(i) + ".";
    EncLayer& l = ctx->enc[i];
    TRY(upload_f32(ctx, mv, p + "attn_ln.weight", &l.attn_ln_w));
    TRY(upload_f32(ctx, mv, p + "attn_ln.bias", &l.attn_ln_b));
    TRY(upload_f32(ctx, mv, p + "mlp_ln.weight", &l.mlp_ln_w));
    TRY(upload_f32(ctx, mv, p + "mlp_ln.bias", &l.mlp_ln_b));
    // Q (+b), K (no bias), V (+b) fused into one [3d][d] weight (1891-1897)
    const std::vector<CatPart> qkv_parts = {{p + "attn.query.weight", p + "attn.query.bias", 1.0f},
                                            {p + "attn.key.weight", "", 1.0f},
                                            {p + "attn.value.weight", p + "attn.value.bias", 1.0f}};
    if (ctx->ln_fold) {   // attn_ln / mlp_ln folded into the weights that consume them (gemm2.cu, LN == 2)
      TRY(upload_cat_ln(ctx, mv, qkv_parts, p + "attn_ln.weight", p + "attn_ln.bias", l.qkv, false));
      TRY(upload_cat_ln(ctx, mv, {{p + "mlp.0.weight", p + "mlp.0.bias", 1.0f}}, p + "mlp_ln.weight", p + "mlp_ln.bias", l.fc1, false));
    } else {
      TRY(upload_cat(ctx, mv, qkv_parts, l.qkv, false));
      TRY(upload_linear(ctx, mv, p + "mlp.0.weight", p + "mlp.0.bias", l.fc1, false));
    }
    TRY(upload_linear(ctx, mv, p + "attn.out.weight", p + "attn.out.bias", l.out, false));
    TRY(upload_linear(ctx, mv, p + "mlp.2.weight", p + "mlp.2.bias", l.fc2, false));
  }
  {
    // cross-attention K/V of every text layer as one GEMM: K rows scaled by (d/H)^-1/4 with no
    // bias (1992-1996), V rows + bias (2013-2016)
    const float ks = powf((float)hp.n_audio_state / (float)hp.n_audio_head, -0.25f);
    std::vector<CatPart> parts;
    for (int i = 0; i < hp.n_text_layer; ++i) {
      const std::string p = "decoder.blocks." + std::to_string(i) + ".cross_attn.";
      parts.push_back({p + "key.weight", "", ks});
      parts.push_back({p + "value.weight", p + "value.bias", 1.0f});
    }
    if (!parts.empty()) TRY(upload_cat(ctx, mv, parts, ctx->cross_kv, false));
  }
  trace("encoder weights");
  TRY(alloc_activations(ctx));
  trace("activations");
  if (cfg.decode_capacity) TRY(decode_setup(ctx, mv));
  trace("decoder weights + state");   // decoder weights + KV cache (wb_decode.cu)
#undef TRY
  if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {
    fail(ctx, WB_ERR_TENSOR_OP, "weight upload", e);
    return bail(WB_ERR_TENSOR_OP);
  }
  ctx->tm.t_load_us =
      std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t_start).count();
  *out = ctx;
  return WB_OK;
}

void wb_ctx_free(wb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  resolve_kernel_clocks(ctx);
  for (cudaEvent_t e : ctx->free_events) cudaEventDestroy(e);
  ctx->free_events.clear();
  if (ctx->step_graph) cudaGraphExecDestroy(ctx->step_graph);
  for (auto& g : ctx->enc_graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  ctx->enc_graphs.clear();
  for (int g = 0; g < WB_MAX_DEC_GROUPS; ++g) {
    if (ctx->dec_group_stream[g]) cudaStreamDestroy(ctx->dec_group_stream[g]);
    if (ctx->dec_group_done[g]) cudaEventDestroy(ctx->dec_group_done[g]);
  }
  if (ctx->dec_fork) cudaEventDestroy(ctx->dec_fork);
  if (ctx->dec_pf_stream) cudaStreamDestroy(ctx->dec_pf_stream);
  if (ctx->dec_pf_fork) cudaEventDestroy(ctx->dec_pf_fork);
  if (ctx->dec_pf_join) cudaEventDestroy(ctx->dec_pf_join);
  for (void* p : ctx->allocs) cudaFree(p);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 2; ++j)
      if (ctx->ev[i][j]) cudaEventDestroy(ctx->ev[i][j]);
  for (int i = 0; i < 2; ++i) {
    if (ctx->ev_copy_done[i]) cudaEventDestroy(ctx->ev_copy_done[i]);
    if (ctx->ev_mel_read[i]) cudaEventDestroy(ctx->ev_mel_read[i]);
  }
  for (int i = 0; i < WB_N_TICKETS; ++i) {
    if (ctx->ev_ticket[i]) cudaEventDestroy(ctx->ev_ticket[i]);
    if (ctx->ev_seg_slot[i]) cudaEventDestroy(ctx->ev_seg_slot[i]);
  }
  if (ctx->h_clip_ids) cudaFreeHost(ctx->h_clip_ids);
  if (ctx->h_offsets) cudaFreeHost(ctx->h_offsets);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
  }
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int wb_get_hparams(const wb_ctx* ctx, int32_t out[11]) {
  if (!ctx || !out) return WB_ERR_UNEXPECTED;
  memcpy(out, &ctx->hp, 44);
  return WB_OK;
}
int wb_get_special_tokens(const wb_ctx* ctx, int32_t out[8]) {
  if (!ctx || !out) return WB_ERR_UNEXPECTED;
  memcpy(out, ctx->special, 32);
  return WB_OK;
}

int wb_token_text(const wb_ctx* ctx, int32_t id, char* out, size_t cap) {
  if (!ctx || id < 0 || (size_t)id >= ctx->vocab.size()) return WB_ERR_UNEXPECTED;
  const std::string& w = ctx->vocab[(size_t)id];
  if (out && cap) {
    const size_t n = w.size() < cap - 1 ? w.size() : cap - 1;
    memcpy(out, w.data(), n);
    out[n] = 0;
  }
  return (int)w.size();
}

int wb_tokens_to_text(const wb_ctx* ctx, const int32_t* ids, int n, char* out, size_t cap) {
  if (!ctx || (!ids && n > 0) || n < 0) return WB_ERR_UNEXPECTED;
  // text tokens are the ids below eot (557-575): specials, language / task and timestamp ids are skipped
  std::string s;
  for (int i = 0; i < n; ++i)
    if (ids[i] >= 0 && ids[i] < ctx->special[0] && (size_t)ids[i] < ctx->vocab.size()) s += ctx->vocab[(size_t)ids[i]];
  if (out && cap) {
    const size_t m = s.size() < cap - 1 ? s.size() : cap - 1;
    memcpy(out, s.data(), m);
    out[m] = 0;
  }
  return (int)s.size();
}

int wb_sync(wb_ctx* ctx) {
  if (!ctx) return WB_ERR_UNEXPECTED;
  WB_CK(cudaStreamSynchronize(ctx->stream));
  return WB_OK;
}

// ------------------------------------------------------------------------------------------------
// whisper_pcm_to_mel (src/main.rs:1681-1707)
// stage 1: log10-mel of n_len frames per clip + the per-clip maximum (1554-1652, 1655-1662)
static int mel_logmel(wb_ctx* ctx, const void* pcm_dev, int is_i16, size_t n_samples, int n_clips, size_t n_len) {
  const int n_mel = ctx->mel_tab.n_mel;
  if (n_clips < 1 || n_clips > ctx->cfg.max_clips || (size_t)n_clips * n_mel * n_len > ctx->d_mel_floats)
    return fail_msg(ctx, WB_ERR_NOT_ENOUGH_SPACE, "not enough space in the context's memory pool\n");
  cudaEventRecord(ctx->ev[0][0], ctx->stream);
  {
    LaunchTimer t(ctx, "fill");
    WB_CK(launch_fill_i32(ctx->d_clip_max, n_clips, mel_enc_ordered_host(-1e20f), ctx->stream));   // mmax = -1e20 (1655)
  }
  {
    LaunchTimer t(ctx, "mel_frames");
    WB_CK(launch_mel_frames(ctx->mel_tab, pcm_dev, is_i16, n_samples, n_clips, (int)n_len, ctx->d_mel, ctx->d_clip_max,
                            ctx->stream));
  }
  ctx->mel_n_len = (int)n_len;
  ctx->mel_n_clips = n_clips;
  ctx->mel_normalized = false;
  ctx->mel_materialized = false;
  return WB_OK;
}

// clamp_and_normalize (1654-1671) of the stored mel in place, with the maxima in d_clip_max
static int mel_materialize(wb_ctx* ctx) {
  if (ctx->mel_materialized || ctx->cfg.norm_scope != WB_NORM_CLIP) return WB_OK;
  LaunchTimer t(ctx, "mel_normalize");
  WB_CK(launch_mel_normalize(ctx->d_mel, ctx->mel_n_clips, (size_t)ctx->mel_tab.n_mel * (size_t)ctx->mel_n_len, ctx->d_clip_max,
                             ctx->stream));
  ctx->mel_materialized = true;
  return WB_OK;
}

// stage 2: clamp_and_normalize (1654-1671) with the maxima in d_clip_max.  The maxima are final from here on; the
// two f32 operations themselves run inside the encoder's window copy (mel_window_kernel), which reads every mel
// value anyway -- the stored mel is rewritten in place only when somebody looks at it (wb_mel_read, checkpoints).
static int mel_norm(wb_ctx* ctx) {
  const int n_mel = ctx->mel_tab.n_mel, n_clips = ctx->mel_n_clips;
  const size_t n_len = (size_t)ctx->mel_n_len;
  ctx->mel_normalized = true;
  if (ctx->cfg.checkpoints) {
    int rc = mel_materialize(ctx);
    if (rc) return rc;
    WB_CK(launch_abs_sum_f32(ctx->d_mel, (long long)n_mel * n_len, (long long)n_mel * n_len,
                             n_clips < ctx->cfg.max_segments ? n_clips : ctx->cfg.max_segments, ctx->d_chk, ctx->d_chk_scratch,
                             ctx->cfg.max_segments, ctx->stream));
    ctx->chk_valid[0] = 1;
  }
  cudaEventRecord(ctx->ev[0][1], ctx->stream);
  ctx->ev_used[0] = true;
  ctx->tm.n_mel_calls += 1;
  return WB_OK;
}

static int mel_run(wb_ctx* ctx, const void* pcm_dev, int is_i16, size_t n_samples, int n_clips) {
  int rc = mel_logmel(ctx, pcm_dev, is_i16, n_samples, n_clips, n_samples / 160);   // n_len, 1575
  return rc ? rc : mel_norm(ctx);
}

int wb_pcm_to_mel_device(wb_ctx* ctx, const float* pcm_dev, size_t n_samples, int n_clips) {
  if (!ctx || !pcm_dev) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  return mel_run(ctx, pcm_dev, 0, n_samples, n_clips);
}

// host PCM -> the staging buffer the mel kernel reads.  If wb_pcm_prefetch already uploaded exactly
// these bytes into the other buffer, switch to it and only wait for that copy.
static int stage_host_pcm(wb_ctx* ctx, const void* pcm, size_t bytes) {
  if (bytes > ctx->d_pcm_bytes)
    return fail_msg(ctx, WB_ERR_NOT_ENOUGH_SPACE, "not enough space in the context's memory pool\n");
  const int other = ctx->pcm_cur ^ 1;
  if (ctx->pf_host[other] == pcm && ctx->pf_bytes[other] == bytes) {
    ctx->pcm_cur = other;
    ctx->pf_host[other] = nullptr;
    WB_CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy_done[other], 0));
  } else {
    // a prefetch that is not consumed by the very next upload is dropped: matching a LATER call by pointer
    // identity would run the mel on whatever the buffer held when it was prefetched
    ctx->pf_host[other] = nullptr;
    ctx->pf_bytes[other] = 0;
    WB_CK(cudaMemcpyAsync(ctx->d_pcm_buf[ctx->pcm_cur], pcm, bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  ctx->d_pcm = ctx->d_pcm_buf[ctx->pcm_cur];
  return WB_OK;
}

int wb_pcm_to_mel(wb_ctx* ctx, const float* pcm, size_t n_samples, int n_clips) {
  if (!ctx || !pcm || n_clips < 1) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  int rc = stage_host_pcm(ctx, pcm, (size_t)n_clips * n_samples * 4);
  if (rc) return rc;
  rc = mel_run(ctx, ctx->d_pcm, 0, n_samples, n_clips);
  if (rc == WB_OK) WB_CK(cudaEventRecord(ctx->ev_mel_read[ctx->pcm_cur], ctx->stream));
  return rc;
}

int wb_pcm16_to_mel(wb_ctx* ctx, const int16_t* pcm, size_t n_samples, int n_clips) {
  if (!ctx || !pcm || n_clips < 1) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  int rc = stage_host_pcm(ctx, pcm, (size_t)n_clips * n_samples * 2);
  if (rc) return rc;
  rc = mel_run(ctx, ctx->d_pcm, 1, n_samples, n_clips);
  if (rc == WB_OK) WB_CK(cudaEventRecord(ctx->ev_mel_read[ctx->pcm_cur], ctx->stream));
  return rc;
}

// Two-phase whisper_pcm_to_mel for a clip whose samples are split across GPUs: the
// whole-clip maximum of clamp_and_normalize (1655-1662) is the one coupling between the parts.
int wb_pcm_to_logmel(wb_ctx* ctx, const float* pcm, size_t n_samples, int n_clips, int n_frames) {
  if (!ctx || !pcm || n_clips < 1 || n_frames < 0) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  int rc = stage_host_pcm(ctx, pcm, (size_t)n_clips * n_samples * 4);
  if (rc) return rc;
  rc = mel_logmel(ctx, ctx->d_pcm, 0, n_samples, n_clips, n_frames ? (size_t)n_frames : n_samples / 160);
  if (rc == WB_OK) WB_CK(cudaEventRecord(ctx->ev_mel_read[ctx->pcm_cur], ctx->stream));
  return rc;
}

int wb_mel_max_read(wb_ctx* ctx, float* out, int n_clips) {
  if (!ctx || !out || n_clips < 1 || n_clips > ctx->mel_n_clips) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  std::vector<int> enc((size_t)n_clips);
  WB_CK(cudaMemcpyAsync(enc.data(), ctx->d_clip_max, sizeof(int) * n_clips, cudaMemcpyDeviceToHost, ctx->stream));
  WB_CK(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n_clips; ++i) out[i] = mel_dec_ordered_host(enc[(size_t)i]);
  return WB_OK;
}

int wb_mel_normalize(wb_ctx* ctx, const float* clip_max, int n_clips) {
  if (!ctx || n_clips != ctx->mel_n_clips || ctx->mel_n_clips < 1) return WB_ERR_UNEXPECTED;
  if (ctx->mel_normalized) return fail_msg(ctx, WB_ERR_UNEXPECTED, "Unexpected error: mel is already normalised\n");
  cudaSetDevice(ctx->device);
  if (clip_max) {
    std::vector<int> enc((size_t)n_clips);
    for (int i = 0; i < n_clips; ++i) enc[(size_t)i] = mel_enc_ordered_host(clip_max[i]);
    WB_CK(cudaMemcpyAsync(ctx->d_clip_max, enc.data(), sizeof(int) * n_clips, cudaMemcpyHostToDevice, ctx->stream));
    WB_CK(cudaStreamSynchronize(ctx->stream));   // enc is a stack-lifetime host buffer
  }
  return mel_norm(ctx);
}

int wb_pcm_prefetch(wb_ctx* ctx, const void* pcm, size_t n_bytes) {
  if (!ctx || !pcm) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  if (n_bytes > ctx->d_pcm_bytes)
    return fail_msg(ctx, WB_ERR_NOT_ENOUGH_SPACE, "not enough space in the context's memory pool\n");
  const int other = ctx->pcm_cur ^ 1;
  // the mel kernel that last read this buffer must have finished before it is overwritten
  WB_CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_mel_read[other], 0));
  WB_CK(cudaMemcpyAsync(ctx->d_pcm_buf[other], pcm, n_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
  WB_CK(cudaEventRecord(ctx->ev_copy_done[other], ctx->copy_stream));
  ctx->pf_host[other] = pcm;
  ctx->pf_bytes[other] = n_bytes;
  return WB_OK;
}

int wb_mel_dims(const wb_ctx* ctx, int* n_mel, int* n_len, int* n_clips) {
  if (!ctx) return WB_ERR_UNEXPECTED;
  if (n_mel) *n_mel = ctx->mel_tab.n_mel;
  if (n_len) *n_len = ctx->mel_n_len;
  if (n_clips) *n_clips = ctx->mel_n_clips;
  return WB_OK;
}

int wb_mel_read(wb_ctx* ctx, int clip, float* out, size_t cap_floats) {
  if (!ctx || !out || clip < 0 || clip >= ctx->mel_n_clips) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  const size_t n = (size_t)ctx->mel_tab.n_mel * ctx->mel_n_len;
  if (cap_floats < n) return fail_msg(ctx, WB_ERR_NOT_ENOUGH_SPACE, "not enough space in the context's memory pool\n");
  if (ctx->mel_normalized) {   // ctx.mel holds normalised values in the reference (1648): apply them now if still pending
    int rc = mel_materialize(ctx);
    if (rc) return rc;
  }
  WB_CK(cudaMemcpyAsync(out, ctx->d_mel + (size_t)clip * n, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  WB_CK(cudaStreamSynchronize(ctx->stream));
  return WB_OK;
}

int wb_mel_write(wb_ctx* ctx, const float* mel, int n_mel, int n_len, int n_clips) {
  if (!ctx || !mel || n_mel != ctx->mel_tab.n_mel || n_clips < 1 || n_clips > ctx->cfg.max_clips) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  const size_t n = (size_t)n_clips * n_mel * n_len;
  if (n > ctx->d_mel_floats) return fail_msg(ctx, WB_ERR_NOT_ENOUGH_SPACE, "not enough space in the context's memory pool\n");
  WB_CK(cudaMemcpyAsync(ctx->d_mel, mel, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  WB_CK(cudaStreamSynchronize(ctx->stream));
  ctx->mel_n_len = n_len;
  ctx->mel_n_clips = n_clips;
  ctx->mel_normalized = true;
  ctx->mel_materialized = true;   // the caller's values are used as they are
  return WB_OK;
}

// ------------------------------------------------------------------------------------------------
// whisper_encode (src/main.rs:1799-2063), batched over segments
int wb_encode(wb_ctx* ctx, const int32_t* clip_ids, const size_t* mel_offsets, int n_seg) {
  if (!ctx) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  const ModelHParams& hp = ctx->hp;
  if (n_seg < 1 || n_seg > ctx->cfg.max_segments)
    return fail_msg(ctx, WB_ERR_NOT_ENOUGH_SPACE, "not enough space in the context's memory pool\n");
  if (ctx->mel_n_clips < 1 || ctx->mel_tab.n_mel != hp.n_mels)   // assert 1813
    return fail_msg(ctx, WB_ERR_UNEXPECTED, "Unexpected: no mel in the context (call wb_pcm_to_mel first)");
  if (!ctx->mel_normalized)
    return fail_msg(ctx, WB_ERR_UNEXPECTED, "Unexpected: the mel is not normalised (call wb_mel_normalize after wb_pcm_to_logmel)");
  // 1803-1807: exp_n_audio_ctx when positive (wb_set_audio_ctx), else the model's audio context
  const int T = ctx->exp_n_audio_ctx > 0 ? ctx->exp_n_audio_ctx : hp.n_audio_ctx;
  const int Tm = 2 * T, d = hp.n_audio_state, H = hp.n_audio_head, L = hp.n_audio_layer;
  const int Lt = hp.n_text_layer, n_mels = hp.n_mels;
  const int M = n_seg * T;
  const bool chk = ctx->cfg.checkpoints != 0;
  cudaStream_t st = ctx->stream;
  const char* terr = "";

  // segment table: through a ring of pinned host slots, so the call never waits for the stream (a caller may
  // queue the next batch while this one runs); a slot is reused only after the copy that read it has executed
  const int slot = ctx->seg_slot_next;
  ctx->seg_slot_next = (slot + 1) % WB_N_TICKETS;
  WB_CK(cudaEventSynchronize(ctx->ev_seg_slot[slot]));
  int* ids = ctx->h_clip_ids + (size_t)slot * ctx->cfg.max_segments;
  long long* offs = ctx->h_offsets + (size_t)slot * ctx->cfg.max_segments;
  for (int s = 0; s < n_seg; ++s) {
    ids[s] = clip_ids ? clip_ids[s] : 0;
    offs[s] = mel_offsets ? (long long)mel_offsets[s] : 0;
    if (ids[s] < 0 || ids[s] >= ctx->mel_n_clips) return fail_msg(ctx, WB_ERR_UNEXPECTED, "Unexpected: clip id out of range");
  }
  cudaEventRecord(ctx->ev[1][0], st);
  WB_CK(cudaMemcpyAsync(ctx->d_clip_ids, ids, sizeof(int) * n_seg, cudaMemcpyHostToDevice, st));
  WB_CK(cudaMemcpyAsync(ctx->d_offsets, offs, sizeof(long long) * n_seg, cudaMemcpyHostToDevice, st));
  WB_CK(cudaEventRecord(ctx->ev_seg_slot[slot], st));

  // ---- tensor maps of the activation operands: they depend only on the batch size and the audio context, so
  // they are encoded once per (n_seg, T) and kept in the handle (17 cuTensorMapEncodeTiled calls per encode
  // otherwise: a third of the call's host time)
  wb::EncodeMaps& em = ctx->enc_maps;
  CUtensorMap &m_conv1 = em.m_conv1, &m_conv2 = em.m_conv2, &m_ln = em.m_ln, &m_att = em.m_att, &m_hid = em.m_hid, &m_enc = em.m_enc;
  CUtensorMap &o_conv1 = em.o_conv1, &o_x3 = em.o_x3, &o_pe = em.o_pe, &o_x = em.o_x, &o_qk = em.o_qk, &o_hid = em.o_hid, &o_cross = em.o_cross;
  AttnProblem& ap = em.ap;
  if (em.n_seg != n_seg || em.T != T) {
  em.n_seg = 0;
  bool ok = tmap_3d_rows(&m_conv1, ctx->conv_in, 3 * n_mels, Tm, n_seg, n_mels, (uint64_t)(Tm + 2) * n_mels, &terr) &&
            tmap_3d_rows(&m_conv2, ctx->h1, 3 * d, T, n_seg, 2 * d, (uint64_t)(Tm + 2) * d, &terr) &&
            tmap_3d_rows(&m_ln, ctx->ln_out, d, M, 1, d, (uint64_t)M * d, &terr) &&
            tmap_3d_rows(&m_att, ctx->attn_out, d, M, 1, d, (uint64_t)M * d, &terr) &&
            tmap_3d_rows(&m_hid, ctx->hidden, 4 * d, M, 1, 4 * d, (uint64_t)M * 4 * d, &terr) &&
            tmap_3d_rows(&m_enc, ctx->enc_f16, d, M, 1, d, (uint64_t)M * d, &terr);
  if (ok) {
    // Q|K buffer [seg*T][2d] viewed as {64, 2H, T, seg}
    const uint64_t dims[4] = {64, (uint64_t)2 * H, (uint64_t)T, (uint64_t)n_seg};
    const uint64_t strd[3] = {128, (uint64_t)2 * d * 2, (uint64_t)T * 2 * d * 2};
    const uint32_t box[4] = {64, 1, 128, 1};
    ok = make_tmap_f16(&ap.qk_map, ctx->qk, 4, dims, strd, box, &terr);
  }
  if (ok) {
    const uint64_t dims[2] = {(uint64_t)ctx->Tp, (uint64_t)n_seg * H * ATTN_VT_HEAD_ROWS};
    const uint64_t strd[1] = {(uint64_t)ctx->Tp * 2};
    const uint32_t box[2] = {64, (uint32_t)ATTN_VT_HEAD_ROWS};
    ok = make_tmap_f16(&ap.vt_map, ctx->vt, 2, dims, strd, box, &terr);
  }
  // epilogue output / residual boxes of the pair GEMM
  ok = ok && tmap_out(&o_conv1, ctx->h1 + d, false, d, Tm, n_seg, d, (uint64_t)(Tm + 2) * d, &terr) &&
       tmap_out(&o_x3, ctx->x, true, d, T, n_seg, d, (uint64_t)T * d, &terr) &&
       tmap_out(&o_pe, ctx->e_pe, true, d, T, 1, d, 0, &terr) &&
       tmap_out(&o_x, ctx->x, true, d, M, 1, d, 0, &terr) &&
       tmap_out(&o_qk, ctx->qk, false, 2 * d, M, 1, 2 * d, 0, &terr) &&
       tmap_out(&o_hid, ctx->hidden, false, 4 * d, M, 1, 4 * d, 0, &terr) &&
       (Lt == 0 || tmap_out(&o_cross, ctx->cross, false, d, M, (uint64_t)2 * Lt, d, ctx->cross_slab, &terr));
  if (!ok) return fail_msg(ctx, WB_ERR_TENSOR_OP, std::string("galois tensor:'") + terr + "'");
  ap.B = n_seg;
  ap.T = T;
  ap.H = H;
  ap.out = ctx->attn_out;
  ap.scale = 1.0f / sqrtf(64.0f);
  em.n_seg = n_seg;
  em.T = T;
  }

  auto probe_f32 = [&](int slot, const float* p, long long per_seg, long long seg_stride) -> int {
    WB_CK(launch_abs_sum_f32(p, per_seg, seg_stride, n_seg, ctx->d_chk + (size_t)slot * ctx->cfg.max_segments, ctx->d_chk_scratch,
                             ctx->cfg.max_segments, st));
    ctx->chk_valid[slot] = 1;
    return WB_OK;
  };
  auto probe_f16 = [&](int slot, const __half* p, int rows, int cols, long long row_stride, long long seg_stride) -> int {
    WB_CK(launch_abs_sum_f16(p, rows, cols, row_stride, seg_stride, n_seg, ctx->d_chk + (size_t)slot * ctx->cfg.max_segments, ctx->d_chk_scratch,
                             ctx->cfg.max_segments, st));
    ctx->chk_valid[slot] = 1;
    return WB_OK;
  };
  int rc;
  // conv_in and h1 hold [(Tm + 2) rows] per segment with zero rows 0 and Tm + 1 (the convolutions' padding); the
  // kernels only ever write rows 1 .. Tm.  When the audio context changes the pad rows move onto memory that held
  // activations of the previous layout, so both buffers are cleared once per change of T.
  if (ctx->pad_T != T) {
    const size_t S = (size_t)ctx->cfg.max_segments, Tm_max = 2 * (size_t)hp.n_audio_ctx;
    WB_CK(cudaMemsetAsync(ctx->conv_in, 0, S * (Tm_max + 2) * n_mels * sizeof(__half), st));
    WB_CK(cudaMemsetAsync(ctx->h1, 0, S * (Tm_max + 2) * d * sizeof(__half), st));
    ctx->pad_T = T;
  }
  const bool fold = ctx->ln_fold;
  const int norm_mode = ctx->mel_materialized ? 0 : ctx->cfg.norm_scope == WB_NORM_SEGMENT ? 2 : 1;

  // every launch of one encode: run directly, or captured once per shape into a CUDA graph and replayed (below)
  auto encode_launches = [&]() -> int {
  // E0: mel window -> token-major F16 rows with zero padding rows (1816-1829)
  // fused with clamp_and_normalize (1654-1671) when the stored mel still holds log10 values: with the clip's maximum
  // (WB_NORM_CLIP, the reference) or the window's own (WB_NORM_SEGMENT)
  if (norm_mode == 2) {
    LaunchTimer t(ctx, "mel_window_max");
    WB_CK(launch_fill_i32(ctx->d_seg_max, n_seg, mel_enc_ordered_host(-1e20f), st));
    WB_CK(launch_mel_window_max(ctx->d_mel, n_mels, ctx->mel_n_len, ctx->d_clip_ids, ctx->d_offsets, n_seg, Tm, ctx->d_seg_max, st));
  }
  {
    LaunchTimer t(ctx, "mel_window");
    WB_CK(launch_mel_window(ctx->d_mel, n_mels, ctx->mel_n_len, ctx->d_clip_ids, ctx->d_offsets, n_seg, Tm, ctx->conv_in, st,
                            norm_mode == 2 ? ctx->d_seg_max : ctx->d_clip_max, norm_mode));
  }
  // E1: conv1 + bias + GELU (1834-1855) as an implicit GEMM: row t = input rows t-1, t, t+1
  {
    GemmEpilogue e;
    e.gelu = 1;
    e.out = ctx->h1 + d;   // skip the leading zero row
    e.out_f16 = 1;
    e.out_bstride = (long long)(Tm + 2) * d;
    e.out_ld = d;
    if ((rc = run_gemm(ctx, m_conv1, Tm, n_seg, ctx->conv1, e, "gemm_conv1", &o_conv1))) return rc;
    if (chk && (rc = probe_f16(1, ctx->h1 + d, Tm, d, d, (long long)(Tm + 2) * d))) return rc;
  }
  // E2 + E3: conv2 (stride 2) + bias + GELU, + positional embedding (1856-1875) -> residual stream
  {
    GemmEpilogue e;
    e.gelu = 1;
    e.residual = ctx->e_pe;
    e.res_bstride = 0;
    e.res_ld = d;
    e.out = ctx->x;
    e.out_f16 = 0;
    e.out_bstride = (long long)T * d;
    e.out_ld = d;
    if ((rc = run_gemm(ctx, m_conv2, T, n_seg, ctx->conv2, e, "gemm_conv2", &o_x3, &o_pe, 1))) return rc;
    if (chk && (rc = probe_f32(2, ctx->x, (long long)T * d, (long long)T * d))) return rc;
  }
  for (int il = 0; il < L; ++il) {   // 1877-1975
    const EncLayer& l = ctx->enc[il];
    if (!fold) {   // E4: attn_ln
      LaunchTimer t(ctx, "layernorm");
      WB_CK(launch_layernorm(ctx->x, l.attn_ln_w, l.attn_ln_b, M, d, ctx->ln_out, nullptr, st, 0, true));
    } else if (il == 0) {
      // folded: layer 0's attn_ln is the one LayerNorm that still runs as a kernel (gamma / beta are in the QKV
      // weights, so plain normalisation): it gives every row the centre the folded LayerNorms downstream
      // start from -- the producers round x - centre to F16, not x
      LaunchTimer t(ctx, "layernorm");
      WB_CK(launch_layernorm(ctx->x, ctx->enc_ones, ctx->enc_zeros, M, d, ctx->ln_out, nullptr, st, 0, true, ctx->ln_center));
    }
    {   // E5 + E6: fused Q|K|V projection, F16 repack; V transposed, time contiguous (1891-1920)
      GemmEpilogue e;
      e.out = ctx->qk;
      e.out_f16 = 1;
      e.out_ld = 2 * d;
      e.vt_out = ctx->vt;
      e.vt_col0 = 2 * d;
      e.vt_heads = H;
      e.vt_head_rows = ATTN_VT_HEAD_ROWS;
      e.vt_ld = ctx->Tp;
      e.vt_T = T;
      if (fold && il > 0) {   // statistics left by the previous layer's fc2
        e.ln_part_in = ctx->ln_part[0];
        e.ln_parts = ctx->ln_parts;
        e.ln_center = ctx->ln_center;
      }
      if ((rc = run_gemm(ctx, m_ln, M, 1, l.qkv, e, "gemm_qkv", &o_qk))) return rc;
    }
    {   // E7: flash attention + head merge (1922-1929)
      LaunchTimer t(ctx, "attention");
      WB_CK(run_attention(ap, st));
    }
    {   // E8: output projection + bias + residual (1936-1942), in place on the residual stream
      GemmEpilogue e;
      e.residual = ctx->x;
      e.res_ld = d;
      e.out = ctx->x;
      e.out_f16 = 0;
      e.out_ld = d;
      if (fold) {   // feeds mlp_ln
        e.ln_part_out = ctx->ln_part[1];
        e.ln_parts = ctx->ln_parts;
        e.ln_center = ctx->ln_center;
        e.x16_out = ctx->ln_out;
        e.x16_ld = d;
      }
      if ((rc = run_gemm(ctx, m_att, M, 1, l.out, e, "gemm_out", &o_x, &o_x))) return rc;
    }
    if (!fold) {   // E9: mlp_ln, fc1 + bias + GELU, fc2 + bias + residual (1948-1968)
      LaunchTimer t(ctx, "layernorm");
      WB_CK(launch_layernorm(ctx->x, l.mlp_ln_w, l.mlp_ln_b, M, d, ctx->ln_out, nullptr, st, 0, true));
    }
    {
      GemmEpilogue e;
      e.gelu = 1;
      e.out = ctx->hidden;
      e.out_f16 = 1;
      e.out_ld = 4 * d;
      if (fold) {
        e.ln_part_in = ctx->ln_part[1];
        e.ln_parts = ctx->ln_parts;
        e.ln_center = ctx->ln_center;
      }
      if ((rc = run_gemm(ctx, m_ln, M, 1, l.fc1, e, "gemm_fc1", &o_hid))) return rc;
    }
    {
      GemmEpilogue e;
      e.residual = ctx->x;
      e.res_ld = d;
      e.out = ctx->x;
      e.out_f16 = 0;
      e.out_ld = d;
      if (fold && il + 1 < L) {   // feeds attn_ln of the next layer (ln_post keeps its own kernel: it has an f32 output)
        e.ln_part_out = ctx->ln_part[0];
        e.ln_parts = ctx->ln_parts;
        e.ln_center = ctx->ln_center;
        e.x16_out = ctx->ln_out;
        e.x16_ld = d;
      }
      if ((rc = run_gemm(ctx, m_hid, M, 1, l.fc2, e, "gemm_fc2", &o_x, &o_x))) return rc;
    }
    if (chk && (rc = probe_f32(3 + il, ctx->x, (long long)T * d, (long long)T * d))) return rc;
  }
  {   // E11: ln_post (1980-1984): f32 copy for read-back, F16 copy as the cross-KV GEMM operand
    LaunchTimer t(ctx, "layernorm");
    WB_CK(launch_layernorm(ctx->x, ctx->ln_post_w, ctx->ln_post_b, M, d, ctx->enc_f16, ctx->enc_out, st, 0, true));
  }
  if (chk && (rc = probe_f32(3 + L, ctx->enc_out, (long long)T * d, (long long)T * d))) return rc;
  if (Lt > 0) {   // E12: cross-attention K/V of all text layers in one GEMM (1990-2030)
    GemmEpilogue e;
    e.out = ctx->cross;
    e.out_f16 = 1;
    e.out_ld = d;
    e.out_slab_cols = d;                       // column block j of the fused weight -> slab j ([rows][d])
    e.out_slab_stride = (long long)ctx->cross_slab;
    if ((rc = run_gemm(ctx, m_enc, M, 1, ctx->cross_kv, e, "gemm_cross", &o_cross))) return rc;
    if (chk) {
      for (int il = 0; il < Lt; ++il) {
        if ((rc = probe_f16(4 + L + 2 * il, ctx->cross + (size_t)(2 * il) * ctx->cross_slab, T, d, d, (long long)T * d))) return rc;
        if ((rc = probe_f16(5 + L + 2 * il, ctx->cross + (size_t)(2 * il + 1) * ctx->cross_slab, T, d, d, (long long)T * d))) return rc;
      }
    }
  }
  return WB_OK;
  };

  // The ~6 L + 6 launches of an encode depend only on (n_seg, T, the mel's frame count, the normalisation mode):
  // captured once per such shape and replayed, so a call costs one graph launch of host time instead of one launch
  // per kernel (the segment table is uploaded outside the graph: its pinned source slot changes from call to call).
  // Not used with per-kernel timing or checkpoints (both add launches that depend on the mode).  WB_ENC_GRAPH=0: off.
  static const bool graph_off = [] { const char* e = getenv("WB_ENC_GRAPH"); return e && e[0] == '0'; }();
  if (!graph_off && !ctx->time_kernels && !chk) {
    // a few shapes are kept (a long clip runs full batches and one ragged tail; a server alternates batch sizes)
    wb::EncodeGraph* found = nullptr;
    for (auto& g : ctx->enc_graphs)
      if (g.n_seg == n_seg && g.T == T && g.mel_n_len == ctx->mel_n_len && g.norm_mode == norm_mode) found = &g;
    if (!found) {
      // first call of a shape: run directly (each kernel instantiation opts in to its shared-memory size at its first
      // launch -- not something to do inside a capture); the next call of the same shape captures
      if (ctx->enc_graphs.size() >= WB_MAX_ENC_GRAPHS) {   // forget the oldest shape
        if (ctx->enc_graphs.front().exec) cudaGraphExecDestroy(ctx->enc_graphs.front().exec);
        ctx->enc_graphs.erase(ctx->enc_graphs.begin());
      }
      wb::EncodeGraph g;
      g.n_seg = n_seg;
      g.T = T;
      g.mel_n_len = ctx->mel_n_len;
      g.norm_mode = norm_mode;
      ctx->enc_graphs.push_back(g);
      if ((rc = encode_launches())) return rc;
    } else {
    wb::EncodeGraph& eg = *found;
    if (!eg.exec) {
      const int64_t l0 = ctx->tm.n_kernel_launches;
      cudaGraph_t graph = nullptr;
      WB_CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      rc = encode_launches();
      const cudaError_t e_end = cudaStreamEndCapture(st, &graph);
      eg.launches = (int)(ctx->tm.n_kernel_launches - l0);
      ctx->tm.n_kernel_launches = l0;   // counted per replay below
      if (rc != WB_OK) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
      }
      if (e_end != cudaSuccess) return fail(ctx, WB_ERR_TENSOR_OP, "cudaStreamEndCapture (encode)", e_end);
      const cudaError_t e_inst = cudaGraphInstantiate(&eg.exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e_inst != cudaSuccess) return fail(ctx, WB_ERR_TENSOR_OP, "cudaGraphInstantiate (encode)", e_inst);
    }
    WB_CK(cudaGraphLaunch(eg.exec, st));
    ctx->tm.n_kernel_launches += eg.launches;
    }
  } else if ((rc = encode_launches())) {
    return rc;
  }
  cudaEventRecord(ctx->ev[1][1], st);
  ctx->ev_used[1] = true;
  ctx->enc_n_seg = n_seg;
  ctx->enc_T = T;
  ctx->tm.n_encode_calls += 1;
  return WB_OK;
}

// exp_n_audio_ctx (src/main.rs:362, read at 1803-1807; the reference initialises it to 0 and never sets it)
int wb_set_audio_ctx(wb_ctx* ctx, int n_ctx) {
  if (!ctx) return WB_ERR_UNEXPECTED;
  if (n_ctx < 0 || n_ctx > ctx->hp.n_audio_ctx)   // the positional embedding has n_audio_ctx rows (960)
    return fail_msg(ctx, WB_ERR_UNEXPECTED, "Unexpected: audio context out of range");
  ctx->exp_n_audio_ctx = n_ctx;
  return WB_OK;
}

int wb_encoder_out_read(wb_ctx* ctx, int seg, float* out) {
  if (!ctx || !out || seg < 0 || seg >= ctx->enc_n_seg) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  const size_t n = (size_t)ctx->enc_T * ctx->hp.n_audio_state;
  WB_CK(cudaMemcpyAsync(out, ctx->enc_out + (size_t)seg * n, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  WB_CK(cudaStreamSynchronize(ctx->stream));
  return WB_OK;
}

int wb_cross_kv_read(wb_ctx* ctx, int seg, int layer, uint16_t* k, uint16_t* v) {
  if (!ctx || seg < 0 || seg >= ctx->enc_n_seg || layer < 0 || layer >= ctx->hp.n_text_layer) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  const size_t n = (size_t)ctx->enc_T * ctx->hp.n_audio_state;
  // the reference's dense [n_ctx][d] slice of memory_cross_k/v (2018-2030)
  const __half* kb = ctx->cross + (size_t)(2 * layer) * ctx->cross_slab + (size_t)seg * n;
  const __half* vb = ctx->cross + (size_t)(2 * layer + 1) * ctx->cross_slab + (size_t)seg * n;
  if (k) WB_CK(cudaMemcpyAsync(k, kb, n * 2, cudaMemcpyDeviceToHost, ctx->stream));
  if (v) WB_CK(cudaMemcpyAsync(v, vb, n * 2, cudaMemcpyDeviceToHost, ctx->stream));
  WB_CK(cudaStreamSynchronize(ctx->stream));
  return WB_OK;
}

int wb_checksum(wb_ctx* ctx, int stage, int layer, int seg, double* abs_sum) {
  if (!ctx || !abs_sum) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  const int slot = chk_slot(ctx, stage, layer);
  if (!ctx->cfg.checkpoints || slot < 0 || !ctx->chk_valid[slot] || seg < 0 || seg >= ctx->cfg.max_segments)
    return fail_msg(ctx, WB_ERR_UNEXPECTED, "Unexpected: checkpoint not recorded (wb_config.checkpoints = 1?)");
  WB_CK(cudaMemcpyAsync(abs_sum, ctx->d_chk + (size_t)slot * ctx->cfg.max_segments + seg, 8, cudaMemcpyDeviceToHost,
                        ctx->stream));
  WB_CK(cudaStreamSynchronize(ctx->stream));
  return WB_OK;
}

int wb_encoder_digest(wb_ctx* ctx, double* out, int cap) {
  if (!ctx || !out || ctx->enc_n_seg < 1 || cap < ctx->enc_n_seg) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  const long long n = (long long)ctx->enc_T * ctx->hp.n_audio_state;
  double* slot = ctx->d_chk + (size_t)(3 + ctx->hp.n_audio_layer) * ctx->cfg.max_segments;   // the LN_POST slot
  {
    LaunchTimer t(ctx, "digest");
    WB_CK(launch_abs_sum_f32(ctx->enc_out, n, n, ctx->enc_n_seg, slot, ctx->d_chk_scratch, ctx->cfg.max_segments, ctx->stream));
  }
  WB_CK(cudaMemcpyAsync(out, slot, sizeof(double) * ctx->enc_n_seg, cudaMemcpyDeviceToHost, ctx->stream));
  WB_CK(cudaStreamSynchronize(ctx->stream));
  return WB_OK;
}

// The same digest without blocking: the read-back is queued on the handle's stream behind the encode
// that produced it, so the caller can submit the next batch before it waits for this one.
int wb_encoder_digest_async(wb_ctx* ctx, double* out, int cap) {
  if (!ctx || !out || ctx->enc_n_seg < 1 || cap < ctx->enc_n_seg) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  const long long n = (long long)ctx->enc_T * ctx->hp.n_audio_state;
  double* slot = ctx->d_chk + (size_t)(3 + ctx->hp.n_audio_layer) * ctx->cfg.max_segments;   // the LN_POST slot
  {
    LaunchTimer t(ctx, "digest");
    WB_CK(launch_abs_sum_f32(ctx->enc_out, n, n, ctx->enc_n_seg, slot, ctx->d_chk_scratch, ctx->cfg.max_segments, ctx->stream));
  }
  WB_CK(cudaMemcpyAsync(out, slot, sizeof(double) * ctx->enc_n_seg, cudaMemcpyDeviceToHost, ctx->stream));
  const int ticket = ctx->ticket_next;
  ctx->ticket_next = (ticket + 1) % WB_N_TICKETS;
  WB_CK(cudaEventRecord(ctx->ev_ticket[ticket], ctx->stream));
  return ticket;
}

int wb_wait(wb_ctx* ctx, int ticket) {
  if (!ctx || ticket < 0 || ticket >= WB_N_TICKETS) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  WB_CK(cudaEventSynchronize(ctx->ev_ticket[ticket]));
  return WB_OK;
}

int wb_timings_get(const wb_ctx* ctx_c, wb_timings* out) {
  wb_ctx* ctx = const_cast<wb_ctx*>(ctx_c);
  if (!ctx || !out) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  // each event pair brackets the most recent call of its kind (device time on the handle's stream)
  int64_t* slots[3] = {&ctx->tm.t_mel_us, &ctx->tm.t_encode_us, &ctx->tm.t_decode_us};
  for (int i = 0; i < 3; ++i) {
    if (!ctx->ev_used[i]) continue;
    float ms = 0.0f;
    if (cudaEventSynchronize(ctx->ev[i][1]) == cudaSuccess &&
        cudaEventElapsedTime(&ms, ctx->ev[i][0], ctx->ev[i][1]) == cudaSuccess)
      *slots[i] = (int64_t)(ms * 1000.0f);
  }
  *out = ctx->tm;
  return WB_OK;
}

int wb_kernel_time_us(const wb_ctx* ctx_c, const char* family, double* total_us, int64_t* launches) {
  wb_ctx* ctx = const_cast<wb_ctx*>(ctx_c);
  if (!ctx || !family) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  if (strcmp(family, "__enable__") == 0) {
    ctx->time_kernels = true;
    return WB_OK;
  }
  if (strcmp(family, "__disable__") == 0) {
    ctx->time_kernels = false;
    return WB_OK;
  }
  cudaStreamSynchronize(ctx->stream);
  resolve_kernel_clocks(ctx);
  if (strcmp(family, "__reset__") == 0) {
    ctx->clocks.clear();
    ctx->tm.n_kernel_launches = 0;
    return WB_OK;
  }
  double tot = 0;
  int64_t n = 0;
  const size_t flen = strlen(family);
  for (auto& kv : ctx->clocks) {
    if (kv.first.compare(0, flen, family) == 0) {   // prefix match: "gemm" covers gemm, gemm_conv, gemm_cross
      tot += kv.second.total_us;
      n += kv.second.launches;
    }
  }
  if (total_us) *total_us = tot;
  if (launches) *launches = n;
  return WB_OK;
}

// ------------------------------------------------------------------------------------------------
// single-op probes over host buffers
int wb_dbg_gemm(wb_ctx* ctx, int M, int N, int K, const uint16_t* a_f16, const uint16_t* w_f16, const float* bias,
                const float* residual, int gelu, float scale, int out_f16, void* out) {
  if (!ctx || !a_f16 || !w_f16 || !out || M < 1 || N < 1 || K < 8 || K % 8 != 0) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  __half *dA = nullptr, *dW = nullptr;
  float *dB = nullptr, *dR = nullptr;
  void* dO = nullptr;
  const size_t osz = (size_t)M * N * (out_f16 ? 2 : 4);
  int rc = WB_OK;
  const char* terr = "";
  cudaError_t e;
  auto cleanup = [&]() {
    cudaFree(dA); cudaFree(dW); cudaFree(dB); cudaFree(dR); cudaFree(dO);
  };
#define DBG_CK(x)                                         \
  do {                                                    \
    e = (x);                                              \
    if (e != cudaSuccess) {                               \
      rc = fail(ctx, WB_ERR_TENSOR_OP, #x, e);            \
      cleanup();                                          \
      return rc;                                          \
    }                                                     \
  } while (0)
  DBG_CK(cudaMalloc(&dA, (size_t)M * K * 2));
  DBG_CK(cudaMalloc(&dW, (size_t)N * K * 2));
  DBG_CK(cudaMalloc(&dO, osz));
  DBG_CK(cudaMemcpy(dA, a_f16, (size_t)M * K * 2, cudaMemcpyHostToDevice));
  DBG_CK(cudaMemcpy(dW, w_f16, (size_t)N * K * 2, cudaMemcpyHostToDevice));
  DBG_CK(cudaMemset(dO, 0, osz));
  if (bias) {
    DBG_CK(cudaMalloc(&dB, (size_t)N * 4));
    DBG_CK(cudaMemcpy(dB, bias, (size_t)N * 4, cudaMemcpyHostToDevice));
  }
  if (residual) {
    DBG_CK(cudaMalloc(&dR, (size_t)M * N * 4));
    DBG_CK(cudaMemcpy(dR, residual, (size_t)M * N * 4, cudaMemcpyHostToDevice));
  }
  Linear l;
  l.w = dW;
  l.N = N;
  l.K = K;
  l.bias = dB;
  CUtensorMap ma;
  if (!make_linear_maps(ctx, l, false) || !tmap_3d_rows(&ma, dA, K, M, 1, K, (uint64_t)M * K, &terr)) {
    if (*terr) fail_msg(ctx, WB_ERR_TENSOR_OP, std::string("galois tensor:'") + terr + "'");
    cleanup();
    return WB_ERR_TENSOR_OP;
  }
  GemmEpilogue ep;
  ep.gelu = gelu;
  ep.scale = scale;
  ep.residual = dR;
  ep.res_ld = N;
  ep.out = dO;
  ep.out_f16 = out_f16;
  ep.out_ld = N;
  CUtensorMap mo, mr;
  if (!tmap_out(&mo, dO, !out_f16, N, M, 1, N, 0, &terr) || (dR && !tmap_out(&mr, dR, true, N, M, 1, N, 0, &terr))) {
    fail_msg(ctx, WB_ERR_TENSOR_OP, std::string("galois tensor:'") + terr + "'");
    cleanup();
    return WB_ERR_TENSOR_OP;
  }
  const char* trace_path = getenv("WB_GEMM_TRACE");   // per-phase clock64() trace of CTA 0 (tools/gemm_trace.py)
  if (trace_path) {
    DBG_CK(cudaMalloc(&g_gemm_trace, 1024 * sizeof(long long)));
    DBG_CK(cudaMemset(g_gemm_trace, 0, 1024 * sizeof(long long)));
  }
  rc = run_gemm(ctx, ma, M, 1, l, ep, "dbg_gemm", &mo, dR ? &mr : nullptr, 0);
  if (trace_path) {
    cudaStreamSynchronize(ctx->stream);
    std::vector<long long> h(1024);
    cudaMemcpy(h.data(), g_gemm_trace, 1024 * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(g_gemm_trace);
    g_gemm_trace = nullptr;
    if (FILE* f = fopen(trace_path, "w")) {
      for (int i = 0; i < 1024; ++i) fprintf(f, "%lld\n", h[i]);
      fclose(f);
    }
  }
  if (rc == WB_OK) {
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, WB_ERR_TENSOR_OP, "gemm kernel", e);
  }
  if (rc == WB_OK) DBG_CK(cudaMemcpy(out, dO, osz, cudaMemcpyDeviceToHost));
  cleanup();
  return rc;
}

int wb_dbg_attention(wb_ctx* ctx, int n_seg, int T, int H, const uint16_t* qkv_f16, uint16_t* out_f16) {
  if (!ctx || !qkv_f16 || !out_f16 || n_seg < 1 || T < 1 || H < 1) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  const int d = H * 64, Tp = (T + 127) / 128 * 128;
  const size_t M = (size_t)n_seg * T;
  __half *dQK = nullptr, *dVt = nullptr, *dO = nullptr;
  int rc = WB_OK;
  cudaError_t e;
  const char* terr = "";
  auto cleanup = [&]() { cudaFree(dQK); cudaFree(dVt); cudaFree(dO); };
  // host repack: Q|K rows [M][2d]; V^T [seg][d][Tp]
  const int VR = ATTN_VT_HEAD_ROWS;
  std::vector<uint16_t> qk(M * 2 * d), vt((size_t)n_seg * H * VR * Tp, 0);
  for (size_t m = 0; m < M; ++m) {
    memcpy(&qk[m * 2 * d], &qkv_f16[m * 3 * d], (size_t)2 * d * 2);
    const size_t seg = m / T, t = m % T;
    for (int c = 0; c < d; ++c) vt[((seg * H + c / 64) * VR + c % 64) * Tp + t] = qkv_f16[m * 3 * d + 2 * d + c];
  }
  for (size_t hb = 0; hb < (size_t)n_seg * H; ++hb)   // the ones row behind every head block
    for (int t = 0; t < Tp; ++t) vt[(hb * VR + 64) * Tp + t] = 0x3C00;
  DBG_CK(cudaMalloc(&dQK, qk.size() * 2));
  DBG_CK(cudaMalloc(&dVt, vt.size() * 2));
  DBG_CK(cudaMalloc(&dO, M * d * 2));
  DBG_CK(cudaMemcpy(dQK, qk.data(), qk.size() * 2, cudaMemcpyHostToDevice));
  DBG_CK(cudaMemcpy(dVt, vt.data(), vt.size() * 2, cudaMemcpyHostToDevice));
  DBG_CK(cudaMemset(dO, 0, M * d * 2));
  AttnProblem ap;
  {
    const uint64_t dims[4] = {64, (uint64_t)2 * H, (uint64_t)T, (uint64_t)n_seg};
    const uint64_t strd[3] = {128, (uint64_t)2 * d * 2, (uint64_t)T * 2 * d * 2};
    const uint32_t box[4] = {64, 1, 128, 1};
    const uint64_t dims2[2] = {(uint64_t)Tp, (uint64_t)n_seg * H * ATTN_VT_HEAD_ROWS};
    const uint64_t strd2[1] = {(uint64_t)Tp * 2};
    const uint32_t box2[2] = {64, (uint32_t)ATTN_VT_HEAD_ROWS};
    if (!make_tmap_f16(&ap.qk_map, dQK, 4, dims, strd, box, &terr) ||
        !make_tmap_f16(&ap.vt_map, dVt, 2, dims2, strd2, box2, &terr)) {
      fail_msg(ctx, WB_ERR_TENSOR_OP, std::string("galois tensor:'") + terr + "'");
      cleanup();
      return WB_ERR_TENSOR_OP;
    }
  }
  ap.B = n_seg;
  ap.T = T;
  ap.H = H;
  ap.out = dO;
  ap.scale = 0.125f;
  long long* d_trace = nullptr;
  const char* trace_path = getenv("WB_ATTN_TRACE");   // developer aid: per-phase clock64() trace of CTA (0,0,0)
  if (trace_path) {
    DBG_CK(cudaMalloc(&d_trace, 1024 * sizeof(long long)));
    DBG_CK(cudaMemset(d_trace, 0, 1024 * sizeof(long long)));
    ap.dbg = d_trace;
  }
  {
    LaunchTimer t(ctx, "dbg_attention");
    DBG_CK(run_attention(ap, ctx->stream));
  }
  DBG_CK(cudaStreamSynchronize(ctx->stream));
  if (d_trace) {
    std::vector<long long> h(1024);
    cudaMemcpy(h.data(), d_trace, 1024 * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(d_trace);
    if (FILE* f = fopen(trace_path, "w")) {
      for (int i = 0; i < 1024; ++i) fprintf(f, "%lld\n", h[i]);
      fclose(f);
    }
  }
  DBG_CK(cudaMemcpy(out_f16, dO, M * d * 2, cudaMemcpyDeviceToHost));
  cleanup();
  return rc;
}

// guard zones around every device buffer (see dev_alloc): returns the number of guard zones that no longer hold
// their fill pattern (0 = no kernel wrote outside its buffers), -1 if the context was created without guards
int wb_dbg_canary_check(wb_ctx* ctx) {
  if (!ctx) return WB_ERR_UNEXPECTED;
  if (!ctx->canary) return -1;
  cudaSetDevice(ctx->device);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return WB_ERR_TENSOR_OP;
  std::vector<uint8_t> h(2 * WB_GUARD);
  int bad = 0;
  for (const auto& g : ctx->guarded) {
    if (cudaMemcpy(h.data(), g.user - WB_GUARD, WB_GUARD, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(h.data() + WB_GUARD, g.user + g.bytes, WB_GUARD, cudaMemcpyDeviceToHost) != cudaSuccess)
      return WB_ERR_TENSOR_OP;
    bool front_ok = true, back_ok = true;
    for (size_t i = 0; i < WB_GUARD; ++i) {
      front_ok = front_ok && h[i] == 0xA5;
      back_ok = back_ok && h[WB_GUARD + i] == 0xA5;
    }
    bad += (front_ok ? 0 : 1) + (back_ok ? 0 : 1);
  }
  return bad;
}

int wb_dbg_layernorm(wb_ctx* ctx, int rows, int d, const float* x, const float* w, const float* b, uint16_t* out_f16) {
  if (!ctx || !x || !w || !b || !out_f16) return WB_ERR_UNEXPECTED;
  cudaSetDevice(ctx->device);
  float *dx = nullptr, *dw = nullptr, *db = nullptr;
  __half* dO = nullptr;
  int rc = WB_OK;
  cudaError_t e;
  auto cleanup = [&]() { cudaFree(dx); cudaFree(dw); cudaFree(db); cudaFree(dO); };
  DBG_CK(cudaMalloc(&dx, (size_t)rows * d * 4));
  DBG_CK(cudaMalloc(&dw, (size_t)d * 4));
  DBG_CK(cudaMalloc(&db, (size_t)d * 4));
  DBG_CK(cudaMalloc(&dO, (size_t)rows * d * 2));
  DBG_CK(cudaMemcpy(dx, x, (size_t)rows * d * 4, cudaMemcpyHostToDevice));
  DBG_CK(cudaMemcpy(dw, w, (size_t)d * 4, cudaMemcpyHostToDevice));
  DBG_CK(cudaMemcpy(db, b, (size_t)d * 4, cudaMemcpyHostToDevice));
  DBG_CK(launch_layernorm(dx, dw, db, rows, d, dO, nullptr, ctx->stream));
  DBG_CK(cudaStreamSynchronize(ctx->stream));
  DBG_CK(cudaMemcpy(out_f16, dO, (size_t)rows * d * 2, cudaMemcpyDeviceToHost));
  cleanup();
  return rc;
}
#undef DBG_CK

}  // extern "C"
