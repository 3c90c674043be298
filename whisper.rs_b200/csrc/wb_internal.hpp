// wb_internal.hpp -- helpers shared by the C-ABI translation units.
#pragma once

#include <stdarg.h>

#include "wb_ctx.hpp"

namespace wb {

int fail(wb_ctx* ctx, int code, const char* what, cudaError_t e);
int fail_msg(wb_ctx* ctx, int code, const std::string& msg);
void set_global_error(const std::string& msg);

#define WB_CK(expr)                                                         \
  do {                                                                      \
    cudaError_t e_ = (expr);                                                \
    if (e_ != cudaSuccess) return wb::fail(ctx, WB_ERR_TENSOR_OP, #expr, e_); \
  } while (0)

// With wb_config.reserved[1] != 0 (or WB_CANARY=1) every device buffer sits between two WB_GUARD-byte guard zones
// filled with 0xA5; wb_dbg_canary_check counts the guard zones a kernel has written into (compute-sanitizer is not
// available on every pool: this catches out-of-bounds WRITES next to any buffer on the real hardware at full speed).
constexpr size_t WB_GUARD = 256;
template <class T>
int dev_alloc(wb_ctx* ctx, T** out, size_t count, bool zero = true) {
  void* p = nullptr;
  const size_t bytes = (count ? count : 1) * sizeof(T);
  const size_t guard = ctx->canary ? WB_GUARD : 0;
  cudaError_t e = cudaMalloc(&p, bytes + 2 * guard);
  if (e != cudaSuccess) return fail(ctx, WB_ERR_NOT_ENOUGH_SPACE, "cudaMalloc", e);
  ctx->allocs.push_back(p);
  uint8_t* user = reinterpret_cast<uint8_t*>(p) + guard;
  if (guard) {
    e = cudaMemsetAsync(p, 0xA5, guard, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(user + bytes, 0xA5, guard, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, WB_ERR_TENSOR_OP, "cudaMemset", e);
    ctx->guarded.push_back({user, bytes});
  }
  if (zero) {
    e = cudaMemsetAsync(user, 0, bytes, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, WB_ERR_TENSOR_OP, "cudaMemset", e);
  }
  *out = reinterpret_cast<T*>(user);
  return WB_OK;
}

// kernel-family device timing: event pairs recorded around launches when ctx->time_kernels
struct LaunchTimer {
  wb_ctx* c;
  const char* fam;
  cudaEvent_t a = nullptr, b = nullptr;
  LaunchTimer(wb_ctx* ctx, const char* family);
  ~LaunchTimer();
};
void resolve_kernel_clocks(wb_ctx* ctx);

bool make_linear_maps(wb_ctx* ctx, Linear& l, bool want_a_map);
bool tmap_2d_rows(CUtensorMap* m, const void* base, uint64_t K, uint64_t rows, uint64_t ld_elems, uint32_t box_rows,
                  const char** err);
bool tmap_3d_rows(CUtensorMap* m, const void* base, uint64_t K, uint64_t rows, uint64_t batch, uint64_t ld_elems,
                  uint64_t bstride_elems, const char** err);

// run one Linear as C = A * W^T with the given A map / epilogue
// out_map (+ res_map for an f32 residual) given: the CTA-pair kernel (gemm2.cu) runs it when the
// shape is eligible; otherwise the single-CTA kernel with the pointers in `epi`
int run_gemm(wb_ctx* ctx, const CUtensorMap& a_map, int M_rows, int batch, const Linear& l, GemmEpilogue epi,
             const char* family = "gemm", const CUtensorMap* out_map = nullptr, const CUtensorMap* res_map = nullptr,
             int res_bcast = 0);
bool tmap_out(CUtensorMap* m, const void* base, bool f32, uint64_t N, uint64_t rows, uint64_t batch, uint64_t ld_elems,
              uint64_t bstride_elems, const char** err);

// weight upload helpers (wb_api.cu)
const HostTensor* find(const ModelFileView& mv, const std::string& n);
void to_f16_host(const HostTensor& t, std::vector<__half>& out);
int upload_f32(wb_ctx* ctx, const ModelFileView& mv, const std::string& name, const float** out);
int upload_f16_vec(wb_ctx* ctx, const std::vector<__half>& h, __half** out);
int upload_f32_vec(wb_ctx* ctx, const std::vector<float>& h, const float** out);
int upload_linear(wb_ctx* ctx, const ModelFileView& mv, const std::string& wname, const std::string& bname, Linear& l,
                  bool want_a_map);
struct CatPart {
  std::string w, b;   // b may be empty (no bias, e.g. the key projections: src/main.rs:675, 704, 718)
  float scale;
};
int upload_cat(wb_ctx* ctx, const ModelFileView& mv, const std::vector<CatPart>& parts, Linear& l, bool want_a_map);
int upload_cat_ln(wb_ctx* ctx, const ModelFileView& mv, const std::vector<CatPart>& parts, const std::string& gamma,
                  const std::string& beta, Linear& l, bool want_a_map);

int decode_setup(wb_ctx* ctx, const ModelFileView& mv);   // wb_decode.cu: decoder weights + KV cache

int mel_enc_ordered_host(float f);
float mel_dec_ordered_host(int i);

}  // namespace wb
