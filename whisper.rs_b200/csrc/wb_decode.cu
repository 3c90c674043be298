// wb_decode.cu -- decoder step (placeholder until the kernels land)
#include "wb_internal.hpp"

namespace wb {
int decode_setup(wb_ctx* ctx, const ModelFileView& mv) {
  (void)ctx;
  (void)mv;
  return WB_OK;
}
}  // namespace wb

extern "C" {
int wb_decode(wb_ctx* ctx, const int32_t*, int, int, int) {
  return wb::fail_msg(ctx, WB_ERR_UNEXPECTED, "Unexpected: wb_decode not built yet");
}
int wb_logits_read(wb_ctx* ctx, int, float*) {
  return wb::fail_msg(ctx, WB_ERR_UNEXPECTED, "Unexpected: wb_decode not built yet");
}
int wb_decode_greedy(wb_ctx* ctx, const int32_t*, int, int, int, int, int32_t*, float*, int32_t*) {
  return wb::fail_msg(ctx, WB_ERR_UNEXPECTED, "Unexpected: wb_decode not built yet");
}
}
