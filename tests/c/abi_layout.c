/* abi_layout.c -- a C (not C++) consumer of include/whisper_b200.h: proves the header compiles as
 * plain C99 and prints the layout of every struct that crosses the ABI, so that the hand-written
 * ctypes mirror (whisper.rs_b200/cabi.py) and the Rust -sys crate can be checked against what the
 * compiler actually lays out (tests/test_host.py::test_abi_struct_layout_matches_ctypes). */
#include <stddef.h>
#include <stdio.h>

#include "whisper_b200.h"

#define FIELD(T, f) printf(#T "." #f " %zu %zu\n", offsetof(T, f), sizeof(((T*)0)->f))

int main(void) {
  printf("sizeof.wb_config %zu\n", sizeof(wb_config));
  FIELD(wb_config, device);
  FIELD(wb_config, max_segments);
  FIELD(wb_config, max_clips);
  FIELD(wb_config, max_clip_samples);
  FIELD(wb_config, norm_scope);
  FIELD(wb_config, checkpoints);
  FIELD(wb_config, stream);
  FIELD(wb_config, decode_capacity);
  FIELD(wb_config, reserved);
  printf("sizeof.wb_timings %zu\n", sizeof(wb_timings));
  FIELD(wb_timings, t_load_us);
  FIELD(wb_timings, t_mel_us);
  FIELD(wb_timings, t_sample_us);
  FIELD(wb_timings, t_encode_us);
  FIELD(wb_timings, t_decode_us);
  FIELD(wb_timings, n_mel_calls);
  FIELD(wb_timings, n_encode_calls);
  FIELD(wb_timings, n_decode_calls);
  FIELD(wb_timings, n_kernel_launches);
  printf("enum.WB_ERR_TENSOR_OP %d\n", (int)WB_ERR_TENSOR_OP);
  printf("enum.WB_NORM_SEGMENT %d\n", (int)WB_NORM_SEGMENT);
  printf("enum.WB_STAGE_CROSS_V %d\n", (int)WB_STAGE_CROSS_V);
  return 0;
}
