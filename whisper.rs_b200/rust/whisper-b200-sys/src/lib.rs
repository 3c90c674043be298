//! Raw bindings: one `extern "C"` item per declaration of include/whisper_b200.h, in the same
//! order.  NOT COMPILED in the build image (no Rust toolchain there); the same symbols are
//! exercised through ctypes (whisper.rs_b200/cabi.py) and checked by
//! tests/test_host.py::test_cabi_exports_every_declared_symbol; a compiled C consumer of the same header
//! (tests/c/main_replay.c) replays the reference's `fn main` against the library on the GPU box.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct wb_ctx {
    _private: [u8; 0],
}

pub const WB_OK: c_int = 0;
pub const WB_ERR_UNEXPECTED: c_int = -1;
pub const WB_ERR_IO: c_int = -2;
pub const WB_ERR_BAD_MAGIC: c_int = -3;
pub const WB_ERR_NOT_ENOUGH_SPACE: c_int = -4;
pub const WB_ERR_UNKNOWN_TENSOR: c_int = -5;
pub const WB_ERR_BAD_REF_TENSOR: c_int = -6;
pub const WB_ERR_WRONG_SIZE_TENSOR: c_int = -7;
pub const WB_ERR_WRONG_SHAPE_TENSOR: c_int = -8;
pub const WB_ERR_WRONG_BYTES_TENSOR: c_int = -9;
pub const WB_ERR_TENSOR_OP: c_int = -10;

pub const WB_NORM_CLIP: i32 = 0;
pub const WB_NORM_SEGMENT: i32 = 1;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct wb_config {
    pub device: i32,
    pub max_segments: i32,
    pub max_clips: i32,
    pub max_clip_samples: i64,
    pub norm_scope: i32,
    pub checkpoints: i32,
    pub stream: *mut c_void,
    pub decode_capacity: i32,
    pub reserved: [i32; 7],
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct wb_timings {
    pub t_load_us: i64,
    pub t_mel_us: i64,
    pub t_sample_us: i64,
    pub t_encode_us: i64,
    pub t_decode_us: i64,
    pub n_mel_calls: i64,
    pub n_encode_calls: i64,
    pub n_decode_calls: i64,
    pub n_kernel_launches: i64,
}

extern "C" {
    pub fn wb_config_default(cfg: *mut wb_config);
    pub fn wb_ctx_create(model_path: *const c_char, cfg: *const wb_config, out: *mut *mut wb_ctx) -> c_int;
    pub fn wb_ctx_free(ctx: *mut wb_ctx);
    pub fn wb_get_hparams(ctx: *const wb_ctx, out: *mut i32) -> c_int;
    pub fn wb_get_special_tokens(ctx: *const wb_ctx, out: *mut i32) -> c_int;
    pub fn wb_pcm_to_mel(ctx: *mut wb_ctx, pcm: *const f32, n_samples: usize, n_clips: c_int) -> c_int;
    pub fn wb_pcm_to_mel_device(ctx: *mut wb_ctx, pcm_dev: *const f32, n_samples: usize, n_clips: c_int) -> c_int;
    pub fn wb_pcm16_to_mel(ctx: *mut wb_ctx, pcm: *const i16, n_samples: usize, n_clips: c_int) -> c_int;
    pub fn wb_pcm_prefetch(ctx: *mut wb_ctx, pcm: *const c_void, n_bytes: usize) -> c_int;
    pub fn wb_pcm_to_logmel(ctx: *mut wb_ctx, pcm: *const f32, n_samples: usize, n_clips: c_int, n_frames: c_int) -> c_int;
    pub fn wb_mel_max_read(ctx: *mut wb_ctx, out: *mut f32, n_clips: c_int) -> c_int;
    pub fn wb_mel_normalize(ctx: *mut wb_ctx, clip_max: *const f32, n_clips: c_int) -> c_int;
    pub fn wb_mel_dims(ctx: *const wb_ctx, n_mel: *mut c_int, n_len: *mut c_int, n_clips: *mut c_int) -> c_int;
    pub fn wb_mel_read(ctx: *mut wb_ctx, clip: c_int, out: *mut f32, cap_floats: usize) -> c_int;
    pub fn wb_mel_write(ctx: *mut wb_ctx, mel: *const f32, n_mel: c_int, n_len: c_int, n_clips: c_int) -> c_int;
    pub fn wb_encode(ctx: *mut wb_ctx, clip_ids: *const i32, mel_offsets: *const usize, n_segments: c_int) -> c_int;
    pub fn wb_set_audio_ctx(ctx: *mut wb_ctx, n_ctx: c_int) -> c_int;
    pub fn wb_encoder_out_read(ctx: *mut wb_ctx, seg: c_int, out: *mut f32) -> c_int;
    pub fn wb_cross_kv_read(ctx: *mut wb_ctx, seg: c_int, layer: c_int, k: *mut u16, v: *mut u16) -> c_int;
    pub fn wb_checksum(ctx: *mut wb_ctx, stage: c_int, layer: c_int, seg: c_int, abs_sum: *mut f64) -> c_int;
    pub fn wb_encoder_digest(ctx: *mut wb_ctx, out: *mut f64, cap: c_int) -> c_int;
    pub fn wb_encoder_digest_async(ctx: *mut wb_ctx, out: *mut f64, cap: c_int) -> c_int;
    pub fn wb_wait(ctx: *mut wb_ctx, ticket: c_int) -> c_int;
    pub fn wb_decode(ctx: *mut wb_ctx, tokens: *const i32, n_tokens: c_int, n_past: c_int, n_seqs: c_int) -> c_int;
    pub fn wb_logits_read(ctx: *mut wb_ctx, seq: c_int, out: *mut f32) -> c_int;
    pub fn wb_decode_greedy(ctx: *mut wb_ctx, prompt: *const i32, n_prompt: c_int, max_new: c_int, eot: c_int,
                            n_seqs: c_int, out_tokens: *mut i32, out_margin: *mut f32, out_len: *mut i32) -> c_int;
    pub fn wb_token_text(ctx: *const wb_ctx, id: i32, out: *mut c_char, cap: usize) -> c_int;
    pub fn wb_tokens_to_text(ctx: *const wb_ctx, ids: *const i32, n: c_int, out: *mut c_char, cap: usize) -> c_int;
    pub fn wb_sync(ctx: *mut wb_ctx) -> c_int;
    pub fn wb_timings_get(ctx: *const wb_ctx, out: *mut wb_timings) -> c_int;
    pub fn wb_last_error(ctx: *const wb_ctx) -> *const c_char;
    pub fn wb_version() -> *const c_char;
    /// guard zones around every device buffer (wb_config.reserved[1] = 1): number of zones written into, -1 = no guards
    pub fn wb_dbg_canary_check(ctx: *mut wb_ctx) -> c_int;
}

// Layout the C compiler gives the two structs (tests/c/abi_layout.c prints it; tests/test_host.py asserts the ctypes
// mirror against it).  A Rust build checks the same numbers at compile time.
const _: () = {
    assert!(std::mem::size_of::<wb_config>() == 72);
    assert!(std::mem::size_of::<wb_timings>() == 72);
};
