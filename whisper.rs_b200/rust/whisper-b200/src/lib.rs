//! Safe wrapper over libwhisper_b200 carrying the reference's own names and signatures
//! (szuwgh/whisper.rs src/main.rs): `WhisperContext::new` (366), `whisper_pcm_to_mel` (1681),
//! `whisper_encode` (1799) and the `whisper_decode` the reference declares state for (351-352)
//! but never implements.  A maintainer replaces the bodies of those functions in main.rs with
//! calls into this crate; `main` (2065-2075) stays as it is.
//!
//! NOT COMPILED in the build image (no Rust toolchain); kept in step with
//! include/whisper_b200.h by review and by the ctypes mirror's symbol test.
use std::ffi::{CStr, CString};
use std::sync::Arc;

use whisper_b200_sys as sys;

/// src/main.rs:50-72 -- same variants.  The C library formats the reference's Display text
/// (wb_last_error), so each variant carries that finished message.
#[derive(thiserror::Error, Debug)]
pub enum WsError {
    #[error("{0}")]
    Unexpected(String),
    #[error("{0}")]
    UnexpectIO(String),
    #[error("{0}")]
    BadMagic(String),
    #[error("not enough space in the context's memory pool\n")]
    NotEnoughSpace,
    #[error("{0}")]
    UnknownTensor(String),
    #[error("{0}")]
    BadRefTensor(String),
    #[error("{0}")]
    WrongSizeTensor(String),
    #[error("{0}")]
    WrongShapeTensor(String),
    #[error("{0}")]
    WrongBytesTensor(String),
    #[error("{0}")]
    WrongGTensor(String),
}
pub type WsResult<T> = Result<T, WsError>;

fn check(rc: i32, ctx: *const sys::wb_ctx) -> WsResult<()> {
    if rc == sys::WB_OK {
        return Ok(());
    }
    let msg = unsafe {
        let p = sys::wb_last_error(ctx);
        if p.is_null() { String::new() } else { CStr::from_ptr(p).to_string_lossy().into_owned() }
    };
    Err(match rc {
        sys::WB_ERR_IO => WsError::UnexpectIO(msg),
        sys::WB_ERR_BAD_MAGIC => WsError::BadMagic(msg),
        sys::WB_ERR_NOT_ENOUGH_SPACE => WsError::NotEnoughSpace,
        sys::WB_ERR_UNKNOWN_TENSOR => WsError::UnknownTensor(msg),
        sys::WB_ERR_BAD_REF_TENSOR => WsError::BadRefTensor(msg),
        sys::WB_ERR_WRONG_SIZE_TENSOR => WsError::WrongSizeTensor(msg),
        sys::WB_ERR_WRONG_SHAPE_TENSOR => WsError::WrongShapeTensor(msg),
        sys::WB_ERR_WRONG_BYTES_TENSOR => WsError::WrongBytesTensor(msg),
        sys::WB_ERR_TENSOR_OP => WsError::WrongGTensor(msg),
        _ => WsError::Unexpected(msg),
    })
}

/// WhisperHparams (607-619), as carried in the file header.
#[derive(Clone, Copy, Debug, Default)]
pub struct WhisperHparams {
    pub n_vocab: i32,
    pub n_audio_ctx: i32,
    pub n_audio_state: i32,
    pub n_audio_head: i32,
    pub n_audio_layer: i32,
    pub n_text_ctx: i32,
    pub n_text_state: i32,
    pub n_text_head: i32,
    pub n_text_layer: i32,
    pub n_mels: i32,
    pub f16: i32,
}

/// WhisperContext (333-363) resident on one B200.  `&mut self` on every compute call keeps the
/// reference's exclusivity; the handle is Send but not Sync.
pub struct WhisperContext {
    h: *mut sys::wb_ctx,
    pub hparams: WhisperHparams,
    pub logits: Vec<f32>,
}
unsafe impl Send for WhisperContext {}

impl WhisperContext {
    /// src/main.rs:366.  One 30 s window, one clip, decoder enabled: the reference's shape.
    pub fn new(fname: &str) -> WsResult<WhisperContext> {
        Self::with_capacity(fname, 0, 1, 1, 480_000)
    }

    pub fn with_capacity(fname: &str, device: i32, max_segments: i32, max_clips: i32, max_clip_samples: i64)
                         -> WsResult<WhisperContext> {
        let path = CString::new(fname).map_err(|e| WsError::Unexpected(e.to_string()))?;
        let mut cfg = unsafe {
            let mut c = std::mem::zeroed::<sys::wb_config>();
            sys::wb_config_default(&mut c);
            c
        };
        cfg.device = device;
        cfg.max_segments = max_segments;
        cfg.max_clips = max_clips;
        cfg.max_clip_samples = max_clip_samples;
        let mut h: *mut sys::wb_ctx = std::ptr::null_mut();
        check(unsafe { sys::wb_ctx_create(path.as_ptr(), &cfg, &mut h) }, std::ptr::null())?;
        let mut hp = [0i32; 11];
        unsafe { sys::wb_get_hparams(h, hp.as_mut_ptr()) };
        let hparams = WhisperHparams {
            n_vocab: hp[0], n_audio_ctx: hp[1], n_audio_state: hp[2], n_audio_head: hp[3], n_audio_layer: hp[4],
            n_text_ctx: hp[5], n_text_state: hp[6], n_text_head: hp[7], n_text_layer: hp[8], n_mels: hp[9], f16: hp[10],
        };
        Ok(WhisperContext { h, hparams, logits: Vec::new() })
    }

    /// ln_post output `cur` (1980-1984), `[n_ctx][d]` f32: the reference drops it with buf_compute.
    pub fn encoder_out(&mut self, seg: i32) -> WsResult<Vec<f32>> {
        let n = (self.hparams.n_audio_ctx * self.hparams.n_audio_state) as usize;
        let mut v = vec![0f32; n];
        check(unsafe { sys::wb_encoder_out_read(self.h, seg, v.as_mut_ptr()) }, self.h)?;
        Ok(v)
    }

    pub fn timings(&self) -> sys::wb_timings {
        let mut t = sys::wb_timings::default();
        unsafe { sys::wb_timings_get(self.h, &mut t) };
        t
    }
}

impl Drop for WhisperContext {
    fn drop(&mut self) {
        unsafe { sys::wb_ctx_free(self.h) }
    }
}

/// src/main.rs:1681 -- the mel stays inside the context, as `ctx.mel` does.
pub fn whisper_pcm_to_mel(ctx: &mut WhisperContext, samples: Arc<Vec<f32>>) -> WsResult<()> {
    check(unsafe { sys::wb_pcm_to_mel(ctx.h, samples.as_ptr(), samples.len(), 1) }, ctx.h)
}

/// Phase 1 of `whisper_pcm_to_mel` for one part of a clip that is split across GPUs: log10 mel
/// (src/main.rs:1554-1652) of exactly `n_frames` frames of `samples`; returns this part's maximum
/// (the partial result of the scan at 1655-1662).
pub fn whisper_pcm_to_logmel(ctx: &mut WhisperContext, samples: &[f32], n_frames: usize) -> WsResult<f32> {
    check(unsafe { sys::wb_pcm_to_logmel(ctx.h, samples.as_ptr(), samples.len(), 1, n_frames as i32) }, ctx.h)?;
    let mut mx = 0f32;
    check(unsafe { sys::wb_mel_max_read(ctx.h, &mut mx, 1) }, ctx.h)?;
    Ok(mx)
}

/// Phase 2: `clamp_and_normalize` (src/main.rs:1654-1671) with the maximum over every part of the clip.
pub fn whisper_mel_normalize(ctx: &mut WhisperContext, clip_max: f32) -> WsResult<()> {
    check(unsafe { sys::wb_mel_normalize(ctx.h, &clip_max, 1) }, ctx.h)
}

/// src/main.rs:1799 -- `n_threads` is accepted and ignored, exactly as in the reference.
pub fn whisper_encode(wctx: &mut WhisperContext, _n_threads: usize, mel_offset: usize) -> WsResult<()> {
    let off = [mel_offset];
    let id = [0i32];
    check(unsafe { sys::wb_encode(wctx.h, id.as_ptr(), off.as_ptr(), 1) }, wctx.h)
}

/// The decode step implied by the reference's state: fills `ctx.logits` (351) for the last token.
pub fn whisper_decode(ctx: &mut WhisperContext, tokens: &[i32], n_past: usize, _n_threads: usize) -> WsResult<()> {
    check(unsafe { sys::wb_decode(ctx.h, tokens.as_ptr(), tokens.len() as i32, n_past as i32, 1) }, ctx.h)?;
    ctx.logits.resize(ctx.hparams.n_vocab as usize, 0.0);
    check(unsafe { sys::wb_logits_read(ctx.h, 0, ctx.logits.as_mut_ptr()) }, ctx.h)
}

/// Device-side greedy loop (arg-max over all logits, stop at `eot` or `max_new`).
pub fn whisper_decode_greedy(ctx: &mut WhisperContext, prompt: &[i32], max_new: usize, eot: i32) -> WsResult<Vec<i32>> {
    let mut toks = vec![0i32; max_new];
    let mut len = 0i32;
    check(unsafe {
        sys::wb_decode_greedy(ctx.h, prompt.as_ptr(), prompt.len() as i32, max_new as i32, eot, 1,
                              toks.as_mut_ptr(), std::ptr::null_mut(), &mut len)
    }, ctx.h)?;
    toks.truncate(len as usize);
    Ok(toks)
}
