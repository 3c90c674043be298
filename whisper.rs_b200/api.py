"""Host-side mirror of the reference's interface for the hot path, over the C-ABI.

The reference's boundary is three crate-private functions sharing one
context, plus the decode call it declares state for but never implements:

    WhisperContext::new(fname) -> WsResult<WhisperContext>              src/main.rs:366
    whisper_pcm_to_mel(ctx, samples) -> WsResult<()>                     src/main.rs:1681
    whisper_encode(ctx, n_threads, mel_offset) -> WsResult<()>           src/main.rs:1799
    whisper_decode(ctx, tokens, n_past, n_threads) -> WsResult<()>       (absent; logits 351-352)

Same names, argument meaning and error behaviour (`WsError` carries the reference's variant
names); results stay inside the context exactly as in the reference and are read back through
explicit accessors.  A Rust `-sys` crate with the same surface ships as source in rust/ (the
image has no Rust toolchain).  Everything below runs on the GPU through libwhisper_b200.so; there
is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import cabi

_VARIANTS = {
    -1: "Unexpected", -2: "UnexpectIO", -3: "BadMagic", -4: "NotEnoughSpace", -5: "UnknownTensor",
    -6: "BadRefTensor", -7: "WrongSizeTensor", -8: "WrongShapeTensor", -9: "WrongBytesTensor",
    -10: "WrongGTensor",
}


class WsError(RuntimeError):
    """src/main.rs:50-72"""

    def __init__(self, code: int, msg: str):
        self.code = code
        self.variant = _VARIANTS.get(code, "Unexpected")
        super().__init__(f"{self.variant}: {msg}")


def _check(rc: int, handle=None) -> None:
    if rc != 0:
        msg = cabi.lib().wb_last_error(handle) or b""
        raise WsError(rc, msg.decode(errors="replace").strip())


def _f32p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class WhisperContext:
    """WhisperContext (src/main.rs:333-363) resident on one B200."""

    def __init__(self, fname: str, max_segments: int = 1, max_clips: int = 1,
                 max_clip_samples: int = 480000, device: int = 0, checkpoints: bool = False,
                 stream: Optional[int] = None, decode_capacity: bool = True, time_kernels: bool = False,
                 norm_scope: int = cabi.NORM_CLIP, canary: bool = False):
        L = cabi.lib()
        cfg = cabi.WbConfig()
        L.wb_config_default(C.byref(cfg))
        cfg.device = device
        cfg.max_segments = max_segments
        cfg.max_clips = max_clips
        cfg.max_clip_samples = max_clip_samples
        cfg.checkpoints = int(checkpoints)
        cfg.norm_scope = int(norm_scope)
        cfg.stream = stream
        cfg.decode_capacity = int(decode_capacity)
        cfg.reserved[0] = int(time_kernels)
        cfg.reserved[1] = int(canary)
        self._h = C.c_void_p()
        _check(L.wb_ctx_create(fname.encode(), C.byref(cfg), C.byref(self._h)))
        hp = (C.c_int32 * 11)()
        L.wb_get_hparams(self._h, hp)
        (self.n_vocab, self.n_audio_ctx, self.n_audio_state, self.n_audio_head, self.n_audio_layer,
         self.n_text_ctx, self.n_text_state, self.n_text_head, self.n_text_layer, self.n_mels,
         self.f16) = list(hp)
        st = (C.c_int32 * 8)()
        L.wb_get_special_tokens(self._h, st)
        (self.token_eot, self.token_sot, self.token_prev, self.token_solm, self.token_not,
         self.token_beg, self.token_translate, self.token_transcribe) = list(st)
        self.max_segments = max_segments
        self.audio_ctx = self.n_audio_ctx        # exp_n_audio_ctx when set (set_audio_ctx), else the model's

    # ---- lifetime
    @classmethod
    def new(cls, fname: str, **kw) -> "WhisperContext":
        return cls(fname, **kw)

    def close(self) -> None:
        if getattr(self, "_h", None):
            cabi.lib().wb_ctx_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_audio_ctx(self, n_ctx: int) -> None:
        """exp_n_audio_ctx (src/main.rs:362, 1803-1807): encode with n_ctx <= n_audio_ctx positions; 0 = the model's."""
        _check(cabi.lib().wb_set_audio_ctx(self._h, n_ctx), self._h)
        self.audio_ctx = n_ctx or self.n_audio_ctx

    def sync(self) -> None:
        _check(cabi.lib().wb_sync(self._h), self._h)

    # ---- accessors (the reference keeps these inside the context)
    def mel(self, clip: int = 0) -> np.ndarray:
        nm, nl, nc = C.c_int(), C.c_int(), C.c_int()
        cabi.lib().wb_mel_dims(self._h, C.byref(nm), C.byref(nl), C.byref(nc))
        out = np.empty((nm.value, nl.value), dtype=np.float32)
        _check(cabi.lib().wb_mel_read(self._h, clip, _f32p(out), out.size), self._h)
        return out

    def set_mel(self, mel: np.ndarray) -> None:
        mel = np.ascontiguousarray(mel, dtype=np.float32)
        if mel.ndim == 2:
            mel = mel[None]
        _check(cabi.lib().wb_mel_write(self._h, _f32p(mel), mel.shape[1], mel.shape[2], mel.shape[0]), self._h)

    def encoder_out(self, seg: int = 0) -> np.ndarray:
        out = np.empty((self.audio_ctx, self.n_audio_state), dtype=np.float32)
        _check(cabi.lib().wb_encoder_out_read(self._h, seg, _f32p(out)), self._h)
        return out

    def cross_kv(self, seg: int, layer: int) -> Tuple[np.ndarray, np.ndarray]:
        k = np.empty((self.audio_ctx, self.n_text_state), dtype=np.float16)
        v = np.empty_like(k)
        u16 = C.POINTER(C.c_uint16)
        _check(cabi.lib().wb_cross_kv_read(self._h, seg, layer, k.ctypes.data_as(u16), v.ctypes.data_as(u16)), self._h)
        return k, v

    def checksum(self, stage: int, layer: int = 0, seg: int = 0) -> float:
        v = C.c_double()
        _check(cabi.lib().wb_checksum(self._h, stage, layer, seg, C.byref(v)), self._h)
        return v.value

    def encoder_digest(self, n_seg: int) -> np.ndarray:
        out = np.zeros(n_seg, dtype=np.float64)
        _check(cabi.lib().wb_encoder_digest(self._h, out.ctypes.data_as(C.POINTER(C.c_double)), n_seg), self._h)
        return out

    def logits(self, seq: int = 0) -> np.ndarray:
        out = np.empty(self.n_vocab, dtype=np.float32)
        _check(cabi.lib().wb_logits_read(self._h, seq, _f32p(out)), self._h)
        return out

    def token_text(self, token_id: int) -> bytes:
        """WhisperVocab::id_to_token (src/main.rs:544)."""
        buf = C.create_string_buffer(256)
        n = cabi.lib().wb_token_text(self._h, int(token_id), buf, 256)
        if n < 0:
            raise WsError(n, "token id out of range")
        return buf.raw[:min(n, 255)]

    def tokens_to_text(self, ids) -> bytes:
        a = np.ascontiguousarray(np.asarray(ids, dtype=np.int32))
        n = cabi.lib().wb_tokens_to_text(self._h, a.ctypes.data_as(C.POINTER(C.c_int32)), a.size, None, 0)
        buf = C.create_string_buffer(n + 1)
        cabi.lib().wb_tokens_to_text(self._h, a.ctypes.data_as(C.POINTER(C.c_int32)), a.size, buf, n + 1)
        return buf.raw[:n]

    def canary_check(self) -> int:
        """Guard zones around the handle's device buffers that a kernel has written into (needs canary=True)."""
        return int(cabi.lib().wb_dbg_canary_check(self._h))

    def timings(self) -> dict:
        t = cabi.WbTimings()
        _check(cabi.lib().wb_timings_get(self._h, C.byref(t)), self._h)
        return {n: getattr(t, n) for n, _ in cabi.WbTimings._fields_}

    def kernel_time_us(self, family: str) -> Tuple[float, int]:
        tot, n = C.c_double(), C.c_int64()
        _check(cabi.lib().wb_kernel_time_us(self._h, family.encode(), C.byref(tot), C.byref(n)), self._h)
        return tot.value, n.value


def whisper_pcm_to_mel(ctx: WhisperContext, samples) -> None:
    """src/main.rs:1681.  `samples`: f32 [n] or [n_clips][n] (host numpy), or int16 of the same
    shapes (the reference's wav path, 1673-1679), or a CUDA torch tensor (f32) already in HBM."""
    L = cabi.lib()
    if hasattr(samples, "is_cuda"):
        if not samples.is_cuda:
            samples = samples.numpy()
        else:
            t = samples.contiguous()
            n_clips = 1 if t.dim() == 1 else t.shape[0]
            _check(L.wb_pcm_to_mel_device(ctx._h, t.data_ptr(), t.shape[-1], n_clips), ctx._h)
            return
    a = np.asarray(samples)
    if a.dtype == np.int16:
        a = np.ascontiguousarray(a)
        n_clips = 1 if a.ndim == 1 else a.shape[0]
        _check(L.wb_pcm16_to_mel(ctx._h, a.ctypes.data, a.shape[-1], n_clips), ctx._h)
        return
    a = np.ascontiguousarray(a, dtype=np.float32)
    n_clips = 1 if a.ndim == 1 else a.shape[0]
    _check(L.wb_pcm_to_mel(ctx._h, a.ctypes.data, a.shape[-1], n_clips), ctx._h)


def whisper_pcm_to_mel_ptr(ctx: WhisperContext, host_ptr: int, n_samples: int, n_clips: int) -> None:
    """Same, from a raw host pointer (pinned staging buffers of the bench's end-to-end leg)."""
    _check(cabi.lib().wb_pcm_to_mel(ctx._h, host_ptr, n_samples, n_clips), ctx._h)


def whisper_pcm_prefetch_ptr(ctx: WhisperContext, host_ptr: int, n_bytes: int) -> None:
    """Begin the H2D upload of the NEXT batch (raw host pointer, pinned) on the context's copy stream;
    the next whisper_pcm_to_mel* call with the same pointer uses it instead of copying again."""
    _check(cabi.lib().wb_pcm_prefetch(ctx._h, host_ptr, n_bytes), ctx._h)


def whisper_pcm_to_logmel(ctx: WhisperContext, samples, n_frames: int = 0) -> np.ndarray:
    """Phase 1 of whisper_pcm_to_mel for a clip split across GPUs: log10 mel (src/main.rs:1554-1652) of
    exactly `n_frames` frames (0 = n_samples / 160) of this part of the clip; returns the local per-clip
    maxima (the partial result of the scan at 1655-1662)."""
    a = np.ascontiguousarray(np.asarray(samples), dtype=np.float32)
    n_clips = 1 if a.ndim == 1 else a.shape[0]
    L = cabi.lib()
    _check(L.wb_pcm_to_logmel(ctx._h, a.ctypes.data, a.shape[-1], n_clips, n_frames), ctx._h)
    mx = np.zeros(n_clips, dtype=np.float32)
    _check(L.wb_mel_max_read(ctx._h, _f32p(mx), n_clips), ctx._h)
    return mx


def whisper_mel_normalize(ctx: WhisperContext, clip_max=None) -> None:
    """Phase 2: clamp_and_normalize (src/main.rs:1654-1671) with the whole-clip maxima `clip_max`
    (MAX over every part of the clip); None = this part's own maxima."""
    L = cabi.lib()
    if clip_max is None:
        n = C.c_int()
        L.wb_mel_dims(ctx._h, None, None, C.byref(n))
        _check(L.wb_mel_normalize(ctx._h, None, n.value), ctx._h)
        return
    mx = np.ascontiguousarray(np.atleast_1d(np.asarray(clip_max, dtype=np.float32)))
    _check(L.wb_mel_normalize(ctx._h, _f32p(mx), mx.size), ctx._h)


def encoder_digest_async(ctx: WhisperContext, out_ptr: int, cap: int) -> int:
    """Queue the per-segment digest and its read-back into pinned host memory at `out_ptr` behind the last
    whisper_encode; returns a ticket for `wait` (the caller may submit the next batch first)."""
    t = cabi.lib().wb_encoder_digest_async(ctx._h, out_ptr, cap)
    _check(min(t, 0), ctx._h)
    return t


def wait(ctx: WhisperContext, ticket: int) -> None:
    _check(cabi.lib().wb_wait(ctx._h, ticket), ctx._h)


def whisper_encode(ctx: WhisperContext, n_threads: int = 1, mel_offset=0,
                   clip_ids: Optional[Sequence[int]] = None) -> None:
    """src/main.rs:1799.  `n_threads` is accepted and ignored, as in the reference (1799, 2074).
    `mel_offset` may be one offset (the reference's call) or a sequence -> one batched call over
    segments; `clip_ids[s]` selects the clip each segment's window is cut from."""
    offs = np.atleast_1d(np.asarray(mel_offset, dtype=np.uint64))
    n_seg = offs.size if clip_ids is None else len(clip_ids)
    if clip_ids is not None and offs.size == 1 and n_seg > 1:
        offs = np.repeat(offs, n_seg)
    ids = np.zeros(n_seg, dtype=np.int32) if clip_ids is None else np.asarray(clip_ids, dtype=np.int32)
    offs = np.ascontiguousarray(offs.astype(np.uint64))
    _check(cabi.lib().wb_encode(ctx._h, ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                offs.ctypes.data_as(C.POINTER(C.c_size_t)), n_seg), ctx._h)


def whisper_decode(ctx: WhisperContext, tokens, n_past: int, n_threads: int = 1) -> None:
    """The decode step implied by the reference's state (logits/probs 351-352): `tokens` is
    [n_tokens] (one sequence) or [n_seqs][n_tokens]; logits of the last position stay on the
    device (read with ctx.logits(seq))."""
    t = np.ascontiguousarray(np.atleast_2d(np.asarray(tokens, dtype=np.int32)))
    _check(cabi.lib().wb_decode(ctx._h, t.ctypes.data_as(C.POINTER(C.c_int32)), t.shape[1], n_past, t.shape[0]), ctx._h)


def whisper_decode_greedy(ctx: WhisperContext, prompt, max_new: int, n_seqs: int = 1,
                          eot: Optional[int] = None):
    """Greedy loop on the device: returns (tokens [n_seqs][max_new], margins, lengths)."""
    p = np.ascontiguousarray(np.asarray(prompt, dtype=np.int32))
    toks = np.zeros((n_seqs, max_new), dtype=np.int32)
    marg = np.zeros((n_seqs, max_new), dtype=np.float32)
    lens = np.zeros(n_seqs, dtype=np.int32)
    i32 = C.POINTER(C.c_int32)
    _check(cabi.lib().wb_decode_greedy(ctx._h, p.ctypes.data_as(i32), p.size, max_new,
                                       ctx.token_eot if eot is None else eot, n_seqs,
                                       toks.ctypes.data_as(i32), _f32p(marg), lens.ctypes.data_as(i32)), ctx._h)
    return toks, marg, lens


# ---- single-kernel probes (tests / bench) ---------------------------------------------------------
def dbg_gemm(ctx: WhisperContext, a: np.ndarray, w: np.ndarray, bias=None, residual=None, gelu=False,
             scale: float = 1.0, out_f16: bool = True) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float16)
    w = np.ascontiguousarray(w, dtype=np.float16)
    M, K = a.shape
    N = w.shape[0]
    out = np.empty((M, N), dtype=np.float16 if out_f16 else np.float32)
    u16 = C.POINTER(C.c_uint16)
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    r = None if residual is None else np.ascontiguousarray(residual, dtype=np.float32)
    _check(cabi.lib().wb_dbg_gemm(ctx._h, M, N, K, a.ctypes.data_as(u16), w.ctypes.data_as(u16),
                                  None if b is None else _f32p(b), None if r is None else _f32p(r),
                                  int(gelu), scale, int(out_f16), out.ctypes.data), ctx._h)
    return out


def dbg_attention(ctx: WhisperContext, qkv: np.ndarray, n_seg: int, T: int, H: int) -> np.ndarray:
    qkv = np.ascontiguousarray(qkv, dtype=np.float16)
    out = np.empty((n_seg * T, H * 64), dtype=np.float16)
    u16 = C.POINTER(C.c_uint16)
    _check(cabi.lib().wb_dbg_attention(ctx._h, n_seg, T, H, qkv.ctypes.data_as(u16), out.ctypes.data_as(u16)), ctx._h)
    return out


def dbg_layernorm(ctx: WhisperContext, x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    out = np.empty(x.shape, dtype=np.float16)
    _check(cabi.lib().wb_dbg_layernorm(ctx._h, x.shape[0], x.shape[1], _f32p(x), _f32p(w), _f32p(b),
                                       out.ctypes.data_as(C.POINTER(C.c_uint16))), ctx._h)
    return out
