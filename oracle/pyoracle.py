"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (see oracle/oracle.h for the parity status of each stage).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib: Optional[C.CDLL] = None

STAGE_MEL, STAGE_CONV1, STAGE_CONV2_POS, STAGE_LAYER, STAGE_LN_POST, STAGE_CROSS_K, STAGE_CROSS_V = range(7)
OPT_ACT_F16_ROUND, OPT_GELU_MODE, OPT_SOFTMAX_EXP, OPT_PROB_F16_ROUND = range(4)


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile (gcc only)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE, "-j8"] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i32p, f32p, u16p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_uint16)
        L.orc_ctx_create.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.orc_ctx_free.argtypes = [vp]
        L.orc_ctx_free.restype = None
        L.orc_get_hparams.argtypes = [vp, i32p]
        L.orc_get_special_tokens.argtypes = [vp, i32p]
        L.orc_set_option.argtypes = [vp, C.c_int, C.c_int]
        L.orc_set_audio_ctx.argtypes = [vp, C.c_int]
        L.orc_pcm_to_mel.argtypes = [vp, f32p, C.c_size_t, C.c_int]
        L.orc_mel_dims.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_mel_read.argtypes = [vp, f32p]
        L.orc_mel_set.argtypes = [vp, f32p, C.c_int, C.c_int]
        L.orc_encode.argtypes = [vp, C.c_int, C.c_size_t]
        L.orc_encoder_out_read.argtypes = [vp, f32p]
        L.orc_cross_kv_read.argtypes = [vp, C.c_int, u16p, u16p]
        L.orc_checksum.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.orc_decode.argtypes = [vp, i32p, C.c_int, C.c_int, C.c_int]
        L.orc_logits_read.argtypes = [vp, f32p]
        L.orc_decode_greedy.argtypes = [vp, i32p, C.c_int, C.c_int, C.c_int, C.c_int, i32p, f32p,
                                        C.POINTER(C.c_int)]
        L.orc_fft.argtypes = [f32p, C.c_int, f32p]
        L.orc_fft.restype = None
        L.orc_dft.argtypes = [f32p, C.c_int, f32p]
        L.orc_dft.restype = None
        for fn in (L.orc_f16_round, L.orc_gelu_lut, L.orc_exp_lut):
            fn.argtypes = [C.c_float]
            fn.restype = C.c_float
        L.orc_last_error.restype = C.c_char_p
        _lib = L
    return _lib


class OracleError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


def _f32p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _check(rc: int) -> None:
    if rc != 0:
        raise OracleError(rc, (lib().orc_last_error() or b"").decode(errors="replace"))


class Oracle:
    """One reference context (WhisperContext, src/main.rs:333-363) on the CPU."""

    def __init__(self, model_path: str, n_threads: int = 0):
        self._h = C.c_void_p()
        _check(lib().orc_ctx_create(model_path.encode(), C.byref(self._h)))
        self.n_threads = n_threads or (os.cpu_count() or 1)
        hp = (C.c_int32 * 11)()
        lib().orc_get_hparams(self._h, hp)
        (self.n_vocab, self.n_audio_ctx, self.n_audio_state, self.n_audio_head, self.n_audio_layer,
         self.n_text_ctx, self.n_text_state, self.n_text_head, self.n_text_layer, self.n_mels,
         self.f16) = list(hp)
        st = (C.c_int32 * 8)()
        lib().orc_get_special_tokens(self._h, st)
        (self.token_eot, self.token_sot, self.token_prev, self.token_solm, self.token_not,
         self.token_beg, self.token_translate, self.token_transcribe) = list(st)

    def close(self) -> None:
        if self._h:
            lib().orc_ctx_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_audio_ctx(self, n_ctx: int) -> None:
        """exp_n_audio_ctx (src/main.rs:362, 1803-1807); 0 = the model's n_audio_ctx."""
        _check(lib().orc_set_audio_ctx(self._h, n_ctx))
        self.audio_ctx = n_ctx or self.n_audio_ctx

    def set_option(self, opt: int, value: int) -> None:
        _check(lib().orc_set_option(self._h, opt, value))

    # whisper_pcm_to_mel (src/main.rs:1681); reference thread count is 4 (1698)
    def pcm_to_mel(self, pcm: np.ndarray, n_threads: int = 4) -> np.ndarray:
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        _check(lib().orc_pcm_to_mel(self._h, _f32p(pcm), pcm.size, n_threads))
        return self.mel()

    def mel(self) -> np.ndarray:
        nm, nl = C.c_int(), C.c_int()
        lib().orc_mel_dims(self._h, C.byref(nm), C.byref(nl))
        out = np.empty((nm.value, nl.value), dtype=np.float32)
        _check(lib().orc_mel_read(self._h, _f32p(out)))
        return out

    def set_mel(self, mel: np.ndarray) -> None:
        mel = np.ascontiguousarray(mel, dtype=np.float32)
        _check(lib().orc_mel_set(self._h, _f32p(mel), mel.shape[0], mel.shape[1]))

    # whisper_encode (src/main.rs:1799)
    def encode(self, mel_offset: int = 0, n_threads: Optional[int] = None) -> np.ndarray:
        _check(lib().orc_encode(self._h, n_threads or self.n_threads, mel_offset))
        out = np.empty((getattr(self, "audio_ctx", self.n_audio_ctx), self.n_audio_state), dtype=np.float32)
        _check(lib().orc_encoder_out_read(self._h, _f32p(out)))
        return out

    def cross_kv(self, layer: int) -> Tuple[np.ndarray, np.ndarray]:
        k = np.empty((getattr(self, "audio_ctx", self.n_audio_ctx), self.n_text_state), dtype=np.float16)
        v = np.empty_like(k)
        _check(lib().orc_cross_kv_read(self._h, layer, k.ctypes.data_as(C.POINTER(C.c_uint16)),
                                       v.ctypes.data_as(C.POINTER(C.c_uint16))))
        return k, v

    def checksum(self, stage: int, layer: int = 0) -> float:
        v = C.c_double()
        _check(lib().orc_checksum(self._h, stage, layer, C.byref(v)))
        return v.value

    def decode(self, tokens, n_past: int, n_threads: Optional[int] = None) -> np.ndarray:
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        _check(lib().orc_decode(self._h, t.ctypes.data_as(C.POINTER(C.c_int32)), t.size, n_past,
                                n_threads or self.n_threads))
        out = np.empty(self.n_vocab, dtype=np.float32)
        _check(lib().orc_logits_read(self._h, _f32p(out)))
        return out

    def decode_greedy(self, prompt, max_new: int, eot: Optional[int] = None,
                      n_threads: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
        p = np.ascontiguousarray(prompt, dtype=np.int32)
        toks = np.zeros(max_new, dtype=np.int32)
        marg = np.zeros(max_new, dtype=np.float32)
        n = C.c_int()
        _check(lib().orc_decode_greedy(self._h, p.ctypes.data_as(C.POINTER(C.c_int32)), p.size, max_new,
                                       self.token_eot if eot is None else eot,
                                       n_threads or self.n_threads,
                                       toks.ctypes.data_as(C.POINTER(C.c_int32)), _f32p(marg), C.byref(n)))
        return toks[:n.value].copy(), marg[:n.value].copy()


def fft(x: np.ndarray) -> np.ndarray:
    """The reference's fft (src/main.rs:1505-1551): returns complex64 [n]."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(2 * x.size, dtype=np.float32)
    lib().orc_fft(_f32p(x), x.size, _f32p(out))
    return out[0::2] + 1j * out[1::2]


def dft(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(2 * x.size, dtype=np.float32)
    lib().orc_dft(_f32p(x), x.size, _f32p(out))
    return out[0::2] + 1j * out[1::2]
