"""ctypes binding of csrc/libwhisper_b200.so -- the C-ABI declared in include/whisper_b200.h.

There is no CPU fallback: if the library is missing, `lib()` raises and tells the caller to run
`__graft_entry__.build()`; if no sm_100 device is usable, `wb_ctx_create` fails with
WB_ERR_TENSOR_OP and the message says so.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
from typing import List, Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# WB_LIB: an alternative build of the same library (A/B runs of kernel variants on the GPU box)
LIB_PATH = os.environ.get("WB_LIB") or os.path.join(CSRC, "libwhisper_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "whisper_b200.h")
_lib: Optional[C.CDLL] = None

(WB_OK, WB_ERR_UNEXPECTED, WB_ERR_IO, WB_ERR_BAD_MAGIC, WB_ERR_NOT_ENOUGH_SPACE, WB_ERR_UNKNOWN_TENSOR,
 WB_ERR_BAD_REF_TENSOR, WB_ERR_WRONG_SIZE_TENSOR, WB_ERR_WRONG_SHAPE_TENSOR, WB_ERR_WRONG_BYTES_TENSOR,
 WB_ERR_TENSOR_OP) = (0, -1, -2, -3, -4, -5, -6, -7, -8, -9, -10)
(STAGE_MEL, STAGE_CONV1, STAGE_CONV2_POS, STAGE_LAYER, STAGE_LN_POST, STAGE_CROSS_K, STAGE_CROSS_V) = range(7)
NORM_CLIP, NORM_SEGMENT = 0, 1


class WbConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("max_segments", C.c_int32), ("max_clips", C.c_int32),
        ("max_clip_samples", C.c_int64), ("norm_scope", C.c_int32), ("checkpoints", C.c_int32),
        ("stream", C.c_void_p), ("decode_capacity", C.c_int32), ("reserved", C.c_int32 * 7),
    ]


class WbTimings(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "t_load_us", "t_mel_us", "t_sample_us", "t_encode_us", "t_decode_us",
        "n_mel_calls", "n_encode_calls", "n_decode_calls", "n_kernel_launches")]


def declared_symbols() -> List[str]:
    """Every function include/whisper_b200.h declares (used by the export test)."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wb_[a-z0-9_]+)\s*\(", src)))


def build(force: bool = False) -> str:
    """Compile the library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.run(["make", "-C", CSRC, "-j8"] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension is not built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback).")
    L = C.CDLL(LIB_PATH)
    vp, i32p, f32p, u16p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_uint16)
    L.wb_config_default.argtypes = [C.POINTER(WbConfig)]
    L.wb_config_default.restype = None
    L.wb_ctx_create.argtypes = [C.c_char_p, C.POINTER(WbConfig), C.POINTER(vp)]
    L.wb_ctx_free.argtypes = [vp]
    L.wb_ctx_free.restype = None
    L.wb_get_hparams.argtypes = [vp, i32p]
    L.wb_get_special_tokens.argtypes = [vp, i32p]
    L.wb_pcm_to_mel.argtypes = [vp, C.c_void_p, C.c_size_t, C.c_int]
    L.wb_pcm_to_mel_device.argtypes = [vp, C.c_void_p, C.c_size_t, C.c_int]
    L.wb_pcm16_to_mel.argtypes = [vp, C.c_void_p, C.c_size_t, C.c_int]
    L.wb_pcm_prefetch.argtypes = [vp, C.c_void_p, C.c_size_t]
    L.wb_pcm_to_logmel.argtypes = [vp, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
    L.wb_mel_max_read.argtypes = [vp, f32p, C.c_int]
    L.wb_mel_normalize.argtypes = [vp, f32p, C.c_int]
    L.wb_token_text.argtypes = [vp, C.c_int32, C.c_char_p, C.c_size_t]
    L.wb_tokens_to_text.argtypes = [vp, C.POINTER(C.c_int32), C.c_int, C.c_char_p, C.c_size_t]
    L.wb_mel_dims.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.wb_mel_read.argtypes = [vp, C.c_int, f32p, C.c_size_t]
    L.wb_mel_write.argtypes = [vp, f32p, C.c_int, C.c_int, C.c_int]
    L.wb_encode.argtypes = [vp, i32p, C.POINTER(C.c_size_t), C.c_int]
    L.wb_set_audio_ctx.argtypes = [vp, C.c_int]
    L.wb_encoder_out_read.argtypes = [vp, C.c_int, f32p]
    L.wb_cross_kv_read.argtypes = [vp, C.c_int, C.c_int, u16p, u16p]
    L.wb_checksum.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.wb_encoder_digest.argtypes = [vp, C.POINTER(C.c_double), C.c_int]
    L.wb_encoder_digest_async.argtypes = [vp, C.c_void_p, C.c_int]
    L.wb_wait.argtypes = [vp, C.c_int]
    L.wb_decode.argtypes = [vp, i32p, C.c_int, C.c_int, C.c_int]
    L.wb_logits_read.argtypes = [vp, C.c_int, f32p]
    L.wb_decode_greedy.argtypes = [vp, i32p, C.c_int, C.c_int, C.c_int, C.c_int, i32p, f32p, i32p]
    L.wb_sync.argtypes = [vp]
    L.wb_timings_get.argtypes = [vp, C.POINTER(WbTimings)]
    L.wb_last_error.argtypes = [vp]
    L.wb_last_error.restype = C.c_char_p
    L.wb_version.restype = C.c_char_p
    L.wb_dbg_gemm.argtypes = [vp, C.c_int, C.c_int, C.c_int, u16p, u16p, f32p, f32p, C.c_int, C.c_float,
                              C.c_int, C.c_void_p]
    L.wb_dbg_attention.argtypes = [vp, C.c_int, C.c_int, C.c_int, u16p, u16p]
    L.wb_dbg_layernorm.argtypes = [vp, C.c_int, C.c_int, f32p, f32p, f32p, u16p]
    L.wb_dbg_canary_check.argtypes = [vp]
    L.wb_kernel_time_us.argtypes = [vp, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    _lib = L
    return L
