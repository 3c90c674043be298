"""Encoder parity (rel-L2 of ln_post output and of the last layer's cross K/V against the CPU oracle) with the
LayerNorm fold on and off: python tools/ln_fold_check.py [arch ...]   (WB_LN_FOLD is read at context creation)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
from whisper_rs_b200 import api  # noqa: E402
from oracle import pyoracle  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


for arch in (sys.argv[1:] or ["tiny", "base"]):
    path = f"/tmp/wb_models/ggml-{arch}.bin"
    os.makedirs("/tmp/wb_models", exist_ok=True)
    if not os.path.exists(path):
        pkg.ggml_file.make_model(path, arch)
    pcm = pkg.synth.make_segment(11, 480000)
    orc = pyoracle.Oracle(path)
    orc.pcm_to_mel(pcm)
    ref = orc.encode(0)
    rk, rv = orc.cross_kv(orc.n_text_layer - 1)
    for fold in ("1", "0"):
        os.environ["WB_LN_FOLD"] = fold
        ctx = api.WhisperContext.new(path, max_segments=1, max_clips=1, max_clip_samples=480000)
        api.whisper_pcm_to_mel(ctx, pcm)
        api.whisper_encode(ctx, 1, 0)
        k, v = ctx.cross_kv(0, ctx.n_text_layer - 1)
        print(f"{arch} fold={fold}: enc rel-L2 {rel(ctx.encoder_out(0), ref):.3e}  cross-K {rel(k, rk):.3e}  cross-V {rel(v, rv):.3e}", flush=True)
        ctx.close()
