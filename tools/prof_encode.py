"""Run mel + encode twice on one synthetic batch (for ncu): python tools/prof_encode.py <arch> <batch>"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
from whisper_rs_b200 import api
arch = sys.argv[1] if len(sys.argv) > 1 else "base"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
os.makedirs("/tmp/wb_models", exist_ok=True)
path = f"/tmp/wb_models/ggml-{arch}.bin"
if not os.path.exists(path):
    pkg.ggml_file.make_model(path, arch)
ctx = api.WhisperContext.new(path, max_segments=B, max_clips=B, max_clip_samples=480000, decode_capacity=False)
pcm = pkg.synth.make_clips(B, first_seg=0, n_samples=480000)
for _ in range(2):
    api.whisper_pcm_to_mel(ctx, pcm)
    api.whisper_encode(ctx, 1, [0] * B, clip_ids=list(range(B)))
d = ctx.encoder_digest(B)
print("ok", float(d[0]))
