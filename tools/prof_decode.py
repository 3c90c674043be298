"""mel + encode + a short greedy decode (for ncu): python tools/prof_decode.py <arch> <batch> <n_new>"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
from whisper_rs_b200 import api
arch = sys.argv[1] if len(sys.argv) > 1 else "small"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
n_new = int(sys.argv[3]) if len(sys.argv) > 3 else 6
os.makedirs("/tmp/wb_models", exist_ok=True)
path = f"/tmp/wb_models/ggml-{arch}.bin"
if not os.path.exists(path):
    pkg.ggml_file.make_model(path, arch)
ctx = api.WhisperContext.new(path, max_segments=B, max_clips=B, max_clip_samples=480000, decode_capacity=True)
pcm = pkg.synth.make_clips(B, first_seg=0, n_samples=480000)
api.whisper_pcm_to_mel(ctx, pcm)
api.whisper_encode(ctx, 1, [0] * B, clip_ids=list(range(B)))
toks, _, lens = api.whisper_decode_greedy(ctx, [ctx.token_sot], n_new, n_seqs=B, eot=-1)
print("ok", toks[0][:n_new].tolist(), ctx.timings()["t_decode_us"])
