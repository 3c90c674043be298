"""Run the fused attention kernel alone (for ncu): 4 segments x 8 heads x T=1500."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
from whisper_rs_b200 import api
path = "/tmp/wb_models/ggml-micro.bin"
os.makedirs("/tmp/wb_models", exist_ok=True)
if not os.path.exists(path):
    pkg.ggml_file.make_model(path, "micro")
ctx = api.WhisperContext.new(path, max_segments=1, decode_capacity=False)
rng = np.random.default_rng(0)
n_seg, T, H = 4, 1500, 8
qkv = rng.standard_normal((n_seg * T, 3 * H * 64)).astype(np.float16)
for _ in range(3):
    out = api.dbg_attention(ctx, qkv, n_seg, T, H)
print("ok", float(np.abs(out.astype(np.float32)).mean()))
