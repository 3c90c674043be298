# sweep the second warpgroup's start offset on the attention probe (4 segments x 8 heads x T=1500)
for ns in 0; do
  WB_ATTN_STAGGER_NS=$ns python - <<PY
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import __graft_entry__ as g
pkg = g.load_package()
from whisper_rs_b200 import api
path = "/tmp/wb_models/ggml-base.bin"
os.makedirs("/tmp/wb_models", exist_ok=True)
if not os.path.exists(path): pkg.ggml_file.make_model(path, "base")
B = 16
ctx = api.WhisperContext.new(path, max_segments=B, max_clips=B, max_clip_samples=480000, decode_capacity=False)
pcm = pkg.synth.make_clips(B, first_seg=0, n_samples=480000)
api.whisper_pcm_to_mel(ctx, pcm)
for _ in range(3): api.whisper_encode(ctx, 1, [0]*B, clip_ids=list(range(B)))
ctx.sync()
ctx.kernel_time_us("__enable__"); ctx.kernel_time_us("__reset__")
for _ in range(5): api.whisper_encode(ctx, 1, [0]*B, clip_ids=list(range(B)))
us, n = ctx.kernel_time_us("attention")
print("stagger_ns", os.environ["WB_ATTN_STAGGER_NS"], "attention us/launch %.1f" % (us / n), "TF %.0f" % (4*1500*1500*512*B/ (us/n*1e-6)/1e12))
PY
done
