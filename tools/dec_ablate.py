"""Step time of the greedy decode loop with kernel families left out (WB_DEC_SKIP): where the step's time goes
inside the CUDA graph.  python tools/dec_ablate.py [arch] [batch]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
from whisper_rs_b200 import api
arch = sys.argv[1] if len(sys.argv) > 1 else "small"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
os.makedirs("/tmp/wb_models", exist_ok=True)
path = f"/tmp/wb_models/ggml-{arch}.bin"
if not os.path.exists(path):
    pkg.ggml_file.make_model(path, arch)
clips = pkg.synth.make_clips(B, first_seg=900)
SKIPS = sys.argv[3].split(";") if len(sys.argv) > 3 else ["", "cross", "self", "linear", "logits", "cross,self", "cross,self,linear", "cross,self,linear,logits"]
for skip in SKIPS:
    os.environ["WB_DEC_SKIP"] = skip
    ctx = api.WhisperContext.new(path, max_segments=B, max_clips=B, max_clip_samples=480000)
    api.whisper_pcm_to_mel(ctx, clips)
    api.whisper_encode(ctx, 1, [0] * B, clip_ids=list(range(B)))
    best = 1e9
    for _ in range(3):
        api.whisper_decode_greedy(ctx, [ctx.token_sot], 224, n_seqs=B, eot=-1)
        best = min(best, ctx.timings()["t_decode_us"])
    print(f"{arch} B={B} skip=[{skip:28s}] {best / 224:8.1f} us/step", flush=True)
    ctx.close()
