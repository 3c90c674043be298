"""Greedy decode throughput of one configuration under the current WB_DEC_GROUPS setting (read once per process):
   WB_DEC_GROUPS=4 python tools/dec_groups.py <arch> <batch> <n_new> [nc]     ("nc": non-collapsing decoder weights)
Prints one JSON line: tokens/s, ms per step, a digest of the token ids (to compare settings)."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
from whisper_rs_b200 import api  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "small"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
n_new = int(sys.argv[3]) if len(sys.argv) > 3 else 224
nc = len(sys.argv) > 4 and sys.argv[4] == "nc"
os.makedirs("/tmp/wb_models", exist_ok=True)
path = f"/tmp/wb_models/ggml-{arch}{'-nc' if nc else ''}.bin"
if not os.path.exists(path):
    if nc:
        hp = pkg.ggml_file.ARCHS[arch]
        t = {}
        for name, a in pkg.ggml_file.random_tensors(hp, pkg.ggml_file.arch_seed(arch)):
            if name == "decoder.token_embedding.weight":
                a = (a.astype(np.float32) * 12.0).astype(a.dtype)
            elif name == "decoder.positional_embedding":
                a = (a * 400.0).astype(np.float32)
            t[name] = a
        pkg.ggml_file.write_model(path, hp, 0, tensors=t)
    else:
        pkg.ggml_file.make_model(path, arch)
ctx = api.WhisperContext.new(path, max_segments=B, max_clips=B, max_clip_samples=480000, decode_capacity=True)
pcm = pkg.synth.make_clips(B, first_seg=0, n_samples=480000)
api.whisper_pcm_to_mel(ctx, pcm)
api.whisper_encode(ctx, 1, [0] * B, clip_ids=list(range(B)))
best, toks = None, None
for rep in range(3):
    toks, marg, lens = api.whisper_decode_greedy(ctx, [ctx.token_sot], n_new, n_seqs=B, eot=-1)
    t = ctx.timings()["t_decode_us"] * 1e-6
    best = t if best is None or (rep > 0 and t < best) else best
print(json.dumps({"arch": arch, "B": B, "n_new": n_new, "groups_env": os.environ.get("WB_DEC_GROUPS"), "tokens_per_s": B * n_new / best,
                  "ms_per_step": best / n_new * 1e3, "distinct_ids_seq0": len(set(toks[0].tolist())),
                  "ids_sha": hashlib.sha1(toks.tobytes()).hexdigest()[:12], "min_margin": float(marg.min())}), flush=True)
ctx.close()
