"""clock64() trace of one CTA of the pair GEMM: WB_GEMM_TRACE=<file> python tools/gemm_trace.py M N K [gelu] [res]
Runs wb_dbg_gemm (f16 out, or f32 out + residual with `res`) and prints per-tile phase times of the
MMA warp and of epilogue warp 4 of CTA 0."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
from whisper_rs_b200 import api
M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
gelu, res = "gelu" in sys.argv, "res" in sys.argv
path = "/tmp/wb_models/ggml-micro.bin"
os.makedirs("/tmp/wb_models", exist_ok=True)
if not os.path.exists(path):
    pkg.ggml_file.make_model(path, "micro")
ctx = api.WhisperContext.new(path, max_segments=1, decode_capacity=False)
rng = np.random.default_rng(0)
a = (rng.standard_normal((M, K)) * 0.5).astype(np.float16)
w = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float16)
bias = rng.standard_normal(N).astype(np.float32)
r = rng.standard_normal((M, N)).astype(np.float32) if res else None
trace = os.environ.get("WB_GEMM_TRACE")
os.environ.pop("WB_GEMM_TRACE", None)
api.dbg_gemm(ctx, a, w, bias=bias, residual=r, gelu=gelu, out_f16=not res)   # warm-up without trace
os.environ["WB_GEMM_TRACE"] = trace
api.dbg_gemm(ctx, a, w, bias=bias, residual=r, gelu=gelu, out_f16=not res)
v = [int(x) for x in open(trace).read().split()]
t0 = min(x for x in v if x > 0)
print(f"M={M} N={N} K={K} gelu={gelu} res={res}; cycles relative to the first stamp")
print("MMA warp: tile | wait_empty_start  acquired  last_commit | wait  issue(all k-blocks)")
for t in range(16):
    a_ = v[t * 4: t * 4 + 3]
    if not a_[0]: break
    print(f"{t:3d} | " + " ".join(f"{x - t0:8d}" for x in a_) + f" | {a_[1]-a_[0]:6d} {a_[2]-a_[1]:6d}")
print("epilogue warp 4: tile | wait_full_start acquired | per chunk: ld_done box_written fenced store_issued ...")
for t in range(16):
    b = v[256 + t * 16: 256 + t * 16 + 16]
    if not b[0]: break
    s = f"{t:3d} | {b[0]-t0:8d} {b[1]-t0:8d} (wait {b[1]-b[0]:6d}) |"
    prev = b[1]
    for k in range(3):
        c = b[2 + 4 * k: 6 + 4 * k]
        if not c[0]: break
        s += f"  ld+{c[0]-prev} math+{c[1]-c[0]} fence+{c[2]-c[1]} store+{c[3]-c[2]}"
        prev = c[3]
    print(s)
