"""BASELINE.json configs[2] and configs[4] at full size on one GPU (this rank's share), timed per stage.

  configs[2]  whisper small, batch 32: mel + encoder + greedy decode to 224 tokens
  configs[4]  whisper large-v3 (128 mel bins, 51866 tokens), one clip of N x 30 s split across G GPUs:
              this process runs rank `--rank` of `--world` (default: rank 0 of 8 -> 15 of the 120 windows) through
              pipeline.transcribe_clip (two-phase mel with the whole-clip maximum, encode, greedy decode)
Prints one JSON line per config.   python tools/run_configs.py [small] [large-v3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
from whisper_rs_b200 import api, pipeline  # noqa: E402


def model(arch):
    os.makedirs("/tmp/wb_models", exist_ok=True)
    p = f"/tmp/wb_models/ggml-{arch}.bin"
    if not os.path.exists(p):
        t0 = time.time()
        pkg.ggml_file.make_model(p, arch)
        print(f"[{arch}] model file written in {time.time() - t0:.1f} s ({os.path.getsize(p) / 2**20:.0f} MiB)", file=sys.stderr)
    return p


def config3(reps=3):
    B, n_new = 32, 224
    ctx = api.WhisperContext.new(model("small"), max_segments=B, max_clips=B, max_clip_samples=480000)
    clips = pkg.synth.make_clips(B, first_seg=500)
    out = []
    for _ in range(reps):
        t0 = time.perf_counter()
        api.whisper_pcm_to_mel(ctx, clips)
        api.whisper_encode(ctx, 1, [0] * B, clip_ids=list(range(B)))
        toks, marg, lens = api.whisper_decode_greedy(ctx, [ctx.token_sot], n_new, n_seqs=B, eot=-1)
        wall = time.perf_counter() - t0
        tm = ctx.timings()
        out.append((wall, tm["t_mel_us"], tm["t_encode_us"], tm["t_decode_us"]))
    wall, mel, enc, dec = min(out)
    ctx.close()
    return {"config": "whisper small full transcribe (mel + encoder + greedy decode to 224 tokens), batch 32, 1 B200",
            "wall_s": wall, "t_mel_us": mel, "t_encode_us": enc, "t_decode_us": dec,
            "segments_per_s_end_to_end": B / wall, "encoder_segments_per_s": B / ((mel + enc) * 1e-6),
            "decoder_tokens_per_s": B * n_new / (dec * 1e-6), "all_lengths_224": bool((lens == n_new).all())}


def config5(rank, world, n_windows, max_new):
    hp = pkg.ggml_file.ARCHS["large-v3"]
    n = n_windows * 480000
    part = pipeline.ClipPart(n, rank, world)
    span = pkg.synth.make_long_clip(len(part.windows), first_seg=part.windows[0])[: part.hi - part.lo]
    if span.size < part.hi - part.lo:                   # the 240-sample halo of the next rank's first window
        span = np.concatenate([span, pkg.synth.make_segment(part.windows[-1] + 1)[: part.hi - part.lo - span.size]])
    B = len(part.windows)
    t0 = time.perf_counter()
    ctx = api.WhisperContext.new(model("large-v3"), max_segments=B, max_clips=1, max_clip_samples=span.size)
    t_load = time.perf_counter() - t0
    assert ctx.n_mels == 128 and ctx.n_vocab == hp.n_vocab
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        wins, toks, lens, marg = pipeline.transcribe_clip(ctx, span, rank=rank, world=world, reduce_max=lambda x: x,
                                                          max_new=max_new, eot=-1, pcm_is_local_span=True, n_samples_total=n)
        wall = time.perf_counter() - t0
        tm = ctx.timings()
        best = min(best, (wall, tm["t_mel_us"], tm["t_encode_us"], tm["t_decode_us"])) if best else (wall, tm["t_mel_us"], tm["t_encode_us"], tm["t_decode_us"])
    wall, mel, enc, dec = best
    ctx.close()
    return {"config": f"whisper large-v3, one {n_windows} x 30 s clip over {world} GPUs: rank {rank}'s {B} windows, "
                      f"mel (two-phase, whole-clip max) + encoder + greedy decode to {max_new} tokens",
            "load_s": t_load, "wall_s": wall, "t_mel_us": mel, "t_encode_us": enc, "t_decode_us": dec,
            "windows": wins, "segments_per_s_end_to_end": B / wall, "encoder_segments_per_s": B / ((mel + enc) * 1e-6),
            "decoder_tokens_per_s": B * max_new / (dec * 1e-6), "text_of_window0": ctx_text(toks[0])}


def ctx_text(t):
    return [int(x) for x in t[:8]]


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["small", "large-v3"])
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--windows", type=int, default=120)
    ap.add_argument("--max-new", type=int, default=224)
    ap.add_argument("--only-load", default=None, help="create one context of this architecture (15 segments) and report the time")
    a = ap.parse_args()
    if a.only_load:
        path = model(a.only_load)
        for rep in range(2):
            t0 = time.perf_counter()
            ctx = api.WhisperContext.new(path, max_segments=15, max_clips=1, max_clip_samples=15 * 480000 + 240)
            dt = time.perf_counter() - t0
            print(json.dumps({"arch": a.only_load, "load_s": dt, "t_load_us": ctx.timings()["t_load_us"], "rep": rep}), flush=True)
            ctx.close()
        sys.exit(0)
    if "small" in a.which:
        print(json.dumps(config3()), flush=True)
    if "large-v3" in a.which:
        print(json.dumps(config5(a.rank, a.world, a.windows, a.max_new)), flush=True)
