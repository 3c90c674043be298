"""BASELINE.json configs[4] for real: whisper large-v3 (128 mel bins), ONE synthetic clip of N x 30 s windows split
across the ranks of a torchrun job -- full mel + encode + greedy decode -- and checked against a single-rank run.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
      tools/run_config5.py [--windows 120] [--max-new 224] [--arch large-v3]

Each rank: its contiguous block of windows (pipeline.ClipPart: the PCM span of the block + the 240-sample halo),
wb_pcm_to_logmel -> one 4-byte NCCL MAX all-reduce (the whole-clip maximum of clamp_and_normalize,
src/main.rs:1655-1662: the only coupling) -> wb_mel_normalize -> wb_encode -> wb_decode_greedy; then one NCCL
all_gather of the token ids.  Rank 0 then transcribes the whole clip alone and compares window by window.
Prints one JSON line (rank 0)."""
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
from whisper_rs_b200 import api, pipeline, shard  # noqa: E402


def model(arch, rank, barrier):
    os.makedirs("/tmp/wb_models", exist_ok=True)
    p = f"/tmp/wb_models/ggml-{arch}.bin"
    if rank == 0 and not os.path.exists(p):
        pkg.ggml_file.make_model(p + ".tmp", arch)
        os.replace(p + ".tmp", p)
    barrier()
    return p


def span_of(part):
    """Synthesize exactly the samples [lo, hi) of the long clip (window w = synth segment w, 0.5 s silent tail)."""
    segs = [pkg.synth.make_segment(w, silent_tail_s=0.5) for w in part.windows]
    need = part.hi - part.lo
    have = sum(s.size for s in segs)
    if have < need:                                     # the halo reaches into the next rank's first window
        segs.append(pkg.synth.make_segment(part.windows[-1] + 1, silent_tail_s=0.5)[: need - have])
    return np.concatenate(segs)[:need]


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="large-v3")
    ap.add_argument("--windows", type=int, default=120)
    ap.add_argument("--max-new", type=int, default=224)
    ap.add_argument("--no-single", action="store_true", help="skip the single-rank comparison run")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    path = model(a.arch, rank, barrier)
    n = a.windows * 480000
    part = pipeline.ClipPart(n, rank, world)
    span = span_of(part)
    B = len(part.windows)
    t0 = time.perf_counter()
    ctx = api.WhisperContext.new(path, max_segments=max(B, 1), max_clips=1, max_clip_samples=max(span.size, 480000), device=local)
    t_load = time.perf_counter() - t0
    reduce_max = pipeline.torch_reduce_max(dev)
    stages = None
    for rep in range(2):                                # the second pass replays the captured graphs
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        wins, toks, lens, marg = pipeline.transcribe_clip(ctx, span, rank=rank, world=world, reduce_max=reduce_max,
                                                          max_new=a.max_new, eot=-1, pcm_is_local_span=True, n_samples_total=n)
        torch.cuda.synchronize()
        t_local = time.perf_counter() - t0
        tm = ctx.timings()
        # the final gather: token ids of every window to every rank (NCCL all_gather)
        all_toks = shard.gather_segment_results(toks, wins, a.windows, device=dev if world > 1 else None)
        all_lens = shard.gather_segment_results(lens, wins, a.windows, device=dev if world > 1 else None)
        torch.cuda.synchronize()
        t_total = time.perf_counter() - t0
        tt = torch.tensor([t_local, t_total, tm["t_mel_us"], tm["t_encode_us"], tm["t_decode_us"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        stages = [float(x) for x in tt.tolist()]
    out = None
    if rank == 0:
        out = {"config": f"whisper {a.arch}: one {a.windows} x 30 s clip over {world} GPU(s), mel (two-phase, whole-clip max by NCCL "
                         f"MAX all-reduce) + encoder + greedy decode to {a.max_new} tokens + NCCL all_gather of the ids",
               "world": world, "windows_per_rank": B, "load_s": t_load, "wall_s_max_over_ranks": stages[1],
               "local_pipeline_s": stages[0], "t_mel_us": stages[2], "t_encode_us": stages[3], "t_decode_us": stages[4],
               "segments_per_s_end_to_end": a.windows / stages[1],
               "decoder_tokens_per_s": a.windows * a.max_new / (stages[4] * 1e-6) if stages[4] else None,
               "all_lengths": int(all_lens.min()), "first_tokens_window0": [int(x) for x in all_toks[0][:8]]}
    ctx.close()
    if rank == 0 and world > 1 and not a.no_single:
        # the same clip on ONE rank, in batches of B windows: ids must be identical window by window (same kernels on the
        # same frames; the whole-clip maximum is the same number)
        whole = pipeline.ClipPart(n, 0, 1)
        pcm = span_of(whole)
        ctx1 = api.WhisperContext.new(path, max_segments=max(B, 1), max_clips=1, max_clip_samples=pcm.size, device=local)
        t0 = time.perf_counter()
        w1, t1, l1, _ = pipeline.transcribe_clip(ctx1, pcm, max_new=a.max_new, eot=-1)
        out["single_rank_wall_s"] = time.perf_counter() - t0
        ctx1.close()
        same = [bool(l1[w] == all_lens[w] and np.array_equal(t1[w][: l1[w]], all_toks[w][: l1[w]])) for w in range(a.windows)]
        out["windows_identical_to_single_rank"] = int(sum(same))
        out["windows"] = a.windows
        out["ok"] = bool(all(same))
        out["speedup_over_single_rank"] = out["single_rank_wall_s"] / out["wall_s_max_over_ranks"]
    barrier()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and out.get("ok") is False:
        sys.exit(4)


if __name__ == "__main__":
    main()
