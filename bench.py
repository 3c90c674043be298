#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config, one JSON line on stdout.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

Metric: 30 s segments/sec through mel + encoder.  A step = one pass of the hot path (log-mel ->
encoder -> cross-attention K/V) over one batch of synthetic 30 s clips per GPU.
Workload at every N = the north-star target, configs[3]: whisper MEDIUM, 64 x 30 s segments per GPU
(">= 64 segments per GPU so tails don't dominate"), random-init weights in the
reference's ggml f16 file layout, synthetic 16 kHz PCM.  N > 1 shards independent segments across
ranks (weak scaling, no collective on the data path; one NCCL all-gather of the per-segment digests
at the end).  configs[1] (whisper base, batch 16) and configs[2] (whisper small greedy decode, batch 32)
ride along as sub-records.

  value      whole-job segments/s with the PCM already resident in HBM when the timed region starts
  e2e        the same metric through the host-facing calls (whisper_pcm_to_mel on pinned HOST buffers +
             whisper_encode + digest read-back): H2D and D2H inside the timed region
  sustained  the same step looped for >= 2 s with the clock sampler inside (steady-state clocks)
  roofline   the dominant kernel family (tcgen05 GEMM): algorithmic FLOPs / its device time, per-launch
             CUDA events on the launching stream, against MEASURED_PEAKS.json's BURST bf16 peak (a bracketed
             kernel runs at burst clocks); the >= 2 s whole-step figure is quoted against the SUSTAINED peak
  parity     sum|x| of the encoder output of the first segments against the CPU oracle run on the same
             clips (the cpu_baseline leg computes them anyway); a miss makes the run exit non-zero
  cpu_baseline  the CPU oracle (a port of the reference's algorithm; the reference itself cannot be
             built here) timed on the host cores on a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

METRIC = "30s segments/sec (mel+encoder)"
UNIT = "segments/s"
PARITY_TOL = 2e-3     # relative difference of sum|x| of the encoder output (encoder rel-L2 bar is 1e-2)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------
def encoder_flops(hp) -> dict:
    """Algorithmic FLOPs per 30 s segment (SURVEY.md section 8d / BASELINE.md section 3)."""
    d, L, Lt, T, nm = hp.n_audio_state, hp.n_audio_layer, hp.n_text_layer, hp.n_audio_ctx, hp.n_mels
    gemm = 2 * (2 * T) * d * 3 * nm + 2 * T * d * d * 3 + L * 24 * T * d * d + Lt * 4 * T * d * d
    attn = L * 4 * T * T * d
    parts = {"gemm_conv1": 2 * (2 * T) * d * 3 * nm, "gemm_conv2": 2 * T * d * d * 3, "gemm_qkv": L * 6 * T * d * d,
             "gemm_out": L * 2 * T * d * d, "gemm_fc1": L * 8 * T * d * d, "gemm_fc2": L * 8 * T * d * d,
             "gemm_cross": Lt * 4 * T * d * d}
    assert sum(parts.values()) == gemm
    return {"gemm": float(gemm), "attention": float(attn), "total": float(gemm + attn), "parts": parts}


def mel_bytes(hp, n_samples: int) -> float:
    return float(n_samples * 4 + hp.n_mels * (n_samples // 160) * 4)   # PCM f32 read + mel f32 write


def ensure_model(pkg, arch: str, rank: int, barrier) -> str:
    d = os.environ.get("WB_MODEL_DIR", "/tmp/wb_models")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, f"ggml-{arch}.bin")
    if rank == 0 and not os.path.exists(path):
        tmp = path + f".tmp{os.getpid()}"
        pkg.ggml_file.make_model(tmp, arch)
        os.replace(tmp, path)
    barrier()
    return path


def measured_peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "MEASURED_PEAKS.json"
        return d
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "_source": "fallback of B200_PROFILING.md"}


def gemm_traffic(arch: str, B: int):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/r0N_gemm_traffic.json, newest round first), or None when no capture exists for this workload."""
    for name in ("r02_gemm_traffic.json", "r01_gemm_traffic.json"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))
            e = d.get(f"{arch}_b{B}")
            if e:
                return {"dram_bytes_per_launch": e["dram_bytes_per_launch"], "unit": "bytes",
                        "algorithmic_bytes_per_launch": e.get("algorithmic_bytes_per_launch_avg"), "source": e["source"]}
        except Exception:
            pass
    return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons, streamed for the whole run; summarised per timed window."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows = []      # (arrival time, csv line)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if not self.proc:
            return
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()

    def window(self, t0: float, t1: float) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.03]
        # a region shorter than one sampling period falls back to the samples closest to it
        near = [r for (t, r) in self.rows if t <= t1 + 0.03][-3:]
        for r in (inside or near):
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": float(np.median(pw)) if pw else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_inside_timed_region": len(inside)}


# ---------------------------------------------------------------------------------------------
def cpu_encode_leg(pkg, model: str, pcm_list, n_threads: int, mel_threads: int = 4, orc=None, warm: bool = True):
    """The reference's CPU algorithm (oracle port) on the host cores: mel (4 threads, src/main.rs:1698) + encode
    of every clip in `pcm_list`.  Returns (segments/s, seconds, [sum|x| of each encoder output]).  `warm`: one
    untimed short-context encode first, so that the one-off widening of the F16 weights to f32 (model load work in
    the reference, 1437) is not billed to the first segment."""
    from oracle import pyoracle
    own = orc is None
    if own:
        orc = pyoracle.Oracle(model, n_threads=n_threads)
        if warm:
            orc.pcm_to_mel(pcm_list[0][:160 * 64], n_threads=mel_threads)
            orc.set_audio_ctx(16)
            orc.encode(0, n_threads=n_threads)
            orc.set_audio_ctx(0)
    digests = []
    t0 = time.perf_counter()
    for pcm in pcm_list:
        orc.pcm_to_mel(pcm, n_threads=mel_threads)
        enc = orc.encode(0, n_threads=n_threads)
        digests.append(enc)
    dt = time.perf_counter() - t0
    digests = [float(np.abs(e.astype(np.float64)).sum()) for e in digests]   # outside the timed region
    if own:
        orc.close()
    return len(pcm_list) / dt, dt, digests


def run_reference(args, pkg, rank: int):
    """--impl reference: the reference's own CPU implementation of the path.  The reference (Rust, absent `galois`
    dependency) cannot be built here, so this arm times the CPU oracle -- the port of its algorithm -- with all the
    host threads, one segment of the same workload per step."""
    if rank != 0:
        return
    hp = pkg.ggml_file.ARCHS[args.arch]
    model = ensure_model(pkg, args.arch, 0, lambda: None)
    cores = os.cpu_count() or 1
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 1))
    from oracle import pyoracle
    pcm = [pkg.synth.make_segment(s, args.samples) for s in range(2)]
    orc = pyoracle.Oracle(model, n_threads=cores)
    orc.pcm_to_mel(pcm[0][:160 * 64])          # widen the weights to f32 once (load-time work), untimed
    orc.set_audio_ctx(16)
    orc.encode(0)
    orc.set_audio_ctx(0)
    times = []
    for it in range(warmup + steps):
        _, dt, _ = cpu_encode_leg(pkg, model, [pcm[it % 2]], cores, orc=orc)
        if it >= warmup:
            times.append(dt)
    orc.close()
    total = float(np.sum(times))
    val = len(times) / total
    sample = (f"1 of the {args.batch} segments of a step per step ({steps} timed + {warmup} warm-up steps): mel 4 threads "
              f"(main.rs:1698) + encoder on {cores} host threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": total / len(times) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16 weights x f16-rounded activations, f32 accumulate (CPU)",
        "data": "synthetic",
        "config": {"workload": f"whisper {args.arch} mel + encoder (+ cross-KV), batch {args.batch} x 30 s segments per GPU, "
                               f"ggml f16 layout, random-init (bounded CPU sample: {sample})",
                   "arch": args.arch, "segments_per_step": 1},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference (szuwgh/whisper.rs) is unbuildable here (no rustc; galois dependency absent): "
                "this arm times the CPU oracle, a port of its algorithm",
    }), flush=True)


def decoder_bytes_per_step(hp, B: int, p_avg: float) -> dict:
    """Algorithmic HBM bytes of one single-token decode step for B sequences: the decoder
    weights once (F16), plus per sequence the cross-attention K/V of every layer and the self-attention
    K/V up to the current position."""
    d, L, T = hp.n_text_state, hp.n_text_layer, hp.n_audio_ctx
    w = (14 * L * d * d + hp.n_vocab * d) * 2
    cross = L * 2 * T * d * 2
    self_kv = L * 2 * p_avg * d * 2
    return {"weights": float(w), "cross_kv_per_seq": float(cross), "self_kv_per_seq": float(self_kv),
            "total": float(w + B * (cross + self_kv))}


def decoder_leg(args, pkg, api, torch, dist, rank, world, local_rank, dev, barrier, peaks):
    """BASELINE.json's second metric, decoder tokens/sec, on configs[2]: whisper small, batch 32, greedy
    decode to 224 tokens (device-side loop, one CUDA graph replayed per position)."""
    arch, B, n_new = args.dec_arch, args.dec_batch, args.dec_tokens
    hp = pkg.ggml_file.ARCHS[arch]
    model = ensure_model(pkg, arch, rank, barrier)
    ctx = api.WhisperContext.new(model, max_segments=B, max_clips=B, max_clip_samples=args.samples, device=local_rank,
                                 decode_capacity=True)
    clips = pkg.synth.make_clips(B, first_seg=7000 + rank * B, n_samples=args.samples)
    pcm = torch.from_numpy(clips).to(dev)
    api.whisper_pcm_to_mel(ctx, pcm)
    api.whisper_encode(ctx, 1, [0] * B, clip_ids=list(range(B)))
    ctx.sync()
    prompt = [ctx.token_sot]
    times, walls = [], []
    toks = marg = None
    for it in range(1 + args.dec_reps):           # first call captures the step graph (warm-up)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        # the user-facing call: prompt H2D, device-side greedy loop, token / margin / length D2H, all inside
        toks, marg, lens = api.whisper_decode_greedy(ctx, prompt, n_new, n_seqs=B, eot=-1)   # eot -1: never stops early
        walls.append(time.perf_counter() - t0)
        times.append(ctx.timings()["t_decode_us"] * 1e-6)
    t, tw = float(np.min(times[1:])), float(np.min(walls[1:]))
    if world > 1:
        tt = torch.tensor([t, tw], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t, tw = float(tt[0].item()), float(tt[1].item())
    ctx.close()
    by = decoder_bytes_per_step(hp, B, (len(prompt) + n_new) / 2.0)
    gbs = by["total"] * n_new / t / 1e9
    out = {
        "metric": "decoder tokens/sec (greedy)", "value": world * B * n_new / t, "unit": "tokens/s",
        "config": {"workload": f"whisper {arch} greedy decode to {n_new} tokens, batch {B} per GPU, after mel + encode "
                               f"of {B} x 30 s segments", "arch": arch, "batch_per_gpu": B, "new_tokens": n_new},
        "ms_per_token_step": t / n_new * 1e3, "timing": "CUDA events around the whole greedy call on its stream "
        "(prompt pass + one CUDA-graph replay per position), best of %d" % args.dec_reps,
        "e2e": {"value": world * B * n_new / tw, "unit": "tokens/s", "h2d_bytes_per_step": 4 * len(prompt) * B,
                "d2h_bytes_per_step": B * (8 * n_new + 4),
                "how": "host clock around whisper_decode_greedy (prompt upload, device loop, token + margin + length "
                       "read-back), best of %d" % args.dec_reps},
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": float(peaks["hbm_gbs"]), "unit": "GB/s",
                     "frac": gbs / float(peaks["hbm_gbs"]), "bytes_per_step": by, "peak_source": peaks["_source"]},
        "all_lengths_equal_new_tokens": bool((np.asarray(lens) == n_new).all()),
        "first_tokens_seq0": [int(x) for x in toks[0][:8]],
    }
    # ---- CPU baseline + parity for the decode leg (rank 0, N = 1): the oracle decodes sequence 0 for a few tokens
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        from oracle import pyoracle
        n_tok = args.dec_cpu_tokens
        cores = os.cpu_count() or 1
        orc = pyoracle.Oracle(model, n_threads=cores)
        orc.pcm_to_mel(clips[0])
        orc.encode(0)
        t0 = time.perf_counter()
        rt, rm = orc.decode_greedy(prompt, n_tok, eot=-1)
        dt = time.perf_counter() - t0
        orc.close()
        n_cmp, first_bad = 0, None
        for i in range(len(rt)):                  # greedy ids equal up to the first low-margin step (north_star rule)
            if int(toks[0][i]) != int(rt[i]):
                if rm[i] >= 2e-2:
                    first_bad = i
                break
            n_cmp += 1
        out["cpu_baseline"] = {"value": len(rt) / dt, "unit": "tokens/s", "cores": cores, "kind": "port",
                               "sample": f"sequence 0 of the {B}, {len(rt)} greedy tokens through the CPU oracle "
                                         f"({dt:.1f} s; per-layer ops on 1 thread, vocabulary projection on {cores})"}
        out["parity"] = {"tokens_equal_to_oracle": n_cmp, "n_checked": len(rt), "ok": first_bad is None,
                         "rule": "ids equal until the first step whose oracle top-1 margin is below 2e-2"}
    return out


# ---------------------------------------------------------------------------------------------
def encoder_leg(args, pkg, api, torch, dist, arch, B, K, W, rank, world, local_rank, dev, barrier, sampler, peaks,
                full: bool):
    """One workload (arch, B segments per GPU): timed device-resident steps, and with `full` the end-to-end leg,
    the >= 2 s sustained loop, the bracketed per-kernel pass, the oracle parity check and the CPU baseline."""
    hp = pkg.ggml_file.ARCHS[arch]
    n_samples = args.samples
    model = ensure_model(pkg, arch, rank, barrier)
    # the context runs on this stream, and the timed regions are bracketed by events recorded on it (torch's
    # default stream has handle 0, which the C-ABI reads as "create your own": events recorded on the default
    # stream would not see the library's kernels)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    t_load = time.perf_counter()
    ctx = api.WhisperContext.new(model, max_segments=B, max_clips=B, max_clip_samples=n_samples, device=local_rank,
                                 stream=stream.cuda_stream, decode_capacity=False)
    t_load = time.perf_counter() - t_load
    # ---- synthetic inputs: n_rot distinct batches so no step re-reads the previous step's PCM
    n_rot = 2 if B >= 32 else 4
    log(f"[rank {rank}] {arch}: context in {t_load:.1f} s; generating {n_rot} x {B} synthetic 30 s clips ...")
    host = [torch.from_numpy(pkg.synth.make_clips(B, first_seg=(rank * n_rot + r) * B, n_samples=n_samples)).pin_memory()
            for r in range(n_rot)]
    devb = [h.to(dev, non_blocking=True) for h in host]
    torch.cuda.synchronize()
    offs = [0] * B
    ids = list(range(B))
    pcm_bytes = B * n_samples * 4

    def step_device(i):
        api.whisper_pcm_to_mel(ctx, devb[i % n_rot])
        api.whisper_encode(ctx, 1, offs, clip_ids=ids)

    def timed(fn, k):
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for i in range(k):
            fn(i)
        e1.record(stream)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, (t0, t1)

    # ---- end-to-end leg through the host-facing calls on HOST buffers.  One context, software-pipelined one step
    # deep: step i+1's upload runs on the context's copy stream (wb_pcm_prefetch) under step i's encoder, and step
    # i+1 is submitted before step i's result is awaited (wb_encoder_digest_async / wb_wait), so the stream never
    # drains.
    res = torch.zeros(8, B, dtype=torch.float64).pin_memory()

    def e2e_submit(i, k):
        h = host[i % n_rot]
        api.whisper_pcm_to_mel_ptr(ctx, h.data_ptr(), n_samples, B)       # H2D of this step's input (prefetched copy)
        if i + 1 < k:
            api.whisper_pcm_prefetch_ptr(ctx, host[(i + 1) % n_rot].data_ptr(), pcm_bytes)
        api.whisper_encode(ctx, 1, offs, clip_ids=ids)
        return api.encoder_digest_async(ctx, res[i % 8].data_ptr(), B)     # D2H of this step's result, queued

    def run_e2e(k):
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        api.whisper_pcm_prefetch_ptr(ctx, host[0].data_ptr(), pcm_bytes)
        tickets = [e2e_submit(0, k)]
        for i in range(k):
            if i + 1 < k:
                tickets.append(e2e_submit(i + 1, k))
            api.wait(ctx, tickets[i])
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        barrier()
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for i in range(W):
        step_device(i)
    torch.cuda.synchronize()

    # ---- timed region 1: device-resident inputs (value)
    l0 = ctx.timings()["n_kernel_launches"]
    ms_dev, win_dev = timed(step_device, K)
    launches = ctx.timings()["n_kernel_launches"] - l0
    fl = encoder_flops(hp)
    rec = {"arch": arch, "B": B, "ms_dev": ms_dev, "win_dev": win_dev, "launches": int(launches), "fl": fl, "hp": hp,
           "t_load_s": t_load, "n_rot": n_rot}
    if not full:
        ctx.close()
        return rec
    # the one collective of the path: final gather of the small per-segment results (contiguous blocks of
    # B segments per rank, NCCL all_gather)
    rec["digests"] = pkg.shard.gather_segment_results(ctx.encoder_digest(B), pkg.shard.segments_for_rank(world * B, rank, world, True),
                                                      world * B, device=dev if world > 1 else None)
    # ---- timed region 2: end to end through the host-facing call
    run_e2e(min(4, max(2, W)))   # warm-up
    t0 = time.perf_counter()
    rec["ms_e2e"] = run_e2e(K)
    rec["win_e2e"] = (t0, time.perf_counter())
    # ---- timed region 3: the same step looped for >= args.sustain_s seconds (steady-state clocks / power cap)
    k_sus = max(K, int(np.ceil(args.sustain_s * 1e3 / (ms_dev / K)))) if args.sustain_s > 0 else 0
    if k_sus:
        rec["ms_sus"], rec["win_sus"] = timed(step_device, k_sus)
        rec["k_sus"] = k_sus
    # ---- timed region 4: same K steps with per-launch CUDA events -> kernel-family device time
    ctx.kernel_time_us("__enable__")
    ctx.kernel_time_us("__reset__")
    rec["ms_prof"], _ = timed(step_device, K)
    rec["fam"] = {f: ctx.kernel_time_us(f) for f in ("gemm", "attention", "mel_frames", "mel_normalize", "mel_window",
                                                      "layernorm", "fill", "gemm_conv1", "gemm_conv2", "gemm_qkv",
                                                      "gemm_out", "gemm_fc1", "gemm_fc2", "gemm_cross")}
    ctx.kernel_time_us("__disable__")
    # ---- parity + CPU baseline (rank 0, N = 1): the oracle on the first clips of batch 0; the GPU digests of the
    # same clips come from one more (untimed) step on batch 0
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        n_cpu = min(args.cpu_segments, B)
        step_device(0)
        gpu_dig = ctx.encoder_digest(B)[:n_cpu]
        cores = os.cpu_count() or 1
        log(f"[rank 0] cpu_baseline + parity: {n_cpu} segment(s) of {arch} through the CPU oracle on {cores} threads ...")
        v, s_total, cpu_dig = cpu_encode_leg(pkg, model, [host[0][c].numpy() for c in range(n_cpu)], cores)
        rel = [abs(g - c) / abs(c) for g, c in zip(gpu_dig, cpu_dig)]
        rec["cpu"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                      "sample": f"{n_cpu} of the {B} segments of one step: mel 4 threads (main.rs:1698) + encoder on "
                                f"{cores} host threads, {s_total:.1f} s of CPU work"}
        rec["parity"] = {"max_rel": float(max(rel)), "n_checked": n_cpu, "tol": PARITY_TOL, "ok": bool(max(rel) <= PARITY_TOL),
                         "what": "sum|x| of the ln_post output per segment (the author's probe, src/main.rs:1836-1849), "
                                 "GPU digest vs CPU oracle on the same clips",
                         "gpu": [float(x) for x in gpu_dig], "oracle": [float(x) for x in cpu_dig]}
    ctx.close()
    return rec


def run_gpu(args, pkg, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from whisper_rs_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    peaks = measured_peaks()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()   # early, so nvidia-smi is already streaming when the timed regions begin
    W, K = max(3, args.warmup), max(1, args.steps)
    arch, B, n_samples = args.arch, args.batch, args.samples
    rec = encoder_leg(args, pkg, api, torch, dist, arch, B, K, W, rank, world, local_rank, dev, barrier, sampler, peaks, True)
    base = None
    if not args.no_base and not (arch == "base" and B == 16):
        base = encoder_leg(args, pkg, api, torch, dist, "base", 16, max(K, 10), W, rank, world, local_rank, dev, barrier,
                           sampler, peaks, False)
    dec = None
    if not args.no_decoder:
        dec = decoder_leg(args, pkg, api, torch, dist, rank, world, local_rank, dev, barrier, peaks)
    cpu_faithful = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the reference-faithful thread setting (mel 4 threads, main.rs:1698; encoder on ONE
        # thread, main.rs:2074 passes n_threads = 1 and 1799 ignores it) on configs[0], the reference's own CPU case
        tiny = ensure_model(pkg, "tiny", 0, lambda: None)
        v, s_total, _ = cpu_encode_leg(pkg, tiny, [pkg.synth.make_segment(0, n_samples)], 1)
        cpu_faithful = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "arch": "tiny",
                        "sample": f"configs[0]: whisper tiny, one 30 s clip, mel 4 threads + encoder 1 thread "
                                  f"(the reference's own setting), {s_total:.1f} s"}
    if rank == 0:
        sampler.stop()
        hp, fl, fam = rec["hp"], rec["fl"], rec["fam"]
        ms_dev, ms_e2e, ms_prof = rec["ms_dev"], rec["ms_e2e"], rec["ms_prof"]
        seg_total = world * B * K
        value = seg_total / (ms_dev / 1e3)
        e2e_val = seg_total / (ms_e2e / 1e3)
        gemm_us, gemm_n = fam["gemm"]
        att_us, att_n = fam["attention"]
        mel_us, mel_n = fam["mel_frames"]
        gemm_tf = (fl["gemm"] * B * K) / (gemm_us * 1e-6) / 1e12 if gemm_us else 0.0
        att_tf = (fl["attention"] * B * K) / (att_us * 1e-6) / 1e12 if att_us else 0.0
        mel_gbs = (mel_bytes(hp, n_samples) * B * K) / (mel_us * 1e-6) / 1e9 if mel_us else 0.0
        burst, sust = float(peaks["bf16_tflops"]), float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
        step_us = ms_prof * 1e3 / K
        shares = {k: (v[0] / K) / step_us for k, v in fam.items() if v[1] and not k.startswith("gemm_")}
        gemm_parts = {k: {"us_per_launch": fam[k][0] / fam[k][1], "launches_per_step": fam[k][1] / K,
                          "tflops": fl["parts"][k] * B * K / (fam[k][0] * 1e-6) / 1e12}
                      for k in fl["parts"] if fam[k][1]}
        whole_tf = fl["total"] * B / (ms_dev / K * 1e-3) / 1e12
        clocks = sampler.window(*rec["win_dev"])
        clocks["e2e_region"] = sampler.window(*rec["win_e2e"])
        sustained = None
        if rec.get("k_sus"):
            ms_sus, k_sus = rec["ms_sus"], rec["k_sus"]
            sus_tf = fl["total"] * B / (ms_sus / k_sus * 1e-3) / 1e12
            sustained = {"value": world * B * k_sus / (ms_sus / 1e3), "unit": UNIT, "steps": k_sus, "seconds": ms_sus / 1e3,
                         "ms_per_step": ms_sus / k_sus, "whole_step_tflops": sus_tf,
                         "whole_step_frac_of_sustained_peak": sus_tf / sust, "whole_step_frac_of_burst_peak": sus_tf / burst,
                         "clocks": sampler.window(*rec["win_sus"])}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands, f32 accumulate (tcgen05 kind::f16); f32 residual stream; f32 mel",
            "data": "synthetic",
            "config": {
                "workload": f"whisper {arch} mel + encoder (+ cross-KV), batch {B} x 30 s segments per GPU, "
                            f"ggml f16 layout, random-init",
                "arch": arch, "segments_per_step_per_gpu": B, "n_samples_per_segment": n_samples,
                "sharding": "independent segments per rank, no data-path collective; one all_gather of digests",
                "cache": f"per-step working set (~{(B * 1500 * hp.n_audio_state * 2 * (8 + 4 * hp.n_text_layer)) / 1e6:.0f} MB of "
                         f"activations) exceeds the 126 MB L2; PCM rotates over {rec['n_rot']} distinct batches",
                "e2e_pipeline": "each e2e step = H2D of the step's pinned host PCM (wb_pcm_prefetch on the context's "
                                "copy stream, issued one step ahead) + whisper_pcm_to_mel + whisper_encode + digest "
                                "read-back (D2H, wb_encoder_digest_async); step i+1 is submitted before step i's "
                                "result is awaited; one context, one compute stream; timed with the host clock between "
                                "device-wide synchronisations, max over ranks",
                "roofline_timing": "a further timed pass of the same K steps with per-launch CUDA events on the launching stream",
                "context_load_s": rec["t_load_s"],
            },
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": B * n_samples * 4,
                    "d2h_bytes_per_step": B * 8, "ms_per_step": ms_e2e / K},
            "gpu_launches": rec["launches"],
            "clocks": clocks,
            "roofline": {
                "kernel": "gemm2_f16_tcgen05_kernel (all tile widths; conv stem, QKV, out-proj, MLP, cross-KV)",
                "bound": "tensor", "achieved": gemm_tf, "peak": burst, "unit": "TFLOP/s",
                "frac": gemm_tf / burst if burst else None, "traffic": gemm_traffic(arch, B),
                "peak_source": f"{peaks['_source']} bf16_tflops (BURST: the kernel times are per-launch event brackets)",
                "frac_of_sustained_peak": gemm_tf / sust if sust else None,
                "launches_per_step": gemm_n / K, "avg_launch_us": gemm_us / gemm_n if gemm_n else None,
                "share_of_step": shares.get("gemm"),
                "algorithmic_flops_per_segment": fl["gemm"],
                # the per-launch event brackets drain the stream (the bracketed pass is slower than the timed step);
                # the same FLOPs over this family's share of the UNbracketed step time, for comparison only
                "achieved_from_share_of_unbracketed_step":
                    (fl["gemm"] * B) / (shares["gemm"] / sum(shares.values()) * (ms_dev / K) * 1e-3) / 1e12
                    if shares.get("gemm") else None,
            },
            "kernels": {
                "attention": {"tflops": att_tf, "frac_of_burst_peak": att_tf / burst if burst else None,
                              "share_of_step": shares.get("attention"), "flops_per_segment": fl["attention"]},
                "mel_frames": {"gbs": mel_gbs, "frac_of_hbm": mel_gbs / float(peaks["hbm_gbs"]),
                               "share_of_step": shares.get("mel_frames"), "bytes_per_segment": mel_bytes(hp, n_samples)},
                "shares_of_step": shares,
                "gemm_by_call_site": gemm_parts,
                "whole_step_tflops": whole_tf,
                "whole_step_frac_of_burst_peak": whole_tf / burst,
                "whole_step_frac_of_sustained_peak": whole_tf / sust,
                "timed_region_s": ms_dev / 1e3,
            },
            "sustained": sustained,
            "parity": rec.get("parity"),
            "cpu_baseline": rec.get("cpu"),
            "cpu_baseline_reference_threads": cpu_faithful,
            "digest_segment0": float(rec["digests"][0]),
        }
        if base is not None:
            bf = base["fl"]
            bms = base["ms_dev"] / max(K, 10)
            out["base_b16"] = {
                "metric": METRIC, "value": world * 16 / (bms / 1e3), "unit": UNIT, "ms_per_step": bms,
                "config": {"workload": "whisper base mel + encoder (+ cross-KV), batch 16 x 30 s segments per GPU (configs[1])"},
                "whole_step_tflops": bf["total"] * 16 / (bms * 1e-3) / 1e12,
                "whole_step_frac_of_burst_peak": bf["total"] * 16 / (bms * 1e-3) / 1e12 / burst,
                "gpu_launches_per_step": base["launches"] / max(K, 10), "clocks": sampler.window(*base["win_dev"]),
            }
        out["decoder"] = dec
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        bad = []
        if rec.get("parity") and not rec["parity"]["ok"]:
            bad.append(f"encoder digest parity: max_rel {rec['parity']['max_rel']:.3e} > {PARITY_TOL}")
        if dec and dec.get("parity") and not dec["parity"]["ok"]:
            bad.append("decoder greedy ids differ from the oracle at a high-margin step")
        if bad:
            log("bench.py: PARITY FAILED: " + "; ".join(bad))
            sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--arch", default="medium")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--samples", type=int, default=480000)
    ap.add_argument("--cpu-segments", type=int, default=2)
    ap.add_argument("--sustain-s", type=float, default=2.0, help="length of the steady-state loop (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle legs (and the parity check)")
    ap.add_argument("--no-decoder", action="store_true", help="skip the decoder tokens/sec leg")
    ap.add_argument("--no-base", action="store_true", help="skip the whisper base / batch 16 sub-record")
    ap.add_argument("--dec-arch", default="small")
    ap.add_argument("--dec-batch", type=int, default=32)
    ap.add_argument("--dec-tokens", type=int, default=224)
    ap.add_argument("--dec-reps", type=int, default=2)
    ap.add_argument("--dec-cpu-tokens", type=int, default=24)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    pkg = graft.load_package()
    if args.impl == "reference":
        run_reference(args, pkg, rank)
        return
    run_gpu(args, pkg, rank, world, local_rank)


if __name__ == "__main__":
    main()
