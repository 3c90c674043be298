"""N > 1 host logic on CPU: segment sharding + the final gather, world_size 2 over gloo."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_segments_for_rank(pkg):
    sh = pkg.shard
    for S in (1, 7, 120):
        for G in (1, 2, 4, 8):
            for contiguous in (False, True):
                got = sorted(s for r in range(G) for s in sh.segments_for_rank(S, r, G, contiguous))
                assert got == list(range(S))          # every window exactly once
    assert sh.segments_for_rank(120, 3, 8) == list(range(3, 120, 8))
    assert sh.segments_for_rank(120, 7, 8, contiguous=True) == list(range(105, 120))
    # one long clip, 120 windows over 8 GPUs: each rank reads its block + the 240-sample halo
    lo, hi = sh.pcm_span_for_segments(sh.segments_for_rank(120, 1, 8, True), 120 * 480000)
    assert (lo, hi) == (15 * 480000, 30 * 480000 + 240)
    lo, hi = sh.pcm_span_for_segments(sh.segments_for_rank(120, 7, 8, True), 120 * 480000)
    assert hi == 120 * 480000


def test_clip_parts_cover_the_clip(pkg):
    """One long clip over G ranks: every window once, every frame once, and every sample a rank's frames read
    lies inside its span or past the end of the clip (where the reference zero-fills, src/main.rs:1596-1600)."""
    import importlib
    pipeline = importlib.import_module("whisper_rs_b200.pipeline")
    for fpw in (3000, 192):
        for n in (100, 160 * fpw - 1, 160 * fpw, 160 * fpw + 1, int(2.6 * 160 * fpw) + 77, 7 * 160 * fpw, 120 * 160 * fpw):
            n_len = n // 160
            for G in (1, 2, 3, 8):
                parts = [pipeline.ClipPart(n, r, G, fpw) for r in range(G)]
                wins = sorted(w for p in parts for w in p.windows)
                assert wins == list(range(pipeline.n_windows(n, fpw)))
                assert sum(p.n_frames for p in parts) == n_len
                for p in parts:
                    if not p.windows or p.n_frames == 0:
                        continue
                    f0 = p.windows[0] * fpw
                    assert p.lo == 160 * f0 and p.local_offset(p.windows[-1]) == (len(p.windows) - 1) * fpw
                    last = 160 * (f0 + p.n_frames - 1) + 400      # one past the last sample the part's frames read
                    assert p.hi >= min(last, n) and p.hi <= n
    # config 5: 1 h of audio = 120 windows over 8 GPUs, 15 each
    parts = [pipeline.ClipPart(3600 * 16000, r, 8) for r in range(8)]
    assert [len(p.windows) for p in parts] == [15] * 8 and parts[1].hi - parts[1].lo == 15 * 480000 + 240


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, %(root)r)
    import __graft_entry__ as g
    pkg = g.load_package()
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    S = 7
    for contiguous in (False, True):
        segs = pkg.shard.segments_for_rank(S, rank, world, contiguous)
        # stand-in for the per-segment greedy tokens: row s = s*100 + position
        local = np.array([[s * 100 + p for p in range(5)] for s in segs], dtype=np.int32).reshape(len(segs), 5)
        full = pkg.shard.gather_segment_results(local, segs, S)
        want = np.array([[s * 100 + p for p in range(5)] for s in range(S)], dtype=np.int32)
        assert (full == want).all(), (rank, contiguous, full)
        dig = pkg.shard.gather_segment_results(np.array([float(s) + 0.5 for s in segs]), segs, S)
        assert np.allclose(dig, np.arange(S) + 0.5)
    # the one coupling of a split clip: the whole-clip maximum (clamp_and_normalize, src/main.rs:1655-1662)
    import importlib
    pipeline = importlib.import_module("whisper_rs_b200.pipeline")
    local = [3.25, -1.5][rank]
    assert pipeline.torch_reduce_max()(local) == 3.25
    assert pipeline.torch_reduce_max()(-1e20 if rank == 0 else -7.0) == -7.0     # a rank with no frames
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_gather_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29671")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29671", str(script)],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2
