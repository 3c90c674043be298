"""Segment sharding across GPUs.

The hot path shards by independent 30 s windows: `whisper_encode` consumes exactly one window per
call (src/main.rs:1799, 1822-1829) and decoding a window needs only that window's cross K/V, so
there is no collective on the data path.  One process per GPU; the model is replicated.  The only
exchange is the final gather of the small per-segment results (token ids, digests) -- NCCL on the
GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def segments_for_rank(n_segments: int, rank: int, world: int, contiguous: bool = False) -> List[int]:
    """Global segment ids handled by `rank`.  Round-robin (s mod G) by default; `contiguous` gives each
    rank one block of ceil(S/G) consecutive windows, so a long clip is read as one PCM span per GPU
    (+ the 240-sample halo frame i needs: samples [160 i, 160 i + 400), src/main.rs:1594-1597)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if contiguous:
        per = -(-n_segments // world)
        return list(range(min(rank * per, n_segments), min((rank + 1) * per, n_segments)))
    return list(range(rank, n_segments, world))


def pcm_span_for_segments(segs: Sequence[int], n_samples_total: int, seg_samples: int = 480000, halo: int = 240):
    """[lo, hi) sample range a rank must read for a contiguous block of windows of one long clip."""
    if not segs:
        return 0, 0
    lo = segs[0] * seg_samples
    hi = min(n_samples_total, (segs[-1] + 1) * seg_samples + halo)
    return lo, hi


def gather_segment_results(local: np.ndarray, seg_ids: Sequence[int], n_segments: int, device=None):
    """All-gather per-segment rows ([n_local, ...]) into the global [n_segments, ...] array on every rank
    (rows in global segment order).  Uses the default torch.distributed group: NCCL when `device` is a
    CUDA device, gloo on CPU.  Single-process (no group initialised) returns the local rows re-ordered."""
    import torch
    import torch.distributed as dist

    local = np.ascontiguousarray(local)
    out = np.zeros((n_segments,) + local.shape[1:], dtype=local.dtype)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out[list(seg_ids)] = local
        return out
    world = dist.get_world_size()
    cap = -(-n_segments // world)                      # rows per rank, padded to a common size
    pad = np.zeros((cap,) + local.shape[1:], dtype=local.dtype)
    pad[: len(seg_ids)] = local
    ids = np.full(cap, -1, dtype=np.int64)
    ids[: len(seg_ids)] = np.asarray(seg_ids, dtype=np.int64)
    t_rows = torch.from_numpy(pad)
    t_ids = torch.from_numpy(ids)
    if device is not None:
        t_rows, t_ids = t_rows.to(device), t_ids.to(device)
    rows = [torch.empty_like(t_rows) for _ in range(world)]
    idl = [torch.empty_like(t_ids) for _ in range(world)]
    dist.all_gather(rows, t_rows)
    dist.all_gather(idl, t_ids)
    for r, i in zip(rows, idl):
        r, i = r.cpu().numpy(), i.cpu().numpy()
        keep = i >= 0
        out[i[keep]] = r[keep]
    return out
