"""One long clip -> token ids per 30 s window, over the reference's three calls.

The reference's driver (`fn main`, src/main.rs:2065-2075) loads a clip, calls whisper_pcm_to_mel once
and whisper_encode once at mel_offset 0.  This module is that driver for a clip of any length: every
3000-frame window of the clip's mel (mel_offset = 3000 s, the window copy of 1816-1829) is encoded and
greedily decoded, in batches of `ctx.max_segments` windows per call.

Sharded over GPUs: rank r takes a contiguous block of windows, reads only the
PCM span of that block plus the 240-sample halo (frame i = samples [160 i, 160 i + 400), 1594-1597) and
runs mel -> encode -> decode locally.  The single coupling between ranks is the whole-clip maximum of
clamp_and_normalize (1655-1662): one MAX all-reduce of 4 bytes between the two phases of the mel.
The final gather of token ids is the only other exchange.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np

from . import api, shard

HOP, FRAMES_PER_WINDOW = 160, 3000     # a window is 2 * n_audio_ctx frames (1816): 3000 for every Whisper model


def n_windows(n_samples: int, fpw: int = FRAMES_PER_WINDOW) -> int:
    """30 s windows of a clip: ceil(n_len / 3000) with n_len = n_samples / 160 (src/main.rs:1575)."""
    n_len = n_samples // HOP
    return max(1, -(-n_len // fpw))


class ClipPart:
    """What one rank reads and computes of a clip of `n_samples` samples."""

    def __init__(self, n_samples: int, rank: int, world: int, fpw: int = FRAMES_PER_WINDOW):
        self.n_samples = n_samples
        self.fpw = fpw
        self.n_len = n_samples // HOP
        self.windows: List[int] = shard.segments_for_rank(n_windows(n_samples, fpw), rank, world, contiguous=True)
        self.lo, self.hi = shard.pcm_span_for_segments(self.windows, n_samples, HOP * fpw, 240)
        if self.windows:
            f0 = self.windows[0] * fpw
            self.n_frames = max(0, min((self.windows[-1] + 1) * fpw, self.n_len) - f0)
        else:
            self.n_frames = 0

    def local_offset(self, window: int) -> int:
        return (window - self.windows[0]) * self.fpw


def default_prompt(ctx: api.WhisperContext) -> List[int]:
    return [ctx.token_sot]


def transcribe_clip(ctx: api.WhisperContext, pcm: np.ndarray, *, rank: int = 0, world: int = 1,
                    reduce_max: Optional[Callable[[float], float]] = None,
                    prompt: Optional[Sequence[int]] = None, max_new: int = 224, eot: Optional[int] = None,
                    pcm_is_local_span: bool = False, n_samples_total: Optional[int] = None):
    """mel + encode + greedy decode of this rank's windows of one clip.

    `pcm` is the whole clip (f32 [n]) or, with `pcm_is_local_span`, only samples [lo, hi) of it as given by
    ClipPart (then `n_samples_total` is the clip length).  `reduce_max(x)` returns the maximum of x over all
    ranks (None for world == 1).  Returns (window ids, tokens [n_local][max_new], lengths, margins)."""
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    total = int(n_samples_total if pcm_is_local_span else pcm.size)
    part = ClipPart(total, rank, world, 2 * ctx.audio_ctx)
    if world > 1 and reduce_max is None:
        raise ValueError("world > 1 needs reduce_max (the whole-clip maximum couples the parts)")
    local_max = -1e20                                   # mmax's initial value (1655)
    if part.windows and part.n_frames > 0:
        span = pcm if pcm_is_local_span else pcm[part.lo:part.hi]
        if span.size != part.hi - part.lo:
            raise ValueError("pcm span does not match ClipPart")
        local_max = float(api.whisper_pcm_to_logmel(ctx, span, part.n_frames)[0])
    clip_max = reduce_max(local_max) if reduce_max is not None else local_max
    toks = np.zeros((len(part.windows), max_new), dtype=np.int32)
    lens = np.zeros(len(part.windows), dtype=np.int32)
    marg = np.zeros((len(part.windows), max_new), dtype=np.float32)
    if part.windows and part.n_frames > 0:
        api.whisper_mel_normalize(ctx, [clip_max])
        B = ctx.max_segments
        pr = list(prompt) if prompt is not None else default_prompt(ctx)
        for b0 in range(0, len(part.windows), B):
            ws = part.windows[b0:b0 + B]
            api.whisper_encode(ctx, 1, [part.local_offset(w) for w in ws], clip_ids=[0] * len(ws))
            t, m, l = api.whisper_decode_greedy(ctx, pr, max_new, n_seqs=len(ws), eot=eot)
            toks[b0:b0 + len(ws)], marg[b0:b0 + len(ws)], lens[b0:b0 + len(ws)] = t, m, l
    return part.windows, toks, lens, marg


def conditioned_prompt(ctx: api.WhisperContext, prompt_past: Sequence[int], prompt_init: Sequence[int]) -> List[int]:
    """The prompt of a window that follows decoded text -- `prompt_past` of WhisperContext (src/main.rs:356; never
    written by the reference, whose decode loop does not exist) with upstream whisper.cpp v1.0.3's rule, the code
    base main.rs transliterates: [token_prev] + the last n_text_ctx / 2 tokens of the text decoded so far, then the
    initial prompt ([sot] ...).  No past text: the initial prompt alone."""
    if not prompt_past:
        return list(prompt_init)
    n_take = min(ctx.n_text_ctx // 2, len(prompt_past))
    return [ctx.token_prev] + [int(t) for t in prompt_past[len(prompt_past) - n_take:]] + list(prompt_init)


def transcribe_long_form(ctx: api.WhisperContext, pcm: np.ndarray, *, prompt_init: Optional[Sequence[int]] = None,
                         max_new: int = 224, eot: Optional[int] = None, condition_on_previous: bool = True):
    """Long-form decoding with `prompt_past` (src/main.rs:356): the windows of one clip are decoded IN ORDER and window
    w's prompt carries the tokens decoded for the windows before it (conditioned_prompt).  This makes the windows of
    a clip sequential: the mel is computed once for the clip, then encode + greedy decode run
    window after window -- independent clips, not windows, are the unit of parallelism here.
    Returns (tokens per window [list of int32 arrays], prompts used per window)."""
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    fpw = 2 * ctx.audio_ctx
    n_win = n_windows(pcm.size, fpw)
    api.whisper_pcm_to_mel(ctx, pcm)
    init = list(prompt_init) if prompt_init is not None else default_prompt(ctx)
    eot_id = ctx.token_eot if eot is None else eot
    past: List[int] = []
    out, prompts = [], []
    for w in range(n_win):
        api.whisper_encode(ctx, 1, w * fpw)                    # the reference's call (2074) at this window's offset
        prompt = conditioned_prompt(ctx, past, init) if condition_on_previous else list(init)
        room = ctx.n_text_ctx - len(prompt)
        t, _, l = api.whisper_decode_greedy(ctx, prompt, min(max_new, room), n_seqs=1, eot=eot_id)
        ids = t[0][: int(l[0])]
        out.append(ids.copy())
        prompts.append(prompt)
        if condition_on_previous:
            # upstream keeps what it took plus the new tokens: the text decoded so far, newest last
            taken = prompt[1:len(prompt) - len(init)] if len(prompt) > len(init) else []
            past = taken + [int(x) for x in ids if int(x) != eot_id]
    return out, prompts


class WhisperTokenData:
    """WhisperTokenData (src/main.rs:317-331): id, the time-stamp token it implies (tid), and the window-relative
    times a time-stamp id carries (t0 / t1 in units of 10 ms, as upstream: id - token_beg steps of 20 ms).  The
    probabilities of the reference's struct (p, pt, ptsum) belong to its sampling code, which it never implemented;
    the greedy loop here reports the top-1 logit margin instead."""

    __slots__ = ("id", "tid", "margin", "t0", "t1")

    def __init__(self, id: int, tid: int, margin: float, t0: int, t1: int):
        self.id, self.tid, self.margin, self.t0, self.t1 = id, tid, margin, t0, t1


class WhisperSegment:
    """WhisperSegment (src/main.rs:599-604): [t0, t1) in units of 10 ms from the start of the clip, the text of the
    window's text tokens, and the tokens themselves."""

    __slots__ = ("t0", "t1", "text", "tokens")

    def __init__(self, t0: int, t1: int, text: bytes, tokens: List[WhisperTokenData]):
        self.t0, self.t1, self.text, self.tokens = t0, t1, text, tokens


def assemble_segments(ctx: api.WhisperContext, windows: Sequence[int], toks: np.ndarray, lens: np.ndarray,
                      margins: Optional[np.ndarray] = None, n_samples: Optional[int] = None) -> List[WhisperSegment]:
    """result_all (src/main.rs:354): one WhisperSegment per decoded 30 s window.  Window w covers
    [3000 w, 3000 (w + 1)) centiseconds (clipped to the clip's length when `n_samples` is given); a time-stamp token
    (id >= token_beg, 568) carries (id - token_beg) * 2 centiseconds relative to its window and ends the run of text
    tokens before it; text = the window's text tokens (ids below eot) through id_to_token (544)."""
    fpw = 2 * ctx.audio_ctx
    out: List[WhisperSegment] = []
    for i, w in enumerate(windows):
        n = int(lens[i])
        ids = [int(x) for x in toks[i][:n]]
        w0 = w * fpw                                       # centiseconds: one mel frame = 10 ms
        w1 = (w + 1) * fpw if n_samples is None else min((w + 1) * fpw, n_samples // HOP)
        tdata: List[WhisperTokenData] = []
        t_cur = w0
        for k, tok in enumerate(ids):
            m = float(margins[i][k]) if margins is not None else 0.0
            if tok >= ctx.token_beg:                       # time-stamp token: moves the clock
                t_cur = min(w0 + (tok - ctx.token_beg) * 2, w1)
                tdata.append(WhisperTokenData(tok, tok, m, t_cur, t_cur))
            else:
                tdata.append(WhisperTokenData(tok, ctx.token_beg + (t_cur - w0) // 2, m, t_cur, w1))
        out.append(WhisperSegment(w0, w1, ctx.tokens_to_text(ids), tdata))
    return out


def torch_reduce_max(device=None) -> Callable[[float], float]:
    """MAX all-reduce of one f32 over the default torch.distributed group (NCCL on `device`, gloo on CPU)."""
    import torch
    import torch.distributed as dist

    def f(x: float) -> float:
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return x
        t = torch.tensor([x], dtype=torch.float32, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return f
