"""GPU: one long clip through mel -> encode -> greedy decode window by window (configs[4]'s shape at
micro scale), on one rank and split over two emulated ranks with the whole-clip maximum exchanged between
the two phases of the mel, against the oracle run on the whole clip."""
import ctypes as C

import numpy as np
import pytest

from conftest import mel_close, rel_l2
from test_gpu_decoder import _check_greedy

pytestmark = pytest.mark.gpu


def _clip(pkg, n_samples, seed=300):
    # loud first part, quiet rest: the whole-clip maximum lives in rank 0's span, so a rank that normalised
    # with its own maximum would be wrong
    x = pkg.synth.make_segment(seed, n_samples, silent_tail_s=0.0)
    x[n_samples // 3:] *= 0.01
    return x


def test_two_phase_mel_equals_one_phase(pkg, model_path):
    from whisper_rs_b200 import api
    n = 50000
    ctx = api.WhisperContext.new(model_path("micro"), max_segments=1, max_clips=2, max_clip_samples=n)
    clips = np.stack([pkg.synth.make_segment(40 + s, n, 0.1) for s in range(2)])
    api.whisper_pcm_to_mel(ctx, clips)
    ref = [ctx.mel(c).copy() for c in range(2)]
    mx = api.whisper_pcm_to_logmel(ctx, clips)
    assert mx.shape == (2,) and np.all(np.isfinite(mx))
    with pytest.raises(api.WsError) as e:              # un-normalised mel must not be encoded
        api.whisper_encode(ctx, 1, 0)
    assert e.value.variant == "Unexpected"
    api.whisper_mel_normalize(ctx)
    for c in range(2):
        assert np.array_equal(ctx.mel(c), ref[c])
        # the reported maximum is the maximum of log10-mel: normalised max = (mx + 4) / 4
        assert abs(float(ref[c].max()) - (mx[c] + 4.0) / 4.0) < 1e-6
    with pytest.raises(api.WsError):                   # normalising twice is an error
        api.whisper_mel_normalize(ctx)
    # an explicit frame count: frames past the samples are zero-filled (1596-1600)
    api.whisper_pcm_to_logmel(ctx, clips[0][:16000], n_frames=120)
    api.whisper_mel_normalize(ctx)
    assert ctx.mel(0).shape == (ctx.n_mels, 120)
    ctx.close()


@pytest.mark.parametrize("n_win_x10", [26, 30])
def test_long_clip_single_rank_and_two_ranks_vs_oracle(pkg, pyoracle, model_path, n_win_x10):
    from whisper_rs_b200 import api, pipeline
    arch = "micro"
    hp = pkg.ggml_file.ARCHS[arch]
    fpw = 2 * hp.n_audio_ctx
    n = int(n_win_x10 * fpw * 160 // 10) + 77            # 2.6 / 3.0 windows (+ a ragged tail)
    pcm = _clip(pkg, n)
    eot = hp.n_vocab - 1
    # ---- oracle on the whole clip
    orc = pyoracle.Oracle(model_path(arch))
    ref_mel = orc.pcm_to_mel(pcm)
    n_win = pipeline.n_windows(n, fpw)
    assert n_win == -(-(n // 160) // fpw)
    ref = []
    for w in range(n_win):
        orc.encode(w * fpw)
        ref.append(orc.decode_greedy([7], 10, eot=eot))
    # ---- one rank
    ctx = api.WhisperContext.new(model_path(arch), max_segments=2, max_clips=1, max_clip_samples=n)
    wins, toks, lens, marg = pipeline.transcribe_clip(ctx, pcm, prompt=[7], max_new=10, eot=eot)
    assert wins == list(range(n_win))
    assert mel_close(ctx.mel(0), ref_mel)
    agree = sum(_check_greedy(toks[i], lens[i], ref[i][0], ref[i][1]) for i in range(n_win))
    assert agree >= n_win                                # not vacuous
    # ---- two emulated ranks on one GPU: phase 1 on each part, MAX exchange, phase 2 + encode + decode
    world = 2
    parts = [pipeline.ClipPart(n, r, world, fpw) for r in range(world)]
    assert sorted(w for p in parts for w in p.windows) == list(range(n_win))
    local_max = []
    for r in range(world):
        p = parts[r]
        local_max.append(float(api.whisper_pcm_to_logmel(ctx, pcm[p.lo:p.hi], p.n_frames)[0]))
    gmax = max(local_max)
    assert local_max[0] > local_max[1] + 1.0             # the coupling matters for this clip
    got = {}
    for r in range(world):
        p = parts[r]
        w2, t2, l2, _ = pipeline.transcribe_clip(ctx, pcm[p.lo:p.hi], rank=r, world=world, reduce_max=lambda x: gmax,
                                                 prompt=[7], max_new=10, eot=eot, pcm_is_local_span=True,
                                                 n_samples_total=n)
        f0 = p.windows[0] * fpw
        assert mel_close(ctx.mel(0), ref_mel[:, f0:f0 + p.n_frames]), r
        for i, w in enumerate(w2):
            got[w] = (t2[i], l2[i])
    for w in range(n_win):                               # same kernels on the same frames: identical ids
        assert got[w][1] == lens[w] and np.array_equal(got[w][0][:lens[w]], toks[w][:lens[w]]), w
    ctx.close()


def test_async_digest_tickets(pkg, model_path):
    import torch
    from whisper_rs_b200 import api
    n = 30720
    ctx = api.WhisperContext.new(model_path("micro"), max_segments=2, max_clips=2, max_clip_samples=n)
    out = torch.zeros(4, 2, dtype=torch.float64).pin_memory()
    want = []
    for i in range(4):
        clips = np.stack([pkg.synth.make_segment(60 + 2 * i + s, n, 0.1) for s in range(2)])
        api.whisper_pcm_to_mel(ctx, clips)
        api.whisper_encode(ctx, 1, [0, 0], clip_ids=[0, 1])
        want.append(ctx.encoder_digest(2).copy())
    tickets = []
    for i in range(4):                                   # four batches in flight before the first wait
        clips = np.stack([pkg.synth.make_segment(60 + 2 * i + s, n, 0.1) for s in range(2)])
        api.whisper_pcm_to_mel(ctx, clips)
        api.whisper_encode(ctx, 1, [0, 0], clip_ids=[0, 1])
        tickets.append(api.encoder_digest_async(ctx, out[i].data_ptr(), 2))
    for i in range(4):
        api.wait(ctx, tickets[i])
        assert np.array_equal(out[i].numpy(), want[i]), i
    ctx.close()


def test_assemble_segments(pkg, model_path):
    """result_all / WhisperSegment (src/main.rs:354, 599-604): window times in centiseconds, text through id_to_token,
    time-stamp tokens (id >= token_beg, 568) moving the clock inside their window."""
    from whisper_rs_b200 import api, pipeline
    ctx = api.WhisperContext.new(model_path("tiny.ml"), max_segments=1, max_clips=1, max_clip_samples=480000)
    fpw = 2 * ctx.n_audio_ctx
    beg = ctx.token_beg
    toks = np.array([[beg, 11, 12, beg + 50, 13, beg + 100, 0, 0],
                     [21, 22, 23, 24, 25, 26, 27, 28]], dtype=np.int32)
    lens = np.array([6, 8], dtype=np.int32)
    segs = pipeline.assemble_segments(ctx, [0, 2], toks, lens, None, n_samples=2 * fpw * 160 + 40 * 160)
    assert [(s.t0, s.t1) for s in segs] == [(0, fpw), (2 * fpw, 2 * fpw + 40)]       # second window clipped to the clip
    assert segs[0].text == ctx.tokens_to_text([11, 12, 13]) and segs[1].text == ctx.tokens_to_text(list(range(21, 29)))
    t = segs[0].tokens
    assert [x.id for x in t] == [beg, 11, 12, beg + 50, 13, beg + 100]
    assert (t[0].t0, t[1].t0, t[3].t0, t[4].t0, t[5].t0) == (0, 0, 100, 100, 200)      # 2 cs per time-stamp step
    assert t[1].tid == beg and t[4].tid == beg + 50
    assert all(x.t0 >= segs[1].t0 for x in segs[1].tokens) and len(segs[1].tokens) == 8
    ctx.close()


def test_long_form_prompt_past_vs_oracle(pkg, pyoracle, model_path):
    """prompt_past (src/main.rs:356): windows decoded in order, each conditioned on the text decoded before it
    ([prev] + the last n_text_ctx / 2 tokens + [sot]).  The oracle replays the same prompts window by window."""
    from whisper_rs_b200 import api, pipeline
    arch = "micro"
    hp = pkg.ggml_file.ARCHS[arch]
    fpw = 2 * hp.n_audio_ctx
    n = 4 * fpw * 160
    pcm = pkg.synth.make_segment(310, n, silent_tail_s=0.0)
    eot = hp.n_vocab - 1
    ctx = api.WhisperContext.new(model_path(arch), max_segments=1, max_clips=1, max_clip_samples=n)
    ctx.token_prev = 3                                    # micro's vocabulary is smaller than the real special ids
    init = [7]
    toks, prompts = pipeline.transcribe_long_form(ctx, pcm, prompt_init=init, max_new=6, eot=eot)
    assert len(toks) == 4 and prompts[0] == init
    assert prompts[1][0] == 3 and prompts[1][-1] == 7 and prompts[1][1:-1] == [int(x) for x in toks[0] if x != eot]
    assert len(prompts[3]) <= 1 + hp.n_text_ctx // 2 + len(init)
    assert prompts[3][1:-1][-len(toks[2]):] == [int(x) for x in toks[2] if x != eot][-len(toks[2]):]
    orc = pyoracle.Oracle(model_path(arch))
    orc.pcm_to_mel(pcm)
    agree = 0
    for w in range(4):
        orc.encode(w * fpw)
        rt, rm = orc.decode_greedy(prompts[w], 6, eot=eot)        # same prompt: the conditioning is what is under test
        agree += _check_greedy(toks[w], len(toks[w]), rt, rm)
    assert agree >= 8
    # unconditioned decoding of the same clip differs from window 1 on (the prompt matters)
    toks_u, prompts_u = pipeline.transcribe_long_form(ctx, pcm, prompt_init=init, max_new=6, eot=eot, condition_on_previous=False)
    assert all(p == init for p in prompts_u)
    ctx.close()
