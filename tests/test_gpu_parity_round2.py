"""GPU: the parity cases VERDICT round 1 asked for, all through the C-ABI against the CPU oracle:
the benchmarked batch shapes, full-depth encoders of every named architecture, the LayerNorm fold on rows
whose mean dwarfs their spread, run-to-run determinism, the decode step deep into the text context, and the
state changes (audio context, prefetch, output strides) the advisor flagged."""
import dataclasses
import os

import numpy as np
import pytest

from conftest import rel_l2
from test_gpu_decoder import LOGIT_TOL, MARGIN_TOL, _check_greedy

pytestmark = pytest.mark.gpu
ENC_TOL = 1e-2


# ------------------------------------------------------------------------------------------------
def test_base_batch16_vs_oracle_and_bit_reproducible(pkg, pyoracle, model_path):
    """configs[1] at its benchmarked shape: 16 x 30 s segments in one call (M = 24000 flattened rows: the
    persistent tile loop, tiles straddling segment boundaries).  Segments 0, 7 and 15 against the oracle; the
    whole batch encoded twice must agree BIT FOR BIT (no atomics on floating-point data anywhere in the encoder:
    LayerNorm statistics and the sum|x| digests are added in a fixed order)."""
    from whisper_rs_b200 import api
    B = 16
    ctx = api.WhisperContext.new(model_path("base"), max_segments=B, max_clips=B, decode_capacity=False)
    clips = pkg.synth.make_clips(B, first_seg=500)
    api.whisper_pcm_to_mel(ctx, clips)
    api.whisper_encode(ctx, 1, [0] * B, clip_ids=list(range(B)))
    first = [ctx.encoder_out(s).copy() for s in range(B)]
    dig1 = ctx.encoder_digest(B).copy()
    kv1 = [np.stack(ctx.cross_kv(s, ctx.n_text_layer - 1)) for s in (0, 7, 15)]
    orc = pyoracle.Oracle(model_path("base"))
    for i, s in enumerate((0, 7, 15)):
        orc.pcm_to_mel(clips[s])
        ref = orc.encode(0)
        assert np.isfinite(first[s]).all()
        assert rel_l2(first[s], ref) < ENC_TOL, (s, rel_l2(first[s], ref))
        rk, rv = orc.cross_kv(orc.n_text_layer - 1)
        assert rel_l2(kv1[i][0], rk) < ENC_TOL and rel_l2(kv1[i][1], rv) < ENC_TOL, s
        # the digest the bench reads back is sum|x| of this output
        assert abs(dig1[s] - np.abs(ref.astype(np.float64)).sum()) < 2e-3 * np.abs(ref).sum()
    for rep in range(2):   # again, from the mel: identical bits
        api.whisper_pcm_to_mel(ctx, clips)
        api.whisper_encode(ctx, 1, [0] * B, clip_ids=list(range(B)))
        for s in range(B):
            assert np.array_equal(ctx.encoder_out(s), first[s]), (rep, s)
        assert np.array_equal(ctx.encoder_digest(B), dig1), rep
        for i, s in enumerate((0, 7, 15)):
            assert np.array_equal(np.stack(ctx.cross_kv(s, ctx.n_text_layer - 1)), kv1[i]), (rep, s)
    ctx.close()


@pytest.mark.parametrize("arch", ["small", "medium", "large-v3"])
def test_full_depth_encoder_vs_oracle(pkg, pyoracle, model_path, arch):
    """configs[2..4] architectures at FULL depth (12 / 24 / 32 layers), one 30 s clip: ln_post output and the last
    text layer's cross K / V against the oracle (tools/ln_fold_check.py of round 1, now in the suite)."""
    from whisper_rs_b200 import api
    ctx = api.WhisperContext.new(model_path(arch), max_segments=1, max_clips=1, decode_capacity=False)
    pcm = pkg.synth.make_segment(11)
    api.whisper_pcm_to_mel(ctx, pcm)
    api.whisper_encode(ctx, 1, 0)
    got = ctx.encoder_out(0)
    k, v = ctx.cross_kv(0, ctx.n_text_layer - 1)
    ctx.close()
    orc = pyoracle.Oracle(model_path(arch))
    orc.pcm_to_mel(pcm)
    ref = orc.encode(0)
    rk, rv = orc.cross_kv(orc.n_text_layer - 1)
    orc.close()
    assert np.isfinite(got).all()
    assert rel_l2(got, ref) < ENC_TOL, (arch, rel_l2(got, ref))
    assert rel_l2(k, rk) < ENC_TOL and rel_l2(v, rv) < ENC_TOL, arch


def test_layernorm_fold_rows_with_large_mean(pkg, pyoracle, model_path, tmp_path):
    """The LayerNorm fold (gemm2.cu) on a residual stream whose rows have mean ~50 and spread ~0.1 -- the case
    where rounding x to F16 before the mean is removed, or forming the variance as E[x^2] - mu^2 in f32, loses
    the signal.  The positional embedding is 50 + 0.1 * sinusoids and the conv stem / block outputs are scaled down,
    so every LayerNorm of the encoder sees such rows; the oracle normalises with f64 statistics."""
    from whisper_rs_b200 import api
    hp = pkg.ggml_file.ARCHS["tiny"]
    t = {}
    for name, a in pkg.ggml_file.random_tensors(hp, 77):
        if name == "encoder.conv2.weight":
            a = (a.astype(np.float32) * 0.05).astype(a.dtype)
        elif name == "encoder.positional_embedding":
            a = (50.0 + a * 0.1).astype(np.float32)
        elif name.startswith("encoder.blocks.") and (name.endswith("attn.out.weight") or name.endswith("mlp.2.weight")):
            a = (a.astype(np.float32) * 0.05).astype(a.dtype)        # the blocks keep the rows near mean 50 / spread 0.1
        t[name] = a
    path = str(tmp_path / "ggml-tiny-mean50.bin")
    pkg.ggml_file.write_model(path, hp, 0, tensors=t)
    pcm = pkg.synth.make_segment(5)
    orc = pyoracle.Oracle(path)
    orc.pcm_to_mel(pcm)
    ref = orc.encode(0)
    # the rows really are of that kind (residual stream after the last block, before ln_post)
    from whisper_rs_b200 import cabi
    ctx = api.WhisperContext.new(path, max_segments=1, decode_capacity=False, checkpoints=True)
    api.whisper_pcm_to_mel(ctx, pcm)
    api.whisper_encode(ctx, 1, 0)
    mean_abs = ctx.checksum(cabi.STAGE_LAYER, hp.n_audio_layer - 1, 0) / (hp.n_audio_ctx * hp.n_audio_state)
    assert 45.0 < mean_abs < 55.0, mean_abs
    got = ctx.encoder_out(0)
    assert np.isfinite(got).all()
    assert rel_l2(got, ref) < ENC_TOL, rel_l2(got, ref)
    k, v = ctx.cross_kv(0, hp.n_text_layer - 1)
    rk, rv = orc.cross_kv(hp.n_text_layer - 1)
    assert rel_l2(k, rk) < ENC_TOL and rel_l2(v, rv) < ENC_TOL
    ctx.close()


# ------------------------------------------------------------------------------------------------
def _non_collapsing_model(pkg, arch, path):
    """Random-init weights whose greedy decode does not settle on one token: token embedding x12 and decoder
    positional embedding x400 (sigma 0.24 / 4), so the hidden state -- and with it the arg-max -- changes from
    position to position and a wrong KV-cache row changes the result."""
    hp = pkg.ggml_file.ARCHS[arch]
    t = {}
    for name, a in pkg.ggml_file.random_tensors(hp, pkg.ggml_file.arch_seed(arch)):
        if name == "decoder.token_embedding.weight":
            a = (a.astype(np.float32) * 12.0).astype(a.dtype)
        elif name == "decoder.positional_embedding":
            a = (a * 400.0).astype(np.float32)
        t[name] = a
    pkg.ggml_file.write_model(path, hp, 0, tensors=t)
    return hp


def test_small_full_depth_decode_to_224_and_deep_positions(pkg, pyoracle, tmp_path):
    """configs[2]: whisper small at full depth.  (1) 224-step free-running greedy decode against the oracle on
    weights that do not collapse to one token (>= 20 distinct ids); (2) teacher-forced on the oracle's tokens,
    single-token steps (the CUDA-graph path's kernels) all the way to position 446: logits against the oracle at
    n_past in {1, 33, 64, 65, 128, 223, 446} and the arg-max at EVERY position whose oracle margin exceeds the
    tolerance."""
    from whisper_rs_b200 import api
    path = str(tmp_path / "ggml-small-nc.bin")
    hp = _non_collapsing_model(pkg, "small", path)
    pcm = pkg.synth.make_segment(3)
    orc = pyoracle.Oracle(path)
    orc.pcm_to_mel(pcm)
    orc.encode(0)
    ctx = api.WhisperContext.new(path, max_segments=2, max_clips=2)
    api.whisper_pcm_to_mel(ctx, np.stack([pcm, pkg.synth.make_segment(4)]))
    api.whisper_encode(ctx, 1, [0, 0], clip_ids=[0, 1])
    prompt = [ctx.token_sot]
    # ---- (1) free run, 224 tokens, two sequences in the batch (sequence 0 is the oracle's clip)
    rt, rm = orc.decode_greedy(prompt, 224, eot=-1, n_threads=1)
    toks, marg, lens = api.whisper_decode_greedy(ctx, prompt, 224, n_seqs=2, eot=-1)
    assert len(rt) == 224 and lens[0] == 224 and lens[1] == 224
    assert len(set(rt.tolist())) >= 20, len(set(rt.tolist()))          # the weights do what they were made for
    agree = _check_greedy(toks[0], lens[0], rt, rm)
    assert agree >= 20, agree
    assert len(set(toks[0][:agree].tolist())) >= 10
    assert not np.array_equal(toks[0], toks[1])                         # the other clip decodes differently
    assert np.abs(marg[0][:agree] - rm[:agree]).max() < 5e-2
    # ---- (2) teacher-forced, one token per call, to the end of the text context
    n_ctx = hp.n_text_ctx                                               # 448
    seq = np.concatenate([prompt, rt, rt])[: n_ctx - 1].astype(np.int32)   # 447 tokens: positions 0 .. 446
    probe = (1, 33, 64, 65, 128, 223, 446)
    ref_logits, lo = {}, 0
    for p in probe:                                                     # the oracle advances chunk by chunk (KV cache)
        ref_logits[p] = orc.decode(seq[lo:p + 1], lo).copy()
        lo = p + 1
    bad = []
    for p in range(len(seq)):
        api.whisper_decode(ctx, np.array([[seq[p]], [seq[p]]], dtype=np.int32), p)
        if p in ref_logits:
            got = ctx.logits(0)
            assert rel_l2(got, ref_logits[p]) < LOGIT_TOL, (p, rel_l2(got, ref_logits[p]))
        if 1 <= p <= 223 and rm[p] > MARGIN_TOL:                        # position p's arg-max is the oracle's token p
            if int(np.argmax(ctx.logits(0))) != int(rt[p]):
                bad.append(p)
    assert not bad, bad
    ctx.close()


# ------------------------------------------------------------------------------------------------
def test_audio_ctx_full_to_short_with_several_segments(pkg, pyoracle, model_path):
    """exp_n_audio_ctx going DOWN with more than one segment: the zero pad rows of the conv stem's buffers move onto
    memory that held activations of the previous layout.  Then back up."""
    from whisper_rs_b200 import api
    arch, short = "tiny", 500
    hp = pkg.ggml_file.ARCHS[arch]
    n = 2 * hp.n_audio_ctx * 160
    pcm = pkg.synth.make_segment(91, n, silent_tail_s=0.1)
    ctx = api.WhisperContext.new(model_path(arch), max_segments=3, max_clips=1, max_clip_samples=n, decode_capacity=False)
    orc = pyoracle.Oracle(model_path(arch))
    api.whisper_pcm_to_mel(ctx, pcm)
    orc.pcm_to_mel(pcm)
    api.whisper_encode(ctx, 1, [0, 0, 0], clip_ids=[0, 0, 0])           # full context: every buffer row written
    full_ref = orc.encode(0)
    assert rel_l2(ctx.encoder_out(2), full_ref) < ENC_TOL
    ctx.set_audio_ctx(short)
    orc.set_audio_ctx(short)
    offs = [0, 2 * short, 4 * short]
    api.whisper_encode(ctx, 1, offs, clip_ids=[0, 0, 0])
    for s, off in enumerate(offs):
        ref = orc.encode(off)
        assert rel_l2(ctx.encoder_out(s), ref) < ENC_TOL, (s, rel_l2(ctx.encoder_out(s), ref))
        # the window edges are where a stale pad row shows
        assert rel_l2(ctx.encoder_out(s)[:4], ref[:4]) < ENC_TOL and rel_l2(ctx.encoder_out(s)[-4:], ref[-4:]) < ENC_TOL, s
    ctx.set_audio_ctx(0)
    orc.set_audio_ctx(0)
    api.whisper_encode(ctx, 1, [0, 0], clip_ids=[0, 0])
    assert rel_l2(ctx.encoder_out(1), full_ref) < ENC_TOL
    ctx.close()


def test_greedy_after_audio_ctx_change_and_output_strides(pkg, pyoracle, model_path, golden):
    """(a) wb_set_audio_ctx between two wb_decode_greedy calls of the same batch shape: the captured step graph
    bakes the audio context into its cross-attention launches and must be re-captured.
    (b) max_new larger than n_text_ctx with several sequences: rows keep the caller's stride."""
    from whisper_rs_b200 import api
    arch = "micro"
    hp = pkg.ggml_file.ARCHS[arch]
    n = 2 * hp.n_audio_ctx * 160
    clips = np.stack([pkg.synth.make_segment(70 + s, n, 0.1) for s in range(2)])
    ctx = api.WhisperContext.new(model_path(arch), max_segments=2, max_clips=2, max_clip_samples=n)
    orcs = [pyoracle.Oracle(model_path(arch)) for _ in range(2)]
    eot = hp.n_vocab - 1
    api.whisper_pcm_to_mel(ctx, clips)
    for s in range(2):
        orcs[s].pcm_to_mel(clips[s])
    for n_ctx in (0, 40, 0, 24):
        ctx.set_audio_ctx(n_ctx)
        api.whisper_encode(ctx, 1, [0, 0], clip_ids=[0, 1])
        toks, marg, lens = api.whisper_decode_greedy(ctx, [7], 12, n_seqs=2, eot=eot)
        agree = 0
        for s in range(2):
            orcs[s].set_audio_ctx(n_ctx)
            orcs[s].encode(0)
            rt, rm = orcs[s].decode_greedy([7], 12, eot=eot)
            agree += _check_greedy(toks[s], lens[s], rt, rm)
        assert agree >= 12, (n_ctx, agree)
    # (b) n_text_ctx = 32 for micro; ask for 50 tokens for 2 sequences
    ctx.set_audio_ctx(0)
    api.whisper_encode(ctx, 1, [0, 0], clip_ids=[0, 1])
    big = 50
    assert big > hp.n_text_ctx
    toks, marg, lens = api.whisper_decode_greedy(ctx, [7], big, n_seqs=2, eot=-1)
    small_t, _, small_l = api.whisper_decode_greedy(ctx, [7], hp.n_text_ctx, n_seqs=2, eot=-1)
    for s in range(2):
        assert lens[s] == small_l[s] <= hp.n_text_ctx
        assert np.array_equal(toks[s][:lens[s]], small_t[s][:lens[s]]), s
        assert np.all(toks[s][hp.n_text_ctx:] == 0)
    ctx.close()


def test_unconsumed_prefetch_is_dropped(pkg, pyoracle, model_path):
    """wb_pcm_prefetch(P) that is NOT consumed by the next upload must not be picked up by a later
    wb_pcm_to_mel(P) after P's contents changed."""
    import torch
    from whisper_rs_b200 import api
    n = 16000
    ctx = api.WhisperContext.new(model_path("micro"), max_segments=1, max_clips=1, max_clip_samples=n, decode_capacity=False)
    a, b, c = (pkg.synth.make_segment(80 + i, n, 0.1) for i in range(3))
    P = torch.from_numpy(a.copy()).pin_memory()
    Q = torch.from_numpy(b.copy()).pin_memory()
    api.whisper_pcm_prefetch_ptr(ctx, P.data_ptr(), n * 4)
    api.whisper_pcm_to_mel_ptr(ctx, P.data_ptr(), n, 1)            # consumed: the prefetched copy is used
    ctx.sync()
    mel_a = ctx.mel(0).copy()
    api.whisper_pcm_prefetch_ptr(ctx, P.data_ptr(), n * 4)         # prefetch P (= a) ...
    api.whisper_pcm_to_mel_ptr(ctx, Q.data_ptr(), n, 1)            # ... but upload Q instead
    ctx.sync()
    mel_b = ctx.mel(0).copy()
    P.copy_(torch.from_numpy(c))                                   # refill P
    api.whisper_pcm_to_mel_ptr(ctx, P.data_ptr(), n, 1)            # must see c, not the stale prefetch of a
    ctx.sync()
    mel_c = ctx.mel(0).copy()
    orc = pyoracle.Oracle(model_path("micro"))
    from conftest import mel_close
    assert mel_close(mel_a, orc.pcm_to_mel(a)) and mel_close(mel_b, orc.pcm_to_mel(b)) and mel_close(mel_c, orc.pcm_to_mel(c))
    assert not np.array_equal(mel_c, mel_a)
    ctx.close()


def test_norm_scope_segment(pkg, pyoracle, model_path):
    """WB_NORM_SEGMENT: every encoder window is clamped against its own maximum.  For the last window of a clip this
    is what the reference computes for a clip made of that window alone; WB_NORM_CLIP (the reference) differs when
    the whole-clip maximum lives in another window."""
    from whisper_rs_b200 import api, cabi
    arch = "micro"
    hp = pkg.ggml_file.ARCHS[arch]
    fpw = 2 * hp.n_audio_ctx
    n = 2 * fpw * 160
    pcm = pkg.synth.make_segment(95, n, silent_tail_s=0.05)
    pcm[n // 2:] *= 0.003                                         # window 1 is 50 dB quieter than window 0
    orc = pyoracle.Oracle(model_path(arch))
    orc.pcm_to_mel(pcm[n // 2:])
    ref_seg = orc.encode(0)                                        # window 1 as a clip of its own
    orc.pcm_to_mel(pcm)
    ref_clip = orc.encode(fpw)                                     # window 1 of the whole clip (reference behaviour)
    assert rel_l2(ref_seg, ref_clip) > 5e-2                        # the scopes really differ for this clip
    for scope, ref in ((cabi.NORM_SEGMENT, ref_seg), (cabi.NORM_CLIP, ref_clip)):
        ctx = api.WhisperContext.new(model_path(arch), max_segments=2, max_clips=1, max_clip_samples=n,
                                     decode_capacity=False, norm_scope=scope)
        api.whisper_pcm_to_mel(ctx, pcm)
        api.whisper_encode(ctx, 1, [0, fpw], clip_ids=[0, 0])
        assert rel_l2(ctx.encoder_out(1), ref) < ENC_TOL, (scope, rel_l2(ctx.encoder_out(1), ref))
        ctx.close()


@pytest.mark.parametrize("arch", ["micro", "tiny"])
def test_no_kernel_writes_outside_its_buffers(pkg, model_path, arch):
    """compute-sanitizer is closed on this GPU pool (profiles/r02_compute_sanitizer.txt), so the library carries its own
    check: with canary=True every device buffer sits between two 256-byte guard zones, and after the whole hot path
    -- mel, encode at two audio contexts with ragged batches, prompt pass, greedy steps -- every guard zone must still
    hold its fill pattern."""
    from whisper_rs_b200 import api
    hp = pkg.ggml_file.ARCHS[arch]
    n = 2 * hp.n_audio_ctx * 160
    S = 3
    ctx = api.WhisperContext.new(model_path(arch), max_segments=S, max_clips=S, max_clip_samples=n, canary=True, checkpoints=True)
    assert ctx.canary_check() == 0
    clips = np.stack([pkg.synth.make_segment(120 + s, n, 0.1) for s in range(S)])
    eot = hp.n_vocab - 1
    for n_ctx, n_seg in ((0, S), (hp.n_audio_ctx // 3, 2), (0, 1), (hp.n_audio_ctx - 1, S)):
        ctx.set_audio_ctx(n_ctx)
        api.whisper_pcm_to_mel(ctx, clips)
        api.whisper_encode(ctx, 1, [0] * n_seg, clip_ids=list(range(n_seg)))
        ctx.encoder_digest(n_seg)
        api.whisper_decode(ctx, np.tile(np.arange(1, 6, dtype=np.int32), (n_seg, 1)), 0)
        api.whisper_decode_greedy(ctx, [7], 9, n_seqs=n_seg, eot=eot)
        assert ctx.canary_check() == 0, (n_ctx, n_seg)
    api.whisper_pcm_to_mel(ctx, (clips[:2, : n // 2 + 77] * 32767).astype(np.int16))    # i16 ingest, ragged length
    ctx.mel(1)
    assert ctx.canary_check() == 0
    ctx.close()
    plain = api.WhisperContext.new(model_path(arch), max_segments=1, decode_capacity=False)
    assert plain.canary_check() == -1                     # no guards unless asked for
    plain.close()


def test_decoder_bit_reproducible(pkg, model_path):
    """The decode step carries no floating-point atomics either (the folded LayerNorm statistics are accumulated as
    64-bit fixed-point integers): logits, greedy ids and margins of repeated runs are identical bit for bit -- which is
    what lets a clip split over several GPUs reproduce the single-GPU ids exactly (tools/run_config5.py)."""
    from whisper_rs_b200 import api
    B = 5
    ctx = api.WhisperContext.new(model_path("tiny"), max_segments=B, max_clips=B)
    clips = pkg.synth.make_clips(B, first_seg=640)
    api.whisper_pcm_to_mel(ctx, clips)
    api.whisper_encode(ctx, 1, [0] * B, clip_ids=list(range(B)))
    prompt = np.tile(np.array([ctx.token_sot, 11, 12], dtype=np.int32), (B, 1))
    ref_logits, ref_greedy = None, None
    for rep in range(3):
        api.whisper_decode(ctx, prompt, 0)                        # many-row prompt pass (LayerNorm kernel path)
        api.whisper_decode(ctx, prompt[:, :1], 3)                 # single-token step (folded statistics)
        lg = np.stack([ctx.logits(s) for s in range(B)])
        toks, marg, lens = api.whisper_decode_greedy(ctx, [ctx.token_sot], 40, n_seqs=B, eot=-1)
        if rep == 0:
            ref_logits, ref_greedy = lg, (toks.copy(), marg.copy(), lens.copy())
        else:
            assert np.array_equal(lg, ref_logits), rep
            assert np.array_equal(toks, ref_greedy[0]) and np.array_equal(marg, ref_greedy[1]) and np.array_equal(lens, ref_greedy[2]), rep
    ctx.close()


def test_decoder_layernorm_fold_rows_with_large_mean(pkg, pyoracle, tmp_path):
    """The decoder's folded LayerNorms on rows of mean ~50 and spread ~0.03 (decoder positional embedding + 50, small
    token embedding): the single-token step keeps an F16 copy of x - centre and exact fixed-point statistics of it, so
    the logits still match the oracle's f64 LayerNorm (rounding x itself to F16 -- ulp 0.03 at 50 -- would not)."""
    from whisper_rs_b200 import api
    hp = pkg.ggml_file.ARCHS["tiny"]
    t = {}
    for name, a in pkg.ggml_file.random_tensors(hp, 78):
        if name == "decoder.positional_embedding":
            a = (50.0 + a * 3.0).astype(np.float32)                      # mean 50, spread 0.03
        elif name.startswith("decoder.blocks.") and (name.endswith("out.weight") or name.endswith("mlp.2.weight")):
            a = (a.astype(np.float32) * 0.05).astype(a.dtype)             # the blocks keep the rows near mean 50
        t[name] = a
    path = str(tmp_path / "ggml-tiny-dec-mean50.bin")
    pkg.ggml_file.write_model(path, hp, 0, tensors=t)
    pcm = pkg.synth.make_segment(6)
    orc = pyoracle.Oracle(path)
    orc.pcm_to_mel(pcm)
    orc.encode(0)
    ctx = api.WhisperContext.new(path, max_segments=2, max_clips=2)
    api.whisper_pcm_to_mel(ctx, np.stack([pcm, pcm]))
    api.whisper_encode(ctx, 1, [0, 0], clip_ids=[0, 1])
    seq = np.array([ctx.token_sot, 17, 900, 3, 64, 511, 5, 2048], dtype=np.int32)
    for p in range(len(seq)):                                             # single-token steps: the folded path
        api.whisper_decode(ctx, np.array([[seq[p]], [seq[p]]], dtype=np.int32), p)
        ref = orc.decode(seq[p:p + 1], p)
        got = ctx.logits(1)
        assert np.isfinite(got).all()
        assert rel_l2(got, ref) < LOGIT_TOL, (p, rel_l2(got, ref))
    ctx.close()


def test_encode_graph_cache_over_alternating_shapes(pkg, pyoracle, model_path):
    """wb_encode replays a captured CUDA graph per shape (n_seg, audio context, mel length); alternating between shapes
    -- full batches and a ragged tail, two audio contexts, two clip lengths -- must keep returning, bit for bit, what
    the first (directly launched) call of each shape returned, and the oracle's result."""
    from whisper_rs_b200 import api
    arch = "micro"
    hp = pkg.ggml_file.ARCHS[arch]
    n = 2 * hp.n_audio_ctx * 160
    ctx = api.WhisperContext.new(model_path(arch), max_segments=3, max_clips=3, max_clip_samples=n, decode_capacity=False)
    clips = np.stack([pkg.synth.make_segment(130 + s, n, 0.1) for s in range(3)])
    shapes = [(3, 0, n), (1, 0, n), (2, 40, n), (2, 0, n // 2), (3, 24, n)]        # (n_seg, audio ctx, samples): 5 > the cache of 4
    first = {}
    launches0 = ctx.timings()["n_kernel_launches"]
    for rep in range(4):
        for sh in shapes:
            n_seg, n_ctx, ns = sh
            ctx.set_audio_ctx(n_ctx)
            api.whisper_pcm_to_mel(ctx, clips[:, :ns])
            api.whisper_encode(ctx, 1, [0] * n_seg, clip_ids=list(range(n_seg)))
            got = np.stack([ctx.encoder_out(s) for s in range(n_seg)])
            if rep == 0:
                first[sh] = got
            else:
                assert np.array_equal(got, first[sh]), (rep, sh)
    assert ctx.timings()["n_kernel_launches"] > launches0
    orc = pyoracle.Oracle(model_path(arch))
    orc.pcm_to_mel(clips[1])
    assert rel_l2(first[(3, 0, n)][1], orc.encode(0)) < ENC_TOL
    ctx.close()
