"""GPU: the decode step (absent in the reference) against the oracle's
restatement of upstream semantics.  north_star: greedy token ids bit-exact wherever the
reference's top-1 logit margin exceeds the tolerance."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu
LOGIT_TOL = 1e-2        # relative L2 on logits (same bar as the encoder outputs)
MARGIN_TOL = 2e-2       # absolute logit margin below which a greedy flip is allowed


def _setup(pkg, pyoracle, model_path, arch, n_seq, n_samples):
    from whisper_rs_b200 import api
    ctx = api.WhisperContext.new(model_path(arch), max_segments=n_seq, max_clips=n_seq, max_clip_samples=n_samples)
    clips = np.stack([pkg.synth.make_segment(20 + s, n_samples, 0.2) for s in range(n_seq)])
    api.whisper_pcm_to_mel(ctx, clips)
    api.whisper_encode(ctx, 1, [0] * n_seq, clip_ids=list(range(n_seq)))
    orcs = []
    for s in range(n_seq):
        o = pyoracle.Oracle(model_path(arch))
        o.pcm_to_mel(clips[s])
        o.encode(0)
        orcs.append(o)
    return ctx, orcs


def test_micro_decode_logits_prompt_and_incremental(pkg, pyoracle, model_path, golden):
    from whisper_rs_b200 import api
    ctx, orcs = _setup(pkg, pyoracle, model_path, "micro", 2, int(golden["n_samples"]))
    toks = np.array([[5, 17, 900, 3, 64, 511], [1, 2, 3, 4, 5, 6]], dtype=np.int32)
    api.whisper_decode(ctx, toks, 0)                        # prompt pass, 6 tokens, 2 sequences
    for s in range(2):
        ref = orcs[s].decode(toks[s], 0)
        assert rel_l2(ctx.logits(s), ref) < LOGIT_TOL, s
    # incremental: 4 tokens then 2 more at n_past = 4 must equal the batched pass (KV cache)
    full = [ctx.logits(s).copy() for s in range(2)]
    api.whisper_decode(ctx, toks[:, :4], 0)
    api.whisper_decode(ctx, toks[:, 4:], 4)
    for s in range(2):
        assert rel_l2(ctx.logits(s), full[s]) < 1e-3
    ctx.close()


def test_micro_long_prompt_many_rows(pkg, pyoracle, model_path, golden):
    """More than 32 rows (sequences x prompt tokens) take the tcgen05 swap-AB GEMM with a plain LayerNorm in front
    of the gamma-folded weights; up to 32 rows take the skinny kernel with the LayerNorm folded into its epilogue.
    Both must give the oracle's logits, and a step after either prompt pass must agree."""
    from whisper_rs_b200 import api
    ctx, orcs = _setup(pkg, pyoracle, model_path, "micro", 2, int(golden["n_samples"]))
    rng = np.random.default_rng(5)
    toks = rng.integers(0, ctx.n_vocab - 1, size=(2, 20)).astype(np.int32)      # 40 rows
    api.whisper_decode(ctx, toks, 0)
    for s in range(2):
        assert rel_l2(ctx.logits(s), orcs[s].decode(toks[s], 0)) < LOGIT_TOL, s
    nxt = np.array([[3], [9]], dtype=np.int32)
    api.whisper_decode(ctx, nxt, 20)                                            # 2 rows: folded path on that cache
    for s in range(2):
        assert rel_l2(ctx.logits(s), orcs[s].decode(nxt[s], 20)) < LOGIT_TOL, s
    ctx.close()


def test_micro_golden_logits(pkg, pyoracle, model_path, golden):
    from whisper_rs_b200 import api
    ctx, _ = _setup(pkg, pyoracle, model_path, "micro", 1, int(golden["n_samples"]))
    # same clip as the fixture (segment seed 0)
    api.whisper_pcm_to_mel(ctx, pkg.synth.make_segment(0, int(golden["n_samples"]), silent_tail_s=0.2))
    api.whisper_encode(ctx, 1, 0)
    api.whisper_decode(ctx, golden["tokens"], 0)
    assert rel_l2(ctx.logits(0), golden["logits_oracle"]) < LOGIT_TOL
    assert rel_l2(ctx.logits(0), golden["logits_hf"]) < LOGIT_TOL
    ctx.close()


def _check_greedy(got_tok, got_len, ref_tok, ref_margin):
    """ids equal up to the first step whose oracle margin is below MARGIN_TOL (after a legal flip
    the suffixes diverge by construction)."""
    n = min(len(ref_tok), int(got_len))
    for i in range(n):
        if got_tok[i] != ref_tok[i]:
            assert ref_margin[i] < MARGIN_TOL, (i, got_tok[i], ref_tok[i], ref_margin[i])
            return i
    assert int(got_len) == len(ref_tok)
    return n


def test_micro_greedy(pkg, pyoracle, model_path, golden):
    from whisper_rs_b200 import api
    ctx, orcs = _setup(pkg, pyoracle, model_path, "micro", 3, int(golden["n_samples"]))
    eot = ctx.n_vocab - 1
    toks, marg, lens = api.whisper_decode_greedy(ctx, [7], 20, n_seqs=3, eot=eot)
    agree = 0
    for s in range(3):
        rt, rm = orcs[s].decode_greedy([7], 20, eot=eot)
        agree += _check_greedy(toks[s], lens[s], rt, rm)
    assert agree >= 20          # not vacuous: most steps are compared
    # margins reported by the device loop are the oracle's margins
    rt, rm = orcs[0].decode_greedy([7], 20, eot=eot)
    k = _check_greedy(toks[0], lens[0], rt, rm)
    assert np.abs(marg[0][:k] - rm[:k]).max() < 5e-2
    # text context full: prompt 1 + at most n_text_ctx - 1 steps
    toks, marg, lens = api.whisper_decode_greedy(ctx, [7], 64, n_seqs=1, eot=eot)
    assert lens[0] <= ctx.n_text_ctx
    ctx.close()


def test_micro_greedy_more_than_32_sequences(pkg, pyoracle, model_path, golden):
    """40 sequences: a single-token step has more than 32 rows, so every linear takes the tcgen05 swap-AB GEMM
    behind a plain LayerNorm and the arg-max reads the full logits (no per-CTA top-2 partials)."""
    from whisper_rs_b200 import api
    n = int(golden["n_samples"])
    S = 40
    ctx = api.WhisperContext.new(model_path("micro"), max_segments=S, max_clips=S, max_clip_samples=n)
    clips = np.stack([pkg.synth.make_segment(200 + s, n, 0.2) for s in range(S)])
    api.whisper_pcm_to_mel(ctx, clips)
    api.whisper_encode(ctx, 1, [0] * S, clip_ids=list(range(S)))
    eot = ctx.n_vocab - 1
    toks, marg, lens = api.whisper_decode_greedy(ctx, [7], 6, n_seqs=S, eot=eot)
    agree = 0
    for s in (0, 17, 33, 39):
        o = pyoracle.Oracle(model_path("micro"))
        o.pcm_to_mel(clips[s])
        o.encode(0)
        rt, rm = o.decode_greedy([7], 6, eot=eot)
        agree += _check_greedy(toks[s], lens[s], rt, rm)
    assert agree >= 8
    ctx.close()


def test_greedy_stops_at_eot(pkg, pyoracle, model_path, golden):
    from whisper_rs_b200 import api
    ctx, orcs = _setup(pkg, pyoracle, model_path, "micro", 1, int(golden["n_samples"]))
    rt, rm = orcs[0].decode_greedy([7], 8, eot=-1)
    eot = int(rt[3])                                    # declare the 4th generated token to be eot
    toks, marg, lens = api.whisper_decode_greedy(ctx, [7], 8, n_seqs=1, eot=eot)
    rt2, _ = orcs[0].decode_greedy([7], 8, eot=eot)
    if np.all(rm[:4] > MARGIN_TOL):
        assert lens[0] == len(rt2) and toks[0][lens[0] - 1] == eot
    ctx.close()


def test_tiny_decode_30s(pkg, pyoracle, model_path):
    """configs[0]: whisper tiny, one 30 s clip, mel + encode + greedy decode."""
    from whisper_rs_b200 import api
    ctx, orcs = _setup(pkg, pyoracle, model_path, "tiny", 1, 480000)
    prompt = [ctx.token_sot]
    api.whisper_decode(ctx, prompt, 0)
    ref = orcs[0].decode(prompt, 0)
    assert rel_l2(ctx.logits(0), ref) < LOGIT_TOL
    toks, marg, lens = api.whisper_decode_greedy(ctx, prompt, 16, n_seqs=1)
    rt, rm = orcs[0].decode_greedy(prompt, 16)
    assert _check_greedy(toks[0], lens[0], rt, rm) >= 1
    ctx.close()


@pytest.mark.parametrize("arch", ["small.2l", "medium.2l", "large-v3.2l"])
def test_full_width_two_layer_pipeline(pkg, pyoracle, model_path, arch):
    """configs[2..4] widths (d = 768 / 1024 / 1280; 12 / 16 / 20 heads; large-v3's 128 mel bins and
    51866-entry vocabulary) at 2 + 2 layers: mel + encode + cross K/V + logits + greedy against the
    oracle on one 30 s clip, two sequences in the batch."""
    from whisper_rs_b200 import api
    ctx, orcs = _setup(pkg, pyoracle, model_path, arch, 2, 480000)
    assert ctx.n_mels == (128 if arch.startswith("large") else 80)
    for s in range(2):
        renc = orcs[s].encode(0)
        assert rel_l2(ctx.encoder_out(s), renc) < 1e-2, (arch, s)
        k, v = ctx.cross_kv(s, ctx.n_text_layer - 1)
        rk, rv = orcs[s].cross_kv(ctx.n_text_layer - 1)
        assert rel_l2(k, rk) < 1e-2 and rel_l2(v, rv) < 1e-2
    prompt = [ctx.token_sot]
    api.whisper_decode(ctx, np.array([prompt, prompt], dtype=np.int32), 0)
    for s in range(2):
        assert rel_l2(ctx.logits(s), orcs[s].decode(prompt, 0)) < LOGIT_TOL, (arch, s)
    toks, marg, lens = api.whisper_decode_greedy(ctx, prompt, 12, n_seqs=2)
    agree = 0
    for s in range(2):
        rt, rm = orcs[s].decode_greedy(prompt, 12)
        agree += _check_greedy(toks[s], lens[s], rt, rm)
    assert agree >= 2
    ctx.close()
