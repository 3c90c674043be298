"""GPU: whisper_encode (src/main.rs:1799-2063) against the oracle on identical synthetic audio and
identical random-init weights.  Tolerances (north_star): encoder outputs <= 1e-2 relative L2."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu
ENC_TOL = 1e-2   # BASELINE.json north_star: "encoder outputs within ... <= 1e-2 relative L2"


def _ctx(model_path, arch, **kw):
    from whisper_rs_b200 import api
    return api.WhisperContext.new(model_path(arch), decode_capacity=False, checkpoints=True, **kw)


def test_micro_encoder_vs_oracle_and_golden(pkg, pyoracle, model_path, golden):
    from whisper_rs_b200 import api, cabi
    ctx = _ctx(model_path, "micro", max_segments=2)
    orc = pyoracle.Oracle(model_path("micro"))
    pcm = pkg.synth.make_segment(0, int(golden["n_samples"]), silent_tail_s=0.2)
    api.whisper_pcm_to_mel(ctx, pcm)
    api.whisper_encode(ctx, 1, 0)
    got = ctx.encoder_out(0)
    orc.pcm_to_mel(pcm)
    ref = orc.encode(0)
    assert np.isfinite(got).all()
    assert rel_l2(got, ref) < ENC_TOL, rel_l2(got, ref)
    assert rel_l2(got, golden["enc_oracle"]) < ENC_TOL          # committed fixture
    assert rel_l2(got, golden["enc_hf"]) < ENC_TOL              # independent fp32 implementation
    # per-stage sum|x| probes localise a divergence (src/main.rs:1836-1849, 1998-2010)
    stages = [(cabi.STAGE_CONV1, 0), (cabi.STAGE_CONV2_POS, 0)] + \
             [(cabi.STAGE_LAYER, i) for i in range(ctx.n_audio_layer)] + [(cabi.STAGE_LN_POST, 0)] + \
             [(cabi.STAGE_CROSS_K, i) for i in range(ctx.n_text_layer)] + \
             [(cabi.STAGE_CROSS_V, i) for i in range(ctx.n_text_layer)]
    for st, layer in stages:
        a, b = ctx.checksum(st, layer, 0), orc.checksum(st, layer)
        assert abs(a - b) <= 2e-3 * abs(b), (st, layer, a, b)
    for layer in range(ctx.n_text_layer):
        k, v = ctx.cross_kv(0, layer)
        rk, rv = orc.cross_kv(layer)
        assert rel_l2(k, rk) < ENC_TOL and rel_l2(v, rv) < ENC_TOL
    ctx.close()


def test_micro_window_offsets_and_batch(pkg, pyoracle, model_path):
    """Batched segments = independent reference calls; windows past the clip end are zero (1822-1829)."""
    from whisper_rs_b200 import api
    ctx = _ctx(model_path, "micro", max_segments=3, max_clips=2, max_clip_samples=160 * 96 * 5)
    orc = pyoracle.Oracle(model_path("micro"))
    n_ctx = ctx.n_audio_ctx
    clips = np.stack([pkg.synth.make_segment(s, 160 * n_ctx * 5, 0.2) for s in (11, 12)])
    api.whisper_pcm_to_mel(ctx, clips)
    segs = [(0, 0), (1, 2 * n_ctx), (0, 4 * n_ctx)]          # (clip, mel_offset); the last is half past the end
    api.whisper_encode(ctx, 1, [o for _, o in segs], clip_ids=[c for c, _ in segs])
    for i, (c, off) in enumerate(segs):
        orc.pcm_to_mel(clips[c])
        ref = orc.encode(off)
        assert rel_l2(ctx.encoder_out(i), ref) < ENC_TOL, i
    ctx.close()


@pytest.mark.parametrize("arch", ["tiny", "base"])
def test_encoder_30s_vs_oracle(pkg, pyoracle, model_path, arch):
    """configs[0] / configs[1] architectures, one 30 s clip (1500-token context, ragged 128-row tiles)."""
    from whisper_rs_b200 import api
    ctx = _ctx(model_path, arch, max_segments=2, max_clips=2)
    orc = pyoracle.Oracle(model_path(arch))
    clips = pkg.synth.make_clips(2)
    api.whisper_pcm_to_mel(ctx, clips)
    api.whisper_encode(ctx, 1, [0, 0], clip_ids=[0, 1])
    for s in range(2 if arch == "tiny" else 1):
        orc.pcm_to_mel(clips[s])
        ref = orc.encode(0)
        got = ctx.encoder_out(s)
        assert np.isfinite(got).all()
        assert rel_l2(got, ref) < ENC_TOL, (arch, s, rel_l2(got, ref))
        k, v = ctx.cross_kv(s, ctx.n_text_layer - 1)
        rk, rv = orc.cross_kv(orc.n_text_layer - 1)
        assert rel_l2(k, rk) < ENC_TOL and rel_l2(v, rv) < ENC_TOL
    ctx.close()


def test_capacity_errors(pkg, model_path):
    from whisper_rs_b200 import api
    ctx = _ctx(model_path, "micro", max_segments=1)
    with pytest.raises(api.WsError) as e:
        api.whisper_encode(ctx, 1, 0)                 # no mel yet
    assert e.value.variant == "Unexpected"
    api.whisper_pcm_to_mel(ctx, pkg.synth.make_segment(0, 160 * 200, 0.1))
    with pytest.raises(api.WsError) as e:
        api.whisper_encode(ctx, 1, [0, 0])            # more segments than the context was sized for
    assert e.value.variant == "NotEnoughSpace"
    with pytest.raises(api.WsError) as e:
        api.whisper_pcm_to_mel(ctx, np.zeros(480001 * 2, np.float32))
    assert e.value.variant == "NotEnoughSpace"
    ctx.close()


def test_vocab_text_and_placeholder_names(pkg, model_path, tmp_path):
    """id_to_token (src/main.rs:544) and the names the reference gives ids the file has no text for
    (442-467), on a multilingual-size vocabulary whose file holds fewer entries than hparams.n_vocab."""
    from whisper_rs_b200 import api
    import dataclasses
    hp = dataclasses.replace(pkg.ggml_file.ARCHS["micro"], n_vocab=51865)
    path = str(tmp_path / "ggml-micro-ml.bin")
    pkg.ggml_file.write_model(path, hp, seed=3, n_vocab_file=50257)
    ctx = api.WhisperContext.new(path, max_segments=1, decode_capacity=False)
    assert ctx.token_eot == 50257 and ctx.token_sot == 50258 and ctx.token_beg == 50364   # +1: multilingual (433-440)
    assert ctx.token_text(0) == b"t0" and ctx.token_text(50256) == b"t50256"
    assert ctx.token_text(ctx.token_eot) == b"[_EOT_]"
    assert ctx.token_text(ctx.token_sot) == b"[_SOT_]"
    assert ctx.token_text(ctx.token_prev) == b"[_PREV_]"
    assert ctx.token_text(ctx.token_not) == b"[_NOT_]"
    assert ctx.token_text(ctx.token_beg) == b"[_BEG_]"
    assert ctx.token_text(ctx.token_beg + 5) == b"[_TT_5]"
    assert ctx.token_text(50300) == b"[_extra_token_50300]"
    assert ctx.tokens_to_text([ctx.token_sot, 5, 17, ctx.token_beg + 3, 9, ctx.token_eot]) == b"t5t17t9"
    with pytest.raises(api.WsError):
        ctx.token_text(51865)
    ctx.close()


@pytest.mark.parametrize("arch,n_ctx", [("micro", 40), ("tiny", 700)])
def test_exp_n_audio_ctx(pkg, pyoracle, model_path, arch, n_ctx):
    """exp_n_audio_ctx (src/main.rs:362, 1803-1807): a shorter audio context -- window of 2 n_ctx frames, the first
    n_ctx rows of the positional embedding -- through mel + encode + cross K/V + logits, against the oracle with the
    same setting; back to the model's context afterwards."""
    from whisper_rs_b200 import api
    hp = pkg.ggml_file.ARCHS[arch]
    n = 2 * hp.n_audio_ctx * 160
    pcm = pkg.synth.make_segment(77, n, silent_tail_s=0.1)
    ctx = api.WhisperContext.new(model_path(arch), max_segments=2, max_clips=1, max_clip_samples=n)
    orc = pyoracle.Oracle(model_path(arch))
    api.whisper_pcm_to_mel(ctx, pcm)
    orc.pcm_to_mel(pcm)
    with pytest.raises(api.WsError):
        ctx.set_audio_ctx(hp.n_audio_ctx + 1)
    ctx.set_audio_ctx(n_ctx)
    orc.set_audio_ctx(n_ctx)
    offs = [0, 2 * n_ctx]                                # two consecutive short windows in one batch
    api.whisper_encode(ctx, 1, offs, clip_ids=[0, 0])
    for s, off in enumerate(offs):
        ref = orc.encode(off)
        assert ref.shape == (n_ctx, hp.n_audio_state)
        assert rel_l2(ctx.encoder_out(s), ref) < 1e-2, (s, rel_l2(ctx.encoder_out(s), ref))
        k, v = ctx.cross_kv(s, hp.n_text_layer - 1)
        rk, rv = orc.cross_kv(hp.n_text_layer - 1)
        assert k.shape == (n_ctx, hp.n_text_state) and rel_l2(k, rk) < 1e-2 and rel_l2(v, rv) < 1e-2
    # the decoder attends to the n_ctx rows of the last encode (sequence 1 <-> the oracle's last window)
    toks = np.array([[3, 5, 8], [3, 5, 8]], dtype=np.int32)
    api.whisper_decode(ctx, toks, 0)
    assert rel_l2(ctx.logits(1), orc.decode(toks[1], 0)) < 1e-2
    ctx.set_audio_ctx(0)
    orc.set_audio_ctx(0)
    api.whisper_encode(ctx, 1, 0)
    assert rel_l2(ctx.encoder_out(0), orc.encode(0)) < 1e-2
    ctx.close()
