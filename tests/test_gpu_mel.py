"""GPU: log-mel front end against the oracle's restatement of src/main.rs:1554-1671.
Tolerance (north_star): mel features within 1e-4, in the well-posed form of conftest.mel_close."""
import numpy as np
import pytest

from conftest import mel_close, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(pkg, model_path):
    from whisper_rs_b200 import api
    c = api.WhisperContext.new(model_path("micro"), max_segments=2, max_clips=4, max_clip_samples=480000,
                               decode_capacity=False, checkpoints=True)
    yield c
    c.close()


def test_mel_one_clip_30s(pkg, pyoracle, model_path, ctx):
    from whisper_rs_b200 import api
    pcm = pkg.synth.make_segment(0)
    api.whisper_pcm_to_mel(ctx, pcm)
    got = ctx.mel(0)
    ref = pyoracle.Oracle(model_path("micro")).pcm_to_mel(pcm)
    assert got.shape == ref.shape == (80, 3000)
    assert mel_close(got, ref), (np.abs(got - ref).max(), rel_l2(got, ref))
    # the author's checksum probe (src/main.rs:1645-1647)
    assert abs(ctx.checksum(0, 0, 0) - np.abs(ref.astype(np.float64)).sum()) < 1e-4 * np.abs(ref).sum()


def test_mel_golden_f64(pkg, ctx, golden):
    from whisper_rs_b200 import api
    pcm = pkg.synth.make_segment(0, int(golden["n_samples"]), silent_tail_s=0.2)
    api.whisper_pcm_to_mel(ctx, pcm)
    assert mel_close(ctx.mel(0), golden["mel_f64"])


def test_mel_batch_of_clips_independent_max(pkg, pyoracle, model_path, ctx):
    """Each clip is normalised with its own whole-clip max (clamp_and_normalize, 1654-1671)."""
    from whisper_rs_b200 import api
    n = 160 * 500
    clips = np.stack([pkg.synth.make_segment(s, n, 0.3) * (0.05 if s == 1 else 1.0) for s in range(3)])
    api.whisper_pcm_to_mel(ctx, clips)
    orc = pyoracle.Oracle(model_path("micro"))
    for c in range(3):
        ref = orc.pcm_to_mel(clips[c])
        assert mel_close(ctx.mel(c), ref), c


def test_mel_edges(pkg, pyoracle, model_path, ctx):
    from whisper_rs_b200 import api
    orc = pyoracle.Oracle(model_path("micro"))
    for n in (1600, 1000, 400 + 160 * 17 + 33, 160 * 16, 160 * 16 + 1):   # ragged tails, partial last CTA
        pcm = pkg.synth.make_segment(9, n, 0.0)
        api.whisper_pcm_to_mel(ctx, pcm)
        ref = orc.pcm_to_mel(pcm)
        got = ctx.mel(0)
        assert got.shape == ref.shape, n
        assert mel_close(got, ref), n
    api.whisper_pcm_to_mel(ctx, np.zeros(3200, np.float32))   # silence -> -1.5 everywhere
    np.testing.assert_allclose(ctx.mel(0), -1.5, atol=1e-6)


def test_mel_int16_input(pkg, pyoracle, model_path, ctx):
    """convert_integer_to_float_audio (1673-1679) fused into the mel kernel."""
    from whisper_rs_b200 import api
    pcm = pkg.synth.make_segment(2, 160 * 300, 0.2)
    s16 = np.round(pcm * 32768.0).clip(-32768, 32767).astype(np.int16)
    api.whisper_pcm_to_mel(ctx, s16)
    ref = pyoracle.Oracle(model_path("micro")).pcm_to_mel(s16.astype(np.float32) / 32768.0)
    assert mel_close(ctx.mel(0), ref)


def test_mel_device_resident_input(pkg, ctx):
    import torch
    from whisper_rs_b200 import api
    pcm = pkg.synth.make_segment(4, 160 * 400, 0.2)
    api.whisper_pcm_to_mel(ctx, pcm)
    a = ctx.mel(0)
    api.whisper_pcm_to_mel(ctx, torch.from_numpy(pcm).cuda())
    np.testing.assert_array_equal(ctx.mel(0), a)


def test_mel_prefetched_upload(pkg, ctx):
    """wb_pcm_prefetch: the next batch's upload runs on the copy stream into the second staging
    buffer; the following whisper_pcm_to_mel with the same host buffer must give the same mel as a
    plain call, and a different buffer must fall back to an ordinary copy."""
    from whisper_rs_b200 import api
    a = np.ascontiguousarray(pkg.synth.make_segment(5, 160 * 500, 0.2))
    b = np.ascontiguousarray(pkg.synth.make_segment(6, 160 * 500, 0.2))
    api.whisper_pcm_to_mel(ctx, a)
    ref_a = ctx.mel(0)
    api.whisper_pcm_to_mel(ctx, b)
    ref_b = ctx.mel(0)
    for _ in range(3):   # alternate so both staging buffers are used as prefetch targets
        api.whisper_pcm_prefetch_ptr(ctx, a.ctypes.data, a.nbytes)
        api.whisper_pcm_to_mel_ptr(ctx, a.ctypes.data, a.size, 1)
        np.testing.assert_array_equal(ctx.mel(0), ref_a)
        api.whisper_pcm_prefetch_ptr(ctx, b.ctypes.data, b.nbytes)
        api.whisper_pcm_to_mel_ptr(ctx, b.ctypes.data, b.size, 1)
        np.testing.assert_array_equal(ctx.mel(0), ref_b)
    api.whisper_pcm_prefetch_ptr(ctx, a.ctypes.data, a.nbytes)   # prefetched but not consumed ...
    api.whisper_pcm_to_mel_ptr(ctx, b.ctypes.data, b.size, 1)    # ... a different buffer is copied normally
    np.testing.assert_array_equal(ctx.mel(0), ref_b)


def test_mel_clips_shorter_than_one_window(pkg, pyoracle, model_path, ctx):
    """n_len = n_samples / 160 (src/main.rs:1575): clips shorter than one FFT window (400 samples) still yield their
    frames (zero-filled past the end, 1596-1600); fewer than 160 samples yield no frame at all, and encoding such a
    clip encodes an all-zero window (1816-1829)."""
    from whisper_rs_b200 import api
    orc = pyoracle.Oracle(model_path("micro"))
    for n in (160, 161, 399, 400, 479, 480):
        pcm = pkg.synth.make_segment(13, 1000, 0.0)[:n].copy()
        api.whisper_pcm_to_mel(ctx, pcm)
        ref = orc.pcm_to_mel(pcm)
        got = ctx.mel(0)
        assert got.shape == ref.shape == (80, n // 160), n
        assert mel_close(got, ref), n
    pcm = pkg.synth.make_segment(13, 1000, 0.0)[:100].copy()       # no frame
    api.whisper_pcm_to_mel(ctx, pcm)
    assert ctx.mel(0).shape == (80, 0)
    api.whisper_encode(ctx, 1, 0)
    orc.pcm_to_mel(pcm)
    assert rel_l2(ctx.encoder_out(0), orc.encode(0)) < 1e-2
