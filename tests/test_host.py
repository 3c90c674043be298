"""CPU: host-side pieces -- ggml file writer/reader, synthetic PCM, C-ABI exports and the
fail-loudly behaviour of the product when no GPU is present."""
import ctypes as C
import os
import struct

import numpy as np
import pytest


def test_ggml_roundtrip(pkg, model_path):
    mf = pkg.ggml_file.read_model(model_path("micro"))
    hp = pkg.ggml_file.ARCHS["micro"]
    assert mf.hparams == hp
    assert mf.filters.shape == (80, 201)
    table = pkg.ggml_file.tensor_table(hp)
    assert set(mf.tensors) == {n for n, _, _ in table}
    # 11 global + 15 per encoder block + 24 per decoder block
    assert len(table) == 11 + 15 * hp.n_audio_layer + 24 * hp.n_text_layer
    for name, ne, kind in table:
        a = mf.tensors[name]
        assert a.shape == tuple(reversed(ne)), name
        assert a.dtype == (np.float16 if kind == "w" else np.float32), name
    # conv biases are stored 2-D [1, d] (src/main.rs:962, 966)
    assert mf.tensors["encoder.conv1.bias"].shape == (hp.n_audio_state, 1)


def test_ggml_header_layout(pkg, model_path):
    raw = open(model_path("micro"), "rb").read()
    assert struct.unpack_from("<I", raw, 0)[0] == 0x67676D6C           # src/main.rs:46
    assert list(struct.unpack_from("<11i", raw, 4)) == pkg.ggml_file.ARCHS["micro"].as_list()
    assert struct.unpack_from("<2i", raw, 48) == (80, 201)


def test_filterbank_matches_transformers(pkg):
    tf = pytest.importorskip("transformers.audio_utils")
    for n_mels in (80, 128):
        ref = tf.mel_filter_bank(num_frequency_bins=201, num_mel_filters=n_mels, min_frequency=0.0,
                                 max_frequency=8000.0, sampling_rate=16000, norm="slaney",
                                 mel_scale="slaney").T
        got = pkg.ggml_file.mel_filterbank(n_mels)
        assert np.abs(got - ref).max() < 1e-6


def test_synth_deterministic(pkg):
    a = pkg.synth.make_segment(5, 48000)
    b = pkg.synth.make_segment(5, 48000)
    np.testing.assert_array_equal(a, b)
    assert a.dtype == np.float32 and np.abs(a).max() < 1.0
    assert np.all(a[-100:] == 0.0)              # silent tail
    assert np.abs(a).max() > 0.05
    assert not np.array_equal(a, pkg.synth.make_segment(6, 48000))


def test_cabi_exports_every_declared_symbol(pkg):
    from whisper_rs_b200 import cabi
    cabi.build()
    L = C.CDLL(cabi.LIB_PATH)
    names = cabi.declared_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    L.wb_version.restype = C.c_char_p
    assert b"sm_100a" in L.wb_version()


def test_cabi_loader_errors_and_no_cpu_fallback(pkg, model_path, tmp_path):
    """Loader errors surface as WsError variants; with no GPU the product refuses to run
    (WsError::WrongGTensor) instead of falling back to a CPU path."""
    import torch
    from whisper_rs_b200 import api
    good = open(model_path("micro"), "rb").read()
    bad = tmp_path / "bad_magic.bin"
    bad.write_bytes(b"\x01\x02\x03\x04" + good[4:])
    with pytest.raises(api.WsError) as e:
        api.WhisperContext.new(str(bad))
    assert e.value.variant == "BadMagic" and "bad magic" in str(e.value)
    with pytest.raises(api.WsError) as e:
        api.WhisperContext.new(str(tmp_path / "nope.bin"))
    assert e.value.variant == "UnexpectIO"
    # unknown tensor name (src/main.rs:1401-1403)
    mf = pkg.ggml_file.read_model(model_path("micro"))
    t = dict(mf.tensors)
    t["encoder.bogus.weight"] = np.zeros((4,), np.float32)
    p = tmp_path / "unknown.bin"
    pkg.ggml_file.write_model(str(p), mf.hparams, 0, tensors=t)
    with pytest.raises(api.WsError) as e:
        api.WhisperContext.new(str(p))
    assert e.value.variant == "UnknownTensor"
    # wrong shape with the right element count (1413-1422)
    t = dict(mf.tensors)
    t["encoder.blocks.0.mlp.0.weight"] = t["encoder.blocks.0.mlp.0.weight"].T.copy()
    p = tmp_path / "shape.bin"
    pkg.ggml_file.write_model(str(p), mf.hparams, 0, tensors=t)
    with pytest.raises(api.WsError) as e:
        api.WhisperContext.new(str(p))
    assert e.value.variant == "WrongShapeTensor"
    if not torch.cuda.is_available():
        with pytest.raises(api.WsError) as e:
            api.WhisperContext.new(model_path("micro"))
        assert e.value.variant == "WrongGTensor" and "no CPU fallback" in str(e.value)
