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


def test_rust_sys_crate_declares_the_header(pkg):
    """The reference-language binding (whisper.rs_b200/rust/whisper-b200-sys, source only: no rustc in this image) has
    to stay in step with include/whisper_b200.h: every exported entry point except the developer hooks is declared, and
    the config struct carries the header's fields in the header's order."""
    import os
    import re
    from whisper_rs_b200 import cabi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rs = open(os.path.join(root, "whisper.rs_b200", "rust", "whisper-b200-sys", "src", "lib.rs")).read()
    hdr = open(os.path.join(root, "include", "whisper_b200.h")).read()
    hooks = {"wb_dbg_attention", "wb_dbg_gemm", "wb_dbg_layernorm", "wb_kernel_time_us"}
    missing = [n for n in cabi.declared_symbols() if n not in hooks and not re.search(r"fn\s+%s\s*\(" % n, rs)]
    assert not missing, missing
    c_fields = re.findall(r"\b(?:int32_t|int64_t|int|void\s*\*)\s*(\w+)\s*(?:\[\d+\])?;",
                          hdr[hdr.index("typedef struct wb_config"):hdr.index("} wb_config;")])
    r_fields = re.findall(r"pub\s+(\w+)\s*:", rs[rs.index("pub struct wb_config"):rs.index("}", rs.index("pub struct wb_config"))])
    assert c_fields and c_fields == r_fields, (c_fields, r_fields)


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


def test_loader_wrong_size_and_wrong_bytes(pkg, model_path, tmp_path):
    """WsError::WrongSizeTensor (src/main.rs:1406-1411: element count differs from the declared tensor) and
    WsError::WrongBytesTensor (1428-1433: the record's ftype gives a byte size other than the declared dtype's),
    through the C-ABI -- both are raised by the file parser, before any device is touched."""
    from whisper_rs_b200 import api
    mf = pkg.ggml_file.read_model(model_path("micro"))
    d = mf.hparams.n_audio_state
    t = dict(mf.tensors)
    t["encoder.ln_post.weight"] = np.ones((d + 1,), np.float32)            # one element too many
    p = tmp_path / "size.bin"
    pkg.ggml_file.write_model(str(p), mf.hparams, 0, tensors=t)
    with pytest.raises(api.WsError) as e:
        api.WhisperContext.new(str(p))
    assert e.value.variant == "WrongSizeTensor" and e.value.code == -7 and "wrong size" in str(e.value)
    t = dict(mf.tensors)
    w = "encoder.blocks.0.mlp.0.weight"
    assert mf.hparams.f16 == 1 and t[w].dtype == np.float16
    t[w] = t[w].astype(np.float32)                                          # right shape, ftype 0 where F16 is declared
    p = tmp_path / "bytes.bin"
    pkg.ggml_file.write_model(str(p), mf.hparams, 0, tensors=t)
    with pytest.raises(api.WsError) as e:
        api.WhisperContext.new(str(p))
    assert e.value.variant == "WrongBytesTensor" and e.value.code == -9 and "wrong bytes" in str(e.value)
    t = dict(mf.tensors)
    t["encoder.ln_post.bias"] = t["encoder.ln_post.bias"].astype(np.float16)   # F16 where f32 is declared
    p = tmp_path / "bytes2.bin"
    pkg.ggml_file.write_model(str(p), mf.hparams, 0, tensors=t)
    with pytest.raises(api.WsError) as e:
        api.WhisperContext.new(str(p))
    assert e.value.variant == "WrongBytesTensor"
    # a tensor of the table that the file never fills (BadRefTensor, 66-67)
    t = dict(mf.tensors)
    del t["decoder.ln.bias"]
    p = tmp_path / "missing.bin"
    pkg.ggml_file.write_model(str(p), mf.hparams, 0, tensors=t)
    with pytest.raises(api.WsError) as e:
        api.WhisperContext.new(str(p))
    assert e.value.variant == "BadRefTensor"
    # truncated file: the last record's data is cut short
    raw = open(model_path("micro"), "rb").read()
    p = tmp_path / "short.bin"
    p.write_bytes(raw[:-100])
    with pytest.raises(api.WsError) as e:
        api.WhisperContext.new(str(p))
    assert e.value.variant == "UnexpectIO"


def test_abi_struct_layout_matches_ctypes(pkg):
    """A plain-C consumer of include/whisper_b200.h (tests/c/abi_layout.c, gcc -std=c99 -pedantic) prints the layout
    the compiler gives every struct that crosses the ABI; the hand-written ctypes mirror must agree field by field
    (a drift in wb_config / wb_timings would otherwise go unnoticed)."""
    import subprocess
    from whisper_rs_b200 import cabi
    cdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
    subprocess.run(["make", "-C", cdir, "abi_layout"], check=True, stdout=subprocess.DEVNULL)
    out = subprocess.run([os.path.join(cdir, "abi_layout")], check=True, capture_output=True, text=True).stdout
    got = {}
    for line in out.splitlines():
        f = line.split()
        got[f[0]] = tuple(int(x) for x in f[1:])
    for name, st in (("wb_config", cabi.WbConfig), ("wb_timings", cabi.WbTimings)):
        assert got[f"sizeof.{name}"] == (C.sizeof(st),), name
        for fname, _ in st._fields_:
            fld = getattr(st, fname)
            assert got[f"{name}.{fname}"] == (fld.offset, fld.size), (name, fname)
        assert len([k for k in got if k.startswith(name + ".")]) == len(st._fields_), name
    assert got["enum.WB_ERR_TENSOR_OP"] == (cabi.WB_ERR_TENSOR_OP,)
    assert got["enum.WB_STAGE_CROSS_V"] == (cabi.STAGE_CROSS_V,)
    assert got["enum.WB_NORM_SEGMENT"] == (cabi.NORM_SEGMENT,)


def test_main_replay_builds_and_fails_loudly_without_gpu(pkg, model_path, tmp_path):
    """tests/c/main_replay.c -- the reference's `fn main` (src/main.rs:2065-2075) in C against the header -- links
    against libwhisper_b200.so with gcc alone; without a B200 it exits with WsError::WrongGTensor (10), not a CPU
    result."""
    import subprocess
    import torch
    from whisper_rs_b200 import cabi
    cabi.build()
    cdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
    subprocess.run(["make", "-C", cdir, "main_replay"], check=True, stdout=subprocess.DEVNULL)
    if torch.cuda.is_available():
        pytest.skip("GPU present: the run itself is tests/test_gpu_cabi.py")
    pcm = tmp_path / "pcm.raw"
    pkg.synth.make_segment(0, 160 * 64, 0.1).tofile(str(pcm))
    r = subprocess.run([os.path.join(cdir, "main_replay"), model_path("micro"), str(pcm), str(tmp_path / "out")],
                       capture_output=True, text=True)
    assert r.returncode == 10 and "no CPU fallback" in r.stderr, (r.returncode, r.stderr)


def test_special_tokens_follow_the_language_count(pkg, pyoracle, tmp_path):
    """Special-token fix-up (src/main.rs:433-440): +1 for a 51865-entry (multilingual) vocabulary; every further
    language token (large-v3: 51866) moves the ids behind the language block once more -- <|notimestamps|> 50364,
    first time stamp 50365, translate / transcribe 50359 / 50360 -- while eot / sot stay at 50257 / 50258."""
    import dataclasses
    micro = pkg.ggml_file.ARCHS["micro"]
    want = {
        51864: (50256, 50257, 50360, 50361, 50362, 50363, 50358, 50359),
        51865: (50257, 50258, 50361, 50362, 50363, 50364, 50358, 50359),
        51866: (50257, 50258, 50362, 50363, 50364, 50365, 50359, 50360),
    }
    for nv, ids in want.items():
        p = str(tmp_path / f"v{nv}.bin")
        pkg.ggml_file.write_model(p, dataclasses.replace(micro, n_vocab=nv), seed=1, n_vocab_file=50257)
        o = pyoracle.Oracle(p)
        assert (o.token_eot, o.token_sot, o.token_prev, o.token_solm, o.token_not, o.token_beg, o.token_translate,
                o.token_transcribe) == ids, nv


def test_cpp_host_header_error_behaviour(pkg, model_path, tmp_path):
    """include/whisper_b200.hpp (the C++ host side with the reference's names) compiled with g++ -std=c++17 -pedantic:
    every loader failure arrives as the reference's WsError variant with its Display text (src/main.rs:50-72), and
    without a B200 `WhisperContext::new_` throws WrongGTensor instead of computing on the CPU."""
    import subprocess
    import torch
    from whisper_rs_b200 import cabi
    cabi.build()
    cdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp")
    subprocess.run(["make", "-C", cdir, "reference_main"], check=True, stdout=subprocess.DEVNULL)
    exe = os.path.join(cdir, "reference_main")
    good = open(model_path("micro"), "rb").read()
    mf = pkg.ggml_file.read_model(model_path("micro"))
    d = tmp_path / "bad"
    d.mkdir()
    (d / "bad_magic.bin").write_bytes(b"\x01\x02\x03\x04" + good[4:])
    t = dict(mf.tensors)
    t["encoder.bogus.weight"] = np.zeros((4,), np.float32)
    pkg.ggml_file.write_model(str(d / "unknown.bin"), mf.hparams, 0, tensors=t)
    t = dict(mf.tensors)
    t["encoder.ln_post.weight"] = np.ones((mf.hparams.n_audio_state + 1,), np.float32)
    pkg.ggml_file.write_model(str(d / "size.bin"), mf.hparams, 0, tensors=t)
    t = dict(mf.tensors)
    t["encoder.blocks.0.mlp.0.weight"] = t["encoder.blocks.0.mlp.0.weight"].T.copy()
    pkg.ggml_file.write_model(str(d / "shape.bin"), mf.hparams, 0, tensors=t)
    t = dict(mf.tensors)
    t["encoder.blocks.0.mlp.0.weight"] = t["encoder.blocks.0.mlp.0.weight"].astype(np.float32)
    pkg.ggml_file.write_model(str(d / "bytes.bin"), mf.hparams, 0, tensors=t)
    r = subprocess.run([exe, "--errors", str(d)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert r.stdout.count(" ok") == 6 and "UNEXPECTED" not in r.stdout
    for variant in ("BadMagic", "UnexpectIO", "UnknownTensor", "WrongSizeTensor", "WrongShapeTensor", "WrongBytesTensor"):
        assert variant in r.stdout
    if not torch.cuda.is_available():
        pcm = tmp_path / "pcm.raw"
        np.zeros(16000, np.int16).tofile(str(pcm))
        r = subprocess.run([exe, model_path("micro"), str(pcm), str(tmp_path / "o")], capture_output=True, text=True)
        assert r.returncode == 10 and r.stderr.startswith("WrongGTensor:") and "no CPU fallback" in r.stderr, (r.returncode, r.stderr)
