"""GPU: a compiled C consumer of include/whisper_b200.h (tests/c/main_replay.c, gcc -std=c99) replays the
reference's `fn main` (src/main.rs:2065-2075: WhisperContext::new -> whisper_pcm_to_mel -> whisper_encode(ctx, 1, 0))
plus one decode step against libwhisper_b200.so; its outputs are checked against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import mel_close, rel_l2

pytestmark = pytest.mark.gpu
CDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")


@pytest.mark.parametrize("arch", ["micro", "tiny"])
def test_main_replay_in_c_vs_oracle(pkg, pyoracle, model_path, tmp_path, arch):
    subprocess.run(["make", "-C", CDIR, "main_replay"], check=True, stdout=subprocess.DEVNULL)
    hp = pkg.ggml_file.ARCHS[arch]
    n = 2 * hp.n_audio_ctx * 160
    pcm = pkg.synth.make_segment(33, n, silent_tail_s=0.2)
    pcm_path, prefix = str(tmp_path / "pcm.raw"), str(tmp_path / "out")
    pcm.tofile(pcm_path)
    r = subprocess.run([os.path.join(CDIR, "main_replay"), model_path(arch), pcm_path, prefix], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "main_replay ok" in r.stdout and "sm_100a" in r.stdout
    orc = pyoracle.Oracle(model_path(arch))
    ref_mel = orc.pcm_to_mel(pcm)
    ref_enc = orc.encode(0)
    ref_logits = orc.decode([orc.token_sot if orc.token_sot < orc.n_vocab else 7], 0)   # main_replay.c's prompt rule
    mel = np.fromfile(prefix + ".mel.f32", dtype=np.float32).reshape(ref_mel.shape)
    enc = np.fromfile(prefix + ".enc.f32", dtype=np.float32).reshape(ref_enc.shape)
    logits = np.fromfile(prefix + ".logits.f32", dtype=np.float32)
    assert mel_close(mel, ref_mel)
    assert rel_l2(enc, ref_enc) < 1e-2
    assert rel_l2(logits, ref_logits) < 1e-2


CPPDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp")


@pytest.mark.parametrize("arch", ["micro", "tiny"])
def test_reference_main_in_cpp_vs_oracle(pkg, pyoracle, model_path, tmp_path, arch):
    """tests/cpp/reference_main.cpp: the reference's `fn main` (src/main.rs:2065-2075) line for line over
    include/whisper_b200.hpp -- i16 samples -> convert_integer_to_float_audio -> WhisperContext::new ->
    whisper_pcm_to_mel -> whisper_encode(ctx, 1, 0) (+ whisper_decode) -- compiled with g++ and checked against the
    oracle fed the same i16 samples."""
    subprocess.run(["make", "-C", CPPDIR, "reference_main"], check=True, stdout=subprocess.DEVNULL)
    hp = pkg.ggml_file.ARCHS[arch]
    n = 2 * hp.n_audio_ctx * 160
    s16 = np.round(pkg.synth.make_segment(34, n, silent_tail_s=0.2) * 32767.0).astype(np.int16)
    pcm_path, prefix = str(tmp_path / "pcm_s16.raw"), str(tmp_path / "out")
    s16.tofile(pcm_path)
    r = subprocess.run([os.path.join(CPPDIR, "reference_main"), model_path(arch), pcm_path, prefix], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert f"len:{n}" in r.stdout and "reference_main ok" in r.stdout
    orc = pyoracle.Oracle(model_path(arch))
    ref_mel = orc.pcm_to_mel(s16.astype(np.float32) / 32768.0)               # convert_integer_to_float_audio (1673-1679)
    ref_enc = orc.encode(0)
    ref_logits = orc.decode([orc.token_sot if orc.token_sot < orc.n_vocab else 7], 0)
    mel = np.fromfile(prefix + ".mel.f32", dtype=np.float32).reshape(ref_mel.shape)
    enc = np.fromfile(prefix + ".enc.f32", dtype=np.float32).reshape(ref_enc.shape)
    logits = np.fromfile(prefix + ".logits.f32", dtype=np.float32)
    assert mel_close(mel, ref_mel)
    assert rel_l2(enc, ref_enc) < 1e-2
    assert rel_l2(logits, ref_logits) < 1e-2
