"""The whole hot path once on the micro architecture -- mel, encode (2 segments), prompt pass, greedy steps -- for
compute-sanitizer:  compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_micro.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
from whisper_rs_b200 import api  # noqa: E402

with tempfile.TemporaryDirectory() as td:
    path = os.path.join(td, "ggml-micro.bin")
    hp = pkg.ggml_file.make_model(path, "micro")
    n = 2 * hp.n_audio_ctx * 160
    clips = np.stack([pkg.synth.make_segment(s, n, 0.2) for s in range(2)])
    ctx = api.WhisperContext.new(path, max_segments=2, max_clips=2, max_clip_samples=n, checkpoints=True)
    api.whisper_pcm_to_mel(ctx, clips)
    api.whisper_encode(ctx, 1, [0, 0], clip_ids=[0, 1])
    d = ctx.encoder_digest(2)
    api.whisper_decode(ctx, np.array([[3, 5, 8], [3, 5, 9]], dtype=np.int32), 0)
    toks, marg, lens = api.whisper_decode_greedy(ctx, [7], 6, n_seqs=2, eot=hp.n_vocab - 1)
    ctx.set_audio_ctx(24)
    api.whisper_encode(ctx, 1, [0, 48], clip_ids=[0, 0])
    toks2, _, _ = api.whisper_decode_greedy(ctx, [7], 4, n_seqs=2, eot=hp.n_vocab - 1)
    ctx.close()
print("sanitize_micro ok: digests", d, "tokens", toks.tolist(), toks2.tolist())
