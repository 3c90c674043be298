"""Generate tests/golden/*.npz -- run in the build container (needs `transformers`; no GPU).

The reference (szuwgh/whisper.rs) holds no golden vectors and cannot be built here (no rustc,
`galois` path dependency absent), so the oracle cannot be pinned against the reference itself.
What CAN be pinned, and is pinned here, is the oracle's dataflow and layout against two
independent evaluations:

  1. mel: an exact f64 evaluation (numpy rfft) of the algorithm of src/main.rs:1554-1671
     (periodic Hann, no centring, zero fill, bin fold, filterbank, log10, whole-clip clamp).
  2. encoder / decoder: HuggingFace `transformers` WhisperModel in fp32 carrying the SAME
     random-init weights (micro architecture).  HF uses erf-GELU and no F16 rounding, so the
     oracle is run with its rounding points switched off (ORC_OPT_*); agreement at ~1e-4 then
     pins every transposition, bias placement, scale and layout of the restatement.  The
     canonical (rounding ON) oracle outputs are stored next to them.

Outputs are small (micro model: d=128, 2+2 layers, n_ctx=96) and committed.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
from oracle import pyoracle  # noqa: E402


def mel_f64(pcm: np.ndarray, filters: np.ndarray) -> np.ndarray:
    n_len = pcm.size // 160
    hann = 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(400) / 400.0))
    padded = np.concatenate([pcm.astype(np.float64), np.zeros(400)])
    idx = np.arange(n_len)[:, None] * 160 + np.arange(400)[None, :]
    frames = padded[idx] * hann[None, :]
    spec = np.fft.fft(frames, axis=1)
    p = (spec.real ** 2 + spec.imag ** 2)
    fold = p[:, :201].copy()
    fold[:, 1:200] += p[:, 399:200:-1]          # p[j] += p[400-j], j = 1..199
    mel = fold @ filters.astype(np.float64).T    # [n_len][n_mel]
    mel = np.log10(np.maximum(mel, 1e-10)).T     # [n_mel][n_len]
    mmax = mel.max() - 8.0
    mel = np.maximum(mel, mmax)
    return (mel + 4.0) / 4.0


def hf_model(mf):
    import torch
    from transformers import WhisperConfig, WhisperModel

    hp = mf.hparams
    cfg = WhisperConfig(
        vocab_size=hp.n_vocab, num_mel_bins=hp.n_mels, encoder_layers=hp.n_audio_layer,
        encoder_attention_heads=hp.n_audio_head, decoder_layers=hp.n_text_layer,
        decoder_attention_heads=hp.n_text_head, d_model=hp.n_audio_state,
        encoder_ffn_dim=4 * hp.n_audio_state, decoder_ffn_dim=4 * hp.n_text_state,
        max_source_positions=hp.n_audio_ctx, max_target_positions=hp.n_text_ctx,
        activation_function="gelu", dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
        pad_token_id=0, bos_token_id=1, eos_token_id=2, decoder_start_token_id=1,
    )
    model = WhisperModel(cfg).eval().float()
    t = {k: torch.from_numpy(v.astype(np.float32)) for k, v in mf.tensors.items()}
    sd = {}
    sd["encoder.conv1.weight"] = t["encoder.conv1.weight"]
    sd["encoder.conv1.bias"] = t["encoder.conv1.bias"].reshape(-1)
    sd["encoder.conv2.weight"] = t["encoder.conv2.weight"]
    sd["encoder.conv2.bias"] = t["encoder.conv2.bias"].reshape(-1)
    sd["encoder.embed_positions.weight"] = t["encoder.positional_embedding"]
    sd["encoder.layer_norm.weight"] = t["encoder.ln_post.weight"]
    sd["encoder.layer_norm.bias"] = t["encoder.ln_post.bias"]
    sd["decoder.embed_tokens.weight"] = t["decoder.token_embedding.weight"]
    sd["decoder.embed_positions.weight"] = t["decoder.positional_embedding"]
    sd["decoder.layer_norm.weight"] = t["decoder.ln.weight"]
    sd["decoder.layer_norm.bias"] = t["decoder.ln.bias"]

    def attn(dst, src):
        sd[dst + "q_proj.weight"] = t[src + "query.weight"]
        sd[dst + "q_proj.bias"] = t[src + "query.bias"]
        sd[dst + "k_proj.weight"] = t[src + "key.weight"]
        sd[dst + "v_proj.weight"] = t[src + "value.weight"]
        sd[dst + "v_proj.bias"] = t[src + "value.bias"]
        sd[dst + "out_proj.weight"] = t[src + "out.weight"]
        sd[dst + "out_proj.bias"] = t[src + "out.bias"]

    for part, n in (("encoder", hp.n_audio_layer), ("decoder", hp.n_text_layer)):
        for i in range(n):
            s, dn = f"{part}.blocks.{i}.", f"{part}.layers.{i}."
            attn(dn + "self_attn.", s + "attn.")
            sd[dn + "self_attn_layer_norm.weight"] = t[s + "attn_ln.weight"]
            sd[dn + "self_attn_layer_norm.bias"] = t[s + "attn_ln.bias"]
            sd[dn + "fc1.weight"] = t[s + "mlp.0.weight"]
            sd[dn + "fc1.bias"] = t[s + "mlp.0.bias"]
            sd[dn + "fc2.weight"] = t[s + "mlp.2.weight"]
            sd[dn + "fc2.bias"] = t[s + "mlp.2.bias"]
            sd[dn + "final_layer_norm.weight"] = t[s + "mlp_ln.weight"]
            sd[dn + "final_layer_norm.bias"] = t[s + "mlp_ln.bias"]
            if part == "decoder":
                attn(dn + "encoder_attn.", s + "cross_attn.")
                sd[dn + "encoder_attn_layer_norm.weight"] = t[s + "cross_attn_ln.weight"]
                sd[dn + "encoder_attn_layer_norm.bias"] = t[s + "cross_attn_ln.bias"]
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [m for m in missing if "k_proj.bias" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    return model


def main() -> None:
    import torch

    out_dir = os.path.dirname(os.path.abspath(__file__))
    arch = "micro"
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "micro.bin")
        pkg.ggml_file.make_model(path, arch)
        mf = pkg.ggml_file.read_model(path)
        hp = mf.hparams
        n_samples = 2 * hp.n_audio_ctx * 160 + 2 * 160      # a clip slightly longer than one window
        pcm = pkg.synth.make_segment(0, n_samples, silent_tail_s=0.2)
        orc = pyoracle.Oracle(path)

        # ---- 1. mel vs exact f64 evaluation
        mel = orc.pcm_to_mel(pcm)
        ref = mel_f64(pcm, mf.filters)
        err = np.abs(mel - ref)
        print("mel: max abs err vs f64 = %.3e, rel-L2 = %.3e" % (err.max(), np.linalg.norm(mel - ref) / np.linalg.norm(ref)))
        assert err.max() < 1e-4

        # ---- 2. encoder / decoder vs HF (rounding points off)
        model = hf_model(mf)
        win = np.zeros((hp.n_mels, 2 * hp.n_audio_ctx), np.float32)
        n = min(mel.shape[1], win.shape[1])
        win[:, :n] = mel[:, :n]
        tokens = np.array([5, 17, 900, 3, 64, 511], dtype=np.int32)
        with torch.no_grad():
            enc_hf = model.encoder(torch.from_numpy(win)[None]).last_hidden_state[0].numpy()
            dec_hf = model.decoder(input_ids=torch.from_numpy(tokens.astype(np.int64))[None],
                                   encoder_hidden_states=torch.from_numpy(enc_hf)[None]).last_hidden_state[0]
            logits_hf = (dec_hf @ model.decoder.embed_tokens.weight.T).numpy()
        for o, v in ((pyoracle.OPT_ACT_F16_ROUND, 0), (pyoracle.OPT_GELU_MODE, 2),
                     (pyoracle.OPT_SOFTMAX_EXP, 1), (pyoracle.OPT_PROB_F16_ROUND, 0)):
            orc.set_option(o, v)
        enc_nr = orc.encode(0)
        lg_nr = orc.decode(tokens, 0)
        rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
        print("encoder (rounding off) vs HF: rel-L2 = %.3e" % rel(enc_nr, enc_hf))
        print("logits  (rounding off) vs HF: rel-L2 = %.3e" % rel(lg_nr, logits_hf[-1]))
        assert rel(enc_nr, enc_hf) < 2e-4
        assert rel(lg_nr, logits_hf[-1]) < 2e-3     # cross-KV stays F16 in the oracle's state
        # incremental decode == batched decode (KV cache)
        orc.decode(tokens[:4], 0)
        lg_inc = orc.decode(tokens[4:], 4)
        print("incremental vs batched logits: max abs = %.3e" % np.abs(lg_inc - lg_nr).max())
        assert np.abs(lg_inc - lg_nr).max() < 1e-4

        # ---- canonical oracle (rounding ON): the vectors the CUDA path is compared against
        for o, v in ((pyoracle.OPT_ACT_F16_ROUND, 1), (pyoracle.OPT_GELU_MODE, 0),
                     (pyoracle.OPT_SOFTMAX_EXP, 0), (pyoracle.OPT_PROB_F16_ROUND, 1)):
            orc.set_option(o, v)
        enc = orc.encode(0)
        ck, cv = orc.cross_kv(hp.n_text_layer - 1)
        lg = orc.decode(tokens, 0)
        toks, marg = orc.decode_greedy([7], 12, eot=hp.n_vocab - 1)
        print("encoder canonical vs HF: rel-L2 = %.3e" % rel(enc, enc_hf))
        print("logits  canonical vs HF: rel-L2 = %.3e" % rel(lg, logits_hf[-1]))
        chk = np.array([orc.checksum(pyoracle.STAGE_MEL), orc.checksum(pyoracle.STAGE_CONV1),
                        orc.checksum(pyoracle.STAGE_CONV2_POS)]
                       + [orc.checksum(pyoracle.STAGE_LAYER, i) for i in range(hp.n_audio_layer)]
                       + [orc.checksum(pyoracle.STAGE_LN_POST)])
        np.savez_compressed(
            os.path.join(out_dir, "micro_golden.npz"),
            n_samples=np.int64(n_samples), mel_f64=ref.astype(np.float32), mel_oracle=mel,
            enc_hf=enc_hf.astype(np.float32), enc_oracle=enc, enc_oracle_noround=enc_nr,
            cross_k_last=ck, cross_v_last=cv, tokens=tokens, logits_hf=logits_hf[-1].astype(np.float32),
            logits_oracle=lg, greedy_tokens=toks, greedy_margin=marg, checksums=chk)
        print("wrote micro_golden.npz")


if __name__ == "__main__":
    main()
