"""GPU: each hand-written kernel in isolation, through the C-ABI probes, against numpy restatements
of the galois/ggml op it replaces."""
import os

import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(pkg, model_path):
    from whisper_rs_b200 import api
    c = api.WhisperContext.new(model_path("micro"), max_segments=2, decode_capacity=False)
    yield c
    c.close()


def _gelu_ref(x):
    x = x.astype(np.float16).astype(np.float32)          # ggml: input rounded to F16 (appendix A)
    return 0.5 * x * (1 + np.tanh(0.79788456080286535588 * x * (1 + 0.044715 * x * x)))


def test_layernorm(ctx):
    from whisper_rs_b200 import api
    rng = np.random.default_rng(1)
    for d in (128, 384, 512, 768, 1024, 1280):
        x = (rng.standard_normal((37, d)) * 3 + 0.5).astype(np.float32)
        w = (1 + 0.1 * rng.standard_normal(d)).astype(np.float32)
        b = (0.1 * rng.standard_normal(d)).astype(np.float32)
        x64 = x.astype(np.float64)
        mu = x64.mean(1, keepdims=True)
        var = ((x64 - mu) ** 2).mean(1, keepdims=True)
        ref = ((x64 - mu) / np.sqrt(var + 1e-5)) * w + b
        got = api.dbg_layernorm(ctx, x, w, b).astype(np.float32)
        assert np.abs(got - ref).max() <= 2e-3 * max(1.0, np.abs(ref).max()), d   # F16 output rounding


@pytest.mark.parametrize("M,N,K", [
    (128, 128, 64), (128, 256, 128), (300, 384, 384), (1500, 1152, 384), (257, 192, 240),
    (1000, 512, 2048), (96, 1536, 512), (130, 1000, 128), (64, 40, 64), (3000, 256, 240),
    (6000, 1024, 256),    # more work items than CTA pairs: the persistent loop and both TMEM accumulators
    (20000, 64, 64), (2050, 320, 128),
])
def test_gemm_plain(ctx, M, N, K):
    from whisper_rs_b200 import api
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    a = (rng.standard_normal((M, K)) * 0.5).astype(np.float16)
    w = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float16)
    ref = a.astype(np.float32) @ w.astype(np.float32).T
    got32 = api.dbg_gemm(ctx, a, w, out_f16=False)
    assert rel_l2(got32, ref) < 1e-5                       # f32 accumulate of exact F16 products
    got16 = api.dbg_gemm(ctx, a, w, out_f16=True).astype(np.float32)
    assert np.abs(got16 - ref).max() <= 2e-3 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("M,N,K", [(515, 768, 384), (9000, 512, 128), (300, 1280, 64)])
def test_gemm_epilogues(ctx, M, N, K):
    from whisper_rs_b200 import api
    rng = np.random.default_rng(5)
    a = (rng.standard_normal((M, K)) * 0.5).astype(np.float16)
    w = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float16)
    bias = (0.3 * rng.standard_normal(N)).astype(np.float32)
    res = rng.standard_normal((M, N)).astype(np.float32)
    acc = a.astype(np.float32) @ w.astype(np.float32).T
    got = api.dbg_gemm(ctx, a, w, bias=bias, out_f16=False)
    assert rel_l2(got, acc + bias) < 1e-5
    got = api.dbg_gemm(ctx, a, w, bias=bias, residual=res, out_f16=False)
    assert rel_l2(got, acc + bias + res) < 1e-5
    got = api.dbg_gemm(ctx, a, w, bias=bias, scale=0.35355339, out_f16=False)
    assert rel_l2(got, (acc + bias) * 0.35355339) < 1e-5
    got = api.dbg_gemm(ctx, a, w, bias=bias, gelu=True, out_f16=True).astype(np.float32)
    ref = _gelu_ref(acc + bias)
    assert np.abs(got - ref).max() <= 3e-3 * max(1.0, np.abs(ref).max())


def _attention_ref(qkv, n_seg, T, H):
    d = H * 64
    q = qkv[:, :d].astype(np.float32).reshape(n_seg, T, H, 64)
    k = qkv[:, d:2 * d].astype(np.float32).reshape(n_seg, T, H, 64)
    v = qkv[:, 2 * d:].astype(np.float32).reshape(n_seg, T, H, 64)
    out = np.empty((n_seg, T, H, 64), np.float32)
    for s in range(n_seg):
        for h in range(H):
            sc = (q[s, :, h] @ k[s, :, h].T) * 0.125
            p = np.exp(sc - sc.max(1, keepdims=True))
            p = (p / p.sum(1, keepdims=True)).astype(np.float16).astype(np.float32)   # P -> F16 (appendix A)
            out[s, :, h] = p @ v[s, :, h]
    return out.reshape(n_seg * T, d)


@pytest.mark.parametrize("n_seg,T,H", [(1, 128, 1), (1, 96, 2), (2, 200, 2), (1, 1500, 6), (2, 333, 3)])
def test_attention(ctx, n_seg, T, H):
    from whisper_rs_b200 import api
    rng = np.random.default_rng(T + H)
    qkv = rng.standard_normal((n_seg * T, 3 * H * 64)).astype(np.float16)
    ref = _attention_ref(qkv, n_seg, T, H)
    got = api.dbg_attention(ctx, qkv, n_seg, T, H).astype(np.float32)
    assert np.isfinite(got).all()
    assert rel_l2(got, ref) < 3e-3
    assert np.abs(got - ref).max() < 2e-2
