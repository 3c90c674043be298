"""ggml-v1 ("lmgg") Whisper model files: writer + reader (host side, numpy only).

The on-disk layout is the one the reference parses, little-endian throughout:

  u32 magic 0x67676d6c                         src/main.rs:46, 368-371
  11 x i32 hparams                             src/main.rs:622-633
  i32 n_mel, i32 n_fft, n_mel*n_fft f32        src/main.rs:513-524
  i32 n_vocab, then per token u32 len + bytes  src/main.rs:430-431, 578-589
  records until EOF:
    i32 n_dims, i32 name_len, i32 ftype(0=f32,1=f16), n_dims x i32 ne[] (ne[0] innermost),
    name bytes, raw data                       src/main.rs:1385-1437

The tensor table (name -> ggml shape, dtype) follows src/main.rs:960-1334.  The writer
exists because the reference ships no model files and no writer;
random-init models of each named architecture are the inputs of every parity test.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, asdict
from typing import Dict, Iterator, List, Tuple

import numpy as np

MAGIC = 0x67676D6C  # src/main.rs:46

HPARAM_FIELDS = (
    "n_vocab", "n_audio_ctx", "n_audio_state", "n_audio_head", "n_audio_layer",
    "n_text_ctx", "n_text_state", "n_text_head", "n_text_layer", "n_mels", "f16",
)  # order of src/main.rs:622-633


@dataclass
class HParams:
    n_vocab: int = 51864
    n_audio_ctx: int = 1500
    n_audio_state: int = 384
    n_audio_head: int = 6
    n_audio_layer: int = 4
    n_text_ctx: int = 448
    n_text_state: int = 384
    n_text_head: int = 6
    n_text_layer: int = 4
    n_mels: int = 80
    f16: int = 1

    def as_list(self) -> List[int]:
        d = asdict(self)
        return [int(d[k]) for k in HPARAM_FIELDS]


# Standard Whisper architectures + a micro config used by the
# CPU-side tests so the whole suite runs in seconds.
ARCHS: Dict[str, HParams] = {
    "micro": HParams(n_vocab=1024, n_audio_ctx=96, n_audio_state=128, n_audio_head=2,
                     n_audio_layer=2, n_text_ctx=32, n_text_state=128, n_text_head=2,
                     n_text_layer=2, n_mels=80),
    "tiny": HParams(51864, 1500, 384, 6, 4, 448, 384, 6, 4, 80, 1),
    "tiny.ml": HParams(51865, 1500, 384, 6, 4, 448, 384, 6, 4, 80, 1),
    "base": HParams(51864, 1500, 512, 8, 6, 448, 512, 8, 6, 80, 1),
    "small": HParams(51864, 1500, 768, 12, 12, 448, 768, 12, 12, 80, 1),
    "medium": HParams(51864, 1500, 1024, 16, 24, 448, 1024, 16, 24, 80, 1),
    "large-v3": HParams(51866, 1500, 1280, 20, 32, 448, 1280, 20, 32, 128, 1),
    # full-width, 2-layer variants: every tile shape / head count / mel width of the big models at a
    # depth the CPU oracle checks in seconds
    "small.2l": HParams(51864, 1500, 768, 12, 2, 448, 768, 12, 2, 80, 1),
    "medium.2l": HParams(51864, 1500, 1024, 16, 2, 448, 1024, 16, 2, 80, 1),
    "large-v3.2l": HParams(51866, 1500, 1280, 20, 2, 448, 1280, 20, 2, 128, 1),
}
ARCH_SEED_INDEX = {"tiny": 0, "base": 1, "small": 2, "medium": 3, "large-v3": 4,
                   "micro": 7, "tiny.ml": 8, "small.2l": 9, "medium.2l": 10, "large-v3.2l": 11}


def tensor_table(hp: HParams) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, ggml ne[] with ne[0] innermost, 'w'|'f32') in the order of src/main.rs:960-1334.

    'w' tensors are F16 when hparams.f16 == 1 (src/main.rs:817-821).
    """
    d, da = hp.n_audio_state, hp.n_audio_state
    dt = hp.n_text_state
    t: List[Tuple[str, Tuple[int, ...], str]] = [
        ("encoder.positional_embedding", (da, hp.n_audio_ctx), "f32"),   # 960
        ("encoder.conv1.weight", (3, hp.n_mels, da), "w"),               # 961
        ("encoder.conv1.bias", (1, da), "f32"),                          # 962 (2-D!)
        ("encoder.conv2.weight", (3, da, da), "w"),                      # 964-965
        ("encoder.conv2.bias", (1, da), "f32"),                          # 966 (2-D!)
        ("encoder.ln_post.weight", (da,), "f32"),                        # 968
        ("encoder.ln_post.bias", (da,), "f32"),                          # 969
    ]
    for i in range(hp.n_audio_layer):                                     # 1006-1136
        p = f"encoder.blocks.{i}."
        t += [
            (p + "mlp_ln.weight", (d,), "f32"), (p + "mlp_ln.bias", (d,), "f32"),
            (p + "mlp.0.weight", (d, 4 * d), "w"), (p + "mlp.0.bias", (4 * d,), "f32"),
            (p + "mlp.2.weight", (4 * d, d), "w"), (p + "mlp.2.bias", (d,), "f32"),
            (p + "attn_ln.weight", (d,), "f32"), (p + "attn_ln.bias", (d,), "f32"),
            (p + "attn.query.weight", (d, d), "w"), (p + "attn.query.bias", (d,), "f32"),
            (p + "attn.key.weight", (d, d), "w"),
            (p + "attn.value.weight", (d, d), "w"), (p + "attn.value.bias", (d,), "f32"),
            (p + "attn.out.weight", (d, d), "w"), (p + "attn.out.bias", (d,), "f32"),
        ]
    t += [
        ("decoder.positional_embedding", (dt, hp.n_text_ctx), "f32"),    # 1139
        ("decoder.token_embedding.weight", (dt, hp.n_vocab), "w"),       # 1141
        ("decoder.ln.weight", (dt,), "f32"), ("decoder.ln.bias", (dt,), "f32"),  # 1142-1143
    ]
    d = dt
    for i in range(hp.n_text_layer):                                      # 1160-1333
        p = f"decoder.blocks.{i}."
        t += [
            (p + "mlp_ln.weight", (d,), "f32"), (p + "mlp_ln.bias", (d,), "f32"),
            (p + "mlp.0.weight", (d, 4 * d), "w"), (p + "mlp.0.bias", (4 * d,), "f32"),
            (p + "mlp.2.weight", (4 * d, d), "w"), (p + "mlp.2.bias", (d,), "f32"),
            (p + "attn_ln.weight", (d,), "f32"), (p + "attn_ln.bias", (d,), "f32"),
            (p + "attn.query.weight", (d, d), "w"), (p + "attn.query.bias", (d,), "f32"),
            (p + "attn.key.weight", (d, d), "w"),
            (p + "attn.value.weight", (d, d), "w"), (p + "attn.value.bias", (d,), "f32"),
            (p + "attn.out.weight", (d, d), "w"), (p + "attn.out.bias", (d,), "f32"),
            (p + "cross_attn_ln.weight", (d,), "f32"), (p + "cross_attn_ln.bias", (d,), "f32"),
            (p + "cross_attn.query.weight", (d, d), "w"),
            (p + "cross_attn.query.bias", (d,), "f32"),
            (p + "cross_attn.key.weight", (d, d), "w"),
            (p + "cross_attn.value.weight", (d, d), "w"),
            (p + "cross_attn.value.bias", (d,), "f32"),
            (p + "cross_attn.out.weight", (d, d), "w"),
            (p + "cross_attn.out.bias", (d,), "f32"),
        ]
    return t


# ---------------------------------------------------------------------------------------
# mel filterbank (slaney scale + slaney norm, the one OpenAI's mel_filters.npz holds)
# ---------------------------------------------------------------------------------------
def _hz_to_mel_slaney(f: np.ndarray) -> np.ndarray:
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz_slaney(m: np.ndarray) -> np.ndarray:
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(n_mels: int, n_fft: int = 400, sr: int = 16000) -> np.ndarray:
    """[n_mels][n_fft//2+1] f32, row-major `[mel][bin]` as src/main.rs:513-524 stores it."""
    n_bins = n_fft // 2 + 1
    fftfreqs = np.linspace(0.0, sr / 2.0, n_bins)
    mel_pts = np.linspace(_hz_to_mel_slaney(0.0), _hz_to_mel_slaney(sr / 2.0), n_mels + 2)
    hz_pts = _mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(hz_pts)
    ramps = hz_pts[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (hz_pts[2:n_mels + 2] - hz_pts[:n_mels])
    w *= enorm[:, None]
    return w.astype(np.float32)


# ---------------------------------------------------------------------------------------
# random-init weights (SURVEY.md section 8d "Synthetic inputs")
# ---------------------------------------------------------------------------------------
def _sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> np.ndarray:
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = np.exp(-inc * np.arange(channels // 2))
    t = np.arange(length)[:, None] * inv[None, :]
    return np.concatenate([np.sin(t), np.cos(t)], axis=1).astype(np.float32)  # [length][channels]


def random_tensors(hp: HParams, seed: int) -> Iterator[Tuple[str, np.ndarray]]:
    """Yields (name, array) with numpy shape = reversed ggml ne[] (C order), dtype f16|f32."""
    rng = np.random.default_rng(seed)
    wdt = np.float16 if hp.f16 == 1 else np.float32
    for name, ne, kind in tensor_table(hp):
        shape = tuple(reversed(ne))
        n = int(np.prod(shape))
        if name == "encoder.positional_embedding":
            a = _sinusoids(hp.n_audio_ctx, hp.n_audio_state)
        elif name == "decoder.positional_embedding":
            a = (rng.standard_normal(n, dtype=np.float32) * 0.01).reshape(shape)
        elif name == "decoder.token_embedding.weight":
            a = (rng.standard_normal(n, dtype=np.float32) * 0.02).reshape(shape)
        elif kind == "w":
            fan_in = int(np.prod(ne[:-1]))  # ne[-1] = out features / out channels
            a = (rng.standard_normal(n, dtype=np.float32) / np.sqrt(fan_in)).reshape(shape)
        elif name.endswith("ln.weight") or name.endswith("ln_post.weight"):
            a = (1.0 + 0.02 * rng.standard_normal(n, dtype=np.float32)).reshape(shape)
        elif name.endswith("ln.bias") or name.endswith("ln_post.bias"):
            a = (0.02 * rng.standard_normal(n, dtype=np.float32)).reshape(shape)
        else:  # linear / conv biases
            a = (0.01 * rng.standard_normal(n, dtype=np.float32)).reshape(shape)
        yield name, np.ascontiguousarray(a.astype(wdt if kind == "w" else np.float32))


def write_model(path: str, hp: HParams, seed: int, n_vocab_file: int | None = None,
                tensors: Dict[str, np.ndarray] | None = None) -> None:
    """Write a ggml-v1 file.  `tensors` overrides the random init (name -> C-order array)."""
    filt = mel_filterbank(hp.n_mels)
    with open(path, "wb") as f:
        f.write(struct.pack("<I", MAGIC))
        f.write(struct.pack("<11i", *hp.as_list()))
        f.write(struct.pack("<2i", hp.n_mels, filt.shape[1]))
        f.write(filt.tobytes())
        nv = hp.n_vocab if n_vocab_file is None else n_vocab_file
        f.write(struct.pack("<i", nv))
        # dummy vocab: token ids as short ascii words (kernels never look at token text)
        chunk = bytearray()
        for i in range(nv):
            w = b"t%d" % i
            chunk += struct.pack("<I", len(w)) + w
        f.write(chunk)
        src = tensors.items() if tensors is not None else random_tensors(hp, seed)
        for name, a in src:
            ne = tuple(reversed(a.shape))
            ftype = 1 if a.dtype == np.float16 else 0
            nb = name.encode()
            f.write(struct.pack("<3i", len(ne), len(nb), ftype))
            f.write(struct.pack("<%di" % len(ne), *ne))
            f.write(nb)
            f.write(np.ascontiguousarray(a).tobytes())


@dataclass
class ModelFile:
    hparams: HParams
    filters: np.ndarray            # [n_mel][n_fft]
    vocab: List[bytes]
    tensors: Dict[str, np.ndarray]  # C-order arrays (shape = reversed ne[])


def read_model(path: str) -> ModelFile:
    """Host-side reader mirroring WhisperContext::new / WhisperModel::load (src/main.rs:366-503,
    809-1483), reading to true EOF."""
    with open(path, "rb") as f:
        buf = f.read()
    off = 0

    def take(fmt: str):
        nonlocal off
        v = struct.unpack_from(fmt, buf, off)
        off += struct.calcsize(fmt)
        return v

    (magic,) = take("<I")
    if magic != MAGIC:
        raise ValueError(f"invalid model file '{path}' (bad magic)")
    hp = HParams(*take("<11i"))
    n_mel, n_fft = take("<2i")
    filt = np.frombuffer(buf, dtype="<f4", count=n_mel * n_fft, offset=off).reshape(n_mel, n_fft).copy()
    off += 4 * n_mel * n_fft
    (nv,) = take("<i")
    vocab = []
    for _ in range(nv):
        (ln,) = take("<I")
        vocab.append(bytes(buf[off:off + ln]))
        off += ln
    expect = {n: (ne, k) for n, ne, k in tensor_table(hp)}
    tensors: Dict[str, np.ndarray] = {}
    while off < len(buf):
        n_dims, name_len, ftype = take("<3i")
        ne = take("<%di" % n_dims)
        name = bytes(buf[off:off + name_len]).decode()
        off += name_len
        if name not in expect:
            raise KeyError(f"unknown tensor '{name}' in model file")
        ene, _ = expect[name]
        if int(np.prod(ne)) != int(np.prod(ene)):
            raise ValueError(f"tensor {name} has wrong size in model file")
        if tuple(ne) != tuple(ene):
            raise ValueError(f"tensor {name} has wrong shape in model file, got:{ne}, expected:{ene}")
        dt = np.dtype("<f2") if ftype == 1 else np.dtype("<f4")
        cnt = int(np.prod(ne))
        tensors[name] = np.frombuffer(buf, dtype=dt, count=cnt, offset=off).reshape(tuple(reversed(ne))).copy()
        off += cnt * dt.itemsize
    return ModelFile(hp, filt, vocab, tensors)


def arch_seed(arch: str) -> int:
    return 20260 + ARCH_SEED_INDEX[arch]


def make_model(path: str, arch: str) -> HParams:
    hp = ARCHS[arch]
    write_model(path, hp, arch_seed(arch))
    return hp
