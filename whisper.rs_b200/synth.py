"""Seeded synthetic 16 kHz PCM (SURVEY.md section 8d "Synthetic inputs").

Each 30 s segment is a mix of 3-5 sinusoids (80 Hz - 7 kHz, amplitude 0.05-0.3), white noise
(sigma 0.02) and a silent last 1.5 s, which exercises the `max - 8` clamp of
clamp_and_normalize (src/main.rs:1654-1671) and the zero fill past the end of the clip
(src/main.rs:1596-1600).  Values stay inside [-1, 1).
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000     # src/main.rs:25
N_FFT = 400             # src/main.rs:26
HOP = 160               # src/main.rs:28
CHUNK_S = 30            # src/main.rs:29
SEG_SAMPLES = SAMPLE_RATE * CHUNK_S


def make_segment(seg: int, n_samples: int = SEG_SAMPLES, silent_tail_s: float = 1.5) -> np.ndarray:
    rng = np.random.default_rng(1000 + seg)
    t = np.arange(n_samples, dtype=np.float64) / SAMPLE_RATE
    x = np.zeros(n_samples, dtype=np.float64)
    for _ in range(int(rng.integers(3, 6))):
        f = float(rng.uniform(80.0, 7000.0))
        a = float(rng.uniform(0.05, 0.3))
        ph = float(rng.uniform(0, 2 * np.pi))
        # slow amplitude modulation so frames differ from each other
        fm = float(rng.uniform(0.2, 3.0))
        x += a * (0.6 + 0.4 * np.sin(2 * np.pi * fm * t)) * np.sin(2 * np.pi * f * t + ph)
    x += 0.02 * rng.standard_normal(n_samples)
    n_tail = min(n_samples, int(silent_tail_s * SAMPLE_RATE))
    if n_tail:
        x[n_samples - n_tail:] = 0.0
    return np.clip(x, -1.0, 1.0 - 2.0 ** -15).astype(np.float32)


def make_clips(n_clips: int, first_seg: int = 0, n_samples: int = SEG_SAMPLES) -> np.ndarray:
    """[n_clips][n_samples] f32, clip c uses seed 1000 + first_seg + c."""
    return np.stack([make_segment(first_seg + c, n_samples) for c in range(n_clips)])


def make_long_clip(n_segments: int, first_seg: int = 0) -> np.ndarray:
    """One clip of n_segments x 30 s (config 5: 120 segments = 1 h)."""
    return np.concatenate([make_segment(first_seg + s, silent_tail_s=0.5) for s in range(n_segments)])
