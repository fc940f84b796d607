"""Seeded synthetic images shared by the tests and bench.py."""
from __future__ import annotations

import numpy as np


def uniform(h, w, c, seed=1234):
    return np.random.default_rng(seed).integers(0, 256, size=(h, w, c), dtype=np.uint8)


def smooth(h, w, c, seed=1234, noise=6):
    """Low-contrast gradient + small noise: keeps Sobel away from the 255 clamp."""
    rng = np.random.default_rng(seed)
    y = np.linspace(0.0, 1.0, h, dtype=np.float32)[:, None, None]
    x = np.linspace(0.0, 1.0, w, dtype=np.float32)[None, :, None]
    ch = np.arange(c, dtype=np.float32)[None, None, :]
    base = 96.0 + 60.0 * np.sin(6.0 * x + ch) * np.cos(4.0 * y) + 40.0 * x * y
    img = base + rng.integers(-noise, noise + 1, size=(h, w, c)).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def white_square(h, w, c):
    """The reference's only synthetic fixture (tests/test_gaussian_blur.cu:22-36)."""
    img = np.zeros((h, w, c), dtype=np.uint8)
    s = w // 4
    y0, x0 = (h - s) // 2, (w - s) // 2
    img[max(y0, 0):y0 + s, max(x0, 0):x0 + s, :] = 255
    return img


def constant(h, w, c, v):
    return np.full((h, w, c), v, dtype=np.uint8)


KINDS = {"uniform": uniform, "smooth": smooth}
