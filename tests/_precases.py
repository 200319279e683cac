"""Seeded synthetic images for the pre-processing parity tests (small, so the golden file stays small)."""
import numpy as np

# (h0, w0, new_shape, auto, scaleup)
CASES = [(97, 131, 160, True, True), (240, 180, 160, True, True), (64, 64, 96, False, True), (50, 200, 128, True, True),
         (300, 300, 160, False, False), (33, 47, 160, True, True), (120, 160, 160, True, True), (160, 120, 160, False, True),
         (75, 75, 160, True, False), (333, 500, 320, False, False)]


def image(i: int, h: int, w: int) -> np.ndarray:
    rng = np.random.default_rng(1000 + i)
    base = rng.integers(0, 256, (h // 4 + 2, w // 4 + 2, 3), dtype=np.uint8)
    img = np.kron(base, np.ones((4, 4, 1), dtype=np.uint8))[:h, :w]      # blocky structure + noise
    noise = rng.integers(-20, 21, (h, w, 3))
    return np.clip(img.astype(int) + noise, 0, 255).astype(np.uint8)
