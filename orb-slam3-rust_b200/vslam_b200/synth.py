"""Deterministic synthetic EuRoC-shaped stereo pairs (SURVEY.md §8 d, BASELINE.json configs).

Scene = three Gaussian-blurred white-noise octaves (sigma 1/3/9, amplitudes 30/50/60, normalised to
0..255) + 150 filled random rectangles (8..40 px) on a canvas (W+128) wide.  left = canvas crop at x=40;
right = the same canvas shifted per row band (disparities 6/14/30 px for three equal bands, i.e. depth
8.4/3.6/1.7 m with the EuRoC-like calibration below) + independent uniform +-3 noise on both eyes.

The generator is plain torch so the same code makes one pair on the CPU (tests, fixtures) or thousands of
pairs on the GPU (bench).  Images are inputs, not results: GPU path and CPU oracle always consume the SAME
bytes (copied, never regenerated), so CPU/GPU float differences in this file cannot affect parity.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# EuRoC-like calibration used by every config (SURVEY.md §8 d).
EUROC_CAMERA = dict(fx=458.654, fy=457.296, cx=367.215, cy=248.375, baseline=0.110074)

_MARGIN = 128
_CROP_X = 40
_DISPARITIES = (6, 14, 30)


def _gauss_kernel(sigma: float, device) -> torch.Tensor:
    r = int(3 * sigma + 0.5)
    x = torch.arange(-r, r + 1, dtype=torch.float32, device=device)
    k = torch.exp(-(x * x) / (2 * sigma * sigma))
    return k / k.sum()


def _blur(img: torch.Tensor, sigma: float) -> torch.Tensor:
    """img: [1,1,H,W] float32; separable Gaussian with reflect padding."""
    k = _gauss_kernel(sigma, img.device)
    r = (k.numel() - 1) // 2
    img = F.conv2d(F.pad(img, (r, r, 0, 0), mode="reflect"), k.view(1, 1, 1, -1))
    img = F.conv2d(F.pad(img, (0, 0, r, r), mode="reflect"), k.view(1, 1, -1, 1))
    return img


def make_stereo_pair(seed: int, width: int = 752, height: int = 480, device="cpu", extra: int = 0, offsets=(0,)):
    """Returns (left, right) uint8 tensors of shape [height, width] on `device`.
    extra / offsets (config 4): the canvas is `extra` px wider and one pair is cut per offset (camera translated
    `offset` px to the right, independent sensor noise per frame): returns ([n, H, W], [n, H, W])."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(1_000_003 * int(seed) + 12345)
    cw = width + _MARGIN + int(extra)
    scale = width / 752.0  # config 3 (1280x720) scales the scene with the image
    canvas = torch.zeros(1, 1, height, cw, dtype=torch.float32, device=dev)
    for sigma, amp in ((1.0, 30.0), (3.0, 50.0), (9.0, 60.0)):
        n = torch.randn(1, 1, height, cw, generator=g, device=dev)
        b = _blur(n, sigma * scale)
        canvas += amp * b / b.std()
    canvas = canvas[0, 0]
    lo, hi = canvas.min(), canvas.max()
    canvas = (canvas - lo) * (255.0 / (hi - lo))

    rng = np.random.default_rng(int(seed))
    lo_sz, hi_sz = max(2, int(8 * scale)), max(3, int(40 * scale))
    for _ in range(150):
        rw, rh = rng.integers(lo_sz, hi_sz + 1, size=2)
        x0 = int(rng.integers(0, cw - rw))
        y0 = int(rng.integers(0, height - rh))
        canvas[y0 : y0 + int(rh), x0 : x0 + int(rw)] = float(rng.integers(0, 256))

    def cut(off):
        x0 = _CROP_X + int(off)
        left = canvas[:, x0 : x0 + width]
        right = torch.empty_like(left)
        bands = len(_DISPARITIES)
        for i, d in enumerate(_DISPARITIES):
            d = int(round(d * scale))
            y0 = (height * i) // bands
            y1 = (height * (i + 1)) // bands
            right[y0:y1] = canvas[y0:y1, x0 + d : x0 + d + width]
        nl = torch.rand(height, width, generator=g, device=dev) * 6.0 - 3.0
        nr = torch.rand(height, width, generator=g, device=dev) * 6.0 - 3.0
        left = (left + nl).round().clamp(0, 255).to(torch.uint8)
        right = (right + nr).round().clamp(0, 255).to(torch.uint8)
        return left.contiguous(), right.contiguous()

    if extra == 0 and tuple(offsets) == (0,):
        return cut(0)
    pairs = [cut(o) for o in offsets]
    return torch.stack([p[0] for p in pairs]), torch.stack([p[1] for p in pairs])


def make_sequence(seed: int, n_frames: int, step_px: int = 2, width: int = 752, height: int = 480, device="cpu"):
    """BASELINE config 4: one scene, the camera translated `step_px` px per frame (image content moves LEFT by step_px per
    frame: a point seen at (u, v) in frame 0 is at (u - step_px * t, v) in frame t).  -> ([n, H, W], [n, H, W]) uint8."""
    return make_stereo_pair(seed, width, height, device, extra=step_px * (n_frames - 1) + 2,
                            offsets=[step_px * t for t in range(n_frames)])


def make_stereo_batch(seeds, width: int = 752, height: int = 480, device="cpu"):
    """Returns (left, right) uint8 tensors of shape [B, height, width]."""
    ls, rs = [], []
    for s in seeds:
        l, r = make_stereo_pair(int(s), width, height, device)
        ls.append(l)
        rs.append(r)
    return torch.stack(ls), torch.stack(rs)
