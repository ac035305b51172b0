"""Synthetic benchmark frames: smooth multi-sinusoid texture, second frame = first frame
shifted by a sub-pixel (dx, dy), both re-quantised to uint8 values.

Modelled on the reference's fallback texture (python/generate_test_frames_natural.py:49-64,
sum of sinusoids -> uint8) and its motion model (:67-73, shifted copy).  uint8-valued
float32 frames are what the verifier feeds the path (optical_flow_verifier.py:61-65), and
they make the fast kernel's separable sums exact.

`make_pairs_numpy` is the CPU generator (tests, small cases); `make_pairs_torch` builds the
same kind of batch directly in device memory for the benchmarks.
"""

from __future__ import annotations

import numpy as np

_FREQS = ((0.021, 0.013, 38.0), (0.057, 0.031, 27.0), (0.113, 0.089, 19.0), (0.241, 0.197, 11.0), (0.47, 0.53, 6.0))


def make_pairs_numpy(batch: int, height: int, width: int, seed: int = 0):
    """-> prev, curr float32 [B, H, W] with integer values in [0, 255], and the (dx, dy) used."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float64)
    prev = np.empty((batch, height, width), np.float32)
    curr = np.empty((batch, height, width), np.float32)
    shifts = rng.uniform(-1.0, 1.0, size=(batch, 2))
    for b in range(batch):
        ph = rng.uniform(0, 2 * np.pi, size=(len(_FREQS), 2))

        def tex(x, y):
            t = np.full(x.shape, 128.0)
            for (fx, fy, amp), (p0, p1) in zip(_FREQS, ph):
                t += amp * np.sin(fx * x + p0) * np.cos(fy * y + p1)
            return np.clip(np.rint(t), 0, 255)

        dx, dy = shifts[b]
        prev[b] = tex(xx, yy)
        curr[b] = tex(xx - dx, yy - dy)  # content moves by (+dx, +dy)
    return prev, curr, shifts.astype(np.float32)


def make_pairs_torch(batch: int, height: int, width: int, device, seed: int = 0, chunk: int = 8):
    """Same texture family generated with torch on `device` (float32 [B, H, W] tensors)."""
    import torch

    gen = torch.Generator(device="cpu").manual_seed(seed)
    shifts = (torch.rand((batch, 2), generator=gen, dtype=torch.float64) * 2.0 - 1.0)
    phases = torch.rand((batch, len(_FREQS), 2), generator=gen, dtype=torch.float64) * (2 * np.pi)
    prev = torch.empty((batch, height, width), dtype=torch.float32, device=device)
    curr = torch.empty((batch, height, width), dtype=torch.float32, device=device)
    yy = torch.arange(height, device=device, dtype=torch.float32).view(1, height, 1)
    xx = torch.arange(width, device=device, dtype=torch.float32).view(1, 1, width)
    for b0 in range(0, batch, chunk):
        b1 = min(b0 + chunk, batch)
        ph = phases[b0:b1].to(device=device, dtype=torch.float32)
        sh = shifts[b0:b1].to(device=device, dtype=torch.float32)
        for dst, moved in ((prev, False), (curr, True)):
            x = xx - sh[:, 0].view(-1, 1, 1) if moved else xx
            y = yy - sh[:, 1].view(-1, 1, 1) if moved else yy
            t = torch.full((b1 - b0, height, width), 128.0, device=device)
            for k, (fx, fy, amp) in enumerate(_FREQS):
                t += amp * torch.sin(fx * x + ph[:, k, 0].view(-1, 1, 1)) * torch.cos(fy * y + ph[:, k, 1].view(-1, 1, 1))
            dst[b0:b1] = torch.clamp(torch.round(t), 0, 255)
    return prev, curr, shifts.to(torch.float32)
