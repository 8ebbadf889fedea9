"""Seeded synthetic inputs for the benchmarks: images with a post-Normalize distribution and blobby binary water masks
(SURVEY.md §8d).  Integer-hash based (no torch RNG), so every rank / run / torch version sees the same data.  The parity
oracle has its own copy of this generator; tests/test_abi_cpu.py checks that both produce identical tensors."""
import numpy as np
import torch
import torch.nn.functional as F


def hash_uniform(n: int, seed: int) -> np.ndarray:
    """n float64 values in [0,1): splitmix64 finaliser over the index, 53 mantissa bits."""
    with np.errstate(over="ignore"):
        z = np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def synthetic_batch(batch, channels, h, w, seed=123):
    """(images fp32 [B,C,H,W] ~ N(0,1) by Box-Muller, masks fp32 [B,1,H,W] in {0,1}: thresholded 8x bilinearly
    up-sampled low-resolution noise, ~50 % water in connected blobs)."""
    n = batch * channels * h * w
    u1, u2 = hash_uniform(n, seed), hash_uniform(n, seed + 17)
    x = np.sqrt(-2.0 * np.log(np.maximum(u1, 1e-12))) * np.cos(2 * np.pi * u2)
    x = torch.from_numpy(x.astype(np.float32).reshape(batch, channels, h, w))
    lh, lw = max(h // 8, 1), max(w // 8, 1)
    low = hash_uniform(batch * lh * lw, seed + 99).reshape(batch, 1, lh, lw)
    m = F.interpolate(torch.from_numpy(low.astype(np.float32)), size=(h, w), mode="bilinear", align_corners=False)
    return x, (m > 0.5).float()
