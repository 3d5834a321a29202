"""synth.py — synthetic int16 waveform generators for the benchmark configs of
BASELINE.md §4 / SURVEY.md §8d (there is no network for real Nab data).

nab_like():  baseline N(0,50) + noise N(0,3) + one exponential pulse per waveform
             (A~U[200,3200], tau=L/10, t0~U[L/4,3L/4]), clipped to 14 bit — the "Nab-like"
             shape of reference docs/Performance.md:14.
gaussian_mix(): equal parts Gaussian sigma in {1,3,10,30,100,1000} (config C4).
"""
from __future__ import annotations

import numpy as np


def nab_like(n_waves: int, L: int, seed: int = 20251018) -> np.ndarray:
    rng = np.random.default_rng(seed)
    out = np.empty((n_waves, L), dtype=np.int16)
    t = np.arange(L, dtype=np.float32)
    tau = np.float32(L / 10.0)
    step = max(1, min(n_waves, (1 << 22) // max(L, 1)))
    for s in range(0, n_waves, step):
        m = min(step, n_waves - s)
        b = rng.normal(0.0, 50.0, (m, 1)).astype(np.float32)
        noise = rng.normal(0.0, 3.0, (m, L)).astype(np.float32)
        t0 = rng.uniform(L / 4.0, 3.0 * L / 4.0, (m, 1)).astype(np.float32)
        A = rng.uniform(200.0, 3200.0, (m, 1)).astype(np.float32)
        dt = t[None, :] - t0
        pulse = np.where(dt >= 0, A * np.exp(-np.maximum(dt, 0) / tau), np.float32(0))
        x = np.rint(b + noise + pulse)
        out[s:s + m] = np.clip(x, -8192, 8191).astype(np.int16)
    return out


def nab_like_torch(n_waves: int, L: int, seed: int, device):
    """Same distribution generated on the GPU with torch (for multi-GB bench inputs)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n_waves, L), dtype=torch.int16, device=device)
    t = torch.arange(L, dtype=torch.float32, device=device)
    tau = L / 10.0
    step = max(1, min(n_waves, (1 << 25) // max(L, 1)))
    for s in range(0, n_waves, step):
        m = min(step, n_waves - s)
        b = torch.randn((m, 1), generator=g, device=device) * 50.0
        noise = torch.randn((m, L), generator=g, device=device) * 3.0
        t0 = (torch.rand((m, 1), generator=g, device=device) * 0.5 + 0.25) * L
        A = torch.rand((m, 1), generator=g, device=device) * 3000.0 + 200.0
        dt = t[None, :] - t0
        pulse = torch.where(dt >= 0, A * torch.exp(-dt.clamp_min(0) / tau), torch.zeros((), device=device))
        x = torch.round(b + noise + pulse).clamp_(-8192, 8191)
        out[s:s + m] = x.to(torch.int16)
    return out


def gaussian_mix(n_waves: int, L: int, sigmas=(1, 3, 10, 30, 100, 1000), seed: int = 7) -> np.ndarray:
    rng = np.random.default_rng(seed)
    out = np.empty((n_waves, L), dtype=np.int16)
    for i in range(n_waves):
        s = sigmas[i % len(sigmas)]
        out[i] = np.clip(np.rint(rng.normal(0, s, L)), -32768, 32767).astype(np.int16)
    return out


def gaussian_mix_torch(n_waves: int, L: int, sigmas=(1, 3, 10, 30, 100, 1000), seed: int = 7, device="cuda"):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n_waves, L), dtype=torch.int16, device=device)
    sig = torch.tensor(sigmas, dtype=torch.float32, device=device)
    step = max(1, min(n_waves, (1 << 25) // max(L, 1)))
    for s in range(0, n_waves, step):
        m = min(step, n_waves - s)
        sg = sig[(torch.arange(s, s + m, device=device) % len(sigmas))][:, None]
        x = torch.round(torch.randn((m, L), generator=g, device=device) * sg).clamp_(-32768, 32767)
        out[s:s + m] = x.to(torch.int16)
    return out
