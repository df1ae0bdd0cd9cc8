"""Deterministic integer-only synthetic images.

For image index ``idx``: s = 0x9E3779B9*(idx+1); iterate c, y, x (planar, c outermost):
s = s*1664525 + 1013904223 (mod 2^32); base = ((2x+3y)>>3)&255; tri = base<128 ? base : 255-base;
v = 64 + tri + 32*(((x>>5)^(y>>5))&1) + (s>>28) + 8c; pixel = min(v, 255).
The LCG is evaluated with a jump-ahead doubling scheme so numpy stays vectorised.
"""
import numpy as np

_A = 1664525
_C = 1013904223
_M = 0xFFFFFFFF


def _lcg_stream(s0: int, n: int) -> np.ndarray:
    """s_1..s_n of s_{k+1} = A*s_k + C mod 2^32 starting from s_0 (s_0 itself excluded)."""
    out = np.empty(n, dtype=np.uint64)
    if n == 0:
        return out.astype(np.uint32)
    out[0] = (s0 * _A + _C) & _M
    a, c = _A, _C  # jump by `have` steps: s -> a*s + c
    have = 1
    while have < n:
        take = min(have, n - have)
        out[have:have + take] = (out[:take] * np.uint64(a) + np.uint64(c)) & np.uint64(_M)
        c = (a * c + c) & _M
        a = (a * a) & _M
        have += take
    return out.astype(np.uint32)


def synth_image(idx: int, width: int, height: int, channels: int) -> np.ndarray:
    """Planar u8 image of shape (channels, height, width)."""
    n = width * height * channels
    s = _lcg_stream((0x9E3779B9 * (idx + 1)) & _M, n).reshape(channels, height, width)
    x = np.arange(width, dtype=np.int64)[None, None, :]
    y = np.arange(height, dtype=np.int64)[None, :, None]
    c = np.arange(channels, dtype=np.int64)[:, None, None]
    base = ((2 * x + 3 * y) >> 3) & 255
    tri = np.where(base < 128, base, 255 - base)
    v = 64 + tri + 32 * (((x >> 5) ^ (y >> 5)) & 1) + (s >> 28).astype(np.int64) + 8 * c
    return np.minimum(v, 255).astype(np.uint8)


def synth_batch_torch(first: int, n: int, width: int, height: int, channels: int, device):
    """The same images, indices first .. first+n-1, generated with torch on `device` (bench plumbing: 32 distinct 4K
    images take seconds with numpy).  Returns u8 (n, channels, height, width); tests pin it to synth_image."""
    import torch
    N = width * height * channels
    s = torch.empty((n, N), dtype=torch.int64, device=device)
    s0 = torch.tensor([(0x9E3779B9 * (first + i + 1)) & _M for i in range(n)], dtype=torch.int64, device=device)
    s[:, 0] = (s0 * _A + _C) & _M
    a, c, have = _A, _C, 1
    while have < N:
        take = min(have, N - have)
        s[:, have:have + take] = (s[:, :take] * a + c) & _M  # (int64 wrap keeps the low 32 bits right)
        c = (a * c + c) & _M
        a = (a * a) & _M
        have += take
    s = s.view(n, channels, height, width)
    x = torch.arange(width, dtype=torch.int64, device=device)[None, None, None, :]
    y = torch.arange(height, dtype=torch.int64, device=device)[None, None, :, None]
    cc = torch.arange(channels, dtype=torch.int64, device=device)[None, :, None, None]
    base = ((2 * x + 3 * y) >> 3) & 255
    tri = torch.where(base < 128, base, 255 - base)
    v = 64 + tri + 32 * (((x >> 5) ^ (y >> 5)) & 1) + (s >> 28) + 8 * cc
    return torch.clamp(v, max=255).to(torch.uint8)
