"""Positional-embedding resampling as a fixed linear operator (host-side, numpy).

DinoVisionTransformer.interpolate_pos_encoding (LT/_models/dinov2_vit/dinov2_vit_src/models/
vision_transformer.py:251-305) calls F.interpolate(mode="bicubic", scale_factor=(w0+offset)/M | size=...,
antialias=...) on the [M, M] grid of patch embeddings in EVERY forward.  Bicubic resampling is linear in its
input, so for a given (grid, crop size) it is a constant matrix W [out*out, M*M]; the device work per step
is one tiny matmul (and its transpose in the backward).  The weights below restate ATen's
upsample_bicubic2d (A=-0.75, align_corners=False, scale passed through when scale_factor is given) and its
anti-aliased variant (_upsample_bicubic2d_aa, A=-0.5, separable, normalised taps).
"""
from __future__ import annotations

import math

import numpy as np


def _cc1(x: np.ndarray, A: float) -> np.ndarray:
    return ((A + 2.0) * x - (A + 3.0)) * x * x + 1.0


def _cc2(x: np.ndarray, A: float) -> np.ndarray:
    return ((A * x - 5.0 * A) * x + 8.0 * A) * x - 4.0 * A


def bicubic_axis_matrix(n_in: int, n_out: int, scale_factor: float | None) -> np.ndarray:
    """[n_out, n_in] weights of 1-D bicubic (A=-0.75) resampling as ATen computes them (float32 math)."""
    A = np.float32(-0.75)
    scale = np.float32(1.0 / scale_factor) if scale_factor else np.float32(n_in / n_out)
    W = np.zeros((n_out, n_in), dtype=np.float32)
    for o in range(n_out):
        src = scale * np.float32(o + 0.5) - np.float32(0.5)
        ix = math.floor(float(src))
        t = np.float32(src - np.float32(ix))
        w = [_cc2(t + np.float32(1.0), A), _cc1(t, A), _cc1(np.float32(1.0) - t, A), _cc2(np.float32(2.0) - t, A)]
        for k in range(4):
            j = min(max(ix - 1 + k, 0), n_in - 1)
            W[o, j] += np.float32(w[k])
    return W


def _aa_filter(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    if x < 2.0:
        return (((x - 5.0) * x + 8.0) * x - 4.0) * a
    return 0.0


def bicubic_aa_axis_matrix(n_in: int, n_out: int, scale_factor: float | None) -> np.ndarray:
    """[n_out, n_in] weights of ATen's anti-aliased bicubic (separable, A=-0.5)."""
    scale = (1.0 / scale_factor) if scale_factor else (n_in / n_out)
    support = 2.0 * scale if scale >= 1.0 else 2.0
    invscale = 1.0 / scale if scale >= 1.0 else 1.0
    W = np.zeros((n_out, n_in), dtype=np.float64)
    for o in range(n_out):
        center = scale * (o + 0.5)
        xmin = max(0, int(center - support + 0.5))
        xsize = min(n_in, int(center + support + 0.5)) - xmin
        ws = np.array([_aa_filter((j + xmin - center + 0.5) * invscale) for j in range(xsize)])
        tot = ws.sum()
        if tot != 0.0:
            ws = ws / tot
        W[o, xmin:xmin + xsize] = ws
    return W.astype(np.float32)


def pos_embed_operator(M: int, w0: int, h0: int, interpolate_offset: float, antialias: bool) -> np.ndarray:
    """W [w0*h0, M*M] with pos_out[(i2,i3)] = sum W[(i2,i3),(m2,m3)] pos_in[(m2,m3)] (vision_transformer.py:276-302)."""
    if interpolate_offset:
        s2, s3 = float(w0 + interpolate_offset) / M, float(h0 + interpolate_offset) / M
    else:
        s2 = s3 = None
    # with scale factors the output size is floor(M*scale); the reference asserts it equals (w0, h0)
    f = bicubic_aa_axis_matrix if antialias else bicubic_axis_matrix
    W2, W3 = f(M, w0, s2), f(M, h0, s3)
    return np.einsum("ab,cd->acbd", W2, W3).reshape(w0 * h0, M * M).astype(np.float32)
