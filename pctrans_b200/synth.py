"""Synthetic inputs for the MSDeformAttn hot path (bench + tests; no dataset, no network).

Everything is generated on the CPU from a seeded ``torch.Generator`` so CPU and GPU legs see
identical bits.  Three sampling-location regimes:

* ``"R"`` realistic (headline): encoder reference points (PD/msdeformattn.py:141-153 with
  valid_ratio 1) plus ``(compass_bias(m,p) + N(0, 1 px)) / (W_l, H_l)``, i.e. what
  ``MSDeformAttn.forward`` produces right after ``_reset_parameters`` (mod.py:66-74,106-109)
  with one pixel of learned jitter.  Requires Lq == S.
* ``"U"`` uniform in [0,1) as in OPS/test.py:37 -- worst-case locality.
* ``"O"`` uniform in [-0.25, 1.25] -- exercises the out-of-range predicate (cuh:293).
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import torch

# BASELINE.json pyramids, low -> high resolution (the order PD/msdeformattn.py:319 produces).
CVPPP_512 = ((16, 16), (32, 32), (64, 64))
CVPPP_448 = ((14, 14), (28, 28), (56, 56))
BBBC_520x696 = ((17, 22), (33, 44), (65, 87))


def level_tensors(shapes: Sequence[Tuple[int, int]], device="cpu"):
    """(spatial_shapes (L,2) int64, level_start_index (L,) int64, S) as the encoder builds them
    (PD/msdeformattn.py:82-83)."""
    spatial_shapes = torch.as_tensor(list(shapes), dtype=torch.long, device=device).reshape(-1, 2)
    level_start_index = torch.cat(
        (spatial_shapes.new_zeros((1,)), spatial_shapes.prod(1).cumsum(0)[:-1]))
    return spatial_shapes, level_start_index, int(spatial_shapes.prod(1).sum())


def encoder_reference_points(shapes: Sequence[Tuple[int, int]]) -> torch.Tensor:
    """(S, 2) normalised (x, y) pixel centres of every pyramid position, valid_ratio = 1."""
    pts = []
    for H, W in shapes:
        ys = (torch.arange(H, dtype=torch.float32) + 0.5) / H
        xs = (torch.arange(W, dtype=torch.float32) + 0.5) / W
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        pts.append(torch.stack((xx.reshape(-1), yy.reshape(-1)), -1))
    return torch.cat(pts, 0)


def compass_bias(n_heads: int, n_levels: int, n_points: int) -> torch.Tensor:
    """(M, L, P, 2) initial sampling-offset bias of MSDeformAttn (mod.py:66-74)."""
    thetas = torch.arange(n_heads, dtype=torch.float32) * (2.0 * math.pi / n_heads)
    g = torch.stack([thetas.cos(), thetas.sin()], -1)
    g = g / g.abs().max(-1, keepdim=True)[0]
    g = g.view(n_heads, 1, 1, 2).repeat(1, n_levels, n_points, 1)
    scale = torch.arange(1, n_points + 1, dtype=torch.float32).view(1, 1, n_points, 1)
    return g * scale


def make_inputs(N: int, shapes: Sequence[Tuple[int, int]], M: int = 8, D: int = 32, P: int = 4,
                regime: str = "R", seed: int = 0, Lq: int | None = None, value_scale: float = 1.0,
                jitter_px: float = 1.0):
    """Returns fp32 CPU tensors (value, spatial_shapes, level_start_index, loc, attn)."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    spatial_shapes, lsi, S = level_tensors(shapes)
    L = len(shapes)
    if Lq is None:
        Lq = S
    value = torch.randn(N, S, M, D, generator=gen) * value_scale
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=gen), -1).view(N, Lq, M, L, P)
    if regime == "R":
        if Lq != S:
            raise ValueError("regime R needs Lq == S (encoder self-attention)")
        ref = encoder_reference_points(shapes)                           # (S,2)
        off = compass_bias(M, L, P)[None, None] + jitter_px * torch.randn(
            N, Lq, M, L, P, 2, generator=gen)
        norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32)   # (L,2) = (W,H)
        loc = ref[None, :, None, None, None, :] + off / norm[None, None, None, :, None, :]
    elif regime == "U":
        loc = torch.rand(N, Lq, M, L, P, 2, generator=gen)
    elif regime == "O":
        loc = torch.rand(N, Lq, M, L, P, 2, generator=gen) * 1.5 - 0.25
    else:
        raise ValueError(f"unknown regime {regime!r}")
    return value, spatial_shapes, lsi, loc.contiguous(), attn.contiguous()


def nudge_off_grid(loc: torch.Tensor, shapes: Sequence[Tuple[int, int]], eps: float = 1e-3):
    """Move sampling locations whose pixel coordinate ``loc*size - 0.5`` lies within ``eps`` of
    an integer away from it.  The op is not differentiable w.r.t. the location on the pixel
    grid (the bilinear cell changes there), so an fp32 and an fp64 evaluation of the same bits
    may legitimately pick different cells; gradient parity tests use inputs off the grid."""
    loc = loc.clone()
    for lvl, (H, W) in enumerate(shapes):
        for axis, size in ((0, W), (1, H)):
            c = loc[:, :, :, lvl, :, axis].double() * size - 0.5
            near = (c - c.round()).abs() < eps
            loc[:, :, :, lvl, :, axis] = torch.where(
                near, loc[:, :, :, lvl, :, axis] + 3.0 * eps / size, loc[:, :, :, lvl, :, axis])
    return loc
