"""The PyTorch ``grid_sample`` formulation of the op (what the reference falls back to,
OPS/functions/ms_deform_attn_func.py:52-72), for GPU baselines in the benchmark scripts only."""
import torch.nn.functional as F


def grid_sample_core(value, shapes, loc, attn):
    N, S, M, D = value.shape
    Lq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    sizes = [(int(h), int(w)) for h, w in shapes.tolist()]
    grid = loc * 2 - 1
    w = attn.permute(0, 2, 1, 3, 4).reshape(N * M, 1, Lq, L, P)
    out, start = None, 0
    for lvl, (H, W) in enumerate(sizes):
        v = value[:, start:start + H * W].permute(0, 2, 3, 1).reshape(N * M, D, H, W)
        start += H * W
        g = grid[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(N * M, Lq, P, 2)
        s = F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False)
        term = (s * w[:, :, :, lvl]).sum(-1)
        out = term if out is None else out + term
    return out.view(N, M * D, Lq).transpose(1, 2).contiguous()
