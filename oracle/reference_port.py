"""Torch-CPU port of the reference's object encoder, op for op.  TEST INFRASTRUCTURE ONLY.

Purpose: the CPU baseline that ``bench.py`` times beside the CUDA path (``cpu_baseline`` and
``--impl reference``, kind "port").  The real reference (a Python module under /root/reference)
cannot travel to the GPU box, so this port issues the same ATen calls in the same order as
``ufvideo/model/layer.py`` -- dense gather, fp32 upcast, bilinear interpolate, the three dense
[q, C, 27, 27] temporaries of the masked mean, the per-object python merge loop with one
comparison per frame, Linear / GELU / Linear -- which is what sets the reference's CPU cost
(index 25 %, mul 20 %, div 19 %, sum 17 %, addmm 9 %).  It is validated against the
real reference in the build container by tests/test_oracle_vs_reference.py (bit-identical
outputs under the same torch build) and is never imported by the product package.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def merge_tokens(x: torch.Tensor, n_remove: int) -> torch.Tensor:
    """layer.py:6-33 -- x [1, n, d]; threshold grouping of adjacent tokens."""
    left = F.normalize(x[:, :-1, :], p=2, dim=-1)
    right = F.normalize(x[:, 1:, :], p=2, dim=-1)
    sim = torch.sum(left * right, dim=-1)
    kth = torch.topk(sim.flatten(), n_remove).values[-1]
    groups, run = [], []
    for i in range(sim.shape[1]):
        run.append(x[:, i:i + 1, :])
        if sim[0, i] < kth:
            groups.append(torch.mean(torch.cat(run, dim=1), dim=1, keepdim=True))
            run = []
    run.append(x[:, sim.shape[1]:sim.shape[1] + 1, :])
    groups.append(torch.mean(torch.cat(run, dim=1), dim=1, keepdim=True))
    return torch.cat(groups, dim=1)


def pool_masks(x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """layer.py:135-152 -- x [b, C, h, w], mask [1, q, H, W] -> [q, C]."""
    if x.shape[-2:] != mask.shape[-2:]:
        mask = F.interpolate(mask, size=x.shape[-2:], mode="bilinear", align_corners=False)
    mask = (mask > 0).to(mask.dtype).permute(1, 0, 2, 3)
    denorm = mask.sum(dim=(-1, -2), keepdim=True) + 1e-8
    return (x * mask / denorm).sum(-1).sum(-1)


def encode(feats: torch.Tensor, masks, ann_indices, k_keep: int, w1, b1, w2, b2, pad_square=False):
    """layer.py:63-128 -- returns (tokens [N_tok, hid] in feats.dtype, list of per-object counts)."""
    per_sample, counts = [], []
    for i in range(len(masks)):
        mask = masks[i].unsqueeze(0).float()
        if mask.shape[1] == 0:
            mask = torch.zeros((1, 1, 336, 336))
        if pad_square:
            h, w = mask.shape[-2:]
            side = max(h, w)
            mask = F.pad(mask, ((side - w) // 2, (side - w) - (side - w) // 2,
                                (side - h) // 2, (side - h) - (side - h) // 2))
        rows = [r for obj in ann_indices[i] for r in obj]
        x = feats[rows]
        n = int(pow(x.shape[1], 0.5))
        x = x.reshape(x.shape[0], n, n, -1).permute(0, 3, 1, 2)
        model_dtype = x.dtype
        pooled = pool_masks(x.to(mask.dtype), mask)
        out, start = [], 0
        for obj in ann_indices[i]:
            tok = pooled[start:start + len(obj), :].unsqueeze(0)
            if tok.shape[1] > k_keep:
                tok = merge_tokens(tok, tok.shape[1] - k_keep)
            counts.append(tok.shape[1])
            out.append(tok)
            start += len(obj)
        per_sample.append(torch.cat(out, dim=1).reshape(-1, pooled.shape[-1]).to(model_dtype))
    tokens = torch.cat(per_sample, dim=0)
    hidden = F.gelu(F.linear(tokens, w1, b1))
    return F.linear(hidden, w2, b2), counts
