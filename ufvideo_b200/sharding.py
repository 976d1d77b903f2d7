"""Clip sharding over ranks and the one result-collection collective.

The reference's eval drivers give each rank ``get_chunk(questions, num_chunks, rank)`` and merge
per-rank JSON files afterwards (eval/inference_PixRQA.py:186,214); clips never interact
(layer.py:68 loops over samples).  Here: contiguous blocks of clips per rank, no communication
during compute, and ONE all-gather (NCCL over NVLink on GPUs) of the padded per-rank token
tensor, with the int32 token counts riding in the tail rows of the same buffer.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def clip_block(n_clips: int, rank: int, world: int) -> range:
    """Contiguous block of clips owned by ``rank`` (sizes differ by at most one), so that
    concatenating ranks in order restores global clip order."""
    base, extra = divmod(n_clips, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def _tail_rows(pad_objs: int, row_bytes: int) -> int:
    return -(-(4 * (2 + pad_objs)) // row_bytes)


def pack_payload(tokens: torch.Tensor, counts: torch.Tensor, pad_rows: int, pad_objs: int) -> torch.Tensor:
    """[pad_rows + tail, hid] buffer: token rows first, then int32 [m, n_obj, counts...]."""
    m, hid = tokens.shape
    n_obj = counts.numel()
    if m > pad_rows or n_obj > pad_objs:
        raise ValueError(f"payload overflow: {m} rows / {n_obj} objects > pad {pad_rows} / {pad_objs}")
    row_bytes = hid * tokens.element_size()
    tail = _tail_rows(pad_objs, row_bytes)
    payload = torch.zeros((pad_rows + tail, hid), dtype=tokens.dtype, device=tokens.device)
    payload[:m] = tokens
    meta = payload[pad_rows:].view(torch.int32).reshape(-1)
    meta[0] = m
    meta[1] = n_obj
    meta[2:2 + n_obj] = counts.to(torch.int32)
    return payload


def unpack_payloads(gathered: torch.Tensor, pad_rows: int):
    """gathered [world, pad_rows + tail, hid] -> (tokens [sum m_r, hid], counts list[int])."""
    meta = gathered[:, pad_rows:].reshape(gathered.shape[0], -1).view(torch.int32).cpu().numpy()
    rows, counts = [], []
    for r in range(gathered.shape[0]):
        m, n_obj = int(meta[r, 0]), int(meta[r, 1])
        rows.append(gathered[r, :m])
        counts.extend(int(c) for c in meta[r, 2:2 + n_obj])
    return torch.cat(rows, dim=0), counts


def all_gather_tokens(tokens: torch.Tensor, counts: torch.Tensor, pad_rows: int, pad_objs: int,
                      group=None):
    """One all-gather of every rank's object tokens.  ``pad_rows`` / ``pad_objs`` are upper bounds
    every rank agrees on (e.g. max clips per rank x objects x K).  Returns the raw gathered buffer
    [world, rows, hid]; ``unpack_payloads`` turns it into (tokens, counts) in global clip order."""
    payload = pack_payload(tokens, counts, pad_rows, pad_objs)
    world = dist.get_world_size(group)
    rows, hid = payload.shape
    out = torch.empty((world * rows, hid), dtype=payload.dtype, device=payload.device)
    dist.all_gather_into_tensor(out, payload, group=group)     # concatenated along dim 0
    return out.view(world, rows, hid)
