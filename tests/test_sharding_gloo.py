"""World-size-2 gloo test (CPU) of clip sharding + the one result-collection all-gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ufvideo_b200 import sharding


def test_clip_blocks_cover_all_clips_in_order():
    for n in (0, 1, 7, 8, 64):
        for w in (1, 2, 3, 8):
            got = [c for r in range(w) for c in sharding.clip_block(n, r, w)]
            assert got == list(range(n))
            sizes = [len(sharding.clip_block(n, r, w)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1


def test_payload_roundtrip_single_process():
    tok = torch.arange(5 * 16, dtype=torch.float32).reshape(5, 16)
    counts = torch.tensor([2, 3], dtype=torch.int32)
    p = sharding.pack_payload(tok, counts, pad_rows=8, pad_objs=4)
    t, c = sharding.unpack_payloads(p[None], 8)
    assert torch.equal(t, tok) and c == [2, 3]
    with pytest.raises(ValueError):
        sharding.pack_payload(tok, counts, pad_rows=4, pad_objs=4)


def _worker(rank, world, port, n_clips, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hid, k, objs = 32, 4, 3
        mine = sharding.clip_block(n_clips, rank, world)
        # per-clip "tokens": rows filled with the clip id; clip c has (c % k) + 1 tokens per object
        rows, counts = [], []
        for c in mine:
            for o in range(objs):
                n = (c + o) % k + 1
                rows.append(torch.full((n, hid), float(c * 10 + o)))
                counts.append(n)
        tokens = torch.cat(rows) if rows else torch.zeros((0, hid))
        pad_clips = -(-n_clips // world)
        gathered = sharding.all_gather_tokens(tokens, torch.tensor(counts, dtype=torch.int32),
                                              pad_rows=pad_clips * objs * k, pad_objs=pad_clips * objs)
        all_tokens, all_counts = sharding.unpack_payloads(gathered, pad_clips * objs * k)
        # the zero-copy variant: "kernels" write padded rows and counts straight into the payload
        slots = np.full(len(counts), k, np.int32)
        payload, tok_view, cnt_view = sharding.new_payload(slots, pad_clips * objs * k, pad_clips * objs, hid,
                                                           torch.float32, torch.device("cpu"))
        tok_view.zero_()
        off = 0
        for row, n in zip(rows, counts):
            tok_view[off:off + n] = row
            off += k
        cnt_view.copy_(torch.tensor(counts, dtype=torch.int32))
        g2, _ = sharding.all_gather_payload(payload)
        padded_tokens, padded_counts = sharding.unpack_padded(g2, pad_clips * objs * k, pad_clips * objs)
        assert padded_counts == all_counts and torch.equal(padded_tokens, all_tokens)
        torch.save((all_tokens, all_counts), os.path.join(result_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_all_gather_restores_global_clip_order(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n_clips, world = 5, 2
    mp.spawn(_worker, args=(world, port, n_clips, str(tmp_path)), nprocs=world, join=True)
    want_rows, want_counts = [], []
    for c in range(n_clips):
        for o in range(3):
            n = (c + o) % 4 + 1
            want_rows.append(torch.full((n, 32), float(c * 10 + o)))
            want_counts.append(n)
    want = torch.cat(want_rows)
    for r in range(world):
        tokens, counts = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert counts == want_counts and torch.equal(tokens, want)
