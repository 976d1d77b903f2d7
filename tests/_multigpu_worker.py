"""Worker of tests/test_multigpu.py: one process per GPU under torch.distributed.run (NCCL).

Checks, on every rank:
  1. the all-gather fused into the last Linear (PeerGather: multimem.st over NVSwitch multicast when the
     fabric offers it, one st per peer otherwise) and its side-stream variant (PeerGather.push of locally written
     rows) deliver bit-identical gathered buffers to an NCCL all-gather of the same payload, over several steps
     with different data and ragged / empty shards;
  2. world-size invariance: tokens and counts gathered from the clip-sharded run
     equal, bit for bit, what ONE GPU computes for the union of the clips.
Exit code 0 and a line "MULTIGPU OK" on rank 0 mean success.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ufvideo_b200 import build_region_encoder, packer, sharding, synth  # noqa: E402

K, HID = 8, 3584


def encoder(dev):
    enc = build_region_encoder(types.SimpleNamespace(mm_hidden_size=1152, hidden_size=HID), "square")
    enc.region_token_num = K
    enc.requires_grad_(False)
    with torch.no_grad():
        for p, w in zip((enc.feat_linear[0].weight, enc.feat_linear[0].bias,
                         enc.feat_linear[2].weight, enc.feat_linear[2].bias), synth.make_weights(0)):
            p.copy_(torch.from_numpy(w))
    return enc.to(dev).bfloat16()


def shard(step, rank, world, n_clips_total, frames, objects, ragged):
    """This rank's contiguous block of the step's clips (sharding.clip_block), as device tensors."""
    block = sharding.clip_block(n_clips_total, rank, world)
    if len(block) == 0:                           # more ranks than clips: an empty shard still takes part
        return np.zeros((1, 729, 1152), np.float32), [], []
    feats, masks, ann = synth.make_batch(len(block), frames, objects, "blob", 96, 128, first_clip=1000 * step + block.start,
                                         ragged=ragged)
    return feats, masks, ann


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    enc = encoder(dev)
    cases = [(5, 6, 3, False), (3, 12, 2, True), (4, 9, 4, True), (5, 6, 3, False)]     # clips, frames, objects, ragged
    pad_objs = max(-(-c // world) * o for c, _, o, _ in cases)
    pad_rows = pad_objs * K
    pg = sharding.PeerGather(pad_rows, pad_objs, HID, torch.bfloat16, dev)
    with torch.inference_mode():
        for step, (n_clips, frames, objects, ragged) in enumerate(cases * 3):          # 12 steps: the ring wraps
            feats, masks, ann = shard(step, rank, world, n_clips, frames, objects, ragged)
            ft = torch.from_numpy(feats).to(dev).bfloat16()
            md = [torch.from_numpy(m).to(dev) for m in masks]
            slots = packer.build_plan(md, ann, ft.shape[0], K, dev).slots
            # (1) fused gather (even steps) / side-stream peer push of locally written rows (odd steps)
            peer, tok_view, cnt_view, s = pg.begin(slots)
            if step % 2 == 0:
                enc.forward_padded(ft, md, ann, out=tok_view, counts_out=cnt_view, peer=peer)
            else:
                enc.forward_padded(ft, md, ann, out=tok_view, counts_out=cnt_view,
                                   after_enqueue=lambda: pg.push(s, peer, int(slots.sum())))
            pg.wait(s)
            fused = pg.gathered(s).clone()
            # (2) NCCL all-gather of the same payload
            payload, tok_view2, cnt_view2 = sharding.new_payload(slots, pad_rows, pad_objs, HID, torch.bfloat16, dev)
            enc.forward_padded(ft, md, ann, out=tok_view2, counts_out=cnt_view2)
            ref, _ = sharding.all_gather_payload(payload)
            torch.cuda.synchronize()
            got_t, got_c = sharding.unpack_padded(fused, pad_rows, pad_objs)
            ref_t, ref_c = sharding.unpack_padded(ref, pad_rows, pad_objs)
            assert got_c == ref_c, (step, rank, got_c, ref_c)
            assert torch.equal(got_t, ref_t), (step, rank, "fused gather != NCCL all-gather")
            # (3) one GPU, all clips of the step at once
            feats_all, masks_all, ann_all = synth.make_batch(n_clips, frames, objects, "blob", 96, 128,
                                                             first_clip=1000 * step, ragged=ragged)
            one_t, one_c = enc(torch.from_numpy(feats_all).to(dev).bfloat16(),
                               [torch.from_numpy(m).to(dev) for m in masks_all], None, ann_all, None)
            assert got_c == one_c, (step, rank, "counts differ from the single-GPU run")
            assert torch.equal(got_t, one_t), (step, rank, "tokens differ from the single-GPU run")
        pg.check()
    dist.barrier()
    if rank == 0:
        print(f"MULTIGPU OK: world {world}, multimem {pg.multimem}, {len(cases) * 3} steps", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
