"""Developer tool: per-stage CUDA-event timings of the hot path on a BASELINE config."""
import argparse
import sys
import os
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ufvideo_b200 import build_region_encoder, layer, packer, synth  # noqa: E402


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in evs)
    return t[len(t) // 2] * 1e3, t[0] * 1e3   # median, min in us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=8)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--objects", type=int, default=4)
    ap.add_argument("--family", default="dense")
    ap.add_argument("--k", type=int, default=8)
    ap.add_argument("--dtype", default="bf16")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    dt = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[a.dtype]
    feats, masks, ann = synth.make_batch(a.clips, a.frames, a.objects, a.family)
    ft = torch.from_numpy(feats).to(dev).to(dt)
    md = [torch.from_numpy(m).to(dev) for m in masks]
    enc = build_region_encoder(types.SimpleNamespace(mm_hidden_size=1152, hidden_size=3584), "square")
    enc.region_token_num = a.k
    enc.requires_grad_(False)
    enc = enc.to(dev).to(dt)
    plan = packer.build_plan(md, ann, ft.shape[0], a.k, dev)
    q = plan.n_masks
    patches = layer.mask_to_patches(plan, dev)
    bits = patches["bits"]
    pooled = layer.mask_pool(ft, plan, patches)
    merged, counts, _ = layer.ttm(pooled, plan, a.k, dt)
    l0, l2 = enc.feat_linear[0], enc.feat_linear[2]
    hid = layer.linear(merged, l0.weight, l0.bias, True)
    print(f"config clips={a.clips} frames={a.frames} objects={a.objects} family={a.family} q={q} "
          f"groups={plan.n_groups} m_pad={plan.m_pad} dtype={a.dtype}")
    on_union = 0
    b = bits.cpu().numpy().view(np.uint32)
    go, gm = plan.host["grp_off"], plan.host["grp_member"]
    for g in range(plan.n_groups):
        u = np.bitwise_or.reduce(b[gm[go[g]:go[g + 1]]], axis=0)
        on_union += int(sum(bin(int(x)).count("1") for x in u))
    feat_bytes = on_union * 1152 * ft.element_size()
    res = {}
    res["k1 patches"] = timed(lambda: layer.mask_to_patches(plan, dev))
    for mdt, mname in ((torch.float32, "f32"), (torch.uint8, "u8")):
        mh = [torch.from_numpy(m).to(mdt).pin_memory() for m in masks]
        mdv = [m.to(dev) for m in mh]
        for mode in ("rows", "taps"):
            packer.READ_MODE = mode
            for where, mm in (("dev", mdv), ("host", mh)):
                pl = packer.build_plan(mm, ann, ft.shape[0], a.k, dev, use_cache=False)
                res[f"k1 {mname} {where} {mode}"] = timed(lambda: layer.mask_to_patches(pl, dev), iters=10)
        packer.READ_MODE = "auto"
    res["k2 pool"] = timed(lambda: layer.mask_pool(ft, plan, patches))
    res["k3 ttm"] = timed(lambda: layer.ttm(pooled, plan, a.k, dt))
    res["k4a linear1+gelu"] = timed(lambda: layer.linear(merged, l0.weight, l0.bias, True))
    res["k4b linear2"] = timed(lambda: layer.linear(hid, l2.weight, l2.bias, False))
    # "next" rows: device-side <region> splice and the pool adjoint (training)
    n_text = 2048
    text = torch.randn((n_text, 3584), device=dev).to(dt)
    pos = torch.arange(plan.n_obj, dtype=torch.int32, device=dev) * (n_text // max(plan.n_obj, 1))
    tok_pad, cnt_dev, _ = enc.encode_padded(ft, md, ann)
    res["splice (f1)"] = timed(lambda: layer.splice_regions(text, pos, tok_pad, cnt_dev, plan))
    w_adj = torch.randn((q, 1152), device=dev)
    go, gm, gr = plan.host["grp_off"], plan.host["grp_member"], plan.host["grp_row"]
    frame_of = np.empty(q, dtype=np.int64)
    for g in range(plan.n_groups):
        frame_of[gm[go[g]:go[g + 1]]] = int(gr[g])
    order = np.argsort(frame_of, kind="stable").astype(np.int32)
    per_row = np.bincount(frame_of, minlength=ft.shape[0])
    row_off = np.concatenate([[0], np.cumsum(per_row)]).astype(np.int32)
    meta = torch.from_numpy(np.concatenate([row_off, order])).to(dev)
    d_feats = torch.empty_like(ft)
    from ufvideo_b200 import _cabi

    def pool_bwd():
        _cabi.check(_cabi.lib().ufv_mask_pool_backward(
            w_adj.data_ptr(), bits.data_ptr(), meta.data_ptr(), meta.data_ptr() + 4 * row_off.size, ft.shape[0],
            int(per_row.max()), 729, 1152, d_feats.data_ptr(), packer.FEAT_DTYPES[dt],
            torch.cuda.current_stream(dev).cuda_stream))
    res["pool adjoint (f3)"] = timed(pool_bwd)
    res["torch linear1"] = timed(lambda: torch.nn.functional.linear(merged, l0.weight, l0.bias))
    res["torch linear2"] = timed(lambda: torch.nn.functional.linear(hid, l2.weight, l2.bias))
    res["plan (host)"] = timed(lambda: packer.build_plan(md, ann, ft.shape[0], a.k, dev))
    res["encode_padded"] = timed(lambda: enc.encode_padded(ft, md, ann))
    res["forward"] = timed(lambda: enc(ft, md, None, ann, None))
    for k, (med, mn) in res.items():
        print(f"{k:20s} median {med:9.1f} us   min {mn:9.1f} us")
    t = res["k2 pool"][0] * 1e-6
    print(f"pool: union feature bytes {feat_bytes/1e6:.1f} MB -> {feat_bytes/t/1e9:.0f} GB/s")
    m = plan.m_pad
    for name, k_, n_ in (("k4a linear1+gelu", 1152, 3584), ("k4b linear2", 3584, 3584)):
        fl = 2.0 * m * k_ * n_
        print(f"{name}: {fl/res[name][0]/1e6:.1f} TFLOP/s   weight-bytes bound {n_*k_*2/res[name][0]/1e3:.0f} GB/s")
    print(f"forward: {q/res['forward'][0]*1e6:.0f} object-frames/s")
    t = res["pool adjoint (f3)"][0] * 1e-6
    print(f"pool adjoint: writes {d_feats.numel() * d_feats.element_size() / 1e6:.0f} MB -> "
          f"{d_feats.numel() * d_feats.element_size() / t / 1e9:.0f} GB/s")
    t = res["splice (f1)"][0] * 1e-6
    moved = 2 * (n_text - plan.n_obj + plan.m_pad) * 3584 * text.element_size()
    print(f"splice: {moved / 1e6:.1f} MB read+written -> {moved / t / 1e9:.0f} GB/s")


if __name__ == "__main__":
    main()
