"""Developer tool: the BASELINE.json configs beyond the bench workload, per GPU, device-resident.

  python tools/sweep.py [c2 c3 c4 c5 ...]

For every configuration: whole-path step time through MaskExtractor.forward (CUDA events, median of
N), object-frames/s, and the per-kernel times (each C-ABI stage called alone) with the pool kernel's
algorithmic GB/s and the projector's TFLOP/s.  Masks are uint8 for the large shapes (the reference
`.float()`s them anyway; kernel 1 reads any of its dtypes in place).  c3 and c4 are the per-GPU
shares of the 8-GPU configs (clips are sharded)."""
import json
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ufvideo_b200 import build_region_encoder, layer, packer, synth  # noqa: E402

CONFIGS = {
    # name: (clips, frames, objects, family, ragged, note)
    "c1": (1, 16, 1, "dense", False, "configs[0]: 1 clip x 16 frames x 1 object, fp32 (the reference's CPU-runnable case)"),
    "c2": (8, 16, 4, "dense", False, "configs[1]: 8 clips x 16 frames x 4 objects"),
    "c2-blob": (8, 16, 4, "blob", False, "configs[1] shape, blob masks"),
    "c3": (8, 32, 8, "sparse", False, "configs[2] per-GPU share: 8 of 64 clips x 32 frames x 8 objects, sparse masks"),
    "c3-dense": (8, 32, 8, "dense", False, "configs[2] shape, dense masks"),
    "c4": (2, 256, 16, "blob", False, "configs[3]: 2 clips x 256 frames x 16 objects (merge-heavy, r = 248)"),
    "c5-small": (16, 8, 1, "blob", False, "configs[4] corner: 16 clips x 8 frames x 1 object"),
    "c5-wide": (1, 64, 64, "blob", True, "configs[4] corner: 1 clip x 64 frames x 64 objects, ragged T_o"),
    "c5-mid": (2, 64, 32, "blob", True, "configs[4] interior: 2 clips x 64 frames x 32 objects, ragged T_o"),
    "c5-long": (1, 512, 4, "dense", False, "configs[4] corner: 1 clip x 512 frames x 4 objects"),
}


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
    return t[len(t) // 2] * 1e3   # median, us


def baselines(name, feats, masks, ann, k, dt, enc, dev):
    """The reference's own op sequence (oracle/reference_port.py, validated bit-identical to the real
    module): on the host CPU (all cores, fp32, a bounded sample of the clips) and as CUDA eager on this GPU."""
    import os
    import time
    from oracle import reference_port
    out = {}
    weights = [p.detach() for p in (enc.feat_linear[0].weight, enc.feat_linear[0].bias,
                                    enc.feat_linear[2].weight, enc.feat_linear[2].bias)]
    # CUDA eager, model dtype, full config (skipped when the three dense fp32 temporaries would not fit)
    q = sum(m.shape[0] for m in masks)
    q_max = max(m.shape[0] for m in masks)
    if q_max * 729 * 1152 * 4 * 4 < 120e9:
        ft = torch.from_numpy(feats).to(dev).to(dt)
        mt = [torch.from_numpy(m).to(dev).float() for m in masks]
        with torch.no_grad():
            for _ in range(2):
                reference_port.encode(ft, mt, ann, k, *weights)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 3
            for _ in range(n):
                reference_port.encode(ft, mt, ann, k, *weights)
            torch.cuda.synchronize()
            out["eager_cuda_obj_frames_per_s"] = q * n / (time.perf_counter() - t0)
        del ft, mt
        torch.cuda.empty_cache()
    # host CPU, fp32, first clip(s) only
    n_clips = 1 if q / len(masks) >= 256 else min(2, len(masks))
    frames_per_clip = feats.shape[0] // len(masks)
    cf = torch.from_numpy(feats[: n_clips * frames_per_clip])
    cm = [torch.from_numpy(m).float() for m in masks[:n_clips]]
    ca = ann[:n_clips]
    cw = [w.float().cpu() for w in weights]
    torch.set_num_threads(os.cpu_count() or 1)
    qs = sum(m.shape[0] for m in cm)
    if qs * 729 * 1152 * 4 * 4 < 60e9:
        with torch.no_grad():
            reference_port.encode(cf, cm, ca, k, *cw)
            t0 = time.perf_counter()
            n = 0
            while n < 2 or time.perf_counter() - t0 < 3.0:
                reference_port.encode(cf, cm, ca, k, *cw)
                n += 1
            out["cpu_obj_frames_per_s"] = qs * n / (time.perf_counter() - t0)
            out["cpu_cores"] = os.cpu_count()
            out["cpu_sample"] = f"{n_clips} clip(s), {qs} object-frames, fp32"
    return out


def run(name, k=8, with_baselines=False):
    clips, frames, objects, family, ragged, note = CONFIGS[name]
    dev = torch.device("cuda:0")
    dt = torch.float32 if name == "c1" else torch.bfloat16
    feats, masks, ann = synth.make_batch(clips, frames, objects, family, ragged=ragged)
    ft = torch.from_numpy(feats).to(dev).to(dt)
    md = [torch.from_numpy(m).to(dev) for m in masks]          # uint8
    enc = build_region_encoder(types.SimpleNamespace(mm_hidden_size=1152, hidden_size=3584), "square")
    enc.region_token_num = k
    enc.requires_grad_(False)                 # forward-only path
    enc = enc.to(dev).to(dt)
    base = baselines(name, feats, masks, ann, k, dt, enc, dev) if with_baselines else {}
    del feats
    plan = packer.build_plan(md, ann, ft.shape[0], k, dev)
    q = plan.n_masks
    patches = layer.mask_to_patches(plan, dev)
    pooled = layer.mask_pool(ft, plan, patches)
    merged, counts, _ = layer.ttm(pooled, plan, k, dt)
    l0, l2 = enc.feat_linear[0], enc.feat_linear[2]
    hid = layer.linear(merged, l0.weight, l0.bias, True)
    # SURVEY 8(d) bytes: union over ALL objects of a frame, however the packer grouped them
    pool_bytes = packer.algorithmic_pool_bytes(plan, patches["bits"].cpu().numpy(), 1152, ft.element_size())
    union = (pool_bytes - q * 1152 * 4 - q * 96) // (1152 * ft.element_size())
    res = {
        "config": name, "note": note, "object_frames": q, "frames": int(ft.shape[0]), "objects": plan.n_obj,
        "tokens": plan.m_pad, "groups": plan.n_groups, "max_len": plan.max_len,
        "union_patch_frac": union / (len(set(plan.host["grp_row"].tolist())) * 729.0),
        "k1_us": timed(lambda: layer.mask_to_patches(plan, dev)),
        "k2_us": timed(lambda: layer.mask_pool(ft, plan, patches)),
        "k3_us": timed(lambda: layer.ttm(pooled, plan, k, dt)),
        "k4a_us": timed(lambda: layer.linear(merged, l0.weight, l0.bias, True)),
        "k4b_us": timed(lambda: layer.linear(hid, l2.weight, l2.bias, False)),
        "forward_us": timed(lambda: enc(ft, md, None, ann, None)),
    }
    res["pool_gbs"] = pool_bytes / res["k2_us"] / 1e3
    res["pool_bytes"] = pool_bytes
    peak_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    peak = json.load(open(peak_path))["hbm_gbs"] if os.path.isfile(peak_path) else 6650.0
    res["pool_frac_of_measured_hbm"] = res["pool_gbs"] / peak
    res["proj_tflops"] = 33947648.0 * plan.m_pad / (res["k4a_us"] + res["k4b_us"]) / 1e6
    res["object_frames_per_s"] = q / res["forward_us"] * 1e6
    res.update(base)
    return res


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    with_baselines = "--baselines" in sys.argv
    names = args or list(CONFIGS)
    out = []
    for n in names:
        r = run(n, with_baselines=with_baselines)
        out.append(r)
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()
    print()
    print(f"{'config':10s} {'obj-fr':>7s} {'tokens':>6s} {'k1':>7s} {'k2':>8s} {'k3':>7s} {'k4a':>6s} {'k4b':>6s} "
          f"{'fwd us':>8s} {'obj-fr/s':>10s} {'pool GB/s':>9s} {'of HBM':>6s} {'proj TF':>7s}")
    for r in out:
        print(f"{r['config']:10s} {r['object_frames']:7d} {r['tokens']:6d} {r['k1_us']:7.1f} {r['k2_us']:8.1f} "
              f"{r['k3_us']:7.1f} {r['k4a_us']:6.1f} {r['k4b_us']:6.1f} {r['forward_us']:8.1f} "
              f"{r['object_frames_per_s']:10.0f} {r['pool_gbs']:9.0f} {r['pool_frac_of_measured_hbm']:6.2f} {r['proj_tflops']:7.1f}"
              + (f"   eager-CUDA {r['eager_cuda_obj_frames_per_s']:9.0f}/s" if "eager_cuda_obj_frames_per_s" in r else "")
              + (f"   CPU({r.get('cpu_cores')}c) {r['cpu_obj_frames_per_s']:7.0f}/s" if "cpu_obj_frames_per_s" in r else ""))


if __name__ == "__main__":
    main()
