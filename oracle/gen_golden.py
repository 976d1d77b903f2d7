"""Generate tests/golden/*.npz by running the REAL reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden
The reference's ``ufvideo/model/layer.py`` is imported standalone (oracle/ref_loader.py) and
executed under the installed torch on seeded inputs (oracle/golden_cases.py).  What is stored
is the reference's OUTPUT only (inputs are regenerated from seeds by the tests).  The script
also cross-checks the numpy restatement against the reference while it runs and refuses to
write a golden file whose bit-exact quantities (patch masks, merge cuts, counts) disagree.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

from oracle import golden_cases as gc
from oracle import ref_loader, restatement as R
from ufvideo_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
TORCH_DTYPE = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}


def save(name, **arrays):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.0f} KiB)")


def gen_resize(ref):
    pool = ref.MaskPooling()
    bits, meta = [], []
    for h, w in gc.RESIZE_SIZES:
        masks = gc.resize_masks(h, w)
        m = torch.from_numpy(masks).float().unsqueeze(0)               # [1, q, H, W]
        if (h, w) != (27, 27):                                         # layer.py:137-139
            m = F.interpolate(m, size=(27, 27), mode="bilinear", align_corners=False)
        on = (m > 0)[0].reshape(masks.shape[0], -1).numpy()
        mine = np.stack([R.mask_to_patches(x) for x in masks])
        assert (on == mine).all(), f"restatement disagrees with F.interpolate at {(h, w)}"
        # cnt through the reference's own pooling on a constant feature map: mean of ones
        x = torch.ones(masks.shape[0], 1, 27, 27)
        pooled = pool(x, torch.from_numpy(masks).float().unsqueeze(0))
        assert torch.equal(pooled[:, 0] > 0.5, torch.from_numpy(on.any(1)))
        bits.append(R.pack_bits(on))
        meta.append({"h": h, "w": w, "n": int(masks.shape[0]), "sha": gc.digest(masks)})
    save("resize.npz", bits=np.concatenate(bits), meta=np.array(json.dumps(meta)))


def gen_pool(ref):
    pool = ref.MaskPooling()
    out = {}
    for name, *_ in gc.POOL_CASES:
        feats, masks, rows = gc.pool_inputs(name)
        x = torch.from_numpy(feats)[torch.from_numpy(rows)]
        x = x.reshape(x.shape[0], 27, 27, -1).permute(0, 3, 1, 2)     # layer.py:100-101
        ref_pooled = pool(x, torch.from_numpy(masks).float().unsqueeze(0)).numpy()
        on = np.stack([R.mask_to_patches(m) for m in masks])
        mine = R.mask_pool(feats, rows, on)
        err = np.abs(mine - ref_pooled).max()
        assert err <= 1e-5, (name, err)
        print(f"pool {name}: restatement vs reference max-abs {err:.2e}")
        out[name] = ref_pooled
        out[name + "_sha"] = np.array(gc.digest(feats, masks, rows))
    save("pool.npz", **out)


def reference_cut(x: torch.Tensor, r: int):
    """The reference's own arithmetic for the cut mask (layer.py:11-18,24), evaluated with the
    same ATen calls, because token_merge does not return its decisions."""
    n1 = F.normalize(x[:, :-1, :], p=2, dim=-1)
    n2 = F.normalize(x[:, 1:, :], p=2, dim=-1)
    sim = torch.sum(n1 * n2, dim=-1)
    kth = torch.topk(sim.flatten(), r).values[-1]
    return sim[0].numpy(), (sim[0] < kth).numpy()


def gen_ttm(ref):
    out, meta = {}, []
    for idx, (fam, t, k, seed) in enumerate(gc.ttm_cases()):
        x = gc.ttm_tokens(fam, t, seed)
        xt = torch.from_numpy(x)[None]
        merged = ref.token_merge(xt, t - k)[0].numpy()
        sims_ref, cut_ref = reference_cut(xt, t - k)
        # sanity: the stored cut reproduces the reference's own output row count
        assert merged.shape[0] == 1 + int(cut_ref.sum())
        tok, cut, sims = R.token_merge(x, k)
        gap = np.abs(np.sort(sims_ref)[::-1][t - k - 1] - sims_ref)
        assert (cut == cut_ref).all(), (fam, t, k, "restatement cut differs from the reference")
        err = np.abs(tok - merged).max()
        assert err <= 1e-5, (fam, t, k, err)
        key = f"case{idx}"
        out[key + "_cut"] = np.packbits(cut_ref)
        out[key + "_merged"] = merged if t <= 64 else merged[:, :64].copy()
        meta.append({"family": fam, "t": t, "k": k, "seed": seed, "sha": gc.digest(x),
                     "rows": int(merged.shape[0]), "full": bool(t <= 64),
                     "min_gap_nonzero": float(gap[gap > 0].min()) if (gap > 0).any() else 0.0})
    save("ttm.npz", meta=np.array(json.dumps(meta)), **out)


def run_reference_module(ref, case, weights):
    dt = TORCH_DTYPE[case["dtype"]]
    enc = ref.build_region_encoder(ref_loader.reference_config(), case["aspect"])
    enc.region_token_num = case["k"]
    with torch.no_grad():
        for p, w in zip((enc.feat_linear[0].weight, enc.feat_linear[0].bias,
                         enc.feat_linear[2].weight, enc.feat_linear[2].bias), weights):
            p.copy_(torch.from_numpy(w))
    enc = enc.to(dt).eval()
    captured = {}
    enc.mask_pooling.register_forward_hook(lambda m, i, o: captured.setdefault("pooled", []).append(o))
    feats = torch.from_numpy(case["feats"]).to(dt)
    masks = [torch.from_numpy(m).float() for m in case["masks"]]
    if case["masks_as_tensor"]:
        masks = torch.stack(masks)                                     # [B, q, H, W]
    with torch.no_grad():
        tokens, counts = enc(feats, masks, feats, case["ann"], None)
    pooled = torch.cat(captured["pooled"]).float().numpy()
    return tokens.float().numpy(), list(counts), pooled


def gen_e2e(ref):
    weights = synth.make_weights(0)
    for name in gc.E2E_NAMES:
        case = gc.e2e_case(name)
        dt = case["dtype"]
        w_rounded = tuple(R.round_to(w, dt) for w in weights)
        tokens, counts, pooled = run_reference_module(ref, case, w_rounded)
        mine = R.encode(R.round_to(case["feats"], dt), case["masks"], case["ann"], case["k"], dt,
                        w_rounded, pad_square=case["aspect"] == "pad")
        assert mine["counts"] == counts, (name, mine["counts"], counts)
        e_pool = np.abs(mine["pooled"] - pooled).max()
        e_tok = np.abs(mine["tokens"] - tokens).max()
        tol = {"f32": 1e-5, "bf16": 1e-2, "f16": 1e-2}[dt]
        print(f"e2e {name}: counts {counts}  pooled max-abs {e_pool:.2e}  tokens max-abs {e_tok:.2e}"
              f"  (|tokens| max {np.abs(tokens).max():.3f})")
        assert e_pool <= 1e-5 and e_tok <= tol, name
        sha = gc.digest(case["feats"], *case["masks"])
        save(f"e2e_{name}.npz", tokens=tokens, counts=np.array(counts), pooled=pooled,
             sha=np.array(sha))


def main():
    ref = ref_loader.load_reference_layer()
    if ref is None:
        sys.exit("reference tree not found: golden vectors can only be generated where "
                 "/root/reference (or $UFV_REF) exists")
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    print("torch", torch.__version__, "threads", torch.get_num_threads())
    gen_resize(ref)
    gen_pool(ref)
    gen_ttm(ref)
    gen_e2e(ref)


if __name__ == "__main__":
    main()
