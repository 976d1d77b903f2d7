"""Case tables + seeded input builders for the golden vectors.  TEST INFRASTRUCTURE ONLY.

Shared by ``oracle/gen_golden.py`` (runs the reference on these inputs, build container only)
and by the tests (re-create the same inputs anywhere and compare against ``tests/golden/``).
Inputs are never stored: they are regenerated from numpy PCG64 seeds, and each golden file
carries a checksum of the regenerated inputs so that generator drift is detected.
"""
from __future__ import annotations

import hashlib

import numpy as np

from ufvideo_b200 import synth

# ---- (1) resize / binarise: SURVEY.md appendix B.1 -------------------------------------------
RESIZE_SIZES = [(384, 384), (336, 336), (378, 378), (720, 1280), (480, 854), (1080, 1920),
                (100, 37), (54, 54), (28, 28), (13, 13), (27, 27), (81, 81), (27, 100),
                (100, 27), (26, 29), (1, 1), (2, 500),
                # sizes where ATen's fused source-index evaluation decides a tap (round 2), and the largest
                # video frames the eval drivers see
                (3, 3), (5, 5), (9, 9), (3, 5), (1, 3), (2049, 2049), (9, 2049), (2160, 3840)]
RESIZE_DENSITIES = [0.5, 0.05, 0.002]


def resize_masks(h: int, w: int) -> np.ndarray:
    """uint8 [n, h, w]: three random densities, all-zero, all-one, four single-corner pixels,
    one centre pixel."""
    out = []
    for d_i, dens in enumerate(RESIZE_DENSITIES):
        g = synth.rng_for(h * 100003 + w * 101 + d_i)
        out.append((g.random((h, w), dtype=np.float32) < dens).astype(np.uint8))
    out.append(np.zeros((h, w), np.uint8))
    out.append(np.ones((h, w), np.uint8))
    for y, x in ((0, 0), (0, w - 1), (h - 1, 0), (h - 1, w - 1), (h // 2, w // 2)):
        m = np.zeros((h, w), np.uint8)
        m[y, x] = 1
        out.append(m)
    return np.stack(out)


# ---- (2) pool ---------------------------------------------------------------------------------
POOL_CASES = [
    # name, frames, objects, family, (h, w)
    ("dense384", 3, 2, "dense", (384, 384)),
    ("blob384", 4, 3, "blob", (384, 384)),
    ("sparse720", 3, 2, "sparse", (720, 1280)),
    ("sparse480", 7, 1, "sparse", (480, 854)),
]


def pool_inputs(name):
    for i, (n, f, o, fam, (h, w)) in enumerate(POOL_CASES):
        if n == name:
            feats, masks, ann = synth.make_clip(9000 + i, f, o, fam, h, w)
            rows = np.array([r for obj in ann for r in obj])
            return feats, masks, rows
    raise KeyError(name)


# ---- (3) temporal token merge -------------------------------------------------------------------
TTM_T = [2, 3, 5, 9, 10, 16, 17, 32, 64, 255, 256, 512]
TTM_K = [1, 4, 8]
TTM_FAMILIES = ["random", "zeros", "somezero", "onehot"]


def ttm_tokens(family: str, t: int, seed: int, c: int = synth.C_SIGLIP) -> np.ndarray:
    g = synth.rng_for(seed)
    if family == "random":
        return g.standard_normal((t, c), dtype=np.float32)
    if family == "zeros":
        return np.zeros((t, c), np.float32)
    if family == "somezero":
        x = g.standard_normal((t, c), dtype=np.float32)
        x[g.random(t) < 0.4] = 0
        return x
    if family == "onehot":       # sims are exactly 0 or exactly 1 -> structural ties
        x = np.zeros((t, c), np.float32)
        x[np.arange(t), g.integers(0, 3, t)] = g.integers(1, 4, t).astype(np.float32)
        return x
    raise ValueError(family)


def ttm_cases():
    """(family, T, K, seed) for every combination with T > K (merge happens) plus a few T <= K."""
    cases, seed = [], 40000
    for fam in TTM_FAMILIES:
        for t in TTM_T:
            for k in TTM_K:
                if t > k and (fam == "random" or t <= 64):
                    cases.append((fam, t, k, seed))
                    seed += 1
    return cases


# ---- (4) end-to-end module calls ------------------------------------------------------------------
def e2e_case(name: str):
    """Returns dict(feats fp32, masks list of uint8 [q,H,W], ann_indices, k, dtype, aspect,
    masks_as_tensor)."""
    base = dict(k=8, dtype="f32", aspect="square", masks_as_tensor=False)
    if name == "c1":            # BASELINE configs[0]: 1 clip x 16 frames x 1 object, K=8, fp32
        feats, masks, ann = synth.make_batch(1, 16, 1, "dense", first_clip=0)
        return dict(base, feats=feats, masks=masks, ann=ann, masks_as_tensor=True)
    if name == "multi":         # ragged objects, per-clip mask sizes, K=4 (reference default)
        f0, m0, a0 = synth.make_clip(100, 6, 3, "blob", 384, 384, row0=0, ragged=True)
        f1, m1, a1 = synth.make_clip(101, 9, 2, "sparse", 480, 854, row0=6, ragged=True)
        f2, m2, a2 = synth.make_clip(102, 5, 4, "dense", 100, 37, row0=15)
        return dict(base, feats=np.concatenate([f0, f1, f2]), masks=[m0, m1, m2],
                    ann=[a0, a1, a2], k=4)
    if name == "bf16":          # UFVideo-7B training dtype
        feats, masks, ann = synth.make_batch(2, 12, 3, "blob", first_clip=200)
        return dict(base, feats=feats, masks=masks, ann=ann, dtype="bf16")
    if name == "f16":           # UFVideo-7B inference dtype (model/__init__.py:62)
        feats, masks, ann = synth.make_batch(2, 10, 2, "dense", first_clip=300)
        return dict(base, feats=feats, masks=masks, ann=ann, dtype="f16")
    if name == "quirk":         # PixRQA driver: q masks, one feature row
        feats, masks, _ = synth.make_clip(400, 2, 3, "blob", 384, 384)
        return dict(base, feats=feats, masks=[masks[:3]], ann=[[[0]]], masks_as_tensor=True)
    if name == "pad":           # image_aspect_ratio == 'pad' (layer.py:77-86)
        feats, masks, ann = synth.make_batch(1, 10, 2, "blob", h=480, w=854, first_clip=500)
        return dict(base, feats=feats, masks=masks, ann=ann, aspect="pad")
    if name == "empty":         # q == 0 -> one zero 336x336 mask (layer.py:73-75)
        feats = synth.features(600, 2)
        return dict(base, feats=feats, masks=[np.zeros((0, 336, 336), np.uint8)], ann=[[[1]]])
    if name == "shared":        # two objects sharing frames, one object listing a frame twice
        feats, masks, _ = synth.make_clip(700, 6, 3, "dense", 54, 54)
        ann = [[0, 1, 2, 3, 4, 5], [2, 2, 3, 5, 0, 1], [4, 4, 4, 1, 0, 3]]
        return dict(base, feats=feats, masks=[masks], ann=[ann], k=4)
    raise KeyError(name)


E2E_NAMES = ["c1", "multi", "bf16", "f16", "quirk", "pad", "empty", "shared"]


def digest(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()
