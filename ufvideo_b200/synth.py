"""Seeded synthetic inputs of the shapes BASELINE.json names.

numpy's PCG64 streams are platform-independent, so the same seed gives the same clip in the
build container (where the golden vectors are generated from the reference) and on the GPU box.
Every clip is seeded with ``1234 + clip_id`` so that sharding clips over ranks never changes
the data of a clip.

Layout produced (the reference's batched input contract, train.py:628-634,689-692 and
eval/inference_videorefer_q_bench.py:96-134):
  feats        [F_total, n_patch, C]   vision-tower patch features of the annotated frames
  masks[i]     [q_i, H, W]             object-major, each object's frames in temporal order
  ann_indices[i][o] = global feature rows of object o of clip i (already offset by clip)
"""
from __future__ import annotations

import numpy as np

N_PATCH = 729
C_SIGLIP = 1152
HID_QWEN2_7B = 3584
CLIP_SEED0 = 1234


def rng_for(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(int(seed)))


def features(seed: int, n_frames: int, n_patch: int = N_PATCH, c: int = C_SIGLIP) -> np.ndarray:
    """Standard-normal fp32 features [n_frames, n_patch, c]."""
    return rng_for(seed).standard_normal((n_frames, n_patch, c), dtype=np.float32)


def masks_dense(seed: int, q: int, h: int, w: int, p: float = 0.3) -> np.ndarray:
    """Bernoulli(p) per pixel -> about 76 % of the 27x27 patches on at p = 0.3, 384x384."""
    return (rng_for(seed).random((q, h, w), dtype=np.float32) < p).astype(np.uint8)


def masks_blob(seed: int, n_obj: int, t: int, h: int, w: int) -> np.ndarray:
    """1-3 axis-aligned ellipses per object covering roughly 2-30 % of the image, drifting by
    at most 5 % of the image size per frame.  Returns uint8 [n_obj * t, h, w], object-major."""
    g = rng_for(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.zeros((n_obj * t, h, w), dtype=np.uint8)
    for o in range(n_obj):
        n_ell = int(g.integers(1, 4))
        cy = g.uniform(0.2, 0.8, n_ell) * h
        cx = g.uniform(0.2, 0.8, n_ell) * w
        ry = g.uniform(0.06, 0.28, n_ell) * h
        rx = g.uniform(0.06, 0.28, n_ell) * w
        for f in range(t):
            m = np.zeros((h, w), dtype=bool)
            for e in range(n_ell):
                m |= ((yy - cy[e]) / ry[e]) ** 2 + ((xx - cx[e]) / rx[e]) ** 2 <= 1.0
            out[o * t + f] = m
            cy = np.clip(cy + g.uniform(-0.05, 0.05, n_ell) * h, 0, h - 1)
            cx = np.clip(cx + g.uniform(-0.05, 0.05, n_ell) * w, 0, w - 1)
    return out


def masks_sparse(seed: int, q: int, h: int, w: int, p: float = 0.002, zero_every: int = 7) -> np.ndarray:
    """Salt noise (p per pixel) plus one thin random polyline per mask; every ``zero_every``-th
    mask is entirely zero (the reference's cnt == 0 case, layer.py:145)."""
    g = rng_for(seed)
    out = (g.random((q, h, w), dtype=np.float32) < p).astype(np.uint8)
    for j in range(q):
        pts = np.stack([g.integers(0, h, 4), g.integers(0, w, 4)], 1)
        for a, b in zip(pts[:-1], pts[1:]):
            n = int(max(abs(b[0] - a[0]), abs(b[1] - a[1]))) + 1
            ys = np.linspace(a[0], b[0], n).round().astype(int)
            xs = np.linspace(a[1], b[1], n).round().astype(int)
            out[j, ys, xs] = 1
        if zero_every and j % zero_every == zero_every - 1:
            out[j] = 0
    return out


def make_masks(family: str, seed: int, n_obj: int, t: int, h: int, w: int) -> np.ndarray:
    if family == "dense":
        return masks_dense(seed, n_obj * t, h, w)
    if family == "blob":
        return masks_blob(seed, n_obj, t, h, w)
    if family == "sparse":
        return masks_sparse(seed, n_obj * t, h, w)
    raise ValueError(f"unknown mask family {family!r}")


def make_clip(clip_id: int, n_frames: int, n_obj: int, family: str = "dense", h: int = 384,
              w: int = 384, row0: int = 0, c: int = C_SIGLIP, n_patch: int = N_PATCH,
              ragged: bool = False, feats: bool = True):
    """One clip: (feats [n_frames, n_patch, c] fp32, masks uint8 [q, h, w], ann_indices_of_clip).

    Every object is annotated on every frame unless ``ragged``, where object o keeps a random
    ascending subset of T_o in [1, n_frames] frames.  ``row0`` is the clip's first global
    feature row (the collator's cumulative offset, train.py:689-692).
    """
    seed = CLIP_SEED0 + clip_id
    feats = features(seed, n_frames, n_patch, c) if feats else None     # feats=False: masks and indices only
    g = rng_for(seed * 7919 + 1)
    if ragged:
        frames = [np.sort(g.choice(n_frames, int(g.integers(1, n_frames + 1)), replace=False))
                  for _ in range(n_obj)]
    else:
        frames = [np.arange(n_frames) for _ in range(n_obj)]
    full = make_masks(family, seed * 7919 + 2, n_obj, n_frames, h, w)
    keep = np.concatenate([o * n_frames + f for o, f in enumerate(frames)])
    masks = full[keep]
    ann = [[int(row0 + r) for r in f] for f in frames]
    return feats, masks, ann


def make_batch(n_clips: int, n_frames: int, n_obj: int, family: str = "dense", h: int = 384,
               w: int = 384, first_clip: int = 0, clip_stride: int = 1, c: int = C_SIGLIP,
               n_patch: int = N_PATCH, ragged: bool = False):
    """``n_clips`` clips with ids first_clip, first_clip+stride, ... concatenated the way the
    reference's collator does.  Returns (feats, masks list, ann_indices)."""
    feats, masks, ann = [], [], []
    for i in range(n_clips):
        f, m, a = make_clip(first_clip + i * clip_stride, n_frames, n_obj, family, h, w,
                            row0=i * n_frames, c=c, n_patch=n_patch, ragged=ragged)
        feats.append(f)
        masks.append(m)
        ann.append(a)
    return np.concatenate(feats), masks, ann


def make_weights(seed: int = 0, c: int = C_SIGLIP, hid: int = HID_QWEN2_7B):
    """nn.Linear-style init (uniform +-1/sqrt(fan_in)) for feat_linear.0 and feat_linear.2
    (reference layer.py:55-59).  Returns fp32 (w1 [hid,c], b1 [hid], w2 [hid,hid], b2 [hid])."""
    g = rng_for(seed)
    k1, k2 = 1.0 / np.sqrt(c), 1.0 / np.sqrt(hid)
    w1 = g.uniform(-k1, k1, (hid, c)).astype(np.float32)
    b1 = g.uniform(-k1, k1, hid).astype(np.float32)
    w2 = g.uniform(-k2, k2, (hid, hid)).astype(np.float32)
    b2 = g.uniform(-k2, k2, hid).astype(np.float32)
    return w1, b1, w2, b2
