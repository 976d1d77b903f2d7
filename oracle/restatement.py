"""CPU restatement of UFVideo's object-encoder hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *oracle*: a from-scratch numpy restatement of the algorithm in the
reference's ``ufvideo/model/layer.py``.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it.  The product package
``ufvideo_b200`` never imports anything under ``oracle/``.

Parity status: PINNED.  The reference holds no tests or golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference module itself,
executed in the build container: ``oracle/gen_golden.py`` imports the reference ``layer.py``
standalone, runs it on seeded inputs and commits the results under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function below against those files.

Where the arithmetic lives: the reference's numerics are PyTorch ATen calls
(``F.interpolate``, ``F.normalize``, ``torch.topk``, ``torch.mean``, ``nn.Linear``, ``nn.GELU``;
torch pinned 2.5.1 in the reference's requirements.txt:223, 2.11.0 installed here).  ATen's fp32
reduction order is not part of its contract, so this restatement fixes ONE documented fp32
evaluation order ("canonical order", below).  The CUDA kernels implement exactly the same
order, so CUDA == oracle bit-for-bit up to the projector input, and oracle vs reference is:
patch bitmasks equal, merge decisions equal, pooled/merged tokens within 1e-5 (fp32).

Canonical order
---------------
pool   : per (object-frame, channel): acc = 0; for p ascending over on-patches:
         acc = fl(acc + x[p]); pooled = fl(acc / fl(fl(cnt) + 1e-8)).
rowsum : the 32-lane strided sum used for norms and dots of a C-vector v:
         lane l owns elements 128*k + 4*l + j (k ascending, then j = 0..3), added one at a time
         into one accumulator per lane; lanes are combined by the xor butterfly
         acc[l] = fl(acc[l] + acc[l ^ off]) for off = 16, 8, 4, 2, 1.  Products are rounded to
         fp32 before they are added (no FMA).
merge  : per (group, channel): acc = 0; for t ascending in the group: acc = fl(acc + x[t]);
         token = fl(acc / fl(n)).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
PATCH = 27          # SigLIP-so400m/14 at 384 px: 27 x 27 patches (reference encoder.py:108)
NORM_EPS = F32(1e-12)   # F.normalize default eps (layer.py:13-14)
DENORM_EPS = F32(1e-8)  # layer.py:145


# ----------------------------------------------------------------------------------------------
# (1) mask resize + binarise  --  layer.py:137-143 (F.interpolate bilinear, align_corners=False, > 0)
# ----------------------------------------------------------------------------------------------
def axis_taps(n_in: int, n_out: int = PATCH):
    """Tap table of ATen's bilinear resize along one axis, in fp32 exactly as ATen computes it.

    Returns (i0, i1, use0, use1): int32 source indices of the two taps of every output index and
    whether their interpolation weight is non-zero.  Follows layer.py:139; ATen semantics are
    restated in SURVEY.md appendix A.1 (scale = in/out in fp32; src = max(scale*(i+.5)-.5, 0)).
    """
    i = np.arange(n_out, dtype=F32)
    scale = F32(n_in) / F32(n_out)
    # ATen's `scale * (i + 0.5) - 0.5` is ONE fused multiply-add on CPU and CUDA (a single rounding).  The
    # product has <= 24 + 6 significant bits, so the float64 expression below is exact and its rounding to
    # fp32 is the fma result.  (Rounding the product first differs at n_in = 3, 5, 9, 2049 for n_out = 27.)
    src = (np.float64(scale) * (i.astype(np.float64) + 0.5) - 0.5).astype(F32)
    src = np.maximum(src, F32(0.0)).astype(F32)
    i0 = np.minimum(np.floor(src).astype(np.int64), n_in - 1)
    lam1 = np.clip(src - i0.astype(F32), F32(0.0), F32(1.0)).astype(F32)
    lam0 = (F32(1.0) - lam1).astype(F32)
    i1 = i0 + (i0 < n_in - 1)
    return i0.astype(np.int32), i1.astype(np.int32), lam0 > 0, lam1 > 0


def pad_to_square_offsets(h: int, w: int):
    """'pad' aspect mode, layer.py:77-86: zero-pad the short side, centred.  Returns
    (side, top, left): the padded side length and where the original image starts."""
    side = max(h, w)
    return side, (side - h) // 2, (side - w) // 2


def mask_to_patches(mask: np.ndarray, n_out: int = PATCH, pad_square: bool = False) -> np.ndarray:
    """[H, W] non-negative mask -> bool [n_out*n_out] 'patch is on' (row-major h, w).

    Equals ``F.interpolate(mask, (n_out, n_out), 'bilinear', align_corners=False) > 0`` for
    non-negative masks: the interpolated value is a sum of non-negative weight*value products,
    so it is positive iff some tap with non-zero weight is positive (layer.py:139,143).
    With H == W == n_out the resize is skipped (layer.py:137) and the test is mask > 0.
    """
    mask = np.asarray(mask)
    h, w = mask.shape
    pos = mask > 0
    if pad_square:
        side, top, left = pad_to_square_offsets(h, w)
        full = np.zeros((side, side), dtype=bool)
        full[top:top + h, left:left + w] = pos
        pos, h, w = full, side, side
    if h == n_out and w == n_out:
        return pos.reshape(-1).copy()
    h0, h1, uh0, uh1 = axis_taps(h, n_out)
    w0, w1, uw0, uw1 = axis_taps(w, n_out)
    on = np.zeros((n_out, n_out), dtype=bool)
    for hi, uh in ((h0, uh0), (h1, uh1)):
        for wi, uw in ((w0, uw0), (w1, uw1)):
            on |= pos[np.ix_(hi, wi)] & uh[:, None] & uw[None, :]
    return on.reshape(-1)


def rle_to_mask(rle: dict) -> np.ndarray:
    """COCO run-length mask -> dense uint8 [H, W] (what pycocotools.mask.decode returns inside the
    reference's annToMask, ufvideo/mm_utils.py:22-33): pixels are numbered column-major, runs alternate
    off / on starting with off.  Uncompressed counts only (lists of ints)."""
    h, w = (int(v) for v in rle["size"])
    flat = np.zeros(h * w, dtype=np.uint8)
    pos, val = 0, 0
    for c in rle["counts"]:
        if val:
            flat[pos:pos + int(c)] = 1
        pos += int(c)
        val ^= 1
    return flat.reshape((h, w), order="F")


def pack_bits(on: np.ndarray) -> np.ndarray:
    """bool [..., n] -> uint32 [..., ceil(n/32)] little-endian bit order (bit p%32 of word p//32)."""
    n = on.shape[-1]
    words = (n + 31) // 32
    padded = np.zeros(on.shape[:-1] + (words * 32,), dtype=np.uint64)
    padded[..., :n] = on
    weights = (np.uint64(1) << np.arange(32, dtype=np.uint64))
    return (padded.reshape(on.shape[:-1] + (words, 32)) * weights).sum(-1).astype(np.uint32)


# ----------------------------------------------------------------------------------------------
# (2) mask pool  --  layer.py:98-104 (gather + upcast) and :144-147 (masked mean)
# ----------------------------------------------------------------------------------------------
def mask_pool(feats: np.ndarray, frame_rows, on: np.ndarray) -> np.ndarray:
    """feats fp32 [F, n_patch, C] (already upcast, layer.py:104); frame_rows int [q];
    on bool [q, n_patch]  ->  pooled fp32 [q, C] in canonical order."""
    feats = np.asarray(feats, dtype=F32)
    frame_rows = np.asarray(frame_rows, dtype=np.int64)
    q, n_patch = on.shape
    acc = np.zeros((q, feats.shape[2]), dtype=F32)
    for p in range(n_patch):
        sel = np.nonzero(on[:, p])[0]
        if sel.size:
            acc[sel] = acc[sel] + feats[frame_rows[sel], p]
    cnt = on.sum(1).astype(F32)
    denorm = (cnt + DENORM_EPS).astype(F32)
    return (acc / denorm[:, None]).astype(F32)


# ----------------------------------------------------------------------------------------------
# (3) temporal token merge  --  layer.py:6-33 and dispatch :110-119
# ----------------------------------------------------------------------------------------------
def _rowsum(v: np.ndarray) -> np.ndarray:
    """Canonical 32-lane strided sum over the last axis of fp32 [..., C] (C % 4 == 0)."""
    v = np.asarray(v, dtype=F32)
    c = v.shape[-1]
    assert c % 4 == 0
    chunks = (c + 127) // 128
    padded = np.zeros(v.shape[:-1] + (chunks * 128,), dtype=F32)
    padded[..., :c] = v
    lanes = padded.reshape(v.shape[:-1] + (chunks, 32, 4))
    live = (np.arange(chunks * 128).reshape(chunks, 32, 4) < c)
    acc = np.zeros(v.shape[:-1] + (32,), dtype=F32)
    for k in range(chunks):
        for j in range(4):
            # lanes whose element lies beyond C skip the add (adding +0.0 could flip a -0.0)
            acc = np.where(live[k, :, j], acc + lanes[..., k, :, j], acc).astype(F32)
    idx = np.arange(32)
    for off in (16, 8, 4, 2, 1):
        acc = (acc + acc[..., idx ^ off]).astype(F32)
    return acc[..., 0]


def ttm_sims(x: np.ndarray) -> np.ndarray:
    """x fp32 [T, C] -> adjacent cosine similarities fp32 [T-1]  (layer.py:11-15)."""
    x = np.asarray(x, dtype=F32)
    ss = _rowsum((x * x).astype(F32))
    norm = np.maximum(np.sqrt(ss).astype(F32), NORM_EPS)
    xn = (x / norm[:, None]).astype(F32)
    return _rowsum((xn[:-1] * xn[1:]).astype(F32))


def _desc_key(s: np.ndarray) -> np.ndarray:
    """Order used by torch.topk: NaN ranks above every number."""
    return np.where(np.isnan(s), np.inf, s)


def ttm_boundaries(sims: np.ndarray, r: int):
    """r-th largest sim (duplicates counted, layer.py:17-18) and the strict-below cut mask
    (layer.py:24).  Returns (kth, cut bool [T-1])."""
    sims = np.asarray(sims, dtype=F32)
    order = np.sort(_desc_key(sims))[::-1]
    kth_key = order[r - 1]
    pick = sims[_desc_key(sims) == kth_key]
    kth = pick[0]
    with np.errstate(invalid="ignore"):
        cut = sims < kth
    return kth, cut


def ttm_merge(x: np.ndarray, cut: np.ndarray) -> np.ndarray:
    """Average maximal runs of tokens between cuts (layer.py:22-33), canonical order."""
    x = np.asarray(x, dtype=F32)
    out, acc, n = [], np.zeros(x.shape[1], dtype=F32), 0
    for t in range(x.shape[0]):
        acc = (acc + x[t]).astype(F32)
        n += 1
        if t == x.shape[0] - 1 or cut[t]:
            out.append((acc / F32(n)).astype(F32))
            acc, n = np.zeros(x.shape[1], dtype=F32), 0
    return np.stack(out)


def token_merge(x: np.ndarray, k_keep: int):
    """Object-level dispatch (layer.py:115-117).  Returns (tokens fp32 [k, C], cut mask or None,
    sims or None)."""
    t = x.shape[0]
    if t <= k_keep:
        return np.asarray(x, dtype=F32).copy(), None, None
    sims = ttm_sims(x)
    _, cut = ttm_boundaries(sims, t - k_keep)
    return ttm_merge(x, cut), cut, sims


# ----------------------------------------------------------------------------------------------
# (4) downcast + projector  --  layer.py:123 and :55-59,126
# ----------------------------------------------------------------------------------------------
def round_to(x: np.ndarray, dtype: str) -> np.ndarray:
    """Round fp32 to the model dtype and return it as fp32 ('f32' | 'bf16' | 'f16')."""
    x = np.asarray(x, dtype=F32)
    if dtype == "f32":
        return x
    if dtype == "f16":
        return x.astype(np.float16).astype(F32)
    if dtype == "bf16":
        u = x.view(np.uint32).astype(np.uint64)
        rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
        out = rounded.astype(np.uint32).view(F32)
        return np.where(np.isnan(x), x, out)
    raise ValueError(dtype)


def gelu_erf(x: np.ndarray) -> np.ndarray:
    from scipy.special import erf
    x64 = np.asarray(x, dtype=np.float64)
    return (0.5 * x64 * (1.0 + erf(x64 / np.sqrt(2.0)))).astype(F32)


def projector(tokens: np.ndarray, w1, b1, w2, b2, dtype: str) -> np.ndarray:
    """Linear -> exact GELU -> Linear with the hidden activation rounded to the model dtype
    after the first Linear and after GELU, as the reference's module chain does (layer.py:55-59).
    All operands are fp32 arrays holding model-dtype-representable values."""
    h = np.asarray(tokens, dtype=F32) @ np.asarray(w1, dtype=F32).T + np.asarray(b1, dtype=F32)
    h = round_to(h, dtype)
    h = round_to(gelu_erf(h), dtype)
    y = h @ np.asarray(w2, dtype=F32).T + np.asarray(b2, dtype=F32)
    return round_to(y, dtype)


# ----------------------------------------------------------------------------------------------
# whole path  --  MaskExtractor.forward, layer.py:63-128
# ----------------------------------------------------------------------------------------------
def encode(feats, masks, ann_indices, k_keep: int = 4, dtype: str = "f32", weights=None,
           pad_square: bool = False, n_out: int = PATCH):
    """Restatement of MaskExtractor.forward.

    feats   fp32 array [F_total, n_patch, C] holding model-dtype-representable values
    masks   list over samples of [q_i, H_i, W_i] arrays (or one [B, q, H, W] array)
    ann_indices[i][o] = list of global frame rows of object o of sample i
    Returns a dict with every intermediate the reference hides.
    """
    feats = np.asarray(feats, dtype=F32)
    on_all, rows_all, obj_len, n_rows = [], [], [], []
    for i in range(len(masks)):
        m = np.asarray(masks[i])
        if m.shape[0] == 0:                       # layer.py:73-75 fallback: one all-zero mask
            m = np.zeros((1, 336, 336), dtype=F32)
        rows = [r for obj in ann_indices[i] for r in obj]          # layer.py:92-95
        on_i = [mask_to_patches(m[j], n_out, pad_square) for j in range(m.shape[0])]
        if len(rows) != len(on_i):                # torch broadcasting of x*mask, layer.py:147
            if len(rows) == 1:                    # PixRQA quirk, SURVEY section 8b
                rows = rows * len(on_i)
            elif len(on_i) == 1:
                on_i = on_i * len(rows)
            else:
                raise ValueError("feature rows and mask count disagree")
        on_all.extend(on_i)
        rows_all.extend(rows)
        n_rows.append(len(rows))
        # object token ranges follow ann_indices lengths with a running offset (layer.py:112-119)
        obj_len.append([len(o) for o in ann_indices[i]])
    on = np.stack(on_all) if on_all else np.zeros((0, n_out * n_out), dtype=bool)
    pooled = mask_pool(feats, rows_all, on)
    tokens, counts, cuts, sims = [], [], [], []
    base = 0
    for i in range(len(masks)):
        start = base
        for t in obj_len[i]:
            tok, cut, s = token_merge(pooled[start:start + t], k_keep)
            tokens.append(tok)
            counts.append(tok.shape[0])
            cuts.append(cut)
            sims.append(s)
            start += t
        base += n_rows[i]
    merged = np.concatenate(tokens) if tokens else np.zeros((0, feats.shape[2]), dtype=F32)
    out = {"on": on, "frame_rows": np.asarray(rows_all), "pooled": pooled, "merged_f32": merged,
           "merged": round_to(merged, dtype), "counts": counts, "cuts": cuts, "sims": sims}
    if weights is not None:
        out["tokens"] = projector(out["merged"], *weights, dtype=dtype)
    return out
