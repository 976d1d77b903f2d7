"""COCO run-length masks as a direct input of the object encoder.

The reference decodes every annotation to a dense H x W mask on the host
(``annToMask`` -> ``pycocotools.mask.decode``, ufvideo/mm_utils.py:22-33) and ships the dense
tensor to the GPU.  Kernel 1 only ever looks at 4 x 27 x 27 pixels of a mask, so the run-length
form itself is enough: the runs' cumulative end positions go to the device (a few hundred bytes
per mask) and each tap is one binary search.  Same patch bits as the dense path, bit for bit.

Format (COCO): pixels are numbered column-major, ``p = x * h + y``; ``counts`` are the lengths of
alternating runs starting with a run of zeros.  ``counts`` may be a list of ints (uncompressed) or
the LEB128-like byte string of ``pycocotools`` (compressed); both are handled here.  pycocotools is
not installed in the build image, so the string codec is checked by round trips only.
"""
from __future__ import annotations

import numpy as np


def counts_from_string(s) -> np.ndarray:
    """Decode the compressed ``counts`` string of pycocotools (maskApi.c: rleFrString)."""
    if isinstance(s, str):
        s = s.encode("ascii")
    out, p, n = [], 0, len(s)
    while p < n:
        x, k, more = 0, 0, True
        while more:
            c = s[p] - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(out) > 2:
            x += out[-2]
        out.append(x)
    return np.asarray(out, dtype=np.int64)


def counts_to_string(counts) -> bytes:
    """Encode run lengths as the compressed string of pycocotools (maskApi.c: rleToString)."""
    counts = [int(c) for c in counts]
    out = bytearray()
    for i, c in enumerate(counts):
        x = c - counts[i - 2] if i > 2 else c
        more = True
        while more:
            ch = x & 0x1F
            x >>= 5
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            out.append(ch + 48)
    return bytes(out)


def encode(mask: np.ndarray) -> dict:
    """Dense [H, W] mask (non-zero = on) -> uncompressed COCO RLE dict {'size': [h, w], 'counts': [...]}."""
    mask = np.asarray(mask)
    h, w = mask.shape
    flat = (mask != 0).reshape(-1, order="F").astype(np.int8)
    change = np.flatnonzero(np.diff(flat)) + 1
    edges = np.concatenate([[0], change, [flat.size]])
    counts = np.diff(edges).tolist()
    if flat.size and flat[0]:          # runs start with zeros: a leading empty zero-run
        counts = [0] + counts
    return {"size": [int(h), int(w)], "counts": counts}


def run_ends(rle: dict) -> tuple:
    """(h, w, int32 cumulative run ends).  Pixel p is on iff the first run whose end exceeds p has an odd index."""
    h, w = (int(v) for v in rle["size"])
    counts = rle["counts"]
    if isinstance(counts, (bytes, str)):
        counts = counts_from_string(counts)
    ends = np.cumsum(np.asarray(counts, dtype=np.int64))
    if ends.size and (ends[-1] > h * w or (np.asarray(counts) < 0).any()):
        raise ValueError("RLE counts do not fit the mask size")
    return h, w, ends.astype(np.int32)


def is_rle_sample(sample) -> bool:
    """A sample given as a list / tuple of COCO RLE dicts, one per object-frame."""
    return isinstance(sample, (list, tuple)) and len(sample) > 0 and isinstance(sample[0], dict) \
        and "counts" in sample[0] and "size" in sample[0]
