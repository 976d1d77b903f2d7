"""Host-side packer: python ``masks`` / ``ann_indices`` lists -> flat descriptor arrays.

Restates the bookkeeping of the reference's forward loop (ufvideo/model/layer.py:66-119) as data:
which mask plane pairs with which feature row (layer.py:92-98), which pooled rows belong to
which object (layer.py:112-119) and where each object's tokens land in the output
(layer.py:121-125).  Pure integer work on the host; everything is uploaded in one buffer, and a
plan is reused when the same batch structure comes back (only the mask base addresses are
patched).
"""
from __future__ import annotations

import ctypes
import itertools
import marshal
import os
import threading
import weakref
from collections import OrderedDict
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _cabi, rle as _rle

_MASK_DTYPES = {torch.uint8: _cabi.UFV_U8, torch.bool: _cabi.UFV_U8, torch.float32: _cabi.UFV_F32,
                torch.bfloat16: _cabi.UFV_BF16, torch.float16: _cabi.UFV_F16}
FEAT_DTYPES = {torch.float32: _cabi.UFV_F32, torch.bfloat16: _cabi.UFV_BF16, torch.float16: _cabi.UFV_F16}

# mirror of struct ufv_mask_desc (include/ufv_b200.h), 32 bytes
MASK_DESC = np.dtype([("addr", "<u8"), ("pitch", "<i4"), ("dtype", "<i4"), ("tap_off", "<i4"),
                      ("group", "<i4"), ("flags", "<i4"), ("aux", "<i4")])
assert MASK_DESC.itemsize == 32

# A pinned host mask tensor is not copied: kernel 1 reads it through its mapped device address and
# touches only the rows its taps need (384x384 fp32: 83 KB of 590 KB).  Developer knobs below.
READ_PINNED_MASKS_IN_PLACE = True
READ_MODE = "auto"          # "auto": row mode for host-resident masks, tap mode in HBM; "rows" / "taps" force one

_tap_cache: dict = {}
_plan_cache: "OrderedDict[tuple, EncodePlan]" = OrderedDict()
PLAN_CACHE_SIZE = 16
RUNS_PER_PLAN = 4           # (module, stream) pairs that keep their own run state on one cached plan
_lock = threading.RLock()   # plan cache, tap cache and the identity fast path are process-wide


def release_run(run: dict) -> None:
    """Destroy the CUDA graph execs a run captured (a graph still executing is released by the driver when
    it completes).  Called when a run is rebound to other pointers and when its plan leaves the cache."""
    for handle in run["graphs"].values():
        _cabi.lib().ufv_encode_graph_destroy(handle)
    run["graphs"].clear()
    run["awaited_calls"] = 0
    run.pop("graph_args", None)


def _evict(plan: "EncodePlan") -> None:
    for run in plan.runs.values():
        release_run(run)
    plan.runs.clear()
    plan.run = None


def tap_table(h: int, w: int, n_out: int, pad_square: bool) -> np.ndarray:
    """int32 [4 * n_out] tap table of an h x w mask (ufv_tap_table; replaces ATen's bilinear
    index computation, layer.py:139, plus the 'pad' mode of layer.py:77-86)."""
    key = (h, w, n_out, bool(pad_square))
    hit = _tap_cache.get(key)
    if hit is None:
        hit = np.empty(4 * n_out, dtype=np.int32)
        _cabi.check(_cabi.lib().ufv_tap_table(h, w, n_out, int(pad_square),
                                              hit.ctypes.data_as(ctypes.c_void_p)))
        _tap_cache[key] = hit
    return hit


@dataclass
class EncodePlan:
    """Everything the kernels need to know about one batch, as host arrays + one device copy."""
    n_masks: int                 # object-frames q (pooled rows)
    n_groups: int
    max_group: int
    n_obj: int
    max_len: int
    m_pad: int                   # token rows reserved: sum over objects of min(T_o, K)
    slots: np.ndarray            # int32 [n_obj] = min(T_o, K): token count unless the merge ties
    host: dict = field(default_factory=dict)      # name -> numpy array
    dev: dict = field(default_factory=dict)       # name -> device pointer (int)
    buffer: torch.Tensor | None = None            # owns the device memory behind ``dev``
    sample_of: np.ndarray | None = None           # int32 [q] sample each object-frame came from
    plane_off: np.ndarray | None = None           # int64 [q] byte offset of its plane in that sample
    base_ptrs: tuple = ()                         # mask tensor base addresses baked into ``buffer``
    run: dict | None = None                       # run state of the latest call (introspection; layer.py)
    layouts: dict = field(default_factory=dict)   # (seq_lens, region_pos) -> RegionLayout + its device maps (layer.forward_into)
    runs: dict = field(default_factory=dict)      # (module id, stream) -> workspace, EncodeArgs, pinned counts
                                                  # words, captured graphs of that pair (layer.py)
    keepalive: object = None                      # buffers the packer itself created for the last call (copies of
                                                  # host / strided / exotic masks, run-length uploads): the
                                                  # descriptors point into them
    rle_rows: list = field(default_factory=list)  # run-length samples: descriptor rows refreshed on every call
    cache_key: object = None                      # key under which the plan sits in the plan cache
    any_row_mode: int = 0                         # some descriptor asks for row mode (kernel 1 variant)
    expect_counts: list = field(default_factory=list)
    slots_bytes: bytes = b""                      # ``slots`` as bytes: the no-ties fast comparison


class RleSample:
    """One sample's masks given as COCO run-length dicts (ufvideo_b200/rle.py), uploaded as the int32
    cumulative run ends of all its object-frames back to back.  Quacks like a mask tensor as far as the
    packer needs; per-mask run counts and offsets replace the uniform row pitch / plane stride."""

    dtype = "rle"

    def __init__(self, rles, device):
        parts = [_rle.run_ends(r) for r in rles]
        sizes = {(h, w) for h, w, _ in parts}
        if len(sizes) != 1:
            raise ValueError("all run-length masks of one sample must share one image size")
        (h, w), = sizes
        self.shape = (len(parts), h, w)
        self.n_runs = np.asarray([p[2].size for p in parts], dtype=np.int32)
        self.run_off = np.concatenate([[0], np.cumsum(self.n_runs)[:-1]]).astype(np.int64)
        ends = np.concatenate([p[2] for p in parts] + [np.zeros(1, np.int32)])      # never empty
        staging = torch.from_numpy(ends)
        if device.type == "cuda":
            staging = staging.pin_memory()
        self.buf = staging.to(device, non_blocking=True)
        self.device = self.buf.device

    def data_ptr(self):
        return self.buf.data_ptr()

    def stride(self, dim=None):
        return ("rle",) if dim is None else 1

    def element_size(self):
        return 4


def _as_mask_list(masks, device):
    """Reference contract: list of [q_i, H_i, W_i] tensors, or one
    [B, q, H, W] tensor whose len() is the sample count.

    Returns (list, temporaries).  ``temporaries`` are the buffers created HERE -- device copies of
    pageable-host / other-device / non-contiguous / exotic-dtype masks, the zero mask that stands in for
    an empty sample, uploaded run-length buffers.  Their addresses go into the plan's descriptors, so the
    plan must keep them alive until its next call replaces them, and a call that needed any of them may not
    be short-cut by the identity fast path (their content has to be copied again)."""
    out, temps = [], []
    for i in range(len(masks)):
        m = masks[i]
        if isinstance(m, RleSample) or _rle.is_rle_sample(m):        # COCO run-length masks, never densified
            if not isinstance(m, RleSample):
                m = RleSample(m, device)
            temps.append(m)                      # also caller-built RleSamples: the plan pins what it points into
            out.append(m)
            continue
        if not torch.is_tensor(m):
            m = torch.as_tensor(m)
        given = m
        shape = m.shape
        if len(shape) != 3:
            raise ValueError(f"masks[{i}] must be [q, H, W], got {tuple(shape)}")
        if shape[0] == 0:                        # layer.py:73-75: substitute one all-zero mask
            m = _zero_mask(device)
        if m.dtype not in _MASK_DTYPES:          # exotic dtypes: binarise once on the device
            m = (m.to(device) > 0).to(torch.uint8)
        if m.device != device:
            in_place = (READ_PINNED_MASKS_IN_PLACE and device.type == "cuda"
                        and m.device.type == "cpu" and m.is_pinned())
            if not in_place:                     # pageable host memory / another device: copy
                m = m.to(device, non_blocking=True)
        if m.stride(-1) != 1:
            m = m.contiguous()
        if m is not given and (not torch.is_tensor(given) or m.data_ptr() != given.data_ptr() or m.device != given.device):
            temps.append(m)
        out.append(m)
    return out, temps


_zero_masks: dict = {}


def _zero_mask(device):
    """The all-zero 336 x 336 mask of layer.py:73-75, one per device for the life of the process (read-only)."""
    z = _zero_masks.get(device)
    if z is None:
        z = _zero_masks[device] = torch.zeros((1, 336, 336), dtype=torch.uint8, device=device)
    return z


def _device_address(m: torch.Tensor) -> int:
    """Address kernel 1 reads the mask at: the tensor's own pointer on the device, the mapped
    device alias of a pinned host tensor otherwise (the kernel then reads it in place over PCIe)."""
    if isinstance(m, RleSample) or m.device.type != "cpu" or not m.is_pinned():
        return m.data_ptr()
    dev = ctypes.c_uint64(0)
    _cabi.check(_cabi.lib().ufv_device_address(ctypes.c_void_p(m.data_ptr()), ctypes.byref(dev)))
    return int(dev.value)


def _same_indices(a, b) -> bool:
    try:
        return bool(a == b)
    except Exception:   # noqa: BLE001 -- e.g. tensors inside the lists: take the slow path
        return False


# Object-frames per pool group (one CTA per group and channel slice).  The kernel takes up to MAX_GROUP = 64, but
# with many objects on a frame it is bound by instructions and warps in flight, not by bytes: sub-groups of 16 (two
# members per consumer warp, six CTAs per SM) measured 74 us against 86 us for 2 x 64 frames x 32 objects and 72
# against 80 us for 64 frames x 64 objects; 8 is worse (285 against 201 us on c4).  UFV_GROUP_SPLIT: developer sweeps.
GROUP_SPLIT = min(_cabi.MAX_GROUP, max(1, int(os.environ.get("UFV_GROUP_SPLIT", 16))))

_last_call: list = [None]     # (mask tensor objects, their data_ptrs and shapes, scalar args, ann bytes, plan)


def build_plan(masks, ann_indices, n_feat_rows: int, k_keep: int, device, pad_square: bool = False,
               n_out: int = _cabi.MAX_PATCH_SIDE, use_cache: bool = True) -> EncodePlan:
    with _lock:
        return _build_plan_locked(masks, ann_indices, n_feat_rows, k_keep, device, pad_square, n_out, use_cache)


def _build_plan_locked(masks, ann_indices, n_feat_rows, k_keep, device, pad_square, n_out, use_cache) -> EncodePlan:
    # Fastest path: the very same mask tensor objects (same storage, same shape) and the same index
    # content as the previous call -> the previous plan, without re-deriving any descriptor.  The index
    # lists are compared by value against a private copy (a nested list compare runs at C speed).
    last = _last_call[0]
    scalars = (n_feat_rows, k_keep, bool(pad_square), n_out, device, READ_MODE)
    if (use_cache and last is not None and last[2] == scalars and len(masks) == len(last[0])
            and _same_indices(ann_indices, last[3])):
        same = True
        for m, (ref, ptr, shape) in zip(masks, last[0]):      # weak references: no mask tensor is kept alive
            if m is not ref() or m.data_ptr() != ptr or m.shape != shape:
                same = False
                break
        if same and _plan_cache.get(last[4].cache_key) is last[4]:
            return last[4]
    ann_bytes = None
    if use_cache:
        try:                                     # lists of python ints: one C-speed serialisation
            ann_bytes = marshal.dumps(ann_indices)
        except ValueError:                       # numpy / tensor scalars inside
            ann_bytes = None
    plan = _lookup_or_build(masks, ann_indices, ann_bytes, n_feat_rows, k_keep, device, pad_square, n_out,
                            use_cache)
    if plan.keepalive:
        # some mask was copied / converted for this call: the next call must take the slow path again so that
        # the copy is refreshed (in-place edits of host masks) and the descriptors point at live memory
        _last_call[0] = None
    elif use_cache and ann_bytes is not None and torch.is_tensor(masks) is False:
        try:
            _last_call[0] = ([(weakref.ref(m), m.data_ptr(), m.shape) for m in masks], None, scalars,
                             marshal.loads(ann_bytes), plan)
        except (AttributeError, TypeError):      # non-tensor mask entries: no identity fast path
            _last_call[0] = None
    return plan


def _lookup_or_build(masks, ann_indices, ann_bytes, n_feat_rows, k_keep, device, pad_square, n_out,
                     use_cache) -> EncodePlan:
    masks, temps = _as_mask_list(masks, device)
    if len(ann_indices) != len(masks):
        raise ValueError("ann_indices and masks disagree on the number of samples")
    ptrs = tuple(_device_address(m) for m in masks)
    key = None
    if use_cache:
        ann_key = ann_bytes
        if ann_key is None:
            ann_key = tuple(tuple(tuple(int(r) for r in o) for o in s) for s in ann_indices)
        key = (ann_key, tuple((m.shape, m.stride(), m.dtype, m.device.type) for m in masks),
               n_feat_rows, k_keep, bool(pad_square), n_out, str(device), READ_MODE)
        plan = _plan_cache.get(key)
        if plan is not None:
            _plan_cache.move_to_end(key)
            if plan.base_ptrs != ptrs or plan.rle_rows:   # same structure, new mask data: patch descriptors
                _patch_addresses(plan, ptrs, device, masks)
            # buffers made for this call replace those of the previous one; the old ones return to the caching
            # allocator, which hands them out again only to work ordered after the launches that read them
            plan.keepalive = temps or None
            return plan
    plan = _build(masks, ann_indices, n_feat_rows, k_keep, device, pad_square, n_out, ptrs)
    plan.keepalive = temps or None
    plan.cache_key = key
    if key is not None:
        _plan_cache[key] = plan
        while len(_plan_cache) > PLAN_CACHE_SIZE:
            _evict(_plan_cache.popitem(last=False)[1])
    return plan


def _build(masks, ann_indices, n_feat_rows, k_keep, device, pad_square, n_out, ptrs) -> EncodePlan:
    # Per-sample scalars first (python), then one np.repeat per descriptor field: the per-row work is numpy's.
    n_s = len(masks)
    s_pitch, s_dtype, s_tap, s_aux, s_plane_stride = (np.zeros(n_s, np.int64) for _ in range(5))
    counts = np.zeros(n_s, np.int64)
    rows_all, planes_all, obj_start, obj_len = [], [], [], []
    rle_rows = []                                # (sample, its descriptor rows, the mask plane of each row)
    taps, tap_chunks, tap_len = {}, [], 0
    base = 0
    chain = itertools.chain.from_iterable
    for i, m in enumerate(masks):
        q, h, w = m.shape
        tkey = (h, w)
        toff = taps.get(tkey)
        if toff is None:
            toff = taps[tkey] = tap_len
            tap_chunks.append(tap_table(h, w, n_out, pad_square))
            tap_len += 4 * n_out
        sample = ann_indices[i]
        lens = np.fromiter(map(len, sample), dtype=np.int64, count=len(sample))
        rows = np.fromiter(chain(sample), dtype=np.int64, count=int(lens.sum()))       # layer.py:92-95
        b = rows.size
        planes = np.arange(q, dtype=np.int64)
        if b != q:                               # torch broadcasting of x * mask, layer.py:147
            if b == 1:
                rows = np.repeat(rows, q)
            elif q == 1:
                planes = np.zeros(b, dtype=np.int64)
            else:
                raise ValueError(f"sample {i}: {b} feature rows cannot pair with {q} masks")
        n_i = rows.size
        counts[i], s_tap[i] = n_i, toff
        if isinstance(m, RleSample):             # per-mask run count / offset instead of pitch / plane stride
            rle_rows.append((i, np.arange(base, base + n_i), planes))
            s_dtype[i], s_aux[i] = _cabi.UFV_RLE, h
        else:
            s_pitch[i], s_dtype[i] = m.stride(1), _MASK_DTYPES[m.dtype]
            s_plane_stride[i] = m.stride(0) * m.element_size()
        rows_all.append(rows)
        planes_all.append(planes)
        starts = np.cumsum(lens) - lens          # running offset over pooled rows, layer.py:112-119
        obj_start.append(base + starts)
        obj_len.append(np.clip(np.minimum(lens, n_i - starts), 0, None))
        base += n_i

    cat = lambda parts, dt: np.concatenate(parts).astype(dt, copy=False) if parts else np.zeros(0, dt)  # noqa: E731
    planes_cat = cat(planes_all, np.int64)
    sample_of = [np.repeat(np.arange(n_s, dtype=np.int32), counts)]
    pitch = [np.repeat(s_pitch, counts).astype(np.int32)]
    dtype_id = [np.repeat(s_dtype, counts).astype(np.int32)]
    tap_off = [np.repeat(s_tap, counts).astype(np.int32)]
    aux = [np.repeat(s_aux, counts).astype(np.int32)]
    plane_off = [planes_cat * np.repeat(s_plane_stride, counts)]
    for i, rows_i, planes in rle_rows:           # run-length samples: per-mask values
        plane_off[0][rows_i] = masks[i].run_off[planes] * 4
        pitch[0][rows_i] = masks[i].n_runs[planes]
    obj_start = cat(obj_start, np.int64).tolist()
    obj_len = cat(obj_len, np.int64).tolist()
    q_total = base
    rows_all = cat(rows_all, np.int64)
    if q_total and (rows_all.min() < 0 or rows_all.max() >= n_feat_rows):
        raise IndexError(f"ann_indices refer to feature rows outside [0, {n_feat_rows})")

    # groups: object-frames that read the same feature row share one staged copy of it
    order = np.argsort(rows_all, kind="stable")
    sorted_rows = rows_all[order]
    if q_total:
        run_start = np.flatnonzero(np.r_[True, sorted_rows[1:] != sorted_rows[:-1]])
    else:
        run_start = np.zeros(0, np.int64)
    run_len = np.diff(np.r_[run_start, q_total])
    if q_total and run_len.max() > GROUP_SPLIT:
        # a frame with more object-frames than one CTA takes is cut into equal sub-groups (each re-reads the rows
        # it needs; they run side by side, so the re-reads are L2 hits)
        pieces = []
        for s, l in zip(run_start.tolist(), run_len.tolist()):
            n = -(-l // GROUP_SPLIT)
            cuts = [s + (l * i) // n for i in range(n + 1)]
            pieces.extend((a, b - a) for a, b in zip(cuts[:-1], cuts[1:]))
        run_start = np.array([p[0] for p in pieces], dtype=np.int64)
        run_len = np.array([p[1] for p in pieces], dtype=np.int64)
    n_groups = int(run_start.size)
    group_of = np.empty(q_total, dtype=np.int32)
    group_of[order] = np.repeat(np.arange(n_groups, dtype=np.int32), run_len)

    plan_sample = cat(sample_of, np.int32)
    desc = np.zeros(q_total, dtype=MASK_DESC)
    desc["pitch"] = cat(pitch, np.int32)
    desc["dtype"] = cat(dtype_id, np.int32)
    desc["tap_off"] = cat(tap_off, np.int32)
    desc["group"] = group_of
    desc["aux"] = cat(aux, np.int32)
    on_host = np.asarray([m.device.type == "cpu" and not isinstance(m, RleSample) for m in masks], dtype=bool)
    rows_mode = on_host[plan_sample] if READ_MODE == "auto" else np.full(q_total, READ_MODE == "rows")
    desc["flags"] = rows_mode.astype(np.int32)
    any_row_mode = int(rows_mode.any())

    obj_len_a = np.asarray(obj_len, dtype=np.int32)
    slots = np.minimum(obj_len_a, k_keep).astype(np.int32)
    slot_off = (np.concatenate([[0], np.cumsum(slots)[:-1]]).astype(np.int32) if slots.size
                else np.zeros(0, np.int32))
    host = {
        "mask_desc": desc,
        "taps": cat(tap_chunks, np.int32),
        "grp_row": sorted_rows[run_start].astype(np.int32) if n_groups else np.zeros(0, np.int32),
        "grp_off": np.r_[run_start, q_total].astype(np.int32),
        "grp_member": order.astype(np.int32),
        "obj_start": np.asarray(obj_start, dtype=np.int32),
        "obj_len": obj_len_a,
        "slot_off": slot_off,
    }
    plan = EncodePlan(n_masks=q_total, n_groups=n_groups,
                      max_group=int(run_len.max()) if n_groups else 1,
                      n_obj=len(obj_len), max_len=int(obj_len_a.max()) if len(obj_len) else 1,
                      m_pad=int(slots.sum()), slots=slots, host=host,
                      sample_of=cat(sample_of, np.int32), plane_off=cat(plane_off, np.int64),
                      expect_counts=[int(s) for s in slots], slots_bytes=slots.tobytes(),
                      any_row_mode=any_row_mode, rle_rows=rle_rows)
    _fill_addresses(plan, ptrs)
    _upload(plan, device)
    return plan


def _fill_addresses(plan: EncodePlan, ptrs, masks=None) -> None:
    if masks is not None:
        for i, rows, planes in plan.rle_rows:    # run-length data changes with every call
            plan.plane_off[rows] = masks[i].run_off[planes] * 4
            plan.host["mask_desc"]["pitch"][rows] = masks[i].n_runs[planes]
    if plan.n_masks:
        base = np.asarray(ptrs, dtype=np.uint64)
        plan.host["mask_desc"]["addr"] = base[plan.sample_of] + plan.plane_off.astype(np.uint64)
    plan.base_ptrs = tuple(ptrs)


def _patch_addresses(plan: EncodePlan, ptrs, device, masks=None) -> None:
    _fill_addresses(plan, ptrs, masks)
    desc = plan.host["mask_desc"]
    staging = torch.from_numpy(desc.view(np.uint8).reshape(-1).copy())
    if device.type == "cuda":
        staging = staging.pin_memory()
    off = plan.dev["mask_desc"] - plan.buffer.data_ptr()
    plan.buffer[off:off + staging.numel()].copy_(staging, non_blocking=True)   # stream-ordered


def _upload(plan: EncodePlan, device) -> None:
    """One pinned staging buffer, one H2D copy; every array starts 16-byte aligned."""
    offsets, total = {}, 0
    for name, arr in plan.host.items():
        offsets[name] = total
        total += (arr.nbytes + 15) // 16 * 16
    staging = torch.empty(max(total, 16), dtype=torch.uint8, pin_memory=device.type == "cuda")
    view = staging.numpy()
    for name, arr in plan.host.items():
        view[offsets[name]:offsets[name] + arr.nbytes] = arr.view(np.uint8).reshape(-1)
    plan.buffer = staging.to(device, non_blocking=True)
    p0 = plan.buffer.data_ptr()
    plan.dev = {name: p0 + off for name, off in offsets.items()}
    _evict(plan)


def algorithmic_pool_bytes(plan: EncodePlan, bits: np.ndarray, c: int, feat_bytes: int) -> int:
    """SURVEY.md section 8(d), pool term: every feature row a frame's objects need, ONCE per frame -- the union
    of the on-patches over ALL object-frames that read that feature row, however the packer grouped them --
    plus the fp32 pooled rows written and the patch bitmasks read.  ``bits`` = uint32 [q, 24] from kernel 1."""
    bits = np.asarray(bits).view(np.uint32).reshape(plan.n_masks, -1)
    go, gm, gr = plan.host["grp_off"], plan.host["grp_member"], plan.host["grp_row"]
    per_row: dict = {}
    for g in range(plan.n_groups):
        u = np.bitwise_or.reduce(bits[gm[go[g]:go[g + 1]]], axis=0)
        r = int(gr[g])
        per_row[r] = u if r not in per_row else (per_row[r] | u)
    union = sum(int(np.unpackbits(u.view(np.uint8)).sum()) for u in per_row.values())
    return union * c * feat_bytes + plan.n_masks * c * 4 + plan.n_masks * bits.shape[1] * 4
