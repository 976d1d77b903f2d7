"""Clip sharding over ranks and the one result-collection collective.

The reference's eval drivers give each rank ``get_chunk(questions, num_chunks, rank)`` and merge
per-rank JSON files afterwards (eval/inference_PixRQA.py:186,214); clips never interact
(layer.py:68 loops over samples).  Here: contiguous blocks of clips per rank, no communication
during compute, and ONE all-gather (NCCL over NVLink on GPUs) of the padded per-rank token
tensor, with the int32 token counts riding in the tail rows of the same buffer.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi


def clip_block(n_clips: int, rank: int, world: int) -> range:
    """Contiguous block of clips owned by ``rank`` (sizes differ by at most one), so that
    concatenating ranks in order restores global clip order."""
    base, extra = divmod(n_clips, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def _tail_rows(pad_objs: int, row_bytes: int) -> int:
    return -(-(4 * (2 + pad_objs)) // row_bytes)


def pack_payload(tokens: torch.Tensor, counts: torch.Tensor, pad_rows: int, pad_objs: int) -> torch.Tensor:
    """[pad_rows + tail, hid] buffer: token rows first, then int32 [m, n_obj, counts...]."""
    m, hid = tokens.shape
    n_obj = counts.numel()
    if m > pad_rows or n_obj > pad_objs:
        raise ValueError(f"payload overflow: {m} rows / {n_obj} objects > pad {pad_rows} / {pad_objs}")
    row_bytes = hid * tokens.element_size()
    tail = _tail_rows(pad_objs, row_bytes)
    payload = torch.zeros((pad_rows + tail, hid), dtype=tokens.dtype, device=tokens.device)
    payload[:m] = tokens
    meta = payload[pad_rows:].view(torch.int32).reshape(-1)
    meta[0] = m
    meta[1] = n_obj
    meta[2:2 + n_obj] = counts.to(torch.int32)
    return payload


def unpack_payloads(gathered: torch.Tensor, pad_rows: int):
    """gathered [world, pad_rows + tail, hid] -> (tokens [sum m_r, hid], counts list[int])."""
    meta = gathered[:, pad_rows:].reshape(gathered.shape[0], -1).view(torch.int32).cpu().numpy()
    rows, counts = [], []
    for r in range(gathered.shape[0]):
        m, n_obj = int(meta[r, 0]), int(meta[r, 1])
        rows.append(gathered[r, :m])
        counts.extend(int(c) for c in meta[r, 2:2 + n_obj])
    return torch.cat(rows, dim=0), counts


def all_gather_tokens(tokens: torch.Tensor, counts: torch.Tensor, pad_rows: int, pad_objs: int,
                      group=None):
    """One all-gather of every rank's object tokens.  ``pad_rows`` / ``pad_objs`` are upper bounds
    every rank agrees on (e.g. max clips per rank x objects x K).  Returns the raw gathered buffer
    [world, rows, hid]; ``unpack_payloads`` turns it into (tokens, counts) in global clip order."""
    payload = pack_payload(tokens, counts, pad_rows, pad_objs)
    world = dist.get_world_size(group)
    rows, hid = payload.shape
    out = torch.empty((world * rows, hid), dtype=payload.dtype, device=payload.device)
    dist.all_gather_into_tensor(out, payload, group=group)     # concatenated along dim 0
    return out.view(world, rows, hid)


# ------------------------------------------------------------------------------------------------
# zero-copy variant: the kernels write straight into the payload
# ------------------------------------------------------------------------------------------------
# Payload of one rank, [pad_rows + tail, hid] of the model dtype:
#   rows [0, m_pad)        padded object tokens, object o at slot_off[o] (written by the projector)
#   tail, viewed as int32  [m_pad, n_obj, slots[pad_objs], counts[pad_objs]]
#                          slots = min(T_o, K) rows reserved per object (static per batch structure),
#                          counts = tokens actually produced (written by the merge kernel)
# Nothing is packed or copied on the way: the all-gather sends what the kernels produced.
_static_tail_cache: dict = {}


def _padded_tail_rows(pad_objs: int, row_bytes: int) -> int:
    return -(-(4 * (2 + 2 * pad_objs)) // row_bytes)


def new_payload(slots, pad_rows: int, pad_objs: int, hid: int, dtype, device):
    """Allocate one payload.  ``slots`` = plan.slots (int32 numpy, rows reserved per object).
    Returns (payload, tokens_view [m_pad, hid], counts_view int32 [n_obj])."""
    n_obj = int(len(slots))
    m_pad = int(slots.sum()) if n_obj else 0
    if m_pad > pad_rows or n_obj > pad_objs:
        raise ValueError(f"payload overflow: {m_pad} rows / {n_obj} objects > pad {pad_rows} / {pad_objs}")
    es = torch.empty((), dtype=dtype).element_size()
    tail = _padded_tail_rows(pad_objs, hid * es)
    payload = torch.empty((pad_rows + tail, hid), dtype=dtype, device=device)
    meta = payload[pad_rows:].view(torch.int32).reshape(-1)
    key = (slots.tobytes(), pad_objs, str(device))
    static = _static_tail_cache.get(key)
    if static is None:
        host = np.zeros(2 + pad_objs, dtype=np.int32)
        host[0], host[1] = m_pad, n_obj
        host[2:2 + n_obj] = slots
        static = torch.from_numpy(host).to(device)
        if len(_static_tail_cache) > 64:
            _static_tail_cache.clear()
        _static_tail_cache[key] = static
    meta[:2 + pad_objs].copy_(static, non_blocking=True)
    return payload, payload[:m_pad], meta[2 + pad_objs:2 + pad_objs + n_obj]


def all_gather_payload(payload: torch.Tensor, group=None, async_op: bool = False):
    """THE collective: one all-gather of every rank's payload.  Returns (gathered [world, rows, hid],
    work handle or None)."""
    world = dist.get_world_size(group)
    rows, hid = payload.shape
    out = torch.empty((world * rows, hid), dtype=payload.dtype, device=payload.device)
    work = dist.all_gather_into_tensor(out, payload, group=group, async_op=async_op)
    return out.view(world, rows, hid), work


def unpack_padded(gathered: torch.Tensor, pad_rows: int, pad_objs: int):
    """gathered [world, pad_rows + tail, hid] -> (tokens [sum counts, hid] in global clip order,
    counts list[int]).  Host-side; drops the zero-filled slots that merge ties left."""
    world = gathered.shape[0]
    meta = gathered[:, pad_rows:].reshape(world, -1).view(torch.int32).cpu().numpy()
    rows, counts = [], []
    for r in range(world):
        n_obj = int(meta[r, 1])
        slots = meta[r, 2:2 + n_obj]
        cnt = meta[r, 2 + pad_objs:2 + pad_objs + n_obj]
        off = 0
        for s, c in zip(slots, cnt):
            rows.append(gathered[r, off:off + int(c)])
            off += int(s)
        counts.extend(int(c) for c in cnt)
    tokens = torch.cat(rows, dim=0) if rows else gathered.new_zeros((0, gathered.shape[2]))
    return tokens, counts


# ------------------------------------------------------------------------------------------------
# the all-gather fused into the last Linear: tiles are stored into every rank's buffer over NVLink
# ------------------------------------------------------------------------------------------------
class PeerGather:
    """Symmetric gathered buffers for ``ufv_linear_gather`` (include/ufv_b200.h).

    Every rank owns ``ring`` copies of the gathered result, [world, pad_rows + tail, hid] each (same
    payload layout as ``new_payload``), allocated as torch symmetric memory so that all ranks can
    address them -- through the NVSwitch multicast alias when the fabric has one (one
    ``multimem.st`` reaches every rank), otherwise through one mapped pointer per peer.  A step:

        peer, tokens_view, counts_view, step = pg.begin(plan.slots)
        enc.forward_padded(feats, masks, ann, out=tokens_view, counts_out=counts_view, peer=peer)
        ...                       # next steps may be issued; nothing here blocks the host
        pg.wait(step)             # stream-ordered: all ranks' rows of `step` have landed
        gathered = pg.gathered(step)

    Flow control: step t may overwrite its ring slot once every rank has raised the flag of step
    ``t - ring + 2`` (whoever raised it has, in stream order, finished reading ``gathered`` of
    anything older than that).  ``begin(s)`` issues the stream-side wait only every ``ring / 2`` steps,
    on the flags of step ``s - 2``, which covers the following ``ring / 2`` steps.  Contract for the
    caller: consume ``gathered(t)`` (in stream order) before this rank's ``begin(t + 2)``.
    torch.distributed is used for the rendezvous only; no collective runs on the data path.
    """

    def __init__(self, pad_rows: int, pad_objs: int, hid: int, dtype, device, group=None, ring: int = 8,
                 use_multicast: bool = True):
        import torch.distributed._symmetric_memory as symm

        if ring < 4 or ring % 2:
            raise ValueError("ring must be even and >= 4")
        self.check_every = ring // 2              # flow-control waits are issued every ring / 2 steps
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _cabi.MAX_PEER_DST:
            raise ValueError(f"at most {_cabi.MAX_PEER_DST} ranks")
        self.pad_rows, self.pad_objs, self.hid, self.ring, self.device = pad_rows, pad_objs, hid, ring, device
        self.es = torch.empty((), dtype=dtype).element_size()
        self.rows = pad_rows + _padded_tail_rows(pad_objs, hid * self.es)
        self.tail_words = 2 + 2 * pad_objs
        self.buf = symm.empty((ring, self.world, self.rows, hid), dtype=dtype, device=device)
        self.flags = symm.empty((ring, self.world), dtype=torch.int32, device=device)
        self.flags.zero_()
        torch.cuda.synchronize(device)
        self.buf_h = symm.rendezvous(self.buf, group)
        self.flag_h = symm.rendezvous(self.flags, group)
        self.flag_h.barrier()
        self.multimem = bool(use_multicast and self.buf_h.has_multicast_support
                             and self.flag_h.has_multicast_support and self.buf_h.multicast_ptr
                             and self.flag_h.multicast_ptr)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=device)
        # time-out word of ufv_wait_flags: pinned host memory the wait kernel writes through its mapped
        # address, so wait() / gathered() / check() notice a time-out without synchronising
        from . import packer
        self.timed_out = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._timed_out_np = self.timed_out.numpy()
        self._timed_out_dev = packer._device_address(self.timed_out)
        self.step = 0
        self._wait_fn = _cabi.lib().ufv_wait_flags
        self._flags_ptr, self._timed_out_ptr = self.flags.data_ptr(), self._timed_out_dev
        self._side = None                         # side stream + events of push(), created on first use
        self._static = {}                         # slot -> (slots bytes, tokens view, counts view)
        self._args = [self._make_args(s) for s in range(ring)]

    def _row_bytes(self):
        return self.hid * self.es

    def _make_args(self, slot: int) -> _cabi.PeerArgs:
        a = _cabi.PeerArgs()
        off = (slot * self.world + self.rank) * self.rows * self._row_bytes()
        tail_off = off + self.pad_rows * self._row_bytes()
        flag_off = (slot * self.world + self.rank) * 4
        if self.multimem:
            bases, flag_bases = [self.buf_h.multicast_ptr], [self.flag_h.multicast_ptr]
        else:
            bases, flag_bases = list(self.buf_h.buffer_ptrs), list(self.flag_h.buffer_ptrs)
        for i, (b, f) in enumerate(zip(bases, flag_bases)):
            a.dst[i], a.tail_dst[i], a.flag[i] = b + off, b + tail_off, f + flag_off
        a.n_dst, a.multimem = len(bases), int(self.multimem)
        a.tail_src = self.buf.data_ptr() + tail_off
        a.ticket = self.ticket.data_ptr()
        a.tail_words = self.tail_words
        return a

    def begin(self, slots: np.ndarray):
        """Start a step.  Returns (peer args, local tokens view [m_pad, hid], counts view int32 [n_obj], step)."""
        step, slot = self.step, self.step % self.ring
        self.step += 1
        if step >= 2 and step % self.check_every == 0:
            self.wait(step - 2)
        key = slots.tobytes()
        cached = self._static.get(slot)
        if cached is None or cached[0] != key:   # views + the static tail part: once per (slot, batch structure)
            n_obj = int(len(slots))
            m_pad = int(slots.sum()) if n_obj else 0
            if m_pad > self.pad_rows or n_obj > self.pad_objs:
                raise ValueError(f"payload overflow: {m_pad} rows / {n_obj} objects > pad {self.pad_rows} / {self.pad_objs}")
            mine = self.buf[slot, self.rank]
            meta = mine[self.pad_rows:].view(torch.int32).reshape(-1)
            host = np.zeros(2 + self.pad_objs, dtype=np.int32)
            host[0], host[1] = m_pad, n_obj
            host[2:2 + n_obj] = slots            # reserved rows per object: static per batch structure
            meta[:2 + self.pad_objs].copy_(torch.from_numpy(host).to(self.device), non_blocking=True)
            cached = self._static[slot] = (key, mine[:m_pad], meta[2 + self.pad_objs:2 + self.pad_objs + n_obj])
        a = self._args[slot]
        a.flag_value = step + 1
        return a, cached[1], cached[2], step

    def push(self, step: int, peer, n_rows: int) -> None:
        """Collection off the critical path: the last Linear wrote this rank's rows LOCALLY (``out=`` the tokens view
        ``begin`` returned, no ``peer=``); push them, the tail and the arrival flag to every rank from a side stream,
        ordered after what the current stream has enqueued so far.  The current stream does not wait for NVLink."""
        if self._side is None:
            self._side = torch.cuda.Stream(self.device)
            self._events = [torch.cuda.Event() for _ in range(self.ring)]
            self._push_fn = _cabi.lib().ufv_peer_push
        slot = step % self.ring
        ev = self._events[slot]
        ev.record(torch.cuda.current_stream(self.device))
        self._side.wait_event(ev)
        src = self.buf.data_ptr() + (slot * self.world + self.rank) * self.rows * self._row_bytes()
        rc = self._push_fn(src, n_rows * self._row_bytes(), ctypes.byref(peer), self._side.cuda_stream)
        if rc:
            _cabi.check(rc)

    def wait(self, step: int) -> None:
        """Make the current stream wait until every rank's rows of ``step`` are in this rank's copy.
        Raises if an earlier wait on this object gave up (a peer never raised its flag): the stream would
        otherwise run on over a half-filled gathered buffer."""
        self.check()
        slot = step % self.ring
        rc = self._wait_fn(self._flags_ptr + slot * self.world * 4, self.world, step + 1, 10000, self._timed_out_ptr,
                           torch._C._cuda_getCurrentRawStream(self.device.index))
        if rc:
            _cabi.check(rc)

    def gathered(self, step: int) -> torch.Tensor:
        """This rank's copy of the gathered result of ``step``: [world, pad_rows + tail, hid]."""
        self.check()
        return self.buf[step % self.ring]

    def check(self) -> None:
        """Raise if a wait kernel that has already run gave up (reads one pinned host word, no sync)."""
        if int(self._timed_out_np[0]):
            raise RuntimeError("PeerGather: timed out waiting for a peer's arrival flag; the gathered "
                               "buffer of that step is incomplete")
