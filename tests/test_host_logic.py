"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared symbol,
the tap-table helper equals the oracle, and the packer restates the reference's bookkeeping."""
import os
import re
import types

import numpy as np
import pytest
import torch

from oracle import restatement as R
from ufvideo_b200 import _cabi, packer, synth

CPU = torch.device("cpu")


def test_library_exports_every_symbol_the_header_declares():
    header = open(_cabi.HEADER).read()
    declared = set(re.findall(r"\b(ufv_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_cabi.EXPORTED), declared ^ set(_cabi.EXPORTED)
    lib = _cabi.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ufv_abi_version() == _cabi.ABI_VERSION


def test_ctypes_struct_mirrors_have_the_sizes_the_library_was_compiled_with():
    lib = _cabi.lib()
    import ctypes
    assert lib.ufv_struct_size(b"ufv_mask_desc") == packer.MASK_DESC.itemsize == 32
    assert lib.ufv_struct_size(b"ufv_peer_args") == ctypes.sizeof(_cabi.PeerArgs)
    assert lib.ufv_struct_size(b"ufv_dyn_args") == ctypes.sizeof(_cabi.DynArgs) == 256
    assert lib.ufv_struct_size(b"ufv_encode_args") == ctypes.sizeof(_cabi.EncodeArgs)
    assert lib.ufv_struct_size(b"nope") == -1


def test_argument_errors_are_reported_not_crashed():
    lib = _cabi.lib()
    assert lib.ufv_tap_table(0, 5, 27, 0, None) == -1            # UFV_E_NULL
    buf = np.zeros(4 * 27, np.int32)
    assert lib.ufv_tap_table(0, 5, 27, 0, buf.ctypes.data) == -2  # UFV_E_SHAPE
    assert b"out of range" in lib.ufv_last_error()
    with pytest.raises(_cabi.UfvError):
        _cabi.check(lib.ufv_linear(None, None, None, None, 4, 8, 8, 1, 0, None, 0, None))


@pytest.mark.parametrize("hw", [(384, 384), (720, 1280), (480, 854), (100, 37), (27, 27), (81, 81),
                                (54, 54), (13, 13), (1, 1), (2, 500), (1080, 1920)])
@pytest.mark.parametrize("pad", [False, True])
def test_tap_table_equals_oracle(hw, pad):
    h, w = hw
    t = packer.tap_table(h, w, 27, pad).reshape(4, 27)
    side, top, left = R.pad_to_square_offsets(h, w) if pad else (None, 0, 0)
    eh, ew = (side, side) if pad else (h, w)
    h0, h1, u0, u1 = R.axis_taps(eh)
    w0, w1, v0, v1 = R.axis_taps(ew)

    def fold(i, use, off, ext):
        s = i - off
        return np.where(use & (s >= 0) & (s < ext), s, -1)

    want = np.stack([fold(h0, u0, top, h), fold(h1, u1, top, h), fold(w0, v0, left, w), fold(w1, v1, left, w)])
    assert np.array_equal(t, want)


def test_plan_groups_objects_by_frame_and_reserves_token_slots():
    masks = [torch.zeros((7, 20, 30), dtype=torch.uint8), torch.zeros((3, 40, 40))]
    ann = [[[0, 1, 2], [2, 3], [1, 1]], [[4, 5, 6]]]
    plan = packer.build_plan(masks, ann, 7, 2, CPU)
    h = plan.host
    assert plan.n_masks == 10 and plan.n_obj == 4 and plan.max_len == 3
    assert h["obj_start"].tolist() == [0, 3, 5, 7] and h["obj_len"].tolist() == [3, 2, 2, 3]
    assert plan.slots.tolist() == [2, 2, 2, 2] and h["slot_off"].tolist() == [0, 2, 4, 6] and plan.m_pad == 8
    # groups: one per distinct feature row, members = object-frames reading it
    rows = [0, 1, 2, 2, 3, 1, 1, 4, 5, 6]
    got = {int(r): sorted(h["grp_member"][a:b].tolist())
           for r, a, b in zip(h["grp_row"], h["grp_off"][:-1], h["grp_off"][1:])}
    want = {r: [j for j, x in enumerate(rows) if x == r] for r in set(rows)}
    assert got == want and plan.max_group == 3
    d = h["mask_desc"]
    assert d["dtype"].tolist() == [_cabi.UFV_U8] * 7 + [_cabi.UFV_F32] * 3
    assert d["pitch"].tolist() == [30] * 7 + [40] * 3 and d["tap_off"].tolist() == [0] * 7 + [108] * 3
    e0 = masks[0].data_ptr() + np.arange(7) * 600
    assert np.array_equal(d["addr"][:7], e0.astype(np.uint64))
    assert d["addr"][7] == masks[1].data_ptr()
    for j in range(10):                                          # every mask knows its pool group
        g = d["group"][j]
        assert j in h["grp_member"][h["grp_off"][g]:h["grp_off"][g + 1]]


def test_plan_splits_frames_with_many_objects():
    masks = [torch.zeros((150, 8, 8), dtype=torch.uint8)]
    plan = packer.build_plan(masks, [[[0]]], 1, 4, CPU)          # PixRQA broadcast: one feature row
    sizes = np.diff(plan.host["grp_off"]).tolist()             # equal sub-groups of at most GROUP_SPLIT members
    n = -(-150 // packer.GROUP_SPLIT)
    assert plan.n_groups == n and sum(sizes) == 150 and max(sizes) - min(sizes) <= 1
    assert plan.max_group == max(sizes) <= packer.GROUP_SPLIT
    assert sorted(plan.host["grp_member"].tolist()) == list(range(150))
    assert plan.host["obj_len"].tolist() == [1] and plan.m_pad == 1


def test_plan_edge_cases():
    empty = packer.build_plan([torch.zeros((0, 50, 50))], [[[1]]], 2, 4, CPU)   # layer.py:73-75
    assert empty.n_masks == 1 and empty.host["mask_desc"]["pitch"].tolist() == [336]
    with pytest.raises(ValueError):
        packer.build_plan([torch.zeros((3, 8, 8))], [[[0, 1]]], 2, 4, CPU)
    with pytest.raises(IndexError):
        packer.build_plan([torch.zeros((1, 8, 8))], [[[5]]], 2, 4, CPU)
    none = packer.build_plan([], [], 0, 4, CPU)
    assert none.n_masks == 0 and none.m_pad == 0


def test_synth_is_deterministic_and_sharding_invariant():
    f_all, m_all, a_all = synth.make_batch(3, 4, 2, "blob", 64, 64)
    f1, m1, a1 = synth.make_batch(1, 4, 2, "blob", 64, 64, first_clip=1)
    assert np.array_equal(f_all[4:8], f1) and np.array_equal(m_all[1], m1[0])
    assert a_all[1] == [[r + 4 for r in o] for o in a1[0]]


def test_plan_cache_reuses_structure_and_patches_mask_addresses():
    ann = [[[0, 1], [1]]]
    a = torch.zeros((3, 16, 16), dtype=torch.uint8)
    b = torch.zeros((3, 16, 16), dtype=torch.uint8)
    p1 = packer.build_plan([a], ann, 2, 4, CPU)
    assert packer.build_plan([a], [[[0, 1], [1]]], 2, 4, CPU) is p1          # equal content -> same plan
    p2 = packer.build_plan([b], ann, 2, 4, CPU)
    assert p2 is p1 and p1.host["mask_desc"]["addr"][0] == b.data_ptr()     # addresses re-pointed
    off = p1.dev["mask_desc"] - p1.buffer.data_ptr()
    on_dev = p1.buffer[off:off + 96].numpy().view(packer.MASK_DESC)
    assert on_dev["addr"].tolist() == [b.data_ptr() + i * 256 for i in range(3)]
    assert packer.build_plan([a], [[[0, 1], [0]]], 2, 4, CPU) is not p1     # different structure


def test_rle_codec_round_trips_and_matches_the_oracle_decode():
    from ufvideo_b200 import rle
    g = synth.rng_for(99)
    for shape in [(5, 7), (384, 384), (3, 1), (1, 9), (27, 27)]:
        for dens in (0.0, 0.3, 1.0):
            m = (g.random(shape) < dens).astype(np.uint8)
            r = rle.encode(m)
            assert sum(r["counts"]) == m.size
            assert np.array_equal(R.rle_to_mask(r), m)                       # oracle's independent decoder
            packed = rle.counts_to_string(r["counts"])
            assert rle.counts_from_string(packed).tolist() == r["counts"]    # pycocotools string codec
            h, w, ends = rle.run_ends({"size": r["size"], "counts": packed})
            assert (h, w) == shape and ends.tolist() == np.cumsum(r["counts"]).tolist()
    with pytest.raises(ValueError):
        rle.run_ends({"size": [2, 2], "counts": [3, 3]})


def test_plan_accepts_run_length_samples_and_refreshes_their_descriptors():
    from ufvideo_b200 import rle
    masks = synth.masks_blob(5, 2, 3, 50, 70)
    ann = [[[0, 1, 2], [0, 1, 2]]]
    plan = packer.build_plan([[rle.encode(m) for m in masks]], ann, 3, 2, CPU)
    d = plan.host["mask_desc"]
    assert (d["dtype"] == _cabi.UFV_RLE).all() and (d["aux"] == 50).all() and not d["flags"].any()
    assert d["pitch"].tolist() == [len(rle.encode(m)["counts"]) for m in masks]
    again = packer.build_plan([[rle.encode(m) for m in masks[::-1]]], ann, 3, 2, CPU)   # same structure, new runs
    assert again is plan and again.host["mask_desc"]["pitch"].tolist() == d["pitch"].tolist()
    assert again.host["mask_desc"]["pitch"].tolist() == [len(rle.encode(m)["counts"]) for m in masks[::-1]]


# ---------------------------------------------------------------------------------------------
# property test: the packer against a direct restatement of the reference's bookkeeping
# ---------------------------------------------------------------------------------------------
from hypothesis import given, settings, strategies as st


@st.composite
def _batches(draw):
    n_samples = draw(st.integers(1, 3))
    n_rows, samples = 0, []
    for _ in range(n_samples):
        frames = draw(st.integers(1, 6))
        n_obj = draw(st.integers(1, 11))
        objs = [draw(st.lists(st.integers(0, frames - 1), min_size=0, max_size=7)) for _ in range(n_obj)]
        if sum(len(o) for o in objs) == 0:
            objs[0] = [0]
        samples.append((frames, [[n_rows + f for f in o] for o in objs]))
        n_rows += frames
    k = draw(st.integers(1, 5))
    return n_rows, samples, k


@settings(max_examples=60, deadline=None)
@given(_batches())
def test_plan_matches_a_direct_restatement_of_the_reference_loop(batch):
    """layer.py:92-95 flattens each sample's index lists (objects in order, frames in order) and pairs
    flattened position j with mask plane j; :112-119 walks the pooled rows object by object; :121-125
    concatenates min(T, K)-or-fewer tokens per object.  The plan must encode exactly that, and its groups
    must partition the object-frames by feature row with at most MAX_GROUP members each."""
    n_rows, samples, k = batch
    masks = [torch.zeros((sum(len(o) for o in objs), 9, 11), dtype=torch.uint8) for _, objs in samples]
    ann = [objs for _, objs in samples]
    plan = packer.build_plan(masks, ann, n_rows, k, CPU, use_cache=False)
    h = plan.host
    want_rows, want_start, want_len = [], [], []
    for _, objs in samples:                                   # the reference's loops, restated
        flat = [r for o in objs for r in o]
        start = len(want_rows)
        for o in objs:
            want_start.append(start)
            want_len.append(len(o))
            start += len(o)
        want_rows.extend(flat)
    assert plan.n_masks == len(want_rows) and plan.n_obj == len(want_len)
    assert h["obj_start"].tolist() == want_start and h["obj_len"].tolist() == want_len
    slots = [min(t, k) for t in want_len]
    assert plan.slots.tolist() == slots and plan.m_pad == sum(slots)
    assert h["slot_off"].tolist() == [sum(slots[:i]) for i in range(len(slots))]
    # groups: a partition of the object-frames, each group on one feature row, at most MAX_GROUP members
    go, gm, gr = h["grp_off"], h["grp_member"], h["grp_row"]
    assert sorted(gm.tolist()) == list(range(len(want_rows))) and go[0] == 0 and go[-1] == len(want_rows)
    for g in range(plan.n_groups):
        members = gm[go[g]:go[g + 1]]
        assert 1 <= len(members) <= _cabi.MAX_GROUP
        assert {want_rows[j] for j in members} == {int(gr[g])}
        assert (h["mask_desc"]["group"][members] == g).all()
    # mask planes: descriptor j points at plane j of its sample (uint8, contiguous: pitch 11, plane 99 bytes)
    off = 0
    for m in masks:
        q = m.shape[0]
        d = h["mask_desc"][off:off + q]
        assert (d["addr"] == m.data_ptr() + np.arange(q, dtype=np.uint64) * 99).all() and (d["pitch"] == 11).all()
        off += q


# ---------------------------------------------------------------------------------------------
# host logic added in round 2
# ---------------------------------------------------------------------------------------------
def _consumer_layout(seq_lens, region_pos, counts):
    """The region part of prepare_inputs_labels_for_multimodal (videorefer_arch.py:291-368) on symbolic rows:
    ('t', i) = text row i of the flattened batch, ('r', j) = flattened object-token row j, None = padding."""
    out, off, row, obj = [], 0, 0, 0
    for n, pos in zip(seq_lens, region_pos):
        if not pos:                                  # no placeholder: the cursor still advances (:263-264, :300-305)
            out.append([("t", off + i) for i in range(n)])
            row += counts[obj]
            obj += 1
        else:
            seq, last = [], 0
            for p in pos:
                seq += [("t", off + i) for i in range(last, p)]
                seq += [("r", row + j) for j in range(counts[obj])]
                row += counts[obj]
                obj += 1
                last = p + 1
            seq += [("t", off + i) for i in range(last, n)]
            out.append(seq)
        off += n
    l_max = max(len(x) for x in out)
    return [x + [None] * (l_max - len(x)) for x in out], l_max


def test_region_layout_equals_the_reference_consumer_on_symbolic_rows():
    from hypothesis import given, settings, strategies as st
    from ufvideo_b200.layer import RegionLayout

    @settings(max_examples=150, deadline=None, derandomize=True)
    @given(st.lists(st.tuples(st.integers(1, 12), st.lists(st.integers(0, 11), max_size=4, unique=True)), min_size=1,
                    max_size=5), st.integers(0, 10_000))
    def check(samples, seed):
        seq_lens = [n for n, _ in samples]
        region_pos = [sorted(p for p in pos if p < n) for n, pos in samples]
        n_obj = sum(max(len(p), 1) for p in region_pos)
        g = np.random.default_rng(seed)
        slots = g.integers(1, 5, n_obj)
        lay = RegionLayout(seq_lens, region_pos, slots)
        want, l_max = _consumer_layout(seq_lens, region_pos, slots.tolist())
        assert lay.l_max == l_max and lay.new_lens == [sum(x is not None for x in row) for row in want]
        src = lay.src_map.reshape(len(seq_lens), l_max)
        slot_off = np.concatenate([[0], np.cumsum(slots)])
        dest_of_token = {int(d): r for r, d in enumerate(lay.token_row_map) if d >= 0}
        for i, row in enumerate(want):
            for l, cell in enumerate(row):
                if cell is None:
                    assert src[i, l] == -1
                elif cell[0] == "t":
                    assert src[i, l] == cell[1]
                else:
                    assert src[i, l] == -2 and dest_of_token[i * l_max + l] == cell[1]
        # tokens of objects consumed by placeholder-less samples are dropped, all others placed exactly once
        placed = sum(len(p) for p in region_pos)
        assert (lay.token_row_map >= 0).sum() == sum(int(slots[o]) for o in _objects_with_placeholder(region_pos))
        assert len(dest_of_token) == (lay.token_row_map >= 0).sum() and placed <= n_obj
        assert slot_off[-1] == lay.token_row_map.size

    check()


def _objects_with_placeholder(region_pos):
    objs, obj = [], 0
    for pos in region_pos:
        if not pos:
            obj += 1
        else:
            objs += list(range(obj, obj + len(pos)))
            obj += len(pos)
    return objs


def test_region_layout_rejects_inconsistent_inputs():
    from ufvideo_b200.layer import RegionLayout
    with pytest.raises(ValueError):
        RegionLayout([4], [[1, 1]], [2, 2])              # duplicate placeholder
    with pytest.raises(ValueError):
        RegionLayout([4], [[5]], [2])                    # outside the sequence
    with pytest.raises(ValueError):
        RegionLayout([4, 4], [[1], [2]], [2])            # fewer objects than placeholders
    with pytest.raises(ValueError):
        RegionLayout([4], [[1]], [2, 2])                 # more objects than the samples consume


def test_algorithmic_pool_bytes_counts_each_feature_row_once_per_frame():
    """union over ALL object-frames of a feature row, however the packer grouped them."""
    masks = [torch.zeros((70, 8, 8), dtype=torch.uint8)]
    plan = packer.build_plan(masks, [[[0]]], 1, 4, CPU)          # 70 objects on one frame: several sub-groups
    assert plan.n_groups == -(-70 // packer.GROUP_SPLIT) > 1
    bits = np.zeros((70, 24), np.uint32)
    bits[:, 0] = 0b1111                                           # every object: patches 0..3
    bits[65, 1] = 1                                               # one object of a later sub-group: patch 32 too
    got = packer.algorithmic_pool_bytes(plan, bits, 1152, 2)
    assert got == 5 * 1152 * 2 + 70 * 1152 * 4 + 70 * 96


@settings(max_examples=40, deadline=None, derandomize=True)
@given(st.lists(st.integers(1, 90), min_size=1, max_size=4), st.integers(1, 64), st.integers(0, 10_000))
def test_plan_groups_partition_object_frames_for_any_split(objs_per_frame, split, seed):
    """Pool groups for any GROUP_SPLIT: every object-frame sits in exactly one group, a group's members read one
    feature row, no group exceeds the split, the sub-groups of a frame differ in size by at most one, and
    SURVEY 8(d)'s algorithmic bytes (union per FRAME) do not depend on the split."""
    import unittest.mock as mock
    g = synth.rng_for(seed)
    n_frames = len(objs_per_frame)
    ann, n_masks = [], 0
    for f, n in enumerate(objs_per_frame):                     # n objects annotated on frame f only
        ann.extend([f] for _ in range(n))
        n_masks += n
    masks = [torch.zeros((n_masks, 8, 8), dtype=torch.uint8)]
    bits = g.integers(0, 2 ** 32, size=(n_masks, 24), dtype=np.uint64).astype(np.uint32)
    with mock.patch.object(packer, "GROUP_SPLIT", split):
        plan = packer.build_plan(masks, [ann], n_frames, 4, CPU, use_cache=False)
    off, mem, row = plan.host["grp_off"], plan.host["grp_member"], plan.host["grp_row"]
    assert sorted(mem.tolist()) == list(range(n_masks))
    sizes = np.diff(off)
    assert sizes.min() >= 1 and sizes.max() == plan.max_group <= split
    frame_of = np.repeat(np.arange(n_frames), objs_per_frame)
    for gi in range(plan.n_groups):
        assert (frame_of[mem[off[gi]:off[gi + 1]]] == row[gi]).all()
    for f, n in enumerate(objs_per_frame):
        mine = sizes[row == f]
        assert mine.sum() == n and mine.size == -(-n // split) and mine.max() - mine.min() <= 1
    with mock.patch.object(packer, "GROUP_SPLIT", 64):
        whole = packer.build_plan(masks, [ann], n_frames, 4, CPU, use_cache=False)
    assert packer.algorithmic_pool_bytes(plan, bits, 1152, 2) == packer.algorithmic_pool_bytes(whole, bits, 1152, 2)


def test_await_counts_host_protocol(monkeypatch):
    """layer._await_counts without a GPU: the merge kernel publishes (epoch << 16) | count per object into pinned
    words; the host returns plan.slots itself on the tie-free fast path (one bytes compare), the real counts when
    an object tied below its reserved count, and keeps polling while any word still carries an older epoch."""
    import threading
    import time
    from ufvideo_b200 import layer
    masks = [torch.zeros((6, 8, 8), dtype=torch.uint8)]
    plan = packer.build_plan(masks, [[[0, 1, 2], [0, 1, 2]]], 3, 2, CPU, use_cache=False)     # 2 objects, T = 3, K = 2
    assert plan.slots.tolist() == [2, 2]
    words = np.zeros(2, np.int32)
    run = {"counts_np": words, "epoch": 7}
    words[:] = (7 << 16) | plan.slots
    assert layer._await_counts(plan, run, CPU) is plan.slots
    words[:] = [(7 << 16) | 2, (7 << 16) | 1]                      # the second object tied: one token
    got = layer._await_counts(plan, run, CPU)
    assert got is not plan.slots and got.tolist() == [2, 1]
    words[:] = [(7 << 16) | 2, (6 << 16) | 2]                      # second word still from the previous call
    idle = []

    def publish():
        time.sleep(0.05)
        words[1] = (7 << 16) | 2
    # every 65536 polls the loop asks the stream whether the launch is still running: here it always is
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: types.SimpleNamespace(query=lambda: False))
    t = threading.Thread(target=publish)
    t.start()
    got = layer._await_counts(plan, run, CPU, idle_work=lambda: idle.append(1))
    t.join()
    assert got is plan.slots and idle == [1]
