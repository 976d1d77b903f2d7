"""The oracle (oracle/restatement.py) against the golden vectors produced by the REAL reference
(oracle/gen_golden.py, run in the build container).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import golden_cases as gc
from oracle import restatement as R
from ufvideo_b200 import synth


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_resize_bits_match_reference(golden_dir):
    g = load(golden_dir, "resize.npz")
    meta = json.loads(str(g["meta"]))
    row = 0
    for m in meta:
        masks = gc.resize_masks(m["h"], m["w"])
        assert gc.digest(masks) == m["sha"], "input generator drifted"
        mine = R.pack_bits(np.stack([R.mask_to_patches(x) for x in masks]))
        assert np.array_equal(mine, g["bits"][row:row + m["n"]]), (m["h"], m["w"])
        row += m["n"]
    assert row == g["bits"].shape[0]


def test_resize_accepts_float_and_bool_masks():
    m = gc.resize_masks(100, 37)[0]
    a = R.mask_to_patches(m)
    assert np.array_equal(a, R.mask_to_patches(m.astype(np.float32)))
    assert np.array_equal(a, R.mask_to_patches(m.astype(bool)))
    assert np.array_equal(a, R.mask_to_patches(m.astype(np.float32) * 0.25))


@pytest.mark.parametrize("name", [c[0] for c in gc.POOL_CASES])
def test_pool_matches_reference(golden_dir, name):
    g = load(golden_dir, "pool.npz")
    feats, masks, rows = gc.pool_inputs(name)
    assert gc.digest(feats, masks, rows) == str(g[name + "_sha"])
    on = np.stack([R.mask_to_patches(m) for m in masks])
    pooled = R.mask_pool(feats, rows, on)
    assert np.abs(pooled - g[name]).max() <= 1e-5          # fp32 tolerance of north_star
    zero_rows = ~on.any(1)
    assert (pooled[zero_rows] == 0).all()                  # cnt == 0 -> exact zero vector


def test_ttm_cuts_and_tokens_match_reference(golden_dir):
    g = load(golden_dir, "ttm.npz")
    meta = json.loads(str(g["meta"]))
    assert len(meta) == len(gc.ttm_cases())
    for idx, m in enumerate(meta):
        x = gc.ttm_tokens(m["family"], m["t"], m["seed"])
        assert gc.digest(x) == m["sha"]
        tok, cut, _ = R.token_merge(x, m["k"])
        ref_cut = np.unpackbits(g[f"case{idx}_cut"])[: m["t"] - 1].astype(bool)
        assert np.array_equal(cut, ref_cut), m                 # merge decisions: bit-exact
        assert tok.shape[0] == m["rows"]
        ref_tok = g[f"case{idx}_merged"]
        mine = tok if m["full"] else tok[:, :64]
        assert np.abs(mine - ref_tok).max() <= 1e-5, m


def test_ttm_passthrough_and_ties():
    x = gc.ttm_tokens("random", 5, 1)
    tok, cut, sims = R.token_merge(x, 8)
    assert cut is None and np.array_equal(tok, x)
    tok, cut, _ = R.token_merge(np.zeros((20, 1152), np.float32), 8)
    assert tok.shape[0] == 1 and not cut.any()              # all sims tie at 0 -> one token


@pytest.mark.parametrize("name", gc.E2E_NAMES)
def test_encode_matches_reference_module(golden_dir, name):
    g = load(golden_dir, f"e2e_{name}.npz")
    case = gc.e2e_case(name)
    assert gc.digest(case["feats"], *case["masks"]) == str(g["sha"])
    dt = case["dtype"]
    weights = tuple(R.round_to(w, dt) for w in synth.make_weights(0))
    out = R.encode(R.round_to(case["feats"], dt), case["masks"], case["ann"], case["k"], dt,
                   weights, pad_square=case["aspect"] == "pad")
    assert out["counts"] == list(g["counts"])
    assert np.abs(out["pooled"] - g["pooled"]).max() <= 1e-5
    tol = 1e-5 if dt == "f32" else 1e-2
    assert out["tokens"].shape == g["tokens"].shape
    assert np.abs(out["tokens"] - g["tokens"]).max() <= tol


# ---------------------------------------------------------------------------------------------
# property tests: the integer restatement of the resize against ATen itself, and the torch-CPU port
# (the timed CPU baseline) against the numpy oracle -- neither needs the reference tree
# ---------------------------------------------------------------------------------------------
from hypothesis import example, given, settings, strategies as st

# input sizes where ATen's fused multiply-add `scale * (i + 0.5) - 0.5` and the separately rounded product
# land on different sides of an integer (n_out = 27): the tap set differs by one tap unless src is evaluated
# with a single rounding, as ATen does
FMA_SENSITIVE_SIZES = (3, 5, 9, 2049)


def aten_axis_taps(n_in, n_out=27):
    """Tap usage of ATen's bilinear resize along one axis, observed through F.interpolate itself: feed the
    n_in one-hot rows and see which outputs each source index reaches with non-zero weight."""
    import torch
    import torch.nn.functional as F
    eye = torch.eye(n_in, dtype=torch.float32)[:, None, :, None].expand(-1, -1, -1, 2).contiguous()   # [n_in, 1, n_in, 2]
    out = F.interpolate(eye, size=(n_out, 2), mode="bilinear", align_corners=False)                    # H resized, W kept
    return (out[:, 0, :, 0] > 0).numpy()            # [source index, output index]


def oracle_axis_taps(n_in, n_out=27):
    i0, i1, u0, u1 = R.axis_taps(n_in, n_out)
    reach = np.zeros((n_in, n_out), dtype=bool)
    o = np.arange(n_out)
    reach[i0[u0], o[u0]] = True
    reach[i1[u1], o[u1]] = True
    return reach


def test_tap_table_equals_aten_for_every_input_size_up_to_2300():
    """Appendix B.1 exhaustively along one axis: every input size 1..2300 plus the video sizes, including the
    four sizes where only the fused evaluation of the source index matches ATen."""
    sizes = list(range(1, 2301)) + [2160, 3840, 4096, 4320, 7680]
    bad = [n for n in sizes if not np.array_equal(oracle_axis_taps(n), aten_axis_taps(n))]
    assert bad == []
    assert all(n in sizes for n in FMA_SENSITIVE_SIZES)


def test_compiled_tap_table_equals_the_oracle_for_every_input_size():
    """ufv_tap_table (host code inside libufv_b200.so, no GPU needed) against the oracle, square and 'pad'."""
    from ufvideo_b200 import packer
    for n in list(range(1, 700)) + [1080, 1920, 2049, 2160, 3840, 4096]:
        t = packer.tap_table(n, n, 27, False).reshape(4, 27)
        i0, i1, u0, u1 = R.axis_taps(n)
        assert np.array_equal(t[0], np.where(u0, i0, -1)) and np.array_equal(t[1], np.where(u1, i1, -1)), n
        assert np.array_equal(t[0], t[2]) and np.array_equal(t[1], t[3]), n
    for h, w in ((3, 5), (5, 3), (9, 2049), (480, 854), (2049, 9)):
        t = packer.tap_table(h, w, 27, True).reshape(4, 27)
        side, top, left = R.pad_to_square_offsets(h, w)
        i0, i1, u0, u1 = R.axis_taps(side)
        for row, idx, use, off, ext in ((0, i0, u0, top, h), (1, i1, u1, top, h), (2, i0, u0, left, w), (3, i1, u1, left, w)):
            want = np.where(use & (idx - off >= 0) & (idx - off < ext), idx - off, -1)
            assert np.array_equal(t[row], want), (h, w, row)


@settings(max_examples=120, deadline=None, derandomize=True)
@given(st.integers(1, 140), st.integers(1, 140), st.floats(0.0, 1.0), st.integers(0, 2 ** 31 - 1), st.booleans())
@example(1, 3, 0.5, 0, False)
@example(3, 3, 0.5, 1, False)
@example(5, 9, 0.5, 2, False)
@example(9, 5, 0.3, 3, True)
@example(3, 2049, 0.5, 4, False)
@example(2049, 2049, 0.01, 5, False)
@example(2049, 5, 0.5, 6, True)
def test_tap_or_equals_aten_bilinear_threshold_for_any_size(h, w, density, seed, pad):
    """layer.py:137-143 on arbitrary sizes (up- and down-sampling, non-square, multiples of 27 where a tap
    weight is exactly zero), with and without the 'pad' mode of layer.py:77-86: interp(mask) > 0 computed by
    ATen equals the OR over the taps with non-zero weight."""
    import torch
    import torch.nn.functional as F
    g = synth.rng_for(seed)
    mask = (g.random((h, w)) < density).astype(np.float32)
    t = torch.from_numpy(mask)[None, None]
    if pad:
        side = max(h, w)
        t = F.pad(t, ((side - w) // 2, (side - w) - (side - w) // 2, (side - h) // 2, (side - h) - (side - h) // 2))
    if t.shape[-2:] != (27, 27):
        t = F.interpolate(t, size=(27, 27), mode="bilinear", align_corners=False)
    want = (t > 0)[0, 0].reshape(-1).numpy()
    assert np.array_equal(R.mask_to_patches(mask, pad_square=pad), want)


@settings(max_examples=12, deadline=None)
@given(st.integers(2, 9), st.integers(1, 3), st.integers(1, 5), st.integers(0, 10_000))
def test_cpu_baseline_port_agrees_with_the_oracle(frames, n_obj, k, seed):
    """oracle/reference_port.py (ATen ops, what bench.py times as the CPU baseline) and
    oracle/restatement.py (numpy, the parity oracle) are two restatements of layer.py:63-128: same counts,
    tokens within the fp32 bar."""
    import torch
    from oracle import reference_port
    feats, masks, ann = synth.make_clip(seed, frames, n_obj, "blob", 48, 60, c=64, n_patch=729, ragged=True)
    g = synth.rng_for(seed + 1)
    w = [g.uniform(-0.1, 0.1, s).astype(np.float32) for s in ((32, 64), (32,), (32, 32), (32,))]
    o = R.encode(feats, [masks], [ann], k, "f32", w)
    with torch.no_grad():
        tok, counts = reference_port.encode(torch.from_numpy(feats), [torch.from_numpy(masks).float()], [ann], k,
                                            *[torch.from_numpy(a) for a in w])
    assert counts == o["counts"]
    assert np.abs(tok.numpy() - o["tokens"]).max() <= 1e-5
