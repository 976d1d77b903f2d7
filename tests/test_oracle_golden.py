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
