"""Parity of the CUDA path (through the C ABI) against the oracle and the reference's golden
vectors.  Needs a B200: every test is marked gpu.

Bars (BASELINE.json north_star): patch indices and merge decisions bit-exact; pooled / merged
tokens within 1e-5 in fp32; projected tokens within 1e-2 in bf16 / fp16, 1e-5 in fp32.  Against
the oracle's canonical evaluation order the pooled rows, similarities, cuts and merged tokens are
additionally required to be BIT-IDENTICAL.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import golden_cases as gc
from oracle import restatement as R
from ufvideo_b200 import layer, packer, synth
from ufvideo_b200.layer import MaskExtractor, MaskPooling, build_region_encoder, token_merge

pytestmark = pytest.mark.gpu

TORCH_DT = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}
TOL = {"f32": 1e-5, "bf16": 1e-2, "f16": 1e-2}


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch.device("cuda:0")


def cfg(c=1152, hid=3584):
    import types
    return types.SimpleNamespace(mm_hidden_size=c, hidden_size=hid)


def make_encoder(dev, dtype="f32", k=8, aspect="square", weights=None, c=1152, hid=3584):
    enc = build_region_encoder(cfg(c, hid), aspect)
    enc.region_token_num = k
    enc.requires_grad_(False)                 # forward-only path
    w = weights if weights is not None else synth.make_weights(0, c, hid)
    with torch.no_grad():
        for p, a in zip((enc.feat_linear[0].weight, enc.feat_linear[0].bias,
                         enc.feat_linear[2].weight, enc.feat_linear[2].bias), w):
            p.copy_(torch.from_numpy(a))
    enc.keep_debug = True
    return enc.to(dev).to(TORCH_DT[dtype])


def bits_to_bool(bits: torch.Tensor, n=729):
    b = bits.cpu().numpy().view(np.uint32)
    return ((b[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool).reshape(b.shape[0], -1)[:, :n]


# ---------------------------------------------------------------------------------------------
# kernel 1
# ---------------------------------------------------------------------------------------------
@pytest.fixture(params=["rows", "taps"])
def read_mode(request):
    """Kernel 1 reads a mask in row mode or tap mode (packer.READ_MODE 'auto' picks by residence)."""
    packer.READ_MODE = request.param
    yield request.param
    packer.READ_MODE = "auto"


@pytest.mark.parametrize("mask_dtype", [torch.uint8, torch.float32, torch.bool, torch.float16, torch.bfloat16])
def test_patch_bits_match_reference_golden(dev, golden_dir, mask_dtype, read_mode):
    g = np.load(os.path.join(golden_dir, "resize.npz"))
    meta = json.loads(str(g["meta"]))
    row = 0
    for m in meta:
        masks = gc.resize_masks(m["h"], m["w"])
        t = torch.from_numpy(masks).to(dev).to(mask_dtype)
        plan = packer.build_plan([t], [[list(range(m["n"]))]], m["n"], 1, dev)
        out = layer.mask_to_patches(plan, dev, 27, want_idx=True)
        bits, cnt, idx = out["bits"], out["cnt"], out["idx"]
        got = bits.cpu().numpy().view(np.uint32)[:, :23]
        want = g["bits"][row:row + m["n"]]
        assert np.array_equal(got, want), (m["h"], m["w"], mask_dtype)
        on = bits_to_bool(bits)
        assert np.array_equal(cnt.cpu().numpy(), on.sum(1))
        idx_np = idx.cpu().numpy()
        for j in range(m["n"]):
            assert np.array_equal(idx_np[j, :on[j].sum()], np.flatnonzero(on[j]))
        row += m["n"]


@pytest.mark.parametrize("pad", [False, True])
def test_patch_bits_equal_cuda_aten_interpolate_threshold(dev, pad):
    """Kernel 1 against ATen's own CUDA bilinear kernel run on this GPU (layer.py:137-143, with the 'pad'
    mode of layer.py:77-86): every input size 1..60 (up-sampling, the multiples of 27 where a tap weight is
    exactly zero, and 3 / 5 / 9 where only the fused evaluation of the source index matches ATen), 81, 2049
    (the fourth such size) and the 4K video frame sizes -- square and paired with another size."""
    import torch.nn.functional as F
    sizes = list(range(1, 61)) + [81, 2049, 2160, 3840]
    shapes = [(n, n) for n in sizes] + [(n, sizes[(7 * i + 3) % len(sizes)]) for i, n in enumerate(sizes)]
    for h, w in shapes:
        g = torch.Generator(device="cpu").manual_seed(h * 4099 + w)
        n = 3 if h * w <= 2049 * 2049 else 2
        dens = torch.tensor([0.5, 0.02, 0.3])[:n, None, None]
        masks = (torch.rand((n, h, w), generator=g) < dens).float().to(dev)
        masks[0, 0, 0] = 1.0
        masks[-1, h - 1, w - 1] = 1.0
        ref = masks[None]
        if pad:
            side = max(h, w)
            ref = F.pad(ref, ((side - w) // 2, (side - w) - (side - w) // 2, (side - h) // 2, (side - h) - (side - h) // 2))
        if ref.shape[-2:] != (27, 27):
            ref = F.interpolate(ref, size=(27, 27), mode="bilinear", align_corners=False)
        want = (ref > 0)[0].reshape(n, -1).cpu().numpy()
        plan = packer.build_plan([masks], [[list(range(n))]], n, 1, dev, pad_square=pad, use_cache=False)
        out = layer.mask_to_patches(plan, dev)
        assert np.array_equal(bits_to_bool(out["bits"]), want), (h, w, pad)
        assert np.array_equal(out["cnt"].cpu().numpy(), want.sum(1)), (h, w, pad)


def test_patch_bits_pad_mode_and_strided_masks(dev):
    masks = synth.masks_blob(3, 2, 3, 480, 854)
    want = np.stack([R.mask_to_patches(m, pad_square=True) for m in masks])
    plan = packer.build_plan([torch.from_numpy(masks).to(dev)], [[list(range(6))]], 6, 1, dev, pad_square=True)
    bits = layer.mask_to_patches(plan, dev)["bits"]
    assert np.array_equal(bits_to_bool(bits), want)
    # a non-contiguous view (row pitch != W) is read in place
    big = torch.zeros((6, 480, 1000), dtype=torch.uint8, device=dev)
    big[:, :, 100:954] = torch.from_numpy(masks).to(dev)
    view = big[:, :, 100:954]
    plan = packer.build_plan([view], [[list(range(6))]], 6, 1, dev)
    bits = layer.mask_to_patches(plan, dev)["bits"]
    assert np.array_equal(bits_to_bool(bits), np.stack([R.mask_to_patches(m) for m in masks]))


@pytest.mark.parametrize("hw", [(384, 384), (480, 854), (720, 1280), (100, 37), (61, 509)])
@pytest.mark.parametrize("mask_dtype", [torch.float32, torch.uint8, torch.float16])
def test_pinned_host_masks_are_read_in_place(dev, hw, mask_dtype, read_mode):
    """A pinned host mask tensor is not copied: kernel 1 reads it over PCIe through its mapped
    address (row mode for narrow masks, tap mode for wide ones) and gives the same bits."""
    h, w = hw
    masks = synth.masks_sparse(h * 7 + w, 5, h, w, p=0.01)
    host = torch.from_numpy(masks).to(mask_dtype).pin_memory()
    # an odd element offset: rows of the view are not 16-byte aligned
    wide = torch.zeros((5, h, w + 7), dtype=mask_dtype).pin_memory()
    wide[:, :, 3:3 + w] = host
    want = np.stack([R.mask_to_patches(m) for m in masks])
    for t in (host, wide[:, :, 3:3 + w]):
        plan = packer.build_plan([t], [[list(range(5))]], 5, 1, dev, use_cache=False)
        assert plan.host["mask_desc"]["addr"][0] != 0
        bits = layer.mask_to_patches(plan, dev)["bits"]
        assert np.array_equal(bits_to_bool(bits), want)


@pytest.mark.parametrize("hw", [(384, 384), (480, 854), (100, 37), (27, 27), (13, 13)])
@pytest.mark.parametrize("pad", [False, True])
def test_run_length_masks_give_the_same_bits_as_dense_masks(dev, hw, pad):
    """COCO RLE dicts go to kernel 1 as cumulative run ends (one binary search per tap): same patch bits
    as the dense mask the reference would have decoded with pycocotools (mm_utils.py:22-33)."""
    from ufvideo_b200 import rle
    h, w = hw
    masks = np.concatenate([synth.masks_blob(h + w, 2, 2, h, w), synth.masks_sparse(h * w, 3, h, w, p=0.01),
                            np.zeros((1, h, w), np.uint8), np.ones((1, h, w), np.uint8)])
    want = np.stack([R.mask_to_patches(R.rle_to_mask(rle.encode(m)), pad_square=pad) for m in masks])
    assert np.array_equal(want, np.stack([R.mask_to_patches(m, pad_square=pad) for m in masks]))
    n = masks.shape[0]
    for sample in ([rle.encode(m) for m in masks],
                   [{"size": [h, w], "counts": rle.counts_to_string(rle.encode(m)["counts"])} for m in masks]):
        plan = packer.build_plan([sample], [[list(range(n))]], n, 1, dev, pad_square=pad, use_cache=False)
        out = layer.mask_to_patches(plan, dev)
        assert np.array_equal(bits_to_bool(out["bits"]), want)
        assert np.array_equal(out["cnt"].cpu().numpy(), want.sum(1))


def test_forward_with_run_length_masks_equals_forward_with_dense_masks(dev):
    from ufvideo_b200 import rle
    case = gc.e2e_case("multi")
    enc = make_encoder(dev, "f32", case["k"])
    feats = torch.from_numpy(case["feats"]).to(dev)
    dense = [torch.from_numpy(m).to(dev) for m in case["masks"]]
    a, na = enc(feats, dense, None, case["ann"], None)
    for _ in range(2):                                       # second call: cached plan, refreshed run data
        b, nb = enc(feats, [[rle.encode(m) for m in ms] for ms in case["masks"]], None, case["ann"], None)
        assert na == nb and torch.equal(a, b)


# ---------------------------------------------------------------------------------------------
# kernel 2
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", [c[0] for c in gc.POOL_CASES])
@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16"])
def test_pool_bit_exact_vs_oracle_and_close_to_reference(dev, golden_dir, name, dtype):
    g = np.load(os.path.join(golden_dir, "pool.npz"))
    feats, masks, rows = gc.pool_inputs(name)
    feats_r = R.round_to(feats, dtype)
    ft = torch.from_numpy(feats_r).to(dev).to(TORCH_DT[dtype])
    n_obj = len(rows) // feats.shape[0]
    ann = [[rows[o * feats.shape[0]:(o + 1) * feats.shape[0]].tolist() for o in range(n_obj)]]
    plan = packer.build_plan([torch.from_numpy(masks).to(dev)], ann, feats.shape[0], 1, dev)
    patches = layer.mask_to_patches(plan, dev)
    pooled = layer.mask_pool(ft, plan, patches).cpu().numpy()
    on = np.stack([R.mask_to_patches(m) for m in masks])
    want = R.mask_pool(feats_r, rows, on)
    assert np.array_equal(pooled, want), np.abs(pooled - want).max()      # canonical order: bit-exact
    if dtype == "f32":
        assert np.abs(pooled - g[name]).max() <= 1e-5                     # vs the real reference
    assert (pooled[~on.any(1)] == 0).all()


@pytest.mark.parametrize("family", ["dense", "blob", "sparse"])
def test_pool_windows_tile_and_row_paths_agree(dev, family):
    """The pool kernel walks a frame in windows of 32 patches and fetches a window either as one 2-D tile or
    row by row, depending on how many of its rows are needed: masks whose windows fall on both sides of the
    threshold (a band of rows on, half-filled windows, empty windows, the ragged last window of 25 patches)
    must give the oracle's bits on either path."""
    feats = R.round_to(synth.features(91, 3), "bf16")
    base = synth.make_masks(family, 92, 4, 3, 54, 54)          # [12, 54, 54]: objects 0..3 on frames 0..2
    masks = base.copy()
    masks[0, :, :] = 0
    masks[0, 10:30, :] = 1                                      # a band: full windows in the middle, empty elsewhere
    masks[1, ::2, :] = 0                                        # every other source row off
    masks[5] = 1                                                # everything on (last window: 25 rows)
    masks[7] = 0                                                # nothing on
    rows = [f for o in range(4) for f in range(3)]
    ann = [[[0, 1, 2] for _ in range(4)]]
    plan = packer.build_plan([torch.from_numpy(masks).to(dev)], ann, 3, 1, dev)
    patches = layer.mask_to_patches(plan, dev)
    pooled = layer.mask_pool(torch.from_numpy(feats).to(dev).bfloat16(), plan, patches).cpu().numpy()
    on = np.stack([R.mask_to_patches(m) for m in masks])
    assert np.array_equal(pooled, R.mask_pool(feats, rows, on))


def test_mask_pooling_module_matches_reference_signature(dev, golden_dir):
    """MaskPooling.forward(x NCHW-view, mask [1,q,H,W]) exactly as layer.py:101,108 calls it."""
    g = np.load(os.path.join(golden_dir, "pool.npz"))
    feats, masks, rows = gc.pool_inputs("dense384")
    x = torch.from_numpy(feats).to(dev)[torch.from_numpy(rows).to(dev)]
    x = x.reshape(x.shape[0], 27, 27, -1).permute(0, 3, 1, 2)
    out = MaskPooling()(x, torch.from_numpy(masks).to(dev).float().unsqueeze(0))
    assert np.abs(out.cpu().numpy() - g["dense384"]).max() <= 1e-5


@pytest.mark.parametrize("split", [16, 64])
@pytest.mark.parametrize("n_obj", [5, 9, 17, 33, 64, 70])
@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16"])
def test_pool_many_objects_on_one_frame(dev, n_obj, split, dtype, monkeypatch):
    """Many objects on one frame (PixRQA broadcast shape): beyond 8 members the bit-iterating consumers run
    (2 / 4 / 8 members per warp, 256-channel slices for 16-bit features up to 32 members); the packer cuts a
    frame into equal sub-groups of at most GROUP_SPLIT members (16 by default; 64 = the kernel's limit, which
    keeps the 4- and 8-members-per-warp variants covered).  A second frame with fewer members rides in the same call."""
    monkeypatch.setattr(packer, "GROUP_SPLIT", split)
    groups = -(-n_obj // split)
    feats = R.round_to(synth.features(77, 2), dtype)
    masks = np.concatenate([synth.masks_blob(78, n_obj, 1, 100, 120), synth.masks_sparse(79, 3, 100, 120)])
    rows = [0] * n_obj + [1] * 3
    ann = [[[r] for r in rows]]
    plan = packer.build_plan([torch.from_numpy(masks).to(dev)], ann, 2, 4, dev, use_cache=False)
    assert plan.n_groups == groups + 1 and plan.max_group == max(-(-n_obj // groups), 3)
    patches = layer.mask_to_patches(plan, dev)
    pooled = layer.mask_pool(torch.from_numpy(feats).to(dev).to(TORCH_DT[dtype]), plan, patches).cpu().numpy()
    on = np.stack([R.mask_to_patches(m) for m in masks])
    assert np.array_equal(pooled, R.mask_pool(feats, rows, on))


# ---------------------------------------------------------------------------------------------
# kernel 3
# ---------------------------------------------------------------------------------------------
def run_ttm(dev, x, k):
    t = x.shape[0]
    host = {"obj_start": np.zeros(1, np.int32), "obj_len": np.full(1, t, np.int32),
            "slot_off": np.zeros(1, np.int32)}
    plan = packer.EncodePlan(n_masks=t, n_groups=0, max_group=1, n_obj=1, max_len=t, m_pad=min(t, k),
                             slots=np.full(1, min(t, k), np.int32), host=host)
    packer._upload(plan, dev)
    tok, counts, ex = layer.ttm(torch.from_numpy(x).to(dev), plan, k, torch.float32, debug=True)
    n = int(counts.item())
    cut = None
    if t > k:
        words = ex["cuts"].cpu().numpy().view(np.uint32)[0]
        cut = ((words[:, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool).reshape(-1)[: t - 1]
    return ex["tokens_f32"].cpu().numpy(), n, cut, ex["sims"].cpu().numpy()[0, : t - 1]


def test_ttm_decisions_bit_exact_vs_reference_golden(dev, golden_dir):
    g = np.load(os.path.join(golden_dir, "ttm.npz"))
    meta = json.loads(str(g["meta"]))
    for idx, m in enumerate(meta):
        x = gc.ttm_tokens(m["family"], m["t"], m["seed"])
        tok, n, cut, sims = run_ttm(dev, x, m["k"])
        ref_cut = np.unpackbits(g[f"case{idx}_cut"])[: m["t"] - 1].astype(bool)
        assert np.array_equal(cut, ref_cut), m                        # the reference's own decisions
        assert n == m["rows"]
        o_tok, o_cut, o_sims = R.token_merge(x, m["k"])
        assert np.array_equal(sims, o_sims), m                        # canonical order: bit-exact
        assert np.array_equal(tok[:n], o_tok), m
        ref_tok = g[f"case{idx}_merged"]
        mine = tok[:n] if m["full"] else tok[:n, :64]
        assert np.abs(mine - ref_tok).max() <= 1e-5, m


@pytest.mark.parametrize("family", ["static", "walk", "duplicates"])
def test_ttm_bit_exact_vs_oracle_on_coherent_tokens(dev, family):
    """Near-tie regimes where ATen's own decisions depend on its reduction order:
    the CUDA kernel and the oracle share one canonical order, so they still agree bit-for-bit."""
    g = synth.rng_for(5150)
    for t, k in [(16, 8), (32, 4), (256, 8), (300, 1)]:
        base = g.standard_normal(1152, dtype=np.float32)
        if family == "static":
            x = base + 1e-3 * g.standard_normal((t, 1152), dtype=np.float32)
        elif family == "walk":
            x = base + np.cumsum(0.05 * g.standard_normal((t, 1152), dtype=np.float32), 0)
        else:
            x = np.repeat(g.standard_normal((t // 4 + 1, 1152), dtype=np.float32), 4, 0)[:t]
        x = x.astype(np.float32)
        tok, n, cut, sims = run_ttm(dev, x, k)
        o_tok, o_cut, o_sims = R.token_merge(x, k)
        assert np.array_equal(sims, o_sims) and np.array_equal(cut, o_cut)
        assert n == o_tok.shape[0] and np.array_equal(tok[:n], o_tok)


@pytest.mark.parametrize("t,k", [(16, 8), (33, 4), (40, 8), (100, 8)])
def test_ttm_quotient_paths_bit_exact_on_mixed_magnitudes(dev, t, k):
    """The merge kernel divides by the row norm with the reciprocal hoisted where that is provably the IEEE quotient
    and with the IEEE division elsewhere (ttm.cu, tools/div_probe.cu).  Rows that hold a zero, a value below 2^-60,
    above 2^41, or whose elements span the whole admitted range sit next to ordinary rows, so adjacent pairs take
    every combination of the two paths; similarities, cuts and merged tokens equal the oracle's bit for bit.
    t = 16 / 33: the one-warp threshold phase (T <= 33); 40: the block-wide one; 100: the split kernels."""
    g = synth.rng_for(977 + t)
    x = g.standard_normal((t, 1152), dtype=np.float32)
    x[1, 7] = 0.0                                             # a zero: IEEE path
    x[2] *= np.float32(2.0 ** -70)                            # whole row below the fast range
    x[3, 100] = np.float32(1e-25)                             # one tiny element
    x[5] *= np.float32(2.0 ** 43)                             # whole row above the fast range
    x[6, 5] = np.float32(3e12)                                # one element just above 2^41
    x[7] *= np.float32(2.0 ** 30)                             # large but inside: fast path
    x[8] *= np.float32(2.0 ** -45)                            # small but inside: fast path
    x[9, ::2] *= np.float32(2.0 ** -55)                       # wide dynamic range inside: fast path
    x[9, 1::2] *= np.float32(2.0 ** 35)
    x[10] = 0.0                                               # all-zero row: norm clamps to 1e-12
    x[12] = x[11]                                             # an exact duplicate: similarity near 1
    tok, n, cut, sims = run_ttm(dev, x, k)
    o_tok, o_cut, o_sims = R.token_merge(x, k)
    assert np.array_equal(sims.view(np.uint32), o_sims.view(np.uint32))
    assert np.array_equal(cut, o_cut) and n == o_tok.shape[0]
    assert np.array_equal(tok[:n].view(np.uint32), o_tok.view(np.uint32))


def test_token_merge_function_matches_reference_signature(dev):
    x = gc.ttm_tokens("random", 16, 40003)
    out = token_merge(torch.from_numpy(x).to(dev)[None], 8)            # r = tokens to remove
    want, _, _ = R.token_merge(x, 8)
    assert out.shape == (1, 8, 1152) and np.array_equal(out[0].cpu().numpy(), want)
    zeros = token_merge(torch.zeros((1, 20, 1152), device=dev), 12)    # all sims tie -> one token
    assert zeros.shape == (1, 1, 1152)


# ---------------------------------------------------------------------------------------------
# kernel 4
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["bf16", "f16", "f32"])
@pytest.mark.parametrize("m", [1, 8, 127, 128, 200, 256, 300])
def test_linear_matches_torch(dev, dtype, m):
    g = synth.rng_for(m * 7 + 1)
    dt = TORCH_DT[dtype]
    w1, b1, w2, b2 = [torch.from_numpy(a).to(dev).to(dt) for a in synth.make_weights(3)]
    x = torch.from_numpy(g.standard_normal((m, 1152), dtype=np.float32) * 0.05).to(dev).to(dt)
    h = layer.linear(x, w1, b1, gelu=True)
    y = layer.linear(h, w2, b2, gelu=False)
    ref_h = torch.nn.functional.gelu(torch.nn.functional.linear(x.float(), w1.float(), b1.float()).to(dt).float()).to(dt)
    ref_y = torch.nn.functional.linear(h.float(), w2.float(), b2.float())
    torch.cuda.synchronize()
    assert (h.float() - ref_h.float()).abs().max().item() <= TOL[dtype]
    assert (y.float() - ref_y).abs().max().item() <= TOL[dtype]


@pytest.mark.parametrize("split,bn", [(2, 256), (4, 256), (8, 256), (3, 128), (5, 256), (2, 64), (4, 128)])
def test_linear_split_k_cluster_variants(dev, split, bn):
    """Every instantiation of the cluster split-K kernel (forced through the developer knobs), both layer
    shapes, with and without GELU, ragged M: equal to the full-K kernel within the 16-bit tolerance, equal to
    torch fp32, and bit-identical from call to call and for any number of tokens in the call (the reduce
    order is fixed: partials are added in k-range order whatever CTA finishes first)."""
    os.environ["UFV_GEMM_SPLIT"], os.environ["UFV_GEMM_SPLIT_BN"] = str(split), str(bn)
    try:
        for m, n, k in ((256, 3584, 3584), (77, 3584, 3584), (200, 3584, 1152), (128, 512, 1152)):
            x = (torch.randn((m, k), device=dev) * 0.05).bfloat16()
            w = (torch.randn((n, k), device=dev) * 0.03).bfloat16()
            b = (torch.randn((n,), device=dev) * 0.03).bfloat16()
            for gelu in (False, True):
                os.environ["UFV_GEMM_SPLIT"] = str(split)
                from ufvideo_b200 import _cabi
                expect_split = (-(-m // 128)) * (-(-n // bn)) * split <= 148
                assert (_cabi.lib().ufv_linear_ws_bytes(m, n, k, _cabi.UFV_BF16) > 0) == expect_split
                y = layer.linear(x, w, b, gelu=gelu)
                again = layer.linear(x, w, b, gelu=gelu)
                part = layer.linear(x[: max(m // 3, 1)], w, b, gelu=gelu)
                os.environ["UFV_GEMM_SPLIT"] = "0"
                full = layer.linear(x, w, b, gelu=gelu)
                ref = torch.nn.functional.linear(x.float(), w.float(), b.float())
                if gelu:
                    ref = torch.nn.functional.gelu(ref.bfloat16().float())
                assert torch.equal(y, again), (m, n, k, gelu)
                part_split = (-(-part.shape[0] // 128)) * (-(-n // bn)) * split <= 148
                if part_split == expect_split:        # same kernel for both calls: same bits for the shared rows
                    assert torch.equal(part, y[: part.shape[0]]), (m, n, k, gelu)
                assert (y.float() - ref).abs().max().item() <= 1e-2, (m, n, k, gelu)
                assert (y.float() - full.float()).abs().max().item() <= 1e-2, (m, n, k, gelu)
    finally:
        os.environ.pop("UFV_GEMM_SPLIT", None)
        os.environ.pop("UFV_GEMM_SPLIT_BN", None)


@pytest.mark.parametrize("bn", [0, 32, 64, 128, 256])
def test_linear_large_m_all_tile_shapes(dev, bn):
    """Every N-tile instantiation of the persistent tcgen05 kernel (0 = the cost model's choice),
    with several tiles per CTA, a ragged M edge and a ragged N edge (n % 32 != 0)."""
    if bn:
        os.environ["UFV_GEMM_BN"] = str(bn)
    try:
        for m, n, k in ((1024, 3584, 1152), (4096 + 77, 3584, 1152), (333, 3584, 3584), (2500, 424, 1152)):
            x = (torch.randn((m, k), device=dev) * 0.05).bfloat16()
            w = (torch.randn((n, k), device=dev) * 0.03).bfloat16()
            b = (torch.randn((n,), device=dev) * 0.03).bfloat16()
            for gelu in (False, True):
                y = layer.linear(x, w, b, gelu=gelu)
                ref = torch.nn.functional.linear(x.float(), w.float(), b.float())
                if gelu:
                    ref = torch.nn.functional.gelu(ref.bfloat16().float())
                assert (y.float() - ref).abs().max().item() <= 1e-2, (bn, m, n, k, gelu)
    finally:
        os.environ.pop("UFV_GEMM_BN", None)


# ---------------------------------------------------------------------------------------------
# whole path
# ---------------------------------------------------------------------------------------------
def to_dev_masks(case, dev):
    masks = [torch.from_numpy(m).to(dev).float() for m in case["masks"]]
    return torch.stack(masks) if case["masks_as_tensor"] else masks


@pytest.mark.parametrize("name", gc.E2E_NAMES)
def test_forward_matches_reference_module_golden(dev, golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"e2e_{name}.npz"))
    case = gc.e2e_case(name)
    dt = case["dtype"]
    enc = make_encoder(dev, dt, case["k"], case["aspect"])
    feats = torch.from_numpy(case["feats"]).to(dev).to(TORCH_DT[dt])
    tokens, counts = enc(feats, to_dev_masks(case, dev), feats, case["ann"], None)
    assert counts == list(g["counts"])                                  # region_token_nums: exact
    assert tokens.dtype == TORCH_DT[dt] and tuple(tokens.shape) == g["tokens"].shape
    assert np.abs(enc._debug["pooled"].cpu().numpy() - g["pooled"]).max() <= 1e-5
    assert np.abs(tokens.float().cpu().numpy() - g["tokens"]).max() <= TOL[dt]
    # and bit-exact against the oracle up to the projector input
    w = tuple(R.round_to(a, dt) for a in synth.make_weights(0))
    o = R.encode(R.round_to(case["feats"], dt), case["masks"], case["ann"], case["k"], dt, w,
                 pad_square=case["aspect"] == "pad")
    assert np.array_equal(enc._debug["pooled"].cpu().numpy(), o["pooled"])
    if counts == list(enc.last_plan.slots):
        assert np.array_equal(enc._debug["merged"].float().cpu().numpy(), o["merged"])
    assert np.abs(tokens.float().cpu().numpy() - o["tokens"]).max() <= TOL[dt]


def test_forward_with_pinned_host_inputs_equals_device_inputs(dev):
    """The e2e path: pinned host feats (copied) and pinned host masks (read in place)."""
    case = gc.e2e_case("bf16")
    enc = make_encoder(dev, "bf16", 8)
    feats = torch.from_numpy(case["feats"]).bfloat16()
    masks = [torch.from_numpy(m).float() for m in case["masks"]]
    a, na = enc(feats.to(dev), [m.to(dev) for m in masks], None, case["ann"], None)
    b, nb = enc(feats.pin_memory(), [m.pin_memory() for m in masks], None, case["ann"], None)
    assert na == nb and torch.equal(a, b)


@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16"])
def test_repeated_calls_replay_and_stay_identical(dev, dtype):
    """From the second call of a batch structure on, bf16 / fp16 forward() replays a captured CUDA graph
    (per-call values travel through the pinned block); results must not change, also with new mask data."""
    case = gc.e2e_case("bf16")
    enc = make_encoder(dev, dtype, 8)
    feats = torch.from_numpy(case["feats"]).to(dev).to(TORCH_DT[dtype])
    masks = [torch.from_numpy(m).to(dev) for m in case["masks"]]
    first, n_first = enc(feats, masks, None, case["ann"], None)
    for _ in range(4):
        again, n_again = enc(feats, masks, None, case["ann"], None)
        assert n_again == n_first and torch.equal(again, first) and again.data_ptr() != first.data_ptr()
    if dtype != "f32":
        assert enc.last_plan.run["graphs"], "the graph should have been captured by now"
    flipped = [m.flip(0).contiguous() for m in masks]        # other mask content, same structure, new pointers
    ref = make_encoder(dev, dtype, 8)
    want, n_want = ref(feats, flipped, None, case["ann"], None)
    got, n_got = enc(feats, flipped, None, case["ann"], None)
    assert n_got == n_want and torch.equal(got, want)


def test_pageable_host_masks_are_recopied_on_every_call(dev):
    """Masks in pageable host memory are copied by the packer: the copy must outlive the launches that read
    it, and a second call with the SAME tensor objects but edited content must see the new content (the
    identity fast path may not skip the copy).  Same for the zero mask that replaces an empty sample."""
    case = gc.e2e_case("bf16")
    enc = make_encoder(dev, "bf16", 8)
    ref = make_encoder(dev, "bf16", 8)
    feats = torch.from_numpy(case["feats"]).to(dev).bfloat16()
    host = [torch.from_numpy(m.copy()).float() for m in case["masks"]]          # pageable CPU tensors
    assert not host[0].is_pinned()
    for round_ in range(3):
        want, n_want = ref(feats, [m.to(dev) for m in host], None, case["ann"], None)
        junk = [torch.randn(1 << 20, device=dev) for _ in range(4)]                # churn the caching allocator
        got, n_got = enc(feats, host, None, case["ann"], None)
        del junk
        assert n_got == n_want and torch.equal(got, want), round_
        for m in host:                                                           # in-place edit, same objects
            m.copy_(m.flip(1) if round_ == 0 else m.roll(17, 2))
    # strided (non-contiguous) device masks are copied too
    wide = torch.zeros((host[0].shape[0], host[0].shape[1], host[0].shape[2] * 2), device=dev)
    wide[:, :, ::2] = host[0].to(dev)
    a, na = enc(feats, [wide[:, :, ::2]] + [m.to(dev) for m in host[1:]], None, case["ann"], None)
    b, nb = ref(feats, [m.to(dev) for m in host], None, case["ann"], None)
    assert na == nb and torch.equal(a, b)


def test_empty_sample_twice_and_next_to_a_real_sample(dev):
    """layer.py:73-75 on repeated calls: the substituted zero mask is process-owned, never a dangling copy."""
    feats_np = synth.features(600, 4)
    feats = torch.from_numpy(feats_np).to(dev)
    real = synth.masks_blob(5, 1, 2, 64, 64)
    enc = make_encoder(dev, "f32", 8)
    empty = torch.zeros((0, 336, 336), dtype=torch.uint8)
    outs = []
    for _ in range(3):
        junk = torch.randn(1 << 20, device=dev)
        t, c = enc(feats, [empty, torch.from_numpy(real).to(dev)], None, [[[1]], [[2, 3]]], None)
        del junk
        outs.append((t.clone(), c))
    o = R.encode(feats_np, [np.zeros((0, 336, 336), np.uint8), real], [[[1]], [[2, 3]]], 8, "f32", synth.make_weights(0))
    for t, c in outs:
        assert c == o["counts"] and np.abs(t.cpu().numpy() - o["tokens"]).max() <= 1e-5
        assert torch.equal(t, outs[0][0])


def test_two_modules_and_two_streams_share_a_plan_without_interfering(dev):
    """Run state is per (module, stream): a second module or stream hitting the same cached plan neither
    drops the first one's captured graph nor shares its pinned counts words."""
    case = gc.e2e_case("bf16")
    a, b = make_encoder(dev, "bf16", 8), make_encoder(dev, "bf16", 8)
    feats = torch.from_numpy(case["feats"]).to(dev).bfloat16()
    masks = [torch.from_numpy(m).to(dev) for m in case["masks"]]
    first, n_first = a(feats, masks, None, case["ann"], None)
    side = torch.cuda.Stream(dev)
    for _ in range(3):
        x, nx = a(feats, masks, None, case["ann"], None)
        y, ny = b(feats, masks, None, case["ann"], None)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            z, nz = a(feats, masks, None, case["ann"], None)
        torch.cuda.current_stream(dev).wait_stream(side)
        assert nx == ny == nz == n_first
        assert torch.equal(x, first) and torch.equal(y, first) and torch.equal(z, first)
    plan = a.last_plan
    mine = [r for key, r in plan.runs.items() if key[0] in (id(a), id(b))]      # the cached plan may also carry runs of
    assert plan is b.last_plan and len(mine) == 3                              # modules from earlier tests
    assert sum(bool(r["graphs"]) for r in mine) >= 2


def test_list_and_tensor_mask_forms_agree(dev):
    case = gc.e2e_case("c1")
    enc = make_encoder(dev, "f32", 8)
    feats = torch.from_numpy(case["feats"]).to(dev)
    m = torch.from_numpy(case["masks"][0]).to(dev)
    a, na = enc(feats, [m], feats, case["ann"], None)
    b, nb = enc(feats, m.float().unsqueeze(0), feats, case["ann"], None)
    assert na == nb and torch.equal(a, b)


def test_ties_compact_the_padded_rows(dev):
    """Identical frames -> every similarity ties -> one token per object, rows compacted."""
    feats = synth.features(9, 1).repeat(6, 0)
    masks = np.ones((6, 27, 27), np.uint8)
    enc = make_encoder(dev, "f32", 4)
    tokens, counts = enc(torch.from_numpy(feats).to(dev), [torch.from_numpy(masks).to(dev)],
                         None, [[[0, 1, 2, 3, 4, 5]]], None)
    assert counts == [1] and tokens.shape[0] == 1


def test_ties_in_the_middle_keep_object_order(dev):
    """Object 1 ties down to one token between two ordinary objects: compaction keeps row order."""
    feats = synth.features(11, 8)
    feats[3:6] = feats[3]                                   # frames 3..5 identical
    masks = np.concatenate([synth.masks_blob(12, 1, 8, 60, 60), np.ones((3, 60, 60), np.uint8),
                            synth.masks_blob(13, 1, 6, 60, 60)])
    ann = [[list(range(8)), [3, 4, 5], [0, 1, 2, 5, 6, 7]]]
    enc = make_encoder(dev, "f32", 2)
    tokens, counts = enc(torch.from_numpy(feats).to(dev), [torch.from_numpy(masks).to(dev)], None, ann, None)
    w = synth.make_weights(0)
    o = R.encode(feats, [masks], ann, 2, "f32", w)
    assert counts == o["counts"] == [2, 1, 2]
    assert tuple(tokens.shape) == (5, 3584)
    assert np.abs(tokens.cpu().numpy() - o["tokens"]).max() <= 1e-5


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("m", [0, 40, 256, 700])
def test_linear_gather_stores_tiles_tail_and_flags_to_every_destination(dev, m, split):
    """ufv_linear_gather on one GPU with two local "ranks" as destinations (one st per peer): both
    copies equal ufv_linear's output, the tail is forwarded, flags carry the step value, the ticket
    resets.  (The multimem.st path needs NVSwitch multicast: covered by bench.py's self-check at N > 1.)"""
    import ctypes
    from ufvideo_b200 import _cabi
    n, k, tail_words = 3584, 3584, 37
    x = (torch.randn((max(m, 1), k), device=dev) * 0.05).bfloat16()[:m]
    w = (torch.randn((n, k), device=dev) * 0.03).bfloat16()
    b = (torch.randn((n,), device=dev) * 0.03).bfloat16()
    copies = [torch.zeros((m + 8, n), dtype=torch.bfloat16, device=dev) for _ in range(2)]
    tails = [torch.zeros(tail_words, dtype=torch.int32, device=dev) for _ in range(2)]
    flags = torch.zeros(2, dtype=torch.int32, device=dev)
    tail_src = torch.arange(1, tail_words + 1, dtype=torch.int32, device=dev)
    ticket = torch.zeros(1, dtype=torch.int32, device=dev)
    timed_out = torch.zeros(1, dtype=torch.int32, device=dev)
    a = _cabi.PeerArgs()
    for i in range(2):
        a.dst[i], a.tail_dst[i], a.flag[i] = copies[i].data_ptr(), tails[i].data_ptr(), flags.data_ptr() + 4 * i
    a.tail_src, a.ticket, a.tail_words, a.n_dst, a.multimem, a.flag_value = (
        tail_src.data_ptr(), ticket.data_ptr(), tail_words, 2, 0, 7)
    stream = torch.cuda.current_stream(dev).cuda_stream
    lib = _cabi.lib()
    os.environ["UFV_GEMM_SPLIT_DEFAULT"] = "1" if split else "0"         # with and without the cluster split-K kernel
    ws_bytes = int(lib.ufv_linear_ws_bytes(m, n, k, _cabi.UFV_BF16))
    assert (ws_bytes > 0) == (split and 0 < m <= 512)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    for rep in range(2):                                     # twice: the ticket must have reset itself
        a.flag_value = 7 + rep
        _cabi.check(lib.ufv_linear_gather(x.data_ptr(), w.data_ptr(), b.data_ptr(), m, n, k, _cabi.UFV_BF16,
                                          ctypes.byref(a), ws.data_ptr(), ws_bytes, stream))
        _cabi.check(lib.ufv_wait_flags(flags.data_ptr(), 2, 7 + rep, 500, timed_out.data_ptr(), stream))
        torch.cuda.synchronize()
        assert int(timed_out.item()) == 0 and flags.tolist() == [7 + rep] * 2 and int(ticket.item()) == 0
        want = layer.linear(x, w, b) if m else x.new_zeros((0, n))
        for c, tl in zip(copies, tails):
            assert torch.equal(c[:m], want) and not c[m:].any()
            assert torch.equal(tl, tail_src)
    os.environ.pop("UFV_GEMM_SPLIT_DEFAULT", None)
    _cabi.check(lib.ufv_wait_flags(flags.data_ptr(), 2, 99, 50, timed_out.data_ptr(), stream))   # never arrives
    torch.cuda.synchronize()
    assert int(timed_out.item()) == 1


def test_kernels_write_straight_into_the_gather_payload(dev):
    """forward_padded(out=, counts_out=) + sharding.unpack_padded == forward(), including an object
    that ties down to fewer tokens than its reserved slots."""
    from ufvideo_b200 import sharding
    feats = synth.features(21, 8)
    feats[3:6] = feats[3]
    masks = np.concatenate([synth.masks_blob(22, 1, 8, 60, 60), np.ones((3, 60, 60), np.uint8),
                            synth.masks_blob(23, 1, 6, 60, 60)])
    ann = [[list(range(8)), [3, 4, 5], [0, 1, 2, 5, 6, 7]]]
    enc = make_encoder(dev, "bf16", 2)
    ft = torch.from_numpy(feats).to(dev).bfloat16()
    md = [torch.from_numpy(masks).to(dev)]
    want_tokens, want_counts = enc(ft, md, None, ann, None)
    slots = enc.last_plan.slots
    payload, tok_view, cnt_view = sharding.new_payload(slots, 16, 5, 3584, torch.bfloat16, dev)
    tokens, counts, plan = enc.forward_padded(ft, md, ann, out=tok_view, counts_out=cnt_view)
    assert tokens.data_ptr() == payload.data_ptr() and counts.tolist() == want_counts == [2, 1, 2]
    torch.cuda.synchronize()
    got_tokens, got_counts = sharding.unpack_padded(payload[None], 16, 5)
    assert got_counts == want_counts and torch.equal(got_tokens, want_tokens)


def test_region_splice_matches_the_reference_consumer_loop(dev):
    """splice_regions == the python loop of videorefer_arch.py:300-311 (text pieces and object tokens
    concatenated at the <region> placeholders), with one object tied down to fewer tokens."""
    feats = synth.features(31, 8)
    feats[3:6] = feats[3]
    masks = np.concatenate([synth.masks_blob(32, 1, 8, 60, 60), np.ones((3, 60, 60), np.uint8),
                            synth.masks_blob(33, 1, 6, 60, 60)])
    ann = [[list(range(8)), [3, 4, 5], [0, 1, 2, 5, 6, 7]]]
    enc = make_encoder(dev, "bf16", 2)
    ft = torch.from_numpy(feats).to(dev).bfloat16()
    md = [torch.from_numpy(masks).to(dev)]
    flat_tokens, nums = enc(ft, md, None, ann, None)                       # the reference contract
    assert nums == [2, 1, 2]
    tokens, counts, plan = enc.encode_padded(ft, md, ann)                   # padded rows + device counts
    n_text = 23
    text = torch.randn((n_text, 3584), device=dev).bfloat16()
    pos = [2, 3, 20]                                                        # adjacent placeholders, one near the end
    out, out_len, row_src = layer.splice_regions(text, torch.tensor(pos, dtype=torch.int32, device=dev),
                                                 tokens, counts, plan, want_row_src=True)
    pieces, cur, prev = [], 0, 0
    for p_, n_ in zip(pos, nums):                                           # videorefer_arch.py:300-311
        pieces.append(text[prev:p_])
        pieces.append(flat_tokens[cur:cur + n_])
        cur += n_
        prev = p_ + 1
    pieces.append(text[prev:])
    want = torch.cat(pieces)
    n = int(out_len.item())
    assert n == want.shape[0] == n_text - 3 + sum(nums)
    assert torch.equal(out[:n], want)
    src = row_src[:n].cpu().numpy()
    assert (src[:2] == [0, 1]).all() and src[2] == -1 and src[-1] == n_text - 1


def reference_consumer(text, labels, seq_lens, region_pos, tokens, nums, ignore=-100):
    """The region part of the reference's prepare_inputs_labels_for_multimodal restated with torch cats
    (videorefer_arch.py:291-368): per sample, text pieces and object tokens concatenated at the <region>
    placeholders (a sample without one still advances the object cursor, :263-264), IGNORE labels on token rows,
    then zero / IGNORE / False padding to the longest sample."""
    embeds, labs, off, row, obj = [], [], 0, 0, 0
    for n, pos in zip(seq_lens, region_pos):
        if not pos:
            embeds.append(text[off:off + n])
            labs.append(labels[off:off + n])
            row += nums[obj]
            obj += 1
        else:
            e, l, last = [], [], 0
            for p_ in pos:
                e += [text[off + last:off + p_], tokens[row:row + nums[obj]]]
                l += [labels[off + last:off + p_], torch.full((nums[obj],), ignore, dtype=labels.dtype, device=labels.device)]
                row += nums[obj]
                obj += 1
                last = p_ + 1
            e.append(text[off + last:off + n])
            l.append(labels[off + last:off + n])
            embeds.append(torch.cat(e))
            labs.append(torch.cat(l))
        off += n
    l_max = max(x.shape[0] for x in embeds)
    out = torch.zeros((len(embeds), l_max, text.shape[1]), dtype=text.dtype, device=text.device)
    lab = torch.full((len(embeds), l_max), ignore, dtype=labels.dtype, device=labels.device)
    att = torch.zeros((len(embeds), l_max), dtype=torch.bool, device=text.device)
    for i, (e, l) in enumerate(zip(embeds, labs)):
        out[i, :e.shape[0]], lab[i, :l.shape[0]], att[i, :e.shape[0]] = e, l, True
    return out, lab, att


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
@pytest.mark.parametrize("case", ["ragged", "equal", "ties"])
def test_forward_into_builds_the_padded_batch_like_the_reference_consumer(dev, dtype, case):
    """forward_into (scatter epilogue of the last Linear + ufv_splice_static) against the reference's consumer
    loop: embeddings, labels and attention mask of the padded batch, including a sample without a placeholder,
    adjacent placeholders, a placeholder at position 0, repeated calls with new text (graph replay) and an object
    that ties down to fewer tokens than reserved (slow path)."""
    k = 4
    f0, m0, a0 = synth.make_clip(40, 6, 2, "blob", 64, 64, row0=0)
    f1, m1, a1 = synth.make_clip(41, 1, 1, "blob", 64, 64, row0=6)          # the dummy object of a region-less sample
    f2, m2, a2 = synth.make_clip(42, 7, 3, "dense", 64, 64, row0=7, ragged=case != "equal")
    feats = np.concatenate([f0, f1, f2])
    if case == "ties":
        feats[0:6] = feats[0]                                                # clip 0: identical frames and masks ->
        m0 = np.repeat(m0[:1], m0.shape[0], axis=0)                          # every similarity ties: one token per object
    masks, ann = [m0, m1, m2], [a0, a1, a2]
    if case == "equal":
        seq_lens, region_pos = [11, 17, 8], [[1, 4], [], [0, 3, 7]]
    else:
        seq_lens, region_pos = [9, 5, 12], [[3, 4], [], [0, 5, 11]]
    enc = make_encoder(dev, dtype, k)
    ft = torch.from_numpy(feats).to(dev).to(TORCH_DT[dtype])
    md = [torch.from_numpy(m).to(dev) for m in masks]
    tokens, nums = enc(ft, md, None, ann, None)
    if case == "ties":
        assert nums[:2] == [1, 1]
    for rep in range(3):                                                     # from the 2nd call on: replayed graph
        g = torch.Generator(device="cpu").manual_seed(rep)
        text = torch.randn((sum(seq_lens), 3584), generator=g).to(dev).to(TORCH_DT[dtype])
        labels = torch.randint(0, 1000, (sum(seq_lens),), generator=g).to(dev)
        out, lab, att, got_nums = enc.forward_into(ft, md, ann, text, seq_lens, region_pos, labels=labels)
        want_out, want_lab, want_att = reference_consumer(text, labels, seq_lens, region_pos, tokens, nums)
        assert got_nums == nums
        assert out.shape == want_out.shape and torch.equal(att, want_att) and torch.equal(lab, want_lab)
        assert torch.equal(out, want_out), (case, rep)
    if case == "equal":
        assert att.all()                                                     # same lengths: no padding row anywhere


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_pool_adjoint_kernel_against_dense_torch(dev, dtype):
    """ufv_mask_pool_backward: rows with <= 8 object-frames (register path), a row with 17 (L1 path) and a
    row nobody pools from (zeros), against d_feats[r, p] = sum_j on[j, p] * w[j] computed densely."""
    from ufvideo_b200 import _cabi
    g = synth.rng_for(4242)
    n_rows, n_patch, c = 4, 729, 1152
    frame_of = np.array([0] * 3 + [2] * 17 + [3] * 8, dtype=np.int64)          # row 1: nobody
    q = frame_of.size
    on = g.random((q, n_patch)) < 0.4
    w = g.standard_normal((q, c), dtype=np.float32)
    bits = torch.from_numpy(R.pack_bits(on).view(np.int32)).to(dev)
    bits24 = torch.zeros((q, 24), dtype=torch.int32, device=dev)
    bits24[:, :23] = bits
    order = np.argsort(frame_of, kind="stable").astype(np.int32)
    row_off = np.concatenate([[0], np.cumsum(np.bincount(frame_of, minlength=n_rows))]).astype(np.int32)
    meta = torch.from_numpy(np.concatenate([row_off, order])).to(dev)
    wt = torch.from_numpy(w).to(dev)
    out = torch.full((n_rows, n_patch, c), 7.0, dtype=TORCH_DT[dtype], device=dev)
    _cabi.check(_cabi.lib().ufv_mask_pool_backward(
        wt.data_ptr(), bits24.data_ptr(), meta.data_ptr(), meta.data_ptr() + 4 * row_off.size, n_rows, 17, n_patch, c,
        out.data_ptr(), packer.FEAT_DTYPES[TORCH_DT[dtype]], torch.cuda.current_stream(dev).cuda_stream))
    want = torch.zeros((n_rows, n_patch, c), dtype=torch.float64, device=dev)
    want.index_add_(0, torch.from_numpy(frame_of).to(dev),
                    torch.from_numpy(on).to(dev).double()[:, :, None] * wt.double()[:, None, :])
    tol = 1e-5 if dtype == "f32" else 0.08
    assert (out.double() - want).abs().max().item() <= tol
    assert not out[1].any()


def test_training_path_gradients_match_the_reference_ops(dev):
    """With grad enabled and trainable parameters / features, forward() is differentiable: gradients of
    a scalar loss w.r.t. the projector weights and the features equal those of the reference's op
    sequence (oracle/reference_port.py under torch autograd), fp32."""
    from oracle import reference_port
    feats_np, masks_np, ann = synth.make_batch(2, 6, 2, "blob", h=96, w=96, first_clip=900, ragged=True)
    enc = make_encoder(dev, "f32", 3)
    enc.requires_grad_(True)
    feats = torch.from_numpy(feats_np).to(dev).requires_grad_(True)
    masks = [torch.from_numpy(m).to(dev) for m in masks_np]
    tokens, nums = enc(feats, masks, None, ann, None)
    assert tokens.requires_grad
    probe = torch.linspace(-1, 1, tokens.numel(), device=dev).reshape(tokens.shape)
    (tokens * probe).sum().backward()
    got = {"feats": feats.grad.clone(), "w1": enc.feat_linear[0].weight.grad.clone(),
           "b2": enc.feat_linear[2].bias.grad.clone()}
    feats2 = torch.from_numpy(feats_np).to(dev).requires_grad_(True)
    ws = [p.detach().clone().requires_grad_(True) for p in (enc.feat_linear[0].weight, enc.feat_linear[0].bias,
                                                            enc.feat_linear[2].weight, enc.feat_linear[2].bias)]
    ref_tokens, ref_nums = reference_port.encode(feats2, [m.float() for m in masks], ann, 3, *ws)
    assert nums == ref_nums and (tokens - ref_tokens).abs().max().item() <= 1e-5
    (ref_tokens * probe).sum().backward()
    scale = feats2.grad.abs().max().item()
    assert (got["feats"] - feats2.grad).abs().max().item() <= 1e-4 * max(scale, 1.0)
    assert (got["w1"] - ws[0].grad).abs().max().item() <= 1e-4 * max(ws[0].grad.abs().max().item(), 1.0)
    assert (got["b2"] - ws[3].grad).abs().max().item() <= 1e-4 * max(ws[3].grad.abs().max().item(), 1.0)
    with torch.no_grad():                                   # and the inference path still agrees
        t2, n2 = enc(feats.detach(), masks, None, ann, None)
    assert n2 == nums and (t2 - tokens.detach()).abs().max().item() <= 1e-5


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("m", [48, 256, 300])
def test_projector_backward_on_the_tensor_core_kernel(dev, dtype, m):
    """_Projector (forward and backward of feat_linear on the tcgen05 kernel: GELU epilogue that keeps the
    pre-activation, dgrad with the GELU-backward epilogue, wgrad over transposed operands, bias column sums)
    against torch autograd over the same 16-bit modules."""
    dt = TORCH_DT[dtype]
    w1, b1, w2, b2 = [torch.from_numpy(a).to(dev).to(dt) for a in synth.make_weights(5)]
    g = torch.Generator(device="cpu").manual_seed(m)
    x0 = (torch.randn((m, 1152), generator=g) * 0.05).to(dev).to(dt)
    probe = (torch.randn((m, 3584), generator=g) * 0.1).to(dev).to(dt)
    ours = [t.detach().clone().requires_grad_(True) for t in (x0, w1, b1, w2, b2)]
    ref = [t.detach().clone().requires_grad_(True) for t in (x0, w1, b1, w2, b2)]
    y = layer._Projector.apply(*ours)
    y_ref = torch.nn.functional.linear(torch.nn.functional.gelu(torch.nn.functional.linear(ref[0], ref[1], ref[2])),
                                       ref[3], ref[4])
    assert (y.float() - y_ref.float()).abs().max().item() <= 1e-2
    (y.float() * probe.float()).sum().backward()
    (y_ref.float() * probe.float()).sum().backward()
    for name, a, b in zip(("dx", "dw1", "db1", "dw2", "db2"), ours, ref):
        ga, gb = a.grad.float(), b.grad.float()
        assert ga.shape == gb.shape and torch.isfinite(ga).all(), name
        scale = gb.abs().max().item()
        # 16-bit gradients of sums over up to 3584 terms: compare against the magnitude of the gradient tensor
        assert (ga - gb).abs().max().item() <= 2e-2 * max(scale, 1e-3), (name, (ga - gb).abs().max().item(), scale)


def test_training_path_in_bf16_uses_the_tensor_core_backward(dev):
    """forward() under autograd in bf16: gradients w.r.t. features and projector parameters agree with the
    same computation when the projector runs under torch autograd / cuBLAS (UFV_TORCH_PROJECTOR_BACKWARD=1)."""
    feats_np, masks_np, ann = synth.make_batch(2, 6, 2, "blob", h=96, w=96, first_clip=910, ragged=True)
    masks = [torch.from_numpy(m).to(dev) for m in masks_np]
    grads = {}
    for mode in ("tcgen05", "torch"):
        if mode == "torch":
            os.environ["UFV_TORCH_PROJECTOR_BACKWARD"] = "1"
        try:
            enc = make_encoder(dev, "bf16", 3)
            enc.requires_grad_(True)
            feats = torch.from_numpy(feats_np).to(dev).bfloat16().requires_grad_(True)
            tokens, nums = enc(feats, masks, None, ann, None)
            probe = torch.linspace(-1, 1, tokens.numel(), device=dev).reshape(tokens.shape)
            (tokens.float() * probe).sum().backward()
            grads[mode] = (feats.grad.float(), enc.feat_linear[0].weight.grad.float(), enc.feat_linear[2].weight.grad.float(),
                           enc.feat_linear[2].bias.grad.float(), nums)
        finally:
            os.environ.pop("UFV_TORCH_PROJECTOR_BACKWARD", None)
    assert grads["tcgen05"][4] == grads["torch"][4]
    for a, b in zip(grads["tcgen05"][:4], grads["torch"][:4]):
        assert (a - b).abs().max().item() <= 3e-2 * max(b.abs().max().item(), 1e-3)


def test_long_objects_spread_over_many_ctas(dev):
    """T = 300 and T = 40 in one batch (K = 8): similarity and merge kernels index by (object, pair/slot)."""
    g = synth.rng_for(77)
    x = g.standard_normal((340, 1152), dtype=np.float32)
    host = {"obj_start": np.array([0, 300], np.int32), "obj_len": np.array([300, 40], np.int32),
            "slot_off": np.array([0, 8], np.int32)}
    plan = packer.EncodePlan(n_masks=340, n_groups=0, max_group=1, n_obj=2, max_len=300, m_pad=16,
                             slots=np.full(2, 8, np.int32), host=host)
    packer._upload(plan, dev)
    tok, counts, ex = layer.ttm(torch.from_numpy(x).to(dev), plan, 8, torch.float32, debug=True)
    for o, (a, b) in enumerate(((0, 300), (300, 340))):
        want, cut, sims = R.token_merge(x[a:b], 8)
        n = int(counts[o])
        assert n == want.shape[0]
        assert np.array_equal(ex["sims"][o, : b - a - 1].cpu().numpy(), sims)
        assert np.array_equal(ex["tokens_f32"][8 * o: 8 * o + n].cpu().numpy(), want)


# ---------------------------------------------------------------------------------------------
# BASELINE.json shapes through forward(), against the oracle (clips are independent, so the oracle runs on
# selected clips / objects only and still pins every quantity of those rows)
# ---------------------------------------------------------------------------------------------
def _forward_vs_oracle(dev, feats, masks, ann, k, clips, mask_dtype=torch.uint8, dtype="bf16"):
    """Run forward() on the whole batch; for every clip index in ``clips`` run the oracle on that clip alone and
    require: region_token_nums equal, pooled rows bit-identical, merged tokens bit-identical, projected tokens
    within the 16-bit tolerance."""
    enc = make_encoder(dev, dtype, k)
    ft = torch.from_numpy(feats).to(dev).to(TORCH_DT[dtype])
    md = [torch.from_numpy(m).to(dev).to(mask_dtype) for m in masks]
    tokens, counts = enc(ft, md, None, ann, None)
    pooled = enc._debug["pooled"].cpu().numpy()
    merged = enc._debug["merged"].float().cpu().numpy()
    plan = enc.last_plan
    slot_off = plan.host["slot_off"]
    tok = tokens.float().cpu().numpy()
    w = tuple(R.round_to(a, dtype) for a in synth.make_weights(0))
    q_off = np.concatenate([[0], np.cumsum([m.shape[0] for m in masks])])
    o_off = np.concatenate([[0], np.cumsum([len(a) for a in ann])])
    t_off = np.concatenate([[0], np.cumsum(counts)])
    assert tok.shape[0] == t_off[-1]
    for i in clips:
        rows = sorted({r for obj in ann[i] for r in obj})
        remap = {r: j for j, r in enumerate(rows)}
        ann_i = [[[remap[r] for r in obj] for obj in ann[i]]]
        o = R.encode(R.round_to(feats[rows], dtype), [masks[i]], ann_i, k, dtype, w)
        a, b = int(o_off[i]), int(o_off[i + 1])
        assert counts[a:b] == o["counts"], i
        assert np.array_equal(pooled[q_off[i]:q_off[i + 1]], o["pooled"]), i
        want_rows = np.concatenate([np.arange(slot_off[j], slot_off[j] + counts[j]) for j in range(a, b)])
        assert np.array_equal(merged[want_rows], o["merged"]), i
        assert np.abs(tok[t_off[a]:t_off[b]] - o["tokens"]).max() <= TOL[dtype], i
    return enc, tokens, counts


@pytest.mark.parametrize("mask_dtype", [torch.uint8, torch.float32])
def test_c3_share_sparse_masks_vs_oracle(dev, mask_dtype):
    """BASELINE configs[2], one GPU's share: 8 of 64 clips x 32 frames x 8 objects, sparse / irregular masks
    (every 7th mask all-zero), K = 8: 2048 object-frames, 512 tokens."""
    feats, masks, ann = synth.make_batch(8, 32, 8, "sparse")
    _, tokens, counts = _forward_vs_oracle(dev, feats, masks, ann, 8, clips=(0, 5, 7), mask_dtype=mask_dtype)
    assert len(counts) == 64 and tokens.shape[1] == 3584


def test_c4_long_clip_16_objects_vs_oracle(dev):
    """BASELINE configs[3], one clip: 256 frames x 16 objects (two or more pool passes per frame, the split
    similarity / merge kernels, r = 248 merges per object)."""
    feats, masks, ann = synth.make_batch(1, 256, 16, "blob")
    enc, tokens, counts = _forward_vs_oracle(dev, feats, masks, ann, 8, clips=(0,))
    assert counts == [8] * 16 and tuple(tokens.shape) == (128, 3584)
    assert enc.last_plan.max_len == 256


def test_c5_wide_ragged_64_objects_vs_oracle(dev):
    """BASELINE configs[4] corner: 1 clip x 64 frames x 64 objects, ragged T_o in [1, 64] (objects shorter
    than K pass through, many objects per frame)."""
    feats, masks, ann = synth.make_batch(1, 64, 64, "blob", ragged=True)
    enc, tokens, counts = _forward_vs_oracle(dev, feats, masks, ann, 8, clips=(0,))
    assert counts == [min(len(o), 8) for o in ann[0]]


def test_c5_long_512_frames_vs_oracle(dev):
    """BASELINE configs[4] corner: 1 clip x 512 frames x 4 objects, dense masks, T = 512 per object."""
    feats, masks, ann = synth.make_batch(1, 512, 4, "dense")
    enc, tokens, counts = _forward_vs_oracle(dev, feats, masks, ann, 8, clips=(0,))
    assert counts == [8] * 4 and enc.last_plan.max_len == 512


def test_c2_full_shape_vs_oracle_every_clip(dev):
    """BASELINE configs[1] at full size, fp32 masks as the reference contract has them: every clip against the
    oracle (counts, pooled, merged bit-exact; projected tokens within 1e-2)."""
    feats, masks, ann = synth.make_batch(8, 16, 4, "dense")
    _, tokens, counts = _forward_vs_oracle(dev, feats, masks, ann, 8, clips=range(8), mask_dtype=torch.float32)
    assert counts == [8] * 32 and tuple(tokens.shape) == (256, 3584)


def test_c2_shape_properties_bf16(dev):
    """BASELINE configs[1] at full size (8 clips x 16 frames x 4 objects, bf16): size-independent
    properties instead of an oracle run."""
    feats, masks, ann = synth.make_batch(8, 16, 4, "dense")
    enc = make_encoder(dev, "bf16", 8)
    ft = torch.from_numpy(feats).to(dev).bfloat16()
    md = [torch.from_numpy(m).to(dev) for m in masks]
    tokens, counts = enc(ft, md, None, ann, None)
    assert counts == [8] * 32 and tuple(tokens.shape) == (256, 3584)
    assert torch.isfinite(tokens.float()).all()
    # sharding invariance: each clip alone gives bit-identical rows
    for i in (0, 5):
        ann_i = [[[r - i * 16 for r in o] for o in ann[i]]]
        t_i, c_i = enc(ft[i * 16:(i + 1) * 16], [md[i]], None, ann_i, None)
        assert torch.equal(t_i, tokens[i * 32:(i + 1) * 32]) and c_i == [8] * 4
    # pooling a constant feature map returns the constant exactly; all-on mask = plain mean
    const = torch.full_like(ft, 0.5)
    enc(const, md, None, ann, None)
    assert (enc._debug["pooled"] == 0.5).all()
    # spot-check pooled rows of one clip against the oracle (bit-exact)
    on = np.stack([R.mask_to_patches(m) for m in masks[3]])
    rows = [r for o in ann[3] for r in o]
    enc(ft, md, None, ann, None)
    want = R.mask_pool(R.round_to(feats, "bf16"), rows, on)
    got = enc._debug["pooled"][3 * 64:4 * 64].cpu().numpy()
    assert np.array_equal(got, want)
