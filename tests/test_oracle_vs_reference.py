"""Build-container-only checks against the REAL reference tree (skipped where /root/reference is
absent, i.e. on the GPU box): the torch-CPU port used as the timed CPU baseline is bit-identical
to the reference module, and the numpy oracle agrees with it on fresh random cases."""
import numpy as np
import pytest
import torch

from oracle import golden_cases as gc
from oracle import ref_loader, reference_port, restatement as R
from ufvideo_b200 import synth

ref = ref_loader.load_reference_layer()
pytestmark = pytest.mark.skipif(ref is None, reason="reference tree not present")


def reference_module(k, aspect, weights, dtype=torch.float32):
    enc = ref.build_region_encoder(ref_loader.reference_config(), aspect)
    enc.region_token_num = k
    with torch.no_grad():
        for p, w in zip((enc.feat_linear[0].weight, enc.feat_linear[0].bias,
                         enc.feat_linear[2].weight, enc.feat_linear[2].bias), weights):
            p.copy_(torch.from_numpy(w))
    return enc.to(dtype).eval()


@pytest.mark.parametrize("name", ["c1", "multi", "pad", "shared"])
def test_port_is_bit_identical_to_reference_module(name):
    case = gc.e2e_case(name)
    weights = synth.make_weights(0)
    enc = reference_module(case["k"], case["aspect"], weights)
    feats = torch.from_numpy(case["feats"])
    masks = [torch.from_numpy(m).float() for m in case["masks"]]
    with torch.no_grad():
        want, counts = enc(feats, masks, feats, case["ann"], None)
        got, got_counts = reference_port.encode(feats, masks, case["ann"], case["k"],
                                                *[torch.from_numpy(w) for w in weights],
                                                pad_square=case["aspect"] == "pad")
    assert got_counts == counts and torch.equal(got, want)


def test_oracle_decisions_equal_reference_on_fresh_random_objects():
    g = synth.rng_for(2718)
    for trial in range(40):
        t = int(g.integers(2, 200))
        k = int(g.choice([1, 4, 8]))
        if t <= k:
            continue
        x = g.standard_normal((t, 1152), dtype=np.float32)
        want = ref.token_merge(torch.from_numpy(x)[None], t - k)[0].numpy()
        tok, cut, _ = R.token_merge(x, k)
        assert tok.shape == want.shape and np.abs(tok - want).max() <= 1e-5
