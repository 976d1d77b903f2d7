"""Developer tool: where the host time of MaskExtractor.forward goes (cProfile + wall clock)."""
import cProfile
import os
import pstats
import sys
import time
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ufvideo_b200 import build_region_encoder, synth  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    feats, masks, ann = synth.make_batch(8, 16, 4, "dense")
    ft = torch.from_numpy(feats).to(dev).bfloat16()
    md = [torch.from_numpy(m).to(dev).float() for m in masks]
    enc = build_region_encoder(types.SimpleNamespace(mm_hidden_size=1152, hidden_size=3584), "square")
    enc.region_token_num = 8
    enc.requires_grad_(False)
    enc = enc.to(dev).bfloat16()
    for _ in range(20):
        enc(ft, md, None, ann, None)
    torch.cuda.synchronize()
    n = 2000
    t0 = time.perf_counter()
    for _ in range(n):
        enc(ft, md, None, ann, None)
    torch.cuda.synchronize()
    print(f"forward wall: {(time.perf_counter() - t0) / n * 1e6:.1f} us/call")
    t0 = time.perf_counter()
    for _ in range(n):
        enc.encode_padded(ft, md, ann)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"encode_padded host-only: {(t1 - t0) / n * 1e6:.1f} us/call, with drain {(time.perf_counter() - t0) / n * 1e6:.1f}")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        enc(ft, md, None, ann, None)
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(25)


if __name__ == "__main__" and not os.environ.get("UFV_FINE"):
    main()


def fine_grained():
    """perf_counter around the pieces of one forward() (no profiler overhead)."""
    import ctypes
    from ufvideo_b200 import _cabi, packer, layer
    dev = torch.device("cuda:0")
    feats, masks, ann = synth.make_batch(8, 16, 4, "dense")
    ft = torch.from_numpy(feats).to(dev).bfloat16()
    md = [torch.from_numpy(m).to(dev).float() for m in masks]
    enc = build_region_encoder(types.SimpleNamespace(mm_hidden_size=1152, hidden_size=3584), "square")
    enc.region_token_num = 8
    enc.requires_grad_(False)
    enc = enc.to(dev).bfloat16()
    for _ in range(10):
        enc(ft, md, None, ann, None)
    torch.cuda.synchronize()
    n = 300
    acc = {"build_plan": 0.0, "ufv_encode (C)": 0.0, "encode_padded total": 0.0, "await counts": 0.0, "forward total": 0.0}
    lib = _cabi.lib()
    for _ in range(n):
        t0 = time.perf_counter()
        plan = packer.build_plan(md, ann, ft.shape[0], 8, dev)
        t1 = time.perf_counter()
        acc["build_plan"] += t1 - t0
        torch.cuda.synchronize()
        run = plan.run
        t0 = time.perf_counter()
        _cabi.check(lib.ufv_encode(run["args_ref"], torch._C._cuda_getCurrentRawStream(0)))
        t1 = time.perf_counter()
        acc["ufv_encode (C)"] += t1 - t0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tokens, counts, plan = enc.encode_padded(ft, md, ann)
        t1 = time.perf_counter()
        layer._await_counts(plan, plan.run, dev)
        t2 = time.perf_counter()
        acc["encode_padded total"] += t1 - t0
        acc["await counts"] += t2 - t1
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        enc(ft, md, None, ann, None)
        acc["forward total"] += time.perf_counter() - t0
        torch.cuda.synchronize()
    for k, v in acc.items():
        print(f"{k:24s} {v / n * 1e6:8.1f} us")
    # cold path: a batch structure never seen before on every call (what a real eval loop does)
    import cProfile, pstats
    n = 40
    variants = []
    for i in range(n):
        a2 = [[[r for r in obj[: 16 - (i % 7)]] for obj in clip] for clip in ann]
        a2[0][0] = a2[0][0][: 3 + i % 5]
        variants.append(a2)
    md2 = []
    for i in range(n):
        md2.append([m[: sum(len(o) for o in variants[i][c])] for c, m in enumerate(md)])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        enc(ft, md2[i], None, variants[i], None)
    torch.cuda.synchronize()
    print(f"cold forward (new structure every call): {(time.perf_counter() - t0) / n * 1e6:.1f} us/call")
    variants = [[[list(o)[::-1][: len(o)] for o in clip] for clip in v] for v in variants]
    pr = cProfile.Profile()
    pr.enable()
    for i in range(n):
        enc(ft, md2[i], None, variants[i], None)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)


if __name__ == "__main__" and os.environ.get("UFV_FINE"):
    fine_grained()
