"""Developer tool: projector GEMM throughput vs token count M, beside torch (cuBLAS) on the same GPU.

Each measurement flushes nothing: weights (8 / 26 MB) are L2-resident for both contenders, which is
also the situation inside the chained path.  Reports TFLOP/s and the fraction of the measured bf16
peak (MEASURED_PEAKS.json)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ufvideo_b200 import layer  # noqa: E402


def timed(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3   # us


_flush = None


def timed_cold(fn, iters=20, warm=3):
    """Like timed(), but the 126 MB L2 is overwritten before every call (a 512 MB fill on the same stream,
    outside the events), so weights and tokens come from HBM -- the situation inside the real step, where
    215 MB of features pass through the L2 between two projector calls."""
    global _flush
    if _flush is None:
        _flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        _flush.fill_(1)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in evs)
    return t[len(t) // 2] * 1e3


def main():
    dev = torch.device("cuda:0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peak = 1632.6
    p = os.path.join(root, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        peak = json.load(open(p))["bf16_tflops"]
    ms = [int(x) for x in sys.argv[1:]] or [32, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]
    print("UFV_GEMM_BN =", os.environ.get("UFV_GEMM_BN", "(cost model)"))
    print(f"{'M':>6} {'layer':>8} {'ours us':>9} {'torch us':>9} {'ours TF':>8} {'torch TF':>8} {'ours/peak':>9} "
          f"{'ours cold':>9} {'torch cold':>10}")
    for m in ms:
        for name, k, n, gelu in (("lin1+gelu", 1152, 3584, True), ("lin2", 3584, 3584, False)):
            x = (torch.randn((m, k), device=dev) * 0.05).bfloat16()
            w = (torch.randn((n, k), device=dev) * 0.03).bfloat16()
            b = (torch.randn((n,), device=dev) * 0.03).bfloat16()
            t_ours = timed(lambda: layer.linear(x, w, b, gelu=gelu))
            if gelu:
                t_torch = timed(lambda: torch.nn.functional.gelu(torch.nn.functional.linear(x, w, b)))
            else:
                t_torch = timed(lambda: torch.nn.functional.linear(x, w, b))
            c_ours = timed_cold(lambda: layer.linear(x, w, b, gelu=gelu))
            if gelu:
                c_torch = timed_cold(lambda: torch.nn.functional.gelu(torch.nn.functional.linear(x, w, b)))
            else:
                c_torch = timed_cold(lambda: torch.nn.functional.linear(x, w, b))
            fl = 2.0 * m * k * n
            print(f"{m:6d} {name:>8} {t_ours:9.1f} {t_torch:9.1f} {fl / t_ours / 1e6:8.1f} {fl / t_torch / 1e6:8.1f} "
                  f"{fl / t_ours / 1e6 / peak:9.3f} {c_ours:9.1f} {c_torch:10.1f}")


if __name__ == "__main__":
    main()
