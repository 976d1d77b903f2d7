"""Developer tool: phase timestamps of CTA 0 of the fused merge kernel inside the real step.

Needs a trace build of the library:
    make -C ufvideo_b200/csrc BUILD=build_trace OUT=../libufv_b200_trace.so EXTRA=-DUFV_TTM_TRACE
    UFV_B200_LIB=ufvideo_b200/libufv_b200_trace.so python tools/ttm_trace.py
"""
import ctypes
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ufvideo_b200 import _cabi, build_region_encoder, synth  # noqa: E402

NAMES = ["entry", "plan scalars loaded", "griddepcontrol.wait returned", "rows staged", "norms done",
         "sims done", "threshold / cuts / run ends done", "means stored"]


def main():
    dev = torch.device("cuda:0")
    feats, masks, ann = synth.make_batch(8, 16, 4, "dense")
    ft = torch.from_numpy(feats).to(dev).bfloat16()
    md = [torch.from_numpy(m).to(dev).float() for m in masks]
    enc = build_region_encoder(types.SimpleNamespace(mm_hidden_size=1152, hidden_size=3584), "square")
    enc.region_token_num = 8
    enc.requires_grad_(False)
    enc = enc.to(dev).bfloat16()
    lib = _cabi.lib()
    fn = lib.ufv_debug_ttm_trace
    fn.restype, fn.argtypes = ctypes.c_int, [ctypes.c_void_p]
    rows = []
    with torch.inference_mode():
        for i in range(30):
            enc(ft, md, None, ann, None)
            torch.cuda.synchronize()
            buf = np.zeros(16, np.uint64)
            assert fn(buf.ctypes.data) == 0
            if i >= 10:
                rows.append(buf[:8].astype(np.int64))
    t = np.stack(rows)
    d = np.diff(t, axis=1)
    print("fused merge kernel, CTA 0, ns between phase boundaries (median over 20 steps):")
    for i in range(7):
        print(f"  {NAMES[i]:34s} -> {NAMES[i + 1]:34s} {np.median(d[:, i]):8.0f}")
    print(f"  total entry -> means stored {np.median(t[:, 7] - t[:, 0]):8.0f};  after the wait {np.median(t[:, 7] - t[:, 2]):8.0f}")


if __name__ == "__main__":
    main()
