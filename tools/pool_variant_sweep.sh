#!/bin/bash
# Developer tool: time kernel 2 for several (rows per stage, stages) builds.  Build the variants first:
#   for cfg in "16 6" "16 4" "32 4"; do set -- $cfg; make -C ufvideo_b200/csrc BUILD=build_r$1s$2 \
#       OUT=../libufv_r$1s$2.so EXTRA="-DUFV_POOL_ROWS=$1 -DUFV_POOL_STAGES=$2"; done
# then run this script on the GPU box (variants are picked up through UFV_B200_LIB).
for v in b200 $(ls ufvideo_b200 | sed -n 's/^libufv_\(r[0-9]*s[0-9]*\)\.so$/\1/p'); do
  echo "variant $v"
  UFV_B200_LIB=$PWD/ufvideo_b200/libufv_$v.so python - <<'PY'
import sys, os, types
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from ufvideo_b200 import layer, packer, synth
dev = torch.device("cuda:0")
def timed(fn, iters=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
for name, (clips, frames, objs, fam) in {"c2": (8, 16, 4, "dense"), "c3d": (8, 32, 8, "dense"), "c2blob": (8, 16, 4, "blob")}.items():
    feats, masks, ann = synth.make_batch(clips, frames, objs, fam)
    ft = torch.from_numpy(feats).to(dev).bfloat16()
    md = [torch.from_numpy(m).to(dev) for m in masks]
    plan = packer.build_plan(md, ann, ft.shape[0], 8, dev, use_cache=False)
    patches = layer.mask_to_patches(plan, dev)
    nu = int(patches["grp_nu"].sum().item())
    t = timed(lambda: layer.mask_pool(ft, plan, patches))
    print(f"  {name}: {t:7.1f} us  {nu * 2304 / t / 1e3:6.0f} GB/s")
PY
done
