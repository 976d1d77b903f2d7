for v in 256 128 64 0; do echo "l2promo=$v"; UFV_TMAP_L2PROMO=$v python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
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
for name, (clips, frames, objs, fam) in {"c2": (8, 16, 4, "dense"), "c3d": (8, 32, 8, "dense")}.items():
    feats, masks, ann = synth.make_batch(clips, frames, objs, fam)
    ft = torch.from_numpy(feats).to(dev).bfloat16()
    md = [torch.from_numpy(m).to(dev) for m in masks]
    plan = packer.build_plan(md, ann, ft.shape[0], 8, dev, use_cache=False)
    patches = layer.mask_to_patches(plan, dev)
    nu = int(patches["grp_nu"].sum().item())
    t = timed(lambda: layer.mask_pool(ft, plan, patches))
    print(f"  {name}: {t:7.1f} us  {nu * 2304 / t / 1e3:6.0f} GB/s")
if True:
    x = torch.empty((128, 729, 1152), dtype=torch.bfloat16, device=dev)
    t = timed(lambda: x.zero_())
    print(f"  memset 215 MB: {t:6.1f} us {x.numel()*2/t/1e3:6.0f} GB/s")
PY
done
