#!/usr/bin/env python
"""bench.py -- object-frames/s of the object-encoder hot path (mask-pool + TTM + projector).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the whole hot path (mask resize/binarise -> mask pool -> temporal token
merge -> projector) over one batch of BASELINE.json's configs[1]:
8 clips x 16 frames x 4 objects per GPU, bf16 features [128, 729, 1152], 384x384 float32 masks
(dense family), K = 8, projector 1152 -> 3584 -> 3584.  Synthetic seeded data (ufvideo_b200/synth.py).

value   device-resident inputs, CUDA events around exactly K steps, max over ranks.
e2e     the same step through the module's public forward() with HOST (pinned) feats and masks:
        the H2D of the inputs and the D2H of the projected tokens are inside the timed region.
roofline  the dominant kernel (segmented mask pool, HBM-bound) timed alone with CUDA events.
cpu_baseline  the reference's CPU implementation of the whole workload on all host cores (rank 0, N=1):
        the real ufvideo/model/layer.py when a reference tree is reachable, else its torch-CPU port.
--impl reference  times that CPU implementation alone, all 8 clips of configs[1] per step.
"""
from __future__ import annotations

import argparse
import collections
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "object_frames_per_s"
UNIT = "object-frames/s"
WORKLOAD = dict(clips_per_gpu=8, frames=16, objects=4, k=8, family="dense", mask_hw=(384, 384))
L2_BYTES = 126 * 1024 * 1024
# result collection at N > 1: "push" (side-stream peer push, overlapped with the next step), "fused" (stores fused into
# the last Linear's epilogue: lowest latency for one step, but the NVLink transfer sits on the stream's critical
# path), "nccl" (one asynchronous all-gather per step), "none"; UFV_BENCH_GATHER overrides
GATHER_DEFAULT = "push"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=WORKLOAD["clips_per_gpu"])
    ap.add_argument("--frames", type=int, default=WORKLOAD["frames"])
    ap.add_argument("--objects", type=int, default=WORKLOAD["objects"])
    ap.add_argument("--family", default=WORKLOAD["family"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true",
                    help="skip the clip-sharded configs[2] / configs[3] runs reported under 'sweep'")
    ap.add_argument("--timeline", default="", metavar="FILE",
                    help="developer knob: trace the steady-state device-resident step with CUPTI (torch.profiler) "
                         "and write the per-kernel timeline of a few steps to FILE instead of benchmarking")
    return ap.parse_args()


def workload_name(a):
    return (f"configs[1]: {a.clips} clips x {a.frames} frames x {a.objects} objects per GPU, bf16, "
            f"K={WORKLOAD['k']}, projector 1152->3584->3584")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi in the background during the timed regions)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, enabled: bool = True):
        self.index = index
        self.enabled = enabled          # only rank 0 samples: one nvidia-smi poller per box is enough
        self.proc = None
        self.file = None

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            self.proc.wait()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.file is None:
            return out
        self.file.flush()
        rows = [r.strip().split(", ") for r in open(self.file.name) if r.strip()]
        os.unlink(self.file.name)
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.strip().lower() == "active":
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's torch-CPU port of the reference on a bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(a, steps: int, warmup: int, sample_clips: int | None = None, min_seconds: float = 0.0):
    """The reference's CPU implementation of the path on all host cores.  The real ``layer.py`` is used
    whenever a reference tree is reachable ($UFV_REF, /root/reference, baseline/_ref: kind "reference");
    otherwise (the GPU box: a Python reference cannot travel) the op-for-op torch-CPU port of it under
    oracle/ (kind "port", validated bit-identical to the real module by tests/test_oracle_vs_reference.py)."""
    import torch

    from oracle import ref_loader, reference_port
    from ufvideo_b200 import synth

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_clips = a.clips if sample_clips is None else min(sample_clips, a.clips)
    feats, masks, ann = synth.make_batch(n_clips, a.frames, a.objects, a.family)
    weights = [torch.from_numpy(w) for w in synth.make_weights(0)]
    ft = torch.from_numpy(feats)                       # fp32: CPU bf16 kernels are not the reference's CPU path
    mt = [torch.from_numpy(m).float() for m in masks]
    q = sum(m.shape[0] for m in masks)
    ref = ref_loader.load_reference_layer()
    if ref is not None:
        kind = "reference"
        enc = ref.build_region_encoder(ref_loader.reference_config(), "square")
        enc.region_token_num = WORKLOAD["k"]
        with torch.no_grad():
            for p, w in zip((enc.feat_linear[0].weight, enc.feat_linear[0].bias,
                             enc.feat_linear[2].weight, enc.feat_linear[2].bias), weights):
                p.copy_(w)
        enc = enc.eval()

        def run():
            return enc(ft, mt, ft, ann, None)
    else:
        kind = "port"

        def run():
            return reference_port.encode(ft, mt, ann, WORKLOAD["k"], *weights)
    with torch.no_grad():
        for _ in range(warmup):
            run()
        t0 = time.perf_counter()
        done = 0
        while done < steps or time.perf_counter() - t0 < min_seconds:
            run()
            done += 1
        dt = time.perf_counter() - t0
    which = ("the reference's own ufvideo/model/layer.py (" + os.path.dirname(os.path.dirname(os.path.dirname(
        ref_loader.reference_layer_path()))) + ")" if kind == "reference"
        else "oracle/reference_port.py, the op-for-op torch-CPU port of layer.py (no reference tree on this box)")
    whole = n_clips == a.clips
    sample = (("the whole workload: " if whole else f"{n_clips} of the workload's {a.clips} clips: ")
              + f"{n_clips} clips x {a.frames} frames x {a.objects} objects = {q} object-frames per pass, fp32, "
              f"{cores} threads, {done} timed passes after {warmup} warm-up; {which}")
    return q * done / dt, dt / done, cores, sample, done, kind, whole


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, sec, cores, sample, done, kind, whole = cpu_reference_run(a, a.steps, max(a.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": done, "warmup": max(a.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a),
                   "step": ("one pass of the reference's CPU implementation over " + sample +
                            "; one rank's share of the workload (weak scaling: every rank holds the same "
                            "amount), timed on rank 0's host cores")},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _parse_cpu_list(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def _topo_cpu_affinity(gpu_index: int):
    """CPU affinity of a GPU as `nvidia-smi topo -m` reports it (works where sysfs hides the NUMA node)."""
    out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    header = None
    for line in out.splitlines():
        cols = [c.strip() for c in line.split("\t") if c.strip()]
        if not cols:
            continue
        if header is None and any("CPU Affinity" in c for c in cols):
            header = cols
            continue
        if header is not None and cols[0] == f"GPU{gpu_index}":
            # data rows carry one more leading column (the row label) than the header
            idx = next(i for i, c in enumerate(header) if "CPU Affinity" in c) + 1
            if idx < len(cols) and cols[idx][0].isdigit():
                return _parse_cpu_list(cols[idx])
    return None


def bind_to_gpu_numa_node(local: int) -> str:
    """Pin this rank (and therefore its first-touch pinned host buffers) to the CPUs next to its GPU: with one
    process per GPU the H2D stream then never crosses the socket interconnect."""
    try:
        import torch
        prop = torch.cuda.get_device_properties(local)
        bdf = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        cpus, how = None, ""
        try:
            node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
            if node >= 0:
                cpus = _parse_cpu_list(open(f"/sys/devices/system/node/node{node}/cpulist").read())
                how = f"sysfs node {node}"
        except OSError:
            pass
        if cpus is None:
            cpus = _topo_cpu_affinity(local)
            how = "nvidia-smi topo"
        if not cpus:
            return "numa: no affinity information"
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return f"numa: {how}: no allowed cpu"
        if allowed == os.sched_getaffinity(0):
            return f"numa: {how}: all {len(allowed)} cpus are local"
        os.sched_setaffinity(0, allowed)
        return f"numa: rank bound to {len(allowed)} cpus ({how})"
    except Exception as exc:   # noqa: BLE001 -- binding is an optimisation, never a requirement
        return f"numa: not bound ({type(exc).__name__})"


def write_timeline(path, step, drain, barrier, rank, world, steps: int = 8):
    """Per-kernel start / duration / gap table of the steady-state step as CUPTI sees it inside the real
    pipeline (graph replay, programmatic dependent launch, concurrent kernels) -- what ncu's serialised,
    cold-cache launch list cannot show.  Times under the tracer are not bench values."""
    import torch
    from torch.profiler import ProfilerActivity, profile

    for _ in range(10):
        step()
    drain()
    barrier()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(steps):
            step()
        drain()
        torch.cuda.synchronize()
    barrier()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    lines = [f"# rank {rank} of {world}: {steps} traced steps; columns: start us (from the first kernel), "
             f"duration us, gap to the previous kernel's end us (negative = overlap), name"]
    if ev:
        t0 = ev[0].time_range.start
        prev_end = t0
        for e in ev:
            st, en = e.time_range.start, e.time_range.end
            lines.append(f"{st - t0:10.1f} {en - st:8.1f} {st - prev_end:8.1f}  {e.name[:110]}")
            prev_end = max(prev_end, en)
        total = prev_end - t0
        lines.append(f"# span {total:.1f} us over {steps} steps = {total / steps:.1f} us per step")
    out = path if world == 1 else f"{path}.rank{rank}"
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    with open(out, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    if rank == 0:
        print(f"[bench] timeline written to {out} ({len(ev)} device events)", file=sys.stderr)


# ------------------------------------------------------------------------------------------------
# clip-sharded configs[2] / configs[3] (strong scaling: the total number of clips is fixed, rank r owns a
# contiguous block of them -- the reference's get_chunk sharding, eval/inference_PixRQA.py:186 -- and the
# per-rank results are collected by the all-gather fused into the last Linear)
# ------------------------------------------------------------------------------------------------
SWEEP = {
    # name: (clips in total, frames, objects, mask family, BASELINE.json entry)
    "c3": (64, 32, 8, "sparse", "configs[2]: PixRQA-style batch, 64 clips x 32 frames x 8 objects, sparse/irregular masks"),
    "c4": (32, 256, 16, "blob", "configs[3]: long video, 256 frames x 16 objects per clip (merge-heavy, r = 248); 32 clips"),
}


def run_sweep(enc, dev, rank, world, peak, steps=10, warmup=3):
    """One entry per config: whole-job ms per step at this world size (CUDA events, max over ranks), the same
    job on ONE GPU timed in the same run (rank 0, gives the scaling efficiency), and the pool kernel's fraction
    of the measured HBM peak by SURVEY 8(d) bytes on this rank's shard.  Data: one clip's seeded synthetic masks
    (uint8 384 x 384) reused for every clip, random bf16 features generated on the device (distinct per frame)."""
    import torch
    import torch.distributed as dist

    from ufvideo_b200 import layer, packer, sharding, synth

    k = WORKLOAD["k"]
    out = []
    for name, (n_clips, frames, objects, family, note) in SWEEP.items():
        _, clip_masks, clip_ann = synth.make_clip(0, frames, objects, family, feats=False)
        mask_dev = torch.from_numpy(clip_masks).to(dev)                                   # shared by all clips
        q_clip = clip_masks.shape[0]

        def build(clips):
            n = len(clips)
            g = torch.Generator(device=dev).manual_seed(4321 + clips.start)
            feats = torch.randn((max(n, 1) * frames, 729, 1152), generator=g, device=dev, dtype=torch.bfloat16)
            ann = [[[r + i * frames for r in obj] for obj in clip_ann] for i in range(n)]
            return feats, [mask_dev] * n, ann

        def time_steps(fn, sync_ranks):
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize()
            if sync_ranks and world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
            if sync_ranks and world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item())

        block = sharding.clip_block(n_clips, rank, world)
        feats, masks, ann = build(block)
        slots = packer.build_plan(masks, ann, feats.shape[0], k, dev).slots
        pg = None
        if world > 1:
            pad_objs = -(-n_clips // world) * objects
            pg = sharding.PeerGather(pad_objs * k, pad_objs, 3584, torch.bfloat16, dev)
        last = [None]

        def step():
            if pg is None:
                enc(feats, masks, None, ann, None)
                return
            peer, tok_view, cnt_view, s = pg.begin(slots)
            if os.environ.get("UFV_BENCH_GATHER", GATHER_DEFAULT) == "push":
                enc.forward_padded(feats, masks, ann, out=tok_view, counts_out=cnt_view,
                                   after_enqueue=lambda: pg.push(s, peer, int(slots.sum())))
            else:
                enc.forward_padded(feats, masks, ann, out=tok_view, counts_out=cnt_view, peer=peer)
            last[0] = s

        def sharded():
            step()

        ms = time_steps(sharded, True)
        if pg is not None:
            pg.wait(last[0])
            torch.cuda.synchronize()
            pg.check()
        # pool kernel of this rank's shard, alone
        plan = packer.build_plan(masks, ann, feats.shape[0], k, dev)
        patches = layer.mask_to_patches(plan, dev)
        pool_bytes = packer.algorithmic_pool_bytes(plan, patches["bits"].cpu().numpy(), 1152, 2)
        ms_pool = time_steps(lambda: layer.mask_pool(feats, plan, patches), False)
        entry = {"config": name, "workload": note, "clips_total": n_clips, "clips_on_rank0": len(block),
                 "object_frames_per_step": n_clips * q_clip, "n_gpus": world, "ms_per_step": ms,
                 "object_frames_per_s": n_clips * q_clip / (ms * 1e-3),
                 "pool": {"us_per_launch": ms_pool * 1e3, "algorithmic_bytes": pool_bytes,
                          "achieved_gbs": pool_bytes / (ms_pool * 1e-3) / 1e9,
                          "frac_of_measured_hbm": pool_bytes / (ms_pool * 1e-3) / 1e9 / peak,
                          "objects_per_frame": objects}}
        del feats, masks, ann, plan, patches, pg
        torch.cuda.empty_cache()
        if world > 1:
            # the whole job on one GPU, same process, for the efficiency of this world size
            ms1 = torch.zeros(1, device=dev)
            if rank == 0:
                f1, m1, a1 = build(range(n_clips))
                ms1[0] = time_steps(lambda: enc(f1, m1, None, a1, None), False)
                del f1, m1, a1
                torch.cuda.empty_cache()
            dist.broadcast(ms1, 0)
            entry["ms_per_step_1gpu_same_run"] = float(ms1.item())
            entry["scaling_efficiency"] = float(ms1.item()) / (world * ms)
        out.append(entry)
    return out


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist

    from ufvideo_b200 import build_region_encoder, layer, packer, sharding, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device -- the object-encoder path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_note = bind_to_gpu_numa_node(local) if world > 1 else "numa: not bound (1 GPU)"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    k = WORKLOAD["k"]
    h, w = WORKLOAD["mask_hw"]

    # ---- data: this rank's clips (weak scaling: clips_per_gpu fixed) -----------------------------
    feats_np, masks_np, ann = synth.make_batch(a.clips, a.frames, a.objects, a.family, h, w,
                                               first_clip=rank * a.clips)
    feats_host = torch.from_numpy(feats_np).bfloat16().pin_memory()
    masks_host = [torch.from_numpy(m).float().pin_memory() for m in masks_np]     # reference contract: float32 0/1
    feats_dev = feats_host.to(dev)
    masks_dev = [m.to(dev) for m in masks_host]
    q = sum(m.shape[0] for m in masks_np)
    n_obj = sum(len(x) for x in ann)

    enc = build_region_encoder(types.SimpleNamespace(mm_hidden_size=1152, hidden_size=3584), "square")
    enc.region_token_num = k
    enc.requires_grad_(False)                 # forward-only path
    with torch.no_grad():
        for p, wt in zip((enc.feat_linear[0].weight, enc.feat_linear[0].bias,
                          enc.feat_linear[2].weight, enc.feat_linear[2].bias), synth.make_weights(0)):
            p.copy_(torch.from_numpy(wt))
    enc = enc.to(dev).bfloat16()
    pad_rows, pad_objs = n_obj * k, n_obj

    # N > 1: the projector and the merge kernel write straight into the all-gather payload
    # (sharding.new_payload); the ONE collective per step is an asynchronous NCCL all-gather of
    # that payload, overlapped with the next step's kernels and drained before the clock stops.
    slots = packer.build_plan(masks_dev, ann, feats_dev.shape[0], k, dev).slots
    pending = collections.deque()
    pg, gather_kind = None, "none (1 GPU)"
    push_mode = os.environ.get("UFV_BENCH_GATHER", GATHER_DEFAULT) == "push"
    m_pad_rows = int(slots.sum())
    if world > 1:
        # Preferred: the all-gather fused into the last Linear (ufv_linear_gather): its epilogue stores
        # every tile into all ranks' gathered buffers over NVLink.  UFV_BENCH_GATHER=nccl (or a failed
        # symmetric-memory rendezvous) falls back to one asynchronous NCCL all-gather of the payload.
        if os.environ.get("UFV_BENCH_GATHER", GATHER_DEFAULT) not in ("nccl", "none"):
            try:
                pg = sharding.PeerGather(pad_rows, pad_objs, 3584, torch.bfloat16, dev)
                how = "multimem.st via NVSwitch multicast" if pg.multimem else "one st per peer"
                if push_mode:
                    gather_kind = ("peer push: the last Linear writes this rank's rows into its slice of the symmetric "
                                   f"buffer, a side-stream kernel (ufv_peer_push) stores them to every rank over NVLink ({how}), "
                                   "arrival flags, overlapped with the next step, no NCCL on the data path")
                else:
                    gather_kind = (f"fused into the last Linear: tcgen05 epilogue stores tiles to every rank over NVLink ({how}), "
                                   "arrival flags, no NCCL on the data path")
            except Exception as exc:   # noqa: BLE001 -- report and fall back
                print(f"[bench] symmetric-memory gather unavailable ({exc!r}); using NCCL", file=sys.stderr)
        if pg is None:
            gather_kind = ("one asynchronous NCCL all-gather per step of the payload the kernels wrote "
                           "(padded tokens + counts), overlapped with the next step")

    def gather_step(feats, masks):
        if pg is not None:
            peer, tok_view, cnt_view, step = pg.begin(slots)
            if push_mode:      # the push is enqueued while the host would otherwise idle waiting for the counts
                enc.forward_padded(feats, masks, ann, out=tok_view, counts_out=cnt_view,
                                   after_enqueue=lambda: pg.push(step, peer, m_pad_rows))
            else:
                enc.forward_padded(feats, masks, ann, out=tok_view, counts_out=cnt_view, peer=peer)
            pending.append(step)
            while len(pending) > 1:
                pending.popleft()
            return step
        payload, tok_view, cnt_view = sharding.new_payload(slots, pad_rows, pad_objs, 3584, torch.bfloat16, dev)
        enc.forward_padded(feats, masks, ann, out=tok_view, counts_out=cnt_view)   # counts also reach the host
        gathered, work = sharding.all_gather_payload(payload, async_op=True)
        pending.append((work, gathered, payload))
        while len(pending) > 2:
            pending.popleft()[0].wait()
        return gathered, work

    def drain():
        """Stream-ordered: every outstanding step's gathered result is complete on this rank."""
        while pending:
            item = pending.popleft()
            if pg is not None:
                pg.wait(item)
            else:
                item[0].wait()

    no_gather = world > 1 and os.environ.get("UFV_BENCH_GATHER") == "none"   # developer knob: ranks without collection

    def step_resident():
        if world > 1 and not no_gather:
            return gather_step(feats_dev, masks_dev)
        return enc(feats_dev, masks_dev, None, ann, None)[0]

    def step_e2e():
        if world > 1:
            res = gather_step(feats_host, masks_host)                           # H2D inside forward_padded
            if pg is not None:
                pg.wait(res)
                return pg.gathered(res).cpu()                                    # D2H of the gathered result
            res[1].wait()
            return res[0].cpu()
        tokens, counts = enc(feats_host, masks_host, None, ann, None)           # H2D inside forward
        return tokens.cpu()                                                      # D2H of the result

    if pg is not None:
        # self-check before timing: the fused gather equals an NCCL all-gather of the same payload
        step = gather_step(feats_dev, masks_dev)
        drain()
        torch.cuda.synchronize()
        got_t, got_c = sharding.unpack_padded(pg.gathered(step), pad_rows, pad_objs)
        payload, tok_view, cnt_view = sharding.new_payload(slots, pad_rows, pad_objs, 3584, torch.bfloat16, dev)
        enc.forward_padded(feats_dev, masks_dev, ann, out=tok_view, counts_out=cnt_view)
        ref, _ = sharding.all_gather_payload(payload)
        torch.cuda.synchronize()
        ref_t, ref_c = sharding.unpack_padded(ref, pad_rows, pad_objs)
        if got_c != ref_c or not torch.equal(got_t, ref_t):
            sys.exit("bench.py: fused all-gather disagrees with the NCCL all-gather")
        pg.check()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        drain()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    if a.timeline:
        with torch.inference_mode():
            write_timeline(a.timeline, step_resident, drain, barrier, rank, world)
        if world > 1:
            dist.destroy_process_group()
        return

    # the reference runs the model under torch.inference_mode() (ufvideo/__init__.py:122)
    with ClockSampler(local, enabled=rank == 0) as clocks, torch.inference_mode():
        ms_step = timed(step_resident, a.steps, max(a.warmup, 3))
        ms_e2e = timed(step_e2e, a.steps, max(a.warmup, 3))
    clock_summary = clocks.summary()

    # ---- roofline of the dominant kernel: segmented mask pool, timed alone ------------------------
    plan = packer.build_plan(masks_dev, ann, feats_dev.shape[0], k, dev)
    patches = layer.mask_to_patches(plan, dev)
    # each needed feature row once per FRAME (union over all its objects) + pooled write + bitmasks
    pool_bytes = packer.algorithmic_pool_bytes(plan, patches["bits"].cpu().numpy(), 1152, 2)
    ms_pool = timed(lambda: layer.mask_pool(feats_dev, plan, patches), max(a.steps, 20), 3)
    peak, peak_src = peaks()
    achieved = pool_bytes / (ms_pool * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "pool_traffic.json")
    if os.path.isfile(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")

    if pg is not None:
        pg.check()
    sweep = None
    if not a.no_sweep:
        pending.clear()
        torch.cuda.empty_cache()
        with torch.inference_mode():
            sweep = run_sweep(enc, dev, rank, world, peak)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        v, sec, cores, sample, done, kind, _ = cpu_reference_run(a, steps=3, warmup=1, min_seconds=10.0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}

    total_q = q * world
    # pinned host masks are read in place by kernel 1 (row mode): only the 2 x 27 source rows of each
    # mask cross PCIe, in 16-byte chunks over the column span of the taps
    mask_bytes_full = sum(m.numel() * 4 for m in masks_host)
    mask_bytes_read = 0
    for m in masks_host:
        t = packer.tap_table(m.shape[1], m.shape[2], 27, False).reshape(4, 27)
        rows = int((t[:2] >= 0).sum())
        cols = t[2:][t[2:] >= 0]
        span = (int(cols.max()) - int(cols.min()) + 1) * 4
        mask_bytes_read += m.shape[0] * rows * ((span + 15) // 16 * 16)
    h2d = feats_host.numel() * 2 + mask_bytes_read + int(plan.buffer.numel())
    if world > 1:
        tail_rows = sharding._padded_tail_rows(pad_objs, 3584 * 2)
        d2h = world * (pad_rows + tail_rows) * 3584 * 2 + n_obj * 4
    else:
        d2h = n_obj * k * 3584 * 2 + n_obj * 4
    line = {
        "metric": METRIC, "value": total_q / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(a), "object_frames_per_step": total_q,
                   "mask_family": a.family, "mask_dtype": "float32", "mask_hw": [h, w],
                   "e2e_inputs": (f"pinned host: features {feats_host.numel() * 2 / 1e6:.0f} MB copied H2D; masks "
                                  f"{mask_bytes_full / 1e6:.0f} MB read in place over PCIe by kernel 1, "
                                  f"{mask_bytes_read / 1e6:.0f} MB touched"),
                   "l2": f"inputs larger than L2: {feats_dev.numel() * 2 / 1e6:.0f} MB of features per step vs 126 MB",
                   "collective": gather_kind, "host_binding": numa_note,
                   "per_step_work": ("all five kernels run on the step's inputs every step (masks re-read, features "
                                     "re-streamed, weights re-read); only the host-side descriptor arrays of the batch "
                                     "structure and the captured launch sequence (CUDA graph) are reused across steps")},
        "e2e": {"value": total_q / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
        "gpu_launches": (6 if pg is not None else 5) * a.steps,   # kernels 1, 2, 3, 4a, 4b (+ peer push / flag wait)
        "roofline": {"kernel": "mask_pool_kernel<bf16>", "bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "algorithmic_bytes_per_launch": pool_bytes, "us_per_launch": ms_pool * 1e3,
                     "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0},
        "clocks": clock_summary,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if sweep is not None:
        line["sweep"] = sweep
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
