"""B200-native object (region) encoder: host-side mirror of the reference's
``ufvideo/model/layer.py``.

Same names, constructor, forward signature, parameter names and return types as the reference
(``MaskExtractor`` layer.py:50-128, ``MaskPooling`` :131-152, ``token_merge`` :6-33,
``build_region_encoder`` :155-161), so it drops in at ``videorefer_arch.py:39,92`` (construction)
and ``videorefer_arch.py:236`` (call).  All arithmetic runs in the hand-written sm_100a kernels
of ``libufv_b200.so`` through the C ABI in ``include/ufv_b200.h``; PyTorch only owns device
memory and the stream.  There is no CPU or eager fallback: tensors must end up on a CUDA device
and the shared library must be built.  Under ``torch.no_grad()`` / ``inference_mode`` the fused
inference path runs (one C call, replayed as a CUDA graph when a batch structure repeats); with
autograd on and trainable parameters or features, a differentiable path runs instead
(``_PoolMerge`` + the projector under torch autograd).
"""
from __future__ import annotations

import contextlib
import ctypes
import os

import numpy as np
import torch
import torch.nn as nn

from . import _cabi, packer

# forward() / forward_padded() replay their launch sequence as a CUDA graph from the second call of a
# batch structure on (UFV_NO_GRAPH=1 disables: developer A/B knob).
USE_CUDA_GRAPH = os.environ.get("UFV_NO_GRAPH") is None

_NULL_CONTEXT = contextlib.nullcontext()

_vp = ctypes.c_void_p
_ELEM_BYTES = {torch.float32: 4, torch.int32: 4, torch.bfloat16: 2, torch.float16: 2, torch.uint8: 1}


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what} must live on a CUDA device: ufvideo_b200 has no CPU path")


def _feat_dtype(t: torch.Tensor) -> int:
    try:
        return packer.FEAT_DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported feature dtype {t.dtype} (float32, bfloat16, float16)") from None


# ------------------------------------------------------------------------------------------------
# stage-level operators (each one C-ABI call); the parity tests drive these directly
# ------------------------------------------------------------------------------------------------
def mask_to_patches(plan: packer.EncodePlan, device, n_out: int = 27, want_idx: bool = False):
    """Kernel 1.  Returns a dict: bits int32 [q, 24], cnt int32 [q], idx int16 [q, 736] | None."""
    q = plan.n_masks
    bits = torch.empty((q, _cabi.BITS_WORDS), dtype=torch.int32, device=device)
    cnt = torch.empty((q,), dtype=torch.int32, device=device)
    idx = torch.zeros((q, 736), dtype=torch.int16, device=device) if want_idx else None
    d = plan.dev
    _cabi.check(_cabi.lib().ufv_mask_to_patches(
        d["mask_desc"], d["taps"], q, n_out, plan.any_row_mode, bits.data_ptr(), cnt.data_ptr(),
        idx.data_ptr() if want_idx else None, 736, _stream_ptr(device)))
    return {"bits": bits, "cnt": cnt, "idx": idx}


def mask_pool(feats: torch.Tensor, plan: packer.EncodePlan, patches: dict):
    """Kernel 2.  feats [F, n_patch, C] + the output of ``mask_to_patches`` -> pooled fp32 [q, C]."""
    _require_cuda(feats, "feats")
    f, n_patch, c = feats.shape
    pooled = torch.empty((plan.n_masks, c), dtype=torch.float32, device=feats.device)
    d = plan.dev
    _cabi.check(_cabi.lib().ufv_mask_pool(
        feats.data_ptr(), _feat_dtype(feats), f, n_patch, c, patches["bits"].data_ptr(), patches["cnt"].data_ptr(),
        d["grp_row"], d["grp_off"], d["grp_member"], plan.n_groups, plan.max_group,
        pooled.data_ptr(), _stream_ptr(feats.device)))
    return pooled


def ttm(pooled: torch.Tensor, plan: packer.EncodePlan, k_keep: int, out_dtype: torch.dtype,
        debug: bool = False):
    """Kernel 3.  Returns (tokens [m_pad, C] out_dtype, counts int32 [n_obj], extras dict)."""
    device = pooled.device
    c = pooled.shape[1]
    tokens = torch.empty((plan.m_pad, c), dtype=out_dtype, device=device)
    counts = torch.empty((plan.n_obj,), dtype=torch.int32, device=device)
    extras = {}
    f32 = cuts = None
    words = (plan.max_len + 31) // 32
    sims = extras["sims"] = torch.zeros((plan.n_obj, max(plan.max_len, 1)), dtype=torch.float32, device=device)
    if debug:
        f32 = extras["tokens_f32"] = torch.empty((plan.m_pad, c), dtype=torch.float32, device=device)
        cuts = extras["cuts"] = torch.zeros((plan.n_obj, words), dtype=torch.int32, device=device)
    d = plan.dev
    _cabi.check(_cabi.lib().ufv_ttm(
        pooled.data_ptr(), c, d["obj_start"], d["obj_len"], d["slot_off"], plan.n_obj, plan.max_len,
        k_keep, tokens.data_ptr(), packer.FEAT_DTYPES[out_dtype],
        f32.data_ptr() if debug else None, counts.data_ptr(), cuts.data_ptr() if debug else None,
        words, sims.data_ptr(), max(plan.max_len, 1), None, 0,
        _stream_ptr(device)))
    return tokens, counts, extras


def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, gelu: bool = False):
    """Kernel 4: act(x @ weight.T + bias) in x.dtype (tcgen05 for bf16 / fp16)."""
    _require_cuda(x, "x")
    if not (x.dtype == weight.dtype == bias.dtype):
        raise TypeError(f"linear: dtype mismatch x={x.dtype} weight={weight.dtype} bias={bias.dtype}")
    x = x.contiguous()
    m, k = x.shape
    n = weight.shape[0]
    y = torch.empty((m, n), dtype=x.dtype, device=x.device)
    lib = _cabi.lib()
    ws_bytes = int(lib.ufv_linear_ws_bytes(m, n, k, _feat_dtype(x)))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device) if ws_bytes else None   # split-K scratch
    _cabi.check(lib.ufv_linear(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), y.data_ptr(), m, n, k,
                               _feat_dtype(x), int(gelu), ws.data_ptr() if ws_bytes else None, ws_bytes,
                               _stream_ptr(x.device)))
    return y


def gather_rows(x: torch.Tensor, row_map: torch.Tensor):
    out = torch.empty((row_map.numel(), x.shape[1]), dtype=x.dtype, device=x.device)
    _cabi.check(_cabi.lib().ufv_gather_rows(x.data_ptr(), row_map.data_ptr(), out.data_ptr(),
                                            row_map.numel(), x.shape[1] * x.element_size(),
                                            _stream_ptr(x.device)))
    return out


def _await_counts(plan: packer.EncodePlan, run: dict, device, idle_work=None) -> np.ndarray:
    """Poll the pinned words kernel 3 writes, (epoch << 16) | count per object, until all of them
    carry this call's epoch; returns the int32 counts (``plan.slots`` itself -- do not modify -- when no
    object tied below its reserved count).  ``idle_work`` (optional callable) runs once first: the host has
    ~60 us to kill here, which is where the next call's output tensor gets allocated.

    The return of this function starts the critical path to the next call's first kernel (the projector of
    this call is still running), so the common case is ONE bytes compare per poll against the words a
    tie-free call must publish, computed before the polling starts."""
    words, epoch = run["counts_np"][:plan.n_obj], run["epoch"]
    if idle_work is not None:
        idle_work()
    want = (plan.slots | np.int32(epoch << 16)).tobytes()
    last = plan.n_obj - 1
    spins = 0
    while True:
        if words.tobytes() == want:              # every object published, none tied: the common case
            return plan.slots
        if (int(words[last]) >> 16) == epoch:    # cheap scalar pre-check before the vector compare
            snap = words.copy()
            if ((snap >> 16) == epoch).all():
                return snap & 0xffff
        spins += 1
        if spins & 0xffff == 0:                  # every few ms: surface a failed launch instead of hanging
            if torch.cuda.current_stream(device).query():
                snap = words.copy()
                if ((snap >> 16) == epoch).all():
                    return snap & 0xffff
                torch.cuda.synchronize(device)
                raise RuntimeError("ufvideo_b200: the merge kernel finished without publishing its counts")


def splice_regions(text_embeds: torch.Tensor, region_pos: torch.Tensor, tokens: torch.Tensor,
                   counts: torch.Tensor, plan: packer.EncodePlan, want_row_src: bool = False):
    """Device-side <region> splice (the consumer loop of videorefer_arch.py:300-311, flattened over
    the batch): ``text_embeds`` [n_text, hid] holds one placeholder row per object at ``region_pos``
    (int32, ascending, on the device); each is replaced by that object's tokens.  ``tokens`` /
    ``counts`` / ``plan`` are what ``encode_padded`` returned, so nothing here waits for the host copy
    of region_token_nums.  Returns (out [n_text - n_obj + m_pad, hid], out_len int32 device scalar,
    row_src or None); rows past out_len are unspecified."""
    _require_cuda(text_embeds, "text_embeds")
    n_text, hid = text_embeds.shape
    if tokens.shape[1] != hid or tokens.dtype != text_embeds.dtype:
        raise ValueError("tokens and text_embeds must share width and dtype")
    if region_pos.numel() != plan.n_obj or region_pos.dtype != torch.int32:
        raise ValueError(f"region_pos must be int32 [{plan.n_obj}]")
    text_embeds, tokens = text_embeds.contiguous(), tokens.contiguous()
    dev = text_embeds.device
    out = torch.empty((n_text - plan.n_obj + plan.m_pad, hid), dtype=text_embeds.dtype, device=dev)
    out_len = torch.empty((1,), dtype=torch.int32, device=dev)
    row_src = torch.full((out.shape[0],), 2 ** 31 - 1, dtype=torch.int32, device=dev) if want_row_src else None
    _cabi.check(_cabi.lib().ufv_splice_rows(
        text_embeds.data_ptr(), n_text, region_pos.data_ptr(), tokens.data_ptr(), plan.dev["slot_off"],
        counts.data_ptr(), plan.n_obj, plan.m_pad, out.data_ptr(), out_len.data_ptr(),
        row_src.data_ptr() if want_row_src else None, hid * text_embeds.element_size(), _stream_ptr(dev)))
    return out, out_len, row_src


class RegionLayout:
    """Where every row of the reference's ``new_input_embeds`` comes from (videorefer_arch.py:291-368), for a
    batch whose objects keep their reserved token counts (``slots`` = min(T_o, K): true unless the merge ties).

    seq_lens[i]    rows of sample i's embedded sequence (mm tokens already expanded by the caller)
    region_pos[i]  ascending local positions of its ``<region>`` placeholders; objects are consumed in order
                   over the batch.  A sample WITHOUT a placeholder still consumes one object whose tokens are
                   dropped -- the reference advances its region cursor by one there (videorefer_arch.py:263-264,
                   :300-305; that object is the dummy mask the collator adds).
    Built on the host with numpy: new_lens, l_max, src_map int32 [B * l_max] (text row index, -1 padding,
    -2 region token) and token_row_map int32 [m_pad] (destination row of every padded token row, -1 dropped)."""

    def __init__(self, seq_lens, region_pos, slots):
        seq_lens = [int(n) for n in seq_lens]
        slots = np.asarray(slots, dtype=np.int64)
        slot_off = np.concatenate([[0], np.cumsum(slots)])
        b = len(seq_lens)
        if len(region_pos) != b:
            raise ValueError("region_pos and seq_lens disagree on the number of samples")
        text_off = np.concatenate([[0], np.cumsum(seq_lens)])
        rows, obj = [], 0
        for i in range(b):
            pos = [int(p) for p in region_pos[i]]
            if any(p < 0 or p >= seq_lens[i] for p in pos) or pos != sorted(set(pos)):
                raise ValueError(f"sample {i}: region positions must be ascending indices into its sequence")
            if not pos:                                   # no placeholder: one object consumed, nothing inserted
                if obj >= slots.size:
                    raise ValueError("fewer objects than the samples consume")
                rows.append((np.arange(seq_lens[i]) + text_off[i], [], [obj]))
                obj += 1
                continue
            if obj + len(pos) > slots.size:
                raise ValueError("fewer objects than <region> placeholders")
            seq, ins = [], []
            last = 0
            for p in pos:
                seq.append(np.arange(last, p) + text_off[i])
                seq.append(np.full(int(slots[obj]), -2, dtype=np.int64))
                ins.append((obj, sum(len(x) for x in seq) - int(slots[obj])))      # (object, local start of its tokens)
                obj += 1
                last = p + 1
            seq.append(np.arange(last, seq_lens[i]) + text_off[i])
            rows.append((np.concatenate(seq), ins, []))
        if obj != slots.size:
            raise ValueError(f"{slots.size} objects but the samples consume {obj}")
        self.new_lens = [int(r[0].size) for r in rows]
        self.l_max = max(self.new_lens) if rows else 0
        self.batch = b
        src = np.full((b, self.l_max), -1, dtype=np.int32)
        tok = np.full(int(slot_off[-1]), -1, dtype=np.int32)
        for i, (seq, ins, _dropped) in enumerate(rows):
            src[i, :seq.size] = seq
            for o, start in ins:
                tok[slot_off[o]:slot_off[o + 1]] = i * self.l_max + start + np.arange(int(slots[o]))
        self.src_map = src.reshape(-1)
        self.token_row_map = tok
        self.n_text = int(text_off[-1])


# ------------------------------------------------------------------------------------------------
# reference-named API
# ------------------------------------------------------------------------------------------------
def token_merge(x: torch.Tensor, r: int) -> torch.Tensor:
    """Drop-in for the reference's ``token_merge(x, r)`` (layer.py:6-33): x fp32 [1, n, d] on a
    CUDA device, r = number of tokens to remove; returns [1, k, d], k <= n - r."""
    _require_cuda(x, "x")
    if x.dim() != 3 or x.shape[0] != 1:
        raise ValueError("token_merge expects x of shape [1, n, d] (the reference's own use)")
    n = x.shape[1]
    k_keep = n - r
    if not 1 <= k_keep < n:
        raise ValueError("r must satisfy 1 <= n - r < n")
    pooled = x[0].to(torch.float32).contiguous()
    host = {"obj_start": np.zeros(1, np.int32), "obj_len": np.full(1, n, np.int32),
            "slot_off": np.zeros(1, np.int32)}
    plan = packer.EncodePlan(n_masks=n, n_groups=0, max_group=1, n_obj=1, max_len=n, m_pad=k_keep,
                             slots=np.full(1, k_keep, np.int32), host=host)
    packer._upload(plan, x.device)
    tokens, counts, _ = ttm(pooled, plan, k_keep, torch.float32)
    return tokens[: int(counts.item())].unsqueeze(0).to(x.dtype)


class MaskPooling(nn.Module):
    """Drop-in for the reference's ``MaskPooling`` (layer.py:131-152)."""

    def forward(self, x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        """x [b, C, h, w] (the NCHW view the reference builds at layer.py:101, i.e. NHWC storage)
        and mask [1, q, H, W]  ->  fp32-accurate pooled [max(b, q), C] in x.dtype."""
        _require_cuda(x, "x")
        b, c, h, w = x.shape
        if h != w or h > _cabi.MAX_PATCH_SIDE:
            raise ValueError(f"patch grid {h}x{w} unsupported (square, side <= {_cabi.MAX_PATCH_SIDE})")
        nhwc = x.permute(0, 2, 3, 1)
        feats = (nhwc if nhwc.is_contiguous() else nhwc.contiguous()).reshape(b, h * w, c)
        if feats.dtype not in packer.FEAT_DTYPES:
            feats = feats.float()
        plan = packer.build_plan([mask[0]], [[list(range(b))]], b, 1, x.device, False, h)
        return mask_pool(feats, plan, mask_to_patches(plan, x.device, h)).to(x.dtype)


class MaskExtractor(nn.Module):
    """Drop-in for the reference's ``MaskExtractor`` (layer.py:50-128).

    State-dict keys are the reference's (``feat_linear.0.weight`` ...), so UFVideo checkpoints
    load unchanged (videorefer_arch.py:120-122).  ``region_token_num`` (K) defaults to 4 as in the
    reference and can be reassigned on the instance.
    """

    def __init__(self, image_aspect_ratio, config, mask_shape=112, depth=2, region_token_num=4):
        super().__init__()
        self.mask_shape = mask_shape              # stored, unused -- as in the reference (layer.py:53)
        self.mask_pooling = MaskPooling()
        modules = [nn.Linear(config.mm_hidden_size, config.hidden_size)]
        for _ in range(1, depth):
            modules.append(nn.GELU())
            modules.append(nn.Linear(config.hidden_size, config.hidden_size))
        self.feat_linear = nn.Sequential(*modules)
        self.image_aspect_ratio = image_aspect_ratio
        self.region_token_num = region_token_num
        self.last_plan = None                     # introspection for tests / bench
        self.keep_debug = False                   # expose intermediates of the last call as _debug

    # -- helpers -------------------------------------------------------------------------------
    def _linears(self):
        cached = self.__dict__.get("_lin_cache")
        if cached is None or cached[0] is not self.feat_linear or cached[1] != len(self.feat_linear):
            cached = (self.feat_linear, len(self.feat_linear),
                      [m for m in self.feat_linear if isinstance(m, nn.Linear)])
            self.__dict__["_lin_cache"] = cached
        return cached[2]

    def encode_padded(self, feats, masks, ann_indices, out=None, counts_out=None, peer=None, _awaited=False,
                      scatter=None):
        """Kernels 1-4 without the host read-back: returns (tokens [m_pad, hid], counts int32
        [n_obj] on the device, plan).  Row r of an object is valid iff r < counts[object].
        ``out`` / ``counts_out`` let the projector and the merge kernel write straight into caller
        memory (e.g. the all-gather payload of ``sharding.new_payload``): contiguous [m_pad, hid]
        of the model dtype and int32 [n_obj] on the module's device.  ``peer`` (a ``_cabi.PeerArgs``
        from ``sharding.PeerGather.begin``) fuses the result-collection all-gather into the last
        Linear: its tiles are stored into every rank's gathered buffer and ``out`` is not written.

        Host work per call is kept to the plan lookup and ONE C call: the workspace, the argument
        struct and the pinned counts buffer live in the cached plan and are reused as long as the
        same pointers come back on the same stream (stream order makes the reuse safe)."""
        linears = self._linears()
        two = len(linears) == 2
        if two:                                   # parameter tensors straight from the module dicts (no __getattr__)
            p0, p1 = linears[0]._parameters, linears[1]._parameters
            w1 = p0["weight"]
            wptrs = (w1.data_ptr(), p0["bias"].data_ptr(), p1["weight"].data_ptr(), p1["bias"].data_ptr())
        else:
            w1 = linears[0].weight
            wptrs = tuple(p.data_ptr() for lin in linears for p in (lin.weight, lin.bias))
        device = w1.device
        if device.type != "cuda":
            raise RuntimeError("MaskExtractor parameters must be on a CUDA device (no CPU path)")
        if not torch.is_tensor(feats):
            raise TypeError("feats must be a tensor [F, n_patch, C]")
        stream = torch._C._cuda_getCurrentRawStream(device.index)
        # everything derived from (features, parameters, stream) is memoised: a repeated call pays one tuple compare
        memo = self.__dict__.get("_feat_memo")
        key = (feats.data_ptr(), feats.shape, feats.dtype, feats.device, feats.is_contiguous(), wptrs, stream)
        if memo is not None and memo[0] == key:
            dt, f, c, side, hid, sig = memo[1]
        else:
            if feats.device != device:
                feats = feats.to(device, non_blocking=True)
            if not feats.is_contiguous():
                feats = feats.contiguous()
            if feats.dim() != 3:
                raise ValueError(f"feats must be [F, n_patch, C], got {tuple(feats.shape)}")
            dt = _feat_dtype(feats)
            if w1.dtype != feats.dtype:
                raise TypeError(f"feature dtype {feats.dtype} != projector dtype {w1.dtype}")
            f, n_patch, c = feats.shape
            side = int(round(n_patch ** 0.5))     # layer.py:100
            if side * side != n_patch or side > _cabi.MAX_PATCH_SIDE:
                raise ValueError(f"n_patch={n_patch} is not a square grid with side <= {_cabi.MAX_PATCH_SIDE}")
            hid = linears[-1].weight.shape[0]
            sig = (feats.data_ptr(), dt, f, c, hid, stream, two, wptrs)
            if feats.data_ptr() == key[0]:        # the caller's own tensor was used as is: safe to memoise
                self.__dict__["_feat_memo"] = (key, (dt, f, c, side, hid, sig))
        k_keep = int(self.region_token_num)
        plan = packer.build_plan(masks, ann_indices, f, k_keep, device,
                                 self.image_aspect_ratio == "pad", side)
        self.last_plan = plan
        q, m_pad = plan.n_masks, plan.m_pad
        # Run state (workspace, argument struct, pinned counts words, captured graphs) belongs to ONE
        # (module, stream) pair: two modules or two streams that meet on the same cached plan neither rebind
        # nor race on each other's buffers.  plan.run is the state of the latest call (introspection).
        run_key = (id(self), stream)
        run = plan.runs.get(run_key)
        if run is None or run["sig"] != sig:
            shape_sig = (dt, f, c, hid, two, k_keep, side)
            if run is not None and run["shape_sig"] == shape_sig:
                # same sizes, other pointers: keep the workspace, rebind, drop captured graphs
                packer.release_run(run)
                run.pop("spare_out", None)
                run["sig"] = sig
                if two:
                    a = run["args"]
                    a.feats = feats.data_ptr()
                    a.w1, a.b1 = linears[0].weight.data_ptr(), linears[0].bias.data_ptr()
                    a.w2, a.b2 = linears[1].weight.data_ptr(), linears[1].bias.data_ptr()
            else:
                if run is not None:
                    packer.release_run(run)
                run = plan.runs[run_key] = self._prepare_run(plan, sig, feats, linears, side, k_keep, device)
                run["shape_sig"] = shape_sig
                while len(plan.runs) > packer.RUNS_PER_PLAN:      # oldest (module, stream) pair goes first
                    packer.release_run(plan.runs.pop(next(iter(plan.runs))))
        plan.run = self._last_run = run
        if peer is not None and not two:
            raise ValueError("peer gather needs the depth-2 projector path")
        if scatter is not None:
            # scatter epilogue: the last Linear stores token row r as row row_map[r] of the caller's buffer
            # (``scatter`` = (buffer [rows, hid], int32 row map [m_pad] on the device)); validated by the caller
            if not two or peer is not None or out is not None:
                raise ValueError("scatter needs the depth-2 projector path and excludes out= / peer=")
            tokens = scatter[0]
        elif out is None:
            tokens = run.pop("spare_out", None)   # allocated while the previous call waited for its counts
            if tokens is None or tokens.dtype != feats.dtype:
                tokens = torch.empty((m_pad, hid), dtype=feats.dtype, device=device)
        else:
            tokens = out
        if out is not None or counts_out is not None:
            # caller-provided destinations (e.g. the 8 ring slots of sharding.PeerGather): validated once each
            seen = (out.data_ptr() if out is not None else 0, counts_out.data_ptr() if counts_out is not None else 0)
            if seen not in run["validated"]:
                if out is not None and (tuple(out.shape) != (m_pad, hid) or out.dtype != feats.dtype
                                        or out.device != device or not out.is_contiguous() or out.data_ptr() % 16):
                    raise ValueError(f"out must be a contiguous, 16-byte aligned [{m_pad}, {hid}] {feats.dtype} "
                                     f"tensor on {device}")
                if counts_out is not None:
                    if (counts_out.numel() != plan.n_obj or counts_out.dtype != torch.int32
                            or counts_out.device != device or not counts_out.is_contiguous()):
                        raise ValueError(f"counts_out must be a contiguous int32 [{plan.n_obj}] tensor on {device}")
                    if not two:
                        raise ValueError("counts_out needs the depth-2 projector path")
                if len(run["validated"]) > 64:
                    run["validated"].clear()
                run["validated"].add(seen)
        lib = _cabi.lib()
        ptr = run["ptr"]
        d = plan.dev
        if two:                                   # the reference's depth=2 projector: one chained call
            epoch = run["epoch"] = run["epoch"] % 32767 + 1        # 1 .. 32767, the tag of this call's counts
            graph = None
            row_map_ptr = scatter[1].data_ptr() if scatter is not None else None
            if _awaited and USE_CUDA_GRAPH and dt != _cabi.UFV_F32 and q > 0 and plan.n_obj > 0 and m_pad > 0:
                # A caller that waits for the counts before its next call (forward, forward_padded) lets
                # the launch sequence be replayed as a CUDA graph: the per-call values travel through the
                # pinned block (kernel 1 forwards it to the device), everything else is constant.
                run["awaited_calls"] += 1
                gkey = (peer is not None, row_map_ptr)           # the row map is a kernel parameter of the capture
                graph = run["graphs"].get(gkey)
                if graph is None and run["awaited_calls"] >= 2:
                    graph = self._capture_graph(run, peer, row_map_ptr)
            if graph is not None:
                dyn = run["dyn"]
                dyn.tokens_out = tokens.data_ptr()
                dyn.counts_out = 0 if counts_out is None else counts_out.data_ptr()
                dyn.epoch = epoch
                if peer is not None:
                    ctypes.memmove(run["dyn_peer_addr"], ctypes.addressof(peer), ctypes.sizeof(_cabi.PeerArgs))
                rc = lib.ufv_encode_graph_launch(graph, stream)
                if rc:
                    _cabi.check(rc)
            else:
                a = run["args"]
                a.tokens_out = tokens.data_ptr()
                a.tokens_row_map = row_map_ptr
                a.counts = ptr["counts"] if counts_out is None else counts_out.data_ptr()
                if peer is None:
                    a.peer = None
                else:
                    ref = getattr(peer, "_as_pointer", None)
                    if ref is None:
                        ref = peer._as_pointer = ctypes.pointer(peer)
                    a.peer = ref
                a.epoch = epoch
                _cabi.check(lib.ufv_encode(run["args_ref"], stream))
        else:                                     # other depths: the same kernels, staged
            _cabi.check(lib.ufv_mask_to_patches(d["mask_desc"], d["taps"], q, side, plan.any_row_mode,
                                                ptr["bits"], ptr["cnt"], None, 0, stream))
            _cabi.check(lib.ufv_mask_pool(feats.data_ptr(), dt, f, side * side, c, ptr["bits"], ptr["cnt"], d["grp_row"],
                                          d["grp_off"], d["grp_member"], plan.n_groups, plan.max_group,
                                          ptr["pooled"], stream))
            _cabi.check(lib.ufv_ttm(ptr["pooled"], c, d["obj_start"], d["obj_len"], d["slot_off"],
                                    plan.n_obj, plan.max_len, k_keep, ptr["merged"], dt, None,
                                    ptr["counts"], None, 0, ptr["sims"], max(plan.max_len, 1), None, 0, stream))
            x = run["view"]("merged", feats.dtype, (m_pad, c))
            for i, lin in enumerate(linears):
                x = linear(x, lin.weight, lin.bias, gelu=i < len(linears) - 1)
            tokens = x
        counts = run["counts"] if counts_out is None else counts_out
        if self.keep_debug:
            view = run["view"]
            self._debug = {"bits": view("bits", torch.int32, (q, _cabi.BITS_WORDS)),
                           "cnt": view("cnt", torch.int32, (q,)),
                           "pooled": view("pooled", torch.float32, (q, c)),
                           "merged": view("merged", feats.dtype, (m_pad, c))}
        return tokens, counts, plan

    def _prepare_run(self, plan, sig, feats, linears, side, k_keep, device):
        """Workspace + argument struct of one (plan, pointers, stream) combination."""
        _, dt, f, c, hid, stream, two, _ = sig
        q, m_pad = plan.n_masks, plan.m_pad
        es = feats.element_size()
        # one workspace allocation, carved into 256-byte aligned pieces
        lib = _cabi.lib()
        gemm_ws = max(int(lib.ufv_linear_ws_bytes(m_pad, hid, c, dt)), int(lib.ufv_linear_ws_bytes(m_pad, hid, hid, dt))) \
            if two and m_pad else 0
        sizes = (("bits", q * _cabi.BITS_WORDS * 4), ("cnt", q * 4), ("pooled", q * c * 4), ("gemm_ws", gemm_ws),
                 ("merged", m_pad * c * es), ("hidden", m_pad * hid * es), ("counts", plan.n_obj * 4),
                 ("sims", plan.n_obj * max(plan.max_len, 1) * 4), ("dyn", 256))
        off, total = {}, 0
        for name, nbytes in sizes:
            off[name] = total
            total += (nbytes + 255) // 256 * 256
        ws = torch.empty((max(total, 256),), dtype=torch.uint8, device=device)
        base = ws.data_ptr()
        ptr = {name: base + o for name, o in off.items()}

        def view(name, dtype, shape):
            n = int(np.prod(shape)) * _ELEM_BYTES[dtype]
            return ws[off[name]:off[name] + n].view(dtype).view(shape)

        run = {"sig": sig, "ws": ws, "ptr": ptr, "view": view,
               "counts": view("counts", torch.int32, (plan.n_obj,)), "graphs": {}, "awaited_calls": 0,
               "validated": set(), "epoch": 0}
        if two:
            d = plan.dev
            # pinned int32 [n_obj]: (epoch << 16) | count per object, written by kernel 3, polled by the host
            run["counts_pinned"] = torch.zeros((max(plan.n_obj, 1),), dtype=torch.int32).pin_memory()
            run["counts_np"] = run["counts_pinned"].numpy()
            counts_dev_addr = packer._device_address(run["counts_pinned"])
            a = _cabi.EncodeArgs(
                feats=feats.data_ptr(), feat_dtype=dt, n_patch_side=side, n_rows=f, c=c, hid=hid,
                mask_desc=d["mask_desc"], taps=d["taps"], n_masks=q, idx_pitch=0,
                any_row_mode=plan.any_row_mode, bits=ptr["bits"],
                cnt=ptr["cnt"], idx=None, grp_row=d["grp_row"],
                grp_off=d["grp_off"], grp_member=d["grp_member"], n_groups=plan.n_groups,
                max_group=plan.max_group, pooled=ptr["pooled"], obj_start=d["obj_start"],
                obj_len=d["obj_len"], slot_off=d["slot_off"], n_obj=plan.n_obj, max_len=plan.max_len,
                k_keep=k_keep, m_pad=m_pad, merged=ptr["merged"], counts=ptr["counts"],
                sims=ptr["sims"], sims_pitch=max(plan.max_len, 1),
                counts_host=counts_dev_addr, epoch=0,
                w1=linears[0].weight.data_ptr(), b1=linears[0].bias.data_ptr(),
                w2=linears[1].weight.data_ptr(), b2=linears[1].bias.data_ptr(),
                hidden=ptr["hidden"], tokens_out=None,
                gemm_ws=ptr["gemm_ws"] if gemm_ws else None, gemm_ws_bytes=gemm_ws)
            run["args"] = a
            run["args_ref"] = ctypes.byref(a)
            # per-call block of the graph-replay mode: pinned, written by the host, forwarded by kernel 1
            pinned = torch.zeros(64, dtype=torch.int32).pin_memory()
            run["dyn_pinned"] = pinned
            run["dyn"] = _cabi.DynArgs.from_address(pinned.data_ptr())
            run["dyn_peer_addr"] = pinned.data_ptr() + _cabi.DynArgs.peer.offset
            run["dyn_src_addr"] = packer._device_address(pinned)
        return run

    def _capture_graph(self, run, peer, row_map_ptr=None):
        """Capture the launch sequence of this run once (include/ufv_b200.h: ufv_encode_graph_create)."""
        a = run["args"]
        g = _cabi.EncodeArgs.from_buffer_copy(a)
        g.dyn_src, g.dyn_dev = run["dyn_src_addr"], run["ptr"]["dyn"]
        g.counts = run["ptr"]["counts"]
        g.tokens_out = None
        g.tokens_row_map = row_map_ptr
        g.peer = ctypes.pointer(peer) if peer is not None else None    # selects the kernel variant only
        handle = ctypes.c_void_p()
        _cabi.check(_cabi.lib().ufv_encode_graph_create(ctypes.byref(g), ctypes.byref(handle)))
        if len(run["graphs"]) >= 8:                                    # bounded: the oldest capture goes
            _cabi.lib().ufv_encode_graph_destroy(run["graphs"].pop(next(iter(run["graphs"]))))
        run["graphs"][(peer is not None, row_map_ptr)] = handle
        run.setdefault("graph_args", []).append(g)
        return handle

    def _on_module_device(self):
        """Context that makes the module's GPU the current CUDA device (the C ABI launches on the current
        device; torch ops guard themselves, raw launches do not).  A no-op in the common single-GPU case."""
        dev = self._linears()[0]._parameters["weight"].device
        if dev.type == "cuda" and torch.cuda.current_device() != dev.index:
            return torch.cuda.device(dev)
        return _NULL_CONTEXT

    def forward_padded(self, feats, masks, ann_indices, out=None, counts_out=None, peer=None, after_enqueue=None):
        """forward() without the compaction: (tokens [m_pad, hidden] with object o's rows at
        plan.host['slot_off'][o], region_token_nums as an int32 numpy array read back from the merge
        kernel, plan).  What the clip-sharded driver gathers (sharding.all_gather_payload).
        ``after_enqueue`` (optional callable) runs right after the kernels have been enqueued and before the
        host starts waiting for the counts -- ~60 us in which the host is otherwise idle, e.g. to enqueue the
        side-stream push of the result (sharding.PeerGather.push) without delaying the next call."""
        with self._on_module_device():
            tokens, counts, plan = self.encode_padded(feats, masks, ann_indices, out, counts_out, peer, _awaited=True)
            if after_enqueue is not None:
                after_enqueue()
            run = self._last_run
            if run.get("args") is not None and plan.n_obj > 0:
                return tokens, _await_counts(plan, run, tokens.device).copy(), plan
            return tokens, counts.cpu().numpy(), plan

    def forward_into(self, feats, masks, ann_indices, text_embeds, seq_lens, region_pos, labels=None,
                     ignore_index: int = -100):
        """The path plus its consumer: builds the reference's ``new_input_embeds`` /
        ``new_labels`` / attention mask of ``prepare_inputs_labels_for_multimodal`` (videorefer_arch.py:291-368)
        for the region part in one go.  The last Linear of the projector stores every object token straight at its
        ``<region>`` position inside the padded [B, L_max, hidden] batch (scatter epilogue); text rows, padding,
        labels and the attention mask are laid down by one small kernel.  No tokens tensor, no per-sample cat.

        text_embeds [sum(seq_lens), hidden]  the samples' embedded sequences back to back (model dtype, device)
        seq_lens, region_pos                 see ``RegionLayout``
        labels (optional) int64 [sum(seq_lens)] -> new labels [B, L_max] with ``ignore_index`` on region and
        padding rows.  Returns (inputs_embeds [B, L_max, hidden], labels or None, attention_mask bool [B, L_max],
        region_token_nums list[int]).  If the merge ties (an object keeps fewer tokens than reserved) the layout is
        rebuilt from the real counts on the slow path, as the reference's python loop would produce it."""
        with self._on_module_device():
            linears = self._linears()
            device = linears[0]._parameters["weight"].device
            _require_cuda(text_embeds, "text_embeds")
            if not torch.is_tensor(feats) or feats.dim() != 3:
                raise ValueError("feats must be a tensor [F, n_patch, C]")
            side = int(round(feats.shape[1] ** 0.5))
            k_keep = int(self.region_token_num)
            plan = packer.build_plan(masks, ann_indices, feats.shape[0], k_keep, device,
                                     self.image_aspect_ratio == "pad", side)
            hid = linears[-1].weight.shape[0]
            if text_embeds.dim() != 2 or text_embeds.shape[1] != hid or text_embeds.dtype != linears[0].weight.dtype:
                raise ValueError(f"text_embeds must be [rows, {hid}] of the projector dtype")
            text_embeds = text_embeds.contiguous()
            lkey = (tuple(int(n) for n in seq_lens), tuple(tuple(int(p) for p in r) for r in region_pos))
            layouts = plan.layouts
            hit = layouts.get(lkey)
            if hit is None:
                lay = RegionLayout(seq_lens, region_pos, plan.slots)
                if lay.n_text != text_embeds.shape[0]:
                    raise ValueError("text_embeds rows != sum(seq_lens)")
                maps = torch.from_numpy(np.concatenate([lay.src_map, lay.token_row_map])).to(device)
                if len(layouts) > 32:
                    layouts.clear()
                hit = layouts[lkey] = (lay, maps[:lay.src_map.size], maps[lay.src_map.size:])
            lay, src_map, row_map = hit
            n_rows = lay.batch * lay.l_max
            out = torch.empty((n_rows, hid), dtype=text_embeds.dtype, device=device)
            new_labels = torch.empty((n_rows,), dtype=torch.int64, device=device) if labels is not None else None
            attn = torch.empty((n_rows,), dtype=torch.uint8, device=device)
            if labels is not None:
                labels = labels.to(device=device, dtype=torch.int64).contiguous()
            _cabi.check(_cabi.lib().ufv_splice_static(
                text_embeds.data_ptr(), labels.data_ptr() if labels is not None else None, src_map.data_ptr(),
                out.data_ptr(), new_labels.data_ptr() if labels is not None else None, attn.data_ptr(), n_rows,
                hid * text_embeds.element_size(), int(ignore_index), _stream_ptr(device)))
            two = len(linears) == 2
            if two and plan.m_pad > 0:
                _, counts, plan = self.encode_padded(feats, masks, ann_indices, _awaited=True, scatter=(out, row_map))
                run = self._last_run
                nums = _await_counts(plan, run, device) if plan.n_obj > 0 else np.zeros(0, np.int32)
            else:                                             # other projector depths: no scatter epilogue
                tokens, nums_list = self._forward_impl(feats, masks, ann_indices)
                nums = np.asarray(nums_list, dtype=np.int32)
                if nums.tobytes() == plan.slots_bytes:
                    live = row_map >= 0
                    out.index_copy_(0, row_map[live].long(), tokens[live])
            if nums.tobytes() != plan.slots_bytes:
                # ties: some object kept fewer tokens than reserved -> every later row moves; rebuild from the counts
                tokens, nums_list = self._forward_impl(feats, masks, ann_indices)
                lay = RegionLayout(seq_lens, region_pos, nums_list)
                n_rows = lay.batch * lay.l_max
                src_map = torch.from_numpy(lay.src_map).to(device)
                out = torch.empty((n_rows, hid), dtype=text_embeds.dtype, device=device)
                new_labels = torch.empty((n_rows,), dtype=torch.int64, device=device) if labels is not None else None
                attn = torch.empty((n_rows,), dtype=torch.uint8, device=device)
                _cabi.check(_cabi.lib().ufv_splice_static(
                    text_embeds.data_ptr(), labels.data_ptr() if labels is not None else None, src_map.data_ptr(),
                    out.data_ptr(), new_labels.data_ptr() if labels is not None else None, attn.data_ptr(), n_rows,
                    hid * text_embeds.element_size(), int(ignore_index), _stream_ptr(device)))
                dest = torch.from_numpy(lay.token_row_map).to(device)
                live = dest >= 0
                out.index_copy_(0, dest[live].long(), tokens[live])
                nums = np.asarray(nums_list, dtype=np.int32)
            shape = (lay.batch, lay.l_max)
            return (out.view(*shape, hid), new_labels.view(shape) if new_labels is not None else None,
                    attn.view(shape).bool(), [int(n) for n in nums])

    def _forward_with_grad(self, feats, masks, ann_indices):
        """Training path (the region encoder is trainable in the reference, videorefer_arch.py:94-96):
        kernels 1-3 through ``_PoolMerge`` (differentiable w.r.t. ``feats``), then the projector through
        its own ``nn.Sequential`` so that autograd records it (cuBLAS GEMMs; the fused tcgen05 projector
        is the inference path)."""
        linears = self._linears()
        device = linears[0].weight.device
        if device.type != "cuda":
            raise RuntimeError("MaskExtractor parameters must be on a CUDA device (no CPU path)")
        feats = feats.to(device).contiguous()
        f, n_patch, c = feats.shape
        side = int(round(n_patch ** 0.5))
        k_keep = int(self.region_token_num)
        plan = packer.build_plan(masks, ann_indices, f, k_keep, device, self.image_aspect_ratio == "pad", side)
        self.last_plan = plan
        merged, counts = _PoolMerge.apply(feats, plan, k_keep, feats.dtype)
        nums = counts.cpu().numpy()
        if nums.tobytes() != plan.slots_bytes:                    # merge ties: drop the zero-filled slots
            keep = np.concatenate([np.arange(s, s + n) for s, n in zip(plan.host["slot_off"], nums)])
            merged = merged[torch.from_numpy(keep).to(device)]
        if (len(linears) == 2 and merged.dtype in (torch.bfloat16, torch.float16) and merged.shape[0] > 0
                and os.environ.get("UFV_TORCH_PROJECTOR_BACKWARD") is None):
            # forward and backward of the projector on the tcgen05 kernel (UFV_TORCH_PROJECTOR_BACKWARD=1: developer
            # A/B knob that puts the projector back under torch autograd / cuBLAS)
            l0, l1 = linears
            return _Projector.apply(merged, l0.weight, l0.bias, l1.weight, l1.bias), [int(n) for n in nums]
        return self.feat_linear(merged), [int(n) for n in nums]

    # -- the reference's forward -----------------------------------------------------------------
    def forward(self, feats, masks, X_features, ann_indices, frame_nums):
        """Same contract as layer.py:63-128: returns (mask_feats [N_tok, hidden], region_token_nums
        list[int]).  ``X_features`` and ``frame_nums`` are accepted and ignored, as in the reference
        (which reads only ``X_features.device`` in its fallbacks).  Under autograd with trainable
        parameters or features the differentiable path runs (``_forward_with_grad``)."""
        with self._on_module_device():
            return self._forward_impl(feats, masks, ann_indices)

    def _forward_impl(self, feats, masks, ann_indices):
        """Same contract as layer.py:63-128: returns (mask_feats [N_tok, hidden], region_token_nums
        list[int]).  ``X_features`` and ``frame_nums`` are accepted and ignored, as in the reference
        (which reads only ``X_features.device`` in its fallbacks).  Forward only: no autograd graph
        is recorded ("next")."""
        if torch.is_grad_enabled() and ((torch.is_tensor(feats) and feats.requires_grad) or any(
                p is not None and p.requires_grad for lin in self._linears() for p in lin._parameters.values())):
            return self._forward_with_grad(feats, masks, ann_indices)
        tokens, counts, plan = self.encode_padded(feats, masks, ann_indices, _awaited=True)
        # the one unavoidable D2H: the caller slices rows by these counts (videorefer_arch.py:307-311)
        run = self._last_run
        if run.get("args") is not None and plan.n_obj > 0:
            # the merge kernel stores the counts straight into pinned host memory, tagged with the call's
            # epoch; the projector may still be running when this returns
            region_token_nums = _await_counts(
                plan, run, tokens.device,
                lambda: run.__setitem__("spare_out", torch.empty_like(tokens)) if "spare_out" not in run else None)
        else:
            region_token_nums = counts.cpu().numpy()
        if region_token_nums is plan.slots or region_token_nums.tobytes() == plan.slots_bytes:
            return tokens, list(plan.expect_counts)           # no ties: every object kept min(T, K) tokens
        # ties at the merge threshold left some object with fewer than min(T, K) tokens:
        # drop the zero-filled slots (rare; exact ties only)
        nums = region_token_nums.tolist()
        packed = torch.empty((sum(nums), tokens.shape[1]), dtype=tokens.dtype, device=tokens.device)
        _cabi.check(_cabi.lib().ufv_compact_rows(
            tokens.data_ptr(), plan.dev["slot_off"], counts.data_ptr(), plan.n_obj, packed.data_ptr(),
            tokens.shape[1] * tokens.element_size(), _stream_ptr(tokens.device)))
        return packed, nums


class _PoolMerge(torch.autograd.Function):
    """Kernels 1-3 as one differentiable op: feats -> merged tokens [m_pad, C] (model dtype).

    Forward runs the CUDA kernels.  Backward (training parity) is the exact adjoint
    of the forward arithmetic: the run means with a few small torch index ops, the masked mean -- the
    part that touches a tensor the size of the features -- with ``ufv_mask_pool_backward``.  The merge
    decisions are piecewise constant, exactly as under the reference's own autograd (its comparisons
    and topk carry no gradient either, layer.py:17-27)."""

    @staticmethod
    def forward(ctx, feats, plan, k_keep, out_dtype):
        dev = feats.device
        patches = mask_to_patches(plan, dev, int(round(feats.shape[1] ** 0.5)))
        pooled = mask_pool(feats, plan, patches)
        merged, counts, extras = ttm(pooled, plan, k_keep, out_dtype, debug=True)
        ctx.plan, ctx.k_keep, ctx.feat_shape, ctx.feat_dtype = plan, k_keep, feats.shape, feats.dtype
        ctx.save_for_backward(patches["bits"], patches["cnt"], extras["cuts"], counts)
        ctx.mark_non_differentiable(counts)
        return merged, counts

    @staticmethod
    def backward(ctx, d_merged, _d_counts):
        bits, cnt, cuts, counts = ctx.saved_tensors
        plan, k_keep = ctx.plan, ctx.k_keep
        dev = d_merged.device
        h = plan.host
        n_obj, q = plan.n_obj, plan.n_masks
        obj_len = torch.from_numpy(h["obj_len"]).to(dev).long()
        obj_start = torch.from_numpy(h["obj_start"]).to(dev).long()
        slot_off = torch.from_numpy(h["slot_off"]).to(dev).long()
        max_len = max(plan.max_len, 1)
        # run id of token t of object o: number of cuts before t (merged objects) or t itself (T <= K)
        t_idx = torch.arange(max_len, device=dev)
        words = cuts.view(n_obj, -1).to(torch.int64) & 0xffffffff
        cut = ((words[:, t_idx // 32] >> (t_idx % 32)) & 1)                     # [n_obj, max_len]
        gid = torch.cumsum(cut, 1) - cut
        gid = torch.where((obj_len <= k_keep)[:, None], t_idx[None, :].expand(n_obj, -1), gid)
        valid = t_idx[None, :] < obj_len[:, None]
        o_of = torch.arange(n_obj, device=dev)[:, None].expand(-1, max_len)[valid]
        t_of = t_idx[None, :].expand(n_obj, -1)[valid]
        g_of = gid[valid]
        pooled_row = obj_start[o_of] + t_of
        token_row = slot_off[o_of] + g_of
        run_size = torch.zeros(d_merged.shape[0], device=dev).index_add_(0, token_row, torch.ones_like(g_of, dtype=torch.float32))
        d_pooled = torch.zeros((q, d_merged.shape[1]), dtype=torch.float32, device=dev)
        d_pooled[pooled_row] = d_merged.float()[token_row] / run_size[token_row][:, None]
        # masked mean: pooled[j] = sum_{p on} feats[row_j, p] / (cnt_j + 1e-8)
        denorm = cnt.float() + 1e-8
        w = (d_pooled / denorm[:, None]).contiguous()
        # adjoint of the masked mean: one streaming kernel writes d_feats (ufv_mask_pool_backward)
        go, gm, gr = h["grp_off"], h["grp_member"], h["grp_row"]
        n_rows, n_patch, c = ctx.feat_shape
        frame_of = np.empty(q, dtype=np.int64)
        for g in range(plan.n_groups):
            frame_of[gm[go[g]:go[g + 1]]] = int(gr[g])
        order = np.argsort(frame_of, kind="stable").astype(np.int32)          # object-frames grouped by feature row
        per_row = np.bincount(frame_of, minlength=n_rows)
        row_off = np.concatenate([[0], np.cumsum(per_row)]).astype(np.int32)
        meta = torch.from_numpy(np.concatenate([row_off, order])).to(dev)
        d_feats = torch.empty(ctx.feat_shape, dtype=ctx.feat_dtype, device=dev)
        _cabi.check(_cabi.lib().ufv_mask_pool_backward(
            w.data_ptr(), bits.data_ptr(), meta.data_ptr(), meta.data_ptr() + 4 * row_off.size,
            n_rows, int(per_row.max()) if q else 0, n_patch, c, d_feats.data_ptr(),
            packer.FEAT_DTYPES[ctx.feat_dtype], _stream_ptr(dev)))
        return d_feats, None, None, None


def _linear_ex(x, w, bias, epilogue=0, aux=None):
    """ufv_linear_ex: epilogue(x @ w.T + bias) on the tcgen05 kernel (bf16 / fp16); see include/ufv_b200.h."""
    m, k = x.shape
    n = w.shape[0]
    y = torch.empty((m, n), dtype=x.dtype, device=x.device)
    _cabi.check(_cabi.lib().ufv_linear_ex(
        x.data_ptr(), w.data_ptr(), bias.data_ptr() if bias is not None else None, y.data_ptr(), m, n, k,
        _feat_dtype(x), epilogue, aux.data_ptr() if aux is not None else None, _stream_ptr(x.device)))
    return y


def _transpose16(x, pad_to=8):
    """[r, c] -> [c, ceil(r / pad_to) * pad_to] (zero-padded columns), 2-byte elements, one kernel."""
    r, c = x.shape
    pitch = -(-r // pad_to) * pad_to
    out = torch.empty((c, pitch), dtype=x.dtype, device=x.device)
    _cabi.check(_cabi.lib().ufv_transpose16(x.data_ptr(), out.data_ptr(), r, c, pitch, _stream_ptr(x.device)))
    return out


def _colsum(x):
    out = torch.empty((x.shape[1],), dtype=x.dtype, device=x.device)
    _cabi.check(_cabi.lib().ufv_colsum(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], _feat_dtype(x),
                                       _stream_ptr(x.device)))
    return out


class _Projector(torch.autograd.Function):
    """feat_linear (Linear -> GELU -> Linear, layer.py:55-59) for training in bf16 / fp16, forward AND backward on
    the tcgen05 kernel.  Forward keeps the rounded pre-activation Z1 (written by the
    GELU epilogue) and H.  Backward: dZ1 = (dY . W2) * GELU'(Z1) in one GEMM with the GELU-backward epilogue,
    dW2 = dY^T . H, dX = dZ1 . W1, dW1 = dZ1^T . X, bias gradients by column sums.  The kernel contracts K-major
    operands, so W2, W1, dY, H, dZ1 and X are transposed by a small kernel first (the weights: 34 MB, ~15 us)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        x = x.contiguous()
        z1 = torch.empty((x.shape[0], w1.shape[0]), dtype=x.dtype, device=x.device)
        h = _linear_ex(x, w1, b1, epilogue=1, aux=z1)
        y = _linear_ex(h, w2, b2, epilogue=0)
        ctx.save_for_backward(x, w1, w2, z1, h)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w1, w2, z1, h = ctx.saved_tensors
        dy = dy.contiguous().to(x.dtype)
        need = ctx.needs_input_grad
        dz1 = _linear_ex(dy, _transpose16(w2), None, epilogue=2, aux=z1)            # [M, hid]
        dw2 = _linear_ex(_transpose16(dy), _transpose16(h), None) if need[3] else None   # [hid, hid]
        db2 = _colsum(dy) if need[4] else None
        dx = _linear_ex(dz1, _transpose16(w1), None) if need[0] else None           # [M, C]
        dw1 = _linear_ex(_transpose16(dz1), _transpose16(x), None) if need[1] else None  # [hid, C]
        db1 = _colsum(dz1) if need[2] else None
        return dx, dw1, db1, dw2, db2


def build_region_encoder(config, image_aspect_ratio):
    """Drop-in for layer.py:155-161."""
    region_encoder_type = getattr(config, "mm_region_encoder_type", "pooling")
    if region_encoder_type == "pooling":
        return MaskExtractor(image_aspect_ratio, config)
    raise ValueError(f"Unknown region encoder type: {region_encoder_type}")
