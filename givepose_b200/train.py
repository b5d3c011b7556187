"""Data-parallel plumbing for the PoseNet path: RoI sharding (inference, no collective) and the training step's
gradient all-reduce (the only collective on the path).

The reference is single-process (``engine/train.py:26,35``); its step is ``forward(do_loss=True) -> loss -> backward ->
clip_grad_norm_(5) -> optimizer.step()`` (``:117-129``).  Here every rank holds a full replica and a contiguous shard of
the RoI batch; one bucketed ``all_reduce(sum) / world`` over a flat fp32 buffer sits between ``backward()`` and the
clip / step.  Backend comes from the process group: NCCL over NVLink on the GPUs, gloo in the CPU tests.
BatchNorm statistics stay per rank (the reference has no SyncBN in use); parameters that never receive a gradient
(``DCNv3_C.bn``, built but unused, ``network/dcnv3.py:29,37``) are tolerated.
"""
from __future__ import annotations

from typing import Dict, Iterable, Tuple

import torch
import torch.distributed as dist
import torch.nn.functional as F


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous RoI shard of ``rank``; shards differ by at most one RoI (remainder goes to the low ranks)."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError(f"bad shard request: total={total} rank={rank} world={world}")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(data: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Slice every per-RoI tensor of the input dict (``datasets/load_data_nocs.py:355-386`` layout).  A camera matrix given
    once as (3,3) is shared.  NOTE: the stride-2 DCNv3 calls couple RoIs inside a batch, so a shard's
    results equal the reference run on that shard, not a slice of the reference run on the whole batch."""
    n = next(v.shape[0] for k, v in data.items() if k != "cam_K" or v.dim() == 3)
    lo, hi = shard_range(n, rank, world)
    return {k: (v if (k == "cam_K" and v.dim() == 2) else v[lo:hi]) for k, v in data.items()}


class GradBucket:
    """One flat fp32 buffer that IS the gradient storage of every trainable parameter (``p.grad`` are views into it, autograd
    accumulates in place) -> zeroing is one memset, the all-reduce one collective, and nothing is copied in or out.
    Parameters that never receive a gradient (``DCNv3_C.bn``) simply keep zeros."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradBucket: no trainable parameters")
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=self.params[0].device)
        self.views, o = [], 0
        for p in self.params:
            chunk = self.flat[o:o + p.numel()]
            # same strides as the parameter (conv weights are kept channels_last): autograd then accumulates without a layout copy
            dense = p.is_contiguous() or p.is_contiguous(memory_format=torch.channels_last) if p.dim() == 4 else p.is_contiguous()
            self.views.append(chunk.as_strided(p.size(), p.stride()) if dense else chunk.view_as(p))
            o += p.numel()
        self.attach()

    def attach(self) -> None:
        """(Re)install the views as ``p.grad`` -- needed after anything that replaced them (``zero_grad(set_to_none=True)``)."""
        for p, v in zip(self.params, self.views):
            if p.dtype == torch.float32 and p.grad is not v:
                p.grad = v

    def zero_(self) -> None:
        self.attach()
        self.flat.zero_()

    def nbytes(self) -> int:
        return self.numel * 4

    @torch.no_grad()
    def allreduce_(self, group=None) -> None:
        """all_reduce(sum) / world of the flat buffer; gradients that autograd placed elsewhere (non-fp32 parameters, a
        replaced ``p.grad``) are folded in first."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        for p, v in zip(self.params, self.views):
            if p.grad is not None and p.grad is not v:
                v.copy_(p.grad)
                p.grad = v if p.dtype == torch.float32 else p.grad
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(world)
            for p, v in zip(self.params, self.views):
                if p.grad is not None and p.grad is not v:
                    p.grad.copy_(v)


    @torch.no_grad()
    def clip_(self, max_norm: float) -> torch.Tensor:
        """``clip_grad_norm_(parameters, max_norm)`` (``engine/train.py:126``) on the flat buffer: one norm, one scale, no host
        sync (the total norm over all parameters IS the norm of the bucket; never-touched parameters hold zeros)."""
        norm = torch.linalg.vector_norm(self.flat)
        self.flat.mul_(torch.clamp(max_norm / (norm + 1e-6), max=1.0))
        return norm


def surrogate_loss(out: Dict[str, torch.Tensor], target: Dict[str, torch.Tensor]) -> torch.Tensor:
    """A dataset-free stand-in for the loss (the reference's ``PoseLoss``, ``losses/pose_loss.py:30-96``, is built in
    ``givepose_b200.loss.PoseLoss`` and is what ``criterion=`` takes; this one needs only prediction-shaped targets, so the
    step can be exercised without ground-truth point clouds / symmetry labels): L1 on rot / trans / size and smooth-L1 on the
    two coordinate maps -- every head and the DCNv3 backward get gradients."""
    return (F.l1_loss(out["rot"], target["rot"]) + F.l1_loss(out["trans"], target["trans"]) + F.l1_loss(out["size"], target["size"])
            + F.smooth_l1_loss(out["nocs_coor"], target["nocs_coor"]) + F.smooth_l1_loss(out["ivfc_coor"], target["ivfc_coor"]))


def make_targets(B: int, device, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(2000 + seed)
    r = lambda *s: torch.randn(*s, generator=g)
    q, _ = torch.linalg.qr(r(B, 3, 3))
    return {k: v.to(device) for k, v in {"rot": q, "trans": r(B, 3) * 0.3, "size": r(B, 3) * 0.1 + 0.5,
                                         "nocs_coor": r(B, 3, 64, 64) * 0.3, "ivfc_coor": r(B, 3, 64, 64) * 0.3}.items()}


def forward_backward(net, data, target, bucket: GradBucket, device, clip: float = 5.0, group=None, criterion=None):
    """Everything of a data-parallel step up to (not including) the optimizer: zero the bucket -> forward (autograd path: torch
    CUDA ops around ``DCNv3Function``) -> loss -> backward (DCNv3 backward kernel) -> gradient all-reduce -> clip
    (``engine/train.py:126``).  Device-side only (no host sync), which is what ``GraphedTrainStep`` captures."""
    net.train()
    bucket.zero_()   # p.grad are views of the flat buffer: one memset instead of optimizer.zero_grad()
    data = dict(data)
    data.setdefault("roi_mask_deform", data["roi_mask"])
    n_local = int(data["roi_img"].shape[0])
    if n_local == 0:
        raise ValueError("train_step: this rank's RoI shard is empty (batch smaller than the world size); give every rank at "
                         "least one RoI or drop the rank from the process group")
    out = net(data, device, do_loss=True)
    loss = surrogate_loss(out, target) if criterion is None else sum(criterion(out, target).values())
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        # every loss term is a mean over THIS rank's RoIs and the bucket averages over ranks: weight the rank by
        # n_local * world / n_global so that unequal shards (shard_range hands the remainder to the low ranks) still give
        # the single-process full-batch mean gradient of engine/train.py:117-125.  One 4-byte collective, device-side only.
        cnt = torch.full((1,), float(n_local), dtype=torch.float32, device=loss.device)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
        scaled = loss * (float(n_local * world) / cnt).squeeze(0)
    else:
        scaled = loss
    scaled.backward()
    bucket.allreduce_(group)
    bucket.clip_(clip)
    return loss.detach()


def train_step(net, data, target, optimizer, bucket: GradBucket, device, clip: float = 5.0, group=None, criterion=None) -> float:
    """One data-parallel step on this rank's shard: ``forward_backward`` (forward -> loss -> backward -> all-reduce -> clip)
    then ``optimizer.step()``.

    ``criterion``: a ``givepose_b200.loss.PoseLoss`` -- then ``target`` is the ground-truth dict of the reference's data
    loader and the loss is the sum of its terms as in ``engine/train.py:120-122``; ``None`` keeps the simple surrogate."""
    loss = forward_backward(net, data, target, bucket, device, clip, group, criterion)
    optimizer.step()
    return loss


class GraphedTrainStep:
    """``forward_backward`` captured ONCE as a CUDA graph and replayed, followed by an EAGER ``optimizer.step()``: zero grads ->
    forward -> loss -> backward (DCNv3 backward kernel) -> NCCL all-reduce of the flat bucket -> clip are ~4400 launches for
    48 RoIs, i.e. the eager step is bound by launch overhead, not by the GPU.  Inputs and targets live in static device
    buffers that ``__call__`` refreshes (H2D or D2D copies on the same stream, ahead of the replay).  Nothing in the captured
    part synchronises with the host (``PoseLoss`` selects the symmetric rotations on device), which makes the capture legal.

    The optimizer step stays OUTSIDE the graph on purpose: a captured step bakes every host-side scalar it reads -- the
    learning rate, Adam's step counter -- into the graph, so a scheduler (``engine/train.py:128`` calls ``scheduler.step()``
    after every optimizer step) or a non-capturable optimizer (the reference's default Ranger, Adam / AdamW without
    ``capturable=True``) would be silently ignored on replay, and torch's SGD reads a tensor learning rate through ``.item()``
    (a sync) outside ``torch.compile``.  Eagerly it is a handful of multi-tensor launches on the gradients the graph left in
    ``bucket`` (``p.grad`` are views of it), works with any optimizer and any scheduler, and costs ~0.1 ms.

    The ``warmup`` eager steps before the capture are real optimisation steps on ``example_data`` (cuDNN algorithm
    selection, momentum buffers, the NCCL communicator all have to exist before capture)."""

    def __init__(self, net, optimizer, bucket: GradBucket, device, example_data, example_target, clip: float = 5.0, group=None,
                 criterion=None, warmup: int = 3):
        dev = torch.device(device)
        self.net, self.optimizer, self.bucket, self.dev = net, optimizer, bucket, dev
        self.clip, self.group, self.criterion = clip, group, criterion
        self.data = {k: v.to(dev).clone() for k, v in example_data.items()}
        self.target = {k: v.to(dev).clone() for k, v in example_target.items()}
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                train_step(net, self.data, self.target, optimizer, bucket, dev, clip, group, criterion)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: the NCCL watchdog thread polls events while we capture
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.loss = forward_backward(net, self.data, self.target, bucket, dev, clip, group, criterion)

    def __call__(self, data=None, target=None) -> torch.Tensor:
        """Refresh the static buffers (skipped for ``None`` / for tensors that already ARE the static buffers), replay the
        captured forward/backward/all-reduce/clip, then step the optimizer eagerly (current learning rate, any optimizer)."""
        for static, new in ((self.data, data), (self.target, target)):
            if new is not None:
                for k, v in new.items():
                    if v is not static[k]:
                        static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        self.optimizer.step()
        return self.loss
