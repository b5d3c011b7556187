"""CPU, world_size 2, gloo: the host logic of the multi-GPU paths -- RoI sharding (no collective) and the gradient
all-reduce bucket of the training step.  No CUDA compute here; the kernels' N>1 run is `bench.py --gpus N` on the box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from givepose_b200.train import GradBucket, shard_batch, shard_range


def test_shard_ranges_partition_the_batch():
    for total in (0, 1, 7, 48, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_range(4096, k, 8) for k in (0, 7)] == [(0, 512), (3584, 4096)]   # multiples of 256 (im2col_step) and of 4
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_shard_batch_slices_per_roi_tensors_and_shares_a_single_camera():
    data = {"roi_img": torch.arange(10.0).view(10, 1), "cam_K": torch.eye(3), "roi_wh": torch.ones(10, 2)}
    s = shard_batch(data, 1, 3)
    assert s["roi_img"].flatten().tolist() == [4.0, 5.0, 6.0] and s["cam_K"].shape == (3, 3) and s["roi_wh"].shape == (3, 2)
    data["cam_K"] = torch.eye(3).repeat(10, 1, 1)
    assert shard_batch(data, 2, 3)["cam_K"].shape == (3, 3, 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)   # identical replicas
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
        unused = torch.nn.BatchNorm1d(3)   # never receives a gradient, like DCNv3_C.bn
        params = list(net.parameters()) + list(unused.parameters())
        bucket = GradBucket(params)
        x = torch.arange(24.0).view(4, 6) / 10
        lo, hi = shard_range(4, rank, world)
        net(x[lo:hi]).square().sum().backward()
        bucket.allreduce_()
        flat = torch.cat([p.grad.flatten() for p in params])
        # reference: mean over ranks of the per-shard gradients == (sum over all rows) / world
        ref_net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
        ref_net.load_state_dict(net.state_dict())
        ref_net(x).square().sum().backward()
        ref = torch.cat([p.grad.flatten() for p in ref_net.parameters()] + [torch.zeros(6)]) / world
        ok = torch.allclose(flat, ref, atol=1e-6) and all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))
        # second step: p.grad are views of the flat buffer, zeroing is one memset and autograd accumulates in place
        bucket.zero_()
        net(x[lo:hi]).square().sum().backward()
        bucket.allreduce_()
        ok = ok and torch.allclose(bucket.flat, ref, atol=1e-6)
        out[rank] = (ok, bucket.nbytes())
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_bucket_gloo_world2():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: (True, (6 * 5 + 5 + 5 * 2 + 2 + 6) * 4), 1: (True, (6 * 5 + 5 + 5 * 2 + 2 + 6) * 4)}


def test_bucket_clip_equals_clip_grad_norm():
    """GradBucket.clip_ == torch.nn.utils.clip_grad_norm_ (engine/train.py:126) on the same gradients, both when the
    norm exceeds the bound and when it does not; parameters without a gradient count as zeros."""
    for scale in (10.0, 1e-3):
        torch.manual_seed(1)
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 2))
        unused = torch.nn.BatchNorm1d(3)
        ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 2))
        ref.load_state_dict(net.state_dict())
        bucket = GradBucket(list(net.parameters()) + list(unused.parameters()))
        x = torch.randn(7, 6) * scale
        net(x).square().sum().backward()
        ref(x).square().sum().backward()
        n_ref = torch.nn.utils.clip_grad_norm_(ref.parameters(), 5.0)
        n = bucket.clip_(5.0)
        assert torch.allclose(n, n_ref, rtol=1e-6)
        for a, b in zip(net.parameters(), ref.parameters()):
            assert torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-8)


class _TinyNet(torch.nn.Module):
    """PoseNet-shaped call signature (data dict, device, do_loss) around one Linear: enough for train_step's host logic."""

    def __init__(self):
        super().__init__()
        self.lin = torch.nn.Linear(6, 3)

    def forward(self, data, device, do_loss=False, pred_scale=None):
        return {"y": self.lin(data["roi_img"])}


def _crit(out, target):
    return {"l2": (out["y"] - target["y"]).square().mean()}   # a mean over the rank's RoIs, like every PoseLoss term


def _worker_unequal(rank, world, port, out):
    from givepose_b200.train import train_step
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net, ref = _TinyNet(), _TinyNet()
        ref.load_state_dict(net.state_dict())
        B = 5                                                   # shards of 3 and 2 RoIs
        x, y = torch.randn(B, 6), torch.randn(B, 3)
        lo, hi = shard_range(B, rank, world)
        bucket = GradBucket(net.parameters())
        opt = torch.optim.SGD(net.parameters(), lr=0.0)
        train_step(net, {"roi_img": x[lo:hi], "roi_mask": x[lo:hi]}, {"y": y[lo:hi]}, opt, bucket, "cpu", clip=1e9, criterion=_crit)
        _crit(ref({"roi_img": x}, "cpu"), {"y": y})["l2"].backward()      # the single-process full-batch step (engine/train.py:117-125)
        want = torch.cat([p.grad.flatten() for p in ref.parameters()])
        ok = torch.allclose(bucket.flat, want, atol=1e-6)
        empty = False
        try:
            train_step(net, {"roi_img": x[:0], "roi_mask": x[:0]}, {"y": y[:0]}, opt, bucket, "cpu", criterion=_crit)
        except ValueError:
            empty = True
        out[rank] = (ok, empty)
    finally:
        dist.destroy_process_group()


def test_unequal_shards_give_the_full_batch_mean_gradient_gloo_world2():
    """shard_range hands the remainder to the low ranks; train_step weights each rank by n_local*world/n_global so the
    averaged gradient equals the single-process full-batch mean (and refuses an empty shard before any collective)."""
    world, port = 2, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker_unequal, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: (True, True), 1: (True, True)}
