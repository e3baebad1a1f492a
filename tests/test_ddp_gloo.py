"""CPU, world_size 2 over gloo: the bucketed gradient reducer (host logic of the data-parallel path)."""
import importlib.util
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_ddp():
    # loaded by path: the reducer is plain torch.distributed and must be testable without a GPU
    spec = importlib.util.spec_from_file_location("maskunet_ddp", os.path.join(ROOT, "maskunet_b200", "ddp.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ddp = _load_ddp()
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.GELU(), torch.nn.Linear(64, 64), torch.nn.GELU(),
                              torch.nn.Linear(64, 8))
    dead = torch.nn.Linear(4, 4)                       # like the reference's emb_layer: never used in forward
    params = list(net.parameters()) + list(dead.parameters())
    if rank == 1:                                      # de-synchronise, then broadcast must repair it
        with torch.no_grad():
            for p in net.parameters():
                p.add_(1.0)
    red = ddp.GradReducer(params, bucket_bytes=8 * 1024)
    red.broadcast_parameters(net)
    gen = torch.Generator().manual_seed(123)
    x_all = torch.randn(3, 8, 16, generator=gen)       # 3 steps, global batch 8
    y_all = torch.randn(3, 8, 8, generator=gen)
    ref = torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.GELU(), torch.nn.Linear(64, 64), torch.nn.GELU(),
                              torch.nn.Linear(64, 8))
    ref.load_state_dict(net.state_dict())
    for step in range(3):
        for p in params:
            p.grad = None
        xs, ys = x_all[step].chunk(world)[rank], y_all[step].chunk(world)[rank]
        torch.nn.functional.mse_loss(net(xs), ys).backward()
        red.finish()
        ref.zero_grad()
        torch.nn.functional.mse_loss(ref(x_all[step]), y_all[step]).backward()
        for p, q in zip(net.parameters(), ref.parameters()):
            assert torch.allclose(p.grad, q.grad, atol=1e-6), f"rank {rank} step {step}"
        assert all(p.grad is None for p in dead.parameters())
    assert len(red.buckets) >= 2 and red.launched == 3 * len(red.buckets)
    # after finish() every gradient is a slice of its bucket's persistent flat buffer (no copy back)
    for b in red.buckets:
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(b.params, b.views))
    # gradient accumulation: two half-batches per rank, the exchange rides on the last backward only
    launched = red.launched
    for p in params:
        p.grad = None
    xs, ys = x_all[0].chunk(world)[rank], y_all[0].chunk(world)[rank]
    for i, (xm, ym) in enumerate(zip(xs.chunk(2), ys.chunk(2))):
        red.accumulate(i == 0)
        (torch.nn.functional.mse_loss(net(xm), ym) / 2).backward()
    red.finish()
    ref.zero_grad()
    torch.nn.functional.mse_loss(ref(x_all[0]), y_all[0]).backward()
    for p, q in zip(net.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-6), f"rank {rank} accumulation"
    assert red.launched == launched + len(red.buckets)
    torch.save(torch.tensor(1), os.path.join(tmp, f"ok{rank}"))
    dist.destroy_process_group()


def test_grad_reducer_world2_gloo(tmp_path):
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def _worker_shared(rank, world, port, tmp):
    """A parameter used twice in the forward fires its hook once per backward: one bucket slot, not two."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ddp = _load_ddp()
    torch.manual_seed(1)
    lin = torch.nn.Linear(8, 8)
    # a channels-last convolution weight (model.to(memory_format=channels_last)): its gradient slice must carry the
    # parameter's strides, or torch's fused AdamW refuses the step ("same dtype, device, and layout")
    conv = torch.nn.Conv2d(4, 6, 3).to(memory_format=torch.channels_last)
    red_c = ddp.GradReducer(list(conv.parameters()), bucket_bytes=1 << 20)
    opt = torch.optim.AdamW(conv.parameters(), lr=1e-3)
    xc = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(3 + rank))
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        conv(xc).square().mean().backward()
        red_c.finish()
        assert conv.weight.grad.stride() == conv.weight.stride() and conv.weight.grad.shape == conv.weight.shape
        opt.step()
    gathered = [torch.zeros_like(conv.weight.grad.contiguous()) for _ in range(world)]
    dist.all_gather(gathered, conv.weight.grad.contiguous())
    assert torch.equal(gathered[0], gathered[1])       # both ranks hold the same mean gradient
    red = ddp.GradReducer(list(lin.parameters()), bucket_bytes=1 << 20)
    try:
        red.finish()                                   # before any backward: refuse, do not finalise an empty discovery
        raise AssertionError("finish() before backward must raise")
    except RuntimeError:
        pass
    x = torch.randn(4, 8, generator=torch.Generator().manual_seed(5))
    for _ in range(2):
        for p in lin.parameters():
            p.grad = None
        lin(lin(x[rank * 2:rank * 2 + 2])).square().mean().backward()
        red.finish()
    assert sum(len(b.params) for b in red.buckets) == 2
    ref = torch.nn.Linear(8, 8)
    ref.load_state_dict(lin.state_dict())
    ref(ref(x)).square().mean().backward()
    for p, q in zip(lin.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-6)
    torch.save(torch.tensor(1), os.path.join(tmp, f"ok{rank}"))
    dist.destroy_process_group()


def test_grad_reducer_shared_parameter_and_empty_finish(tmp_path):
    port = 29300 + os.getpid() % 300
    mp.spawn(_worker_shared, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
