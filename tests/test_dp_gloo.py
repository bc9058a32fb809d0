"""Data-parallel host logic on CPU with gloo, world_size 2: the flat per-optimiser gradient
all-reduce hooked on ``optimizer.step`` (ideas_b200/train_step.py) keeps replicas identical and
equals single-process training on the concatenated batch; it also fires on the reference loop's
second backward without a forward (train.py:214-216), which stock DDP does not cover."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _Net(torch.nn.Module):
    """Equalised-lr style layers: parameters enter the graph only through ``w * scale``
    temporaries (as in stylegan2/model.py:115,153), so a second backward after an in-place
    optimiser step is legal -- the property the reference loop relies on."""

    def __init__(self):
        super().__init__()
        self.w1 = torch.nn.Parameter(torch.randn(5, 6))
        self.w2 = torch.nn.Parameter(torch.randn(1, 5))

    def forward(self, x):
        h = torch.tanh(torch.nn.functional.linear(x, self.w1 * 0.4))
        return torch.nn.functional.linear(h, self.w2 * 0.45)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out, bound=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ideas_b200.train_step import FlatGradAllReduce
    torch.manual_seed(0)
    net = _Net()
    opt = torch.optim.Adam(net.parameters(), lr=1e-2, betas=(0.0, 0.99))
    red = FlatGradAllReduce(list(net.parameters()))
    opt.register_step_pre_hook(red)
    if bound:                      # grads become views into the flat bucket; zero_grad must keep them
        red.bind()
        opt.zero_grad = lambda *a, **k: red.zero()
    g = torch.Generator().manual_seed(123)
    X = torch.randn(8, 6, generator=g)
    Y = torch.randn(8, 1, generator=g)
    xs, ys = X[rank::world], Y[rank::world]
    for _ in range(3):
        pred = net(xs)
        loss1 = (pred - ys).pow(2).mean()
        loss2 = (pred - ys).abs().mean()
        opt.zero_grad()
        loss1.backward(retain_graph=True)
        opt.step()
        opt.zero_grad()
        loss2.backward()               # second backward, no forward in between
        opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        torch.save(dict(params=gathered, calls=red.calls), out)
    dist.destroy_process_group()


def _single():
    torch.manual_seed(0)
    net = _Net()
    opt = torch.optim.Adam(net.parameters(), lr=1e-2, betas=(0.0, 0.99))
    g = torch.Generator().manual_seed(123)
    X = torch.randn(8, 6, generator=g)
    Y = torch.randn(8, 1, generator=g)
    for _ in range(3):
        pred = net(X)
        loss1 = (pred - Y).pow(2).mean()
        loss2 = (pred - Y).abs().mean()
        opt.zero_grad()
        loss1.backward(retain_graph=True)
        opt.step()
        opt.zero_grad()
        loss2.backward()
        opt.step()
    return torch.cat([p.detach().reshape(-1) for p in net.parameters()])


def test_flat_allreduce_world2(tmp_path):
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["calls"] == 6
    assert torch.equal(r["params"][0], r["params"][1])                     # replicas stay identical
    assert torch.allclose(r["params"][0], _single(), atol=1e-6, rtol=1e-5)  # == big-batch training


def test_flat_allreduce_in_place_bucket_world2(tmp_path):
    """bind(): p.grad are views into the bucket, autograd accumulates in place, the all-reduce needs no copies."""
    out = str(tmp_path / "dp_bound.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, True), nprocs=2, join=True)
    r = torch.load(out)
    assert r["calls"] == 6
    assert torch.equal(r["params"][0], r["params"][1])
    assert torch.allclose(r["params"][0], _single(), atol=1e-6, rtol=1e-5)


def test_flat_allreduce_is_noop_without_process_group():
    from ideas_b200.train_step import FlatGradAllReduce
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    FlatGradAllReduce([p])()
    assert torch.equal(p.grad, torch.full((3,), 2.0))
