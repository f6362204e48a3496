"""CPU, world_size 2 (gloo): the data-parallel plumbing of the optimiser — one all-reduce(SUM) of the flat LoRA-gradient
buffer, mean folded in as 1/world — reproduces single-process training on the concatenated batch (SURVEY 8e / test
pyramid item 4).  The fused CUDA update is replaced by its torch formula (test infrastructure)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torch_step(self, handle=None):
    if handle is not None:
        handle.wait()
    self.step_count += 1
    g = self.grad / self.world()
    if self.max_norm > 0:
        g = g * min(1.0, self.max_norm / (float(g.norm()) + 1e-6))
    b1, b2 = self.betas
    self.flat.mul_(1 - self.lr * self.wd)
    self.m.mul_(b1).add_(g, alpha=1 - b1)
    self.v.mul_(b2).addcmul_(g, g, value=1 - b2)
    denom = self.v.sqrt() / (1 - b2 ** self.step_count) ** 0.5 + self.eps
    self.flat.addcdiv_(self.m, denom, value=-self.lr / (1 - b1 ** self.step_count))


def _model():
    torch.manual_seed(0)
    return [torch.nn.Parameter(torch.randn(8, 5)), torch.nn.Parameter(torch.randn(3, 8))]


def _loss(params, x, y):
    return ((torch.tanh(x @ params[0].t()) @ params[1].t() - y) ** 2).mean()


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from comat_b200 import optim
    optim.FlatAdamW.step = _torch_step
    params = _model()
    opt = optim.FlatAdamW(params, lr=1e-2, max_grad_norm=0.5)
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(8, 5, generator=g), torch.randn(8, 3, generator=g)
    for _ in range(3):
        xs, ys = X[rank::world], Y[rank::world]            # prompts i = r (mod world)
        opt.zero_grad()
        _loss(params, xs, ys).backward()
        opt.step(opt.all_reduce())
    if rank == 0:
        torch.save(opt.flat.clone(), out)
    dist.destroy_process_group()


def test_two_rank_flat_allreduce_matches_single_process(tmp_path):
    port = 29500 + os.getpid() % 2000
    out = str(tmp_path / "flat.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    sys.path.insert(0, ROOT)
    from comat_b200 import optim
    old = optim.FlatAdamW.step
    optim.FlatAdamW.step = _torch_step
    try:
        params = _model()
        opt = optim.FlatAdamW(params, lr=1e-2, max_grad_norm=0.5)
        g = torch.Generator().manual_seed(1)
        X, Y = torch.randn(8, 5, generator=g), torch.randn(8, 3, generator=g)
        for _ in range(3):
            opt.zero_grad()
            # mean over ranks of per-rank means == mean over the whole batch (equal shard sizes)
            (0.5 * _loss(params, X[0::2], Y[0::2]) + 0.5 * _loss(params, X[1::2], Y[1::2])).backward()
            opt.step()
    finally:
        optim.FlatAdamW.step = old
    assert torch.allclose(got, opt.flat, atol=1e-6), (got - opt.flat).abs().max()
