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


# ---- the training entry point under DP (comat_b200/train.py): sharded prompts, one gradient all-reduce per optimiser, rank-0
# ---- checkpoints, rank-averaged logs
class _Patch:
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def _loop_worker(rank, world, port, prompts, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(4)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import cpu_ops_emulation as EMU
    from tests.test_trainer_logic_cpu import _emulate_cuda_only
    _emulate_cuda_only(_Patch())
    EMU.install_blip(_Patch())
    from comat_b200 import optim, synthetic
    from comat_b200.train import Trainer
    optim.FlatAdamW.step = _torch_step
    a = synthetic.default_args(pretrain_model_name="sd_1_5", train_batch_size=1, K=1, total_step=2, resolution=64,
                               training_prompts=prompts, output_dir=out, max_train_steps=2, validation_steps=100,
                               resume_from_checkpoint=None, seed=3, gradient_accumulation_steps=1)
    tr = Trainer(a, None, torch.device("cpu"), rank, world, weights="synthetic_tiny", dtype=torch.float32)
    seen = [t for b in tr.loader for t in b["text"]]
    tr.train()
    torch.save({"flat": tr.core.optimizer.flat.clone(), "seen": seen}, os.path.join(out, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_training_loop(tmp_path):
    import json
    prompts = tmp_path / "prompts.txt"
    prompts.write_text("\n".join(["a red apple", "two dogs on a sofa", "a blue car", "snow on a hill", "a green bench"]) + "\n")
    out = str(tmp_path / "run")
    os.makedirs(out)
    port = 31500 + os.getpid() % 2000
    mp.spawn(_loop_worker, args=(2, port, str(prompts), out), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(out, f"rank{r}.pt")) for r in range(2))
    assert torch.equal(r0["flat"], r1["flat"])                                   # replicas stay identical: same averaged gradient
    assert len(r0["seen"]) == len(r1["seen"]) == 2 and not set(r0["seen"]) & set(r1["seen"])
    assert sorted(d for d in os.listdir(out) if d.startswith("checkpoint")) == ["checkpoint-2"]      # written once, by rank 0
    logs = [json.loads(l) for l in open(os.path.join(out, "train_log.jsonl"))]
    assert [l["step"] for l in logs] == [1, 2] and all(abs(l["step_loss"]) > 0 for l in logs)
