"""Flat-buffer optimiser for the LoRA parameters: one NCCL all-reduce + one fused clip/AdamW launch per optimiser step.

Mirrors the reference's optimiser semantics (torch.optim.AdamW + clip_grad_norm_, training_script.py:215-275, 658-664):
same hyper-parameters, same update rule; the data-parallel mean (DDP allreduce-avg) is folded in as
``grad_scale = 1 / world_size``.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib

_vp, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
_lib.register_signature("comat_grad_sumsq", [_vp, _ll, _vp, _vp, _vp])
_lib.register_signature("comat_adamw_clip", [_vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _i, _f, _f, _vp, _vp])


class FlatAdamW:
    def __init__(self, params: List[torch.nn.Parameter], lr, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8, max_grad_norm=0.0,
                 process_group=None):
        self.params = list(params)
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.n = n
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:                               # re-point parameters and their grads at the flat buffers
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p.data)
            p.grad = self.grad[off:off + k].view_as(p.data)
            off += k
        self.lr, self.betas, self.wd, self.eps, self.max_norm = lr, betas, weight_decay, eps, max_grad_norm
        self.step_count = 0
        self.pg = process_group
        self._scratch = torch.empty(1024 + 1, dtype=torch.float32, device=dev) if dev.type == "cuda" else None

    def zero_grad(self):
        self.grad.zero_()
        for p, view in zip(self.params, self._grad_views()):
            p.grad = view

    def _grad_views(self):
        off = 0
        for p in self.params:
            k = p.numel()
            yield self.grad[off:off + k].view_as(p.data)
            off += k

    def world(self):
        return dist.get_world_size(self.pg) if dist.is_available() and dist.is_initialized() else 1

    def all_reduce(self):
        """the single data-path collective of the step: sum over ranks of the flat LoRA gradient."""
        if self.world() > 1:
            return dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        return None

    def step(self, handle=None):
        if handle is not None:
            handle.wait()
        self.step_count += 1
        scale = 1.0 / self.world()
        if not self.flat.is_cuda:
            raise _lib.ComatError("FlatAdamW.step runs the fused CUDA kernel only (no CPU fallback)")
        L = _lib.lib()
        st = _lib.stream_ptr()
        sumsq = self._scratch[1024:]
        if self.max_norm > 0:
            _lib.check(L.comat_grad_sumsq(self.grad.data_ptr(), self.n, self._scratch.data_ptr(), sumsq.data_ptr(), st), "grad_sumsq")
            _lib.count_launch(2)
        _lib.check(L.comat_adamw_clip(self.flat.data_ptr(), self.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.n,
                                      self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.step_count,
                                      float(self.max_norm), scale, sumsq.data_ptr(), st), "adamw_clip")
        _lib.count_launch()

    def grad_norm(self) -> torch.Tensor:
        return self._scratch[1024:].sqrt() / self.world()
