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

_vp, _i, _f, _ll, _d = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_double
_lib.register_signature("comat_grad_sumsq", [_vp, _ll, _vp, _vp, _vp])
_lib.register_signature("comat_adamw_clip", [_vp, _vp, _vp, _vp, _ll, _d, _d, _d, _d, _d, _f, _f, _vp, _vp, _vp, _vp])


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
        self._step_host = 0
        self.pg = process_group
        cuda = dev.type == "cuda"
        self._scratch = torch.zeros(1024 + 1 + 4, dtype=torch.float32, device=dev) if cuda else None
        # device-side step bookkeeping {steps taken, steps skipped on a non-finite gradient, last step skipped}: the overflow guard
        # (GradScaler semantics) never needs the host; ``poll_overflow`` reads it back one step late without a sync
        self.counters = torch.zeros(3, dtype=torch.int32, device=dev) if cuda else None
        self._counters_host = torch.zeros(3, dtype=torch.int32).pin_memory() if cuda else None
        self._counters_ev = None
        self._skipped_seen = 0

    @property
    def step_count(self) -> int:
        """optimiser steps actually taken (skipped overflow steps do not count, as with GradScaler).  On CUDA this reads the
        device counter (a sync: checkpointing / tests only)."""
        return int(self.counters[0].item()) if self.counters is not None else self._step_host

    @step_count.setter
    def step_count(self, v: int):
        if self.counters is not None:
            self.counters[0] = int(v)
        self._step_host = int(v)

    def poll_overflow(self):
        """non-blocking: returns (new overflow-skipped steps since the last poll, steps taken) once the previous snapshot has
        landed, else None; then queues the next snapshot of the device counters on the current stream."""
        if self.counters is None:
            return None
        res = None
        if self._counters_ev is None or self._counters_ev.query():
            if self._counters_ev is not None:
                skipped = int(self._counters_host[1])
                res = (skipped - self._skipped_seen, int(self._counters_host[0]))
                self._skipped_seen = skipped
            self._counters_host.copy_(self.counters, non_blocking=True)
            self._counters_ev = torch.cuda.Event()
            self._counters_ev.record()
        return res

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
        scale = 1.0 / self.world()
        if not self.flat.is_cuda:
            raise _lib.ComatError("FlatAdamW.step runs the fused CUDA kernel only (no CPU fallback)")
        self._step_host += 1
        L = _lib.lib()
        st = _lib.stream_ptr()
        sumsq, state = self._scratch[1024:1025], self._scratch[1025:]
        # always reduced: the clip coefficient and the overflow guard (non-finite gradient -> the whole update is skipped) read it
        _lib.check(L.comat_grad_sumsq(self.grad.data_ptr(), self.n, self._scratch.data_ptr(), sumsq.data_ptr(), st), "grad_sumsq")
        _lib.check(L.comat_adamw_clip(self.flat.data_ptr(), self.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.n,
                                      self.lr, self.betas[0], self.betas[1], self.eps, self.wd, float(self.max_norm), scale,
                                      sumsq.data_ptr(), state.data_ptr(), self.counters.data_ptr(), st), "adamw_clip")
        _lib.count_launch(4)

    def grad_norm(self) -> torch.Tensor:
        return self._scratch[1024:1025].sqrt() / self.world()
