"""diffusers-shaped callables backed by the explicit executors (comat_b200/engine.py).

``EngineUNet`` / ``EngineVAE`` expose exactly the surface the reference pipelines call —
``unet(x, t, encoder_hidden_states=..., added_cond_kwargs=..., return_dict=False)[0]`` (TrainableSDPipeline.py:144-150)
and ``vae.decode(z, return_dict=False)[0]`` (:220) — so they drop into ``TrainableSDPipeline`` unchanged.  Each call is ONE
node in torch's autograd graph (plumbing for the latent chain); inside it the forward and backward are our own kernels
driven by an explicit tape.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch

from . import engine as E
from . import ops


class _UNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod: "EngineUNet", x, t, ehs, added, capture, ctx_kv, *lora_params):
        eng = mod.engine
        tape = E.Tape()
        xv = E.Var(ops.latent_to_nhwc(x, eng.dtype, 64), needs_grad=ctx.needs_input_grad[1])
        ctx.wgrad = bool(mod.train_lora and any(p.requires_grad for p in lora_params))
        ctx.product = ctx.wgrad and eng.lora_train_impl == "product"
        ctx16, kv = ctx_kv if ctx_kv is not None else (ehs.to(eng.dtype), None)
        # d(encoder_hidden_states) only when the caller's embeddings require grad (never on the frozen-text-encoder CoMat path)
        cvar = E.Var(ctx16, needs_grad=True) if ctx.needs_input_grad[3] else ctx16
        # no LoRA weight gradient wanted (the discriminator's generator-side pass, gan_sdxl.py:52-89): LoRA folded into the weights
        out = eng.forward(tape, xv, t, cvar, capture=capture, added_cond=added,
                          lora_mode="train" if ctx.wgrad else "frozen", cross_kv=kv if (ctx.product or not ctx.wgrad) else None)
        eps = ops.nhwc_to_nchw_f32(out.v, mod.out_channels)
        pvars = []
        if capture is not None:
            for place in ("down", "mid", "up"):
                pvars.extend(capture.store[place])
        ctx.tape, ctx.xv, ctx.out, ctx.pvars, ctx.mod = tape, xv, out, pvars, mod
        ctx.cvar = cvar if isinstance(cvar, E.Var) else None
        ctx.taped_key = (mod._taped_key(x, ehs, added, capture, bool(ctx.needs_input_grad[1]), ctx.wgrad)
                         if mod.graph_taped and not ctx.needs_input_grad[3] else None)
        ctx.x_dtype, ctx.ehs_meta = x.dtype, (ehs.dtype, ehs.shape)
        return (eps.to(x.dtype), *[p.v for p in pvars])

    @staticmethod
    def backward(ctx, g_eps, *g_probs):
        eng = ctx.mod.engine
        S = ctx.mod.grad_scale          # static loss scaling of the fp16 backward (the reference relies on GradScaler, node8.yaml:8)
        ctx.out.g = (g_eps.float() * S).permute(0, 2, 3, 1).contiguous().to(eng.dtype)
        for p, g in zip(ctx.pvars, g_probs):
            if g is not None:
                p.g = (g.float() * S).contiguous()
        eng.zero_lora_grads()
        eng.set_lora_wgrad(ctx.wgrad)
        # direct mode (the trainer owns zero_grad + the flat gradient buffer): weight gradients accumulate in place into
        # param.grad, so autograd gets None for them
        direct = bool(ctx.wgrad and ctx.mod.direct_lora_grads and all(
            p.grad is not None and p.grad.dtype == torch.float32 and p.grad.is_contiguous() for p in eng.lora_params()))
        for l in eng.loras:
            l.direct, l.inv_scale = direct, 1.0 / S
        ctx.tape.backward()
        gx = None
        if ctx.needs_input_grad[1] and ctx.xv.g is not None:
            gx = ops.nhwc_to_nchw_f32(ctx.xv.g, ctx.mod.in_channels, 1.0 / S).to(ctx.x_dtype)
        if ctx.product:
            # the products dy^T x of this pass sit in the engine's accumulators: the trainer projects them once per optimiser
            # step (EngineUNet.finalize_lora_grads); without a trainer they are projected here and handed to autograd
            eng.G_dirty = True
            lg = [None] * len(eng.lora_grads()) if direct else eng.finalize_lora_grads(1.0 / S, into_param_grads=False)
        else:
            lg = eng.lora_grads() if (ctx.wgrad and not direct) else [None] * len(eng.lora_grads())
            if ctx.wgrad and not direct and S != 1.0:
                torch._foreach_mul_([g for g in lg if g is not None], 1.0 / S)
        g_ehs = None
        if ctx.cvar is not None and ctx.cvar.g is not None:
            g_ehs = (ctx.cvar.g.float() / S).reshape(ctx.ehs_meta[1]).to(ctx.ehs_meta[0])
        ctx.tape = ctx.xv = ctx.out = ctx.pvars = ctx.cvar = None
        if ctx.taped_key is not None:          # this signature has now run forward + backward eagerly: later calls may be captured
            ctx.mod._taped.setdefault(ctx.taped_key, {"warm": False, "inst": [], "off": False})["warm"] = True
        return (None, gx, None, g_ehs, None, None, None, *lg)


class _GraphedForward:
    """One captured no-grad UNet forward (fixed shapes): ~700 launches replayed by a single cudaGraphLaunch.
    CUDA graphs instead of a tracing compiler — the executor's Python runs once, at capture."""

    def __init__(self, mod: "EngineUNet", sample, t, ehs, added=None):
        from . import _lib
        eng = mod.engine
        self.mod = mod
        # SDXL's pooled-text / time-id conditioning (TrainableSDPipeline.py:772-784): static copies, refreshed per call (tiny)
        self.added = None if added is None else {k: v.detach().clone() for k, v in added.items()}
        self.x = torch.empty_like(sample)
        self.t = torch.zeros((), dtype=torch.int64, device=sample.device)
        self.x.copy_(sample); self.t.copy_(t.reshape(()))
        # the text context and its per-layer k|v projections are graph INPUTS (static buffers refreshed only when the
        # caller's context tensor or the LoRA weights change): the 32 context GEMMs are not part of the replayed work
        self.ehs16 = ehs.to(eng.dtype).contiguous()
        self.kv = eng.cross_kv(self.ehs16)
        self._key = (ehs, ehs._version, eng.lora_version)

        def run():
            out = eng.forward(None, E.Var(ops.latent_to_nhwc(self.x, eng.dtype, 64), False), self.t, self.ehs16, cross_kv=self.kv,
                              added_cond=self.added)
            return ops.nhwc_to_nchw_f32(out.v, mod.out_channels)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run()                                   # eager warm-up on the side stream (one-time attribute / entry-point setup)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        l0 = _lib.LAUNCH_COUNT
        with torch.cuda.graph(self.graph):
            self.out = run()
        self.launches = _lib.LAUNCH_COUNT - l0

    def __call__(self, sample, t, ehs, added=None):
        from . import _lib
        eng = self.mod.engine
        eng.ensure_merged()                                   # eager, before the replay reads the folded weights
        self.x.copy_(sample); self.t.copy_(t.reshape(()))
        if self.added is not None:
            for k, v in self.added.items():
                v.copy_(added[k])
        if not (self._key[0] is ehs and self._key[1] == ehs._version and self._key[2] == eng.lora_version):
            self.ehs16.copy_(ehs)
            eng.cross_kv(self.ehs16, out=self.kv)
            self._key = (ehs, ehs._version, eng.lora_version)
        self.graph.replay()
        _lib.count_launch(self.launches)
        return self.out.clone()


class _GraphedTaped:
    """One captured TAPED UNet call (fixed shapes): a forward graph and a backward graph that share a private memory pool, so the
    activations the forward leaves behind are exactly what the backward replay reads.  The executor's Python (tape, closures,
    ~2 000 ctypes launches for forward + backward) runs once, at capture; a training step then costs two cudaGraphLaunch calls per
    back-propagated UNet call.  Matters where a call's GPU time is below the host's issue time - SDXL at batch 1 / GPU
    (BASELINE configs[3]) and the discriminator's SD1.5 UNet at n = 1 / 2 - see DESIGN.md section 6.

    Static inputs: sample, timestep, text context (+ its per-layer k|v projections), SDXL's added conditioning, and - for the
    backward - the loss-scaled 16-bit gradient of the noise prediction and the fp32 gradients of the exported probabilities.
    LoRA weight gradients accumulate into the engine's product-gradient arena (persistent buffer, fp32 atomics), exactly as in
    the eager 'product' path; the trainer projects them once per optimiser step.  An instance is busy from its forward replay
    until its backward replay: K back-propagated sampler steps hold K instances."""

    def __init__(self, mod: "EngineUNet", sample, t, ehs, added, capture, needs_xgrad: bool, wgrad: bool):
        from . import _lib
        eng = mod.engine
        self.mod, self.busy, self.wgrad, self.needs_xgrad = mod, False, wgrad, needs_xgrad
        self.x = sample.detach().clone()
        self.t = torch.zeros((), dtype=torch.int64, device=sample.device)
        self.t.copy_(t.reshape(()))
        self.added = None if added is None else {k: v.detach().clone() for k, v in added.items()}
        self.ehs16 = ehs.detach().to(eng.dtype).contiguous()
        eng.ensure_merged(transposed=True)
        if wgrad:
            eng._ensure_G()                                 # persistent arena: must not be born inside the graph's pool
        self.kv = eng.cross_kv(self.ehs16)
        self._key = (ehs, ehs._version, eng.lora_version)
        tape = E.Tape()
        if capture is not None:
            capture.reset()
        l0 = _lib.LAUNCH_COUNT
        self.g_fwd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fwd):
            xv = E.Var(ops.latent_to_nhwc(self.x, eng.dtype, 64), needs_grad=needs_xgrad)
            out = eng.forward(tape, xv, self.t, self.ehs16, capture=capture, added_cond=self.added,
                              lora_mode="train" if wgrad else "frozen", cross_kv=self.kv)
            self.eps = ops.nhwc_to_nchw_f32(out.v, mod.out_channels)
        self.fwd_launches = _lib.LAUNCH_COUNT - l0
        self.places, flat = [], []
        if capture is not None:
            for place in ("down", "mid", "up"):
                self.places.append((place, len(capture.store[place])))
                flat.extend(capture.store[place])
            self.count = capture.count
        self.probs = [p.v for p in flat]
        self.g_out = torch.zeros_like(out.v)                # (n, h, w, 4) 16-bit, loss-scaled by the caller
        self.g_probs = [torch.zeros_like(p) for p in self.probs]
        l0 = _lib.LAUNCH_COUNT
        self.g_bwd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_bwd, pool=self.g_fwd.pool()):
            out.g = self.g_out
            for p, g in zip(flat, self.g_probs):
                p.g = g
            eng.zero_lora_grads()
            eng.set_lora_wgrad(wgrad)
            for l in eng.loras:
                l.direct, l.inv_scale = False, 1.0
            tape.backward()
            self.gx = ops.nhwc_to_nchw_f32(xv.g, mod.in_channels, 1.0) if (needs_xgrad and xv.g is not None) else None
        self.bwd_launches = _lib.LAUNCH_COUNT - l0

    def load_inputs(self, sample, t, ehs, added):
        eng = self.mod.engine
        eng.ensure_merged(transposed=True)                  # eager, before a replay reads the folded weights
        self.x.copy_(sample.detach())
        self.t.copy_(t.reshape(()))
        if self.added is not None:
            for k, v in self.added.items():
                v.copy_(added[k])
        if not (self._key[0] is ehs and self._key[1] == ehs._version and self._key[2] == eng.lora_version):
            self.ehs16.copy_(ehs.detach())
            eng.cross_kv(self.ehs16, out=self.kv)
            self._key = (ehs, ehs._version, eng.lora_version)


class _GraphedUNetFn(torch.autograd.Function):
    """autograd node around one _GraphedTaped instance (same contract as _UNetFn with the trainer's settings: LoRA weight gradients
    stay in the engine's product arena, so autograd gets None for them)"""

    @staticmethod
    def forward(ctx, inst: _GraphedTaped, x, *lora_params):
        from . import _lib
        inst.g_fwd.replay()
        _lib.count_launch(inst.fwd_launches)
        ctx.inst, ctx.x_dtype = inst, x.dtype
        return (inst.eps.clone().to(x.dtype), *[p.detach() for p in inst.probs])

    @staticmethod
    def backward(ctx, g_eps, *g_probs):
        from . import _lib
        inst = ctx.inst
        S = inst.mod.grad_scale
        inst.g_out.copy_((g_eps.float() * S).permute(0, 2, 3, 1))
        for buf, g in zip(inst.g_probs, g_probs):
            if g is None:
                buf.zero_()
            else:
                torch.mul(g, S, out=buf)
        inst.g_bwd.replay()
        _lib.count_launch(inst.bwd_launches)
        if inst.wgrad:
            inst.mod.engine.G_dirty = True                  # the products of this pass sit in the arena (finalize_lora_grads)
        gx = None
        if ctx.needs_input_grad[1] and inst.gx is not None:
            gx = (inst.gx / S).to(ctx.x_dtype)
        inst.busy = False
        return (None, gx, *[None] * (len(ctx.needs_input_grad) - 2))


class EngineUNet(torch.nn.Module):
    """Wraps a diffusers-shaped UNet2DConditionModel (parameters + LoRA layers live there, state-dict compatible) and
    executes it with the B200 kernels."""

    def __init__(self, unet: torch.nn.Module, dtype=torch.float16, train_lora: bool = True):
        super().__init__()
        self.ref = unet                       # parameter owner (LoRA fp32 masters are trained in place)
        self.config = unet.config
        self.engine = E.UNetEngine(unet, dtype)
        self.in_channels = unet.config.in_channels
        self.out_channels = unet.config.out_channels
        self.train_lora = train_lora
        self.capture: Optional[E.AttnCapture] = None
        self.last_probs = None
        self._dtype = dtype
        self.grad_scale = 4096.0 if dtype == torch.float16 else 1.0   # power of two: exact scale / unscale around the 16-bit backward
        self.use_graphs = False                # bench / trainer switch: CUDA-graph the no-grad forwards of the rollout
        self._graphs = {}
        self.direct_lora_grads = False         # trainer switch: LoRA weight gradients accumulate straight into param.grad
        # trainer / bench switch: CUDA-graph the TAPED calls too (forward + backward graph pairs, _GraphedTaped).  Needs the
        # trainer's protocol (direct_lora_grads + finalize_lora_grads once per step) and holds each instance's activations for
        # the life of the module, so it is opt-in
        self.graph_taped = False
        self.max_taped_instances = 8
        self._taped = {}
        self._kv_key, self._kv_val = None, None   # cached 16-bit context + per-layer k|v projections (eager no-grad path)

    @property
    def dtype(self):
        return torch.float32                  # interface dtype of latents / embeddings (fp32 chain, SURVEY 'numerics')

    @property
    def device(self):
        return self.engine.te1.w.device

    def lora_parameters(self):
        return self.engine.lora_params()

    def refresh_lora(self, rebuild_folded: bool = False):
        """call after every optimiser step: re-materialise the 16-bit LoRA operands from the fp32 masters.
        ``rebuild_folded``: also rebuild the LoRA-folded projection weights now (the trainer does, on its side stream) instead of
        lazily at the next forward."""
        self.engine.refresh_lora()
        if rebuild_folded and self.engine._merged_version >= 0:
            self.engine.ensure_merged(transposed=self.engine._merged_t)

    def finalize_lora_grads(self):
        """project the product gradients accumulated by the backward passes since the last call onto the LoRA factors and
        add them to ``param.grad`` (call once after ``loss.backward()``, before the optimiser's all-reduce / step)."""
        if self.engine.G_dirty:
            self.engine.finalize_lora_grads(1.0 / self.grad_scale, into_param_grads=True)

    def _context_kv(self, ehs, store=True):
        """16-bit text context and its k|v projections for every cross-attention layer, cached while the caller keeps passing
        the SAME tensor object (unmodified: ``_version``) and the LoRA weights are unchanged - true for every step of a
        rollout (TrainableSDPipeline.py:132-150 passes one ``prompt_embeds`` to all S UNet calls).  The key holds a
        reference to the tensor, so its storage cannot be recycled under the cache."""
        eng = self.engine
        k = self._kv_key
        if k is not None and k[0] is ehs and k[1] == ehs._version and k[2] == eng.lora_version:
            return self._kv_val
        ctx16 = ehs.detach().to(eng.dtype).contiguous()
        val = (ctx16, eng.cross_kv(ctx16))
        if store:                                  # one-off contexts (the attrcon half-batches) do not evict the rollout's entry
            self._kv_val = val
            self._kv_key = (ehs, ehs._version, eng.lora_version)
        return val

    def enable_gradient_checkpointing(self):
        pass                                  # never recomputes: activations of K steps fit in 180 GB

    def new_step(self):
        """trainer hook (start of an optimiser step): graph instances whose backward never ran (a forward without a backward)
        become available again"""
        for slot in self._taped.values():
            for inst in slot["inst"]:
                inst.busy = False

    @staticmethod
    def _taped_key(sample, ehs, added, capture, needs_xgrad, wgrad):
        return (tuple(sample.shape), tuple(ehs.shape), sample.dtype, ehs.dtype, needs_xgrad, wgrad,
                None if capture is None else (tuple(capture.places), capture.sample_from),
                None if added is None else tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(added.items())))

    def _taped_instance(self, sample, t, ehs, added, capture, needs_xgrad, wgrad):
        """a free captured (forward, backward) graph pair for this call signature, or None -> run the call eagerly.  A signature is
        captured only after one EAGER forward + backward of it has completed (``_UNetFn.backward`` marks it): every kernel variant
        the pass launches has then been loaded and configured - CUDA's lazy module loading inside a stream capture invalidates the
        capture (cudaErrorStreamCaptureInvalidated, profiles/r02_gpu_tests_run7.log)."""
        eng = self.engine
        if not (self.graph_taped and sample.is_cuda and not ehs.requires_grad):
            return None
        if wgrad and not (eng.lora_train_impl == "product" and self.direct_lora_grads):
            return None
        key = self._taped_key(sample, ehs, added, capture, needs_xgrad, wgrad)
        slot = self._taped.setdefault(key, {"warm": False, "inst": [], "off": False})
        if not slot["warm"]:
            return None
        inst = next((i for i in slot["inst"] if not i.busy), None)
        if inst is None:
            if slot["off"] or len(slot["inst"]) >= self.max_taped_instances:
                return None
            try:
                inst = _GraphedTaped(self, sample, t, ehs, added, capture, needs_xgrad, wgrad)
            except torch.cuda.OutOfMemoryError:
                slot["off"] = True            # no room for another resident activation set: this signature stays eager
                torch.cuda.empty_cache()
                return None
            slot["inst"].append(inst)
        inst.load_inputs(sample, t, ehs, added)
        inst.busy = True
        return inst

    def forward(self, sample, timestep, encoder_hidden_states, cross_attention_kwargs=None, added_cond_kwargs=None,
                return_dict=False):
        t = timestep if torch.is_tensor(timestep) else torch.tensor(timestep, device=sample.device)
        params = self.engine.lora_params()
        want_grad = torch.is_grad_enabled() and (sample.requires_grad or encoder_hidden_states.requires_grad or
                                                 (self.train_lora and any(p.requires_grad for p in params)))
        capture = self.capture
        if capture is not None:
            capture.reset()
        inst = None
        if want_grad and self.graph_taped:
            wgrad = bool(self.train_lora and any(p.requires_grad for p in params))
            inst = self._taped_instance(sample, t, encoder_hidden_states, added_cond_kwargs, capture, bool(sample.requires_grad), wgrad)
        if inst is not None:
            outs = _GraphedUNetFn.apply(inst, sample, *params)
            eps, probs = outs[0], outs[1:]
            if capture is not None:           # the AttentionStore protocol on the replayed call: same counts, Vars over the outputs
                capture.reset()
                capture.count = inst.count
                i = 0
                for place, cnt in inst.places:
                    capture.store[place] = [E.Var(probs[i + j]) for j in range(cnt)]
                    i += cnt
        elif want_grad:
            ckv = self._context_kv(encoder_hidden_states, store=capture is None)
            outs = _UNetFn.apply(self, sample, t, encoder_hidden_states, added_cond_kwargs, capture, ckv, *params)
            eps, probs = outs[0], outs[1:]
        elif self.use_graphs and capture is None and sample.is_cuda:
            key = (tuple(sample.shape), tuple(encoder_hidden_states.shape), sample.dtype, encoder_hidden_states.dtype,
                   None if added_cond_kwargs is None else tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(added_cond_kwargs.items())))
            if key not in self._graphs:
                self.engine.ensure_merged()
                self._graphs[key] = _GraphedForward(self, sample.detach(), t, encoder_hidden_states, added_cond_kwargs)
            eps = self._graphs[key](sample.detach(), t, encoder_hidden_states, added_cond_kwargs).to(sample.dtype)
            probs = ()
        else:
            eng = self.engine
            ctx16, kv = self._context_kv(encoder_hidden_states)
            out = eng.forward(None, E.Var(ops.latent_to_nhwc(sample, eng.dtype, 64), False), t,
                              ctx16, capture=capture, added_cond=added_cond_kwargs, cross_kv=kv)
            eps = ops.nhwc_to_nchw_f32(out.v, self.out_channels).to(sample.dtype)
            probs = ()
        if capture is not None:
            # re-point the captured Vars' values at the autograd-visible tensors so losses differentiate through them
            i = 0
            for place in ("down", "mid", "up"):
                for p in capture.store[place]:
                    if i < len(probs):
                        p.v = probs[i]
                    i += 1
        if not return_dict:
            return (eps,)
        return SimpleNamespace(sample=eps)


class _VAEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod: "EngineVAE", z):
        eng = mod.engine
        tape = E.Tape()
        zv = E.Var(ops.latent_to_nhwc(z, eng.dtype, 64))
        out = eng.forward(tape, zv)
        ctx.tape, ctx.zv, ctx.out, ctx.mod, ctx.zdtype = tape, zv, out, mod, z.dtype
        return ops.nhwc_to_nchw_f32(out.v, 3).to(z.dtype)

    @staticmethod
    def backward(ctx, g):
        eng = ctx.mod.engine
        S = ctx.mod.grad_scale
        ctx.out.g = (g.float() * S).permute(0, 2, 3, 1).contiguous().to(eng.dtype)
        ctx.tape.backward()
        gz = ops.nhwc_to_nchw_f32(ctx.zv.g, 4, 1.0 / S).to(ctx.zdtype)
        ctx.tape = ctx.zv = ctx.out = None
        return None, gz


class EngineVAE(torch.nn.Module):
    def __init__(self, vae: torch.nn.Module, dtype=torch.float16):
        super().__init__()
        self.ref = vae
        self.config = vae.config
        self.engine = E.VAEDecoderEngine(vae, dtype)
        self.grad_scale = 4096.0 if dtype == torch.float16 else 1.0

    @property
    def dtype(self):
        return torch.float32

    def decode(self, z, return_dict=False):
        if torch.is_grad_enabled() and z.requires_grad:
            x = _VAEFn.apply(self, z)
        else:
            eng = self.engine
            out = eng.forward(None, E.Var(ops.latent_to_nhwc(z, eng.dtype, 64), False))
            x = ops.nhwc_to_nchw_f32(out.v, 3).to(z.dtype)
        if not return_dict:
            return (x,)
        return SimpleNamespace(sample=x)
