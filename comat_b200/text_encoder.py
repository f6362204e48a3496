"""CLIP text encoders of the step: what ``encode_prompt`` (TrainableSDPipeline.py:227-424; the SDXL twin is the
un-vendored diffusers ``StableDiffusionXLPipeline.encode_prompt``) calls once per optimiser step under ``no_grad``.

CLIP-L (12 x 768, 12 heads, quick-GELU) for SD1.5 / SDXL encoder 1 and OpenCLIP-bigG (32 x 1280, 20 heads, GELU, text
projection) for SDXL encoder 2 run on the same tcgen05 GEMM / attention kernels and HBM-bound LayerNorm kernels as the UNet:
pre-LN blocks, q|k|v as ONE GEMM whose column slices the causal attention kernel reads in place, head dim 64.

quick-GELU ``x * sigmoid(1.702 x)`` needs no kernel of its own: it equals ``silu(1.702 x) / 1.702``, so fc1 runs with
``alpha = 1.702`` (bias pre-scaled) and the SiLU epilogue, and fc2 with ``alpha = 1 / 1.702``.

The parameter owner is an HF ``CLIPTextModel`` / ``CLIPTextModelWithProjection`` (real checkpoint or random init); weights are
packed once.

``--train_text_encoder_lora`` (training_script.py:227-255; SD1.5 form): ``install_text_lora`` puts a ``LoRALinearLayer(rank)`` on
q_proj / k_proj / v_proj / out_proj of every block (the projections diffusers' ``LoraLoaderMixin._modify_text_encoder`` patches -
un-vendored), and a grad-enabled call runs ``forward_taped``: the same kernels behind an explicit tape (projections un-fused so
each carries its LoRA branch, attention backward with the causal mask), LoRA weight gradients returned to autograd, the data
gradient arriving from the UNet executor's d(encoder_hidden_states).  Host logic checked against HF autograd on emulated ops
(tests/test_text_encoder_cpu.py); this training path has NOT been run on the B200 yet.  ``--tune_text_encoder`` (full weights)
is not implemented.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch

from . import attention as attn_ops
from . import engine as E
from . import ops

_QG = 1.702


class _FusedQKV:
    def __init__(self, attn, dtype):
        w = torch.cat([attn.q_proj.weight.detach(), attn.k_proj.weight.detach(), attn.v_proj.weight.detach()], 0)
        b = torch.cat([attn.q_proj.bias.detach(), attn.k_proj.bias.detach(), attn.v_proj.bias.detach()], 0)
        self.w = w.to(dtype).contiguous()            # [3C, C] K-major
        self.bias = b.float().contiguous()


def install_text_lora(model, rank: int, up_std: float = 0.0):
    """LoRA(rank) on the attention projections of every CLIP block; fp32 masters, down ~ N(0, 1/rank), up = 0 (``up_std`` > 0 only
    for gradient checks).  Returns the trainable parameters in module order (what ``get_trainable_parameters`` collects as
    ``text_lora_parameters``, training_utils/pipeline.py:172-181)."""
    from .containers import LoRALinearLayer
    params = []
    for lyr in model.text_model.encoder.layers:
        for lin in (lyr.self_attn.q_proj, lyr.self_attn.k_proj, lyr.self_attn.v_proj, lyr.self_attn.out_proj):
            l = LoRALinearLayer(lin.in_features, lin.out_features, rank).to(lin.weight.device, torch.float32)
            if up_std > 0:
                torch.nn.init.normal_(l.up.weight, std=up_std)
            lin.lora_layer = l
            params.extend(l.parameters())
    return params


def _quick_gelu(tape, x: E.Var) -> E.Var:
    """x * sigmoid(1.702 x) = silu(1.702 x) / 1.702 with the elementwise kernels; d/dx = silu'(1.702 x)."""
    s_ = ops.elementwise("scale", x.v, alpha=_QG)
    out = E.Var(ops.elementwise("scale", ops.elementwise("silu", s_), alpha=1.0 / _QG))
    if tape is not None:
        def bwd():
            if out.g is not None:
                E._acc(x, ops.elementwise("silu_bwd", s_, out.g))
        tape.record(bwd)
    return out


class ClipTextEngine:
    """``forward``: the frozen encoder of the CoMat path (training_script.py:210-226) - fused q|k|v, no tape, CUDA-graph friendly.
    ``forward_taped``: the same network with un-fused projections and explicit LoRA branches for ``--train_text_encoder_lora``."""

    def __init__(self, model, dtype=torch.float16):
        cfg = model.config
        tm = model.text_model
        self.dtype = dtype
        self.heads = cfg.num_attention_heads
        self.dim = cfg.hidden_size
        self.eos_token_id = cfg.eos_token_id
        act = cfg.hidden_act
        if act not in ("quick_gelu", "gelu"):
            raise NotImplementedError(f"CLIP text hidden_act {act!r} (SD1.5 / SDXL use quick_gelu and gelu)")
        self.quick = act == "quick_gelu"
        self.tok = tm.embeddings.token_embedding.weight.detach()
        self.pos = tm.embeddings.position_embedding.weight.detach()
        self.layers = []
        for lyr in tm.encoder.layers:
            fc1, fc2 = E.LinW(lyr.mlp.fc1, dtype), E.LinW(lyr.mlp.fc2, dtype)
            if self.quick:
                fc1.bias = (fc1.bias * _QG).contiguous()
            self.layers.append(dict(n1=E.NormW(lyr.layer_norm1), qkv=_FusedQKV(lyr.self_attn, dtype),
                                    out=E.LinW(lyr.self_attn.out_proj, dtype), n2=E.NormW(lyr.layer_norm2), fc1=fc1, fc2=fc2))
        # training path: un-fused projections with their LoRA branches (built only when install_text_lora ran before packing)
        self.loras, self.tlayers, self._arenas = [], [], None
        if getattr(tm.encoder.layers[0].self_attn.q_proj, "lora_layer", None) is not None:
            for lyr in tm.encoder.layers:
                att = lyr.self_attn
                lins = (att.q_proj, att.k_proj, att.v_proj, att.out_proj)
                lw = [E.LoRAW(l.lora_layer, dtype) for l in lins]
                self.loras.extend(lw)
                self.tlayers.append(dict(lin=[E.LinW(l, dtype) for l in lins], lora=lw, fc1=E.LinW(lyr.mlp.fc1, dtype)))
        self.final_ln = E.NormW(tm.final_layer_norm)
        proj = getattr(model, "text_projection", None)
        self.proj = None if proj is None else proj.weight.detach().to(dtype).contiguous()      # [proj_dim, C], no bias

    def refresh_lora(self):
        """16-bit operand images of the LoRA factors follow the fp32 masters (call after every optimiser step)."""
        if self.loras:
            if self._arenas is None:
                self._arenas = E._LoRAArenas(self.loras, self.dtype)
            self._arenas.refresh()

    def lora_params(self):
        return [p for l in self.loras for p in (l.down, l.up)]

    def forward_taped(self, tape, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> E.Var:
        """taped forward with the LoRA branches explicit -> Var last_hidden_state (B, T, C); backward leaves d down / d up of
        every adapter in ``lora.g_down / g_up``."""
        from .blip_engine import _gelu, _mha
        if not self.loras:
            raise RuntimeError("forward_taped needs install_text_lora(model, rank) before the executor is built")
        if self._arenas is None:
            self.refresh_lora()
        B, T = input_ids.shape
        emb = (self.tok[input_ids] + self.pos[:T][None]).to(self.dtype).contiguous()
        h = E.Var(emb, needs_grad=False)                      # embeddings are frozen: the tape ends at the first block's adapters
        lens = (torch.full((B,), T, dtype=torch.int32, device=emb.device) if attention_mask is None
                else attention_mask.sum(1).to(torch.int32).contiguous())
        for L, TL in zip(self.layers, self.tlayers):
            (lq, lk, lv, lo), (aq, ak, av, ao) = TL["lin"], TL["lora"]
            y = E.layernorm(tape, h, L["n1"])
            a = _mha(tape, E.linear(tape, y, lq, aq), E.linear(tape, y, lk, ak), E.linear(tape, y, lv, av), self.heads,
                     kv_lens=lens, causal=True)
            h = E.linear(tape, a, lo, ao, residual=h)
            y = E.layernorm(tape, h, L["n2"])
            f = E.linear(tape, y, TL["fc1"])
            f = _quick_gelu(tape, f) if self.quick else _gelu(tape, f)
            h = E.linear(tape, f, L["fc2"], residual=h)
        return E.layernorm(tape, h, self.final_ln)

    def final_layer_norm(self, h: torch.Tensor) -> torch.Tensor:
        x = h.to(self.dtype).contiguous()
        return ops.layernorm_fwd(x, self.final_ln.gamma, self.final_ln.beta, self.final_ln.eps)[0]

    def forward(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None, all_hidden: bool = False):
        """input_ids (B, T) int64 -> (last_hidden_state (B,T,C) 16-bit, pooled (B,C), text_embeds | None, hidden_states | None)."""
        B, T = input_ids.shape
        C_ = self.dim
        h = (self.tok[input_ids] + self.pos[:T][None]).to(self.dtype).contiguous()
        # right-padded prompts (the CLIP tokenizer pads after EOS): a per-sample key length on top of the causal mask
        lens = (torch.full((B,), T, dtype=torch.int32, device=h.device) if attention_mask is None
                else attention_mask.sum(1).to(torch.int32).contiguous())
        hidden = [h] if all_hidden else None
        for L in self.layers:
            y = ops.layernorm_fwd(h, L["n1"].gamma, L["n1"].beta, L["n1"].eps)[0]
            qkv = ops.gemm([y.reshape(B * T, C_)], [L["qkv"].w], bias=L["qkv"].bias).reshape(B, T, 3 * C_)
            a, _, _ = attn_ops.attention_fwd_native(qkv[..., :C_], qkv[..., C_:2 * C_], qkv[..., 2 * C_:], self.heads,
                                                    kv_lens=lens, causal=True)
            h = ops.gemm([a.reshape(B * T, C_)], [L["out"].w], bias=L["out"].bias, residual=h.reshape(B * T, C_)).reshape(B, T, C_)
            y = ops.layernorm_fwd(h, L["n2"].gamma, L["n2"].beta, L["n2"].eps)[0].reshape(B * T, C_)
            if self.quick:
                f = ops.gemm([y], [L["fc1"].w], bias=L["fc1"].bias, alpha=_QG, act="silu")
                h = ops.gemm([f], [L["fc2"].w], bias=L["fc2"].bias, alpha=1.0 / _QG, residual=h.reshape(B * T, C_)).reshape(B, T, C_)
            else:
                f = ops.gemm([y], [L["fc1"].w], bias=L["fc1"].bias, act="gelu")
                h = ops.gemm([f], [L["fc2"].w], bias=L["fc2"].bias, residual=h.reshape(B * T, C_)).reshape(B, T, C_)
            if all_hidden:
                hidden.append(h)
        last = ops.layernorm_fwd(h, self.final_ln.gamma, self.final_ln.beta, self.final_ln.eps)[0]
        ids = input_ids.to(torch.int32)
        # pooled token: transformers CLIPTextTransformer.forward — the legacy configs (eos_token_id == 2, what the SD / SDXL
        # checkpoints ship) take the arg-max id (EOS = 49407 is the largest), newer ones the first EOS position
        eos = ids.argmax(-1) if self.eos_token_id == 2 else (ids == self.eos_token_id).int().argmax(-1)
        pooled = last[torch.arange(B, device=last.device), eos].contiguous()
        embeds = None
        if self.proj is not None:
            embeds = ops.gemm([pooled], [self.proj])
        return last, pooled, embeds, hidden


class _TextOutput:
    """ordered-field result with the tuple indexing the reference relies on (``out[0]``; ``out[-1][-(clip_skip + 1)]`` when
    hidden states were requested, TrainableSDPipeline.py:325-335) and HF attribute names."""

    def __init__(self, **fields):
        self._keys = [k for k, v in fields.items() if v is not None]
        for k, v in fields.items():
            setattr(self, k, v)

    def __getitem__(self, i):
        if isinstance(i, str):
            return getattr(self, i)
        return getattr(self, self._keys[i])

    def keys(self):
        return list(self._keys)

    def __len__(self):
        return len(self._keys)


class _GraphedEncoder:
    """One captured encoder forward for a fixed (batch, length, mask?, hidden?) signature: ~100 launches of a launch-bound pass
    (M = B x 77 rows) replayed by one cudaGraphLaunch.  Token ids / mask are static inputs, outputs are static buffers."""

    def __init__(self, eng: ClipTextEngine, ids, mask, all_hidden):
        from . import _lib
        self.ids = ids.clone()
        self.mask = None if mask is None else mask.clone()
        run = lambda: eng.forward(self.ids, self.mask, all_hidden=all_hidden)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run()                                   # eager warm-up off the capturing stream (one-time entry-point setup)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        l0 = _lib.LAUNCH_COUNT
        with torch.cuda.graph(self.graph):
            self.out = run()
        self.launches = _lib.LAUNCH_COUNT - l0

    def __call__(self, ids, mask):
        from . import _lib
        self.ids.copy_(ids)
        if self.mask is not None:
            self.mask.copy_(mask)
        self.graph.replay()
        _lib.count_launch(self.launches)
        return self.out                             # static buffers: the caller converts (copies) before the next replay


class _TextLoRAFn(torch.autograd.Function):
    """one autograd node for a grad-enabled encoder call: explicit tape inside, LoRA weight gradients out."""

    @staticmethod
    def forward(ctx, mod: "EngineCLIPText", ids, mask, *lora_params):
        eng = mod.engine
        tape = E.Tape()
        for l in eng.loras:
            l.wgrad, l.direct, l.g_down, l.g_up = True, False, None, None
        out = eng.forward_taped(tape, ids, mask)
        ctx.tape, ctx.out, ctx.mod = tape, out, mod
        return out.v.float()

    @staticmethod
    def backward(ctx, g):
        eng = ctx.mod.engine
        S = ctx.mod.grad_scale                       # static loss scaling of the 16-bit backward, as for the UNet
        ctx.out.g = (g.float() * S).to(eng.dtype).contiguous()
        ctx.tape.backward()
        grads = []
        for l in eng.loras:
            grads += [None if l.g_down is None else l.g_down / S, None if l.g_up is None else l.g_up / S]
            l.g_down = l.g_up = None
        ctx.tape = ctx.out = None
        return (None, None, None, *grads)


class EngineCLIPText(torch.nn.Module):
    """HF call surface ``text_encoder(input_ids, attention_mask=None, output_hidden_states=False)`` backed by ``ClipTextEngine``.
    Outputs are fp32 (the pipelines' interface dtype, like ``EngineUNet.dtype``); the arithmetic is 16-bit on the tensor cores.
    ``use_graphs`` (default on for CUDA inputs): one captured CUDA graph per call signature, at most ``MAX_GRAPHS`` kept."""

    MAX_GRAPHS = 8

    def __init__(self, model, dtype=torch.float16):
        super().__init__()
        self.ref = model
        self.config = model.config
        self.engine = ClipTextEngine(model, dtype)
        self.with_projection = self.engine.proj is not None
        self.text_model = SimpleNamespace(final_layer_norm=lambda h: self.engine.final_layer_norm(h).float())
        self.use_graphs = True
        self._graphs = {}
        self.grad_scale = 4096.0 if dtype == torch.float16 else 1.0

    def lora_parameters(self):
        return self.engine.lora_params()

    def refresh_lora(self):
        """after an optimiser step on the text LoRA: new operand images; captured graphs of the (LoRA-free) frozen forward stay
        valid only for encoders without adapters, so they are dropped here."""
        self.engine.refresh_lora()
        self._graphs = {}

    @property
    def dtype(self):
        return torch.float32

    @property
    def device(self):
        return self.engine.tok.device

    def forward(self, input_ids, attention_mask=None, output_hidden_states: bool = False, **_):
        ids = input_ids.to(self.device)
        mask = None if attention_mask is None else attention_mask.to(self.device)
        params = self.engine.lora_params()
        if params:
            if output_hidden_states or self.with_projection:
                raise NotImplementedError("text-encoder LoRA is implemented for the SD1.5 call form (last_hidden_state) only")
            if torch.is_grad_enabled() and any(p.requires_grad for p in params):
                last = _TextLoRAFn.apply(self, ids, mask, *params)
            else:                                     # adapters present but no gradient wanted: same executor, no tape kept
                with torch.no_grad():
                    last = self.engine.forward_taped(None, ids, mask).v.float()
            eos = ids.to(torch.int32).argmax(-1) if self.engine.eos_token_id == 2 else (ids == self.engine.eos_token_id).int().argmax(-1)
            pooled = last.detach()[torch.arange(ids.shape[0], device=last.device), eos]
            return _TextOutput(last_hidden_state=last, pooler_output=pooled, hidden_states=None)
        with torch.no_grad():
            return self._frozen_forward(ids, mask, output_hidden_states)

    def _frozen_forward(self, ids, mask, output_hidden_states):
        if self.use_graphs and ids.is_cuda:
            key = (tuple(ids.shape), mask is not None, bool(output_hidden_states))
            g = self._graphs.get(key)
            if g is None:
                if len(self._graphs) >= self.MAX_GRAPHS:
                    self._graphs.pop(next(iter(self._graphs)))
                g = self._graphs[key] = _GraphedEncoder(self.engine, ids, mask, bool(output_hidden_states))
            last, pooled, embeds, hidden = g(ids, mask)
        else:
            last, pooled, embeds, hidden = self.engine.forward(ids, mask, all_hidden=bool(output_hidden_states))
        hs = None if hidden is None else tuple(x.float() for x in hidden)       # .float() copies out of the graph's buffers
        if self.with_projection:           # CLIPTextModelOutput: text_embeds first (SDXL reads out[0] as the pooled vector)
            return _TextOutput(text_embeds=embeds.float(), last_hidden_state=last.float(), hidden_states=hs)
        return _TextOutput(last_hidden_state=last.float(), pooler_output=pooled.float(), hidden_states=hs)
