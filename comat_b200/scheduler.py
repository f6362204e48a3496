"""DDPM ancestral sampler used inside the training step (the reference allows only DDPM: training_utils/arguments.py:283-288,
training_utils/pipeline.py:50-59) + classifier-free-guidance rescale.  Scalar schedule on the host, latent math in fp32.

Config = the SD checkpoint's scheduler config: scaled-linear betas 0.00085 -> 0.012 over 1000 steps,
steps_offset 1, "leading" spacing, epsilon prediction, variance "fixed_small", no sample clipping.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch


class DDPMScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1, variance_type="fixed_small"):
        if variance_type in ("learned", "learned_range"):          # training_utils/pipeline.py:51-59
            variance_type = "fixed_small"
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                      beta_schedule="scaled_linear", steps_offset=steps_offset, timestep_spacing="leading",
                                      clip_sample=False, prediction_type="epsilon", variance_type=variance_type)
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)

    @classmethod
    def from_config(cls, config, **kw):
        return cls(**kw)

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (torch.arange(0, num_inference_steps) * ratio).flip(0).to(torch.int64) + self.config.steps_offset
        self._ts_host = ts.tolist()
        if device is None:
            self.timesteps = ts
        else:
            # one pageable H2D copy is a stream sync; the schedule only depends on (steps, device), so keep the device copy
            key = (num_inference_steps, str(device))
            cache = self.__dict__.setdefault("_ts_dev_cache", {})
            if key not in cache:
                cache[key] = ts.to(device)
            self.timesteps = cache[key]

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step_coefficients(self, t: int):
        """x_prev = c_eps * eps + c_x * x + sigma * z   (epsilon prediction folded in)."""
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[t])
        a_prev = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else 1.0
        b_t, b_prev = 1.0 - a_t, 1.0 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1.0 - cur_alpha
        c_x0 = (a_prev ** 0.5) * cur_beta / b_t
        c_xt = (cur_alpha ** 0.5) * b_prev / b_t
        var = max(b_prev / b_t * cur_beta, 1e-20)
        # x0 = (x - sqrt(b_t) eps) / sqrt(a_t)
        c_x = c_x0 / a_t ** 0.5 + c_xt
        c_eps = -c_x0 * (b_t ** 0.5) / a_t ** 0.5
        return c_eps, c_x, (var ** 0.5 if t > 0 else 0.0), a_t, b_t

    def step(self, model_output, timestep, sample, generator=None, return_dict=True, variance_noise=None, **_):
        t = int(timestep)
        c_eps, c_x, sigma, a_t, b_t = self.step_coefficients(t)
        prev = c_eps * model_output + c_x * sample
        if sigma > 0.0:
            if variance_noise is None:
                variance_noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype,
                                             device=generator.device if generator is not None else model_output.device)
            prev = prev + sigma * variance_noise.to(prev.device)
        if not return_dict:
            return (prev,)
        x0 = (sample - (b_t ** 0.5) * model_output) / a_t ** 0.5
        return SimpleNamespace(prev_sample=prev, pred_original_sample=x0)


def rescale_noise_cfg(noise_cfg, noise_pred_text, guidance_rescale=0.0):
    """arXiv 2305.08891 sec. 3.4 (diffusers pipeline_stable_diffusion.rescale_noise_cfg; TrainableSDPipeline.py:159-161)."""
    dims = list(range(1, noise_pred_text.ndim))
    std_text = noise_pred_text.std(dim=dims, keepdim=True)
    std_cfg = noise_cfg.std(dim=dims, keepdim=True)
    rescaled = noise_cfg * (std_text / std_cfg)
    return guidance_rescale * rescaled + (1 - guidance_rescale) * noise_cfg


class _CfgDdpmStep(torch.autograd.Function):
    """x_prev = c_x x + c_eps (e_u + s (e_c - e_u)) + sigma z  — one fused launch forward, one backward (csrc/optim.cu)."""

    @staticmethod
    def forward(ctx, eps, x, noise, s, c_eps, c_x, sigma, cfg):
        import ctypes as C
        from . import _lib
        eps, x = eps.float().contiguous(), x.float().contiguous()
        out = torch.empty_like(x)
        z = noise.float().contiguous() if (noise is not None and sigma > 0) else None
        _lib.check(_lib.lib().comat_cfg_ddpm_step_fwd(C.c_void_p(eps.data_ptr()), C.c_void_p(x.data_ptr()),
                                                      C.c_void_p(z.data_ptr() if z is not None else None), C.c_void_p(out.data_ptr()),
                                                      C.c_longlong(x.numel()), C.c_float(s), C.c_float(c_eps), C.c_float(c_x),
                                                      C.c_float(sigma), int(cfg), _lib.stream_ptr()), "cfg_ddpm_step_fwd")
        _lib.count_launch()
        ctx.k = (s, c_eps, c_x, cfg, eps.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        import ctypes as C
        from . import _lib
        s, c_eps, c_x, cfg, eshape = ctx.k
        g = g.float().contiguous()
        d_eps = torch.empty(eshape, dtype=torch.float32, device=g.device) if ctx.needs_input_grad[0] else None
        dx = torch.empty_like(g) if ctx.needs_input_grad[1] else None
        _lib.check(_lib.lib().comat_cfg_ddpm_step_bwd(C.c_void_p(g.data_ptr()), C.c_void_p(d_eps.data_ptr() if d_eps is not None else None),
                                                      C.c_void_p(dx.data_ptr() if dx is not None else None), C.c_longlong(g.numel()),
                                                      C.c_float(s), C.c_float(c_eps), C.c_float(c_x), int(cfg), _lib.stream_ptr()),
                   "cfg_ddpm_step_bwd")
        _lib.count_launch()
        return d_eps, dx, None, None, None, None, None, None


def fused_cfg_ddpm_step(scheduler: DDPMScheduler, eps, t, latents, guidance_scale, cfg, noise=None, generator=None):
    """CFG combine + scheduler.step in one kernel (CUDA tensors only).  ``eps`` is the raw UNet output ([uncond; cond] if cfg)."""
    c_eps, c_x, sigma, _, _ = scheduler.step_coefficients(int(t))
    if sigma > 0 and noise is None:
        noise = torch.randn(latents.shape, generator=generator, dtype=torch.float32,
                            device=generator.device if generator is not None else latents.device).to(latents.device)
    return _CfgDdpmStep.apply(eps, latents, noise, float(guidance_scale), c_eps, c_x, sigma, bool(cfg))
