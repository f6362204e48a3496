"""Image-space glue of the reward path: crop -> bicubic-antialias resize to 384 -> CLIP normalise (fwd + bwd).

TODO(native): `comat_resize_bicubic_aa` (separable two-pass HBM-bound kernel).  Until it lands this uses
aten's `_upsample_bicubic2d_aa` on the GPU (library call, counted in LIBRARY_CALLS)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

LIBRARY_CALLS = 0


def resize_bicubic_aa_normalize(images: torch.Tensor, size: int, mean, std) -> torch.Tensor:
    global LIBRARY_CALLS
    LIBRARY_CALLS += 1
    x = F.interpolate(images.float(), size=(size, size), mode="bicubic", antialias=True, align_corners=False)
    m = torch.tensor(mean, device=x.device, dtype=x.dtype).view(1, -1, 1, 1)
    s = torch.tensor(std, device=x.device, dtype=x.dtype).view(1, -1, 1, 1)
    return (x - m) / s
