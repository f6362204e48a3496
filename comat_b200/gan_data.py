"""GAN ground-truth data: the producer ``tools/gan_gt_generate.py`` and the reader ``training_utils/gan_dataset.py``.

Producer (gan_gt_generate.py:109-193): for each batch of prompts, a plain 50-step DDPM sampling at guidance 7.5 on the *frozen
base* pipeline, ``output_type='latent'``; every latent is written as ``<save_dir>/latents/<uid>.pt`` - ``torch.save`` of an fp32
``(4, H/8, W/8)`` CPU tensor - and one JSON line ``{"prompt": ..., "file_path": ...}`` per sample is appended to the jsonl
index; ``use_cache`` skips prompts already in the index (:92-95).  It is a pure-inference use of the step's own UNet kernels:
the no-grad forwards replay one captured CUDA graph.

Reader (gan_dataset.py:28-74): index as ``.txt`` / ``.jsonl`` / ``.json``; an item is ``{'text': prompt, 'latents': tensor, **other
keys}`` with ``file_path`` lists sampled at random.  The reference fetches the bytes through a Ceph client (``aoss_client``,
out of scope); here the opener is a parameter, local disk by default.
"""
from __future__ import annotations

import io
import json
import os
import random
import uuid
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import torch

_ALPHABET = "23456789ABCDEFGHJKLMNPQRSTUVWXYZabcdefghijkmnopqrstuvwxyz"


def short_uid(u: Optional[uuid.UUID] = None) -> str:
    """``shortuuid.uuid()`` (gan_gt_generate.py:183; the package is not in this image): a uuid4 in base 57 over shortuuid's
    alphabet, most significant digit first, padded to 22 characters."""
    n = (u or uuid.uuid4()).int
    digits = []
    while n:
        n, r = divmod(n, 57)
        digits.append(_ALPHABET[r])
    return "".join(reversed(digits)).rjust(22, _ALPHABET[0])


def read_jsonl(path: str) -> List[dict]:
    out = []
    if os.path.exists(path):
        with open(path, "r") as f:
            for line in f:
                if line.strip():
                    out.append(json.loads(line))
    return out


def read_prompts(data_path: str, start: int = 0, end: Optional[int] = None) -> List[str]:
    """gan_gt_generate.py:82-91: one prompt per line of a .txt, or a JSON list; sliced [start:end]."""
    if "txt" in data_path:
        with open(data_path, "r") as f:
            ann = [line.strip() for line in f]
    elif "json" in data_path:
        ann = json.load(open(data_path, "r"))
    else:
        raise NotImplementedError(data_path)
    return ann[start:end]


@torch.no_grad()
def generate_gan_ground_truth(pipeline, prompts: Sequence[str], save_prompt_path: str, batch_size: int = 8,
                              num_inference_steps: int = 50, guidance_scale: float = 7.5, height: int = 512, width: int = 512,
                              generator: Optional[torch.Generator] = None, use_cache: bool = False,
                              uid_fn: Callable[[], str] = short_uid, use_graphs: bool = True, **call_kwargs) -> int:
    """the loop of gan_gt_generate.py:169-193.  ``pipeline`` is a TrainableSD(XL)Pipeline built with text encoder(s), or
    pass ``prompt_embeds_fn(prompts) -> dict`` in ``call_kwargs`` to feed pre-computed embeddings.  Returns #samples written."""
    save_dir = os.path.dirname(save_prompt_path)
    os.makedirs(os.path.join(save_dir, "latents"), exist_ok=True)
    prompts = list(prompts)
    if use_cache:                                                                   # :92-95 (set difference: order not kept)
        done = {inst["prompt"] for inst in read_jsonl(save_prompt_path)}
        prompts = [p for p in dict.fromkeys(prompts) if p not in done]
    embeds_fn = call_kwargs.pop("prompt_embeds_fn", None)
    unet = pipeline.unet
    prev_graphs = getattr(unet, "use_graphs", False)
    if use_graphs and hasattr(unet, "use_graphs"):
        unet.use_graphs = True
    written = 0
    try:
        for i in range(0, len(prompts), batch_size):
            chunk = prompts[i:i + batch_size]
            extra = embeds_fn(chunk) if embeds_fn is not None else {}
            latents = pipeline(None if "prompt_embeds" in extra else chunk, height=height, width=width,
                               num_inference_steps=num_inference_steps, generator=generator, guidance_scale=guidance_scale,
                               guidance_rescale=0.0, output_type="latent", **extra, **call_kwargs).images
            host = latents.detach().float().cpu()                                    # ONE D2H copy per batch
            lines = []
            for j, prompt in enumerate(chunk):
                path = os.path.join(save_dir, "latents", f"{uid_fn()}.pt")
                torch.save(host[j].clone(), path)                                    # :184 fp32 (4, H/8, W/8)
                lines.append(json.dumps({"prompt": prompt, "file_path": path}))     # :185-188
            with open(save_prompt_path, "a") as f:                                   # :40-43, :191
                f.write("\n".join(lines) + "\n")
            written += len(chunk)
    finally:
        if hasattr(unet, "use_graphs"):
            unet.use_graphs = prev_graphs
    return written


def _open_local(path: str) -> bytes:
    with open(path, "rb") as f:
        return f.read()


class Gan_Dataset(torch.utils.data.Dataset):
    """gan_dataset.py:28-74.  ``args.training_prompts`` names the index; ``opener(path) -> bytes`` replaces the Ceph client."""

    def __init__(self, args, opener: Callable[[str], bytes] = _open_local, rng: Optional[random.Random] = None):
        self.args = args
        src = args.training_prompts
        if "txt" in src:
            with open(src, "r") as f:
                self.ann = [line.strip() for line in f]
        elif "jsonl" in src:
            self.ann = read_jsonl(src)
        elif "json" in src:
            self.ann = json.load(open(src, "r"))
        else:
            raise NotImplementedError(src)
        self.opener = opener
        self.rng = rng or random

    def __len__(self):
        return len(self.ann)

    def __getitem__(self, index) -> Dict:
        rec = self.ann[index]
        example = {"text": rec["prompt"]}
        fp = rec["file_path"]
        path = fp if not isinstance(fp, list) else self.rng.choice(fp)               # :60
        with io.BytesIO(self.opener(path)) as f:
            example["latents"] = torch.load(f)
        for k in rec.keys():                                                         # :69-72
            if k not in ("prompt", "file_path", "image"):
                example[k] = rec[k]
        return example


def collate_gan_batch(examples: Iterable[Dict]) -> Dict:
    """default-collate equivalent for the keys the trainer reads: ``text`` list + stacked ``latents`` -> the trainer batch's
    ``real_latents`` (training_script.py:683 ``batch=batch`` -> gan_sdxl.py:46-48)."""
    examples = list(examples)
    lat = torch.stack([e["latents"].float() for e in examples])
    return {"text": [e["text"] for e in examples], "latents": lat, "real_latents": lat}


def main(argv=None) -> int:
    """CLI of tools/gan_gt_generate.py (:45-55): same flags; ``--weights`` picks the random-init stand-ins (this image has no Hub
    access - a deployment builds ``pipeline`` from real modules and calls ``generate_gan_ground_truth`` directly)."""
    import argparse
    p = argparse.ArgumentParser(description="GAN ground-truth latents (50-step cfg-7.5 sampling) -> latents/<uid>.pt + jsonl")
    p.add_argument("--start", type=int, default=0)
    p.add_argument("--end", type=int, default=10000)
    p.add_argument("--unet-path", type=str)
    p.add_argument("--use-cache", action="store_true")
    p.add_argument("--prompt-path", type=str, default="merged_data/abc5k_hrs10k_t2icompall_20k.txt")
    p.add_argument("--save-prompt-path", type=str, default="train_data/gan_train_data.jsonl")
    p.add_argument("--model-path", type=str, default="runwayml/stable-diffusion-v1-5")
    p.add_argument("--model-type", type=str, default="sd_1_5", choices=["sd_1_5", "sdxl", "sdxl_unet"])
    p.add_argument("--batch-size", type=int, default=8)
    p.add_argument("--weights", choices=["synthetic", "synthetic_tiny", "pretrained"], default="synthetic",
                   help="pretrained: --model-path is a local diffusers-layout directory")
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--resolution", type=int, default=512)
    a = p.parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("comat_b200.gan_data needs a GPU: the package has no CPU fallback")
    import time
    from . import synthetic
    from .train import synthetic_components
    name = {"sd_1_5": "sd_1_5", "sdxl": "sdxl", "sdxl_unet": "sdxl"}[a.model_type]
    args = synthetic.default_args(pretrain_model_name=name, gan_loss=False, seed=42)
    dev = torch.device("cuda", torch.cuda.current_device())
    if a.weights == "pretrained":
        from . import pipelines as P
        from .loading import load_unet
        cls = P.TrainableSDXLPipeline if name == "sdxl" else P.TrainableSDPipeline
        unet = load_unet(a.unet_path, dev) if (a.model_type == "sdxl_unet" and a.unet_path) else None      # :75-79
        pipe = cls.from_pretrained(a.model_path, unet=unet, device=dev)
    else:
        pipe = synthetic_components(args, dev, tiny=(a.weights == "synthetic_tiny"), with_caption=False)["pipeline"]
    gen = torch.Generator(device=dev).manual_seed(int(time.time()))                  # :166 (time-seeded in the reference too)
    n = generate_gan_ground_truth(pipe, read_prompts(a.prompt_path, a.start, a.end), a.save_prompt_path, batch_size=a.batch_size,
                                  num_inference_steps=a.steps, guidance_scale=7.5, height=a.resolution, width=a.resolution,
                                  generator=gen, use_cache=a.use_cache)
    print(f"wrote {n} latents, index {a.save_prompt_path}")
    return 0


if __name__ == "__main__":
    import sys
    sys.exit(main())
