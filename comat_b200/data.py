"""Prompt data for the training loop: ``get_dataset_dataloader`` (training_utils/dataset.py:10-57) without the HF ``datasets`` /
image-folder machinery the CoMat scripts never use (prompts only; ``Gan_Dataset`` when the GAN loss is on), and the
data-parallel sharding of SURVEY 8e: every epoch ONE seeded permutation shared by all ranks, rank r takes positions
``i = r (mod world)`` - no prompt is seen twice in an epoch.  (The reference builds a shuffling DataLoader, dataset.py:39,47-52, and
passes it through ``accelerator.prepare`` (training_script.py:324-330), which shards it across ranks as well.)"""
from __future__ import annotations

import json
import random
from typing import Dict, Iterator, List, Optional

import torch

from .gan_data import Gan_Dataset, collate_gan_batch


class PromptDataset(torch.utils.data.Dataset):
    """``load_dataset('text' | 'json', data_files=args.training_prompts)['train']`` (dataset.py:13-16): items ``{'text': prompt}``."""

    def __init__(self, path: str, max_train_samples: Optional[int] = None):
        if path.endswith("txt"):
            with open(path, "r") as f:
                rows = [{"text": line.rstrip("\n")} for line in f if line.strip()]
        elif path.endswith("json"):
            rows = json.load(open(path, "r"))
            rows = [r if isinstance(r, dict) else {"text": r} for r in rows]
        else:
            raise NotImplementedError(f"training_prompts must be .txt or .json (dataset.py:13-16): {path}")
        self.rows = rows[:max_train_samples] if max_train_samples is not None else rows

    def __len__(self):
        return len(self.rows)

    def __getitem__(self, i) -> Dict:
        return self.rows[i]


def get_dataset(args):
    """dataset.py:11-16."""
    if args.gan_loss:
        return Gan_Dataset(args)
    return PromptDataset(args.training_prompts, getattr(args, "max_train_samples", None))


class ShardedBatches:
    """One rank's batches of one epoch.  ``len()`` is the same on every rank (the tail that does not fill one batch on EVERY
    rank is dropped, so no rank waits in the gradient all-reduce for a peer that ran out of data)."""

    def __init__(self, dataset, batch_size: int, rank: int = 0, world: int = 1, seed: int = 0, shuffle: bool = True):
        self.dataset, self.batch_size, self.rank, self.world, self.seed, self.shuffle = dataset, batch_size, rank, world, seed, shuffle
        self.epoch = 0

    def set_epoch(self, epoch: int):
        self.epoch = epoch

    def __len__(self):
        return len(self.dataset) // (self.batch_size * self.world)

    def indices(self) -> List[int]:
        order = list(range(len(self.dataset)))
        if self.shuffle:
            random.Random(self.seed * 1000003 + self.epoch).shuffle(order)
        usable = len(self) * self.batch_size * self.world
        return order[:usable][self.rank::self.world]

    def __iter__(self) -> Iterator[Dict]:
        idx = self.indices()
        for b in range(len(self)):
            items = [self.dataset[i] for i in idx[b * self.batch_size:(b + 1) * self.batch_size]]
            if "latents" in items[0]:
                batch = collate_gan_batch(items)
                for k in items[0]:
                    if k not in batch:
                        batch[k] = [it[k] for it in items]
            else:
                batch = {k: [it[k] for it in items] for k in items[0]}
            yield batch
