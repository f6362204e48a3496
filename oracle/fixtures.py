"""Seeded synthetic inputs shared by the pinning script, the tests and the bench.

ORACLE / TEST INFRASTRUCTURE.  All draws come from CPU ``torch.Generator``s so the same tensors are produced in
the build container and on the GPU box (same image, same torch).
"""
from __future__ import annotations

import random
from typing import List

import torch

from . import sd_modules as sdm
from . import comat_ref as R


def case_key(case: dict) -> str:
    return ",".join(f"{k}={case[k]}" for k in sorted(case))


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def random_prob_maps(g, n, res, T=77, sharp=2.0):
    """(n,res,res,T) rows summing to 1 over T — what a cross-attention softmax stores."""
    return torch.softmax(torch.randn(n, res, res, T, generator=g) * sharp, dim=-1)


def random_mask(g, size=512, empty=False):
    """(1,1,size,size) bool: union of 1-2 axis-aligned rectangles covering 5-40 %."""
    m = torch.zeros(1, 1, size, size, dtype=torch.bool)
    if empty:
        return m
    for _ in range(int(torch.randint(1, 3, (1,), generator=g))):
        h = int(torch.randint(size // 5, size * 6 // 10, (1,), generator=g))
        w = int(torch.randint(size // 5, size * 6 // 10, (1,), generator=g))
        y = int(torch.randint(0, size - h, (1,), generator=g))
        x = int(torch.randint(0, size - w, (1,), generator=g))
        m[..., y:y + h, x:x + w] = True
    return m


def random_words(g, n_words, T=77, max_tok=3, lo=1, hi=40):
    words = []
    perm = (torch.randperm(hi - lo, generator=g) + lo).tolist()
    for _ in range(n_words):
        k = int(torch.randint(1, max_tok + 1, (1,), generator=g))
        words.append([perm.pop() for _ in range(k)])
    return words


# ------------------------------------------------------------------ layer loss (tc_loss_utils.py:66-173)
LAYER_LOSS_CASES = [
    dict(seed=1, res=8, heads=8, n_maps=1, n_words=1, empty=-1, msize=512),
    dict(seed=2, res=16, heads=8, n_maps=3, n_words=3, empty=1, msize=512),
    dict(seed=3, res=32, heads=8, n_maps=3, n_words=2, empty=-1, msize=512),
    dict(seed=4, res=64, heads=8, n_maps=3, n_words=3, empty=-1, msize=512),
    dict(seed=5, res=16, heads=20, n_maps=5, n_words=1, empty=-1, msize=512),   # SDXL-like head count
    dict(seed=6, res=16, heads=8, n_maps=2, n_words=0, empty=-1, msize=512),    # W == 0 -> python zeros
    dict(seed=7, res=8, heads=4, n_maps=2, n_words=2, empty=-1, msize=64),      # tiny-pipeline geometry
]


def layer_loss_inputs(seed, res, heads, n_maps, n_words, empty, msize):
    g = _gen(seed)
    maps = [random_prob_maps(g, heads, res) for _ in range(n_maps)]
    masks = [random_mask(g, msize, empty=(i == empty)) for i in range(n_words)]
    words = random_words(g, n_words)
    return maps, masks, words, res


# ------------------------------------------------------------------ mask loss (gsam_interface.py:140-228)
MASK_LOSS_CASES = [
    dict(seed=11, B=2, heads=8, layers="mid_8,up_16,up_32", n_t=2, msize=512),
    dict(seed=12, B=3, heads=4, layers="up_8,up_16", n_t=1, msize=128),
]
_NOUNS = ["dog", "cat", "table", "sky", "car", "apple", "hat", "dog"]      # 'sky' is in the stop list; 'dog' repeats


def mask_loss_inputs(seed, B, heads, layers, n_t, msize):
    g = _gen(seed)
    layer_ls = layers.split(",")
    n_per = {"mid": 1, "up": 3, "down": 2}
    attn_dict = {}
    for ti in range(n_t):
        d = {}
        for ly in layer_ls:
            place, res = ly.split("_")
            d[ly] = [random_prob_maps(g, B * heads, int(res)) for _ in range(n_per[place])]
        attn_dict[str(951 - 100 * ti)] = d
    subtrees, idx2wp, masks_by_sample = [], [], []
    rr = random.Random(seed)
    for b in range(B):
        n_groups = rr.randint(0 if b == B - 1 else 1, 3)
        pos = list(range(1, 30))
        rr.shuffle(pos)
        groups, wp = [], {}
        for _ in range(n_groups):
            noun = rr.choice(_NOUNS)
            npos = [pos.pop()] if rr.random() < 0.7 else [pos.pop(), pos.pop()]
            if len(npos) == 1:
                wp[npos[0]] = noun
            else:
                wp[npos[0]], wp[npos[1]] = noun[:2], noun[2:]
            mods = []
            for _ in range(rr.randint(0, 2)):
                p = pos.pop()
                wp[p] = "big"
                mods.append(p)
            groups.append(mods + [npos if len(npos) > 1 else npos[0]])
        subtrees.append(groups)
        idx2wp.append(wp)
        _, attrs = R.words_from_subtrees(groups, wp, None)
        # one mask per *surviving* noun is what get_mask returns; generate generously, slice later
        masks_by_sample.append([random_mask(g, msize, empty=(rr.random() < 0.1)) for _ in range(max(1, len(attrs)))])
    # masks must line up with the nouns that survive update_nouns_attributes: recompute with the restated filter
    for b in range(B):
        nouns, attrs = R.words_from_subtrees(subtrees[b], idx2wp[b], update_nouns_attributes)
        masks_by_sample[b] = masks_by_sample[b][: max(1, len(nouns))]
    return attn_dict, subtrees, idx2wp, masks_by_sample, layer_ls, B


_INVALID_NOUNS = set(
    "scene surface area atmosphere noise place kitchen dream interior exterior meal background bathroom room scent "
    "street hillside mountain sky sea ocean lost language skill one night day morning space environment conditions "
    "field shore restroom party grass snow meadow water shadow waves song cycle sunlight mysteries wall salon range "
    "cry speech tone thing about activity air advertisement airport also".split())


def update_nouns_attributes(nouns, attributes):
    """Restated gsam_interface.py:232-261: drop nouns that occur more than once, then stop-listed nouns
    (also when the noun minus its last character — a plural — is stop-listed)."""
    keep = [(n, a) for n, a in zip(nouns, attributes) if nouns.count(n) == 1]
    keep = [(n, a) for n, a in keep if n not in _INVALID_NOUNS and n[:-1] not in _INVALID_NOUNS]
    return [n for n, _ in keep], [a for _, a in keep]


# ------------------------------------------------------------------ pipeline (tiny geometry)
PIPELINE_CASES = [
    dict(seed=21, B=2, S=4, K=2, hw=32, rank=4),
    dict(seed=22, B=1, S=5, K=1, hw=32, rank=4, rescale=0.7),
]


SDXL_PIPELINE_CASES = [
    dict(seed=41, B=2, S=4, K=2, hw=32, rank=4),
    dict(seed=42, B=1, S=3, K=1, hw=32, rank=4, rescale=0.7),
]


def make_tiny_unet(seed, rank=4, sdxl=False, up_std=0.05, width=32, ctx=64):
    torch.manual_seed(seed)
    unet = sdm.UNet2DConditionModel(**sdm.tiny_unet_config(sdxl=sdxl, width=width, cross_attention_dim=ctx))
    unet.requires_grad_(False)
    if rank:
        sdm.install_lora(unet, rank, up_std=up_std, seed=seed + 1)
    return unet


def make_tiny_vae(seed, sdxl=False):
    torch.manual_seed(seed)
    vae = sdm.AutoencoderKL(block_out_channels=(32, 32, 64, 64), scaling_factor=0.13025 if sdxl else 0.18215)
    vae.requires_grad_(False)
    return vae


def pipeline_world(seed, B, S, K, hw, rank, rescale=0.0, sdxl=False):
    g = _gen(seed)
    rr = random.Random(seed)
    steps, attrcon = R.select_training_steps(S, K, rr, 2)
    return dict(
        make_unet=lambda: make_tiny_unet(seed, rank, sdxl),
        vae=make_tiny_vae(seed + 5, sdxl),
        prompt_embeds=torch.randn(B, 77, 64, generator=g),
        null_embeds=torch.randn(1, 77, 64, generator=g).expand(B, -1, -1).contiguous(),
        latents=torch.randn(B, 4, hw, hw, generator=g),
        training_steps=steps, attrcon_steps=attrcon,
        train_layer_ls=["up_8", "up_16", "up_32"],
    )


# ------------------------------------------------------------------ BLIP
BLIP_CASES = [dict(seed=31, B=2, size=254, L=9), dict(seed=32, B=3, size=190, L=14)]


def blip_token_batch(g, B, L_mean, vocab=30524):
    """[CLS] + 'a photography of' stand-in ids + random wordpieces + [SEP], right-padded with 0."""
    rows = []
    for _ in range(B):
        L = int(torch.clamp(torch.randn((), generator=g) * 4 + L_mean, 4, 40).round())
        body = torch.randint(1000, 30000, (L,), generator=g)
        rows.append(torch.cat([torch.tensor([101, 1037, 5855, 1997]), body, torch.tensor([102])]))
    T = max(len(r) for r in rows)
    ids = torch.zeros(B, T, dtype=torch.long)
    mask = torch.zeros(B, T, dtype=torch.long)
    for i, r in enumerate(rows):
        ids[i, : len(r)] = r
        mask[i, : len(r)] = 1
    return ids, mask


def blip_inputs(seed, B, size, L):
    g = _gen(seed)
    model = R.make_blip(large=False, seed=seed)
    images = torch.rand(B, 3, size, size, generator=g)
    ids, mask = blip_token_batch(g, B, L)
    return model, images, ids, mask


# ------------------------------------------------------------------ GAN
GAN_CASES = [dict(seed=41, B=2, S=20, hw=16)]


def gan_world(seed, B, S, hw):
    g = _gen(seed)
    d_unet = make_tiny_unet(seed, rank=4)
    torch.manual_seed(seed + 3)
    head = torch.nn.Sequential(torch.nn.Linear(4, 1))
    return dict(d_unet=d_unet, head=head, fake=torch.randn(B, 4, hw, hw, generator=g),
                real=torch.randn(B, 4, hw, hw, generator=g), null=torch.randn(B, 77, 64, generator=g))


# ---- prompt encoding: no CLIP vocabulary on disk -> a framing-exact stand-in tokenizer
class ClipTokenizerStub:
    """CLIPTokenizer call protocol and framing (BOS 49406 + one id per whitespace word + EOS 49407, padded with ``pad_token_id``);
    word ids are an FNV-1a hash into [1000, 49000).  ORACLE-side twin of comat_b200.synthetic.SyntheticClipTokenizer (written
    independently; tests/test_text_encoder_cpu.py checks both against the ids stored in tests/golden/encode_prompt.pt)."""

    def __init__(self, model_max_length=77, pad_token_id=49407):
        self.model_max_length, self.pad_token_id = model_max_length, pad_token_id

    def __call__(self, text, padding="max_length", max_length=None, truncation=True, return_tensors="pt", **_):
        from types import SimpleNamespace
        texts = [text] if isinstance(text, str) else list(text)
        rows = []
        for t in texts:
            r = [49406]
            for w in t.lower().split():
                h = 0x811C9DC5
                for byte in w.encode("utf-8"):
                    h = ((h ^ byte) * 0x01000193) % (1 << 32)
                r.append(1000 + h % 48000)
            r.append(49407)
            cap = max_length or self.model_max_length
            if truncation and len(r) > cap:
                r = r[:cap - 1] + [49407]
            rows.append(r)
        width = (max_length or self.model_max_length) if padding == "max_length" else max(map(len, rows))
        ids = torch.full((len(rows), width), self.pad_token_id, dtype=torch.long)
        mask = torch.zeros(len(rows), width, dtype=torch.long)
        for i, r in enumerate(rows):
            ids[i, :len(r)] = torch.tensor(r)
            mask[i, :len(r)] = 1
        return SimpleNamespace(input_ids=ids, attention_mask=mask)

    def batch_decode(self, ids):
        return [" ".join(str(int(x)) for x in row) for row in ids]


ENCODE_PROMPT_CASES = [
    dict(prompts=["a red apple on a wooden table", "two dogs"], n_per=1, cfg=True, negative=None, clip_skip=None, seed=7),
    dict(prompts=["the quick brown fox jumps over the lazy dog"], n_per=2, cfg=True, negative="blurry low quality", clip_skip=None, seed=8),
    dict(prompts=["a blue car and a green bench", "a cat", "snow"], n_per=1, cfg=False, negative=None, clip_skip=1, seed=9),
    dict(prompts=[" ".join(["word%d" % i for i in range(90)])], n_per=1, cfg=True, negative=None, clip_skip=None, seed=10),   # truncated at 77
]


# ---- prompt parsing / CLIP alignment: spaCy and the CLIP vocabulary are not in this image -> hand-written parses
class FakeToken:
    """spaCy ``Token`` protocol subset the reference reads: ``.text .pos_ .dep_ .head .children`` (children in sentence order)."""

    def __init__(self, i, text, pos, dep):
        self.i, self.text, self.pos_, self.dep_ = i, text, pos, dep
        self.head, self.children = self, []

    def __repr__(self):
        return f"{self.text}/{self.pos_}/{self.dep_}"


def fake_doc(spec):
    """spec: [(text, pos, dep, head_index), ...] (head_index == own index for the root) -> list of FakeToken."""
    toks = [FakeToken(i, t, p, d) for i, (t, p, d, _) in enumerate(spec)]
    for tok, (_, _, _, h) in zip(toks, spec):
        tok.head = toks[h]
        if h != tok.i:
            toks[h].children.append(tok)
    return toks


class BpeStub:
    """``tokenizer(prompt).input_ids`` + ``convert_ids_to_tokens`` with CLIP's conventions: BOS / EOS strings, ``</w>`` on the last
    piece of a word, configurable multi-piece words."""

    def __init__(self, splits=None):
        self.splits = splits or {}
        self.vocab = ["<|startoftext|>", "<|endoftext|>"]

    def _id(self, piece):
        if piece not in self.vocab:
            self.vocab.append(piece)
        return self.vocab.index(piece)

    def __call__(self, prompt, **_):
        from types import SimpleNamespace
        ids = [0]
        for w in prompt.lower().split():
            pieces = self.splits.get(w, [w])
            ids += [self._id(p + ("</w>" if j == len(pieces) - 1 else "")) for j, p in enumerate(pieces)]
        return SimpleNamespace(input_ids=ids + [1])

    def convert_ids_to_tokens(self, ids):
        return [self.vocab[i] for i in ids]


ATTR_ALIGN_CASES = {
    "a red apple and a blue car": [
        ("a", "DET", "det", 2), ("red", "ADJ", "amod", 2), ("apple", "NOUN", "ROOT", 2), ("and", "CCONJ", "cc", 2),
        ("a", "DET", "det", 6), ("blue", "ADJ", "amod", 6), ("car", "NOUN", "conj", 2)],
    "a dog that is red": [
        ("a", "DET", "det", 1), ("dog", "NOUN", "ROOT", 1), ("that", "PRON", "nsubj", 3), ("is", "AUX", "relcl", 1),
        ("red", "ADJ", "acomp", 3)],
    "the strawberry cake is pink and fluffy": [
        ("the", "DET", "det", 2), ("strawberry", "NOUN", "compound", 2), ("cake", "NOUN", "nsubj", 3), ("is", "AUX", "ROOT", 3),
        ("pink", "ADJ", "acomp", 3), ("and", "CCONJ", "cc", 4), ("fluffy", "ADJ", "conj", 4)],
    "a red red red bear": [
        ("a", "DET", "det", 4), ("red", "ADJ", "amod", 4), ("red", "ADJ", "amod", 4), ("red", "ADJ", "amod", 4),
        ("bear", "NOUN", "ROOT", 4)],
    "two cats and two dogs": [
        ("two", "NUM", "nummod", 1), ("cats", "NOUN", "ROOT", 1), ("and", "CCONJ", "cc", 1), ("two", "NUM", "nummod", 4),
        ("dogs", "NOUN", "conj", 1)],
    "a big old red wooden table": [
        ("a", "DET", "det", 5), ("big", "ADJ", "amod", 5), ("old", "ADJ", "amod", 5), ("red", "ADJ", "amod", 5),
        ("wooden", "ADJ", "amod", 5), ("table", "NOUN", "ROOT", 5)],
    "a red bear and a red car near a skateboard": [
        ("a", "DET", "det", 2), ("red", "ADJ", "amod", 2), ("bear", "NOUN", "ROOT", 2), ("and", "CCONJ", "cc", 2),
        ("a", "DET", "det", 6), ("red", "ADJ", "amod", 6), ("car", "NOUN", "conj", 2), ("near", "ADP", "prep", 6),
        ("a", "DET", "det", 9), ("skateboard", "NOUN", "pobj", 7)],
    "a dark blue metal bench that looks very old": [
        ("a", "DET", "det", 4), ("dark", "ADJ", "amod", 2), ("blue", "ADJ", "amod", 4), ("metal", "NOUN", "compound", 4),
        ("bench", "NOUN", "ROOT", 4), ("that", "PRON", "nsubj", 6), ("looks", "VERB", "relcl", 4), ("very", "ADV", "advmod", 8),
        ("old", "ADJ", "acomp", 6)],
    "a strawberry cake on a skateboard": [
        ("a", "DET", "det", 2), ("strawberry", "NOUN", "compound", 2), ("cake", "NOUN", "ROOT", 2), ("on", "ADP", "prep", 2),
        ("a", "DET", "det", 5), ("skateboard", "NOUN", "pobj", 3)],
    "the wooden skateboard is fluffy": [
        ("the", "DET", "det", 2), ("wooden", "ADJ", "amod", 2), ("skateboard", "NOUN", "nsubj", 3), ("is", "AUX", "ROOT", 3),
        ("fluffy", "ADJ", "acomp", 3)],
    "a fluffy cat and a fluffy dog": [
        ("a", "DET", "det", 2), ("fluffy", "ADJ", "amod", 2), ("cat", "NOUN", "ROOT", 2), ("and", "CCONJ", "cc", 2),
        ("a", "DET", "det", 6), ("fluffy", "ADJ", "amod", 6), ("dog", "NOUN", "conj", 2)],
}
ATTR_ALIGN_SPLITS = {"strawberry": ["straw", "berry"], "skateboard": ["skate", "board"], "fluffy": ["flu", "ffy"]}
