"""Prompt -> (modifier ..., noun) token-index groups: the host-side producer of ``get_mask_loss``'s ``all_subtree_indices`` and
``attn_map_idx_to_wp_all`` inputs (training_script.py:628-629).

Mirrors ``attribute_concen_utils.py`` (dependency-subtree extraction over a spaCy ``Doc``, CLIP word-piece alignment) and
``AttrConcenTrainableSDPipeline._extract_attribution_indices / _align_indices / unify_lists`` (:281-338, :539-564).  The parser itself
(spaCy ``en_core_web_trf``) is an external model: any object with spaCy's token protocol (``.text .pos_ .dep_ .children``) works,
and the tokenizer needs ``tokenizer(prompt).input_ids`` + ``convert_ids_to_tokens``.  Pure host string / tree logic - it runs on the
CPU beside the step (one prompt batch ahead), never on the GPU path.

The three extractors of the reference differ only in which heads they start from, which dependents they follow and which
nodes they keep, so they are one tree walk here, parameterised by predicates.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence

START_TOKEN, END_TOKEN = "<|startoftext|>", "<|endoftext|>"
MODIFIERS = ("amod", "nmod", "compound", "npadvmod", "advmod", "acomp")
_NOUNS, _VERBAL = ("NOUN", "PROPN"), ("AUX", "VERB")


def _walk(head, follow_first: Callable, keep_first: Callable):
    """first level: the direct children of ``head`` left to right -> (kept nodes, stack of their children to visit)."""
    group, stack = [], []
    for child in head.children:
        if follow_first(child):
            if keep_first(child):
                group.append(child)
            stack.extend(child.children)
    return group, stack


def _drain(group: list, stack: list, follow: Callable, keep: Callable) -> list:
    """deeper levels, in the reference's visiting order (LIFO over the pending descendants)."""
    while stack:
        node = stack.pop()
        if follow(node):
            if keep(node):
                group.append(node)
            stack.extend(node.children)
    return group


def extract_attribution_indices(doc) -> list:
    """attribute_concen_utils.py:39-62 - nouns with adjectival / compound modifiers ("a red apple", "strawberry cake")."""
    out = []
    for w in doc:
        if w.pos_ not in _NOUNS or w.dep_ in MODIFIERS:
            continue
        is_mod = lambda n: n.dep_ in MODIFIERS
        group, stack = _walk(w, is_mod, lambda n: True)
        group = _drain(group, stack, lambda n: n.dep_ in MODIFIERS or n.dep_ == "conj", lambda n: True)
        if group:
            out.append(group + [w])
    return out


def extract_attribution_indices_with_verbs(doc) -> Optional[list]:
    """attribute_concen_utils.py:64-93 - a verb between noun and modifier ("a dog that is red"): relative clauses are followed,
    verbs / auxiliaries themselves are not kept.  Reference quirk kept: the function returns after the FIRST candidate noun
    (its ``return`` sits inside the loop), and returns None when the prompt has none."""
    mods = MODIFIERS + ("relcl",)
    not_verbal = lambda n: n.pos_ not in _VERBAL
    for w in doc:
        if w.pos_ not in _NOUNS or w.dep_ in mods:
            continue
        group, stack = _walk(w, lambda n: n.dep_ in mods, not_verbal)
        group = _drain(group, stack, lambda n: n.dep_ in mods or n.dep_ == "conj", not_verbal)
        return [group + [w]] if group else []
    return None


def extract_attribution_indices_with_verb_root(doc) -> list:
    """attribute_concen_utils.py:95-131 - copular sentences ("the cake is pink and fluffy"): start from an auxiliary that has both a
    noun child and a modifier child; the auxiliary itself is never part of the group."""
    out = []
    for w in doc:
        if w.pos_ != "AUX" or w.dep_ in MODIFIERS:
            continue
        group, stack = _walk(w, lambda n: n.dep_ in MODIFIERS or n.pos_ in _NOUNS, lambda n: n.pos_ not in _VERBAL)
        if len(group) < 2:
            continue
        group = _drain(group, stack, lambda n: n.dep_ in MODIFIERS or n.dep_ == "conj", lambda n: n.pos_ != "AUX")
        out.append(group)
    return out


def unify_lists(*lists: Sequence[list]) -> list:
    """AttrConcenTrainableSDPipeline.py:539-564: concatenate, order by length (stable), drop exact repeats and every group that is
    a strict subset of a later (longer or equal-position) group."""
    ordered = sorted([g for l in lists for g in l], key=len)
    seen, result = set(), []
    for i, g in enumerate(ordered):
        if tuple(g) in seen:
            continue
        if any(len(g) < len(h) and all(x in h for x in g) for h in ordered[i + 1:]):
            continue
        result.append(g)
        seen.add(tuple(g))
    return result


def get_indices(tokenizer, prompt: str) -> Dict[int, str]:
    """attribute_concen_utils.py:134-143: position -> word-piece string of the tokenised prompt (BOS / EOS included)."""
    ids = tokenizer(prompt).input_ids
    return dict(enumerate(tokenizer.convert_ids_to_tokens(ids)))


def get_attention_map_index_to_wordpiece(tokenizer, prompt: str) -> Dict[int, str]:
    """attribute_concen_utils.py:145-155: the same without BOS / EOS and without the ``</w>`` end-of-word marks."""
    wp = get_indices(tokenizer, prompt)
    return {i: wp[i].replace("</w>", "") for i in list(wp.keys())[1:-1]}


def align_wordpieces_indices(wordpieces2indices: Dict[int, str], start_idx: int, target_word: str) -> List[int]:
    """attribute_concen_utils.py:11-36: positions of the consecutive word pieces that spell ``target_word`` starting at
    ``start_idx``; [] when the pieces after ``start_idx`` stop matching before the word is complete."""
    got = [start_idx]
    built = wordpieces2indices[start_idx].replace("</w>", "")
    for j in range(start_idx + 1, len(wordpieces2indices)):
        if built == target_word:
            break
        nxt = wordpieces2indices[j].replace("</w>", "")
        if target_word.startswith(built + nxt) and nxt != target_word:
            built += nxt
            got.append(j)
        else:
            return []
    return got


def align_indices(tokenizer, prompt: str, spacy_pairs: Iterable[Sequence]) -> list:
    """AttrConcenTrainableSDPipeline.py:296-338: every parser token of every group -> its CLIP position (an int) or positions
    (a list, for words split into several pieces); a position is handed out once across the whole prompt, so repeated words
    ("a red bear and a red car") land on successive occurrences."""
    wp = get_indices(tokenizer, prompt)
    taken, paired = set(), []
    for pair in spacy_pairs:
        current = []
        for member in pair:
            for idx, piece in wp.items():
                if piece in (START_TOKEN, END_TOKEN):
                    continue
                piece = piece.replace("</w>", "")
                if member.text == piece:
                    if idx not in current and idx not in taken:
                        current.append(idx)
                        break
                elif member.text.startswith(piece) and piece != member.text:
                    span = align_wordpieces_indices(wp, idx, member.text)
                    if span and span not in current and all(j not in taken for j in span):
                        current.append(span)
                        break
        for c in current:
            taken.update(c if isinstance(c, list) else [c])
        paired.append(current)
    return paired


def extract_attribution_indices_for_prompt(doc, tokenizer, prompt: str, max_group: int = 4) -> list:
    """AttrConcenTrainableSDPipeline.py:281-294: the three extractors unified, groups of ``max_group`` or more tokens dropped,
    aligned to CLIP positions.  ``doc`` = ``parser(prompt)``."""
    pairs = unify_lists(extract_attribution_indices(doc) or [], extract_attribution_indices_with_verb_root(doc) or [],
                        extract_attribution_indices_with_verbs(doc) or [])
    pairs = [p for p in pairs if len(p) < max_group]
    return align_indices(tokenizer, prompt, pairs)


def words_for_prompts(parser: Callable, tokenizer, prompts: Sequence[str]):
    """what the trainer needs per batch (training_script.py:628-629 + gsam_interface.py:163-196): per prompt, the noun strings
    (to be segmented by the mask model) and the token-position list of each noun with its attributes."""
    from .attn_loss import words_from_subtrees
    nouns_all, words_all = [], []
    for p in prompts:
        groups = extract_attribution_indices_for_prompt(parser(p), tokenizer, p)
        nouns, words = words_from_subtrees(groups, get_attention_map_index_to_wordpiece(tokenizer, p))
        nouns_all.append(nouns)
        words_all.append(words)
    return nouns_all, words_all
