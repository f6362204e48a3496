"""CPU: prompt -> (modifier ..., noun) token groups against the reference's own ``attribute_concen_utils`` and
``AttrConcenTrainableSDPipeline._extract_attribution_indices`` run verbatim (tests/golden/attr_align.json, written by
oracle/pin_against_reference.py) on hand-written dependency parses and a CLIP-convention word-piece stub."""
import json
import os

import pytest

from oracle import fixtures as FX
from tests.conftest import GOLDEN

GOLD = json.load(open(os.path.join(GOLDEN, "attr_align.json")))


def _idx(groups):
    return None if groups is None else [[t.i for t in g] for g in groups]


@pytest.mark.parametrize("prompt", list(FX.ATTR_ALIGN_CASES))
def test_extractors_and_alignment_match_reference(prompt):
    from comat_b200 import attr_align as AA
    gold = GOLD[prompt]
    doc = FX.fake_doc(FX.ATTR_ALIGN_CASES[prompt])
    tok = FX.BpeStub(FX.ATTR_ALIGN_SPLITS)
    assert _idx(AA.extract_attribution_indices(doc)) == gold["plain"]
    assert _idx(AA.extract_attribution_indices_with_verbs(doc)) == gold["with_verbs"]
    assert _idx(AA.extract_attribution_indices_with_verb_root(doc)) == gold["verb_root"]
    assert AA.extract_attribution_indices_for_prompt(doc, tok, prompt) == gold["aligned"]
    assert {str(k): v for k, v in AA.get_attention_map_index_to_wordpiece(tok, prompt).items()} == gold["idx_to_wp"]


def test_golden_covers_the_interesting_branches():
    assert set(GOLD) == set(FX.ATTR_ALIGN_CASES)
    assert GOLD["two cats and two dogs"]["with_verbs"] == [] and GOLD["two cats and two dogs"]["aligned"] == []
    assert any(isinstance(e, list) for g in GOLD["a strawberry cake on a skateboard"]["aligned"] for e in g)     # split word
    assert GOLD["a big old red wooden table"]["plain"] and GOLD["a big old red wooden table"]["aligned"] == []   # >= 4 tokens dropped
    assert GOLD["a fluffy cat and a fluffy dog"]["aligned"] == [[[2, 3], 4], [[7, 8], 9]]                        # repeats -> next occurrence


def test_words_for_prompts_feeds_the_mask_loss_inputs():
    """parser + tokenizer -> what CoMatTrainer's batch carries as ``words`` (gsam_interface.py:163-196 via words_from_subtrees)."""
    from comat_b200 import attr_align as AA
    tok = FX.BpeStub(FX.ATTR_ALIGN_SPLITS)
    parser = lambda p: FX.fake_doc(FX.ATTR_ALIGN_CASES[p])
    nouns, words = AA.words_for_prompts(parser, tok, ["a red apple and a blue car", "a fluffy cat and a fluffy dog", "two cats and two dogs"])
    assert nouns == [["apple", "car"], ["cat", "dog"], []]
    assert words == [[[2, 3], [6, 7]], [[2, 3, 4], [7, 8, 9]], []]


def test_unify_lists_drops_repeats_and_subsets():
    from comat_b200.attr_align import unify_lists
    a, b, c, d = "a", "b", "c", "d"
    assert unify_lists([[a, b]], [[a, b, c]], [[a, b]]) == [[a, b, c]]
    assert unify_lists([[a, b], [c, d]], [], [[c, d]]) == [[a, b], [c, d]]
    assert unify_lists([], [], []) == []
