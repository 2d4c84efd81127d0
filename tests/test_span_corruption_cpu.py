"""N3: span-corruption target construction against golden vectors produced by the unmodified reference loader
(tests/golden/make_span_golden.py), plus size-independent properties."""
import json
import os

import torch

from klab_multimodalmodel_b200.data import restore, span_corrupt, split_words

GOLD = os.path.join(os.path.dirname(__file__), "golden", "span_corruption.json")


def test_matches_reference_golden_vectors():
    cases = json.load(open(GOLD))
    assert len(cases) >= 20
    for c in cases:
        torch.manual_seed(c["seed"])
        src, tgt = span_corrupt(c["caption"])
        assert (src, tgt) == (c["source"], c["target"]), c
        g = torch.Generator().manual_seed(c["seed"])               # an explicit generator draws the same permutation
        assert span_corrupt(c["caption"], g) == (c["source"], c["target"])


def test_properties_hold_for_long_and_degenerate_captions():
    g = torch.Generator().manual_seed(3)
    words = [f"w{i}" for i in range(400)]
    cap = " ".join(words) + " !"
    src, tgt = span_corrupt(cap, g)
    n = len(split_words(cap))
    k = int(n * 0.15) + 1
    assert src.count("<extra_id_") == k and tgt.count("<extra_id_") == k + 1
    ids = [int(t[len("<extra_id_"):-1]) for t in src.split() if t.startswith("<extra_id_")]
    assert ids == list(range(k)), "sentinels are numbered in order of position"
    assert restore(src, tgt) == " ".join(split_words(cap))
    # a single word is always masked; an empty caption yields the bare opening sentinel
    assert span_corrupt("cat", torch.Generator().manual_seed(0)) == ("<extra_id_0>", "<extra_id_0> cat <extra_id_1>")
    assert span_corrupt("", torch.Generator().manual_seed(0)) == ("", "<extra_id_0>")
