"""The oracle vs outputs of the UNMODIFIED reference MyModel (tests/golden/*.npz, made by
tests/golden/make_golden.py in the build container).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle.caption_model import caption_embeddings, caption_generate, caption_loss, seeded_inputs, seeded_state_dicts
from tests.golden.make_golden import CASES, EXTRA_CASES, dims_of, sample_index

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", sorted(CASES) + sorted(EXTRA_CASES))
def test_oracle_matches_reference_golden(name):
    case = CASES.get(name) or EXTRA_CASES[name]
    gold = np.load(os.path.join(GOLD, f"{name}.npz"))
    swin, t5 = dims_of(case)
    sds = seeded_state_dicts(t5, swin, t5, seed=0)
    leaves = {}
    for scope in ("transformer", "image_model"):
        train = scope == "transformer" or case["train_swin"]
        uniq = {}
        for k, v in sds[scope].items():                      # tied tensors share one leaf
            if id(v) not in uniq:
                uniq[id(v)] = v.clone().requires_grad_(train)
            sds[scope][k] = uniq[id(v)]
            leaves[(scope, k)] = uniq[id(v)]
    px, src, tgt = seeded_inputs(case["batch"], swin, t5.vocab_size, case["l_src"], case["l_tgt"],
                                 ignore_tail=case["ignore_tail"])
    loss = caption_loss(px, src, tgt, sds, t5, swin, t5)
    assert abs(loss.item() - float(gold["loss"])) <= 2e-5 * abs(float(gold["loss"]))
    loss.backward()
    checked = 0
    for key in gold.files:
        if not key.startswith("gnorm/"):
            continue
        _, scope, pname = key.split("/", 2)
        g = leaves[(scope, pname)].grad
        assert g is not None, key
        g = g.double().flatten()
        ref_norm = float(gold[key])
        assert abs(g.norm().item() - ref_norm) <= 1e-4 * ref_norm + 1e-9, key
        samp = g[torch.from_numpy(sample_index(g.numel()))].numpy()
        np.testing.assert_allclose(samp, gold[f"gsamp/{scope}/{pname}"], rtol=2e-3, atol=1e-5 * ref_norm + 1e-9)
        checked += 1
    assert checked >= 50
    with torch.no_grad():
        emb = caption_embeddings(px, src, sds, t5, swin)
    n_img = swin.out_tokens
    np.testing.assert_allclose(emb[:, :n_img, :8].numpy(), gold["img_emb"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(emb[:, n_img:, :8].numpy(), gold["lang_emb"], rtol=1e-4, atol=1e-5)
    ids = caption_generate(px, src, sds, t5, swin, t5)
    gen = gold["generated"]
    assert ids.shape[1] <= gen.shape[1]
    np.testing.assert_array_equal(ids.numpy(), gen[:, :ids.shape[1]])
    assert (gen[:, ids.shape[1]:] == 0).all()
