"""The oracle against the real `transformers` classes the reference instantiates (/root/reference/models/model.py:14-17),
module by module: Swinv2Model, T5EncoderModel, T5ForConditionalGeneration (loss, gradients, greedy generate).  CPU only;
skipped where `transformers` does not import (it is the un-vendored dependency that holds the reference's arithmetic,
SURVEY.md 8c).  Together with tests/test_oracle_golden.py (outputs of the unmodified reference MyModel) this pins the oracle."""
import pytest
import torch

transformers = pytest.importorskip("transformers")

from oracle.caption_model import seeded_inputs, seeded_state_dicts  # noqa: E402
from oracle.swinv2 import swinv2_forward  # noqa: E402
from oracle.t5 import t5_greedy_decode, t5_lm_loss, t5_stack  # noqa: E402
from tests.golden.make_golden import CASES, EXTRA_CASES, dims_of  # noqa: E402


def _hf_models(swin, t5):
    from transformers import Swinv2Config, Swinv2Model, T5Config, T5EncoderModel, T5ForConditionalGeneration
    def cfg():                   # one config object per model: T5EncoderModel edits the one it is given
        return T5Config(vocab_size=t5.vocab_size, d_model=t5.d_model, d_kv=t5.d_kv, d_ff=t5.d_ff, num_layers=t5.num_layers,
                        num_decoder_layers=t5.n_dec, num_heads=t5.num_heads, decoder_start_token_id=0)
    scfg = Swinv2Config(image_size=swin.image_size, patch_size=swin.patch_size, embed_dim=swin.embed_dim, depths=list(swin.depths),
                        num_heads=list(swin.num_heads), window_size=swin.window_size,
                        pretrained_window_sizes=list(swin.pretrained_window_sizes))
    return T5EncoderModel(cfg()).eval(), Swinv2Model(scfg).eval(), T5ForConditionalGeneration(cfg()).eval()


@pytest.mark.parametrize("name", ["tiny_a", "tiny_b", "mid"])
def test_oracle_modules_match_transformers(name):
    case = CASES.get(name) or EXTRA_CASES[name]
    swin, t5 = dims_of(case)
    sds = seeded_state_dicts(t5, swin, t5, seed=3)
    lm, im, tr = _hf_models(swin, t5)
    lm.load_state_dict(sds["language_model"], strict=True)
    im.load_state_dict(sds["image_model"], strict=True)
    tr.load_state_dict(sds["transformer"], strict=True)
    px, src, tgt = seeded_inputs(case["batch"], swin, t5.vocab_size, case["l_src"], case["l_tgt"], ignore_tail=case["ignore_tail"], seed=77)
    with torch.no_grad():
        # Swinv2Model.forward (HF/models/swinv2/modeling_swinv2.py:933-987)
        ref_img = im(pixel_values=px).last_hidden_state
        got_img = swinv2_forward(px, sds["image_model"], swin)
        torch.testing.assert_close(got_img, ref_img, rtol=1e-4, atol=1e-5)
        # T5EncoderModel.forward (HF/models/t5/modeling_t5.py:1164-1208)
        ref_lang = lm(input_ids=src).last_hidden_state
        got_lang = t5_stack(torch.nn.functional.embedding(src, sds["language_model"]["shared.weight"]), sds["language_model"], "encoder.", t5)
        torch.testing.assert_close(got_lang, ref_lang, rtol=1e-4, atol=1e-5)
    # T5ForConditionalGeneration(inputs_embeds, labels).loss + gradients (HF/models/t5/modeling_t5.py:992-1133)
    emb = torch.cat((ref_img, ref_lang), dim=1)
    ref_loss = tr(inputs_embeds=emb, labels=tgt).loss
    ref_loss.backward()
    osd, uniq = dict(sds["transformer"]), {}
    for k, v in osd.items():
        if id(v) not in uniq:
            uniq[id(v)] = v.clone().requires_grad_(True)
        osd[k] = uniq[id(v)]
    got_loss = t5_lm_loss(emb, tgt, osd, t5)
    got_loss.backward()
    assert abs(got_loss.item() - ref_loss.item()) <= 2e-5 * abs(ref_loss.item())
    checked = 0
    for k, p in tr.named_parameters():
        g, r = osd[k].grad, p.grad
        assert (g - r).norm().item() <= 1e-4 * r.norm().item() + 1e-9, k
        checked += 1
    assert checked >= 20
    # greedy generate (HF/generation/utils.py:2658-2800)
    with torch.no_grad():
        ref_ids = tr.generate(inputs_embeds=emb)                    # the reference's call, all defaults (models/model.py:28)
        got_ids = t5_greedy_decode(emb, sds["transformer"], t5, 20)
    n = got_ids.shape[1]
    assert torch.equal(got_ids, ref_ids[:, :n]) and (ref_ids[:, n:] == 0).all()
