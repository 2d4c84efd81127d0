"""Generate tests/golden/*.npz from the UNMODIFIED reference (`MyModel` in /root/reference).

Run in the build container only (needs /root/reference and `transformers`):
    python tests/golden/make_golden.py

What it does, per case:
  1. builds random-init HF checkpoints in a temp dir under the names the reference's
     `from_pretrained` calls expect (/root/reference/models/model.py:14-17), chdir's there;
  2. instantiates the reference `MyModel(args)` and loads the oracle's numpy-seeded weights
     (oracle.seeded_state_dicts) into its three sub-models, strict=True;
  3. runs `model(images, source_encoding, target_encoding)` + `.backward()` in eval mode (dropout
     off) and `model(..., return_loss=False)` (greedy decode);
  4. stores loss, every parameter-gradient's L2 norm and 8 sampled elements, a slice of the
     concatenated embeddings, and the generated ids.
The weights / inputs are NOT stored: tests regenerate them from the same seeds.
"""
import argparse
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.caption_model import seeded_inputs, seeded_state_dicts  # noqa: E402
from oracle.swinv2 import SwinDims  # noqa: E402
from oracle.t5 import T5Dims  # noqa: E402

CASES = {
    # 3 Swin stages: 16x16 / 8x8 (shifted windows + mask) and 4x4 (grid == window -> shift forced to 0)
    "tiny_a": dict(swin=dict(image_size=64, embed_dim=32, depths=(2, 2, 2), num_heads=(1, 2, 4), window_size=4),
                   t5=dict(vocab_size=512, d_model=128, d_ff=256, num_layers=2, num_heads=2),
                   batch=2, l_src=9, l_tgt=12, ignore_tail=False, train_swin=True),
    # non-power-of-two window (N=49 like Swin-T/224), ragged lengths, ignore-index labels, frozen Swin
    "tiny_b": dict(swin=dict(image_size=56, embed_dim=32, depths=(2, 2), num_heads=(1, 2), window_size=7,
                             pretrained_window_sizes=(0, 0)),
                   t5=dict(vocab_size=300, d_model=64, d_ff=160, num_layers=3, num_heads=1, num_decoder_layers=2),
                   batch=3, l_src=5, l_tgt=7, ignore_tail=True, train_swin=False),
}


# Closer to the real geometries; the oracle is pinned to the reference on them too (tests/test_oracle_golden.py).  The GPU
# step-parity tests compare these cases with the oracle only (tests/test_step_parity_gpu.py: EXTRA_CASES).
EXTRA_CASES = {
    # window 8 with shift 4, head_dim 32, d_kv 64, 96-token encoder sequence
    "mid": dict(swin=dict(image_size=128, embed_dim=32, depths=(2, 2, 2), num_heads=(1, 2, 4), window_size=8),
                t5=dict(vocab_size=1000, d_model=128, d_ff=512, num_layers=2, num_heads=2),
                batch=2, l_src=16, l_tgt=24, ignore_tail=True, train_swin=True),
    # BASELINE configs 3 / 4 in miniature: 12 x 12 windows with shift 6 (N = 144: the CUDA-core window-attention kernel), a
    # 64-token source (encoder length 144 + 64 = 208: the multi-tile T5 attention kernels) and 128-token targets
    "hires": dict(swin=dict(image_size=96, embed_dim=32, depths=(2, 2), num_heads=(1, 2), window_size=12,
                            pretrained_window_sizes=(0, 0)),
                  t5=dict(vocab_size=600, d_model=64, d_ff=128, num_layers=2, num_heads=1),
                  batch=2, l_src=64, l_tgt=128, ignore_tail=True, train_swin=True),
}


def dims_of(case):
    sw = dict(case["swin"])
    sw.setdefault("pretrained_window_sizes", (0,) * len(sw["depths"]))
    swin = SwinDims(**sw)
    t5 = T5Dims(**case["t5"])
    return swin, t5


def sample_index(numel: int, k: int = 8) -> np.ndarray:
    return (np.arange(k, dtype=np.int64) * 2654435761 % max(numel, 1)).astype(np.int64)


def run_reference(name: str, case: dict) -> dict:
    from transformers import Swinv2Config, Swinv2Model, T5Config, T5EncoderModel, T5ForConditionalGeneration
    swin, t5 = dims_of(case)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        os.environ["HF_HUB_OFFLINE"] = "1"
        cfg = T5Config(vocab_size=t5.vocab_size, d_model=t5.d_model, d_kv=t5.d_kv, d_ff=t5.d_ff,
                       num_layers=t5.num_layers, num_decoder_layers=t5.n_dec, num_heads=t5.num_heads,
                       decoder_start_token_id=0)
        T5ForConditionalGeneration(cfg).save_pretrained("t5-small")
        scfg = Swinv2Config(image_size=swin.image_size, patch_size=swin.patch_size, embed_dim=swin.embed_dim,
                            depths=list(swin.depths), num_heads=list(swin.num_heads), window_size=swin.window_size,
                            pretrained_window_sizes=list(swin.pretrained_window_sizes))
        Swinv2Model(scfg).save_pretrained("swin-local")
        sys.path.insert(0, "/root/reference")
        from models.model import MyModel  # the reference, unmodified
        args = types.SimpleNamespace(result_dir=tmp, language_model_name="t5-small", image_model_name="swin-local",
                                     image_model_train=case["train_swin"], transformer_model_name="t5-small")
        model = MyModel(args)
        os.chdir(cwd)
    sds = seeded_state_dicts(t5, swin, t5, seed=0)
    model.language_model.load_state_dict(sds["language_model"], strict=True)
    model.image_model.load_state_dict(sds["image_model"], strict=True)
    model.transformer.load_state_dict(sds["transformer"], strict=True)
    model.eval()
    px, src, tgt = seeded_inputs(case["batch"], swin, t5.vocab_size, case["l_src"], case["l_tgt"],
                                 ignore_tail=case["ignore_tail"])
    loss = model({"pixel_values": px}, {"input_ids": src}, {"input_ids": tgt})
    loss.backward()
    out = {"loss": np.float64(loss.item())}
    for scope, mod in (("transformer", model.transformer), ("image_model", model.image_model)):
        for k, p in mod.named_parameters():
            if p.grad is None:
                continue
            g = p.grad.detach().double().flatten()
            out[f"gnorm/{scope}/{k}"] = np.float64(g.norm().item())
            out[f"gsamp/{scope}/{k}"] = g[torch.from_numpy(sample_index(g.numel()))].numpy()
    with torch.no_grad():
        lang = model.language_model(src).last_hidden_state
        img = model.image_model(pixel_values=px).last_hidden_state
        out["img_emb"] = img[:, :, :8].numpy()
        out["lang_emb"] = lang[:, :, :8].numpy()
        ids = model({"pixel_values": px}, {"input_ids": src}, return_loss=False)
    out["generated"] = ids.numpy()
    print(name, "loss", out["loss"], "generated", tuple(ids.shape), "n_grads",
          sum(k.startswith("gnorm/") for k in out))
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.dirname(os.path.abspath(__file__)))
    a = ap.parse_args()
    torch.manual_seed(0)
    for name, case in {**CASES, **EXTRA_CASES}.items():
        np.savez_compressed(os.path.join(a.out, f"{name}.npz"), **run_reference(name, case))
