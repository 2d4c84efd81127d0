"""BASELINE config 5 in numbers: greedy caption decode, Swin-S/256 + T5-base, batch 256, 20 new tokens with EOS disabled
(fixed length), KV-cached single-token steps vs the prefix-recompute schedule.  CUDA-event timings after a warm-up run.
python scripts/decode_probe.py [--batch 256]"""
import argparse
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from klab_multimodalmodel_b200.generation import greedy_generate, greedy_generate_recompute
from klab_multimodalmodel_b200.modeling import Swinv2Config, T5Config, init_swin_, init_t5_
from klab_multimodalmodel_b200.models.model import MyModel

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
a = ap.parse_args()
dev = torch.device("cuda", 0)
tcfg = T5Config(d_model=768, d_ff=3072, num_layers=12, num_heads=12)
scfg = Swinv2Config(image_size=256, embed_dim=96, depths=(2, 2, 18, 2), num_heads=(3, 6, 12, 24), window_size=8, pretrained_window_sizes=(0, 0, 0, 0))
args = types.SimpleNamespace(result_dir="/tmp", language_model_name=tcfg, image_model_name=scfg, image_model_train=False,
                             transformer_model_name=tcfg, compute_dtype="bf16")
model = MyModel(args)
init_t5_(model.language_model, seed=1); init_swin_(model.image_model, seed=2); init_t5_(model.transformer, seed=3)
model = model.to(dev).eval()
model.transformer.config.eos_token_id = -1                      # never finishes early: 20 tokens per sample
g = torch.Generator().manual_seed(0)
px = torch.randn(a.batch, 3, 256, 256, generator=g).to(dev)
src = torch.randint(2, 32000, (a.batch, 32), generator=g).to(dev)


def timed(fn, warm=3):
    for _ in range(warm):                                       # eager warm-up, graph capture, first replay
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    out = fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e), out


with torch.no_grad():
    t_enc, (emb, B, Le) = timed(lambda: model._concat_embeddings({"pixel_values": px}, {"input_ids": src}))
    t_kv, ids = timed(lambda: greedy_generate(model.transformer, emb, B, Le))
    t_rc, ids2 = timed(lambda: greedy_generate_recompute(model.transformer, emb, B, Le), warm=2)
    t_all, ids3 = timed(lambda: model({"pixel_values": px}, {"input_ids": src}, return_loss=False))
new_tokens = a.batch * (ids.shape[1] - 1)
print(json.dumps({"workload": "greedy decode, Swin-S/256 + T5-base, bf16", "batch": a.batch, "new_tokens_per_sample": ids.shape[1] - 1,
                  "towers_ms": t_enc, "kv_cached_decode_ms": t_kv, "prefix_recompute_decode_ms": t_rc, "end_to_end_ms": t_all,
                  "kv_cached_tokens_per_s": new_tokens / (t_kv * 1e-3), "end_to_end_tokens_per_s": new_tokens / (t_all * 1e-3),
                  "ids_equal_to_recompute_prefix": float((ids[:, :ids2.shape[1]] == ids2[:, :ids.shape[1]]).float().mean())}))
