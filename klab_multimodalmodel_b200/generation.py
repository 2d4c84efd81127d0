"""Greedy caption decoding: `transformer.generate(inputs_embeds=...)` of /root/reference/models/model.py:28 with the defaults
the reference relies on (HF/generation/utils.py:765-804,2658-2800): greedy, 20 new tokens, start id 0, EOS 1, pad 0, finished
rows keep emitting pad, stop as soon as every row has finished.

This first version re-runs the decoder on the whole prefix every step (same arithmetic as HF's KV cache, O(T^2) work for
T <= 20); the cached single-token path reuses the same kernels with q_offset and is the next step.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import functional as Fn
from . import ops as O


@torch.no_grad()
def greedy_generate(tr, embeds, B, Le, max_new_tokens: int = 20):
    cfg = tr.config
    cd = embeds.dtype
    dev = embeds.device
    enc = tr.encoder.run_blocks(embeds, B, Le, tr.cache)
    enc, _ = O.rmsnorm_fwd(enc, tr.encoder.final_layer_norm.weight, cfg.layer_norm_epsilon, save_stats=False)
    tab = tr.cache.get([tr.shared.weight], cd)
    V, d = tr.shared.weight.shape
    ids = torch.full((B, max_new_tokens + 1), cfg.pad_token_id, dtype=torch.int64, device=dev)
    ids[:, 0] = cfg.decoder_start_token_id
    unfinished = torch.ones(B, dtype=torch.int32, device=dev)
    n_out = 1
    for t in range(max_new_tokens):
        Lt = t + 1
        prefix = ids[:, :Lt].contiguous()
        x = O.embedding_fwd(prefix, tab)
        x = tr.decoder.run_blocks(x, B, Lt, tr.cache, enc_out=enc, Le=Le)
        last = x.view(B, Lt, d)[:, -1, :]                               # strided rows: the norm kernel takes a row stride
        n, _ = O.rmsnorm_fwd(last, tr.decoder.final_layer_norm.weight, cfg.layer_norm_epsilon, save_stats=False)
        logits = O.linear_fwd(n, tab, alpha=d ** -0.5, out_dtype=torch.float32)
        L.check(L.lib().klab_greedy_step(torch.cuda.current_stream().cuda_stream, B, V, logits.data_ptr(), logits.stride(0),
                                         ids.data_ptr(), ids.stride(0), Lt, unfinished.data_ptr(), cfg.pad_token_id, cfg.eos_token_id))
        n_out = Lt + 1
        if not bool(unfinished.cpu().any()):                           # HF checks the stopping criteria every step too
            break
    return ids[:, :n_out]
