"""Oracle: T5 encoder / decoder / LM-head loss / greedy decode, restated in plain torch fp32.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  HF/ = site-packages/transformers (5.5.0).
All functions take a flat HF-style state dict `sd` (key names of T5ForConditionalGeneration /
T5EncoderModel) so the same weights feed the oracle, the reference and the CUDA build.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class T5Dims:
    """Subset of HF/models/t5/configuration_t5.py:44-83 the path reads."""
    vocab_size: int = 32128
    d_model: int = 512
    d_kv: int = 64
    d_ff: int = 2048
    num_layers: int = 6
    num_decoder_layers: int | None = None
    num_heads: int = 8
    num_buckets: int = 32
    max_distance: int = 128
    eps: float = 1e-6
    pad_token_id: int = 0
    eos_token_id: int = 1
    decoder_start_token_id: int = 0

    @property
    def n_dec(self) -> int:
        return self.num_layers if self.num_decoder_layers is None else self.num_decoder_layers

    @staticmethod
    def named(name: str) -> "T5Dims":
        table = {
            "t5-small": dict(d_model=512, d_ff=2048, num_layers=6, num_heads=8),
            "t5-base": dict(d_model=768, d_ff=3072, num_layers=12, num_heads=12),
            "t5-large": dict(d_model=1024, d_ff=4096, num_layers=24, num_heads=16),
        }
        return T5Dims(**table[name])


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """T5LayerNorm.forward, HF/models/t5/modeling_t5.py:55-68: no mean, no bias, fp32 variance."""
    var = x.float().pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + eps))


def relative_position_bucket(rel: torch.Tensor, bidirectional: bool, num_buckets: int,
                             max_distance: int) -> torch.Tensor:
    """T5Attention._relative_position_bucket, HF/models/t5/modeling_t5.py:189-234 (integer)."""
    buckets = torch.zeros_like(rel)
    if bidirectional:
        num_buckets //= 2
        buckets = buckets + (rel > 0).long() * num_buckets
        rel = rel.abs()
    else:
        rel = -torch.minimum(rel, torch.zeros_like(rel))
    max_exact = num_buckets // 2
    is_small = rel < max_exact
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact)
                         * (num_buckets - max_exact)).long()
    large = torch.minimum(large, torch.full_like(large, num_buckets - 1))
    return buckets + torch.where(is_small, rel, large)


def t5_bias(table: torch.Tensor, lq: int, lk: int, bidirectional: bool, dims: T5Dims,
            q_offset: int = 0) -> torch.Tensor:
    """T5Attention.compute_bias, HF/models/t5/modeling_t5.py:236-251 -> (h, lq, lk)."""
    ctx = torch.arange(lq)[:, None] + q_offset
    mem = torch.arange(lk)[None, :]
    bucket = relative_position_bucket(mem - ctx, bidirectional, dims.num_buckets, dims.max_distance)
    return table[bucket].permute(2, 0, 1)


def t5_attention(x: torch.Tensor, kv_src: torch.Tensor, sd: dict, prefix: str, dims: T5Dims,
                 bias: torch.Tensor | None, causal: bool) -> torch.Tensor:
    """T5Attention.forward, HF/models/t5/modeling_t5.py:253-344.

    q/k/v/o have no bias (:178-181); scores are NOT scaled by 1/sqrt(d) (:308); the additive
    position bias (and, for the decoder, the finfo.min causal mask, :704) is added before an fp32
    softmax (:331); cross-attention uses a zero bias (:312-315).
    """
    b, lq, _ = x.shape
    lk = kv_src.shape[1]
    h, dk = dims.num_heads, dims.d_kv
    q = (x @ sd[prefix + "q.weight"].T).view(b, lq, h, dk).transpose(1, 2)
    k = (kv_src @ sd[prefix + "k.weight"].T).view(b, lk, h, dk).transpose(1, 2)
    v = (kv_src @ sd[prefix + "v.weight"].T).view(b, lk, h, dk).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    if bias is not None:
        s = s + bias[None]
    if causal:
        qi = torch.arange(lq)[:, None] + (lk - lq)
        ki = torch.arange(lk)[None, :]
        s = s + torch.where(ki > qi, torch.finfo(s.dtype).min, 0.0)
    p = torch.softmax(s.float(), dim=-1).to(s.dtype)
    o = (p @ v).transpose(1, 2).reshape(b, lq, h * dk)
    return o @ sd[prefix + "o.weight"].T


def _ff(x, sd, prefix, dims):
    """T5LayerFF + T5DenseActDense, HF/models/t5/modeling_t5.py:92-103,146-150 (ReLU, no biases)."""
    n = rms_norm(x, sd[prefix + "layer_norm.weight"], dims.eps)
    hdn = torch.relu(n @ sd[prefix + "DenseReluDense.wi.weight"].T)
    return x + hdn @ sd[prefix + "DenseReluDense.wo.weight"].T


def t5_stack(x: torch.Tensor, sd: dict, stack: str, dims: T5Dims,
             enc_out: torch.Tensor | None = None) -> torch.Tensor:
    """T5Stack.forward (HF/models/t5/modeling_t5.py:637-792) in eval mode (dropout off).

    `stack` is "encoder." or "decoder." (prefix inside `sd`).  Position bias is computed from block
    0's table and shared by all blocks (:758).  Pre-norm residual wiring per T5LayerSelfAttention
    :356-377, T5LayerCrossAttention :387-408.
    """
    is_dec = enc_out is not None
    n_layers = dims.n_dec if is_dec else dims.num_layers
    l = x.shape[1]
    table = sd[f"{stack}block.0.layer.0.SelfAttention.relative_attention_bias.weight"]
    bias = t5_bias(table, l, l, bidirectional=not is_dec, dims=dims)
    for i in range(n_layers):
        p = f"{stack}block.{i}.layer."
        n = rms_norm(x, sd[p + "0.layer_norm.weight"], dims.eps)
        x = x + t5_attention(n, n, sd, p + "0.SelfAttention.", dims, bias, causal=is_dec)
        if is_dec:
            n = rms_norm(x, sd[p + "1.layer_norm.weight"], dims.eps)
            x = x + t5_attention(n, enc_out, sd, p + "1.EncDecAttention.", dims, None, causal=False)
            x = _ff(x, sd, p + "2.", dims)
        else:
            x = _ff(x, sd, p + "1.", dims)
    return rms_norm(x, sd[f"{stack}final_layer_norm.weight"], dims.eps)


def shift_right(labels: torch.Tensor, dims: T5Dims) -> torch.Tensor:
    """T5PreTrainedModel._shift_right, HF/models/t5/modeling_t5.py:595-614 (integer)."""
    out = labels.new_zeros(labels.shape)
    out[..., 1:] = labels[..., :-1]
    out[..., 0] = dims.decoder_start_token_id
    return out.masked_fill(out == -100, dims.pad_token_id)


def t5_lm_loss(enc_in: torch.Tensor, labels: torch.Tensor, sd: dict, dims: T5Dims) -> torch.Tensor:
    """T5ForConditionalGeneration.forward(inputs_embeds, labels).loss,
    HF/models/t5/modeling_t5.py:1070-1117: encoder on inputs_embeds, decoder on shift_right(labels),
    decoder output * d_model**-0.5 (:1107-1108), tied LM head (:956-960), CE(ignore_index=-100)."""
    enc = t5_stack(enc_in, sd, "encoder.", dims)
    dec_in = F.embedding(shift_right(labels, dims), sd["shared.weight"])
    dec = t5_stack(dec_in, sd, "decoder.", dims, enc_out=enc)
    logits = (dec * dims.d_model ** -0.5) @ sd["lm_head.weight"].T
    return F.cross_entropy(logits.view(-1, logits.shape[-1]), labels.reshape(-1), ignore_index=-100)


def t5_greedy_decode(enc_in: torch.Tensor, sd: dict, dims: T5Dims, max_new_tokens: int = 20) -> torch.Tensor:
    """Greedy `generate(inputs_embeds=...)`, HF/generation/utils.py:2658-2800 with the defaults the
    reference relies on (max_length 20 new tokens, start id 0, EOS 1, pad 0; finished rows keep
    emitting pad).  Recomputes the full decoder prefix each step: same arithmetic as the KV cache."""
    enc = t5_stack(enc_in, sd, "encoder.", dims)
    b = enc_in.shape[0]
    ids = torch.full((b, 1), dims.decoder_start_token_id, dtype=torch.long)
    unfinished = torch.ones(b, dtype=torch.bool)
    for _ in range(max_new_tokens):
        dec = t5_stack(F.embedding(ids, sd["shared.weight"]), sd, "decoder.", dims, enc_out=enc)
        logits = (dec[:, -1] * dims.d_model ** -0.5) @ sd["lm_head.weight"].T
        nxt = logits.float().argmax(-1)
        nxt = torch.where(unfinished, nxt, torch.full_like(nxt, dims.pad_token_id))
        ids = torch.cat([ids, nxt[:, None]], dim=1)
        unfinished = unfinished & (nxt != dims.eos_token_id)
        if not unfinished.any():
            break
    return ids
