"""Oracle: the reference's `MyModel.forward` (/root/reference/models/model.py:19-28) end to end,
plus the deterministic synthetic weights / inputs every parity test and the bench share.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import zlib

import numpy as np
import torch

from .swinv2 import SwinDims, swinv2_forward
from .t5 import T5Dims, t5_greedy_decode, t5_lm_loss, t5_stack
import torch.nn.functional as F


def caption_embeddings(pixel_values, src_ids, sds: dict, lm: T5Dims, swin: SwinDims):
    """models/model.py:20-23: frozen text encoder (no grad), Swin encoder, concat along the sequence."""
    with torch.no_grad():
        lm_sd = sds["language_model"]
        lang = t5_stack(F.embedding(src_ids, lm_sd["shared.weight"]), lm_sd, "encoder.", lm)
    img = swinv2_forward(pixel_values, sds["image_model"], swin)
    return torch.cat((img, lang), dim=1)


def caption_loss(pixel_values, src_ids, tgt_ids, sds: dict, lm: T5Dims, swin: SwinDims, tr: T5Dims):
    """models/model.py:25-26: loss of the trainable T5 on [image tokens; text tokens]."""
    return t5_lm_loss(caption_embeddings(pixel_values, src_ids, sds, lm, swin), tgt_ids, sds["transformer"], tr)


def caption_generate(pixel_values, src_ids, sds: dict, lm: T5Dims, swin: SwinDims, tr: T5Dims,
                     max_new_tokens: int = 20):
    """models/model.py:28: greedy `generate(inputs_embeds=...)`."""
    with torch.no_grad():
        emb = caption_embeddings(pixel_values, src_ids, sds, lm, swin)
        return t5_greedy_decode(emb, sds["transformer"], tr, max_new_tokens)


# ----------------------------------------------------------------------------------------------
# Deterministic synthetic weights and inputs (numpy PCG64: reproducible on any box, no checkpoints)
# ----------------------------------------------------------------------------------------------

def t5_param_shapes(d: T5Dims, encoder_only: bool = False) -> dict:
    """Key -> shape for T5ForConditionalGeneration / T5EncoderModel state dicts."""
    inner = d.num_heads * d.d_kv
    out = {"shared.weight": (d.vocab_size, d.d_model)}

    def attn(p, bias):
        for n in "qkv":
            out[f"{p}{n}.weight"] = (inner, d.d_model)
        out[f"{p}o.weight"] = (d.d_model, inner)
        if bias:
            out[f"{p}relative_attention_bias.weight"] = (d.num_buckets, d.num_heads)

    def ff(p):
        out[f"{p}DenseReluDense.wi.weight"] = (d.d_ff, d.d_model)
        out[f"{p}DenseReluDense.wo.weight"] = (d.d_model, d.d_ff)
        out[f"{p}layer_norm.weight"] = (d.d_model,)

    out["encoder.embed_tokens.weight"] = (d.vocab_size, d.d_model)
    for i in range(d.num_layers):
        p = f"encoder.block.{i}.layer."
        attn(p + "0.SelfAttention.", i == 0)
        out[p + "0.layer_norm.weight"] = (d.d_model,)
        ff(p + "1.")
    out["encoder.final_layer_norm.weight"] = (d.d_model,)
    if encoder_only:
        return out
    out["decoder.embed_tokens.weight"] = (d.vocab_size, d.d_model)
    for i in range(d.n_dec):
        p = f"decoder.block.{i}.layer."
        attn(p + "0.SelfAttention.", i == 0)
        out[p + "0.layer_norm.weight"] = (d.d_model,)
        attn(p + "1.EncDecAttention.", False)
        out[p + "1.layer_norm.weight"] = (d.d_model,)
        ff(p + "2.")
    out["decoder.final_layer_norm.weight"] = (d.d_model,)
    out["lm_head.weight"] = (d.vocab_size, d.d_model)
    return out


def swin_param_shapes(d: SwinDims) -> dict:
    """Key -> shape for the Swinv2Model state dict (persistent keys only)."""
    out = {
        "embeddings.patch_embeddings.projection.weight": (d.embed_dim, d.num_channels, d.patch_size, d.patch_size),
        "embeddings.patch_embeddings.projection.bias": (d.embed_dim,),
        "embeddings.norm.weight": (d.embed_dim,),
        "embeddings.norm.bias": (d.embed_dim,),
    }
    c = d.embed_dim
    for s, depth in enumerate(d.depths):
        h = d.num_heads[s]
        for i in range(depth):
            p = f"encoder.layers.{s}.blocks.{i}."
            a = p + "attention.self."
            out[a + "logit_scale"] = (h, 1, 1)
            out[a + "continuous_position_bias_mlp.0.weight"] = (d.cpb_hidden, 2)
            out[a + "continuous_position_bias_mlp.0.bias"] = (d.cpb_hidden,)
            out[a + "continuous_position_bias_mlp.2.weight"] = (h, d.cpb_hidden)
            out[a + "query.weight"] = (c, c)
            out[a + "query.bias"] = (c,)
            out[a + "key.weight"] = (c, c)
            out[a + "value.weight"] = (c, c)
            out[a + "value.bias"] = (c,)
            out[p + "attention.output.dense.weight"] = (c, c)
            out[p + "attention.output.dense.bias"] = (c,)
            out[p + "layernorm_before.weight"] = (c,)
            out[p + "layernorm_before.bias"] = (c,)
            out[p + "intermediate.dense.weight"] = (4 * c, c)
            out[p + "intermediate.dense.bias"] = (4 * c,)
            out[p + "output.dense.weight"] = (c, 4 * c)
            out[p + "output.dense.bias"] = (c,)
            out[p + "layernorm_after.weight"] = (c,)
            out[p + "layernorm_after.bias"] = (c,)
        if s < len(d.depths) - 1:
            p = f"encoder.layers.{s}.downsample."
            out[p + "reduction.weight"] = (2 * c, 4 * c)
            out[p + "norm.weight"] = (2 * c,)
            out[p + "norm.bias"] = (2 * c,)
            c *= 2
    out["layernorm.weight"] = (c,)
    out["layernorm.bias"] = (c,)
    return out


_TIED = ("encoder.embed_tokens.weight", "decoder.embed_tokens.weight", "lm_head.weight")


def _hf_std(scope: str, key: str, shape, dims) -> float | None:
    """Standard deviations of HF's own initialisers (T5: HF/models/t5/modeling_t5.py:540-593 with factor 1;
    Swin-V2: HF/models/swinv2/modeling_swinv2.py:883-902, initializer_range 0.02)."""
    if scope == "image_model":
        return 0.02 if len(shape) > 1 else None
    d, dk, h, dff = dims.d_model, dims.d_kv, dims.num_heads, dims.d_ff
    if key.endswith(".q.weight"):
        return (d * dk) ** -0.5
    if key.endswith((".k.weight", ".v.weight", "wi.weight", "relative_attention_bias.weight")):
        return d ** -0.5
    if key.endswith(".o.weight"):
        return (h * dk) ** -0.5
    if key.endswith("wo.weight"):
        return dff ** -0.5
    return None


def _seeded_tensor(seed: int, scope: str, key: str, shape, style: str = "hot", dims=None) -> torch.Tensor:
    """style "hot": 1/sqrt(fan_in) matrices (unscaled T5 attention then has logits of std ~ sqrt(d_kv): sharply peaked
    softmaxes, a demanding fp32 test); style "hf": HF's initialiser scales (well conditioned: what bf16 parity is quoted on)."""
    rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(f"{scope}/{key}".encode())]))
    x = rng.standard_normal(size=shape, dtype=np.float32)
    std = _hf_std(scope, key, shape, dims) if style == "hf" else None
    if std is not None:
        x = x * np.float32(std)
    elif key.endswith("logit_scale"):
        x = np.float32(np.log(10.0)) + np.float32(0.3) * x
    elif key.endswith("relative_attention_bias.weight"):
        x = np.float32(0.5) * x
    elif key == "shared.weight":
        pass
    elif len(shape) == 1 and key.endswith("weight"):          # norm gains
        x = np.float32(1.0) + np.float32(0.1) * x
    elif key.endswith("bias"):
        x = np.float32(0.05) * x
    else:                                                      # matrices / conv: 1/sqrt(fan_in)
        fan_in = int(np.prod(shape[1:]))
        x = x * np.float32(fan_in ** -0.5)
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))


def seeded_state_dict(shapes: dict, seed: int, scope: str, style: str = "hot", dims=None) -> dict:
    sd = {}
    for key, shape in shapes.items():
        if key in _TIED:
            continue
        sd[key] = _seeded_tensor(seed, scope, key, shape, style, dims)
    for key in _TIED:
        if key in shapes:
            sd[key] = sd["shared.weight"]
    return sd


def seeded_state_dicts(lm: T5Dims, swin: SwinDims, tr: T5Dims, seed: int = 0, style: str = "hot") -> dict:
    """Weights for the three sub-models of MyModel (models/model.py:14-17), fp32, CPU."""
    return {
        "language_model": seeded_state_dict(t5_param_shapes(lm, encoder_only=True), seed, "language_model", style, lm),
        "image_model": seeded_state_dict(swin_param_shapes(swin), seed, "image_model", style, swin),
        "transformer": seeded_state_dict(t5_param_shapes(tr), seed, "transformer", style, tr),
    }


def seeded_inputs(batch: int, swin: SwinDims, vocab: int, l_src: int, l_tgt: int, seed: int = 1234,
                  ignore_tail: bool = False):
    """Synthetic inputs per SURVEY.md 8(d): pixels ~ N(0,1); ids ~ U{2..vocab-29}; last target id = EOS;
    `ignore_tail` sets the trailing 25% of every target to -100 (ignore-index variant)."""
    rng = np.random.Generator(np.random.PCG64([seed, 7]))
    px = rng.standard_normal(size=(batch, swin.num_channels, swin.image_size, swin.image_size), dtype=np.float32)
    hi = max(3, vocab - 28)
    src = rng.integers(2, hi, size=(batch, l_src), dtype=np.int64)
    tgt = rng.integers(2, hi, size=(batch, l_tgt), dtype=np.int64)
    tgt[:, -1] = 1
    if ignore_tail:
        tgt[:, l_tgt - max(1, l_tgt // 4):] = -100
    return torch.from_numpy(px), torch.from_numpy(src), torch.from_numpy(tgt)
