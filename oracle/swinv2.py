"""Oracle: Swin-V2 encoder restated in plain torch fp32.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  HF/ = site-packages/transformers (5.5.0);
all line numbers refer to HF/models/swinv2/modeling_swinv2.py.  Takes a flat HF-style state dict
with `Swinv2Model` key names.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F


@dataclass
class SwinDims:
    """Subset of HF/models/swinv2/configuration_swinv2.py:56-73 the path reads."""
    image_size: int = 256
    patch_size: int = 4
    num_channels: int = 3
    embed_dim: int = 128
    depths: tuple = (2, 2, 18, 2)
    num_heads: tuple = (4, 8, 16, 32)
    window_size: int = 8
    pretrained_window_sizes: tuple = (0, 0, 0, 0)
    eps: float = 1e-5
    cpb_hidden: int = 512

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def out_width(self) -> int:
        return self.embed_dim * 2 ** (len(self.depths) - 1)

    @property
    def out_tokens(self) -> int:
        return (self.grid // 2 ** (len(self.depths) - 1)) ** 2

    @staticmethod
    def named(name: str, image_size: int = 256, window_size: int = 8) -> "SwinDims":
        table = {
            "swinv2-tiny": dict(embed_dim=96, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24)),
            "swinv2-small": dict(embed_dim=96, depths=(2, 2, 18, 2), num_heads=(3, 6, 12, 24)),
            "swinv2-base": dict(embed_dim=128, depths=(2, 2, 18, 2), num_heads=(4, 8, 16, 32)),
        }
        return SwinDims(image_size=image_size, window_size=window_size, **table[name])


def window_and_shift(res: int, window: int, block_idx: int) -> tuple[int, int]:
    """Swinv2Layer._compute_window_shift :622-625 with the stage's shift rule :733."""
    w = min(res, window)
    s = 0 if block_idx % 2 == 0 else window // 2
    if res <= w:
        s = 0
    return w, s


def coords_table(w: int, pretrained_w: int) -> torch.Tensor:
    """create_coords_table_and_index :489-508 -> ((2w-1)^2, 2) fp32."""
    r = torch.arange(-(w - 1), w, dtype=torch.int64).float()
    t = torch.stack(torch.meshgrid([r, r], indexing="ij")).permute(1, 2, 0).contiguous()
    if pretrained_w > 0:
        t = t / (pretrained_w - 1)
    elif w > 1:
        t = t / (w - 1)
    t = t * 8
    t = torch.sign(t) * torch.log2(torch.abs(t) + 1.0) / math.log2(8)
    return t.view(-1, 2)


def position_index(w: int) -> torch.Tensor:
    """create_coords_table_and_index :510-522 -> (w*w, w*w) int64."""
    c = torch.stack(torch.meshgrid([torch.arange(w), torch.arange(w)], indexing="ij")).flatten(1)
    rel = (c[:, :, None] - c[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += w - 1
    rel[:, :, 1] += w - 1
    rel[:, :, 0] *= 2 * w - 1
    return rel.sum(-1)


def cpb_bias(sd: dict, prefix: str, w: int, pretrained_w: int, heads: int) -> torch.Tensor:
    """Continuous position bias :450-460: 16*sigmoid(MLP(coords)[index]) -> (h, N, N)."""
    t = coords_table(w, pretrained_w)
    hid = torch.relu(t @ sd[prefix + "continuous_position_bias_mlp.0.weight"].T
                     + sd[prefix + "continuous_position_bias_mlp.0.bias"])
    tab = hid @ sd[prefix + "continuous_position_bias_mlp.2.weight"].T        # ((2w-1)^2, h)
    n = w * w
    bias = tab[position_index(w).view(-1)].view(n, n, heads).permute(2, 0, 1)
    return 16 * torch.sigmoid(bias)


def shift_mask(h: int, w_: int, window: int, shift: int) -> torch.Tensor | None:
    """Swinv2Layer.get_attn_mask :627-653 -> (nW, N, N) with values {0, -100}."""
    if shift == 0:
        return None
    img = torch.zeros(h, w_)
    cnt = 0
    for hs in (slice(0, -window), slice(-window, -shift), slice(-shift, None)):
        for ws in (slice(0, -window), slice(-window, -shift), slice(-shift, None)):
            img[hs, ws] = cnt
            cnt += 1
    mw = img.view(h // window, window, w_ // window, window).permute(0, 2, 1, 3).reshape(-1, window * window)
    diff = mw[:, None, :] - mw[:, :, None]
    return torch.where(diff != 0, -100.0, 0.0)


def _partition(x, w):
    """window_partition :146-155."""
    b, h, w_, c = x.shape
    return x.view(b, h // w, w, w_ // w, w, c).permute(0, 1, 3, 2, 4, 5).reshape(-1, w * w, c)


def _reverse(win, w, h, w_):
    """window_reverse :159-166."""
    c = win.shape[-1]
    return win.view(-1, h // w, w_ // w, w, w, c).permute(0, 1, 3, 2, 4, 5).reshape(-1, h, w_, c)


def swinv2_layer(x: torch.Tensor, sd: dict, prefix: str, res: int, heads: int, window: int,
                 shift: int, pretrained_w: int, eps: float) -> torch.Tensor:
    """Swinv2Layer.forward :662-715 (res divisible by window: maybe_pad :655-660 is a no-op for every
    BASELINE geometry) with Swinv2SelfAttention.forward :421-487 inlined."""
    b, l, c = x.shape
    d = c // heads
    a = prefix + "attention.self."
    xs = x.view(b, res, res, c)
    if shift > 0:
        xs = torch.roll(xs, shifts=(-shift, -shift), dims=(1, 2))
    win = _partition(xs, window)                                       # (b*nW, N, c)
    bw, n, _ = win.shape
    q = (win @ sd[a + "query.weight"].T + sd[a + "query.bias"]).view(bw, n, heads, d).transpose(1, 2)
    k = (win @ sd[a + "key.weight"].T).view(bw, n, heads, d).transpose(1, 2)           # no bias :417
    v = (win @ sd[a + "value.weight"].T + sd[a + "value.bias"]).view(bw, n, heads, d).transpose(1, 2)
    s = F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-1, -2)              # cosine :445-447
    s = s * torch.clamp(sd[a + "logit_scale"], max=math.log(1.0 / 0.01)).exp()        # :448-449
    s = s + cpb_bias(sd, a, window, pretrained_w, heads)[None]
    mask = shift_mask(res, res, window, shift)
    if mask is not None:                                                                # added twice :465-468
        nw = mask.shape[0]
        s = (s.view(bw // nw, nw, heads, n, n) + 2 * mask[None, :, None]).view(bw, heads, n, n)
    p = torch.softmax(s, dim=-1)
    ctx = (p @ v).transpose(1, 2).reshape(bw, n, c)
    ctx = ctx @ sd[prefix + "attention.output.dense.weight"].T + sd[prefix + "attention.output.dense.bias"]
    ctx = _reverse(ctx, window, res, res)
    if shift > 0:
        ctx = torch.roll(ctx, shifts=(shift, shift), dims=(1, 2))
    ctx = ctx.reshape(b, l, c)
    hdn = x + F.layer_norm(ctx, (c,), sd[prefix + "layernorm_before.weight"],
                           sd[prefix + "layernorm_before.bias"], eps)                  # res-post-norm :707-708
    m = F.gelu(hdn @ sd[prefix + "intermediate.dense.weight"].T + sd[prefix + "intermediate.dense.bias"])
    m = m @ sd[prefix + "output.dense.weight"].T + sd[prefix + "output.dense.bias"]
    return hdn + F.layer_norm(m, (c,), sd[prefix + "layernorm_after.weight"],
                              sd[prefix + "layernorm_after.bias"], eps)                # :712


def patch_merging(x, sd, prefix, res, eps):
    """Swinv2PatchMerging.forward :365-388: 2x2 concat order (0,0),(1,0),(0,1),(1,1); reduce THEN norm."""
    b, l, c = x.shape
    g = x.view(b, res, res, c)
    cat = torch.cat([g[:, 0::2, 0::2], g[:, 1::2, 0::2], g[:, 0::2, 1::2], g[:, 1::2, 1::2]], -1)
    y = cat.view(b, -1, 4 * c) @ sd[prefix + "reduction.weight"].T
    return F.layer_norm(y, (2 * c,), sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], eps)


def swinv2_forward(pixel_values: torch.Tensor, sd: dict, dims: SwinDims) -> torch.Tensor:
    """Swinv2Model.forward(pixel_values).last_hidden_state :933-974 (eval mode: drop-path off)."""
    x = F.conv2d(pixel_values, sd["embeddings.patch_embeddings.projection.weight"],
                 sd["embeddings.patch_embeddings.projection.bias"], stride=dims.patch_size)  # :329
    x = x.flatten(2).transpose(1, 2)
    c = dims.embed_dim
    x = F.layer_norm(x, (c,), sd["embeddings.norm.weight"], sd["embeddings.norm.bias"], dims.eps)  # :273
    res = dims.grid
    for s, depth in enumerate(dims.depths):
        for i in range(depth):
            w, sh = window_and_shift(res, dims.window_size, i)
            x = swinv2_layer(x, sd, f"encoder.layers.{s}.blocks.{i}.", res, dims.num_heads[s], w, sh,
                             dims.pretrained_window_sizes[s], dims.eps)
        if s < len(dims.depths) - 1:
            x = patch_merging(x, sd, f"encoder.layers.{s}.downsample.", res, dims.eps)
            res //= 2
            c *= 2
    return F.layer_norm(x, (c,), sd["layernorm.weight"], sd["layernorm.bias"], dims.eps)    # :969
