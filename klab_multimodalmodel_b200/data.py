"""N3: the wire format of the RedCaps span-corruption pre-training step (BASELINE config 3) -- how a raw caption
becomes the (source text, target text) pair that the tokenizer turns into `source_encoding` / `target_encoding` of
`MyModel.forward`.  Restates /root/reference/modules/loader.py:56-72 (RedCapsDatasetLoader.__getitem__):

  * a space is put in front of every `.`, `,`, `!`, `?`, then the caption is split on whitespace into words (:58-59);
  * `int(0.15 * n_words) + 1` word positions are drawn as the head of `torch.randperm(n_words)` (:61) -- so even a one-word
    caption loses a word, and every masked WORD gets its own sentinel (adjacent masked words are not merged into one span,
    unlike canonical T5 span corruption);
  * walking the words left to right, the j-th masked word is replaced by `<extra_id_j>` in the source; the target is
    `<extra_id_0> word_0 <extra_id_1> word_1 ... <extra_id_k>` (:63-72).

Pure host code (it runs inside data-loader workers); the random draw goes through torch's generator exactly as the reference's
does, so under the same seed the output is identical (tests/golden/span_corruption.json comes from the unmodified reference).
"""
from __future__ import annotations

import torch

MASK_FRACTION = 0.15
_PUNCTUATION = ".,!?"


def split_words(caption: str) -> list[str]:
    for ch in _PUNCTUATION:
        caption = caption.replace(ch, " " + ch)
    return caption.split()


def span_corrupt(caption: str, generator: torch.Generator | None = None) -> tuple[str, str]:
    """-> (source text with sentinels, target text).  `generator=None` draws from torch's global generator, as the reference does."""
    words = split_words(caption)
    n = len(words)
    perm = torch.randperm(n) if generator is None else torch.randperm(n, generator=generator)
    masked = set(perm[:int(n * MASK_FRACTION) + 1].tolist())
    target = ["<extra_id_0>"]
    j = 0
    for i in range(n):
        if i in masked:
            target += [words[i], f"<extra_id_{j + 1}>"]
            words[i] = f"<extra_id_{j}>"
            j += 1
    return " ".join(words), " ".join(target)


def restore(source: str, target: str) -> str:
    """Inverse of `span_corrupt` (up to the whitespace normalisation of `split_words`): fills the sentinels of the source with
    the words the target carries.  A size-independent property for tests: restore(*span_corrupt(c)) == ' '.join(split_words(c))."""
    fills, toks = {}, target.split()
    for a, b in zip(toks[0::2], toks[1::2]):
        fills[a] = b
    return " ".join(fills.get(w, w) for w in source.split())


# ------------------------------------------------------------------------------------------------------------------
# N2: the host -> device input path of /root/reference/train.py:55-59
#   images = image_processor(images, return_tensors="pt").to(device_id)          # resize / rescale 1/255 / normalise ON THE HOST
#   source_encoding = tokenizer(...).to(device_id); target_encoding = tokenizer(...).to(device_id)
#   loss = model(images, source_encoding, target_encoding); loss_counter.add_loss('train', loss.item())
# The drop-in keeps the call shapes but moves the per-pixel arithmetic to the GPU (klab_image_normalize) and the copies off the
# critical path (pinned staging buffers, a side stream, one batch of look-ahead).
# ------------------------------------------------------------------------------------------------------------------
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class DeviceBatch(dict):
    """What `GpuImageProcessor.__call__` returns: a mapping with `pixel_values` like transformers' BatchFeature, so
    `model(images, ...)` can unpack it (`**images`, /root/reference/models/model.py:22) and `.to(device)` keeps working."""

    def to(self, device, non_blocking: bool = False):
        dev = torch.device(device) if not isinstance(device, torch.device) else device
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        proc = self.pop("_klab_processor", None)
        if proc is None:
            return DeviceBatch({k: (v.to(dev, non_blocking=non_blocking) if torch.is_tensor(v) else v) for k, v in self.items()})
        return DeviceBatch({"pixel_values": proc.finish(self["_raw"], dev)})


class GpuImageProcessor:
    """Drop-in for the `AutoImageProcessor` of /root/reference/train.py:39,55 (transformers ViTImageProcessor semantics for
    Swin-V2 checkpoints: `do_rescale` by 1/255, `do_normalize` with the ImageNet mean / std; the dataset already delivers
    256 x 256 images, /root/reference/modules/loader.py:15, so `do_resize` to the same size is the identity -- other sizes raise).

        image_processor = GpuImageProcessor.from_pretrained(args.image_model_name)      # instead of AutoImageProcessor
        images = image_processor(images, return_tensors="pt").to(device_id)             # train.py:55, unchanged

    `__call__` only stacks the raw batch (uint8 or float, as the data loader made it) into a pinned staging buffer; `.to(device)`
    issues ONE asynchronous copy of the raw pixels and runs the rescale + normalise kernel on the device
    (ops.image_normalize: same rounding sequence as the numpy code of the host processor)."""

    def __init__(self, size=None, rescale_factor: float = 1.0 / 255.0, image_mean=IMAGENET_MEAN, image_std=IMAGENET_STD,
                 do_rescale: bool = True, do_normalize: bool = True):
        self.size = size
        self.rescale_factor = float(rescale_factor) if do_rescale else 1.0
        self.image_mean = tuple(float(v) for v in image_mean) if do_normalize else None
        self.image_std = tuple(float(v) for v in image_std) if do_normalize else None
        self._pinned: dict = {}
        self._flip = 0

    @classmethod
    def from_pretrained(cls, source, **kw):
        """Reads preprocessor_config.json of a local checkpoint directory / cached snapshot when there is one (rescale factor,
        mean, std, size); otherwise the ViTImageProcessor defaults that every microsoft/swinv2-* checkpoint ships."""
        import json
        import os

        from .modeling import resolve_snapshot
        path = source if isinstance(source, str) and os.path.isdir(source) else (resolve_snapshot(source) if isinstance(source, str) else None)
        cfg = {}
        if path and os.path.exists(os.path.join(path, "preprocessor_config.json")):
            with open(os.path.join(path, "preprocessor_config.json")) as fh:
                cfg = json.load(fh)
        size = cfg.get("size")
        if isinstance(size, dict):
            size = (size.get("height", size.get("shortest_edge")), size.get("width", size.get("shortest_edge")))
        args = dict(size=size, rescale_factor=cfg.get("rescale_factor", 1.0 / 255.0), image_mean=cfg.get("image_mean", IMAGENET_MEAN),
                    image_std=cfg.get("image_std", IMAGENET_STD), do_rescale=cfg.get("do_rescale", True),
                    do_normalize=cfg.get("do_normalize", True))
        args.update(kw)
        return cls(**args)

    def _stage(self, images) -> torch.Tensor:
        if isinstance(images, (list, tuple)):
            images = torch.stack([torch.as_tensor(im) for im in images])
        images = torch.as_tensor(images)
        if images.dim() == 3:
            images = images[None]
        if images.dim() != 4:
            raise ValueError(f"GpuImageProcessor: expected (B, C, H, W) images, got shape {tuple(images.shape)}")
        if images.shape[-1] in (1, 3) and images.shape[1] not in (1, 3):                 # channels-last input (PIL / numpy order)
            images = images.permute(0, 3, 1, 2)
        if images.dtype not in (torch.uint8, torch.float32):
            images = images.float()
        if self.size and tuple(images.shape[-2:]) != tuple(self.size):
            raise NotImplementedError(f"GpuImageProcessor: images are {tuple(images.shape[-2:])} but the model expects {tuple(self.size)}; "
                                      "resize in the dataset (the reference's loader does, loader.py:15)")
        if images.is_cuda or (images.is_pinned() and images.is_contiguous()):
            return images.contiguous()                     # already on the device / already page-locked: no staging copy
        key = (tuple(images.shape), images.dtype)
        bufs = self._pinned.get(key)
        if bufs is None:                                   # two pinned staging buffers per geometry: the copy of batch i may still be
            bufs = self._pinned[key] = [torch.empty(images.shape, dtype=images.dtype).pin_memory() for _ in range(2)]   # in flight
        self._flip ^= 1
        buf = bufs[self._flip]
        buf.copy_(images)
        return buf

    def __call__(self, images, return_tensors="pt", **_):
        raw = self._stage(images)
        out = DeviceBatch({"_raw": raw})
        out["_klab_processor"] = self
        return out

    def finish(self, raw: torch.Tensor, device) -> torch.Tensor:
        from . import ops as O
        with torch.cuda.device(device):
            dev_raw = raw if raw.is_cuda else raw.to(device, non_blocking=True)
            return O.image_normalize(dev_raw, self.rescale_factor, self.image_mean, self.image_std)

    def preprocess_on_host(self, images) -> torch.Tensor:
        """The same arithmetic with numpy on the host (what transformers' slow ViTImageProcessor does): the parity reference."""
        import numpy as np
        x = torch.as_tensor(images).numpy()
        r = (x.astype(np.float64) * self.rescale_factor).astype(np.float32)
        if self.image_mean is not None:
            m = np.asarray(self.image_mean, dtype=np.float32)[None, :, None, None]
            s = np.asarray(self.image_std, dtype=np.float32)[None, :, None, None]
            r = (r - m) / s
        return torch.from_numpy(r)


class DevicePrefetcher:
    """Iterates over batches of HOST tensors one step ahead of the consumer: while step i computes, the tensors of batch i + 1
    are copied (pinned -> device, asynchronously) on a side stream and, for raw images, normalised there.  `for batch in
    DevicePrefetcher(loader_like, device, transform)`; `transform(host_batch) -> device_batch` runs inside the side stream."""

    def __init__(self, batches, device, transform):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.transform = transform
        self.stream = torch.cuda.Stream(self.device)
        self._next = None
        self._preload()

    def _preload(self):
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        with torch.cuda.stream(self.stream):
            out = self.transform(host)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._next = (out, ev)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        out, ev = self._next
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)                                   # device-side wait: the host does not block
        for t in _tensors_of(out):
            t.record_stream(cur)
        self._preload()
        return out


def _tensors_of(obj):
    if torch.is_tensor(obj):
        yield obj
    elif isinstance(obj, dict):
        for v in obj.values():
            yield from _tensors_of(v)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            yield from _tensors_of(v)
