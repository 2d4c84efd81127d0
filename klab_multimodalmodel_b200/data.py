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
