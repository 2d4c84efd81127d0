"""Golden vectors for the span-corruption target construction: run the UNMODIFIED reference
`RedCapsDatasetLoader.__getitem__` (/root/reference/modules/loader.py:56-77) on fixed captions under fixed torch seeds and
record (source text, target text).  The loader's heavy imports (pycocotools, image files) are stubbed: only its text logic
runs.  Usage (build container only; /root/reference does not exist on the GPU box):
    python tests/golden/make_span_golden.py > tests/golden/span_corruption.json"""
import json
import sys
import types

import torch

CAPTIONS = [
    "a dog catches a frisbee in the park.",
    "my grandmother's 1962 kitchen, restored! what do you think?",
    "itap of a bridge at sunset, somewhere in norway",
    "two cats",
    "cat",
    "first snow of the year, taken from my window this morning. it was cold, really cold!",
    "   leading and   irregular   spaces ,already spaced punctuation .",
]
SEEDS = [0, 1, 7]


def main():
    sys.path.insert(0, "/root/reference")
    stub = types.ModuleType("pycocotools")
    stub.coco = types.ModuleType("pycocotools.coco")
    stub.coco.COCO = object
    sys.modules.setdefault("pycocotools", stub)
    sys.modules.setdefault("pycocotools.coco", stub.coco)
    import importlib

    from PIL import Image
    pkg = types.ModuleType("modules")                             # the package WITHOUT its __init__ (which pulls matplotlib etc.)
    pkg.__path__ = ["/root/reference/modules"]
    sys.modules["modules"] = pkg
    ref_loader = importlib.import_module("modules.loader")        # the reference, unmodified
    real_open = Image.open
    Image.open = lambda *_a, **_k: Image.new("RGB", (8, 8))      # the image branch is not under test
    try:
        out = []
        for cap in CAPTIONS:
            for seed in SEEDS:
                ds = object.__new__(ref_loader.RedCapsDatasetLoader)
                ref_loader.DatasetLoader.__init__(ds)
                ds.images, ds.src_texts = ["unused.jpg"], [cap]
                torch.manual_seed(seed)
                _, src, tgt = ds[0]
                out.append({"caption": cap, "seed": seed, "source": src, "target": tgt})
    finally:
        Image.open = real_open
    json.dump(out, sys.stdout, indent=1, ensure_ascii=False)


if __name__ == "__main__":
    main()
