"""Host-side logic that needs no GPU: the operand-slot registry the fused optimizer uses to refresh bf16 operand copies, the
backward-phase switches of the data-parallel mode, and the JSON contract of bench.py's reference arm."""
import gc
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_operand_slot_registry_tracks_groups_and_dies_with_its_cache():
    from klab_multimodalmodel_b200 import functional as Fn
    ps = [torch.nn.Parameter(torch.randn(4, 8)) for _ in range(3)]
    solo = torch.nn.Parameter(torch.randn(5, 8))
    cache = Fn.OperandCache()
    e = cache._entry(ps, torch.bfloat16)                     # q|k|v group: one [12, 8] buffer (no conversion kernel on the CPU)
    e2 = cache._entry([solo], torch.bfloat16)
    assert e[1].shape == (12, 8) and e2[1].shape == (5, 8)
    for i, p in enumerate(ps):
        (entry, view), = Fn.operand_slots(p)
        assert entry is e and view.data_ptr() == e[1][4 * i:4 * i + 4].data_ptr() and view.shape == (4, 8)
    assert Fn.operand_slots(solo)[0][0] is e2
    # freshness: marking an entry stores the CURRENT versions; an in-place update makes it stale again
    Fn.mark_operands_fresh([e])
    assert e[0] == Fn.OperandCache._version(ps)
    with torch.no_grad():
        ps[1].add_(1.0)
    assert e[0] != Fn.OperandCache._version(ps)
    # an unrelated parameter has no slots; a direct (same dtype, 2-D, single) parameter is never copied
    assert Fn.operand_slots(torch.nn.Parameter(torch.zeros(2, 2))) == []
    assert cache.get([solo], torch.float32).data_ptr() == solo.data_ptr()
    # the registry holds the cache weakly: slots vanish with the model that owned them
    del cache, e, e2, entry, view
    gc.collect()
    assert all(Fn.operand_slots(p) == [] for p in ps + [solo])


def test_backward_phase_switches_are_scoped_to_backward():
    """functional._backward_phase: under data parallelism the library is told to distribute work dynamically / leave SMs to the
    collective only while a backward function runs, and the switches are restored even if it raises."""
    from klab_multimodalmodel_b200 import _lib as L
    from klab_multimodalmodel_b200 import functional as Fn
    calls = []

    class FakeLib:
        def klab_set_dynamic_sched(self, v):
            calls.append(("dyn", v))

        def klab_set_sm_reserve(self, v):
            calls.append(("res", v))

    real = L.lib
    L.lib = lambda: FakeLib()
    try:
        @Fn._backward_phase
        def bwd(ctx, g):
            calls.append(("run", g))
            if g < 0:
                raise ValueError("boom")
            return g + 1

        assert bwd(None, 1) == 2 and calls == [("run", 1)]                      # single GPU: nothing is switched
        calls.clear()
        Fn.DP_BACKWARD.update(on=True, reserve=16)
        assert bwd(None, 2) == 3
        assert calls == [("dyn", 1), ("res", 16), ("run", 2), ("dyn", 0), ("res", 0)]
        calls.clear()
        try:
            bwd(None, -1)
        except ValueError:
            pass
        assert calls[-2:] == [("dyn", 0), ("res", 0)]
    finally:
        L.lib = real
        Fn.DP_BACKWARD.update(on=False, reserve=0)
    # every block / loss / embedding backward is wrapped
    for cls in (Fn.T5BlockFn, Fn.SwinBlockFn, Fn.LMHeadLossFn, Fn.PatchEmbedFn, Fn.PatchMergeFn, Fn.DecoderEmbedFn, Fn.ConcatEmbeddingsFn):
        assert cls.backward.__name__ == "backward" and cls.backward.__closure__ is not None


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port on the host cores) prints ONE JSON line with the contract's keys."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] and d["metric"].startswith("train samples/sec")


def test_checkpoint_save_load_round_trip_in_reference_format(tmp_path):
    """N4: MyModel.save / .load write and read the reference's file format (/root/reference/models/model.py:30-42):
    {'transformer': state_dict[, 'image_model': state_dict]} with HF key names, so checkpoints are interchangeable."""
    import types

    from klab_multimodalmodel_b200.modeling import Swinv2Config, T5Config, init_swin_, init_t5_
    from klab_multimodalmodel_b200.models.model import MyModel
    tcfg = T5Config(vocab_size=64, d_model=128, d_ff=64, num_layers=1, num_heads=2)
    scfg = Swinv2Config(image_size=32, embed_dim=32, depths=(1, 1, 1), num_heads=(1, 2, 4), window_size=4, pretrained_window_sizes=(0, 0, 0))

    def make(train_swin, seed):
        args = types.SimpleNamespace(result_dir=str(tmp_path), language_model_name=tcfg, image_model_name=scfg,
                                     image_model_train=train_swin, transformer_model_name=tcfg)
        m = MyModel(args)
        init_t5_(m.transformer, seed=seed)
        init_swin_(m.image_model, seed=seed + 1)
        return m

    for train_swin in (True, False):
        a, b = make(train_swin, 10), make(train_swin, 20)
        a.save(result_name="ckpt.pth")
        blob = torch.load(os.path.join(str(tmp_path), "ckpt.pth"))
        assert set(blob) == ({"transformer", "image_model"} if train_swin else {"transformer"})
        assert "shared.weight" in blob["transformer"] and "lm_head.weight" in blob["transformer"]
        assert "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight" in blob["transformer"]
        if train_swin:
            assert "encoder.layers.0.blocks.0.attention.self.continuous_position_bias_mlp.0.weight" in blob["image_model"]
        b.load(result_name="ckpt.pth")
        for (k, pa), (_, pb) in zip(a.transformer.state_dict().items(), b.transformer.state_dict().items()):
            assert torch.equal(pa, pb), k
        same_swin = all(torch.equal(x, y) for x, y in zip(a.image_model.state_dict().values(), b.image_model.state_dict().values()))
        assert same_swin == train_swin                      # the image model travels only when it is trained (model.py:33-34,41-42)
