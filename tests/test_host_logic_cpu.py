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
    # kind "reference": the reference's own engine (transformers) ran; "port": the oracle restatement stood in (no transformers)
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
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


def test_from_pretrained_resolution(tmp_path, monkeypatch):
    """/root/reference/models/model.py:14-17 calls from_pretrained(name).  A bare name resolves through the local HF cache; a
    task-head checkpoint ('swinv2.' prefix + classifier, what microsoft/swinv2-* publishes) loads into the base model; a name with
    no weights on disk raises (opt-in: KLAB_ALLOW_RANDOM_INIT=1 warns); a directory with a config but no weights raises."""
    import json

    import pytest
    from safetensors.torch import save_file

    from klab_multimodalmodel_b200.modeling import (Swinv2Config, Swinv2Model, T5Config, T5EncoderModel, T5ForConditionalGeneration,
                                                      config_dict, init_swin_, init_t5_)
    scfg = Swinv2Config(image_size=32, embed_dim=32, depths=(1, 1), num_heads=(1, 2), window_size=4, pretrained_window_sizes=(0, 0))
    src = Swinv2Model(scfg)
    init_swin_(src, seed=5)
    # (1) HF-style cache layout, keys prefixed as Swinv2ForImageClassification writes them
    monkeypatch.setenv("HF_HOME", str(tmp_path / "hf"))
    for var in ("HF_HUB_CACHE", "HUGGINGFACE_HUB_CACHE", "TRANSFORMERS_CACHE", "KLAB_ALLOW_RANDOM_INIT"):
        monkeypatch.delenv(var, raising=False)
    snap = tmp_path / "hf" / "hub" / "models--microsoft--swinv2-base-patch4-window8-256" / "snapshots" / "abc123"
    snap.mkdir(parents=True)
    (snap / "config.json").write_text(json.dumps(config_dict(scfg)))
    sd = {"swinv2." + k: v.detach().clone() for k, v in src.state_dict().items()}
    sd["classifier.weight"] = torch.zeros(10, 64)
    sd["classifier.bias"] = torch.zeros(10)
    save_file(sd, str(snap / "model.safetensors"))
    with pytest.warns(RuntimeWarning, match="classifier"):
        got = Swinv2Model.from_pretrained("microsoft/swinv2-base-patch4-window8-256")
    assert got.config.embed_dim == 32 and not got.training
    for (k, a), (_, b) in zip(src.state_dict().items(), got.state_dict().items()):
        assert torch.equal(a, b), k
    # (2) a T5ForConditionalGeneration checkpoint under the legacy name feeds both the frozen encoder and the trainable model
    tcfg = T5Config(vocab_size=64, d_model=32, d_ff=64, num_layers=1, num_heads=1, d_kv=32)
    t5 = T5ForConditionalGeneration(tcfg)
    init_t5_(t5, seed=6)
    snap5 = tmp_path / "hf" / "hub" / "models--google-t5--t5-small" / "snapshots" / "r1"
    snap5.mkdir(parents=True)
    (snap5 / "config.json").write_text(json.dumps(config_dict(tcfg)))
    torch.save(t5.state_dict(), str(snap5 / "pytorch_model.bin"))
    full = T5ForConditionalGeneration.from_pretrained("t5-small")
    assert all(torch.equal(a, b) for a, b in zip(t5.state_dict().values(), full.state_dict().values()))
    with pytest.warns(RuntimeWarning, match="decoder"):
        enc = T5EncoderModel.from_pretrained("t5-small")
    assert torch.equal(enc.shared.weight, t5.shared.weight)
    assert torch.equal(enc.encoder.block[0].layer[0].SelfAttention.q.weight, t5.encoder.block[0].layer[0].SelfAttention.q.weight)
    # (3) known name, nothing on disk: the reference would fail; so does the drop-in unless random init is opted into
    with pytest.raises(FileNotFoundError, match="KLAB_ALLOW_RANDOM_INIT"):
        T5ForConditionalGeneration.from_pretrained("t5-base")
    monkeypatch.setenv("KLAB_ALLOW_RANDOM_INIT", "1")
    with pytest.warns(RuntimeWarning, match="RANDOMLY INITIALISED"):
        m = T5EncoderModel.from_pretrained("t5-base", num_layers=1)
    assert m.config.d_model == 768
    # (4) config without weights
    bare = tmp_path / "bare"
    bare.mkdir()
    (bare / "config.json").write_text(json.dumps(config_dict(tcfg)))
    with pytest.raises(FileNotFoundError, match="model.safetensors"):
        T5EncoderModel.from_pretrained(str(bare))
    with pytest.raises(FileNotFoundError):
        T5EncoderModel.from_pretrained("no-such-model")


def test_true_resume_round_trip(tmp_path):
    """N4, optional half: weights + optimizer + scheduler + progress in one file that the reference's own `load` still reads
    (extra keys only); restoring it continues the optimizer state exactly."""
    import types

    from klab_multimodalmodel_b200.modeling import Swinv2Config, T5Config, init_t5_
    from klab_multimodalmodel_b200.models.model import MyModel
    tcfg = T5Config(vocab_size=64, d_model=128, d_ff=64, num_layers=1, num_heads=2)
    scfg = Swinv2Config(image_size=32, embed_dim=32, depths=(1, 1, 1), num_heads=(1, 2, 4), window_size=4, pretrained_window_sizes=(0, 0, 0))
    args = types.SimpleNamespace(result_dir=str(tmp_path), language_model_name=tcfg, image_model_name=scfg, image_model_train=True,
                                 transformer_model_name=tcfg)
    a, b = MyModel(args), MyModel(args)
    init_t5_(a.transformer, seed=1)
    init_t5_(b.transformer, seed=2)
    opt_a = torch.optim.Adam(a.transformer.parameters(), lr=1e-3)
    sch_a = torch.optim.lr_scheduler.StepLR(opt_a, step_size=10, gamma=0.1)
    for p in a.transformer.parameters():
        p.grad = torch.full_like(p, 0.01)
    opt_a.step()
    sch_a.step()
    a.save_state("last.pth", optimizer=opt_a, scheduler=sch_a, epoch=3, step=1234)
    blob = torch.load(os.path.join(str(tmp_path), "last.pth"))
    assert {"transformer", "image_model", "optimizer", "scheduler", "progress"} <= set(blob)
    b.load("last.pth")                                              # the reference-format reader ignores the extra keys
    assert all(torch.equal(x, y) for x, y in zip(a.transformer.state_dict().values(), b.transformer.state_dict().values()))
    opt_b = torch.optim.Adam(b.transformer.parameters(), lr=1e-3)
    sch_b = torch.optim.lr_scheduler.StepLR(opt_b, step_size=10, gamma=0.1)
    prog = b.load_state("last.pth", optimizer=opt_b, scheduler=sch_b)
    assert prog == {"epoch": 3, "step": 1234} and sch_b.last_epoch == sch_a.last_epoch
    for pa, pb in zip(a.transformer.parameters(), b.transformer.parameters()):
        assert torch.equal(opt_a.state[pa]["exp_avg"], opt_b.state[pb]["exp_avg"]) and opt_b.state[pb]["step"] == opt_a.state[pa]["step"]
    a.save("weights_only.pth")                                      # a weights-only file (the reference's writer) resumes with empty progress
    assert b.load_state("weights_only.pth") == {"epoch": None, "step": None}


def test_optimizer_side_stream_registry():
    """optim.allow_overlap / _overlap_ok: the fused Adam may leave the compute stream only when EVERY parameter of its step was
    declared by its owner (MyModel declares the trainable transformer, whose weights forward() reads after wait_pending_updates());
    a stale id of a freed parameter must not count."""
    import gc

    from klab_multimodalmodel_b200 import optim as KO
    ps = [torch.nn.Parameter(torch.zeros(3)) for _ in range(3)]
    assert not KO._overlap_ok(ps)
    KO.allow_overlap(ps[:2])
    assert KO._overlap_ok(ps[:2]) and not KO._overlap_ok(ps)
    KO.allow_overlap(ps)
    assert KO._overlap_ok(ps) and KO._overlap_ok(iter(ps))
    stale = id(ps[2])
    del ps[2]
    gc.collect()
    assert KO._OVERLAP_SAFE[stale]() is None                  # the weak reference died with the parameter
    impostor = torch.nn.Parameter(torch.zeros(3))
    KO._OVERLAP_SAFE[id(impostor)] = KO._OVERLAP_SAFE[stale]  # what an id reused by a new tensor would find
    assert not KO._overlap_ok([impostor])
    KO.wait_pending_updates(None)                              # nothing pending: a no-op without a GPU
    for k in [id(p) for p in ps] + [stale, id(impostor)]:
        KO._OVERLAP_SAFE.pop(k, None)
