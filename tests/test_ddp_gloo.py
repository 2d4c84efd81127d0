"""World-size-2 data-parallel plumbing on the CPU (gloo): the parameter trees the drop-in exposes to DistributedDataParallel
(/root/reference/train.py:26) -- tied embeddings, frozen text encoder, no buffers -- broadcast from rank 0 and average their
gradients across ranks; bench.py's rank-sharded synthetic data and max-over-ranks timing behave.  (The arithmetic itself
needs a B200; `bench.py --gpus N` covers it on the device.)"""
import os
import socket
import sys
import types

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench
        from klab_multimodalmodel_b200.modeling import Swinv2Config, T5Config, init_swin_, init_t5_
        from klab_multimodalmodel_b200.models.model import MyModel
        from torch.nn.parallel import DistributedDataParallel as DDP
        tcfg = T5Config(vocab_size=64, d_model=128, d_ff=64, num_layers=1, num_heads=2)
        scfg = Swinv2Config(image_size=32, embed_dim=32, depths=(1, 1, 1), num_heads=(1, 2, 4), window_size=4, pretrained_window_sizes=(0, 0, 0))
        args = types.SimpleNamespace(result_dir="/tmp", language_model_name=tcfg, image_model_name=scfg, image_model_train=True,
                                     transformer_model_name=tcfg)
        model = MyModel(args)
        init_t5_(model.transformer, seed=100 + rank)          # ranks start DIFFERENT; DDP must broadcast rank 0's values
        init_swin_(model.image_model, seed=200 + rank)
        assert not list(model.buffers()), "the drop-in must not register buffers (DDP would broadcast them every forward)"
        ddp = DDP(model)
        ref = [torch.zeros_like(p) for p in model.parameters()]
        for r, p in zip(ref, model.parameters()):
            r.copy_(p.detach())
            dist.broadcast(r, src=0)
            assert torch.equal(r, p.detach()), "parameters were not broadcast from rank 0"
        # the module took the gradient exchange over: DDP's own reducer keeps exactly one trainable tensor
        assert model._klab_reducer is not None
        assert sum(p.requires_grad for p in ddp._module_parameters) == 1
        assert "transformer.decoder.final_layer_norm.weight" not in ddp.parameters_to_ignore
        trainable = [p for p in ddp.parameters() if p.requires_grad]
        assert len(model._klab_reducer.params) == len(trainable) - 1
        assert all(not p.requires_grad for p in model.language_model.parameters())
        assert len({id(p) for p in model.transformer.parameters()}) == len(list(model.transformer.parameters()))   # tied weight once
        # gradients: a rank-dependent surrogate loss through DDP's hooks -> every rank ends with the mean over ranks
        with ddp.no_sync():
            pass
        out = sum((p * (rank + 1.0)).sum() for p in trainable)
        ddp.reducer.prepare_for_backward([])                  # we bypass ddp.forward (no CPU arithmetic path): arm the reducer
        out.backward()
        for p in trainable:
            assert p.grad is not None and torch.allclose(p.grad, torch.full_like(p.grad, (1.0 + world) / 2.0)), "gradients not averaged"
        assert model._klab_reducer.buckets_last_backward >= 1 and not model._klab_reducer._works
        # second micro-step without zero_grad (the reference accumulates and all-reduces every time, train.py:61-67): the
        # accumulated value is averaged again, which leaves the already-averaged part unchanged
        out = sum((p * (10.0 * (rank + 1.0))).sum() for p in trainable)
        ddp.reducer.prepare_for_backward([])
        out.backward()
        for p in trainable:
            assert torch.allclose(p.grad, torch.full_like(p.grad, 11.0 * (1.0 + world) / 2.0)), "accumulated gradients not averaged"
        # no_sync: local accumulation only
        for p in trainable:
            p.grad = None
        with model._klab_reducer.no_sync(), ddp.no_sync():
            sum((p * (rank + 1.0)).sum() for p in trainable).backward()
        assert all(torch.allclose(p.grad, torch.full_like(p.grad, rank + 1.0)) for p in trainable)
        # DDP's own no_sync() is honoured too: every forward tells the reducer whether its backward synchronises
        for p in trainable:
            p.grad = None
        with ddp.no_sync():
            model._klab_reducer.begin_step(ddp)
            sum((p * (rank + 1.0)).sum() for p in trainable).backward()
        assert all(torch.allclose(p.grad, torch.full_like(p.grad, rank + 1.0)) for p in trainable)
        model._klab_reducer.begin_step(ddp)
        assert model._klab_reducer.sync_next_backward
        # small buckets: several grouped all-reduces per backward, same result
        model._klab_reducer.bucket_bytes = 64 << 10
        for p in trainable:
            p.grad = None
        ddp.reducer.prepare_for_backward([])
        sum((p * (rank + 1.0)).sum() for p in trainable).backward()
        assert model._klab_reducer.buckets_last_backward > 1
        assert all(torch.allclose(p.grad, torch.full_like(p.grad, (1.0 + world) / 2.0)) for p in trainable)
        # a backward that died half way leaves no stale state behind: begin_step() (called by every forward) clears it
        red = model._klab_reducer
        red._in_backward, red._pending = True, [trainable[0].grad]
        red.begin_step()
        assert not red._in_backward and not red._pending and not red._works
        for p in trainable:
            p.grad = None
        ddp.reducer.prepare_for_backward([])
        sum((p * (rank + 1.0)).sum() for p in trainable).backward()
        assert all(torch.allclose(p.grad, torch.full_like(p.grad, (1.0 + world) / 2.0)) for p in trainable)
        # flat per-block gradient buffers (functional._flat_grads): the gradients of a block are slices of one buffer, which is
        # reduced as ONE tensor once its last parameter has reported; a parameter that accumulated elsewhere is reduced on its own
        from klab_multimodalmodel_b200 import functional as Fn

        class FlatBlock(torch.autograd.Function):
            @staticmethod
            def forward(ctx, scale, *ps):
                ctx.ps, ctx.scale = ps, scale
                return torch.zeros((), requires_grad=True) + sum(p.detach().sum() for p in ps) * 0

            @staticmethod
            def backward(ctx, g):
                flat, views = Fn._flat_grads([tuple(p.shape) for p in ctx.ps], "cpu")
                flat.zero_()
                for v in views:
                    v.fill_(ctx.scale)
                Fn._publish_flat(ctx.ps, flat, all(p.grad is None for p in ctx.ps))
                return (None,) + tuple(v.detach() for v in views)

        model._klab_reducer.bucket_bytes = 1 << 30
        group_ps = trainable[:5]
        for p in trainable:
            p.grad = None
        ddp.reducer.prepare_for_backward([])
        FlatBlock.apply(rank + 1.0, *group_ps).backward()
        assert model._klab_reducer.flats_last_backward == 1, model._klab_reducer.flats_last_backward
        assert all(torch.allclose(p.grad, torch.full_like(p.grad, (1.0 + world) / 2.0)) for p in group_ps)
        # second backward without zeroing: autograd adds into the existing .grad, the flat buffer is NOT what gets reduced
        ddp.reducer.prepare_for_backward([])
        FlatBlock.apply(10.0 * (rank + 1.0), *group_ps).backward()
        assert model._klab_reducer.flats_last_backward == 0
        assert all(torch.allclose(p.grad, torch.full_like(p.grad, 11.0 * (1.0 + world) / 2.0)) for p in group_ps)
        # bench helpers: per-rank shards differ; timing is the max over ranks
        w = dict(bench.WORKLOADS["tiny"])
        px, src, tgt = bench.synth_batch(w, 512, 1234 + rank, pin=False)
        gathered = [torch.zeros_like(src) for _ in range(world)]
        dist.all_gather(gathered, src)
        assert not torch.equal(gathered[0], gathered[1])
        t = torch.tensor([10.0 * (rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == 10.0 * world
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_ddp_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
