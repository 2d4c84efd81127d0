"""N1: fused multi-tensor Adam -- the optimizer step that follows the hot path every iteration
(/root/reference/train.py:28 `torch.optim.Adam(model.module.transformer.parameters(), lr=args.lr)`, :66 `optimizer.step()`).

Drop-in for `torch.optim.Adam` (same constructor arguments for the features the reference uses: lr, betas, eps, weight_decay;
same state keys `step`, `exp_avg`, `exp_avg_sq`, so `state_dict()` round-trips with torch's and LR schedulers work unchanged).
One CUDA launch (csrc/optim.cu) updates every parameter tensor instead of torch's ~250 multi_tensor_apply launches:
the step is pure HBM traffic (28 B per parameter) and runs at memory speed.  No CPU path: parameters must live on an
sm_100 device.
"""
from __future__ import annotations

import os
import weakref

import torch

from . import _lib as L

# The update of step i only has to be complete before the TRAINABLE transformer runs in step i + 1; the image tower and the frozen
# text tower that open every forward (/root/reference/models/model.py:20-22) do not read what Adam writes (train.py:28: the
# optimizer owns model.transformer only).  The fused kernel is pure HBM traffic (3.4 ms for T5-large), so it is issued on a side
# stream and overlaps the next step's towers; whoever is about to touch the transformer's weights on the compute stream calls
# `wait_pending_updates()` first (MyModel.forward before the transformer, T5ForConditionalGeneration.state_dict / load_state_dict,
# greedy decoding).  KLAB_ADAM_OVERLAP=0 keeps everything on the current stream.
OVERLAP = os.environ.get("KLAB_ADAM_OVERLAP", "1") != "0"
_SIDE: dict = {}
_DONE: dict = {}
# Parameters whose owner has declared that it waits (wait_pending_updates) before it reads them: only a step made of such
# parameters alone may leave the compute stream.  An optimizer built over anything else -- say over model.parameters(), image
# tower included, whose forward runs BEFORE the wait -- stays on the current stream, as torch.optim.Adam would.
_OVERLAP_SAFE: dict = {}


def allow_overlap(params) -> None:
    """Called by a model for the parameters it only reads after `wait_pending_updates()` (MyModel: the trainable transformer)."""
    for p in params:
        _OVERLAP_SAFE[id(p)] = weakref.ref(p)


def _overlap_ok(params) -> bool:
    for p in params:
        r = _OVERLAP_SAFE.get(id(p))
        if r is None or r() is not p:
            return False
    return True


def wait_pending_updates(device=None) -> None:
    """Make the current stream wait (on the device, not the host) for optimizer steps still running on the side stream."""
    if not _DONE:
        return
    idx = torch.cuda.current_device() if device is None or device.index is None else device.index
    ev = _DONE.pop(idx, None)
    if ev is not None:
        torch.cuda.current_stream(idx).wait_event(ev)


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False, *, maximize=False,
                 grad_scale=1.0, overlap=True):
        if amsgrad or maximize:
            raise NotImplementedError("klab Adam: amsgrad / maximize are not used by the reference and not implemented")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or not 0.0 <= weight_decay:
            raise ValueError("klab Adam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                                      grad_scale=grad_scale))
        self._tables: dict = {}
        self.overlap = overlap              # run the update on a side stream (see wait_pending_updates); False: on the current stream

    def _table(self, key, items):
        """items: list of (p, g, m, v, w16 pointer or 0).  Device tables are rebuilt only when a pointer changed (with CUDA graphs the gradient
        buffers are static, so this happens once).  Rebuilds go through two alternating pinned staging buffers and an event, so
        they never synchronise the device."""
        sig = tuple((p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), w) for p, g, m, v, w in items)
        ent = self._tables.get(key)
        if ent is not None and ent["sig"] == sig:
            return ent["table"], ent["blockmap"], ent["nblocks"]
        chunk = int(L.lib().klab_adam_chunk_elems())
        dev = items[0][0].device
        shape_sig = tuple(it[0].numel() for it in items)
        if ent is None or ent["shape_sig"] != shape_sig:
            bm = []
            for ti, n in enumerate(shape_sig):
                bm.extend([ti, c] for c in range((n + chunk - 1) // chunk))
            ent = {"shape_sig": shape_sig, "nblocks": len(bm),
                   "blockmap": torch.tensor(bm, dtype=torch.int32).to(dev),
                   "table": torch.empty(len(items), 6, dtype=torch.int64, device=dev),
                   "host": [torch.empty(len(items), 6, dtype=torch.int64).pin_memory() for _ in range(2)],
                   "events": [None, None], "flip": 0}
            self._tables[key] = ent
        i = ent["flip"]
        ent["flip"] = 1 - i
        if ent["events"][i] is not None:
            ent["events"][i].synchronize()                     # the copy that last used this staging buffer (two rebuilds ago)
        host = ent["host"][i]
        host.copy_(torch.tensor([[p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), w, p.numel()] for p, g, m, v, w in items],
                                dtype=torch.int64))
        ent["table"].copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        ent["events"][i] = ev
        ent["sig"] = sig
        self.rebuilds = getattr(self, "rebuilds", 0) + 1
        return ent["table"], ent["blockmap"], ent["nblocks"]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        from .functional import mark_operands_fresh, operand_slots
        for gi, group in enumerate(self.param_groups):
            by_step: dict = {}
            written: dict = {}                          # id(entry) -> [entry, parameter ids whose bf16 copy this step rewrites]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise RuntimeError("klab Adam: fp32 CUDA parameters and gradients only (there is no CPU path)")
                if p.grad.is_sparse:
                    raise RuntimeError("klab Adam does not support sparse gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0                      # python int (torch.optim.Adam.__setstate__ accepts it on load_state_dict)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("klab Adam: parameters and gradients must be contiguous")
                t = int(st["step"]) + 1
                st["step"] = t
                w16 = 0                                 # bf16 operand copy of p (what the tensor cores read): refreshed in the same pass
                for e, view in operand_slots(p):
                    if view.dtype == torch.bfloat16 and view.is_contiguous() and view.numel() == p.numel():
                        w16 = view.data_ptr()
                        written.setdefault(id(e), [e, set()])[1].add(id(p))
                        break
                by_step.setdefault(t, []).append((p, p.grad, st["exp_avg"], st["exp_avg_sq"], w16))
            b1, b2 = group["betas"]
            for t, items in by_step.items():
                table, blockmap, nblocks = self._table((gi, len(by_step) > 1 and t), items)
                dev = items[0][0].device
                main = torch.cuda.current_stream(dev)
                if OVERLAP and self.overlap and _overlap_ok(it[0] for it in items):
                    side = _SIDE.get(dev.index)
                    if side is None:
                        side = _SIDE[dev.index] = torch.cuda.Stream(dev)
                    wait_pending_updates(dev)                      # (an earlier update nobody waited for: keep the steps ordered)
                    side.wait_stream(main)                         # gradients (and the device tables) are ready
                    for it in items:
                        it[1].record_stream(side)                  # zero_grad(set_to_none) may free the gradient right after this call
                    table.record_stream(side)
                    run_on = side
                else:
                    run_on = main
                L.check(L.lib().klab_adam_step(run_on.cuda_stream, table.data_ptr(), blockmap.data_ptr(), nblocks, float(group["lr"]), float(b1),
                                               float(b2), float(group["eps"]), float(group["weight_decay"]), t,
                                               float(group.get("grad_scale", 1.0))))
                if run_on is not main:
                    ev = torch.cuda.Event()
                    ev.record(run_on)
                    _DONE[dev.index] = ev
                # the kernel wrote the masters through raw pointers: tell autograd (and every operand cache) that they changed
                torch.autograd.graph.increment_version([it[0] for it in items])
            mark_operands_fresh([e for e, ids in written.values() if ids == set(e[5])])
        return loss
