"""D1: data-parallel gradient averaging, overlapped with backward, without bucket copies.

The reference wraps the model in `DistributedDataParallel` (/root/reference/train.py:26).  DDP's reducer copies every gradient
into a flat 25 MiB bucket, all-reduces the bucket and copies it back: for this model that is ~2 000 extra small kernels per
step (980 parameter tensors, 3.3 GB in and out) on the stream the backward pass runs on.  The drop-in keeps the `DDP(model)`
line working unchanged but takes the gradient exchange over itself:

  * `MyModel._ddp_params_and_buffers_to_ignore` (the attribute DDP's constructor reads) names every trainable parameter
    except one small one, so DDP builds its reducer for that one tensor only (it refuses a module without any);
    reading the attribute is also what arms this reducer and broadcasts the ignored parameters from rank 0, as DDP's
    constructor would have done -- a model that is never wrapped in DDP never averages anything;
  * a post-accumulate-grad hook per parameter collects `p.grad` tensors as backward produces them (block by block: every
    block is one autograd node) and, once `bucket_bytes` have accumulated, issues ONE grouped NCCL all-reduce (ncclGroupStart /
    End over the tensors, ReduceOp.AVG) IN PLACE on the gradients, asynchronously on the process group's stream;
  * a callback queued on the autograd engine at the first hook waits (stream-wise, not host-wise) for all outstanding
    all-reduces at the end of backward, so `optimizer.step()` sees averaged gradients exactly as with DDP.

Semantics match DDP's: every backward all-reduces (the reference never uses no_sync, train.py:61-67), the value reduced is
the ACCUMULATED p.grad, the result is the mean over ranks.  `no_sync()` is provided for completeness.
"""
from __future__ import annotations

import contextlib
import os

import torch
import torch.distributed as dist


_HAS_COALESCING = hasattr(dist, "_coalescing_manager")          # private API of torch.distributed (present in 2.1 ... 2.11)


class GradReducer:
    def __init__(self, params, group=None, bucket_bytes: int | None = None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.params = [p for p in params if p.requires_grad]
        if bucket_bytes is None:
            bucket_bytes = int(float(os.environ.get("KLAB_BUCKET_MB", "64")) * (1 << 20))
        self.bucket_bytes = bucket_bytes
        self.enabled = True
        self.sync_next_backward = True          # DDP's require_backward_grad_sync at the time of the forward (model.no_sync())
        self.use_flat = os.environ.get("KLAB_FLAT_GRAD_REDUCE", "1") != "0"
        self._pending: list = []
        self._pending_bytes = 0
        self._open_groups: dict = {}            # flat gradient buffers some (not yet all) of whose parameters have reported
        self._flats = 0
        self.flats_last_backward = 0
        self._works: list = []
        self._in_backward = False
        self._native_avg = dist.get_backend(group) == "nccl"        # gloo has no AVG: SUM, then scale at the end of backward
        self._summed: list = []
        self.buckets_last_backward = 0
        self._count = 0
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    # ------------------------------------------------------------------------------------------
    def _on_grad(self, p):
        if not self.enabled or not self.sync_next_backward or p.grad is None or not dist.is_initialized():
            return                                # (no_sync(), or the process group is already torn down: local run)
        if not self._in_backward:
            self._in_backward = True
            self._count = 0
            torch.autograd.Variable._execution_engine.queue_callback(self._finalize)
        g = p.grad
        # Block-level functions produce all parameter gradients of a block as slices of ONE flat buffer (functional._flat_grads) and
        # say so on the parameters (`_klab_flat`): the whole buffer is reduced as a single tensor once the last of its parameters
        # has reported -- ~100 large all-reduces per step instead of ~1 000 small ones, whose per-operation latency (not bandwidth)
        # is what a grouped NCCL all-reduce pays for.
        group = getattr(p, "_klab_flat", None)
        if group is not None and self.use_flat:
            flat = group["flat"]
            if g.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr():
                p._klab_flat = None
                group["left"] -= 1
                if group["left"] > 0:
                    self._open_groups[id(group)] = group
                    return
                self._open_groups.pop(id(group), None)
                self._flats += 1
                g = flat
            else:
                p._klab_flat = None                  # the gradient was accumulated elsewhere: reduce it on its own
        self._pending.append(g)
        self._pending_bytes += g.numel() * g.element_size()
        if self._pending_bytes >= self.bucket_bytes:
            self._flush()

    def _flush(self):
        if not self._pending:
            return
        tensors, self._pending, self._pending_bytes = self._pending, [], 0
        op = dist.ReduceOp.AVG if self._native_avg else dist.ReduceOp.SUM
        dev = tensors[0].device if tensors[0].is_cuda else None
        if _HAS_COALESCING:
            with dist._coalescing_manager(self.group, dev, async_ops=True) as cm:          # one ncclGroupStart / End around the bucket
                for t in tensors:
                    dist.all_reduce(t, op=op, group=self.group)
            self._works.append(cm)
        else:                                             # torch without the (private) coalescing manager: one async collective per tensor
            self._works.extend(dist.all_reduce(t, op=op, group=self.group, async_op=True) for t in tensors)
        if not self._native_avg:
            self._summed.extend(tensors)
        self._count += 1

    def _finalize(self):
        try:
            for group in self._open_groups.values():          # a parameter of the block got no gradient this time: reduce what there is
                self._pending.append(group["flat"])
                self._flats += 1
            self._open_groups = {}
            self._flush()
            for w in self._works:
                w.wait()                     # CUDA: the current stream waits for the collective's stream; the host does not block
            if self._summed:
                torch._foreach_div_(self._summed, float(self.world))
        finally:
            self._works, self._summed = [], []
            self._in_backward = False
            self.buckets_last_backward = self._count
            self.flats_last_backward, self._flats = self._flats, 0

    def begin_step(self, ddp=None):
        """Called at the start of every forward (`ddp`: the DistributedDataParallel wrapper whose forward is running, if any --
        inside `with ddp_model.no_sync():` it has require_backward_grad_sync = False and this backward must not all-reduce): if the previous backward died half way (an exception inside a kernel wrapper,
        a KeyboardInterrupt), its end-of-backward callback never ran -- drop the stale bookkeeping instead of skipping the
        callback of every later backward.  A no-op in normal operation."""
        if self._in_backward or self._pending or self._works:
            for w in self._works:
                try:
                    w.wait()
                except Exception:                 # noqa: BLE001
                    pass
            self._pending, self._pending_bytes, self._works, self._summed = [], 0, [], []
            self._in_backward = False
        self._open_groups = {}
        self.sync_next_backward = True if ddp is None else bool(getattr(ddp, "require_backward_grad_sync", True))

    # ------------------------------------------------------------------------------------------
    @contextlib.contextmanager
    def no_sync(self):
        old, self.enabled = self.enabled, False
        try:
            yield
        finally:
            self.enabled = old

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []


def broadcast_from_rank0(tensors, group=None, bucket_bytes: int = 250 << 20):
    """What DDP's constructor does for the parameters it owns (`_sync_module_states`): rank 0's values everywhere."""
    tensors = [t.detach() for t in tensors]
    if not tensors:
        return
    pg = group or dist.distributed_c10d._get_default_group()
    try:
        dist._broadcast_coalesced(pg, tensors, bucket_bytes, 0)
    except Exception:                         # noqa: BLE001  (backend without the coalesced path)
        for t in tensors:
            dist.broadcast(t, src=0, group=group)
