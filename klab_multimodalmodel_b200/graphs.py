"""CUDA-graph regions: the kernel sequence of one block (forward or backward) is captured once per shape signature and replayed
afterwards, so a training step costs ~170 graph launches on the host instead of ~4000 kernel launches.

The reference's step is launch-bound on the host once the kernels are fast (thousands of small launches per step); a tracing
compiler is not an option here, CUDA graphs are (the kernels are static given the shapes).  Design:

  * a region is `fn(*inputs, *consts) -> tuple of tensors`; `inputs` are activations / gradients, `consts` are parameters,
    prepared bf16 operands, LUTs and non-tensor arguments (addresses assumed stable, verified before every replay);
  * first call with a new key runs eagerly (warm-up: lazy attribute setup inside the library, and the GEMM's per-signature
    timing of its one-CTA / CTA-pair candidates, which must not happen under capture), the second call captures, later calls
    replay; a region may fork work onto a second stream (functional._Side: weight gradients) -- the fork and the join are
    captured as a parallel branch of the graph;
  * outputs (and every activation kept for backward) are static buffers owned by the graph's private memory pool; an input
    that is itself the output of another region is consumed in place, anything else is copied into a staging buffer first;
  * dropout masks come from a device-side counter (klab_seed_advance), so replays draw fresh masks with no host involvement;
  * hazards are handled conservatively: if a region's forward is called again before its backward ran (two forwards, one
    backward), or a parameter's .grad still aliases a static gradient buffer (gradient accumulation), the call falls back to
    eager execution or detaches the alias first -- results never depend on whether graphs are on.

KLAB_CUDA_GRAPHS=0 disables the mechanism (everything runs eagerly).
"""
from __future__ import annotations

import collections
import os

import torch


class _Region:
    __slots__ = ("graph", "outs", "cap_inputs", "staging", "const_ptrs", "out_ptrs", "calls", "n_kernels", "nbytes", "tick")

    def __init__(self):
        self.graph = None
        self.outs = None
        self.cap_inputs = None
        self.staging = None
        self.const_ptrs = None
        self.out_ptrs = ()
        self.calls = 0
        self.n_kernels = 0
        self.nbytes = 0                            # device memory this region pins (private pool + staging buffers)
        self.tick = 0                              # step in which the region was last used


def _default_budget_bytes() -> int:
    gb = os.environ.get("KLAB_GRAPH_POOL_GB")
    if gb is not None:
        return int(float(gb) * (1 << 30))
    try:
        if torch.cuda.is_available():
            return int(0.6 * torch.cuda.get_device_properties(torch.cuda.current_device()).total_memory)
    except Exception:                              # noqa: BLE001
        pass
    return 96 << 30


class GraphPool:
    """Captured regions live in an LRU map bounded by BYTES (KLAB_GRAPH_POOL_GB, default 60 % of the device memory): the
    reference pads every batch to its longest sequence (/root/reference/train.py:56-57, padding="longest"), so real training
    walks through dozens of (L_src, L_tgt) signatures, and every signature pins a private pool with the saved activations and
    static gradient buffers of every block.  When a new capture would exceed the budget, the least recently used regions that
    were not touched in the current or the previous step are dropped (their pools are returned to the allocator once autograd
    lets go of the tensors); if nothing can be dropped the call simply runs eagerly."""

    def __init__(self):
        self.enabled = os.environ.get("KLAB_CUDA_GRAPHS", "1") != "0"
        self.regions: "collections.OrderedDict" = collections.OrderedDict()
        self.static_ptrs: set = set()
        self.side_stream = None
        self.replays = 0
        self.captures = 0
        self.eager_calls = 0
        self.evictions = 0
        self.replayed_kernels = 0                  # kernels of libklab_b200 launched through graph replays
        self.max_regions = int(os.environ.get("KLAB_GRAPH_MAX_REGIONS", "4096"))
        self.budget_bytes = None                   # resolved lazily (needs the device)
        self.pinned_bytes = 0
        self.tick = 0

    def begin_step(self):
        """Called once per model forward: regions used in this or the previous step are never evicted."""
        self.tick += 1

    # ------------------------------------------------------------------------------------------
    def run(self, key, fn, inputs, consts, allow_graph=True):
        """Run fn(*inputs, *consts); returns (outputs tuple, used_graph)."""
        if not (self.enabled and allow_graph):
            self.eager_calls += 1
            return fn(*inputs, *consts), False
        reg = self.regions.get(key)
        if reg is None:
            if len(self.regions) >= self.max_regions and not self._evict(1 << 62, want_slots=1):
                self.eager_calls += 1                             # too many live signatures and none is old enough to go
                return fn(*inputs, *consts), False
            self.regions[key] = reg = _Region()
        else:
            self.regions.move_to_end(key)
        reg.calls += 1
        reg.tick = self.tick
        if reg.calls == 1:                                    # warm-up run
            self.eager_calls += 1
            return fn(*inputs, *consts), False
        if reg.graph is not None and not self._consts_match(reg, consts):
            self._drop(reg)
        if reg.graph is not None:
            for t, cap, stage in zip(inputs, reg.cap_inputs, reg.staging):
                if t is None:
                    continue
                if stage is not None:
                    if t.data_ptr() != stage.data_ptr():
                        stage.copy_(t)
                elif t.data_ptr() != cap.data_ptr():          # a static producer changed its buffer: capture again
                    self._drop(reg)
                    break
        if reg.graph is None:
            if self.budget_bytes is None:
                self.budget_bytes = _default_budget_bytes()
            if self.pinned_bytes >= self.budget_bytes and not self._evict(self.pinned_bytes - self.budget_bytes + 1):
                self.eager_calls += 1                             # over budget and everything live is in use: no new capture
                return fn(*inputs, *consts), False
            self._capture(reg, fn, inputs, consts)
        reg.graph.replay()
        self.replays += 1
        self.replayed_kernels += reg.n_kernels
        return reg.outs, True

    # ------------------------------------------------------------------------------------------
    def _consts_match(self, reg, consts) -> bool:
        for c, p in zip(consts, reg.const_ptrs):
            if torch.is_tensor(c):
                if c.data_ptr() != p:
                    return False
            elif c != p:
                return False
        return True

    def _drop(self, reg):
        for p in reg.out_ptrs:
            self.static_ptrs.discard(p)
        self.pinned_bytes -= reg.nbytes
        reg.nbytes = 0
        reg.graph = None
        reg.outs = None
        reg.cap_inputs = None
        reg.staging = None
        reg.const_ptrs = None
        reg.out_ptrs = ()

    def _evict(self, need_bytes, want_slots=0) -> bool:
        """Drop least-recently-used regions (not used in this or the previous step) until `need_bytes` are released (and
        `want_slots` map entries are free).  Returns False if nothing could be released."""
        freed, slots = 0, 0
        for key in list(self.regions):
            if freed >= need_bytes or (want_slots and slots >= want_slots):
                break
            reg = self.regions[key]
            if reg.tick >= self.tick - 1:
                break                                             # LRU order: everything after this one is in use as well
            freed += reg.nbytes
            slots += 1
            self._drop(reg)
            del self.regions[key]
            self.evictions += 1
        return freed > 0 or slots > 0

    def drop_keys(self, pred) -> int:
        """Drop every region whose key satisfies `pred` (the owner of the static buffers it bakes in is going away)."""
        n = 0
        for key in [k for k in self.regions if pred(k)]:
            self._drop(self.regions.pop(key))
            n += 1
        return n

    def clear(self):
        self.drop_keys(lambda k: True)

    def _capture(self, reg, fn, inputs, consts):
        dev = torch.cuda.current_device()
        r0, a0 = torch.cuda.memory_reserved(dev), torch.cuda.memory_allocated(dev)
        cap_inputs, staging = [], []
        for t in inputs:
            if t is None:
                cap_inputs.append(None)
                staging.append(None)
            elif t.untyped_storage().data_ptr() in self.static_ptrs and t.is_contiguous():
                cap_inputs.append(t)
                staging.append(None)
            else:
                buf = torch.empty_like(t, memory_format=torch.contiguous_format)
                buf.copy_(t)
                cap_inputs.append(buf)
                staging.append(buf)
        reg.const_ptrs = [c.data_ptr() if torch.is_tensor(c) else c for c in consts]
        if self.side_stream is None:
            self.side_stream = torch.cuda.Stream()
        s = self.side_stream
        cur = torch.cuda.current_stream()
        s.wait_stream(cur)
        from . import ops as _ops
        g = torch.cuda.CUDAGraph()
        n0 = _ops.launch_count()
        with torch.cuda.stream(s):
            g.capture_begin()
            try:
                outs = fn(*cap_inputs, *consts)
            finally:
                g.capture_end()
        cur.wait_stream(s)
        reg.n_kernels = _ops.launch_count() - n0
        reg.graph, reg.outs, reg.cap_inputs, reg.staging = g, outs, cap_inputs, staging
        reg.out_ptrs = tuple(o.untyped_storage().data_ptr() for o in outs if torch.is_tensor(o))
        self.static_ptrs.update(reg.out_ptrs)
        reg.nbytes = max(torch.cuda.memory_reserved(dev) - r0, torch.cuda.memory_allocated(dev) - a0, 0)
        self.pinned_bytes += reg.nbytes
        self.captures += 1

    def owns(self, t) -> bool:
        """True if `t` (or the buffer it is a view of) is a static output of some captured region."""
        return t is not None and t.untyped_storage().data_ptr() in self.static_ptrs

    def stats(self) -> dict:
        return {"regions": len(self.regions), "captures": self.captures, "replays": self.replays, "eager_calls": self.eager_calls,
                "evictions": self.evictions, "pinned_gb": round(self.pinned_bytes / (1 << 30), 2)}


POOL = GraphPool()
