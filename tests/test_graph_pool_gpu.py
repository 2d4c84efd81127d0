"""CUDA-graph region pool under variable sequence lengths (B200 only).  The reference pads every batch to its longest sequence
(/root/reference/train.py:56-57), so training walks through many (L_src, L_tgt) signatures; every signature pins a private
pool of activations and static gradient buffers.  The pool must stay within its byte budget by evicting least-recently-used
signatures, memory must plateau, and results must not depend on what was evicted."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.caption_model import seeded_inputs  # noqa: E402
from tests.golden.make_golden import EXTRA_CASES  # noqa: E402
from tests.test_step_parity_gpu import build  # noqa: E402


def test_graph_pool_is_bounded_and_memory_plateaus():
    from klab_multimodalmodel_b200.graphs import POOL
    if not POOL.enabled:
        pytest.skip("CUDA graphs disabled")
    case = EXTRA_CASES["mid"]
    model, sds, swin, t5 = build(case, "bf16", style="hf")
    combos = [(4 + 2 * (i % 6), 6 + 3 * (i // 6)) for i in range(30)]          # 30 distinct (L_src, L_tgt) signatures
    old_budget = POOL.budget_bytes
    POOL.clear()
    first_loss, mem = {}, []
    try:
        budget = None
        for n, (ls, lt) in enumerate(combos):
            px, src, tgt = [t.cuda() for t in seeded_inputs(case["batch"], swin, t5.vocab_size, ls, lt, seed=50 + n)]
            for rep in range(3):                                              # eager warm-up, capture, replay
                for p in model.parameters():
                    p.grad = None
                loss = model({"pixel_values": px}, {"input_ids": src}, {"input_ids": tgt})
                loss.backward()
                if rep == 0:
                    first_loss[n] = loss.item()
                else:
                    assert abs(loss.item() - first_loss[n]) <= 1e-3 * abs(first_loss[n])       # graph replay == eager
            if n == 2:                                                        # budget: what three signatures pin
                budget = POOL.budget_bytes = max(POOL.pinned_bytes, 1)
            torch.cuda.synchronize()
            mem.append(torch.cuda.memory_allocated())
        assert POOL.evictions > 0, "nothing was evicted although 30 signatures went through a 3-signature budget"
        one_sig = budget / 3
        assert POOL.pinned_bytes <= budget + 3.5 * one_sig, (POOL.pinned_bytes, budget)    # current + previous step are protected
        assert max(mem[20:]) <= 1.25 * max(mem[6:14]), (mem[6:14], mem[20:])              # plateau, not growth
        # an evicted signature comes back (eager once, then captured again) with the same result
        px, src, tgt = [t.cuda() for t in seeded_inputs(case["batch"], swin, t5.vocab_size, combos[0][0], combos[0][1], seed=50)]
        for _ in range(3):
            for p in model.parameters():
                p.grad = None
            loss = model({"pixel_values": px}, {"input_ids": src}, {"input_ids": tgt})
            loss.backward()
            assert abs(loss.item() - first_loss[0]) <= 1e-3 * abs(first_loss[0])
    finally:
        POOL.budget_bytes = old_budget
        POOL.clear()
