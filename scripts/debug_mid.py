import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_step_parity_gpu import EXTRA_CASES, build, oracle_grads
from oracle.caption_model import seeded_inputs
case = EXTRA_CASES["mid"]
model, sds, swin, t5 = build(case, "bf16", style="hf")
px, src, tgt = seeded_inputs(case["batch"], swin, t5.vocab_size, case["l_src"], case["l_tgt"], ignore_tail=case["ignore_tail"])
loss = model({"pixel_values": px.cuda()}, {"input_ids": src.cuda()}, {"input_ids": tgt.cuda()})
loss.backward()
ref_loss, leaves = oracle_grads(case, sds, swin, t5, px, src, tgt)
ac_loss, ac_leaves = oracle_grads(case, sds, swin, t5, px, src, tgt, autocast=True)
print("loss", loss.item(), ref_loss, ac_loss)
for k, p in model.image_model.named_parameters():
    if "attention.self" not in k and "layers.2" not in k:
        continue
    ref = leaves[("image_model", k)].grad
    g = p.grad.detach().float().cpu()
    nref = max(ref.norm().item(), 1e-12)
    err = (g - ref).norm().item() / nref
    err_ac = (ac_leaves[("image_model", k)].grad.float() - ref).norm().item() / nref
    print(f"{k:80s} |ref| {nref:.3e} err {err:.3e} autocast {err_ac:.3e} ratio {err / max(err_ac, 1e-9):.2f}")
