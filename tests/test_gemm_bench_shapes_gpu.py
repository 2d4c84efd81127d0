"""Parity of the tensor-core GEMM at the sizes the benchmark actually runs (B200 only).

The signatures below are the ones that carry the most time in one training step of bench workload 2a (Swin-B/256 + T5-large,
64 samples; profiles/r01_gemm_signatures_final.txt): shape, operand majors AND fused epilogue.  Each one is run four ways --
library's own choice, forced one-CTA kernel, forced CTA-pair kernel (tcgen05.mma.cta_group::2), and, for the weight gradients,
forced split-K -- through the test-only pin `klab_gemm_set_force`, and compared with an fp32 `A @ B^T` (+ the same epilogue in
torch) at 2e-2 of the result's scale (bf16 operands, fp32 accumulation).  `klab_gemm_last_config` proves the pinned kernel ran.
Dropout epilogues are checked through a size-independent property: every output element is either 0 or the undropped value / keep.
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF, F32 = torch.bfloat16, torch.float32
# (M, N, K, a_mn, b_mn, out dtype, bias, act, residual, aux_in, aux_out, dropout p)      act: 0 none 1 relu 2 gelu 3 relu' 4 gelu'
SIGNATURES = [
    (6144, 4096, 1024, 0, 1, BF, False, 3, False, True, False, 0.1),     # T5 FFN dgrad with ReLU' + dropout (x24, 1.44 ms)
    (6144, 4096, 1024, 0, 0, BF, False, 1, False, False, False, 0.1),    # T5 FFN wi forward, ReLU + dropout
    (6144, 1024, 4096, 0, 0, BF, False, 0, True, False, False, 0.1),     # T5 FFN wo forward, dropout + residual
    (16384, 2048, 512, 0, 1, BF, False, 4, False, True, False, 0.0),     # Swin stage-3 fc2 dgrad with GELU'
    (6144, 1024, 4096, 0, 1, BF, False, 0, False, False, False, 0.0),    # T5 FFN wi dgrad
    (1024, 1024, 2048, 1, 1, F32, False, 0, False, False, False, 0.0),   # decoder / text-tower wgrad (x72)
    (1024, 4096, 6144, 1, 1, F32, False, 0, False, False, False, 0.0),   # T5 FFN wgrad
    (4096, 1024, 6144, 1, 1, F32, False, 0, False, False, False, 0.0),
    (3072, 1024, 6144, 1, 1, F32, False, 0, False, False, False, 0.0),   # encoder q|k|v wgrad
    (16384, 2048, 512, 0, 0, BF, True, 2, False, False, True, 0.0),      # Swin fc1 forward: bias + GELU + pre-activation copy
    (6144, 1024, 3072, 0, 1, BF, False, 0, False, False, False, 0.0),    # encoder q|k|v dgrad
    (2048, 3072, 1024, 0, 0, BF, False, 0, False, False, False, 0.0),    # decoder / text-tower q|k|v forward
    (2048, 1024, 1024, 0, 1, BF, False, 0, False, False, False, 0.0),    # decoder o / cross-q dgrad (x72)
    (2048, 1024, 4096, 0, 0, BF, False, 0, True, False, False, 0.1),
    (2048, 1024, 1024, 0, 0, BF, False, 0, True, False, False, 0.1),     # decoder o-projection, dropout + residual (x48)
    (512, 2048, 16384, 1, 1, F32, False, 0, False, False, False, 0.0),   # Swin fc wgrad, K = 16384 (split-K in every step)
    (2048, 512, 16384, 1, 1, F32, False, 0, False, False, False, 0.0),
    (1536, 512, 16384, 1, 1, F32, False, 0, False, False, False, 0.0),   # Swin q|k|v wgrad
    (16384, 512, 2048, 0, 1, BF, False, 0, True, False, False, 0.0),     # Swin fc1 dgrad + residual
    (16384, 1536, 512, 0, 0, BF, True, 0, False, False, False, 0.0),     # Swin q|k|v forward + bias
    (262144, 512, 128, 0, 1, BF, False, 4, False, True, False, 0.0),     # Swin stage-1 fc2 dgrad with GELU' (K = 128)
    (262144, 384, 128, 0, 0, BF, True, 0, False, False, False, 0.0),     # Swin stage-1 q|k|v forward
    (512, 512, 16384, 1, 1, F32, False, 0, False, False, False, 0.0),
    (2048, 4096, 1024, 0, 0, BF, False, 1, False, False, False, 0.1),
]
MODES = ["auto", "one_cta", "cta_pair", "split_k"]


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale).to(BF)


def _last_config(lib):
    bn, sp, c2 = C.c_int(0), C.c_int(0), C.c_int(0)
    lib.klab_gemm_last_config(C.byref(bn), C.byref(sp), C.byref(c2))
    return bn.value, sp.value, c2.value


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("sig", SIGNATURES, ids=lambda s: "M%dN%dK%d_%d%d_act%d%s" % (s[0], s[1], s[2], s[3], s[4], s[7], "_drop" if s[11] else ""))
def test_gemm_at_bench_signature(sig, mode):
    from klab_multimodalmodel_b200 import _lib as L
    from klab_multimodalmodel_b200 import ops as O
    M, N, K, a_mn, b_mn, od, has_bias, act, has_res, has_auxin, has_auxout, p = sig
    wgrad = a_mn and b_mn
    if mode == "split_k" and not wgrad:
        pytest.skip("split-K is used for weight gradients (fp32 output, linear epilogue) only")
    lib = L.lib()
    A = _rand((K, M) if a_mn else (M, K), 1)
    B = _rand((K, N) if b_mn else (N, K), 2, scale=K ** -0.5)
    kw = dict(a_mn=bool(a_mn), b_mn=bool(b_mn), out_dtype=od, act=act)
    bias = torch.randn(N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)) if has_bias else None
    res = _rand((M, N), 4) if has_res else None
    aux_in = _rand((M, N), 5) if has_auxin else None
    if has_bias:
        kw["bias"] = bias
    if has_res:
        kw["residual"] = res
    if has_auxin:
        kw["aux_in"] = aux_in
    force = {"auto": (-1, -1, -1), "one_cta": (0, -1, -1), "cta_pair": (1, -1, -1), "split_k": (-1, -1, 4)}[mode]
    lib.klab_gemm_set_force(*force)
    try:
        pre = torch.empty(M, N, dtype=od, device="cuda") if has_auxout else None
        out = O.gemm(A, B, M, N, K, aux_out=pre, **kw)
        bn, splits, cta2 = _last_config(lib)
        out_p = O.gemm(A, B, M, N, K, dropout_p=p, seed=1234, **kw) if p > 0 else None
    finally:
        lib.klab_gemm_set_force(-1, -1, -1)
    if mode == "one_cta":
        assert cta2 == 0
    if mode == "cta_pair":
        assert cta2 == 1, "the CTA-pair kernel was not used"
    if mode == "split_k":
        assert splits >= 2, "split-K was not used"
    A32 = A.float().t() if a_mn else A.float()
    B32 = B.float() if b_mn else B.float().t()
    base = A32 @ B32                                   # fp32 reference (torch matmul, TF32 off by default)
    if has_bias:
        base = base + bias
    ref = base
    if act == 1:
        ref = torch.relu(base)
    elif act == 2:
        ref = F.gelu(base)
    elif act == 3:
        ref = base * (aux_in.float() > 0)
    elif act == 4:
        x = aux_in.float()
        ref = base * (0.5 * (1 + torch.erf(x * 0.7071067811865476)) + x * torch.exp(-0.5 * x * x) * 0.3989422804014327)
    core = ref                                         # value the dropout mask multiplies
    if has_res:
        ref = ref + res.float()
    scale = ref.abs().max().item()
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-2 * scale + 1e-6, f"{mode} (bn={bn} splits={splits} cta2={cta2}): max err {err:.3e} vs scale {scale:.3e}"
    if has_auxout:
        assert (pre.float() - base).abs().max().item() <= 2e-2 * base.abs().max().item()
    if out_p is not None:
        # dropout: out_p = keep_mask * core / keep (+ residual); masks are a pure function of (seed, element index)
        d = out_p.float() - (res.float() if has_res else 0.0)
        keep = 1.0 - p
        tol = 2e-2 * scale + 1e-6
        kept = (d - core / keep).abs() <= tol
        dropped = d.abs() <= tol
        assert bool((kept | dropped).all()), "an element is neither dropped nor scaled by 1 / keep"
        sure = core.abs() > 4 * tol                    # elements large enough to tell the two cases apart
        frac = (dropped & sure).float().sum().item() / max(sure.float().sum().item(), 1.0)
        assert abs(frac - p) < 0.01, frac
