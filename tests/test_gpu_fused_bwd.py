"""The fused persistent backward launch (csrc/bwd_fused.cuh) against the block loop it replaces and against float64:
square and row-sharded (rectangular, offset diagonal) problems, D = 256 / 512 / 1024, several block shapes, with and
without d logit_scale."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

def _setenv(**kw):
    """Plan knobs of the fused backward (mmg_tune); no arguments = back to the defaults."""
    from mmgclip_b200 import ops
    ops.set_tuning()
    if kw:
        ops.set_tuning(**kw)


def _problem(rows, cols, d, off, seed=3):
    from mmgclip_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.nn.functional.normalize(torch.randn(rows, d, device="cuda", generator=gen), dim=1)
    b = torch.nn.functional.normalize(torch.randn(cols, d, device="cuda", generator=gen), dim=1)
    b[off:off + rows] = torch.nn.functional.normalize(b[off:off + rows] + 0.7 * a, dim=1)
    ab, bb = ops.cast_bf16(a), ops.cast_bf16(b)
    s = torch.tensor(1 / 0.07, device="cuda")
    rs, cs, diag = ops.infonce_forward_raw(ab, bb, s, off, "bf16")
    return ops, a, b, ab, bb, s, rs, cs, diag


@pytest.mark.parametrize("rows,cols,d,off,cfg", [
    (256, 256, 256, 0, {}),
    (256, 1024, 256, 768, {}),
    (512, 512, 256, 0, {}),
    (768, 768, 1024, 0, {}),
    (1024, 2048, 512, 512, {}),
    (1024, 2048, 512, 1024, {"fused_rb": 256, "fused_cb": 512, "fused_nbuf": 3, "fused_ksl": 2}),
    (4096, 8192, 512, 4096, {}),
    (4096, 4096, 512, 0, {"fused_rb": 1024, "fused_cb": 1024, "fused_nbuf": 5, "fused_ksl": 4, "fused_ksl_t": 16}),
])
def test_fused_backward_matches_block_loop(rows, cols, d, off, cfg):
    ops, a, b, ab, bb, s, rs, cs, diag = _problem(rows, cols, d, off)
    one = torch.ones((), device="cuda")
    b32 = b[off:off + rows].contiguous()
    run = lambda: ops.infonce_backward_raw(ab, bb, s, rs, cs, one, 0.5 / cols, off, "bf16", a32=a, b32=b32, diag=diag)  # noqa: E731
    try:
        _setenv(fused=0)
        dA0, dB0, dl0 = run()
        torch.cuda.synchronize()
        _setenv(**cfg)
        n0 = ops._lib.load().mmg_kernel_launch_count()
        dA1, dB1, dl1 = run()
        torch.cuda.synchronize()
        launches = ops._lib.load().mmg_kernel_launch_count() - n0
    finally:
        _setenv()
    assert launches == 2, "prep + matching-pair init (one launch) + ONE fused launch expected"
    # same bf16 operands and coefficients; only the fp32 accumulation order differs
    assert rel_err(dA1.cpu(), dA0.cpu()) < 2e-4
    assert rel_err(dB1.cpu(), dB0.cpu()) < 2e-4
    assert abs(dl1.item() - dl0.item()) <= 2e-4 * max(abs(dl0.item()), 1e-3)


def test_fused_backward_vs_float64():
    rows = cols = 512
    d = 256
    ops, a, b, ab, bb, s, rs, cs, diag = _problem(rows, cols, d, 0, seed=9)
    one = torch.ones((), device="cuda")
    dA, dB, dls = ops.infonce_backward_raw(ab, bb, s, rs, cs, one, 0.5 / cols, 0, "bf16", a32=a, b32=b, diag=diag)
    torch.cuda.synchronize()
    A, Bm = a.double().cpu().numpy(), b.double().cpu().numpy()
    sv = float(s.item())
    cos = A @ Bm.T
    e = np.exp(sv * cos - sv)
    coef = sv * 0.5 / cols
    g = e * (coef / e.sum(1)[:, None] + coef / e.sum(0)[None, :])
    g[np.arange(rows), np.arange(rows)] -= 2 * coef
    # bf16 operand rounding: same bars as tests/test_gpu_parity.py (GRAD bf16 = 4e-3)
    assert rel_err(dA.cpu(), g @ Bm) < 4e-3
    assert rel_err(dB.cpu(), g.T @ A) < 4e-3
    assert abs(dls.item() - float((g * cos).sum())) < 4e-3 * max(1.0, abs(float((g * cos).sum())))


@pytest.mark.parametrize("world,rank,n_parts", [(2, 1, 1), (4, 2, 1), (8, 5, 1), (2, 0, 2), (4, 3, 2)])
def test_owner_distributed_column_gradients(world, rank, n_parts):
    """mmg_infonce_bwd_owners on one GPU: the column-side gradient is spread over `world` separate buffers (what the ranks'
    NVLink-mapped buffers are in the multi-GPU run), optionally in `n_parts` launches that each cover one part of every
    owner's columns; stitched together they equal the ordinary dB."""
    from mmgclip_b200 import ops
    rows, d = 512, 256
    cols = rows * world
    off = rank * rows
    ops_, a, b, ab, bb, s, rs, cs, diag = _problem(rows, cols, d, off, seed=21)
    one = torch.ones((), device="cuda")
    b32 = b[off:off + rows].contiguous()
    dA0, dB0, dl0 = ops.infonce_backward_raw(ab, bb, s, rs, cs, one, 0.5 / cols, off, "bf16", a32=a, b32=b32, diag=diag)
    rp = rows // n_parts
    # bufs[part][owner]: the other owners start from zero, this rank's own buffers are initialised by the call
    bufs = [[torch.zeros((rp, d), device="cuda") for _ in range(world)] for _ in range(n_parts)]
    events = []
    try:
        dA1, owns, dl1 = ops.infonce_backward_owners(
            ab, bb, s, rs, cs, one, 0.5 / cols, off,
            [(bufs[i][rank], [t.data_ptr() for t in bufs[i]]) for i in range(n_parts)],
            lambda: events.append("pre"), lambda: events.append("post"), lambda i: events.append(i), a32=a, b32=b32,
            diag=diag)
        torch.cuda.synchronize()
    finally:
        _setenv()
    assert events == ["pre"] + list(range(n_parts)) + ["post"] and owns[0] is bufs[0][rank]
    assert rel_err(dA1.cpu(), dA0.cpu()) < 2e-4
    stitched = torch.cat([torch.cat([bufs[i][o] for i in range(n_parts)]) for o in range(world)])
    assert rel_err(stitched.cpu(), dB0.cpu()) < 2e-4
    assert abs(dl1.item() - dl0.item()) <= 2e-4 * max(abs(dl0.item()), 1e-3)


def test_gemm_split_single_launch_matches_three_products():
    """mmg_gemm_split: A_hi.B_hi + A_hi.B_lo + A_lo.B_hi in one contraction over three K segments (all four operand
    layouts, bias + ReLU epilogue, split-K)."""
    from mmgclip_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 300, 520, 200
    for a_mn in (False, True):
        for b_mn in (False, True):
            if (a_mn and M % 8) or (b_mn and N % 8):
                M2, N2 = 304, 520
            else:
                M2, N2 = M, N
            A = torch.randn((K, M2) if a_mn else (M2, K), device="cuda", generator=gen)
            B = torch.randn((K, N2) if b_mn else (N2, K), device="cuda", generator=gen)
            Ao, Bo = ops._Operand(*ops.cast_bf16_split(A)), ops._Operand(*ops.cast_bf16_split(B))
            bias = torch.randn(N2, device="cuda", generator=gen)
            out = ops.gemm_heads(Ao, Bo, M2, N2, K, a_mn=a_mn, b_mn=b_mn, bias=bias, relu=True)
            f = lambda t, mn: (t.t() if mn else t).double()  # noqa: E731
            ref = (f(Ao.hi, a_mn) @ f(Bo.hi, b_mn).t() + f(Ao.hi, a_mn) @ f(Bo.lo, b_mn).t()
                   + f(Ao.lo, a_mn) @ f(Bo.hi, b_mn).t() + bias.double()).clamp_min(0)
            assert rel_err(out.cpu(), ref.cpu()) < 2e-6
            exact = (f(A, a_mn) @ f(B, b_mn).t() + bias.double()).clamp_min(0)
            assert rel_err(out.cpu(), exact.cpu()) < 3e-5          # the split keeps ~16 mantissa bits per operand
    # split-K over the concatenated range (the dW shape: short and wide, long K)
    M, N, K = 256, 384, 4096
    A = torch.randn(K, M, device="cuda", generator=gen)
    B = torch.randn(K, N, device="cuda", generator=gen)
    Ao, Bo = ops._Operand(*ops.cast_bf16_split(A)), ops._Operand(*ops.cast_bf16_split(B))
    out = ops.gemm_heads(Ao, Bo, M, N, K, a_mn=True, b_mn=True, k_splits=12)
    assert rel_err(out.cpu(), (A.double().t() @ B.double()).cpu()) < 3e-5
