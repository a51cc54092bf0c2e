"""Where does the 16-bit path's gradient error come from?  (analysis script, CPU only, not a pytest file)

    python tests/emulate_operand_rounding.py [batch]

Float64 statement of the benchmark step (heads -> normalise -> symmetric InfoNCE -> head-weight gradients, the arithmetic
of mmgclip_model.py:124-136 + losses.py:36-44 as restated in SURVEY.md s3.5) in which ONE class of tensor-core operand at a
time is rounded the way the kernels round it:

    cos   = the embeddings as operands of the cosine contractions (forward sums and the backward's recomputation)
    gemm  = the embeddings as the B operand of the two gradient contractions  dI = g.T,  dT = g^T.I
    g     = the gradient coefficients (bf16 in true units, or fp16 in 2^14-scaled units: MMG_PREC_F16)

Printed: max-abs / max-abs and Frobenius errors of dI, dT, dW_image, dW_text against the unrounded float64 step.  The
"bf16 bf16 bf16" row reproduces the GPU's `parity` block to three digits (batch 4096: dW_image 1.74e-3 / 2.13e-3), which is
what licenses reading the other rows as predictions; DESIGN.md s2 quotes them."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import clip_oracle as oc  # noqa: E402  (test infrastructure: same synthetic inputs as the bench)


def rnd(x, kind):
    if kind == "bf16":
        return x.float().bfloat16().double()
    if kind == "f16":
        return x.float().half().double()
    return x


def run(xi, xt, wi, wt, s, cos_kind, gemm_kind, g_kind):
    B = xi.shape[0]
    ui, ut = xi @ wi.T, xt @ wt.T
    ni, nt = ui.norm(dim=1, keepdim=True), ut.norm(dim=1, keepdim=True)
    I, T = ui / ni, ut / nt
    Ic, Tc = rnd(I, cos_kind), rnd(T, cos_kind)
    Ig, Tg = rnd(I, gemm_kind), rnd(T, gemm_kind)
    E = torch.exp(s * (Ic @ Tc.T) - s)
    rs, cs = E.sum(1), E.sum(0)
    coef = s * 0.5 / B
    if g_kind == "f16-scaled":
        g = E * (16384.0 / rs[:, None] + 16384.0 / cs[None, :])
        g.fill_diagonal_(0)
        g = g.float().half().double() * (coef / 16384.0)
    else:
        g = E * (coef / rs[:, None] + coef / cs[None, :])
        g.fill_diagonal_(0)
        g = rnd(g, g_kind)
    dI, dT = g @ Tg, g.T @ Ig
    diag = (I * T).sum(1) * s  # the matching pair is applied in fp32 by mmg_infonce_bwd_prep_diag
    gd = torch.exp(diag - s) * (coef / rs + coef / cs) - 2 * coef
    dI += gd[:, None] * T
    dT += gd[:, None] * I
    dui = (dI - I * (I * dI).sum(1, keepdim=True)) / ni
    dut = (dT - T * (T * dT).sum(1, keepdim=True)) / nt
    return dI, dT, dui.T @ xi, dut.T @ xt


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    torch.set_num_threads(os.cpu_count() or 1)
    xi, xt = oc.synthetic_features(B, 768, 768, seed=42)
    wi, wt = oc.synthetic_head_weights(512, 768, 768, seed=43)
    xi, xt, wi, wt = (torch.from_numpy(np.asarray(a)).double() for a in (xi, xt, wi, wt))
    s = 1 / 0.07
    ref = run(xi, xt, wi, wt, s, None, None, None)

    def err(a, b):
        return ((a - b).abs().max() / b.abs().max()).item(), ((a - b).norm() / b.norm()).item()

    print(f"batch {B}: operand rounding (cos, gemm, g) -> error vs float64 (max-abs / Frobenius)")
    for ck, gk, gg in (("bf16", "bf16", "bf16"), (None, None, "bf16"), ("f16", "bf16", "bf16"), ("bf16", "f16", "bf16"),
                       ("f16", "f16", "bf16"), ("f16", "f16", "f16-scaled")):
        out = run(xi, xt, wi, wt, s, ck, gk, gg)
        e = [err(o, r) for o, r in zip(out, ref)]
        print(f"  {str(ck):5s} {str(gk):5s} {str(gg):10s} dI %.2e/%.2e  dT %.2e/%.2e  dW_image %.2e/%.2e  dW_text %.2e/%.2e"
              % tuple(v for pair in e for v in pair))


if __name__ == "__main__":
    main()
