"""Stored-E variant of the bf16 InfoNCE (opt-in: MMGCLIP_B200_STORE_E_MB or ops.set_store_e_budget_mb; off by default): the forward keeps E = exp(logit - s) as bf16 and
the fused backward transforms it instead of recomputing the cosines.  Checked against the recompute path of the same
library and against the float64 closed form of the oracle -- same tolerances as the default bf16 path (loss 2e-3,
embedding gradients 4e-3 max-abs/max-abs)."""
import math
import os

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import clip_oracle as oc

pytestmark = pytest.mark.gpu

def _setenv(**kw):
    """Plan knobs of the fused backward (mmg_tune); no arguments = back to the defaults."""
    from mmgclip_b200 import ops
    ops.set_tuning()
    if kw:
        ops.set_tuning(**kw)


def _embeddings(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1)
    b = torch.nn.functional.normalize(torch.randn(n, d, generator=g) + 0.7 * a, dim=1)
    return a, b


@pytest.mark.parametrize("n,d,cfg", [
    (256, 256, {}),
    (1024, 512, {}),
    (2048, 256, {"fused_rb": 512, "fused_cb": 256, "fused_nbuf": 3, "fused_ksl": 2}),
    (4096, 512, {"fused_rb": 1024, "fused_cb": 1024, "fused_sr": 2, "fused_sc": 2}),
])
def test_stored_e_matches_recompute_and_float64(n, d, cfg):
    from mmgclip_b200 import ops
    a, b = _embeddings(n, d, seed=n + d)
    s = 1 / 0.07
    ref = oc.closed_form_info_nce(a.double().numpy(), b.double().numpy(), s)
    out = {}
    saved = ops.get_store_e_budget_mb()
    try:
        _setenv(**cfg)
        for mode, mb in (("recompute", 0), ("stored", 1 << 14)):
            ops.set_store_e_budget_mb(mb)
            assert ops.want_store_e(n, n, d, "bf16", False) == (mode == "stored")
            ac, bc = a.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
            loss = ops.info_nce(ac, bc, torch.tensor(s, device="cuda"), prec="bf16")
            loss.backward()
            torch.cuda.synchronize()
            out[mode] = (loss.item(), ac.grad.cpu().double().numpy(), bc.grad.cpu().double().numpy())
    finally:
        ops.set_store_e_budget_mb(saved)
        _setenv()
    for mode in ("recompute", "stored"):
        l, da, db = out[mode]
        assert abs(l - ref["loss"]) / ref["loss"] < 2e-3
        assert rel_err(da, ref["da"]) < 4e-3
        assert rel_err(db, ref["db"]) < 4e-3
    assert abs(out["stored"][0] - out["recompute"][0]) <= 1e-5 * abs(out["recompute"][0])
    assert rel_err(out["stored"][1], out["recompute"][1]) < 2e-3
    assert rel_err(out["stored"][2], out["recompute"][2]) < 2e-3


def test_stored_e_is_skipped_when_logit_scale_is_trained_or_shape_is_not_covered():
    from mmgclip_b200 import ops
    saved = ops.get_store_e_budget_mb()
    try:
        ops.set_store_e_budget_mb(1 << 14)
        assert not ops.want_store_e(1024, 1024, 512, "bf16", True)      # d/d logit_scale needs the cosines
        assert not ops.want_store_e(1000, 1000, 512, "bf16", False)     # not multiples of 256: block loop
        assert not ops.want_store_e(1024, 1024, 512, "fp32", False)
        ops.set_store_e_budget_mb(1)
        assert not ops.want_store_e(1024, 1024, 512, "bf16", False)     # 2 MiB of E does not fit a 1 MiB budget
        ops.set_store_e_budget_mb(1 << 14)
        # a trained scale silently takes the recompute path and still gets its gradient
        a, b = _embeddings(512, 256, seed=5)
        ac, bc = a.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
        sc = torch.tensor(1 / 0.07, device="cuda", requires_grad=True)
        ops.info_nce(ac, bc, sc, prec="bf16").backward()
        ref = oc.closed_form_info_nce(a.double().numpy(), b.double().numpy(), 1 / 0.07)
        assert abs(sc.grad.item() - ref["ds"]) <= 1e-2 * abs(ref["ds"]) + 1e-6
    finally:
        ops.set_store_e_budget_mb(saved)
