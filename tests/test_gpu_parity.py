"""GPU parity: the CUDA path (through the Python boundary -> C ABI) against the golden vectors frozen from the
reference's own source files and against the oracle on seeded inputs.

Tolerances (BASELINE.json north_star): fp32 path 1e-5 relative; bf16 path: loss 2e-3 relative; labels/indices bit-exact.
"Relative" is max-abs error over the max-abs of the reference tensor (loss: relative to the loss value).

bf16 gradients: the O(B^2) contractions see operands rounded to bf16 (unit round-off 2^-9 = 1.95e-3 each), so a gradient
element can carry two such roundings: GRAD_BF16 = 4e-3 (max-abs and Frobenius).  Embeddings and everything computed by
the projection heads are fp32-faithful even in bf16 mode (three-pass bf16x3 contractions): EMB_BF16 = 2e-5.  Head-weight
gradients inherit the embedding-gradient error through a cancelling sum over the batch: DW_BF16 = 8e-3 max-abs, 6e-3
Frobenius.  (Measured on these cases: loss <= 3.5e-4, dI/dT <= 3.2e-3, dW <= 5.1e-3.)"""
import math

import numpy as np
import pytest
import torch

from conftest import fro_err, rel_err
from oracle import clip_oracle as oc

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-5, "bf16": 2e-3}          # loss
GRAD = {"fp32": 1e-5, "bf16": 4e-3}         # embedding gradients, d logit_scale
EMB = {"fp32": 2e-6, "bf16": 2e-5}          # embeddings
DW = {"fp32": 1e-5, "bf16": 8e-3}           # head-weight gradients (max-abs); Frobenius: 6e-3 in bf16
DW_FRO = {"fp32": 1e-5, "bf16": 6e-3}


def cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.fixture(scope="module")
def mm():
    from mmgclip_b200 import losses, model, ops, projection
    return type("M", (), {"ops": ops, "losses": losses, "projection": projection, "model": model})


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_linear_heads_clip_loss_vs_reference_fixture(mm, golden, prec):
    g = golden("clip_linear_small")
    P = mm.projection
    hi = P.LinearProjectionLayer(96, 64, precision=prec).cuda()
    ht = P.LinearProjectionLayer(80, 64, precision=prec).cuda()
    hi.load_state_dict({"layer.weight": cuda(g["w_image"])})
    ht.load_state_dict({"layer.weight": cuda(g["w_text"])})
    ls = torch.tensor(float(g["logit_scale_log"]), device="cuda", requires_grad=True)
    ie = hi.forward_normalized(cuda(g["xi"]))
    te = ht.forward_normalized(cuda(g["xt"]))
    ie.retain_grad(); te.retain_grad()
    loss, labels = mm.losses.CLIPLoss(precision=prec)(image_embeddings=ie, text_embeddings=te, logit_scale=ls.exp(),
                                                      logits_per_image=None, logits_per_text=None)
    loss.backward()
    tol = TOL[prec]
    assert labels.dtype == torch.int64 and labels.is_cuda and labels.tolist() == g["labels"].tolist()
    assert abs(loss.item() - float(g["loss"])) / float(g["loss"]) < tol
    assert rel_err(ie.detach().cpu(), g["image_embeddings"]) < EMB[prec]
    assert rel_err(te.detach().cpu(), g["text_embeddings"]) < EMB[prec]
    assert rel_err(ie.grad.cpu(), g["d_image_embeddings"]) < GRAD[prec]
    assert rel_err(te.grad.cpu(), g["d_text_embeddings"]) < GRAD[prec]
    assert fro_err(ie.grad.cpu(), g["d_image_embeddings"]) < GRAD[prec]
    assert fro_err(te.grad.cpu(), g["d_text_embeddings"]) < GRAD[prec]
    assert rel_err(hi.layer.weight.grad.cpu(), g["dw_image"]) < DW[prec]
    assert rel_err(ht.layer.weight.grad.cpu(), g["dw_text"]) < DW[prec]
    assert fro_err(hi.layer.weight.grad.cpu(), g["dw_image"]) < DW_FRO[prec]
    assert fro_err(ht.layer.weight.grad.cpu(), g["dw_text"]) < DW_FRO[prec]
    assert abs(ls.grad.item() - float(g["dlogit_scale_log"])) < GRAD[prec] * max(1.0, abs(float(g["dlogit_scale_log"])))


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_cfg1_shape_b32_768_512(mm, golden, prec):
    """BASELINE config 1: batch 32, 768-d features, projection 512 (train_prompt_clf shape)."""
    g = golden("clip_cfg1_b32_768_512")
    xi, xt = oc.synthetic_features(32, 768, 768, seed=int(g["seed_inputs"]))
    wi, wt = oc.synthetic_head_weights(512, 768, 768, seed=int(g["seed_weights"]))
    P = mm.projection
    hi, ht = P.LinearProjectionLayer(768, 512, precision=prec).cuda(), P.LinearProjectionLayer(768, 512, precision=prec).cuda()
    hi.load_state_dict({"layer.weight": cuda(wi)})
    ht.load_state_dict({"layer.weight": cuda(wt)})
    ls = torch.tensor(math.log(1 / 0.07), device="cuda", requires_grad=True)
    ie, te = hi.forward_normalized(cuda(xi)), ht.forward_normalized(cuda(xt))
    loss, _ = mm.losses.CLIPLoss(precision=prec)(image_embeddings=ie, text_embeddings=te, logit_scale=ls.exp())
    loss.backward()
    tol = TOL[prec]
    assert abs(loss.item() - float(g["loss"])) / float(g["loss"]) < tol
    gi = hi.layer.weight.grad.cpu().numpy()
    gt = ht.layer.weight.grad.cpu().numpy()
    assert abs(np.linalg.norm(gi.astype(np.float64)) / float(g["dw_image_fro"]) - 1) < DW_FRO[prec]
    assert abs(np.linalg.norm(gt.astype(np.float64)) / float(g["dw_text_fro"]) - 1) < DW_FRO[prec]
    scale = np.abs(gi).max()
    assert np.abs(gi[:32, :32] - g["dw_image_block"]).max() / scale < DW[prec]
    assert rel_err(ie.detach().cpu().numpy()[:, :16], g["image_embeddings_head"]) < EMB[prec]
    assert abs(ls.grad.item() - float(g["dlogit_scale_log"])) < GRAD[prec] * max(1.0, abs(float(g["dlogit_scale_log"])))


def _load_head(head, g, tag):
    head.load_state_dict({k[len(f"p_{tag}."):]: cuda(v) for k, v in g.items() if k.startswith(f"p_{tag}.")})


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("kind", ["multilinear", "mlp"])
def test_deep_heads_vs_reference_fixture(mm, golden, prec, kind):
    g = golden(f"clip_{kind}_small")
    P = mm.projection
    if kind == "multilinear":
        dims = g["dims"].tolist()
        hi, ht = P.MultiLinearHead(48, dims, dropout=0.5, precision=prec), P.MultiLinearHead(40, dims, dropout=0.5, precision=prec)
    else:
        hi, ht = P.MLPProjectionHead(48, 32, dropout=0.5, precision=prec), P.MLPProjectionHead(40, 32, dropout=0.5, precision=prec)
    hi, ht = hi.cuda().eval(), ht.cuda().eval()
    _load_head(hi, g, "i"); _load_head(ht, g, "t")
    ie = mm.ops.l2_normalize(hi(cuda(g["xi"])), prec=prec)
    te = mm.ops.l2_normalize(ht(cuda(g["xt"])), prec=prec)
    s = torch.tensor(math.log(1 / 0.07), device="cuda").exp()
    loss, _ = mm.losses.CLIPLoss(precision=prec)(image_embeddings=ie, text_embeddings=te, logit_scale=s)
    loss.backward()
    tol = TOL[prec]
    assert rel_err(ie.detach().cpu(), g["image_embeddings"]) < 2 * EMB[prec]
    assert rel_err(te.detach().cpu(), g["text_embeddings"]) < 2 * EMB[prec]
    assert abs(loss.item() - float(g["loss"])) / float(g["loss"]) < tol
    for tag, head in (("i", hi), ("t", ht)):
        for k, p in head.named_parameters():
            ref = g[f"g_{tag}.{k}"]
            err = fro_err(p.grad.cpu(), ref)
            # toy widths (24 x 48 -> 56 -> 32): few terms to average the bf16 rounding of the O(B^2) part over
            assert err < (2e-5 if prec == "fp32" else 1.5e-2), (k, err)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_mmgclip_loss_vs_reference_fixture(mm, golden, prec):
    g = golden("mmgclip_loss_small")
    ie, te, te2 = (cuda(g[k]).requires_grad_() for k in ("image_embeddings", "text_embeddings", "text_embeddings2"))
    s = cuda(g["logit_scale"]).requires_grad_()
    loss, labels = mm.losses.MMGCLIPLoss(t2t_weight=0.5, precision=prec)(
        image_embeddings=ie, text_embeddings=te, text_embeddings2=te2, logit_scale=s, logits_per_image=None)
    loss.backward()
    tol = TOL[prec]
    assert labels.tolist() == list(range(20))
    assert abs(loss.item() - float(g["loss"])) / float(g["loss"]) < tol
    assert rel_err(ie.grad.cpu(), g["d_image_embeddings"]) < GRAD[prec]
    assert rel_err(te.grad.cpu(), g["d_text_embeddings"]) < GRAD[prec]
    assert rel_err(te2.grad.cpu(), g["d_text_embeddings2"]) < GRAD[prec]
    assert abs(s.grad.item() - float(g["d_logit_scale"])) < GRAD[prec] * max(1.0, abs(float(g["d_logit_scale"])))


def test_literal_logit_signature_known_answers(mm, golden):
    """CLIPLoss(logits_per_image, logits_per_text) exactly as the reference is called; KAT 2.1586060524."""
    k = golden("reference_kats")
    lg = cuda(k["logits8"]).requires_grad_()
    loss, labels = mm.losses.CLIPLoss()(lg, lg.t().contiguous())
    assert abs(loss.item() - 2.1586060524) < 2e-6
    assert labels.tolist() == list(range(8)) and labels.is_cuda
    loss.backward()
    ref = torch.from_numpy(k["logits8"]).requires_grad_()
    l_ref, _ = oc.torch_clip_loss(ref, ref.t())
    l_ref.backward()
    assert rel_err(lg.grad.cpu(), ref.grad) < 1e-5
    # AveragedMedicalCLIPLoss: notebook KAT (averaged CE 1.2048) and a full forward frozen from the reference
    am = mm.losses.AveragedMedicalCLIPLoss()
    avg = am._average_logits(cuda(k["logits8"]), k["notebook_labels"].tolist())
    assert rel_err(avg.cpu(), k["notebook_avg_logits"]) < 1e-6
    ce = mm.ops.cross_entropy(avg, cuda(k["notebook_labels"]))
    assert abs(ce.item() - 1.2048) < 5e-5
    loss_am, lab = mm.losses.AveragedMedicalCLIPLoss(0.65)(
        cuda(k["am_image_embeddings"]), cuda(k["am_text_embeddings"]), cuda(k["am_logit_scale"]),
        cuda(k["am_logits_per_image"]), cuda(k["am_logits_per_text"]))
    assert lab.tolist() == k["am_labels"].tolist()
    assert abs(loss_am.item() - float(k["am_loss"])) < 1e-5 * float(k["am_loss"])


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("n,d", [(1, 64), (7, 40), (129, 256), (1000, 512), (4096, 512)])
def test_infonce_vs_closed_form_ragged_sizes(mm, prec, n, d):
    """Empty-ish / ragged / multi-tile batches against the float64 closed form (seeded inputs)."""
    rng = np.random.RandomState(n * 31 + d)
    a = rng.standard_normal((n, d)); a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = rng.standard_normal((n, d)) + 0.7 * a; b /= np.linalg.norm(b, axis=1, keepdims=True)
    s32 = float(np.float32(1 / 0.07))
    ref = oc.closed_form_info_nce(a, b, s32)
    at, bt = cuda(a.astype(np.float32)).requires_grad_(), cuda(b.astype(np.float32)).requires_grad_()
    st = torch.tensor(s32, device="cuda", requires_grad=True)
    loss = mm.ops.info_nce(at, bt, st, prec=prec)
    (2.0 * loss).backward()  # non-unit upstream gradient
    tol = TOL[prec]
    assert abs(loss.item() - ref["loss"]) <= tol * max(ref["loss"], 1.0)
    # n = 1: the exact gradient is zero (softmax of one logit); compare against the size of its two cancelling parts
    floor = 2.0 * s32 / n * 0.05
    assert rel_err(at.grad.cpu(), 2.0 * ref["da"], floor=floor) < GRAD[prec]
    assert rel_err(bt.grad.cpu(), 2.0 * ref["db"], floor=floor) < GRAD[prec]
    assert abs(st.grad.item() - 2.0 * ref["ds"]) < GRAD[prec] * max(abs(2.0 * ref["ds"]), 1e-2)


def test_zero_row_gives_nan_like_the_reference(mm):
    """Quirk Q3: no epsilon in the normalisation -- an all-zero projected row is NaN there and here."""
    u = torch.randn(4, 32, device="cuda")
    u[2] = 0
    y = mm.ops.l2_normalize(u, prec="fp32")
    assert torch.isnan(y[2]).all() and torch.isfinite(y[[0, 1, 3]]).all()


def test_bf16_full_size_properties(mm):
    """BASELINE config 2 size (B = 4096, D = 512) and a 5-block ragged size: size-independent properties.

    (1) a<->b swap symmetry, (2) invariance to the gradient block shape, (3) row-sharded evaluation (two 'ranks' on one
    GPU through the raw offsets API) == unsharded, (4) sampled rows against float64, (5) no O(B^2) allocation."""
    ops = mm.ops
    for n in (4096, 9000):
        d = 512
        gen = torch.Generator(device="cuda").manual_seed(n)
        a = torch.nn.functional.normalize(torch.randn(n, d, device="cuda", generator=gen), dim=1)
        b = torch.nn.functional.normalize(torch.randn(n, d, device="cuda", generator=gen) + 0.5 * a, dim=1)
        s = torch.tensor(float(np.float32(1 / 0.07)), device="cuda")
        ab, bb = ops.cast_bf16(a), ops.cast_bf16(b)
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        rs, cs, dg = ops.infonce_forward_raw(ab, bb, s, 0, "bf16")
        loss = ops.infonce_loss_raw(rs, cs, dg, s, 0.5 / n)
        one = torch.ones((), device="cuda")
        dA, dB, dls = ops.infonce_backward_raw(ab, bb, s, rs, cs, one, 0.5 / n, 0, "bf16", a32=a, b32=b, diag=dg)
        peak = torch.cuda.max_memory_allocated() - base
        assert peak < 8192 * 8192 * 2 + 64 * n * d, peak     # one g block + O(B*D), independent of B*B
        # (1) swap symmetry
        rs2, cs2, dg2 = ops.infonce_forward_raw(bb, ab, s, 0, "bf16")
        loss2 = ops.infonce_loss_raw(rs2, cs2, dg2, s, 0.5 / n)
        assert abs(loss.item() - loss2.item()) < 1e-5 * loss.item()
        assert rel_err(rs.cpu(), cs2.cpu()) < 1e-5 and rel_err(cs.cpu(), rs2.cpu()) < 1e-5
        # (2) block-shape invariance of the backward
        dA2, dB2, dls2 = ops.infonce_backward_raw(ab, bb, s, rs, cs, one, 0.5 / n, 0, "bf16", 1024, 2048, a32=a, b32=b,
                                                  diag=dg)
        assert rel_err(dA2.cpu(), dA.cpu()) < 2e-4 and rel_err(dB2.cpu(), dB.cpu()) < 2e-4
        assert abs(dls2.item() - dls.item()) < 1e-3 * abs(dls.item()) + 1e-6
        # (3) two row shards with offsets, partial sums added
        h = n // 2
        rsa, csa, dga = ops.infonce_forward_raw(ab[:h], bb, s, 0, "bf16")
        rsb, csb, dgb = ops.infonce_forward_raw(ab[h:], bb, s, h, "bf16", colsum=csa)
        assert rel_err(torch.cat([rsa, rsb]).cpu(), rs.cpu()) < 1e-5
        assert rel_err(csb.cpu(), cs.cpu()) < 1e-5
        assert rel_err(torch.cat([dga, dgb]).cpu(), dg.cpu()) < 1e-6
        dAb, dBb, _ = ops.infonce_backward_raw(ab[h:], bb, s, rs[h:], cs, one, 0.5 / n, h, "bf16", a32=a[h:], b32=b[h:],
                                               diag=dg[h:])
        assert rel_err(dAb.cpu(), dA[h:].cpu()) < 2e-4
        # (4) 48 sampled rows against float64 on the host (operands as the kernel sees them: bf16-rounded)
        idx = np.random.RandomState(0).choice(n, 48, replace=False)
        a64, b64 = ab.double().cpu().numpy(), bb.double().cpu().numpy()
        sv = float(s.item())
        cos = a64[idx] @ b64.T
        e = np.exp(sv * cos - sv)
        assert rel_err(rs.cpu().numpy()[idx], e.sum(1)) < 2e-4
        assert rel_err(dg.cpu().numpy()[idx], sv * cos[np.arange(48), idx]) < 1e-5
        cs64 = cs.double().cpu().numpy()
        coef = sv * 0.5 / n
        g = e * (coef / e.sum(1)[:, None] + coef / cs64[None, :])
        g[np.arange(48), idx] -= 2 * coef
        ref_da = g @ b.double().cpu().numpy()
        assert rel_err(dA.cpu().numpy()[idx], ref_da) < GRAD["bf16"]


@pytest.mark.parametrize("rows,cols,d,off", [(32768, 32768, 512, 0), (16384, 131072, 1024, 5 * 16384)])
def test_bf16_baseline_full_sizes(mm, rows, cols, d, off):
    """BASELINE.json's full sizes -- config 3 (global batch 32768, D = 512) and one rank's share of config 5 (16384 local
    rows x 131072 global columns, D = 1024, rank 5 of 8): the fused path end to end with size-independent checks --
    sampled rows against float64, one fused backward launch, scratch independent of rows x cols."""
    ops = mm.ops
    gen = torch.Generator(device="cuda").manual_seed(rows + d)
    a = torch.nn.functional.normalize(torch.randn(rows, d, device="cuda", generator=gen), dim=1)
    b = torch.nn.functional.normalize(torch.randn(cols, d, device="cuda", generator=gen), dim=1)
    b[off:off + rows] = torch.nn.functional.normalize(b[off:off + rows] + 0.5 * a, dim=1)
    s = torch.tensor(float(np.float32(1 / 0.07)), device="cuda")
    ab, bb = ops.cast_bf16(a), ops.cast_bf16(b)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    rs, cs, dg = ops.infonce_forward_raw(ab, bb, s, off, "bf16")
    one = torch.ones((), device="cuda")
    n0 = ops._lib.load().mmg_kernel_launch_count()
    b32 = b[off:off + rows].contiguous()
    dA, dB, dls = ops.infonce_backward_raw(ab, bb, s, rs, cs, one, 0.5 / cols, off, "bf16", a32=a, b32=b32, diag=dg)
    torch.cuda.synchronize()
    assert ops._lib.load().mmg_kernel_launch_count() - n0 == 2      # prep + matching-pair init (one launch) + ONE fused launch
    peak = torch.cuda.max_memory_allocated() - base
    # outputs + paired-row copy (O(B*D)) + the coefficient scratch; a bf16 logit block alone would be rows*cols*2
    assert peak < (2 * rows + cols) * d * 4 + (256 << 20), peak
    assert peak < rows * cols * 2 or rows * cols * 2 < (1 << 30)
    assert torch.isfinite(dA).all() and torch.isfinite(dB).all()
    idx = np.random.RandomState(1).choice(rows, 32, replace=False)
    a64, b64 = ab[idx].double().cpu().numpy(), bb.double().cpu().numpy()
    sv = float(s.item())
    cos = a64 @ b64.T
    e = np.exp(sv * cos - sv)
    assert rel_err(rs.cpu().numpy()[idx], e.sum(1)) < 2e-4
    assert rel_err(dg.cpu().numpy()[idx], sv * cos[np.arange(32), off + idx]) < 1e-5
    coef = sv * 0.5 / cols
    g = e * (coef / e.sum(1)[:, None] + coef / cs.double().cpu().numpy()[None, :])
    g[np.arange(32), off + idx] -= 2 * coef
    # the matching-pair element uses the fp32 embeddings in the kernel path; everything else the bf16-rounded ones
    ref_da = g @ b64
    ref_da += (g[np.arange(32), off + idx])[:, None] * (b[off + idx].double().cpu().numpy() - b64[off + idx])
    assert rel_err(dA.cpu().numpy()[idx], ref_da) < GRAD["bf16"]
    assert abs(dls.item() - 0.0) < 1e3  # finite, sane magnitude (exact value checked at small sizes)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_multilinear_head_training_dropout(mm, prec):
    """MultiLinearHead in train mode: Linear -> ReLU -> Dropout(p) per hidden layer (projection.py:54-61).  The keep mask
    is drawn inside the kernel (Philox, mmg_dropout_draw_apply); the oracle is handed the mask that was drawn, and
    re-seeding the stream reproduces it."""
    P = mm.projection
    torch.manual_seed(5)
    head = P.MultiLinearHead(40, [48, 24], dropout=0.25, precision=prec).cuda().train()
    x = torch.randn(64, 40, device="cuda")
    mm.ops.seed_dropout(123)
    y = head(x)
    keep = mm.ops.last_dropout_mask.bool()
    assert keep.shape == (64, 48) and 0.6 < keep.float().mean().item() < 0.9
    mm.ops.seed_dropout(123)
    y_again = head(x)
    assert torch.equal(mm.ops.last_dropout_mask.bool(), keep) and torch.equal(y_again, y)   # same seed, same masks
    assert not torch.equal(head(x), y)                                                       # the stream advances
    w = [l.weight.detach().cpu() for l in head.layers]
    b = [l.bias.detach().cpu() for l in head.layers]
    ref = oc.torch_multi_linear_head(x.cpu(), w, b, keep_masks=[keep.cpu()], p=0.25)
    assert rel_err(y.detach().cpu(), ref) < EMB[prec] * 5
    frac_dropped = ((y.detach() != 0).float().mean().item())
    assert frac_dropped > 0.5  # sanity: not everything was zeroed
    # gradients flow only through kept, active units
    wts = torch.randn(64, 24, device="cuda")
    (y * wts).sum().backward()
    ws = [t.clone().requires_grad_() for t in w]
    bs = [t.clone().requires_grad_() for t in b]
    yr = oc.torch_multi_linear_head(x.cpu(), ws, bs, keep_masks=[keep.cpu()], p=0.25)
    (yr * wts.cpu()).sum().backward()
    tol = 1e-5 if prec == "fp32" else 2e-5
    for i, layer in enumerate(head.layers):
        assert rel_err(layer.weight.grad.cpu(), ws[i].grad) < tol * 5
        assert rel_err(layer.bias.grad.cpu(), bs[i].grad) < tol * 5
    head.eval()
    y_eval = head(x)
    ref_eval = oc.torch_multi_linear_head(x.cpu(), w, b)
    assert rel_err(y_eval.detach().cpu(), ref_eval) < EMB[prec] * 5
