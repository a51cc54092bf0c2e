"""The oracle against every golden vector: the reference's own known answers (notebooks/loss.ipynb, losses.py
docstrings) and outputs of the reference's source files frozen by tests/golden/make_golden.py.  CPU only."""
import math

import numpy as np
import torch

from conftest import fro_err, rel_err
from oracle import clip_oracle as oc

T = torch.from_numpy


def test_reference_known_answers(golden):
    k = golden("reference_kats")
    lg = T(k["logits8"])
    loss, labels = oc.torch_clip_loss(lg, lg.t())
    assert abs(loss.item() - 2.1586060524) < 1e-6            # CLIPLoss(lpi, lpi^T) on the notebook matrix
    assert abs(loss.item() - float(k["clip_loss8"])) < 1e-7
    assert labels.tolist() == list(range(8))
    avg = oc.average_logits(lg, k["notebook_labels"].tolist())
    assert rel_err(avg.numpy(), k["notebook_avg_logits"]) < 1e-7
    ce = torch.nn.functional.cross_entropy(avg, T(k["notebook_labels"]))
    assert abs(ce.item() - 1.2048) < 5e-5                      # value printed in notebooks/loss.ipynb (cells 13-18)
    assert abs(ce.item() - float(k["notebook_avg_ce"])) < 1e-7
    assert oc.assign_labels(k["doc_cosine"].tolist(), 0.65) == k["doc_labels"].tolist() == [0, 1, 0, 1, 0, 1, 0, 1]
    loss_am, lab_am = oc.torch_averaged_medical_clip_loss(T(k["am_text_embeddings"]), T(k["am_logits_per_image"]),
                                                          T(k["am_logits_per_text"]), 0.65)
    assert lab_am.tolist() == k["am_labels"].tolist()
    assert abs(loss_am.item() - float(k["am_loss"])) < 1e-6


def test_linear_head_clip_loss_small(golden):
    g = golden("clip_linear_small")
    r = oc.torch_train_step(T(g["xi"]), T(g["xt"]), T(g["w_image"]), T(g["w_text"]), T(g["logit_scale_log"]))
    assert abs(r["loss"].item() - float(g["loss"])) < 1e-6
    assert rel_err(r["image_embeddings"], g["image_embeddings"]) < 1e-6
    assert rel_err(r["dw_image"], g["dw_image"]) < 1e-5
    assert rel_err(r["dw_text"], g["dw_text"]) < 1e-5
    assert abs(r["dlogit_scale_log"].item() - float(g["dlogit_scale_log"])) < 1e-6
    lpi, lpt = oc.torch_logits(T(g["image_embeddings"]), T(g["text_embeddings"]), T(g["logit_scale"]))
    assert np.array_equal(lpi.numpy(), g["logits_per_image"])  # same ops, same machine class: bit-exact
    assert np.array_equal(lpt.numpy(), g["logits_per_text"])
    # independent float64 closed form agrees with the reference's autograd
    c = oc.closed_form_train_step(g["xi"], g["xt"], g["w_image"], g["w_text"], float(g["logit_scale_log"]))
    assert abs(c["loss"] - float(g["loss"])) < 2e-6
    assert rel_err(c["dw_image"], g["dw_image"]) < 2e-5
    assert rel_err(c["dw_text"], g["dw_text"]) < 2e-5
    assert rel_err(c["da"], g["d_image_embeddings"]) < 2e-5
    assert rel_err(c["db"], g["d_text_embeddings"]) < 2e-5
    assert abs(c["dlogit_scale_log"] - float(g["dlogit_scale_log"])) < 1e-5


def test_cfg1_shape(golden):
    g = golden("clip_cfg1_b32_768_512")
    xi, xt = oc.synthetic_features(32, 768, 768, seed=int(g["seed_inputs"]))
    wi, wt = oc.synthetic_head_weights(512, 768, 768, seed=int(g["seed_weights"]))
    r = oc.torch_train_step(T(xi), T(xt), T(wi), T(wt), torch.tensor(math.log(1 / 0.07)))
    assert abs(r["loss"].item() - float(g["loss"])) < 1e-6
    assert rel_err(r["dw_image"].numpy()[:32, :32], g["dw_image_block"]) < 1e-5
    assert rel_err(r["dw_text"].numpy()[:32, :32], g["dw_text_block"]) < 1e-5
    assert abs(np.linalg.norm(r["dw_image"].double().numpy()) / float(g["dw_image_fro"]) - 1) < 1e-5
    assert abs(np.linalg.norm(r["dw_text"].double().numpy()) / float(g["dw_text_fro"]) - 1) < 1e-5
    assert abs(r["dlogit_scale_log"].item() - float(g["dlogit_scale_log"])) < 1e-6
    c = oc.closed_form_train_step(xi, xt, wi, wt, math.log(1 / 0.07))
    assert abs(c["loss"] - float(g["loss"])) < 2e-6
    assert rel_err(c["dw_image"][:32, :32], g["dw_image_block"]) < 5e-5


def _head_params(g, tag):
    return {k[len(f"p_{tag}."):]: T(v) for k, v in g.items() if k.startswith(f"p_{tag}.")}


def test_multilinear_and_mlp_heads(golden):
    g = golden("clip_multilinear_small")
    outs = {}
    for tag, x in (("i", g["xi"]), ("t", g["xt"])):
        p = _head_params(g, tag)
        n_layers = len([k for k in p if k.endswith("weight")])
        ws = [p[f"layers.{i}.weight"] for i in range(n_layers)]
        bs = [p[f"layers.{i}.bias"] for i in range(n_layers)]
        outs[tag] = oc.torch_normalize(oc.torch_multi_linear_head(T(x), ws, bs))
    assert rel_err(outs["i"], g["image_embeddings"]) < 1e-6
    assert rel_err(outs["t"], g["text_embeddings"]) < 1e-6
    lpi, lpt = oc.torch_logits(outs["i"], outs["t"], torch.tensor(math.log(1 / 0.07)).exp())
    assert abs(oc.torch_clip_loss(lpi, lpt)[0].item() - float(g["loss"])) < 1e-6

    g = golden("clip_mlp_small")
    outs = {}
    for tag, x in (("i", g["xi"]), ("t", g["xt"])):
        p = _head_params(g, tag)
        outs[tag] = oc.torch_normalize(oc.torch_mlp_projection_head(
            T(x), p["projection.weight"], p["projection.bias"], p["fc.weight"], p["fc.bias"], p["layer_norm.weight"],
            p["layer_norm.bias"]))
    assert rel_err(outs["i"], g["image_embeddings"]) < 1e-6
    assert rel_err(outs["t"], g["text_embeddings"]) < 1e-6


def test_mmgclip_loss(golden):
    g = golden("mmgclip_loss_small")
    loss, _ = oc.torch_mmgclip_loss(T(g["image_embeddings"]), T(g["text_embeddings"]), T(g["text_embeddings2"]),
                                    T(g["logit_scale"]), 0.5)
    assert abs(loss.item() - float(g["loss"])) < 1e-6
    s = float(g["logit_scale"])
    c1 = oc.closed_form_info_nce(g["image_embeddings"], g["text_embeddings"], s)
    c2 = oc.closed_form_info_nce(g["text_embeddings2"], g["text_embeddings"], s)
    assert abs(c1["loss"] + 0.5 * c2["loss"] - float(g["loss"])) < 2e-6
    assert rel_err(c1["da"], g["d_image_embeddings"]) < 2e-5
    assert rel_err(c1["db"] + 0.5 * c2["db"], g["d_text_embeddings"]) < 2e-5
    assert rel_err(0.5 * c2["da"], g["d_text_embeddings2"]) < 2e-5
    assert abs(c1["ds"] + 0.5 * c2["ds"] - float(g["d_logit_scale"])) < 1e-5 * max(1.0, abs(float(g["d_logit_scale"])))


def test_zeroshot(golden):
    g = golden("zeroshot_small")
    lg, pr, am = oc.torch_zeroshot(T(g["img"]), T(g["txt"]), T(g["logit_scale"]))
    assert np.array_equal(lg.numpy(), g["logits"])
    assert np.array_equal(am.numpy(), g["argmax"])
    assert np.array_equal(g["argmax"], g["argmax_numpy"])      # torch and NumPy/SciPy paths of the reference agree
    assert rel_err(pr.numpy(), g["probs"]) < 1e-7
    c = oc.closed_form_zeroshot(g["img"], g["txt"], float(g["logit_scale"]), k=5)
    assert np.array_equal(c["argmax"], g["argmax"])            # duplicated prompt 2 == 5: first index wins
    assert fro_err(c["probs"], g["probs"]) < 1e-6
    assert not np.any(c["argmax"] == 5)


def test_eos_pool_and_adamw_statements():
    """The two 'next'-row helpers of the oracle: pooling is an index expression (checked on hand-made masks incl. the
    all-zero wrap-around); the float64 AdamW closed form is pinned against torch.optim.AdamW itself on the CPU."""
    g = torch.Generator().manual_seed(7)
    hidden = torch.randn(5, 9, 12, generator=g)
    mask = torch.zeros(5, 9, dtype=torch.int64)
    lens = [9, 1, 4, 0, 7]
    for r, n in enumerate(lens):
        mask[r, :n] = 1
    pooled = oc.torch_eos_pool(hidden, mask)
    for r, n in enumerate(lens):
        assert torch.equal(pooled[r], hidden[r, n - 1])  # n = 0 -> index -1 -> last position

    params = [torch.randn(33, 17, generator=g), torch.randn(17, generator=g)]
    grads = [[torch.randn(33, 17, generator=g) * 0.1, torch.randn(17, generator=g) * 0.1] for _ in range(6)]
    lrs = [1e-3, 2e-3, 3e-3, 2e-3, 1e-3, 5e-4]
    got = oc.torch_adamw_steps(params, grads, lrs, weight_decay=1e-2)
    want = oc.closed_form_adamw_steps([p.numpy() for p in params], [[x.numpy() for x in gs] for gs in grads], lrs, 1e-2)
    for a, b in zip(got, want):
        np.testing.assert_allclose(a.numpy(), b, rtol=2e-5, atol=2e-6)
