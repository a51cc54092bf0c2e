"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not shipped, not a fallback.

CPU restatement of the reference's arithmetic for the contrastive hot path of abdel-habib/mmg-clip, used only as the
checker by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py``.  Nothing under ``mmgclip_b200/`` imports this module.

Two independent statements of the same maths:

* ``torch_*``  -- fp32 eager PyTorch on the CPU, operation for operation what the reference executes (same library
  calls: ``nn.functional.linear``, ``Tensor.norm``, ``@``, ``F.cross_entropy``, autograd).  Each function cites the
  reference lines it follows.  This is also the "port" that is timed as the CPU baseline.
* ``closed_form_*`` -- NumPy float64 closed-form loss and gradients, sharing no code with the above.

Pinning: the reference ships no tests; its only known-answer vectors are in notebooks/loss.ipynb and
the docstrings of losses.py.  The oracle is pinned against (a) those, and (b) outputs of the reference's own source
files executed in the build container (``tests/golden/make_golden.py`` imports /root/reference/mmgclip/loss/losses.py
and networks/projection.py by path and freezes inputs/outputs into ``tests/golden/*.npz``).
``tests/test_oracle_golden.py`` checks every fixture.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------------------
# torch fp32 restatement
# --------------------------------------------------------------------------------------------------------------


def torch_linear_projection(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """LinearProjectionLayer.forward: bias-free nn.Linear (mmgclip/networks/projection.py:17,33)."""
    return F.linear(x, weight)


def torch_multi_linear_head(x, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor],
                            keep_masks: Optional[Sequence[Optional[torch.Tensor]]] = None, p: float = 0.0):
    """MultiLinearHead.forward (projection.py:54-61): Linear -> ReLU -> Dropout for all but the last layer.
    ``keep_masks[i]`` (0/1) replaces nn.Dropout's RNG so a run can be reproduced exactly; None = eval mode."""
    last = len(weights) - 1
    for i, (w, b) in enumerate(zip(weights, biases)):
        x = F.linear(x, w, b)
        if i < last:
            x = torch.relu(x)
            if keep_masks is not None and keep_masks[i] is not None:
                x = x * keep_masks[i].to(x.dtype) / (1.0 - p)
    return x


def torch_mlp_projection_head(x, w_proj, b_proj, w_fc, b_fc, ln_w, ln_b, keep_mask=None, p: float = 0.0,
                              eps: float = 1e-5):
    """MLPProjectionHead.forward (projection.py:94-101): LN(Dropout(fc(GELU(proj(x)))) + proj(x))."""
    projected = F.linear(x, w_proj, b_proj)
    h = F.gelu(projected)
    h = F.linear(h, w_fc, b_fc)
    if keep_mask is not None:
        h = h * keep_mask.to(h.dtype) / (1.0 - p)
    h = h + projected
    return F.layer_norm(h, (h.shape[-1],), ln_w, ln_b, eps)


def torch_normalize(x: torch.Tensor) -> torch.Tensor:
    """x / x.norm(dim=1, keepdim=True) -- no epsilon (mmgclip/networks/mmgclip_model.py:128-129)."""
    return x / x.norm(dim=1, keepdim=True)


def torch_logits(image_embeddings, text_embeddings, logit_scale):
    """logit_scale * I @ T.t() and logit_scale * T @ I.t(): Python precedence scales first (mmgclip_model.py:135-136).
    ``logit_scale`` is the exponentiated value (mmgclip_model.py:132)."""
    lpi = logit_scale * image_embeddings @ text_embeddings.t()
    lpt = logit_scale * text_embeddings @ image_embeddings.t()
    return lpi, lpt


def torch_clip_loss(logits_per_image, logits_per_text):
    """CLIPLoss.forward (mmgclip/loss/losses.py:36-44) with the .cuda() of the labels dropped."""
    n, _ = logits_per_image.shape
    labels = torch.arange(n)
    loss_i = F.cross_entropy(logits_per_image, labels)
    loss_t = F.cross_entropy(logits_per_text, labels)
    return (loss_i + loss_t) / 2, labels


def torch_mmgclip_loss(image_embeddings, text_embeddings, text_embeddings2, logit_scale, t2t_weight: float = 0.5):
    """MMGCLIPLoss.forward (losses.py:63-96)."""
    lpi, lpt = torch_logits(image_embeddings, text_embeddings, logit_scale)
    loss_clip, labels = torch_clip_loss(lpi, lpt)
    l21 = logit_scale * text_embeddings2 @ text_embeddings.t()
    l12 = logit_scale * text_embeddings @ text_embeddings2.t()
    loss_t2t = (F.cross_entropy(l21, labels) + F.cross_entropy(l12, labels)) / 2.0
    return loss_clip + loss_t2t * t2t_weight, labels


def assign_labels(cosine_sim_matrix, threshold: float = 0.65) -> List[int]:
    """AveragedMedicalCLIPLoss._assign_labels (losses.py:141-162): greedy first-come clustering."""
    n = len(cosine_sim_matrix)
    labels = [-1] * n
    current = 0
    for i in range(n):
        if labels[i] == -1:
            labels[i] = current
            for j in range(i + 1, n):
                if float(cosine_sim_matrix[i][j]) >= threshold and labels[j] == -1:
                    labels[j] = current
            current += 1
    return labels


def average_logits(logits: torch.Tensor, list_labels: Sequence[int]) -> torch.Tensor:
    """AveragedMedicalCLIPLoss._average_logits (losses.py:164-186): mean of the columns sharing a label."""
    cols = []
    for label in sorted(set(list_labels)):
        idx = [i for i, l in enumerate(list_labels) if l == label]
        cols.append(logits[:, idx].mean(dim=1))
    return torch.stack(cols, dim=1)


def torch_averaged_medical_clip_loss(text_embeddings, logits_per_image, logits_per_text, threshold: float = 0.65):
    """AveragedMedicalCLIPLoss.forward (losses.py:189-216); cos_sim = normalised dot products (losses.py:119)."""
    unit = F.normalize(text_embeddings, dim=1)
    list_labels = assign_labels((unit @ unit.t()).tolist(), threshold)
    avg = average_logits(logits_per_image, list_labels)
    labels = torch.tensor(list_labels)
    return (F.cross_entropy(avg, labels) + F.cross_entropy(logits_per_text, labels)) / 2, labels


def torch_zeroshot(image_embeddings, text_embeddings, logit_scale):
    """PromptClassifier scoring: logits_per_image.softmax(-1) then argmax of the probabilities
    (mmgclip_model.py:201-209; NumPy twin evaluator.py:354-368)."""
    logits = logit_scale * image_embeddings @ text_embeddings.t()
    probs = logits.softmax(dim=-1)
    return logits, probs, torch.argmax(probs, dim=-1)


def torch_eos_pool(hidden: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """'eos' text pooling, statement for statement (mmgclip/networks/mmgclip_model.py:108-111): the hidden state at
    index ``attention_mask.sum(-1) - 1`` of every sequence (an all-zero mask yields -1, i.e. the last position)."""
    eos_token_indices = attention_mask.sum(dim=-1) - 1
    return hidden[torch.arange(hidden.shape[0]), eos_token_indices]


def torch_adamw_steps(params: Sequence[torch.Tensor], grads_per_step: Sequence[Sequence[torch.Tensor]], lr,
                      weight_decay: float, betas=(0.9, 0.999), eps: float = 1e-8) -> List[torch.Tensor]:
    """The reference's optimiser, by the same library call: ``torch.optim.AdamW(model.parameters(), lr=..., weight_decay=...)``
    followed by ``optimizer.step()`` per iteration (mmgclip/experiments/ClassifierExperiment.py:74,118), single-tensor
    implementation on the CPU.  ``lr`` may be a float or one value per step (scheduler)."""
    ps = [torch.nn.Parameter(p.detach().clone()) for p in params]
    lr0 = lr if isinstance(lr, (int, float)) else lr[0]
    opt = torch.optim.AdamW(ps, lr=lr0, weight_decay=weight_decay, betas=betas, eps=eps, foreach=False)
    for t, grads in enumerate(grads_per_step):
        if not isinstance(lr, (int, float)):
            for g in opt.param_groups:
                g["lr"] = lr[t]
        for p, g in zip(ps, grads):
            p.grad = g.detach().clone()
        opt.step()
    return [p.detach() for p in ps]


def torch_train_step(image_features, text_features, w_image, w_text, logit_scale_log) -> Dict[str, torch.Tensor]:
    """One step of the reference hot path with LinearProjectionLayer heads and CLIPLoss, forward and backward:
    heads -> normalise -> exp -> two logit GEMMs -> CLIPLoss -> backward (ClassifierExperiment.py:109-115).
    Leaves: the two head weights and the log-scale.  fp32, eager, on the CPU -- this is what ``cpu_baseline`` times."""
    w_i = w_image.detach().clone().requires_grad_(True)
    w_t = w_text.detach().clone().requires_grad_(True)
    ls = logit_scale_log.detach().clone().requires_grad_(True)
    ie = torch_normalize(torch_linear_projection(image_features, w_i))
    te = torch_normalize(torch_linear_projection(text_features, w_t))
    s = ls.exp()
    lpi, lpt = torch_logits(ie, te, s)
    loss, _ = torch_clip_loss(lpi, lpt)
    loss.backward()
    return {"loss": loss.detach(), "dw_image": w_i.grad, "dw_text": w_t.grad, "dlogit_scale_log": ls.grad,
            "image_embeddings": ie.detach(), "text_embeddings": te.detach()}


# --------------------------------------------------------------------------------------------------------------
# NumPy float64 closed form -- independent of autograd
# --------------------------------------------------------------------------------------------------------------


def closed_form_info_nce(a: np.ndarray, b: np.ndarray, s: float):
    """Loss and gradients of (CE(L, arange) + CE(L^T, arange))/2, L = s * a b^T, for unit-row a, b (float64).

    G = dloss/dL = (softmax_rows(L) + softmax_cols(L) - 2I) / (2n);  da = s G b;  db = s G^T a;  ds = sum G * (a b^T).
    """
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    n = a.shape[0]
    cos = a @ b.T
    L = s * cos
    r_lse = _logsumexp(L, axis=1)
    c_lse = _logsumexp(L, axis=0)
    diag = np.diag(L)
    loss = 0.5 * ((r_lse - diag).mean() + (c_lse - diag).mean())
    G = (np.exp(L - r_lse[:, None]) + np.exp(L - c_lse[None, :]) - 2.0 * np.eye(n)) / (2.0 * n)
    return {"loss": loss, "da": s * G @ b, "db": s * G.T @ a, "ds": float((G * cos).sum()), "G": G}


def closed_form_normalize_backward(u: np.ndarray, dy: np.ndarray) -> np.ndarray:
    """du for y = u/||u||: (dy - y <y, dy>) / ||u||."""
    u = np.asarray(u, dtype=np.float64)
    nrm = np.linalg.norm(u, axis=1, keepdims=True)
    y = u / nrm
    return (dy - y * (y * dy).sum(axis=1, keepdims=True)) / nrm


def closed_form_train_step(xi, xt, wi, wt, logit_scale_log: float):
    """float64 closed form of torch_train_step."""
    xi, xt, wi, wt = (np.asarray(t, dtype=np.float64) for t in (xi, xt, wi, wt))
    ui, ut = xi @ wi.T, xt @ wt.T
    a = ui / np.linalg.norm(ui, axis=1, keepdims=True)
    b = ut / np.linalg.norm(ut, axis=1, keepdims=True)
    s = math.exp(logit_scale_log)
    r = closed_form_info_nce(a, b, s)
    dui = closed_form_normalize_backward(ui, r["da"])
    dut = closed_form_normalize_backward(ut, r["db"])
    return {"loss": r["loss"], "dw_image": dui.T @ xi, "dw_text": dut.T @ xt, "dlogit_scale_log": s * r["ds"],
            "image_embeddings": a, "text_embeddings": b, "da": r["da"], "db": r["db"], "ds": r["ds"]}


def closed_form_zeroshot(img: np.ndarray, txt: np.ndarray, s: float, k: int = 0):
    """float64 logits / softmax / argmax (first occurrence) / top-k ordered (value desc, index asc)."""
    logits = (s * np.asarray(img, np.float64)) @ np.asarray(txt, np.float64).T
    z = logits - logits.max(axis=1, keepdims=True)
    probs = np.exp(z)
    probs /= probs.sum(axis=1, keepdims=True)
    out = {"logits": logits, "probs": probs, "argmax": np.argmax(logits, axis=1)}
    if k:
        order = np.lexsort((np.broadcast_to(np.arange(logits.shape[1]), logits.shape), -logits), axis=1)
        out["topk_idx"] = order[:, :k]
    return out


def closed_form_adamw_steps(params, grads_per_step, lr, weight_decay: float, betas=(0.9, 0.999), eps: float = 1e-8):
    """float64 AdamW (Loshchilov & Hutter; the algorithm box of torch.optim.AdamW's documentation), independent of
    torch's implementation: theta <- theta(1 - lr*wd); m, v moments; theta <- theta - lr * m_hat / (sqrt(v_hat) + eps)."""
    th = [np.asarray(p, np.float64).copy() for p in params]
    m = [np.zeros_like(x) for x in th]
    v = [np.zeros_like(x) for x in th]
    b1, b2 = betas
    for t, grads in enumerate(grads_per_step, start=1):
        lr_t = lr if isinstance(lr, (int, float)) else lr[t - 1]
        for i, g in enumerate(grads):
            g = np.asarray(g, np.float64)
            th[i] *= 1.0 - lr_t * weight_decay
            m[i] = b1 * m[i] + (1.0 - b1) * g
            v[i] = b2 * v[i] + (1.0 - b2) * g * g
            m_hat = m[i] / (1.0 - b1 ** t)
            v_hat = v[i] / (1.0 - b2 ** t)
            th[i] -= lr_t * m_hat / (np.sqrt(v_hat) + eps)
    return th


def _logsumexp(x: np.ndarray, axis: int) -> np.ndarray:
    m = x.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(x - m).sum(axis=axis, keepdims=True))).squeeze(axis)


# --------------------------------------------------------------------------------------------------------------
# synthetic inputs: encoder-feature-like statistics, portable NumPy generator
# --------------------------------------------------------------------------------------------------------------


def synthetic_features(batch: int, e_image: int = 768, e_text: int = 768, seed: int = 42):
    """Image features ~ clamp(1 + 0.35 N(0,1), min=0) (ConvNeXt avg-pool-like); text features ~ 0.5 N(0,1) (BERT-like)."""
    rng = np.random.RandomState(seed)
    xi = np.maximum(1.0 + 0.35 * rng.standard_normal((batch, e_image)), 0.0).astype(np.float32)
    xt = (0.5 * rng.standard_normal((batch, e_text))).astype(np.float32)
    return xi, xt


def synthetic_head_weights(d: int, e_image: int = 768, e_text: int = 768, seed: int = 43):
    """nn.Linear default init (kaiming_uniform(a=sqrt 5) = U(-1/sqrt(E), 1/sqrt(E))) from a portable generator."""
    rng = np.random.RandomState(seed)
    wi = rng.uniform(-1.0, 1.0, (d, e_image)).astype(np.float32) / np.float32(math.sqrt(e_image))
    wt = rng.uniform(-1.0, 1.0, (d, e_text)).astype(np.float32) / np.float32(math.sqrt(e_text))
    return wi, wt
