"""Losses with the reference's names and call signatures (mmgclip/loss/losses.py), computed by the fused kernels.

The experiment calls ``loss, labels = criterion(**outputs)`` with *every* key of the model's output dict
(ClassifierExperiment.py:112,172), so a loss here can pick the normalised embeddings and the scale and never touch a
materialised logit matrix.  When only logits are supplied (the literal reference signature) the same cross-entropy is
evaluated on them by the row-LSE kernel.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops


def _arange_labels(n: int, like: torch.Tensor) -> torch.Tensor:
    # reference: torch.arange(n).cuda() -> int64 on the default CUDA device (losses.py:39,78)
    return torch.arange(n, device=like.device if like.is_cuda else "cuda")


class CLIPLoss(nn.Module):
    """Symmetric cross-entropy over image->text and text->image logits with labels ``arange(n)`` (losses.py:6-44)."""

    def __init__(self, precision=None):
        super().__init__()
        self.precision = precision

    def forward(self, logits_per_image=None, logits_per_text=None, **kwargs):
        ie = kwargs.get("image_embeddings")
        te = kwargs.get("text_embeddings")
        s = kwargs.get("logit_scale")
        if ie is not None and te is not None and s is not None and ie.shape == te.shape:
            # fused path: (CE(L) + CE(L^T)) / 2 with L = s * I T^T never written to memory
            loss = ops.info_nce(ie, te, s, prec=self.precision)
            return loss, _arange_labels(ie.shape[0], ie)
        if logits_per_image is None or logits_per_text is None:
            raise TypeError("CLIPLoss.forward() needs logits_per_image and logits_per_text, or image_embeddings, "
                            "text_embeddings and logit_scale")
        n, _ = logits_per_image.shape
        loss_i = ops.cross_entropy(logits_per_image)
        loss_t = ops.cross_entropy(logits_per_text)
        return (loss_i + loss_t) / 2, _arange_labels(n, logits_per_image)


class MMGCLIPLoss(nn.Module):
    """CLIP loss plus a text<->text InfoNCE between two text views, weighted by ``t2t_weight`` (losses.py:46-96)."""

    def __init__(self, t2t_weight=0.5, precision=None):
        super().__init__()
        self.t2t_weight = t2t_weight
        self.precision = precision

    def forward(self, image_embeddings, text_embeddings, text_embeddings2, logit_scale, **kwargs):
        loss_clip = ops.info_nce(image_embeddings, text_embeddings, logit_scale, prec=self.precision)
        # losses.py:85-91: rows = second text view, columns = first text view (and the transpose)
        loss_t2t = ops.info_nce(text_embeddings2, text_embeddings, logit_scale, prec=self.precision)
        loss = loss_clip + (loss_t2t * self.t2t_weight)
        return loss, _arange_labels(image_embeddings.shape[0], image_embeddings)


class AveragedMedicalCLIPLoss(nn.Module):
    """CLIP loss whose image->text logits are averaged over groups of near-duplicate texts (losses.py:98-216).

    Texts whose cosine similarity to an earlier, still-unlabelled text reaches ``similarity_threshold`` share its label
    (greedy, first-come order, losses.py:141-162); ``logits_per_image`` columns of one label are averaged
    (losses.py:164-186); both cross-entropies then use those labels.  The O(n^2) clustering runs on the host from one
    device->host copy of the similarity matrix (the reference syncs once per matrix element); the similarity matrix,
    the column averaging and the cross-entropies are kernels.
    """

    def __init__(self, similarity_threshold=0.65, precision=None):
        super().__init__()
        self.similarity_threshold = similarity_threshold
        self.precision = precision

    def _mesaure_embeddings_similarity(self, embeddings):
        unit = ops.l2_normalize(embeddings.detach().to(torch.float32), prec="fp32")
        return ops.similarity_logits(unit, unit, 1.0, prec="fp32")

    def _assign_labels(self, cosine_sim_matrix, threshold=0.65):
        sim = cosine_sim_matrix.detach().cpu().tolist() if torch.is_tensor(cosine_sim_matrix) else cosine_sim_matrix
        n = len(sim)
        labels = [-1] * n
        next_label = 0
        for i in range(n):
            if labels[i] != -1:
                continue
            labels[i] = next_label
            row = sim[i]
            for j in range(i + 1, n):
                if labels[j] == -1 and row[j] >= threshold:
                    labels[j] = next_label
            next_label += 1
        return labels

    def _average_logits(self, logits, list_labels):
        n_groups = max(list_labels) + 1
        counts = [0] * n_groups
        for lab in list_labels:
            counts[lab] += 1
        # averaging as one contraction: avg[:, g] = sum_c logits[:, c] * W[g, c],  W[g, c] = [label_c == g] / count_g
        w = torch.zeros((n_groups, len(list_labels)), dtype=torch.float32)
        for c, lab in enumerate(list_labels):
            w[lab, c] = 1.0 / counts[lab]
        return ops.linear(logits.to(torch.float32), w.to(logits.device), None, prec="fp32")

    def forward(self, image_embeddings, text_embeddings, logit_scale, logits_per_image=None, logits_per_text=None,
                **kwargs):
        if logits_per_image is None:  # a caller that skipped the materialisation (mmgclip_model.py:135-136)
            logits_per_image = ops.similarity_logits(image_embeddings, text_embeddings, logit_scale, prec=self.precision)
        if logits_per_text is None:
            logits_per_text = ops.similarity_logits(text_embeddings, image_embeddings, logit_scale, prec=self.precision)
        sim = self._mesaure_embeddings_similarity(text_embeddings)
        list_labels = self._assign_labels(sim, threshold=self.similarity_threshold)
        averaged = self._average_logits(logits=logits_per_image, list_labels=list_labels)
        labels = torch.tensor(list_labels).to(logits_per_image.device)
        loss_i = ops.cross_entropy(averaged, labels)
        loss_t = ops.cross_entropy(logits_per_text, labels)
        return (loss_i + loss_t) / 2, labels
