"""Zero-shot prompt scoring at scale -- the arithmetic of ``Evaluator.zeroshot_label_prompt`` / ``zeroshot_eval`` /
``clf_conf_matrix`` (mmgclip/evaluator.py:147-256, 258-319, 321-478) without the per-batch NumPy round trip.

    similarities = logit_scale * image_embeddings @ text_embeddings.T      (evaluator.py:282-285, 354-357)
    similarities = softmax(similarities, axis=1)                           (evaluator.py:290, 362)
    y_pred       = argmax(similarities, axis=-1)                           (evaluator.py:299, 368)

ROC / AUC / bootstrap CIs stay on the host with sklearn, as in the reference (out of scope).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


def _to_device(x, device):
    if torch.is_tensor(x):
        return x.to(device=device, dtype=torch.float32)
    return torch.as_tensor(x, dtype=torch.float32).to(device)  # numpy arrays from Evaluator.encode_image / encode_text


def score_prompts(image_embeddings, text_embeddings, logit_scale, top_k: int = 0, device: Optional[str] = None,
                  chunk_rows: int = 1 << 22):
    """Probabilities [N, C], argmax [N] (int64) and optional top-k for N image embeddings against C prompts.

    Inputs are L2-normalised embeddings (torch tensors or NumPy arrays, as ``Evaluator.encode_*`` return) and the
    exponentiated ``logit_scale``.  Rows are processed in chunks so N is bounded only by memory for the outputs.
    """
    dev = torch.device(device or "cuda")
    img = _to_device(image_embeddings, dev)
    txt = _to_device(text_embeddings, dev)
    if torch.is_tensor(logit_scale):
        s = logit_scale.detach().to(device=dev, dtype=torch.float32).reshape(())
    else:
        s = torch.tensor(float(logit_scale), dtype=torch.float32, device=dev)
    outs = []
    for r0 in range(0, max(img.shape[0], 1), chunk_rows):
        outs.append(ops.zeroshot_score(img[r0:r0 + chunk_rows], txt, s, k=top_k))
    if len(outs) == 1:
        return outs[0]
    cat = lambda key: None if outs[0][key] is None else torch.cat([o[key] for o in outs], 0)  # noqa: E731
    return {k: cat(k) for k in outs[0]}


def zeroshot_label_prompt(image_embeddings, text_embeddings, logit_scale):
    """(probabilities, y_pred) as NumPy arrays -- what evaluator.py:354-368 computes before the sklearn metrics."""
    out = score_prompts(image_embeddings, text_embeddings, logit_scale)
    return out["probs"].cpu().numpy(), out["argmax"].cpu().numpy()
