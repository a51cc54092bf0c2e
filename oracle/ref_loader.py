"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Loader for the reference's OWN source files, where they exist.

``/root/reference`` (or ``$MMG_REFERENCE_ROOT``) is present in the build container only; it never travels to the GPU
box.  When it is there, ``mmgclip/loss/losses.py`` and ``mmgclip/networks/projection.py`` are loaded by file path
(``import mmgclip`` itself needs fuzzywuzzy, an nltk download and hydra -- SURVEY.md s8c) with a one-function stub for
``sentence_transformers.util`` and, on a GPU-less box, ``Tensor.cuda`` made an identity (losses.py:39,78 call ``.cuda()``
unconditionally).  The nine arithmetic lines of mmgclip_model.py:124-136 cannot be imported (prettytable, HF downloads)
and are restated in :func:`reference_train_step`, next to their line numbers.

Used by ``tests/golden/make_golden*.py`` (fixture generation) and by ``bench.py --impl reference`` / its
``cpu_baseline`` leg (kind "reference" when this loader finds the files, else the port in ``clip_oracle.py``).
Nothing under ``mmgclip_b200/`` imports this module.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
import torch.nn.functional as F

_cache = None


def reference_root():
    for root in (os.environ.get("MMG_REFERENCE_ROOT"), "/root/reference"):
        if root and os.path.isfile(os.path.join(root, "mmgclip", "loss", "losses.py")):
            return root
    return None


def load_reference():
    """(losses module, projection module) of the reference, or None when its sources are not on this machine."""
    global _cache
    if _cache is not None:
        return _cache
    root = reference_root()
    if root is None:
        return None
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # losses.py:39,78
    st, ut = types.ModuleType("sentence_transformers"), types.ModuleType("sentence_transformers.util")
    ut.cos_sim = lambda a, b: F.normalize(a, dim=1) @ F.normalize(b, dim=1).t()  # only used at losses.py:119
    st.util = ut
    sys.modules.setdefault("sentence_transformers", st)
    sys.modules.setdefault("sentence_transformers.util", ut)

    def load(path, name):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m

    _cache = (load(os.path.join(root, "mmgclip/loss/losses.py"), "ref_losses"),
              load(os.path.join(root, "mmgclip/networks/projection.py"), "ref_projection"))
    return _cache


class ReferenceStep:
    """The reference's training step of the hot path built from ITS classes: two ``LinearProjectionLayer`` heads, the
    arithmetic of mmgclip_model.py:124-136, ``CLIPLoss`` and ``loss.backward()`` (ClassifierExperiment.py:109-115)."""

    def __init__(self, w_image: torch.Tensor, w_text: torch.Tensor, logit_scale_log: float):
        mods = load_reference()
        if mods is None:
            raise RuntimeError("the reference sources are not on this machine")
        ref_losses, ref_proj = mods
        d, e_i = w_image.shape
        e_t = w_text.shape[1]
        self.head_i = ref_proj.LinearProjectionLayer(e_i, d)
        self.head_t = ref_proj.LinearProjectionLayer(e_t, d)
        with torch.no_grad():
            self.head_i.layer.weight.copy_(w_image)
            self.head_t.layer.weight.copy_(w_text)
        self.logit_scale = torch.tensor(float(logit_scale_log), dtype=torch.float32)
        self.criterion = ref_losses.CLIPLoss()

    def __call__(self, xi: torch.Tensor, xt: torch.Tensor):
        self.head_i.layer.weight.grad = None
        self.head_t.layer.weight.grad = None
        ie = self.head_i(xi)                               # mmgclip_model.py:124
        te = self.head_t(xt)                               # :125
        ie = ie / ie.norm(dim=1, keepdim=True)             # :128
        te = te / te.norm(dim=1, keepdim=True)             # :129
        s = self.logit_scale.exp()                         # :132
        lpi = s * ie @ te.t()                              # :135
        lpt = s * te @ ie.t()                              # :136
        loss, _ = self.criterion(image_embeddings=ie, text_embeddings=te, logit_scale=s, logits_per_image=lpi,
                                 logits_per_text=lpt)     # ClassifierExperiment.py:112
        loss.backward()                                    # :115
        return {"loss": loss.detach(), "dw_image": self.head_i.layer.weight.grad, "dw_text": self.head_t.layer.weight.grad}
