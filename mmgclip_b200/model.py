"""Model shell around the hot path -- ``MMGCLIP`` and ``PromptClassifier`` with the reference's call signatures and
output-dict keys (mmgclip/networks/mmgclip_model.py:12-166, 168-249).

What is in scope here is the arithmetic between the encoders' features and the loss: projection heads, L2
normalisation, ``exp(logit_scale)`` and the similarity logits.  The encoders themselves (BERT, ConvNeXt, ResNet --
SURVEY.md s2 rows 9-10) are upstream feature producers and are *injected*: pass any ``nn.Module`` with the reference's
encoder interface (``forward(tokens) -> [n, seq, H]`` and ``.model_output_dimension``), or feed pre-pooled features.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Any, Optional

import torch
from torch import nn

from . import ops
from .projection import LinearProjectionLayer
from .projection_controller import get_projection_head


def as_config(d: Any) -> Any:
    """dict -> attribute-style config (the reference wraps its Hydra config in AttrDict, train.py:14)."""
    if isinstance(d, dict):
        return SimpleNamespace(**{k: as_config(v) for k, v in d.items()})
    if isinstance(d, (list, tuple)):
        return type(d)(as_config(v) for v in d)
    return d


def _cfg(obj: Any, path: str, default: Any = None) -> Any:
    for key in path.split("."):
        if obj is None:
            return default
        obj = obj.get(key, None) if isinstance(obj, dict) else getattr(obj, key, None)
    return default if obj is None else obj


class MMGCLIP(nn.Module):
    """Drop-in for the reference ``MMGCLIP``: ``forward(batch, **kwargs) -> dict`` with keys ``image_embeddings``,
    ``text_embeddings``, ``logit_scale`` (already exponentiated), ``logits_per_image``, ``logits_per_text`` and, for
    ``MMGCLIPLoss`` during training, ``text_embeddings2`` (mmgclip_model.py:146-164).

    Differences that are deliberate and documented:
      * Q1 -- on CUDA the reference's ``logit_scale`` silently stops being a Parameter.  ``trainable_logit_scale=False``
        (default) reproduces that: a constant tensor, absent from ``parameters()``/``state_dict()``.  ``True`` registers
        it (what the reference does on CPU).  ``load_state_dict`` accepts checkpoints of either flavour.
      * ``logits_per_image`` / ``logits_per_text`` are materialised when they are cheap or needed (evaluation,
        ``validation=True``, n != m, n*m <= ``materialize_logits_below``, or a configured loss other than ``CLIPLoss`` /
        ``MMGCLIPLoss``); for large paired training batches under the fused losses they are ``None`` and the loss
        consumes the embeddings instead -- the B x B matrix is never built.
    """

    def __init__(self, config=None, text_encoder: Optional[nn.Module] = None, image_encoder: Optional[nn.Module] = None,
                 trainable_logit_scale: bool = False, precision: Optional[str] = None,
                 materialize_logits_below: int = 1 << 20):
        super().__init__()
        assert config is not None, 'Error in initializing the model. Missing training config object.'
        self.config = config
        if not torch.cuda.is_available():
            raise RuntimeError("mmgclip_b200.MMGCLIP needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda")
        self.precision = precision
        self.materialize_logits_below = materialize_logits_below

        self.image_encoder = image_encoder.to(self.device) if image_encoder is not None else None
        self.text_encoder = text_encoder.to(self.device) if text_encoder is not None else None

        name = _cfg(config, "projection.config.projection_name")
        image_dim = _cfg(config, "networks.image_encoder.image_features_dimension")
        text_dim = getattr(self.text_encoder, "model_output_dimension", None) or _cfg(
            config, "networks.text_encoder.model_output_dimension", 768)
        if name != "ZeroProjection":
            out_dim = _cfg(config, "projection.config.output_projection_dimension")
            dropout = _cfg(config, "networks.dropout.config.dropout", 0)
            head = get_projection_head(name)
            self.image_projection_layer = head(embedding_dim=image_dim, projection_dim=out_dim, dropout=dropout).to(self.device)
            self.text_projection_layer = head(embedding_dim=text_dim, projection_dim=out_dim, dropout=dropout).to(self.device)
            for h in (self.image_projection_layer, self.text_projection_layer):
                h.precision = precision
        else:
            self.image_projection_layer = None
            self.text_projection_layer = None

        init = torch.ones([]) * math.log(1 / _cfg(config, "networks.logit_temperature", 0.07))
        self.trainable_logit_scale = trainable_logit_scale
        if trainable_logit_scale:
            self.logit_scale = nn.Parameter(init.to(self.device))
        else:
            self.logit_scale = init.to(self.device)  # plain tensor: what `.to('cuda')` leaves behind in the reference

    # -- checkpoints -------------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        sd = dict(state_dict)
        if not self.trainable_logit_scale and "logit_scale" in sd:
            self.logit_scale = sd.pop("logit_scale").detach().to(self.device, torch.float32).reshape(())
        if self.trainable_logit_scale and "logit_scale" not in sd:
            sd["logit_scale"] = self.logit_scale.detach().clone()
        return super().load_state_dict(sd, strict=strict, **kw)

    def count_parameters(self, model=None):
        model = model if model is not None else self
        return sum(p.numel() for p in model.parameters() if p.requires_grad)

    # -- feature producers (out of scope arithmetic; interface kept) -------------------------------------------------
    def encode_images(self, batch):
        """[n, 1, F, 1, 1] (or any [n, ...]) pre-extracted features -> [n, F] on the device (mmgclip_model.py:76-93)."""
        flat = torch.flatten(batch['image_features'].to(self.device), 1)
        if self.image_encoder is not None and _cfg(self.config, "networks.image_encoder.name") == "ResNet50Encoder":
            return self.image_encoder(flat)
        return flat

    def encode_text(self, batch, text_pooling='eos'):
        """Text encoder + 'eos' pooling = hidden state of the last attended token (mmgclip_model.py:95-115).
        A batch may instead carry pre-pooled ``text_features`` [n, H] (synthetic-feature benchmarks, cached encoders)."""
        if 'text_features' in batch and batch['text_features'] is not None:
            return batch['text_features'].to(self.device)
        if self.text_encoder is None:
            raise RuntimeError("no text encoder was injected and the batch has no 'text_features'")
        tokens = batch['text_tokens']
        tokens = tokens.to(self.device) if hasattr(tokens, "to") else {k: v.to(self.device) for k, v in tokens.items()}
        hidden = self.text_encoder(tokens)
        if text_pooling != 'eos':
            raise NotImplementedError(f"{text_pooling} method is not implemented...")
        return ops.eos_pool(hidden.to(torch.float32), tokens['attention_mask'])

    # -- the hot path ------------------------------------------------------------------------------------------
    def _embed(self, head, features):
        features = features.to(torch.float32)
        if head is None:
            return ops.l2_normalize(features, prec=self.precision)
        if isinstance(head, LinearProjectionLayer):
            return head.forward_normalized(features)  # projection + normalise fused
        return ops.l2_normalize(head(features), prec=self.precision)

    def forward(self, batch, **kwargs):
        image_features = self.encode_images(batch)
        text_features = self.encode_text(batch, text_pooling='eos')

        image_embeddings = self._embed(self.image_projection_layer, image_features)
        text_embeddings = self._embed(self.text_projection_layer, text_features)
        logit_scale = self.logit_scale.exp()

        n, m = image_embeddings.shape[0], text_embeddings.shape[0]
        validation = kwargs.get('validation', False) is True
        # only the fused losses can do without the [n, m] logits; any other criterion (AveragedMedicalCLIPLoss, a caller's
        # own) reads them, as does reference-side code that looks at outputs['logits_per_image'] in the train loop
        fused_loss = _cfg(self.config, "loss.config.loss_name") in ("CLIPLoss", "MMGCLIPLoss")
        materialise = ((not self.training) or validation or n != m or n * m <= self.materialize_logits_below
                       or not fused_loss)
        logits_per_image = logits_per_text = None
        if materialise:
            logits_per_image = ops.similarity_logits(image_embeddings, text_embeddings, logit_scale, prec=self.precision)
            logits_per_text = ops.similarity_logits(text_embeddings, image_embeddings, logit_scale, prec=self.precision)

        output = {
            "image_embeddings": image_embeddings,
            "text_embeddings": text_embeddings,
            "logit_scale": logit_scale,
            "logits_per_image": logits_per_image,
            "logits_per_text": logits_per_text,
        }

        if _cfg(self.config, "loss.config.loss_name") == "MMGCLIPLoss" and not validation:
            # second text view (the report's impression section) for the text<->text term (mmgclip_model.py:154-164)
            if batch.get('text_features2') is not None:
                text_features2 = batch['text_features2'].to(self.device)
            else:
                second = dict(batch)
                second['text_tokens'] = batch['image_impression_tokens']
                second['text_features'] = None
                text_features2 = self.encode_text(second, text_pooling='eos')
            output['text_embeddings2'] = self._embed(self.text_projection_layer, text_features2)
        return output


class PromptClassifier(nn.Module):
    """Zero-shot classifier over a list of prompts (mmgclip_model.py:168-249): softmax of the image->prompt logits and
    its argmax for the first image.  ``tokenizer`` is injected (the reference builds a HuggingFace AutoTokenizer, which
    needs network access); it must be callable as ``tokenizer(class_list, padding=..., truncation=..., return_tensors="pt",
    max_length=...)``."""

    def __init__(self, model=None, tokenizer=None):
        super().__init__()
        self.model = model
        self.device = torch.device("cuda")
        self.tokenizer = tokenizer

    def forward(self, image_features, class_list, visualize=True, image_id=None, ground_truth=None):
        if self.tokenizer is None:
            raise RuntimeError("PromptClassifier needs a tokenizer")
        seq_len = _cfg(self.model.config, "tokenizer.config.sequence_length", 256)
        inputs = {
            "image_features": image_features,
            "text_tokens": self.tokenizer(class_list, padding="max_length", truncation=True, return_tensors="pt",
                                          max_length=seq_len),
        }
        self.model.eval()
        with torch.no_grad():
            out = self.model(inputs)
            scored = ops.zeroshot_score(out["image_embeddings"], out["text_embeddings"], out["logit_scale"])
        classes_similarities = scored["probs"]
        outputs = {
            "classes_similarities": classes_similarities,
            "similarities_argmax": scored["argmax"][0].item(),
            "class_list": class_list,
        }
        if visualize:
            # the reference's default (mmgclip_model.py:188,213-247) draws a matplotlib bar chart; plotting is outside the
            # accelerated path, so the call keeps its signature and its assertion, returns the same dict and says so
            assert image_id is not None, "For visualizing results, image_id value is required."
            import warnings
            warnings.warn("mmgclip_b200.PromptClassifier: visualize=True does not plot (plotting is out of scope); plot "
                          "outputs['classes_similarities'] against outputs['class_list'] on the caller's side",
                          stacklevel=2)
        return outputs
