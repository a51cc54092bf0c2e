"""Host-side mirror of the reference's plugin interface: names, constructor signatures, state-dict keys, error
behaviour -- everything that can be checked without launching a kernel (reference: projection_controller.py:3-24,
loss_controller.py:3-23, projection.py, losses.py)."""
import inspect

import pytest
import torch

from mmgclip_b200 import ops
from mmgclip_b200.loss_controller import create_loss
from mmgclip_b200.losses import AveragedMedicalCLIPLoss, CLIPLoss, MMGCLIPLoss
from mmgclip_b200.model import _cfg, as_config
from mmgclip_b200.projection import LinearProjectionLayer, MLPProjectionHead, MultiLinearHead
from mmgclip_b200.projection_controller import get_projection_head


def test_projection_controller_names():
    assert get_projection_head("LinearProjectionLayer") is LinearProjectionLayer
    assert get_projection_head("MultiLinearHead") is MultiLinearHead
    assert get_projection_head("MLPProjectionHead") is MLPProjectionHead
    for bad in ("ZeroProjection", "ProjectionHead", "nope", "ops"):
        with pytest.raises(ValueError, match=f"Invalid network_name: {bad}"):
            get_projection_head(bad)


def test_loss_controller_names():
    assert create_loss("CLIPLoss") is CLIPLoss
    assert create_loss("MMGCLIPLoss") is MMGCLIPLoss
    assert create_loss("AveragedMedicalCLIPLoss") is AveragedMedicalCLIPLoss
    with pytest.raises(ValueError, match="Invalid network_name: Foo"):
        create_loss("Foo")
    # ClassifierExperiment.py:70 builds the loss with no arguments
    assert isinstance(create_loss("CLIPLoss")(), CLIPLoss)
    assert create_loss("MMGCLIPLoss")().t2t_weight == 0.5
    assert create_loss("AveragedMedicalCLIPLoss")().similarity_threshold == 0.65


def test_state_dict_keys_and_shapes_match_reference_layout():
    h = LinearProjectionLayer(embedding_dim=768, projection_dim=512, dropout=0.5)
    assert {k: tuple(v.shape) for k, v in h.state_dict().items()} == {"layer.weight": (512, 768)}
    assert all(p.requires_grad for p in h.parameters())
    m = MultiLinearHead(embedding_dim=768, projection_dim=[768, 512], dropout=0.2)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {
        "layers.0.weight": (768, 768), "layers.0.bias": (768,), "layers.1.weight": (512, 768), "layers.1.bias": (512,)}
    assert m.dropout.p == 0.2
    p = MLPProjectionHead(embedding_dim=768, projection_dim=256)
    assert {k: tuple(v.shape) for k, v in p.state_dict().items()} == {
        "projection.weight": (256, 768), "projection.bias": (256,), "fc.weight": (256, 256), "fc.bias": (256,),
        "layer_norm.weight": (256,), "layer_norm.bias": (256,)}


def test_constructor_signatures():
    sig = inspect.signature(LinearProjectionLayer.__init__).parameters
    assert list(sig)[:4] == ["self", "embedding_dim", "projection_dim", "dropout"]
    assert sig["projection_dim"].default == 512 and sig["dropout"].default == 0
    sig = inspect.signature(MultiLinearHead.__init__).parameters
    assert sig["projection_dim"].default == [] and sig["dropout"].default == 0.5
    with pytest.raises(TypeError):                      # quirk Q6: an int projection_dim is not subscriptable
        MultiLinearHead(embedding_dim=768, projection_dim=512)
    sig = inspect.signature(MMGCLIPLoss.forward).parameters
    assert list(sig)[:5] == ["self", "image_embeddings", "text_embeddings", "text_embeddings2", "logit_scale"]
    sig = inspect.signature(CLIPLoss.forward).parameters
    assert list(sig)[:3] == ["self", "logits_per_image", "logits_per_text"] and "kwargs" in sig


def test_no_cpu_fallback():
    x = torch.randn(4, 8)
    w = torch.randn(6, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.linear(x, w)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.info_nce(x, x, 10.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.zeroshot_score(x, x, 10.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        LinearProjectionLayer(8, 6)(x)
    with pytest.raises(ValueError):
        ops.set_default_precision("fp16")


def test_config_helpers():
    cfg = as_config({"projection": {"config": {"projection_name": "LinearProjectionLayer",
                                               "output_projection_dimension": [768, 512]}}})
    assert cfg.projection.config.projection_name == "LinearProjectionLayer"
    assert _cfg(cfg, "projection.config.output_projection_dimension") == [768, 512]
    assert _cfg(cfg, "networks.logit_temperature", 0.07) == 0.07
    assert _cfg({"a": {"b": 3}}, "a.b") == 3


def test_assign_labels_host_logic(golden):
    k = golden("reference_kats")
    loss = AveragedMedicalCLIPLoss()
    assert loss._assign_labels(k["doc_cosine"].tolist(), threshold=0.65) == [0, 1, 0, 1, 0, 1, 0, 1]
    assert loss._assign_labels(torch.from_numpy(k["doc_cosine"]), threshold=0.65) == k["doc_labels"].tolist()


def test_bench_synthetic_recipe_matches_oracle():
    """bench.py generates the measured arm's inputs itself (nothing under oracle/ on the product path); the CPU arm uses
    the oracle's generator.  Both must produce the same arrays so that the two arms see identical data."""
    import importlib.util
    import os
    import numpy as np
    from oracle import clip_oracle as oc
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    a, b = bench.synthetic_features(64, 48, 40, seed=42), oc.synthetic_features(64, 48, 40, seed=42)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    a, b = bench.synthetic_head_weights(32, 48, 40, seed=43), oc.synthetic_head_weights(32, 48, 40, seed=43)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
