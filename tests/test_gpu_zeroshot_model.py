"""GPU tests: zero-shot prompt scoring (bit-exact indices), the MMGCLIP model shell / PromptClassifier boundary and
checkpoint compatibility."""
import math

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import clip_oracle as oc

pytestmark = pytest.mark.gpu


def cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def test_zeroshot_vs_reference_fixture(golden):
    from mmgclip_b200 import ops, zeroshot
    g = golden("zeroshot_small")
    out = ops.zeroshot_score(cuda(g["img"]), cuda(g["txt"]), cuda(g["logit_scale"]), k=5)
    assert out["argmax"].dtype == torch.int64
    assert np.array_equal(out["argmax"].cpu().numpy(), g["argmax"])          # bit-exact, incl. the duplicated prompt
    assert np.array_equal(out["argmax"].cpu().numpy(), g["argmax_numpy"])    # ... and vs the NumPy/SciPy twin
    assert rel_err(out["logits"].cpu(), g["logits"]) < 1e-6
    assert rel_err(out["probs"].cpu(), g["probs"]) < 1e-6
    ref = oc.closed_form_zeroshot(g["img"], g["txt"], float(g["logit_scale"]), k=5)
    assert np.array_equal(out["topk_idx"].cpu().numpy(), ref["topk_idx"])    # value desc, index asc (2 before 5)
    assert np.array_equal(out["topk_idx"][:, 0].cpu().numpy(), g["argmax"])
    probs, pred = zeroshot.zeroshot_label_prompt(g["img"], g["txt"], float(g["logit_scale"]))  # NumPy in, NumPy out
    assert np.array_equal(pred, g["argmax_numpy"]) and rel_err(probs, g["probs_numpy"]) < 1e-6


def test_zeroshot_tensor_core_path_vs_reference_fixture(golden):
    """The 3xTF32 tensor-core kernel (what large batches use) on the reference fixture: logits fp32-faithful (2e-6),
    indices bit-exact including the duplicated prompt (identical columns give identical logits -> lowest index)."""
    from mmgclip_b200 import ops
    g = golden("zeroshot_small")
    out = ops.zeroshot_score(cuda(g["img"]), cuda(g["txt"]), cuda(g["logit_scale"]), k=5, impl="tc")
    assert np.array_equal(out["argmax"].cpu().numpy(), g["argmax"])
    assert rel_err(out["logits"].cpu(), g["logits"]) < 2e-6
    assert rel_err(out["probs"].cpu(), g["probs"]) < 5e-6
    ref = oc.closed_form_zeroshot(g["img"], g["txt"], float(g["logit_scale"]), k=5)
    assert np.array_equal(out["topk_idx"].cpu().numpy(), ref["topk_idx"])
    assert np.array_equal(out["topk_idx"][:, 0].cpu().numpy(), g["argmax"])
    with pytest.raises(ValueError):
        ops.zeroshot_score(cuda(g["img"])[:, :62].contiguous(), cuda(g["txt"])[:, :62].contiguous(), 1.0, impl="tc")


@pytest.mark.parametrize("impl", ["ffma", "tc"])
@pytest.mark.parametrize("n,c,d,k", [(1, 2, 512, 0), (3, 1, 64, 1), (1000, 7, 200, 3), (4097, 64, 512, 5), (300, 64, 36, 8)])
def test_zeroshot_shapes_and_edge_cases(n, c, d, k, impl):
    from mmgclip_b200 import ops
    rng = np.random.RandomState(n + c)
    img = rng.standard_normal((n, d)).astype(np.float32); img /= np.linalg.norm(img, axis=1, keepdims=True)
    txt = rng.standard_normal((c, d)).astype(np.float32); txt /= np.linalg.norm(txt, axis=1, keepdims=True)
    s = float(np.float32(1 / 0.07))
    out = ops.zeroshot_score(cuda(img), cuda(txt), s, k=k, impl=impl)
    ref = oc.closed_form_zeroshot(img, txt, s, k=k)
    srt = np.sort(ref["logits"], axis=1)
    safe = np.ones(n, bool) if c == 1 else (srt[:, -1] - srt[:, -2]) > 1e-4   # rows without a float32 near-tie
    assert np.array_equal(out["argmax"].cpu().numpy()[safe], ref["argmax"][safe])
    assert rel_err(out["logits"].cpu(), ref["logits"]) < 1e-5
    assert rel_err(out["probs"].cpu(), ref["probs"]) < 1e-5
    if k:
        tail = srt[:, -(k + 1):] if c > k else srt
        gaps = np.min(np.diff(tail, axis=1), axis=1) if tail.shape[1] > 1 else np.full(n, np.inf)
        ok = gaps > 1e-4
        assert np.array_equal(out["topk_idx"].cpu().numpy()[ok], ref["topk_idx"][ok])
        assert (~ok).sum() <= max(2, n // 200)


def test_zeroshot_full_size_cfg4():
    """BASELINE config 4: 1M image embeddings x 64 prompts, D = 512; indices exact against float64 on sampled rows and
    everywhere self-consistent (argmax == top-1, probabilities sum to one, rows are independent of chunking)."""
    from mmgclip_b200 import ops
    n, c, d = 1 << 20, 64, 512
    gen = torch.Generator(device="cuda").manual_seed(4)
    img = torch.nn.functional.normalize(torch.randn(n, d, device="cuda", generator=gen), dim=1)
    txt = torch.nn.functional.normalize(torch.randn(c, d, device="cuda", generator=gen), dim=1)
    s = float(np.float32(1 / 0.07))
    out = ops.zeroshot_score(img, txt, s, k=5, want_logits=False)
    assert torch.equal(out["argmax"], out["topk_idx"][:, 0])
    assert (out["probs"].sum(1) - 1).abs().max().item() < 1e-5
    assert (out["topk_val"][:, :-1] >= out["topk_val"][:, 1:]).all()
    idx = torch.from_numpy(np.random.RandomState(1).choice(n, 4096, replace=False)).cuda()
    sub = ops.zeroshot_score(img[idx].contiguous(), txt, s, k=5)
    assert torch.equal(sub["argmax"], out["argmax"][idx]) and torch.equal(sub["topk_idx"], out["topk_idx"][idx])
    ref = oc.closed_form_zeroshot(img[idx].cpu().numpy(), txt.cpu().numpy(), s, k=5)
    srt = np.sort(ref["logits"], axis=1)
    ok = np.min(np.diff(srt[:, -6:], axis=1), axis=1) > 2e-5
    assert ok.mean() > 0.98
    assert np.array_equal(sub["topk_idx"].cpu().numpy()[ok], ref["topk_idx"][ok])
    assert np.array_equal(sub["argmax"].cpu().numpy()[ok], ref["argmax"][ok])


class FakeTextEncoder(torch.nn.Module):
    """Stands in for BertEncoder (encoder.py:121-156): tokens -> [n, seq, H]."""
    model_output_dimension = 48

    def __init__(self):
        super().__init__()
        self.emb = torch.nn.Embedding(100, 48)

    def forward(self, tokens):
        return self.emb(tokens["input_ids"])


class FakeTokenizer:
    def __call__(self, texts, padding=None, truncation=None, return_tensors=None, max_length=16):
        ids = torch.zeros((len(texts), max_length), dtype=torch.long)
        mask = torch.zeros((len(texts), max_length), dtype=torch.long)
        for i, t in enumerate(texts):
            toks = [(ord(ch) % 97) + 1 for ch in t][:max_length]
            ids[i, :len(toks)] = torch.tensor(toks)
            mask[i, :len(toks)] = 1
        return {"input_ids": ids, "attention_mask": mask}


def make_config(head="LinearProjectionLayer", out=32, loss="CLIPLoss"):
    from mmgclip_b200.model import as_config
    return as_config({
        "networks": {"image_encoder": {"name": "ConvNextTiny", "image_features_dimension": 64},
                     "text_encoder": {"name": "BertEncoder"}, "dropout": {"config": {"dropout": 0.0}},
                     "logit_temperature": 0.07},
        "projection": {"config": {"projection_name": head, "output_projection_dimension": out}},
        "loss": {"config": {"loss_name": loss}},
        "tokenizer": {"config": {"tokenizer_name": "fake", "sequence_length": 16}},
    })


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_model_shell_forward_contract_and_training_step(prec):
    from mmgclip_b200.loss_controller import create_loss
    from mmgclip_b200.model import MMGCLIP
    torch.manual_seed(0)
    model = MMGCLIP(make_config(), text_encoder=FakeTextEncoder(), precision=prec)
    tok = FakeTokenizer()
    texts = [f"birads score of {i} mass shape {i * 7}" for i in range(12)]
    batch = {"image_features": torch.randn(12, 1, 64, 1, 1).abs(), "text_tokens": tok(texts, max_length=16)}
    model.train()
    out = model(batch)
    assert list(out) == ["image_embeddings", "text_embeddings", "logit_scale", "logits_per_image", "logits_per_text"]
    assert out["image_embeddings"].shape == (12, 32) and out["logits_per_image"].shape == (12, 12)
    assert abs(out["logit_scale"].item() - 1 / 0.07) < 1e-4
    assert "logit_scale" not in model.state_dict()          # quirk Q1: not a Parameter on CUDA
    assert sorted(model.state_dict()) == ["image_projection_layer.layer.weight", "text_encoder.emb.weight",
                                          "text_projection_layer.layer.weight"]
    criterion = create_loss("CLIPLoss")().cuda()
    criterion.precision = prec
    loss, labels = criterion(**out)
    loss.backward()
    # same numbers from plain torch on the same features (eos pooling + heads + loss restated by the oracle)
    xi = torch.flatten(batch["image_features"], 1)
    hidden = model.text_encoder({k: v.cuda() for k, v in batch["text_tokens"].items()}).detach().cpu()
    last = batch["text_tokens"]["attention_mask"].sum(-1) - 1
    xt = hidden[torch.arange(12), last]
    ref = oc.torch_train_step(xi, xt, model.image_projection_layer.layer.weight.detach().cpu(),
                              model.text_projection_layer.layer.weight.detach().cpu(), torch.tensor(math.log(1 / 0.07)))
    tol = {"fp32": 1e-5, "bf16": 2e-3}[prec]
    assert abs(loss.item() - ref["loss"].item()) < tol * ref["loss"].item()
    # tiny model (n=12, D=32): few terms to average the bf16 operand rounding over -> 1.5e-2 on the weight gradient
    assert rel_err(model.image_projection_layer.layer.weight.grad.cpu(), ref["dw_image"]) < (1e-5 if prec == "fp32" else 1.5e-2)
    lpi, lpt = oc.torch_logits(ref["image_embeddings"], ref["text_embeddings"], torch.tensor(math.log(1 / 0.07)).exp())
    assert rel_err(out["logits_per_image"].detach().cpu(), lpi) < 2 * tol
    assert rel_err(out["logits_per_text"].detach().cpu(), lpt) < 2 * tol
    # the literal reference signature (logits only) gives the same loss
    loss2, _ = create_loss("CLIPLoss")()(out["logits_per_image"], out["logits_per_text"])
    assert abs(loss2.item() - loss.item()) < 2 * tol * loss.item()
    # non-square evaluation forward (validate(): prompts as text side, ClassifierExperiment.py:192-229)
    model.eval()
    with torch.no_grad():
        ev = model({"image_features": batch["image_features"], "text_tokens": tok(["benign", "malignant"], max_length=16)})
    assert ev["logits_per_image"].shape == (12, 2) and ev["logits_per_text"].shape == (2, 12)


def test_mmgclip_loss_branch_and_checkpoint_flavours():
    from mmgclip_b200.loss_controller import create_loss
    from mmgclip_b200.model import MMGCLIP
    torch.manual_seed(1)
    cfg = make_config(head="MultiLinearHead", out=[40, 24], loss="MMGCLIPLoss")
    model = MMGCLIP(cfg, text_encoder=None, trainable_logit_scale=True, precision="fp32")
    assert "logit_scale" in model.state_dict()
    batch = {"image_features": torch.randn(10, 1, 64, 1, 1), "text_features": torch.randn(10, 768),
             "text_features2": torch.randn(10, 768)}
    model.train()
    out = model(batch)
    assert "text_embeddings2" in out
    loss, _ = create_loss("MMGCLIPLoss")()(**out)
    loss.backward()
    assert model.logit_scale.grad is not None and torch.isfinite(model.logit_scale.grad)
    assert "text_embeddings2" not in model(batch, validation=True)
    # checkpoints: CPU-trained reference (has logit_scale) -> device-faithful model, and the reverse
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sd["logit_scale"] = torch.tensor(3.0)
    plain = MMGCLIP(cfg, precision="fp32")
    plain.load_state_dict(sd)
    assert abs(plain.logit_scale.item() - 3.0) < 1e-6 and "logit_scale" not in plain.state_dict()
    model.load_state_dict(plain.state_dict())  # checkpoint without logit_scale into the trainable flavour


def test_prompt_classifier_contract():
    from mmgclip_b200.model import MMGCLIP, PromptClassifier
    torch.manual_seed(2)
    model = MMGCLIP(make_config(), text_encoder=FakeTextEncoder(), precision="fp32")
    clf = PromptClassifier(model, tokenizer=FakeTokenizer())
    feats = torch.randn(1, 64)
    classes = ["BIRADS unknown."] + [f"BIRADS score of {i}." for i in range(7)]
    res = clf(image_features=feats, class_list=classes, visualize=False)
    assert set(res) == {"classes_similarities", "similarities_argmax", "class_list"}
    assert res["classes_similarities"].shape == (1, 8) and isinstance(res["similarities_argmax"], int)
    assert abs(res["classes_similarities"].sum().item() - 1) < 1e-5
    with torch.no_grad():
        out = model({"image_features": feats, "text_tokens": FakeTokenizer()(classes, max_length=16)})
    probs = out["logits_per_image"].softmax(-1)
    assert res["similarities_argmax"] == int(torch.argmax(probs, -1)[0].item())
    assert rel_err(res["classes_similarities"].cpu(), probs.cpu()) < 1e-5
