"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES (build container only).

    python tests/golden/make_golden.py            # needs /root/reference (read-only); writes next to this file

mmgclip/loss/losses.py and mmgclip/networks/projection.py are loaded by file path (``import mmgclip`` itself needs
fuzzywuzzy / nltk downloads / hydra -- SURVEY.md s8c) with a one-function stub for ``sentence_transformers.util`` and,
on this GPU-less box, ``Tensor.cuda`` made an identity (the reference hard-codes ``.cuda()`` at losses.py:39,78).  The
nine arithmetic lines of mmgclip_model.py:124-136 cannot be imported (prettytable, HF downloads) and are restated
inline below, next to their line numbers.  Inputs come from a NumPy RandomState so they are reproducible anywhere;
the reference's outputs are frozen.  The .npz files and this script are committed; the reference is never copied.
"""
import importlib.util
import json
import os
import re
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("MMG_REFERENCE_ROOT", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # losses.py:39,78 call .cuda() unconditionally
    st, ut = types.ModuleType("sentence_transformers"), types.ModuleType("sentence_transformers.util")
    ut.cos_sim = lambda a, b: F.normalize(a, dim=1) @ F.normalize(b, dim=1).t()  # only used at losses.py:119
    st.util = ut
    sys.modules["sentence_transformers"] = st
    sys.modules["sentence_transformers.util"] = ut

    def load(path, name):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m

    return (load(os.path.join(REF, "mmgclip/loss/losses.py"), "ref_losses"),
            load(os.path.join(REF, "mmgclip/networks/projection.py"), "ref_projection"))


def model_forward(head_i, head_t, xi, xt, logit_scale_log):
    """mmgclip_model.py:124-136 restated (that file cannot be imported)."""
    ie = head_i(xi)                                   # :124
    te = head_t(xt)                                   # :125
    ie = ie / ie.norm(dim=1, keepdim=True)            # :128
    te = te / te.norm(dim=1, keepdim=True)            # :129
    s = logit_scale_log.exp()                         # :132
    lpi = s * ie @ te.t()                             # :135
    lpt = s * te @ ie.t()                             # :136
    return ie, te, s, lpi, lpt


def rng_inputs(rng, n, e_i, e_t):
    xi = np.maximum(1.0 + 0.35 * rng.standard_normal((n, e_i)), 0.0).astype(np.float32)
    xt = (0.5 * rng.standard_normal((n, e_t))).astype(np.float32)
    return xi, xt


def set_params(module, rng, scale=1.0):
    """Overwrite parameters with portable NumPy-generated values (state-dict order); returns them as a dict."""
    out = {}
    with torch.no_grad():
        for k, p in module.state_dict().items():
            fan_in = p.shape[-1] if p.dim() > 1 else p.shape[0]
            v = (rng.uniform(-1, 1, tuple(p.shape)) / np.sqrt(fan_in) * scale).astype(np.float32)
            if k.endswith("layer_norm.weight"):
                v = (1.0 + 0.1 * v).astype(np.float32)
            p.copy_(torch.from_numpy(v))
            out[k] = v
    return out


def grads(module):
    return {k: p.grad.detach().numpy().copy() for k, p in module.named_parameters()}


def main():
    ref_losses, ref_proj = load_reference()
    torch.set_num_threads(1)

    # ---- 1. LinearProjectionLayer + CLIPLoss, small (tensors stored) ------------------------------------------
    rng = np.random.RandomState(1234)
    n, e_i, e_t, d = 32, 96, 80, 64
    xi, xt = rng_inputs(rng, n, e_i, e_t)
    hi, ht = ref_proj.LinearProjectionLayer(e_i, d), ref_proj.LinearProjectionLayer(e_t, d)
    pi, pt = set_params(hi, rng), set_params(ht, rng)
    ls = torch.tensor(np.log(1 / 0.07), dtype=torch.float32, requires_grad=True)
    ie, te, s, lpi, lpt = model_forward(hi, ht, torch.from_numpy(xi), torch.from_numpy(xt), ls)
    loss, labels = ref_losses.CLIPLoss()(image_embeddings=ie, text_embeddings=te, logit_scale=s, logits_per_image=lpi,
                                         logits_per_text=lpt)
    ie.retain_grad(); te.retain_grad()
    loss.backward()
    np.savez(os.path.join(OUT, "clip_linear_small.npz"), xi=xi, xt=xt, w_image=pi["layer.weight"],
             w_text=pt["layer.weight"], logit_scale_log=np.float32(np.log(1 / 0.07)),
             image_embeddings=ie.detach().numpy(), text_embeddings=te.detach().numpy(), logit_scale=s.detach().numpy(),
             logits_per_image=lpi.detach().numpy(), logits_per_text=lpt.detach().numpy(), loss=loss.detach().numpy(),
             labels=labels.numpy(), d_image_embeddings=ie.grad.numpy(), d_text_embeddings=te.grad.numpy(),
             dw_image=hi.layer.weight.grad.numpy(), dw_text=ht.layer.weight.grad.numpy(),
             dlogit_scale_log=ls.grad.numpy())

    # ---- 2. BASELINE config 1 shape: B=32, 768 -> 512 (inputs regenerated from the seed; summaries stored) -----
    rng = np.random.RandomState(42)
    n, e, d = 32, 768, 512
    xi, xt = rng_inputs(rng, n, e, e)
    rngw = np.random.RandomState(43)
    wi = (rngw.uniform(-1.0, 1.0, (d, e)).astype(np.float32) / np.float32(np.sqrt(e)))
    wt = (rngw.uniform(-1.0, 1.0, (d, e)).astype(np.float32) / np.float32(np.sqrt(e)))
    hi, ht = ref_proj.LinearProjectionLayer(e, d), ref_proj.LinearProjectionLayer(e, d)
    with torch.no_grad():
        hi.layer.weight.copy_(torch.from_numpy(wi)); ht.layer.weight.copy_(torch.from_numpy(wt))
    ls = torch.tensor(np.log(1 / 0.07), dtype=torch.float32, requires_grad=True)
    ie, te, s, lpi, lpt = model_forward(hi, ht, torch.from_numpy(xi), torch.from_numpy(xt), ls)
    loss, _ = ref_losses.CLIPLoss()(logits_per_image=lpi, logits_per_text=lpt)
    loss.backward()
    gi, gt = hi.layer.weight.grad.numpy(), ht.layer.weight.grad.numpy()
    np.savez(os.path.join(OUT, "clip_cfg1_b32_768_512.npz"), seed_inputs=42, seed_weights=43, loss=loss.detach().numpy(),
             dlogit_scale_log=ls.grad.numpy(), dw_image_block=gi[:32, :32].copy(), dw_text_block=gt[:32, :32].copy(),
             dw_image_fro=np.float64(np.linalg.norm(gi.astype(np.float64))),
             dw_text_fro=np.float64(np.linalg.norm(gt.astype(np.float64))),
             dw_image_sum=np.float64(gi.astype(np.float64).sum()), dw_text_sum=np.float64(gt.astype(np.float64).sum()),
             image_embeddings_head=ie.detach().numpy()[:, :16].copy(), logits_diag=np.diag(lpi.detach().numpy()).copy())

    # ---- 3. MultiLinearHead (eval: dropout inactive) + CLIPLoss ----------------------------------------------------
    rng = np.random.RandomState(77)
    n, e_i, e_t, dims = 24, 48, 40, [56, 32]
    xi, xt = rng_inputs(rng, n, e_i, e_t)
    hi, ht = ref_proj.MultiLinearHead(e_i, dims, dropout=0.5), ref_proj.MultiLinearHead(e_t, dims, dropout=0.5)
    hi.eval(); ht.eval()
    pi, pt = set_params(hi, rng), set_params(ht, rng)
    ls = torch.tensor(np.log(1 / 0.07), dtype=torch.float32)
    ie, te, s, lpi, lpt = model_forward(hi, ht, torch.from_numpy(xi), torch.from_numpy(xt), ls)
    loss, _ = ref_losses.CLIPLoss()(logits_per_image=lpi, logits_per_text=lpt)
    loss.backward()
    blob = {"xi": xi, "xt": xt, "loss": loss.detach().numpy(), "image_embeddings": ie.detach().numpy(),
            "text_embeddings": te.detach().numpy(), "dims": np.array(dims)}
    for tag, params, head in (("i", pi, hi), ("t", pt, ht)):
        for k, v in params.items():
            blob[f"p_{tag}.{k}"] = v
        for k, v in grads(head).items():
            blob[f"g_{tag}.{k}"] = v
    np.savez(os.path.join(OUT, "clip_multilinear_small.npz"), **blob)

    # ---- 4. MLPProjectionHead (eval) + CLIPLoss ------------------------------------------------------------------
    rng = np.random.RandomState(78)
    n, e_i, e_t, d = 24, 48, 40, 32
    xi, xt = rng_inputs(rng, n, e_i, e_t)
    hi, ht = ref_proj.MLPProjectionHead(e_i, d, dropout=0.5), ref_proj.MLPProjectionHead(e_t, d, dropout=0.5)
    hi.eval(); ht.eval()
    pi, pt = set_params(hi, rng), set_params(ht, rng)
    ls = torch.tensor(np.log(1 / 0.07), dtype=torch.float32)
    ie, te, s, lpi, lpt = model_forward(hi, ht, torch.from_numpy(xi), torch.from_numpy(xt), ls)
    loss, _ = ref_losses.CLIPLoss()(logits_per_image=lpi, logits_per_text=lpt)
    loss.backward()
    blob = {"xi": xi, "xt": xt, "loss": loss.detach().numpy(), "image_embeddings": ie.detach().numpy(),
            "text_embeddings": te.detach().numpy()}
    for tag, params, head in (("i", pi, hi), ("t", pt, ht)):
        for k, v in params.items():
            blob[f"p_{tag}.{k}"] = v
        for k, v in grads(head).items():
            blob[f"g_{tag}.{k}"] = v
    np.savez(os.path.join(OUT, "clip_mlp_small.npz"), **blob)

    # ---- 5. MMGCLIPLoss on three embedding sets -----------------------------------------------------------------------
    rng = np.random.RandomState(79)
    n, d = 20, 48
    mk = lambda: torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32))  # noqa: E731
    raw = [mk().requires_grad_(True) for _ in range(3)]
    ie, te, te2 = (r / r.norm(dim=1, keepdim=True) for r in raw)
    s = torch.tensor(1 / 0.07, dtype=torch.float32, requires_grad=True)
    loss, _ = ref_losses.MMGCLIPLoss(t2t_weight=0.5)(image_embeddings=ie, text_embeddings=te, text_embeddings2=te2,
                                                     logit_scale=s)
    for t in (ie, te, te2):
        t.retain_grad()
    loss.backward()
    np.savez(os.path.join(OUT, "mmgclip_loss_small.npz"), image_embeddings=ie.detach().numpy(),
             text_embeddings=te.detach().numpy(), text_embeddings2=te2.detach().numpy(), logit_scale=s.detach().numpy(),
             loss=loss.detach().numpy(), d_image_embeddings=ie.grad.numpy(), d_text_embeddings=te.grad.numpy(),
             d_text_embeddings2=te2.grad.numpy(), d_logit_scale=s.grad.numpy())

    # ---- 6. known-answer vectors held by the reference itself --------------------------------------------------------
    # 8x8 logits literal: docstring of the commented-out AveragedBinaryCLIPLoss (losses.py:241-250); the same matrix is
    # used in notebooks/loss.ipynb (cells 13-18) with cluster labels [0,1,0,0,0,1,0,2] -> averaged CE 1.2048.
    src = open(os.path.join(REF, "mmgclip/loss/losses.py")).read().splitlines()[240:250]
    rows = [[float(v) for v in re.findall(r"-?\d+\.\d+", line)] for line in src]
    rows = [r for r in rows if len(r) == 8]
    lpi8 = np.array(rows, dtype=np.float32)
    assert lpi8.shape == (8, 8), lpi8.shape
    t8 = torch.from_numpy(lpi8)
    clip8, _ = ref_losses.CLIPLoss()(logits_per_image=t8, logits_per_text=t8.t())
    nb_labels = [0, 1, 0, 0, 0, 1, 0, 2]
    avg = ref_losses.AveragedMedicalCLIPLoss()._average_logits(t8, nb_labels)
    avg_ce = F.cross_entropy(avg, torch.tensor(nb_labels))
    # docstring example of _assign_labels (losses.py:126-139)
    doc = np.full((8, 8), -0.0237, dtype=np.float32)
    for i in range(8):
        for j in range(8):
            if (i - j) % 2 == 0:
                doc[i, j] = 1.0
    doc_labels = ref_losses.AveragedMedicalCLIPLoss()._assign_labels(torch.from_numpy(doc), threshold=0.65)
    # full AveragedMedicalCLIPLoss forward on a seeded case with duplicated texts
    rng = np.random.RandomState(80)
    base = rng.standard_normal((5, 24)).astype(np.float32)
    txt = torch.from_numpy(base[[0, 1, 0, 2, 3, 1, 4, 0]] + 0.01 * rng.standard_normal((8, 24)).astype(np.float32))
    img = torch.from_numpy(rng.standard_normal((8, 24)).astype(np.float32))
    ie = img / img.norm(dim=1, keepdim=True)
    te = txt / txt.norm(dim=1, keepdim=True)
    s = torch.tensor(1 / 0.07)
    lpi, lpt = s * ie @ te.t(), s * te @ ie.t()
    am_loss, am_labels = ref_losses.AveragedMedicalCLIPLoss(0.65)(ie, te, s, lpi, lpt)
    np.savez(os.path.join(OUT, "reference_kats.npz"), logits8=lpi8, clip_loss8=clip8.numpy(),
             notebook_labels=np.array(nb_labels), notebook_avg_logits=avg.numpy(), notebook_avg_ce=avg_ce.numpy(),
             notebook_avg_ce_printed=np.float32(1.2048), doc_cosine=doc, doc_labels=np.array(doc_labels),
             am_image_embeddings=ie.numpy(), am_text_embeddings=te.numpy(), am_logit_scale=s.numpy(),
             am_logits_per_image=lpi.numpy(), am_logits_per_text=lpt.numpy(), am_loss=am_loss.numpy(),
             am_labels=am_labels.numpy())

    # ---- 7. zero-shot scoring: torch path (mmgclip_model.py:201-209) and NumPy/SciPy twin (evaluator.py:354-368) ----
    from scipy.special import softmax as sp_softmax
    rng = np.random.RandomState(81)
    nimg, c, d = 257, 8, 64
    img = rng.standard_normal((nimg, d)).astype(np.float32)
    txt = rng.standard_normal((c, d)).astype(np.float32)
    img /= np.linalg.norm(img, axis=1, keepdims=True)
    txt /= np.linalg.norm(txt, axis=1, keepdims=True)
    txt[5] = txt[2]  # duplicated prompt: exact tie -> lowest index must win
    s = torch.tensor(np.log(1 / 0.07), dtype=torch.float32).exp()
    lg = s * torch.from_numpy(img) @ torch.from_numpy(txt).t()                 # mmgclip_model.py:135
    pr = lg.softmax(dim=-1)                                                   # mmgclip_model.py:204
    am = torch.argmax(pr, dim=-1)                                             # mmgclip_model.py:209
    s_np = s.detach().cpu().numpy()                                           # evaluator.py:354-355
    sim = s_np * img @ np.transpose(txt)                                      # evaluator.py:357
    sim = sp_softmax(sim, axis=1)                                             # evaluator.py:362
    am_np = np.argmax(sim, axis=-1)                                           # evaluator.py:368
    np.savez(os.path.join(OUT, "zeroshot_small.npz"), img=img, txt=txt, logit_scale=s.numpy(), logits=lg.numpy(),
             probs=pr.numpy(), argmax=am.numpy(), probs_numpy=sim.astype(np.float32), argmax_numpy=am_np)

    meta = {"generated_from": REF, "torch": torch.__version__, "numpy": np.__version__,
            "files": sorted(f for f in os.listdir(OUT) if f.endswith(".npz"))}
    json.dump(meta, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1)
    print(json.dumps(meta, indent=1))
    print("CLIPLoss(logits8, logits8.T) =", float(clip8), " notebook averaged CE =", float(avg_ce), " doc labels =",
          doc_labels, " AveragedMedicalCLIPLoss =", float(am_loss), am_labels.tolist())


if __name__ == "__main__":
    main()
