"""Round-2 additions to tests/golden (build container only; needs /root/reference, read-only):

    python tests/golden/make_golden_r2.py

* reference_configs.json -- the parsed contents of the reference's Hydra groups configs/projection/*.yaml and
  configs/loss/*.yaml (the construction contract of the hot path), so that the config-driven
  construction tests also run where /root/reference does not exist (the GPU box).
* clip_multilinear_768_512.npz -- the reference's MultiLinearHead at the SHIPPED shape of configs/projection/2xLinear512.yaml
  ([768, 512] on 768-d features, eval mode) + CLIPLoss forward/backward, executed from the reference's own projection.py /
  losses.py exactly like make_golden.py does.  Inputs and parameters are regenerated from seeds by the test (they would
  be 8 MB); outputs and gradient summaries are stored.
"""
import glob
import json
import os

import numpy as np
import torch
import yaml

from make_golden import OUT, REF, grads, load_reference, model_forward, rng_inputs, set_params


def main():
    cfgs = {}
    for group in ("projection", "loss"):
        for path in sorted(glob.glob(os.path.join(REF, "configs", group, "*.yaml"))):
            with open(path) as f:
                cfgs[f"{group}/{os.path.basename(path)}"] = yaml.safe_load(f)
    with open(os.path.join(OUT, "reference_configs.json"), "w") as f:
        json.dump(cfgs, f, indent=1, sort_keys=True)

    ref_losses, ref_proj = load_reference()
    torch.set_num_threads(1)
    rng = np.random.RandomState(2025)
    n, e, dims = 48, 768, [768, 512]
    xi, xt = rng_inputs(rng, n, e, e)
    hi, ht = ref_proj.MultiLinearHead(e, dims, dropout=0.5), ref_proj.MultiLinearHead(e, dims, dropout=0.5)
    hi.eval(); ht.eval()
    set_params(hi, rng); set_params(ht, rng)
    ls = torch.tensor(np.log(1 / 0.07), dtype=torch.float32)
    ie, te, s, lpi, lpt = model_forward(hi, ht, torch.from_numpy(xi), torch.from_numpy(xt), ls)
    loss, _ = ref_losses.CLIPLoss()(logits_per_image=lpi, logits_per_text=lpt)
    loss.backward()
    blob = {"seed": 2025, "n": n, "dims": np.array(dims), "loss": loss.detach().numpy(),
            "image_embeddings": ie.detach().numpy(), "text_embeddings": te.detach().numpy()}
    for tag, head in (("i", hi), ("t", ht)):
        for k, g in grads(head).items():
            g64 = g.astype(np.float64)
            blob[f"g_{tag}.{k}.fro"] = np.float64(np.linalg.norm(g64))
            blob[f"g_{tag}.{k}.block"] = (g[:24, :24] if g.ndim == 2 else g[:64]).copy()
    np.savez(os.path.join(OUT, "clip_multilinear_768_512.npz"), **blob)
    print("wrote reference_configs.json, clip_multilinear_768_512.npz; loss =", float(loss))


if __name__ == "__main__":
    main()
