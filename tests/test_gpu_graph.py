"""GraphedStep: a whole captured step (heads -> normalise -> fused InfoNCE -> backward) replays to the same loss and
head-weight gradients as the eager step, for every input set, and follows the data in its static input buffers."""
import math

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import clip_oracle as oc

pytestmark = pytest.mark.gpu


def _make(prec):
    from mmgclip_b200.losses import CLIPLoss
    from mmgclip_b200.projection import LinearProjectionLayer
    wi, wt = oc.synthetic_head_weights(128, 96, 80, seed=5)
    hi = LinearProjectionLayer(96, 128, precision=prec).cuda()
    ht = LinearProjectionLayer(80, 128, precision=prec).cuda()
    with torch.no_grad():
        hi.layer.weight.copy_(torch.from_numpy(wi)); ht.layer.weight.copy_(torch.from_numpy(wt))
    scale = torch.tensor(math.log(1 / 0.07), device="cuda").exp()
    crit = CLIPLoss(precision=prec)

    def step(xi, xt):
        hi.layer.weight.grad = None
        ht.layer.weight.grad = None
        ie, te = hi.forward_normalized(xi), ht.forward_normalized(xt)
        loss, _ = crit(image_embeddings=ie, text_embeddings=te, logit_scale=scale)
        loss.backward()
        return loss

    return hi, ht, step, (wi, wt)


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_graphed_step_matches_eager_and_oracle(prec):
    from mmgclip_b200.graph import GraphedStep
    hi, ht, step, (wi, wt) = _make(prec)
    sets = []
    for seed in (11, 12):
        xi, xt = oc.synthetic_features(300, 96, 80, seed=seed)
        sets.append((torch.from_numpy(xi).cuda(), torch.from_numpy(xt).cuda()))
    eager = []
    for xi, xt in sets:
        l = step(xi, xt)
        eager.append((l.item(), hi.layer.weight.grad.clone(), ht.layer.weight.grad.clone()))
    # a live loss from an eager step keeps its autograd graph -- and the parameters' AccumulateGrad nodes, which are
    # bound to the (legacy) stream they were created on -- alive; capture needs them re-created on the capture stream
    del l
    g = GraphedStep(step, sets, params=list(hi.parameters()) + list(ht.parameters()))
    assert len(g) == 2 and g.kernel_launches > 0
    for rep in range(2):
        for i in range(2):
            l = g(i)
            torch.cuda.synchronize()
            assert abs(l.item() - eager[i][0]) <= 1e-6 * abs(eager[i][0]) + 1e-7
            assert rel_err(hi.layer.weight.grad.cpu(), eager[i][1].cpu()) < 1e-5
            assert rel_err(ht.layer.weight.grad.cpu(), eager[i][2].cpu()) < 1e-5
    # new data copied into a static input set is what the replay computes on
    xi, xt = oc.synthetic_features(300, 96, 80, seed=13)
    g.copy_in(0, torch.from_numpy(xi).cuda(), torch.from_numpy(xt).cuda())
    l = g(0)
    torch.cuda.synchronize()
    ref = oc.closed_form_train_step(xi, xt, wi, wt, math.log(1 / 0.07))
    tol = 2e-3 if prec == "bf16" else 1e-5
    assert abs(l.item() - ref["loss"]) / ref["loss"] < tol
    assert rel_err(hi.layer.weight.grad.double().cpu().numpy(), ref["dw_image"]) < (8e-3 if prec == "bf16" else 1e-5)


def test_graphed_step_rejects_host_tensors():
    from mmgclip_b200.graph import GraphedStep
    with pytest.raises(RuntimeError):
        GraphedStep(lambda x: x, [(torch.zeros(4),)])
    with pytest.raises(ValueError):
        GraphedStep(lambda x: x, [])
