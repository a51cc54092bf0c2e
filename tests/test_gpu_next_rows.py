"""SURVEY s8(f) rows built after the hot path: 'eos' text pooling (N1) and the fused AdamW step for the head weights
(N4), each against the oracle's statement of the reference lines (mmgclip_model.py:108-111;
ClassifierExperiment.py:74,118 = torch.optim.AdamW).  Pooling is a copy: bit-exact.  AdamW is fp32 arithmetic in
torch's operation order: 1e-6 relative after several steps (tolerance written below)."""
import math

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import clip_oracle as oc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,seq,H", [(1, 1, 4), (7, 33, 5), (64, 256, 768), (300, 19, 1024)])
def test_eos_pool_bit_exact_and_backward(n, seq, H):
    from mmgclip_b200 import ops
    g = torch.Generator().manual_seed(n * 131 + seq)
    hidden = torch.randn(n, seq, H, generator=g)
    lens = torch.randint(0, seq + 1, (n,), generator=g)  # 0 = empty mask -> wraps to the last position
    lens[0] = seq
    mask = (torch.arange(seq)[None, :] < lens[:, None]).to(torch.int64)
    want = oc.torch_eos_pool(hidden, mask)
    h = hidden.cuda().requires_grad_(True)
    got, idx = ops.eos_pool(h, mask.cuda(), return_index=True)
    assert torch.equal(got.cpu(), want)
    assert torch.equal(idx.cpu(), torch.where(lens > 0, lens - 1, torch.full_like(lens, seq - 1)))
    # int32 / bool masks (tokenizers differ) are accepted
    assert torch.equal(ops.eos_pool(h.detach(), mask.cuda().to(torch.int32)).cpu(), want)
    # backward = scatter of the pooled rows' gradients
    dy = torch.randn(n, H, generator=g)
    got.backward(dy.cuda())
    href = hidden.clone().requires_grad_(True)
    oc.torch_eos_pool(href, mask).backward(dy)
    assert torch.equal(h.grad.cpu(), href.grad)


def test_eos_pool_rejects_bad_input():
    from mmgclip_b200 import ops
    with pytest.raises(RuntimeError):
        ops.eos_pool(torch.zeros(2, 3, 4), torch.ones(2, 3, dtype=torch.int64))  # host tensors: no CPU path
    with pytest.raises(ValueError):
        ops.eos_pool(torch.zeros(2, 3, 4, device="cuda"), torch.ones(2, 4, dtype=torch.int64, device="cuda"))


def test_model_encode_text_uses_eos_pool():
    from mmgclip_b200.model import MMGCLIP, as_config

    class Enc(torch.nn.Module):
        model_output_dimension = 48

        def __init__(self):
            super().__init__()
            self.emb = torch.nn.Embedding(100, 48)

        def forward(self, tokens):
            return self.emb(tokens["input_ids"])

    cfg = as_config({"projection": {"config": {"projection_name": "LinearProjectionLayer",
                                               "output_projection_dimension": 32}},
                     "networks": {"image_encoder": {"image_features_dimension": 40, "name": "ConvNeXt"},
                                  "dropout": {"config": {"dropout": 0}}, "logit_temperature": 0.07},
                     "loss": {"config": {"loss_name": "CLIPLoss"}}})
    torch.manual_seed(3)
    model = MMGCLIP(cfg, text_encoder=Enc())
    ids = torch.randint(0, 100, (6, 11))
    mask = (torch.arange(11)[None, :] < torch.tensor([11, 3, 1, 7, 5, 2])[:, None]).to(torch.int64)
    pooled = model.encode_text({"text_tokens": {"input_ids": ids, "attention_mask": mask}})
    want = oc.torch_eos_pool(model.text_encoder.emb.weight.detach().cpu()[ids], mask)
    assert torch.equal(pooled.detach().cpu(), want)


def _problem(shapes, steps, seed):
    g = torch.Generator().manual_seed(seed)
    params = [torch.randn(*s, generator=g) for s in shapes]
    grads = [[torch.randn(*s, generator=g) * 0.05 for s in shapes] for _ in range(steps)]
    return params, grads


ADAMW_RTOL, ADAMW_ATOL = 2e-6, 2e-7


@pytest.mark.parametrize("shapes", [
    [(512, 768)],                                                   # one head weight
    [(768, 768), (768,), (512, 768), (512,), (1,), (3, 5), (1027,)],  # MultiLinearHead-like + odd sizes (tail paths)
    [(9, 7)] * 70 + [(2051,)],                                      # > 64 tensors: two launches, one step advance
])
def test_fused_adamw_matches_torch_adamw(shapes):
    from mmgclip_b200.optim import FusedAdamW
    steps = 7
    params, grads = _problem(shapes, steps, seed=len(shapes))
    lrs = [1e-3 * (1 + 0.5 * math.sin(t)) for t in range(steps)]
    want = oc.torch_adamw_steps(params, grads, lrs, weight_decay=1e-2)
    want64 = oc.closed_form_adamw_steps([p.numpy() for p in params], [[x.numpy() for x in gs] for gs in grads], lrs, 1e-2)
    ps = [torch.nn.Parameter(p.clone().cuda()) for p in params]
    opt = FusedAdamW(ps, lr=lrs[0], weight_decay=1e-2)
    for t in range(steps):
        for grp in opt.param_groups:
            grp["lr"] = lrs[t]
        for p, g in zip(ps, grads[t]):
            p.grad = g.cuda()
        opt.step()
    assert opt.steps_taken() == steps
    assert opt.kernel_launches == steps * (2 if len(shapes) > 64 else 1)
    for p, w, w64 in zip(ps, want, want64):
        np.testing.assert_allclose(p.detach().cpu().numpy(), w.numpy(), rtol=ADAMW_RTOL, atol=ADAMW_ATOL)
        np.testing.assert_allclose(p.detach().cpu().numpy(), w64, rtol=2e-5, atol=2e-6)


def test_fused_adamw_state_dict_interchange_with_torch():
    """A torch.optim.AdamW checkpoint continues under FusedAdamW (and the other way round) as if nothing happened."""
    from mmgclip_b200.optim import FusedAdamW
    shapes = [(64, 48), (48,)]
    params, grads = _problem(shapes, 6, seed=9)
    want = oc.torch_adamw_steps(params, grads, 2e-3, weight_decay=1e-4)
    ps_t = [torch.nn.Parameter(p.clone().cuda()) for p in params]
    topt = torch.optim.AdamW(ps_t, lr=2e-3, weight_decay=1e-4)
    for t in range(3):
        for p, g in zip(ps_t, grads[t]):
            p.grad = g.cuda()
        topt.step()
    ps = [torch.nn.Parameter(p.detach().clone()) for p in ps_t]
    opt = FusedAdamW(ps, lr=2e-3, weight_decay=1e-4)
    opt.load_state_dict(topt.state_dict())
    for t in range(3, 6):
        for p, g in zip(ps, grads[t]):
            p.grad = g.cuda()
        opt.step()
    assert opt.steps_taken() == 6
    for p, w in zip(ps, want):
        np.testing.assert_allclose(p.detach().cpu().numpy(), w.numpy(), rtol=ADAMW_RTOL, atol=ADAMW_ATOL)
    sd = opt.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    assert "_mmg_state" not in sd["param_groups"][0]


def test_fused_adamw_in_cuda_graph_with_scheduler():
    """capturable=True: the step counter and the learning rate live on the device, so one recorded launch replays as
    steps 1, 2, 3, ... with whatever lr the scheduler wrote before the replay."""
    from mmgclip_b200.optim import FusedAdamW
    shapes = [(96, 64), (64,)]
    steps = 6
    params, grads = _problem(shapes, steps, seed=21)
    lrs = [1e-3 * (t + 1) / 3 if t < 3 else 1e-3 * 0.5 ** (t - 2) for t in range(steps)]
    want = oc.torch_adamw_steps(params, grads, lrs, weight_decay=1e-2)
    ps = [torch.nn.Parameter(p.clone().cuda()) for p in params]
    static_g = [torch.zeros_like(p) for p in ps]
    for p, g in zip(ps, static_g):
        p.grad = g
    opt = FusedAdamW(ps, lr=lrs[0], weight_decay=1e-2, capturable=True)
    # one eager step allocates the state; undo it so that the graph starts from step 0
    snap = [p.detach().clone() for p in ps]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        opt.step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    with torch.no_grad():
        for p, s in zip(ps, snap):
            p.copy_(s)
            opt.state[p]["exp_avg"].zero_(); opt.state[p]["exp_avg_sq"].zero_()
        opt.state[ps[0]]["step"].zero_()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        opt.step()
    for t in range(steps):
        for sg, g in zip(static_g, grads[t]):
            sg.copy_(g)
        for grp in opt.param_groups:
            grp["lr"] = lrs[t]
        opt.sync_lr()
        graph.replay()
    torch.cuda.synchronize()
    assert opt.steps_taken() == steps
    for p, w in zip(ps, want):
        np.testing.assert_allclose(p.detach().cpu().numpy(), w.numpy(), rtol=ADAMW_RTOL, atol=ADAMW_ATOL)


def test_graphed_train_step_with_optimizer_matches_eager():
    """forward + loss + backward + FusedAdamW recorded by GraphedStep == the same loop run eagerly."""
    from mmgclip_b200.graph import GraphedStep
    from mmgclip_b200.losses import CLIPLoss
    from mmgclip_b200.optim import FusedAdamW
    from mmgclip_b200.projection import LinearProjectionLayer

    def build():
        wi, wt = oc.synthetic_head_weights(64, 96, 80, seed=5)
        hi = LinearProjectionLayer(96, 64, precision="fp32").cuda()
        ht = LinearProjectionLayer(80, 64, precision="fp32").cuda()
        with torch.no_grad():
            hi.layer.weight.copy_(torch.from_numpy(wi)); ht.layer.weight.copy_(torch.from_numpy(wt))
        params = list(hi.parameters()) + list(ht.parameters())
        opt = FusedAdamW(params, lr=1e-3, weight_decay=1e-2, capturable=True)
        crit = CLIPLoss(precision="fp32")
        scale = torch.tensor(math.log(1 / 0.07), device="cuda").exp()

        def step(xi, xt):
            opt.zero_grad(set_to_none=True)
            loss, _ = crit(image_embeddings=hi.forward_normalized(xi), text_embeddings=ht.forward_normalized(xt),
                           logit_scale=scale)
            loss.backward()
            opt.step()
            return loss
        return params, opt, step

    xi, xt = oc.synthetic_features(256, 96, 80, seed=31)
    xi, xt = torch.from_numpy(xi).cuda(), torch.from_numpy(xt).cuda()
    warm, steps = 2, 5
    params_e, opt_e, step_e = build()
    losses_e = [step_e(xi, xt).item() for _ in range(warm + steps)]
    params_g, opt_g, step_g = build()
    g = GraphedStep(step_g, [(xi, xt)], warmup=warm, params=params_g)   # the warm-up steps are real steps
    losses_g = []
    for _ in range(steps):
        losses_g.append(g(0).item())
    assert opt_g.steps_taken() == warm + steps
    assert losses_g[-1] < losses_g[0]  # it trains
    np.testing.assert_allclose(losses_g, losses_e[warm:], rtol=1e-5)
    for a, b in zip(params_g, params_e):
        assert rel_err(a.detach().cpu(), b.detach().cpu()) < 1e-5
