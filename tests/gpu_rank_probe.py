"""Compute-only time of ONE rank's step in an R-way row-sharded run (no collectives), replayed as a CUDA graph on one GPU
(not a pytest file): what the per-rank kernels cost at the 2/4/8-GPU shapes of the bench, so the rest of a measured
multi-GPU step can be attributed to the exchange steps."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmgclip_b200 import ops  # noqa: E402
from mmgclip_b200.graph import GraphedStep  # noqa: E402
from mmgclip_b200.projection import LinearProjectionLayer  # noqa: E402

dev = torch.device("cuda:0")
B, E, D = 32768, 768, 512
hi, ht = LinearProjectionLayer(E, D).to(dev), LinearProjectionLayer(E, D).to(dev)
scale = torch.tensor(math.log(1 / 0.07), device=dev).exp()
for world in (1, 2, 4, 8):
    bl = B // world
    off = (world // 2) * bl if world > 1 else 0
    gen = torch.Generator(device=dev).manual_seed(world)
    xi = torch.randn(bl, E, device=dev, generator=gen).abs() + 0.5
    xt = torch.randn(bl, E, device=dev, generator=gen) * 0.5
    b_other = torch.nn.functional.normalize(torch.randn(B, D, device=dev, generator=gen), dim=1).bfloat16()

    def step(xi, xt):
        hi.layer.weight.grad = None
        ht.layer.weight.grad = None
        te = ht.forward_normalized(xt)
        ie = hi.forward_normalized(xi)
        b_all = b_other.clone()                       # stands in for the all-gather output (same bytes written)
        b_all[off:off + bl] = ops._operand(te, "bf16")
        a_op = ops._operand(ie, "bf16")
        rs, cs, dg = ops.infonce_forward_raw(a_op, b_all, scale, off, "bf16")
        loss = ops.infonce_loss_raw(rs, cs[off:off + bl], dg, scale, 0.5 / B)
        one = torch.ones((), device=dev)
        dA, dB, _ = ops.infonce_backward_raw(a_op, b_all, scale, rs, cs, one, 0.5 / B, off, "bf16", a32=ie.detach(),
                                             b32=te.detach(), diag=dg, need_dscale=False)
        torch.autograd.backward([ie, te], [dA, dB[off:off + bl].contiguous()])
        return loss

    g = GraphedStep(step, [(xi, xt)], params=list(hi.parameters()) + list(ht.parameters()))
    for _ in range(3):
        g(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g(0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"world {world}: {bl} rows x {B} cols per rank: compute-only step {ms:.3f} ms ({g.kernel_launches} launches) -> "
          f"{B / ms / 1e3:.1f} M pairs/s if the exchange were free", flush=True)
    del g
