"""Probe (not a pytest): accuracy and cost of the head-weight gradient with three K segments (hi.hi + hi.lo + lo.hi)
versus one (hi.hi) at the benchmark shape, against the fp32 (FFMA) mode of the same library on the same inputs.
    python tests/gpu_dw_probe.py [batch]
"""
import math
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402  (synthetic recipe)
from mmgclip_b200 import ops  # noqa: E402
from mmgclip_b200.losses import CLIPLoss  # noqa: E402
from mmgclip_b200.projection import LinearProjectionLayer  # noqa: E402


def run(prec, xi, xt, wi, wt, reps=0):
    hi = LinearProjectionLayer(768, wi.shape[0], precision=prec).cuda()
    ht = LinearProjectionLayer(768, wt.shape[0], precision=prec).cuda()
    with torch.no_grad():
        hi.layer.weight.copy_(wi); ht.layer.weight.copy_(wt)
    crit = CLIPLoss(precision=prec)
    scale = torch.tensor(math.log(1 / 0.07), device="cuda").exp()

    def step():
        hi.layer.weight.grad = None; ht.layer.weight.grad = None
        loss, _ = crit(image_embeddings=hi.forward_normalized(xi), text_embeddings=ht.forward_normalized(xt),
                       logit_scale=scale)
        loss.backward()
        return loss
    loss = step()
    ms = None
    if reps:
        for _ in range(3):
            step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps):
            step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    return loss.item(), hi.layer.weight.grad.double().cpu().numpy(), ht.layer.weight.grad.double().cpu().numpy(), ms


def err(a, b):
    return np.abs(a - b).max() / np.abs(b).max(), np.linalg.norm(a - b) / np.linalg.norm(b)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    xi, xt = bench.synthetic_features(B)
    wi, wt = bench.synthetic_head_weights(512)
    xi, xt, wi, wt = (torch.from_numpy(t).cuda() for t in (xi, xt, wi, wt))
    l32, gi32, gt32, _ = run("fp32", xi, xt, wi, wt)
    print(f"batch {B}: fp32 loss {l32:.7f}")
    for split in (True, False):
        ops.set_split_dw(split)
        l, gi, gt, ms = run("bf16", xi, xt, wi, wt, reps=20)
        (mi, fi), (mt, ft) = err(gi, gi32), err(gt, gt32)
        print(f"split_dw={int(split)}: loss rel {abs(l - l32) / l32:.2e} | dW_image max {mi:.2e} fro {fi:.2e} | "
              f"dW_text max {mt:.2e} fro {ft:.2e} | eager step {ms:.3f} ms")


if __name__ == "__main__":
    main()
