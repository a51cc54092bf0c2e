"""Multi-GPU equivalence check (not a pytest file; run under torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/gpu_dist_check.py

R-rank sharded loss / gradients (NCCL: all-gather, all-reduce, asynchronous reduce-scatter, flat head-grad all-reduce)
== the 1-GPU result on the same global batch, eager and as a CUDA graph."""
import math
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmgclip_b200 import ops  # noqa: E402
from mmgclip_b200.distributed import allreduce_gradients, gather_columns_async, sharded_info_nce  # noqa: E402
from mmgclip_b200.graph import GraphedStep  # noqa: E402
from mmgclip_b200.projection import LinearProjectionLayer  # noqa: E402
from oracle import clip_oracle as oc  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, E, D = 2048 * world, 768, 512
    bl = B // world
    xi_g, xt_g = oc.synthetic_features(B, E, E, seed=42)
    wi, wt = oc.synthetic_head_weights(D, E, E, seed=43)
    hi, ht = LinearProjectionLayer(E, D).to(dev), LinearProjectionLayer(E, D).to(dev)
    with torch.no_grad():
        hi.layer.weight.copy_(torch.from_numpy(wi)); ht.layer.weight.copy_(torch.from_numpy(wt))
    scale = torch.tensor(math.log(1 / 0.07), device=dev).exp()
    xi = torch.from_numpy(xi_g[rank * bl:(rank + 1) * bl].copy()).to(dev)
    xt = torch.from_numpy(xt_g[rank * bl:(rank + 1) * bl].copy()).to(dev)

    def step(xi, xt):
        hi.layer.weight.grad = None
        ht.layer.weight.grad = None
        te = ht.forward_normalized(xt)
        gathered = gather_columns_async(te)
        ie = hi.forward_normalized(xi)
        loss = sharded_info_nce(ie, te, scale, gathered=gathered)
        loss.backward()
        allreduce_gradients(hi, ht)
        return loss

    loss = step(xi, xt)
    torch.cuda.synchronize()
    got = (loss.item(), hi.layer.weight.grad.clone(), ht.layer.weight.grad.clone())
    del loss
    g = GraphedStep(step, [(xi, xt)], params=list(hi.parameters()) + list(ht.parameters()))
    lg = g(0)
    torch.cuda.synchronize()
    got_graph = (lg.item(), hi.layer.weight.grad.clone(), ht.layer.weight.grad.clone())
    ok = True
    if rank == 0:
        # single-GPU reference on the whole batch (no process group involved: plain info_nce)
        hi.layer.weight.grad = None
        ht.layer.weight.grad = None
        ie = hi.forward_normalized(torch.from_numpy(xi_g).to(dev))
        te = ht.forward_normalized(torch.from_numpy(xt_g).to(dev))
        ref = ops.info_nce(ie, te, scale)
        ref.backward()
        torch.cuda.synchronize()
        for name, (l, gi, gt) in (("eager", got), ("graph", got_graph)):
            el = abs(l - ref.item()) / ref.item()
            ei = ((gi - hi.layer.weight.grad).abs().max() / hi.layer.weight.grad.abs().max()).item()
            et = ((gt - ht.layer.weight.grad).abs().max() / ht.layer.weight.grad.abs().max()).item()
            print(f"{name}: world {world} loss {l:.6f} vs {ref.item():.6f} rel {el:.2e}  dW_image {ei:.2e}  dW_text {et:.2e}", flush=True)
            ok = ok and el < 1e-5 and ei < 2e-4 and et < 2e-4
        print("DIST CHECK", "OK" if ok else "FAILED", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
