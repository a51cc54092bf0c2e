"""Feasibility probe (torchrun, >= 2 GPUs; not a pytest file): TMA reduce-add (the EpiStoreF32 atomic mode of the tcgen05
GEMM) straight into ANOTHER rank's buffer through torch symmetric memory (NVLink peer mapping)."""
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmgclip_b200 import _lib, ops  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    M, N, K = 2048, 512, 4096
    c = symm_mem.empty((M, N), dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(c, group=dist.group.WORLD)
    print(f"[{rank}] rendezvous ok: world {hdl.world_size} ptrs {[hex(p) for p in hdl.buffer_ptrs]}", flush=True)
    c.zero_()
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    A = torch.randn(M, K, device=dev, generator=g).bfloat16()
    B = torch.randn(N, K, device=dev, generator=g).bfloat16()
    torch.cuda.synchronize()
    hdl.barrier(channel=0)
    peer = (rank + 1) % world
    cp = hdl.get_buffer(peer, (M, N), torch.float32)
    t0 = time.time()
    for _ in range(3):
        ops.gemm(A, B, M, N, K, out=cp, mode=_lib.MMG_ATOMIC_ADD, k_splits=4, prec="bf16")   # TMA reduce-add into the peer
    torch.cuda.synchronize()
    print(f"[{rank}] 3 remote reduce-add GEMMs done in {time.time() - t0:.3f}s", flush=True)
    hdl.barrier(channel=0)
    # my buffer now holds 3 x (A.B^T of rank-1)
    src = (rank - 1) % world
    g2 = torch.Generator(device=dev).manual_seed(100 + src)
    A2 = torch.randn(M, K, device=dev, generator=g2).bfloat16()
    B2 = torch.randn(N, K, device=dev, generator=g2).bfloat16()
    ref = 3 * (A2.float() @ B2.float().t())
    err = ((c - ref).abs().max() / ref.abs().max()).item()
    print(f"[{rank}] remote reduce-add result rel err {err:.2e} -> {'OK' if err < 1e-4 else 'FAILED'}", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
