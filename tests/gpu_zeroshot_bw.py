"""Achieved HBM bandwidth of the tensor-core zero-shot kernel for different row lengths at equal bytes (not a pytest file)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmgclip_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
c, k = 64, 5
s = torch.tensor(1 / 0.07, device=dev)
for (n, d) in [(1 << 22, 128), (1 << 21, 256), (1 << 20, 512), (1 << 19, 1024), (1 << 18, 2048)]:
    gen = torch.Generator(device=dev).manual_seed(4)
    sets = [torch.randn(n, d, device=dev, generator=gen) for _ in range(2)]
    txt = torch.randn(c, d, device=dev, generator=gen)
    for impl in ("tc",):
        fn = lambda i: ops.zeroshot_score(sets[i % 2], txt, s, k=k, want_logits=False, want_probs=False, impl=impl)  # noqa: E731
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"  N={n} D={d} {impl}: {ms:.3f} ms, {n * d * 4 / ms / 1e6:.0f} GB/s of embeddings", flush=True)
    del sets
