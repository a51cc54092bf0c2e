"""Tensor-core zero-shot scoring vs the FFMA kernel and float64 (not a pytest file): errors, index agreement, timing."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmgclip_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def main():
    g = torch.Generator(device="cpu").manual_seed(5)
    s = torch.tensor(float(np.float32(1 / 0.07)), device=dev)
    for (N, C, D, k) in [(1, 2, 512, 0), (100, 8, 512, 5), (129, 64, 64, 8), (5000, 64, 512, 5), (777, 7, 256, 3), (4097, 33, 200, 5),
                         (65536, 64, 512, 5)]:
        img = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=1).to(dev)
        txt = torch.nn.functional.normalize(torch.randn(C, D, generator=g), dim=1).to(dev)
        if C >= 8:
            txt[5] = txt[2]  # duplicated prompt: exact ties
        a = ops.zeroshot_score(img, txt, s, k=k, impl="ffma")
        b = ops.zeroshot_score(img, txt, s, k=k, impl="tc")
        torch.cuda.synchronize()
        L64 = (s.double() * img.double()) @ txt.double().t()
        msg = (f"  N={N} C={C} D={D}: logits tc-vs-f64 {rel(b['logits'], L64):.2e} ffma-vs-f64 {rel(a['logits'], L64):.2e} "
               f"probs tc-vs-ffma {rel(b['probs'], a['probs']):.2e} argmax mismatch {(a['argmax'] != b['argmax']).sum().item()}")
        if k:
            msg += f" topk rows differing {(a['topk_idx'] != b['topk_idx']).any(1).sum().item()}"
            msg += f" topk_val {rel(b['topk_val'], a['topk_val']):.2e}"
        print(msg, flush=True)
    n, c, d, k = 1 << 20, 64, 512, 5
    gen = torch.Generator(device=dev).manual_seed(4)
    sets = [torch.nn.functional.normalize(torch.randn(n, d, device=dev, generator=gen), dim=1) for _ in range(2)]
    txt = torch.nn.functional.normalize(torch.randn(c, d, device=dev, generator=gen), dim=1)
    for impl in ("ffma", "tc"):
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
        byt = n * d * 4 + c * d * 4 + n * (8 + k * 12)
        print(f"  {impl}: {ms:.3f} ms per 1M x 64 x 512 ({byt / ms / 1e6:.0f} GB/s algorithmic)", flush=True)
    a = ops.zeroshot_score(sets[0], txt, s, k=k, want_logits=False, want_probs=False, impl="ffma")
    b = ops.zeroshot_score(sets[0], txt, s, k=k, want_logits=False, want_probs=False, impl="tc")
    print(f"  1M rows: argmax mismatch {(a['argmax'] != b['argmax']).sum().item()} topk rows differing "
          f"{(a['topk_idx'] != b['topk_idx']).any(1).sum().item()}")


if __name__ == "__main__":
    main()
