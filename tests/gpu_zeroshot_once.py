"""One tensor-core zero-shot scoring call at 1M x 64 x 512 (for ncu captures; not a pytest file)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmgclip_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
n, c, d, k = 1 << 20, 64, 512, 5
gen = torch.Generator(device=dev).manual_seed(4)
img = torch.nn.functional.normalize(torch.randn(n, d, device=dev, generator=gen), dim=1)
txt = torch.nn.functional.normalize(torch.randn(c, d, device=dev, generator=gen), dim=1)
s = torch.tensor(1 / 0.07, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    ops.zeroshot_score(img, txt, s, k=k, want_logits=False, want_probs=False, impl="tc")
torch.cuda.synchronize()
print("ok")
