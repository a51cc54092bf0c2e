"""One fused-backward call at the bench shape (for ncu captures; not a pytest file)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmgclip_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
cols = int(sys.argv[2]) if len(sys.argv) > 2 else rows
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
d = 512
gen = torch.Generator(device=dev).manual_seed(7)
a = torch.nn.functional.normalize(torch.randn(rows, d, device=dev, generator=gen), dim=1)
b = torch.nn.functional.normalize(torch.randn(cols, d, device=dev, generator=gen), dim=1)
ab, bb = ops.cast_bf16(a), ops.cast_bf16(b)
s = torch.tensor(1 / 0.07, device=dev)
one = torch.ones((), device=dev)
rs, cs, _ = ops.infonce_forward_raw(ab, bb, s, 0, "bf16")
for _ in range(reps):
    ops.infonce_backward_raw(ab, bb, s, rs, cs, one, 0.5 / cols, 0, "bf16", need_dscale=False)
torch.cuda.synchronize()
print("ok")
