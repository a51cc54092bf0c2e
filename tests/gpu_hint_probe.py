"""Plan / L2-hint sweep of the fused InfoNCE backward through mmg_tune (product library; not a pytest file).

    python tests/gpu_hint_probe.py [rows] [cols] [once]

`once`: one launch per configuration (for `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:infonce_bwd_fused`:
launch i of the log is configuration i of the list printed here)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmgclip_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")

CONFIGS = [
    dict(),
    dict(fused_hints=1),
    dict(fused_hints=3),
    dict(fused_hints=3 + 64),
    dict(fused_hints=3 + 4),
    dict(fused_hints=3 + 64 + 8),
    dict(fused_hints=3 + 64 + 16),
    dict(fused_hints=3 + 64 + 32),
    dict(fused_hints=3 + 64 + 16 + 32),
    dict(fused_ksl_t=64),
    dict(fused_ksl_t=64, fused_hints=3 + 64),
    dict(fused_cb=1024),
    dict(fused_cb=1024, fused_hints=3 + 64),
    dict(fused_cb=1024, fused_ksl_t=64, fused_hints=3 + 64),
    dict(fused_rb=2048, fused_hints=3 + 64),
    dict(fused_nbuf=3, fused_hints=3 + 64),
    dict(fused_rb=2048, fused_cb=1024, fused_nbuf=6, fused_hints=3 + 64),
]


def timeit(fn, iters=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    cols = int(sys.argv[2]) if len(sys.argv) > 2 else rows
    once = len(sys.argv) > 3
    d = 512
    off = 0 if rows == cols else min(cols // 2 // 256 * 256, cols - rows)
    gen = torch.Generator(device=dev).manual_seed(7)
    a = torch.nn.functional.normalize(torch.randn(rows, d, device=dev, generator=gen), dim=1)
    b = torch.nn.functional.normalize(torch.randn(cols, d, device=dev, generator=gen), dim=1)
    b[off:off + rows] = torch.nn.functional.normalize(b[off:off + rows] + 0.5 * a, dim=1)
    ab, bb = ops.cast_bf16(a), ops.cast_bf16(b)
    s = torch.tensor(1 / 0.07, device=dev)
    one = torch.ones((), device=dev)
    f = 2.0 * rows * cols * d
    rs, cs, _ = ops.infonce_forward_raw(ab, bb, s, off, "bf16")
    bwd = lambda: ops.infonce_backward_raw(ab, bb, s, rs, cs, one, 0.5 / cols, off, "bf16", need_dscale=False)  # noqa: E731
    ops.set_tuning()
    dA0, dB0, _ = bwd()
    torch.cuda.synchronize()
    for i, c in enumerate(CONFIGS):
        c = {k: v for k, v in c.items() if not (k == "fused_rb" and rows % v) and not (k == "fused_cb" and cols % v)}
        ops.set_tuning()
        ops.set_tuning(**c) if c else None
        if once:
            bwd()
            torch.cuda.synchronize()
            print(f"config {i}: {c or 'default'}", flush=True)
            continue
        dA, dB, _ = bwd()
        torch.cuda.synchronize()
        ea = ((dA - dA0).abs().max() / dA0.abs().max()).item()
        eb = ((dB - dB0).abs().max() / dB0.abs().max()).item()
        t = timeit(bwd)
        print(f"{rows}x{cols} {str(c or 'default'):70s} {t:.3f} ms ({3 * f / t / 1e9:.0f} TF exec) | vs default dA {ea:.1e} dB {eb:.1e}",
              flush=True)
    ops.set_tuning()


if __name__ == "__main__":
    main()
