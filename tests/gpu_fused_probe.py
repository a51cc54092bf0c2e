"""Fused persistent backward vs the block loop: agreement and timing (not a pytest file).

    python tests/gpu_fused_probe.py [rows] [cols] [quick]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmgclip_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, iters=5):
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


def setenv(**kw):
    for k in ("MMG_BWD_FUSED", "MMG_FUSED_RB", "MMG_FUSED_CB", "MMG_FUSED_NBUF", "MMG_FUSED_KSL", "MMG_FUSED_KSL_T",
              "MMG_FUSED_BN", "MMG_FUSED_EPI_WARPS"):
        os.environ.pop(k, None)
    for k, v in kw.items():
        os.environ[k] = str(v)


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    cols = int(sys.argv[2]) if len(sys.argv) > 2 else rows
    quick = len(sys.argv) > 3
    d = 512
    off = 0 if rows == cols else cols // 2 // 256 * 256  # sharded case: local rows pair with a column range
    off = min(off, cols - rows)
    gen = torch.Generator(device=dev).manual_seed(7)
    a = torch.nn.functional.normalize(torch.randn(rows, d, device=dev, generator=gen), dim=1)
    b = torch.nn.functional.normalize(torch.randn(cols, d, device=dev, generator=gen), dim=1)
    b[off:off + rows] = torch.nn.functional.normalize(b[off:off + rows] + 0.5 * a, dim=1)  # correlated pairs
    ab, bb = ops.cast_bf16(a), ops.cast_bf16(b)
    s = torch.tensor(1 / 0.07, device=dev)
    one = torch.ones((), device=dev)
    f = 2.0 * rows * cols * d
    rs, cs, _ = ops.infonce_forward_raw(ab, bb, s, off, "bf16")
    # the column sums of a shard are partial; good enough as positive normalisers for a backward comparison
    bwd = lambda dls: ops.infonce_backward_raw(ab, bb, s, rs, cs, one, 0.5 / cols, off, "bf16", need_dscale=dls)  # noqa: E731

    setenv(MMG_BWD_FUSED=0)
    dA0, dB0, dl0 = bwd(True)
    torch.cuda.synchronize()
    t_loop = timeit(lambda: bwd(False))
    print(f"{rows}x{cols}: block loop {t_loop:.3f} ms ({3 * f / t_loop / 1e9:.0f} TF exec)", flush=True)
    cfgs = [dict()]
    if not quick:
        cfgs += [dict(MMG_FUSED_EPI_WARPS=16), dict(), dict(MMG_FUSED_EPI_WARPS=16), dict(MMG_FUSED_EPI_WARPS=16, MMG_FUSED_KSL=16),
                 dict(MMG_FUSED_EPI_WARPS=16, MMG_FUSED_NBUF=3)]
    for c in cfgs:
        c = {k: v for k, v in c.items() if not (k == "MMG_FUSED_RB" and rows % v) and not (k == "MMG_FUSED_CB" and cols % v)}
        setenv(**c)
        dA, dB, dl = bwd(True)
        torch.cuda.synchronize()
        ea = ((dA - dA0).abs().max() / dA0.abs().max()).item()
        eb = ((dB - dB0).abs().max() / dB0.abs().max()).item()
        el = abs(dl.item() - dl0.item()) / max(abs(dl0.item()), 1e-30)
        t = timeit(lambda: bwd(False))
        t1 = timeit(lambda: bwd(True))
        print(f"  fused {c or 'default'}: {t:.3f} ms ({3 * f / t / 1e9:.0f} TF exec; with dscale {t1:.3f} ms) | vs loop: dA {ea:.2e} dB {eb:.2e} "
              f"dls {el:.2e}", flush=True)
    setenv()


if __name__ == "__main__":
    main()
