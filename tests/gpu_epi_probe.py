"""Per-phase timing of the fused InfoNCE at the bench shape (not a pytest file).

    MMG_EPI_WARPS=8|16  MMG_EPI_DBG=0..3  MMG_TC_DUAL_SPLIT=0|1  python tests/gpu_epi_probe.py [rows] [cols]

Prints forward, coefficient-launch (EpiGrad) and gradient-GEMM launch times measured with CUDA events."""
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


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    cols = int(sys.argv[2]) if len(sys.argv) > 2 else rows
    d = 512
    gen = torch.Generator(device=dev).manual_seed(7)
    a = torch.nn.functional.normalize(torch.randn(rows, d, device=dev, generator=gen), dim=1)
    b = torch.nn.functional.normalize(torch.randn(cols, d, device=dev, generator=gen), dim=1)
    ab, bb = ops.cast_bf16(a), ops.cast_bf16(b)
    s = torch.tensor(1 / 0.07, device=dev)
    one = torch.ones((), device=dev)
    f = 2.0 * rows * cols * d
    t_f = timeit(lambda: ops.infonce_forward_raw(ab, bb, s, 0, "bf16"))
    rs, cs, _ = ops.infonce_forward_raw(ab, bb, s, 0, "bf16")
    M = 8192
    A = torch.randn(M, M, device=dev).bfloat16()
    Bm = torch.randn(M, M, device=dev).bfloat16()
    C = torch.empty(M, M, device=dev)
    t_p = timeit(lambda: ops.gemm(A, Bm, M, M, M, prec="bf16", out=C))
    print(f"plain tc_gemm 8192^3: {t_p:.3f} ms ({2.0 * M ** 3 / t_p / 1e9:.0f} TF)")
    del A, Bm, C
    for dls in (False,) if os.environ.get("MMG_PROBE_QUICK") else (False, True):
        bwd = lambda: ops.infonce_backward_raw(ab, bb, s, rs, cs, one, 0.5 / cols, 0, "bf16", need_dscale=dls)  # noqa: E731
        t_b = timeit(bwd)
        os.environ["MMG_BWD_PHASES"] = "1"
        t_c = timeit(bwd)
        os.environ["MMG_BWD_PHASES"] = "2"
        t_g = timeit(bwd)
        os.environ.pop("MMG_BWD_PHASES", None)
        tag = " ".join(f"{k}={os.environ[k]}" for k in ("MMG_EPI_WARPS", "MMG_EPI_DBG", "MMG_TC_DUAL_SPLIT", "MMGCLIP_B200_LIB", "MMGCLIP_B200_BLOCK_ROWS",
                                                             "MMGCLIP_B200_BLOCK_COLS") if k in os.environ)
        print(f"[{tag or 'default'}] dls={int(dls)} {rows}x{cols}: fwd {t_f:.3f} ms ({f / t_f / 1e9:.0f} TF) | coef {t_c:.3f} ms "
              f"({f / t_c / 1e9:.0f} TF) | grad {t_g:.3f} ms ({2 * f / t_g / 1e9:.0f} TF) | bwd {t_b:.3f} ms "
              f"({3 * f / t_b / 1e9:.0f} TF exec)", flush=True)


if __name__ == "__main__":
    main()
