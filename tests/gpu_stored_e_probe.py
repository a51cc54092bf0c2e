"""Probe (not a pytest): stored-E forward/backward against the recompute path.
    python tests/gpu_stored_e_probe.py check          # correctness at small / ragged-free shapes
    python tests/gpu_stored_e_probe.py time [B] [D]   # timings at the benchmark shape
    python tests/gpu_stored_e_probe.py variants       # opt-in kernel variants (panel counters, deferred publish, 16
                                                      # transform warps): correctness vs the default, then launch times
    python tests/gpu_stored_e_probe.py trace [B] [prefix]   # MMG_FUSED_TRACE timelines -> <prefix>_{stored,recompute}.npy
"""
import math
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from mmgclip_b200 import ops  # noqa: E402


def embeddings(n, D, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.nn.functional.normalize(torch.randn(n, D, generator=g) + 0.3, dim=1).cuda()
    b = torch.nn.functional.normalize(torch.randn(n, D, generator=g) + 0.3 * a.cpu(), dim=1).cuda()
    return a, b


def rel(x, y):
    return float((x.double() - y.double()).abs().max() / y.double().abs().max())


def run(a, b, s, stored, reps=0):
    n, D = a.shape
    ab, bb = ops.cast_bf16(a), ops.cast_bf16(b)
    E = torch.empty((n, n), dtype=torch.bfloat16, device="cuda") if stored else None
    gl = torch.ones((), device="cuda")

    def fwd():
        return ops.infonce_forward_raw(ab, bb, s, 0, "bf16", e_out=E)

    def bwd(rowsum, colsum, diag):
        return ops.infonce_backward_raw(ab, bb, s, rowsum, colsum, gl, 0.5 / n, 0, "bf16", a32=a, b32=b, diag=diag,
                                        need_dscale=False, e_stored=E)
    rowsum, colsum, diag = fwd()
    dA, dB, _ = bwd(rowsum, colsum, diag)
    torch.cuda.synchronize()
    t = None
    if reps:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = tb = 0.0
        for _ in range(reps):
            ev[0].record(); r = fwd(); ev[1].record(); bwd(*r); ev[2].record()
            torch.cuda.synchronize()
            tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
        t = (tf / reps, tb / reps)
    return rowsum, colsum, diag, dA, dB, E, t


def check():
    s = torch.tensor(math.log(1 / 0.07), device="cuda").exp()
    for n, D in [(256, 256), (512, 256), (1024, 512), (4096, 512), (8192, 256), (4096, 1024)]:
        a, b = embeddings(n, D, seed=n + D)
        r0 = run(a, b, s, False)
        r1 = run(a, b, s, True)
        ab, bb = ops.cast_bf16(a).float(), ops.cast_bf16(b).float()
        e_ref = torch.exp(s * (ab @ bb.t()) - s)
        e_err = float(((r1[5].float() - e_ref).abs() / e_ref).max())
        print(f"n={n} D={D}: rowsum {rel(r1[0], r0[0]):.1e} colsum {rel(r1[1], r0[1]):.1e} diag {rel(r1[2], r0[2]):.1e} "
              f"| E rel err {e_err:.2e} | dA {rel(r1[3], r0[3]):.2e} dB {rel(r1[4], r0[4]):.2e}", flush=True)


def time_(B, D):
    s = torch.tensor(math.log(1 / 0.07), device="cuda").exp()
    a, b = embeddings(B, D, seed=1)
    probe = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    ops.set_backward_probe(probe)
    ab, bb = ops.cast_bf16(a), ops.cast_bf16(b)
    gl = torch.ones((), device="cuda")
    E = torch.empty((B, B), dtype=torch.bfloat16, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    import os
    orders = ((False, True), (True, False))
    if os.environ.get("PROBE_ONLY") == "stored":
        orders = ((True,),)
    elif os.environ.get("PROBE_ONLY") == "recompute":
        orders = ((False,),)
    elif os.environ.get("PROBE_ONLY") == "tw":  # stored-E with 8, then 16, then 8 transform warps (env is read per launch)
        orders = ((True,), (True,), (True,))
    tw_cycle = ["8", "16", "8"]
    for oi, order in enumerate(orders):
        if os.environ.get("PROBE_ONLY") == "tw":
            os.environ["MMG_STORED_TW"] = tw_cycle[oi]
            print("MMG_STORED_TW=" + tw_cycle[oi], flush=True)
        for stored in order:
            e = E if stored else None
            tf, tb = [], []
            for _ in range(int(os.environ.get("PROBE_REPS", "8"))):
                ev[0].record()
                r = ops.infonce_forward_raw(ab, bb, s, 0, "bf16", e_out=e)
                ev[1].record()
                ops.infonce_backward_raw(ab, bb, s, r[0], r[1], gl, 0.5 / B, 0, "bf16", a32=a, b32=b, diag=r[2],
                                         need_dscale=False, e_stored=e)
                torch.cuda.synchronize()
                tf.append(ev[0].elapsed_time(ev[1])); tb.append(probe[0].elapsed_time(probe[1]))
            print(f"B={B} D={D} stored={int(stored)}: fwd " + " ".join(f"{x:.3f}" for x in tf) + " | fused bwd launch "
                  + " ".join(f"{x:.3f}" for x in tb), flush=True)


def trace(B, D, out_prefix):
    """MMG_FUSED_TRACE=1 timelines of one fused backward launch per mode -> <out_prefix>_{stored,recompute}.npy
    (int64 [CTAs, roles, records, 2] = {globaltimer ns, tag}; see mmg_debug_fused_trace_region in the header)."""
    import ctypes
    import os
    os.environ["MMG_FUSED_TRACE"] = "1"
    lib = ops._lib.load()
    s = torch.tensor(math.log(1 / 0.07), device="cuda").exp()
    a, b = embeddings(B, D, seed=1)
    ab, bb = ops.cast_bf16(a), ops.cast_bf16(b)
    gl = torch.ones((), device="cuda")
    E = torch.empty((B, B), dtype=torch.bfloat16, device="cuda")
    off, nb = ctypes.c_size_t(), ctypes.c_size_t()
    per, roles = ctypes.c_int(), ctypes.c_int()
    assert lib.mmg_debug_fused_trace_region(B, B, D, ctypes.byref(off), ctypes.byref(nb), ctypes.byref(per),
                                            ctypes.byref(roles)) == 1
    nbytes = lib.mmg_infonce_workspace_bytes(ops._PREC["bf16"], B, B, D)
    for mode, e in (("stored", E), ("recompute", None)):
        for _ in range(3):
            r = ops.infonce_forward_raw(ab, bb, s, 0, "bf16", e_out=e)
            ops.infonce_backward_raw(ab, bb, s, r[0], r[1], gl, 0.5 / B, 0, "bf16", a32=a, b32=b, diag=r[2],
                                     need_dscale=False, e_stored=e)
        torch.cuda.synchronize()
        ws = ops._workspace(a.device, nbytes)
        raw = ws[off.value:off.value + nb.value].clone().view(torch.int64).cpu().numpy()
        raw = raw.reshape(-1, roles.value, per.value, 2)
        np.save(f"{out_prefix}_{mode}.npy", raw)
        used = (raw[..., 0] != 0).sum(axis=2)
        t = raw[..., 0][raw[..., 0] != 0]
        print(f"{mode}: CTAs {raw.shape[0]}, records per role (max) {used.max(axis=0).tolist()}, span "
              f"{(t.max() - t.min()) / 1e6:.3f} ms", flush=True)


VARIANTS = [
    ("recompute", False, {}),
    ("recompute+panel", False, {"MMG_FUSED_PANEL": "1"}),
    ("stored", True, {}),
    ("stored+panel", True, {"MMG_FUSED_PANEL": "1"}),
    ("stored+defer", True, {"MMG_STORED_DEFER": "1"}),
    ("stored+defer+panel", True, {"MMG_STORED_DEFER": "1", "MMG_FUSED_PANEL": "1"}),
    ("stored+tw16", True, {"MMG_STORED_TW": "16"}),
]


def variants(B=32768, D=512):
    """Every opt-in kernel variant of the fused backward against the default recompute kernel: correctness at a few
    shapes (the library reads the switches at launch time, so one process can toggle them), then launch times at B."""
    import os
    keys = ("MMG_FUSED_PANEL", "MMG_STORED_DEFER", "MMG_STORED_TW")
    s = torch.tensor(math.log(1 / 0.07), device="cuda").exp()

    def setenv(env):
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update(env)

    for n, d in [(256, 256), (1024, 512), (4096, 512), (8192, 256)]:
        a, b = embeddings(n, d, seed=n + d)
        setenv({})
        ref = run(a, b, s, False)
        for name, stored, env in VARIANTS[1:]:
            setenv(env)
            r = run(a, b, s, stored)
            tol = 2e-3 if stored else 1e-5
            da, db = rel(r[3], ref[3]), rel(r[4], ref[4])
            print(f"n={n} D={d} {name:20s} dA {da:.2e} dB {db:.2e} {'ok' if max(da, db) < tol else 'MISMATCH'}", flush=True)
    a, b = embeddings(B, D, seed=1)
    ab, bb = ops.cast_bf16(a), ops.cast_bf16(b)
    gl = torch.ones((), device="cuda")
    E = torch.empty((B, B), dtype=torch.bfloat16, device="cuda")
    probe = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    ops.set_backward_probe(probe)
    for rnd in range(2):
        for name, stored, env in VARIANTS:
            setenv(env)
            e = E if stored else None
            ts = []
            for _ in range(5):
                r = ops.infonce_forward_raw(ab, bb, s, 0, "bf16", e_out=e)
                ops.infonce_backward_raw(ab, bb, s, r[0], r[1], gl, 0.5 / B, 0, "bf16", a32=a, b32=b, diag=r[2],
                                         need_dscale=False, e_stored=e)
                torch.cuda.synchronize()
                ts.append(probe[0].elapsed_time(probe[1]))
            print(f"B={B} round {rnd} {name:20s} fused bwd launch " + " ".join(f"{x:.3f}" for x in ts[1:]), flush=True)
    setenv({})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "variants":
        variants()
    elif len(sys.argv) > 1 and sys.argv[1] == "trace":
        trace(int(sys.argv[2]) if len(sys.argv) > 2 else 32768, 512, sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/trace")
    elif len(sys.argv) > 1 and sys.argv[1] == "time":
        time_(int(sys.argv[2]) if len(sys.argv) > 2 else 32768, int(sys.argv[3]) if len(sys.argv) > 3 else 512)
    else:
        check()
