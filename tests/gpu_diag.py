"""GPU bring-up diagnostics (not a pytest file): prints max errors of every kernel family against plain torch.

Run on a B200:  python tests/gpu_diag.py [section ...]
Each section is isolated in try/except so one failing kernel does not hide the others; a CUDA fault (sticky error)
aborts the remaining sections, which is reported.
"""
import math
import os
import sys
import time
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmgclip_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def sec_gemm_fp32():
    g = torch.Generator(device="cpu").manual_seed(0)
    for (M, N, K) in [(64, 64, 64), (100, 70, 33), (257, 130, 515)]:
        for a_mn in (False, True):
            for b_mn in (False, True):
                A = torch.randn((K, M) if a_mn else (M, K), generator=g).to(dev)
                B = torch.randn((K, N) if b_mn else (N, K), generator=g).to(dev)
                C = ops.gemm(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, prec="fp32")
                ref = (A.t() if a_mn else A).double() @ (B if b_mn else B.t()).double()
                print(f"  sgemm M{M} N{N} K{K} a_mn={int(a_mn)} b_mn={int(b_mn)} rel_err={rel(C, ref):.3e}")


def sec_gemm_bf16():
    g = torch.Generator(device="cpu").manual_seed(1)
    shapes = [(128, 256, 64), (128, 128, 64), (128, 256, 512), (256, 512, 768), (4096, 512, 768), (200, 264, 136),
              (32, 32, 512), (1000, 520, 72)]
    for (M, N, K) in shapes:
        for a_mn in (False, True):
            for b_mn in (False, True):
                if a_mn and M % 8:
                    continue
                if b_mn and N % 8:
                    continue
                A = torch.randn((K, M) if a_mn else (M, K), generator=g).to(dev).bfloat16()
                B = torch.randn((K, N) if b_mn else (N, K), generator=g).to(dev).bfloat16()
                C = ops.gemm(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, prec="bf16")
                torch.cuda.synchronize()
                ref = (A.t() if a_mn else A).double() @ (B if b_mn else B.t()).double()
                print(f"  tc_gemm M{M} N{N} K{K} a_mn={int(a_mn)} b_mn={int(b_mn)} rel_err={rel(C, ref):.3e}")
    # split-K + atomic, accumulate, bias + relu
    M, N, K = 512, 768, 4096
    A = torch.randn(K, M, generator=g).to(dev).bfloat16()
    B = torch.randn(K, N, generator=g).to(dev).bfloat16()
    ref = A.t().double() @ B.double()
    C = ops.gemm(A, B, M, N, K, a_mn=True, b_mn=True, prec="bf16", k_splits=12, mode=_lib.MMG_ATOMIC_ADD)
    print(f"  tc_gemm splitK=12 atomic rel_err={rel(C, ref):.3e}")
    C2 = torch.ones(M, N, device=dev)
    ops.gemm(A, B, M, N, K, a_mn=True, b_mn=True, prec="bf16", out=C2, mode=_lib.MMG_ACCUMULATE)
    print(f"  tc_gemm accumulate rel_err={rel(C2, ref + 1):.3e}")
    bias = torch.randn(N, generator=g).to(dev)
    A3 = torch.randn(M, 256, generator=g).to(dev).bfloat16()
    B3 = torch.randn(N, 256, generator=g).to(dev).bfloat16()
    C3 = ops.gemm(A3, B3, M, N, 256, prec="bf16", bias=bias, relu=True)
    ref3 = torch.relu(A3.double() @ B3.double().t() + bias.double())
    print(f"  tc_gemm bias+relu rel_err={rel(C3, ref3):.3e}")


def torch_clip(a, b, s):
    L = s * a @ b.t()
    n = a.shape[0]
    lab = torch.arange(n, device=a.device)
    return (torch.nn.functional.cross_entropy(L, lab) + torch.nn.functional.cross_entropy(L.t(), lab)) / 2


def sec_infonce(prec, sizes=((8, 64), (32, 512), (200, 256), (1024, 512), (4096, 512), (5000, 512))):
    g = torch.Generator(device="cpu").manual_seed(2)
    for (n, D) in sizes:
        a = torch.nn.functional.normalize(torch.randn(n, D, generator=g), dim=1).to(dev)
        b = torch.nn.functional.normalize(torch.randn(n, D, generator=g) + 0.5 * a.cpu(), dim=1).to(dev)
        s = torch.tensor(1 / 0.07, device=dev)
        a64, b64, s64 = a.double().requires_grad_(), b.double().requires_grad_(), s.double().requires_grad_()
        ref = torch_clip(a64, b64, s64)
        ref.backward()
        a1, b1, s1 = a.clone().requires_grad_(), b.clone().requires_grad_(), s.clone().requires_grad_()
        loss = ops.info_nce(a1, b1, s1, prec=prec)
        loss.backward()
        torch.cuda.synchronize()
        print(f"  infonce[{prec}] n={n} D={D} loss={loss.item():.6f} ref={ref.item():.6f} "
              f"rel={abs(loss.item() - ref.item()) / abs(ref.item()):.2e} dA={rel(a1.grad, a64.grad):.2e} "
              f"dB={rel(b1.grad, b64.grad):.2e} ds={abs(s1.grad.item() - s64.grad.item()) / (abs(s64.grad.item()) + 1e-12):.2e}")


def sec_small():
    g = torch.Generator(device="cpu").manual_seed(3)
    u = torch.randn(37, 512, generator=g).to(dev).requires_grad_()
    y = ops.l2_normalize(u, prec="bf16")
    ref = u.detach().double().requires_grad_()
    yr = ref / ref.norm(dim=1, keepdim=True)
    w = torch.randn(37, 512, generator=g).to(dev)
    (y * w).sum().backward()
    (yr * w.double()).sum().backward()
    print(f"  l2norm fwd={rel(y, yr):.2e} bwd={rel(u.grad, ref.grad):.2e} bf16copy={rel(y._mmg_bf16.float(), yr):.2e}")
    x = torch.randn(50, 64, generator=g).to(dev).requires_grad_()
    xr = x.detach().double().requires_grad_()
    yg = ops.gelu(x)
    ygr = torch.nn.functional.gelu(xr)
    yg.sum().backward()
    ygr.sum().backward()
    print(f"  gelu fwd={rel(yg, ygr):.2e} bwd={rel(x.grad, xr.grad):.2e}")
    gam = torch.randn(64, generator=g).to(dev).requires_grad_()
    bet = torch.randn(64, generator=g).to(dev).requires_grad_()
    x2 = torch.randn(50, 64, generator=g).to(dev).requires_grad_()
    yl = ops.layer_norm(x2, gam, bet)
    x2r, gr, br = (t.detach().double().requires_grad_() for t in (x2, gam, bet))
    ylr = torch.nn.functional.layer_norm(x2r, (64,), gr, br, 1e-5)
    wt = torch.randn(50, 64, generator=g).to(dev)
    (yl * wt).sum().backward()
    (ylr * wt.double()).sum().backward()
    print(f"  layernorm fwd={rel(yl, ylr):.2e} dx={rel(x2.grad, x2r.grad):.2e} dg={rel(gam.grad, gr.grad):.2e} "
          f"db={rel(bet.grad, br.grad):.2e}")
    L = (torch.randn(8, 8, generator=g) * 5).to(dev).requires_grad_()
    Lr = L.detach().double().requires_grad_()
    lab = torch.arange(8, device=dev)
    out = ops.ce_arange(L, 1.0 / 8)
    outr = torch.nn.functional.cross_entropy(Lr, lab)
    out.backward()
    outr.backward()
    print(f"  ce_arange fwd={abs(out.item() - outr.item()):.2e} bwd={rel(L.grad, Lr.grad):.2e}")


def sec_linear(prec):
    g = torch.Generator(device="cpu").manual_seed(4)
    for (Bn, E, D) in [(32, 768, 512), (300, 768, 512), (4096, 768, 512)]:
        x = torch.randn(Bn, E, generator=g).to(dev).requires_grad_()
        W = (torch.randn(D, E, generator=g) / math.sqrt(E)).to(dev).requires_grad_()
        b = torch.randn(D, generator=g).to(dev).requires_grad_()
        xr, Wr, br = (t.detach().double().requires_grad_() for t in (x, W, b))
        y = ops.linear(x, W, b, relu=True, prec=prec)
        yr = torch.relu(xr @ Wr.t() + br)
        wt = torch.randn(Bn, D, generator=g).to(dev)
        (y * wt).sum().backward()
        (yr * wt.double()).sum().backward()
        print(f"  linear[{prec}] B={Bn} fwd={rel(y, yr):.2e} dx={rel(x.grad, xr.grad):.2e} dW={rel(W.grad, Wr.grad):.2e} "
              f"db={rel(b.grad, br.grad):.2e}")
        x2 = x.detach().clone()
        W2 = W.detach().clone().requires_grad_()
        y2 = ops.project_normalize(x2, W2, prec=prec)
        W2r = W2.detach().double().requires_grad_()
        u = x2.double() @ W2r.t()
        y2r = u / u.norm(dim=1, keepdim=True)
        (y2 * wt).sum().backward()
        (y2r * wt.double()).sum().backward()
        print(f"  projnorm[{prec}] B={Bn} fwd={rel(y2, y2r):.2e} dW={rel(W2.grad, W2r.grad):.2e}")


def sec_zeroshot():
    g = torch.Generator(device="cpu").manual_seed(5)
    for (N, C, D, k) in [(1, 2, 512, 0), (100, 8, 512, 5), (5000, 64, 512, 5), (777, 7, 256, 3)]:
        img = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=1).to(dev)
        txt = torch.nn.functional.normalize(torch.randn(C, D, generator=g), dim=1).to(dev)
        s = torch.tensor(1 / 0.07, device=dev)
        out = ops.zeroshot_score(img, txt, s, k=k)
        L = (s * img) @ txt.t()
        P = L.softmax(-1)
        am = torch.argmax(P, -1)
        msg = f"  zeroshot N={N} C={C} logits={rel(out['logits'], L):.2e} probs={rel(out['probs'], P):.2e} " \
              f"argmax_mismatch={(out['argmax'] != am).sum().item()}"
        if k:
            tv, ti = torch.topk(L, k, dim=-1)
            msg += f" topk_mismatch_rows={(out['topk_idx'] != ti).any(1).sum().item()}"
        print(msg)


def sec_perf():
    # quick timing of the main pieces at the bench shape (device-resident), CUDA events
    def timeit(fn, iters=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    g = torch.Generator(device="cpu").manual_seed(6)
    for n in (4096, 32768):
        D = 512
        a = torch.nn.functional.normalize(torch.randn(n, D, generator=g), dim=1).to(dev)
        b = torch.nn.functional.normalize(torch.randn(n, D, generator=g), dim=1).to(dev)
        ab, bb = a.bfloat16(), b.bfloat16()
        s = torch.tensor(1 / 0.07, device=dev)
        t_f = timeit(lambda: ops.infonce_forward_raw(ab, bb, s, 0, "bf16"))
        rs, cs, dg = ops.infonce_forward_raw(ab, bb, s, 0, "bf16")
        gl = torch.ones((), device=dev)
        t_b = timeit(lambda: ops.infonce_backward_raw(ab, bb, s, rs, cs, gl, 0.5 / n, 0, "bf16"), iters=3)
        fl = 2.0 * n * n * D
        print(f"  perf n={n}: fwd {t_f:.3f} ms ({fl / t_f / 1e9:.1f} TFLOP/s), bwd {t_b:.3f} ms "
              f"({3 * fl / t_b / 1e9:.1f} TFLOP/s executed)")
        for (br, bc) in ((2048, 2048), (4096, 2048), (8192, 4096)):
            if br > n:
                continue
            t_b = timeit(lambda: ops.infonce_backward_raw(ab, bb, s, rs, cs, gl, 0.5 / n, 0, "bf16", br, bc), iters=3)
            print(f"    bwd block {br}x{bc}: {t_b:.3f} ms ({3 * fl / t_b / 1e9:.1f} TFLOP/s executed)")
    M, N, K = 8192, 8192, 8192
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    C = torch.empty(M, N, device=dev)
    t = timeit(lambda: ops.gemm(A, B, M, N, K, prec="bf16", out=C))
    print(f"  perf plain tc_gemm 8192^3: {t:.3f} ms ({2.0 * M * N * K / t / 1e9:.1f} TFLOP/s)")
    t = timeit(lambda: torch.matmul(A, B.t()))
    print(f"  perf torch.matmul bf16 8192^3: {t:.3f} ms ({2.0 * M * N * K / t / 1e9:.1f} TFLOP/s)")


def sec_bf16_err():
    """max-abs and Frobenius relative errors of the bf16 path vs float64 closed form, for tolerance setting."""
    import numpy as np
    from oracle import clip_oracle as oc
    from mmgclip_b200.projection import LinearProjectionLayer

    def fro(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return np.linalg.norm(a - b) / np.linalg.norm(b)

    def mx(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return np.abs(a - b).max() / np.abs(b).max()

    for (n, e, d) in [(32, 96, 64), (32, 768, 512), (256, 768, 512), (2048, 768, 512), (8192, 768, 512)]:
        xi, xt = oc.synthetic_features(n, e, e, seed=n)
        wi, wt = oc.synthetic_head_weights(d, e, e, seed=n + 1)
        ref = oc.closed_form_train_step(xi, xt, wi, wt, math.log(1 / 0.07))
        hi, ht = LinearProjectionLayer(e, d, precision="bf16").cuda(), LinearProjectionLayer(e, d, precision="bf16").cuda()
        with torch.no_grad():
            hi.layer.weight.copy_(torch.from_numpy(wi)); ht.layer.weight.copy_(torch.from_numpy(wt))
        ls = torch.tensor(math.log(1 / 0.07), device="cuda", requires_grad=True)
        ie = hi.forward_normalized(torch.from_numpy(xi).cuda()); te = ht.forward_normalized(torch.from_numpy(xt).cuda())
        ie.retain_grad(); te.retain_grad()
        loss = ops.info_nce(ie, te, ls.exp(), prec="bf16")
        loss.backward()
        print(f"  bf16 n={n} e={e} d={d}: loss_rel={abs(loss.item() - ref['loss']) / ref['loss']:.2e} "
              f"emb max={mx(ie.detach().cpu(), ref['image_embeddings']):.2e} "
              f"dA max={mx(ie.grad.cpu(), ref['da']):.2e} fro={fro(ie.grad.cpu(), ref['da']):.2e} "
              f"dB max={mx(te.grad.cpu(), ref['db']):.2e} fro={fro(te.grad.cpu(), ref['db']):.2e} "
              f"dWi max={mx(hi.layer.weight.grad.cpu(), ref['dw_image']):.2e} fro={fro(hi.layer.weight.grad.cpu(), ref['dw_image']):.2e} "
              f"dWt max={mx(ht.layer.weight.grad.cpu(), ref['dw_text']):.2e} fro={fro(ht.layer.weight.grad.cpu(), ref['dw_text']):.2e} "
              f"dls={abs(ls.grad.item() - ref['dlogit_scale_log']) / max(1, abs(ref['dlogit_scale_log'])):.2e}")


def sec_blocks():
    def timeit(fn, iters=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    n, D = 32768, 512
    g = torch.Generator(device="cpu").manual_seed(6)
    a = torch.nn.functional.normalize(torch.randn(n, D, generator=g), dim=1).to(dev)
    b = torch.nn.functional.normalize(torch.randn(n, D, generator=g), dim=1).to(dev)
    ab, bb = a.bfloat16(), b.bfloat16()
    s = torch.tensor(1 / 0.07, device=dev)
    rs, cs, dg = ops.infonce_forward_raw(ab, bb, s, 0, "bf16")
    gl = torch.ones((), device=dev)
    fl = 2.0 * n * n * D
    for (br, bc) in ((4096, 4096), (8192, 4096), (8192, 8192), (16384, 4096), (16384, 8192), (32768, 4096), (4736, 4736)):
        t_b = timeit(lambda: ops.infonce_backward_raw(ab, bb, s, rs, cs, gl, 0.5 / n, 0, "bf16", br, bc))
        import time as _t
        t0 = _t.perf_counter()
        ops.infonce_backward_raw(ab, bb, s, rs, cs, gl, 0.5 / n, 0, "bf16", br, bc)
        host = (_t.perf_counter() - t0) * 1e3
        torch.cuda.synchronize()
        print(f"    bwd block {br}x{bc}: {t_b:.3f} ms ({3 * fl / t_b / 1e9:.1f} TFLOP/s executed) host-side enqueue {host:.3f} ms")


def sec_one_bwd():
    """one forward + one backward at n=32768 (for the ncu launch list)"""
    n, D = 32768, 512
    g = torch.Generator(device="cpu").manual_seed(6)
    a = torch.nn.functional.normalize(torch.randn(n, D, generator=g), dim=1).to(dev)
    b = torch.nn.functional.normalize(torch.randn(n, D, generator=g), dim=1).to(dev)
    ab, bb = a.bfloat16(), b.bfloat16()
    s = torch.tensor(1 / 0.07, device=dev)
    rs, cs, dg = ops.infonce_forward_raw(ab, bb, s, 0, "bf16")
    gl = torch.ones((), device=dev)
    ops.infonce_backward_raw(ab, bb, s, rs, cs, gl, 0.5 / n, 0, "bf16")
    torch.cuda.synchronize()


SECTIONS = {
    "bf16_err": sec_bf16_err,
    "blocks": sec_blocks,
    "one_bwd": sec_one_bwd,
    "gemm_fp32": sec_gemm_fp32,
    "small": sec_small,
    "infonce_fp32": lambda: sec_infonce("fp32", ((8, 64), (32, 512), (200, 256), (1024, 512), (2500, 512))),
    "linear_fp32": lambda: sec_linear("fp32"),
    "zeroshot": sec_zeroshot,
    "gemm_bf16": sec_gemm_bf16,
    "infonce_bf16": lambda: sec_infonce("bf16"),
    "linear_bf16": lambda: sec_linear("bf16"),
    "perf": sec_perf,
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(SECTIONS)
    print("device:", torch.cuda.get_device_name(0), "lib version", _lib.load().mmg_version(), "info", _lib.device_info())
    for name in names:
        print(f"[{name}]")
        t0 = time.time()
        try:
            SECTIONS[name]()
            torch.cuda.synchronize()
            print(f"  ok ({time.time() - t0:.1f}s)")
        except Exception:  # noqa: BLE001
            traceback.print_exc()
            print(f"  FAILED ({time.time() - t0:.1f}s)")
            try:
                torch.cuda.synchronize()
            except Exception as e:  # sticky CUDA error: nothing after this can run
                print("  CUDA context is dead:", e)
                break
    sys.stdout.flush()
