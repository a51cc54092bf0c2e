#!/usr/bin/env python
"""Headline benchmark: CLIP loss fwd+bwd pairs/sec at global batch 32768 on 1/2/4/8 B200 (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N ...            # the reference arithmetic on the host CPU (oracle port)

A *step* is one pass of the hot path over one global batch of synthetic encoder features:
    LinearProjectionLayer heads (768 -> 512, image and text) -> L2 normalise -> exp(logit_scale) -> symmetric InfoNCE
    (CLIPLoss) forward -> backward down to the head-weight gradients            (reference: mmgclip_model.py:124-136,
    losses.py:36-44, ClassifierExperiment.py:109-115; optimizer step excluded, as in the metric's definition)
Rows are sharded across ranks (strong scaling: the global batch is fixed at 32768); text embeddings are all-gathered
(NCCL), column sums / loss / head gradients all-reduced (symmetric-memory kernels over NVLink), and the text-side gradients
reach their owner ranks inside the fused backward kernel (TMA reduce-add into NVLink peer memory; NCCL reduce-scatter
as the fallback).  The step is replayed as one CUDA graph.

`value`  : inputs already resident in HBM (fp32 features), timed with CUDA events on the launching stream, max over ranks.
`e2e`    : same call with features in pinned HOST memory, as a steady-state pipeline: every timed iteration issues one
           host->device batch upload (the next step's, on a copy stream), the step, and a device->host read of the loss;
           K uploads, K steps and K read-backs complete inside the timed region.
`roofline`: algorithmic FLOPs (6 B^2 D + 4 B (E_i+E_t) D, SURVEY s8d / DESIGN.md) / step time / GPUs vs the measured
           bf16 tensor peak in MEASURED_PEAKS.json (sustained figure: the kernels are timed inside a long step).
`cpu_baseline`: the oracle port of the reference step timed on this box's host cores on a bounded sample (rank 0, N=1).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GLOBAL_BATCH = 32768
E_IMG = E_TXT = 768
D_PROJ = 512
METRIC = "CLIP loss fwd+bwd pairs/sec (global batch 32k)"


def alg_flops(b, e_i, e_t, d):
    return 6.0 * b * b * d + 4.0 * b * (e_i + e_t) * d


def ncu_traffic_bytes(stored_e=False):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
    captures (profiles/r01c_ncu_top_kernels.md, r01d_ncu_stored_e_bwd.md); None if the summary is missing."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            key = "dominant_kernel_dram_bytes_per_launch" + ("_stored_e" if stored_e else "")
            return float(json.load(f)[key])
    except Exception:  # noqa: BLE001
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return {"burst": float(p["bf16_tflops"]), "sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "hbm": float(p["hbm_gbs"]), "source": "measured"}
    except Exception:  # noqa: BLE001
        return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


# ----------------------------------------------------------------------------------------------------------------
# synthetic inputs, generated here so that the measured arm touches nothing under oracle/:
# image features ~ clamp(1 + 0.35 N(0,1), min=0) (ConvNeXt avg-pool-like), text features ~ 0.5 N(0,1) (BERT-like), head
# weights ~ nn.Linear's default U(-1/sqrt(E), 1/sqrt(E)).  tests/test_boundary_cpu.py checks that the oracle's recipe
# (used by the CPU arm) yields the same arrays, so both arms see identical data.
# ----------------------------------------------------------------------------------------------------------------
def synthetic_features(batch, e_image=768, e_text=768, seed=42):
    import numpy as np
    rng = np.random.RandomState(seed)
    xi = np.maximum(1.0 + 0.35 * rng.standard_normal((batch, e_image)), 0.0).astype(np.float32)
    xt = (0.5 * rng.standard_normal((batch, e_text))).astype(np.float32)
    return xi, xt


def synthetic_head_weights(d, e_image=768, e_text=768, seed=43):
    import numpy as np
    rng = np.random.RandomState(seed)
    wi = rng.uniform(-1.0, 1.0, (d, e_image)).astype(np.float32) / np.float32(math.sqrt(e_image))
    wt = rng.uniform(-1.0, 1.0, (d, e_text)).astype(np.float32) / np.float32(math.sqrt(e_text))
    return wi, wt


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference step on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference(sample_batch, steps, warmup):
    """Time the reference's step on the host cores: its own projection.py / losses.py loaded by path where the reference
    tree exists (kind "reference"), else the operation-for-operation port in oracle/clip_oracle.py (kind "port")."""
    import torch
    from oracle import clip_oracle as oc
    from oracle import ref_loader
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    xi, xt = oc.synthetic_features(sample_batch, E_IMG, E_TXT, seed=42)
    wi, wt = oc.synthetic_head_weights(D_PROJ, E_IMG, E_TXT, seed=43)
    xi, xt, wi, wt = (torch.from_numpy(t) for t in (xi, xt, wi, wt))
    ls = torch.tensor(math.log(1 / 0.07))
    if ref_loader.load_reference() is not None:
        kind = "reference"
        fn = ref_loader.ReferenceStep(wi, wt, math.log(1 / 0.07))
        step = lambda: fn(xi, xt)  # noqa: E731
    else:
        kind = "port"
        step = lambda: oc.torch_train_step(xi, xt, wi, wt, ls)  # noqa: E731
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"pairs_per_s": sample_batch / dt, "ms_per_step": dt * 1e3, "cores": cores, "loss": float(out["loss"]),
            "threads": torch.get_num_threads(), "kind": kind, "steps": steps, "warmup": warmup}


def cpu_kind_text(kind):
    return ("the reference's own projection.py / losses.py loaded by path + mmgclip_model.py:124-136 restated"
            if kind == "reference" else "oracle port of the reference step (operation for operation, fp32 eager PyTorch)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = min(args.cpu_sample_batch or 16384, GLOBAL_BATCH)
    r = cpu_reference(sample, max(1, args.steps), max(0, args.warmup))  # exactly the steps / warm-ups that are printed
    desc = (f"{cpu_kind_text(r['kind'])} on the host CPU at batch {sample} of {GLOBAL_BATCH}, {r['threads']} threads, "
            f"{r['steps']} timed steps after {r['warmup']} warm-ups; cost per pair grows ~linearly with the batch (B x B "
            f"logits), so the full {GLOBAL_BATCH} batch would be ~{GLOBAL_BATCH / sample:.0f}x slower per pair")
    cfg = workload_config(args.gpus, sample_batch=sample)
    if sample != GLOBAL_BATCH:
        cfg["workload"] += f" -- CPU arm timed on a bounded sample: batch {sample}"
    line = {
        "impl": "reference", "metric": METRIC, "value": r["pairs_per_s"], "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "loss": r["loss"],
        "cpu_baseline": {"value": r["pairs_per_s"], "unit": "pairs/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": desc},
        "e2e": {"value": r["pairs_per_s"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------------
# parity of the timed step, checked outside the timed regions on the bench's own inputs (every N)
# ----------------------------------------------------------------------------------------------------------------
def parity_fp64(torch, dev, xi_g, xt_g, wi, wt, loss_gpu, dwi_gpu, dwt_gpu, chunk=2048):
    """Float64 closed form of the WHOLE step on the full global batch (heads mmgclip/networks/
    projection.py:33, normalise + logits mmgclip_model.py:128-136, CLIPLoss losses.py:36-44 and their gradients down to the
    head weights), evaluated with torch float64 on the GPU in row chunks (cuBLAS dgemm -- an independent implementation; a
    few hundred ms at B = 32768) and compared with what the timed step produced: the loss and both head-weight gradients
    (after the cross-rank all-reduce at N > 1).  Errors are max-abs / max-abs and Frobenius-relative."""
    f64 = torch.float64
    Xi, Xt = torch.from_numpy(xi_g).to(dev, f64), torch.from_numpy(xt_g).to(dev, f64)
    Wi, Wt = torch.from_numpy(wi).to(dev, f64), torch.from_numpy(wt).to(dev, f64)
    s = float(torch.tensor(math.log(1 / 0.07), dtype=torch.float32).exp())  # the fp32 scale the step uses
    ui, ut = Xi @ Wi.t(), Xt @ Wt.t()
    ni, nt = ui.norm(dim=1, keepdim=True), ut.norm(dim=1, keepdim=True)
    I, T = ui / ni, ut / nt
    B = I.shape[0]
    diag = s * (I * T).sum(1)
    rowsum = torch.empty(B, dtype=f64, device=dev)
    colsum = torch.zeros(B, dtype=f64, device=dev)
    for r0 in range(0, B, chunk):
        E = torch.exp(s * (I[r0:r0 + chunk] @ T.t()) - s)
        rowsum[r0:r0 + chunk] = E.sum(1)
        colsum += E.sum(0)
    loss = float(((torch.log(rowsum) + s - diag) + (torch.log(colsum) + s - diag)).sum() / (2 * B))
    coef = s / (2 * B)
    rinv, cinv = coef / rowsum, coef / colsum
    dI, dT = torch.empty_like(I), torch.zeros_like(T)
    for r0 in range(0, B, chunk):
        r1 = min(B, r0 + chunk)
        G = torch.exp(s * (I[r0:r1] @ T.t()) - s) * (rinv[r0:r1, None] + cinv[None, :])
        idx = torch.arange(r1 - r0, device=dev)
        G[idx, r0 + idx] -= 2 * coef
        dI[r0:r1] = G @ T
        dT += G.t() @ I[r0:r1]
    del G, E
    dui = (dI - I * (I * dI).sum(1, keepdim=True)) / ni
    dut = (dT - T * (T * dT).sum(1, keepdim=True)) / nt
    dWi, dWt = dui.t() @ Xi, dut.t() @ Xt

    def err(got, ref):
        d = got.to(f64) - ref
        return {"max_abs_over_max_abs": float(d.abs().max() / ref.abs().max()),
                "frobenius_rel": float(d.norm() / ref.norm())}
    out = {"checker": "float64 closed form of the whole step on the full global batch (torch float64 on the GPU, row "
                      "chunks; bench.py parity_fp64), outside the timed regions",
           "tolerance": 2e-3, "loss": loss_gpu, "loss_fp64": loss, "loss_rel_err": abs(loss_gpu - loss) / abs(loss),
           "dw_image": err(dwi_gpu, dWi), "dw_text": err(dwt_gpu, dWt)}
    out["ok"] = bool(out["loss_rel_err"] < 2e-3 and out["dw_image"]["frobenius_rel"] < 2e-3
                     and out["dw_text"]["frobenius_rel"] < 2e-3)
    return out


def gpu_eager_reference(torch, dev, batches, iters=5):
    """Informational: the reference's eager arithmetic (mmgclip/networks/projection.py:33, mmgclip_model.py:124-136,
    losses.py:36-44, loss.backward()) on the same GPU under the installed torch -- fp32 as the reference runs it, and under
    bf16 autocast.  Stock ATen / cuBLAS kernels; it materialises >= 6 [B, B] fp32 matrices (24+ GiB at B = 32768)."""
    import torch.nn.functional as F
    out = {}
    for Bn in batches:
        try:
            xi_h, xt_h = synthetic_features(Bn, E_IMG, E_TXT, seed=42)
            wi_h, wt_h = synthetic_head_weights(D_PROJ, E_IMG, E_TXT, seed=43)
            xi, xt = torch.from_numpy(xi_h).to(dev), torch.from_numpy(xt_h).to(dev)
            wi = torch.from_numpy(wi_h).to(dev).requires_grad_()
            wt = torch.from_numpy(wt_h).to(dev).requires_grad_()
            ls = torch.tensor(math.log(1 / 0.07), device=dev)
            labels = torch.arange(Bn, device=dev)

            def step():
                wi.grad = wt.grad = None
                ie, te = F.linear(xi, wi), F.linear(xt, wt)
                ie = ie / ie.norm(dim=1, keepdim=True)
                te = te / te.norm(dim=1, keepdim=True)
                s = ls.exp()
                lpi = s * ie @ te.t()
                lpt = s * te @ ie.t()
                loss = (F.cross_entropy(lpi, labels) + F.cross_entropy(lpt, labels)) / 2
                loss.backward()
                return loss
            for mode in ("fp32", "bf16_autocast"):
                ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode != "fp32" else torch.autocast("cuda", enabled=False)
                with ctx:
                    for _ in range(2):
                        step()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(iters):
                        loss = step()
                    e1.record()
                    torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters
                out[f"b{Bn}_{mode}"] = {"ms_per_step": ms, "pairs_per_s": Bn / (ms * 1e-3), "loss": float(loss.detach())}
            del xi, xt, wi, wt
            torch.cuda.empty_cache()
        except Exception as exc:  # noqa: BLE001  (informational block: never lose the bench line over it)
            out[f"b{Bn}_error"] = repr(exc)[:200]
    out["what"] = ("the reference's eager ops on cuda:0 under torch " + torch.__version__ + " (stock ATen/cuBLAS kernels), "
                   "device-resident inputs, CUDA-event timing; informational, outside the timed regions")
    return out


def workload_config(n_gpus, sample_batch=None, n_sets=None, graph=None, peer=None, symm=None, stored_e=None):
    cfg = {"workload": f"mmg-clip hot path: LinearProjectionLayer heads {E_IMG}->{D_PROJ} (image, text) + L2 normalise + "
                       f"symmetric CLIPLoss fwd+bwd to head-weight grads, global batch {GLOBAL_BATCH}, synthetic "
                       f"ConvNeXt-like / BERT-like features",
           "global_batch": GLOBAL_BATCH, "embedding_dim": E_IMG, "projection_dim": D_PROJ,
           "parallelism": f"row-sharded x{n_gpus} (all-gather text embeddings, all-reduce column sums, " + (
               "dT slices reduce-added into their owner's buffer over NVLink peer memory by the fused backward kernel)"
               if peer else "reduce-scatter dT)"),
           "l2": "step inputs rotate over distinct buffer sets; each step also streams > 126 MB of intermediates "
                 "(the 64 MiB coefficient scratch, 128 MiB of fp32 gradients, 2 x 32 MiB operand copies), so nothing survives in the "
                 "126 MB L2 between steps"}
    if symm:
        cfg["small_allreduces"] = ("torch symmetric memory one-shot / two-shot kernels over NVLink (column sums + the loss's "
                                   "row part in ONE sum, head grads)")
    if n_sets is not None:
        cfg["input_sets"] = n_sets
    if graph is not None:
        cfg["launch"] = ("one CUDA graph replay per step (mmgclip_b200.graph.GraphedStep)" if graph
                         else "eager (one host launch per kernel)")
    if stored_e is not None:
        cfg["backward"] = ("stored-E: the forward keeps exp(logit - s) as bf16 [rows, cols] (2*rows*cols bytes of HBM) and "
                           "the fused backward transforms it into the gradient coefficients" if stored_e else
                           "recompute: the fused backward recomputes the cosines on the tensor cores (O(B*D) memory)")
    if sample_batch is not None:
        cfg["cpu_sample_batch"] = sample_batch
    return cfg


# ----------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock, power and clock-event (throttle) reasons of one GPU from a background thread through NVML
    while the timed regions run (the recipe's nvidia-smi query, without the process start-up latency)."""

    def __init__(self, gpu_index, period_s=0.005):
        import threading
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.ok = False
        self._stop = threading.Event()
        self._active = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                try:
                    phys = int(vis.split(",")[gpu_index])
                except (ValueError, IndexError):
                    phys = gpu_index
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False
        self.period = period_s
        self.thread = threading.Thread(target=self._run, daemon=True)
        if self.ok:
            self.thread.start()

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
            nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                    self.samples.append((mhz, pw))
                    r = int(get_reasons(self.h))
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:  # noqa: BLE001
                    pass
            time.sleep(self.period)

    def begin(self):
        self._active.set()

    def end(self):
        self._active.clear()

    def stop(self):
        self._stop.set()
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.samples:
            mhz = sorted(m for m, _ in self.samples)
            out["sm_mhz"] = mhz[len(mhz) // 2]
            out["sm_mhz_min"] = mhz[0]
            out["power_w_max"] = max(p for _, p in self.samples)
            out["how"] = "NVML polled every 5 ms from a thread during both timed regions (median)"
        return out


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from mmgclip_b200 import _lib, ops
    from mmgclip_b200.distributed import (allreduce_gradients, gather_columns_async, peer_reduce_active,
                                          push_gather_active, sharded_info_nce, symm_allreduce_active)
    from mmgclip_b200.projection import LinearProjectionLayer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} processes (one per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: mmgclip_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    if not args.pdl:
        ops.set_tuning(pdl=0)
    if args.emb_f16 is None:
        args.emb_f16 = int(os.environ.get("WORLD_SIZE", "1")) == 1 and args.batch <= 8192
    ops.set_embedding_f16(bool(args.emb_f16))

    B = args.batch
    if B % world:
        raise SystemExit("global batch must divide evenly across ranks")
    bl = B // world
    prec = args.precision

    # identical global tensors on every rank, contiguous row shard per rank
    xi_g, xt_g = synthetic_features(B, E_IMG, E_TXT, seed=42)
    wi, wt = synthetic_head_weights(D_PROJ, E_IMG, E_TXT, seed=43)
    xi_h = torch.from_numpy(xi_g[rank * bl:(rank + 1) * bl].copy())
    xt_h = torch.from_numpy(xt_g[rank * bl:(rank + 1) * bl].copy())
    if rank != 0 or not args.parity:
        del xi_g, xt_g  # rank 0 keeps the global batch on the host for the float64 parity check after the timed regions
    step_bytes = (xi_h.numel() + xt_h.numel()) * 4
    n_sets = max(2, -(-(256 << 20) // step_bytes))
    n_sets = min(n_sets, 16)
    dev_sets = [(xi_h.to(dev), xt_h.to(dev)) for _ in range(n_sets)]
    host_sets = [(xi_h.clone().pin_memory(), xt_h.clone().pin_memory()) for _ in range(min(n_sets, 4))]

    head_i = LinearProjectionLayer(E_IMG, D_PROJ, precision=prec).to(dev)
    head_t = LinearProjectionLayer(E_TXT, D_PROJ, precision=prec).to(dev)
    with torch.no_grad():
        head_i.layer.weight.copy_(torch.from_numpy(wi))
        head_t.layer.weight.copy_(torch.from_numpy(wt))
    logit_scale = torch.tensor(math.log(1 / 0.07), device=dev).exp()
    group = None

    if args.head_overlap is None:
        args.head_overlap = bl <= 8192
    side = torch.cuda.Stream(device=dev) if args.head_overlap else None

    def step(xi, xt):
        head_i.layer.weight.grad = None
        head_t.layer.weight.grad = None
        ops.mark("start")
        if side is not None:
            # the two heads are independent: the text head runs on a side stream so its bandwidth-bound kernels (cast,
            # normalise) overlap the other head's contraction -- and autograd replays each head's backward on the
            # stream its forward ran on, so the same overlap happens there
            cur = torch.cuda.current_stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                te = head_t.forward_normalized(xt)
                gathered = gather_columns_async(te, group=group, prec=prec) if world > 1 else None
            ie = head_i.forward_normalized(xi)
            cur.wait_stream(side)
            te.record_stream(cur)
            if getattr(te, "_mmg_bf16", None) is not None:
                te._mmg_bf16.record_stream(cur)
        else:
            te = head_t.forward_normalized(xt)
            ops.mark("head_t")
            gathered = gather_columns_async(te, group=group, prec=prec) if world > 1 else None  # overlaps the image head
            ie = head_i.forward_normalized(xi)
            ops.mark("head_i")
        loss = sharded_info_nce(ie, te, logit_scale, group=group, prec=prec, gathered=gathered)
        ops.mark("loss")
        loss.backward()
        ops.mark("backward")
        if world > 1:
            allreduce_gradients(head_i, head_t, group=group)
        ops.mark("grad_ar")
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    W, K = max(args.warmup, 3), args.steps
    # live timing of the dominant launch (the fused backward) inside the replayed graphs: external events are captured as
    # event-record nodes, so after a replay they hold that replay's timestamps
    probe = None
    if rank == 0:
        try:
            probe = (torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True))
            ops.set_backward_probe(probe)
        except Exception:  # noqa: BLE001  (older torch: no external events)
            probe = None
    timeline = {} if args.timeline else None
    ops.set_timeline(timeline)
    # e2e staging buffers (the H2D copies land here); allocated before any capture so they can be graph inputs
    stage = [(torch.empty_like(dev_sets[0][0]), torch.empty_like(dev_sets[0][1])) for _ in range(2)]
    # ---- one CUDA graph per input set: a step is a single cudaGraphLaunch (mmgclip_b200/graph.py) ----
    gstep = None
    if args.graph:
        from mmgclip_b200.graph import GraphedStep
        n_sets = min(n_sets, 4)
        dev_sets = dev_sets[:n_sets]
        gstep = GraphedStep(step, list(dev_sets) + stage, warmup=3,
                            params=list(head_i.parameters()) + list(head_t.parameters()))

    def run_step(i):
        if gstep is not None:
            return gstep(i % n_sets)
        return step(*dev_sets[i % n_sets])

    def run_stage_step(s):
        if gstep is not None:
            return gstep(n_sets + s)
        return step(*stage[s])

    # ---- device-resident timing ("value") ----
    for i in range(W):
        loss = run_step(i)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler is not None:
        sampler.begin()
    launches0 = _lib.load().mmg_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        loss = run_step(i)
    e1.record()
    barrier()
    ms_value = max_over_ranks(e0.elapsed_time(e1) / K)
    if sampler is not None:
        sampler.end()
    launches = _lib.load().mmg_kernel_launch_count() - launches0
    if gstep is not None:
        launches = K * gstep.kernel_launches  # replays issue no host-side launches; counted while recording
    live_bwd_ms = None
    if probe is not None:
        try:
            live_bwd_ms = probe[0].elapsed_time(probe[1])  # the LAST timed step's backward launch, no profiler
        except Exception:  # noqa: BLE001
            live_bwd_ms = None
    loss_value = float(loss.item())
    dw_value = (head_i.layer.weight.grad.detach().clone(), head_t.layer.weight.grad.detach().clone())
    if timeline is not None:
        # phase boundaries of the LAST timed step on every rank (ms since that rank's "start" mark; external events
        # captured into the replayed graph) -> gpurun_out/timeline_n<N>.json.  Diagnostic output, not part of the JSON line.
        torch.cuda.synchronize()
        t0 = timeline.get("start")
        mine = {}
        for k, ev in timeline.items():
            try:
                mine[k] = t0.elapsed_time(ev)
            except Exception:  # noqa: BLE001  (a mark that the last replay did not pass)
                pass
        allm = [mine]
        if world > 1:
            allm = [None] * world
            dist.all_gather_object(allm, mine)
        if rank == 0:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", f"timeline_n{world}.json"), "w") as f:
                json.dump({"ms_per_step": ms_value, "ranks": [dict(sorted(m.items(), key=lambda kv: kv[1])) for m in allm]}, f,
                          indent=1)
        ops.set_timeline(None)

    # ---- end-to-end timing: pinned host features -> H2D (prefetched one step ahead) -> step -> loss D2H ----
    copy_stream = torch.cuda.Stream(device=dev)
    compute = torch.cuda.current_stream(dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    loss_host = torch.zeros(K + W + 2, dtype=torch.float32).pin_memory()

    def prefetch(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            hx, ht = host_sets[i % len(host_sets)]
            stage[s][0].copy_(hx, non_blocking=True)
            stage[s][1].copy_(ht, non_blocking=True)
            ready[s].record(copy_stream)

    def e2e_loop(n_steps, base, first_in_flight):
        """n_steps steps of the steady-state pipeline: while step i computes, the copy stream uploads batch i+1.  Every
        iteration issues exactly one H2D batch copy (the NEXT step's) and one loss D2H; the batch of the first step is
        already in flight when `first_in_flight` (it was issued by the previous loop's last iteration)."""
        if not first_in_flight:
            prefetch(base)
        for i in range(n_steps):
            s = (base + i) % 2
            prefetch(base + i + 1)
            compute.wait_event(ready[s])
            l = run_stage_step(s)
            consumed[s].record(compute)
            loss_host[base + i].copy_(l.detach(), non_blocking=True)
        # the upload issued by the last iteration belongs to this loop's bytes: its completion is inside the region
        compute.wait_event(ready[(base + n_steps) % 2])

    for s in range(2):
        consumed[s].record(compute)
    W2 = W + (W % 2)  # keep the double-buffer parity aligned
    e2e_loop(W2, 0, False)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler is not None:
        sampler.begin()
    e2.record()
    e2e_loop(K, W2, True)
    e3.record()
    barrier()
    if sampler is not None:
        sampler.end()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3) / K)
    clocks = sampler.stop() if sampler is not None else None
    # peak device memory of the whole run (max over ranks): O(B*D) by design -- a [rows, B] fp32 logit matrix alone would be
    # 4*bl*B bytes per rank (the reference materialises >= 6 such matrices at world = 1)
    peak_mem = float(torch.cuda.max_memory_allocated(dev))
    if world > 1:
        t = torch.tensor([peak_mem], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        peak_mem = float(t.item())

    # ---- per-launch timing of the dominant kernels (informational; outside the timed regions) ----
    kernels = None
    if rank == 0 and args.kernel_breakdown:
        kernels = kernel_breakdown(torch, ops, dev, bl, B, D_PROJ,
                                   stored_e=world == 1 and ops.want_store_e(bl, B, D_PROJ, prec, False))
    parity = None
    if rank == 0 and args.parity:
        try:
            parity = parity_fp64(torch, dev, xi_g, xt_g, wi, wt, loss_value, dw_value[0], dw_value[1])
        except Exception as exc:  # noqa: BLE001  (never lose the bench line over the checker)
            parity = {"error": repr(exc)[:300]}
        del xi_g, xt_g
        torch.cuda.empty_cache()
    eager = None
    if rank == 0 and world == 1 and args.gpu_eager and prec == "bf16":
        eager = gpu_eager_reference(torch, dev, sorted({4096, B}) if B <= 32768 else [4096])

    def finish(code=0):
        # Multi-GPU teardown: NCCL communicators that were captured into CUDA graphs do not always unwind cleanly
        # (destroy_process_group was seen to hang after the result line was printed), so every rank synchronises, flushes
        # and leaves without running the communicator destructors.
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            os._exit(code)
        return code

    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
    if rank != 0:
        return finish(0)

    peaks = measured_peaks()
    fl = alg_flops(B, E_IMG, E_TXT, D_PROJ)
    ach = fl / (ms_value * 1e-3) / world / 1e12
    dom = kernels["backward_fused"] if kernels else None
    # stored-E mode (ops.want_store_e): the single-GPU loss keeps E in the forward and the backward does no recomputation
    stored_e = world == 1 and ops.want_store_e(bl, B, D_PROJ, prec, False)
    if live_bwd_ms is not None and live_bwd_ms > 0:
        f_bwd = 4.0 * bl * B * D_PROJ  # dI + dT of this rank's rows (the recomputed cosines are not counted)
        dom = {"kernel": "infonce_bwd_fused_kernel (one persistent launch = the whole InfoNCE backward of the step"
                         + ("; stored-E mode: coefficients transformed from the forward's bf16 E, no recomputation)"
                            if stored_e else ")"),
               "tflops": f_bwd / (live_bwd_ms * 1e-3) / 1e12, "ms_per_launch": live_bwd_ms,
               "flops_per_launch": f_bwd,
               "tflops_executed": (1.0 if stored_e else 1.5) * f_bwd / (live_bwd_ms * 1e-3) / 1e12,
               "how": "CUDA events recorded around the launch on its own stream inside the timed region (captured into "
                      "the replayed graph as external event-record nodes); value of the last timed step"}
    line = {
        "metric": METRIC, "value": B / (ms_value * 1e-3), "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_value, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": ("f16" if ops.get_embedding_f16() else "bf16") if prec == "bf16" else "f32", "data": "synthetic",
        "config": dict(workload_config(world, n_sets=n_sets, graph=gstep is not None, peer=peer_reduce_active(),
                                       symm=symm_allreduce_active(), stored_e=stored_e),
                       heads="the two heads run on two streams (forward and backward)" if side is not None
                       else "the two heads run back to back on one stream",
                       operands=("InfoNCE contractions: fp16 normalised embeddings x fp16 (2^14-scaled) gradient coefficients, "
                                 "fp32 accumulation; heads bf16x3" if (prec == "bf16" and ops.get_embedding_f16()) else
                                 ("InfoNCE contractions: bf16 embeddings x bf16 coefficients, fp32 accumulation; heads bf16x3"
                                  if prec == "bf16" else "fp32 FFMA")),
                       gather=("push over NVLink peer memory (mmg_push_rows into every rank's symmetric buffer + barrier) on a "
                               "communication stream" if push_gather_active() else
                               ("NCCL all-gather (asynchronous)" if world > 1 else "none (one GPU)"))),
        "loss": loss_value,
        "parity": parity,
        "clocks": clocks,
        "e2e": {"value": B / (ms_e2e * 1e-3), "unit": "pairs/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": step_bytes * world, "d2h_bytes_per_step": 4 * world,
                "how": "steady-state pipeline through the public modules: K iterations, each issuing one pinned-host -> "
                       "device batch upload (the next step's, on a copy stream) + the step + the loss read-back; K "
                       "uploads complete inside the timed region (CUDA events, max over ranks)"},
        "gpu_launches": int(launches),
        "memory": {"max_allocated_bytes_per_rank": peak_mem, "one_logit_block_fp32_bytes_per_rank": 4.0 * bl * B,
                   "note": "max over ranks of torch.cuda.max_memory_allocated for the whole run (all input sets, staging "
                           "buffers and graphs included); no [rows, B] matrix is ever allocated on the default path"},
        # dominant kernel = the fused persistent backward launch (~65% of the step): algorithmic FLOPs per launch (dI + dT;
        # the recomputed cosines are not counted) / live CUDA-event launch duration, against the sustained bf16 peak (it
        # runs inside a long step).  `whole_step` is the same ratio for the entire step (all kernels, algorithmic FLOPs
        # only) -- the number the metric's "% of peak" means.
        "roofline": {"bound": "tensor",
                     "kernel": dom["kernel"] if dom else "gemm_tc_kernel",
                     "achieved": dom["tflops"] if dom else ach, "peak": peaks["burst"], "unit": "TFLOP/s",
                     "frac": (dom["tflops"] if dom else ach) / peaks["burst"],
                     "frac_of_sustained_peak": (dom["tflops"] if dom else ach) / peaks["sustained"],
                     "peak_source": peaks["source"] + " (MEASURED_PEAKS.json bf16_tflops = burst cuBLAS figure, the one "
                                    "SURVEY s8d defines the fraction against; sustained = " + str(peaks["sustained"]) + ")",
                     "traffic": ncu_traffic_bytes(stored_e),
                     "dominant_kernel_live": dom if (live_bwd_ms is not None and live_bwd_ms > 0) else None,
                     "whole_step": {"algorithmic_flops_per_step": fl, "achieved": ach, "frac": ach / peaks["burst"],
                                    "frac_of_sustained_peak": ach / peaks["sustained"]},
                     "kernels": kernels},
    }
    if eager is not None:
        line["reference_gpu_eager"] = eager
    if world == 1 and not args.no_cpu_baseline:
        sample = min(args.cpu_sample_batch or 8192, B)
        r = cpu_reference(sample, 3, 1)
        line["cpu_baseline"] = {
            "value": r["pairs_per_s"], "unit": "pairs/s", "cores": r["cores"], "kind": r["kind"],
            "sample": f"{cpu_kind_text(r['kind'])}, {r['threads']} threads, at batch {sample} of {B}, 3 timed steps after "
                      f"1 warm-up; per-pair cost grows ~linearly with the batch"}
    print(json.dumps(line))
    return finish(0)


def kernel_breakdown(torch, ops, dev, rows, cols, d, stored_e=False):
    """Live CUDA-event timing (no profiler) of each tcgen05 kernel family at the bench shape (local rows x global
    columns), outside the timed regions."""
    gen = torch.Generator(device=dev).manual_seed(7)
    a = torch.nn.functional.normalize(torch.randn(rows, d, device=dev, generator=gen), dim=1)
    b = torch.nn.functional.normalize(torch.randn(cols, d, device=dev, generator=gen), dim=1)
    ab, bb = ops.cast_embedding(a), ops.cast_embedding(b)
    s = torch.tensor(1 / 0.07, device=dev)
    one = torch.ones((), device=dev)

    def timeit(fn, iters=5):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    t_f = timeit(lambda: ops.infonce_forward_raw(ab, bb, s, 0, "bf16"))
    rs, cs, _ = ops.infonce_forward_raw(ab, bb, s, 0, "bf16")
    bwd = lambda: ops.infonce_backward_raw(ab, bb, s, rs, cs, one, 0.5 / cols, 0, "bf16", need_dscale=False)  # noqa: E731
    t_b = timeit(bwd)                       # prep + zero fills + the fused persistent launch
    f = 2.0 * rows * cols * d
    out = {
        "forward_lse": {"kernel": "gemm_tc_kernel<256, EpiLse, 2>", "launches": 1, "ms_per_launch": t_f,
                        "flops_per_launch": f, "tflops": f / t_f / 1e9},
        "backward_fused": {"kernel": "infonce_bwd_fused_kernel (coefficient tiles + dI/dT slices, one persistent launch)",
                           "launches": 1, "ms_per_launch": t_b, "flops_per_launch": 2 * f,
                           "flops_per_launch_executed": 3 * f, "tflops": 2 * f / t_b / 1e9,
                           "tflops_executed": 3 * f / t_b / 1e9,
                           "note": "algorithmic = dI + dT (4 rows cols D); executed adds the recomputed cosines; the "
                                   "timing includes the prep kernel and the zero fills of the call"},
    }
    if stored_e:
        # the step itself ran in stored-E mode (ops.want_store_e): the two entries above are the recompute kernels, kept
        # for comparison; these are the kernels of the step
        out["forward_lse"]["mode"] = out["backward_fused"]["mode"] = "recompute (not what the timed step ran)"
        try:
            e_mat = torch.empty((rows, cols), dtype=torch.bfloat16, device=dev)
            t_fs = timeit(lambda: ops.infonce_forward_raw(ab, bb, s, 0, "bf16", e_out=e_mat))
            rs2, cs2, dg2 = ops.infonce_forward_raw(ab, bb, s, 0, "bf16", e_out=e_mat)
            t_bs = timeit(lambda: ops.infonce_backward_raw(ab, bb, s, rs2, cs2, one, 0.5 / cols, 0, "bf16", a32=a, b32=b,
                                                           diag=dg2, need_dscale=False, e_stored=e_mat))
            out["forward_lse_store_e"] = {
                "kernel": "gemm_tc_kernel<256, EpiLse<store E>, 2> (row/column sums + bf16 E through TMA stores)",
                "launches": 1, "ms_per_launch": t_fs, "flops_per_launch": f, "tflops": f / t_fs / 1e9,
                "bytes_written": 2.0 * rows * cols}
            out["backward_fused_stored_e"] = {
                "kernel": "infonce_bwd_fused_kernel<256, 8, 8> (E -> coefficients on transform warps + dI/dT slices)",
                "launches": 1, "ms_per_launch": t_bs, "flops_per_launch": 2 * f, "tflops": 2 * f / t_bs / 1e9,
                "note": "no recomputation: executed = algorithmic; the timing includes the prep and matching-pair kernels"}
            del e_mat
        except Exception as exc:  # noqa: BLE001  (informational section: never lose the bench line over it)
            out["stored_e_breakdown_error"] = repr(exc)
    # the block loop the fused launch replaces (general shapes still use it)
    try:
        ops.set_tuning(fused=0)
        t_loop = timeit(bwd)
    finally:
        ops.set_tuning()
    block = 8192
    n_blocks = (-(-rows // block)) * (-(-cols // block))
    out["backward_block_loop"] = {
        "total_ms": t_loop, "launches": 2 * n_blocks,
        "kernels": "gemm_tc_kernel<256, EpiGrad, 2> (coefficients of one 8192 x 8192 block) + gemm_tc_kernel<256, EpiStoreF32, 2> "
                   "(its dI and dT) per block"}
    return out


def run_zeroshot(args):
    """BASELINE config 4 (secondary line, not the headline): zero-shot prompt scoring of N image embeddings against 64
    prompts, D = 512, fp32 -> argmax + top-5.  HBM-roofline: algorithmic bytes N*D*4 + C*D*4 + N*(8 + 5*12)."""
    import torch
    from mmgclip_b200 import _lib, ops
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    n, c, d, k = args.zeroshot_rows, 64, 512, 5
    gen = torch.Generator(device=dev).manual_seed(4)
    sets = [torch.nn.functional.normalize(torch.randn(n, d, device=dev, generator=gen), dim=1) for _ in range(2)]
    txt = torch.nn.functional.normalize(torch.randn(c, d, device=dev, generator=gen), dim=1)
    s = torch.tensor(1 / 0.07, device=dev)
    impl = args.zeroshot_impl
    fn = lambda i: ops.zeroshot_score(sets[i % 2], txt, s, k=k, want_logits=False, want_probs=False, impl=impl)  # noqa: E731
    for i in range(max(args.warmup, 3)):
        fn(i)
    torch.cuda.synchronize()
    n0 = _lib.load().mmg_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    peaks = measured_peaks()
    alg_bytes = n * d * 4 + c * d * 4 + n * (8 + k * 12)
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    print(json.dumps({
        "metric": "zero-shot prompt scoring rows/sec (1M x 64 prompts, D=512, fp32, argmax + top-5)",
        "value": n / (ms * 1e-3), "unit": "rows/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": f"zero-shot scoring N={n} C={c} D={d} k={k}",
                                         "l2": "two 2 GiB embedding sets alternate (larger than L2)"},
        "gpu_launches": int(_lib.load().mmg_kernel_launch_count() - n0),
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                     "traffic": None, "algorithmic_bytes_per_launch": alg_bytes,
                     "kernel": "zeroshot_kernel (fp32 FFMA)" if impl == "ffma" else
                               "zeroshot_tc_kernel (tcgen05 kind::tf32, 3xTF32 split in shared memory; + a 2 us prompt prep)"}}))
    return 0


def main():
    global D_PROJ, GLOBAL_BATCH
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=GLOBAL_BATCH, help="global batch (default: the metric's 32768)")
    ap.add_argument("--dim", type=int, default=D_PROJ,
                    help="projection dimension D (default: the metric's 512; BASELINE config 5 uses 1024)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-sample-batch", type=int, default=0,
                    help="batch of the CPU arm's bounded sample (default: 8192 for the cpu_baseline block, 16384 for "
                         "--impl reference)")
    ap.add_argument("--no-parity", dest="parity", action="store_false", default=True,
                    help="skip the float64 parity check of the timed step (rank 0, outside the timed regions)")
    ap.add_argument("--no-gpu-eager", dest="gpu_eager", action="store_false", default=True,
                    help="skip the informational reference_gpu_eager block (N = 1 only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-breakdown", action="store_true", default=True)
    ap.add_argument("--no-kernel-breakdown", dest="kernel_breakdown", action="store_false")
    ap.add_argument("--graph", dest="graph", action="store_true", default=os.environ.get("MMGCLIP_BENCH_GRAPH", "1") != "0",
                    help="replay the step as a CUDA graph (default)")
    ap.add_argument("--no-graph", dest="graph", action="store_false")
    ho_env = os.environ.get("MMGCLIP_BENCH_HEAD_OVERLAP")
    ap.add_argument("--head-overlap", dest="head_overlap", action="store_true",
                    default=None if ho_env is None else ho_env == "1",
                    help="run the text head on a side stream (forward and, through autograd, backward); default: on when a "
                         "rank holds <= 8192 rows (launch-bound heads: 0.212 -> 0.184 ms/step at batch 4096), off above "
                         "(bandwidth-bound heads: no gain measured)")
    ap.add_argument("--no-head-overlap", dest="head_overlap", action="store_false")
    ef_env = os.environ.get("MMGCLIP_B200_EMB_F16")
    ap.add_argument("--emb-f16", dest="emb_f16", action="store_true", default=None if ef_env is None else ef_env == "1",
                    help="fp16 tensor-core operands for the normalised embeddings and the (2^14-scaled) gradient "
                         "coefficients instead of bf16 (MMG_PREC_F16).  Default: fp16 on one GPU up to batch 8192 (there the "
                         "bf16 operand rounding puts the head-weight gradients at the 2e-3 bar -- 2.1e-3 at batch 4096 -- "
                         "and the step is launch-bound, so fp16 costs nothing: 0.171 vs 0.170 ms), bf16 above (fp16 "
                         "multipliers draw more power: 4.35 vs 4.15 ms per step at batch 32768 under the power cap, where "
                         "bf16 is already inside the bar)")
    ap.add_argument("--emb-bf16", dest="emb_f16", action="store_false")
    ap.add_argument("--no-pdl", dest="pdl", action="store_false", default=True,
                    help="A/B switch: launch without programmatic dependent launch (mmg_tune pdl=0)")
    ap.add_argument("--timeline", action="store_true",
                    help="diagnostic: record phase-boundary events in the step and dump them to gpurun_out/timeline_n<N>.json")
    ap.add_argument("--workload", default="clip", choices=["clip", "zeroshot"],
                    help="clip = the headline metric (default); zeroshot = BASELINE config 4 (secondary line)")
    ap.add_argument("--zeroshot-rows", type=int, default=1 << 20)
    ap.add_argument("--zeroshot-impl", default="auto", choices=["auto", "tc", "ffma"])
    args = ap.parse_args()
    D_PROJ = args.dim
    GLOBAL_BATCH = args.batch  # the config block of the JSON line names the batch that was actually run
    if args.workload == "zeroshot":
        return run_zeroshot(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
