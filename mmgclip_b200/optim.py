"""Fused multi-tensor AdamW for the head weights.

The reference trains with ``torch.optim.AdamW(self.model.parameters(), lr=..., weight_decay=...)`` and calls
``self.optimizer.step()`` right after ``loss.backward()`` (mmgclip/experiments/ClassifierExperiment.py:74,118).
:class:`FusedAdamW` is a ``torch.optim.Optimizer`` with the same constructor and the same update (torch's operation
order, bias corrections in double) that updates *all* tensors of a parameter group in one kernel launch through
``mmg_adamw_step``.  The step counter -- and, with ``capturable=True``, the learning rate -- live on the device, so the
whole training step (forward, loss, backward, optimizer) can be recorded once by :class:`mmgclip_b200.graph.GraphedStep`
and replayed; LR schedulers keep working: they write ``group['lr']`` as usual and :meth:`sync_lr` (called by ``step()``
whenever the stream is not capturing) mirrors it into the device scalar.

There is no CPU path: parameters must be fp32 CUDA tensors.
"""
from __future__ import annotations

import ctypes
from typing import Iterable, Tuple

import torch

from . import _lib
from ._lib import check


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, capturable: bool = False):
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        if not 0.0 <= weight_decay:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, capturable=capturable))
        self.kernel_launches = 0

    # per-group device scalars: step_state int64[2] = {t, ticket}, lr fp32.  Kept outside param_groups so that
    # state_dict() stays what torch.optim.AdamW's is: per-parameter {step, exp_avg, exp_avg_sq} + plain group options.
    def _group_state(self, gi, group, params, device):
        dev_state = self.__dict__.setdefault("_mmg_dev", {})
        st = dev_state.get(gi)
        if st is None or st["step"].device != device:
            t0 = 0
            for p in params:  # a loaded checkpoint (ours or torch.optim.AdamW's) carries the step count per parameter
                step = self.state.get(p, {}).get("step")
                if step is not None:
                    t0 = max(t0, int(step.item()) if torch.is_tensor(step) else int(step))
            step_state = torch.zeros(2, dtype=torch.int64, device=device)
            step_state[0] = t0
            st = {"step": step_state,
                  "lr": torch.full((), float(group["lr"]), dtype=torch.float32, device=device),
                  "lr_host": float(group["lr"])}
            dev_state[gi] = st
        return st

    def sync_lr(self) -> None:
        """Mirror every group's ``lr`` into its device scalar (call between graph replays after a scheduler step)."""
        for gi, group in enumerate(self.param_groups):
            st = self.__dict__.get("_mmg_dev", {}).get(gi)
            if st is not None and st["lr_host"] != float(group["lr"]):
                st["lr"].fill_(float(group["lr"]))
                st["lr_host"] = float(group["lr"])

    def steps_taken(self, group_index: int = 0) -> int:
        st = self.__dict__.get("_mmg_dev", {}).get(group_index)
        return 0 if st is None else int(st["step"][0].item())

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self.__dict__["_mmg_dev"] = {}  # device counters are rebuilt from the loaded per-parameter step

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing:
            self.sync_lr()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            dev = ps[0].device
            for p in ps:
                if not (p.is_cuda and p.dtype == torch.float32 and p.grad.dtype == torch.float32):
                    raise RuntimeError("FusedAdamW needs fp32 CUDA parameters and gradients (no CPU fallback)")
                if p.device != dev:
                    raise RuntimeError("all parameters of a group must live on one device")
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdamW does not support sparse gradients")
                if not p.is_contiguous():
                    raise RuntimeError("FusedAdamW needs contiguous parameters")
                state = self.state[p]
                if len(state) == 0:
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st = self._group_state(gi, group, ps, dev)
            for p in ps:
                self.state[p]["step"] = st["step"][0]  # view of the group's device counter
            grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in ps]
            n = len(ps)
            arr = ctypes.c_void_p * n
            P = arr(*[p.data_ptr() for p in ps])
            G = arr(*[g.data_ptr() for g in grads])
            M = arr(*[self.state[p]["exp_avg"].data_ptr() for p in ps])
            V = arr(*[self.state[p]["exp_avg_sq"].data_ptr() for p in ps])
            N = (ctypes.c_longlong * n)(*[p.numel() for p in ps])
            b1, b2 = group["betas"]
            lr_dev = st["lr"].data_ptr() if group["capturable"] else None
            n0 = lib.mmg_kernel_launch_count()
            with torch.cuda.device(dev):
                check(lib.mmg_adamw_step(P, G, M, V, N, n, float(group["lr"]), lr_dev, float(b1), float(b2),
                                         float(group["eps"]), float(group["weight_decay"]), st["step"].data_ptr(),
                                         torch.cuda.current_stream().cuda_stream), "mmg_adamw_step")
            self.kernel_launches += int(lib.mmg_kernel_launch_count() - n0)
        return loss
