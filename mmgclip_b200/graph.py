"""CUDA-graph capture of a whole training step of the hot path.

One step of the path (two projection heads -> L2 normalise -> fused InfoNCE forward -> backward down to the head-weight
gradients, plus the exchange steps of the row-sharded variant) is ~60 kernel launches and, on 8 GPUs, well under a
millisecond of device time: issued one by one from Python the host is the bottleneck.  :class:`GraphedStep` records the
step once per *input set* into a ``torch.cuda.CUDAGraph`` (the C ABI launches on the stream it is handed, so the capture
stream sees every kernel, memset and NCCL collective) and afterwards a step is a single ``cudaGraphLaunch``.

The reference's loop (ClassifierExperiment.py:109-118: forward, ``criterion(**outputs)``, ``loss.backward()``) is what gets
captured -- it is the same Python, run once under capture.  Usage::

    step = GraphedStep(lambda xi, xt: train_step(xi, xt), input_sets=[(xi0, xt0), (xi1, xt1)],
                       params=list(head_i.parameters()) + list(head_t.parameters()))
    loss = step(0)          # replays the graph recorded for input set 0 (reads whatever xi0 / xt0 hold now)
    step.copy_in(1, new_xi, new_xt); loss = step(1)

Drop every loss tensor of earlier *eager* steps before constructing a :class:`GraphedStep` (``del loss``): a live loss
keeps its autograd graph and with it the parameters' gradient-accumulator nodes, which stay bound to the stream they
were created on (usually the legacy default stream) -- CUDA then refuses the capture with "operation would make the legacy
stream depend on a capturing blocking stream".  The constructor collects garbage first for the same reason.

Outputs (the loss, ``param.grad``) are static tensors owned by the graph's memory pool: read or consume them before the
same set is replayed again.  All sets share one pool, so the graphs must be replayed on one stream, one at a time.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

from . import _lib


def _detach(out):
    if torch.is_tensor(out):
        return out.detach()
    if isinstance(out, (tuple, list)):
        return type(out)(_detach(o) for o in out)
    return out


class GraphedStep:
    """``fn(*inputs) -> tensor (or tuple of tensors)`` captured once per input set and replayed with ``step(i)``."""

    def __init__(self, fn: Callable, input_sets: Sequence[Sequence[torch.Tensor]], warmup: int = 3,
                 params: Optional[Sequence[torch.nn.Parameter]] = None,
                 before_capture: Optional[Callable[[], None]] = None):
        if not input_sets:
            raise ValueError("GraphedStep needs at least one input set")
        for s in input_sets:
            for t in s:
                if not (torch.is_tensor(t) and t.is_cuda):
                    raise RuntimeError("GraphedStep inputs must be CUDA tensors (mmgclip_b200 has no CPU fallback)")
        self.fn = fn
        self.input_sets: List[tuple] = [tuple(s) for s in input_sets]
        self.graphs: List[torch.cuda.CUDAGraph] = []
        self.outputs: List = []
        self.replays = 0
        # every recorded graph writes the gradients into its own static tensors; `params` lets a replay re-point
        # `param.grad` at the tensors of the graph that just ran (an optimizer then sees the right ones)
        self.params = list(params) if params is not None else []
        self.grads: List[list] = []
        lib = _lib.load()
        import gc
        gc.collect()  # dead autograd graphs of earlier eager steps (see the module docstring)
        # eager warm-up on a side stream: lazy one-time work (cudaFuncSetAttribute, workspace growth, NCCL communicator
        # set-up, autograd buffers) must happen before anything is recorded
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(max(warmup, 1)):
                out = fn(*self.input_sets[i % len(self.input_sets)])
            del out
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        pool = None
        self.launches_per_replay: List[int] = []
        for s in self.input_sets:
            if before_capture is not None:
                before_capture()
            g = torch.cuda.CUDAGraph()
            n0 = lib.mmg_kernel_launch_count()
            with torch.cuda.graph(g, pool=pool):
                out = fn(*s)
            self.launches_per_replay.append(int(lib.mmg_kernel_launch_count() - n0))
            if pool is None:
                pool = g.pool()
            self.graphs.append(g)
            # detached views of the static outputs: the recorded step's autograd graph is not needed after capture
            self.outputs.append(_detach(out))
            self.grads.append([p.grad for p in self.params])
            del out
        torch.cuda.synchronize()

    def __len__(self) -> int:
        return len(self.graphs)

    def copy_in(self, i: int, *tensors: torch.Tensor) -> None:
        """Asynchronously overwrite input set ``i`` (device or pinned-host sources) on the current stream."""
        for dst, src in zip(self.input_sets[i], tensors):
            dst.copy_(src, non_blocking=True)

    def __call__(self, i: int = 0):
        self.graphs[i].replay()
        self.replays += 1
        for p, g in zip(self.params, self.grads[i]):
            p.grad = g
        return self.outputs[i]

    @property
    def kernel_launches(self) -> int:
        """Kernels of libmmgclip_b200.so per replayed step (counted while recording; replays issue no host launches)."""
        return self.launches_per_replay[0]
