"""Row-sharded symmetric InfoNCE across the GPUs of one NVSwitch domain (the reference has no
multi-GPU code at all).  One process per GPU, ``torch.distributed`` (NCCL over NVLink 5) for the exchange steps:

    rank r owns rows [r*Bl, (r+1)*Bl) of both modalities
    1. all-gather the column-side embeddings T (bf16 on the tensor-core path)          -> T_all [B, D]
    2. fused forward on local rows x all columns: rowsum (complete), colsum (partial)  -> all-reduce colsum [B]
    3. loss partial over own rows / own columns                                        -> all-reduce scalar
    4. backward: dI_local complete, dT partial for ALL columns                         -> reduce-scatter to owners
    5. d logit_scale partial                                                           -> all-reduce scalar

Only the column side is gathered: with the fixed softmax shift m = s (|cos| <= 1) partial column sums add directly,
so no running-max exchange is needed.  Head-weight gradients are summed with :func:`allreduce_gradients`.
The collectives are abstracted behind ``torch.distributed`` so the same code runs under ``gloo`` on CPU tensors in
the host-logic tests (tests/test_dist_gloo.py) with the kernels replaced by the oracle -- the product path always
uses the CUDA kernels.
"""
from __future__ import annotations

import os
import warnings
from typing import Callable, Optional

import torch
import torch.distributed as dist

from . import ops

# ---------------------------------------------------------------------------------------------------------------------
# Column-side gradients over NVLink peer memory: every rank owns one [B/R, D] fp32 buffer in torch symmetric memory; the
# fused backward's gradient slices TMA-reduce-add straight into the owner's buffer (mmg_infonce_bwd_owners), so the
# gradient GEMM *is* the reduce-scatter.  Results are identical to the staging-buffer + NCCL reduce-scatter path
# (tests/gpu_dist_check.py, eager and graph).  What makes it pay is that the kernel takes whole-K dB slices when the
# owners are remote, so every gradient element crosses NVLink once per row block: 8 x B200 0.775 ms per step against
# 0.817 ms with the asynchronous NCCL reduce-scatter (and 0.945 ms with the 32-block slices of the single-GPU schedule,
# whose remote 128-byte reduce-adds lengthen the slice epilogues).  MMGCLIP_B200_PEER_REDUCE=0, a non-NCCL backend or a
# failed set-up (a collective decision) select the NCCL path.
# ---------------------------------------------------------------------------------------------------------------------
_peer_cache = {}


class _PeerBuffers:
    """A ring of symmetric [rows, D] fp32 buffers (one per in-flight column-side gradient) with their peer mappings.

    The backward hands the ring slot itself to autograd as dB (no copy): a slot is written again four sharded backward
    calls of the same shape later, long after the gradient has been consumed by the column side's producer.  Callers that
    may keep the tensor (a leaf or ``retain_grad()`` column side) get a private copy instead."""

    RING = 4

    def __init__(self, rows, D, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        grp = group if group is not None else dist.group.WORLD
        self.slots = []
        for _ in range(self.RING):
            own = symm_mem.empty((rows, D), dtype=torch.float32, device=device)
            handle = symm_mem.rendezvous(own, group=grp)
            self.slots.append((own, handle, [int(p) for p in handle.buffer_ptrs]))
        self.i = -1
        self.own, self.handle, self.ptrs = self.slots[0]

    def next(self):
        self.i = (self.i + 1) % self.RING
        self.own, self.handle, self.ptrs = self.slots[self.i]
        return self

    def pre_sync(self):
        self.handle.barrier(channel=0)

    def post_sync(self):
        self.handle.barrier(channel=1)


# ---------------------------------------------------------------------------------------------------------------------
# Push all-gather of the column-side embeddings: every rank owns a ring of [B, D] bf16 buffers in symmetric memory; a rank
# copies its shard straight into the same rows of every rank's buffer (mmg_push_rows: one launch, a few CTAs per
# destination, no protocol) and a cross-rank barrier follows.  Both run on a communication stream, so the gather overlaps
# the other head's projection exactly like the asynchronous NCCL all-gather it replaces.  Measured (global batch 32768 on
# 8 GPUs / 8192 on 2, CUDA-graph step, profiles/r02d_*): 2 x B200 0.306 ms per step against 0.320 ms with NCCL;
# 8 x B200 0.712 ms against 0.678 ms with NCCL (seven remote destinations from 32 CTAs do not reach the line rate NCCL's
# tuned all-gather does).  Default therefore: push on two ranks, NCCL above; MMGCLIP_B200_PUSH_GATHER=1 / 0 forces one.
# A ring slot is reused only after the loss that read it has issued its backward (or, without gradients, its forward):
# the collectives inside those (column-sum all-reduce, peer reduction barriers / reduce-scatter) order every peer's last
# read of the slot before this rank's next push into it.  A busy ring, a non-bf16 operand or a non-NCCL backend select
# the NCCL all-gather.
# ---------------------------------------------------------------------------------------------------------------------
_gather_rings = {}
_comm_streams = {}


class _GatherSlot:
    def __init__(self, buf, handle):
        self.buf, self.handle = buf, handle
        self.ptrs = [int(p) for p in handle.buffer_ptrs]
        self.busy = False
        self.gen = 0  # bumped by every acquire: a stale release (a finalizer of an old loss) must not free a newer gather

    def release(self, gen):
        if gen == self.gen:
            self.busy = False


class _GatherRing:
    RING = 4

    def __init__(self, B, D, device, group, dtype):
        import torch.distributed._symmetric_memory as symm_mem
        grp = group if group is not None else dist.group.WORLD
        self.slots = []
        for _ in range(self.RING):
            buf = symm_mem.empty((B, D), dtype=dtype, device=device)
            self.slots.append(_GatherSlot(buf, symm_mem.rendezvous(buf, group=grp)))
        self.i = -1

    def acquire(self):
        """The next slot in ring order, or None while it is still being read (the same decision on every rank: the ranks
        run the same program)."""
        nxt = (self.i + 1) % self.RING
        slot = self.slots[nxt]
        if slot.busy:
            return None
        self.i = nxt
        slot.busy = True
        slot.gen += 1
        return slot


def _gather_ring(B, D, device, group, dtype):
    want = os.environ.get("MMGCLIP_B200_PUSH_GATHER", "auto")
    world = dist.get_world_size(group)
    if want == "0" or device.type != "cuda" or (want != "1" and world > 2):
        return None
    if dist.get_backend(group) != "nccl" or world > 8 or (D * 2) % 16 != 0:
        return None
    key = (B, D, device.index, id(group), dtype)
    if key not in _gather_rings:
        try:
            _gather_rings[key] = _GatherRing(B, D, device, group, dtype)
        except Exception as e:  # noqa: BLE001
            warnings.warn(f"mmgclip_b200: push all-gather unavailable ({e}); using the NCCL all-gather")
            _gather_rings[key] = None
    return _gather_rings[key]


def push_gather_active() -> bool:
    return any(v is not None for v in _gather_rings.values())


def _comm_stream(device):
    st = _comm_streams.get(device.index)
    if st is None:
        st = _comm_streams[device.index] = torch.cuda.Stream(device=device)
    return st


# stored-E (ops.want_store_e) in the sharded loss: opt-in until measured on several GPUs
_STORE_E_DIST = os.environ.get("MMGCLIP_B200_STORE_E_DIST", "0") == "1"


def peer_reduce_active() -> bool:
    """True once a sharded backward has run with the NVLink peer reduction (i.e. not on the NCCL fallback)."""
    return any(v is not None for v in _peer_cache.values())


def _peer_buffers(rows, D, device, group):
    if os.environ.get("MMGCLIP_B200_PEER_REDUCE", "1") == "0":
        return None
    world = dist.get_world_size(group)
    if dist.get_backend(group) != "nccl" or world > 8 or rows % 256 != 0 or D % 256 != 0 or D < 256:
        return None
    key = (rows, D, device.index, id(group))
    if key not in _peer_cache:
        try:
            _peer_cache[key] = _PeerBuffers(rows, D, device, group)
        except Exception as e:  # noqa: BLE001  (no symmetric-memory support on this platform / torch build)
            warnings.warn(f"mmgclip_b200: NVLink peer reduction unavailable ({e}); using NCCL reduce-scatter")
            _peer_cache[key] = None
    return _peer_cache[key]


# ---------------------------------------------------------------------------------------------------------------------
# Small all-reduces (column sums [B], the loss scalar, the flat head gradients) through torch symmetric memory: one-shot /
# two-shot kernels that read the peers' buffers over NVLink cost a few microseconds, where an NCCL all-reduce inside a
# replayed graph costs tens (launch + 8-way rendezvous) -- and an 8-GPU step has three of them on its critical path.
# MMGCLIP_B200_SYMM_ALLREDUCE=0, a non-NCCL backend or a failed set-up (a collective decision) select dist.all_reduce.
# ---------------------------------------------------------------------------------------------------------------------
_symm_vectors = {}


class _SymmVector:
    """A persistent fp32 vector in symmetric memory with an in-place-style sum across the group."""

    def __init__(self, numel, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.numel = (numel + 1023) // 1024 * 1024
        self.buf = symm_mem.empty(self.numel, dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.buf, group=self.group)
        self.name = self.group.group_name

    def all_reduce(self) -> torch.Tensor:
        """Sum of every rank's buffer (a new tensor for small vectors, the buffer itself for large ones)."""
        if self.numel * 4 <= (1 << 20):
            return torch.ops.symm_mem.one_shot_all_reduce(self.buf, "sum", self.name)
        return torch.ops.symm_mem.two_shot_all_reduce_(self.buf, "sum", self.name)


def _symm_vector(tag, numel, device, group):
    if os.environ.get("MMGCLIP_B200_SYMM_ALLREDUCE", "1") == "0" or device.type != "cuda":
        return None
    if dist.get_backend(group) != "nccl" or dist.get_world_size(group) > 8:
        return None
    key = (tag, numel, device.index, id(group))
    if key not in _symm_vectors:
        try:
            _symm_vectors[key] = _SymmVector(numel, device, group)
        except Exception as e:  # noqa: BLE001
            warnings.warn(f"mmgclip_b200: symmetric-memory all-reduce unavailable ({e}); using NCCL")
            _symm_vectors[key] = None
    return _symm_vectors[key]


def symm_allreduce_active() -> bool:
    return any(v is not None for v in _symm_vectors.values())


def _reduce_scatter_sum(out: torch.Tensor, full: torch.Tensor, group) -> None:
    """out = this rank's slice of sum_over_ranks(full).  NCCL: one reduce-scatter; gloo (CPU host-logic tests) has no
    reduce-scatter, so all-reduce and slice."""
    if dist.get_backend(group) == "gloo":
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
        n = out.shape[0]
        r = dist.get_rank(group)
        out.copy_(full[r * n:(r + 1) * n])
    else:
        dist.reduce_scatter_tensor(out, full, op=dist.ReduceOp.SUM, group=group)


class _Kernels:
    """The four local compute steps the sharded loss is made of (swappable for the gloo host-logic tests)."""

    @staticmethod
    def operand(t, prec):
        return ops._operand(t, prec)

    @staticmethod
    def forward(a_op, b_all_op, scale, diag_offset, prec, colsum=None, e_out=None):
        return ops.infonce_forward_raw(a_op, b_all_op, scale, diag_offset, prec, colsum=colsum, e_out=e_out)

    @staticmethod
    def row_part(rowsum, diag, out=None):
        return ops.infonce_row_part_raw(rowsum, diag, out=out)

    @staticmethod
    def loss_cols(colsum, scale, row_part, inv_two_b):
        return ops.infonce_loss_cols_raw(colsum, scale, row_part, inv_two_b)

    @staticmethod
    def backward(a_op, b_all_op, scale, rowsum, colsum, grad_loss, inv_two_b, diag_offset, prec, a32, b32_paired,
                 diag, need_dscale=True):
        return ops.infonce_backward_raw(a_op, b_all_op, scale, rowsum, colsum, grad_loss, inv_two_b, diag_offset, prec,
                                        a32=a32, b32=b32_paired, diag=diag, need_dscale=need_dscale)


class GatheredColumns:
    """Column-side embeddings being all-gathered in the background (see :func:`gather_columns_async`)."""

    def __init__(self, local, b_all, work=None, event=None, slot=None):
        self.local, self.b_all, self.work, self.event, self.slot = local, b_all, work, event, slot
        self.gen = slot.gen if slot is not None else 0

    def wait(self):
        """Make the current stream wait for the gathered matrix (no host sync)."""
        if self.work is not None:
            self.work.wait()
            self.work = None
        if self.event is not None:
            torch.cuda.current_stream().wait_event(self.event)
            self.event = None
        return self.b_all

    def release(self):
        """The gathered matrix will not be read any more (its ring slot may take the next gather)."""
        if self.slot is not None:
            self.slot.release(self.gen)
            self.slot = None


def gather_columns_async(b_local: torch.Tensor, group=None, prec: Optional[str] = None, _kernels=None):
    """Start the all-gather of the column-side embeddings now and let it run while the caller projects the other
    modality; pass the result to :func:`sharded_info_nce` as ``gathered=``."""
    kernels = _kernels or _Kernels
    prec = ops._resolve(prec)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    world = dist.get_world_size(group)
    b_op = kernels.operand(b_local, prec).contiguous()
    bl, D = b_op.shape
    if kernels is _Kernels and b_op.is_cuda and b_op.dtype in (torch.bfloat16, torch.float16):
        ring = _gather_ring(bl * world, D, b_op.device, group, b_op.dtype)
        slot = ring.acquire() if ring is not None else None
        if slot is not None:
            comm = _comm_stream(b_op.device)
            comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(comm):
                ops.push_rows(b_op, slot.ptrs, dist.get_rank(group) * bl * D * 2)
                slot.handle.barrier(channel=0)
                ev = torch.cuda.Event()
                ev.record(comm)
            b_op.record_stream(comm)
            return GatheredColumns(b_local, slot.buf, event=ev, slot=slot)
    b_all = torch.empty((bl * world, D), dtype=b_op.dtype, device=b_op.device)
    work = dist.all_gather_into_tensor(b_all, b_op, group=group, async_op=True)
    return GatheredColumns(b_local, b_all, work=work)


class _ShardedInfoNCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a_local, b_local, scale, group, prec, kernels, gathered=None, pending=None, b_key=None):
        ctx.pending = pending
        key = b_key if b_key is not None else b_local
        ctx.private_db = bool(key.is_leaf or key.retains_grad)
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        bl, D = a_local.shape
        if b_local.shape != a_local.shape:
            raise ValueError("each rank must hold the same number of rows of both modalities")
        B = bl * world
        s = scale.detach().reshape(()).to(device=a_local.device, dtype=torch.float32).contiguous()
        a_op = kernels.operand(a_local, prec)
        if gathered is None or not (gathered.local is b_local or gathered.local is b_key):
            gathered = gather_columns_async(b_local if b_key is None else b_key, group=group, prec=prec, _kernels=kernels)
        b_all = gathered.wait()
        ops.mark("gather_wait")
        ctx.gathered = gathered
        off = rank * bl
        # stored-E (ops.want_store_e) for the sharded loss: opt-in (MMGCLIP_B200_STORE_E_DIST=1) until it has been measured
        # on several GPUs; only with the peer-memory backward, which is the one that reaches mmg_infonce_bwd_stored
        e_mat = None
        if (kernels is _Kernels and _STORE_E_DIST and a_local.is_cuda and peer_reduce_active() and
                (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]) and
                a_op.dtype == torch.bfloat16 and ops.want_store_e(bl, B, D, prec, ctx.needs_input_grad[2], n_owners=world)):
            e_mat = torch.empty((bl, B), dtype=torch.bfloat16, device=a_local.device)
        ctx.e_mat = e_mat
        fwd_kw = {"e_out": e_mat} if e_mat is not None else {}
        # ONE cross-rank sum per forward: the partial column sums [B] and, in element B of the same vector, this rank's part
        # of the loss that is known before the exchange (sum over its rows of log rowsum - 2 diag).  Afterwards every rank
        # holds the global column sums and finishes the loss locally (sum over ALL columns of log colsum) -- no second
        # all-reduce for the scalar, and the value is bit-identical on every rank.
        cvec = _symm_vector("colsum", B + 1, a_local.device, group) if kernels is _Kernels else None
        if cvec is not None:
            cvec.buf.zero_()  # the partial column sums accumulate straight into the symmetric buffer
            rowsum, _, diag = kernels.forward(a_op, b_all, s, off, prec, colsum=cvec.buf[:B], **fwd_kw)
            kernels.row_part(rowsum, diag, out=cvec.buf[B:B + 1])
            ops.mark("lse_fwd")
            red = cvec.all_reduce()
            if red.data_ptr() == cvec.buf.data_ptr():
                red = red.clone()  # very large B: the two-shot kernel reduces in place, and the next forward clears the buffer
            colsum, part = red[:B], red[B:B + 1]
            ops.mark("colsum_ar")
        else:
            rowsum, colsum, diag = kernels.forward(a_op, b_all, s, off, prec, **fwd_kw)
            vec = torch.cat([colsum, kernels.row_part(rowsum, diag).to(colsum.dtype).reshape(1)])
            dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
            colsum, part = vec[:B], vec[B:B + 1]
        loss = kernels.loss_cols(colsum, s, part, 0.5 / B)
        ctx.group, ctx.prec, ctx.kernels, ctx.off, ctx.B = group, prec, kernels, off, B
        ctx.scale_shape = scale.shape
        if not any(ctx.needs_input_grad[:3]):
            gathered.release()  # forward only: the column-sum all-reduce above ordered every rank's read of the slot
        elif gathered.slot is not None:
            # a loss that never runs its backward must not pin the ring: free the slot when the autograd node dies
            try:
                import weakref
                weakref.finalize(ctx, gathered.slot.release, gathered.gen)
            except TypeError:
                pass  # not weak-referenceable on this torch: the ring then falls back to NCCL once it is full
        ctx.save_for_backward(a_op, b_all, s, rowsum, colsum, a_local.detach(), b_local.detach(), diag)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        a_op, b_all, s, rowsum, colsum, a32, b32_local, diag = ctx.saved_tensors
        group, prec, kernels = ctx.group, ctx.prec, ctx.kernels
        gathered, ctx.gathered = ctx.gathered, None
        bl, D = a_op.shape
        # fp32 embeddings for the matching-pair term: the columns paired with this rank's rows are its own b rows
        if not (prec == "bf16" and a32.dtype == torch.float32):
            a32 = b32_local = None
        peer = None
        if kernels is _Kernels and prec == "bf16" and a_op.is_cuda:
            peer = _peer_buffers(bl, D, a_op.device, group)
        if peer is not None:
            peer.next()
            # gradient GEMM + reduce-scatter in one kernel: the slices add into their owners' buffers over NVLink
            dA, owns, dls = ops.infonce_backward_owners(a_op, b_all, s, rowsum, colsum, grad_loss, 0.5 / ctx.B, ctx.off,
                                                        [(peer.own, peer.ptrs)], peer.pre_sync, peer.post_sync, a32=a32,
                                                        b32=b32_local, diag=diag, need_dscale=ctx.needs_input_grad[2],
                                                        e_stored=ctx.e_mat)
            ctx.e_mat = None
            dB = owns[0].clone() if ctx.private_db else owns[0]  # a ring slot (see _PeerBuffers) unless it may be kept
            ops.mark("clone")
            if gathered is not None:
                gathered.release()  # the peer reduction's closing barrier ordered every rank's last read of the slot
            dscale = None
            if ctx.needs_input_grad[2]:
                dist.all_reduce(dls, op=dist.ReduceOp.SUM, group=group)
                dscale = (dls / s).reshape(ctx.scale_shape)
            return dA, dB, dscale, None, None, None, None, None, None
        dA, dB_all, dls = kernels.backward(a_op, b_all, s, rowsum, colsum, grad_loss, 0.5 / ctx.B, ctx.off, prec, a32,
                                           b32_local, diag, need_dscale=ctx.needs_input_grad[2])
        dB = torch.empty((bl, D), dtype=dB_all.dtype, device=dB_all.device)
        pending = ctx.pending
        if pending is not None and pending.get("deferred") and dist.get_backend(group) != "gloo":
            # The column-side gradients travel while the caller's row-side (image head) backward runs: the current
            # stream only waits for the reduce-scatter in the backward of the identity node sharded_info_nce() put in
            # front of b_local (_AwaitColumnGrad), i.e. before the gradient meets any other consumer's.
            pending["work"] = dist.reduce_scatter_tensor(dB, dB_all, op=dist.ReduceOp.SUM, group=group, async_op=True)
            pending["keep"] = dB_all  # stays referenced until the wait
        else:
            _reduce_scatter_sum(dB, dB_all, group)
        if gathered is not None:
            gathered.release()  # the reduce-scatter issued above orders every rank's last read of the slot
        dscale = None
        if ctx.needs_input_grad[2]:
            dist.all_reduce(dls, op=dist.ReduceOp.SUM, group=group)
            dscale = (dls / s).reshape(ctx.scale_shape)
        return dA, dB, dscale, None, None, None, None, None, None


def sharded_info_nce(a_local: torch.Tensor, b_local: torch.Tensor, logit_scale, group=None, prec: Optional[str] = None,
                     _kernels=_Kernels, gathered: Optional[GatheredColumns] = None) -> torch.Tensor:
    """Global symmetric InfoNCE for a batch sharded by rows; every rank gets the same (global mean) loss and, on
    backward, the exact gradient of that global loss with respect to ITS rows (no 1/world rescaling is needed)."""
    prec = ops._resolve(prec)
    if not torch.is_tensor(logit_scale):
        logit_scale = torch.tensor(float(logit_scale), dtype=torch.float32, device=a_local.device)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return ops.info_nce(a_local, b_local, logit_scale, prec=prec)
    pending = {}
    b_in = b_local
    if torch.is_grad_enabled() and b_local.requires_grad:
        # The NCCL-fallback backward returns dB while its asynchronous reduce-scatter is still in flight.  b_local is
        # routed through a private identity node so that this gradient has exactly ONE producer and the wait happens in
        # that node's backward -- i.e. before autograd can add it to the gradients b_local receives from other consumers
        # (MMGCLIPLoss uses the text embeddings as the column side of two losses).
        b_in = _AwaitColumnGrad.apply(b_local, pending)
        bf = getattr(b_local, "_mmg_bf16", None)
        if bf is not None:
            b_in._mmg_bf16 = bf  # keep the bf16 operand copy the normalise kernel attached
        pending["deferred"] = True
    return _ShardedInfoNCEFn.apply(a_local, b_in, logit_scale, group, prec, _kernels, gathered, pending, b_local)


class _AwaitColumnGrad(torch.autograd.Function):
    """Identity on the column-side embeddings; its backward makes the current stream wait for the asynchronous
    reduce-scatter that fills the incoming gradient (stream-ordered, no host sync)."""

    @staticmethod
    def forward(ctx, b_local, pending):
        ctx.pending = pending
        return b_local.view_as(b_local)

    @staticmethod
    def backward(ctx, grad):
        work = ctx.pending.pop("work", None)
        if work is not None:
            work.wait()
        ctx.pending.pop("keep", None)
        return grad, None


def allreduce_gradients(*modules: torch.nn.Module, group=None) -> None:
    """Sum parameter gradients across ranks (one flat all-reduce for all given modules): with the global loss above each
    rank's head gradient covers only its own rows.

    On the symmetric-memory path the summed gradients are handed back as VIEWS of the persistent flat buffer (no copy
    out): ``p.grad`` is valid until the next call with the same parameters overwrites it -- what an optimizer step or a
    graph replay needs.  A gradient that is already such a view (gradients accumulated in place over micro-steps) is
    summed where it lies."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    params = [p for m in modules for p in m.parameters() if p.grad is not None]
    grads = [p.grad for p in params]
    if not grads:
        return
    n = sum(g.numel() for g in grads)
    vec = _symm_vector("grads", n, grads[0].device, group) if all(g.dtype == torch.float32 for g in grads) else None
    if vec is not None:
        o = 0
        in_place = True
        for g in grads:
            in_place = in_place and g.is_contiguous() and g.data_ptr() == vec.buf.data_ptr() + 4 * o
            o += g.numel()
        if not in_place:
            torch.cat([g.reshape(-1) for g in grads], out=vec.buf[:n])
        flat = vec.all_reduce()
        o = 0
        if flat.data_ptr() == vec.buf.data_ptr():
            for p, g in zip(params, grads):
                p.grad = flat[o:o + g.numel()].view_as(g)
                o += g.numel()
            return
    else:
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()
