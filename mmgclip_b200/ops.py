"""Host-side operators of the hot path: torch tensors in, C-ABI calls on the current CUDA stream, torch tensors out.

PyTorch is used for what it is good at here -- device memory, streams and the autograd graph that ties the operators
to ``loss.backward()`` in the reference's training loop (mmgclip/experiments/ClassifierExperiment.py:109-118).  All
arithmetic happens in ``libmmgclip_b200.so``; nothing in this file computes on the CPU or through ATen kernels.

Arithmetic follows the reference lines cited in ``include/mmgclip_b200.h``; closed-form gradients are SURVEY.md s3.5.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import MMG_ACCUMULATE, MMG_ATOMIC_ADD, MMG_PREC_BF16, MMG_PREC_F16, MMG_PREC_FP32, MMG_STORE, check

_PREC = {"fp32": MMG_PREC_FP32, "bf16": MMG_PREC_BF16}
_default_precision = os.environ.get("MMGCLIP_B200_PRECISION", "bf16")
_NUM_SMS = None


def set_default_precision(p: str) -> None:
    """'bf16' (tcgen05 tensor cores, 2e-3 parity) or 'fp32' (FFMA, 1e-5 parity)."""
    global _default_precision
    if p not in _PREC:
        raise ValueError(f"Invalid precision: {p}")
    _default_precision = p


def get_default_precision() -> str:
    return _default_precision


def _resolve(prec: Optional[str]) -> str:
    p = prec or _default_precision
    if p not in _PREC:
        raise ValueError(f"Invalid precision: {p}")
    return p


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("mmgclip_b200 has no CPU fallback: expected CUDA tensors, got a tensor on " + str(t.device))


def num_sms() -> int:
    global _NUM_SMS
    if _NUM_SMS is None:
        _NUM_SMS = _lib.device_info()[0]
    return _NUM_SMS


_workspaces = {}

# Optional timing probe around the dominant launch (the InfoNCE backward contraction): a pair of CUDA events recorded on
# the launching stream right before / after it.  With `torch.cuda.Event(enable_timing=True, external=True)` the records are
# captured into CUDA graphs as well, so a replayed step can be timed per kernel without a profiler (bench.py's live
# roofline).
_bwd_probe = None


def set_backward_probe(events) -> None:
    """events: (start, end) CUDA events or None."""
    global _bwd_probe
    _bwd_probe = events


# Optional phase timeline of a step (bench.py --timeline): named external CUDA events recorded on the current stream at
# phase boundaries; like the probe above they are captured into CUDA graphs, so after a replay they hold that replay's
# timestamps.  None (the default) = no events are recorded anywhere.
_timeline = None


def set_timeline(marks) -> None:
    """marks: dict name -> event (filled on first use) or None."""
    global _timeline
    _timeline = marks


def mark(name: str) -> None:
    if _timeline is None:
        return
    ev = _timeline.get(name)
    if ev is None:
        ev = _timeline[name] = torch.cuda.Event(enable_timing=True, external=True)
    ev.record()


_TUNE_KEYS = ("fused", "fused_rb", "fused_cb", "fused_nbuf", "fused_ksl", "fused_ksl_t", "fused_sr", "fused_sc", "fused_rot", "pdl")


def set_tuning(**kw) -> None:
    """Plan knobs of the fused InfoNCE backward (mmg_tune): ``fused`` (0 = block loop), ``fused_rb`` / ``fused_cb`` (block
    shape), ``fused_nbuf`` (scratch buffers), ``fused_ksl`` / ``fused_ksl_t`` (K blocks per gradient slice), ``fused_sr`` /
    ``fused_sc`` (block-order super-tile).  Every setting computes the same result; the defaults are the measured best.
    ``set_tuning()`` without arguments restores them; a value of None unsets one knob."""
    lib = _lib.load()
    if not kw:
        check(lib.mmg_tune(b"reset", 0), "mmg_tune")
        return
    for k, v in kw.items():
        if k not in _TUNE_KEYS:
            raise ValueError(f"Invalid tuning key: {k}")
        check(lib.mmg_tune(k.encode(), -1 if v is None else int(v)), "mmg_tune")


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Grow-only scratch per device (the L2-resident gradient-coefficient block lives here)."""
    key = (device.type, device.index)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


# ----------------------------------------------------------------------------------------------------------------
# raw wrappers
# ----------------------------------------------------------------------------------------------------------------
def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    x = x.contiguous()
    if x.dtype == torch.bfloat16:
        return x
    if x.dtype != torch.float32:
        raise ValueError("cast_bf16 expects float32 input")
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(_lib.load().mmg_cast_f32_to_bf16(_p(x), _p(y), x.numel(), _stream()), "mmg_cast_f32_to_bf16")
    return y


def cast_f16(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> fp16 (operand copy of L2-normalised embeddings)."""
    _need_cuda(x)
    x = x.contiguous()
    if x.dtype == torch.float16:
        return x
    if x.dtype != torch.float32:
        raise ValueError("cast_f16 expects float32 input")
    y = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    check(_lib.load().mmg_cast_f32_to_f16(_p(x), _p(y), x.numel(), _stream()), "mmg_cast_f32_to_f16")
    return y


# 16-bit operand format of the L2-NORMALISED EMBEDDINGS on the tensor-core ("bf16") path of the fused InfoNCE.  Unit-norm
# rows have |x| <= 1, so IEEE fp16 (11-bit significand) represents them 8x more finely than bf16 (8-bit) at no cost in range
# or tensor-core rate; the gradient-coefficient operand stays bf16 (it spans e^-2s .. 1 times 1/B).  Measured on the
# benchmark step (float64 closed form, batch 4096): head-weight gradient error 2.1e-3 -> 2.7e-4 (Frobenius), column-side
# embedding gradients 2.9e-3 -> 2.9e-4 -- the embeddings' bf16 rounding was the whole error budget.
# MMGCLIP_B200_EMB_F16=0 (or set_embedding_f16(False)) keeps bf16 embedding operands.
_emb_f16 = os.environ.get("MMGCLIP_B200_EMB_F16", "0") == "1"


def set_embedding_f16(flag: bool) -> None:
    global _emb_f16
    _emb_f16 = bool(flag)


def get_embedding_f16() -> bool:
    return _emb_f16


def cast_embedding(x: torch.Tensor) -> torch.Tensor:
    """The 16-bit tensor-core operand of an fp32 embedding matrix (fp16 by default, see above)."""
    return cast_f16(x) if _emb_f16 else cast_bf16(x)


def _tc_prec(prec: str, a: torch.Tensor, b: torch.Tensor) -> int:
    """C-ABI precision code of a fused-InfoNCE call from the operands the caller holds."""
    if prec == "fp32":
        return MMG_PREC_FP32
    if a.dtype != b.dtype or a.dtype not in (torch.bfloat16, torch.float16):
        raise ValueError(f"tensor-core InfoNCE operands must both be bfloat16 or both float16, got {a.dtype} / {b.dtype}")
    return MMG_PREC_F16 if a.dtype == torch.float16 else MMG_PREC_BF16


def cast_bf16_split(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(hi, lo) bf16 pair with hi + lo ~= x to 16 mantissa bits (operands of the bf16x3 head contractions)."""
    _need_cuda(x)
    x = x.contiguous()
    if x.dtype != torch.float32:
        raise ValueError("cast_bf16_split expects float32 input")
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(_lib.load().mmg_cast_f32_to_bf16_split(_p(x), _p(hi), _p(lo), x.numel(), _stream()),
          "mmg_cast_f32_to_bf16_split")
    return hi, lo


# Projection heads on the bf16 tensor pipe use three passes (hi.hi + hi.lo + lo.hi) by default: they are a few percent
# of the step's FLOPs at large batch, and it keeps embeddings and head-weight gradients fp32-faithful (~1e-5) instead of
# carrying the 2^-9 bf16 operand rounding into every downstream quantity.  MMGCLIP_B200_SPLIT_HEADS=0 = single pass.
_split_heads = os.environ.get("MMGCLIP_B200_SPLIT_HEADS", "1") != "0"


def set_split_heads(flag: bool) -> None:
    global _split_heads
    _split_heads = bool(flag)


# The head-weight gradient dW = du^T x can drop the lo terms independently of the forward: du already carries the bf16
# rounding of the InfoNCE coefficients, so hi.hi alone is inside the bf16 gradient bar (tests/gpu_dw_probe.py measures
# both).  MMGCLIP_B200_SPLIT_DW: "1" = three K segments, "0" = one.
_split_dw = os.environ.get("MMGCLIP_B200_SPLIT_DW", "1") != "0"


def set_split_dw(flag: bool) -> None:
    global _split_dw
    _split_dw = bool(flag)


def _dw_operands(dz: "_Operand", xo: "_Operand"):
    if _split_dw:
        return dz, xo
    return _Operand(dz.hi), _Operand(xo.hi)


class _Operand:
    """A contraction operand for a given precision: fp32 tensor, bf16 tensor, or a (hi, lo) bf16 pair."""

    __slots__ = ("hi", "lo")

    def __init__(self, hi, lo=None):
        self.hi, self.lo = hi, lo

    @staticmethod
    def of(x: torch.Tensor, prec: str) -> "_Operand":
        if prec == "fp32":
            return _Operand(x.contiguous())
        if _split_heads:
            return _Operand(*cast_bf16_split(x))
        return _Operand(cast_bf16(x))

    def tensors(self):
        return (self.hi,) if self.lo is None else (self.hi, self.lo)

    @staticmethod
    def from_tensors(ts):
        return _Operand(*ts)


def gemm_heads(A: "_Operand", B: "_Operand", M: int, N: int, K: int, *, a_mn=False, b_mn=False, bias=None, relu=False,
               k_splits: int = 1, prec: str = "bf16", zeroed_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """alpha=1 contraction of head operands.  Split operands (hi, lo) contract as hi.hi + hi.lo + lo.hi in ONE launch over
    the concatenated K range (mmg_gemm_split); bias / ReLU ride in its epilogue."""
    if A.lo is None and B.lo is None:
        if k_splits > 1:
            if bias is not None or relu:
                raise ValueError("bias/ReLU cannot be combined with split-K")
            out = zeroed_out if zeroed_out is not None else torch.zeros((M, N), dtype=torch.float32, device=A.hi.device)
            return gemm(A.hi, B.hi, M, N, K, a_mn=a_mn, b_mn=b_mn, out=out, mode=MMG_ATOMIC_ADD, k_splits=k_splits,
                        prec=prec)
        return gemm(A.hi, B.hi, M, N, K, a_mn=a_mn, b_mn=b_mn, bias=bias, relu=relu, prec=prec)
    if prec != "bf16":
        raise ValueError("split operands exist only on the bf16 path")
    _need_cuda(A.hi, A.lo, B.hi, B.lo, bias)
    for t in A.tensors() + B.tensors():
        if t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1:
            raise ValueError("gemm_heads expects 2-D bf16 operands with a contiguous last dimension")
    exp_a = (K, M) if a_mn else (M, K)
    exp_b = (K, N) if b_mn else (N, K)
    if tuple(A.hi.shape) != exp_a or tuple(B.hi.shape) != exp_b:
        raise ValueError(f"gemm shape mismatch: A {tuple(A.hi.shape)} vs {exp_a}, B {tuple(B.hi.shape)} vs {exp_b}")
    if k_splits > 1:
        if bias is not None or relu:
            raise ValueError("bias/ReLU cannot be combined with split-K")
        out = zeroed_out if zeroed_out is not None else torch.zeros((M, N), dtype=torch.float32, device=A.hi.device)
        mode = MMG_ATOMIC_ADD
    else:
        out = torch.empty((M, N), dtype=torch.float32, device=A.hi.device)
        mode = MMG_STORE
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N or not bias.is_contiguous()):
        raise ValueError("gemm bias must be contiguous float32 [N]")
    check(_lib.load().mmg_gemm_split(_p(A.hi), _p(A.lo), A.hi.stride(0), int(a_mn), _p(B.hi), _p(B.lo), B.hi.stride(0),
                                     int(b_mn), _p(out), out.stride(0), M, N, K, 1.0, _p(bias), int(relu), mode,
                                     k_splits, _stream()), "mmg_gemm_split")
    return out


def gemm(A: torch.Tensor, B: torch.Tensor, M: int, N: int, K: int, *, a_mn: bool = False, b_mn: bool = False,
         out: Optional[torch.Tensor] = None, alpha: float = 1.0, alpha_dev: Optional[torch.Tensor] = None,
         bias: Optional[torch.Tensor] = None,
         relu: bool = False, mode: int = MMG_STORE, k_splits: int = 1, prec: Optional[str] = None) -> torch.Tensor:
    """C[M,N] (op)= alpha * A . B^T.  A is [M,K] (or [K,M] if a_mn), B is [N,K] (or [K,N] if b_mn)."""
    prec = _resolve(prec)
    _need_cuda(A, B, out, bias)
    want = torch.bfloat16 if prec == "bf16" else torch.float32
    if A.dtype != want or B.dtype != want:
        raise ValueError(f"gemm({prec}) expects {want} operands, got {A.dtype} / {B.dtype}")
    if A.dim() != 2 or B.dim() != 2 or A.stride(1) != 1 or B.stride(1) != 1:
        raise ValueError("gemm operands must be 2-D with a contiguous last dimension")
    exp_a = (K, M) if a_mn else (M, K)
    exp_b = (K, N) if b_mn else (N, K)
    if tuple(A.shape) != exp_a or tuple(B.shape) != exp_b:
        raise ValueError(f"gemm shape mismatch: A {tuple(A.shape)} vs {exp_a}, B {tuple(B.shape)} vs {exp_b}")
    if out is None:
        if mode != MMG_STORE:
            out = torch.zeros((M, N), dtype=torch.float32, device=A.device)
        else:
            out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    if out.dtype != torch.float32 or tuple(out.shape) != (M, N) or out.stride(1) != 1:
        raise ValueError("gemm output must be float32 [M, N] with contiguous rows")
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N or not bias.is_contiguous()):
        raise ValueError("gemm bias must be contiguous float32 [N]")
    check(_lib.load().mmg_gemm(_PREC[prec], _p(A), A.stride(0), int(a_mn), _p(B), B.stride(0), int(b_mn), _p(out),
                               out.stride(0), M, N, K, float(alpha), _p(alpha_dev), _p(bias), int(relu), mode, k_splits,
                               _stream()),
          "mmg_gemm")
    return out


def _n_segments(A: "_Operand", B: "_Operand") -> int:
    """K segments of a split-precision contraction: hi.hi (+ hi.lo) (+ lo.hi)."""
    return 1 + (B.lo is not None) + (A.lo is not None)


def _split_k_for(M: int, N: int, K: int) -> int:
    """Split-K factor that fills the SMs for a short-and-wide output (the dW contraction has K = batch)."""
    tiles = ((M + 127) // 128) * ((N + 255) // 256)
    nkb = (K + 63) // 64
    return max(1, min(nkb // 4 if nkb >= 8 else 1, num_sms() // max(tiles, 1)))


def l2norm_fwd(u: torch.Tensor, want_bf16: bool) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """y = u/||u||, 1/||u||, and (``want_bf16``) the 16-bit operand copy of y -- fp16 unless set_embedding_f16(False)."""
    _need_cuda(u)
    u = u.contiguous()
    B, D = u.shape
    y = torch.empty_like(u)
    inv = torch.empty(B, dtype=torch.float32, device=u.device)
    f16 = _emb_f16
    yb = torch.empty((B, D), dtype=torch.float16 if f16 else torch.bfloat16, device=u.device) if want_bf16 else None
    if B > 0:
        check(_lib.load().mmg_l2norm_fwd(_p(u), B, D, _p(y), _p(inv), _p(yb), int(f16), _stream()), "mmg_l2norm_fwd")
    return y, inv, yb


def l2norm_bwd(dy: torch.Tensor, y: torch.Tensor, inv: torch.Tensor, want_f32: bool, want_bf16: bool,
               want_lo: bool = False, zero: Optional[torch.Tensor] = None):
    """``zero``: an fp32 buffer (numel % 4 == 0) the same launch clears -- the split-K output of the contraction that follows."""
    _need_cuda(dy, y, inv, zero)
    dy = dy.contiguous()
    B, D = y.shape
    du = torch.empty_like(y) if want_f32 else None
    dub = torch.empty((B, D), dtype=torch.bfloat16, device=y.device) if want_bf16 else None
    dul = torch.empty((B, D), dtype=torch.bfloat16, device=y.device) if (want_bf16 and want_lo) else None
    if B > 0:
        check(_lib.load().mmg_l2norm_bwd(_p(dy), _p(y), _p(inv), B, D, _p(du), _p(dub), _p(dul), _p(zero),
                                         0 if zero is None else zero.numel(), _stream()), "mmg_l2norm_bwd")
    elif zero is not None:
        zero.zero_()
    if want_lo:
        return du, dub, dul
    return du, dub


# Stored-E mode of the bf16 InfoNCE -- OPT-IN, off by default.  The default path never materialises anything of size
# rows x cols in HBM (O(B*D) memory; the backward recomputes the cosines on the tensor cores).  With a budget
# (MMGCLIP_B200_STORE_E_MB > 0, or set_store_e_budget_mb) the forward keeps E = exp(logit - s) as bf16 [rows, cols] and the
# backward transforms it into the gradient coefficients instead of recomputing (4 instead of 6 B^2 D FLOPs in the backward)
# at the price of 2*rows*cols bytes of HBM per live loss.  Only taken when E fits the budget, the shape is covered by the
# fused backward and logit_scale carries no gradient (the cosines are gone, so sum g*cos cannot be formed).
_store_e_mb = int(os.environ.get("MMGCLIP_B200_STORE_E_MB", "0"))


def get_store_e_budget_mb() -> int:
    return _store_e_mb


def set_store_e_budget_mb(mb: int) -> None:
    global _store_e_mb
    _store_e_mb = int(mb)


def want_store_e(rows: int, cols: int, D: int, prec: str, need_dscale: bool, n_owners: int = 1, n_parts: int = 1) -> bool:
    if prec != "bf16" or need_dscale or _store_e_mb <= 0 or rows * cols * 2 > (_store_e_mb << 20):
        return False
    return bool(_lib.load().mmg_infonce_stored_supported(rows, cols, D, n_owners, n_parts))


def infonce_forward_raw(a, b, scale, diag_offset: int, prec: str, rowsum=None, colsum=None, e_out=None):
    """Accumulate row/column sums of exp(s*cos - s) for local rows `a` against all columns `b`.
    ``e_out`` (bf16 [rows, cols], bf16 path only): also keep E for the stored-E backward."""
    _need_cuda(a, b, scale, e_out)
    rows, D = a.shape
    cols = b.shape[0]
    dev = a.device
    if rowsum is None and colsum is None:
        both = torch.zeros(rows + cols, dtype=torch.float32, device=dev)  # one fill launch for the two accumulators
        rowsum, colsum = both[:rows], both[rows:]
    if rowsum is None:
        rowsum = torch.zeros(rows, dtype=torch.float32, device=dev)
    if colsum is None:
        colsum = torch.zeros(cols, dtype=torch.float32, device=dev)
    diag = torch.empty(rows, dtype=torch.float32, device=dev)
    lib = _lib.load()
    if e_out is not None:
        if prec != "bf16" or e_out.dtype != torch.bfloat16 or tuple(e_out.shape) != (rows, cols) or e_out.stride(1) != 1:
            raise ValueError("e_out must be a bf16 [rows, cols] tensor with contiguous rows (bf16 path only)")
        check(lib.mmg_infonce_fwd_store(_tc_prec(prec, a, b), _p(a), _p(b), rows, cols, D, diag_offset, _p(scale),
                                        _p(rowsum), _p(colsum), _p(diag), _p(e_out), e_out.stride(0), _stream()),
              "mmg_infonce_fwd_store")
        return rowsum, colsum, diag
    nbytes = lib.mmg_infonce_workspace_bytes(_PREC[prec], rows, cols, D)
    ws = _workspace(dev, nbytes)
    check(lib.mmg_infonce_fwd(_tc_prec(prec, a, b), _p(a), _p(b), rows, cols, D, diag_offset, _p(scale), _p(rowsum),
                              _p(colsum), _p(diag), _p(ws), ws.numel(), _stream()), "mmg_infonce_fwd")
    return rowsum, colsum, diag


def infonce_loss_raw(rowsum, colsum_slice, diag, scale, inv_two_b: float) -> torch.Tensor:
    out = torch.empty((), dtype=torch.float32, device=rowsum.device)
    check(_lib.load().mmg_infonce_loss(_p(rowsum), _p(colsum_slice), _p(diag), rowsum.numel(), _p(scale),
                                       float(inv_two_b), _p(out), _stream()), "mmg_infonce_loss")
    return out


def push_rows(src: torch.Tensor, dst_ptrs, dst_offset_bytes: int) -> None:
    """Copy the contiguous tensor ``src`` to ``dst_ptrs[i] + dst_offset_bytes`` for every destination (mmg_push_rows): the
    push half of the all-gather over NVLink peer memory (``dst_ptrs``: device pointers of every rank's symmetric buffer)."""
    _need_cuda(src)
    if not src.is_contiguous():
        raise ValueError("push_rows expects a contiguous source")
    n = len(dst_ptrs)
    ptrs = (ctypes.c_void_p * n)(*[int(x) for x in dst_ptrs])
    check(_lib.load().mmg_push_rows(_p(src), src.numel() * src.element_size(), ptrs, n, int(dst_offset_bytes), _stream()),
          "mmg_push_rows")


def infonce_row_part_raw(rowsum, diag, out=None) -> torch.Tensor:
    """out[0] = sum_r (log rowsum[r] - 2*diag[r]) over the local rows (mmg_infonce_row_part): the part of the sharded loss
    that is known before the exchange.  ``out``: a 1-element fp32 view to write into (e.g. the tail of the symmetric
    column-sum buffer, so that one all-reduce carries both)."""
    if out is None:
        out = torch.empty(1, dtype=torch.float32, device=rowsum.device)
    check(_lib.load().mmg_infonce_row_part(_p(rowsum), _p(diag), rowsum.numel(), _p(out), _stream()),
          "mmg_infonce_row_part")
    return out


def infonce_loss_cols_raw(colsum, scale, row_part, inv_two_b: float) -> torch.Tensor:
    """loss = inv_two_b * (row_part + sum_c log colsum[c] + 2*cols*s) from the GLOBAL column sums and the summed row
    parts (mmg_infonce_loss_cols); every rank computes the same value."""
    out = torch.empty((), dtype=torch.float32, device=colsum.device)
    check(_lib.load().mmg_infonce_loss_cols(_p(colsum), colsum.numel(), _p(scale), _p(row_part), float(inv_two_b), _p(out),
                                            _stream()), "mmg_infonce_loss_cols")
    return out


def infonce_backward_raw(a, b, scale, rowsum, colsum, grad_loss, inv_two_b: float, diag_offset: int, prec: str,
                         block_rows: int = 0, block_cols: int = 0, a32=None, b32=None, diag=None,
                         need_dscale: bool = True, e_stored=None):
    """Returns (dA [rows,D], dB_partial [cols,D], sum g*cos) -- see mmg_infonce_bwd in the header.

    ``need_dscale=False`` (the scale carries no gradient -- the reference's CUDA behaviour) skips the
    ``sum g*cos`` accumulation in the epilogue and returns ``None`` for it.

    When the fp32 embeddings (a32 [rows,D]; b32 [rows,D] = the column-side rows paired with the local rows) are supplied
    to the bf16 path, the matching-pair term of the
    gradient is applied from them in fp32 (mmg_infonce_bwd_diag) instead of through the bf16 contraction.

    ``e_stored`` (bf16 [rows, cols] written by ``infonce_forward_raw(e_out=...)``): stored-E backward, no ``sum g*cos``."""
    if e_stored is not None and need_dscale:
        raise ValueError("the stored-E backward cannot produce d/d logit_scale")
    rows, D = a.shape
    cols = b.shape[0]
    dev = a.device
    lib = _lib.load()
    rinv = torch.empty(rows, dtype=torch.float32, device=dev)
    cinv = torch.empty(cols, dtype=torch.float32, device=dev)
    scal = torch.empty(4, dtype=torch.float32, device=dev)
    gl = grad_loss.reshape(()).to(torch.float32).contiguous()
    diag_fp32 = prec == "bf16" and a32 is not None and b32 is not None and diag is not None
    pcode = _tc_prec(prec, a, b)
    if e_stored is not None and pcode != MMG_PREC_BF16:
        raise ValueError("the stored-E backward takes bf16 operands (set_embedding_f16(False))")
    dls = torch.zeros((), dtype=torch.float32, device=dev) if need_dscale else None
    if diag_fp32:
        # the fp32 matching-pair term is the FIRST writer of the gradient rows (no zero-fill, no read-modify-write);
        # the contraction kernels then accumulate on top.  Column rows without a local partner start from zero.
        # The same launch writes rinv / cinv / scal (mmg_infonce_bwd_prep_diag).
        dA = torch.empty((rows, D), dtype=torch.float32, device=dev)
        if cols == rows:
            dB = torch.empty((cols, D), dtype=torch.float32, device=dev)
        else:
            dB = torch.zeros((cols, D), dtype=torch.float32, device=dev)
        dBm = dB[diag_offset:diag_offset + rows]
        check(lib.mmg_infonce_bwd_prep_diag(pcode, _p(rowsum), rows, _p(colsum), cols, diag_offset, _p(scale), _p(gl),
                                            float(inv_two_b), _p(rinv), _p(cinv), _p(scal), _p(a32), _p(b32), D, _p(diag),
                                            _p(dA), _p(dBm), _p(dls), _stream()), "mmg_infonce_bwd_prep_diag")
    else:
        check(lib.mmg_infonce_bwd_prep(pcode, _p(rowsum), rows, _p(colsum), cols, _p(scale), _p(gl), float(inv_two_b),
                                       0, _p(rinv), _p(cinv), _p(scal), _stream()), "mmg_infonce_bwd_prep")
        dA = torch.zeros((rows, D), dtype=torch.float32, device=dev)
        dB = torch.zeros((cols, D), dtype=torch.float32, device=dev)
    block_rows = block_rows or int(os.environ.get("MMGCLIP_B200_BLOCK_ROWS", "0"))
    block_cols = block_cols or int(os.environ.get("MMGCLIP_B200_BLOCK_COLS", "0"))
    nbytes = lib.mmg_infonce_workspace_bytes(_PREC[prec], rows, cols, D)
    if block_rows or block_cols:
        esz = 2 if prec == "bf16" else 4
        br = min(rows, block_rows or 4096)
        bc = (min(cols, block_cols or 4096) + 63) // 64 * 64
        nbytes = max(nbytes, br * bc * esz + 256)
    ws = _workspace(dev, nbytes)
    if _bwd_probe is not None:
        _bwd_probe[0].record()
    if e_stored is not None:
        owners = (ctypes.c_void_p * 1)(dB.data_ptr())
        check(lib.mmg_infonce_bwd_stored(pcode, _p(a), _p(b), _p(e_stored), e_stored.stride(0), rows, cols,
                                         D, diag_offset, _p(scale), _p(rinv), _p(cinv), _p(scal), _p(dA), owners, 1, 1, 0,
                                         _p(ws), ws.numel(), _stream()), "mmg_infonce_bwd_stored")
    else:
        check(lib.mmg_infonce_bwd(pcode, _p(a), _p(b), rows, cols, D, diag_offset, _p(scale), _p(rinv), _p(cinv),
                                  _p(scal), _p(dA), _p(dB), _p(dls), block_rows, block_cols, _p(ws), ws.numel(),
                                  _stream()), "mmg_infonce_bwd")
    if _bwd_probe is not None:
        _bwd_probe[1].record()
    return dA, dB, dls


def infonce_backward_owners(a, b, scale, rowsum, colsum, grad_loss, inv_two_b: float, diag_offset: int, parts,
                            pre_sync=None, post_sync=None, after_part=None, a32=None, b32=None, diag=None,
                            need_dscale: bool = True, e_stored=None):
    """Row-sharded backward whose column-side gradient goes straight to its owners (mmg_infonce_bwd_owners).

    ``parts`` is a list of ``(own, owner_ptrs)``, one entry per column part (usually one): ``own`` is this rank's fp32
    ``[cols/world/n_parts, D]`` gradient buffer for that part of its columns and ``owner_ptrs`` the device pointers of
    every rank's buffer for the part, in rank order (NVLink peer mappings, or slices of a local staging buffer that a
    reduce-scatter sends home).  ``pre_sync`` / ``post_sync``: stream-ordered cross-rank barriers for the peer-memory
    case (every owner's buffer is initialised before anybody adds into it; all adds have landed before anybody reads its
    own).  ``after_part(i)`` runs right after part i has been launched (e.g. to start its reduce-scatter).
    ``e_stored`` (bf16 [rows, cols] kept by the forward): stored-E backward (mmg_infonce_bwd_stored), no ``sum g*cos``.
    Returns (dA [rows, D], [own buffers], sum g*cos or None)."""
    if e_stored is not None and need_dscale:
        raise ValueError("the stored-E backward cannot produce d/d logit_scale")
    rows, D = a.shape
    cols = b.shape[0]
    n_parts = len(parts)
    world = len(parts[0][1])
    rp = rows // n_parts
    if rows * world != cols or diag_offset % rows != 0 or rp * n_parts != rows:
        raise ValueError("infonce_backward_owners: every rank must own cols/world rows paired with its local rows")
    for own, ptrs in parts:
        if tuple(own.shape) != (rp, D) or own.dtype != torch.float32 or len(ptrs) != world:
            raise ValueError("infonce_backward_owners: each part needs an fp32 [rows/n_parts, D] buffer per owner")
    dev = a.device
    lib = _lib.load()
    rinv = torch.empty(rows, dtype=torch.float32, device=dev)
    cinv = torch.empty(cols, dtype=torch.float32, device=dev)
    scal = torch.empty(4, dtype=torch.float32, device=dev)
    gl = grad_loss.reshape(()).to(torch.float32).contiguous()
    diag_fp32 = a32 is not None and b32 is not None and diag is not None
    pcode = _tc_prec("bf16", a, b)
    if e_stored is not None and pcode != MMG_PREC_BF16:
        raise ValueError("the stored-E backward takes bf16 operands (set_embedding_f16(False))")
    dls = torch.zeros((), dtype=torch.float32, device=dev) if need_dscale else None
    if diag_fp32 and n_parts == 1:
        # one launch: rinv / cinv / scal + the matching-pair term as first writer of dA and of this rank's own buffer
        dA = torch.empty((rows, D), dtype=torch.float32, device=dev)
        check(lib.mmg_infonce_bwd_prep_diag(pcode, _p(rowsum), rows, _p(colsum), cols, diag_offset, _p(scale), _p(gl),
                                            float(inv_two_b), _p(rinv), _p(cinv), _p(scal), _p(a32), _p(b32), D, _p(diag),
                                            _p(dA), _p(parts[0][0]), _p(dls), _stream()), "mmg_infonce_bwd_prep_diag")
    else:
        check(lib.mmg_infonce_bwd_prep(pcode, _p(rowsum), rows, _p(colsum), cols, _p(scale), _p(gl), float(inv_two_b),
                                       int(diag_fp32), _p(rinv), _p(cinv), _p(scal), _stream()), "mmg_infonce_bwd_prep")
    if diag_fp32 and n_parts == 1:
        pass
    elif diag_fp32:
        # the matching-pair term is the first writer of dA and of this rank's own buffers (part i = local rows
        # [i*rp, (i+1)*rp), whose partners are rows [0, rp) of the part's buffer)
        dA = torch.empty((rows, D), dtype=torch.float32, device=dev)
        for i, (own, _) in enumerate(parts):
            r0 = i * rp
            cinvm = cinv[diag_offset + r0:diag_offset + r0 + rp]
            check(lib.mmg_infonce_bwd_diag(_p(a32[r0:r0 + rp]), _p(b32[r0:r0 + rp]), rp, D, _p(diag[r0:r0 + rp]), _p(scale),
                                           _p(rinv[r0:r0 + rp]), _p(cinvm), _p(scal), _p(dA[r0:r0 + rp]), _p(own), _p(dls),
                                           1, _stream()), "mmg_infonce_bwd_diag")
    else:
        dA = torch.zeros((rows, D), dtype=torch.float32, device=dev)
        for own, _ in parts:
            own.zero_()
    nbytes = lib.mmg_infonce_workspace_bytes(_PREC["bf16"], rows, cols, D)
    ws = _workspace(dev, nbytes)
    mark("bwd_prep")
    if pre_sync is not None:
        pre_sync()
    mark("pre_sync")
    for i, (_, owner_ptrs) in enumerate(parts):
        ptrs = (ctypes.c_void_p * world)(*[int(x) for x in owner_ptrs])
        if e_stored is not None:
            check(lib.mmg_infonce_bwd_stored(pcode, _p(a), _p(b), _p(e_stored), e_stored.stride(0), rows,
                                             cols, D, diag_offset, _p(scale), _p(rinv), _p(cinv), _p(scal), _p(dA), ptrs,
                                             world, n_parts, i, _p(ws), ws.numel(), _stream()), "mmg_infonce_bwd_stored")
        else:
            check(lib.mmg_infonce_bwd_owners(pcode, _p(a), _p(b), rows, cols, D, diag_offset, _p(scale),
                                             _p(rinv), _p(cinv),
                                             _p(scal), _p(dA), ptrs, world, n_parts, i, _p(dls), _p(ws), ws.numel(),
                                             _stream()), "mmg_infonce_bwd_owners")
        if after_part is not None:
            after_part(i)
    mark("bwd_fused")
    if post_sync is not None:
        post_sync()
    mark("post_sync")
    return dA, [own for own, _ in parts], dls


# ----------------------------------------------------------------------------------------------------------------
# autograd operators
# ----------------------------------------------------------------------------------------------------------------
# nn.Dropout state of the in-kernel generator (mmg_dropout_draw_apply): {seed, offset, ticket} as three int64 on each
# device.  The seed is taken from torch.initial_seed() when a device first draws (so torch.manual_seed() before building
# the model fixes the masks); seed_dropout() re-seeds explicitly.  The offset advances on the device, launch by launch.
_dropout_states = {}


def seed_dropout(seed: int, device=None) -> None:
    """Restart the dropout stream of ``device`` (default: current CUDA device) from ``seed``."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    _dropout_states[dev.index] = torch.tensor([int(seed) & 0x7FFFFFFFFFFFFFFF, 0, 0], dtype=torch.int64, device=dev)


def _dropout_state(dev: torch.device) -> torch.Tensor:
    st = _dropout_states.get(dev.index)
    if st is None:
        seed_dropout(torch.initial_seed(), dev)
        st = _dropout_states[dev.index]
    return st


def dropout_draw_apply(y: torch.Tensor, p: float) -> torch.Tensor:
    """In place: y = dropout(y, p) with the keep mask drawn in the kernel; returns the uint8 mask (for the backward)."""
    _need_cuda(y)
    mask = torch.empty(y.shape, dtype=torch.uint8, device=y.device)
    if y.numel() > 0:
        check(_lib.load().mmg_dropout_draw_apply(_p(y), _p(mask), float(p), y.numel(), _p(_dropout_state(y.device)),
                                                 _stream()), "mmg_dropout_draw_apply")
    return mask


class _LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) [* dropout mask]; nn.Linear semantics (projection.py:17,45-59,88-97).  The mask is either given
    (``mask`` uint8 + ``keep_scale``) or drawn in the kernel (``drop_p`` > 0)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu: bool, mask, keep_scale: float, prec: str, drop_p: float = 0.0):
        _need_cuda(x, weight, bias, mask)
        if x.dim() != 2:
            raise ValueError("projection heads take [B, E] inputs")
        Bn, E = x.shape
        D = weight.shape[0]
        if weight.shape[1] != E:
            raise ValueError(f"mat1 and mat2 shapes cannot be multiplied ({Bn}x{E} and {weight.shape[1]}x{D})")
        x = x.contiguous()
        w = weight.contiguous()
        b = bias.contiguous() if bias is not None else None
        if Bn == 0:
            ctx.empty = True
            return x.new_zeros((0, D))
        ctx.empty = False
        xo, wo = _Operand.of(x, prec), _Operand.of(w, prec)
        y = gemm_heads(xo, wo, Bn, D, E, bias=b, relu=relu, prec=prec)
        if mask is not None:
            check(_lib.load().mmg_dropout_apply(_p(y), _p(mask), float(keep_scale), y.numel(), _stream()),
                  "mmg_dropout_apply")
        elif drop_p > 0.0:
            mask = dropout_draw_apply(y, drop_p)
            keep_scale = 1.0 / (1.0 - drop_p) if drop_p < 1.0 else 0.0
            ctx.mask_drawn = mask  # readable by tests / callers that want the mask (weak convenience, not an output)
        ctx.prec, ctx.relu, ctx.keep_scale = prec, relu, keep_scale
        ctx.has_bias = bias is not None
        ctx.nx, ctx.nw = len(xo.tensors()), len(wo.tensors())
        ctx.shape = (Bn, E, D)
        ctx.save_for_backward(*xo.tensors(), *wo.tensors(), y if (relu or mask is not None) else None, mask)
        return y

    @staticmethod
    def backward(ctx, dy):
        if ctx.empty:
            return None, None, None, None, None, None, None, None
        saved = ctx.saved_tensors
        xo = _Operand.from_tensors(saved[:ctx.nx])
        wo = _Operand.from_tensors(saved[ctx.nx:ctx.nx + ctx.nw])
        y, mask = saved[ctx.nx + ctx.nw], saved[ctx.nx + ctx.nw + 1]
        prec = ctx.prec
        Bn, E, D = ctx.shape
        dy = dy.contiguous()
        if ctx.relu or mask is not None:
            dz = torch.empty_like(dy)
            yy = y if ctx.relu else None
            check(_lib.load().mmg_relu_dropout_bwd(_p(dy), _p(yy), _p(mask), float(ctx.keep_scale), _p(dz), dz.numel(),
                                                   _stream()), "mmg_relu_dropout_bwd")
        else:
            dz = dy
        dzo = _Operand.of(dz, prec)
        dx = dw = db = None
        if ctx.needs_input_grad[1]:
            da, xa = _dw_operands(dzo, xo)
            ks = _split_k_for(D, E, Bn * _n_segments(da, xa)) if prec == "bf16" else 1
            dw = gemm_heads(da, xa, D, E, Bn, a_mn=True, b_mn=True, prec=prec, k_splits=ks)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.empty(D, dtype=torch.float32, device=dy.device)
            check(_lib.load().mmg_colsum(_p(dz), Bn, D, _p(db), _stream()), "mmg_colsum")
        if ctx.needs_input_grad[0]:
            dx = gemm_heads(dzo, wo, Bn, E, D, b_mn=True, prec=prec)
        return dx, dw, db, None, None, None, None, None


# the keep mask of the most recent in-kernel dropout draw (what a test or a debugging session compares an oracle against)
last_dropout_mask = None


def linear(x, weight, bias=None, relu=False, mask=None, keep_scale=1.0, prec=None, drop_p=0.0):
    """``drop_p`` > 0: inverted dropout after the activation with the mask drawn in the kernel (mmg_dropout_draw_apply)."""
    global last_dropout_mask
    y = _LinearFn.apply(x, weight, bias, relu, mask, keep_scale, _resolve(prec), float(drop_p))
    if drop_p > 0.0 and mask is None and y.grad_fn is not None:
        last_dropout_mask = getattr(y.grad_fn, "mask_drawn", None)
    return y


class _L2NormFn(torch.autograd.Function):
    """x / x.norm(dim=1, keepdim=True)  (mmgclip_model.py:128-129), no epsilon."""

    @staticmethod
    def forward(ctx, u, want_bf16: bool):
        y, inv, yb = l2norm_fwd(u, want_bf16)
        ctx.save_for_backward(y, inv)
        if yb is None:
            yb = y.new_empty(0)
        ctx.mark_non_differentiable(yb)
        ctx.set_materialize_grads(False)  # no zero-filled [B, D] gradient for the bf16 copy (an 11 us fill at B = 32768)
        return y, yb

    @staticmethod
    def backward(ctx, dy, _unused):
        if dy is None:
            return None, None
        y, inv = ctx.saved_tensors
        du, _ = l2norm_bwd(dy, y, inv, True, False)
        return du, None


def l2_normalize(u, prec=None):
    """Row-normalise; the bf16 operand copy the loss kernel needs is produced in the same pass and attached."""
    prec = _resolve(prec)
    y, yb = _L2NormFn.apply(u, prec == "bf16")
    if prec == "bf16":
        y._mmg_bf16 = yb
    return y


class _ProjNormFn(torch.autograd.Function):
    """Bias-free projection + L2 normalise in one operator (LinearProjectionLayer followed by mmgclip_model.py:128).

    Forward: u = x W^T (tensor cores) -> y = u/||u|| (+ bf16 copy).  Backward: du = (dy - y<y,dy>)/||u|| written
    directly as the bf16 MN-major operand of dW = du^T x (split-K over the batch)."""

    @staticmethod
    def forward(ctx, x, weight, prec: str):
        _need_cuda(x, weight)
        Bn, E = x.shape
        D = weight.shape[0]
        if weight.shape[1] != E:
            raise ValueError(f"mat1 and mat2 shapes cannot be multiplied ({Bn}x{E} and {weight.shape[1]}x{D})")
        x = x.contiguous()
        w = weight.contiguous()
        xo, wo = _Operand.of(x, prec), _Operand.of(w, prec)
        u = gemm_heads(xo, wo, Bn, D, E, prec=prec)
        y, inv, yb = l2norm_fwd(u, prec == "bf16")
        ctx.prec = prec
        ctx.nx, ctx.nw = len(xo.tensors()), len(wo.tensors())
        ctx.shape = (Bn, E, D)
        ctx.save_for_backward(*xo.tensors(), *wo.tensors(), y, inv)
        if yb is None:
            yb = y.new_empty(0)
        ctx.mark_non_differentiable(yb)
        ctx.set_materialize_grads(False)  # no zero-filled [B, D] gradient for the bf16 copy (an 11 us fill at B = 32768)
        return y, yb

    @staticmethod
    def backward(ctx, dy, _unused):
        if dy is None:
            return None, None, None
        saved = ctx.saved_tensors
        xo = _Operand.from_tensors(saved[:ctx.nx])
        wo = _Operand.from_tensors(saved[ctx.nx:ctx.nx + ctx.nw])
        y, inv = saved[ctx.nx + ctx.nw], saved[ctx.nx + ctx.nw + 1]
        prec = ctx.prec
        Bn, E, D = ctx.shape
        need_dx = ctx.needs_input_grad[0]
        dw_buf = None
        if prec == "bf16":
            split = xo.lo is not None and (_split_dw or need_dx)
            ks = 1
            if ctx.needs_input_grad[1]:
                nseg = (1 + 2 * int(split)) if _split_dw else 1
                ks = _split_k_for(D, E, Bn * nseg)
                if ks > 1 and (D * E) % 4 == 0:
                    # the split-K weight gradient accumulates with reduce-adds: its output is cleared by the normalise
                    # backward launch that precedes it instead of a fill launch of its own
                    dw_buf = torch.empty((D, E), dtype=torch.float32, device=dy.device)
            res = l2norm_bwd(dy, y, inv, False, True, want_lo=split, zero=dw_buf)
            dz = _Operand(res[1], res[2] if split else None)
        else:
            dz = _Operand(l2norm_bwd(dy, y, inv, True, False)[0])
        dw = dx = None
        if ctx.needs_input_grad[1]:
            da, xa = _dw_operands(dz, xo)
            ks = _split_k_for(D, E, Bn * _n_segments(da, xa)) if prec == "bf16" else 1
            dw = gemm_heads(da, xa, D, E, Bn, a_mn=True, b_mn=True, prec=prec, k_splits=ks,
                            zeroed_out=dw_buf if ks > 1 else None)
        if need_dx:
            dx = gemm_heads(dz, wo, Bn, E, D, b_mn=True, prec=prec)
        return dx, dw, None


def project_normalize(x, weight, prec=None):
    prec = _resolve(prec)
    y, yb = _ProjNormFn.apply(x, weight, prec)
    if prec == "bf16":
        y._mmg_bf16 = yb
    return y


def _operand(t: torch.Tensor, prec: str) -> torch.Tensor:
    """The kernel operand for an embedding matrix: itself (fp32) or its bf16 copy (reused if a producer attached one)."""
    if prec == "fp32":
        if t.dtype != torch.float32:
            raise ValueError("fp32 path expects float32 embeddings")
        return t.detach().contiguous()
    if t.dtype in (torch.bfloat16, torch.float16):
        return t.detach().contiguous()
    cached = getattr(t, "_mmg_bf16", None)  # the 16-bit copy the normalise kernel attached (fp16 or bf16)
    if cached is not None and cached.shape == t.shape and cached.device == t.device:
        return cached
    return cast_embedding(t.detach())


def _operand_bf16(t: torch.Tensor, prec: str) -> torch.Tensor:
    """Like _operand, for consumers that contract in bf16 only (the materialised-logits contraction, mmg_gemm)."""
    o = _operand(t, prec)
    if prec == "fp32" or o.dtype == torch.bfloat16:
        return o
    return cast_bf16(t.detach().to(torch.float32))


class _InfoNCEFn(torch.autograd.Function):
    """(CE(L, arange) + CE(L^T, arange)) / 2 with L = s * A B^T, never materialising L (losses.py:28-44, 73-82)."""

    @staticmethod
    def forward(ctx, a_hat, b_hat, scale, a_op, b_op, prec: str):
        n, D = a_hat.shape
        if b_hat.shape != a_hat.shape:
            raise ValueError(f"paired InfoNCE needs equal shapes, got {tuple(a_hat.shape)} and {tuple(b_hat.shape)}")
        s = scale.detach().reshape(()).to(device=a_hat.device, dtype=torch.float32).contiguous()
        e_mat = None
        if (a_op.dtype == torch.bfloat16 and want_store_e(n, n, D, prec, ctx.needs_input_grad[2])
                and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1])):
            try:
                e_mat = torch.empty((n, n), dtype=torch.bfloat16, device=a_hat.device)
            except torch.OutOfMemoryError:
                e_mat = None  # not enough free HBM for E: the recompute path needs O(B*D) only
        rowsum, colsum, diag = infonce_forward_raw(a_op, b_op, s, 0, prec, e_out=e_mat)
        loss = infonce_loss_raw(rowsum, colsum, diag, s, 0.5 / n)
        ctx.prec = prec
        ctx.e_mat = e_mat  # not an input or output of the function: kept on the ctx
        # fp32 embeddings (when that is what the caller holds) serve the matching-pair term of the backward
        keep32 = prec == "bf16" and a_hat.dtype == torch.float32 and b_hat.dtype == torch.float32
        ctx.keep32 = keep32
        if keep32:
            ctx.save_for_backward(a_op, b_op, s, rowsum, colsum, a_hat.detach().contiguous(), b_hat.detach().contiguous(),
                                  diag)
        else:
            ctx.save_for_backward(a_op, b_op, s, rowsum, colsum)
        ctx.scale_shape = scale.shape
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        if ctx.keep32:
            a_op, b_op, s, rowsum, colsum, a32, b32, diag = ctx.saved_tensors
        else:
            a_op, b_op, s, rowsum, colsum = ctx.saved_tensors
            a32 = b32 = diag = None
        n = a_op.shape[0]
        dA, dB, dls = infonce_backward_raw(a_op, b_op, s, rowsum, colsum, grad_loss, 0.5 / n, 0, ctx.prec,
                                           a32=a32, b32=b32, diag=diag, need_dscale=ctx.needs_input_grad[2],
                                           e_stored=ctx.e_mat)
        ctx.e_mat = None
        dscale = None
        if ctx.needs_input_grad[2]:
            dscale = (dls / s).reshape(ctx.scale_shape)  # d loss / d s ; sum g*cos = s * dloss/ds
        return (dA if ctx.needs_input_grad[0] else None, dB if ctx.needs_input_grad[1] else None, dscale, None, None,
                None)


# Range of the fixed softmax shift.  The fused kernels use m = s for every row and column: E = exp(s*cos - s).  That needs
# |cos| <= 1 (L2-normalised inputs) and is unconditionally safe while 2*s stays below -log(FLT_MIN) = 87.3: no term can
# underflow, so no row / column sum can vanish.  Beyond that a row whose best cosine is far from 1 can lose ALL its terms
# (s * (1 - max cos) > 87), where the reference's per-row-max cross-entropy stays finite.  Policy:
#   * scale known on the host (float / CPU tensor) and above the bound: the materialised, row-max-stabilised path
#     (similarity_logits + cross_entropy) up to FIXED_SHIFT_FALLBACK_ROWS rows, ValueError above;
#   * scale on the device (no host sync on the hot path): the loss kernel turns a vanished / overflowed sum into a NaN
#     loss instead of silent inf gradients (simt_kernels.cu, infonce_loss_kernel).
FIXED_SHIFT_MAX_SCALE = 43.0
FIXED_SHIFT_FALLBACK_ROWS = 16384


def _host_scale(logit_scale):
    """The scale's value if it can be had without a device sync, else None."""
    if not torch.is_tensor(logit_scale):
        return float(logit_scale)
    if not logit_scale.is_cuda:
        return float(logit_scale.detach().reshape(()))
    return None


def info_nce(a_hat: torch.Tensor, b_hat: torch.Tensor, logit_scale: torch.Tensor, prec: Optional[str] = None):
    """Symmetric InfoNCE of L2-normalised [n, D] embeddings; `logit_scale` is the already-exponentiated scale.

    Inputs must be L2-normalised and the scale moderate (see FIXED_SHIFT_MAX_SCALE above; CLIP's 1/0.07 = 14.3 and its
    usual clamp region up to 43 are inside)."""
    prec = _resolve(prec)
    _need_cuda(a_hat, b_hat)
    s_host = _host_scale(logit_scale)
    if s_host is not None and not (0.0 < s_host <= FIXED_SHIFT_MAX_SCALE):
        if s_host <= 0.0 or s_host != s_host:
            raise ValueError(f"info_nce: logit_scale must be a positive number (got {s_host})")
        n = a_hat.shape[0]
        if n > FIXED_SHIFT_FALLBACK_ROWS:
            raise ValueError(f"info_nce: logit_scale {s_host:.1f} exceeds the fixed-shift range ({FIXED_SHIFT_MAX_SCALE}) and "
                             f"{n} rows are too many for the materialised fallback (<= {FIXED_SHIFT_FALLBACK_ROWS})")
        lpi = similarity_logits(a_hat, b_hat, logit_scale, prec=prec)
        lpt = similarity_logits(b_hat, a_hat, logit_scale, prec=prec)
        return (cross_entropy(lpi) + cross_entropy(lpt)) / 2
    if not torch.is_tensor(logit_scale):
        logit_scale = torch.tensor(float(logit_scale), dtype=torch.float32, device=a_hat.device)
    return _InfoNCEFn.apply(a_hat, b_hat, logit_scale, _operand(a_hat, prec), _operand(b_hat, prec), prec)


class _CEFn(torch.autograd.Function):
    """coef * sum_r (logsumexp(logits[r]) - logits[r, labels[r]]); labels None = arange(n)."""

    @staticmethod
    def forward(ctx, logits, labels, coef: float):
        _need_cuda(logits, labels)
        if logits.dim() != 2:
            raise ValueError("cross entropy expects logits [n, m]")
        if labels is None and logits.shape[1] < logits.shape[0]:
            raise ValueError("cross entropy with labels arange(n) needs logits [n, m] with m >= n")
        lg = logits.detach().to(torch.float32).contiguous()
        n, m = lg.shape
        lab = None
        if labels is not None:
            lab = labels.detach().to(device=lg.device, dtype=torch.int64).contiguous()
            if lab.numel() != n:
                raise ValueError(f"Expected input batch_size ({n}) to match target batch_size ({lab.numel()}).")
        lse = torch.empty(n, dtype=torch.float32, device=lg.device)
        out = torch.zeros((), dtype=torch.float32, device=lg.device)
        check(_lib.load().mmg_ce_fwd(_p(lg), lg.stride(0), n, m, _p(lab), float(coef), _p(lse), _p(out), _stream()),
              "mmg_ce_fwd")
        ctx.coef = coef
        ctx.has_labels = lab is not None
        if lab is not None:
            ctx.save_for_backward(lg, lse, lab)
        else:
            ctx.save_for_backward(lg, lse)
        return out

    @staticmethod
    def backward(ctx, grad):
        if ctx.has_labels:
            lg, lse, lab = ctx.saved_tensors
        else:
            lg, lse = ctx.saved_tensors
            lab = None
        n, m = lg.shape
        d = torch.empty_like(lg)
        gl = grad.reshape(()).to(torch.float32).contiguous()
        check(_lib.load().mmg_ce_bwd(_p(lg), lg.stride(0), n, m, _p(lab), _p(lse), _p(gl), float(ctx.coef), _p(d),
                                     d.stride(0), _stream()), "mmg_ce_bwd")
        return d, None, None


def cross_entropy(logits: torch.Tensor, labels: Optional[torch.Tensor] = None) -> torch.Tensor:
    """F.cross_entropy(logits, labels) with mean reduction; labels None = arange(n)."""
    return _CEFn.apply(logits, labels, 1.0 / logits.shape[0])


def ce_arange(logits: torch.Tensor, coef: float) -> torch.Tensor:
    return _CEFn.apply(logits, None, coef)


class _LogitsFn(torch.autograd.Function):
    """L = (s * A) @ B^T materialised (mmgclip_model.py:135-136) -- the n != m / evaluation / small-n case."""

    @staticmethod
    def forward(ctx, a, b, scale, prec: str):
        _need_cuda(a, b)
        n, D = a.shape
        m = b.shape[0]
        if b.shape[1] != D:
            raise ValueError(f"mat1 and mat2 shapes cannot be multiplied ({n}x{D} and {b.shape[1]}x{m})")
        s = scale.detach().reshape(()).to(device=a.device, dtype=torch.float32).contiguous()
        ao, bo = _operand_bf16(a, prec), _operand_bf16(b, prec)
        logits = gemm(ao, bo, n, m, D, alpha_dev=s, prec=prec)  # s applied in the contraction's epilogue
        ctx.prec = prec
        ctx.scale_shape = scale.shape
        ctx.save_for_backward(ao, bo, s, logits)
        return logits

    @staticmethod
    def backward(ctx, dL):
        ao, bo, s, logits = ctx.saved_tensors
        prec = ctx.prec
        n, D = ao.shape
        m = bo.shape[0]
        dL = dL.contiguous()
        dLo = cast_bf16(dL) if prec == "bf16" else dL
        dA = dB = dS = None
        if ctx.needs_input_grad[0]:
            dA = gemm(dLo, bo, n, D, m, b_mn=True, alpha_dev=s, prec=prec)
        if ctx.needs_input_grad[1]:
            dB = gemm(dLo, ao, m, D, n, a_mn=True, b_mn=True, alpha_dev=s, prec=prec)
        if ctx.needs_input_grad[2]:
            out = torch.empty((), dtype=torch.float32, device=dL.device)
            check(_lib.load().mmg_dot_sum(_p(dL), _p(logits), dL.numel(), _p(out), _stream()), "mmg_dot_sum")
            dS = (out / s).reshape(ctx.scale_shape)  # sum dL*logits = s * sum dL*cos; 0-d glue only
        return dA, dB, dS, None


def similarity_logits(a_hat, b_hat, logit_scale, prec: Optional[str] = None) -> torch.Tensor:
    prec = _resolve(prec)
    if not torch.is_tensor(logit_scale):
        logit_scale = torch.tensor(float(logit_scale), dtype=torch.float32, device=a_hat.device)
    return _LogitsFn.apply(a_hat, b_hat, logit_scale, prec)


class _GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _need_cuda(x)
        x = x.contiguous()
        y = torch.empty_like(x)
        check(_lib.load().mmg_gelu_fwd(_p(x), _p(y), x.numel(), _stream()), "mmg_gelu_fwd")
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        check(_lib.load().mmg_gelu_bwd(_p(dy), _p(x), _p(dx), x.numel(), _stream()), "mmg_gelu_bwd")
        return dx


def gelu(x):
    return _GeluFn.apply(x)


class _AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        _need_cuda(x, y)
        if x.shape != y.shape:
            raise ValueError("residual_add expects equal shapes")
        x, y = x.contiguous(), y.contiguous()
        out = torch.empty_like(x)
        check(_lib.load().mmg_add(_p(x), _p(y), _p(out), x.numel(), _stream()), "mmg_add")
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


def residual_add(x, y):
    return _AddFn.apply(x, y)


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps: float):
        _need_cuda(x, gamma, beta)
        x = x.contiguous()
        rows, cols = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        g, b = gamma.contiguous(), beta.contiguous()
        check(_lib.load().mmg_layernorm_fwd(_p(x), _p(g), _p(b), rows, cols, float(eps), _p(y), _p(mean), _p(rstd),
                                            _stream()), "mmg_layernorm_fwd")
        ctx.save_for_backward(x, g, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, g, mean, rstd = ctx.saved_tensors
        rows, cols = x.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dg = torch.empty(cols, dtype=torch.float32, device=x.device)
        db = torch.empty(cols, dtype=torch.float32, device=x.device)
        check(_lib.load().mmg_layernorm_bwd(_p(dy), _p(x), _p(g), _p(mean), _p(rstd), rows, cols, _p(dx), _p(dg),
                                            _p(db), _stream()), "mmg_layernorm_bwd")
        return dx, dg, db, None


def layer_norm(x, gamma, beta, eps=1e-5):
    return _LayerNormFn.apply(x, gamma, beta, eps)


# ----------------------------------------------------------------------------------------------------------------
# zero-shot scoring (inference only)
# ----------------------------------------------------------------------------------------------------------------
_ZEROSHOT_TC_MIN_ROWS = 4096


def zeroshot_score(image_embeddings: torch.Tensor, text_embeddings: torch.Tensor, logit_scale, k: int = 0,
                   want_logits: bool = True, want_probs: bool = True, impl: str = "auto"):
    """logits = (s*I) @ T^T, softmax(-1), argmax (ties -> lowest index), top-k (value desc, index asc).

    Returns dict(logits, probs, argmax[int64], topk_idx[int64, k], topk_val).  k <= 8; any number of prompts (more than
    64 take a one-warp-per-row kernel instead of the tiled ones).
    (mmgclip_model.py:201-209; evaluator.py:182-188, 282-299, 354-368)

    ``impl``: "ffma" = fp32 FFMA kernel (what small batches use); "tc" = tensor-core kernel (3xTF32 split, logits within
    ~1e-6 of the FFMA evaluation, HBM-bound at large N); "auto" = "tc" from 4096 rows when the layout allows it."""
    _need_cuda(image_embeddings, text_embeddings)
    img = image_embeddings.detach().to(torch.float32).contiguous()
    txt = text_embeddings.detach().to(torch.float32).contiguous()
    N, D = img.shape
    C = txt.shape[0]
    if txt.shape[1] != D:
        raise ValueError(f"embedding dims differ: {D} vs {txt.shape[1]}")
    dev = img.device
    if torch.is_tensor(logit_scale):
        s = logit_scale.detach().reshape(()).to(device=dev, dtype=torch.float32).contiguous()
    else:
        s = torch.tensor(float(logit_scale), dtype=torch.float32, device=dev)
    logits = torch.empty((N, C), dtype=torch.float32, device=dev) if (want_logits or C > 64) else None
    probs = torch.empty((N, C), dtype=torch.float32, device=dev) if want_probs else None
    amax = torch.empty(N, dtype=torch.int64, device=dev)
    tki = torch.empty((N, k), dtype=torch.int64, device=dev) if k > 0 else None
    tkv = torch.empty((N, k), dtype=torch.float32, device=dev) if k > 0 else None
    if impl not in ("auto", "tc", "ffma"):
        raise ValueError(f"Invalid impl: {impl}")
    lib = _lib.load()
    tc_ok = N > 0 and C <= 64 and D % 4 == 0 and img.data_ptr() % 16 == 0
    if impl == "tc" and not tc_ok:
        raise ValueError("zeroshot_score(impl='tc') needs C <= 64, D % 4 == 0 and 16-byte aligned embeddings")
    if tc_ok and (impl == "tc" or (impl == "auto" and N >= _ZEROSHOT_TC_MIN_ROWS)):
        ws = _workspace(dev, lib.mmg_zeroshot_workspace_bytes(C, D))
        check(lib.mmg_zeroshot_score_tc(_p(img), _p(txt), N, C, D, _p(s), _p(logits), _p(probs), _p(amax), k, _p(tki),
                                        _p(tkv), _p(ws), ws.numel(), _stream()), "mmg_zeroshot_score_tc")
    else:
        check(lib.mmg_zeroshot_score(_p(img), _p(txt), N, C, D, _p(s), _p(logits), _p(probs), _p(amax), k, _p(tki),
                                     _p(tkv), _stream()), "mmg_zeroshot_score")
    return {"logits": logits if want_logits else None, "probs": probs, "argmax": amax, "topk_idx": tki, "topk_val": tkv}


# -------------------------------------------------------------------------------------------------------------------
# 'eos' text pooling: hidden state of the last attended token, mmgclip_model.py:108-111
# -------------------------------------------------------------------------------------------------------------------
class _EosPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, attention_mask):
        _need_cuda(hidden, attention_mask)
        if hidden.dim() != 3 or attention_mask.dim() != 2 or attention_mask.shape != hidden.shape[:2]:
            raise ValueError("eos_pool expects hidden [n, seq, H] and attention_mask [n, seq]")
        if hidden.dtype != torch.float32:
            raise ValueError("eos_pool expects fp32 hidden states (the reference encoders run in fp32)")
        hidden = hidden.contiguous()
        mask = attention_mask.to(torch.int64).contiguous()
        n, seq, H = hidden.shape
        out = torch.empty((n, H), dtype=torch.float32, device=hidden.device)
        idx = torch.empty((n,), dtype=torch.int64, device=hidden.device)
        check(_lib.load().mmg_eos_pool(_p(hidden), _p(mask), n, seq, H, _p(out), _p(idx), _stream()), "mmg_eos_pool")
        ctx.save_for_backward(idx)
        ctx.shape = (n, seq, H)
        ctx.mark_non_differentiable(idx)
        return out, idx

    @staticmethod
    def backward(ctx, g, _g_idx):
        (idx,) = ctx.saved_tensors
        n, seq, H = ctx.shape
        dh = torch.empty((n, seq, H), dtype=torch.float32, device=g.device)
        check(_lib.load().mmg_eos_pool_bwd(_p(g.contiguous()), _p(idx), n, seq, H, _p(dh), _stream()),
              "mmg_eos_pool_bwd")
        return dh, None


def eos_pool(hidden: torch.Tensor, attention_mask: torch.Tensor, return_index: bool = False):
    """``hidden[arange(n), attention_mask.sum(-1) - 1]`` (mmgclip_model.py:108-111) in one launch that touches only the
    pooled rows.  An all-zero mask row wraps to the last position, as Python's ``-1`` index does in the reference."""
    out, idx = _EosPoolFn.apply(hidden, attention_mask)
    return (out, idx) if return_index else out
