"""Host logic of the row-sharded InfoNCE (mmgclip_b200/distributed.py) with world_size 2 over gloo on the CPU.

The CUDA kernels are replaced by a NumPy float64 statement of the same four local steps (this is the checker standing
in for the kernels, never a product path); what is under test is the sharding, the collectives and the offsets:
sharded loss/gradients on 2 ranks == closed form on the whole batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpyKernels:
    @staticmethod
    def operand(t, prec):
        return t.detach().double().contiguous()

    @staticmethod
    def forward(a, b_all, s, off, prec):
        cos = a.numpy() @ b_all.numpy().T
        sv = float(s)
        e = np.exp(sv * cos - sv)
        rows = a.shape[0]
        diag = sv * cos[np.arange(rows), off + np.arange(rows)]
        return torch.from_numpy(e.sum(1)), torch.from_numpy(e.sum(0)), torch.from_numpy(diag)

    @staticmethod
    def row_part(rowsum, diag, out=None):
        return (torch.log(rowsum) - 2 * diag).sum().reshape(1)

    @staticmethod
    def loss_cols(colsum, s, row_part, inv_two_b):
        sv = float(s)
        return (inv_two_b * (row_part.sum() + torch.log(colsum).sum() + 2 * colsum.numel() * sv)).reshape(())

    @staticmethod
    def backward(a, b_all, s, rowsum, colsum, grad_loss, inv_two_b, off, prec, a32, b32, diag, need_dscale=True):
        sv, gl = float(s), float(grad_loss)
        cos = a.numpy() @ b_all.numpy().T
        e = np.exp(sv * cos - sv)
        coef = sv * gl * inv_two_b
        g = e * (coef / rowsum.numpy()[:, None] + coef / colsum.numpy()[None, :])
        rows = a.shape[0]
        g[np.arange(rows), off + np.arange(rows)] -= 2 * coef
        return (torch.from_numpy(g @ b_all.numpy()), torch.from_numpy(g.T @ a.numpy()),
                torch.tensor(float((g * cos).sum()), dtype=torch.float64))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mmgclip_b200 import distributed as D
    rng = np.random.RandomState(5)
    n, d = 12, 16
    a = rng.standard_normal((n, d)); a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = rng.standard_normal((n, d)); b /= np.linalg.norm(b, axis=1, keepdims=True)
    bl = n // world
    al = torch.from_numpy(a[rank * bl:(rank + 1) * bl]).requires_grad_()
    blt = torch.from_numpy(b[rank * bl:(rank + 1) * bl]).requires_grad_()
    s = torch.tensor(float(np.float32(1 / 0.07)), dtype=torch.float64, requires_grad=True)  # the op carries s as fp32
    loss = D._ShardedInfoNCEFn.apply(al, blt, s, None, "fp32", NumpyKernels)
    (loss * 3.0).backward()
    lin = torch.nn.Linear(4, 2).double()
    torch.manual_seed(rank)
    lin.weight.grad = torch.full_like(lin.weight, float(rank + 1))
    lin.bias.grad = torch.full_like(lin.bias, float(10 * (rank + 1)))
    D.allreduce_gradients(lin, group=None)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), loss=loss.detach().numpy(), da=al.grad.numpy(), db=blt.grad.numpy(),
             ds=s.grad.numpy(), wg=lin.weight.grad.numpy(), bg=lin.bias.grad.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_infonce_two_ranks_equals_closed_form(tmp_path):
    from oracle import clip_oracle as oc
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.RandomState(5)
    n, d = 12, 16
    a = rng.standard_normal((n, d)); a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = rng.standard_normal((n, d)); b /= np.linalg.norm(b, axis=1, keepdims=True)
    ref = oc.closed_form_info_nce(a, b, float(np.float32(1 / 0.07)))
    bl = n // world
    for r in range(world):
        out = np.load(os.path.join(str(tmp_path), f"r{r}.npz"))
        assert abs(float(out["loss"]) - ref["loss"]) < 1e-12           # same global loss on every rank
        assert np.allclose(out["da"], 3.0 * ref["da"][r * bl:(r + 1) * bl], rtol=1e-10, atol=1e-13)
        assert np.allclose(out["db"], 3.0 * ref["db"][r * bl:(r + 1) * bl], rtol=1e-10, atol=1e-13)
        assert abs(float(out["ds"]) - 3.0 * ref["ds"]) < 1e-10
        assert np.allclose(out["wg"], 3.0) and np.allclose(out["bg"], 30.0)  # gradients are summed across ranks


def _worker_two_losses(rank, world, port, out_dir):
    """The column-side embeddings feed TWO sharded losses (MMGCLIPLoss: CLIP term + text<->text term, losses.py:73-91) and a
    third, ordinary consumer: the gradients of all three must add up -- the situation in which the NCCL-fallback path's
    deferred reduce-scatter used to race with autograd's accumulation (every column gradient now has one producer, the
    private identity node of sharded_info_nce)."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mmgclip_b200 import distributed as D
    rng = np.random.RandomState(11)
    n, d = 12, 16
    unit = lambda x: x / np.linalg.norm(x, axis=1, keepdims=True)  # noqa: E731
    a, b, c = (unit(rng.standard_normal((n, d))) for _ in range(3))
    bl = n // world
    sl = slice(rank * bl, (rank + 1) * bl)
    al, blt, cl = (torch.from_numpy(x[sl]).requires_grad_() for x in (a, b, c))
    s = float(np.float32(1 / 0.07))
    st = torch.tensor(s, dtype=torch.float64)
    loss = (D.sharded_info_nce(al, blt, st, prec="fp32", _kernels=NumpyKernels)
            + 0.5 * D.sharded_info_nce(cl, blt, st, prec="fp32", _kernels=NumpyKernels) + 0.25 * (blt * blt).sum())
    loss.backward()
    np.savez(os.path.join(out_dir, f"t{rank}.npz"), loss=loss.detach().numpy(), da=al.grad.numpy(), db=blt.grad.numpy(),
             dc=cl.grad.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_column_side_feeding_two_sharded_losses(tmp_path):
    from oracle import clip_oracle as oc
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_two_losses, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.RandomState(11)
    n, d = 12, 16
    unit = lambda x: x / np.linalg.norm(x, axis=1, keepdims=True)  # noqa: E731
    a, b, c = (unit(rng.standard_normal((n, d))) for _ in range(3))
    s = float(np.float32(1 / 0.07))
    r1, r2 = oc.closed_form_info_nce(a, b, s), oc.closed_form_info_nce(c, b, s)
    bl = n // world
    for r in range(world):
        out = np.load(os.path.join(str(tmp_path), f"t{r}.npz"))
        sl = slice(r * bl, (r + 1) * bl)
        want_loss = r1["loss"] + 0.5 * r2["loss"] + 0.25 * float((b[sl] ** 2).sum())
        assert abs(float(out["loss"]) - want_loss) < 1e-12
        assert np.allclose(out["da"], r1["da"][sl], rtol=1e-10, atol=1e-13)
        assert np.allclose(out["dc"], 0.5 * r2["da"][sl], rtol=1e-10, atol=1e-13)
        assert np.allclose(out["db"], r1["db"][sl] + 0.5 * r2["db"][sl] + 0.5 * b[sl], rtol=1e-10, atol=1e-13)


def test_single_process_path_needs_no_process_group():
    from mmgclip_b200 import distributed as D
    x = torch.randn(4, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        D.sharded_info_nce(x, x, 10.0)  # falls through to the single-GPU operator, which refuses CPU tensors
    D.allreduce_gradients(torch.nn.Linear(2, 2))  # no-op without a process group
