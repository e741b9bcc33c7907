"""CPU (-m "not gpu"): the row-sharded multi-GPU step's HOST logic on two gloo ranks.

The CUDA kernels are replaced by an oracle-backed ``ops`` object (numpy float64 closed forms of SURVEY.md P1/P2),
so this exercises exactly what ``clica_b200.sharded`` adds on top of the kernels: shard ownership, the all-gather
of encoder outputs and row statistics, the 1/B_global gradient scaling, and the SUM all-reduce of parameter
gradients -- against the single-process full-batch result of the reference formulation (oracle.torch_port).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

LN2 = float(np.log(2.0))


class OracleOps:
    """Drop-in for sharded.CudaOps on CPU tensors (test infrastructure)."""

    @staticmethod
    def _dist(a, b, p):
        return (np.abs(a[:, None, :] - b[None, :, :]) ** p).sum(-1)

    @staticmethod
    def local_forward(z1, z2, z_all, p, tau, alpha, include_pos):
        a, b, n = (t.detach().double().numpy() for t in (z1, z2, z_all))
        D = OracleOps._dist(a, n, p)
        pos = (np.abs(a - b) ** p).sum(-1)
        logits = -D / tau
        if include_pos:
            logits = np.concatenate([logits, (-pos / tau)[:, None]], axis=1)
        mx = logits.max(1)
        lse = mx + np.log(np.exp(logits - mx[:, None]).sum(1))
        shown = lse if include_pos else lse - np.log(n.shape[0])
        loss_i = 2.0 * (alpha * pos / tau + (1 - alpha) * shown)
        rowstat = np.stack([lse / LN2, np.zeros_like(lse)], axis=1)      # (m2, ls) with m2 = log2-domain lse, ls = 0
        f = lambda x: torch.tensor(x, dtype=z1.dtype)
        return f(loss_i), f(shown), f(pos), f(rowstat)

    @staticmethod
    def local_backward(z1, z2, z_all, rowstat_all, pos, row0, p, tau, alpha, include_pos, g_scale=None):
        a, b, n = (t.detach().double().numpy() for t in (z1, z2, z_all))
        M, B = n.shape[0], a.shape[0]
        g = 1.0 if g_scale is None else float(g_scale)
        lse_all = (rowstat_all.double().numpy()[:, 0] + rowstat_all.double().numpy()[:, 1]) * LN2
        gl = g / M
        dG = lambda t: p * np.sign(t) * np.abs(t) ** (p - 1)
        # anchor role: local anchors vs all rows
        t = a[:, None, :] - n[None, :, :]
        W = np.exp(-(np.abs(t) ** p).sum(-1) / tau - lse_all[row0:row0 + B, None])
        g1 = -(2 * gl * (1 - alpha) / tau) * (W[:, :, None] * dG(t)).sum(1)
        # column role: every global anchor i vs local row k as a negative
        t2 = n[:, None, :] - a[None, :, :]
        W2 = np.exp(-(np.abs(t2) ** p).sum(-1) / tau - lse_all[:, None])
        g1 += (2 * gl * (1 - alpha) / tau) * (W2[:, :, None] * dG(t2)).sum(0)
        # positive pair
        wpos = np.exp(-pos.double().numpy() / tau - lse_all[row0:row0 + B]) if include_pos else 0.0
        cpos = 2 * gl * (alpha - (1 - alpha) * wpos) / tau
        gp = cpos[:, None] * dG(a - b)
        f = lambda x: torch.tensor(x, dtype=z1.dtype)
        return f(g1 + gp), f(-gp)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, p, result_q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from clica_b200 import sharded
    from oracle import torch_port as tp
    n, B = 5, 48
    torch.manual_seed(3)
    f = tp.build_encoder(n).double()
    g = tp.build_mixing(n, 3, seed=1).double()
    z1, z2 = tp.synth_latents(B, n, "real", c_param=0.3, seed=9, dtype=torch.float64)
    z1, z2 = z1 * 3, z2 * 3
    Bl = B // world
    sl = slice(rank * Bl, (rank + 1) * Bl)
    a, b = f(g(z1[sl])), f(g(z2[sl]))
    loss, loss_i, parts = sharded.sharded_lp_infonce(a, b, p, tau=0.8, alpha=0.5, include_pos=True, ops=OracleOps)
    loss.backward()
    sharded.allreduce_grads(list(f.parameters()))
    if rank == 0:
        result_q.put((loss.item(), [prm.grad.numpy().copy() for prm in f.parameters()], parts.detach().numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("p", [2, 3])
def test_two_rank_sharded_step_equals_full_batch(p):
    from oracle import torch_port as tp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, p, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    loss_sh, grads_sh, parts_sh = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0

    # single-process full-batch reference (the reference's formulation, z3 = roll(z1_rec))
    n, B = 5, 48
    torch.manual_seed(3)
    f = tp.build_encoder(n).double()
    g = tp.build_mixing(n, 3, seed=1).double()
    z1, z2 = tp.synth_latents(B, n, "real", c_param=0.3, seed=9, dtype=torch.float64)
    z1, z2 = z1 * 3, z2 * 3
    a, b = f(g(z1)), f(g(z2))
    total, _, parts = tp.lp_infonce(a, b, torch.roll(a, 1, 0), p, tau=0.8, alpha=0.5, compat=True)
    total.backward()
    assert abs(loss_sh - total.item()) <= 1e-10 * max(1.0, abs(total.item()))
    assert abs(parts_sh[0] - parts[0].item()) <= 1e-10 and abs(parts_sh[1] - parts[1].item()) <= 1e-10
    gscale = max(float(prm.grad.abs().max()) for prm in f.parameters())   # (the last bias' gradient is analytically 0)
    for gs, prm in zip(grads_sh, f.parameters()):
        assert np.abs(gs - prm.grad.numpy()).max() <= 1e-9 * gscale
