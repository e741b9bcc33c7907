"""Row-sharded InfoNCE step: one process per GPU, NCCL over NVLink.

The reference is single-device (``main_mlp.py:14-18``); its only multi-GPU construct gathers encoder
outputs and evaluates the loss over the full batch (``main_3dident.py:373,480-492``).  This module keeps
exactly those semantics -- every anchor sees ALL negatives of the global batch -- while sharding the work:

  rank r owns anchors/positives [r*B/W, (r+1)*B/W)
  fwd : a_r = f(g(z1_r)), b_r = f(g(z2_r))          encoder on the local shard
        z_all   = all_gather(a_r)                   (B_global x d, 0.25 .. 1.3 MB: latency-bound)
        loss_i, lse, pos, rowstat = lpnce_fwd(a_r, b_r, z_all)   local rows x all columns
        rowstat_all = all_gather(rowstat)           (B_global x 2 floats: row max + log2 sum)
        loss    = all_reduce(sum_i loss_i) / B_global
  bwd : g_a, g_b = lpnce_bwd_sharded(...)           anchor-role + column-role + positive terms for the
                                                    LOCAL rows, already scaled by 1/B_global: no
                                                    reduce-scatter of a B x d gradient is needed
        encoder backward on the local shard, then ONE all_reduce(SUM) over the flattened parameter grads.

The kernels are reached through ``ops`` (default: the CUDA C-ABI); tests inject an oracle-backed ``ops`` to
exercise this host logic on CPU with the gloo backend.
"""
import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib


# ---- kernel access (CUDA) ---------------------------------------------------------------------------------
def local_forward(z1_local, z2_local, z_all, p, tau, alpha, include_pos) -> Tuple[torch.Tensor, ...]:
    """(loss_i, lse, pos, rowstat[B,2]) of the local anchors against all gathered rows (clica_lpnce_fwd)."""
    from . import functional as F
    lib = _lib.load()
    z1, z2, za = F._as_rows(z1_local, "z1_local"), F._as_rows(z2_local, "z2_local"), F._as_rows(z_all, "z_all")
    B, d = z1.shape
    M = za.shape[0]
    dev = z1.device
    with torch.cuda.device(dev):
        out = torch.empty(5 * B + 3, dtype=torch.float32, device=dev)
        ws = F._workspace(lib.clica_lpnce_workspace_bytes(B, M, d), dev, "lpnce")
        rc = lib.clica_lpnce_fwd(z1.data_ptr(), F._ld(z1), z2.data_ptr(), F._ld(z2), za.data_ptr(), F._ld(za),
                                 B, M, d, float(p), float(tau), float(alpha), int(include_pos), 1,
                                 out[2 * B:].data_ptr(), out[3 * B:].data_ptr(), out[4 * B:].data_ptr(),
                                 out.data_ptr(), out[5 * B:].data_ptr(), ws.data_ptr(), ws.numel(), F._stream_ptr(dev))
        _lib.check(rc, "clica_lpnce_fwd")
    return out[2 * B:3 * B], out[3 * B:4 * B], out[4 * B:5 * B], out[:2 * B].view(B, 2)


def local_backward(z1_local, z2_local, z_all, rowstat_all, pos_local, row0, p, tau, alpha, include_pos,
                   g_scale: Optional[torch.Tensor] = None):
    """Gradient of the GLOBAL mean loss w.r.t. the local anchors / positives (clica_lpnce_bwd_sharded)."""
    from . import functional as F
    lib = _lib.load()
    z1, z2, za = F._as_rows(z1_local, "z1_local"), F._as_rows(z2_local, "z2_local"), F._as_rows(z_all, "z_all")
    B, d = z1.shape
    M = za.shape[0]
    dev = z1.device
    rowstat_all = rowstat_all.contiguous()
    pos_local = pos_local.contiguous()
    with torch.cuda.device(dev):
        g1 = torch.empty((B, d), dtype=torch.float32, device=dev)
        g2 = torch.empty((B, d), dtype=torch.float32, device=dev)
        if g_scale is not None:
            g_scale = g_scale.to(device=dev, dtype=torch.float32).contiguous()
        ws = F._workspace(lib.clica_lpnce_bwd_sharded_workspace_bytes(B, M, d), dev, "lpnce_bwd")
        rc = lib.clica_lpnce_bwd_sharded(z1.data_ptr(), F._ld(z1), z2.data_ptr(), F._ld(z2), za.data_ptr(), F._ld(za),
                                         rowstat_all.data_ptr(), pos_local.data_ptr(), B, M, d, int(row0), float(p),
                                         float(tau), float(alpha), int(include_pos),
                                         None if g_scale is None else g_scale.data_ptr(),
                                         g1.data_ptr(), d, g2.data_ptr(), d, ws.data_ptr(), ws.numel(),
                                         F._stream_ptr(dev))
        _lib.check(rc, "clica_lpnce_bwd_sharded")
    return g1, g2


class CudaOps:
    local_forward = staticmethod(local_forward)
    local_backward = staticmethod(local_backward)


# ---- collectives ------------------------------------------------------------------------------------------
def _all_gather_rows(t: torch.Tensor, group) -> torch.Tensor:
    world = dist.get_world_size(group)
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out


class _ShardedInfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a_local, b_local, p, tau, alpha, include_pos, group, ops):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        a_det, b_det = a_local.detach(), b_local.detach()
        z_all = _all_gather_rows(a_det, group)
        loss_i, lse, pos, rowstat = ops.local_forward(a_det, b_det, z_all, p, tau, alpha, include_pos)
        rowstat_all = _all_gather_rows(rowstat, group)
        stats = torch.stack([loss_i.sum(), pos.sum() / tau, lse.sum()])
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
        stats = stats / z_all.shape[0]
        ctx.save_for_backward(a_det, b_det, z_all, rowstat_all, pos)
        ctx.cfg = (p, tau, alpha, include_pos, rank * a_local.shape[0], ops)
        parts = stats[1:].clone()
        ctx.mark_non_differentiable(loss_i, parts)
        return stats[0], loss_i, parts

    @staticmethod
    def backward(ctx, g_mean, _g_li, _g_parts):
        a, b, z_all, rowstat_all, pos = ctx.saved_tensors
        p, tau, alpha, include_pos, row0, ops = ctx.cfg
        g1, g2 = ops.local_backward(a, b, z_all, rowstat_all, pos, row0, p, tau, alpha, include_pos, g_mean)
        return g1, g2, None, None, None, None, None, None


def sharded_lp_infonce(a_local, b_local, p, tau=1.0, alpha=0.5, include_pos=True, group=None, ops=CudaOps):
    """Global-batch Lp-InfoNCE from local shards. Returns (global mean loss, local per-item loss,
    tensor([global pos_mean, global neg_mean])).  Equal shard sizes on every rank are required."""
    if float(p) < 1.0:
        raise ValueError("sharded_lp_infonce: p < 1 (losses.py:433-442) is not implemented by the sharded loss")
    if float(tau) <= 0.0:
        raise ValueError("sharded_lp_infonce: tau must be > 0")
    if a_local.dim() != 2 or a_local.shape != b_local.shape:
        raise ValueError(f"sharded_lp_infonce: expected two [B_local, d] tensors, got {tuple(a_local.shape)}, {tuple(b_local.shape)}")
    if ops is CudaOps and a_local.shape[1] > 320:
        raise ValueError(f"sharded_lp_infonce: feature width {a_local.shape[1]} > 320 is outside the CUDA kernels' domain")
    return _ShardedInfoNCE.apply(a_local, b_local, float(p), float(tau), float(alpha), bool(include_pos), group, ops)


def allreduce_grads(params: List[torch.nn.Parameter], group=None) -> None:
    """SUM the parameter gradients over ranks in one flattened bucket (NVLS in-switch reduce on NVSwitch).

    The sharded loss already scales local gradients by 1/B_global, so the sum -- not the mean -- equals the
    single-device full-batch gradient."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


class overlapped_grad_allreduce:
    """Context manager for ``loss.backward()``: the CUDA encoder's backward all-reduces (SUM) its parameter gradients
    bucket by bucket as soon as each bucket is complete -- on NCCL's own stream, overlapping the remaining backward
    GEMMs (``functional.grad_sync``).  On exit, gradients that did not pass through that hook (parameters outside the
    fused Linear+LeakyReLU stack, or an encoder that ran on torch's own kernels) are reduced in one flattened bucket."""

    def __init__(self, params, group=None):
        from . import functional as F
        self.params, self.group = list(params), group
        self.sync = F.grad_sync(lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True))

    def __enter__(self):
        # SMs the persistent tcgen05 GEMMs of the backward leave free for NCCL's reduction kernels, which otherwise only
        # get an SM when a GEMM CTA retires (CLICA_ALLREDUCE_SM_RESERVE, default 0: measured in profiles/r2_scaling.md)
        self.reserve = int(os.environ.get("CLICA_ALLREDUCE_SM_RESERVE", "0"))
        if self.reserve > 0:
            _lib.check(_lib.load().clica_tc_set_sm_reserve(self.reserve), "clica_tc_set_sm_reserve")
        self.sync.__enter__()
        return self

    def __exit__(self, exc_type, exc, tb):
        self.sync.__exit__(exc_type, exc, tb)
        if self.reserve > 0:
            _lib.load().clica_tc_set_sm_reserve(0)
        if exc_type is None:
            rest = [p for p in self.params if id(p) not in self.sync.covered]
            allreduce_grads(rest, self.group)
        return False


def sharded_train_step(f, g, optimizer, z1_local, z2_local, p, tau=1.0, alpha=0.5, group=None, ops=CudaOps,
                       z12_local=None, overlap=True):
    """The body of ``main_mlp.py:258-285`` (unsupervised branch) on one rank's shard of the global batch.

    Returns (global mean loss tensor, tensor([pos_mean, neg_mean])) -- 0-dim / 2-element device tensors; the
    caller decides when to ``.item()`` them."""
    optimizer.zero_grad()
    if z12_local is not None:
        # anchors and positives pre-concatenated by the caller ([2B, n]): one encoder pass over 2B rows
        ab = f(g(z12_local))
        B = z12_local.shape[0] // 2
        a, b = ab[:B], ab[B:]
    else:
        a = f(g(z1_local))
        b = f(g(z2_local))
    loss, _, parts = sharded_lp_infonce(a, b, p, tau, alpha, True, group, ops)
    if overlap:
        with overlapped_grad_allreduce(f.parameters(), group):     # the encoder backward reduces its own gradient buckets
            loss.backward()
    else:
        loss.backward()
        allreduce_grads([prm for prm in f.parameters()], group)
    optimizer.step()
    return loss.detach(), parts
