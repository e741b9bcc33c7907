"""torch.autograd.Function wrappers around the C-ABI kernels (device memory, streams: PyTorch; math: CUDA).

* :func:`lp_infonce`     <- reference ``losses.py:443-477`` forward + autograd backward
* :func:`mlp_forward`    <- reference ``encoders.py:36-58`` (Linear+LeakyReLU stack) forward + backward
* :func:`adam_step`      <- ``torch.optim.Adam.step`` for a parameter list (main_mlp.py:283)

All tensors must be CUDA fp32; anything else raises.  Nothing here synchronises the host.
"""
import ctypes
import weakref
from typing import List, Optional, Sequence

import torch

from . import _lib

_vp = ctypes.c_void_p

# ---- per-(device, stream) scratch ---------------------------------------------------------------------
_workspaces = {}
_retired = []          # outgrown workspaces, kept alive (see _workspace)

# ---- gradient synchronisation hook of the encoder backward (multi-GPU) ----------------------------------
_grad_sync = None
# Gradient buckets of the multi-GPU step (CLICA_GRAD_BUCKET_MB overrides, read when a backward runs).  Measured on 8 x B200
# at BASELINE config 3 (profiles/r2_n8_allreduce_tuning.md): ONE bucket -- the whole backward as one chained launch, then
# one all-reduce -- beats 2 MB / 12 MB buckets (1.58 vs 1.76 / 1.66 ms per step): NCCL's CTAs cannot run beside the
# persistent one-CTA-per-SM GEMM grid, so "overlapped" buckets only break the chain without hiding the transfer.
GRAD_BUCKET_BYTES = 1 << 40


class grad_sync:
    """Context manager: while active, ``_MLP.backward`` issues the encoder backward bucket by bucket (layers from the
    output side; a bucket closes once it holds >= ``GRAD_BUCKET_BYTES`` of gradients) and calls ``fn(flat_slice)``
    on each finished bucket -- a contiguous fp32 slice holding that bucket's dW / db.  ``fn`` may return an object
    with ``.wait()`` (e.g. ``dist.all_reduce(..., async_op=True)``): all of them are waited for (stream-wise) before
    the backward returns, so the reduction of the last layers' gradients overlaps the GEMMs of the earlier ones."""

    def __init__(self, fn):
        self.fn = fn
        self.covered = set()       # id() of every parameter whose gradient went through `fn`

    def __call__(self, flat_slice):
        return self.fn(flat_slice)

    def __enter__(self):
        global _grad_sync
        self.prev, _grad_sync = _grad_sync, self
        return self

    def __exit__(self, *exc):
        global _grad_sync
        _grad_sync = self.prev
        return False



def _workspace(nbytes: int, device: torch.device, tag: str) -> torch.Tensor:
    """A cached byte buffer, private to (device, current stream, tag); grows geometrically.

    Stream-ordered reuse is safe because every kernel that touches it is launched on that stream."""
    stream = torch.cuda.current_stream(device).cuda_stream
    key = (device.index, stream, tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            # a CUDA graph recorded on this stream may have baked the old pointer in (GraphedTrainStep); PyTorch hands
            # out stream handles from a small pool, so an unrelated later user of the "same" stream must not free it
            _retired.append(buf)
        # zero-filled: the loss kernels keep self-resetting arrival counters at the head of their workspace
        buf = torch.zeros(max(int(nbytes * 1.25), 1 << 16), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _as_rows(t: torch.Tensor, what: str) -> torch.Tensor:
    """fp32 CUDA matrix with unit column stride (row stride may exceed the width: strided views are fine)."""
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (the CUDA path has no CPU fallback)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{what}: expected float32, got {t.dtype}")
    if t.dim() != 2:
        raise RuntimeError(f"{what}: expected a 2-D tensor, got shape {tuple(t.shape)}")
    if t.shape[0] > 1 and (t.stride(1) != 1 or t.stride(0) < t.shape[1]):
        t = t.contiguous()
    elif t.shape[0] <= 1 and t.stride(1) != 1:
        t = t.contiguous()
    return t


def _ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


# ---- fused Lp-InfoNCE loss ----------------------------------------------------------------------------
class _LpInfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z1, z2, z3, p, tau, alpha, include_pos, symmetric=False):
        lib = _lib.load()
        z1, z2, z3 = _as_rows(z1, "z1_rec"), _as_rows(z2, "z2_con_z1_rec"), _as_rows(z3, "z3_rec")
        B, d = z1.shape
        M = z3.shape[0]
        if z2.shape != (B, d) or z3.shape[1] != d:
            raise RuntimeError(f"shape mismatch: z1 {tuple(z1.shape)}, z2 {tuple(z2.shape)}, z3 {tuple(z3.shape)}")
        dev = z1.device
        with torch.cuda.device(dev):
            # one allocation: rowstat[B][2] first (8-byte aligned), then loss_i, lse, pos, scalars[3]
            out = torch.empty(5 * B + 3, dtype=torch.float32, device=dev)
            rowstat, loss_i, lse, pos, scal = out[:2 * B], out[2 * B:3 * B], out[3 * B:4 * B], out[4 * B:5 * B], out[5 * B:]
            nbytes = lib.clica_lpnce_workspace_bytes(B, M, d)
            ws = _workspace(nbytes, dev, "lpnce")
            rc = lib.clica_lpnce_fwd(z1.data_ptr(), _ld(z1), z2.data_ptr(), _ld(z2), z3.data_ptr(), _ld(z3),
                                     B, M, d, float(p), float(tau), float(alpha), int(include_pos), 1,
                                     loss_i.data_ptr(), lse.data_ptr(), pos.data_ptr(), rowstat.data_ptr(),
                                     scal.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
            _lib.check(rc, "clica_lpnce_fwd")
        ctx.save_for_backward(z1, z2, z3, rowstat, pos)
        ctx.cfg = (float(p), float(tau), float(alpha), int(include_pos))
        ctx.symmetric = bool(symmetric) and z3.shape[0] == B
        ctx.set_materialize_grads(False)
        mean, pos_mean, neg_mean = scal[0], scal[1], scal[2]
        ctx.mark_non_differentiable(pos_mean, neg_mean)
        return mean, loss_i, pos_mean, neg_mean

    @staticmethod
    def backward(ctx, g_mean, g_loss_i, _g_pos, _g_neg):
        lib = _lib.load()
        z1, z2, z3, rowstat, pos = ctx.saved_tensors
        p, tau, alpha, include_pos = ctx.cfg
        B, d = z1.shape
        M = z3.shape[0]
        dev = z1.device
        need1, need2, need3 = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        if not (need1 or need2 or need3) or (g_mean is None and g_loss_i is None):
            return (None,) * 8
        if ctx.symmetric and g_loss_i is None:
            # z3 is a row permutation of z1 (torch.roll, main_mlp.py:272): the pair distance is symmetric, so the
            # anchor-role and the column-role gradients come out of ONE merged pass (clica_lpnce_bwd_sharded with
            # the whole batch as the local shard); the roll's own backward then receives no gradient for z3.
            with torch.cuda.device(dev):
                g1 = torch.empty((B, d), dtype=torch.float32, device=dev)
                g2 = torch.empty((B, d), dtype=torch.float32, device=dev)
                g_mean = g_mean.to(device=dev, dtype=torch.float32).contiguous()
                nbytes = lib.clica_lpnce_bwd_sharded_workspace_bytes(B, B, d)
                ws = _workspace(nbytes, dev, "lpnce_bwd")
                rc = lib.clica_lpnce_bwd_sharded(z1.data_ptr(), _ld(z1), z2.data_ptr(), _ld(z2), z1.data_ptr(), _ld(z1),
                                                 rowstat.data_ptr(), pos.data_ptr(), B, B, d, 0, p, tau, alpha,
                                                 include_pos, g_mean.data_ptr(), g1.data_ptr(), d, g2.data_ptr(), d,
                                                 ws.data_ptr(), ws.numel(), _stream_ptr(dev))
                _lib.check(rc, "clica_lpnce_bwd_sharded")
            return g1 if need1 else None, g2 if need2 else None, None, None, None, None, None, None
        with torch.cuda.device(dev):
            g1 = torch.empty_like(z1, memory_format=torch.contiguous_format) if need1 else None
            g2 = torch.empty_like(z2, memory_format=torch.contiguous_format) if need2 else None
            g3 = torch.empty_like(z3, memory_format=torch.contiguous_format) if need3 else None
            if g_mean is not None:
                g_mean = g_mean.to(device=dev, dtype=torch.float32).contiguous()
            if g_loss_i is not None:
                g_loss_i = g_loss_i.to(device=dev, dtype=torch.float32).contiguous()
            nbytes = lib.clica_lpnce_bwd_workspace_bytes(B, M, d)
            ws = _workspace(nbytes, dev, "lpnce_bwd")
            rc = lib.clica_lpnce_bwd(z1.data_ptr(), _ld(z1), z2.data_ptr(), _ld(z2), z3.data_ptr(), _ld(z3),
                                     B, M, d, p, tau, alpha, include_pos, 1, rowstat.data_ptr(), pos.data_ptr(),
                                     _ptr(g_mean), _ptr(g_loss_i), _ptr(g1), d, _ptr(g2), d, _ptr(g3), d,
                                     ws.data_ptr(), ws.numel(), _stream_ptr(dev))
            _lib.check(rc, "clica_lpnce_bwd")
        return g1, g2, g3, None, None, None, None, None


class _LpInfoNCEPairs(torch.autograd.Function):
    """The training step's own form of the loss: ``ab = [z1_rec; z2_rec]`` as ONE [2B, d] encoder output and the
    negatives ``z3_rec = roll(z1_rec, 1, 0)`` of main_mlp.py:272 taken as what they are for a sum over all of them -- the
    rows of ``z1_rec`` in another order.  No roll kernel, no slice / concatenate nodes in the autograd graph: the forward
    streams ``z1_rec`` itself, the merged backward writes both gradient halves into one [2B, d] tensor."""

    @staticmethod
    def forward(ctx, ab, p, tau, alpha, include_pos):
        lib = _lib.load()
        ab = _as_rows(ab, "encoder output [z1_rec; z2_rec]").contiguous()
        if ab.shape[0] % 2:
            raise RuntimeError("lp_infonce_pairs: expected [2B, d] rows (anchors first, then positives)")
        B, d = ab.shape[0] // 2, ab.shape[1]
        z1, z2 = ab[:B], ab[B:]
        dev = ab.device
        with torch.cuda.device(dev):
            out = torch.empty(5 * B + 3, dtype=torch.float32, device=dev)
            rowstat, loss_i, lse, pos, scal = out[:2 * B], out[2 * B:3 * B], out[3 * B:4 * B], out[4 * B:5 * B], out[5 * B:]
            ws = _workspace(lib.clica_lpnce_workspace_bytes(B, B, d), dev, "lpnce")
            rc = lib.clica_lpnce_fwd(z1.data_ptr(), d, z2.data_ptr(), d, z1.data_ptr(), d, B, B, d, float(p), float(tau),
                                     float(alpha), int(include_pos), 1, loss_i.data_ptr(), lse.data_ptr(), pos.data_ptr(),
                                     rowstat.data_ptr(), scal.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
            _lib.check(rc, "clica_lpnce_fwd")
        ctx.save_for_backward(ab, rowstat, pos)
        ctx.cfg = (float(p), float(tau), float(alpha), int(include_pos))
        ctx.set_materialize_grads(False)
        mean, pos_mean, neg_mean = scal[0], scal[1], scal[2]
        ctx.mark_non_differentiable(loss_i, pos_mean, neg_mean)
        return mean, loss_i, pos_mean, neg_mean

    @staticmethod
    def backward(ctx, g_mean, _g_li, _g_pos, _g_neg):
        if g_mean is None or not ctx.needs_input_grad[0]:
            return (None,) * 5
        lib = _lib.load()
        ab, rowstat, pos = ctx.saved_tensors
        p, tau, alpha, include_pos = ctx.cfg
        B, d = ab.shape[0] // 2, ab.shape[1]
        dev = ab.device
        with torch.cuda.device(dev):
            g = torch.empty_like(ab)
            g_mean = g_mean.to(device=dev, dtype=torch.float32).contiguous()
            ws = _workspace(lib.clica_lpnce_bwd_sharded_workspace_bytes(B, B, d), dev, "lpnce_bwd")
            z1, z2 = ab[:B], ab[B:]
            rc = lib.clica_lpnce_bwd_sharded(z1.data_ptr(), d, z2.data_ptr(), d, z1.data_ptr(), d, rowstat.data_ptr(),
                                             pos.data_ptr(), B, B, d, 0, p, tau, alpha, include_pos, g_mean.data_ptr(),
                                             g[:B].data_ptr(), d, g[B:].data_ptr(), d, ws.data_ptr(), ws.numel(),
                                             _stream_ptr(dev))
            _lib.check(rc, "clica_lpnce_bwd_sharded")
        return g, None, None, None, None


def lp_infonce_pairs(ab, p, tau=1.0, alpha=0.5, include_pos=True):
    """Fused Lp-InfoNCE (``p = 0``: dot-product similarity) of ``ab = [z1_rec; z2_rec]`` with every anchor as a negative
    (= the reference's ``z3_rec = roll(z1_rec, 1, 0)``, main_mlp.py:272).  Returns ``(mean, per_item, pos_mean,
    neg_mean)``; only ``mean`` carries grad."""
    return _LpInfoNCEPairs.apply(ab, p, tau, alpha, include_pos)


def is_row_roll_of(z3, z1) -> bool:
    """True when autograd itself says ``z3 = torch.roll(z1, k, 0)`` (the reference's negatives, main_mlp.py:272)."""
    fn = getattr(z3, "grad_fn", None)
    if fn is None or type(fn).__name__ != "RollBackward0" or z1.dim() != 2 or z3.shape != z1.shape:
        return False
    try:
        dims = tuple(int(x) for x in fn._saved_dims)
        src, out_nr = fn.next_functions[0]
    except Exception:
        return False
    if dims not in ((0,), (-2,)):
        return False
    if z1.grad_fn is not None:
        return src is z1.grad_fn and out_nr == z1.output_nr
    return getattr(src, "variable", None) is z1


def lp_infonce(z1_rec, z2_rec, z3_rec, p, tau=1.0, alpha=0.5, include_pos=True, symmetric=None):
    """Fused Lp-InfoNCE. Returns ``(mean, per_item[B], pos_mean, neg_mean)``; mean and per_item carry grad.

    ``symmetric`` (default: detected from the autograd graph) declares that ``z3_rec`` is a row permutation of
    ``z1_rec``; the backward then needs a single pass over the B x B pairs instead of two."""
    if symmetric is None:
        symmetric = is_row_roll_of(z3_rec, z1_rec)
    return _LpInfoNCE.apply(z1_rec, z2_rec, z3_rec, p, tau, alpha, include_pos, bool(symmetric))


# ---- encoder stack ------------------------------------------------------------------------------------
def _ptr_array(tensors: Sequence[Optional[torch.Tensor]]):
    arr = (_vp * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


# packed (GEMM-operand-format) copies of the encoder weights, keyed by the identity of the weight tensors (weak
# references: a new tensor that happens to reuse a freed tensor's address and version must NOT hit) and refreshed
# whenever any weight's version counter or storage moves (i.e. once per optimizer step, while the encoder runs
# four times per step)
_packed_cache = {}


def _packed_weights(lib, Ws, widths, mode, dev):
    """(device pointer (1024-byte aligned), owning buffer) of the packed weight planes; re-packed only on change."""
    key = (dev.index, int(mode), tuple(id(W) for W in Ws))
    sig = tuple((W.data_ptr(), W._version) for W in Ws)
    stream = _stream_ptr(dev)
    hit = _packed_cache.get(key)
    alive = hit is not None and all(r() is W for r, W in zip(hit[4], Ws))
    if alive and hit[0] == sig and hit[2] == stream:
        return hit[3], hit[1]
    L = len(Ws)
    cw = (ctypes.c_int * (L + 1))(*widths)
    nbytes = lib.clica_mlp_packed_weight_bytes(L, cw, int(mode))
    if alive and hit[1].numel() >= nbytes + 1024 and hit[2] == stream:
        buf = hit[1]
    else:
        buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)     # the allocator only guarantees 512 B
    ptr = (buf.data_ptr() + 1023) // 1024 * 1024
    rc = lib.clica_mlp_pack_weights(L, cw, _ptr_array(Ws), int(mode), ptr, nbytes, stream)
    _lib.check(rc, "clica_mlp_pack_weights")
    if len(_packed_cache) >= 16:                                            # drop entries whose tensors died
        for k in [k for k, v in _packed_cache.items() if any(r() is None for r in v[4])]:
            del _packed_cache[k]
        if len(_packed_cache) >= 16:
            _packed_cache.clear()
    _packed_cache[key] = (sig, buf, stream, ptr, [weakref.ref(W) for W in Ws])
    return ptr, buf


class _MLP(torch.autograd.Function):
    """y = Linear_{L-1}( LeakyReLU( ... LeakyReLU(Linear_0(x)) ) ) through clica_mlp_fwd / clica_mlp_bwd."""

    @staticmethod
    def forward(ctx, x, slope, mode, packed, *params):
        # `packed` = (pointer, owning buffer) of the packed weight planes, resolved by mlp_forward() on the caller's
        # own tensor objects; the backward reuses it (autograd guarantees the weights did not change in between)
        lib = _lib.load()
        x = _as_rows(x, "encoder input").contiguous()
        L = len(params) // 2
        Ws = [params[2 * l] for l in range(L)]
        bs = [params[2 * l + 1] for l in range(L)]
        widths = [Ws[0].shape[1]] + [W.shape[0] for W in Ws]
        if x.shape[1] != widths[0]:
            raise RuntimeError(f"encoder input has {x.shape[1]} features, first Linear expects {widths[0]}")
        for l, (W, b) in enumerate(zip(Ws, bs)):
            if not (W.is_cuda and W.dtype == torch.float32 and W.is_contiguous() and W.shape[1] == widths[l]):
                raise RuntimeError(f"layer {l}: weight must be a contiguous CUDA fp32 [{widths[l + 1]}, {widths[l]}] tensor")
            if b is None or not (b.is_cuda and b.dtype == torch.float32 and b.is_contiguous()):
                raise RuntimeError(f"layer {l}: bias must be a contiguous CUDA fp32 tensor")
        M = x.shape[0]
        dev = x.device
        with torch.cuda.device(dev):
            # hidden activations are opaque buffers in the GEMM operand format of `mode` (see include/clica.h)
            acts = [x]
            for w in widths[1:-1]:
                acts.append(torch.empty(lib.clica_mlp_act_floats(M, w, int(mode)), dtype=torch.float32, device=dev))
            acts.append(torch.empty((M, widths[-1]), dtype=torch.float32, device=dev))
            cw = (ctypes.c_int * (L + 1))(*widths)
            nbytes = lib.clica_mlp_workspace_bytes(M, L, cw, mode)
            ws = _workspace(nbytes, dev, "mlp")
            rc = lib.clica_mlp_fwd(L, cw, _ptr_array(Ws), _ptr_array(bs), _ptr_array(acts), M, float(slope),
                                   int(mode), packed[0], ws.data_ptr(), ws.numel(), _stream_ptr(dev))
            _lib.check(rc, "clica_mlp_fwd")
        ctx.save_for_backward(*acts[:-1], *Ws)
        ctx.cfg = (L, widths, float(slope), int(mode), [b is not None for b in bs])
        ctx.param_ids = [id(t) for t in params]
        ctx.packed = packed
        return acts[-1]

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        L, widths, slope, mode, _ = ctx.cfg
        saved = ctx.saved_tensors
        acts, Ws = list(saved[:L]), list(saved[L:])
        M = acts[0].shape[0]
        dev = gy.device
        gy = gy.contiguous()
        with torch.cuda.device(dev):
            # all parameter gradients live in ONE zero-initialised flat buffer (split-K / fused column sums accumulate),
            # laid out in backward order [W_{L-1}, b_{L-1}, ..., W_0, b_0] so that the layers finished first form a
            # contiguous prefix (gradient buckets of the multi-GPU step are plain slices)
            offs_w, offs_b, tot = [0] * L, [0] * L, 0
            for l in range(L - 1, -1, -1):
                offs_w[l] = tot
                tot += (Ws[l].numel() + 3) // 4 * 4            # keep every tensor 16-byte aligned
                offs_b[l] = tot
                tot += (Ws[l].shape[0] + 3) // 4 * 4
            flat = torch.zeros(tot, dtype=torch.float32, device=dev)
            dWs = [flat[offs_w[l]:offs_w[l] + Ws[l].numel()].view_as(Ws[l]) for l in range(L)]
            dbs = [flat[offs_b[l]:offs_b[l] + Ws[l].shape[0]] for l in range(L)]
            g_in = torch.empty_like(acts[0]) if ctx.needs_input_grad[0] else None
            cw = (ctypes.c_int * (L + 1))(*widths)
            nbytes = lib.clica_mlp_workspace_bytes(M, L, cw, mode)
            ws = _workspace(nbytes, dev, "mlp")
            # buckets: [(l_first, l_last, flat_begin, flat_end)]
            if _grad_sync is None:
                buckets = [(L - 1, 0, 0, tot)]
            else:
                buckets, l_first, begin = [], L - 1, 0
                import os
                bucket_bytes = int(float(os.environ.get("CLICA_GRAD_BUCKET_MB", GRAD_BUCKET_BYTES / 2 ** 20)) * 2 ** 20)
                for l in range(L - 1, -1, -1):
                    end = offs_b[l] + (Ws[l].shape[0] + 3) // 4 * 4
                    if (end - begin) * 4 >= bucket_bytes or l == 0:
                        buckets.append((l_first, l, begin, end))
                        l_first, begin = l - 1, end
            works = []
            for (l_first, l_last, begin, end) in buckets:
                rc = lib.clica_mlp_bwd_range(L, cw, _ptr_array(Ws), _ptr_array(acts), gy.data_ptr(), _ptr_array(dWs),
                                             _ptr_array(dbs), _ptr(g_in), M, slope, mode, ctx.packed[0], 1,
                                             l_first, l_last, ws.data_ptr(), ws.numel(), _stream_ptr(dev))
                _lib.check(rc, "clica_mlp_bwd_range")
                if _grad_sync is not None:
                    w = _grad_sync(flat[begin:end])
                    if w is not None:
                        works.append(w)
            for w in works:
                w.wait()
            if _grad_sync is not None:
                _grad_sync.covered.update(ctx.param_ids)
        grads = []
        for l in range(L):
            grads.append(dWs[l] if ctx.needs_input_grad[4 + 2 * l] else None)
            grads.append(dbs[l] if ctx.needs_input_grad[5 + 2 * l] else None)
        return (g_in, None, None, None, *grads)


def mlp_forward(x, weights: List[torch.Tensor], biases: List[torch.Tensor], slope: float = 0.01,
                mode: Optional[int] = None):
    """Linear+LeakyReLU(slope) stack, identity after the last Linear. Differentiable in x, weights, biases."""
    if mode is None:
        mode = _lib.gemm_mode_from_env()
    flat = []
    for W, b in zip(weights, biases):
        flat += [W, b]
    if not weights or not weights[0].is_cuda:
        raise RuntimeError("mlp_forward: expected CUDA weights (the CUDA path has no CPU fallback)")
    dev = weights[0].device
    widths = [weights[0].shape[1]] + [W.shape[0] for W in weights]
    with torch.cuda.device(dev):
        packed = _packed_weights(_lib.load(), list(weights), widths, int(mode), dev)
    return _MLP.apply(x, slope, mode, packed, *flat)


# ---- fused Adam -----------------------------------------------------------------------------------------
def adam_step(params, grads, exp_avgs, exp_avg_sqs, lr, beta1, beta2, eps, step, grad_scale=1.0):
    """One fused multi-tensor Adam update (in place) for contiguous CUDA fp32 tensors on one device."""
    lib = _lib.load()
    if not params:
        return
    dev = params[0].device
    for group in (params, grads, exp_avgs, exp_avg_sqs):
        for t in group:
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.device == dev):
                raise RuntimeError("adam_step: all tensors must be contiguous CUDA fp32 on one device")
    n = len(params)
    numel = (ctypes.c_int64 * n)(*[p.numel() for p in params])
    with torch.cuda.device(dev):
        rc = lib.clica_adam_step(n, _ptr_array(params), _ptr_array(grads), _ptr_array(exp_avgs),
                                 _ptr_array(exp_avg_sqs), numel, float(lr), float(beta1), float(beta2),
                                 float(eps), int(step), float(grad_scale), _stream_ptr(dev))
        _lib.check(rc, "clica_adam_step")


def adam_step_capturable(params, grads, exp_avgs, exp_avg_sqs, lr, beta1, beta2, eps, step_state, grad_scale=1.0,
                         pack=None):
    """Fused Adam whose step count lives on the device (``step_state``: int64[2] CUDA tensor, zero before the
    first step).  Safe to record into a CUDA graph: every replay advances the count and applies one update.

    ``pack``: optional list (one entry per parameter) of ``None`` or ``(hi_ptr, lo_ptr_or_None, cols, ld)`` -- the
    updated weight matrix is then also written in the tensor-core operand format (``packed_weight_targets``), which
    replaces the separate re-pack pass after the step (``clica_adam_step_capturable_packed``)."""
    lib = _lib.load()
    if not params:
        return
    dev = params[0].device
    for group in (params, grads, exp_avg_sqs, exp_avgs):
        for t in group:
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.device == dev):
                raise RuntimeError("adam_step_capturable: all tensors must be contiguous CUDA fp32 on one device")
    if not (step_state.is_cuda and step_state.dtype == torch.int64 and step_state.numel() >= 2 and step_state.device == dev):
        raise RuntimeError("adam_step_capturable: step_state must be an int64[2] CUDA tensor on the parameters' device")
    n = len(params)
    numel = (ctypes.c_int64 * n)(*[p.numel() for p in params])
    with torch.cuda.device(dev):
        if pack is not None and any(e is not None for e in pack):
            if len(pack) != n:
                raise RuntimeError("adam_step_capturable: `pack` needs one entry per parameter")
            hi, lo, cols, ld = (_vp * n)(), (_vp * n)(), (ctypes.c_int * n)(), (ctypes.c_int * n)()
            for i, e in enumerate(pack):
                hi[i], lo[i], cols[i], ld[i] = (None, None, 1, 1) if e is None else (e[0], e[1], int(e[2]), int(e[3]))
            rc = lib.clica_adam_step_capturable_packed(n, _ptr_array(params), _ptr_array(grads), _ptr_array(exp_avgs),
                                                       _ptr_array(exp_avg_sqs), numel, float(lr), float(beta1),
                                                       float(beta2), float(eps), step_state.data_ptr(),
                                                       float(grad_scale), hi, lo, cols, ld, _stream_ptr(dev))
            _lib.check(rc, "clica_adam_step_capturable_packed")
            return
        rc = lib.clica_adam_step_capturable(n, _ptr_array(params), _ptr_array(grads), _ptr_array(exp_avgs),
                                            _ptr_array(exp_avg_sqs), numel, float(lr), float(beta1), float(beta2),
                                            float(eps), step_state.data_ptr(), float(grad_scale), _stream_ptr(dev))
        _lib.check(rc, "clica_adam_step_capturable")


def packed_weight_targets(weights, mode, packed_ptr):
    """{weight tensor id: (hi_ptr, lo_ptr or None, cols, ld)} for the layers that live in the packed-weight buffer at
    ``packed_ptr`` (``clica_mlp_packed_weight_layout``); layers outside it (first / last at small n) are absent."""
    lib = _lib.load()
    L = len(weights)
    widths = [weights[0].shape[1]] + [W.shape[0] for W in weights]
    cw = (ctypes.c_int * (L + 1))(*widths)
    hi, lo, ld = (ctypes.c_longlong * L)(), (ctypes.c_longlong * L)(), (ctypes.c_int * L)()
    _lib.check(lib.clica_mlp_packed_weight_layout(L, cw, int(mode), hi, lo, ld), "clica_mlp_packed_weight_layout")
    out = {}
    for l, W in enumerate(weights):
        if hi[l] >= 0:
            out[id(W)] = (packed_ptr + hi[l], None if lo[l] < 0 else packed_ptr + lo[l], W.shape[1], ld[l])
    return out


def repack_weights_into(packed_ptr, weights, mode):
    """Re-pack ``weights`` into an existing packed buffer on the current stream (``clica_mlp_pack_weights``)."""
    lib = _lib.load()
    L = len(weights)
    widths = [weights[0].shape[1]] + [W.shape[0] for W in weights]
    cw = (ctypes.c_int * (L + 1))(*widths)
    nbytes = lib.clica_mlp_packed_weight_bytes(L, cw, int(mode))
    dev = weights[0].device
    with torch.cuda.device(dev):
        rc = lib.clica_mlp_pack_weights(L, cw, _ptr_array(list(weights)), int(mode), packed_ptr, nbytes, _stream_ptr(dev))
        _lib.check(rc, "clica_mlp_pack_weights")


def invalidate_packed_weights():
    """Drop every cached packed-weight buffer (the next encoder call re-packs from the live parameters)."""
    _packed_cache.clear()


# ---- frozen mixing network (see include/clica.h, clica_mixing_fwd) ------------------------------------------
def mixing_plan(g):
    """(weights, slope) when ``g`` is the frozen mixing stack of invertible_network_utils.py:87-123 -- bias-free
    square ``nn.Linear`` layers with one LeakyReLU slope in between, nothing requiring grad -- else None."""
    from torch import nn
    if not isinstance(g, nn.Sequential) or len(g) == 0:
        return None
    weights, slope, expect_linear = [], None, True
    for m in g:
        if expect_linear:
            if not isinstance(m, nn.Linear) or m.bias is not None or m.weight.requires_grad:
                return None
            if m.weight.shape[0] != m.weight.shape[1] or (weights and m.weight.shape != weights[0].shape):
                return None
            weights.append(m.weight)
        else:
            if not isinstance(m, nn.LeakyReLU):
                return None
            s = float(m.negative_slope)
            if slope is not None and s != slope:
                return None
            slope = s
        expect_linear = not expect_linear
    if expect_linear or not weights:          # must end on a Linear
        return None
    n = weights[0].shape[0]
    if len(weights) > 8 or n > 48 or len(weights) * n * n * 4 > 48 * 1024:
        return None
    return weights, (0.2 if slope is None else slope)


def mixing_forward(x, weights, slope):
    """y = W_{L-1} lrelu( ... lrelu(W_0 x)) for frozen weights; no autograd graph is recorded (forward only)."""
    lib = _lib.load()
    x = _as_rows(x, "mixing input")
    n = weights[0].shape[0]
    if x.shape[1] != n:
        raise RuntimeError(f"mixing_forward: input has {x.shape[1]} features, the mixing layers are {n} x {n}")
    for W in weights:
        if not (W.is_cuda and W.dtype == torch.float32 and W.is_contiguous() and W.device == x.device):
            raise RuntimeError("mixing_forward: weights must be contiguous CUDA fp32 tensors on the input's device")
    dev = x.device
    with torch.cuda.device(dev):
        y = torch.empty((x.shape[0], n), dtype=torch.float32, device=dev)
        rc = lib.clica_mixing_fwd(x.data_ptr(), _ld(x), _ptr_array([W.detach() for W in weights]), len(weights), n,
                                  x.shape[0], float(slope), y.data_ptr(), n, _stream_ptr(dev))
        _lib.check(rc, "clica_mixing_fwd")
    return y
