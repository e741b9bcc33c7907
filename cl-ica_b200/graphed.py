"""GraphedTrainStep: the whole InfoNCE training step recorded once into a CUDA graph and replayed per step.

The step is the unsupervised branch of the reference's ``train_step`` (``main_mlp.py:258-285``):

    zero_grad -> z1_rec = h(z1), z2_rec = h(z2) -> z3_rec = roll(z1_rec, 1, 0) -> LpSimCLRLoss -> backward -> Adam.step

with ``h = f o g`` (``main_mlp.py:313``).  At the reference's sizes the step is a chain of ~45 short kernels
(0.9 ms of device time at n = 10, B = 6144) and the eager Python/autograd dispatch of that chain costs about as
much again on the host.  Recording it once (``torch.cuda.graph``: device memory and streams are PyTorch's, every
kernel is this library's) removes the host from the loop:

* the (anchor, positive) batch is staged in ONE pinned host buffer ``[2B, n]``; the graph starts with its
  host->device copy and ends with the device->host copy of ``(loss, pos_mean, neg_mean)`` into pinned memory
  (``host_io=True``), so one replay is a complete step from host data to host scalars;
* anchors and positives go through the encoder as one 2B-row batch (identical arithmetic per row);
* Adam runs as ``FusedAdam(capturable=True)``: its step count lives on the device, so consecutive replays are
  consecutive optimizer steps.

Warm-up steps needed before capture are rolled back (parameters and optimizer state are restored), so a freshly
built object has taken zero optimizer steps.

Multi-GPU (``group`` given, one process per GPU): the recorded step is the row-sharded one of ``sharded.py`` -- the
NCCL all-gathers of the encoder outputs / row statistics and the all-reduce of the parameter gradients are captured
into the graph with the kernels (every rank records and replays the same sequence).
"""
import os
from typing import Optional, Tuple

import torch

from . import _lib
from . import functional as F
from .optim import FusedAdam


def pairs_config(criterion):
    """``(p, tau, alpha, include_pos)`` for ``functional.lp_infonce_pairs`` when ``criterion`` is one of the two fused
    losses with arguments inside the kernels' domain -- ``LpSimCLRLoss`` with ``p >= 1`` and ``pow=True``
    (losses.py:416-431), or ``SimCLRLoss`` without ``normalize`` (losses.py:172-175; p = 0 selects the dot-product form)
    -- else ``None`` (the criterion is then called the way the script calls it)."""
    name = type(criterion).__name__
    try:
        if name == "LpSimCLRLoss":
            if not getattr(criterion, "pow", True) or float(criterion.p) < 1.0:
                return None
            return (float(criterion.p), float(criterion.tau), float(criterion.alpha),
                    bool(criterion.simclr_compatibility_mode))
        if name == "SimCLRLoss":
            if getattr(criterion, "normalize", False):
                return None
            return (0.0, float(criterion.tau), float(criterion.alpha), True)
    except (AttributeError, TypeError, ValueError):
        return None
    return None


class GraphedTrainStep:
    def __init__(self, f: torch.nn.Module, g: Optional[torch.nn.Module], criterion, batch_size: int, n_in: int,
                 lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, host_io: bool = True,
                 device: Optional[torch.device] = None, warmup: int = 3, group=None):
        _lib.load()
        params = [p for p in f.parameters() if p.requires_grad]
        if not params or not params[0].is_cuda:
            raise RuntimeError("GraphedTrainStep: the encoder must live on a CUDA device (no CPU path)")
        self.device = params[0].device if device is None else torch.device(device)
        self.f, self.g, self.criterion = f, g, criterion
        self.B, self.n = int(batch_size), int(n_in)
        self.host_io = bool(host_io)
        self.params = params
        self.group = group
        self.world = 1
        if group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(group)
        if self.world > 1:
            # the row-sharded loss only exists as CUDA kernels (no delegation to the reference's torch code): refuse
            # what those kernels do not implement instead of silently training a different objective
            for attr in ("p", "tau", "alpha", "simclr_compatibility_mode"):
                if not hasattr(criterion, attr):
                    raise TypeError("GraphedTrainStep(group=...): the criterion must be an LpSimCLRLoss "
                                    f"(missing attribute {attr!r})")
            if not getattr(criterion, "pow", True):
                raise ValueError("GraphedTrainStep(group=...): LpSimCLRLoss(pow=False) is not implemented by the sharded CUDA loss")
            if float(criterion.p) < 1.0:
                raise ValueError("GraphedTrainStep(group=...): p < 1 is not implemented by the sharded CUDA loss")
        # the frozen mixing net as one fused kernel (CLICA_FUSED_MIXING=0: torch's L GEMMs + L-1 activations)
        self._mix = None
        if g is not None and os.environ.get("CLICA_FUSED_MIXING", "1") == "1":
            self._mix = F.mixing_plan(g)
        self.optimizer = FusedAdam(params, lr=lr, betas=betas, eps=eps, capturable=True)
        # the recorded Adam step also writes the hidden layers' weights in the tensor-core operand format (no re-pack pass
        # at the start of the next step); CLICA_ADAM_PACK=0: the graph re-packs the weights before every forward
        self._pack_weights, self._packed_ptr, self._packed_buf, self._pack_versions = None, None, None, None
        if os.environ.get("CLICA_ADAM_PACK", "1") == "1" and hasattr(f, "_plan"):
            plan = f._plan()
            if plan is not None and all(l.weight.requires_grad for l in plan[0]):
                self._pack_weights = [l.weight for l in plan[0]]
        self._pairs_cfg = pairs_config(criterion) if os.environ.get("CLICA_GRAPH_PAIRS", "1") == "1" else None
        dev = self.device
        self.z_dev = torch.zeros((2 * self.B, self.n), dtype=torch.float32, device=dev)
        self.out_dev = torch.zeros(3, dtype=torch.float32, device=dev)
        self.z_host = torch.zeros((2 * self.B, self.n), dtype=torch.float32).pin_memory() if host_io else None
        self.out_host = torch.zeros(3, dtype=torch.float32).pin_memory() if host_io else None
        self.stream = torch.cuda.Stream(dev)
        self.launches_per_replay = 0
        self.steps_taken = 0
        self._replayed = torch.cuda.Event()      # recorded after every replay (guards the pinned staging buffer)
        self._capture(max(int(warmup), 1))

    # the recorded body -------------------------------------------------------------------------------------
    def _body(self):
        if self.host_io:
            self.z_dev.copy_(self.z_host, non_blocking=True)
        self.optimizer.zero_grad(set_to_none=True)
        if self.g is None:
            x = self.z_dev
        elif self._mix is not None:
            x = F.mixing_forward(self.z_dev, *self._mix)       # one kernel instead of L GEMMs + L-1 activations
        else:
            x = self.g(self.z_dev)
        ab = self.f(x)
        if self.world == 1 and self._pairs_cfg is not None and ab.shape == (2 * self.B, ab.shape[-1]) and ab.shape[-1] <= 320:
            # anchors and positives as one tensor, the anchors themselves as the (rolled) negatives: one forward and one
            # backward launch with no roll / slice / concatenate kernels around them
            total, _, pos_mean, neg_mean = F.lp_infonce_pairs(ab, *self._pairs_cfg)
            total.backward()
            self.optimizer.step()
            torch.stack([total.detach(), pos_mean, neg_mean], out=self.out_dev)
            if self.host_io:
                self.out_host.copy_(self.out_dev, non_blocking=True)
            return
        a, b = ab[:self.B], ab[self.B:]
        if self.world > 1:
            from . import sharded
            c = self.criterion
            total, _, parts = sharded.sharded_lp_infonce(a, b, float(c.p), float(c.tau), float(c.alpha),
                                                         bool(c.simclr_compatibility_mode), self.group)
            with sharded.overlapped_grad_allreduce(self.params, self.group):
                total.backward()
        else:
            total, _, parts = self.criterion(None, None, None, a, b, torch.roll(a, 1, 0))
            total.backward()
        self.optimizer.step()
        torch.stack([total.detach(), parts[0].detach(), parts[1].detach()], out=self.out_dev)
        if self.host_io:
            self.out_host.copy_(self.out_dev, non_blocking=True)

    def _capture(self, warmup):
        lib = _lib.load()
        dev = self.device
        with torch.no_grad():
            saved = [p.detach().clone() for p in self.params]
            # module buffers too (e.g. BatchNorm running statistics, which the warm-up batches would pollute)
            saved_buffers = [(b, b.detach().clone()) for b in self.f.buffers()]
        cur = torch.cuda.current_stream(dev)
        self.stream.wait_stream(cur)
        with torch.cuda.device(dev), torch.cuda.stream(self.stream):
            for _ in range(warmup):          # sizes every workspace / cuBLAS handle on the capture stream
                self._body()
            self.stream.synchronize()
            F.invalidate_packed_weights()
            gemm_mode = self._gemm_mode = _lib.gemm_mode_from_env()      # the recording is tied to this numeric mode
            if self._pack_weights is not None:
                # pack once, outside the recording: the cached planes stay valid through the capture (no parameter changes
                # before the recorded Adam step), so the recorded forward contains no re-pack, and the recorded Adam step
                # keeps the planes current from then on
                widths = [self._pack_weights[0].shape[1]] + [W.shape[0] for W in self._pack_weights]
                self._packed_ptr, self._packed_buf = F._packed_weights(lib, self._pack_weights, widths, gemm_mode, dev)
                self.optimizer.set_packed_targets(F.packed_weight_targets(self._pack_weights, gemm_mode, self._packed_ptr))
            # (otherwise the weight re-pack is part of the recording: the cache was just invalidated)
            self.graph = torch.cuda.CUDAGraph()
            n0 = lib.clica_launch_count(-1)
            # NCCL's watchdog thread polls CUDA events while the collectives are being recorded: only this thread's
            # calls may be policed during a multi-GPU capture
            mode = "thread_local" if self.world > 1 else "global"
            with torch.cuda.graph(self.graph, stream=self.stream, capture_error_mode=mode):
                self._body()
            self.launches_per_replay = int(lib.clica_launch_count(-1) - n0)
            # roll the warm-up back: parameters, moments and the device step count
            with torch.no_grad():
                for p, s in zip(self.params, saved):
                    p.copy_(s)
                for b, s in saved_buffers:
                    b.copy_(s)
                for st in self.optimizer.state.values():
                    st["exp_avg"].zero_()
                    st["exp_avg_sq"].zero_()
                for group in self.optimizer.param_groups:
                    if group.get("_step_state") is not None:
                        group["_step_state"].zero_()
            if self._pack_weights is not None:
                F.repack_weights_into(self._packed_ptr, self._pack_weights, gemm_mode)      # planes of the restored weights
            self.stream.synchronize()
        cur.wait_stream(self.stream)
        torch.autograd.graph.increment_version(self.params)
        self._note_versions()

    def _note_versions(self):
        if self._pack_weights is not None:
            self._pack_versions = [W._version for W in self._pack_weights]

    # per-step API ----------------------------------------------------------------------------------------------
    def stage(self, z1: torch.Tensor, z2: torch.Tensor) -> None:
        """Put this step's (anchor, positive) latents where the graph reads them: host tensors go into the pinned
        staging buffer (``host_io=True``), device tensors into the device-resident input."""
        B = self.B
        if z1.shape != (B, self.n) or z2.shape != (B, self.n):
            raise RuntimeError(f"GraphedTrainStep: expected two [{B}, {self.n}] tensors, got {tuple(z1.shape)}, {tuple(z2.shape)}")
        if z1.is_cuda != z2.is_cuda:
            raise RuntimeError("GraphedTrainStep: z1 and z2 must both be host or both be device tensors")
        if not z1.is_cuda:
            if not self.host_io:
                raise RuntimeError("GraphedTrainStep(host_io=False) takes device tensors")
            self._replayed.synchronize()         # the previous replay's H2D copy has finished reading the buffer
            self.z_host[:B].copy_(z1)
            self.z_host[B:].copy_(z2)
        else:
            if self.host_io:
                raise RuntimeError("GraphedTrainStep(host_io=True) takes host tensors (the graph copies them to the device)")
            self.z_dev[:B].copy_(z1)
            self.z_dev[B:].copy_(z2)

    def replay(self) -> torch.Tensor:
        """One training step on the staged inputs (asynchronous). Returns the device tensor
        ``[loss, pos_mean, neg_mean]`` that the step overwrites."""
        if self._pack_weights is not None and any(W._version != v for W, v in zip(self._pack_weights, self._pack_versions)):
            # somebody else changed the weights since the last replay (checkpoint load, manual copy_): the planes the graph
            # reads are stale -- refresh them before the step
            F.repack_weights_into(self._packed_ptr, self._pack_weights, self._gemm_mode)
        self.graph.replay()
        self._replayed.record(torch.cuda.current_stream(self.device))
        self.steps_taken += 1
        torch.autograd.graph.increment_version(self.params)     # the graph updated the parameters in place
        self._note_versions()
        return self.out_dev

    def __call__(self, z1: Optional[torch.Tensor] = None, z2: Optional[torch.Tensor] = None) -> torch.Tensor:
        if z1 is not None:
            self.stage(z1, z2)
        return self.replay()

    def pinned_inputs(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """The two [B, n] halves of the pinned staging buffer the graph's host->device copy reads.  A producer (sampler,
        data loader) that writes the batch straight into them saves the host-side staging copy: call ``step_host()``
        without arguments afterwards.  They may be overwritten as soon as ``step_host`` has returned."""
        if not self.host_io:
            raise RuntimeError("pinned_inputs needs host_io=True")
        return self.z_host[:self.B], self.z_host[self.B:]

    def step_host(self, z1: Optional[torch.Tensor] = None, z2: Optional[torch.Tensor] = None) -> Tuple[float, float, float]:
        """Host tensors in, host floats out (what ``main_mlp.py:285`` reads with ``.item()``): stage, replay,
        wait for the graph's own device->host copy.  Without arguments the batch is taken from ``pinned_inputs()``."""
        if not self.host_io:
            raise RuntimeError("step_host needs host_io=True")
        if z1 is not None:
            self.stage(z1, z2)
        self.replay()
        torch.cuda.current_stream(self.device).synchronize()
        o = self.out_host
        return float(o[0]), float(o[1]), float(o[2])
