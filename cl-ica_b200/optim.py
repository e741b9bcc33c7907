"""FusedAdam: ``torch.optim.Optimizer``-compatible Adam on ``clica_adam_step`` (one launch per 32 tensors).

Same update rule and defaults as ``torch.optim.Adam(params, lr)`` used by ``main_mlp.py:312`` (betas
(0.9, 0.999), eps 1e-8, no weight decay, no amsgrad).  ``main_mlp.py`` itself keeps ``torch.optim.Adam``
(the script runs unchanged); this class is what ``bench.py``, the sharded step and the CUDA-graph step use.

``capturable=True`` keeps the step count on the device (``clica_adam_step_capturable``) so that ``step()`` can
be recorded into a CUDA graph and replayed: every replay is one more Adam step.

Checkpoints interchange with ``torch.optim.Adam``: ``state_dict()`` has torch's layout (per-parameter ``step`` as a
0-dim float32 tensor, ``exp_avg``, ``exp_avg_sq``; param_groups with lr / betas / eps plus torch's remaining keys at
their defaults); the device-side counter of a capturable group is written into ``step`` on save and restored from it
on ``load_state_dict()``.
"""
import torch

from . import functional as F


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, capturable=False):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, capturable=bool(capturable)))
        self._packed_targets = None

    def set_packed_targets(self, targets):
        """``{id(weight): (hi_ptr, lo_ptr, cols, ld)}`` (``functional.packed_weight_targets``): a capturable step then also
        writes those weights in the encoder GEMMs' operand format, so no re-pack pass follows the update.  ``None``
        switches it off.  The caller owns the packed buffer and keeps it alive."""
        self._packed_targets = dict(targets) if targets else None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps, gs, ms, vs, owners = [], [], [], [], []
            step = None
            capturable = group.get("capturable", False)
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)       # torch.optim.Adam's layout
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                if not capturable:
                    st["step"] = st["step"] + 1 if torch.is_tensor(st["step"]) else torch.tensor(float(st["step"]) + 1.0)
                    cur = int(st["step"])
                    step = cur if step is None else step
                    if cur != step:
                        raise RuntimeError("FusedAdam: parameters of one group must share their step count")
                g = p.grad
                if not g.is_contiguous():
                    if not getattr(self, "_warned_noncontig", False):
                        import warnings
                        warnings.warn("FusedAdam: a non-contiguous .grad is copied on every step", stacklevel=2)
                        self._warned_noncontig = True
                    g = g.contiguous()
                ps.append(p.data), gs.append(g), ms.append(st["exp_avg"]), vs.append(st["exp_avg_sq"])
                owners.append(p)
            if not ps:
                continue
            if capturable:
                # {int64 step, float, float} on the device, shared by the group; advanced by the kernel itself
                state = group.get("_step_state")
                if state is None:
                    state = torch.zeros(2, dtype=torch.int64, device=ps[0].device)
                    pending = int(self.state[owners[0]]["step"])               # e.g. restored by load_state_dict
                    if pending:
                        state[0] = pending
                    group["_step_state"] = state
                pack = None
                if self._packed_targets:
                    pack = [self._packed_targets.get(id(p)) for p in owners]
                F.adam_step_capturable(ps, gs, ms, vs, group["lr"], group["betas"][0], group["betas"][1],
                                       group["eps"], state, pack=pack)
            else:
                F.adam_step(ps, gs, ms, vs, group["lr"], group["betas"][0], group["betas"][1], group["eps"], step)
            # the kernel wrote the parameters through raw pointers: tell autograd / the packed-weight cache
            # (functional._packed_weights keys on the version counter) that they changed
            torch.autograd.graph.increment_version(owners)
        return loss

    # ---- checkpoints in torch.optim.Adam's layout ------------------------------------------------------
    _TORCH_GROUP_DEFAULTS = dict(weight_decay=0, amsgrad=False, maximize=False, foreach=None, differentiable=False,
                                 fused=None, decoupled_weight_decay=False)

    def state_dict(self):
        for gi, group in enumerate(self.param_groups):
            if group.get("capturable") and group.get("_step_state") is not None:
                count = float(self.device_step_count(gi))
                for p in group["params"]:
                    if p in self.state and self.state[p]:
                        self.state[p]["step"] = torch.tensor(count, dtype=torch.float32)
        sd = super().state_dict()
        for g in sd["param_groups"]:
            g.pop("_step_state", None)
            for k, v in self._TORCH_GROUP_DEFAULTS.items():
                g.setdefault(k, v)
        return sd

    def load_state_dict(self, state_dict):
        live = [g.get("_step_state") for g in self.param_groups]
        super().load_state_dict(state_dict)
        for group, st in zip(self.param_groups, live):
            for k in self._TORCH_GROUP_DEFAULTS:
                group.pop(k, None)
            steps = [int(self.state[p]["step"]) for p in group["params"] if p in self.state and "step" in self.state[p]]
            for p in group["params"]:
                if p in self.state and "step" in self.state[p]:
                    self.state[p]["step"] = torch.tensor(float(int(self.state[p]["step"])), dtype=torch.float32)
            if group.get("capturable"):
                if st is not None:
                    st[0] = steps[0] if steps else 0           # keep the buffer a recorded CUDA graph points at
                    group["_step_state"] = st
                else:
                    group.pop("_step_state", None)             # created (from `step`) by the next step()

    def device_step_count(self, group_index=0):
        """Step count of a capturable group (reads the device counter; synchronises)."""
        state = self.param_groups[group_index].get("_step_state")
        return 0 if state is None else int(state[0].item())
