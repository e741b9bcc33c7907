"""FusedAdam: ``torch.optim.Optimizer``-compatible Adam on ``clica_adam_step`` (one launch per 32 tensors).

Same update rule and defaults as ``torch.optim.Adam(params, lr)`` used by ``main_mlp.py:312`` (betas
(0.9, 0.999), eps 1e-8, no weight decay, no amsgrad).  ``main_mlp.py`` itself keeps ``torch.optim.Adam``
(the script runs unchanged); this class is what ``bench.py``, the sharded step and the CUDA-graph step use.

``capturable=True`` keeps the step count on the device (``clica_adam_step_capturable``) so that ``step()`` can
be recorded into a CUDA graph and replayed: every replay is one more Adam step.
"""
import torch

from . import functional as F


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, capturable=False):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, capturable=bool(capturable)))

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
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                if not capturable:
                    st["step"] += 1
                    step = st["step"] if step is None else step
                    if st["step"] != step:
                        raise RuntimeError("FusedAdam: parameters of one group must share their step count")
                ps.append(p.data), gs.append(p.grad.contiguous()), ms.append(st["exp_avg"]), vs.append(st["exp_avg_sq"])
                owners.append(p)
            if not ps:
                continue
            if capturable:
                # {int64 step, float, float} on the device, shared by the group; advanced by the kernel itself
                state = group.get("_step_state")
                if state is None:
                    state = torch.zeros(2, dtype=torch.int64, device=ps[0].device)
                    group["_step_state"] = state
                F.adam_step_capturable(ps, gs, ms, vs, group["lr"], group["betas"][0], group["betas"][1],
                                       group["eps"], state)
            else:
                F.adam_step(ps, gs, ms, vs, group["lr"], group["betas"][0], group["betas"][1], group["eps"], step)
            # the kernel wrote the parameters through raw pointers: tell autograd / the packed-weight cache
            # (functional._packed_weights keys on the version counter) that they changed
            torch.autograd.graph.increment_version(owners)
        return loss

    def device_step_count(self, group_index=0):
        """Step count of a capturable group (reads the device counter; synchronises)."""
        state = self.param_groups[group_index].get("_step_state")
        return 0 if state is None else int(state[0].item())
