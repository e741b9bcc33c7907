"""GPU (-m gpu): arguments outside the CUDA kernels' domain -- p < 1 (the reference's transposed branch,
losses.py:433-442) and pow=False (losses.py:453-457) -- on CUDA tensors: the drop-in class hands them explicitly to the
reference's own torch code, so results and gradients are the reference's, bit for bit."""
import sys
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kwargs", [dict(p=0.5), dict(p=2, pow=False)])
def test_out_of_domain_arguments_run_the_reference_code_on_cuda(kwargs, cuda_device):
    import clica_b200
    if clica_b200.DROPIN_DIR not in sys.path:
        sys.path.insert(0, clica_b200.DROPIN_DIR)
    import losses
    from _reference import load_reference_module
    ref = load_reference_module("losses")
    if ref is None:
        pytest.skip("no reference checkout reachable (baseline/_ref is made by build())")
    g = torch.Generator().manual_seed(11)
    base = [torch.randn(96, 6, generator=g) for _ in range(3)]
    ours_in = [t.clone().to(cuda_device).requires_grad_(True) for t in base]
    ref_in = [t.clone().to(cuda_device).requires_grad_(True) for t in base]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")              # the one-time "delegating to the reference" UserWarning
        mean, per_item, parts = losses.LpSimCLRLoss(tau=0.7, simclr_compatibility_mode=True, **kwargs)(
            None, None, None, *ours_in)
    mean_r, per_item_r, parts_r = ref.LpSimCLRLoss(tau=0.7, simclr_compatibility_mode=True, **kwargs)(
        None, None, None, *ref_in)
    mean.backward()
    mean_r.backward()
    assert torch.equal(mean, mean_r) and torch.equal(per_item, per_item_r)
    assert all(torch.equal(a, b) for a, b in zip(parts, parts_r))
    for a, b in zip(ours_in, ref_in):
        assert torch.equal(a.grad, b.grad)
