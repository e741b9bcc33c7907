"""CPU (-m "not gpu"): host-side mirror of the reference interface (drop-in losses / encoders modules)."""
import os
import sys

import numpy as np
import pytest
import torch

import clica_b200
from conftest import load_golden

REF = os.environ.get("CLICA_REFERENCE_DIR", "/root/reference")


@pytest.fixture(scope="module")
def dropin():
    if clica_b200.DROPIN_DIR not in sys.path:
        sys.path.insert(0, clica_b200.DROPIN_DIR)
    if os.path.isdir(REF):
        os.environ["CLICA_REFERENCE_DIR"] = REF
    for name in ("losses", "encoders"):
        sys.modules.pop(name, None)
    import encoders
    import losses
    assert os.path.dirname(os.path.abspath(losses.__file__)) == clica_b200.DROPIN_DIR
    return losses, encoders


def test_get_mlp_matches_reference_structure_and_init(dropin):
    _, encoders = dropin
    g = load_golden("mlp_init_n5")
    torch.manual_seed(1234)
    layers = [50, 250, 250, 250, 250, 50]
    f = encoders.get_mlp(n_in=5, n_out=5, layers=layers)
    assert layers[-1] == 5 and len(layers) == 7          # the reference appends n_out to the caller's list
    assert isinstance(f, torch.nn.Sequential)
    sd = f.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    assert [type(m).__name__ for m in f] == [str(t) for t in g["module_types"]]
    assert [m.negative_slope for m in f if isinstance(m, torch.nn.LeakyReLU)] == list(g["leaky_slope"])
    for k, v in sd.items():
        assert np.array_equal(v.numpy().ravel()[:8], g["head_" + k]), k
    y = f(torch.tensor(g["x"]))                           # CPU input: torch modules (config-1 plumbing)
    assert np.allclose(y.detach().numpy(), g["y"], rtol=1e-5, atol=1e-7)
    assert isinstance(f[-1], torch.nn.Linear) and f[-1].out_features == 5   # indexable like nn.Sequential
    assert f._plan() is not None and f._plan()[2] == 13


def test_get_mlp_rejects_bad_output_normalization(dropin):
    _, encoders = dropin
    with pytest.raises(ValueError):
        encoders.get_mlp(4, 4, [8], output_normalization="nope")


def test_layer_normalization_stack_is_left_to_torch(dropin):
    _, encoders = dropin
    f = encoders.get_mlp(4, 4, [8, 8], layer_normalization="bn")
    assert any(isinstance(m, torch.nn.BatchNorm1d) for m in f)
    assert f._plan() is None


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present on this box")
def test_cpu_inputs_delegate_to_the_reference_loss(dropin):
    losses, _ = dropin
    g = load_golden("lpnce_roll_p2_d10")
    assert hasattr(losses, "SimCLRLoss") and hasattr(losses, "CLLoss")   # re-exported, main_mlp.py:147
    crit = losses.LpSimCLRLoss(p=2, tau=1.0, simclr_compatibility_mode=True)
    a = torch.tensor(g["z1"], requires_grad=True)
    b = torch.tensor(g["z2"], requires_grad=True)
    with pytest.warns(UserWarning):
        mean, per_item, parts = crit(None, None, None, a, b, torch.roll(a, 1, 0))
    mean.backward()
    assert np.allclose(per_item.detach().numpy(), g["loss_i_32"], rtol=1e-5, atol=1e-6)
    assert np.allclose(a.grad.numpy(), g["g1_32"], rtol=1e-4, atol=1e-7)
    assert len(parts) == 2 and parts[0].dim() == 0


def test_cuda_path_fails_loudly_without_a_gpu():
    """No silent fallback: the functional API refuses CPU tensors instead of computing something else."""
    from clica_b200 import functional as F
    a = torch.randn(4, 3)
    with pytest.raises(RuntimeError):
        F.lp_infonce(a, a, a, 2.0)
    with pytest.raises(RuntimeError):
        F.mlp_forward(a, [torch.randn(5, 3)], [torch.randn(5)])


def test_graphed_step_refuses_a_cpu_model():
    """GraphedTrainStep is CUDA-only: a CPU encoder must raise, not run some other path."""
    import pytest
    import torch
    from clica_b200.graphed import GraphedTrainStep
    f = torch.nn.Sequential(torch.nn.Linear(4, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        GraphedTrainStep(f, None, object(), 8, 4)


def test_mixing_plan_recognises_only_the_frozen_bias_free_stack():
    import torch
    from clica_b200 import functional as F
    from clica_b200 import synth
    g = synth.build_mixing(6, 3)
    ws, slope = F.mixing_plan(g)
    assert len(ws) == 3 and slope == 0.2
    assert F.mixing_plan(torch.nn.Sequential(torch.nn.Linear(6, 6))) is None                  # has a bias
    g2 = synth.build_mixing(6, 2)
    g2[0].weight.requires_grad = True
    assert F.mixing_plan(g2) is None                                                          # trainable
    assert F.mixing_plan(torch.nn.Sequential(*list(g)[:2])) is None                           # ends on an activation
