"""CPU (no GPU needed): host-side decisions that the GPU tests only exercise indirectly -- the roofline accounting of
bench.py, the kernel-family classifier of its breakdown, and which criteria the graphed step
records in the pairs form of the loss."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_algorithmic_work_matches_the_survey_figures():
    import bench
    c2 = bench.algorithmic_work(bench.WORKLOADS["c2"], 6144, 6144, 1)
    c3 = bench.algorithmic_work(bench.WORKLOADS["c3"], 8192, 8192, 1)
    assert c2["mac_row"] == 852_000 and c3["mac_row"] == 13_632_000                 # sum K*N of the 7 layers
    assert c2["enc_flops"] == 2 * 6144 * (6 * 852_000 - 2 * 10 * 100)               # fwd + dX + dW, 2 calls, no layer-0 dX
    assert abs(c2["enc_flops"] - 62.8e9) < 0.05e9 and abs(c3["enc_flops"] - 1.34e12) < 0.005e12
    assert c2["loss_bytes"] == 24 * 6144 * 10 + 8 * 6144 == 1_523_712               # bytes_min of the 3-input API
    assert c2["loss_ops_fwd"] == 2 * 6144 * 6144 * 10 and c2["loss_ops_bwd"] == 4 * 6144 * 6144 * 10
    assert c3["loss_ops_fwd"] == 3 * 8192 * 8192 * 40 and c3["loss_ops_bwd"] == 5 * 8192 * 8192 * 40
    # row-sharded: local rows x all columns; bytes = own rows in / out + the gathered negatives
    sh = bench.algorithmic_work(bench.WORKLOADS["c3"], 1024, 8192, 8)
    assert sh["loss_ops_fwd"] == 3 * 1024 * 8192 * 40
    assert sh["loss_bytes"] == 12 * 1024 * 40 + 4 * 8192 * 40 + 8 * 1024


@pytest.mark.parametrize("name,family", [
    ("void clica::lpnce_fwd_kernel<2, 5, 4, 4, 1>(clica::FwdParams)", "loss_fwd"),
    ("void clica::lpnce_bwd_kernel<3, 20, 1, 4, 1>(clica::BwdParams)", "loss_bwd"),
    ("clica::lpnce_reduce_kernel(clica::ReduceParams)", "loss_aux"),
    ("void clica::(anonymous namespace)::gemm_tc_kernel<2, false>(clica::ChainParams)", "gemm_tc"),
    ("void clica::skinny_kin_kernel<16>(clica::SkinnyKinParams)", "gemm_simt"),
    ("void clica::mixing_fwd_kernel<16>(clica::MixParams)", "gemm_simt"),
    ("clica::adam_kernel(clica::AdamArgs)", "adam"),
    ("clica::split_planes_multi_kernel(clica::SplitMultiArgs)", "misc"),
    ("ncclDevKernel_AllReduce_Sum_f32_RING_LL(ncclDevKernelArgsStorage<4096ul>)", "nccl"),
    ("Memcpy HtoD (Pinned -> Device)", "memops"),
    ("void at::native::vectorized_elementwise_kernel<4, ...>", "torch_other"),
])
def test_kernel_family_classifier(name, family):
    import bench
    assert bench.family_of(name) == family


def test_pairs_config_only_for_criteria_inside_the_kernels_domain():
    from clica_b200.graphed import pairs_config

    class LpSimCLRLoss:
        def __init__(self, p, tau=1.0, alpha=0.5, simclr_compatibility_mode=False, pow=True):
            self.p, self.tau, self.alpha, self.simclr_compatibility_mode, self.pow = p, tau, alpha, simclr_compatibility_mode, pow

    class SimCLRLoss:
        def __init__(self, normalize=False, tau=1.0, alpha=0.5):
            self.normalize, self.tau, self.alpha = normalize, tau, alpha

    class Other:
        pass

    assert pairs_config(LpSimCLRLoss(2, 0.7, 0.3, True)) == (2.0, 0.7, 0.3, True)
    assert pairs_config(LpSimCLRLoss(3)) == (3.0, 1.0, 0.5, False)
    assert pairs_config(LpSimCLRLoss(0.5)) is None            # p < 1: the reference's transposed branch (losses.py:433-442)
    assert pairs_config(LpSimCLRLoss(2, pow=False)) is None
    assert pairs_config(SimCLRLoss(False, 0.5)) == (0.0, 0.5, 0.5, True)
    assert pairs_config(SimCLRLoss(True)) is None             # normalisation stays torch code in front of the kernel
    assert pairs_config(Other()) is None


def _reference_spaces():
    """The reference's spaces.py (the checkout, or the read-only copy build() vendors); None when neither is present."""
    import importlib.util
    import warnings
    for d in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        path = os.path.join(d, "spaces.py")
        if os.path.isfile(path):
            sys.path.insert(0, d)                    # spaces.py imports spaces_utils / vmf by bare name
            try:
                spec = importlib.util.spec_from_file_location("_ref_spaces_for_test", path)
                mod = importlib.util.module_from_spec(spec)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore", SyntaxWarning)
                    spec.loader.exec_module(mod)
                return mod
            finally:
                sys.path.remove(d)
    return None


def test_sampler_classes_keep_the_reference_surface_and_delegate_on_cpu():
    """clica_b200.samplers: same constructors / attributes as spaces.py:47-351; for a CPU device every draw
    is the reference's own method (bit-identical under the same seed); the module object re-exports the other names."""
    import torch
    from clica_b200 import samplers
    ref = _reference_spaces()
    if ref is None:
        pytest.skip("no reference checkout reachable")
    mod = samplers.build_module(ref)
    assert mod.__clica_device_samplers__ and mod.Space is ref.Space
    for name in ("NRealSpace", "NSphereSpace", "NBoxSpace"):
        assert issubclass(getattr(mod, name), getattr(ref, name)) and getattr(mod, name) is not getattr(ref, name)
    ours, theirs = mod.NSphereSpace(7), ref.NSphereSpace(7)
    assert (ours.n, ours.r, ours.dim) == (theirs.n, theirs.r, theirs.dim)
    torch.manual_seed(5)
    a = ours.uniform(64, device="cpu")
    torch.manual_seed(5)
    b = theirs.uniform(64, device="cpu")
    assert torch.equal(a, b)
    torch.manual_seed(6)
    a2 = ours.normal(a, 0.05, 64, device="cpu")
    torch.manual_seed(6)
    b2 = theirs.normal(b, 0.05, 64, device="cpu")
    assert torch.equal(a2, b2)
    box_o, box_t = mod.NBoxSpace(4, -1.0, 1.0), ref.NBoxSpace(4, -1.0, 1.0)
    torch.manual_seed(7)
    c = box_o.uniform(32, device="cpu")
    torch.manual_seed(7)
    d = box_t.uniform(32, device="cpu")
    assert torch.equal(c, d) and c.min() >= -1 and c.max() <= 1
    real_o, real_t = mod.NRealSpace(3), ref.NRealSpace(3)
    torch.manual_seed(8)
    e = real_o.normal(torch.zeros(3), 0.5, 16, device="cpu")
    torch.manual_seed(8)
    f = real_t.normal(torch.zeros(3), 0.5, 16, device="cpu")
    assert torch.equal(e, f)
    # tensor-valued scales stay with the reference even for a CUDA device string (no kernel for them)
    assert samplers._scalar(0.05) and not samplers._scalar(torch.ones(3))
