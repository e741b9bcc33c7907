"""GPU (-m gpu): BASELINE config 2 -- the reference's UNCHANGED main_mlp.py on the B200, once against the drop-in
modules (fused CUDA loss + tcgen05 encoder) and once as the plain reference (torch eager), same seed.

Compared: every value `train_step` (main_mlp.py:258-285) returns, supervised phase (MSE through the CUDA encoder) and
unsupervised phase (fused Lp-InfoNCE).  The first step of each phase sees identical parameters and identical samples,
so it is a direct parity check of forward values (1e-5 relative); later steps compare two fp32 Adam trajectories
whose gradients differ by rounding, held to 1e-5 relative (measured 2.2e-6).  Needs baseline/_ref (placed by `build()`; it
travels to the GPU box with the snapshot)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ["--n", "10", "--space-type", "sphere", "--p", "2", "--tau", "1.0", "--batch-size", "6144", "--n-steps", "20",
        "--seed", "0", "--num-eval-batches", "1"]


def _run(arm, ref, dump, extra=()):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_main_mlp.py"), "--arm", arm, "--reference",
                          ref, "--dump", dump, *extra, "--"] + ARGS, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    with open(dump) as fh:
        return json.load(fh), out


def test_config2_unchanged_script_on_the_gpu_matches_the_plain_reference(cuda_device, tmp_path):
    sys.path.insert(0, ROOT)
    from clica_b200 import vendor
    ref = vendor.vendored_dir()
    if ref is None:
        pytest.skip("baseline/_ref is absent (run __graft_entry__.build() where /root/reference exists)")
    ours, out = _run("ours", ref, str(tmp_path / "ours.json"))
    assert "delegating to the reference" not in out.stderr, out.stderr[-2000:]     # the CUDA path ran, not the torch one
    plain, _ = _run("plain", ref, str(tmp_path / "plain.json"))
    n_sup, n_unsup = 20, 60
    assert ours["n_steps"] == plain["n_steps"] == n_sup + n_unsup
    a, b = ours["total"], plain["total"]
    rel = [abs(x - y) / max(abs(y), 1e-30) for x, y in zip(a, b)]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "script_c2_trajectories.json"), "w") as fh:
        json.dump({"ours": ours, "plain": plain, "max_rel": max(rel)}, fh)
    assert rel[0] <= 1e-5 and rel[n_sup] <= 1e-5, (rel[0], rel[n_sup])          # first step of each phase
    assert max(rel) <= 1e-5, max(rel)            # measured on B200: 2.2e-6 over the 80 steps
    for pa, pb in zip(ours["parts"][n_sup:], plain["parts"][n_sup:]):                # [pos_mean, neg_mean]
        assert all(abs(x - y) <= 2e-4 * max(1.0, abs(y)) for x, y in zip(pa, pb))


def test_config2_script_with_device_samplers(cuda_device, tmp_path):
    """Same unchanged script with clica_b200.samplers installed behind `import spaces` (a different random stream, so no
    step-by-step comparison): the first unsupervised step of a fresh encoder must show KA1 (loss = ln(B + 1), SURVEY
    section 4), the run must stay finite and close to the host-sampled run's loss level, and it must not be slower."""
    import math
    sys.path.insert(0, ROOT)
    from clica_b200 import vendor
    ref = vendor.vendored_dir()
    if ref is None:
        pytest.skip("baseline/_ref is absent")
    dev, out = _run("ours", ref, str(tmp_path / "dev.json"), extra=("--device-samplers",))
    host, _ = _run("ours", ref, str(tmp_path / "host.json"))
    n_sup = 20
    assert dev["n_steps"] == host["n_steps"] == 80
    assert all(math.isfinite(x) for x in dev["total"])
    assert abs(dev["total"][n_sup] - math.log(6145)) < 5e-3
    tail_dev = sum(dev["total"][-10:]) / 10
    tail_host = sum(host["total"][-10:]) / 10
    assert abs(tail_dev - tail_host) <= 0.02 * abs(tail_host), (tail_dev, tail_host)
    import numpy as np
    dt_dev = float(np.median(np.diff(dev["t_rel"][n_sup + 5:])))
    dt_host = float(np.median(np.diff(host["t_rel"][n_sup + 5:])))
    with open(os.path.join(ROOT, "gpurun_out", "script_c2_device_samplers.json"), "w") as fh:
        json.dump({"device_samplers_ms_per_step": dt_dev * 1e3, "host_samplers_ms_per_step": dt_host * 1e3,
                   "pairs_per_s_device_samplers": 6144 / dt_dev, "pairs_per_s_host_samplers": 6144 / dt_host}, fh)
    assert dt_dev <= 1.1 * dt_host


def test_config4_conv_encoder_step_matches_the_reference_loss(cuda_device, tmp_path):
    """BASELINE config 4: the reference's own 64x64 ConvNet (kitti_masks/model.py:28-56, via cuDNN) feeding
    the fused loss through strided views, as kitti_masks/solver.py:60-75 does; 5-step loss trajectory vs the reference."""
    sys.path.insert(0, ROOT)
    from clica_b200 import vendor
    if vendor.vendored_dir() is None:
        pytest.skip("baseline/_ref is absent")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "config4_bench.py"), "--batch", "256", "--steps", "3",
                          "--out", str(tmp_path / "c4.json")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    with open(tmp_path / "c4.json") as fh:
        d = json.load(fh)
    assert d["rel_loss_diff_per_step"][0] <= 1e-5 and d["max_rel_loss_diff_first_5_steps"] <= 2e-3
