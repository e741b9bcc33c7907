"""GPU (-m gpu): BASELINE config 2 -- the reference's UNCHANGED main_mlp.py on the B200, once against the drop-in
modules (fused CUDA loss + tcgen05 encoder) and once as the plain reference (torch eager), same seed.

Compared: every value `train_step` (main_mlp.py:258-285) returns, supervised phase (MSE through the CUDA encoder) and
unsupervised phase (fused Lp-InfoNCE).  The first step of each phase sees identical parameters and identical samples,
so it is a direct parity check of forward values (1e-5 relative); later steps compare two fp32 Adam trajectories
whose gradients differ by rounding, held to 2e-4 relative.  Needs baseline/_ref (placed by `build()`; it
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
    assert max(rel) <= 2e-4, max(rel)
    for pa, pb in zip(ours["parts"][n_sup:], plain["parts"][n_sup:]):                # [pos_mean, neg_mean]
        assert all(abs(x - y) <= 2e-4 * max(1.0, abs(y)) for x, y in zip(pa, pb))
