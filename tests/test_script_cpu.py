"""CPU (-m "not gpu"): BASELINE config 1 -- the reference's UNCHANGED main_mlp.py (`--n 5 --space-type box
--n-mixing-layer 3 --batch-size 512`, "reference plumbing, no GPU") driven through the drop-in modules.

On CPU tensors the drop-in LpSimCLRLoss / get_mlp hand over to the reference's own torch code,
so the per-step values must be IDENTICAL to the plain reference run with the same seed.  Needs a reference checkout
(/root/reference in the build container or the read-only copy `build()` places under baseline/_ref)."""
import json
import math
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ["--n", "5", "--space-type", "box", "--n-mixing-layer", "3", "--batch-size", "512", "--n-steps", "3",
        "--seed", "5", "--num-eval-batches", "1"]     # seed 5: the mixing net's rejection sampling ends in ~1 s


def _reference_dir():
    sys.path.insert(0, ROOT)
    from clica_b200 import vendor
    if os.path.isfile("/root/reference/main_mlp.py"):
        return "/root/reference"
    return vendor.vendored_dir()


def _run(arm, ref, dump):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")      # config 1 is the CPU configuration
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_main_mlp.py"), "--arm", arm, "--reference",
                          ref, "--dump", dump, "--"] + ARGS, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    with open(dump) as fh:
        return json.load(fh), out.stdout


def test_config1_runs_unchanged_and_matches_the_plain_reference(tmp_path):
    ref = _reference_dir()
    if ref is None:
        pytest.skip("no reference checkout reachable")
    ours, log = _run("ours", ref, str(tmp_path / "ours.json"))
    plain, _ = _run("plain", ref, str(tmp_path / "plain.json"))
    assert ours["n_steps"] == plain["n_steps"] == 3 + 3 * 3          # supervised, then 3x unsupervised
    assert ours["total"] == plain["total"] and ours["parts"] == plain["parts"]
    # KA1: a freshly initialised encoder maps everything near one point -> loss = ln(B + 1)
    assert abs(ours["total"][3] - math.log(513)) < 1e-4
    assert "Linear(in_features=5, out_features=50" in log and "supervised test: False" in log
