#!/usr/bin/env python
"""Run the reference's UNCHANGED main_mlp.py and record what every ``train_step`` call returned.

    python tools/run_main_mlp.py --arm {ours,plain} [--reference DIR] [--dump out.json] -- <main_mlp.py args>

* ``--arm plain``: ``sys.path = [<reference>, ...]`` -- the script imports the reference's own losses / encoders.
* ``--arm ours`` : ``sys.path = [cl-ica_b200/dropin, <reference>, ...]`` -- exactly what ``python -m clica_b200.launch``
  does: ``import losses, encoders`` resolve to the drop-in modules, everything else to the reference.

The script file is executed byte-identical with ``runpy.run_path``; the per-step values ``(total_loss, [parts])`` that
``train_step`` (main_mlp.py:258-285) returns are observed from outside with ``sys.setprofile`` (return events of the
code object named ``train_step`` in that file) -- no patching of either arm.  Wall time of each phase (first to last
train_step return) is recorded too, so the dump doubles as the script-level pairs/s measurement.
"""
import argparse
import json
import os
import runpy
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", choices=["ours", "plain"], required=True)
    ap.add_argument("--reference", default=None)
    ap.add_argument("--dump", default=None)
    ap.add_argument("--device-samplers", action="store_true",
                    help="ours only: install clica_b200.samplers' device-side samplers behind the reference's spaces API")
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    args = ap.parse_args()
    rest = args.rest[1:] if args.rest and args.rest[0] == "--" else args.rest
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import clica_b200
    from clica_b200 import vendor
    ref = args.reference or os.environ.get("CLICA_REFERENCE_DIR") or vendor.vendored_dir() or "/root/reference"
    ref = os.path.abspath(ref)
    script = os.path.join(ref, "main_mlp.py")
    if not os.path.isfile(script):
        raise SystemExit(f"{script} not found")
    head = [clica_b200.DROPIN_DIR, ref] if args.arm == "ours" else [ref]
    if args.arm == "ours":
        os.environ["CLICA_REFERENCE_DIR"] = ref
    sys.path[:] = head + [p for p in sys.path if p not in head]
    if args.arm == "ours" and args.device_samplers:
        from clica_b200 import samplers
        samplers.install()
    warnings.simplefilter("ignore", SyntaxWarning)

    steps = []          # (t_return, total, parts)
    t_first = {}

    def prof(frame, event, arg):
        if event == "return" and frame.f_code.co_name == "train_step" and frame.f_code.co_filename == script:
            if isinstance(arg, tuple) and len(arg) == 2:
                steps.append((time.perf_counter(), float(arg[0]), [float(x) for x in arg[1]]))
        return None

    sys.argv = [script] + rest
    sys.setprofile(prof)
    t0 = time.perf_counter()
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        sys.setprofile(None)
    wall = time.perf_counter() - t0
    out = {"arm": args.arm, "argv": rest, "wall_s": wall, "n_steps": len(steps),
           "total": [s[1] for s in steps], "parts": [s[2] for s in steps],
           "t_rel": [s[0] - t0 for s in steps]}
    if args.dump:
        os.makedirs(os.path.dirname(os.path.abspath(args.dump)), exist_ok=True)
        with open(args.dump, "w") as fh:
            json.dump(out, fh)
    print(f"[run_main_mlp] arm={args.arm} steps={len(steps)} wall={wall:.1f}s first={out['total'][:1]} last={out['total'][-1:]}")


if __name__ == "__main__":
    main()
