"""Run the reference's byte-identical ``main_mlp.py`` against the drop-in ``losses`` / ``encoders`` modules.

    python -m clica_b200.launch --reference /root/reference -- --n 10 --space-type sphere --p 2 ...

``sys.path`` becomes ``[<dropin dir>, <reference dir>, ...]`` so ``import losses, encoders`` inside the script
resolve to ``cl-ica_b200/dropin`` while every other module (spaces, latent_spaces, layers, ...) still comes
from the reference checkout.  The script file itself is executed unmodified with ``runpy.run_path``.
"""
import argparse
import os
import runpy
import sys
import warnings


def _default_reference():
    if os.environ.get("CLICA_REFERENCE_DIR"):
        return os.environ["CLICA_REFERENCE_DIR"]
    if os.path.isfile("/root/reference/main_mlp.py"):
        return "/root/reference"
    from clica_b200 import vendor
    return vendor.vendored_dir() or "/root/reference"


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", default=_default_reference(),
                    help="reference checkout (default: $CLICA_REFERENCE_DIR, /root/reference, then baseline/_ref)")
    ap.add_argument("--script", default="main_mlp.py")
    ap.add_argument("--device-samplers", action="store_true",
                    help="install clica_b200.samplers behind the reference's `spaces` module (one launch per draw, no host syncs)")
    ap.add_argument("script_args", nargs=argparse.REMAINDER)
    args = ap.parse_args(argv)
    ref = os.path.abspath(args.reference)
    script = os.path.join(ref, args.script)
    if not os.path.isfile(script):
        raise SystemExit(f"{script} not found (pass --reference)")
    import clica_b200
    os.environ["CLICA_REFERENCE_DIR"] = ref
    sys.path[:] = [clica_b200.DROPIN_DIR, ref] + [p for p in sys.path if p not in (clica_b200.DROPIN_DIR, ref)]
    if args.device_samplers:
        from clica_b200 import samplers
        samplers.install()
    rest = args.script_args
    if rest and rest[0] == "--":
        rest = rest[1:]
    sys.argv = [script] + rest
    warnings.simplefilter("ignore", SyntaxWarning)
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
