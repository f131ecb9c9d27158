"""Launcher for UNMODIFIED DefTet scripts on top of the drop-in tree:

    cd /path/to/DefTet && python -m deftet_b200.run train_multigpu.py --res 70 --batch_size 8 ...
    cd /path/to/DefTet/diff_render/diftet_6_subdiv/6_optim && python -m deftet_b200.run optim_with_mask_subdiv_from_gridmov.py --expname hotdog
    ... python -m deftet_b200.run --leaf optim_with_mask_subdiv_from_gridmov.py ...    (keep the reference's own Deftet model and
                                                                                       rendermeshcolor, replace only the leaf modules)

`python script.py` puts the script's own directory FIRST on sys.path, so the checkout's `layers/` and `utils/` packages would
win over any PYTHONPATH entry and the reference's JIT extensions would be imported.  This launcher runs the script with
  sys.path = [deftet_b200/dropin, (deftet_b200/dropin/diff_render for the diff_render scripts), <repo root>, <script dir>, ...]
and DEFTET_REFERENCE_ROOT pointing at the checkout (found by walking up from the script to the directory that holds
train_multigpu.py, unless the variable is already set), so that the drop-in modules resolve first and everything they do not
replace falls through to the checkout.  The script itself is executed unmodified (runpy, __name__ == "__main__")."""
import os
import runpy
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
DROPIN = os.path.join(_HERE, "dropin")
REPO = os.path.dirname(_HERE)


def _find_checkout(start):
    d = os.path.abspath(start)
    while True:
        if os.path.isfile(os.path.join(d, "train_multigpu.py")) and os.path.isdir(os.path.join(d, "layers")):
            return d
        up = os.path.dirname(d)
        if up == d:
            return None
        d = up


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 0
    leaf = argv[0] == "--leaf"
    if leaf:
        argv = argv[1:]
    script = os.path.abspath(argv[0])
    if not os.path.isfile(script):
        raise SystemExit("deftet_b200.run: no such script: %s" % argv[0])
    script_dir = os.path.dirname(script)
    if not os.environ.get("DEFTET_REFERENCE_ROOT"):
        root = _find_checkout(script_dir)
        if root:
            os.environ["DEFTET_REFERENCE_ROOT"] = root
    front = [DROPIN]
    if os.sep + "diff_render" + os.sep in script + os.sep:
        front.append(os.path.join(DROPIN, "diff_render_leaf" if leaf else "diff_render"))
    front += [REPO, script_dir]
    sys.path[:] = front + [p for p in sys.path if p not in front and p not in ("", os.getcwd())]
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
