#!/usr/bin/env python
"""Stage the UNMODIFIED reference modules of the hot path under ``baseline/_ref/``.

``baseline/_ref/`` is git-ignored (no reference source enters the history) but not gpurun-ignored, so the copy travels to the GPU
box, where ``bench.py --impl reference`` and the ``cpu_baseline`` / ``eager_b200`` legs import ``look2hear.models.TasNet`` /
``Sepformer`` and ``look2hear.losses.PITLossWrapper`` from it: the reference's own classes, on the host cores (and, for the
second bar, eagerly on the B200 through cuDNN / cuBLAS).  The reference is pure Python without a setup.py / pyproject, so
``pip install`` has nothing to build; the copy of the two sub-packages it imports is the whole "install".  Run in the build
container (``/root/reference`` does not exist on the GPU box): ``python baseline/make_ref.py``; ``__graft_entry__.build()`` calls it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("DUALPATH_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def stage(verbose=True) -> bool:
    pkg = os.path.join(SRC, "look2hear")
    if not os.path.isdir(pkg):
        if verbose:
            print(f"make_ref: {pkg} not present (GPU box?) - keeping whatever is staged under {DST}")
        return os.path.isdir(os.path.join(DST, "look2hear", "models"))
    out = os.path.join(DST, "look2hear")
    if os.path.isdir(out):
        shutil.rmtree(out)
    os.makedirs(out)
    for sub in ("models", "losses"):
        shutil.copytree(os.path.join(pkg, sub), os.path.join(out, sub), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    open(os.path.join(out, "__init__.py"), "w").close()   # the reference relies on namespace-package lookup; an empty marker is equivalent
    cfg = os.path.join(DST, "configs")
    if os.path.isdir(cfg):
        shutil.rmtree(cfg)
    shutil.copytree(os.path.join(SRC, "configs"), cfg)
    if verbose:
        print(f"make_ref: staged look2hear/{{models,losses}} + configs under {DST}")
    return True


def import_reference():
    """(models, losses) modules of the staged reference, or raise ImportError."""
    if not os.path.isdir(os.path.join(DST, "look2hear", "models")):
        raise ImportError(f"{DST}/look2hear is not staged (run baseline/make_ref.py in the build container)")
    if DST not in sys.path:
        sys.path.insert(0, DST)
    import look2hear.losses as L
    import look2hear.models as M

    if not os.path.abspath(M.__file__).startswith(os.path.abspath(DST)):
        raise ImportError(f"look2hear resolved to {M.__file__}, not to the staged reference")
    return M, L


if __name__ == "__main__":
    ok = stage()
    if ok:
        M, L = import_reference()
        print("make_ref: import ok:", M.TasNet.__name__, M.Sepformer.__name__, L.PITLossWrapper.__name__)
