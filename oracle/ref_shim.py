"""Import the UNMODIFIED reference (hnuzhy/SemiUHPE) from ``/root/reference``.

TEST INFRASTRUCTURE ONLY, and usable only in the build container:
``/root/reference`` does not exist on the GPU box, so nothing in ``-m gpu``
tests, ``smoke()`` or ``bench.py`` may call :func:`load`.  Its one consumer is
``tests/golden/make_golden.py`` (fixture generator) plus the ``not gpu`` tests
that re-check the oracle against the live reference when it is present.

Three shims are needed to import the reference's rotation math without
touching its sources:

1. ``matplotlib.pyplot`` is imported but unused (src/fisher/fisher_utils.py:7);
2. ``pytorch3d.transforms`` is absent -> ``oracle.pytorch3d_restated``;
3. the reference calls ``.cuda()`` unconditionally
   (src/fisher/fisher_utils.py:35,47,65; src/fisher/between_bingham_fisher.py:81):
   on a CPU-only host ``Tensor.cuda`` is made the identity while loading/using it.
"""
import importlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("SEMIUHPE_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "fisher", "fisher_utils.py"))


class Reference(types.SimpleNamespace):
    """Bag of the reference's modules: fisher_utils, torch_norm_factor, bbf,
    bingham_utils, rotation_laplace, utils_euler (function), grids (dict)."""


_CACHE = None


def load():
    global _CACHE
    if _CACHE is not None:
        return _CACHE
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    from . import pytorch3d_restated as p3d

    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    pkg = types.ModuleType("pytorch3d")
    tr = types.ModuleType("pytorch3d.transforms")
    for fn in ("so3_relative_angle", "so3_rotation_angle", "matrix_to_quaternion",
               "quaternion_to_matrix", "acos_linear_extrapolation"):
        setattr(tr, fn, getattr(p3d, fn))
    pkg.transforms = tr
    sys.modules.setdefault("pytorch3d", pkg)
    sys.modules.setdefault("pytorch3d.transforms", tr)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    # src/__init__ may not exist; import by file location under the names the
    # reference itself uses (``from src.fisher.fisher_utils import ...``).
    def _mod(dotted):
        return importlib.import_module(dotted)

    fisher_utils = _mod("src.fisher.fisher_utils")
    ref = Reference(
        fisher_utils=fisher_utils,
        torch_norm_factor=sys.modules["torch_norm_factor"],
        bbf=sys.modules["between_bingham_fisher"],
        bingham_utils=sys.modules["bingham_utils"],
        rotation_laplace=_mod("src.laplace.rotation_laplace"),
    )
    # src/utils.py imports cv2 etc.; pull the two pure functions by exec of their source slice
    ref.euler_from_matrices, ref.limit_angle, ref.get_6DRepNet_Rot = _load_utils_functions()
    ref.grid_path = lambda name="eq_grids2.npy": os.path.join(REFERENCE_ROOT, "src", "laplace", name)
    _CACHE = ref
    return ref


def _load_utils_functions():
    """``src/utils.py`` drags in cv2/matplotlib at import time; try the plain
    import first and fall back to compiling only the three function bodies we
    need out of the file (still the reference's own code, executed unmodified)."""
    try:
        u = importlib.import_module("src.utils")
        return (u.compute_euler_angles_from_rotation_matrices, u.limit_angle, u.get_6DRepNet_Rot)
    except Exception:
        import ast
        import math
        import numpy as np
        path = os.path.join(REFERENCE_ROOT, "src", "utils.py")
        tree = ast.parse(open(path).read())
        want = {"compute_euler_angles_from_rotation_matrices", "limit_angle",
                "get_6DRepNet_Rot", "rot_euler_6DRepNet"}
        body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
        ns = {"torch": torch, "np": np, "math": math}
        exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
        return (ns["compute_euler_angles_from_rotation_matrices"], ns["limit_angle"],
                ns["get_6DRepNet_Rot"])
