"""Module injection: run the UNMODIFIED reference on the B200 path (option B).

The reference imports its rotation math by module path (src/agent.py:14-20,
predict.py:15, image.py:23, video.py:23, eval_DAD3DHeads.py:14).  Registering
the mirrors under those names before the reference's own packages are imported
makes every later ``from src.fisher.fisher_utils import vmf_loss`` resolve to the
CUDA path; nothing in the reference tree is edited.
"""
import importlib
import sys
import types

_MIRRORS = {
    "src.fisher.fisher_utils": "semiuhpe_b200.fisher.fisher_utils",
    "src.fisher.torch_norm_factor": "semiuhpe_b200.fisher.torch_norm_factor",
    "src.laplace.rotation_laplace": "semiuhpe_b200.laplace.rotation_laplace",
}


def _package(name):
    """The already-imported package ``name`` or an empty namespace stand-in for it."""
    mod = sys.modules.get(name)
    if mod is None:
        try:
            mod = importlib.import_module(name)
        except Exception:
            mod = types.ModuleType(name)
            mod.__path__ = []
            sys.modules[name] = mod
    return mod


def patch_reference(patch_utils=True, patch_agent=True):
    """Register the mirrors under the reference's module names; returns the list of
    names that were patched.  Call before ``import train`` / ``from src.agent import ...``
    (modules that already did ``from ... import name`` keep their old binding)."""
    done = []
    for ref_name, ours_name in _MIRRORS.items():
        ours = importlib.import_module(ours_name)
        parent_name, _, leaf = ref_name.rpartition(".")
        if parent_name.count("."):
            _package(parent_name.rpartition(".")[0])
        parent = _package(parent_name)
        sys.modules[ref_name] = ours
        setattr(parent, leaf, ours)
        done.append(ref_name)
    if patch_utils:
        # src/utils.py keeps its logging helpers; only the Euler function is swapped (src/utils.py:232)
        from . import utils as ours_utils
        ref_utils = sys.modules.get("src.utils")
        if ref_utils is None:
            try:
                ref_utils = importlib.import_module("src.utils")
            except Exception:
                ref_utils = None
        if ref_utils is not None:
            ref_utils.compute_euler_angles_from_rotation_matrices = \
                ours_utils.compute_euler_angles_from_rotation_matrices
            done.append("src.utils.compute_euler_angles_from_rotation_matrices")
    if patch_agent and "src.agent" in sys.modules:
        patch_agent_class(sys.modules["src.agent"].SSLAgent)
        done.append("src.agent.SSLAgent")
    return done


def patch_agent_class(cls):
    """Bind the filter / metric slice of ``SSLAgent`` (src/agent.py:357-455) to the CUDA path."""
    from . import agent as ours
    cls.compute_dynamic_entropy_threshold = ours.compute_dynamic_entropy_threshold
    cls.compute_err_deg_from_matrices = staticmethod(ours.compute_err_deg_from_matrices)
    cls.compute_err_deg_from_quats = staticmethod(ours.compute_err_deg_from_quats)
    return cls
