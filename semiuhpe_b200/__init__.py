"""semiuhpe_b200 -- B200-native (sm_100a) rotation-distribution hot path of SemiUHPE.

Drop-in for the reference's Python functions on this path (same names,
argument meaning and error behaviour), backed by hand-written CUDA kernels
behind a C ABI (``include/semiuhpe_b200.h``).  Module layout mirrors the
reference so its ``src/agent.py`` switches by changing import lines only
(see INTEGRATION.md):

=====================================  =========================================
reference                              here
=====================================  =========================================
``src.fisher.fisher_utils``            ``semiuhpe_b200.fisher.fisher_utils``
``src.fisher.torch_norm_factor``       ``semiuhpe_b200.fisher.torch_norm_factor``
``src.laplace.rotation_laplace``       ``semiuhpe_b200.laplace.rotation_laplace``
``src.utils`` (Euler helpers)          ``semiuhpe_b200.utils``
``SSLAgent`` filter / metric methods   ``semiuhpe_b200.agent``
(none: single GPU)                     ``semiuhpe_b200.distributed``
=====================================  =========================================

There is no CPU path and no fallback: tensors must be CUDA fp32 and the
library must have been built (``__graft_entry__.build()``).
"""
from . import _capi

__all__ = ["set_error_checking", "error_checking", "library_path", "set_quadrature_cut_bits"]

_CHECK = True


def set_error_checking(enabled):
    """The reference raises on non-finite input (``torch.svd`` -> LinAlgError) and
    on out-of-range traces (pytorch3d ``ValueError``).  Reproducing that needs one
    4-byte device->host read per call (a sync the reference pays anyway at every
    ``.cpu()``).  Disable for sync-free / CUDA-graph use; the kernels then only
    leave NaNs in the outputs of the offending rows."""
    global _CHECK
    _CHECK = bool(enabled)


def error_checking():
    return _CHECK


def set_quadrature_cut_bits(bits):
    """Negligible-node cut of the Fisher quadrature (``suhpe_set_quadrature_cut_bits``): nodes
    whose total contribution is provably below ``2**-bits`` of the normaliser sum are skipped.
    Default 26; ``0`` evaluates all 512 nodes of every integral.  Returns the previous value."""
    return _capi.lib().suhpe_set_quadrature_cut_bits(int(bits))


def library_path():
    from . import _build
    return _build.LIB_PATH
