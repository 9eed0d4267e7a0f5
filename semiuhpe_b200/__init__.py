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

__all__ = ["set_error_checking", "error_checking", "library_path", "set_quadrature_cut_bits", "quadrature_cut_bits"]

_CHECK = True
_CUT_BITS = _capi.CUT_BITS_DEFAULT


def set_error_checking(enabled):
    """The reference raises on non-finite input (``torch.svd`` -> LinAlgError) and
    on out-of-range traces (pytorch3d ``ValueError``).  Reproducing that needs one
    4-byte device->host read per call (a sync the reference pays anyway at every
    ``.cpu()``).  Disable for sync-free / CUDA-graph use; the kernels then only
    leave NaNs in the outputs of the offending rows.  Switching it back on clears
    whatever the unchecked launches left in the status words, so a later call is
    never blamed for an earlier one."""
    global _CHECK
    enabled = bool(enabled)
    if enabled and not _CHECK:
        from . import _ops
        _ops.reset_status()
    _CHECK = enabled


def error_checking():
    return _CHECK


def set_quadrature_cut_bits(bits):
    """Default ``cut_bits`` the Python mirrors pass to the C ABI (the library itself keeps no setting: it is a
    per-call argument of ``suhpe_fisher_fused_f32``).  Nodes whose total contribution is provably below
    ``2**-bits`` of the normaliser sum are skipped; default 26; ``0`` evaluates all 512 nodes of every
    integral.  Returns the previous value."""
    global _CUT_BITS
    bits = int(bits)
    if bits > 60:
        raise ValueError("cut_bits must be <= 60")
    prev, _CUT_BITS = _CUT_BITS, max(bits, 0)
    return prev


def quadrature_cut_bits():
    return _CUT_BITS


def library_path():
    from . import _build
    return _build.LIB_PATH
