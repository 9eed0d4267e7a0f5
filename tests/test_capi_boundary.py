"""The C-ABI boundary without a GPU: the library loads, exports exactly what
include/semiuhpe_b200.h declares, the ctypes table matches the header, and the host
side refuses to run without CUDA (no CPU fallback)."""
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "semiuhpe_b200.h")
PROBE_HEADER = os.path.join(ROOT, "include", "semiuhpe_b200_probe.h")


def declared_functions(header=HEADER):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = re.findall(r"\b(suhpe_[a-z0-9_]+)\s*\(([^;{]*)\)\s*;", text)
    return {name: args for name, args in decls}


def test_header_symbols_are_exported(built):
    from semiuhpe_b200 import _build
    out = subprocess.run(["nm", "-D", "--defined-only", _build.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (suhpe_[a-z0-9_]+)", out))
    declared = set(declared_functions())
    probes = set(declared_functions(PROBE_HEADER))
    assert declared and probes, "no declarations parsed"
    assert not (declared & probes), "a measurement entry point leaked into the drop-in header"
    assert "suhpe_fp32_probe" in probes and not any("probe" in name for name in declared)
    declared |= probes
    assert declared == exported, f"header-only: {declared - exported}, library-only: {exported - declared}"


def test_ctypes_table_matches_header(built):
    from semiuhpe_b200 import _capi
    decl = declared_functions()
    assert set(_capi.SIGNATURES) == set(decl)
    probe_decl = declared_functions(PROBE_HEADER)
    assert set(_capi.PROBE_SIGNATURES) == set(probe_decl)
    for table, d in ((_capi.SIGNATURES, decl), (_capi.PROBE_SIGNATURES, probe_decl)):
        for name, args in d.items():
            n_args = 0 if args.strip() in ("", "void") else len(args.split(","))
            assert len(table[name][1]) == n_args, name
    handle = _capi.lib()                      # dlopen works without a GPU
    assert handle.suhpe_abi_version() == _capi.ABI_VERSION == 2
    assert not hasattr(handle, "suhpe_set_quadrature_cut_bits")     # no process-wide settings: cut_bits is per call
    assert handle.suhpe_error_string(0) == b"ok"
    assert handle.suhpe_error_string(_capi.EINVAL) == b"invalid argument"
    # argument validation happens before any CUDA call
    assert handle.suhpe_fisher_fused_f32(None, None, 5, 1.0, 26, None, None, None, None, None, None, None, None, None, None) == _capi.EINVAL
    assert handle.suhpe_scale_rows_f32(None, 3, 9, None, None, None, None, None) == _capi.EINVAL
    assert handle.suhpe_scale_rows_f32(None, 0, 9, None, None, None, None, None) == 0
    assert handle.suhpe_ssl_step_f32(None, None, None, 1, None, None, 0, None, 0, None, 0.0, 1.0, 1.0, 0, 26, None,
                                     None, None, None, None, None, None, None, None, None, None, None) == _capi.EINVAL
    # scratch of the one-call SSL step: every region a multiple of 4 floats, so the carving stays 16-byte aligned
    assert handle.suhpe_ssl_step_workspace_floats(32, 128) == 4 + 32 + 128 + 1280 + 1152 + 1152 + 128 + 32
    assert handle.suhpe_ssl_step_workspace_floats(5, 0) == 4 + 8 and handle.suhpe_ssl_step_workspace_floats(1, 1) % 4 == 0
    assert handle.suhpe_entropy_threshold_f32(None, 0, 0, None, None, None, None) == _capi.EINVAL
    assert handle.suhpe_laplace_nll_f32(None, None, 1, None, 0, None, None, None, None, None, None) == _capi.EINVAL
    assert handle.suhpe_ema_update_f32(None, None, None, 3, 0.5, 0.5, 0, None) == _capi.EINVAL
    assert handle.suhpe_ema_update_f32(None, None, None, 0, 0.5, 0.5, 2, None) == _capi.EINVAL
    assert handle.suhpe_ema_update_f32(None, None, None, 0, 0.5, 0.5, 1, None) == 0        # empty list: nothing to launch


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (and as C++) on its own, warning-free."""
    src = tmp_path / "use_header.c"
    src.write_text('#include "semiuhpe_b200.h"\n#include "semiuhpe_b200_probe.h"\nint main(void) { return suhpe_abi_version() < 0; }\n')
    inc = os.path.join(ROOT, "include")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)],
                ["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)]):
        proc = subprocess.run(cmd, capture_output=True, text=True)
        assert proc.returncode == 0, proc.stderr


def test_c_host_links_and_calls_the_library(built, tmp_path):
    """A plain C host program linked against the shared library (what a non-Python host does): version,
    error strings and argument validation work without a GPU."""
    from semiuhpe_b200 import _build
    src = tmp_path / "host.c"
    src.write_text(
        '#include <stdio.h>\n#include "semiuhpe_b200.h"\n'
        'int main(void) {\n'
        '  printf("%d|%s|%s|", suhpe_abi_version(), suhpe_error_string(0), suhpe_error_string(SUHPE_EINVAL));\n'
        '  printf("%d|", suhpe_fisher_fused_f32(NULL, NULL, 4, 1.0f, SUHPE_CUT_BITS_DEFAULT, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL) == SUHPE_EINVAL);\n'
        '  printf("%d\\n", suhpe_ema_update_f32(NULL, NULL, NULL, 0, 0.5f, 0.5f, 1, NULL));\n'
        '  return 0;\n}\n')
    exe = tmp_path / "host"
    lib_dir = os.path.dirname(_build.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", lib_dir, "-lsemiuhpe_b200", f"-Wl,-rpath,{lib_dir}"], check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip()
    assert out == "2|ok|invalid argument|1|0", out


def test_sass_is_sm100a_only(built):
    from semiuhpe_b200 import _build
    out = subprocess.run(["cuobjdump", "-lelf", _build.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback(built):
    """CPU tensors are an error, not a slow path."""
    from semiuhpe_b200.fisher.fisher_utils import vmf_loss, fisher_entropy, batch_torch_A_to_R, fisher_CE
    from semiuhpe_b200.laplace.rotation_laplace import NLL_loss
    from semiuhpe_b200.agent import compute_err_deg_from_matrices, entropy_threshold, entropy_mask
    from semiuhpe_b200.agent import update_ema_variables
    with pytest.raises(RuntimeError):
        update_ema_variables(torch.nn.Linear(3, 2), torch.nn.Linear(3, 2), True, 0.999, 10)
    A, R = torch.randn(4, 9), torch.eye(3).repeat(4, 1, 1)
    for call in (lambda: vmf_loss(A, R), lambda: fisher_entropy(A), lambda: batch_torch_A_to_R(A),
                 lambda: NLL_loss("RLaplace", A, R, R), lambda: compute_err_deg_from_matrices(R, R),
                 lambda: fisher_CE(A, A),
                 lambda: entropy_threshold(torch.randn(10), 0.5), lambda: entropy_mask(torch.randn(10), 0.0)):
        with pytest.raises(RuntimeError, match="CUDA"):
            call()
    with pytest.raises(NotImplementedError):              # the target is a constant (src/agent.py:107)
        fisher_CE(A.clone().requires_grad_(True), A)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under semiuhpe_b200/ may reference it."""
    pkg = os.path.join(ROOT, "semiuhpe_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "ref_shim" not in text and "/root/reference" not in text, f


def test_pool_index_semantics():
    from semiuhpe_b200.agent import pool_index
    assert pool_index(128, 0.95) == 121
    assert pool_index(2 ** 26, 0.95) == 63753420
    assert pool_index(64000000, 0.95) == 60800000
    assert pool_index(2 ** 26, 0.75) == 50331648
    assert pool_index(10, 0.0) == 0
    with pytest.raises(IndexError):
        pool_index(128, 1.0)
    with pytest.raises(IndexError):
        pool_index(0, 0.5)


def test_host_helpers_match_oracle(golden):
    from oracle import so3_oracle as orc
    from semiuhpe_b200 import utils
    import numpy as np
    g = golden("metrics")
    assert [utils.limit_angle(a) for a in g["limit_in"]] == list(g["limit_out"])
    np.testing.assert_allclose(utils.get_6DRepNet_Rot(0.3, -0.2, 1.1), orc.rot_from_euler(0.3, -0.2, 1.1), atol=1e-15)
    # rot_euler_6DRepNet (src/utils.py:263-286): the per-sample numpy twin of the batched Euler function
    import torch
    rng = np.random.default_rng(3)
    for full_range in (False, True):
        for _ in range(50):
            p, y, r = rng.uniform(-np.pi, np.pi, 3) * np.array([0.49, 0.99 if full_range else 0.49, 0.99])
            R = utils.get_6DRepNet_Rot(p, y, r)
            want = orc.euler_from_matrices(torch.from_numpy(R)[None], full_range=full_range)[0].numpy()
            np.testing.assert_allclose(utils.rot_euler_6DRepNet(R, full_range), want, atol=1e-12)
    Rs = np.array([[0.0, 0.0, 1.0], [0.0, 1.0, 0.0], [-1.0, 0.0, 0.0]])       # gimbal lock: sy = 0
    np.testing.assert_allclose(utils.rot_euler_6DRepNet(Rs), [0.0, np.pi / 2, 0.0], atol=1e-12)
