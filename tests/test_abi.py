"""The C-ABI library builds, loads, and exports every symbol include/draupnir_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "draupnir_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(drp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from draupnir_asr_b200 import _lib
    names = _declared()
    assert len(names) >= 17
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} declared in the header but not bound in _lib.SIGNATURES"
    assert set(_lib.SIGNATURES) == set(names), "bindings and header disagree"
    assert _lib.lib.drp_abi_version() == _lib.ABI_VERSION


def test_argument_validation_without_a_gpu():
    """Bad-argument paths return before any CUDA call, so they are checkable on a CPU-only box."""
    from draupnir_asr_b200._lib import lib
    assert lib.drp_potrf_batched_f64(None, 4, 1, 0, None, None, None, 0, None, None, None, None) == -1
    assert lib.drp_syrk_update_f64(None, 8, 64, 1, 0, 128, 128, 512, 512, 6, None, 0, None) == -1
    assert lib.drp_ou_cov_fwd_f64(None, 4, 4, None, None, None, 1, None, None) == -1
    assert lib.drp_cat_loglik_fwd_f64(None, None, 1, 10, 21, None, None, None, None) == -1
    assert lib.drp_ou_workspace_ld(19, 19) == 40 and lib.drp_ou_workspace_ld(2000, 2000) == 4008
    assert lib.drp_ou_workspace_elems(19, 0, 30) == 30 * 20 * 24 + (lib.drp_factor_workspace_bytes(20, 30) + 7) // 8
    # two slice buffers + the precision gate's index lists + the block inverses of the fused panel kernels (4 x 64 x 64 per matrix)
    assert lib.drp_factor_workspace_bytes(2000, 30) == (2 * ((30 * 2176 * (12 * 128 + 8) + 2048 + 1023) // 1024 * 1024) + 256
                                                        + 30 * 4 * 64 * 64 * 8 + 256)
    assert lib.drp_ozaki_cond_limit_f64() == 1e-5 * 2.0 ** 35
    assert lib.drp_timing_classes() == 21 and lib.drp_timing_class_name(8) == b"oz_syrk_kernel"
    assert lib.drp_timing_class_name(19) == b"potrf_diag_kernel" and lib.drp_timing_class_name(20) == b"panel_solve_kernel"
    # every weight-gradient shape of the DRAUPNIR networks (M = rows of dY, N = rows of X) is taken by the split-K kernel
    for M, N in ((180, 21), (21, 120), (21, 21), (120, 21), (21, 180), (60, 60), (30, 60), (1, 1), (64, 64)):
        assert lib.drp_gram_tn_supported(M, N) == 1, (M, N)
    for M, N in ((193, 21), (21, 193), (192, 192), (0, 5)):
        assert lib.drp_gram_tn_supported(M, N) == 0, (M, N)
    assert lib.drp_gru_hidden_size() == 60
    assert lib.drp_gru_fwd_f64(None, None, None, None, 4, 4, 60, None, None, None) == -1
    assert lib.drp_cat_loglik_partials(257) == 2
    assert lib.drp_lca_workspace_bytes(10, 100) == 0 and lib.drp_lca_workspace_bytes(10, 2000) == 10 * 2001 * 24


def test_sources_are_sm100a_only():
    from draupnir_asr_b200 import build
    assert "arch=compute_100a,code=sm_100a" in build.NVCC_FLAGS and "-lineinfo" in build.NVCC_FLAGS
