"""ctypes binding of libdraupnir_b200.so (C ABI in include/draupnir_b200.h).

There is deliberately no fallback: if the shared library is missing or a symbol is absent the import of any product
module fails loudly.  `build.py` (called by `__graft_entry__.build()`) produces the library in-tree.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DRP_LIB: another build of the same ABI (profiles/build_debug.sh: the library with the tcgen05 kernel's debug hooks)
LIB_PATH = os.environ.get("DRP_LIB") or os.path.join(_HERE, "libdraupnir_b200.so")

ABI_VERSION = 5

_i32, _i64, _f64, _p = ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/draupnir_b200.h one to one
SIGNATURES = {
    "drp_abi_version": (_i32, []),
    "drp_launch_count": (_i64, []),
    "drp_timing_classes": (_i32, []),
    "drp_timing_class_name": (ctypes.c_char_p, [_i32]),
    "drp_timing_class_is_flops": (_i32, [_i32]),
    "drp_timing_enable": (_i32, [_i32]),
    "drp_timing_read": (_i32, [_p, _p, _p]),
    "drp_tree_prepare": (_i32, [_p, _p, _i32, _p, _p, _p, _p, _p, _p]),
    "drp_lca_workspace_bytes": (_i64, [_i32, _i32]),
    "drp_lca_patristic_f64": (_i32, [_p, _p, _p, _p, _p, _i32, _i32, _p, _i32, _p, _i32, _p, _p, _p, _p, _i64, _p]),
    "drp_ou_cov_fwd_f64": (_i32, [_p, _i64, _i32, _p, _p, _p, _i32, _p, _p]),
    "drp_ou_cov_bwd_f64": (_i32, [_p, _i64, _i32, _p, _i32, _p, _p, _p, _p]),
    "drp_factor_workspace_bytes": (_i64, [_i32, _i32]),
    "drp_ozaki_cond_limit_f64": (_f64, []),
    "drp_potrf_batched_f64": (_i32, [_p, _i32, _i32, _i32, _p, _p, _p, _i64, _p, _p, _p, _p]),
    "drp_syrk_update_f64": (_i32, [_p, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _i64, _p]),
    "drp_syrk_update_range_f64": (_i32, [_p, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _i64, _i32, _i32, _p]),
    "drp_ou_workspace_ld": (_i64, [_i32, _i32]),
    "drp_ou_workspace_elems": (_i64, [_i32, _i32, _i32]),
    "drp_ou_prior_f64": (_i32, [_p, _i64, _i32, _p, _p, _p, _p, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "drp_ou_prior_factor_f64": (_i32, [_p, _i64, _i32, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _p]),
    "drp_ou_prior_apply_workspace_elems": (_i64, [_i32, _i32]),
    "drp_ou_prior_apply_f64": (_i32, [_p, _i64, _i32, _p, _p, _i32, _p, _p, _p, _p, _p, _p, _p]),
    "drp_mvn_logprob_f64": (_i32, [_p, _p, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "drp_mvn_grad_cov_f64": (_i32, [_p, _i32, _i32, _p, _p, _p]),
    "drp_potri_batched_f64": (_i32, [_p, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p]),
    "drp_ou_conditional_f64": (_i32, [_p, _i64, _p, _i32, _i32, _p, _p, _p, _p, _i32, _f64, _p, _p, _p, _p, _p, _p, _p, _p]),
    "drp_gru_hidden_size": (_i32, []),
    "drp_gru_fwd_f64": (_i32, [_p, _p, _p, _p, _i32, _i32, _i32, _p, _p, _p]),
    "drp_gru_fwd_dir0_f64": (_i32, [_p, _p, _p, _p, _i32, _i32, _i32, _p, _p, _p]),
    "drp_gru_fwd_uv_f64": (_i32, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _p, _p, _p]),
    "drp_gru_bwd_f64": (_i32, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _p, _p, _p]),
    "drp_gru_bwd_dir0_f64": (_i32, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _p, _p, _p]),
    "drp_gru_bwd_last_dir0_f64": (_i32, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _p, _p, _p]),
    "drp_gru_input_sums_blocks": (_i32, [_i32]),
    "drp_gru_input_sums_chunks": (_i32, [_i32, _i32]),
    "drp_gru_input_sums_f64": (_i32, [_p, _i32, _i32, _i32, _p, _p, _p]),
    "drp_gru_wgrad_dir0_f64": (_i32, [_p, _p, _p, _p, _i32, _i32, _i32, _p, _p, _p, _i64, _p]),
    "drp_gru_wgrad_workspace_elems": (_i64, []),
    "drp_gru_wgrad_f64": (_i32, [_p, _p, _p, _p, _i32, _i32, _i32, _p, _p, _p, _i64, _p]),
    "drp_gram_workspace_elems": (_i64, [_i32, _i32]),
    "drp_gram_tn_supported": (_i32, [_i32, _i32]),
    "drp_gram_tn_f64": (_i32, [_p, _p, _i64, _i32, _i32, _p, _p, _p, _i64, _p]),
    "drp_tall_gemm_f64": (_i32, [_p, _p, _p, _i64, _i32, _i32, _p, _p]),
    "drp_site_logp_workspace_elems": (_i64, []),
    "drp_site_logp_f64": (_i32, [_p, _p, _p, _p, _i64, _i64, _i32, _p, _p, _p]),
    "drp_site_logp_bwd_f64": (_i32, [_p, _p, _p, _p, _i64, _i64, _i32, _p, _p, _p, _p, _p]),
    "drp_clipped_adam_max_tensors": (_i32, []),
    "drp_clipped_adam_f64": (_i32, [_p, _p, _p, _i32, _i32, _p, _f64, _f64, _f64, _f64, _f64, _f64, _p]),
    "drp_cat_loglik_partials": (_i64, [_i64]),
    "drp_cat_loglik_fwd_f64": (_i32, [_p, _p, _i32, _i64, _i32, _p, _p, _p, _p]),
    "drp_cat_loglik_bwd_f64": (_i32, [_p, _p, _i32, _p, _p, _i64, _i32, _p, _p]),
}


class DraupnirB200LibraryError(ImportError):
    pass


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise DraupnirB200LibraryError(
            f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise DraupnirB200LibraryError(f"{LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype, fn.argtypes = res, args
    for name in ("drp_ozaki_set_trace", "drp_ozaki_set_prof"):      # only a -DDRP_DEBUG_HOOKS build has them
        if hasattr(lib, name):
            getattr(lib, name).restype, getattr(lib, name).argtypes = None, [_p]
    v = lib.drp_abi_version()
    if v != ABI_VERSION:
        raise DraupnirB200LibraryError(f"{LIB_PATH} has ABI {v}, this package expects {ABI_VERSION}; rebuild it")
    return lib


lib = _load()
