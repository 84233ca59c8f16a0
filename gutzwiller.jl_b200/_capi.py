"""ctypes binding of libgutzwiller_b200.so (include/gutzwiller_b200.h).

The library is the product; this module only declares its entry points.  If the shared
library has not been built the import fails loudly -- there is no Python or CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GWZ_LIB_PATH: developer override (kernel variants built side by side for A/B timing)
LIB_PATH = os.environ.get("GWZ_LIB_PATH") or os.path.join(_HERE, "lib", "libgutzwiller_b200.so")

MAX_WORDS = 4
NUM_PARAMS = 1
NCCL_ID_BYTES = 128
NUM_ACC = 13
NUM_BACC = 38

OK, ERR_INVALID, ERR_CUDA, ERR_NCCL, ERR_NONFINITE, ERR_NOMEM = range(6)
KERNEL_ENUMERATE, KERNEL_BITPARALLEL = 0, 1
RUN_KEEP_ACC = 1
RUN_POOLED = 2
RUN_SERIES = 4
RUN_BLOCKING = 8

ACC_NAMES = ("walkers", "sum_val", "sum_val2", "sum_grad", "sum_grad2", "sum_t", "sum_tE", "sum_tG",
             "sum_tGE", "sum_tEE", "hops", "steps", "sum_var")


class GutzwillerError(RuntimeError):
    """A CUDA / NCCL failure reported by the library."""


class BlockingSummary(C.Structure):
    """gwz_blocking_summary: val_err_and_grad(::KineticVQMCResult) on the device (state.jl:150-164)"""
    _fields_ = [("mean", C.c_double), ("err", C.c_double), ("err_err", C.c_double), ("err_ok", C.c_double),
                ("grad", C.c_double * NUM_PARAMS), ("walkers", C.c_uint64), ("failed", C.c_uint64),
                ("k_min", C.c_int32), ("k_max", C.c_int32), ("series_steps", C.c_uint64),
                ("resample_len", C.c_uint64), ("kernel_ms", C.c_float), ("reserved", C.c_float)]


class VqmcResult(C.Structure):
    _fields_ = [("val", C.c_double), ("err", C.c_double), ("grad", C.c_double * NUM_PARAMS),
                ("grad_err", C.c_double * NUM_PARAMS), ("pooled_val", C.c_double),
                ("pooled_grad", C.c_double * NUM_PARAMS), ("pooled_var", C.c_double),
                ("walker_var", C.c_double), ("total_time", C.c_double), ("walkers", C.c_uint64), ("steps", C.c_uint64),
                ("hops", C.c_uint64), ("acc", C.c_double * NUM_ACC), ("kernel_ms", C.c_float),
                ("reserved", C.c_float), ("blocking", BlockingSummary)]


class GenResult(C.Structure):
    _fields_ = [("val", C.c_double), ("err", C.c_double), ("pooled_val", C.c_double), ("pooled_var", C.c_double),
                ("total_time", C.c_double), ("walkers", C.c_uint64), ("steps", C.c_uint64), ("hops", C.c_uint64),
                ("kernel_ms", C.c_float), ("np", C.c_int32)]


ANSATZ_GUTZWILLER, ANSATZ_EXT_GUTZWILLER, ANSATZ_MULTINOMIAL, ANSATZ_DENSITY_PROFILE, ANSATZ_RELATIVE_JASTROW, \
    ANSATZ_JASTROW = range(6)
ANSATZ_COMBINATION = 6


class BlockingResult(C.Structure):
    _fields_ = [("mean", C.c_double), ("err", C.c_double), ("err_err", C.c_double), ("p_cov", C.c_double),
                ("k", C.c_int32), ("blocks", C.c_int64)]


class GdOpts(C.Structure):
    _fields_ = [("maxiter", C.c_int32), ("use_amsgrad", C.c_int32), ("early_stop", C.c_int32),
                ("reserved", C.c_int32), ("alpha", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
                ("grad_tol", C.c_double), ("param_tol", C.c_double), ("val_tol", C.c_double),
                ("err_tol", C.c_double), ("first_moment_init", C.POINTER(C.c_double)),
                ("second_moment_init", C.POINTER(C.c_double)), ("fix_params", C.POINTER(C.c_uint8))]


class GdTrace(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("converged", C.c_int32), ("reason", C.c_int32), ("np", C.c_int32),
                ("param", C.POINTER(C.c_double)), ("value", C.POINTER(C.c_double)),
                ("error", C.POINTER(C.c_double)), ("gradient", C.POINTER(C.c_double)),
                ("first_moment", C.POINTER(C.c_double)), ("second_moment", C.POINTER(C.c_double)),
                ("param_delta", C.POINTER(C.c_double))]


OBJECTIVE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double),
                           C.POINTER(C.c_double), C.POINTER(C.c_double))

_vp, _dp, _u64p, _i32p = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)

# name -> (restype, argtypes); every symbol declared in include/gutzwiller_b200.h
SIGNATURES = {
    "gwz_last_error": (C.c_char_p, []),
    "gwz_version": (C.c_int, []),
    "gwz_launch_count": (C.c_int, [_u64p]),
    "gwz_context_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "gwz_context_destroy": (C.c_int, [_vp]),
    "gwz_context_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                   C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "gwz_context_stream": (C.c_int, [_vp, C.POINTER(_vp)]),
    "gwz_context_synchronize": (C.c_int, [_vp]),
    "gwz_nccl_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "gwz_context_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(C.c_uint8)]),
    "gwz_context_comm_init_all": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "gwz_context_comm_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gwz_model_create_hubbard1d_gutzwiller": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_double,
                                                        C.POINTER(_vp)]),
    "gwz_model_destroy": (C.c_int, [_vp]),
    "gwz_model_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), _dp, _dp, C.POINTER(C.c_int),
                                 C.POINTER(C.c_int)]),
    "gwz_model_near_uniform": (C.c_int, [_vp, _u64p]),
    "gwz_model_from_onr": (C.c_int, [_vp, _i32p, _u64p]),
    "gwz_model_to_onr": (C.c_int, [_vp, _u64p, _i32p]),
    "gwz_model_dimension": (C.c_int, [_vp, _dp, _u64p]),
    "gwz_probe_addresses": (C.c_int, [_vp, _vp, C.c_size_t, _dp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gwz_ansatz_val_and_grad": (C.c_int, [_vp, _vp, C.c_size_t, _dp, _vp, _vp]),
    "gwz_walkers_create": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _u64p, C.c_uint64, C.POINTER(_vp)]),
    "gwz_walkers_destroy": (C.c_int, [_vp]),
    "gwz_walkers_count": (C.c_int, [_vp, _u64p, _u64p]),
    "gwz_walkers_get_addresses": (C.c_int, [_vp, _vp]),
    "gwz_walkers_set_addresses": (C.c_int, [_vp, _vp]),
    "gwz_walkers_set_step_counter": (C.c_int, [_vp, C.c_uint64]),
    "gwz_vqmc_run": (C.c_int, [_vp, _dp, C.c_uint64, C.c_uint64, C.c_int, C.c_uint, C.POINTER(VqmcResult)]),
    "gwz_vqmc_run_group": (C.c_int, [C.POINTER(_vp), C.c_int, _dp, C.c_uint64, C.c_uint64, C.c_int, C.c_uint,
                                     C.POINTER(VqmcResult)]),
    "gwz_walkers_estimates": (C.c_int, [_vp, _vp, _vp]),
    "gwz_vqmc_run_record": (C.c_int, [_vp, _dp, C.c_uint64, C.c_int, C.c_uint, _vp, _vp, _vp,
                                      C.POINTER(VqmcResult)]),
    "gwz_vqmc_run_trace": (C.c_int, [_vp, _dp, C.c_uint64, C.c_uint, _vp, _vp, _vp, _vp, C.POINTER(VqmcResult)]),
    "gwz_walkers_blocking": (C.c_int, [_vp, C.c_uint64, C.POINTER(BlockingSummary), _vp, _vp, _vp, _vp, _vp]),
    "gwz_walkers_series": (C.c_int, [_vp, _u64p, _vp, _vp]),
    "gwz_series_blocking": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(BlockingSummary),
                                      _vp, _vp, _vp, _vp, _vp, _vp]),
    "gwz_vqmc_sample_vector": (C.c_int, [_vp, _dp, C.c_uint64, C.c_int, C.c_uint, C.c_uint64, _vp, _vp, _u64p,
                                         C.POINTER(VqmcResult)]),
    "gwz_sweep_create": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(_vp)]),
    "gwz_sweep_destroy": (C.c_int, [_vp]),
    "gwz_sweep_dim": (C.c_int, [_vp, _u64p]),
    "gwz_sweep_val": (C.c_int, [_vp, _dp, _dp]),
    "gwz_sweep_val_and_grad": (C.c_int, [_vp, _dp, _dp, _dp, _dp]),
    "gwz_sweep_terms": (C.c_int, [_vp, _dp, C.c_uint64, C.c_uint64, _vp]),
    "gwz_sweep_states": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp]),
    "gwz_sweep_representative_range": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64),
                                                 C.POINTER(C.c_uint64), C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]),
    "gwz_sweep_last_kernel_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "gwz_resample": (C.c_int, [_vp, C.c_size_t, _vp, _vp, C.c_size_t]),
    "gwz_blocking_analysis": (C.c_int, [_vp, C.c_size_t, C.POINTER(BlockingResult)]),
    "gwz_gd_opts_default": (C.c_int, [C.POINTER(GdOpts)]),
    "gwz_adaptive_gradient_descent": (C.c_int, [OBJECTIVE_FN, OBJECTIVE_FN, _vp, _dp, C.c_int, C.POINTER(GdOpts),
                                                C.POINTER(GdTrace)]),
    "gwz_amsgrad_vqmc": (C.c_int, [_vp, _dp, C.POINTER(GdOpts), C.c_uint64, C.c_uint64, C.c_int, C.c_uint,
                                   C.POINTER(GdTrace), C.POINTER(VqmcResult)]),
    "gwz_amsgrad_sweep": (C.c_int, [_vp, _dp, C.POINTER(GdOpts), C.POINTER(GdTrace)]),
    # ---- ansatz-generic engine ----
    "gwz_model_create_ext_hubbard1d": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                                 C.POINTER(_vp)]),
    "gwz_model_info_ext": (C.c_int, [_vp, _dp]),
    "gwz_ansatz_create": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_vp)]),
    "gwz_ansatz_create_combination": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "gwz_ansatz_destroy": (C.c_int, [_vp]),
    "gwz_ansatz_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gwz_gen_ansatz_eval": (C.c_int, [_vp, _vp, C.c_size_t, _dp, _vp, _vp]),
    "gwz_gen_probe_addresses": (C.c_int, [_vp, _vp, C.c_size_t, _dp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gwz_gwalkers_create": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _u64p, C.c_uint64, C.POINTER(_vp)]),
    "gwz_gwalkers_destroy": (C.c_int, [_vp]),
    "gwz_gwalkers_count": (C.c_int, [_vp, _u64p, _u64p]),
    "gwz_gwalkers_get_addresses": (C.c_int, [_vp, _vp]),
    "gwz_gwalkers_set_addresses": (C.c_int, [_vp, _vp]),
    "gwz_gwalkers_set_step_counter": (C.c_int, [_vp, C.c_uint64]),
    "gwz_gen_vqmc_run": (C.c_int, [_vp, _dp, C.c_uint64, C.c_uint64, C.c_uint, C.POINTER(GenResult), _vp, _vp, _vp]),
    "gwz_gen_vqmc_run_record": (C.c_int, [_vp, _dp, C.c_uint64, C.c_uint, _vp, _vp, _vp, C.POINTER(GenResult), _vp,
                                          _vp, _vp]),
    "gwz_gwalkers_estimates": (C.c_int, [_vp, _vp, _vp]),
    "gwz_gsweep_create": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(_vp)]),
    "gwz_gsweep_destroy": (C.c_int, [_vp]),
    "gwz_gsweep_dim": (C.c_int, [_vp, _u64p]),
    "gwz_gsweep_val_and_grad": (C.c_int, [_vp, _dp, _dp, _vp]),
    "gwz_gsweep_terms": (C.c_int, [_vp, _dp, C.c_uint64, C.c_uint64, _vp]),
    "gwz_gsweep_states": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp]),
    "gwz_gsweep_last_kernel_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    # ---- VectorAnsatz ----
    "gwz_vector_ansatz_create": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.POINTER(_vp)]),
    "gwz_vector_ansatz_destroy": (C.c_int, [_vp]),
    "gwz_vector_ansatz_lookup": (C.c_int, [_vp, _vp, C.c_size_t, _vp]),
    "gwz_vector_ansatz_rayleigh": (C.c_int, [_vp, _dp]),
    "gwz_vwalkers_create": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _u64p, C.c_uint64, C.POINTER(_vp)]),
    "gwz_vwalkers_destroy": (C.c_int, [_vp]),
    "gwz_vwalkers_get_addresses": (C.c_int, [_vp, _vp]),
    "gwz_vwalkers_run": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint, C.POINTER(GenResult)]),
    "gwz_vwalkers_run_record": (C.c_int, [_vp, C.c_uint64, C.c_uint, _vp, _vp, _vp, C.POINTER(GenResult)]),
    # ---- stored operators (any Hamiltonian, any ansatz) ----
    "gwz_sparse_create": (C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    "gwz_sparse_destroy": (C.c_int, [_vp]),
    "gwz_sparse_info": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                  C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "gwz_sparse_layout": (C.c_int, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]),
    "gwz_sparse_set_table": (C.c_int, [_vp, _vp, _vp, C.c_int32]),
    "gwz_sparse_set_gutzwiller": (C.c_int, [_vp, C.c_double]),
    "gwz_sparse_get_table": (C.c_int, [_vp, _vp, _vp]),
    "gwz_sparse_probe": (C.c_int, [_vp, _vp, C.c_size_t, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gwz_sparse_sweep": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _dp, _vp]),
    "gwz_sparse_sweep_terms": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp]),
    "gwz_sparse_last_kernel_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "gwz_swalkers_create": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.POINTER(_vp)]),
    "gwz_swalkers_destroy": (C.c_int, [_vp]),
    "gwz_swalkers_get_indices": (C.c_int, [_vp, _vp]),
    "gwz_swalkers_set_indices": (C.c_int, [_vp, _vp]),
    "gwz_swalkers_set_step_counter": (C.c_int, [_vp, C.c_uint64]),
    "gwz_swalkers_run": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint, C.POINTER(GenResult), _vp, _vp, _vp]),
    "gwz_swalkers_run_record": (C.c_int, [_vp, C.c_uint64, C.c_uint, _vp, _vp, _vp, C.POINTER(GenResult), _vp, _vp, _vp]),
    "gwz_swalkers_sample_vector": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint, _vp, C.POINTER(GenResult)]),
    "gwz_swalkers_estimates": (C.c_int, [_vp, _vp, _vp]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C gutzwiller.jl_b200/csrc`. There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    """Map a gwz_status to the exception the reference would raise."""
    if rc == OK:
        return
    msg = lib().gwz_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID:
        raise ValueError(msg)  # Julia: ArgumentError (state.jl:195-197)
    if rc == ERR_NONFINITE:
        raise FloatingPointError(msg)  # Julia: error("Infinite residence time ...") (qmc.jl:34-36)
    if rc == ERR_NOMEM:
        raise MemoryError(msg)
    raise GutzwillerError(f"[status {rc}] {msg}")
