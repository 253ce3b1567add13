"""ctypes binding of libdpe_b200.so (include/dpe_b200.h). There is no CPU fallback: if the library is
missing this module raises, it never routes anywhere else."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "csrc" / "libdpe_b200.so"
DPE_MAX_ITER = 8
MODE_FORWARD, MODE_LAPLACIAN = 0, 1
STAGE_NAMES = ("features", "el_ion_stream", "pair_stream", "h_map", "schnet_conv", "spin_mean", "mean_term_gemm", "main_layer",
               "orbitals", "det_factor", "det_trace", "combine", "mcmc")


class DpeDims(C.Structure):
    _fields_ = [("n_el", C.c_int32), ("n_up", C.c_int32), ("n_ion", C.c_int32), ("n_iterations", C.c_int32),
                ("n_hidden_one_el", C.c_int32 * DPE_MAX_ITER), ("n_hidden_two_el", C.c_int32 * DPE_MAX_ITER),
                ("emb_dim", C.c_int32), ("n_ion_features", C.c_int32), ("n_dets", C.c_int32),
                ("z_min", C.c_int32), ("z_max", C.c_int32), ("use_taos", C.c_int32)]


class DpeMcmcConfig(C.Structure):
    _fields_ = [("max_age", C.c_int32), ("stepsize_update_interval", C.c_int32), ("target_acceptance_rate", C.c_float),
                ("min_stepsize_scale", C.c_float), ("max_stepsize_scale", C.c_float), ("proposal", C.c_int32),
                ("r_min", C.c_float), ("r_max", C.c_float), ("langevin_scale", C.c_float)]


class DpeXlaDescriptor(C.Structure):
    """include/dpe_b200.h dpe_xla_descriptor: the `opaque` of the XLA custom calls."""
    _fields_ = [("model", C.c_uint64), ("workspace_bytes", C.c_uint64), ("n_walkers", C.c_int32), ("n_steps", C.c_int32),
                ("recompute_log_psi", C.c_int32), ("run_controller", C.c_int32), ("mcmc", DpeMcmcConfig)]


class DpeMcmcState(C.Structure):
    _fields_ = [("r_dev", C.c_void_p), ("log_psi_sqr_dev", C.c_void_p), ("walker_age_dev", C.c_void_p),
                ("rng_state_dev", C.c_void_p), ("stepsize_dev", C.c_void_p), ("step_nr_dev", C.c_void_p),
                ("acc_rate_dev", C.c_void_p)]


# every exported symbol of include/dpe_b200.h: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "dpe_version": (C.c_char_p, []),
    "dpe_last_error": (C.c_char_p, []),
    "dpe_model_create": (C.c_int, [C.POINTER(DpeDims), C.POINTER(_P)]),
    "dpe_model_destroy": (None, [_P]),
    "dpe_param_count": (C.c_int64, [_P]),
    "dpe_param_leaf_count": (C.c_int32, [_P]),
    "dpe_param_leaf": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "dpe_model_set_params": (C.c_int, [_P, _P, C.c_int64, _P]),
    "dpe_model_set_geometry": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int32), _P]),
    "dpe_model_set_geometry_dev": (C.c_int, [_P, _P, _P, _P]),
    "dpe_model_geometry_status": (C.c_int, [_P, _P]),
    "dpe_model_set_tao_cache": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "dpe_workspace_bytes": (C.c_size_t, [_P, C.c_int32, C.c_int32]),
    "dpe_log_psi_sqr": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, C.c_size_t, _P]),
    "dpe_local_energy": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "dpe_mcmc_steps": (C.c_int, [_P, C.POINTER(DpeMcmcState), C.c_int32, C.c_int32, C.POINTER(DpeMcmcConfig), C.c_int32, C.c_int32, _P, _P, C.c_size_t, _P]),
    "dpe_mcmc_controller": (C.c_int, [C.POINTER(DpeMcmcState), _P, C.c_int32, C.c_int64, C.POINTER(DpeMcmcConfig), _P]),
    "dpe_energy_moments1": (C.c_int, [_P, C.c_int32, _P, C.c_int32, _P, _P, _P]),
    "dpe_energy_moments2": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P]),
    "dpe_energy_median": (C.c_int, [_P, C.c_int32, _P, _P]),
    "dpe_energy_width": (C.c_int, [_P, C.c_int32, _P, C.c_int32, _P, _P]),
    "dpe_set_mcmc_graph": (C.c_int, [_P, C.c_int32]),
    "dpe_get_mcmc_graph": (C.c_int, [_P]),
    "dpe_kfac_layer_count": (C.c_int32, [_P]),
    "dpe_kfac_floats": (C.c_int64, [_P]),
    "dpe_kfac_layer": (C.c_int, [_P, C.c_int32, C.c_char_p, C.c_int32] + [C.POINTER(C.c_int32)] * 4 + [C.POINTER(C.c_int64)] * 2),
    "dpe_gradient_workspace_bytes": (C.c_size_t, [_P, C.c_int32]),
    "dpe_param_gradient": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "dpe_xla_log_psi_sqr": (None, [_P, C.POINTER(_P), C.c_char_p, C.c_size_t]),
    "dpe_xla_local_energy": (None, [_P, C.POINTER(_P), C.c_char_p, C.c_size_t]),
    "dpe_xla_mcmc_steps": (None, [_P, C.POINTER(_P), C.c_char_p, C.c_size_t]),
    "dpe_xla_last_status": (C.c_int, []),
    "dpe_threefry_mcmc_randoms": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "dpe_threefry_bits": (C.c_int, [C.POINTER(C.c_uint32), C.c_int32, _P, _P]),
    "dpe_threefry_normal": (C.c_int, [C.POINTER(C.c_uint32), C.c_int32, _P, _P]),
    "dpe_debug_ws_offset": (C.c_int64, [_P, C.c_int32, C.c_int32, C.c_char_p]),
    "dpe_set_gemm_path": (C.c_int, [_P, C.c_int32]),
    "dpe_debug_gemm": (C.c_int, [_P, C.c_int32, _P, C.c_int32, _P, _P, C.c_int32] + [C.c_int32] * 9 + [_P]),
    "dpe_get_gemm_path": (C.c_int, [_P]),
    "dpe_set_det_path": (C.c_int, [_P, C.c_int32]),
    "dpe_get_det_path": (C.c_int, [_P]),
    "dpe_profile_enable": (C.c_int, [_P, C.c_int32]),
    "dpe_profile_collect": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "dpe_profile_launches": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int32, C.POINTER(C.c_int32)]),
    "dpe_profile_stages": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32]),
    "dpe_launch_count": (C.c_int64, [_P]),
}

_lib = None


class DpeError(RuntimeError):
    pass


def load():
    """Loads the CUDA library or raises. No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(needs nvcc). deeperwin_b200 has no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().dpe_last_error().decode()
        raise DpeError(f"{what} failed with status {rc}: {msg}")
