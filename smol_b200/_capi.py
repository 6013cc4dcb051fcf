"""ctypes binding of ``liblmc.so`` (declared in ``include/lmc.h``).

The product path has no CPU fallback: if the CUDA library is missing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

LMC_ABI_VERSION = 16
LMC_MAX_CLUSTER_SITES = 4
LMC_MAX_SUBLATTICES = 8
LMC_MAX_CODES = 8
LMC_MAX_FLIPS = 4
LMC_MAX_DIMS = 16
LMC_MAX_TABLE_FLIPS = 8
LMC_MAX_COMPOSITE = 4
LMC_MAX_BIAS_ROWS = 4
LMC_USHER_FLIP, LMC_USHER_SWAP, LMC_USHER_TABLEFLIP, LMC_USHER_COMPOSITE, LMC_USHER_MULTISTEP = 0, 1, 2, 3, 4
LMC_KERNEL_METROPOLIS, LMC_KERNEL_WANGLANDAU = 0, 1
LMC_BIAS_NONE, LMC_BIAS_TABLE_SUM, LMC_BIAS_SQUARE_SUM = 0, 1, 2

_P = C.c_void_p


class LmcModelDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("num_sites", C.c_int32),
        ("num_features", C.c_int32),
        ("num_ce_features", C.c_int32),
        ("supercell_size", C.c_int32),
        ("feature0", C.c_double),
        ("natural_parameters", _P),
        ("num_orbits", C.c_int32),
        ("orb_tab_off", _P),
        ("orb_tab_len", _P),
        ("orb_nfunc", _P),
        ("orb_fidx", _P),
        ("orb_csize", _P),
        ("orb_stride", _P),
        ("orb_weight", _P),
        ("ftab", _P),
        ("ftab_len", C.c_int64),
        ("orb_row_off", _P),
        ("full_rows", _P),
        ("num_classes", C.c_int32),
        ("cls_orbit", _P),
        ("cls_stride", _P),
        ("site_rec_off", _P),
        ("site_rec", _P),
        ("site_seg_off", _P),
        ("site_seg", _P),
        ("ewald_size", C.c_int32),
        ("ewald_width", C.c_int32),
        ("ewald_matrix", _P),
        ("ewald_inds", _P),
        ("ewald_feature", C.c_int32),
        ("mu_width", C.c_int32),
        ("mu_table", _P),
        ("mu_feature", C.c_int32),
        ("num_sublattices", C.c_int32),
        ("sl_site_off", _P),
        ("sl_sites", _P),
        ("sl_ncodes", _P),
        ("sl_codes", _P),
        ("sl_prob", _P),
        ("tf_num_dims", C.c_int32),
        ("tf_num_flips", C.c_int32),
        ("tf_table", _P),
        ("tf_weights", _P),
        ("tf_max_n", _P),
        ("tf_dim_sl", _P),
        ("tf_dim_code", _P),
        ("tf_swap_weight", C.c_double),
    ]


class LmcWangLandau(C.Structure):
    _fields_ = [
        ("min_enthalpy", C.c_double), ("max_enthalpy", C.c_double), ("bin_size", C.c_double),
        ("flatness", C.c_double), ("mod_update", C.c_double),
        ("num_bins", C.c_int32), ("check_period", C.c_int32), ("update_period", C.c_int32),
        ("reserved", C.c_int32),
        ("entropy_dev", _P), ("histogram_dev", _P), ("occurrences_dev", _P),
        ("mean_features_dev", _P), ("mod_factor_dev", _P), ("steps_counter_dev", _P),
        ("trace_entropy_dev", _P), ("trace_histogram_dev", _P), ("trace_occurrences_dev", _P),
        ("trace_mean_features_dev", _P), ("trace_mod_factor_dev", _P),
        ("mod_table_dev", _P), ("mod_table_len", C.c_int32), ("reserved2", C.c_int32),
    ]


class LmcRunConfig(C.Structure):
    _fields_ = [
        ("num_walkers", C.c_int32), ("walker_id_base", C.c_int32), ("usher", C.c_int32),
        ("kernel", C.c_int32), ("num_samples", C.c_int64), ("thin_by", C.c_int32),
        ("group_size", C.c_int32), ("block_threads", C.c_int32), ("spec_mode", C.c_int32),
        ("step_begin", C.c_uint64), ("seeds_dev", _P), ("beta_dev", _P),
        ("occ_dev", _P), ("features_dev", _P), ("enthalpy_dev", _P),
        ("trace_occ_dev", _P), ("trace_features_dev", _P), ("trace_enthalpy_dev", _P),
        ("trace_accepted_dev", _P), ("trace_naccepted_dev", _P), ("ewald_field_dev", _P),
        ("bias_mode", C.c_int32), ("bias_width", C.c_int32), ("bias_rows", C.c_int32), ("bias_penalty", C.c_double),
        ("bias_table_dev", _P), ("bias_dev", _P), ("bias_sum_dev", _P), ("trace_bias_dev", _P),
        ("dist_mode", C.c_int32), ("dist_num_groups", C.c_int32), ("dist_tol", C.c_double),
        ("dist_target_dev", _P), ("dist_group_off_dev", _P), ("dist_group_idx_dev", _P),
        ("dist_group_diam_dev", _P), ("dist_vector_dev", _P),
        ("comp_num", C.c_int32), ("comp_usher", C.c_int32 * LMC_MAX_COMPOSITE),
        ("comp_cum", C.c_double * LMC_MAX_COMPOSITE),
        ("comp_sl_cum", (C.c_double * LMC_MAX_SUBLATTICES) * LMC_MAX_COMPOSITE),
        ("ms_usher", C.c_int32), ("ms_num", C.c_int32), ("ms_len", C.c_int32 * LMC_MAX_COMPOSITE),
        ("ms_cum", C.c_double * LMC_MAX_COMPOSITE),
        ("wl", LmcWangLandau),
        ("walker_mask_dev", _P), ("accept_offset_dev", _P), ("spec_env_dev", _P),
    ]


EXPORTS = (
    "lmc_version", "lmc_last_error", "lmc_row_stride", "lmc_model_create", "lmc_model_destroy",
    "lmc_model_num_features", "lmc_cast_i32_to_i8", "lmc_cast_i8_to_i32", "lmc_full_features",
    "lmc_delta_features", "lmc_run", "lmc_launch_count", "lmc_env_launch_count", "lmc_c64_launch_count", "lmc_spec_tables_host", "lmc_spec_env_host", "lmc_spec_c64_host", "lmc_model_info",
    "lmc_ewald_field", "lmc_bias_init", "lmc_ewald_site_kernel", "lmc_distance_init", "lmc_full_features_field",
)

_LIB = None


def lib_path() -> str:
    return os.environ.get("LMC_LIBRARY") or os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                         "_lib", "liblmc.so")


def load():
    """Load liblmc.so; raises (no fallback) when the CUDA extension has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build the CUDA extension with `python -m smol_b200.build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(path)
    lib.lmc_version.restype = C.c_int
    lib.lmc_last_error.restype = C.c_char_p
    lib.lmc_row_stride.argtypes = [C.c_int]
    lib.lmc_model_create.argtypes = [C.POINTER(LmcModelDesc), C.POINTER(_P)]
    lib.lmc_model_destroy.argtypes = [_P]
    lib.lmc_model_num_features.argtypes = [_P]
    lib.lmc_cast_i32_to_i8.argtypes = [_P, _P, C.c_int, C.c_int, _P]
    lib.lmc_cast_i8_to_i32.argtypes = [_P, _P, C.c_int64, C.c_int, C.c_int, _P]
    lib.lmc_full_features.argtypes = [_P, _P, C.c_int, _P, _P, _P]
    lib.lmc_delta_features.argtypes = [_P, _P, C.c_int, _P, _P, C.c_int, _P, _P]
    lib.lmc_run.argtypes = [_P, C.POINTER(LmcRunConfig), _P]
    lib.lmc_model_info.argtypes = [_P, C.POINTER(C.c_int32), C.c_int]
    lib.lmc_ewald_field.argtypes = [_P, _P, C.c_int, _P, _P]
    lib.lmc_full_features_field.argtypes = [_P, _P, C.c_int, _P, _P, _P, _P]
    lib.lmc_bias_init.argtypes = [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _P, _P, _P, _P, _P]
    lib.lmc_ewald_site_kernel.argtypes = [_P, C.c_int, _P, C.c_int, _P, _P, C.c_int, _P, C.c_int, C.c_double,
                                          C.c_double, C.c_double, _P, _P]
    lib.lmc_distance_init.argtypes = [_P, C.c_int, _P, _P, _P, _P, C.c_double, C.c_int, _P, _P, _P, _P]
    lib.lmc_launch_count.restype = C.c_int64
    lib.lmc_env_launch_count.restype = C.c_int64
    lib.lmc_c64_launch_count.restype = C.c_int64
    lib.lmc_spec_tables_host.argtypes = [C.POINTER(LmcModelDesc), C.POINTER(C.c_int32), _P, C.c_int64, _P, C.c_int64]
    lib.lmc_spec_c64_host.argtypes = [C.POINTER(LmcModelDesc), C.POINTER(C.c_int32), _P, C.c_int64, _P, C.c_int64,
                                      _P, C.c_int64, _P, C.c_int64]
    lib.lmc_spec_env_host.argtypes = [C.POINTER(LmcModelDesc), C.POINTER(C.c_int32), _P, C.c_int64, _P, C.c_int64,
                                      _P, C.c_int64]
    if lib.lmc_version() != LMC_ABI_VERSION:
        raise RuntimeError("liblmc.so ABI version mismatch; rebuild with python -m smol_b200.build")
    _LIB = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise RuntimeError("liblmc: " + load().lmc_last_error().decode())
