"""ctypes loader for libpsim_b200.so (include/psim_b200.h).

There is no CPU fallback: if the library cannot be loaded, or no CUDA device is present when a
context is created, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

PSIM_OK = 0
ERRORS = {-1: "PSIM_E_CUDA", -2: "PSIM_E_ARG", -3: "PSIM_E_OOM", -4: "PSIM_E_NODE_OVERFLOW",
          -5: "PSIM_E_STATE", -6: "PSIM_E_NCCL"}
BUILD_CONTAINING, BUILD_DOMAIN = 0, 1
SR_LJ, SR_REPULSION, SR_STACK_PRESSURE = 1, 2, 4


class PsimError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("theta", C.c_float), ("epsilon", C.c_float), ("leaf_capacity", C.c_uint32),
        ("thread_capacity", C.c_uint32), ("lj_force_max", C.c_float), ("collision_passes", C.c_uint32),
        ("stack_pressure_enabled", C.c_uint32), ("stack_pressure", C.c_float),
        ("stack_pressure_decay", C.c_float), ("parity_mode", C.c_uint32), ("node_factor", C.c_float),
        ("strict_centres", C.c_uint32), ("reserved", C.c_uint32 * 4),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("n_bodies", C.c_uint64), ("n_electrons", C.c_uint64), ("compact_nodes", C.c_uint64),
        ("reference_nodes", C.c_uint64), ("max_depth", C.c_uint32), ("depth_cap", C.c_uint32),
        ("zero_leaves", C.c_uint32), ("cap_leaves", C.c_uint32), ("root_center", C.c_float * 2),
        ("root_size", C.c_float), ("grid_x", C.c_uint32), ("grid_y", C.c_uint32),
        ("traversal_warp_steps", C.c_uint64), ("kernel_launches", C.c_uint64),
    ]


class StepParams(C.Structure):
    _fields_ = [
        ("hw", C.c_float), ("hh", C.c_float), ("hd", C.c_float), ("dt", C.c_float),
        ("damping_base", C.c_float), ("k_e", C.c_float), ("bg_x", C.c_float), ("bg_y", C.c_float),
        ("density_threshold", C.c_float), ("enable_out_of_plane", C.c_uint32),
        ("do_short_range", C.c_uint32), ("do_electrons", C.c_uint32), ("do_iterate", C.c_uint32),
        ("do_polar", C.c_uint32), ("reserved", C.c_uint32 * 2),
    ]


SPECIES_DTYPE = np.dtype([
    ("mass", "<f4"), ("radius", "<f4"), ("damping", "<f4"), ("lj_epsilon", "<f4"),
    ("lj_sigma", "<f4"), ("lj_cutoff", "<f4"), ("polar_offset", "<f4"), ("polar_charge", "<f4"),
    ("repulsion_strength", "<f4"), ("repulsion_cutoff", "<f4"), ("lj_enabled", "<u4"),
    ("repulsion_enabled", "<u4"),
], align=True)
NODE_DTYPE = np.dtype([
    ("children", "<u8"), ("next", "<u8"), ("pos", "<f4", (2,)), ("mass", "<f4"),
    ("quad_center", "<f4", (2,)), ("quad_size", "<f4"), ("bodies_start", "<u8"),
    ("bodies_end", "<u8"), ("charge", "<f4"), ("_pad", "<u4"),
], align=True)

# every symbol include/psim_b200.h declares: (restype, argtypes)
_vp, _f, _i32, _u32, _u64 = C.c_void_p, C.c_float, C.c_int32, C.c_uint32, C.c_uint64
SIGNATURES = {
    "psim_default_config": (None, [_vp]),
    "psim_default_species_table": (None, [_vp]),
    "psim_create": (_i32, [_i32, _u64, _u64, _vp, C.POINTER(_vp)]),
    "psim_destroy": (_i32, [_vp]),
    "psim_last_error": (C.c_char_p, [_vp]),
    "psim_set_config": (_i32, [_vp, _vp]),
    "psim_set_stream": (_i32, [_vp, _u64]),
    "psim_sync": (_i32, [_vp]),
    "psim_stats_get": (_i32, [_vp, _vp]),
    "psim_reset_counters": (_i32, [_vp]),
    "psim_field_counters": (_i32, [_vp, _vp]),
    "psim_build_info": (_i32, [_vp, _vp]),
    "psim_fp32_peak": (_i32, [_vp, C.POINTER(C.c_float), C.POINTER(_i32)]),
    "psim_upload_species_table": (_i32, [_vp, _vp, _u32]),
    "psim_upload_bodies": (_i32, [_vp, _u64] + [_vp] * 8),
    "psim_update_state": (_i32, [_vp, _u64, _vp, _vp, _vp]),
    "psim_update_positions": (_i32, [_vp, _u64, _vp]),
    "psim_update_charges": (_i32, [_vp, _u64, _vp]),
    "psim_upload_electrons": (_i32, [_vp, _u64, _vp, _vp, _vp]),
    "psim_download_bodies": (_i32, [_vp] + [_vp] * 12),
    "psim_download_electrons": (_i32, [_vp, _vp, _vp, _vp]),
    "psim_build": (_i32, [_vp, _i32, _f, _f]),
    "psim_build_async": (_i32, [_vp, _i32, _f, _f]),
    "psim_build_status": (_i32, [_vp]),
    "psim_get_permutation": (_i32, [_vp, _vp]),
    "psim_get_keys": (_i32, [_vp, _vp]),
    "psim_download_nodes": (_i32, [_vp, _vp, _u64, C.POINTER(_u64)]),
    "psim_field": (_i32, [_vp, _f, _f, _f, _i32, _vp, _vp]),
    "psim_acc_points": (_i32, [_vp, _u64, _vp, _vp, _vp, _f, _vp]),
    "psim_update_electrons": (_i32, [_vp, _f, _f, _f, _f]),
    "psim_collide": (_i32, [_vp, _f, _f, _f, _u32, _u32, _f, _i32, _i32, C.POINTER(_u64)]),
    "psim_hop_alignment": (_i32, [_vp, _u64, _vp, _vp, _vp, _f, _f, _f, _f, _vp, _vp]),
    "psim_cell_build": (_i32, [_vp, _f, _f, _f]),
    "psim_cell_download": (_i32, [_vp, C.POINTER(_u64), C.POINTER(_u64), _vp, _vp]),
    "psim_neighbors_within": (_i32, [_vp, _u64, _vp, _f, _i32, _vp, _vp, _u64, C.POINTER(_u64)]),
    "psim_reset_acc": (_i32, [_vp]),
    "psim_use_cell_list": (_i32, [_vp, _f, _f, _f]),
    "psim_prepare_spatial_structures": (_i32, [_vp, _f, _f, _f]),
    "psim_short_range": (_i32, [_vp, _u32]),
    "psim_apply_polar_forces": (_i32, [_vp, _f, _i32]),
    "psim_iterate": (_i32, [_vp, _f, _f, _f, _f, _f, _i32]),
    "psim_update_surrounded_flags": (_i32, [_vp, _f, _f, _u64, _f, _u64]),
    "psim_get_surrounded": (_i32, [_vp, _vp, _vp, _vp]),
    "psim_enforce_metal_z_boundaries": (_i32, [_vp, _f, _f, _f]),
    "psim_shard_init": (_i32, [_vp, _u32, _u32]),
    "psim_shard_phase": (_i32, [_vp, _i32, _i32, _f, _f, _vp]),
    "psim_shard_ptrs": (_i32, [_vp, _vp]),
    "psim_step": (_i32, [_vp, _vp]),
    "psim_shard_capacity": (_u64, [_u64, _u32]),
    "psim_comm_unique_id": (_i32, [_vp]),
    "psim_comm_init": (_i32, [_vp, _vp, _u32, _u32]),
    "psim_comm_destroy": (_i32, [_vp]),
    "psim_comm_stats": (_i32, [_vp, _vp]),
    "psim_build_sharded": (_i32, [_vp, _i32, _f, _f]),
    "psim_step_sharded": (_i32, [_vp, _vp]),
    "psim_step_host": (_i32, [_vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "psim_phase_times": (_i32, [_vp, _vp]),
    "psim_set_target_range": (_i32, [_vp, _u64, _u64]),
    "psim_set_electron_range": (_i32, [_vp, _u64, _u64]),
    "psim_device_ptrs": (_i32, [_vp, _vp]),
    "psim_mark_positions_changed": (_i32, [_vp]),
}

_LIB = None


def lib_path() -> str:
    return _build.LIB


def load(rebuild_if_stale: bool = True) -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        if not rebuild_if_stale:
            raise RuntimeError(f"{path} is missing; run particlesim_b200/build.py (needs nvcc)")
        _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = the library does not export the ABI
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def default_config(**kw) -> Config:
    cfg = Config()
    load().psim_default_config(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def default_species_table() -> np.ndarray:
    t = np.zeros(21, dtype=SPECIES_DTYPE)
    load().psim_default_species_table(t.ctypes.data)
    return t
