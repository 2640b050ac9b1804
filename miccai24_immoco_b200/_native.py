"""ctypes binding of the C ABI declared in include/immoco_b200.h (libimmoco_b200.so).

There is NO CPU fallback: if the CUDA library is missing every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
# IMMOCO_LIB_PATH / IMMOCO_NVCC_FLAGS: build-variant experiments (tools/), not used by the product path
LIB_PATH = os.environ.get("IMMOCO_LIB_PATH") or os.path.join(_HERE, "libimmoco_b200.so")
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["hashgrid.cu", "hashgrid_csr.cu", "mlp_tc.cu", "forward_model.cu", "fit.cu", "metrics.cu", "simulate.cu", "unet.cu",
           "unet_tc.cu", "autofocus.cu"]

MAX_LEVELS = 16
ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2
ERR_BAD_ARG, ERR_UNSUPPORTED = -1, -2
LAYOUT_LUT = 0xFFFFFFFF
PROFILE_SLOTS = ["hashgrid_fwd_image", "mlp_fwd_image", "hashgrid_fwd_motion", "mlp_fwd_motion", "fft_rows",
                 "motion_rows_fwd", "colpass_loss", "grad_entropy", "fft_rows_adj", "motion_rows_bwd",
                 "mlp_bwd_motion", "hashgrid_bwd_motion", "mlp_bwd_image", "hashgrid_bwd_image", "adam_motion",
                 "adam_image"]


class GridDesc(C.Structure):
    _fields_ = [
        ("n_dims", C.c_int32),
        ("n_levels", C.c_int32),
        ("scale", C.c_float * MAX_LEVELS),
        ("resolution", C.c_uint32 * MAX_LEVELS),
        ("entries", C.c_uint32 * MAX_LEVELS),
        ("offset", C.c_uint32 * (MAX_LEVELS + 1)),
        ("hashed", C.c_uint32 * MAX_LEVELS),
        ("swizzle", C.c_uint32 * MAX_LEVELS),
        ("layout_lut", C.c_void_p),
    ]


class GridCsr(C.Structure):
    _fields_ = [
        ("row_ptr", C.c_void_p),
        ("taps", C.c_void_p),
        ("n_taps", C.c_int64),
        ("n_points", C.c_int64),
    ]


class GridTaps(C.Structure):
    _fields_ = [
        ("rows", C.c_void_p),
        ("first_level", C.c_int32),
        ("reserved", C.c_int32),
        ("n_points", C.c_int64),
        ("n_active_rows", C.c_int64),
    ]


class Lines(C.Structure):
    _fields_ = [
        ("n_groups", C.c_int32),
        ("group_ofs", C.c_void_p),
        ("line_idx", C.c_void_p),
        ("line_w", C.c_void_p),
        ("static_w", C.c_void_p),
        ("max_lines", C.c_int32),
    ]


class Fit(C.Structure):
    _fields_ = [
        ("h", C.c_int32), ("w", C.c_int32), ("m", C.c_int32),
        ("grid_image", GridDesc), ("grid_motion", GridDesc),
        ("width_image", C.c_int32), ("act_image", C.c_int32),
        ("width_motion", C.c_int32), ("act_motion", C.c_int32),
        ("n_motion", C.c_int64), ("n_image", C.c_int64),
        ("params", C.c_void_p), ("grads", C.c_void_p),
        ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
        ("coords_image", C.c_void_p), ("coords_motion", C.c_void_p),
        ("lines", Lines),
        ("tw_h", C.c_void_p), ("tw_w", C.c_void_p),
        ("k_in", C.c_void_p),
        ("enc_image", C.c_void_p), ("d_enc_image", C.c_void_p),
        ("enc_motion", C.c_void_p), ("d_enc_motion", C.c_void_p),
        ("image", C.c_void_p), ("d_image", C.c_void_p),
        ("disp", C.c_void_p), ("d_disp", C.c_void_p),
        ("c_tmp", C.c_void_p), ("d_c", C.c_void_p), ("k_out", C.c_void_p),
        ("loss", C.c_void_p),
        ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
        ("loss_slots", C.c_void_p),
        ("deterministic", C.c_int32), ("fuse_adam", C.c_int32),
        ("csr_image", GridCsr), ("csr_motion", GridCsr),
        ("mlp_part_image", C.c_void_p), ("mlp_part_motion", C.c_void_p),
        ("d_image_fx", C.c_void_p), ("dc_max_bits", C.c_void_p),
        ("taps_image", GridTaps),
    ]


_P = C.c_void_p
_SIGNATURES = {
    "immoco_hashgrid_fwd": (C.c_int, [C.POINTER(GridDesc), _P, _P, _P, C.c_int64, _P]),
    "immoco_hashgrid_bwd": (C.c_int, [C.POINTER(GridDesc), _P, _P, _P, C.c_int64, _P]),
    "immoco_hashgrid_fwd_levels": (C.c_int, [C.POINTER(GridDesc), _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P]),
    "immoco_hashgrid_bwd_levels": (C.c_int, [C.POINTER(GridDesc), _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P]),
    "immoco_mlp_fwd": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P]),
    "immoco_mlp_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P]),
    "immoco_tanh_bwd": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "immoco_hashgrid_csr_workspace_bytes": (C.c_int64, [C.POINTER(GridDesc), C.c_int64]),
    "immoco_hashgrid_csr_build": (C.c_int, [C.POINTER(GridDesc), _P, C.c_int64, _P, _P, _P, C.c_int64, _P]),
    "immoco_hashgrid_bwd_csr": (C.c_int, [C.POINTER(GridDesc), C.POINTER(GridCsr), _P, _P, _P]),
    "immoco_hashgrid_bwd_csr_adam": (C.c_int, [C.POINTER(GridDesc), C.POINTER(GridCsr), _P, _P, _P, _P, _P, C.c_double,
                                               C.c_double, C.c_double, C.c_double, C.c_int32, _P]),
    "immoco_hashgrid_fwd_grouped": (C.c_int, [C.POINTER(GridDesc), _P, _P, _P, C.c_int64, C.c_int32, _P]),
    "immoco_hashgrid_bwd_grouped": (C.c_int, [C.POINTER(GridDesc), _P, _P, _P, C.c_int64, C.c_int32, _P]),
    "immoco_hashgrid_tap_rows": (C.c_int, [C.POINTER(GridDesc), _P, C.c_int64, C.c_int32, C.c_int32, _P, _P]),
    "immoco_hashgrid_fwd_taps": (C.c_int, [C.POINTER(GridDesc), C.POINTER(GridTaps), _P, _P, _P, C.c_int64, _P]),
    "immoco_hashgrid_bwd_taps": (C.c_int, [C.POINTER(GridDesc), C.POINTER(GridTaps), _P, _P, _P, C.c_int64, _P]),
    "immoco_mlp_bwd_partials": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P]),
    "immoco_mlp_bwd_scatter": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.POINTER(GridDesc), _P, _P, C.c_int64, C.c_int32,
                                         C.c_int32, _P]),
    "immoco_hashgrid_bwd_dense_levels": (C.c_int, [C.POINTER(GridDesc), _P, _P, _P, C.c_int64, _P]),
    "immoco_set_fused_scatter": (C.c_int, [C.c_int32]),
    "immoco_get_fused_scatter": (C.c_int, []),
    "immoco_mlp_bwd_partial_count": (C.c_int, [C.c_int64]),
    "immoco_adam_step_partials": (C.c_int, [_P, _P, C.c_int32, _P, _P, C.c_int64, C.c_double, C.c_double, C.c_double,
                                            C.c_double, C.c_int32, _P]),
    "immoco_fit_loss_slots": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "immoco_set_deterministic": (C.c_int, [C.c_int32]),
    "immoco_get_deterministic": (C.c_int, []),
    "immoco_release_streams": (C.c_int, []),
    "immoco_launches_per_iteration_mode": (C.c_int, [C.c_int32, C.c_int32, C.c_int32]),
    "immoco_set_hashgrid_impl": (C.c_int, [C.c_int32]),
    "immoco_set_hashgrid_ctas_per_sm": (C.c_int, [C.c_int32]),
    "immoco_set_hashgrid_bwd_ctas_per_sm": (C.c_int, [C.c_int32]),
    "immoco_set_adam_tuning": (C.c_int, [C.c_int32, C.c_int32]),
    "immoco_fft2c": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, C.c_int32,
                               C.c_float, _P]),
    "immoco_forward_model": (C.c_int, [_P, _P, _P, C.POINTER(Lines), _P, _P, _P, _P, C.c_int32,
                                       C.c_int32, _P]),
    "immoco_forward_model_bwd": (C.c_int, [_P, _P, _P, _P, C.POINTER(Lines), _P, _P, _P, _P, _P,
                                           C.c_int32, C.c_int32, C.c_int32, _P]),
    "immoco_colpass_loss": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "immoco_grad_entropy": (C.c_int, [_P, C.c_float, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "immoco_adam_step": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_double, C.c_double, C.c_double,
                                   C.c_double, C.c_int32, C.c_int32, _P]),
    "immoco_fit_run": (C.c_int, [C.POINTER(Fit), C.c_int32, C.c_int32, C.POINTER(C.c_float), _P, _P,
                                 C.c_int32]),
    "immoco_fit_run_batched": (C.c_int, [C.POINTER(C.POINTER(Fit)), C.c_int32, C.c_int32, C.c_int32,
                                         C.POINTER(C.c_float), _P, _P, C.c_int32]),
    "immoco_max_fit_batch": (C.c_int, []),
    "immoco_metrics2d": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, _P, C.c_int64, C.c_int64, C.c_int32,
                                   C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    "immoco_haarpsi": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                 C.c_int32, C.c_int32, C.c_float, C.c_float, _P, _P, _P, _P]),
    "immoco_rigid_resample": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "immoco_replace_lines": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "immoco_unet_conv3x3": (C.c_int, [_P, C.c_int32, _P, C.c_int32, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_int32, _P]),
    "immoco_unet_pack_conv3x3": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P]),
    "immoco_unet_conv3x3_tc": (C.c_int, [_P, C.c_int32, _P, C.c_int32, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32, _P]),
    "immoco_unet_convt2x2": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "immoco_unet_instnorm_lrelu": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, _P]),
    "immoco_unet_conv1x1": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "immoco_rigid_bicubic_fwd": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "immoco_rigid_bicubic_bwd_theta": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "immoco_profile_create": (_P, [C.c_int32]),
    "immoco_profile_destroy": (None, [_P]),
    "immoco_profile_read": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "immoco_profile_timeline": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "immoco_set_profile_overlap": (C.c_int, [C.c_int32]),
    "immoco_set_branch_overlap": (C.c_int, [C.c_int32]),
    "immoco_set_pdl": (C.c_int, [C.c_int32]),
    "immoco_set_fused_rows": (C.c_int, [C.c_int32]),
    "immoco_set_deferred_zero": (C.c_int, [C.c_int32]),
    "immoco_abi_version": (C.c_int, []),
    "immoco_launches_per_iteration": (C.c_int, [C.c_int32]),
    "immoco_struct_sizes": (None, [C.POINTER(C.c_int32)]),
}
EXPORTED_SYMBOLS: List[str] = sorted(_SIGNATURES)

_lib = None


def nvcc_command(out_path: str = LIB_PATH) -> List[str]:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    return [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--threads", "0",
            "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(_ROOT, "include"),
            "-o", out_path] + os.environ.get("IMMOCO_NVCC_FLAGS", "").split() + [os.path.join(CSRC, s) for s in SOURCES]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_ROOT, "include", "immoco_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into the in-tree shared library."""
    if force or needs_build():
        cmd = nvcc_command()
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the IM-MoCo CUDA extension has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)     # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        sizes = (C.c_int32 * 5)()
        handle.immoco_struct_sizes(sizes)
        want = (C.sizeof(GridDesc), C.sizeof(Lines), C.sizeof(Fit), C.sizeof(GridCsr), C.sizeof(GridTaps))
        if tuple(sizes) != want:
            raise RuntimeError(f"struct layout mismatch: library {tuple(sizes)} vs binding {want}")
        _lib = handle
    return _lib


class NativeError(RuntimeError):
    pass


def check(code: int, what: str) -> None:
    if code == 0:
        return
    if code == ERR_BAD_ARG:
        raise ValueError(f"{what}: invalid argument")
    if code == ERR_UNSUPPORTED:
        raise NotImplementedError(f"{what}: configuration not supported by the CUDA path")
    raise NativeError(f"{what}: CUDA error {code}")
