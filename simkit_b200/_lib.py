"""ctypes binding of the C ABI in ``include/simkit_b200.h``.

The product path is the CUDA library: if it is missing or no GPU is usable this
module raises -- there is no CPU fallback (BASELINE.json north_star).
"""

import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_TAG = os.environ.get("SKB_LIB_TAG", "")  # development: load an A/B build (see build.py)
LIB_PATH = os.path.join(HERE, "libsimkit_b200" + ("_" + _TAG if _TAG else "") + ".so")

SKB_OK = 0
SKB_EINVAL = -1

MATERIAL_IDS = {
    "stable_neo_hookean": 0,
    "neo_hookean": 1,
    "arap": 2,
    "stvk": 3,
    "linear_elasticity": 4,
    "fcr": 5,
    "macklin_mueller_neo_hookean": 6,
}
PSD_NONE, PSD_AFTER_VOL, PSD_BEFORE_VOL = 0, 1, 2

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_dbl = ctypes.c_double


class NewtonOpts(ctypes.Structure):
    _fields_ = [
        ("material", _int), ("psd_mode", _int), ("max_iter", _int), ("do_line_search", _int),
        ("tolerance", _dbl), ("ls_alpha", _dbl), ("ls_beta", _dbl), ("ls_max_iter", _int),
        ("ls_threshold", _dbl), ("pcg_rtol", _dbl), ("pcg_max_iter", _int),
    ]


class NewtonInfo(ctypes.Structure):
    _fields_ = [
        ("iters", _int), ("pcg_iters_total", _int), ("last_alpha", _dbl), ("last_step_norm", _dbl),
        ("last_pcg_relres", _dbl), ("alphas", _dbl * 64),
    ]


class DistPcgArgs(ctypes.Structure):
    """skb_dist_pcg_args (include/simkit_b200.h)."""
    _fields_ = [
        ("vals", _vp), ("diag", _vp), ("rhs", _vp), ("x", _vp),
        ("dinv", _vp), ("r", _vp), ("z", _vp), ("p", _vp), ("q", _vp),
        ("s", _vp), ("work", _vp), ("Ac", _vp), ("rc", _vp), ("zc", _vp), ("stream", _vp),
        ("rtol", _dbl), ("v0", ctypes.c_int32), ("v1", ctypes.c_int32), ("max_iter", ctypes.c_int32),
        ("check_every", ctypes.c_int32), ("use_graph", ctypes.c_int32), ("reserved", ctypes.c_int32),
    ]


class DistPcg2Args(ctypes.Structure):
    """skb_dist_pcg2_args (include/simkit_b200.h)."""
    _fields_ = [
        ("vals", _vp), ("diag", _vp), ("rhs", _vp), ("x", _vp), ("stream", _vp), ("rtol", _dbl),
        ("v0", ctypes.c_int32), ("v1", ctypes.c_int32), ("max_iter", ctypes.c_int32), ("check_every", ctypes.c_int32),
        ("use_graph", ctypes.c_int32), ("use_coarse", ctypes.c_int32), ("transport", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


_MAT = [_vp, _i64, _vp, _i64, _vp, _i64]  # mu, mu_n, lam, lam_n, vol, vol_n

SIGNATURES = {
    "skb_last_error": (ctypes.c_char_p, []),
    "skb_device_count": (_int, []),
    "skb_version": (ctypes.c_char_p, []),
    "skb_plan_create": (_int, [_vp, _vp, _int, _i64, _i64, _int, _int, _int, ctypes.POINTER(_vp)]),
    "skb_plan_create_from_operator": (_int, [_vp, _vp, _int, _i64, _i64, _int, _int, _int, ctypes.POINTER(_vp)]),
    "skb_plan_create_sharded": (_int, [_vp, _vp, _int, _i64, _i64, _i64, _i64, _int, _int, _int, ctypes.POINTER(_vp)]),
    "skb_plan_destroy": (None, [_vp]),
    "skb_plan_info": (_int, [_vp, _vp]),
    "skb_plan_csr_pattern": (_int, [_vp, _vp, _vp]),
    "skb_plan_block_pattern": (_int, [_vp, _vp, _vp]),
    "skb_plan_slot_map": (_int, [_vp, _vp]),
    "skb_plan_element_D": (_int, [_vp, _vp]),
    "skb_plan_volume": (_int, [_vp, _vp]),
    "skb_plan_element_order": (_int, [_vp, _vp]),
    "skb_plan_vertex_masses": (_int, [_vp, _vp, _i64, _vp]),
    "skb_energy": (_int, [_vp, _int, _vp, _vp] + _MAT + [_vp]),
    "skb_gradient": (_int, [_vp, _int, _vp, _vp] + _MAT + [_vp]),
    "skb_hessian": (_int, [_vp, _int, _int, _vp, _vp] + _MAT + [_vp]),
    "skb_gradient_hessian": (_int, [_vp, _int, _int, _vp, _vp] + _MAT + [_vp, _vp]),
    "skb_set_materials": (_int, [_vp] + _MAT),
    "skb_set_materials_dev": (_int, [_vp] + _MAT + [_vp]),
    "skb_energy_dev": (_int, [_vp, _int, _vp, _vp, _vp, _vp]),
    "skb_gradient_hessian_dev": (_int, [_vp, _int, _int, _vp, _vp, _vp, _vp, _vp]),
    "skb_last_launch_count": (_int, [_vp]),
    "skb_kernel_timing": (_int, [_vp, _int]),
    "skb_kernel_times": (_int, [_vp, _vp, _vp]),
    "skb_gather_dev": (_int, [_vp, _vp, _i64, _vp, _vp]),
    "skb_scatter_add_dev": (_int, [_vp, _vp, _i64, _vp, _vp]),
    "skb_scatter_dev": (_int, [_vp, _vp, _i64, _vp, _vp]),
    "skb_dist_pcg_init_dev": (_int, [_vp, _vp, _vp, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "skb_dist_spmv_dot_dev": (_int, [_vp, _vp, _vp, _int, _int, _vp, _vp, _vp, _vp, _vp]),
    "skb_dist_pcg_update_dev": (_int, [_vp, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "skb_dist_pcg_direction_dev": (_int, [_vp, _int, _int, _vp, _vp, _vp, _vp]),
    "skb_dist_newton_rhs_dev": (_int, [_vp, _int, _int, _vp, _vp, _vp, _vp, _dbl, _vp, _vp, _vp, _vp, _vp, _vp]),
    "skb_dist_newton_terms_dev": (_int, [_vp, _int, _int, _vp, _vp, _dbl, _vp, _vp, _vp, _dbl, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "skb_dist_coarse_set": (_int, [_vp, _i64, _vp, _vp, _int, _int]),
    "skb_dist_coarse_assemble_dev": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "skb_dist_coarse_invert_dev": (_int, [_vp, _vp, _vp]),
    "skb_dist_coarse_restrict_dev": (_int, [_vp, _vp, _vp, _vp]),
    "skb_dist_coarse_correct_dev": (_int, [_vp, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp, _vp]),
    "skb_fp64_peak": (_int, [_int, ctypes.POINTER(_dbl)]),
    "skb_dmma_peak": (_int, [_int, ctypes.POINTER(_dbl)]),
    "skb_element_energy": (_int, [_int, _int, _i64, _vp, _vp, _i64, _vp, _i64, _vp]),
    "skb_element_gradient": (_int, [_int, _int, _i64, _vp, _vp, _i64, _vp, _i64, _vp]),
    "skb_element_hessian": (_int, [_int, _int, _i64, _vp, _vp, _i64, _vp, _i64, _vp]),
    "skb_psd_project": (_int, [_i64, _int, _vp, _int, _vp]),
    "skb_svd_rv": (_int, [_int, _i64, _vp, _vp, _vp, _vp]),
    "skb_polar": (_int, [_int, _i64, _vp, _vp, _vp]),
    "skb_rotation_gradient": (_int, [_int, _i64, _vp, _vp]),
    "skb_stretch_gradient": (_int, [_int, _i64, _vp, _vp]),
    "skb_pcg": (_int, [_vp, _vp, _vp, _vp, _dbl, _int, _vp, _vp, _vp]),
    "skb_pcg_dev": (_int, [_vp, _vp, _vp, _vp, _dbl, _int, _vp, _vp, _vp, _vp]),
    "skb_csr_pcg": (_int, [_i64, _vp, _vp, _vp, _int, _vp, _dbl, _int, _vp, _vp, _vp]),
    "skb_dense_solve": (_int, [_i64, _vp, _vp, _vp]),
    "skb_contact_springs_plane": (_int, [_int, _i64, _vp, _dbl, _vp, _vp, _vp, ctypes.POINTER(_dbl), _vp, _vp, _vp]),
    "skb_newton_set_contact_plane": (_int, [_vp, _dbl, _vp, _vp, _vp]),
    "skb_contact_springs_sphere": (_int, [_int, _i64, _vp, _dbl, _vp, _dbl, _vp, ctypes.POINTER(_dbl), _vp, _vp, _vp]),
    "skb_newton_set_contact_sphere": (_int, [_vp, _dbl, _vp, _dbl, _vp]),
    "skb_nccl_unique_id": (_int, [_vp, _i64]),
    "skb_nccl_init": (_int, [_vp, _vp, _i64, _int, _int]),
    "skb_nccl_set_halo": (_int, [_vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "skb_nccl_finalize": (_int, [_vp]),
    "skb_dist_pcg_native": (_int, [_vp, ctypes.POINTER(DistPcgArgs), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(_dbl)]),
    "skb_dist_contact_dev": (_int, [_vp, _int, _int, _vp, _int, _dbl, _vp, _vp, _dbl, _vp, _vp, _vp, _vp, _vp, _vp]),
    "skb_dist_pcg2": (_int, [_vp, ctypes.POINTER(DistPcg2Args), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(_dbl)]),
    "skb_dist_pcg2_times": (_int, [_vp, _vp]),
    "skb_pcg2_peer_export": (_int, [_vp, _vp, _vp, _i64]),
    "skb_pcg2_peer_import": (_int, [_vp, _vp, _vp]),
    "skb_quadratic": (_int, [_i64, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(_dbl), _vp]),
    "skb_newton_set_quadratic": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "skb_buf_alloc": (_int, [_int, _i64, ctypes.POINTER(_vp)]),
    "skb_buf_free": (None, [_int, _vp]),
    "skb_buf_copy": (_int, [_int, _vp, _vp, _i64]),
    "skb_buf_axpy": (_int, [_int, _vp, _dbl, _vp, _i64]),
    "skb_buf_scale": (_int, [_int, _vp, _dbl, _i64]),
    "skb_buf_index_add": (_int, [_int, _vp, _vp, _vp, _i64, _dbl]),
    "skb_buf_add_diagonal": (_int, [_vp, _vp, _vp, _dbl]),
    "skb_buf_download": (_int, [_int, _vp, _i64, _vp]),
    "skb_host_alloc": (_int, [_i64, ctypes.POINTER(_vp)]),
    "skb_host_free": (None, [_vp]),
    "skb_gradient_hessian_resident": (_int, [_vp, _int, _int, _vp, _vp] + _MAT + [_vp, _vp]),
    "skb_qr_thin": (_int, [_i64, _i64, _vp, _vp, _vp]),
    "skb_weighted_gram": (_int, [_i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "skb_average_onto_simplex": (_int, [_i64, _i64, _i64, _int, _vp, _vp, _vp]),
    "skb_kmeans2_pp": (_int, [_i64, _i64, _i64, _int, _i64, _vp, _vp, _vp, _vp, _vp]),
    "skb_cubature_pick": (_int, [_i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "skb_pcg_vals_dev": (_int, [_vp, _vp, _vp, _vp, _dbl, _int, _vp, ctypes.POINTER(_int), ctypes.POINTER(_dbl)]),
    "skb_plan_value_positions": (_int, [_vp, _i64, _vp, _vp, _vp]),
    "skb_pcg_set_coarse": (_int, [_vp, _i64, _vp, _vp]),
    "skb_spmv_dev": (_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "skb_newton": (_int, [_vp, ctypes.POINTER(NewtonOpts), _vp, _vp, _vp, _dbl, _vp, _vp, _vp, _vp, ctypes.POINTER(NewtonInfo)]),
    "skb_reduced_gradient_hessian": (_int, [_int, _int, _int, _i64, _i64, _vp, _vp, _vp] + _MAT + [_vp, _vp, _vp]),
    "skb_reduced_hessian_from_basis": (_int, [_vp, _int, _int, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "skb_reduced_last_times": (_int, [_vp]),
    "skb_plan_set_basis": (_int, [_vp, _i64, _vp]),
    "skb_fst_precompute": (_int, [_int, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "skb_fst_eval": (_int, [_int, _i64, _i64, _i64, _vp, _vp, _vp]),
}

_lib = None


class SimkitB200Error(RuntimeError):
    pass


def load(require_gpu=True):
    """Load the CUDA library (once).  Raises if it is absent or no GPU is visible."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SimkitB200Error(
                "simkit_b200: CUDA library %s is missing. Build it with "
                "`python -m simkit_b200.build` (needs nvcc). There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        missing = [name for name in SIGNATURES if not hasattr(lib, name)]
        if missing:
            raise SimkitB200Error("simkit_b200: %s lacks symbols %s (stale build?)" % (LIB_PATH, missing))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    if require_gpu and _lib.skb_device_count() <= 0:
        raise SimkitB200Error("simkit_b200: no CUDA device is visible; this library has no CPU fallback")
    return _lib


def check(rc):
    if rc == SKB_OK:
        return
    msg = load(require_gpu=False).skb_last_error().decode("utf-8", "replace")
    if rc == SKB_EINVAL:
        raise ValueError(msg)
    raise SimkitB200Error("simkit_b200 error %d: %s" % (rc, msg))


def f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def material_arg(a, t, name):
    """(array, count) for a material parameter: scalar broadcast or per element."""
    if a is None:
        return None, 0
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    if arr.size not in (1, t):
        raise ValueError("%s must be a scalar or have one entry per element (%d), got %d" % (name, t, arr.size))
    return arr, arr.size
