"""ctypes binding of libsspbo.so (C ABI: include/sspbo.h).

The shared library is built in-tree by `__graft_entry__.build()` /
`ssp_bayesopt_b200/csrc/Makefile`.  There is no CPU fallback: if the library is
missing, or no CUDA device is present when a context is requested, this module
raises instead of degrading.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SSPBO_LIB") or os.path.join(_HERE, "lib", "libsspbo.so")   # SSPBO_LIB: kernel-variant experiments

F32, F64 = 0, 1
ENGINE_AUTO, ENGINE_SIMT, ENGINE_UMMA, ENGINE_EXACT, ENGINE_UMMA_LR, ENGINE_UMMA_F8 = 0, 1, 2, 3, 4, 5
ENGINE_NAMES = {ENGINE_AUTO: "auto", ENGINE_SIMT: "simt", ENGINE_UMMA: "umma", ENGINE_EXACT: "exact",
                ENGINE_UMMA_LR: "umma-lowrank", ENGINE_UMMA_F8: "umma-factor-f8"}
CAND_UNIFORM, CAND_GRID = 1, 2

_vp, _i, _i64, _u64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
_pd, _pi = C.POINTER(C.c_double), C.POINTER(C.c_int)

# name -> (restype, argtypes); mirrors include/sspbo.h one to one
SIGNATURES = {
    "sspbo_abi_version": (_i, []),
    "sspbo_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "sspbo_ctx_destroy": (None, [_vp]),
    "sspbo_last_error": (C.c_char_p, [_vp]),
    "sspbo_space_set": (_i, [_vp, _pd, _pd, _i, _i]),
    "sspbo_space_set_classes": (_i, [_vp, _pd, _pd, _i, _i, _i]),
    "sspbo_space_set_traj": (_i, [_vp, _pd, _i]),
    "sspbo_space_set_waypoints": (_i, [_vp, _pd, _pi, _i, _i]),
    "sspbo_space_dims": (_i, [_vp, _pi, _pi, _pi, _pi]),
    "sspbo_space_set_xbound": (_i, [_vp, _d]),
    "sspbo_score_stats": (_i, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), _pi, _i, _vp]),
    "sspbo_score_finish": (_i, [_vp, _vp, C.POINTER(_u64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64),
                                _pi, _i, _vp]),
    "sspbo_eval_host": (_i, [_vp, _pd, _i, _d, _d, _pd, _pd, _pd, _vp]),
    "sspbo_lowrank_info": (_i, [_vp, _pi, _pi, _pi, _pi]),
    "sspbo_set_policy": (_i, [_vp, _i, _i]),
    "sspbo_encode": (_i, [_vp, _vp, _i, _i64, _vp, _vp]),
    "sspbo_encode_deriv": (_i, [_vp, _vp, _i, _i64, _vp, _vp, _vp]),
    "sspbo_rff_encode": (_i, [_vp, _vp, _vp, _i, _i, _d, _vp, _i, _i64, _vp, _vp]),
    "sspbo_encode_pitch": (_i, [_vp]),
    "sspbo_encode_f32": (_i, [_vp, _vp, _i, _i64, _vp, _vp]),
    "sspbo_blr_init": (_i, [_vp, _d]),
    "sspbo_blr_init_dim": (_i, [_vp, _i, _d]),
    "sspbo_blr_set_beta": (_i, [_vp, _d]),
    "sspbo_blr_get_beta": (_i, [_vp, _pd, _pi]),
    "sspbo_blr_update": (_i, [_vp, _vp, _pd, _i, _vp]),
    "sspbo_blr_update_x": (_i, [_vp, _pd, _pd, _i, _vp]),
    "sspbo_observe": (_i, [_vp, _pd, _d, _pd, _pd, _vp]),
    "sspbo_set_update_path": (_i, [_vp, _i]),
    "sspbo_blr_get": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "sspbo_blr_export_size": (_i64, [_vp]),
    "sspbo_blr_export": (_i, [_vp, _vp, _vp]),
    "sspbo_blr_import": (_i, [_vp, _vp, _d, _vp]),
    "sspbo_blr_import_ranked": (_i, [_vp, _vp, _d, _i, _vp]),
    "sspbo_blr_rank": (_i, [_vp, _pi]),
    "sspbo_blr_checkpoint": (_i, [_vp, _vp]),
    "sspbo_blr_rollback": (_i, [_vp, _vp]),
    "sspbo_dist_init": (_i, [_vp, _i, _i, _vp]),
    "sspbo_dist_connect": (_i, [_vp, _vp]),
    "sspbo_dist_ready": (_i, [_vp]),
    "sspbo_dist_slots": (_i, [_vp, C.POINTER(C.c_void_p)]),
    "sspbo_dist_connect_ptrs": (_i, [_vp, C.POINTER(C.c_void_p)]),
    "sspbo_dist_exchange": (_i, [_vp, _vp, _vp]),
    "sspbo_blr_predict": (_i, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "sspbo_score": (_i, [_vp, _vp, _i, _i64, _i64, _d, _d, _vp, _vp, _vp, _vp, _i, _vp]),
    "sspbo_score_generated": (_i, [_vp, _i, _u64, _vp, _pi, _i64, _i64, _d, _d, _vp, _vp, _vp, _vp, _i, _vp]),
    "sspbo_set_operand_format": (_i, [_vp, _i]),
    "sspbo_key_unpack": (None, [_u64, C.POINTER(C.c_float), C.POINTER(_i64)]),
    "sspbo_key_pack": (_u64, [C.c_float, _i64]),
    "sspbo_last_engine": (_i, [_vp]),
    "sspbo_launch_count": (_i64, [_vp]),
    "sspbo_set_timing": (_i, [_vp, _i]),
    "sspbo_timing_read": (_i, [_vp, _pd, _pd, C.POINTER(_i64), _i, _vp]),
    "sspbo_from_set_argmax": (_i, [_vp, _vp, _i64, _vp, _i, _vp, _vp]),
    "sspbo_decode_optim": (_i, [_vp, _vp, _i, _vp, _pd, _vp, _vp]),
    "sspbo_acq_ascent": (_i, [_vp, _vp, _i, _d, _d, _d, _i, _d, _vp, _vp, _vp, _vp]),
    "sspbo_gram": (_i, [_vp, _vp, _i, _vp, _vp]),
    "sspbo_score_batch": (_i, [C.POINTER(_vp), _i, _vp, _i, _i64, _pd, _pd, _vp, C.POINTER(_vp), _i, _pi]),
    "sspbo_score_batch_status": (_i, [C.POINTER(_vp), _i, _vp, C.POINTER(_vp), _i, _pi]),
    "sspbo_update_batch": (_i, [C.POINTER(_vp), _i, _pd, _pd, C.POINTER(_vp), _i, _pi]),
    "sspbo_bind": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp]),
    "sspbo_fill_uniform": (_i, [_vp, _u64, _i64, _i64, _i, _vp, _vp]),
    "sspbo_grid_points": (_i, [_vp, _vp, _pi, _i, _i64, _i64, _vp, _vp]),
}

_lib = None


class SspboError(RuntimeError):
    pass


def load():
    """Load libsspbo.so and bind every symbol the header declares.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SspboError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C ssp_bayesopt_b200/csrc`).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so is stale
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(lib, ctx, rc, what):
    if rc != 0:
        msg = lib.sspbo_last_error(ctx)
        raise SspboError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def key_unpack(key: int):
    lib = load()
    s, i = C.c_float(), C.c_int64()
    lib.sspbo_key_unpack(C.c_uint64(key & 0xFFFFFFFFFFFFFFFF), C.byref(s), C.byref(i))
    return s.value, i.value


def key_pack(score: float, index: int) -> int:
    return int(load().sspbo_key_pack(C.c_float(score), C.c_int64(index)))
