"""ctypes binding of libalphapig_b200.so (include/alphapig_b200.h).

No CPU fallback: importing the symbols works anywhere the library file exists,
but creating an engine without a CUDA device raises ``EngineError``.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libalphapig_b200.so")

AP_OK, AP_ERR_BAD_ARG, AP_ERR_ILLEGAL_MOVE, AP_ERR_POOL_EXHAUSTED = 0, -1, -2, -3
AP_ERR_CUDA, AP_ERR_NO_NET, AP_ERR_BAD_HANDLE = -4, -5, -6
AP_META_INTS = 8
AP_FLAG_HIGH_PRIORITY_STREAM = 1
AP_ARCH_SIMPLE, AP_ARCH_RESNET, AP_ARCH_INCEPTION = 0, 1, 2
AP_NET_SPLIT = 0x100  # OR into arch: hi + lo fp16 operand pairs (near-fp32), residual net only
AP_NET_SPLIT_ACT = 0x200  # OR into arch: hi + lo activations x error-diffusion-rounded fp16 weights, residual net only


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "alphapig_b200 error %d: %s" % (code, msg))
        self.code = code


class ApConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("n_in_row", C.c_int32), ("n_games", C.c_int32),
                ("node_capacity", C.c_int32), ("n_playout_hint", C.c_int32), ("device", C.c_int32),
                ("flags", C.c_int32), ("c_puct", C.c_double)]


class ApTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("numel", C.c_int64)]


_P = C.c_void_p
_I = C.c_int32
# every exported symbol of include/alphapig_b200.h: name -> (restype, argtypes)
SIGNATURES = {
    "ap_engine_create": (C.c_int, [C.POINTER(ApConfig), C.POINTER(_P)]),
    "ap_engine_destroy": (C.c_int, [_P]),
    "ap_last_error": (C.c_char_p, [_P]),
    "ap_version": (C.c_char_p, []),
    "ap_sync": (C.c_int, [_P]),
    "ap_engine_memory": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "ap_engine_node_capacity": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "ap_boards_reset": (C.c_int, [_P, _P, _I, _P]),
    "ap_boards_do_move": (C.c_int, [_P, _P, _P, _I, _P]),
    "ap_boards_status": (C.c_int, [_P, _P, _I, _P, _P]),
    "ap_boards_legal": (C.c_int, [_P, _P, _I, _P]),
    "ap_boards_features": (C.c_int, [_P, _P, _I, _P]),
    "ap_boards_features_packed": (C.c_int, [_P, _P, _I, _P]),
    "ap_boards_export": (C.c_int, [_P, _P, _I, _P, _P]),
    "ap_boards_import": (C.c_int, [_P, _P, _I, _P, _P]),
    "ap_search_select": (C.c_int, [_P, _P, _P, _P]),
    "ap_search_leaf_export": (C.c_int, [_P, _P, _P]),
    "ap_search_leaf_features": (C.c_int, [_P, _P]),
    "ap_search_expand_backup": (C.c_int, [_P, _P, _P, _P, _P]),
    "ap_search_expand_backup_dense": (C.c_int, [_P, _P, _P]),
    "ap_search_run": (C.c_int, [_P, _I]),
    "ap_search_run_vl": (C.c_int, [_P, _I, _I]),
    "ap_search_root": (C.c_int, [_P, _P, _I, _P, _P, _P, _P, _P]),
    "ap_search_root_probs": (C.c_int, [_P, C.c_double, _P]),
    "ap_search_advance": (C.c_int, [_P, _P, _I, _P]),
    "ap_search_set_active": (C.c_int, [_P, _P]),
    "ap_search_stats": (C.c_int, [_P, _P]),
    "ap_selfplay_pick": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_uint32, _P, _P, _P]),
    "ap_pure_run": (C.c_int, [_P, _I, C.c_uint64, _I, _P]),
    "ap_rollout_eval": (C.c_int, [_P, C.c_uint64, _P, _P]),
    "ap_rollout_eval2": (C.c_int, [_P, C.c_uint64, _I, _P, _P]),
    "ap_rollout_eval_keys": (C.c_int, [_P, _P, _P, _P]),
    "ap_rollout_hash": (C.c_int, [_P, _P]),
    "ap_net_load": (C.c_int, [_P, _I, _I, _I, C.POINTER(ApTensor), _I]),
    "ap_net_forward": (C.c_int, [_P, _P, _I, _P, _P]),
    "ap_net_forward_precise": (C.c_int, [_P, _P, _I, _P, _P]),
    "ap_net_forward_leaves": (C.c_int, [_P, _I, _P, _P]),
    "ap_net_weights_ptr": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "ap_net_refresh": (C.c_int, [_P]),
    "ap_net_layout": (C.c_int, [_P, _I, _P, _P, _P]),
    "ap_search_timing": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "ap_search_profile": (C.c_int, [_P, _I, _P, _I]),
    "ap_launch_count": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "ap_replay_create": (C.c_int, [_P, C.c_int64]),
    "ap_replay_push": (C.c_int, [_P, _P, _P, _P, _I]),
    "ap_replay_size": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ap_replay_gather": (C.c_int, [_P, _P, _I, _P, _P, _P, _I]),
    "ap_replay_push_sgf": (C.c_int, [_P, _P, _I, _P, _P, _I, _P]),
    "ap_replay_push_packed": (C.c_int, [_P, _P, C.c_int64, _I]),
    "ap_traj_create": (C.c_int, [_P, _I, C.c_int64]),
    "ap_traj_append_forced": (C.c_int, [_P, _P, _I, _P]),
    "ap_traj_finish": (C.c_int, [_P, _P, _I, _P]),
    "ap_traj_discard": (C.c_int, [_P, _P, _I]),
    "ap_traj_outbox": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "ap_traj_outbox_clear": (C.c_int, [_P]),
}

_lib = None


def load():
    """Load the shared library (building it is ``alphapig_b200.build.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(AP_ERR_CUDA, "%s is missing: run `python -m alphapig_b200.build` (needs nvcc). "
                          "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
