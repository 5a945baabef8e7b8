"""ctypes binding of libba_b200.so (the C ABI declared in include/ba_b200.h).

There is deliberately no fallback: if the shared library has not been built, or a call
returns a CUDA error, this module raises.  The bundle-adjustment path never runs on the CPU.
"""
import ctypes
import os

from . import build as _build

BA_OK = 0
BA_ERR_ILLCONDITIONED = 1
BA_ERR_CUDA = 2
BA_ERR_NCCL = 3
BA_ERR_BAD_ARGUMENT = 4
BA_ERR_NOT_BOUND = 5
BA_ERR_NONFINITE = 6
BA_ERR_TIMEOUT = 7

(BA_OPT_SPIN_TIMEOUT_MS, BA_OPT_STRICT_FLAGS, BA_OPT_DIST_SOLVE_MIN_TILES, BA_OPT_DIST_BAND,
 BA_OPT_SOLVE_GRID_CAP, BA_OPT_SOLVER_PROFILE, BA_OPT_FUSE_COST_REDUCTION,
 BA_OPT_TC_MIN_TILES, BA_OPT_TC_SLICES, BA_OPT_TC_WINDOW, BA_OPT_TC_BK, BA_OPT_TC_OVER_DIST_MAX_WORLD) = range(12)
SOLVER_PROFILE_SLOTS = ["panel_tasks", "wait_k_operands", "last_step_polls", "panel_idle_rounds", "wait_diag_task",
                        "peer_flags", "peer_contributions", "wait_y", "backward_waits", "start_barrier", "launch",
                        "chain_tasks", "diag_tasks"]

BA_WANT_BLOCKS = 1
BA_WANT_SCHUR = 2

(BA_ARR_HCC, BA_ARR_HPP, BA_ARR_HCP, BA_ARR_BC, BA_ARR_BP, BA_ARR_HPP_INV, BA_ARR_DC,
 BA_ARR_DP, BA_ARR_RESIDUAL, BA_ARR_JC, BA_ARR_JP) = range(11)

_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_double_p = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p

# name -> (restype, argtypes); kept in one table so tests can check it against the header.
SIGNATURES = {
    "ba_version": (ctypes.c_char_p, []),
    "ba_last_error": (ctypes.c_char_p, [_vp]),
    "ba_system_ld": (ctypes.c_int, [ctypes.c_int]),
    "ba_system_size": (ctypes.c_size_t, [ctypes.c_int]),
    "ba_create": (ctypes.c_int, [ctypes.c_int] * 6 + [ctypes.POINTER(_vp)]),
    "ba_destroy": (ctypes.c_int, [_vp]),
    "ba_set_intrinsics": (ctypes.c_int, [_vp, _c_double_p]),
    "ba_set_sensor_model": (ctypes.c_int, [_vp, ctypes.c_int, _c_double_p]),
    "ba_bind_structure": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "ba_bind_state": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "ba_bind_candidate": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "ba_bind_system": (ctypes.c_int, [_vp, _vp]),
    "ba_upload_system": (ctypes.c_int, [_vp, _vp, _vp]),
    "ba_get_system": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t, _vp]),
    "ba_set_option": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_double]),
    "ba_dist_solve_active": (ctypes.c_int, [_vp]),
    "ba_tc_solve_active": (ctypes.c_int, [_vp]),
    "ba_linearize_eliminate": (ctypes.c_int, [_vp, ctypes.c_double, ctypes.c_double, ctypes.c_int, _vp]),
    "ba_solve": (ctypes.c_int, [_vp, _vp, _vp]),
    "ba_backsub_retract_cost": (ctypes.c_int, [_vp, _vp]),
    "ba_cost": (ctypes.c_int, [_vp, _vp]),
    "ba_accept": (ctypes.c_int, [_vp]),
    "ba_read_scalars": (ctypes.c_int, [_vp, _c_double_p, _c_double_p, _c_int_p, _vp]),
    "ba_trial_host": (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_double, ctypes.c_double, _vp, _vp, _vp,
                                     _c_double_p, _c_double_p, _c_int_p, _vp]),
    "ba_trial_host_packed": (ctypes.c_int, [_vp, _vp, ctypes.c_double, ctypes.c_double, _vp, _vp, _vp]),
    "ba_scalars_ptr": (ctypes.c_int, [_vp, ctypes.POINTER(_vp)]),
    "ba_comm_create": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp]),
    "ba_comm_connect": (ctypes.c_int, [_vp, _vp]),
    "ba_comm_disconnect": (ctypes.c_int, [_vp]),
    "ba_comm_system_ptr": (ctypes.c_int, [_vp, ctypes.POINTER(_vp)]),
    "ba_allreduce_system": (ctypes.c_int, [_vp, _vp]),
    "ba_allreduce_costs": (ctypes.c_int, [_vp, _vp]),
    "ba_eval_observations": (ctypes.c_int, [_vp, _vp]),
    "ba_get_array": (ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_size_t, _vp]),
    "ba_retract": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "ba_triangulate": (ctypes.c_int, [_vp, _vp]),
    "ba_pack_observations": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                            _vp, _c_int_p, _vp]),
    "ba_set_solution": (ctypes.c_int, [_vp, _vp, _vp]),
    "ba_sync": (ctypes.c_int, [_vp, _vp]),
    "ba_solver_profile": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp]),
    "ba_tc_solve_profile": (ctypes.c_int, [_vp, _vp, ctypes.c_int]),
    "ba_launch_count": (ctypes.c_longlong, [_vp]),
    "ba_tc_trailing_update_host": (ctypes.c_int, [ctypes.c_int] * 5 + [_vp] * 6),
}

_lib = None


class BAError(RuntimeError):
    pass


def library_path():
    return _build.LIB_PATH


def load():
    """Load (once) and return the ctypes library; raise loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.isfile(path):
        raise BAError(
            "%s is missing: build it with `python -m pysfm_b200.build` (needs nvcc). "
            "pysfm_b200 has no CPU fallback for the bundle-adjustment path." % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(handle, rc, what):
    """Translate a status code into the exception the reference's callers expect."""
    if rc == BA_OK:
        return
    lib = load()
    detail = lib.ba_last_error(handle).decode() if handle else ""
    if rc == BA_ERR_BAD_ARGUMENT:
        raise AssertionError("%s: bad argument %s" % (what, detail))
    names = {BA_ERR_CUDA: "CUDA error", BA_ERR_NCCL: "NCCL error", BA_ERR_NOT_BOUND: "buffers not bound",
             BA_ERR_ILLCONDITIONED: "ill-conditioned", BA_ERR_NONFINITE: "non-finite cost",
             BA_ERR_TIMEOUT: "a device-side wait timed out (peer lost or kernel fault)"}
    raise BAError("%s failed: %s %s" % (what, names.get(rc, "status %d" % rc), detail))
