"""
ctypes binding of the C ABI declared in ``include/dfdm_b200.h``.

This is the stub a maintainer of the reference would add to call the CUDA path from
``EquilibriumModel.__call__`` / ``Loss.__call__`` / ``Optimizer.minimize``.
The library is built in-tree by ``python -m dfdm_b200.build``; there is no fallback implementation:
if the shared object cannot be loaded, importing this module raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DFDM_B200_LIB") or os.path.join(HERE, "libdfdm_b200.so")

OK, EINVAL, ECUDA, ENOMEM, EUNSUPPORTED = 0, -1, -2, -3, -4
STATUS_OK, STATUS_SINGULAR, STATUS_NOT_CONVERGED, STATUS_NONFINITE = 0, 1, 2, 3
SOLVER_AUTO, SOLVER_DIRECT, SOLVER_PCG = 0, 1, 2
SOLVERS = {"auto": SOLVER_AUTO, "direct": SOLVER_DIRECT, "pcg": SOLVER_PCG}
PRECONDS = {"auto": 0, "jacobi": 1, "multigrid": 2}

ARR_FREE, ARR_FIXED, ARR_FREEFIXED, ARR_ROWPTR, ARR_COLIDX, ARR_EDGE_SLOTS, ARR_SOLVER_PERM = range(7)
ARR_NODE_INC_PTR, ARR_NODE_INC_EDGE, ARR_NODE_INC_SIGN = 7, 8, 9
ARR_MG_ROWS, ARR_MG_NNZ, ARR_PCG_PERM, ARR_MG_AGG = 10, 11, 12, 100
FLAG_SPMV_MATRIX_FREE = 1
FLAG_MG_NO_TAIL = 2
PCG_RTOL_DEFAULT, PCG_RTOL_ADJOINT_DEFAULT, PCG_RTOL_ADJOINT_JACOBI = 1e-12, 5e-9, 5e-10     # include/dfdm_b200.h DFDM_PCG_RTOL_*

GOAL_KINDS = {"NodePoint": 0, "NodeLine": 1, "NodePlane": 2, "NodeResidualForce": 3, "NodeResidualVector": 4,
              "NodeResidualDirection": 5, "EdgeLength": 6, "EdgeForce": 7, "EdgeLoadPath": 8, "EdgeVectorAngle": 9,
              "EdgeDirection": 10, "NetworkLoadPath": 11}
GOAL_NPARAMS = 6
TERM_KINDS = {"SquaredError": 0, "MeanSquaredError": 1, "PredictionError": 2}

# every symbol include/dfdm_b200.h declares
SYMBOLS = ["dfdm_last_error", "dfdm_abi_version", "dfdm_plan_create", "dfdm_plan_destroy", "dfdm_plan_get_info",
           "dfdm_plan_array", "dfdm_goalprog_create", "dfdm_goalprog_destroy", "dfdm_workspace_bytes", "dfdm_forward",
           "dfdm_loss_and_grad", "dfdm_vjp", "dfdm_forward_host", "dfdm_loss_and_grad_host", "dfdm_launch_count",
           "dfdm_last_pcg_iterations", "dfdm_last_direct_kernel", "dfdm_profile_enable", "dfdm_profile_read", "dfdm_peer_alloc", "dfdm_peer_open",
           "dfdm_peer_close", "dfdm_peer_free", "dfdm_peer_push"]


class Options(C.Structure):
    _fields_ = [("solver", C.c_int32), ("pcg_maxiter", C.c_int32), ("pcg_rtol", C.c_double),
                ("pcg_check_every", C.c_int32), ("pcg_precond", C.c_int32), ("mg_omega", C.c_double), ("mg_over", C.c_double),
                ("mg_w_levels", C.c_int32), ("flags", C.c_int32), ("mg_over_w", C.c_double), ("pcg_rtol_adjoint", C.c_double)]


class PlanInfo(C.Structure):
    _fields_ = [("n", C.c_int32), ("n_edges", C.c_int32), ("n_free", C.c_int32), ("n_fixed", C.c_int32),
                ("nnz", C.c_int32), ("bandwidth_natural", C.c_int32), ("bandwidth_solver", C.c_int32),
                ("direct_fits", C.c_int32), ("direct_smem_bytes", C.c_int64), ("auto_solver", C.c_int32),
                ("mg_levels", C.c_int32), ("mg_n_coarse", C.c_int32), ("mg_n_coarsest", C.c_int32)]


class DfdmError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("dfdm_b200: %s is missing; build it with `python -m dfdm_b200.build` "
                          "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.dfdm_last_error.restype = C.c_char_p
    lib.dfdm_abi_version.restype = i32
    lib.dfdm_plan_create.argtypes = [i32, i32, vp, vp, C.POINTER(vp)]
    lib.dfdm_plan_destroy.argtypes = [vp]
    lib.dfdm_plan_destroy.restype = None
    lib.dfdm_plan_get_info.argtypes = [vp, C.POINTER(PlanInfo)]
    lib.dfdm_plan_array.argtypes = [vp, i32, C.POINTER(i64)]
    lib.dfdm_plan_array.restype = C.POINTER(i32)
    lib.dfdm_goalprog_create.argtypes = [vp, i32, vp, vp, vp, vp, vp, i32, vp, vp, dbl, C.POINTER(vp)]
    lib.dfdm_goalprog_destroy.argtypes = [vp]
    lib.dfdm_goalprog_destroy.restype = None
    lib.dfdm_workspace_bytes.argtypes = [vp, i32, C.POINTER(Options)]
    lib.dfdm_workspace_bytes.restype = i64
    lib.dfdm_forward.argtypes = [vp, i32] + [vp] * 3 + [vp] * 5 + [vp, vp, i64, C.POINTER(Options), vp]
    lib.dfdm_loss_and_grad.argtypes = [vp, vp, i32] + [vp] * 3 + [vp, vp] + [vp] * 5 + [vp, vp, i64, C.POINTER(Options), vp]
    lib.dfdm_vjp.argtypes = [vp, i32] + [vp] * 3 + [vp] * 5 + [vp, vp, vp, i64, C.POINTER(Options), vp]
    lib.dfdm_forward_host.argtypes = [vp, i32] + [vp] * 3 + [vp] * 5 + [vp, C.POINTER(Options)]
    lib.dfdm_loss_and_grad_host.argtypes = [vp, vp, i32] + [vp] * 3 + [vp, vp, vp, C.POINTER(Options)]
    lib.dfdm_launch_count.restype = i64
    lib.dfdm_last_pcg_iterations.argtypes = [C.POINTER(i32), C.POINTER(i32)]
    lib.dfdm_last_pcg_iterations.restype = None
    lib.dfdm_last_direct_kernel.argtypes = []
    lib.dfdm_last_direct_kernel.restype = C.c_char_p
    lib.dfdm_profile_enable.argtypes = [i32]
    lib.dfdm_profile_enable.restype = None
    lib.dfdm_profile_read.argtypes = [C.POINTER(dbl), C.POINTER(i64)]
    lib.dfdm_profile_read.restype = None
    lib.dfdm_peer_alloc.argtypes = [i64, C.POINTER(vp), C.c_char_p]
    lib.dfdm_peer_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    lib.dfdm_peer_close.argtypes = [vp]
    lib.dfdm_peer_free.argtypes = [vp]
    lib.dfdm_peer_push.argtypes = [vp, i64, i64, C.POINTER(vp), i32, vp]
    for name in SYMBOLS:
        getattr(lib, name)
    if lib.dfdm_abi_version() != 3:
        raise ImportError("dfdm_b200: ABI version mismatch between _lib.py and libdfdm_b200.so")
    return lib


lib = _load()


def check(rc):
    if rc != OK:
        raise DfdmError("dfdm_b200 error %d: %s" % (rc, (lib.dfdm_last_error() or b"").decode()))


def make_options(solver="auto", pcg_maxiter=0, pcg_rtol=0.0, pcg_check_every=0, pcg_precond="auto", mg_omega=0.0, mg_over=0.0,
                 mg_w_levels=0, mg_over_w=0.0, flags=0, pcg_rtol_adjoint=0.0):
    if isinstance(solver, str):
        solver = SOLVERS[solver]
    if isinstance(pcg_precond, str):
        pcg_precond = PRECONDS[pcg_precond]
    return Options(solver, int(pcg_maxiter), float(pcg_rtol), int(pcg_check_every), int(pcg_precond), float(mg_omega), float(mg_over),
                   int(mg_w_levels), int(flags), float(mg_over_w), float(pcg_rtol_adjoint))


def plan_array(handle, which):
    length = C.c_int64(0)
    ptr = lib.dfdm_plan_array(handle, which, C.byref(length))
    if not ptr:
        raise DfdmError((lib.dfdm_last_error() or b"").decode())
    if length.value == 0:
        return np.zeros(0, dtype=np.int32)
    return np.ctypeslib.as_array(ptr, shape=(length.value,)).copy()


def launch_count():
    return int(lib.dfdm_launch_count())


def last_direct_kernel():
    return lib.dfdm_last_direct_kernel().decode()


def last_pcg_iterations():
    f, a = C.c_int32(0), C.c_int32(0)
    lib.dfdm_last_pcg_iterations(C.byref(f), C.byref(a))
    return f.value, a.value


def profile_enable(on=True):
    lib.dfdm_profile_enable(1 if on else 0)


def profile_read():
    ms = (C.c_double * 8)()
    count = (C.c_int64 * 8)()
    lib.dfdm_profile_read(ms, count)
    return list(ms), list(count)
