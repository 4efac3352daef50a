"""ctypes binding of libddp.so (include/ddp.h).  Fails loudly when the CUDA library is missing:
there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libddp.so")

c_double_p = C.POINTER(C.c_double)


class Tensor(C.Structure):          # ddp_tensor
    _fields_ = [("ptr", C.c_void_p), ("stride_b", C.c_int64), ("stride_t", C.c_int64)]


class BoxQPOpts(C.Structure):       # ddp_boxqp_opts
    _fields_ = [("max_iter", C.c_int32), ("min_grad", C.c_double), ("min_rel_improve", C.c_double),
                ("step_dec", C.c_double), ("min_step", C.c_double), ("armijo", C.c_double)]


class BackPassArgs(C.Structure):    # ddp_back_pass_args
    _fields_ = [("cx", Tensor), ("cu", Tensor), ("cxx", Tensor), ("cxu", Tensor), ("cuu", Tensor),
                ("fx", Tensor), ("fu", Tensor), ("lam", C.c_void_p), ("reg_type", C.c_int32),
                ("lims", C.c_void_p), ("lims_stride_t", C.c_int64), ("u", Tensor), ("active", C.c_void_p),
                ("fxx", Tensor), ("fxu", Tensor), ("fuu", Tensor),
                ("diverge", C.c_void_p), ("K", C.c_void_p), ("k", C.c_void_p), ("Vx", C.c_void_p),
                ("Vxx", C.c_void_p), ("Vxx1", C.c_void_p), ("Quu", C.c_void_p), ("dV", C.c_void_p),
                ("Vxx_tri", C.c_void_p), ("Quu_tri", C.c_void_p),
                ("qp", BoxQPOpts)]


class GpsArgs(C.Structure):         # ddp_gps_args
    _fields_ = [("K_prev", Tensor), ("k_prev", Tensor), ("Sigi_prev", Tensor), ("eta", C.c_void_p),
                ("Quui", C.c_void_p), ("Quui_tri", C.c_void_p)]


class Model(C.Structure):           # ddp_model
    _fields_ = [("kind", C.c_int32), ("A", Tensor), ("Bm", Tensor), ("Q", Tensor), ("R", Tensor),
                ("goal", C.c_void_p), ("p", C.c_double * 8), ("terminal_cost", C.c_int32), ("flags", C.c_int32)]


class ForwardPassArgs(C.Structure):  # ddp_forward_pass_args
    _fields_ = [("K", C.c_void_p), ("k", C.c_void_p), ("x0", Tensor), ("x", Tensor), ("u", Tensor),
                ("alpha", C.c_void_p), ("alpha_scalar", C.c_double), ("u_scale", C.c_double),
                ("lims", C.c_void_p), ("lims_stride_t", C.c_int64), ("active", C.c_void_p),
                ("xnew", C.c_void_p), ("unew", C.c_void_p), ("cost", C.c_void_p), ("cost_t", C.c_void_p),
                ("cx", C.c_void_p), ("cu", C.c_void_p)]


class KlArgs(C.Structure):          # ddp_kl_args
    _fields_ = [("fx", Tensor), ("R1", Tensor), ("xnew", C.c_void_p), ("xold", C.c_void_p),
                ("K_new", C.c_void_p), ("k_new", C.c_void_p), ("Sig_new", C.c_void_p),
                ("K_prev", Tensor), ("k_prev", Tensor), ("Sig_prev", Tensor), ("Sigi_prev", Tensor),
                ("kl_t", C.c_void_p), ("kl_mean", C.c_void_p), ("Sx_tri", C.c_void_p), ("Sx_mode", C.c_int32), ("pad_", C.c_int32), ("Sx_count", C.c_int64)]


class IlqgOpts(C.Structure):        # ddp_ilqg_opts
    _fields_ = [("n_alpha", C.c_int32), ("alpha", C.c_double * 16), ("tol_fun", C.c_double),
                ("tol_grad", C.c_double), ("max_iter", C.c_int32), ("lam", C.c_double),
                ("dlam", C.c_double), ("lam_factor", C.c_double), ("lam_max", C.c_double),
                ("lam_min", C.c_double), ("reg_type", C.c_int32), ("reduce_ratio_min", C.c_double),
                ("lims", C.c_void_p), ("x_init", C.c_void_p), ("cost_init", C.c_void_p),
                ("trace", C.c_void_p), ("trace_cap", C.c_int32)]


class IlqgTrace(C.Structure):       # ddp_ilqg_trace
    _fields_ = [("lam", C.c_double), ("dlam", C.c_double), ("cost", C.c_double), ("alpha", C.c_double),
                ("grad_norm", C.c_double), ("improvement", C.c_double), ("reduce_ratio", C.c_double),
                ("accepted", C.c_int32), ("bp_retries", C.c_int32)]


class IlqgState(C.Structure):       # ddp_ilqg_state
    _fields_ = [("lam", C.c_double), ("dlam", C.c_double), ("cost", C.c_double), ("g_norm", C.c_double),
                ("last_dcost", C.c_double), ("last_alpha", C.c_double), ("iter", C.c_int32),
                ("accepted_iter", C.c_int32), ("status", C.c_int32), ("pad", C.c_int32)]


class IlqgklOpts(C.Structure):      # ddp_ilqgkl_opts
    _fields_ = [("kl_step", C.c_double), ("max_iter", C.c_int32), ("eta_bracket", C.c_double * 3),
                ("del0", C.c_double), ("max_eta_retries", C.c_int32), ("lims", C.c_void_p), ("no_covariance_cache", C.c_int32), ("pad_", C.c_int32)]


class IlqgklState(C.Structure):     # ddp_ilqgkl_state
    _fields_ = [("eta_min", C.c_double), ("eta", C.c_double), ("eta_max", C.c_double), ("del0", C.c_double),
                ("divergence", C.c_double), ("dcost", C.c_double), ("expected", C.c_double), ("cost", C.c_double),
                ("iter", C.c_int32), ("status", C.c_int32), ("retries", C.c_int32), ("pad", C.c_int32)]


class IlqgklArgs(C.Structure):      # ddp_ilqgkl_args
    _fields_ = [("x", C.c_void_p), ("u", C.c_void_p), ("cost", C.c_void_p),
                ("K_prev", Tensor), ("Sig_prev", Tensor), ("Sigi_prev", Tensor), ("fx_model", Tensor), ("R1", Tensor),
                ("xnew", C.c_void_p), ("unew", C.c_void_p), ("K", C.c_void_p), ("k", C.c_void_p),
                ("Sig", C.c_void_p), ("Sigi", C.c_void_p), ("Vx", C.c_void_p), ("Vxx1", C.c_void_p),
                ("costnew", C.c_void_p), ("state", C.c_void_p)]


class IterHostArgs(C.Structure):    # ddp_iter_host_args
    _fields_ = [("fx", C.c_void_p), ("fu", C.c_void_p), ("cx", C.c_void_p), ("cu", C.c_void_p),
                ("x", C.c_void_p), ("u", C.c_void_p), ("lam", C.c_void_p),
                ("Q", C.c_void_p), ("R", C.c_void_p), ("cxu", C.c_void_p),
                ("reg_type", C.c_int32), ("q_diagonal", C.c_int32), ("alpha", C.c_double),
                ("xnew", C.c_void_p), ("unew", C.c_void_p), ("cost", C.c_void_p), ("dV", C.c_void_p),
                ("diverge", C.c_void_p), ("chunk", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("keep_policy", C.c_int32), ("inputs_resident", C.c_int32), ("commit_accepted", C.c_int32), ("pad_", C.c_int32),
                ("cost_prev", C.c_void_p)]


class IterArgs(C.Structure):        # ddp_iter_args
    _fields_ = [("x", C.c_void_p), ("u", C.c_void_p), ("lam", C.c_void_p), ("alpha", C.c_void_p),
                ("alpha_scalar", C.c_double), ("reg_type", C.c_int32), ("pad_", C.c_int32),
                ("lims", C.c_void_p), ("active", C.c_void_p),
                ("xnew", C.c_void_p), ("unew", C.c_void_p), ("cost", C.c_void_p), ("dV", C.c_void_p),
                ("diverge", C.c_void_p), ("K", C.c_void_p), ("k", C.c_void_p), ("Vx", C.c_void_p),
                ("chunk", C.c_int64), ("n_chunks", C.c_int64)]


# every symbol include/ddp.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("ddp_version", C.c_int, []),
    ("ddp_device_count", C.c_int, []),
    ("ddp_create", C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_uint32]),
    ("ddp_destroy", C.c_int, [C.c_void_p]),
    ("ddp_last_error", C.c_char_p, [C.c_void_p]),
    ("ddp_set_stream", C.c_int, [C.c_void_p, C.c_void_p]),
    ("ddp_synchronize", C.c_int, [C.c_void_p]),
    ("ddp_kernel_variant", C.c_char_p, [C.c_void_p]),
    ("ddp_launch_count", C.c_int64, [C.c_void_p]),
    ("ddp_malloc", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_size_t]),
    ("ddp_free", C.c_int, [C.c_void_p, C.c_void_p]),
    ("ddp_memset", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t]),
    ("ddp_upload", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    ("ddp_download", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    ("ddp_host_alloc", C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    ("ddp_host_free", C.c_int, [C.c_void_p]),
    ("ddp_back_pass_f64", C.c_int, [C.c_void_p, C.POINTER(BackPassArgs)]),
    ("ddp_back_pass_gps_f64", C.c_int, [C.c_void_p, C.POINTER(BackPassArgs), C.POINTER(GpsArgs)]),
    ("ddp_boxqp_f64", C.c_int, [C.c_void_p, C.c_int64] + [C.c_void_p] * 5 + [C.POINTER(BoxQPOpts)] + [C.c_void_p] * 5),
    ("ddp_boxqp_large_f64", C.c_int, [C.c_void_p, C.c_int32, C.c_int64] + [C.c_void_p] * 5 + [C.POINTER(BoxQPOpts)] + [C.c_void_p] * 5),
    ("ddp_forward_pass_f64", C.c_int, [C.c_void_p, C.POINTER(Model), C.POINTER(ForwardPassArgs)]),
    ("ddp_batch_stats_f64", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    ("ddp_kl_div_f64", C.c_int, [C.c_void_p, C.POINTER(KlArgs)]),
    ("ddp_model_derivs_f64", C.c_int, [C.c_void_p, C.POINTER(Model)] + [C.c_void_p] * 6),
    ("ddp_ilqg_solve_f64", C.c_int, [C.c_void_p, C.POINTER(Model), C.POINTER(IlqgOpts)] + [C.c_void_p] * 9 +
     [C.POINTER(C.c_int32)]),
    ("ddp_forward_costs_multi_f64", C.c_int, [C.c_void_p, C.POINTER(Model), C.POINTER(ForwardPassArgs), C.c_int32,
                                              C.POINTER(C.c_double), C.c_void_p]),
    ("ddp_comm_unique_id", C.c_int, [C.c_void_p]),
    ("ddp_comm_init", C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    ("ddp_comm_allreduce_stats_f64", C.c_int, [C.c_void_p, C.c_void_p]),
    ("ddp_comm_destroy", C.c_int, [C.c_void_p]),
    ("ddp_ilqgkl_solve_f64", C.c_int, [C.c_void_p, C.POINTER(Model), C.POINTER(IlqgklOpts), C.POINTER(IlqgklArgs),
                                       C.POINTER(C.c_int32)]),
    ("ddp_ilqg_iter_host_f64", C.c_int, [C.c_void_p, C.POINTER(IterHostArgs)]),
    ("ddp_iter_host_policy", C.c_int, [C.c_void_p] + [C.POINTER(C.c_void_p)] * 5),
    ("ddp_ilqg_iter_f64", C.c_int, [C.c_void_p, C.POINTER(Model), C.POINTER(IterArgs)]),
    ("ddp_selftest_peak_f64", C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
]

# Python mirror <-> C struct, for the layout tests (sizeof and offsetof of every field, tests/test_cpu_host.py)
STRUCTS = {"ddp_tensor": Tensor, "ddp_boxqp_opts": BoxQPOpts, "ddp_back_pass_args": BackPassArgs, "ddp_gps_args": GpsArgs,
           "ddp_model": Model, "ddp_forward_pass_args": ForwardPassArgs, "ddp_kl_args": KlArgs, "ddp_ilqg_opts": IlqgOpts,
           "ddp_ilqg_trace": IlqgTrace, "ddp_ilqg_state": IlqgState, "ddp_ilqgkl_opts": IlqgklOpts,
           "ddp_ilqgkl_state": IlqgklState, "ddp_ilqgkl_args": IlqgklArgs, "ddp_iter_host_args": IterHostArgs,
           "ddp_iter_args": IterArgs}
# ctypes field name -> C field name where they differ (Python keywords / shorter names)
FIELD_ALIASES = {"lam": "lambda", "dlam": "dlambda", "lam_factor": "lambda_factor", "lam_max": "lambda_max", "lam_min": "lambda_min"}

_lib = None


class DDPLibraryMissing(ImportError):
    pass


def load() -> C.CDLL:
    """Load libddp.so.  Raises DDPLibraryMissing when it has not been built: the package has no
    CPU or PyTorch fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DDPLibraryMissing(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)      # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class DDPError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libddp error {code}: {msg}")
        self.code = code
