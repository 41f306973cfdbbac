"""ctypes binding of libbinest.so (include/binest.h).  There is no CPU fallback: loading fails loudly
when the CUDA library has not been built, and every compute call fails when no B200 is present."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BINEST_LIB") or os.path.join(_HERE, "libbinest.so")  # BINEST_LIB: A/B builds (csrc/Makefile)

LOGZERO = -1.7976931348623157e308  # -$MaxMachineNumber, what BU:47 evaluates to on IEEE hardware

STATUS = {0: "OK", 1: "TYPE", 2: "RANK", 3: "DIMENSION", 4: "NUMERICAL", 5: "MEMORY", 6: "FUNCTION", 7: "CUDA",
          8: "BAD_LIKELIHOOD"}


class BinestError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"binest error {code} ({STATUS.get(code, '?')}): {msg}")
        self.code = code


class Options(C.Structure):
    """binest_options — the flattened option set of nestedSampling (BS:837-851)."""
    _fields_ = [("pool_size", C.c_int64), ("batch_k", C.c_int64), ("mc_steps", C.c_int64),
                ("max_iter", C.c_int64), ("min_iter", C.c_int64), ("term_frac", C.c_double),
                ("acc_min", C.c_double), ("acc_max", C.c_double), ("seed", C.c_uint64),
                ("first_run_id", C.c_int64), ("n_runs", C.c_int64), ("loglmax", C.c_double)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes): every symbol include/binest.h declares
SIGNATURES = {
    "binest_version": (C.c_int, []),
    "binest_last_error": (C.c_char_p, []),
    "binest_init": (C.c_int, [C.c_double, C.c_int]),
    "binest_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "binest_default_options": (None, [C.POINTER(Options)]),
    "binest_measure_fp64_peak": (C.c_int, [_dp, _dp]),
    "binest_launch_count": (C.c_int64, []),
    "binest_problem_create": (C.c_int, [C.c_int, _ip, _dp, C.c_int64, C.c_int64, _dp, C.c_int64, C.c_int64, _i32p,
                                        _dp, _dp, _dp, _dp, C.POINTER(_vp)]),
    "binest_problem_free": (C.c_int, [_vp]),
    "binest_problem_dim": (C.c_int, [_vp, _ip]),
    "binest_loglike": (C.c_int, [_vp, _dp, C.c_int64, _dp]),
    "binest_logprior": (C.c_int, [_vp, _dp, C.c_int64, _dp]),
    "binest_predictive_width": (C.c_int, [_vp, _ip]),
    "binest_predictive_components": (C.c_int, [_vp, _dp, C.c_int64, _dp, C.c_int64, _dp]),
    "binest_gp_predict": (C.c_int, [_vp, _dp, C.c_int64, _dp, C.c_int64, _dp, _dp]),
    "binest_sample_prior": (C.c_int, [_vp, C.c_int64, C.c_uint64, C.c_int64, _dp]),
    "binest_run_create": (C.c_int, [_vp, C.POINTER(Options), _dp, C.POINTER(_vp)]),
    "binest_run_advance": (C.c_int, [_vp, C.c_int64, _i32p]),
    "binest_run_sizes": (C.c_int, [_vp, C.c_int64, _ip, _ip, _ip, _ip]),
    "binest_run_fetch": (C.c_int, [_vp, C.c_int64, _dp, _dp, _dp, _dp, _ip, _dp, _dp, _dp]),
    "binest_run_estimates": (C.c_int, [_vp, C.c_int64, _dp, _dp]),
    "binest_run_free": (C.c_int, [_vp]),
    "binest_chain_create": (C.c_int, [_vp, _dp, C.c_int64, _dp, C.c_int64, C.c_uint64, C.POINTER(_vp)]),
    "binest_chain_iterate": (C.c_int, [_vp, C.c_int64, _dp]),
    "binest_chain_state": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _ip, _ip]),
    "binest_chain_free": (C.c_int, [_vp]),
    "binest_evidence_sampling": (C.c_int, [C.c_int64, C.c_int64, _dp, _dp, _ip, C.c_int64, C.c_int64, C.c_uint64,
                                           _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "binest_crude_weights": (C.c_int, [C.c_int64, _dp, _ip, C.c_int64, _dp, _dp, _dp]),
    "binest_merge_runs": (C.c_int, [C.c_int64, _ip, C.c_int64, _dp, _dp, _dp, _dp, _ip, _ip, _dp, _dp, _dp, _dp, _ip, _ip, _ip, _ip]),
    "binest_combine_runs": (C.c_int, [C.c_int64, _ip, C.c_int64, _dp, _dp, _dp, _dp, _ip, _ip, C.c_int32, C.c_int64,
                                      C.c_int64, C.c_uint64, _dp, _dp, _ip, _dp, _dp, _dp, _dp, _ip, _ip]),
    "binest_run_merge_size": (C.c_int, [_vp, _ip]),
    "binest_run_merge": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _ip, _ip, _ip, _ip]),
    "binest_run_combine": (C.c_int, [_vp, C.c_int32, C.c_int64, C.c_uint64, _dp, _dp, _ip, _dp, _dp, _dp, _dp, _ip, _ip]),
    "binest_run_merge_dev": (C.c_int, [_vp, _vp, _ip, _ip]),
    "binest_combine_runs_dev": (C.c_int, [C.c_int64, _ip, C.c_int64, _vp, C.c_int32, C.c_int64, C.c_int64, C.c_uint64,
                                          _dp, _dp, _ip, _dp, _dp, _dp, _dp, _ip, _ip]),
    "binest_bench_loglike": (C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int64, C.c_int, _dp, _dp]),
    "binest_run_timing": (C.c_int, [_vp, _dp, _ip, _ip]),
    "binest_run_path": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "binest_problem_stream": (C.c_int, [_vp, C.POINTER(C.c_void_p)]),
    "binest_comm_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "binest_comm_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_uint8), C.POINTER(_vp)]),
    "binest_comm_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "binest_comm_free": (C.c_int, [_vp]),
    "binest_problem_shard": (C.c_int, [_vp, _vp]),
    "binest_problem_shard_batch": (C.c_int, [_vp, _vp]),
    "binest_comm_stats": (C.c_int, [_vp, _ip, _ip, C.POINTER(C.c_int)]),
}
COMM_ID_BYTES = 128

_lib = None


def load():
    """Load libbinest.so; raises if it is missing (build with __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not built — run `python -c 'import __graft_entry__ as g; g.build()'`; "
                              "there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status):
    if status != 0:
        raise BinestError(status, load().binest_last_error().decode())


def dptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def iptr(a):
    return None if a is None else a.ctypes.data_as(_ip)
