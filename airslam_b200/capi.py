"""ctypes binding of include/airfe_c.h.  Fails loudly if libairfe.so is missing: there is no fallback."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libairfe.so")


class AirfeError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AirfeError("libairfe.so not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                             "there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.airfe_last_error.restype = C.c_char_p
        _declare(_lib)
    return _lib


def check(rc):
    if rc != 0:
        raise AirfeError("airfe error %d: %s" % (rc, lib().airfe_last_error().decode()))


vp, i32, i64, f32p = C.c_void_p, C.c_int, C.c_longlong, C.c_void_p


def _declare(L):
    L.airfe_op_tc_gemm.argtypes = [vp, i32, i32, i32, i32, i64, i64, i64,
                                   vp, i32, i32, i64, i64, i32, i32,
                                   i32, i32, i32, f32p, i32, i32,
                                   vp, i64, i64, i64, i32, i32, i32, i32, vp]
    L.airfe_op_tc_gemm.restype = i32
