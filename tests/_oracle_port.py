"""ctypes wrapper of oracle/_ref/lib/libncb200_oracle.so -- the oracle's plain-C restatement of the
reference algorithm (oracle/oracle_*.c).  TEST INFRASTRUCTURE: checker only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_ref", "lib", "libncb200_oracle.so")
_dp = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)


def _d(a):
    return a.ctypes.data_as(_dp)


_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
        L = C.CDLL(LIB)
        L.orc_load.restype = C.c_void_p
        L.orc_load.argtypes = [C.c_char_p, C.c_uint64]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_error.restype = C.c_char_p
        L.orc_error.argtypes = [C.c_void_p]
        L.orc_ncomp.argtypes = [C.c_void_p]
        L.orc_xs_iso_many.argtypes = [C.c_void_p, _dp, C.c_uint64, _dp]
        L.orc_sample_iso_many.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, _dp, C.c_uint64, _dp, _dp, _u32p, _i32p]
        L.orc_xs_many.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_uint64, _dp]
        L.orc_sample_many.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, _dp, _dp, _dp, _dp, C.c_uint64,
                                      _dp, _dp, _dp, _dp, _u32p, _i32p]
        L.orc_sab_xscheck.argtypes = [C.c_void_p, C.c_int, _dp]
        L.orc_sab_sampler_dump.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.orc_bench.restype = C.c_double
        L.orc_bench.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, C.c_uint64, _dp, _dp]
        _lib = L
    return _lib


class PortOracle:
    kind = "port"

    def __init__(self, blob):
        L = lib()
        self.h = L.orc_load(blob, len(blob))
        if not self.h:
            raise RuntimeError("oracle: could not load compiled material")
        err = L.orc_error(self.h)
        if err:
            raise RuntimeError("oracle: %s" % err.decode())

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_free(self.h)
            self.h = None

    def xs_iso(self, ekin):
        ekin = np.ascontiguousarray(ekin, dtype=np.float64)
        out = np.empty_like(ekin)
        lib().orc_xs_iso_many(self.h, _d(ekin), ekin.size, _d(out))
        return out

    def sample_iso(self, ekin, seed, first_index=0):
        ekin = np.ascontiguousarray(ekin, dtype=np.float64)
        eo, mu = np.empty_like(ekin), np.empty_like(ekin)
        nd = np.zeros(ekin.size, dtype=np.uint32)
        er = np.zeros(ekin.size, dtype=np.int32)
        lib().orc_sample_iso_many(self.h, seed, first_index, _d(ekin), ekin.size, _d(eo), _d(mu),
                                  nd.ctypes.data_as(_u32p), er.ctypes.data_as(_i32p))
        return eo, mu, nd, er

    def xs(self, ekin, ux, uy, uz):
        ekin, ux, uy, uz = [np.ascontiguousarray(a, dtype=np.float64) for a in (ekin, ux, uy, uz)]
        out = np.empty_like(ekin)
        lib().orc_xs_many(self.h, _d(ekin), _d(ux), _d(uy), _d(uz), ekin.size, _d(out))
        return out

    def sample(self, ekin, ux, uy, uz, seed, first_index=0):
        ekin, ux, uy, uz = [np.ascontiguousarray(a, dtype=np.float64) for a in (ekin, ux, uy, uz)]
        eo, ox, oy, oz = [np.empty_like(ekin) for _ in range(4)]
        nd = np.zeros(ekin.size, dtype=np.uint32)
        er = np.zeros(ekin.size, dtype=np.int32)
        lib().orc_sample_many(self.h, seed, first_index, _d(ekin), _d(ux), _d(uy), _d(uz), ekin.size,
                              _d(eo), _d(ox), _d(oy), _d(oz), nd.ctypes.data_as(_u32p), er.ctypes.data_as(_i32p))
        return eo, ox, oy, oz, nd, er

    def sab_sampler_dump(self, c, iE, nbeta):
        from _libs import _sab_dump
        return _sab_dump(lib().orc_sab_sampler_dump, self.h, c, iE, nbeta)

    def sab_xscheck(self, c, negrid):
        out = np.zeros(negrid)
        n = lib().orc_sab_xscheck(self.h, c, _d(out))
        return out[:n]

    def bench(self, mode, nthreads, ekin):
        ekin = np.ascontiguousarray(ekin, dtype=np.float64)
        o0, o1 = np.empty_like(ekin), np.empty_like(ekin)
        return lib().orc_bench(self.h, mode, nthreads, _d(ekin), ekin.size, _d(o0), _d(o1))
