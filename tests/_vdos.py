"""Helpers of the VDOS -> S(alpha,beta) tests: one ctypes binding of the reference's own C entry points
(ncrystal_raw_vdos2kernel / ncrystal_raw_vdos2gn, ncrystal.h:885-925) that is pointed either at the unmodified
reference (oracle/_ref/lib/libNCrystal.so), at the product (libncrystal_b200.so) or at the TEST-ONLY host build, plus
the golden-vector cases (tests/golden/vdos_reference.npz, made by tests/golden/make_golden_vdos.py)."""
import ctypes as C
import hashlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFLIB_PATH = os.path.join(ROOT, "oracle", "_ref", "lib", "libNCrystal.so")
GOLDEN = os.path.join(ROOT, "tests", "golden", "vdos_reference.npz")
_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint)
WEIGHT_FCT = C.CFUNCTYPE(C.c_double, C.c_uint)


def _d(a):
    return a.ctypes.data_as(_dp)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


class RawVdosAPI:
    """ncrystal_raw_vdos2kernel / ncrystal_raw_vdos2gn of any library that exports them."""

    def __init__(self, lib):
        self.L = lib
        lib.ncrystal_raw_vdos2kernel.restype = None
        lib.ncrystal_raw_vdos2kernel.argtypes = [_dp, _dp, C.c_uint, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_uint,
                                                 C.c_void_p, _up, _up, C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_dp),
                                                 C.c_double, _dp]
        lib.ncrystal_raw_vdos2gn.restype = None
        lib.ncrystal_raw_vdos2gn.argtypes = [_dp, _dp, C.c_uint, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_uint,
                                             _dp, _dp, _up, C.POINTER(_dp)]
        lib.ncrystal_dealloc_doubleptr.restype = None
        lib.ncrystal_dealloc_doubleptr.argtypes = [_dp]

    def _take(self, p, n):
        a = np.ctypeslib.as_array(p, (n,)).copy()
        self.L.ncrystal_dealloc_doubleptr(p)
        return a

    def kernel(self, egrid, density, sigma, mass, temperature, vdoslux, target_emax=0.0, weight=None):
        egrid = np.ascontiguousarray(egrid, dtype=np.float64)
        density = np.ascontiguousarray(density, dtype=np.float64)
        na, nb = C.c_uint(0), C.c_uint(0)
        pa, pb, ps = _dp(), _dp(), _dp()
        sug = C.c_double(-1.0)
        cb = C.cast(WEIGHT_FCT(weight), C.c_void_p) if weight is not None else None
        self._keep = cb
        self.L.ncrystal_raw_vdos2kernel(_d(egrid), _d(density), egrid.size, density.size, sigma, mass, temperature,
                                        vdoslux, cb, C.byref(na), C.byref(nb), C.byref(pa), C.byref(pb),
                                        C.byref(ps), target_emax, C.byref(sug))
        if not pa or not pb or not ps:
            raise RuntimeError("ncrystal_raw_vdos2kernel failed")
        alpha = self._take(pa, na.value)
        beta = self._take(pb, nb.value)
        sab = self._take(ps, na.value * nb.value)
        return alpha, beta, sab, sug.value

    def gn(self, egrid, density, sigma, mass, temperature, order):
        egrid = np.ascontiguousarray(egrid, dtype=np.float64)
        density = np.ascontiguousarray(density, dtype=np.float64)
        xmin, xmax, n, p = C.c_double(), C.c_double(), C.c_uint(0), _dp()
        self.L.ncrystal_raw_vdos2gn(_d(egrid), _d(density), egrid.size, density.size, sigma, mass, temperature, order,
                                    C.byref(xmin), C.byref(xmax), C.byref(n), C.byref(p))
        if not p:
            raise RuntimeError("ncrystal_raw_vdos2gn failed")
        return xmin.value, xmax.value, self._take(p, n.value)


def have_reference():
    return os.path.exists(REFLIB_PATH)


def reference_api():
    return RawVdosAPI(C.CDLL(REFLIB_PATH))


class HostSimVdos:
    """Same call shapes over the TEST-ONLY host build (tests/hostsim/hostsim_vdos.cpp)."""

    def __init__(self):
        from _libs import HostSim
        L = HostSim.lib()
        L.hostsim_vdos_lasterror.restype = C.c_char_p
        L.hostsim_vdos_expand.argtypes = [_dp, C.c_uint, _dp, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_uint,
                                          C.c_double, _dp, C.POINTER(C.c_int), _dp, C.POINTER(C.c_int), _dp, C.c_int, _dp]
        L.hostsim_vdos_gn.argtypes = [_dp, C.c_uint, _dp, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_int, _dp,
                                      C.c_int, _dp]
        self.L = L

    def kernel(self, egrid, density, sigma, mass, temperature, vdoslux, target_emax=0.0, weight=None):
        assert weight is None
        egrid = np.ascontiguousarray(egrid, dtype=np.float64)
        density = np.ascontiguousarray(density, dtype=np.float64)
        nbmax = 100 * (1 << vdoslux)
        a, b, s, m = np.zeros(nbmax), np.zeros(nbmax), np.zeros(nbmax * nbmax // 2), np.zeros(5)
        ia, ib = C.c_int(), C.c_int()
        rc = self.L.hostsim_vdos_expand(_d(egrid), egrid.size, _d(density), density.size, temperature, mass, sigma,
                                        vdoslux, target_emax, _d(a), C.byref(ia), _d(b), C.byref(ib), _d(s), s.size, _d(m))
        if rc:
            raise RuntimeError("hostsim_vdos_expand: %d %s" % (rc, self.L.hostsim_vdos_lasterror().decode()))
        return a[:ia.value].copy(), b[:ib.value].copy(), s[:ia.value * ib.value].copy(), m[0]

    def gn(self, egrid, density, sigma, mass, temperature, order):
        egrid = np.ascontiguousarray(egrid, dtype=np.float64)
        density = np.ascontiguousarray(density, dtype=np.float64)
        spec, m = np.zeros(1 << 20), np.zeros(4)
        n = self.L.hostsim_vdos_gn(_d(egrid), egrid.size, _d(density), density.size, temperature, mass, sigma, order,
                                   _d(spec), spec.size, _d(m))
        if n < 0:
            raise RuntimeError("hostsim_vdos_gn: %d %s" % (n, self.L.hostsim_vdos_lasterror().decode()))
        return m[0], m[1], spec[:n].copy()


def synthetic_curves():
    """VDOS curves that are not in the reference's data library: an irregular grid (forces regulariseVDOSGrid), a
    two-point grid that needs a slight emax correction, a coarse curve (forces the thickening of G_1's grid)."""
    rng = np.random.Generator(np.random.Philox(key=2024))
    e1 = np.cumsum(0.0005 + 0.0004 * rng.random(60)) + 0.004
    d1 = np.sin(np.linspace(0.05, 3.0, 60)) ** 2 + 0.3 * np.exp(-((e1 - 0.02) / 0.003) ** 2)
    e2 = np.array([0.0123, 0.0456])
    d2 = np.abs(np.sin(np.linspace(0.0, 7.0, 83))) + 0.01
    e3 = np.array([0.01, 0.03])
    d3 = np.array([0.2, 0.7, 1.0, 0.6, 0.4, 0.9, 0.1])
    return {"irregular": (e1, d1), "twopoint": (e2, d2), "coarse": (e3, d3)}


def weight_halve_evens(order):
    return 0.5 if order % 2 == 0 else 1.0


# name -> (curve key, sigma, mass, T, vdoslux, target_emax, weight fct or None, G_n order to pin)
CASES = {
    "Al_lux3": ("Al", None, None, None, 3, 0.0, None, 7),
    "Al_lux0": ("Al", None, None, None, 0, 0.0, None, 2),
    "CH2_H_lux3": ("CH2_H", None, None, None, 3, 0.0, None, 33),
    "Be_lux1_emax": ("Be", None, None, None, 1, 0.7, None, 5),
    "Be_lux2_weights": ("Be", None, None, None, 2, 0.0, weight_halve_evens, 12),
    "irregular_lux2": ("irregular", 5.0, 12.0, 350.0, 2, 0.0, None, 9),
    "twopoint_lux1": ("twopoint", 2.2, 55.8, 150.0, 1, 0.0, None, 3),
    "coarse_lux1": ("coarse", 80.0, 1.008, 293.15, 1, 0.0, None, 6),
}
LIBRARY_CURVES = {"Al": ("Al_sg225.ncmat;temp=293.15K", 0), "CH2_H": ("Polyethylene_CH2.ncmat", 0), "Be": ("Be_sg194.ncmat", 0)}
SUBSAMPLE = 97


def load_golden():
    return np.load(GOLDEN, allow_pickle=False)


def case_inputs(g, name):
    curve, sigma, mass, T, lux, emax, weight, order = CASES[name]
    egrid, density = g["in_%s_egrid" % curve], g["in_%s_density" % curve]
    if sigma is None:
        sigma, mass, T = [float(x) for x in g["in_%s_meta" % curve]]
    return egrid, density, sigma, mass, T, lux, emax, weight, order


def check_against_golden(api, g, name, with_gn=True):
    """Bit-for-bit comparison of one case with the committed reference results."""
    egrid, density, sigma, mass, T, lux, emax, weight, order = case_inputs(g, name)
    alpha, beta, sab, sug = api.kernel(egrid, density, sigma, mass, T, lux, emax, weight)
    assert np.array_equal(alpha, g["out_%s_alpha" % name]), name
    assert np.array_equal(beta, g["out_%s_beta" % name]), name
    assert sab.size == alpha.size * beta.size
    assert np.array_equal(sab[::SUBSAMPLE], g["out_%s_sab_sub" % name]), name
    assert sha(sab) == str(g["out_%s_sab_sha" % name]), name
    assert sug == float(g["out_%s_emax" % name]), name
    if with_gn:
        xmin, xmax, spec = api.gn(egrid, density, sigma, mass, T, order)
        assert (xmin, xmax) == tuple(g["out_%s_gn_range" % name]), name
        assert sha(spec) == str(g["out_%s_gn_sha" % name]), name
