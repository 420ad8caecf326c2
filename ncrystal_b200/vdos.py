"""VDOS -> S(alpha,beta) on the device: host-side mirror of the reference's Python helpers `extractGn` / `extractKnl`
(ref: ncrystal_python/src/NCrystal/vdos.py:90-157, bound like _chooks.py:305-362) over the library's
ncrystal_raw_vdos2gn / ncrystal_raw_vdos2kernel (same C symbols and signatures as the reference's C-API,
include/NCrystal/cinterface/ncrystal.h:885-925).  The phonon-order convolutions and the sum over orders run in CUDA
kernels (csrc/ncb_vdos_dev.cuh); the results equal the reference's bit for bit.  No CPU fallback."""
import ctypes as C

import numpy as np

from . import _lib
from .core import _check_error

_dblp = C.POINTER(C.c_double)
_ORDERWEIGHTFCT = C.CFUNCTYPE(C.c_double, C.c_uint)


def _take(L, p, n):
    a = np.ctypeslib.as_array(p, (n,)).copy() if n else np.zeros(0)
    L.ncrystal_dealloc_doubleptr(p)
    return a


def _curve(vdos):
    egrid, density = vdos
    egrid = np.ascontiguousarray(egrid, dtype=np.float64)
    density = np.ascontiguousarray(density, dtype=np.float64)
    return egrid, density


def extractGn(vdos, n, mass_amu, temperature, scatxs=1.0, expand_egrid=True):
    """Sjolander's G_n of order n for the curve vdos = (egrid, density); egrid has two points or one per density value."""
    assert 1 <= n <= 99999
    L = _lib.lib()
    egrid, density = _curve(vdos)
    xmin, xmax, ny, y = C.c_double(), C.c_double(), C.c_uint(0), _dblp()
    L.ncrystal_raw_vdos2gn(egrid.ctypes.data_as(_dblp), density.ctypes.data_as(_dblp), egrid.size, density.size,
                           float(scatxs), float(mass_amu), float(temperature), int(n),
                           C.byref(xmin), C.byref(xmax), C.byref(ny), C.byref(y))
    _check_error()
    gn = _take(L, y, ny.value)
    if not expand_egrid:
        return (xmin.value, xmax.value), gn
    return np.linspace(xmin.value, xmax.value, len(gn)), gn


def extractKnl(vdos, mass_amu, temperature, vdoslux=3, scatxs=1.0, order_weight_fct=None, target_emax=None):
    """Expand the curve vdos = (egrid, density) to a scattering kernel; returns the reference's dictionary (alpha, beta,
    sab, mass_amu, temperature, scatxs, suggested_emax -- None when an order weight function is given)."""
    L = _lib.lib()
    egrid, density = _curve(vdos)
    cb = None
    if order_weight_fct:
        cb = _ORDERWEIGHTFCT(lambda order: float(order_weight_fct(int(order))))
    emax = float(target_emax) if target_emax and target_emax > 0.0 else 0.0
    na, nb, sug = C.c_uint(0), C.c_uint(0), C.c_double(0.0)
    pa, pb, ps = _dblp(), _dblp(), _dblp()
    L.ncrystal_raw_vdos2kernel(egrid.ctypes.data_as(_dblp), density.ctypes.data_as(_dblp), egrid.size, density.size,
                               float(scatxs), float(mass_amu), float(temperature), int(vdoslux),
                               C.cast(cb, C.c_void_p) if cb else None, C.byref(na), C.byref(nb),
                               C.byref(pa), C.byref(pb), C.byref(ps), emax, C.byref(sug))
    _check_error()
    return dict(alpha=_take(L, pa, na.value), beta=_take(L, pb, nb.value), sab=_take(L, ps, na.value * nb.value),
                mass_amu=mass_amu, temperature=temperature, scatxs=scatxs, suggested_emax=float(sug.value) or None)
