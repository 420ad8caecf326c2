#!/usr/bin/env python
"""Generates tests/golden/vdos_reference.npz: results of the REFERENCE's VDOS -> S(alpha,beta) expansion (NCrystal 4.4.2
in oracle/_ref, its own C-API ncrystal_raw_vdos2kernel / ncrystal_raw_vdos2gn) for the cases of tests/_vdos.py.  Inputs
are stored in full, of each S(alpha,beta) table the grids, every 97th value and the SHA-256 of its bytes (the product
and the host build must reproduce the tables bit for bit).  Run in the build container only.

    python tests/golden/make_golden_vdos.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _vdos  # noqa: E402
from _libs import RefDrv, _d  # noqa: E402

api = _vdos.reference_api()
L = RefDrv.lib()
L.refdrv_vdos_data.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
out = {}
for key, (cfg, k) in _vdos.LIBRARY_CURVES.items():
    meta, dens = np.zeros(5), np.zeros(100000)
    n = L.refdrv_vdos_data(cfg.encode(), k, _d(meta), _d(dens), dens.size)
    assert n > 0, (cfg, n)
    out["in_%s_egrid" % key] = meta[:2].copy()
    out["in_%s_density" % key] = dens[:n].copy()
    out["in_%s_meta" % key] = np.array([meta[4], meta[3], meta[2]])   # bound xs, mass, temperature
for key, (e, d) in _vdos.synthetic_curves().items():
    out["in_%s_egrid" % key] = e
    out["in_%s_density" % key] = d
for name in _vdos.CASES:
    egrid, density, sigma, mass, T, lux, emax, weight, order = _vdos.case_inputs(out, name)
    alpha, beta, sab, sug = api.kernel(egrid, density, sigma, mass, T, lux, emax, weight)
    xmin, xmax, spec = api.gn(egrid, density, sigma, mass, T, order)
    out["out_%s_alpha" % name] = alpha
    out["out_%s_beta" % name] = beta
    out["out_%s_sab_sub" % name] = sab[::_vdos.SUBSAMPLE].copy()
    out["out_%s_sab_sha" % name] = np.array(_vdos.sha(sab))
    out["out_%s_emax" % name] = np.array(sug)
    out["out_%s_gn_range" % name] = np.array([xmin, xmax])
    out["out_%s_gn_sha" % name] = np.array(_vdos.sha(spec))
    out["out_%s_gn_sub" % name] = spec[::7].copy()
    print(name, alpha.size, beta.size, sug, spec.size, "nonzero", int((sab > 0).sum()))
np.savez_compressed(_vdos.GOLDEN, **out)
print(os.path.getsize(_vdos.GOLDEN), "bytes")
