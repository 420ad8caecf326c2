#!/usr/bin/env python
"""Generates tests/golden/abs_reference.npz: absorption cross sections of the REFERENCE (NCrystal 4.4.2 in
oracle/_ref, C-API ncrystal_create_absorption + ncrystal_crosssection_nonoriented_many) for the benchmark
materials.  Run in the build container only.

    python tests/golden/make_golden_abs.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from __graft_entry__ import CONFIGS  # noqa: E402


class H(C.Structure):
    _fields_ = [("internal", C.c_void_p)]


L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "lib", "libNCrystal.so"))
L.ncrystal_create_absorption.restype = H
L.ncrystal_create_absorption.argtypes = [C.c_char_p]
L.ncrystal_cast_abs2proc.restype = H
L.ncrystal_cast_abs2proc.argtypes = [H]
dp = C.POINTER(C.c_double)
L.ncrystal_crosssection_nonoriented_many.argtypes = [H, dp, C.c_ulong, C.c_ulong, dp]
rng = np.random.default_rng(7)
ekin = np.concatenate([[0.0, 5e-324, 1e-300, 1e-12, 1e-5, 0.0253, 1.0, 10.0, 1e6, 1e300],
                       10.0 ** rng.uniform(-7, 3, 500)])
out = {"ekin": ekin}
for key, cfg in CONFIGS.items():
    a = L.ncrystal_create_absorption(cfg.encode())
    p = L.ncrystal_cast_abs2proc(a)
    xs = np.empty_like(ekin)
    L.ncrystal_crosssection_nonoriented_many(p, ekin.ctypes.data_as(dp), len(ekin), 1, xs.ctypes.data_as(dp))
    out[key] = xs
    print(key, xs[5])
np.savez_compressed(os.path.join(HERE, "abs_reference.npz"), **out)
