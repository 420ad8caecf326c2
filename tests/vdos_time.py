"""Timing of the VDOS -> S(alpha,beta) expansion: ncrystal_raw_vdos2kernel of the product (device) beside the same call
of the unmodified reference (host, its own worker threads) on the GPU box.  One JSON line per case."""
import json
import sys
import time

import numpy as np

import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _vdos  # noqa: E402

g = _vdos.load_golden()
from ncrystal_b200 import _lib  # noqa: E402
prod = _vdos.RawVdosAPI(_lib.lib())
ref = _vdos.reference_api() if _vdos.have_reference() else None
L = _lib.lib()
cases = [("Al", 293.15, 3), ("CH2_H", 293.15, 3), ("Be", 293.15, 4), ("Al", 900.0, 5), ("CH2_H", 77.0, 5)]
prod.kernel(g["in_Be_egrid"], g["in_Be_density"], 7.6, 9.0, 300.0, 0)   # context, module load
for curve, T, lux in cases:
    egrid, density = g["in_%s_egrid" % curve], g["in_%s_density" % curve]
    sigma, mass, _ = [float(x) for x in g["in_%s_meta" % curve]]
    best = 1e9
    for _ in range(3):
        l0 = L.ncb200_kernel_launch_count() if hasattr(L, "ncb200_launch_count") else 0
        t = time.perf_counter()
        a, b, s, e = prod.kernel(egrid, density, sigma, mass, T, lux)
        best = min(best, time.perf_counter() - t)
    rec = {"curve": curve, "temperature": T, "vdoslux": lux, "nalpha": int(a.size), "nbeta": int(b.size), "device_call_s": best}
    if ref is not None:
        rb = 1e9
        for _ in range(2):
            t = time.perf_counter()
            ra, rbeta, rs, re_ = ref.kernel(egrid, density, sigma, mass, T, lux)
            rb = min(rb, time.perf_counter() - t)
        rec.update(reference_call_s=rb, speedup=rb / best, bit_identical=bool(np.array_equal(s, rs) and np.array_equal(a, ra) and np.array_equal(b, rbeta)))
    print(json.dumps(rec), flush=True)
