"""VDOS -> S(alpha,beta) sweep over the reference's whole data library (evidence script, not a pytest): every phonon
density of states of every NCMAT file the compiled reference embeds is expanded with ncrystal_raw_vdos2kernel by

   python tests/vdos_sweep.py device [vdoslux]     the product (libncrystal_b200.so, CUDA)
   python tests/vdos_sweep.py host   [vdoslux]     the TEST-ONLY host build of the same code (no GPU needed)
   (optional third argument n: every n-th file of the library only)

and by the live reference; the tables must be bit-identical.  One JSON line per curve, a summary line at the end."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _vdos  # noqa: E402
from _libs import RefDrv, _d  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "host"
lux = int(sys.argv[2]) if len(sys.argv) > 2 else 1
every = int(sys.argv[3]) if len(sys.argv) > 3 else 1      # take every n-th file only
if which == "device":
    from ncrystal_b200 import _lib
    api = _vdos.RawVdosAPI(_lib.lib())
else:
    api = _vdos.HostSimVdos()
ref = _vdos.reference_api()
NC = ref.L
n, strs = C.c_uint(0), C.POINTER(C.c_char_p)()
NC.ncrystal_get_file_list.argtypes = [C.POINTER(C.c_uint), C.POINTER(C.POINTER(C.c_char_p))]
NC.ncrystal_get_file_list(C.byref(n), C.byref(strs))
names = sorted({strs[i].decode() for i in range(0, n.value, 4) if strs[i].decode().endswith(".ncmat")})
L = RefDrv.lib()
L.refdrv_vdos_data.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
ncurves = nbad = nfail = 0
t_dev = t_ref = 0.0
for name in names[::every]:
    for k in range(16):
        meta, dens = np.zeros(5), np.zeros(200000)
        m = L.refdrv_vdos_data(name.encode(), k, _d(meta), _d(dens), dens.size)
        if m <= 0:
            break
        egrid, density = meta[:2].copy(), dens[:m].copy()
        sigma, mass, T = float(meta[4]), float(meta[3]), float(meta[2])
        rec = {"file": name, "vdos": k, "npts": int(m), "mass": mass, "T": T, "vdoslux": lux}
        try:
            t0 = time.perf_counter(); r = ref.kernel(egrid, density, sigma, mass, T, lux); t1 = time.perf_counter()
            g = api.kernel(egrid, density, sigma, mass, T, lux); t2 = time.perf_counter()
            same = all(a.shape == b.shape and np.array_equal(a, b) for a, b in zip(r[:3], g[:3])) and r[3] == g[3]
            rec.update(nalpha=int(r[0].size), nbeta=int(r[1].size), identical=bool(same), ref_s=t1 - t0, s=t2 - t1)
            t_ref += t1 - t0; t_dev += t2 - t1
            nbad += 0 if same else 1
        except Exception as e:  # noqa: BLE001
            rec["error"] = str(e); nfail += 1
        ncurves += 1
        print(json.dumps(rec), flush=True)
print(json.dumps({"summary": which, "vdoslux": lux, "files": len(names), "curves": ncurves, "not_identical": nbad, "errors": nfail,
                  "seconds": t_dev, "reference_seconds": t_ref}))
