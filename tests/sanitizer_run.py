"""Small driver for compute-sanitizer (evidence, not a pytest): every launch sequence of the host API on three
materials, with the staged free-gas kernels forced, plus the virtual-API style single-neutron calls.
usage: compute-sanitizer --tool memcheck python tests/sanitizer_run.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ncrystal_b200 as nc
from __graft_entry__ import CONFIGS, EXTRA_CONFIGS
from _libs import loguniform_energies, isotropic_directions

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
for key, cfg in list(CONFIGS.items()) + [("gas", EXTRA_CONFIGS["gas"])]:
    if key in ("CH2", "YAG"):
        continue
    sc = nc.Scatter(cfg, seed=3)
    e = loguniform_energies(n, seed=11)
    for nmin in (1, 1 << 40):
        sc._L.ncb200_set_fg_staged_min(nmin)
        if sc.isOriented():
            d = isotropic_directions(n, seed=12)
            xs = sc.crossSection(e, d)
            eo, _ = sc.sampleScatter(e, d)
        else:
            xs = sc.crossSectionIsotropic(e)
            eo, _ = sc.sampleScatterIsotropic(e)
        print(key, "staged" if nmin == 1 else "single", float(np.sum(xs)), float(np.sum(eo)))
    sc._L.ncb200_set_fg_staged_min(4000000)
    if key == "Al":
        res = sc.minimc("sphere;r=0.01", "constant;ekin=0.0253;z=-0.01;n=20000", "tally=mu")
        print("minimc tallied", res["output"]["metadata"]["tallied"]["count"])
