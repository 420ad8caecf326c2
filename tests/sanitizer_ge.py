"""compute-sanitizer driver for the oriented (single-crystal) kernels only; see tests/sanitizer_run.py."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ncrystal_b200 as nc
from __graft_entry__ import CONFIGS
from _libs import loguniform_energies, isotropic_directions
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
sc = nc.Scatter(CONFIGS["Ge"], seed=3)
e = loguniform_energies(n, seed=11); d = isotropic_directions(n, seed=12)
print(float(np.sum(sc.crossSection(e, d))), float(np.sum(sc.sampleScatter(e, d)[0])))
res = sc.minimc("sphere;r=0.005", "constant;wl=3.2;z=-0.005;n=20000", "tally=mu")
print("minimc tallied", res["output"]["metadata"]["tallied"]["count"])
