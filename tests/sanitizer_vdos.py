"""Driver for compute-sanitizer over the round-2 additions (evidence, not a pytest): the VDOS -> S(alpha,beta) kernels
(raw expansion of two curves incl. one whose low orders take the global-memory FFT stages, G_n, a material with VDOS
leaves), the layered-crystal kernels, and a Ge transport run that ends in the one-launch tail kernel.
usage: compute-sanitizer --tool memcheck python tests/sanitizer_vdos.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ncrystal_b200 as nc
from ncrystal_b200 import vdos
import _vdos
from __graft_entry__ import CONFIGS, EXTRA_CONFIGS
from _libs import loguniform_energies, isotropic_directions

g = _vdos.load_golden()
for curve, lux in (("Be", 1), ("Al", 2), ("coarse", 1)):
    egrid, density = g["in_%s_egrid" % curve], g["in_%s_density" % curve]
    sigma, mass, T = (float(x) for x in g["in_%s_meta" % curve]) if curve != "coarse" else (80.0, 1.008, 293.15)
    k = vdos.extractKnl((egrid, density), mass, T, vdoslux=lux, scatxs=sigma)
    (a, b), gn = vdos.extractGn((egrid, density), 9, mass, T, scatxs=sigma, expand_egrid=False)
    print(curve, k["sab"].shape, float(k["sab"].sum()), gn.size)
path = os.path.join(ROOT, "ncrystal_b200", "data", "solid__V_6.1gcm3_TDebye390K.vdos.ncb")
if os.path.exists(path):
    sc = nc.Scatter.fromBlob(open(path, "rb").read(), seed=2)
    e = loguniform_energies(20000, seed=3)
    print("V (VDOS leaf)", float(sc.crossSectionIsotropic(e).sum()), float(sc.sampleScatterIsotropic(e)[0].sum()))
if len(sys.argv) > 1 and sys.argv[1] == "all":
    sc = nc.Scatter(EXTRA_CONFIGS["PG"], seed=4)
    n = 4000
    e, d = loguniform_energies(n, seed=5, lo=1e-3, hi=0.1), isotropic_directions(n, seed=6)
    print("PG", float(sc.crossSection(e, d).sum()), float(sc.sampleScatter(e, d)[0].sum()))
    sc = nc.Scatter(CONFIGS["Ge"], seed=1)
    res = sc.minimc("sphere;r=0.005", "constant;wl=3.2;z=-0.005;n=30000", "tally=mu,theta")
    print("Ge minimc", res["output"]["metadata"]["tallied"]["count"], res["b200"]["steps"])
