import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ncrystal_b200 as nc
from __graft_entry__ import CONFIGS
key = sys.argv[1] if len(sys.argv) > 1 else "Al"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
wl = float(sys.argv[4]) if len(sys.argv) > 4 else 1.8
s = nc.Scatter(CONFIGS[key], seed=1)
r = {"Al": 0.05, "H2O": 0.002, "Ge": 0.005}[key]
for k in range(reps):
    t0 = time.perf_counter()
    res = s.minimc("sphere;r=%g" % r, "constant;wl=%g;z=%g;n=%d" % (wl, -r, n), "tally=theta,mu")
    print("wall %.1f ms" % (1e3 * (time.perf_counter() - t0)), res["b200"], res["output"]["metadata"]["tallied"])
