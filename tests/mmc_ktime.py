"""Per-kernel times of one transport run (ncb200_kernel_timing): python tests/mmc_ktime.py Ge 1e6"""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ncrystal_b200 as nc
from ncrystal_b200 import _lib
from __graft_entry__ import CONFIGS
from _mmc import Scenario
key = sys.argv[1] if len(sys.argv) > 1 else "Ge"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
sc = {"Ge": Scenario("ge", "Ge", ("sphere", {"r": 0.005}), "constant", ("wl", 3.2), n, pos=(0, 0, -0.005)),
      "Al": Scenario("al", "Al", ("sphere", {"r": 0.05}), "constant", ("wl", 1.8), n, pos=(0, 0, -0.05)),
      "H2O": Scenario("h2o", "H2O", ("sphere", {"r": 0.002}), "circular", ("wl", 1.8), n, pos=(0, 0, -0.002), radius=0.002)}[key]
s = nc.Scatter(CONFIGS[key], seed=1)
if len(sys.argv) > 3:
    _lib.lib().ncb200_set_mmc_tail_mode(int(sys.argv[3]))
s.minimc(sc.geomcfg, sc.srccfg(), sc.enginecfg())
L = _lib.lib()
L.ncb200_kernel_timing(1)
t0 = time.perf_counter()
res = s.minimc(sc.geomcfg, sc.srccfg(), sc.enginecfg())
dt = time.perf_counter() - t0
buf = C.create_string_buffer(1 << 16)
L.ncb200_kernel_timing_report(buf, len(buf))
L.ncb200_kernel_timing(0)
kt = json.loads(buf.value.decode())
tot = {k: v["launches"] * v["ms_avg"] for k, v in kt.items()}
print(json.dumps({"tail_mode": sys.argv[3] if len(sys.argv) > 3 else "default", "material": key, "n": n, "wall_ms_with_timing": dt * 1e3, "device_ms": res["b200"]["device_ms"], "steps": res["b200"]["steps"],
                  "launches": res["b200"]["kernel_launches"],
                  "kernel_total_ms": dict(sorted(tot.items(), key=lambda kv: -kv[1])), "kernel_launches": {k: v["launches"] for k, v in kt.items()}}))
