"""Throughput of the device-resident transport step next to the reference's MiniMC on the host cores.
usage: python tests/mmc_bench.py [n_device] [n_reference]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ncrystal_b200 as nc
from __graft_entry__ import CONFIGS
from _mmc import Scenario, reference_minimc

n_dev = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
n_ref = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
cases = [("Al", Scenario("al", "Al", ("sphere", {"r": 0.05}), "constant", ("wl", 1.8), n_dev, pos=(0, 0, -0.05))),
         ("H2O", Scenario("h2o", "H2O", ("sphere", {"r": 0.002}), "circular", ("wl", 1.8), n_dev, pos=(0, 0, -0.002), radius=0.002)),
         ("Ge", Scenario("ge", "Ge", ("sphere", {"r": 0.005}), "constant", ("wl", 3.2), n_dev // 10, pos=(0, 0, -0.005)))]
for key, sc in cases:
    s = nc.Scatter(CONFIGS[key], seed=1)
    s.minimc(sc.geomcfg, sc.srccfg(), sc.enginecfg())   # warm-up (allocations, module load)
    t0 = time.perf_counter()
    res = s.minimc(sc.geomcfg, sc.srccfg(), sc.enginecfg())
    dt = time.perf_counter() - t0
    md = res["output"]["metadata"]
    line = dict(material=key, n=sc.n, wall_s=dt, device_ms=res["b200"]["device_ms"], steps=res["b200"]["steps"],
                launches=res["b200"]["kernel_launches"], histories_per_s=sc.n / dt,
                records_per_s=md["tallied"]["count"] / dt, tallied_weight_frac=md["tallied"]["weight"] / sc.n)
    try:
        nr = n_ref if key != "Ge" else n_ref // 10
        t0 = time.perf_counter()
        js = json.loads(reference_minimc(CONFIGS[key], sc, nthreads=os.cpu_count() or 2, n=nr))
        dtr = time.perf_counter() - t0
        line.update(ref_n=nr, ref_wall_s=dtr, ref_histories_per_s=nr / dtr, ref_threads=os.cpu_count(),
                    ref_tallied_weight_frac=js["output"]["metadata"]["tallied"]["weight"] / nr,
                    speedup=(sc.n / dt) / (nr / dtr))
    except Exception as e:  # noqa: BLE001
        line["ref_error"] = str(e)
    print(json.dumps(line))
