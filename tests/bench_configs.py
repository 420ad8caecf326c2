"""Device-resident throughput of every BASELINE.json config (not a bench.py line: parity-test configs,
measured for the record).  Usage: python tests/bench_configs.py [n] -> JSON lines."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import ncrystal_b200 as nc  # noqa: E402
from __graft_entry__ import CONFIGS  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream(dev)
sp = C.c_void_p(st.cuda_stream)


def timeit(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(reps):
        fn()
    b.record(st)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for key, cfg in CONFIGS.items():
    sc = nc.Scatter(cfg, seed=1)
    L = sc._L
    out = {"config": key, "cfg": cfg, "n": n, "table_MB": sc.tableBytes() / 1e6, "components": sc.components()}
    if sc.isOriented():
        m = min(n, 4_000_000)
        e, (ux, uy, uz) = nc.generateSource(m, directions=True, device=dev)
        xs = torch.empty_like(e)
        eo, ox, oy, oz = [torch.empty_like(e) for _ in range(4)]
        t_xs = timeit(lambda: L.ncb200_crosssection_many_dev(sc._p, e.data_ptr(), ux.data_ptr(), uy.data_ptr(), uz.data_ptr(), m, xs.data_ptr(), sp))
        t_sm = timeit(lambda: L.ncb200_samplescatter_manydir_dev(sc._h, e.data_ptr(), ux.data_ptr(), uy.data_ptr(), uz.data_ptr(), m,
                                                                 eo.data_ptr(), ox.data_ptr(), oy.data_ptr(), oz.data_ptr(), sp))
        out.update(n=m, xs_per_s=m / t_xs * 1e3, samples_per_s=m / t_sm * 1e3, ms_xs=t_xs, ms_sample=t_sm)
    else:
        e = nc.generateSource(n, device=dev)
        xs, eo, mu = [torch.empty_like(e) for _ in range(3)]
        t_xs = timeit(lambda: L.ncb200_crosssection_nonoriented_many_dev(sc._p, e.data_ptr(), n, xs.data_ptr(), sp))
        t_sm = timeit(lambda: L.ncb200_samplescatterisotropic_many_dev(sc._h, e.data_ptr(), n, eo.data_ptr(), mu.data_ptr(), sp))
        t_fu = timeit(lambda: L.ncb200_xs_and_samplescatterisotropic_many_dev(sc._h, e.data_ptr(), n, xs.data_ptr(), eo.data_ptr(), mu.data_ptr(), sp))
        out.update(xs_per_s=n / t_xs * 1e3, samples_per_s=n / t_sm * 1e3, fused_per_s=n / t_fu * 1e3, ms_xs=t_xs, ms_sample=t_sm, ms_fused=t_fu)
    out["device_error_flags"] = sc.checkDeviceErrors(dev)
    print(json.dumps(out), flush=True)
