"""Device-resident throughput of the layered-crystal (LCBragg) path next to the reference on the host cores.
    python tests/lc_time.py [n]        -> one JSON line"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import ncrystal_b200 as nc
from __graft_entry__ import EXTRA_CONFIGS
m = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
cfg = EXTRA_CONFIGS["PG"]
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream(dev); sp = C.c_void_p(st.cuda_stream)
sc = nc.Scatter(cfg, seed=1); L = sc._L
e, (ux, uy, uz) = nc.generateSource(m, directions=True, device=dev)
xs = torch.empty_like(e); eo, ox, oy, oz = [torch.empty_like(e) for _ in range(4)]
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(reps): fn()
    b.record(st); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
txs = t(lambda: L.ncb200_crosssection_many_dev(sc._p, e.data_ptr(), ux.data_ptr(), uy.data_ptr(), uz.data_ptr(), m, xs.data_ptr(), sp))
tsm = t(lambda: L.ncb200_samplescatter_manydir_dev(sc._h, e.data_ptr(), ux.data_ptr(), uy.data_ptr(), uz.data_ptr(), m, eo.data_ptr(), ox.data_ptr(), oy.data_ptr(), oz.data_ptr(), sp))
L.ncb200_kernel_timing(1)
L.ncb200_crosssection_many_dev(sc._p, e.data_ptr(), ux.data_ptr(), uy.data_ptr(), uz.data_ptr(), m, xs.data_ptr(), sp)
L.ncb200_samplescatter_manydir_dev(sc._h, e.data_ptr(), ux.data_ptr(), uy.data_ptr(), uz.data_ptr(), m, eo.data_ptr(), ox.data_ptr(), oy.data_ptr(), oz.data_ptr(), sp)
buf = C.create_string_buffer(8192); kt = {}
if L.ncb200_kernel_timing_report(buf, 8192) > 0: kt = json.loads(buf.value.decode())
L.ncb200_kernel_timing(0)
out = {"config": cfg, "n": m, "xs_ms": txs, "xs_per_s": m / txs * 1e3, "sample_ms": tsm, "samples_per_s": m / tsm * 1e3,
       "xs_checksum": float(xs.sum()), "flags": sc.checkDeviceErrors(dev), "kernel_ms": {k: round(v["ms_avg"], 3) for k, v in kt.items()}}
try:
    from _libs import RefDrv, have_refdrv
    if have_refdrv():
        nt = len(os.sched_getaffinity(0)); nc_ = 40000
        Lr = RefDrv.lib(); dp = C.POINTER(C.c_double)
        h = [a[:nc_].cpu().numpy().copy() for a in (e, ux, uy, uz)]; o = [np.empty(nc_) for _ in range(4)]
        p = lambda a: a.ctypes.data_as(dp)
        null = C.cast(None, dp)
        t_xs = Lr.refdrv_bench_capi(cfg.encode(), 2, nt, 1, p(h[0]), p(h[1]), p(h[2]), p(h[3]), nc_, p(o[0]), null, null, null)
        t_sm = Lr.refdrv_bench_capi(cfg.encode(), 3, nt, 1, p(h[0]), p(h[1]), p(h[2]), p(h[3]), nc_, p(o[0]), p(o[1]), p(o[2]), p(o[3]))
        out["cpu_reference"] = {"threads": nt, "n": nc_, "xs_per_s": nc_ / t_xs, "samples_per_s": nc_ / t_sm}
        out["speedup_xs"] = out["xs_per_s"] / (nc_ / t_xs); out["speedup_sample"] = out["samples_per_s"] / (nc_ / t_sm)
except Exception as ex:  # noqa: BLE001
    out["cpu_reference"] = {"error": str(ex)}
print(json.dumps(out))
