"""Pipeline tuning helper: device time of the xs / sample launch sequences vs batch size, and e2e time of the
host-pointer calls (run under different NCB200_CHUNK0 / NCB200_CHUNK / NCB200_CHUNK_GROWTH settings)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ncrystal_b200 as nc
from __graft_entry__ import CONFIGS
n = 10_000_000
torch.cuda.set_device(0)
sc = nc.Scatter(CONFIGS["Al"], seed=1)
L = sc._L
e = nc.generateSource(n)
tag = "c0=%s max=%s g=%s" % tuple(os.environ.get(k, "-") for k in ("NCB200_CHUNK0", "NCB200_CHUNK", "NCB200_CHUNK_GROWTH"))
if "--tn" in sys.argv:
    d_xs, d_eo, d_mu = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(3)]
    st = torch.cuda.current_stream()
    sp = C.c_void_p(st.cuda_stream)
    for m in (65536, 131072, 262144, 524288, 1 << 20, 1 << 21, 1 << 22, n):
        for name, f in (("xs", lambda: L.ncb200_crosssection_nonoriented_many_dev(sc._p, e.data_ptr(), m, d_xs.data_ptr(), sp)),
                        ("sample", lambda: L.ncb200_samplescatterisotropic_many_dev(sc._h, e.data_ptr(), m, d_eo.data_ptr(), d_mu.data_ptr(), sp))):
            for _ in range(3): f()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record(st)
            for _ in range(20): f()
            b.record(st); torch.cuda.synchronize()
            print("t(n) %s n=%d: %.4f ms (%.4f ms per 1M)" % (name, m, a.elapsed_time(b) / 20, a.elapsed_time(b) / 20 / m * 1e6))
h_e = torch.empty(n, dtype=torch.float64).pin_memory(); h_e.copy_(e)
h_xs, h_eo, h_mu = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(3)]
dp = C.POINTER(C.c_double)
def xs(): L.ncrystal_crosssection_nonoriented_many(sc._p, C.cast(h_e.data_ptr(), dp), n, 1, C.cast(h_xs.data_ptr(), dp))
def sm(): L.ncrystal_samplescatterisotropic_many(sc._h, C.cast(h_e.data_ptr(), dp), n, 1, C.cast(h_eo.data_ptr(), dp), C.cast(h_mu.data_ptr(), dp))
for f in (xs, sm): f(); f()
res = []
for name, f in (("xs", xs), ("sample", sm)):
    best = 1e9
    for _ in range(7):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); best = min(best, time.perf_counter() - t0)
    res.append(best)
print("%s : xs %.3f ms  sample %.3f ms  total %.3f ms -> %.3e neutrons/s" % (tag, res[0] * 1e3, res[1] * 1e3, sum(res) * 1e3, n / sum(res)))
