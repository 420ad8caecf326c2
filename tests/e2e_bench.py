"""e2e-only timing (host-pointer C-API path with pinned buffers) for pipeline tuning."""
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
h_e = torch.empty(n, dtype=torch.float64).pin_memory(); h_e.copy_(e)
h_xs, h_eo, h_mu = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(3)]
dp = C.POINTER(C.c_double)
def xs(): L.ncrystal_crosssection_nonoriented_many(sc._p, C.cast(h_e.data_ptr(), dp), n, 1, C.cast(h_xs.data_ptr(), dp))
def sm(): L.ncrystal_samplescatterisotropic_many(sc._h, C.cast(h_e.data_ptr(), dp), n, 1, C.cast(h_eo.data_ptr(), dp), C.cast(h_mu.data_ptr(), dp))
for f in (xs, sm): f()
for name, f in (("xs", xs), ("sample", sm)):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): f()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print("%s %s: %.3f ms  -> %.3e /s ; PCIe traffic %.1f GB/s" % (os.environ.get("NCB200_CHUNK", "-"), name, dt * 1e3, n / dt, (16 if name == "xs" else 24) * n / dt / 1e9))
# raw copy bandwidth reference
d = torch.empty(n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): d.copy_(h_e, non_blocking=True)
torch.cuda.synchronize(); print("H2D 80MB: %.1f GB/s" % (5 * 8 * n / (time.perf_counter() - t0) / 1e9))
t0 = time.perf_counter()
for _ in range(5): h_xs.copy_(d, non_blocking=True)
torch.cuda.synchronize(); print("D2H 80MB: %.1f GB/s" % (5 * 8 * n / (time.perf_counter() - t0) / 1e9))
# both directions at once in 8 MB pieces on two streams: what the link gives a 1:2 (in:out) pipeline such as the sampling call
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty(2 * n, dtype=torch.float64, device="cuda")
h_out = torch.empty(2 * n, dtype=torch.float64).pin_memory()
piece = 1 << 20
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    for k in range(0, n, piece):
        with torch.cuda.stream(s1):
            d[k:k + piece].copy_(h_e[k:k + piece], non_blocking=True)
        with torch.cuda.stream(s2):
            h_out[2 * k:2 * (k + piece)].copy_(d2[2 * k:2 * (k + piece)], non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print("80 MB in + 160 MB out concurrently: %.3f ms (%.1f GB/s in total)" % (dt * 1e3, 24 * n / dt / 1e9))
