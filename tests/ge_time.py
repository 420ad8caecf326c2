import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ncrystal_b200 as nc
from __graft_entry__ import CONFIGS
m = 4_000_000
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream(dev); sp = C.c_void_p(st.cuda_stream)
sc = nc.Scatter(CONFIGS["Ge"], seed=1); L = sc._L
e, (ux, uy, uz) = nc.generateSource(m, directions=True, device=dev)
xs = torch.empty_like(e); eo, ox, oy, oz = [torch.empty_like(e) for _ in range(4)]
def t(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(reps): fn()
    b.record(st); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
txs = t(lambda: L.ncb200_crosssection_many_dev(sc._p, e.data_ptr(), ux.data_ptr(), uy.data_ptr(), uz.data_ptr(), m, xs.data_ptr(), sp))
tsm = t(lambda: L.ncb200_samplescatter_manydir_dev(sc._h, e.data_ptr(), ux.data_ptr(), uy.data_ptr(), uz.data_ptr(), m, eo.data_ptr(), ox.data_ptr(), oy.data_ptr(), oz.data_ptr(), sp))
print("Ge 4M: xs %.2f ms (%.3e /s)  sample %.2f ms (%.3e /s)  xs checksum %.12e" % (txs, m / txs * 1e3, tsm, m / tsm * 1e3, float(xs.sum())))
import json
buf = C.create_string_buffer(1 << 16)
for name, fn in (("xs", lambda: L.ncb200_crosssection_many_dev(sc._p, e.data_ptr(), ux.data_ptr(), uy.data_ptr(), uz.data_ptr(), m, xs.data_ptr(), sp)),
                 ("sample", lambda: L.ncb200_samplescatter_manydir_dev(sc._h, e.data_ptr(), ux.data_ptr(), uy.data_ptr(), uz.data_ptr(), m, eo.data_ptr(), ox.data_ptr(), oy.data_ptr(), oz.data_ptr(), sp))):
    L.ncb200_kernel_timing(1)
    fn(); torch.cuda.synchronize()
    L.ncb200_kernel_timing_report(buf, len(buf))
    L.ncb200_kernel_timing(0)
    kt = json.loads(buf.value.decode())
    print(name, {k: round(v["launches"] * v["ms_avg"], 3) for k, v in kt.items()})
