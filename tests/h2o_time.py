import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ncrystal_b200 as nc
from __graft_entry__ import CONFIGS
key = sys.argv[1] if len(sys.argv) > 1 else "H2O"
n = 10_000_000
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream(dev); sp = C.c_void_p(st.cuda_stream)
sc = nc.Scatter(CONFIGS[key], seed=1); L = sc._L
e = nc.generateSource(n, device=dev)
eo, mu = torch.empty_like(e), torch.empty_like(e)
def t(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(reps): fn()
    b.record(st); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
L.ncb200_kernel_timing(1)
ts = t(lambda: L.ncb200_samplescatterisotropic_many_dev(sc._h, e.data_ptr(), n, eo.data_ptr(), mu.data_ptr(), sp))
buf = C.create_string_buffer(4096); L.ncb200_kernel_timing_report(buf, 4096)
print(key, os.environ.get("NCB200_SORT", "-"), "sample %.3f ms" % ts, buf.value.decode())
