"""Probe (GPU box): cost of cudaHostRegister/Unregister for an 80 MB malloc'd array, multi-threaded memcpy bandwidth
pageable->pinned, and pageable cudaMemcpy -- inputs for the design of the pageable-buffer path."""
import ctypes as C, time, threading, numpy as np, torch
rt = torch.cuda.cudart()
torch.cuda.init(); torch.zeros(1, device="cuda")
n = 10_000_000
a = np.random.rand(n)               # pageable, touched
lib = C.CDLL("libcudart.so.12") if False else None
for flags in (0, 8):   # 8 = cudaHostRegisterReadOnly
    ts = []
    for rep in range(4):
        t0 = time.perf_counter(); r = rt.cudaHostRegister(a.ctypes.data, a.nbytes, flags); t1 = time.perf_counter()
        rt.cudaHostUnregister(a.ctypes.data); t2 = time.perf_counter()
        ts.append((t1 - t0, t2 - t1, int(r)))
    print("cudaHostRegister flags=%d 80MB: " % flags, ["reg %.2f ms unreg %.2f ms rc %d" % (x * 1e3, y * 1e3, z) for x, y, z in ts])
# chunked registration (8 MB pieces)
t0 = time.perf_counter()
for o in range(0, a.nbytes, 8 << 20):
    rt.cudaHostRegister(a.ctypes.data + o, min(8 << 20, a.nbytes - o), 0)
t1 = time.perf_counter()
for o in range(0, a.nbytes, 8 << 20):
    rt.cudaHostUnregister(a.ctypes.data + o)
t2 = time.perf_counter()
print("10 x 8MB pieces: reg %.2f ms unreg %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
pin = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
for nt in (1, 2, 4, 8, 12, 16):
    parts = np.array_split(np.arange(n), nt)
    def work(p):
        pin[p[0]:p[-1] + 1] = a[p[0]:p[-1] + 1]
    best = 1e9
    for rep in range(5):
        th = [threading.Thread(target=work, args=(p,)) for p in parts]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        best = min(best, time.perf_counter() - t0)
    print("memcpy pageable->pinned %2d threads: %.1f GB/s" % (nt, a.nbytes / best / 1e9))
d = torch.empty(n, dtype=torch.float64, device="cuda")
ta = torch.from_numpy(a)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(ta); torch.cuda.synchronize(); t1 = time.perf_counter()
print("pageable H2D 80MB: %.1f GB/s" % (a.nbytes / (t1 - t0) / 1e9))
tp = torch.from_numpy(pin)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(tp, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print("pinned H2D 80MB: %.1f GB/s" % (a.nbytes / (t1 - t0) / 1e9))
import os
print("cores", len(os.sched_getaffinity(0)))
