#!/usr/bin/env python
"""bench.py -- throughput of the hot path on B200: neutrons/s through
crossSection + sampleScatter on synthetic log-uniform 1e-5..10 eV isotropic batches
(BASELINE.json: metric / configs[0], Al_sg225 at 293.15 K, 1e7 neutrons per GPU per step).

    python bench.py --gpus N --steps K --warmup W        (N>1: under torch.distributed.run)
    python bench.py --impl reference ...                 (reference CPU path, oracle/_ref)

One "step" = one pass of the hot path over one batch: batched cross sections, batched scatter
sampling and the mu tally histogram for 1e7 neutrons per GPU.  `value` = neutrons/s of the whole
job with inputs resident in HBM; `e2e` = the same through the reference-facing host-pointer
C entry points (ncrystal_crosssection_nonoriented_many + ncrystal_samplescatterisotropic_many)
with pinned host buffers, copies inside the timed region.  Ranks shard the global neutron index
range (weak scaling, no data-path collective); NCCL only reduces the tally histogram.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CFG_KEY = "Al"
N_PER_GPU = 10_000_000
SEED = 12345
BYTES_XS, BYTES_SAMPLE, BYTES_TALLY = 16, 24, 8   # algorithmic HBM bytes per neutron
NBINS = 200


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference(cfg, n, nthreads, nrep, seed=SEED):
    """The reference's own C-API *_many calls on the host cores (oracle/_ref), one cloned handle per thread."""
    import numpy as np
    from _libs import RefDrv, loguniform_energies
    L = RefDrv.lib()
    ekin = loguniform_energies(n, seed=seed)
    o = [np.empty(n) for _ in range(2)]
    dp = C.POINTER(C.c_double)
    null = C.cast(None, dp)
    t_xs = L.refdrv_bench_capi(cfg.encode(), 0, nthreads, nrep, ekin.ctypes.data_as(dp), null, null, null, n,
                               o[0].ctypes.data_as(dp), null, null, null)
    t_sm = L.refdrv_bench_capi(cfg.encode(), 1, nthreads, nrep, ekin.ctypes.data_as(dp), null, null, null, n,
                               o[0].ctypes.data_as(dp), o[1].ctypes.data_as(dp), null, null)
    return t_xs, t_sm


def cpu_port_baseline(cfg, n=4_000_000):
    """Fallback when the compiled reference is absent: the plain-C oracle port on all host threads, bounded sample."""
    from _libs import loguniform_energies
    try:
        from oracle_check import oracle_for
        o = oracle_for(cfg, prefer="port")
        nt = host_threads()
        ekin = loguniform_energies(n, seed=SEED)
        t_xs = o.bench(0, nt, ekin)
        t_sm = o.bench(1, nt, ekin)
        return {"value": n / (t_xs + t_sm), "unit": "neutrons/s", "cores": nt, "kind": "port",
                "sample": "%d neutrons, xs + sample, oracle/ C port" % n, "xs_per_s": n / t_xs, "samples_per_s": n / t_sm}
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": "neutrons/s", "cores": 0, "kind": "port", "sample": "oracle unavailable: %s" % e}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from _libs import have_refdrv
    nthreads = host_threads()
    n = N_PER_GPU
    if not have_refdrv():
        # the compiled reference did not travel: time the oracle port instead (kind "port")
        for _ in range(max(args.warmup, 1)):
            cb = cpu_port_baseline(cfg)
        vals = [cpu_port_baseline(cfg) for _ in range(args.steps)]
        if vals[0]["value"] is None:
            print(json.dumps({"impl": "reference", "unavailable": vals[0]["sample"]}))
            return
        val = sum(v["value"] for v in vals) / len(vals)
        cb = dict(vals[0], value=val)
        print(json.dumps({
            "impl": "reference", "metric": "neutrons/sec (xs eval + sampleScatter)", "value": val, "unit": "neutrons/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * 4_000_000 / val,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Al_sg225.ncmat;temp=293.15K powder, isotropic crossSection + sampleScatter, "
                                   "log-uniform 1e-5..10 eV", "neutrons_per_step": 4_000_000},
            "cpu_baseline": cb,
            "e2e": {"value": val, "unit": "neutrons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    # each step: a bounded sample of the 1e7-neutron workload through ncrystal_crosssection_nonoriented_many +
    # ncrystal_samplescatterisotropic_many on all host cores; the sample is sized from a probe so that the
    # whole --steps run stays within ~2 minutes (full 1e7 when that fits).
    a, b = cpu_reference(cfg, 1_000_000, nthreads, 1)
    rate = 1_000_000 / (a + b)
    n = int(min(N_PER_GPU, max(200_000, 120.0 * rate / max(args.steps + args.warmup, 1))))
    for _ in range(max(args.warmup, 1)):
        cpu_reference(cfg, n, nthreads, 1)
    t0 = time.perf_counter()
    txs = tsm = 0.0
    for _ in range(args.steps):
        a, b = cpu_reference(cfg, n, nthreads, 1)
        txs += a; tsm += b
    wall = time.perf_counter() - t0
    t_step = (txs + tsm) / args.steps
    val = n / t_step
    print(json.dumps({
        "impl": "reference", "metric": "neutrons/sec (xs eval + sampleScatter)", "value": val, "unit": "neutrons/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Al_sg225.ncmat;temp=293.15K powder, isotropic crossSection + sampleScatter, "
                               "log-uniform 1e-5..10 eV (BASELINE.json configs[0])", "neutrons_per_step": n,
                   "xs_per_s": n * args.steps / txs, "samples_per_s": n * args.steps / tsm,
                   "note": "reference NCrystal 4.4.2 C-API *_many on host cores; wall %.1fs incl. handle setup" % wall},
        "cpu_baseline": {"value": val, "unit": "neutrons/s", "cores": nthreads, "kind": "reference",
                         "sample": "%d neutrons per step (of the 1e7-neutron workload)" % n},
        "e2e": {"value": val, "unit": "neutrons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=N_PER_GPU, help="neutrons per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    from __graft_entry__ import CONFIGS
    cfg = CONFIGS[CFG_KEY]
    if args.impl == "reference":
        run_reference(args, cfg)
        return
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    import ncrystal_b200 as nc
    from ncrystal_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    L = _lib.lib()
    n = args.n
    sc = nc.Scatter(cfg, seed=SEED)
    stream = torch.cuda.current_stream(dev)
    sp = C.c_void_p(stream.cuda_stream)

    # inputs: 3 rotating device buffers (3 x 80 MB) so that a step's input is never L2-resident from
    # the previous step; per-step working set (in+out) = 320 MB > 126 MB L2.
    NBUF = 3
    from ncrystal_b200.sharding import shard_range, merge_tallies
    first = shard_range(world * n, rank, world)[0]   # weak scaling: n neutrons per GPU, contiguous global index ranges
    d_e = [nc.generateSource(n, seed=SEED + b, first_index=first, device=dev) for b in range(NBUF)]
    d_xs = torch.empty(n, dtype=torch.float64, device=dev)
    d_eo = torch.empty(n, dtype=torch.float64, device=dev)
    d_mu = torch.empty(n, dtype=torch.float64, device=dev)
    d_hist = torch.zeros(NBINS + 2, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    step_counter = [0]

    def step(events=None):
        # One pass of the hot path over one batch: total cross section AND sampled scattering for every
        # neutron (the fused entry point, cf. the reference's evalXSAndSampleScatterIsotropic batch ABI,
        # NCABIUtils.hh:78-100), then the mu tally.
        k = step_counter[0]
        step_counter[0] += 1
        e = d_e[k % NBUF]
        sc.setRNGStream(SEED, 0, k * world * n + first)
        if events is not None:
            events[0].record(stream)
        L.ncb200_xs_and_samplescatterisotropic_many_dev(sc._h, e.data_ptr(), n, d_xs.data_ptr(), d_eo.data_ptr(),
                                                        d_mu.data_ptr(), sp)
        if events is not None:
            events[1].record(stream)
        L.ncb200_tally_hist_dev(d_mu.data_ptr(), None, n, -1.0, 1.0, NBINS, d_hist.data_ptr(), None, sp)
        if events is not None:
            events[2].record(stream)

    def timed_loop(fn, reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn(0)
        torch.cuda.synchronize()
        a.record(stream)
        for r in range(reps):
            fn(r)
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    d_hist.zero_()
    barrier()

    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    launches0 = nc.kernelLaunchCount()
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin.record(stream)
    for k in range(args.steps):
        step(ev[k])
    merge_tallies(d_hist)                 # the only collective: tally merge (NCCL all-reduce over NVLink)
    t_end.record(stream)
    barrier()
    launches = nc.kernelLaunchCount() - launches0
    flags = sc.checkDeviceErrors(dev)
    ms_total = t_begin.elapsed_time(t_end)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_fused = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    ms_ta = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    # the two reference calls separately (outside the headline timed region): xs-evaluations/s, scatter-samples/s
    ms_xs = timed_loop(lambda r: L.ncb200_crosssection_nonoriented_many_dev(
        sc._p, d_e[r % NBUF].data_ptr(), n, d_xs.data_ptr(), sp), args.steps)
    ms_sm = timed_loop(lambda r: L.ncb200_samplescatterisotropic_many_dev(
        sc._h, d_e[r % NBUF].data_ptr(), n, d_eo.data_ptr(), d_mu.data_ptr(), sp), args.steps)
    hist_total = float(d_hist.sum().item())
    clk = clocks.stop() if clocks else None   # sampled over the timed region + the two single-call loops (all under load)

    # ---- per-kernel durations, live: CUDA events recorded by the library on the launching stream around
    # each of its kernels (ncb200_kernel_timing), over a short extra loop of the same step.
    ktimes, qcounts = {}, None
    L.ncb200_kernel_timing(1)
    for _ in range(min(args.steps, 20)):
        step()
    buf = C.create_string_buffer(4096)
    if L.ncb200_kernel_timing_report(buf, 4096) > 0:
        ktimes = json.loads(buf.value.decode())
    L.ncb200_kernel_timing(0)
    qc = (C.c_uint32 * 3)()
    if L.ncb200_last_queue_counts(sc._h, qc) == 0:
        qcounts = [int(qc[i]) for i in range(3)]

    # ---- e2e: reference-facing C entry points, pinned host buffers, copies inside the timed region
    h_e = torch.empty(n, dtype=torch.float64).pin_memory()
    h_e.copy_(d_e[0])
    h_xs = torch.empty(n, dtype=torch.float64).pin_memory()
    h_eo = torch.empty(n, dtype=torch.float64).pin_memory()
    h_mu = torch.empty(n, dtype=torch.float64).pin_memory()
    dp = C.POINTER(C.c_double)

    def e2e_step():
        L.ncrystal_crosssection_nonoriented_many(sc._p, C.cast(h_e.data_ptr(), dp), n, 1, C.cast(h_xs.data_ptr(), dp))
        L.ncrystal_samplescatterisotropic_many(sc._h, C.cast(h_e.data_ptr(), dp), n, 1, C.cast(h_eo.data_ptr(), dp),
                                               C.cast(h_mu.data_ptr(), dp))

    e2e_steps = max(2, min(args.steps, 10))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    nc.core._check_error()
    e2e_val = world * n * e2e_steps / e2e_s
    mean_mu = float(h_mu.mean())

    # ---- for information: the same result through ONE fused host-pointer call (the reference's experimental batch
    # ABI has such an entry, NCABIUtils.hh:78-100; its C-API does not): 8 B in + 24 B out per neutron over the bus
    def e2e_fused_step():
        L.ncb200_xs_and_samplescatterisotropic_many(sc._h, C.cast(h_e.data_ptr(), dp), n, C.cast(h_xs.data_ptr(), dp),
                                                    C.cast(h_eo.data_ptr(), dp), C.cast(h_mu.data_ptr(), dp))
    e2e_fused_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_fused_step()
    barrier()
    tf = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
    nc.core._check_error()
    e2e_fused_val = world * n * e2e_steps / float(tf.item())

    # ---- secondary figure: the device-resident transport step (ncb200_minimc_run) on the same
    # material: 1e7 source neutrons per GPU through a 10 cm Al sphere, tallies all-reduced over the ranks (NCCL)
    transport = None
    try:
        from ncrystal_b200.sharding import minimc_sharded
        n_src = world * 10_000_000
        geom, src = "sphere;r=0.05", "constant;wl=1.8;z=-0.05;n=%d" % n_src
        eng = "tally=theta,mu;seed=%d" % SEED
        minimc_sharded(sc, geom, src, eng, device=dev)   # warm-up (population buffers, NCCL)
        barrier()
        from ncrystal_b200.sharding import shard_range as _sr, merge_minimc_results
        b0, b1 = _sr(n_src, rank, world)
        t0 = time.perf_counter()
        part = sc.minimc(geom, src, eng, first=b0, count=b1 - b0)
        t_local = time.perf_counter() - t0
        res = merge_minimc_results(part, device=dev)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt, t_local], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, t_local = [float(x) for x in tt.tolist()]
        md = res["output"]["metadata"]
        transport = {"workload": "Al sphere r=5cm, pencil beam 1.8 Aa, %d source neutrons, tallies theta+mu" % n_src,
                     "histories_per_s": n_src / dt, "tally_records_per_s": md["tallied"]["count"] / dt,
                     "seconds": dt, "seconds_slice_max": t_local, "tallied_weight_fraction": md["tallied"]["weight"] / n_src}
    except Exception as e:  # noqa: BLE001
        transport = {"error": str(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * n * args.steps / (ms_total * 1e-3)
    peak, peak_kind = peaks()
    BYTES_FUSED = 32   # 8 in + 8 xs + 8 E' + 8 mu (fused xs+sample_iso)
    ach_seq = n * BYTES_FUSED / (ms_fused * 1e-3) / 1e9
    # dominant kernel: k_sample_sab_refill (S(alpha,beta)-table sampling of the neutrons queued for it):
    # algorithmic bytes per unit = BYTES_SAMPLE (8 B energy in, 16 B (E', mu) out), units = its queue length.
    dom = ktimes.get("k_sab_classes") or ktimes.get("k_sample_sab_refill")
    if dom and qcounts:
        dom_units = qcounts[0]
        dom_ms = dom["ms_avg"]
        ach = dom_units * BYTES_SAMPLE / (dom_ms * 1e-3) / 1e9
        dom_bytes = dom_units * BYTES_SAMPLE
    else:
        dom_units, dom_ms, ach, dom_bytes = n, ms_fused, ach_seq, n * BYTES_FUSED
    # DRAM traffic of that kernel from the committed ncu --set full capture (profiles/r1_kernels_final_ncu_full.csv,
    # same command, 1e7-neutron Al batch): dram__bytes_read.sum + dram__bytes_write.sum per launch.
    traffic = 370.7e6 if n == N_PER_GPU else None
    # FP64 side of the reading (north_star: "FP64 pipe utilisation for the sampling kernels against B200 peak"): the
    # vector-FP64 FMA rate measured here with a DFMA probe kernel; the dominant kernel's FP64 pipe utilisation is the
    # ncu figure of the committed capture (sm__inst_executed_pipe_fp64, profiles/r1_kernels_final_ncu_full.csv)
    try:
        fp64_peak = float(L.ncb200_fp64_fma_probe())
    except Exception:  # noqa: BLE001
        fp64_peak = None
    out = {
        "metric": "neutrons/sec (xs eval + sampleScatter)", "value": value, "unit": "neutrons/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": "Al_sg225.ncmat;temp=293.15K powder, 1e7 isotropic crossSection + sampleScatter per GPU per step, "
                        "log-uniform 1e-5..10 eV (BASELINE.json configs[0])",
            "neutrons_per_gpu_per_step": n, "parallelism": "neutron index range sharded over %d GPU(s)" % world,
            "l2": "3 rotating 80 MB input buffers; per-step working set 320 MB > 126 MB L2",
            "xs_per_s": world * n / (ms_xs * 1e-3), "samples_per_s": world * n / (ms_sm * 1e-3),
            "step": "fused xs+sample (ncb200_xs_and_samplescatterisotropic_many_dev) + mu tally",
            "ms_fused_xs_sample": ms_fused, "ms_xs": ms_xs, "ms_sample": ms_sm, "ms_tally": ms_ta,
            "rng": "Philox4x32-10 per-neutron streams", "device_error_flags": flags,
            "tally_total": hist_total, "mean_mu_e2e": mean_mu,
            "transport_step": transport,
        },
        "e2e": {"value": e2e_val, "unit": "neutrons/s", "h2d_bytes_per_step": 2 * 8 * n, "d2h_bytes_per_step": 3 * 8 * n,
                "steps": e2e_steps, "api": "ncrystal_crosssection_nonoriented_many + ncrystal_samplescatterisotropic_many",
                "fused_call": {"value": e2e_fused_val, "api": "ncb200_xs_and_samplescatterisotropic_many (extension)",
                               "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 3 * 8 * n}},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"bound": "hbm", "kernel": "k_sample_sab_refill",
                     "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "traffic": traffic, "peak_source": peak_kind,
                     "algorithmic_bytes_per_launch": dom_bytes, "units_per_launch": dom_units, "ms_per_launch": dom_ms,
                     "note": "rejection sampling in fp64: latency/issue bound, not HBM bound (SURVEY 8d); "
                             "ncu stall and pipe evidence under profiles/",
                     "fp64": {"fma_peak_tflops_measured": fp64_peak, "pipe_pct_dominant_kernel": 23.1,
                              "source": "DFMA probe (live) / ncu capture (committed)"},
                     "kernel_ms": ktimes, "queue_units": qcounts,
                     "launch_sequence": {"achieved": ach_seq, "frac": ach_seq / peak,
                                         "algorithmic_bytes_per_step": n * BYTES_FUSED, "ms": ms_fused},
                     "xs_kernel": {"achieved": n * BYTES_XS / (ms_xs * 1e-3) / 1e9,
                                   "frac": n * BYTES_XS / (ms_xs * 1e-3) / 1e9 / peak}},
    }
    if not args.no_cpu_baseline and world == 1:
        from _libs import have_refdrv
        if have_refdrv():
            nt = host_threads()
            t_xs, t_sm = cpu_reference(cfg, n, nt, 2)
            out["cpu_baseline"] = {"value": n / (t_xs + t_sm), "unit": "neutrons/s", "cores": nt, "kind": "reference",
                                   "sample": "full workload (1e7 neutrons), best of 2 after warm-up",
                                   "xs_per_s": n / t_xs, "samples_per_s": n / t_sm}
        else:
            out["cpu_baseline"] = cpu_port_baseline(cfg)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
