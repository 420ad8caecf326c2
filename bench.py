#!/usr/bin/env python
"""bench.py -- throughput of the hot path on B200: neutrons/s through crossSection + sampleScatter on
synthetic log-uniform 1e-5..10 eV batches, for the five BASELINE.json configs.

    python bench.py --gpus N --steps K --warmup W [--config Al|CH2|H2O|YAG|Ge]   (N>1: under torch.distributed.run)
    python bench.py --impl reference ...                                          (reference CPU path, oracle/_ref)

Default config: Al (BASELINE.json configs[0], Al_sg225 at 293.15 K, 1e7 neutrons per GPU per step, weak scaling).
Ge (configs[4]): 1e9 neutrons with per-neutron directions sharded over the ranks (strong scaling) through the
batched oriented entry points.

One "step" = one pass of the hot path over one batch: cross sections and sampled scatterings for every neutron
(+ the mu tally histogram for the isotropic configs).  `value` = neutrons/s of the whole job with inputs resident in
HBM; `e2e` = the same through the reference-facing host-pointer C entry points with host buffers, copies inside
the timed region.  Ranks shard the global neutron index range (no data-path collective); NCCL only reduces the
tally histogram.  The default line also carries the device-resident figures of the other four configs
(`config.other_configs`).  No figure in the line is a constant: ncu-derived ones (`roofline.traffic`, FP64 pipe
share) are read from profiles/ncu_kernel_metrics.json (made by profiles/extract_ncu.py from a committed capture)
and are null when that file has no entry for the kernel and batch size measured here.
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

SEED = 12345
NBINS = 200
# algorithmic HBM bytes per neutron (tables are cache resident and excluded)
BYTES = {"xs_iso": 16, "sample_iso": 24, "fused_iso": 32, "tally": 8, "xs_aniso": 40, "sample_aniso": 64}
METRIC = "neutrons/sec (xs eval + sampleScatter)"

WORKLOADS = {
    "Al": dict(kind="iso", n=10_000_000, idx=0),
    "CH2": dict(kind="iso", n=10_000_000, idx=1),
    "H2O": dict(kind="iso", n=10_000_000, idx=2),
    "YAG": dict(kind="iso", n=10_000_000, idx=3),
    "Ge": dict(kind="aniso", n_total=1_000_000_000, idx=4),
}
# which of the library's kernels the algorithmic bytes of a call are attributed to (the dominant one)
SAMPLER_KERNELS = ("k_sab_classes", "k_sample_sab_refill")


def workload_text(key):
    """One string per config, used verbatim by both arms (this library and --impl reference)."""
    from __graft_entry__ import CONFIGS
    w = WORKLOADS[key]
    if w["kind"] == "iso":
        return ("%s powder/isotropic: 1e7 crossSection + sampleScatter calls per GPU per step, log-uniform 1e-5..10 eV "
                "(BASELINE.json configs[%d])" % (CONFIGS[key], w["idx"]))
    return ("%s oriented single crystal: crossSection + sampleScatter with per-neutron isotropic directions, 1e9 neutrons "
            "sharded over the GPUs, log-uniform 1e-5..10 eV (BASELINE.json configs[%d])" % (CONFIGS[key], w["idx"]))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_metrics():
    """Per-kernel figures of the committed ncu --set full capture (profiles/extract_ncu.py), or {}."""
    p = os.path.join(ROOT, "profiles", "ncu_kernel_metrics.json")
    try:
        return json.load(open(p))
    except Exception:  # noqa: BLE001
        return {}


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


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------ reference (CPU) arm

def cpu_reference(key, n, nthreads, nrep, seed=SEED):
    """The reference's own C-API on the host cores (oracle/_ref), one cloned handle per thread.  Isotropic configs:
    ncrystal_crosssection_nonoriented_many + ncrystal_samplescatterisotropic_many; Ge: per-neutron
    ncrystal_crosssection + ncrystal_samplescatter loops.  Returns (t_xs, t_sample) seconds."""
    import numpy as np
    from _libs import RefDrv, loguniform_energies
    from __graft_entry__ import CONFIGS
    cfg = CONFIGS[key]
    L = RefDrv.lib()
    ekin = loguniform_energies(n, seed=seed)
    dp = C.POINTER(C.c_double)
    null = C.cast(None, dp)

    def p(a):
        return a.ctypes.data_as(dp)
    if WORKLOADS[key]["kind"] == "iso":
        o = [np.empty(n) for _ in range(2)]
        t_xs = L.refdrv_bench_capi(cfg.encode(), 0, nthreads, nrep, p(ekin), null, null, null, n, p(o[0]), null, null, null)
        t_sm = L.refdrv_bench_capi(cfg.encode(), 1, nthreads, nrep, p(ekin), null, null, null, n, p(o[0]), p(o[1]), null, null)
        return t_xs, t_sm
    rng = np.random.Generator(np.random.Philox(key=seed + 1))
    z = 2.0 * rng.random(n) - 1.0
    phi = 2.0 * np.pi * rng.random(n)
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    ux, uy, uz = r * np.cos(phi), r * np.sin(phi), z
    o = [np.empty(n) for _ in range(4)]
    t_xs = L.refdrv_bench_capi(cfg.encode(), 2, nthreads, nrep, p(ekin), p(ux), p(uy), p(uz), n, p(o[0]), null, null, null)
    t_sm = L.refdrv_bench_capi(cfg.encode(), 3, nthreads, nrep, p(ekin), p(ux), p(uy), p(uz), n, p(o[0]), p(o[1]), p(o[2]), p(o[3]))
    return t_xs, t_sm


def cpu_port_baseline(key, n=2_000_000):
    """Fallback when the compiled reference is absent: the plain-C oracle port on all host threads, bounded sample."""
    from _libs import loguniform_energies
    from __graft_entry__ import CONFIGS
    try:
        from oracle_check import oracle_for
        o = oracle_for(CONFIGS[key], prefer="port")
        nt = host_threads()
        ekin = loguniform_energies(n, seed=SEED)
        t_xs = o.bench(0, nt, ekin)
        t_sm = o.bench(1, nt, ekin)
        return {"value": n / (t_xs + t_sm), "unit": "neutrons/s", "cores": nt, "kind": "port",
                "sample": "%d neutrons, xs + sample, oracle/ C port" % n, "xs_per_s": n / t_xs, "samples_per_s": n / t_sm}
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": "neutrons/s", "cores": 0, "kind": "port", "sample": "oracle unavailable: %s" % e}


def cpu_baseline_block(key, budget_s=12.0):
    """cpu_baseline of the bench line: the reference on all host threads and on one thread, each on a bounded
    sample of the workload sized from a probe (about budget_s seconds of CPU work in total)."""
    from _libs import have_refdrv
    if not have_refdrv():
        return cpu_port_baseline(key)
    nt = host_threads()
    probe = 100_000 if WORKLOADS[key]["kind"] == "iso" else 20_000
    a, b = cpu_reference(key, probe, nt, 1)
    rate = probe / (a + b)
    n_full = WORKLOADS[key].get("n", 10_000_000)
    n = int(min(n_full, max(probe, 0.35 * budget_s * rate)))    # (a call = warm-up pass + nrep timed passes)
    t_xs, t_sm = cpu_reference(key, n, nt, 1)
    n1 = int(min(n, max(probe // 4, 0.15 * budget_s * rate / max(nt, 1))))
    s_xs, s_sm = cpu_reference(key, n1, 1, 1)
    return {"value": n / (t_xs + t_sm), "unit": "neutrons/s", "cores": nt, "kind": "reference",
            "sample": "%d neutrons of the workload on %d threads (1 warm-up + 1 timed pass); 1 thread: %d neutrons" % (n, nt, n1),
            "xs_per_s": n / t_xs, "samples_per_s": n / t_sm,
            "one_thread": {"value": n1 / (s_xs + s_sm), "xs_per_s": n1 / s_xs, "samples_per_s": n1 / s_sm, "cores": 1}}


def run_reference(args, key):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from _libs import have_refdrv
    nthreads = host_threads()
    wl = workload_text(key)
    base = {"impl": "reference", "metric": METRIC, "unit": "neutrons/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True,
            "scaling": "weak" if WORKLOADS[key]["kind"] == "iso" else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
    if not have_refdrv():
        # the compiled reference did not travel: time the oracle port instead (kind "port")
        for _ in range(max(args.warmup, 1)):
            cb = cpu_port_baseline(key)
        vals = [cpu_port_baseline(key) for _ in range(args.steps)]
        if vals[0]["value"] is None:
            print(json.dumps({"impl": "reference", "unavailable": vals[0]["sample"]}))
            return
        val = sum(v["value"] for v in vals) / len(vals)
        cb = dict(vals[0], value=val)
        print(json.dumps(dict(base, value=val, ms_per_step=1e3 * 2_000_000 / val,
                              config={"workload": wl, "neutrons_per_step": 2_000_000}, cpu_baseline=cb,
                              e2e={"value": val, "unit": "neutrons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})))
        return
    # each step: a bounded sample of the workload through the reference's C-API on all host cores; the sample is
    # sized from a probe so that the whole --steps run stays within ~2 minutes (the full 1e7 when that fits).
    probe = 1_000_000 if WORKLOADS[key]["kind"] == "iso" else 50_000
    a, b = cpu_reference(key, probe, nthreads, 1)
    rate = probe / (a + b)
    n_full = WORKLOADS[key].get("n", 10_000_000)
    n = int(min(n_full, max(probe // 5, 60.0 * rate / max(args.steps + args.warmup, 1))))   # (each call: warm-up + timed pass)
    for _ in range(max(args.warmup, 1)):
        cpu_reference(key, n, nthreads, 1)
    t0 = time.perf_counter()
    txs = tsm = 0.0
    for _ in range(args.steps):
        a, b = cpu_reference(key, n, nthreads, 1)
        txs += a; tsm += b
    wall = time.perf_counter() - t0
    t_step = (txs + tsm) / args.steps
    val = n / t_step
    api = ("ncrystal_crosssection_nonoriented_many + ncrystal_samplescatterisotropic_many" if WORKLOADS[key]["kind"] == "iso"
           else "per-neutron ncrystal_crosssection + ncrystal_samplescatter")
    print(json.dumps(dict(
        base, value=val, ms_per_step=1e3 * t_step,
        config={"workload": wl, "neutrons_per_step": n, "xs_per_s": n * args.steps / txs, "samples_per_s": n * args.steps / tsm,
                "note": "reference NCrystal 4.4.2 C-API (%s) on host cores; wall %.1fs incl. handle setup" % (api, wall)},
        cpu_baseline={"value": val, "unit": "neutrons/s", "cores": nthreads, "kind": "reference",
                      "sample": "%d neutrons per step (bounded sample of the workload)" % n},
        e2e={"value": val, "unit": "neutrons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})))


# ------------------------------------------------------------------------------------------------- this library

class Ctx:
    """torch plumbing shared by the measurements: device, stream, NCCL, event timing."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.current_stream(self.dev)
        self.sp = C.c_void_p(self.stream.cuda_stream)
        # library messages (e.g. the reference's "reverts to isotropic model" warning, raised on the device) go to
        # stderr: stdout carries the one JSON line
        from ncrystal_b200 import _lib
        self._msgh = C.CFUNCTYPE(None, C.c_char_p, C.c_uint)(
            lambda m, t: sys.stderr.write("[ncrystal_b200 %s] %s\n" % (("info", "warning", "raw")[min(t, 2)], m.decode(errors="replace"))))
        _lib.lib().ncrystal_setmsghandler(C.cast(self._msgh, C.c_void_p))

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def timed_loop(self, fn, reps):
        a, b = self.event(), self.event()
        fn(0)
        self.torch.cuda.synchronize()
        a.record(self.stream)
        for r in range(reps):
            fn(r)
        b.record(self.stream)
        self.torch.cuda.synchronize()
        return a.elapsed_time(b) / reps


def kernel_times(L, fn, reps):
    """Per-kernel durations, live: CUDA events recorded by the library on the launching stream around each of its
    kernels (ncb200_kernel_timing), over a short extra loop of the same step."""
    L.ncb200_kernel_timing(1)
    for r in range(reps):
        fn(r)
    buf = C.create_string_buffer(8192)
    kt = {}
    if L.ncb200_kernel_timing_report(buf, 8192) > 0:
        kt = json.loads(buf.value.decode())
    L.ncb200_kernel_timing(0)
    return kt


def queue_counts(L, sc):
    qc = (C.c_uint32 * 3)()
    if L.ncb200_last_queue_counts(sc._h, qc) == 0:
        return [int(qc[i]) for i in range(3)]
    return None


def dominant(ktimes):
    """(name, ms per launch) of the kernel with the largest share of the step."""
    best = None
    for k, v in ktimes.items():
        if best is None or v["ms_avg"] > best[1]:
            best = (k, v["ms_avg"])
    return best


def roofline_block(ktimes, qcounts, n, peak, peak_kind, kind, ms_sequence, seq_bytes):
    """roofline object for the dominant kernel of the measured launch sequence.  achieved = algorithmic bytes of the
    units that kernel processed per launch / its live CUDA-event duration; traffic and the FP64 pipe share come from
    the committed ncu capture when it has this kernel at this batch size."""
    dom = dominant(ktimes)
    name, ms = dom if dom else ("launch sequence", ms_sequence)
    if name in SAMPLER_KERNELS and qcounts:
        units, per_unit = qcounts[0], BYTES["sample_iso"]
        unit_what = "neutrons queued for the S(alpha,beta) table sampler x 24 B (8 B energy in, 16 B (E',mu) out)"
    elif kind == "aniso":
        xs_side = name in ("k_sc_find", "k_sc_eval", "k_sc_scan", "k_xs_aniso_pre")
        units, per_unit = n, (BYTES["xs_aniso"] if xs_side else BYTES["sample_aniso"])
        unit_what = "neutrons of the block x %d B (SURVEY 8d %s)" % (per_unit, "xs_aniso" if xs_side else "sample_aniso")
    else:
        units, per_unit = n, BYTES["sample_iso"]
        unit_what = "neutrons of the batch x 24 B"
    ach = units * per_unit / (ms * 1e-3) / 1e9
    m = ncu_metrics().get(name, {})
    same = bool(m) and int(m.get("batch_neutrons", -1)) == int(n)
    return {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": (m.get("dram_bytes_per_launch") if same else None), "peak_source": peak_kind,
            "algorithmic_bytes_per_launch": units * per_unit, "units_per_launch": units, "units": unit_what,
            "ms_per_launch": ms,
            "ncu": ({k: m.get(k) for k in ("capture", "commit", "fp64_pipe_pct", "issue_active_pct", "l2_hit_pct",
                                            "lanes_per_inst", "achieved_occupancy_pct", "dram_bytes_per_launch")} if same else None),
            "note": "rejection sampling in fp64 is gather-latency / issue bound, not HBM bound (SURVEY 8d); ncu evidence under profiles/",
            "kernel_ms": ktimes, "queue_units": qcounts,
            "launch_sequence": {"achieved": seq_bytes / (ms_sequence * 1e-3) / 1e9, "frac": seq_bytes / (ms_sequence * 1e-3) / 1e9 / peak,
                                "algorithmic_bytes_per_step": seq_bytes, "ms": ms_sequence}}


def measure_iso(cx, key, n, steps, warmup, headline):
    """Device-resident measurement of an isotropic config.  headline=True also runs the K-step timed region with
    barriers / tally merge (the bench line's `value`)."""
    import ncrystal_b200 as nc
    from ncrystal_b200 import _lib
    from ncrystal_b200.sharding import shard_range, merge_tallies
    from __graft_entry__ import CONFIGS
    torch = cx.torch
    L = _lib.lib()
    sc = nc.Scatter(CONFIGS[key], seed=SEED)
    dev, sp, stream, world, rank = cx.dev, cx.sp, cx.stream, cx.world, cx.rank
    # inputs: 3 rotating device buffers (3 x 80 MB) so that a step's input is never L2-resident from the previous
    # step; per-step working set (in+out) = 320 MB > 126 MB L2.
    NBUF = 3
    first = shard_range(world * n, rank, world)[0]   # weak scaling: n neutrons per GPU, contiguous global index ranges
    d_e = [nc.generateSource(n, seed=SEED + b, first_index=first, device=dev) for b in range(NBUF)]
    d_xs, d_eo, d_mu = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)]
    d_hist = torch.zeros(NBINS + 2, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    counter = [0]

    def step(events=None):
        # total cross section AND sampled scattering for every neutron (the fused entry point, cf. the reference's
        # evalXSAndSampleScatterIsotropic batch ABI, NCABIUtils.hh:78-100), then the mu tally
        k = counter[0]
        counter[0] += 1
        e = d_e[k % NBUF]
        sc.setRNGStream(SEED, 0, k * world * n + first)
        if events is not None:
            events[0].record(stream)
        L.ncb200_xs_and_samplescatterisotropic_many_dev(sc._h, e.data_ptr(), n, d_xs.data_ptr(), d_eo.data_ptr(), d_mu.data_ptr(), sp)
        if events is not None:
            events[1].record(stream)
        L.ncb200_tally_hist_dev(d_mu.data_ptr(), None, n, -1.0, 1.0, NBINS, d_hist.data_ptr(), None, sp)
        if events is not None:
            events[2].record(stream)

    for _ in range(warmup):
        step()
    d_hist.zero_()
    out = {"key": key, "n": n}
    if headline:
        cx.barrier()
        clocks = ClockSampler(cx.local_rank) if rank == 0 else None
        if clocks:
            clocks.start()
        ev = [[cx.event() for _ in range(3)] for _ in range(steps)]
        launches0 = nc.kernelLaunchCount()
        t_begin, t_end = cx.event(), cx.event()
        cx.barrier()
        t_begin.record(stream)
        for k in range(steps):
            step(ev[k])
        merge_tallies(d_hist)                 # the only collective: tally merge (NCCL all-reduce over NVLink)
        t_end.record(stream)
        cx.barrier()
        out["launches"] = nc.kernelLaunchCount() - launches0
        out["ms_total"] = cx.max_over_ranks(t_begin.elapsed_time(t_end))[0]
        out["ms_fused"] = sum(e[0].elapsed_time(e[1]) for e in ev) / steps
        out["ms_tally"] = sum(e[1].elapsed_time(e[2]) for e in ev) / steps
    else:
        clocks = None
        out["ms_fused"] = cx.timed_loop(lambda r: step(), steps)
    # the two reference calls separately (outside the headline timed region): xs-evaluations/s, scatter-samples/s
    out["ms_xs"] = cx.timed_loop(lambda r: L.ncb200_crosssection_nonoriented_many_dev(
        sc._p, d_e[r % NBUF].data_ptr(), n, d_xs.data_ptr(), sp), steps)
    out["ms_sample"] = cx.timed_loop(lambda r: L.ncb200_samplescatterisotropic_many_dev(
        sc._h, d_e[r % NBUF].data_ptr(), n, d_eo.data_ptr(), d_mu.data_ptr(), sp), steps)
    out["flags"] = sc.checkDeviceErrors(dev)
    out["hist_total"] = float(d_hist.sum().item())
    out["clocks"] = clocks.stop() if clocks else None   # sampled over the timed region + the two single-call loops (all under load)
    out["ktimes"] = kernel_times(L, lambda r: step(), min(steps, 10))
    out["qcounts"] = queue_counts(L, sc)
    out["table_MB"] = sc.tableBytes() / 1e6
    out["sc"], out["d_e"] = sc, d_e
    return out


def measure_aniso(cx, key, n_rank, first, steps, warmup, headline, block=1 << 25):
    """Device-resident measurement of the oriented config: crossSection(E,dir) + sampleScatter(E,dir) for every
    neutron of this rank's share, in blocks of 2^25 neutrons (bounded scratch; results do not depend on the blocking)."""
    import ncrystal_b200 as nc
    from ncrystal_b200 import _lib
    from __graft_entry__ import CONFIGS
    torch = cx.torch
    L = _lib.lib()
    sc = nc.Scatter(CONFIGS[key], seed=SEED)
    dev, sp, stream = cx.dev, cx.sp, cx.stream
    n = n_rank
    # inputs resident in HBM before the timed region: 32 B in + 40 B out per neutron (9 GB per GPU at N=8, 72 GB at N=1)
    e, ux, uy, uz = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(4)]
    for b0 in range(0, n, block):
        m = min(block, n - b0)
        ee, (xx, yy, zz) = nc.generateSource(m, seed=SEED, first_index=first + b0, directions=True, device=dev)
        e[b0:b0 + m], ux[b0:b0 + m], uy[b0:b0 + m], uz[b0:b0 + m] = ee, xx, yy, zz
        del ee, xx, yy, zz
    xs, eo, ox, oy, oz = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(5)]
    torch.cuda.synchronize()
    counter = [0]

    def xs_pass():
        for b0 in range(0, n, block):
            m = min(block, n - b0)
            o = 8 * b0
            L.ncb200_crosssection_many_dev(sc._p, e.data_ptr() + o, ux.data_ptr() + o, uy.data_ptr() + o, uz.data_ptr() + o, m,
                                           xs.data_ptr() + o, sp)

    def sample_pass(k):
        for b0 in range(0, n, block):
            m = min(block, n - b0)
            o = 8 * b0
            sc.setRNGStream(SEED, 0, k * 1_000_000_000 + first + b0)
            L.ncb200_samplescatter_manydir_dev(sc._h, e.data_ptr() + o, ux.data_ptr() + o, uy.data_ptr() + o, uz.data_ptr() + o, m,
                                               eo.data_ptr() + o, ox.data_ptr() + o, oy.data_ptr() + o, oz.data_ptr() + o, sp)

    def step(events=None):
        k = counter[0]
        counter[0] += 1
        if events is not None:
            events[0].record(stream)
        xs_pass()
        if events is not None:
            events[1].record(stream)
        sample_pass(k)
        if events is not None:
            events[2].record(stream)

    for _ in range(warmup):
        step()
    out = {"key": key, "n": n}
    cx.barrier()
    clocks = ClockSampler(cx.local_rank) if (cx.rank == 0 and headline) else None
    if clocks:
        clocks.start()
    ev = [[cx.event() for _ in range(3)] for _ in range(steps)]
    launches0 = nc.kernelLaunchCount()
    t_begin, t_end = cx.event(), cx.event()
    cx.barrier()
    t_begin.record(stream)
    for k in range(steps):
        step(ev[k])
    t_end.record(stream)
    cx.barrier()
    out["launches"] = nc.kernelLaunchCount() - launches0
    out["ms_total"] = cx.max_over_ranks(t_begin.elapsed_time(t_end))[0]
    out["ms_xs"] = sum(x[0].elapsed_time(x[1]) for x in ev) / steps
    out["ms_sample"] = sum(x[1].elapsed_time(x[2]) for x in ev) / steps
    out["flags"] = sc.checkDeviceErrors(dev)
    out["clocks"] = clocks.stop() if clocks else None
    # per-kernel times on ONE block of the batch (a launch = one block)
    m = min(block, n)
    out["kt_block"] = m

    def one_block(r):
        L.ncb200_crosssection_many_dev(sc._p, e.data_ptr(), ux.data_ptr(), uy.data_ptr(), uz.data_ptr(), m, xs.data_ptr(), sp)
        L.ncb200_samplescatter_manydir_dev(sc._h, e.data_ptr(), ux.data_ptr(), uy.data_ptr(), uz.data_ptr(), m,
                                           eo.data_ptr(), ox.data_ptr(), oy.data_ptr(), oz.data_ptr(), sp)
    out["ktimes"] = kernel_times(L, one_block, 2)
    out["qcounts"] = None
    out["table_MB"] = sc.tableBytes() / 1e6
    out["mean_eo"] = float(eo[: min(n, 1 << 20)].mean().item())
    out["sc"], out["inputs"] = sc, (e, ux, uy, uz)
    return out


def copy_ceiling(cx, bytes_in, bytes_out, reps=3, chunk=8 << 20):
    """Bare-copy ceiling of a host-pointer call: the same bytes (bytes_in H2D, bytes_out D2H) as concurrent pinned
    copies in 8 MB pieces on two streams, no kernels, all ranks at once.  Seconds (max over ranks)."""
    torch = cx.torch
    hin = torch.empty(bytes_in, dtype=torch.uint8).pin_memory()
    hout = torch.empty(bytes_out, dtype=torch.uint8).pin_memory()
    din = torch.empty(bytes_in, dtype=torch.uint8, device=cx.dev)
    dout = torch.empty(bytes_out, dtype=torch.uint8, device=cx.dev)
    s1, s2 = torch.cuda.Stream(cx.dev), torch.cuda.Stream(cx.dev)

    def once():
        with torch.cuda.stream(s1):
            for o in range(0, bytes_in, chunk):
                din[o:o + chunk].copy_(hin[o:o + chunk], non_blocking=True)
        with torch.cuda.stream(s2):
            for o in range(0, bytes_out, chunk):
                hout[o:o + chunk].copy_(dout[o:o + chunk], non_blocking=True)
        s1.synchronize(); s2.synchronize()
    once()
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    cx.barrier()
    return cx.max_over_ranks((time.perf_counter() - t0) / reps)[0]


def e2e_iso(cx, sc, d_e0, n, steps):
    """The reference-facing host-pointer C entry points on HOST buffers, copies inside the timed region: pinned
    buffers (headline), pageable (malloc'd numpy) buffers, the fused extension call, and the bare-copy ceiling."""
    import numpy as np
    import ncrystal_b200 as nc
    from ncrystal_b200 import _lib
    torch = cx.torch
    L = _lib.lib()
    dp = C.POINTER(C.c_double)
    h = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(4)]
    h[0].copy_(d_e0)
    pg = [np.empty(n, dtype=np.float64) for _ in range(4)]     # pageable
    pg[0][:] = h[0].numpy()

    def ptrs(bufs):
        if isinstance(bufs[0], np.ndarray):
            return [b.ctypes.data_as(dp) for b in bufs]
        return [C.cast(b.data_ptr(), dp) for b in bufs]

    def two_calls(p):
        L.ncrystal_crosssection_nonoriented_many(sc._p, p[0], n, 1, p[1])
        L.ncrystal_samplescatterisotropic_many(sc._h, p[0], n, 1, p[2], p[3])

    def fused(p):
        L.ncb200_xs_and_samplescatterisotropic_many(sc._h, p[0], n, p[1], p[2], p[3])

    def timeit(fn, p):
        fn(p)
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn(p)
        cx.barrier()
        s = cx.max_over_ranks(time.perf_counter() - t0)[0]
        nc.core._check_error()
        return cx.world * n * steps / s
    v_pinned = timeit(two_calls, ptrs(h))
    v_page = timeit(two_calls, ptrs(pg))
    v_fused = timeit(fused, ptrs(h))
    mean_mu = float(h[3].mean())
    # ceiling: xs call (8 B in, 8 B out per neutron) then sampling call (8 B in, 16 B out), as bare copies
    t_ceiling = copy_ceiling(cx, 8 * n, 8 * n) + copy_ceiling(cx, 8 * n, 16 * n)
    v_ceiling = cx.world * n / t_ceiling
    return {"value": v_pinned, "unit": "neutrons/s", "h2d_bytes_per_step": 2 * 8 * n, "d2h_bytes_per_step": 3 * 8 * n,
            "steps": steps, "api": "ncrystal_crosssection_nonoriented_many + ncrystal_samplescatterisotropic_many",
            "host_buffers": "pinned",
            "pageable": {"value": v_page, "host_buffers": "pageable (numpy/malloc)", "frac_of_pinned": v_page / v_pinned},
            "copy_ceiling": {"value": v_ceiling, "what": "same bytes as bare concurrent pinned copies, no kernels, all ranks at once",
                             "frac": v_pinned / v_ceiling},
            "fused_call": {"value": v_fused, "api": "ncb200_xs_and_samplescatterisotropic_many (extension)",
                           "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 3 * 8 * n}}, mean_mu


def e2e_aniso(cx, sc, inputs, n, steps):
    """Host-pointer oriented batch calls on a bounded slice of this rank's share (host buffers of 72 B/neutron)."""
    import ncrystal_b200 as nc
    from ncrystal_b200 import _lib
    torch = cx.torch
    L = _lib.lib()
    dp = C.POINTER(C.c_double)
    m = min(n, 1 << 24)
    h = [torch.empty(m, dtype=torch.float64).pin_memory() for _ in range(9)]
    for k in range(4):
        h[k].copy_(inputs[k][:m])
    p = [C.cast(b.data_ptr(), dp) for b in h]

    def once():
        L.ncb200_crosssection_many(sc._p, p[0], p[1], p[2], p[3], m, p[4])
        L.ncb200_samplescatter_manydir(sc._h, p[0], p[1], p[2], p[3], m, p[5], p[6], p[7], p[8])
    once()
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        once()
    cx.barrier()
    s = cx.max_over_ranks(time.perf_counter() - t0)[0]
    nc.core._check_error()
    t_ceiling = copy_ceiling(cx, 32 * m, 8 * m) + copy_ceiling(cx, 32 * m, 32 * m)
    val = cx.world * m * steps / s
    return {"value": val, "unit": "neutrons/s", "h2d_bytes_per_step": 2 * 32 * m, "d2h_bytes_per_step": 40 * m,
            "steps": steps, "api": "ncb200_crosssection_many + ncb200_samplescatter_manydir (batched forms of ncrystal_crosssection / "
                                    "ncrystal_samplescatter; the reference's C-API has no batch call with per-neutron directions)",
            "host_buffers": "pinned", "sample": "%d neutrons per rank per step (bounded slice of the rank's share)" % m,
            "copy_ceiling": {"value": cx.world * m / t_ceiling, "frac": val / (cx.world * m / t_ceiling)}}


def other_configs(cx, skip, steps=4):
    """Device-resident figures of the configs that are not this line's workload (short runs, same code path)."""
    peak, _ = peaks()
    res = {}
    for key, w in WORKLOADS.items():
        if key == skip:
            continue
        try:
            if w["kind"] == "iso":
                r = measure_iso(cx, key, w["n"], steps, 2, headline=False)
                dom = dominant(r["ktimes"])
                units = r["qcounts"][0] if (dom and dom[0] in SAMPLER_KERNELS and r["qcounts"]) else r["n"]
                res[key] = {"workload": workload_text(key), "neutrons_per_gpu": r["n"],
                            "xs_per_s": cx.world * r["n"] / (r["ms_xs"] * 1e-3), "samples_per_s": cx.world * r["n"] / (r["ms_sample"] * 1e-3),
                            "neutrons_per_s_fused_step": cx.world * r["n"] / (r["ms_fused"] * 1e-3),
                            "xs_frac_of_hbm": r["n"] * BYTES["xs_iso"] / (r["ms_xs"] * 1e-3) / 1e9 / peak,
                            "dominant_kernel": dom[0] if dom else None, "dominant_ms": dom[1] if dom else None,
                            "dominant_frac": (units * BYTES["sample_iso"] / (dom[1] * 1e-3) / 1e9 / peak) if dom else None,
                            "kernel_ms": {k: round(v["ms_avg"], 4) for k, v in r["ktimes"].items()},
                            "queue_units": r["qcounts"], "table_MB": r["table_MB"], "device_error_flags": r["flags"]}
            else:
                # BASELINE.json configs[4] as it is stated: 1e9 neutrons sharded over the ranks (strong scaling; 72 GB of
                # resident arrays on one GPU, 9 GB per GPU on eight), one warm-up pass + `ge_steps` timed passes.
                # NCB200_BENCH_GE_TOTAL overrides the total (smaller boxes).
                from ncrystal_b200.sharding import shard_range
                n_total = int(os.environ.get("NCB200_BENCH_GE_TOTAL", w["n_total"]))
                b0, b1 = shard_range(n_total, cx.rank, cx.world)
                n, ge_steps = b1 - b0, 2
                r = measure_aniso(cx, key, n, b0, ge_steps, 1, headline=False)
                dom = dominant(r["ktimes"])
                res[key] = {"workload": workload_text(key), "neutrons_total_per_step": n_total, "neutrons_per_gpu": n, "steps": ge_steps,
                            "scaling": "strong", "ms_per_step": r["ms_total"] / ge_steps,
                            "neutrons_per_s": n_total * ge_steps / (r["ms_total"] * 1e-3),
                            "xs_per_s": n_total / (cx.max_over_ranks(r["ms_xs"])[0] * 1e-3),
                            "samples_per_s": n_total / (cx.max_over_ranks(r["ms_sample"])[0] * 1e-3),
                            "xs_frac_of_hbm": n * BYTES["xs_aniso"] / (r["ms_xs"] * 1e-3) / 1e9 / peak,
                            "sample_frac_of_hbm": n * BYTES["sample_aniso"] / (r["ms_sample"] * 1e-3) / 1e9 / peak,
                            "dominant_kernel": dom[0] if dom else None, "dominant_ms": dom[1] if dom else None,
                            "kernel_ms_block": {k: round(v["ms_avg"], 4) for k, v in r["ktimes"].items()}, "kernel_block_neutrons": r["kt_block"],
                            "table_MB": r["table_MB"], "device_error_flags": r["flags"]}
                r.pop("inputs", None); r.pop("sc", None)
            del r
            cx.torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            res[key] = {"error": "%s: %s" % (type(e).__name__, e)}
    return res


def transport_figure(cx, sc):
    """Secondary figure: the device-resident transport step (ncb200_minimc_run) on the same
    material: 1e7 source neutrons per GPU through a 10 cm Al sphere, tallies all-reduced over the ranks (NCCL)."""
    try:
        from ncrystal_b200.sharding import minimc_sharded, shard_range, merge_minimc_results
        n_src = cx.world * 10_000_000
        geom, src = "sphere;r=0.05", "constant;wl=1.8;z=-0.05;n=%d" % n_src
        eng = "tally=theta,mu;seed=%d" % SEED
        minimc_sharded(sc, geom, src, eng, device=cx.dev)   # warm-up (population buffers, NCCL)
        cx.barrier()
        b0, b1 = shard_range(n_src, cx.rank, cx.world)
        t0 = time.perf_counter()
        part = sc.minimc(geom, src, eng, first=b0, count=b1 - b0)
        t_local = time.perf_counter() - t0
        res = merge_minimc_results(part, device=cx.dev)
        cx.barrier()
        dt, t_local = cx.max_over_ranks(time.perf_counter() - t0, t_local)
        md = res["output"]["metadata"]
        return {"workload": "Al sphere r=5cm, pencil beam 1.8 Aa, %d source neutrons, tallies theta+mu" % n_src,
                "histories_per_s": n_src / dt, "tally_records_per_s": md["tallied"]["count"] / dt,
                "seconds": dt, "seconds_slice_max": t_local, "tallied_weight_fraction": md["tallied"]["weight"] / n_src}
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)}


def vdos_figure(cx):
    """Secondary figure (rank 0): VDOS -> S(alpha,beta) expansion of the Al curve through
    ncrystal_raw_vdos2kernel of the library -- FFT convolutions and the sum over phonon orders on the device -- beside the
    same call of the compiled reference on the host, and whether the two tables are bit-identical.  The input curve is
    read from the committed golden file (tests/golden/vdos_reference.npz)."""
    if cx.rank != 0:
        return None
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import numpy as np
        import _vdos
        from ncrystal_b200 import _lib
        g = _vdos.load_golden()
        egrid, density = g["in_Al_egrid"], g["in_Al_density"]
        sigma, mass, T = [float(x) for x in g["in_Al_meta"]]
        prod = _vdos.RawVdosAPI(_lib.lib())
        prod.kernel(egrid, density, sigma, mass, T, 1)     # warm-up (module load, memory pool)
        l0 = int(_lib.lib().ncb200_kernel_launch_count())
        best, out = 1e9, None
        for _ in range(3):
            t0 = time.perf_counter()
            out = prod.kernel(egrid, density, sigma, mass, T, 3)
            best = min(best, time.perf_counter() - t0)
        launches = (int(_lib.lib().ncb200_kernel_launch_count()) - l0) // 3
        rec = {"workload": "Al_sg225 VDOS (%d points), 293.15 K, vdoslux 3 -> %dx%d table" % (density.size, out[0].size, out[1].size),
               "api": "ncrystal_raw_vdos2kernel", "seconds": best, "expansions_per_s": 1.0 / best, "gpu_launches": launches,
               "sab_sha256_matches_reference_golden": _vdos.sha(out[2]) == str(g["out_Al_lux3_sab_sha"])}
        if _vdos.have_reference():
            ref = _vdos.reference_api()
            rb = 1e9
            for _ in range(2):
                t0 = time.perf_counter()
                r = ref.kernel(egrid, density, sigma, mass, T, 3)
                rb = min(rb, time.perf_counter() - t0)
            rec.update(reference_seconds=rb, reference_threads="reference's own thread pool",
                       bit_identical_to_live_reference=bool(np.array_equal(out[2], r[2])))
        return rec
    except Exception as e:  # noqa: BLE001
        return {"error": "%s: %s" % (type(e).__name__, e)}


def run_iso(args, key, cx):
    from ncrystal_b200 import _lib
    L = _lib.lib()
    n = args.n or WORKLOADS[key]["n"]
    r = measure_iso(cx, key, n, args.steps, args.warmup, headline=True)
    e2e, mean_mu = e2e_iso(cx, r["sc"], r["d_e"][0], n, max(2, min(args.steps, 10)))
    transport = transport_figure(cx, r["sc"]) if key == "Al" else None
    vdosfig = vdos_figure(cx) if key == "Al" else None
    others = other_configs(cx, key) if not args.no_other_configs else None
    if cx.rank != 0:
        return None
    world = cx.world
    peak, peak_kind = peaks()
    value = world * n * args.steps / (r["ms_total"] * 1e-3)
    roof = roofline_block(r["ktimes"], r["qcounts"], n, peak, peak_kind, "iso", r["ms_fused"], n * BYTES["fused_iso"])
    try:
        fp64_peak = float(L.ncb200_fp64_fma_probe())
    except Exception:  # noqa: BLE001
        fp64_peak = None
    roof["fp64"] = {"fma_peak_tflops_measured": fp64_peak,
                    "pipe_pct_dominant_kernel": roof["ncu"].get("fp64_pipe_pct") if roof.get("ncu") else None,
                    "source": "DFMA probe (live) / profiles/ncu_kernel_metrics.json (committed capture; null if none for this kernel and batch)"}
    xm = ncu_metrics().get("k_xs_iso", {})
    roof["xs_kernel"] = {"achieved": n * BYTES["xs_iso"] / (r["ms_xs"] * 1e-3) / 1e9,
                         "frac": n * BYTES["xs_iso"] / (r["ms_xs"] * 1e-3) / 1e9 / peak,
                         "traffic": xm.get("dram_bytes_per_launch") if int(xm.get("batch_neutrons", -1)) == n else None}
    return {
        "metric": METRIC, "value": value, "unit": "neutrons/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_total"] / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": workload_text(key),
            "neutrons_per_gpu_per_step": n, "parallelism": "neutron index range sharded over %d GPU(s)" % world,
            "l2": "3 rotating 80 MB input buffers; per-step working set 320 MB > 126 MB L2",
            "xs_per_s": world * n / (r["ms_xs"] * 1e-3), "samples_per_s": world * n / (r["ms_sample"] * 1e-3),
            "step": "fused xs+sample (ncb200_xs_and_samplescatterisotropic_many_dev) + mu tally",
            "ms_fused_xs_sample": r["ms_fused"], "ms_xs": r["ms_xs"], "ms_sample": r["ms_sample"], "ms_tally": r["ms_tally"],
            "rng": "Philox4x32-10 per-neutron streams", "device_error_flags": r["flags"],
            "tally_total": r["hist_total"], "mean_mu_e2e": mean_mu, "table_MB": r["table_MB"],
            "transport_step": transport,
            "vdos_expansion": vdosfig,
            "other_configs": others,
        },
        "e2e": e2e, "gpu_launches": int(r["launches"]), "clocks": r["clocks"], "roofline": roof,
    }


def run_aniso(args, key, cx):
    from ncrystal_b200.sharding import shard_range
    n_total = args.n * cx.world if args.n else WORKLOADS[key]["n_total"]
    b0, b1 = shard_range(n_total, cx.rank, cx.world)
    r = measure_aniso(cx, key, b1 - b0, b0, args.steps, args.warmup, headline=True)
    e2e = e2e_aniso(cx, r["sc"], r["inputs"], b1 - b0, max(2, min(args.steps, 5)))
    if cx.rank != 0:
        return None
    peak, peak_kind = peaks()
    value = n_total * args.steps / (r["ms_total"] * 1e-3)
    m = r["kt_block"]
    n_rank = b1 - b0
    roof = roofline_block(r["ktimes"], None, m, peak, peak_kind, "aniso",
                          (r["ms_xs"] + r["ms_sample"]) * m / n_rank, m * (BYTES["xs_aniso"] + BYTES["sample_aniso"]))
    roof["xs_call"] = {"achieved": n_rank * BYTES["xs_aniso"] / (r["ms_xs"] * 1e-3) / 1e9,
                       "frac": n_rank * BYTES["xs_aniso"] / (r["ms_xs"] * 1e-3) / 1e9 / peak}
    roof["sample_call"] = {"achieved": n_rank * BYTES["sample_aniso"] / (r["ms_sample"] * 1e-3) / 1e9,
                           "frac": n_rank * BYTES["sample_aniso"] / (r["ms_sample"] * 1e-3) / 1e9 / peak}
    return {
        "metric": METRIC, "value": value, "unit": "neutrons/s",
        "n_gpus": cx.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_total"] / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": workload_text(key), "neutrons_total_per_step": n_total, "neutrons_per_gpu_per_step": n_rank,
            "parallelism": "global neutron index range sharded contiguously over %d GPU(s); no collective on the path" % cx.world,
            "l2": "inputs 32 B + outputs 40 B per neutron, %d MB per GPU >> 126 MB L2" % (72 * n_rank // 1000000),
            "blocks": "entry points called per block of 2^25 neutrons (bounded scratch)",
            "xs_per_s": n_total / (r["ms_xs"] * 1e-3), "samples_per_s": n_total / (r["ms_sample"] * 1e-3),
            "step": "ncb200_crosssection_many_dev + ncb200_samplescatter_manydir_dev over the rank's share",
            "ms_xs": r["ms_xs"], "ms_sample": r["ms_sample"], "rng": "Philox4x32-10 per-neutron streams",
            "device_error_flags": r["flags"], "mean_ekin_out_sample": r["mean_eo"], "table_MB": r["table_MB"],
        },
        "e2e": e2e, "gpu_launches": int(r["launches"]), "clocks": r["clocks"], "roofline": roof,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="Al", choices=list(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="neutrons per GPU per step (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    args = ap.parse_args()
    key = args.config
    if args.impl == "reference":
        run_reference(args, key)
        return
    args.warmup = max(args.warmup, 3)
    cx = Ctx()
    out = run_iso(args, key, cx) if WORKLOADS[key]["kind"] == "iso" else run_aniso(args, key, cx)
    if cx.rank == 0:
        if not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_block(key)
        print(json.dumps(out))
    if cx.world > 1:
        cx.dist.barrier()
        cx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
