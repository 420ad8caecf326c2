"""Multi-GPU sharding of a neutron batch: every neutron is independent, tables are
replicated per GPU, the GLOBAL neutron index range is partitioned contiguously over the ranks and
the random stream of a neutron is keyed by its global index -- so the union of the shards is
bit-identical to a single-GPU run for any world size.  The only collective is the element-wise sum
of tally arrays (the analogue of the reference's Tally::merge, NCMMC_Tally.hh:40-62)."""


def shard_range(n_total, rank, world):
    """Contiguous [begin, end) of rank `rank`; sizes differ by at most one, order preserved."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def merge_tallies(hist, group=None):
    """In-place sum of a tally tensor over all ranks (NCCL on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


# ---- transport step (ncb200_minimc_run_slice): ranks simulate contiguous slices of the source, tallies add

def source_count(srccfg):
    """number of source neutrons of a MiniMC source cfg string (parameter n, default 1e6)"""
    for tok in srccfg.split(";"):
        k, _, v = tok.partition("=")
        if k.strip() == "n":
            return int(float(v) + 0.5)
    return 1000000


def _hist_arrays(h):
    b = h["bindata"]
    s = h["stats"]
    integral = s.get("integral") or 0.0
    mean = s.get("mean") or 0.0
    rms = s.get("rms") or 0.0
    sums = [integral, integral * mean, integral * (rms * rms + mean * mean)]
    return ([b["underflow"]] + list(b["content"]) + [b["overflow"]],
            [b["underflow_errorsq"]] + list(b["errorsq"]) + [b["overflow_errorsq"]], sums,
            s.get("minfilled"), s.get("maxfilled"))


def _set_hist(h, c, e, sums, lo, hi):
    b = h["bindata"]
    b["underflow"], b["overflow"] = c[0], c[-1]
    b["content"] = list(c[1:-1])
    b["underflow_errorsq"], b["overflow_errorsq"] = e[0], e[-1]
    b["errorsq"] = list(e[1:-1])
    if sums[0] > 0.0:
        mean = sums[1] / sums[0]
        h["stats"] = dict(integral=sums[0], mean=mean, rms=max(0.0, sums[2] / sums[0] - mean * mean) ** 0.5,
                          minfilled=lo, maxfilled=hi)
    else:
        h["stats"] = dict(integral=0.0, mean=None, rms=None, minfilled=None, maxfilled=None)


def merge_minimc_results(res, device=None, group=None):
    """All-reduce the tallies and counters of per-rank transport results (result dictionaries of
    Scatter.minimc(..., first, count) for disjoint source slices) -- the analogue of the reference's
    TallyStdHists::merge (NCMMC_StdTallies.cc) + Hist1D::merge / RunningStats1D::merge (NCHists.hh:532-539).
    Returns the merged dictionary (identical on all ranks)."""
    import copy
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return res
    out = copy.deepcopy(res)
    hists = []
    for name, t in out["output"]["tally"].items():
        hists.append(t["total"])
        hists.extend(t.get("breakdown", {}).values())
    flat, mins, maxs = [], [], []
    for h in hists:
        c, e, sums, lo, hi = _hist_arrays(h)
        flat += c + e + sums
        mins.append(float("inf") if lo is None else lo)
        maxs.append(float("-inf") if hi is None else hi)
    md = out["output"]["metadata"]
    flat += [md[k][f] for k in ("provided", "miss", "tallied") for f in ("count", "weight")]
    kw = dict(dtype=torch.float64, device=device)
    tsum, tmin, tmax = torch.tensor(flat, **kw), torch.tensor(mins, **kw), torch.tensor(maxs, **kw)
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX, group=group)
    v = tsum.cpu().tolist()
    pos = 0
    for k, h in enumerate(hists):
        nb2 = h["bindata"]["nbins"] + 2
        c, e, sums = v[pos:pos + nb2], v[pos + nb2:pos + 2 * nb2], v[pos + 2 * nb2:pos + 2 * nb2 + 3]
        pos += 2 * nb2 + 3
        _set_hist(h, c, e, sums, float(tmin[k]), float(tmax[k]))
    for key in ("provided", "miss", "tallied"):
        md[key]["count"] = int(round(v[pos]))
        md[key]["weight"] = v[pos + 1]
        pos += 2
    return out


def minimc_sharded(scatter, geomcfg, srccfg, enginecfg="", device=None, group=None):
    """Scatter.minimc over all ranks of the process group: rank k simulates its contiguous slice of the source
    neutrons, the tallies are all-reduced (NCCL over NVLink on GPUs)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    b, e = shard_range(source_count(srccfg), rank, world)
    res = scatter.minimc(geomcfg, srccfg, enginecfg, first=b, count=e - b)
    return merge_minimc_results(res, device=device, group=group)
