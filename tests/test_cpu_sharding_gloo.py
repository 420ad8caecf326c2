"""world_size=2 gloo test of the N>1 host logic (no GPU): index-range sharding + per-neutron streams keyed
by GLOBAL index reproduce the single-process result bit for bit, and the tally merge (all_reduce sum)
equals the single-process histogram.  The per-shard compute stand-in is the CPU oracle (checker code)."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT

N = 6001
SEED = 4711
NBINS = 50


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _hist(mu):
    h = np.zeros(NBINS + 2)
    rel = (mu + 1.0) * (NBINS / 2.0)
    idx = np.where(rel < 0, 0, np.where(rel >= NBINS, NBINS + 1, 1 + np.floor(rel))).astype(int)
    np.add.at(h, idx, 1.0)
    return h


def _worker(rank, world, port, blob_path, outdir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from ncrystal_b200.sharding import shard_range, merge_tallies
    from _oracle_port import PortOracle
    from _libs import loguniform_energies
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    e = loguniform_energies(N, seed=99)
    b, en = shard_range(N, rank, world)
    o = PortOracle(open(blob_path, "rb").read())
    eo, mu, nd, er = o.sample_iso(e[b:en], seed=SEED, first_index=b)   # stream index = global index
    hist = torch.from_numpy(_hist(mu))
    merge_tallies(hist)
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), b=b, en=en, eo=eo, mu=mu, hist=hist.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partition():
    from ncrystal_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_matches_single_process(tmp_path, configs):
    import torch.multiprocessing as mp
    from oracle_check import material_path
    from _oracle_port import PortOracle
    from _libs import loguniform_energies
    blob_path = material_path(configs["Al"])
    if not os.path.exists(blob_path):
        pytest.skip("compiled Al material not present")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, blob_path, str(tmp_path)), nprocs=2, join=True)
    e = loguniform_energies(N, seed=99)
    o = PortOracle(open(blob_path, "rb").read())
    eo, mu, nd, er = o.sample_iso(e, seed=SEED, first_index=0)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(2)]
    assert parts[0]["b"] == 0 and parts[0]["en"] == parts[1]["b"] and parts[1]["en"] == N
    assert np.array_equal(np.concatenate([p["eo"] for p in parts]), eo)
    assert np.array_equal(np.concatenate([p["mu"] for p in parts]), mu)
    full = _hist(mu)
    for p in parts:
        assert np.array_equal(p["hist"], full) and p["hist"].sum() == N


# ---- transport step: per-rank source slices + tally merge (the per-rank compute stand-in is the CPU oracle)

def _mmc_worker(rank, world, port, outdir):
    import json
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from ncrystal_b200.sharding import shard_range, merge_minimc_results, source_count
    from _mmc import all_scenarios, cached_oracle, run_oracle, result_dict_from_layout
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    sc = all_scenarios()["box_yag"]
    assert source_count(sc.srccfg()) == sc.n
    b, e = shard_range(sc.n, rank, world)
    o, _ = cached_oracle(sc.material)
    h, meta = run_oracle(o, sc, first=b, count=e - b)
    merged = merge_minimc_results(result_dict_from_layout(h, meta, sc.tallies, e - b))
    json.dump(merged, open(os.path.join(outdir, "mmc_rank%d.json" % rank), "w"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_transport_tally_merge(tmp_path, configs):
    import json
    import torch.multiprocessing as mp
    from oracle_check import material_path
    from _mmc import all_scenarios, cached_oracle, run_oracle, result_dict_from_layout
    if not os.path.exists(material_path(configs["YAG"])):
        pytest.skip("compiled YAG material not present")
    port = _free_port()
    mp.spawn(_mmc_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sc = all_scenarios()["box_yag"]
    o, _ = cached_oracle(sc.material)
    h, meta = run_oracle(o, sc)
    whole = result_dict_from_layout(h, meta, sc.tallies, sc.n)
    parts = [json.load(open(os.path.join(str(tmp_path), "mmc_rank%d.json" % r))) for r in range(2)]
    assert parts[0] == parts[1]
    m = parts[0]["output"]
    assert m["metadata"]["provided"]["count"] == sc.n
    assert m["metadata"]["miss"]["count"] == whole["output"]["metadata"]["miss"]["count"]
    assert m["metadata"]["tallied"]["count"] == whole["output"]["metadata"]["tallied"]["count"]
    for name, *_ in sc.tallies:
        for which in ["total"] + list(m["tally"][name]["breakdown"]):
            a = m["tally"][name][which] if which == "total" else m["tally"][name]["breakdown"][which]
            w = whole["output"]["tally"][name][which] if which == "total" else whole["output"]["tally"][name]["breakdown"][which]
            assert np.allclose(a["bindata"]["content"], w["bindata"]["content"], rtol=1e-12, atol=1e-9)
            assert np.allclose(a["bindata"]["errorsq"], w["bindata"]["errorsq"], rtol=1e-12, atol=1e-9)
            for k in ("integral", "mean", "rms", "minfilled", "maxfilled"):
                if w["stats"][k] is None:
                    assert a["stats"][k] is None
                else:
                    assert abs(a["stats"][k] - w["stats"][k]) <= 1e-9 * max(1.0, abs(w["stats"][k]))
