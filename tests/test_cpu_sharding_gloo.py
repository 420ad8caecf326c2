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
