"""CPU tests of the transport-step oracle (oracle/oracle_mmc.c): it must reproduce the REFERENCE's MiniMC
statistically -- the reference's own acceptance criterion for MiniMC changes (chi-square compatibility of the exit
tallies, p > 0.001; tests/pypath/NCTestUtils/minimc_ref.py:196-209) -- against tallies the reference produced here
at 10x statistics (tests/golden/mmc_reference.json, made by tests/golden/make_golden_mmc.py) and, when the
reference tree is present, against the reference's own stored histograms (tests/data/mmcref/*.json)."""
import json
import os

import numpy as np
import pytest

from _mmc import (all_scenarios, cached_oracle, run_oracle, load_golden, chi2_pvalue, compatible, CLASS_NAMES)

SCEN = ["al_4Aa", "al_1Aa", "circ_h2o", "slab_ch2", "box_yag", "cyl_al", "cylinf_h2o", "scge", "iso_al", "isoshell_ch2",
        "isopoint_box_ch2", "thermal_h2o"]
PMIN = 0.001   # the reference's threshold


@pytest.fixture(scope="module")
def results():
    return {}


def _run(results, key):
    if key not in results:
        sc = all_scenarios()[key]
        o, _ = cached_oracle(sc.material)
        results[key] = (sc,) + run_oracle(o, sc)
    return results[key]


@pytest.mark.parametrize("key", SCEN)
def test_oracle_matches_reference_minimc(results, key):
    sc, h, meta = _run(results, key)
    g = load_golden()[key]
    for name, nb, lo, hi in sc.tallies:
        c = h[name]["content"].sum(axis=0)
        e = h[name]["errsq"].sum(axis=0)
        ok, msg = compatible(name, c, e, g["tallies"][name]["content"], g["tallies"][name]["errsq"], PMIN)
        assert ok, "%s/%s: %s" % (key, name, msg)
    # per-neutron averages of the run metadata
    n_ref = g["n"]
    assert abs(meta["tallied_weight"] / sc.n - g["metadata"]["tallied"]["weight"] / n_ref) < 5e-3
    assert abs(meta["tallied_count"] / sc.n - g["metadata"]["tallied"]["count"] / n_ref) < 0.05
    assert abs(meta["miss_count"] / sc.n - g["metadata"]["miss"]["count"] / n_ref) < 0.01
    # breakdown of the first tally into scattering histories
    first = sc.tallies[0][0]
    mine = h[first]["content"].sum(axis=1) / sc.n
    ref = np.array(g["tallies"][first]["class_integrals"]) / n_ref
    assert np.all(np.abs(mine - ref) < 5e-3), (mine, ref)


def test_unscattered_weight_is_deterministic(results):
    # pencil beam through the centre of a sphere: the NOSCAT weight is exp(-(Sigma_s+Sigma_a)*2r) per neutron, no
    # randomness involved => equal to the reference's to rounding
    sc, h, meta = _run(results, "al_4Aa")
    g = load_golden()["al_4Aa"]
    mine = h["theta"]["content"][0].sum() / sc.n
    ref = g["tallies"]["theta"]["class_integrals"][0] / g["n"]
    assert abs(mine - ref) < 1e-12 * ref


def test_slices_add_up(results):
    sc, h, meta = _run(results, "cyl_al")
    o, _ = cached_oracle(sc.material)
    parts = [run_oracle(o, sc, first=a, count=b) for a, b in ((0, 12345), (12345, sc.n - 12345))]
    tot = parts[0][0]["theta"]["content"] + parts[1][0]["theta"]["content"]
    assert np.allclose(tot, h["theta"]["content"], rtol=1e-12, atol=1e-9)
    assert parts[0][1]["tallied_count"] + parts[1][1]["tallied_count"] == meta["tallied_count"]


REF_MMCREF = "/root/reference/tests/data/mmcref"


@pytest.mark.skipif(not os.path.isdir(REF_MMCREF), reason="reference tree not present")
@pytest.mark.parametrize("key,fn", [("al_4Aa", "mmc_al_4Aa.json"), ("al_1Aa", "mmc_al_0.8Aa.json"),
                                    ("circ_h2o", "mmc_circ_h20.json")])
def test_oracle_matches_reference_stored_histograms(results, key, fn):
    # the very files the reference's tests/scripts/mmc_al.py and mmc_circ.py compare against
    sc, h, meta = _run(results, key)
    ref = json.load(open(os.path.join(REF_MMCREF, fn)))
    b = ref["bindata"]
    assert (b["nbins"], b["xmin"], b["xmax"]) == (90, 0.0, 180.0)
    c_ref = np.array([b["underflow"]] + b["content"] + [b["overflow"]])
    e_ref = np.array([b["underflow_errorsq"]] + b["errorsq"] + [b["overflow_errorsq"]])
    c = h["theta"]["content"].sum(axis=0)
    e = h["theta"]["errsq"].sum(axis=0)
    p, chi, k = chi2_pvalue(c, e, c_ref, e_ref)
    assert p > PMIN, "%s vs %s: chi2=%.1f dof=%d p=%.2g" % (key, fn, chi, k, p)
    assert abs(c.sum() / sc.n - ref["stats"]["integral"] / 1e5) < 5e-3


@pytest.mark.parametrize("key", ["al_4Aa", "slab_ch2", "box_yag", "cylinf_h2o", "scge", "iso_al", "isopoint_box_ch2",
                                 "thermal_h2o"])
def test_product_transport_physics_on_host_equals_oracle(results, key):
    # the product's __host__ __device__ transport code (ncb_mmc.cuh + the scatter physics), compiled by the host
    # compiler in tests/hostsim, must reproduce the oracle restatement history by history: identical record counts
    # and (same libm, no FMA contraction on either side) identical histogram contents
    from _libs import HostSim
    from _mmc import run_hostsim
    from oracle_check import material_path
    from __graft_entry__ import CONFIGS
    sc, h, meta = _run(results, key)
    o, hdr = cached_oracle(sc.material)
    hs = HostSim(open(material_path(CONFIGS[sc.material]), "rb").read())
    n = min(sc.n, 20000)
    ho, mo = run_oracle(o, sc, first=0, count=n)
    hh, mh = run_hostsim(hs, hdr, sc, first=0, count=n)
    assert mh["tallied_count"] == mo["tallied_count"] and mh["miss_count"] == mo["miss_count"]
    assert abs(mh["tallied_weight"] - mo["tallied_weight"]) <= 1e-12 * mo["tallied_weight"]
    for name, *_ in sc.tallies:
        assert np.allclose(hh[name]["content"], ho[name]["content"], rtol=1e-12, atol=1e-12), (key, name)
        assert np.allclose(hh[name]["errsq"], ho[name]["errsq"], rtol=1e-12, atol=1e-12), (key, name)
