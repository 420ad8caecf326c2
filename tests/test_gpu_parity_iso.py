"""GPU parity tests (isotropic configs): the CUDA path, called through the C ABI, against
 (a) the committed golden vectors generated from the unmodified reference, and
 (b) the live oracle (reference via oracle/_ref when present, else the C restatement).
Tolerances are north_star's: cross sections 1e-12 relative; replayed scatter outcomes 1e-10."""
import numpy as np
import pytest

from conftest import CONFIG_KEYS_ISO, golden
from _parity import assert_replay

pytestmark = pytest.mark.gpu

XS_RTOL = 1e-12
SAMPLE_TOL = 1e-10


def _scatter(cfg, seed=0):
    import ncrystal_b200 as nc
    return nc.Scatter(cfg, seed=seed)


def _match(eo, mu, eo_ref, mu_ref):
    ok_e = np.abs(eo - eo_ref) <= SAMPLE_TOL * np.maximum(np.abs(eo_ref), 1e-300)
    ok_m = np.abs(mu - mu_ref) <= SAMPLE_TOL
    return ok_e & ok_m


@pytest.mark.parametrize("key", CONFIG_KEYS_ISO)
def test_xs_vs_golden(key, configs):
    g = golden(key)
    sc = _scatter(configs[key])
    xs = sc.crossSectionIsotropic(g["ekin"])
    ref = g["xs"]
    finite = np.isfinite(ref) & (ref != 0)
    rel = np.abs(xs[finite] - ref[finite]) / np.abs(ref[finite])
    assert rel.max() <= XS_RTOL, "max rel err %g" % rel.max()
    assert np.array_equal(xs[~finite], ref[~finite])
    # repeat semantics: results[r*n+i]
    xs3 = sc.crossSectionIsotropic(g["ekin"][:100], repeat=3)
    assert np.array_equal(xs3, np.tile(xs[:100], 3))
    # scalar entry point
    assert sc.crossSectionIsotropic(float(g["ekin"][2])) == xs[2]


@pytest.mark.parametrize("key", CONFIG_KEYS_ISO)
def test_sample_replay_vs_golden(key, configs):
    import torch
    import ncrystal_b200 as nc
    g = golden(key)
    seed = int(g["seed"])
    sc = _scatter(configs[key], seed=seed)
    sc.setRNGStream(seed, 0, 0)
    eo, mu = sc.sampleScatterIsotropic(g["ekin"])
    ok = _match(eo, mu, g["ekin_out"], g["mu"])
    # draw counts through the device path
    d_e = torch.from_numpy(g["ekin"]).cuda()
    nd = torch.zeros(d_e.numel(), dtype=torch.int32, device="cuda")
    comp = torch.zeros(d_e.numel(), dtype=torch.int32, device="cuda")
    sc.setRNGStream(seed, 0, 0)
    sc._L.ncb200_set_diagnostics_dev(sc._h, nd.data_ptr(), comp.data_ptr())
    eo2, mu2 = sc.sampleScatterIsotropic(d_e)
    sc.checkDeviceErrors()
    assert np.array_equal(eo2.cpu().numpy(), eo) and np.array_equal(mu2.cpu().numpy(), mu)
    nd = nd.cpu().numpy().astype(np.uint32)
    flips = nd != g["ndraws"]
    print("%s: match %.6f, branch flips (draw count differs) %d, numeric-only mismatches %d, bit-exact %.6f"
          % (key, ok.mean(), flips.sum(), (~ok & ~flips).sum(),
             ((eo == g["ekin_out"]) & (mu == g["mu"])).mean()))
    assert_replay((eo, mu), (g["ekin_out"], g["mu"]), nd, g["ndraws"], key)
    assert np.all(np.abs(mu) <= 1.0) and np.all(eo >= 0.0)


@pytest.mark.parametrize("key", CONFIG_KEYS_ISO)
def test_vs_live_oracle_larger(key, configs):
    from oracle_check import oracle_for
    from _libs import loguniform_energies
    orc = oracle_for(configs[key])
    n = 300000
    e = loguniform_energies(n, seed=4242)
    sc = _scatter(configs[key], seed=5)
    xs = sc.crossSectionIsotropic(e)
    ref = orc.xs_iso(e)
    rel = np.abs(xs - ref) / np.abs(ref)
    assert rel.max() <= XS_RTOL
    sc.setRNGStream(5, 0, 1000)
    eo, mu = sc.sampleScatterIsotropic(e)
    eo_r, mu_r, nd_r = orc.sample_iso(e, seed=5, first_index=1000)[:3]
    ok = _match(eo, mu, eo_r, mu_r)
    print("%s (%s oracle): xs max rel %.2e; replay match %.7f over %d" % (key, orc.kind, rel.max(), ok.mean(), n))
    assert_replay((eo, mu), (eo_r, mu_r), None, None, "%s vs live oracle" % key)


@pytest.mark.parametrize("key", CONFIG_KEYS_ISO)
def test_sab_tables_built_on_device(key, configs):
    """The S(alpha,beta) sampler tables built by the device kernels reproduce the xs grid that the
    reference's SABIntegrator computed (stored in the compiled material) to 1e-13."""
    import ctypes as C
    sc = _scatter(configs[key])
    for ic, (kind, scale) in enumerate(sc.components()):
        if kind != 3:
            continue
        out = np.zeros(4096)
        n = sc._L.ncb200_sab_xscheck(sc._p, ic, out.ctypes.data_as(C.POINTER(C.c_double)), out.size)
        assert n > 10
        g = golden(key)
        # xs grid from the blob: evaluate the SAB xs exactly at the grid energies through hostsim-free route:
        # compare against the reference-computed grid shipped in the compiled material
        from oracle_check import material_path
        from test_cpu_blob import sab_grids
        egrid, xsgrid = sab_grids(open(material_path(configs[key]), "rb").read(), ic)
        assert n == egrid.size
        rel = np.abs(out[:n] - xsgrid) / np.abs(xsgrid)
        assert rel.max() < 1e-13, rel.max()


def test_empty_and_single(configs):
    sc = _scatter(configs["Al"], seed=3)
    assert sc.crossSectionIsotropic(np.zeros(0)).size == 0
    eo, mu = sc.sampleScatterIsotropic(np.zeros(0))
    assert eo.size == 0 and mu.size == 0
    e1, m1 = sc.sampleScatterIsotropic(0.025)
    assert e1 >= 0 and -1 <= m1 <= 1


def test_streams_shard_invariant(configs):
    """Results depend only on (seed, stream, global neutron index): evaluating a batch in two shards
    with the matching first_index reproduces the single-call result bit for bit (the property
    multi-GPU sharding relies on)."""
    from _libs import loguniform_energies
    e = loguniform_energies(50001, seed=11)
    sc = _scatter(configs["Al"], seed=9)
    sc.setRNGStream(9, 0, 0)
    eo, mu = sc.sampleScatterIsotropic(e)
    k = 20011
    sc.setRNGStream(9, 0, 0)
    eo_a, mu_a = sc.sampleScatterIsotropic(e[:k])
    assert sc.getRNGStream() == (9, 0, k)
    eo_b, mu_b = sc.sampleScatterIsotropic(e[k:])
    assert np.array_equal(np.concatenate([eo_a, eo_b]), eo)
    assert np.array_equal(np.concatenate([mu_a, mu_b]), mu)
    # clones share tables but draw from an independent stream
    c = sc.clone()
    c_eo, c_mu = c.sampleScatterIsotropic(e[:1000])
    assert not np.array_equal(c_mu, mu[:1000])
    st = sc.getRNGState()
    x1 = sc.sampleScatterIsotropic(e[:100])
    sc.setRNGState(st)
    x2 = sc.sampleScatterIsotropic(e[:100])
    assert np.array_equal(x1[1], x2[1])


@pytest.mark.parametrize("key", CONFIG_KEYS_ISO)
def test_free_gas_staged_kernels_replay_vs_golden(key, configs):
    # the free-gas queue has two implementations (one kernel, neutron per lane / five staged kernels with
    # attempt-level scheduling; the batch size picks one): both must reproduce the reference's replayed outcomes
    # and draw counts, and each other bit for bit
    import torch
    g = golden(key)
    seed = int(g["seed"])
    sc = _scatter(configs[key], seed=seed)
    d_e = torch.from_numpy(g["ekin"]).cuda()
    res = {}
    try:
        for name, nmin in (("staged", 1), ("single", 1 << 40)):
            sc._L.ncb200_set_fg_staged_min(nmin)
            nd = torch.zeros(d_e.numel(), dtype=torch.int32, device="cuda")
            sc.setRNGStream(seed, 0, 0)
            sc._L.ncb200_set_diagnostics_dev(sc._h, nd.data_ptr(), None)
            eo, mu = sc.sampleScatterIsotropic(d_e)
            sc.checkDeviceErrors()
            res[name] = (eo.cpu().numpy(), mu.cpu().numpy(), nd.cpu().numpy().astype(np.uint32))
    finally:
        sc._L.ncb200_set_fg_staged_min(4000000)
    for a, b in zip(res["staged"], res["single"]):
        assert np.array_equal(a, b)
    eo, mu, nd = res["staged"]
    assert_replay((eo, mu), (g["ekin_out"], g["mu"]), nd, g["ndraws"], "%s staged free gas" % key)
