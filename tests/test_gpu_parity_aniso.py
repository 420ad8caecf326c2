"""GPU parity tests for the oriented single-crystal config (Ge, SCBragg) through the C ABI."""
import os

import numpy as np
import pytest

from conftest import HERE

from _parity import assert_replay

pytestmark = pytest.mark.gpu


def _g():
    return np.load(os.path.join(HERE, "golden", "aniso_Ge.npz"))


def test_xs_oriented_vs_golden(configs):
    import ncrystal_b200 as nc
    g = _g()
    sc = nc.Scatter(configs["Ge"], seed=1)
    assert sc.isOriented()
    xs = sc.crossSection(g["ekin"], (g["ux"], g["uy"], g["uz"]))
    ref = g["xs"]
    nz = ref != 0
    rel = np.abs(xs[nz] - ref[nz]) / np.abs(ref[nz])
    print("Ge xs: max rel %.3e over %d (%d Bragg-dominated)" % (rel.max(), nz.sum(), (ref > 10).sum()))
    assert rel.max() <= 1e-12
    assert np.array_equal(xs[~nz], ref[~nz])
    # the reference's own known answers (_testimpl.py:253-254)
    e0 = 0.081804209605330899 / 1.54 ** 2
    assert sc.crossSection(e0, (0., 1., 1.)) == pytest.approx(591.0263476502018, rel=1e-6)
    assert sc.crossSection(e0, (1., 1., 0.)) == pytest.approx(1.667600586136298, rel=1e-6)
    # isotropic entry points must refuse oriented processes, like the reference
    with pytest.raises(nc.NCLogicError):
        sc.crossSectionIsotropic(np.array([0.025]))


def test_sample_oriented_replay_vs_golden(configs):
    import torch
    import ncrystal_b200 as nc
    g = _g()
    seed = int(g["seed"])
    sc = nc.Scatter(configs["Ge"], seed=seed)
    sc.setRNGStream(seed, 0, 0)
    eo, (ox, oy, oz) = sc.sampleScatter(g["ekin"], (g["ux"], g["uy"], g["uz"]))
    tol = 1e-10
    ok = (np.abs(eo - g["ekin_out"]) <= tol * np.maximum(np.abs(g["ekin_out"]), 1e-300))
    for a, b in ((ox, g["ox"]), (oy, g["oy"]), (oz, g["oz"])):
        ok &= np.abs(a - b) <= tol
    d = [torch.from_numpy(np.ascontiguousarray(g[k])).cuda() for k in ("ekin", "ux", "uy", "uz")]
    nd = torch.zeros(d[0].numel(), dtype=torch.int32, device="cuda")
    sc.setRNGStream(seed, 0, 0)
    sc._L.ncb200_set_diagnostics_dev(sc._h, nd.data_ptr(), None)
    eo2, (ox2, oy2, oz2) = sc.sampleScatter(d[0], (d[1], d[2], d[3]))
    sc.checkDeviceErrors()
    assert np.array_equal(eo2.cpu().numpy(), eo) and np.array_equal(ox2.cpu().numpy(), ox)
    flips = nd.cpu().numpy().astype(np.uint32) != g["ndraws"]
    print("Ge replay: match %.6f, branch flips %d, numeric-only mismatches %d" % (ok.mean(), flips.sum(), (~ok & ~flips).sum()))
    assert_replay((eo, ox, oy, oz), (g["ekin_out"], g["ox"], g["oy"], g["oz"]), nd.cpu().numpy(), g["ndraws"], "Ge")
    nrm = ox * ox + oy * oy + oz * oz
    assert np.all(np.abs(nrm - 1) < 1e-9)
    # fixed (E,dir) repeated sampling entry point (ncrystal_samplescatter_many)
    e0 = 0.081804209605330899 / 1.54 ** 2
    er, (rx, ry, rz) = sc.sampleScatter(e0, (0., 1., 1.), repeat=2000)
    assert er.shape == (2000,) and np.all(np.abs(rx * rx + ry * ry + rz * rz - 1) < 1e-9)
    assert (er == e0).mean() > 0.9          # Bragg (elastic) dominated at this orientation


def test_oriented_api_on_isotropic_material(configs):
    """crossSection(E,dir)/sampleScatter(E,dir) on an isotropic material go through the isotropic
    leaves plus a uniformly random azimuth (ScatterIsotropicMat, NCProcImpl.cc:29-37)."""
    import ncrystal_b200 as nc
    from _libs import loguniform_energies, isotropic_directions
    from oracle_check import oracle_for
    orc = oracle_for(configs["Al"])
    if orc.kind != "reference":
        pytest.skip("needs oracle/_ref")
    n = 20000
    e = loguniform_energies(n, seed=31)
    ux, uy, uz = isotropic_directions(n, seed=32)
    sc = nc.Scatter(configs["Al"], seed=77)
    xs = sc.crossSection(e, (ux, uy, uz))
    assert np.abs(xs / orc.xs(e, ux, uy, uz) - 1).max() < 1e-12
    sc.setRNGStream(77, 0, 0)
    eo, (ox, oy, oz) = sc.sampleScatter(e, (ux, uy, uz))
    r = orc.sample(e, ux, uy, uz, seed=77)
    ok = (np.abs(eo - r[0]) <= 1e-10 * np.abs(r[0])) & (np.abs(ox - r[1]) <= 1e-10) & (np.abs(oy - r[2]) <= 1e-10) & (np.abs(oz - r[3]) <= 1e-10)
    print("Al oriented API replay match %.6f" % ok.mean())
    assert_replay((eo, ox, oy, oz), (r[0], r[1], r[2], r[3]), None, None, "Al through the oriented API")
