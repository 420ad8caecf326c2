"""GPU parity for the materials beyond BASELINE.json's five (__graft_entry__.EXTRA_CONFIGS) through the C ABI, against
golden vectors from the unmodified reference: cross sections to 1e-12 relative, replayed scatterings to 1e-10."""
import os

import numpy as np
import pytest

from conftest import HERE
from _parity import assert_replay

pytestmark = pytest.mark.gpu
EXTRA_ISO = ["Be", "D2O", "AlBe", "gas", "CH2_77K", "V", "YAGCor"]
EXTRA_ANISO = ["Cu_sc", "PG"]


def _scatter(key, seed):
    import ncrystal_b200 as nc
    from __graft_entry__ import EXTRA_CONFIGS
    from oracle_check import material_path
    if not os.path.exists(material_path(EXTRA_CONFIGS[key])):
        pytest.skip("compiled material for %s not present" % key)
    return nc.Scatter(EXTRA_CONFIGS[key], seed=seed)


def _xs_check(xs, ref):
    fin = np.isfinite(ref) & (ref != 0)
    rel = np.abs(xs[fin] - ref[fin]) / np.abs(ref[fin])
    assert rel.max() <= 1e-12, rel.max()
    assert np.array_equal(xs[~fin], ref[~fin])
    return rel.max()


@pytest.mark.parametrize("key", EXTRA_ISO)
def test_extra_isotropic(key):
    import torch
    g = np.load(os.path.join(HERE, "golden", "iso_%s.npz" % key))
    seed = int(g["seed"])
    sc = _scatter(key, seed)
    worst = _xs_check(sc.crossSectionIsotropic(g["ekin"]), g["xs"])
    d_e = torch.from_numpy(g["ekin"]).cuda()
    nd = torch.zeros(d_e.numel(), dtype=torch.int32, device="cuda")
    sc.setRNGStream(seed, 0, 0)
    sc._L.ncb200_set_diagnostics_dev(sc._h, nd.data_ptr(), None)
    eo, mu = [t.cpu().numpy() for t in sc.sampleScatterIsotropic(d_e)]
    sc.checkDeviceErrors()
    ok = (np.abs(eo - g["ekin_out"]) <= 1e-10 * np.maximum(np.abs(g["ekin_out"]), 1e-300)) & (np.abs(mu - g["mu"]) <= 1e-10)
    flips = nd.cpu().numpy().astype(np.uint32) != g["ndraws"]
    print("%s: xs max rel %.2e; replay match %.6f, branch flips %d" % (key, worst, ok.mean(), flips.sum()))
    assert_replay((eo, mu), (g["ekin_out"], g["mu"]), nd.cpu().numpy(), g["ndraws"], key)


@pytest.mark.parametrize("key", ["D2O", "gas", "CH2_77K", "YAGCor"])
def test_extra_isotropic_staged_free_gas(key):
    # the staged free-gas kernels (used for batches >= 4e6 neutrons) forced on a small batch: same golden vectors
    import torch
    g = np.load(os.path.join(HERE, "golden", "iso_%s.npz" % key))
    seed = int(g["seed"])
    sc = _scatter(key, seed)
    d_e = torch.from_numpy(g["ekin"]).cuda()
    nd = torch.zeros(d_e.numel(), dtype=torch.int32, device="cuda")
    sc._L.ncb200_set_fg_staged_min(1)
    try:
        sc.setRNGStream(seed, 0, 0)
        sc._L.ncb200_set_diagnostics_dev(sc._h, nd.data_ptr(), None)
        eo, mu = [t.cpu().numpy() for t in sc.sampleScatterIsotropic(d_e)]
        sc.checkDeviceErrors()
    finally:
        sc._L.ncb200_set_fg_staged_min(4000000)
    ok = (np.abs(eo - g["ekin_out"]) <= 1e-10 * np.maximum(np.abs(g["ekin_out"]), 1e-300)) & (np.abs(mu - g["mu"]) <= 1e-10)
    assert_replay((eo, mu), (g["ekin_out"], g["mu"]), nd.cpu().numpy(), g["ndraws"], "%s staged free gas" % key)


@pytest.mark.parametrize("key", EXTRA_ANISO)
def test_extra_oriented(key):
    import torch
    g = np.load(os.path.join(HERE, "golden", "aniso_%s.npz" % key))
    seed = int(g["seed"])
    sc = _scatter(key, seed)
    assert sc.isOriented()
    worst = _xs_check(sc.crossSection(g["ekin"], (g["ux"], g["uy"], g["uz"])), g["xs"])
    d = [torch.from_numpy(np.ascontiguousarray(g[k])).cuda() for k in ("ekin", "ux", "uy", "uz")]
    nd = torch.zeros(d[0].numel(), dtype=torch.int32, device="cuda")
    sc.setRNGStream(seed, 0, 0)
    sc._L.ncb200_set_diagnostics_dev(sc._h, nd.data_ptr(), None)
    eo, (ox, oy, oz) = sc.sampleScatter(d[0], (d[1], d[2], d[3]))
    sc.checkDeviceErrors()
    eo, ox, oy, oz = [t.cpu().numpy() for t in (eo, ox, oy, oz)]
    ok = np.abs(eo - g["ekin_out"]) <= 1e-10 * np.maximum(np.abs(g["ekin_out"]), 1e-300)
    for a, b in ((ox, g["ox"]), (oy, g["oy"]), (oz, g["oz"])):
        ok &= np.abs(a - b) <= 1e-10
    flips = nd.cpu().numpy().astype(np.uint32) != g["ndraws"]
    print("%s: xs max rel %.2e; replay match %.6f, branch flips %d" % (key, worst, ok.mean(), flips.sum()))
    assert_replay((eo, ox, oy, oz), (g["ekin_out"], g["ox"], g["oy"], g["oz"]), nd.cpu().numpy(), g["ndraws"], key)
    # a batch large enough for the two-kernel candidate search (small batches use the combined kernel)
    n = 200000
    rep = [np.tile(g[k], n // g["ekin"].size + 1)[:n] for k in ("ekin", "ux", "uy", "uz")]
    xs_big = sc.crossSection(rep[0], (rep[1], rep[2], rep[3]))
    assert np.array_equal(xs_big[:g["ekin"].size], sc.crossSection(g["ekin"], (g["ux"], g["uy"], g["uz"])))


def test_layered_crystal_live_reference():
    """LCBragg through the C ABI against the live reference (oracle/_ref) on a batch large enough for the warp-per-neutron
    kernels, for the golden file's material and one with another mosaicity and layer axis."""
    import ncrystal_b200 as nc
    from _libs import RefDrv, have_refdrv, loguniform_energies, isotropic_directions
    from __graft_entry__ import EXTRA_CONFIGS
    if not have_refdrv():
        pytest.skip("needs oracle/_ref")
    for cfg, n in ((EXTRA_CONFIGS["PG"], 40000),
                   ("C_sg194_pyrolytic_graphite.ncmat;mos=0.5deg;dir1=@crys_hkl:0,0,1@lab:0,1,1;dir2=@crys_hkl:1,0,0@lab:1,0,0;lcaxis=0,0,1", 20000)):
        r = RefDrv(cfg)
        sc = nc.Scatter.fromBlob(r.compile(), seed=9)
        e = loguniform_energies(n, seed=51)
        ux, uy, uz = isotropic_directions(n, seed=52)
        xs = sc.crossSection(e, (ux, uy, uz))
        ref = r.xs(e, ux, uy, uz)
        rel = np.abs(xs - ref) / np.maximum(np.abs(ref), 1e-300)
        print("LC xs max rel %.2e" % rel.max())
        # Tolerance 1e-10 (not 1e-12) for this leaf, measured 1.2e-13 at mos=2deg and 3.1e-12 at mos=0.5deg: the ROI
        # limits come out of acos(), where CUDA's and glibc's results differ in the last bit; the mosaic Gaussian is
        # evaluated through cosines of small angles, which amplifies an absolute 1e-16 by 1/sigma_mos^2 (7e4 at 0.5deg).
        # The host build of the same device functions (glibc on both sides) is bit-identical to the reference
        # (test_cpu_extra_materials.py); the reference's own phi integration is accurate to 1e-3.
        assert rel.max() <= 1e-10
        sc.setRNGStream(9, 0, 0)
        eo, (ox, oy, oz) = sc.sampleScatter(e, (ux, uy, uz))
        a = r.sample(e, ux, uy, uz, seed=9, first_index=0)
        assert_replay((eo, ox, oy, oz), (a[0], a[1], a[2], a[3]), None, None, "LCBragg live")
