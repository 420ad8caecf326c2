"""VDOS -> S(alpha,beta) expansion on the device, through the reference's own C entry points
ncrystal_raw_vdos2kernel / ncrystal_raw_vdos2gn as exported by libncrystal_b200.so.  Parity bar: BIT-EXACT -- the
tables must equal the reference's (committed golden results of the reference's C-API, and the live reference when
oracle/_ref is present): the device does only +,-,*,/ and sqrt in the reference's order, every transcendental set-up
value is computed on the host with the same libm."""
import ctypes as C
import os

import numpy as np
import pytest

import _vdos
from _parity import assert_replay

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    return _vdos.load_golden()


@pytest.fixture(scope="module")
def product():
    from ncrystal_b200 import _lib
    return _vdos.RawVdosAPI(_lib.lib())


@pytest.mark.parametrize("name", list(_vdos.CASES))
def test_device_expansion_reproduces_reference_tables(product, golden, name):
    from ncrystal_b200 import _lib
    n0 = _lib.lib().ncb200_vdos_expansion_count()
    _vdos.check_against_golden(product, golden, name)
    assert _lib.lib().ncb200_vdos_expansion_count() == n0 + 1      # (the CUDA path ran: there is no other)


@pytest.mark.skipif(not _vdos.have_reference(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("curve,temperature,lux,emax", [("Al", 77.0, 3, 0.0), ("Al", 600.0, 4, 0.0), ("CH2_H", 293.15, 2, 2.0),
                                                        ("Be", 5.0, 1, 0.0), ("twopoint", 2000.0, 2, 0.0)])
def test_device_expansion_vs_live_reference(product, golden, curve, temperature, lux, emax):
    """Temperature / vdoslux / target-Emax sweep against the live reference (4 to ~600 phonon orders)."""
    egrid, density = golden["in_%s_egrid" % curve], golden["in_%s_density" % curve]
    sigma, mass = (float(golden["in_%s_meta" % curve][0]), float(golden["in_%s_meta" % curve][1])) if curve != "twopoint" else (2.2, 55.8)
    ref = _vdos.reference_api().kernel(egrid, density, sigma, mass, temperature, lux, emax)
    got = product.kernel(egrid, density, sigma, mass, temperature, lux, emax)
    for a, b in zip(ref[:3], got[:3]):
        assert a.shape == b.shape and np.array_equal(a, b)
    assert ref[3] == got[3]
    for order in (1, 2, 5, 40):
        r, g = _vdos.reference_api().gn(egrid, density, sigma, mass, temperature, order), product.gn(egrid, density, sigma, mass, temperature, order)
        assert r[0] == g[0] and r[1] == g[1] and np.array_equal(r[2], g[2]), order


def test_python_mirror_and_error_convention(golden):
    """extractKnl / extractGn (names of the reference's vdos.py) and the C-API error convention on bad input."""
    import ncrystal_b200 as nc
    from ncrystal_b200 import vdos
    egrid, density, sigma, mass, T, lux, emax, weight, order = _vdos.case_inputs(golden, "Be_lux2_weights")
    k = vdos.extractKnl((egrid, density), mass, T, vdoslux=lux, scatxs=sigma, order_weight_fct=weight)
    assert k["suggested_emax"] is None and _vdos.sha(k["sab"]) == str(golden["out_Be_lux2_weights_sab_sha"])
    (xmin, xmax), gn = vdos.extractGn((egrid, density), order, mass, T, scatxs=sigma, expand_egrid=False)
    assert (xmin, xmax) == tuple(golden["out_Be_lux2_weights_gn_range"]) and _vdos.sha(gn) == str(golden["out_Be_lux2_weights_gn_sha"])
    # obsolete spelling ncrystal_raw_vdos2knl (ncrystal.h:897-907): same table, no Emax arguments
    L = nc._lib.lib()
    dp = C.POINTER(C.c_double)
    na, nb, pa, pb, ps = C.c_uint(0), C.c_uint(0), dp(), dp(), dp()
    L.ncrystal_raw_vdos2knl(egrid.ctypes.data_as(dp), density.ctypes.data_as(dp), egrid.size, density.size, sigma, mass, T, 1,
                            None, C.byref(na), C.byref(nb), C.byref(pa), C.byref(pb), C.byref(ps))
    k1 = vdos.extractKnl((egrid, density), mass, T, vdoslux=1, scatxs=sigma)
    assert na.value == k1["alpha"].size and nb.value == k1["beta"].size
    assert np.array_equal(np.ctypeslib.as_array(ps, (na.value * nb.value,)), k1["sab"])
    for p_ in (pa, pb, ps):
        L.ncrystal_dealloc_doubleptr(p_)
    with pytest.raises(nc.NCBadInput):
        vdos.extractKnl((np.array([1e-7, 0.03]), density), mass, T)
    with pytest.raises(nc.NCBadInput):
        vdos.extractKnl((egrid, density[:3]), mass, T)
    with pytest.raises(nc.NCException):
        vdos.extractKnl((egrid, density), mass, T, vdoslux=9)


def _data_blob(cfg, suffix):
    from ncrystal_b200 import _lib
    buf = C.create_string_buffer(512)
    _lib.lib().ncb200_cfg_to_filestem(cfg.encode(), buf, 512)
    path = os.path.join(_lib.DATA_DIR, buf.value.decode() + suffix)
    if not os.path.exists(path):
        pytest.skip("compiled material %s not built" % os.path.basename(path))
    with open(path, "rb") as f:
        return f.read()


@pytest.mark.parametrize("key", ["Al", "CH2", "YAG", "V"])
def test_material_with_vdos_leaves(key, configs):
    """Compiled materials whose S(alpha,beta) leaves arrive as phonon densities of states (material compiler --vdos,
    ncb_blob.h NCB_KIND_SABVDOS): the library expands them on the device when the material is created and must then
    return what the material with the reference's expanded tables returns -- identical energy grids, cross sections to
    1e-12 (north_star), identical replayed samples -- and what the reference itself returns (golden vectors)."""
    import ncrystal_b200 as nc
    from conftest import golden as golden_iso
    from __graft_entry__ import EXTRA_CONFIGS
    from _libs import loguniform_energies
    cfg = configs.get(key) or EXTRA_CONFIGS[key]
    plain, vd = _data_blob(cfg, ".ncb"), _data_blob(cfg, ".vdos.ncb")
    assert len(vd) < len(plain) / 5
    n0 = nc._lib.lib().ncb200_vdos_expansion_count()
    a, b = nc.Scatter.fromBlob(plain, seed=3), nc.Scatter.fromBlob(vd, seed=3)
    assert nc._lib.lib().ncb200_vdos_expansion_count() > n0
    e = loguniform_energies(200000, seed=77)
    xa, xb = a.crossSectionIsotropic(e), b.crossSectionIsotropic(e)
    assert np.max(np.abs(xa - xb) / xa) < 1e-12
    a.setRNGStream(9, 0, 0); b.setRNGStream(9, 0, 0)
    assert_replay(b.sampleScatterIsotropic(e), a.sampleScatterIsotropic(e), what=key)
    g = golden_iso(key)
    xs = b.crossSectionIsotropic(g["ekin"])
    ok = np.isfinite(g["xs"]) & (g["xs"] != 0)
    assert np.max(np.abs(xs[ok] - g["xs"][ok]) / np.abs(g["xs"][ok])) <= 1e-12
    b.setRNGStream(int(g["seed"]), 0, 0)
    assert_replay(b.sampleScatterIsotropic(g["ekin"]), (g["ekin_out"], g["mu"]), what=key + " vs golden")


@pytest.mark.skipif(not _vdos.have_reference(), reason="compiled reference (oracle/_ref) not present")
def test_data_library_sweep_subset(product):
    """Every 7th phonon density of states of the reference's embedded data library (tests/vdos_sweep.py runs all 211;
    results under profiles/): device table == reference table, bit for bit."""
    from _libs import RefDrv, _d
    ref = _vdos.reference_api()
    n, strs = C.c_uint(0), C.POINTER(C.c_char_p)()
    ref.L.ncrystal_get_file_list.argtypes = [C.POINTER(C.c_uint), C.POINTER(C.POINTER(C.c_char_p))]
    ref.L.ncrystal_get_file_list(C.byref(n), C.byref(strs))
    names = sorted({strs[i].decode() for i in range(0, n.value, 4) if strs[i].decode().endswith(".ncmat")})
    L = RefDrv.lib()
    L.refdrv_vdos_data.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
    done = 0
    for name in names[::7]:
        meta, dens = np.zeros(5), np.zeros(200000)
        m = L.refdrv_vdos_data(name.encode(), 0, _d(meta), _d(dens), dens.size)
        if m <= 0:
            continue
        args = (meta[:2].copy(), dens[:m].copy(), float(meta[4]), float(meta[3]), float(meta[2]), 1)
        r, g = ref.kernel(*args), product.kernel(*args)
        assert all(a.shape == b.shape and np.array_equal(a, b) for a, b in zip(r[:3], g[:3])) and r[3] == g[3], name
        done += 1
    assert done >= 10


def test_create_scatter_falls_back_to_the_vdos_form(configs, tmp_path):
    """ncrystal_create_scatter(cfg) looks for <stem>.ncb and then for <stem>.vdos.ncb: a data directory that holds only
    the density-of-states form of a material (43 KB instead of 2.6 MB for Al) serves the cfg string."""
    import shutil
    import ncrystal_b200 as nc
    from ncrystal_b200 import _lib
    from _libs import loguniform_energies
    L = _lib.lib()
    cfg = configs["Al"]
    alias = cfg + ";vdoslux=3"                       # same material, a spelling no compiled file exists for
    buf = C.create_string_buffer(512)
    L.ncb200_cfg_to_filestem(cfg.encode(), buf, 512)
    src = os.path.join(_lib.DATA_DIR, buf.value.decode() + ".vdos.ncb")
    if not os.path.exists(src):
        pytest.skip("vdos form not built")
    L.ncb200_cfg_to_filestem(alias.encode(), buf, 512)
    shutil.copy(src, tmp_path / (buf.value.decode() + ".vdos.ncb"))
    with pytest.raises(nc.NCFileNotFound):
        nc.Scatter(alias, seed=1)
    e = loguniform_energies(5000, seed=8)
    ref = nc.Scatter(cfg, seed=1).crossSectionIsotropic(e)
    n0 = L.ncb200_vdos_expansion_count()
    L.ncb200_set_data_path((str(tmp_path) + ":" + _lib.DATA_DIR).encode())
    try:
        xs = nc.Scatter(alias, seed=1).crossSectionIsotropic(e)
    finally:
        L.ncb200_set_data_path(os.environ.get("NCB200_DATA_PATH", _lib.DATA_DIR).encode())
    assert L.ncb200_vdos_expansion_count() == n0 + 1
    assert np.max(np.abs(xs - ref) / ref) < 1e-12
