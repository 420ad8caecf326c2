"""CPU tests (no GPU):
 * the Philox stream restated in oracle/philox_ref.h == the product's ncb_rng.cuh (host build),
   pinned by the Random123 known-answer vectors;
 * oracle/_ref (unmodified reference built from /root/reference) reproduces the reference's own golden
   vectors for this path (pins the oracle);
 * the host compilation of the product's device functions (tests/hostsim) reproduces the committed
   golden vectors bit for bit (kernel logic, table builder), for every config.
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import CONFIG_KEYS_ISO, HERE, ROOT, golden
from _libs import HostSim, RefDrv, have_refdrv, loguniform_energies

needs_ref = pytest.mark.skipif(not have_refdrv(), reason="oracle/_ref not built (needs /root/reference)")


def _blob(cfg):
    from oracle_check import material_path
    p = material_path(cfg)
    if os.path.exists(p):
        return open(p, "rb").read()
    if have_refdrv():
        return RefDrv(cfg).compile()
    pytest.skip("compiled material %s not available" % p)


def test_philox_known_answers(tmp_path):
    src = tmp_path / "kat.c"
    src.write_text(r'''
#include "philox_ref.h"
#include <stdio.h>
int main(){ uint32_t o[4];
 { uint32_t c[4]={0,0,0,0},k[2]={0,0}; ncb_philox4x32_10(c,k,o); printf("%08x %08x %08x %08x\n",o[0],o[1],o[2],o[3]); }
 { uint32_t c[4]={~0u,~0u,~0u,~0u},k[2]={~0u,~0u}; ncb_philox4x32_10(c,k,o); printf("%08x %08x %08x %08x\n",o[0],o[1],o[2],o[3]); }
 { uint32_t c[4]={0x243f6a88,0x85a308d3,0x13198a2e,0x03707344},k[2]={0xa4093822,0x299f31d0}; ncb_philox4x32_10(c,k,o); printf("%08x %08x %08x %08x\n",o[0],o[1],o[2],o[3]); }
 ncb_stream_t s; ncb_stream_init(&s, 0x123456789abcdefULL, 77); for(int i=0;i<7;i++) printf("%.17g\n", ncb_stream_next(&s));
 return 0; }''')
    exe = tmp_path / "kat"
    subprocess.check_call(["gcc", "-O1", "-I", os.path.join(ROOT, "oracle"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split("\n")
    # Random123 kat_vectors (philox4x32 10 rounds)
    assert out[0] == "6627e8d5 e169c58d bc57ac4c 9b00dbd8"
    assert out[1] == "408f276d 41c83b0e a20bc7c6 6d5451fd"
    assert out[2] == "d16cfe09 94fdcceb 5001e420 24126ea1"
    ref = np.array([float(x) for x in out[3:10]])
    # the product's generator (ncb_rng.cuh compiled for the host) draws the same uniforms
    hs = HostSim.__new__(HostSim)
    got = HostSim.uniforms(hs, 0x123456789abcdef, 77, 7)
    assert np.array_equal(ref, got)
    assert np.all((got > 0) & (got <= 1))


@needs_ref
def test_reference_golden_vectors_pin_the_oracle():
    """Golden values held by the reference's own tests for this path."""
    def wl2ekin(wl):
        return 0.081804209605330899 / wl ** 2
    # ncrystal_python/src/NCrystal/_testimpl.py:119-133 (PowderBragg only)
    r = RefDrv("stdlib::Al_sg225.ncmat;dcutoff=1.4;incoh_elas=0;inelas=0")
    assert r.compnames() == ["PowderBragg"]
    assert r.xs_iso([wl2ekin(4.0)])[0] == pytest.approx(1.632435821586171, rel=1e-6, abs=1e-6)
    assert r.xs_iso([wl2ekin(5.0)])[0] == 0.0
    # _testimpl.py:148-158 (Ni composite)
    r = RefDrv("stdlib::Ni_sg225.ncmat;dcutoff=0.6;vdoslux=2")
    assert r.xs_iso([wl2ekin(1.2)])[0] == pytest.approx(16.76474410391571, rel=1e-6)
    assert r.xs_iso([wl2ekin(5.0)])[0] == pytest.approx(5.958467463288343, rel=1e-6)
    # _testimpl.py:245-254 (Ge single crystal)
    r = RefDrv("stdlib::Ge_sg227.ncmat;dcutoff=0.5;mos=40.0arcsec;dir1=@crys_hkl:5,1,1@lab:0,0,1;dir2=@crys_hkl:0,-1,1@lab:0,1,0")
    e = np.array([wl2ekin(1.54)] * 2)
    xs = r.xs(e, [0., 1.], [1., 1.], [1., 0.])
    assert xs[0] == pytest.approx(591.0263476502018, rel=1e-6)
    assert xs[1] == pytest.approx(1.667600586136298, rel=1e-6)
    # ncrystal_core/app_test/main.cc:28-58: 60 Al cross sections, |delta| < 0.01
    refxs = [1.39667, 1.39437, 1.38778, 1.37679, 1.36133, 1.342, 1.32158, 1.30743, 1.30778, 1.32854, 1.37329, 1.37085,
             1.38283, 1.37629, 1.31355, 1.37167, 1.35602, 1.34481, 1.44825, 1.2065, 1.29534, 1.32806, 1.42857, 1.53696,
             1.50743, 1.11265, 1.1864, 1.26397, 1.34521, 1.00269, 1.06099, 1.12169, 1.18474, 1.2501, 1.31772, 1.38759,
             1.45968, 1.53396, 1.61043, 1.68906, 1.76985, 1.18806, 1.2403, 1.29387, 1.34877, 1.40498, 1.46251, 0.143847,
             0.144769, 0.145723, 0.146725, 0.147773, 0.148863, 0.14999, 0.151153, 0.152347, 0.153569, 0.154815,
             0.156124, 0.157444]
    r = RefDrv("stdlib::Al_sg225.ncmat;dcutoff=0.5;temp=25C")
    wl = np.arange(60) * 0.1
    with np.errstate(divide="ignore"):
        e = np.where(wl > 0, 0.081804209605330899 / np.maximum(wl, 1e-300) ** 2, np.inf)
    xs = r.xs_iso(e[1:])
    assert np.all(np.abs(xs - np.array(refxs[1:])) < 0.01)
    # the same cfg through the product's device code (host build)
    h = HostSim(r.compile())
    assert np.array_equal(h.xs_iso(e[1:]), xs)
    # app_test/main.cc:52-121: component fractions of sampled scatterings at 3.5 Aa within 4 sigma
    n = 400000
    ek = np.full(n, wl2ekin(3.5))
    eo, mu, nd, er = h.sample_iso(ek, seed=424242)
    elastic = eo == ek
    ang = np.degrees(np.arccos(mu))
    n_inel = (~elastic).sum()
    n1 = (elastic & (np.abs(ang - 96.9) < 0.05)).sum()
    n2 = (elastic & (np.abs(ang - 119.6) < 0.05)).sum()
    n_incel = elastic.sum() - n1 - n2
    frac_bragg, frac_incelas, frac_inel, rel1 = 0.899618, 0.00554598, 0.0948364, 0.6119430177909819
    for cnt, exp in ((n_inel, frac_inel), (n_incel, frac_incelas), (n1, frac_bragg * rel1), (n2, frac_bragg * (1 - rel1))):
        x, dev = cnt / n, np.sqrt(cnt) / n
        assert abs(x - exp) <= 4.0 * dev, (cnt, exp)


@needs_ref
def test_survey_known_answers():
    """cross sections generated from the reference at E = 1e-5, 1e-3, 0.0253, 1, 10 eV."""
    from __graft_entry__ import CONFIGS
    known = {
        "Al": [1.7346733918473856, 0.20149926201309404, 1.4466377182937247, 1.395229354732082, 1.3964556425986621],
        "CH2": [233.23206140559097, 73.505929073044328, 39.502578053352217, 15.958236243758972, 15.30214878367225],
        "H2O": [352.20714153502928, 81.259681625063905, 35.921633489406631, 15.60960755157306, 14.953329355273882],
        "YAG": [2.53212368256112, 0.37187594948733754, 3.5280112030621913, 3.7244145049714072, 3.725130579773011],
    }
    e = np.array([1e-5, 1e-3, 0.0253, 1.0, 10.0])
    for k, v in known.items():
        xs = RefDrv(CONFIGS[k]).xs_iso(e)
        assert np.allclose(xs, v, rtol=1e-13, atol=0), k


@pytest.mark.parametrize("key", CONFIG_KEYS_ISO)
def test_hostsim_reproduces_golden(key, configs):
    g = golden(key)
    h = HostSim(_blob(configs[key]))
    xs = h.xs_iso(g["ekin"])
    assert np.array_equal(xs, g["xs"]), "host build of the device xs code must be bit-exact (same libm, no FMA)"
    assert np.array_equal(h.xs_iso_components(g["ekin"]), g["xs_components"])
    eo, mu, nd, er = h.sample_iso(g["ekin"], seed=int(g["seed"]))
    assert np.array_equal(nd, g["ndraws"])
    assert np.array_equal(eo, g["ekin_out"]) and np.array_equal(mu, g["mu"])
    assert not np.any(er & ~16)


def test_hostsim_reproduces_golden_oriented(configs):
    g = np.load(os.path.join(HERE, "golden", "aniso_Ge.npz"))
    h = HostSim(_blob(configs["Ge"]))
    xs = h.xs(g["ekin"], g["ux"], g["uy"], g["uz"])
    assert np.array_equal(xs, g["xs"])
    eo, ox, oy, oz, nd, er = h.sample(g["ekin"], g["ux"], g["uy"], g["uz"], seed=int(g["seed"]))
    assert np.array_equal(nd, g["ndraws"])
    for a, b in ((eo, g["ekin_out"]), (ox, g["ox"]), (oy, g["oy"]), (oz, g["oz"])):
        assert np.array_equal(a, b)
    # outgoing directions are unit vectors; SCBragg and the elastic leaves conserve energy
    nrm = ox * ox + oy * oy + oz * oz
    assert np.all(np.abs(nrm - 1) < 1e-12)


@needs_ref
@pytest.mark.parametrize("key", CONFIG_KEYS_ISO)
def test_native_sab_table_builder_matches_reference_internals(key, configs):
    """The product derives the S(alpha,beta) sampler tables itself; they must equal the reference's
    private SABSamplerAtE_Alg1 state (dumped from oracle/_ref) entry by entry."""
    r = RefDrv(configs[key])
    h = HostSim(r.compile())
    names = r.compnames()
    checked = 0
    for c, nm in enumerate(names):
        if nm != "SABScatter":
            continue
        for iE in list(range(0, 300, 7)) + [299]:
            try:
                a = r.sab_sampler_dump(c, iE, 1000)
            except RuntimeError:
                break
            b = h.sab_sampler_dump(c, iE, 1000)
            assert a["n"] == b["n"] and a["ibeta_off"] == b["ibeta_off"] and a["first_bin"] == b["first_bin"]
            for k in ("x", "pdf", "cdf", "infos"):
                assert np.array_equal(a[k], b[k]), (key, c, iE, k)
            checked += 1
    assert checked > 40


@needs_ref
def test_edge_energies_vs_reference(configs):
    """Edge cases: far below / above the tabulated grids, exactly on grid points and thresholds."""
    r = RefDrv(configs["Al"])
    h = HostSim(r.compile())
    from test_cpu_blob import sab_grids, parse_header
    blob = r.compile()
    egrid, _ = sab_grids(blob, 2)
    e = np.concatenate([[1e-12, 1e-9, 1e-7, 50.0, 1e3, 1e6], egrid[[0, 1, 150, 298, 299]],
                        np.nextafter(egrid[[0, 299]], 0), np.nextafter(egrid[[0, 299]], 1e9)])
    assert np.array_equal(r.xs_iso(e), h.xs_iso(e))
    a = r.sample_iso(e, seed=99)
    b = h.sample_iso(e, seed=99)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


@pytest.mark.parametrize("key", ["Al", "CH2", "H2O", "YAG"])
def test_short_chain_sampler_equals_plain_sampler(key, configs):
    """The short-chain variants of the S(alpha,beta) table sampler (SabBPoint / SabHead / SabTail / SabPoint gather
    records, log guide: what k_sample_sab_refill runs) consume the same uniforms and give bit-identical outcomes as
    the plain restatement, which the goldens pin against the reference -- for every S(alpha,beta) leaf."""
    h = HostSim(_blob(configs[key]))
    ekin = np.concatenate([loguniform_energies(30000, seed=5), [1e-9, 1e-7, 4.99999, 5.0, 7.5]])
    nleaf = 0
    for c in range(h.ncomp):
        if h.component_kind(c) != 3:
            continue
        nleaf += 1
        a = h.sample_iso(ekin, seed=77, leaf=c)
        b = h.sample_sab_staged(ekin, c, seed=77)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    assert nleaf >= 1


@pytest.mark.parametrize("key", ["Al", "YAG", "H2O", "Ge"])
def test_energy_key_lut_gives_the_exact_upper_bound(key, configs):
    """The energy-key luts that replace the whole-table bisections of the cross-section path (PowderBragg 2dE table,
    S(alpha,beta) energy grid) must give std::upper_bound exactly -- at, just below and just above every table value,
    for random energies, and for the inputs outside the table (negative, zero, subnormal, inf, NaN)."""
    import ctypes as C
    h = HostSim(_blob(configs[key]))
    L = HostSim.lib()
    L.hostsim_keyed_upper_bound.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_uint64,
                                            C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.hostsim_table_values.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int]
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    nlut = 0
    for c in range(h.ncomp):
        if h.component_kind(c) not in (1, 3):
            continue
        tab = np.empty(70000)
        n = L.hostsim_table_values(h.h, c, tab.ctypes.data_as(dp), tab.size)
        tab = tab[:n]
        v = np.concatenate([tab, np.nextafter(tab, 0.0), np.nextafter(tab, np.inf), loguniform_energies(20000, seed=9, lo=1e-7, hi=1e3),
                            [0.0, -0.0, -1.0, 5e-324, 1e-300, 1e300, np.inf, -np.inf, np.nan, tab[0] * 0.5, tab[-1] * 2]])
        a, b = np.empty(v.size, dtype=np.int32), np.empty(v.size, dtype=np.int32)
        nk = L.hostsim_keyed_upper_bound(h.h, c, v.ctypes.data_as(dp), v.size, a.ctypes.data_as(ip), b.ctypes.data_as(ip))
        assert nk >= 0
        nlut += nk > 0
        assert np.array_equal(a, b)
        fin = np.isfinite(v)
        assert np.array_equal(b[fin], np.searchsorted(tab, v[fin], side="right"))
    assert nlut >= 1


@pytest.mark.parametrize("key", ["Al", "H2O", "CH2"])
def test_sab_energy_grid_determined_from_the_kernel_alone(key):
    """SABIntegrator::setupEnergyGrid / determineEMin / determineEMax (NCSABIntegrator.cc:147-283) restated
    (csrc/ncb_sabgrid.h): a compiled material stripped of the reference's energy grids, grid cross sections and
    extension constants must come out identical to the unstripped one -- grid points bit for bit, cross sections and
    replayed samples likewise (host build: same libm as the reference)."""
    from _libs import HostSim, strip_sab_energy_grids, loguniform_energies
    from oracle_check import material_path
    from __graft_entry__ import CONFIGS
    blob = open(material_path(CONFIGS[key]), "rb").read()
    # (the water file fixes Emax itself: "egrid 4.02", data/LiquidWaterH2O_T293.6K.ncmat:95; Emin is determined)
    a, b = HostSim(blob), HostSim(strip_sab_energy_grids(blob, emax_request=4.02 if key == "H2O" else 0.0))
    e = loguniform_energies(3000, seed=5)
    e[:6] = [1e-9, 1e-7, 4.999, 5.001, 9.0, 100.0]
    assert np.array_equal(a.xs_iso_components(e), b.xs_iso_components(e))
    ra, rb = a.sample_iso(e, seed=3), b.sample_iso(e, seed=3)
    for x, y in zip(ra, rb):
        assert np.array_equal(np.asarray(x), np.asarray(y))


@pytest.mark.parametrize("key", ["H2O", "D2O", "V"])
def test_sab_energy_grid_fully_automatic_vs_live_reference(key):
    """No Emax request and (water kernels) no suggested Emax in the table: determineEMax's walk down from the kinematic
    limit decides the upper end.  The reference's SABIntegrator is run here on the same kernel with no "egrid" request."""
    from _libs import HostSim, RefDrv, have_refdrv, strip_sab_energy_grids
    from __graft_entry__ import CONFIGS, EXTRA_CONFIGS
    if not have_refdrv():
        pytest.skip("reference not built here")
    cfg = CONFIGS.get(key) or EXTRA_CONFIGS[key]
    r = RefDrv(cfg)
    h = HostSim(strip_sab_energy_grids(r.compile()))
    names = r.compnames()
    nchecked = 0
    for c, nm in enumerate(names):
        if nm != "SABScatter":
            continue
        eg_ref, xs_ref = r.sab_auto_egrid(c)
        eg = h.sab_egrid(c, 1000)
        assert np.array_equal(eg, eg_ref), (key, c, eg[[0, -1]], eg_ref[[0, -1]])
        assert np.array_equal(h.sab_xscheck(c, 1000), xs_ref)
        nchecked += 1
    assert nchecked >= 1
