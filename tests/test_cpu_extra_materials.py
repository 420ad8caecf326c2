"""CPU tier for the materials beyond BASELINE.json's five (__graft_entry__.EXTRA_CONFIGS: hexagonal powder, heavy
water from a file kernel, a multiphase mix, a pure gas mixture, a 77 K polymer, a Debye-model solid, a copper single
crystal with 0.3 deg mosaicity): the oracle's C restatement and the product's host-compiled device code against
golden vectors from the unmodified reference (tests/golden/make_golden.py), bit for bit."""
import os

import numpy as np
import pytest

from conftest import HERE
from _libs import HostSim
from _oracle_port import PortOracle

EXTRA_ISO = ["Be", "D2O", "AlBe", "gas", "CH2_77K", "V", "YAGCor"]
EXTRA_ANISO = ["Cu_sc", "PG"]


def _blob(key):
    from __graft_entry__ import EXTRA_CONFIGS
    from oracle_check import material_path
    p = material_path(EXTRA_CONFIGS[key])
    if not os.path.exists(p):
        pytest.skip("compiled material %s not present (build() makes it where the reference is available)" % p)
    return open(p, "rb").read()


@pytest.mark.parametrize("impl", ["oracle", "hostsim"])
@pytest.mark.parametrize("key", EXTRA_ISO)
def test_extra_isotropic_vs_golden(key, impl):
    g = np.load(os.path.join(HERE, "golden", "iso_%s.npz" % key))
    o = PortOracle(_blob(key)) if impl == "oracle" else HostSim(_blob(key))
    assert np.array_equal(o.xs_iso(g["ekin"]), g["xs"])
    eo, mu, nd, er = o.sample_iso(g["ekin"], seed=int(g["seed"]))[:4]
    assert np.array_equal(nd, g["ndraws"])
    assert np.array_equal(eo, g["ekin_out"]) and np.array_equal(mu, g["mu"])


@pytest.mark.parametrize("impl", ["oracle", "hostsim"])
@pytest.mark.parametrize("key", EXTRA_ANISO)
def test_extra_oriented_vs_golden(key, impl):
    if key == "PG" and impl == "oracle":
        pytest.skip("the C restatement has no layered-crystal leaf; the reference itself (oracle/_ref) is the oracle there")
    g = np.load(os.path.join(HERE, "golden", "aniso_%s.npz" % key))
    o = PortOracle(_blob(key)) if impl == "oracle" else HostSim(_blob(key))
    assert np.array_equal(o.xs(g["ekin"], g["ux"], g["uy"], g["uz"]), g["xs"])
    eo, ox, oy, oz, nd, er = o.sample(g["ekin"], g["ux"], g["uy"], g["uz"], seed=int(g["seed"]))[:6]
    assert np.array_equal(nd, g["ndraws"])
    for a, b in ((eo, g["ekin_out"]), (ox, g["ox"]), (oy, g["oy"]), (oz, g["oz"])):
        assert np.array_equal(a, b)


def test_material_compiler_refuses_leaves_outside_the_scope():
    # layered crystals are restated in their default mode (lcmode=0, LCHelper); the reference's two validation models
    # (lcmode != 0: an SCBragg rotated in steps / at random, NCLCRefModels.cc) are not: the reference-side material
    # compiler says so instead of producing tables the kernels would misread
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "bin", "ncb200_matcompile")
    if not os.path.exists(exe):
        pytest.skip("reference not built here")
    cfg = ("C_sg194_pyrolytic_graphite.ncmat;mos=1deg;dir1=@crys_hkl:0,0,1@lab:0,0,1;"
           "dir2=@crys_hkl:1,0,0@lab:1,0,0;lcaxis=0,0,1;lcmode=10")
    out = subprocess.run([exe, cfg, os.devnull], capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "lcmode!=0" in (out.stdout + out.stderr)


def test_layered_crystal_live_reference():
    """LCBragg (pyrolytic graphite, other mosaicity / orientation than the golden file): host build of the device
    functions against the live reference, bit for bit, incl. neutrons along the layer axis."""
    from _libs import RefDrv, have_refdrv, loguniform_energies, isotropic_directions
    if not have_refdrv():
        pytest.skip("reference not built here")
    cfg = ("C_sg194_pyrolytic_graphite.ncmat;mos=0.5deg;dir1=@crys_hkl:0,0,1@lab:0,1,1;"
           "dir2=@crys_hkl:1,0,0@lab:1,0,0;lcaxis=0,0,1")
    r = RefDrv(cfg)
    assert "LCBragg" in r.compnames()
    h = HostSim(r.compile())
    n = 3000
    e = loguniform_energies(n, seed=41)
    ux, uy, uz = isotropic_directions(n, seed=42)
    s = 1 / np.sqrt(2.0)
    ux[:20] = 0.0; uy[:20] = s; uz[:20] = s          # along the layer axis (lab (0,1,1))
    uy[10:20] = -s; uz[10:20] = -s
    assert np.array_equal(h.xs(e, ux, uy, uz), r.xs(e, ux, uy, uz))
    a = r.sample(e, ux, uy, uz, seed=11, first_index=5)
    b = h.sample(e, ux, uy, uz, seed=11, first_index=5)
    for k in range(5):
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k
    assert not np.asarray(b[5]).any()
