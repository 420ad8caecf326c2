"""Pins the oracle's plain-C restatement (oracle/oracle_*.c) against the golden vectors generated
from the unmodified reference (tests/golden/*.npz, made with oracle/_ref by make_golden.py):
cross sections and replayed scatter outcomes must agree BIT FOR BIT on this machine (same libm, no
FMA contraction), for all five configs incl. the oriented one.  Also cross-checks its sequential
S(alpha,beta) table builder against the reference's internal sampler tables when oracle/_ref exists."""
import os

import numpy as np
import pytest

from conftest import CONFIG_KEYS_ISO, HERE, golden
from _libs import RefDrv, have_refdrv
from _oracle_port import PortOracle


def _blob(cfg):
    from oracle_check import material_path
    p = material_path(cfg)
    if os.path.exists(p):
        return open(p, "rb").read()
    if have_refdrv():
        return RefDrv(cfg).compile()
    pytest.skip("compiled material %s not available" % p)


@pytest.mark.parametrize("key", CONFIG_KEYS_ISO)
def test_port_vs_golden_iso(key, configs):
    g = golden(key)
    o = PortOracle(_blob(configs[key]))
    assert np.array_equal(o.xs_iso(g["ekin"]), g["xs"])
    eo, mu, nd, er = o.sample_iso(g["ekin"], seed=int(g["seed"]))
    assert np.array_equal(nd, g["ndraws"])
    assert np.array_equal(eo, g["ekin_out"]) and np.array_equal(mu, g["mu"])
    assert not er.any()


def test_port_vs_golden_oriented(configs):
    g = np.load(os.path.join(HERE, "golden", "aniso_Ge.npz"))
    o = PortOracle(_blob(configs["Ge"]))
    assert np.array_equal(o.xs(g["ekin"], g["ux"], g["uy"], g["uz"]), g["xs"])
    eo, ox, oy, oz, nd, er = o.sample(g["ekin"], g["ux"], g["uy"], g["uz"], seed=int(g["seed"]))
    assert np.array_equal(nd, g["ndraws"])
    for a, b in ((eo, g["ekin_out"]), (ox, g["ox"]), (oy, g["oy"]), (oz, g["oz"])):
        assert np.array_equal(a, b)


@pytest.mark.skipif(not have_refdrv(), reason="oracle/_ref not built")
def test_port_table_builder_vs_reference_internals(configs):
    r = RefDrv(configs["H2O"])
    o = PortOracle(r.compile())
    for iE in range(0, 300, 11):
        try:
            a = r.sab_sampler_dump(0, iE, 1000)
        except RuntimeError:
            break
        b = o.sab_sampler_dump(0, iE, 1000)
        assert a["n"] == b["n"] and a["ibeta_off"] == b["ibeta_off"] and a["first_bin"] == b["first_bin"]
        for k in ("x", "pdf", "cdf", "infos"):
            assert np.array_equal(a[k], b[k]), (iE, k)
    # the integrator's total xs per energy point reproduces the reference's xs grid (from the blob)
    from test_cpu_blob import sab_grids
    egrid, xsgrid = sab_grids(r.compile(), 0)
    assert np.allclose(o.sab_xscheck(0, egrid.size), xsgrid, rtol=1e-14, atol=0)
