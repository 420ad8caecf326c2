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
EXTRA_ANISO = ["Cu_sc"]


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
    g = np.load(os.path.join(HERE, "golden", "aniso_%s.npz" % key))
    o = PortOracle(_blob(key)) if impl == "oracle" else HostSim(_blob(key))
    assert np.array_equal(o.xs(g["ekin"], g["ux"], g["uy"], g["uz"]), g["xs"])
    eo, ox, oy, oz, nd, er = o.sample(g["ekin"], g["ux"], g["uy"], g["uz"], seed=int(g["seed"]))[:6]
    assert np.array_equal(nd, g["ndraws"])
    for a, b in ((eo, g["ekin_out"]), (ox, g["ox"]), (oy, g["oy"]), (oz, g["oz"])):
        assert np.array_equal(a, b)
