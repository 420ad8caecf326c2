"""GPU: absorption handles (ncrystal_create_absorption / ncrystal_cast_abs2proc) through the C ABI against the
reference's absorption cross sections (bit-exact: one sqrt and one division in fp64)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "abs_reference.npz"))


@pytest.mark.parametrize("key", ["Al", "CH2", "H2O", "YAG", "Ge"])
def test_absorption_xs_bit_exact(key):
    import ncrystal_b200 as nc
    from __graft_entry__ import CONFIGS
    a = nc.createAbsorption(CONFIGS[key])
    assert a.getName() == "AbsOOV" and a.isNonOriented()
    assert a.domain() == (0.0, float("inf"))
    xs = a.crossSectionIsotropic(GOLD["ekin"])
    assert np.array_equal(np.asarray(xs), GOLD[key])
    # oriented call on an isotropic process gives the same numbers (ncrystal_crosssection)
    assert float(a.crossSection(0.0253, (0.0, 0.0, 1.0))) == GOLD[key][5]


def test_absorption_cannot_be_sampled_or_cast_to_scatter():
    import ctypes as C
    import ncrystal_b200 as nc
    from ncrystal_b200 import _lib
    from __graft_entry__ import CONFIGS
    a = nc.createAbsorption(CONFIGS["Al"])
    L = _lib.lib()
    assert not L.ncrystal_cast_proc2scat(a._p).internal
    assert L.ncrystal_cast_proc2abs(a._p).internal
    s = nc.Scatter(CONFIGS["Al"], seed=1)
    assert not L.ncrystal_cast_proc2abs(s._p).internal
    fake = _lib.ncrystal_scatter_t(a._h.internal)
    e, mu = C.c_double(), C.c_double()
    L.ncrystal_sethaltonerror(0)
    L.ncrystal_setquietonerror(1)
    L.ncrystal_samplescatterisotropic(fake, 0.0253, C.byref(e), C.byref(mu))
    assert L.ncrystal_error() and e.value == -1.0
    L.ncrystal_clearerror()
