"""CPU: the oracle's 1/v absorption (oracle_api.c::orc_abs_xs_many, constant from the compiled material) against the
reference's absorption cross sections (tests/golden/abs_reference.npz, made by make_golden_abs.py)."""
import ctypes as C
import os

import numpy as np
import pytest

from _mmc import cached_oracle

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "abs_reference.npz"))


@pytest.mark.parametrize("key", ["Al", "CH2", "H2O", "YAG", "Ge"])
def test_oracle_absorption_bit_exact(key):
    from _oracle_port import lib
    o, hdr = cached_oracle(key)
    L = lib()
    dp = C.POINTER(C.c_double)
    L.orc_abs_xs_many.argtypes = [C.c_void_p, dp, C.c_uint64, dp]
    ekin = np.ascontiguousarray(GOLD["ekin"])
    out = np.empty_like(ekin)
    L.orc_abs_xs_many(o.h, ekin.ctypes.data_as(dp), len(ekin), out.ctypes.data_as(dp))
    assert np.array_equal(out, GOLD[key])
    assert np.isinf(out[0]) and hdr["abs_c"] > 0
