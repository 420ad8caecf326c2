"""VDOS -> S(alpha,beta) expansion, CPU tier: the host orchestration of csrc/ncb_vdos.h with the
NCB_HD device functions of csrc/ncb_vdos_dev.cuh run in plain loops (tests/hostsim, TEST-ONLY) must reproduce the
reference's tables BIT FOR BIT -- against the committed golden results of the reference's own C-API
(ncrystal_raw_vdos2kernel / ncrystal_raw_vdos2gn, tests/golden/vdos_reference.npz) and, when the compiled reference is
present, against the live reference."""
import numpy as np
import pytest

import _vdos
from _libs import HostSim, RefDrv, have_refdrv, loguniform_energies


@pytest.fixture(scope="module")
def golden():
    return _vdos.load_golden()


@pytest.mark.parametrize("name", [n for n in _vdos.CASES if _vdos.CASES[n][6] is None])
def test_host_build_reproduces_reference_tables(golden, name):
    _vdos.check_against_golden(_vdos.HostSimVdos(), golden, name)


@pytest.mark.skipif(not _vdos.have_reference(), reason="compiled reference (oracle/_ref) not present")
def test_golden_file_matches_live_reference(golden):
    api = _vdos.reference_api()
    for name in ("Be_lux1_emax", "Be_lux2_weights", "irregular_lux2", "coarse_lux1"):
        _vdos.check_against_golden(api, golden, name)


@pytest.mark.skipif(not _vdos.have_reference(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("temperature,lux", [(20.0, 2), (1200.0, 1)])
def test_host_build_vs_live_reference_other_temperatures(golden, temperature, lux):
    """A temperature sweep of one curve: low T (few orders, detailed balance factors underflow) and high T."""
    egrid, density = golden["in_Be_egrid"], golden["in_Be_density"]
    sigma, mass, _ = golden["in_Be_meta"]
    ref = _vdos.reference_api().kernel(egrid, density, sigma, mass, temperature, lux)
    got = _vdos.HostSimVdos().kernel(egrid, density, sigma, mass, temperature, lux)
    for a, b in zip(ref[:3], got[:3]):
        assert np.array_equal(a, b)
    assert ref[3] == got[3]


def test_invalid_input_is_refused():
    h = _vdos.HostSimVdos()
    d = np.array([0.1, 0.5, 1.0, 0.5, 0.2, 0.1])
    with pytest.raises(RuntimeError):
        h.kernel(np.array([1e-7, 0.03]), d, 5.0, 12.0, 300.0, 1)        # grid starts below 1e-5 eV
    with pytest.raises(RuntimeError):
        h.kernel(np.array([0.01, 0.03]), d, 5.0, 12.0, 300.0, 7)        # vdoslux out of range
    with pytest.raises(RuntimeError):
        h.kernel(np.array([0.01, 0.02, 0.03]), d, 5.0, 12.0, 300.0, 1)  # egrid length neither 2 nor len(density)


@pytest.mark.skipif(not have_refdrv(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("cfg", ["Be_sg194.ncmat;vdoslux=1", "solid::V/6.1gcm3/TDebye390K", "Polyethylene_CH2.ncmat;vdoslux=2"])
def test_material_with_vdos_leaves_equals_material_with_expanded_leaves(cfg):
    """A compiled material whose S(alpha,beta) leaves arrive as phonon densities of states (NCB_KIND_SABVDOS, the
    material compiler's --vdos form) is expanded at load time and must then behave exactly like the material compiled
    with the reference's expanded tables: same energy grids, cross sections to 1e-14 (the grid cross sections are
    integrated by the library instead of copied), replayed samples identical."""
    r = RefDrv(cfg)
    plain, vd = r.compile(), r.compile(flags=1)
    assert len(vd) < len(plain) / 5
    a, b = HostSim(plain), HostSim(vd)
    assert [a.component_kind(c) for c in range(a.ncomp)] == [b.component_kind(c) for c in range(b.ncomp)]
    e = loguniform_energies(4000, seed=5)
    xa, xb = a.xs_iso(e), b.xs_iso(e)
    assert np.max(np.abs(xa - xb) / xa) < 1e-13
    ea, ma, na = a.sample_iso(e, seed=11)[:3]
    eb, mb, nb = b.sample_iso(e, seed=11)[:3]
    same = (np.abs(ea - eb) <= 1e-10 * np.abs(ea)) & (np.abs(ma - mb) <= 1e-10) & (na == nb)
    assert same.all(), "%d of %d replayed samples differ" % ((~same).sum(), e.size)


def test_twiddle_table_by_doubling_equals_reference_recipe():
    """csrc/ncb_vdos.h builds the FFT twiddle table of size 2N from the table of size N (one pass); it must equal the
    table the reference's recursive recipe gives (NCFastConvolve.cc:183-215,466-564), entry for entry."""
    L = HostSim.lib()
    for log2size in (1, 2, 5, 12, 16, 18):
        assert L.hostsim_vdos_twiddle_check(log2size) == 0


@pytest.mark.skipif(not have_refdrv(), reason="compiled reference (oracle/_ref) not present")
def test_corrupt_vdos_leaf_is_refused_not_read_out_of_bounds():
    """The array count and the parameters inside a VDOS leaf come from the (untrusted) buffer: a truncated or
    inconsistent leaf must raise, never read past the payload (same rule as for the other leaf kinds, ncb_loader.h)."""
    import struct
    blob = bytearray(RefDrv("Be_sg194.ncmat;vdoslux=0").compile(flags=1))
    ncomp = struct.unpack_from("<I", blob, 12)[0]
    hdr_fixed = struct.calcsize("<QIIIIQdddddd176s")
    comp_sz = struct.calcsize("<IIdddQQ")
    leaf = None
    for i in range(ncomp):
        kind, _r, _s, _lo, _hi, off, nbytes = struct.unpack_from("<IIdddQQ", blob, hdr_fixed + i * comp_sz)
        if kind == 8:
            leaf = (i, off, nbytes)
    assert leaf is not None
    i, off, nbytes = leaf
    o_vdoslux = off + 14 * 8          # 14 doubles, then vdoslux, negrid, ndensity, reserved
    HostSim(bytes(blob))              # the intact material loads

    def broken(mutate):
        b = bytearray(blob)
        mutate(b)
        with pytest.raises(RuntimeError):
            HostSim(bytes(b))

    broken(lambda b: struct.pack_into("<Q", b, o_vdoslux + 16, 10**9))                   # ndensity beyond the payload
    broken(lambda b: struct.pack_into("<Q", b, o_vdoslux + 16, 1))                       # too few density points
    broken(lambda b: struct.pack_into("<Q", b, o_vdoslux, 9))                            # vdoslux out of range
    broken(lambda b: struct.pack_into("<Q", b, o_vdoslux + 8, 3))                        # energy grid too short
    broken(lambda b: struct.pack_into("<Q", b, hdr_fixed + i * comp_sz + 40, 64))        # payload shorter than its header
    broken(lambda b: struct.pack_into("<d", b, off + 9 * 8, 1e-9))                       # emin below 1e-5 eV
    broken(lambda b: struct.pack_into("<d", b, off + 8, -5.0))                           # negative temperature
