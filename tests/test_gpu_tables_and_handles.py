"""GPU tests of what the device builds and keeps: the S(alpha,beta) sampler tables against the reference's private
tables, the derived gather tables / guides, several materials alive together (shared-memory attribute regression),
pageable caller buffers, and the large replay sweep."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import CONFIG_KEYS_ISO
from _parity import assert_replay

pytestmark = pytest.mark.gpu
_dp = C.POINTER(C.c_double)


def _dump(sc, c, iE, nbeta=1000):
    from _libs import _sab_dump
    return _sab_dump(lambda h, *a: sc._L.ncb200_sab_sampler_dump(sc._p, *a), None, c, iE, nbeta)


@pytest.mark.parametrize("key", CONFIG_KEYS_ISO + ["Ge"])
def test_device_built_sampler_tables_vs_reference(key, configs):
    """Entry-by-entry diff of the S(alpha,beta) sampler tables the device kernels build (ncb200_sab_sampler_dump)
    against the reference's private SABSamplerAtE_Alg1 state (refdrv_sab_sampler_dump), for every 7th energy point of
    every S(alpha,beta) leaf.  Integers and grid values must be identical; values that went through the device's
    exp/log (cumulative integrals, tail points) to 1e-12."""
    import ncrystal_b200 as nc
    from _libs import RefDrv, have_refdrv
    if not have_refdrv():
        pytest.skip("needs oracle/_ref")
    r = RefDrv(configs[key])
    sc = nc.Scatter(configs[key])
    checked = 0
    for c, nm in enumerate(r.compnames()):
        if nm != "SABScatter":
            continue
        for iE in list(range(0, 300, 7)) + [299]:
            try:
                a = r.sab_sampler_dump(c, iE, 1000)
            except RuntimeError:
                break
            b = _dump(sc, c, iE)
            assert a["n"] == b["n"] and a["ibeta_off"] == b["ibeta_off"], (key, c, iE)
            if a["n"] == 0:
                continue
            assert np.array_equal(a["x"][1:], b["x"][1:])                      # beta grid values
            assert abs(a["x"][0] - b["x"][0]) <= 1e-13 * abs(a["x"][0]) and abs(a["first_bin"] - b["first_bin"]) <= 1e-13 * abs(a["first_bin"])
            for k in ("pdf", "cdf"):
                assert np.all(np.abs(a[k] - b[k]) <= 1e-12 * np.maximum(np.abs(a[k]), 1e-300)), (key, c, iE, k)
            ia, ib = a["infos"], b["infos"]
            assert np.array_equal(ia[:, [3, 7]], ib[:, [3, 7]])                # alpha grid indices
            for col in (0, 1, 4, 5, 8, 9):                                     # alpha, S values, probabilities
                assert np.all(np.abs(ia[:, col] - ib[:, col]) <= 1e-12 * np.maximum(np.abs(ia[:, col]), 1e-290)), (key, c, iE, col)
            for col in (2, 6):                                                 # log S of the tail points
                fin = np.isfinite(ia[:, col])
                assert np.array_equal(fin, np.isfinite(ib[:, col]))
                assert np.all(np.abs(ia[fin, col] - ib[fin, col]) <= 1e-11 * np.maximum(np.abs(ia[fin, col]), 1.0))
            checked += 1
    assert checked > 40


@pytest.mark.parametrize("key", CONFIG_KEYS_ISO)
def test_gather_tables_and_guides_selfcheck(key, configs):
    import ncrystal_b200 as nc
    sc = nc.Scatter(configs[key])
    n = 0
    for c, (kind, _scale) in enumerate(sc.components()):
        if kind == 3:
            assert sc._L.ncb200_sab_selfcheck(sc._p, c) == 0
            n += 1
    assert n >= 1


def test_materials_alive_together(configs):
    """the dynamic shared-memory limit of a kernel was set from the material loaded LAST, so a
    handle created earlier with larger tables failed to launch afterwards.  Ge (73/77 KB of staged single-crystal
    tables) then Cu (54/68 KB), then YAG and Al; every handle is used after all were created, the first one last."""
    import ncrystal_b200 as nc
    from __graft_entry__ import EXTRA_CONFIGS
    from oracle_check import material_path
    from _libs import loguniform_energies, isotropic_directions
    cfgs = [configs["Ge"]]
    if os.path.exists(material_path(EXTRA_CONFIGS["Cu_sc"])):
        cfgs.append(EXTRA_CONFIGS["Cu_sc"])
    cfgs += [configs["YAG"], configs["Al"]]
    n = 40000
    e = loguniform_energies(n, seed=3)
    d = isotropic_directions(n, seed=4)
    first = {}
    handles = []
    for cfg in cfgs:                      # results right after creation
        sc = nc.Scatter(cfg, seed=12)
        handles.append(sc)
        sc.setRNGStream(12, 0, 0)
        if sc.isOriented():
            first[cfg] = (sc.crossSection(e, d), sc.sampleScatter(e, d))
        else:
            first[cfg] = (sc.crossSectionIsotropic(e), sc.sampleScatterIsotropic(e))
    for sc, cfg in reversed(list(zip(handles, cfgs))):     # ... and again with every other material loaded
        sc.setRNGStream(12, 0, 0)
        if sc.isOriented():
            xs, (eo, dirs) = sc.crossSection(e, d), sc.sampleScatter(e, d)
            assert np.array_equal(xs, first[cfg][0]) and np.array_equal(eo, first[cfg][1][0])
            for a, b in zip(dirs, first[cfg][1][1]):
                assert np.array_equal(a, b)
        else:
            xs, (eo, mu) = sc.crossSectionIsotropic(e), sc.sampleScatterIsotropic(e)
            assert np.array_equal(xs, first[cfg][0]) and np.array_equal(eo, first[cfg][1][0]) and np.array_equal(mu, first[cfg][1][1])
        assert np.all(xs >= 0) and not np.any(eo == -1.0)


def test_pageable_and_pinned_host_buffers_agree(configs):
    """malloc'd (pageable) caller arrays go through the pinned bounce ring of the host pipeline, pinned ones are
    copied directly: same results, for a call long enough to wrap the ring several times and with a ragged tail."""
    import torch
    import ncrystal_b200 as nc
    from _libs import loguniform_energies
    n = 9_000_001
    e = loguniform_energies(n, seed=21)
    sc = nc.Scatter(configs["Al"], seed=8)
    L = sc._L
    pin = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(4)]
    pin[0].copy_(torch.from_numpy(e))
    pp = [C.cast(t.data_ptr(), _dp) for t in pin]
    pg = [e] + [np.empty(n) for _ in range(3)]
    gp = [a.ctypes.data_as(_dp) for a in pg]
    for p in (pp, gp):
        sc.setRNGStream(8, 0, 0)
        L.ncrystal_crosssection_nonoriented_many(sc._p, p[0], n, 1, p[1])
        L.ncrystal_samplescatterisotropic_many(sc._h, p[0], n, 1, p[2], p[3])
        nc.core._check_error()
    for k in (1, 2, 3):
        assert np.array_equal(pin[k].numpy(), pg[k])
    # mixed: pageable input, pinned outputs
    sc.setRNGStream(8, 0, 0)
    L.ncrystal_samplescatterisotropic_many(sc._h, gp[0], n, 1, pp[2], pp[3])
    assert np.array_equal(pin[3].numpy(), pg[3])


@pytest.mark.parametrize("key", CONFIG_KEYS_ISO + ["Ge"])
def test_large_replay_sweep_vs_live_reference(key, configs):
    """tests/parity_sweep.py as a test: 5e6 neutrons per isotropic config (1e6 for the single crystal) against the
    live reference on the same per-neutron streams: zero mismatches beyond 1e-10 (or draw-count flips <= 1e-6)."""
    import ncrystal_b200 as nc
    from _libs import RefDrv, have_refdrv, loguniform_energies, isotropic_directions
    if not have_refdrv():
        pytest.skip("needs oracle/_ref")
    r = RefDrv(configs[key])
    sc = nc.Scatter(configs[key], seed=1)
    if sc.isOriented():
        n = 1_000_000
        e = loguniform_energies(n, seed=606)
        d = isotropic_directions(n, seed=607)
        xs, xr = sc.crossSection(e, d), r.xs(e, *d)
        sc.setRNGStream(31337, 0, 0)
        eo, (ox, oy, oz) = sc.sampleScatter(e, d)
        ref = r.sample(e, *d, seed=31337, first_index=0)
        outs, refs = (eo, ox, oy, oz), ref[:4]
    else:
        n = 5_000_000
        e = loguniform_energies(n, seed=505)
        xs, xr = sc.crossSectionIsotropic(e), r.xs_iso(e)
        sc.setRNGStream(31337, 0, 0)
        outs = sc.sampleScatterIsotropic(e)
        refs = r.sample_iso(e, seed=31337, first_index=0)[:2]
    nz = xr != 0
    assert np.array_equal(xs == 0, xr == 0)
    assert np.max(np.abs(xs[nz] - xr[nz]) / np.abs(xr[nz])) <= 1e-12
    assert_replay(outs, refs, None, None, "%s sweep" % key)


def test_plain_c_caller_uses_all_visible_gpus(tmp_path, configs):
    """tests/capi_multigpu.c: one process, host arrays, ncb200_set_devices(all) + the reference's *_many calls; the
    results must be bit-identical to the one-device results, the NCCL-merged tally equal to the one-device tally.
    (On a one-GPU box this degenerates to one device; gpurun --gpus 2/8 evidence is under profiles/.)"""
    import subprocess
    from conftest import ROOT
    exe = os.path.join(str(tmp_path), "capi_multigpu")
    libdir = os.path.join(ROOT, "ncrystal_b200", "lib")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-D_POSIX_C_SOURCE=200809L", "-pedantic", "-Wall", "-Werror", "-I",
                           os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "capi_multigpu.c"), "-o", exe,
                           "-L", libdir, "-lncrystal_b200", "-lm", "-Wl,-rpath," + libdir])
    r = subprocess.run([exe, configs["Al"], "0", "3000001"], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert '"identical_to_one_device": true' in r.stdout


@pytest.mark.parametrize("key", ["Al", "H2O", "CH2"])
def test_sab_energy_grid_determined_on_the_device(key):
    """A compiled material stripped of the reference's S(alpha,beta) energy grids / grid cross sections / extension
    constants: the library determines them itself (csrc/ncb_sabgrid.h; all probe energies of determineEMin /
    determineEMax integrated in one pass of the table-build kernels) and must behave like the unstripped material."""
    import ncrystal_b200 as nc
    from _libs import strip_sab_energy_grids, loguniform_energies
    from _parity import assert_replay
    from oracle_check import material_path
    from __graft_entry__ import CONFIGS
    blob = open(material_path(CONFIGS[key]), "rb").read()
    a = nc.Scatter.fromBlob(blob, seed=4)
    b = nc.Scatter.fromBlob(strip_sab_energy_grids(blob, emax_request=4.02 if key == "H2O" else 0.0), seed=4)
    e = loguniform_energies(200000, seed=6)
    e[:6] = [1e-9, 1e-7, 4.999, 5.001, 9.0, 100.0]
    xa, xb = a.crossSectionIsotropic(e), b.crossSectionIsotropic(e)
    rel = np.abs(xa - xb) / np.abs(xa)
    print("%s: xs max rel %.2e between given and device-determined energy grid" % (key, rel.max()))
    assert rel.max() <= 1e-12
    a.setRNGStream(4, 0, 0); b.setRNGStream(4, 0, 0)
    ra, rb = a.sampleScatterIsotropic(e), b.sampleScatterIsotropic(e)
    assert_replay(rb, ra, None, None, "%s auto energy grid" % key)
