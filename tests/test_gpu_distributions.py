"""Independent-stream statistics: with the product's own per-neutron Philox streams (no replay) the
distributions of mu and of the energy transfer must be statistically compatible with the
reference's (its own C-API, builtin xoroshiro RNG) -- chi2 two-sample test on histograms and a KS
test, north_star's third correctness leg, on north_star's 1e8 samples per config (45 s for the four isotropic
configs on the GPU box; NCB200_DIST_N overrides)."""
import os

import numpy as np
import pytest

from conftest import CONFIG_KEYS_ISO
from _libs import RefDrv, have_refdrv, loguniform_energies

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_refdrv(), reason="needs oracle/_ref for the reference arm")]

N = int(os.environ.get("NCB200_DIST_N", "100000000"))
P_MIN = 1e-4


def _chi2_two_sample(a, b):
    from scipy.stats import chi2
    m = (a + b) >= 25
    a, b = a[m].astype(float), b[m].astype(float)
    k1, k2 = np.sqrt(b.sum() / a.sum()), np.sqrt(a.sum() / b.sum())
    stat = (((k1 * a - k2 * b) ** 2) / (a + b)).sum()
    dof = m.sum() - 1
    return stat, dof, chi2.sf(stat, dof)


@pytest.mark.parametrize("key", CONFIG_KEYS_ISO)
def test_mu_and_deltae_distributions(key, configs):
    import ncrystal_b200 as nc
    from scipy.stats import ks_2samp
    cfg = configs[key]
    tot = {"mu": None, "de": None}
    ks_mu_a, ks_mu_b = [], []
    chunk = min(N, 10_000_000)
    done = 0
    sc = nc.Scatter(cfg, seed=987654321)
    mu_edges = np.linspace(-1, 1, 201)
    # energy transfer: ratio E_out/E_in on a log grid (+ exact-elastic bin)
    r_edges = np.concatenate([[-np.inf], np.linspace(-6, 6, 241), [np.inf]])
    while done < N:
        m = min(chunk, N - done)
        e = loguniform_energies(m, seed=1000 + done)
        eo_g, mu_g = sc.sampleScatterIsotropic(e)
        eo_r, mu_r = RefDrv.capi_sample_iso(cfg, e)
        assert np.all(eo_g >= 0) and np.all(np.abs(mu_g) <= 1)
        for name, (xg, xr) in {"mu": (mu_g, mu_r), "de": (None, None)}.items():
            if name == "mu":
                hg, hr = np.histogram(xg, mu_edges)[0], np.histogram(xr, mu_edges)[0]
            else:
                el_g, el_r = eo_g == e, eo_r == e
                with np.errstate(divide="ignore"):
                    lg = np.log10(np.maximum(eo_g[~el_g], 1e-300) / e[~el_g])
                    lr = np.log10(np.maximum(eo_r[~el_r], 1e-300) / e[~el_r])
                hg = np.concatenate([[el_g.sum()], np.histogram(lg, r_edges)[0]])
                hr = np.concatenate([[el_r.sum()], np.histogram(lr, r_edges)[0]])
            tot[name] = (hg, hr) if tot[name] is None else (tot[name][0] + hg, tot[name][1] + hr)
        if len(ks_mu_a) * chunk < 2_000_000:
            ks_mu_a.append(mu_g[:500000]); ks_mu_b.append(mu_r[:500000])
        done += m
    for name in ("mu", "de"):
        stat, dof, p = _chi2_two_sample(*tot[name])
        print("%s %s: chi2/dof = %.1f/%d  p = %.3g  (N = %d)" % (key, name, stat, dof, p, N))
        assert p > P_MIN, (key, name, stat, dof, p)
    ks = ks_2samp(np.concatenate(ks_mu_a), np.concatenate(ks_mu_b))
    print("%s KS(mu): D = %.2e p = %.3g" % (key, ks.statistic, ks.pvalue))
    assert ks.pvalue > P_MIN


def test_oriented_distributions(configs):
    """Same for the oriented single crystal: scattering-angle cosine, energy transfer and the azimuth of the outgoing
    direction around the incoming one, against the reference's per-neutron ncrystal_samplescatter (2e7 samples)."""
    import ncrystal_b200 as nc
    from _libs import isotropic_directions
    cfg = configs["Ge"]
    n_tot = int(os.environ.get("NCB200_DIST_N_ORIENTED", "20000000"))
    sc = nc.Scatter(cfg, seed=24680)
    mu_edges = np.linspace(-1, 1, 201)
    r_edges = np.concatenate([[-np.inf], np.linspace(-6, 6, 241), [np.inf]])
    phi_edges = np.linspace(-np.pi, np.pi, 73)
    tot = {}
    done = 0
    while done < n_tot:
        m = min(5_000_000, n_tot - done)
        e = loguniform_energies(m, seed=77 + done)
        ux, uy, uz = isotropic_directions(m, seed=99 + done)
        res = {"g": sc.sampleScatter(e, (ux, uy, uz)), "r": RefDrv.capi_sample_aniso(cfg, e, ux, uy, uz)}
        for who, (eo, (ox, oy, oz)) in (("g", (res["g"][0], res["g"][1])), ("r", (res["r"][0], res["r"][1:]))):
            mu = np.clip(ux * ox + uy * oy + uz * oz, -1, 1)
            el = eo == e
            with np.errstate(divide="ignore"):
                lr = np.log10(np.maximum(eo[~el], 1e-300) / e[~el])
            # azimuth of the outgoing direction in a frame fixed to the LAB z axis (the crystal is oriented)
            phi = np.arctan2(oy, ox)
            h = (np.histogram(mu, mu_edges)[0], np.concatenate([[el.sum()], np.histogram(lr, r_edges)[0]]),
                 np.histogram(phi, phi_edges)[0])
            tot[who] = h if who not in tot else tuple(a + b for a, b in zip(tot[who], h))
        done += m
    for k, name in enumerate(("mu", "de", "phi_lab")):
        stat, dof, p = _chi2_two_sample(tot["g"][k], tot["r"][k])
        print("Ge %s: chi2/dof = %.1f/%d  p = %.3g  (N = %d)" % (name, stat, dof, p, n_tot))
        assert p > P_MIN, (name, stat, dof, p)
