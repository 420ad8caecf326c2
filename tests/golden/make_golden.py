"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run here (the reference sources are not on the GPU box):

    python tests/golden/make_golden.py

Per isotropic config: N energies (log-uniform 1e-5..10 eV + edge cases), total and per-component
cross sections, and scatter outcomes under replay of the per-neutron Philox streams
(seed=GOLDEN_SEED, index=i) through the reference's RNG interface, incl. the number of
uniforms each neutron consumed.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _libs import RefDrv, loguniform_energies, isotropic_directions  # noqa: E402
from __graft_entry__ import CONFIGS, EXTRA_CONFIGS  # noqa: E402

GOLDEN_SEED = 20261017
N = 4000


def energies(n):
    e = loguniform_energies(n, seed=777)
    # edge cases the reference tests exercise: below/above SAB grid, Bragg thresholds, tiny and large
    e[:12] = [1e-5, 1e-3, 0.0253, 1.0, 10.0, 1e-9, 1e-7, 4.9999, 5.0001, 0.0037, 0.00375, 100.0]
    return e


def oriented_inputs(n):
    """Isotropic directions + log-uniform energies, with a quarter of the neutrons placed close to
    Bragg conditions of the Ge config (dir ~ (0,1,1)/sqrt2 at 1.54 Aa, within ~3e-4 rad) so that the
    mosaic Gaussian, both circle-integral branches and the scatter generation are exercised."""
    e = loguniform_energies(n, seed=778)
    ux, uy, uz = isotropic_directions(n, seed=779)
    k = n // 4
    rng = np.random.Generator(np.random.Philox(key=780))
    e0 = 0.081804209605330899 / 1.54 ** 2
    ux[:k] = rng.normal(0, 3e-4, k)
    uy[:k] = 1 / np.sqrt(2) + rng.normal(0, 3e-4, k)
    uz[:k] = 1 / np.sqrt(2) + rng.normal(0, 3e-4, k)
    e[:k] = e0 * (1 + rng.normal(0, 1e-3, k))
    # exact known-answer points of the reference's own tests (_testimpl.py:253-254)
    e[k:k + 3] = e0
    ux[k:k + 3] = [0, 1, 0]
    uy[k:k + 3] = [1, 1, 0]
    uz[k:k + 3] = [1, 0, 1]
    return e, ux, uy, uz


def oriented_inputs_lc(n):
    """Layered crystal (lcaxis = lab z): isotropic directions + log-uniform energies, plus the degenerate cases of the
    ROI search -- neutrons along +-lcaxis (all crystallite rotations equivalent), almost along it, perpendicular to
    it, non-normalised directions -- and energies at / just around the Bragg threshold."""
    e = loguniform_energies(n, seed=781)
    ux, uy, uz = isotropic_directions(n, seed=782)
    k = 64
    e[:k] = np.geomspace(1.5e-3, 8.0, k)
    ux[:k] = 0.0; uy[:k] = 0.0; uz[:k] = 1.0
    uz[k // 2:k] = -1.0
    e[k:2 * k] = np.geomspace(1.5e-3, 8.0, k)
    ux[k:2 * k] = 1e-11; uy[k:2 * k] = 0.0; uz[k:2 * k] = 1.0
    e[2 * k:3 * k] = np.geomspace(1.5e-3, 8.0, k)
    ux[2 * k:3 * k] = 0.6; uy[2 * k:3 * k] = 0.8; uz[2 * k:3 * k] = 0.0
    ux[3 * k:4 * k] *= 3.0; uy[3 * k:4 * k] *= 3.0; uz[3 * k:4 * k] *= 3.0
    e[4 * k:4 * k + 6] = [0.0018163, 0.0018164, 0.00181636, 0.002, 1e-5, 100.0]
    return e, ux, uy, uz


def main():
    only = sys.argv[1:]
    for key, cfg in list(CONFIGS.items()) + list(EXTRA_CONFIGS.items()):
        if only and key not in only:
            continue
        r = RefDrv(cfg)
        if RefDrv.lib().refdrv_isoriented(r.h):
            e, ux, uy, uz = oriented_inputs_lc(N) if "LCBragg" in r.compnames() else oriented_inputs(N)
            xs = r.xs(e, ux, uy, uz)
            eo, ox, oy, oz, nd = r.sample(e, ux, uy, uz, seed=GOLDEN_SEED, first_index=0)
            out = os.path.join(HERE, "aniso_%s.npz" % key)
            np.savez_compressed(out, cfg=cfg, seed=GOLDEN_SEED, ekin=e, ux=ux, uy=uy, uz=uz, xs=xs,
                                comp_names=np.array(r.compnames()), ekin_out=eo, ox=ox, oy=oy, oz=oz, ndraws=nd)
            print(out, os.path.getsize(out), "bytes; mean draws %.2f; n(xs>10 barn)=%d" % (nd.mean(), (xs > 10).sum()))
            continue
        e = energies(N)
        xs = r.xs_iso(e)
        xsc = r.xs_iso_components(e)
        eo, mu, nd = r.sample_iso(e, seed=GOLDEN_SEED, first_index=0)
        out = os.path.join(HERE, "iso_%s.npz" % key)
        np.savez_compressed(out, cfg=cfg, seed=GOLDEN_SEED, ekin=e, xs=xs, xs_components=xsc,
                            comp_names=np.array(r.compnames()), ekin_out=eo, mu=mu, ndraws=nd)
        print(out, os.path.getsize(out), "bytes; mean draws %.2f" % nd.mean())


if __name__ == "__main__":
    main()
