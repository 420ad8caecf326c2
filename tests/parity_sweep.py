"""One-off large replay sweep (evidence, not a pytest): product on the GPU vs the live reference (oracle/_ref) on the
same per-neutron Philox streams.  usage: python tests/parity_sweep.py [n_iso] [n_oriented]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ncrystal_b200 as nc
from __graft_entry__ import CONFIGS, EXTRA_CONFIGS
from _libs import RefDrv, loguniform_energies, isotropic_directions
n_iso = int(float(sys.argv[1])) if len(sys.argv) > 1 else 5_000_000
n_or = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
for key, cfg in list(CONFIGS.items()) + list(EXTRA_CONFIGS.items()):
    if only and key not in only:
        continue
    try:
        r = RefDrv(cfg); sc = nc.Scatter(cfg, seed=1)
    except Exception as e:  # noqa: BLE001
        print(json.dumps(dict(config=key, skipped=str(e)[:80]))); continue
    t0 = time.time()
    if sc.isOriented():
        n = n_or
        e = loguniform_energies(n, seed=606); ux, uy, uz = isotropic_directions(n, seed=607)
        xs, xr = sc.crossSection(e, (ux, uy, uz)), r.xs(e, ux, uy, uz)
        sc.setRNGStream(31337, 0, 0)
        eo, (ox, oy, oz) = sc.sampleScatter(e, (ux, uy, uz))
        eo_r, ox_r, oy_r, oz_r, nd = r.sample(e, ux, uy, uz, seed=31337, first_index=0)
        ok = np.abs(eo - eo_r) <= 1e-10 * np.maximum(np.abs(eo_r), 1e-300)
        for a, b in ((ox, ox_r), (oy, oy_r), (oz, oz_r)): ok &= np.abs(a - b) <= 1e-10
        exact = (eo == eo_r) & (ox == ox_r) & (oy == oy_r) & (oz == oz_r)
    else:
        n = n_iso
        e = loguniform_energies(n, seed=505)
        xs, xr = sc.crossSectionIsotropic(e), r.xs_iso(e)
        sc.setRNGStream(31337, 0, 0)
        eo, mu = sc.sampleScatterIsotropic(e)
        eo_r, mu_r, nd = r.sample_iso(e, seed=31337, first_index=0)
        ok = (np.abs(eo - eo_r) <= 1e-10 * np.maximum(np.abs(eo_r), 1e-300)) & (np.abs(mu - mu_r) <= 1e-10)
        exact = (eo == eo_r) & (mu == mu_r)
    nz = xr != 0
    rel = float(np.max(np.abs(xs[nz] - xr[nz]) / np.abs(xr[nz])))
    print(json.dumps(dict(config=key, n=n, xs_max_rel_err=rel, xs_zero_pattern_equal=bool(np.array_equal(xs == 0, xr == 0)),
                          replay_match_1e10=float(ok.mean()), mismatches=int((~ok).sum()), bit_identical=float(exact.mean()),
                          mean_draws=float(nd.mean()), seconds=round(time.time() - t0, 1))), flush=True)
