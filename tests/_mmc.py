"""Test helpers for the device-resident transport step (ncb200_minimc_run): scenario definitions mirroring the
reference's MiniMC unit tests (tests/scripts/mmc_al.py, mmc_circ.py, mmc_scge.py via
tests/pypath/NCTestUtils/minimc_ref.py), the oracle binding (oracle/oracle_mmc.c), the reference runner
(ncrystal_jsonquery of oracle/_ref/lib/libNCrystal.so) and the reference's histogram compatibility test
(Hist1D.check_compat, ncrystal_python/src/NCrystal/hist.py:810-905)."""
import ctypes as C
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

TALLY_TYPES = ["theta", "mu", "nscat", "nscat_uw", "w", "e", "l", "de", "q"]
CLASS_NAMES = ["NOSCAT", "SINGLESCAT_ELAS", "SINGLESCAT_INELAS", "MULTISCAT_PUREELAS", "MULTISCAT_OTHER"]
NCLASS, NSTAT = 5, 5
WL2EKIN = 0.081804209605330899


class OrcMmcCfg(C.Structure):
    _fields_ = [("geom_kind", C.c_int), ("ga", C.c_double), ("gb", C.c_double), ("gc", C.c_double),
                ("src_kind", C.c_int), ("pos", C.c_double * 3), ("dir", C.c_double * 3), ("radius", C.c_double),
                ("emode", C.c_int), ("e0", C.c_double), ("e1", C.c_double), ("weight", C.c_double),
                ("roul_psurv", C.c_double), ("roul_wthr", C.c_double), ("roul_nscat", C.c_int),
                ("nscatlimit", C.c_int), ("ignore_miss", C.c_int), ("include_abs", C.c_int), ("seed", C.c_uint64)]


class Scenario:
    """One transport set-up, expressed once and rendered (a) as the reference's cfg strings, (b) as the oracle's
    numeric struct."""

    def __init__(self, key, material, geom, src, energy, n, tallies=(("theta", 90, 0.0, 180.0),), seed=0,
                 nscatlimit=None, ignore_miss=False, absorption=True, roulette=None, pos=(0.0, 0.0, 0.0),
                 direction=(0.0, 0.0, 1.0), radius=0.0):
        self.key, self.material, self.geom, self.src, self.energy, self.n = key, material, geom, src, energy, int(n)
        self.tallies, self.seed, self.nscatlimit = list(tallies), seed, nscatlimit
        self.ignore_miss, self.absorption, self.roulette = ignore_miss, absorption, roulette
        self.pos, self.direction, self.radius = tuple(pos), tuple(direction), radius

    # ---- cfg strings (reference vocabulary)
    @property
    def geomcfg(self):
        name, pars = self.geom
        return name + "".join(";%s=%.17g" % kv for kv in pars.items())

    def srccfg(self, n=None):
        kind, val = self.energy
        if kind == "thermal":
            es = "ekin=thermal:%.17gK" % val
        elif isinstance(val, tuple) and len(val) == 3:      # (mean, rms, "lognormal")
            es = "%s=%.17g+-%.17g" % (kind, val[0], val[1])
        elif isinstance(val, tuple):
            es = "%s=%.17g-%.17g" % (kind, val[0], val[1])
        else:
            es = "%s=%.17g" % (kind, val)
        nn = self.n if n is None else int(n)
        if self.src == "isotropic":
            s = "isotropic;%s;x=%.17g;y=%.17g;z=%.17g;n=%d" % ((es,) + self.pos + (nn,))
            if self.radius:
                s += ";r=%.17g" % self.radius
            return s
        s = "%s;%s;x=%.17g;y=%.17g;z=%.17g;ux=%.17g;uy=%.17g;uz=%.17g;n=%d" % (
            (self.src, es) + self.pos + self.direction + (nn,))
        if self.src == "circular":
            s += ";r=%.17g" % self.radius
        return s

    def enginecfg(self, extra=""):
        s = "tally=%s;tallybins=%s;seed=%d" % (",".join(t[0] for t in self.tallies),
                                               ",".join("%s:%d:%.17g:%.17g" % t for t in self.tallies), self.seed)
        if self.nscatlimit is not None:
            s += ";nscatlimit=%d" % self.nscatlimit
        if self.ignore_miss:
            s += ";ignoremiss=1"
        if not self.absorption:
            s += ";absorption=0"
        if self.roulette:
            s += ";roulette=%.17g,%.17g,%d" % self.roulette
        return s + extra

    # ---- oracle struct
    def orc_cfg(self):
        c = OrcMmcCfg()
        name, p = self.geom
        c.geom_kind = {"sphere": 1, "slab": 2, "box": 3, "cyl": 4}[name]
        if name == "sphere":
            c.ga = p["r"]
        elif name == "slab":
            c.gc = p["dz"]
        elif name == "box":
            c.ga, c.gb, c.gc = p["dx"], p["dy"], p["dz"]
        else:
            c.ga, c.gb = p["r"], p.get("dy", 0.0)
        c.src_kind = {"constant": 1, "circular": 2, "isotropic": 3}[self.src]
        for k in range(3):
            c.pos[k] = self.pos[k]
            c.dir[k] = self.direction[k]
        c.radius = self.radius
        kind, val = self.energy
        if kind == "thermal":
            c.emode, c.e0, c.e1 = 5, val, 0.0
        elif isinstance(val, tuple) and len(val) == 3:
            c.emode = 3 if kind == "ekin" else 4
            c.e0, c.e1 = val[0], val[1]
        elif isinstance(val, tuple):
            c.emode = 1 if kind == "ekin" else 2
            c.e0, c.e1 = val
        else:
            c.emode = 0
            c.e0 = c.e1 = val if kind == "ekin" else WL2EKIN / (val * val)
        c.weight = 1.0
        r = self.roulette or (0.1, 1e-2, 2)
        c.roul_psurv, c.roul_wthr, c.roul_nscat = r
        c.nscatlimit = -1 if self.nscatlimit is None else self.nscatlimit
        c.ignore_miss = int(self.ignore_miss)
        c.include_abs = int(self.absorption)
        c.seed = self.seed
        return c


def hist_doubles(nbins):
    return NCLASS * (2 * (nbins + 2) + NSTAT)


def split_device_layout(buf, tallies):
    """device/oracle tally layout -> {name: dict(content[class,nb+2], errsq[class,nb+2], stats[class,5])}"""
    out, off = {}, 0
    for name, nb, lo, hi in tallies:
        nb2 = nb + 2
        g = np.asarray(buf[off:off + hist_doubles(nb)])
        out[name] = dict(content=g[:NCLASS * nb2].reshape(NCLASS, nb2).copy(),
                         errsq=g[NCLASS * nb2:2 * NCLASS * nb2].reshape(NCLASS, nb2).copy(),
                         stats=g[2 * NCLASS * nb2:].reshape(NCLASS, NSTAT).copy(), nbins=nb, xmin=lo, xmax=hi)
        off += hist_doubles(nb)
    return out


_orc = None


def oracle_lib():
    global _orc
    if _orc is None:
        from _oracle_port import lib
        L = lib()
        L.orc_minimc_run.restype = C.c_int
        L.orc_minimc_run.argtypes = [C.c_void_p, C.POINTER(OrcMmcCfg), C.c_uint64, C.c_uint64, C.c_int,
                                     C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _orc = L
    return _orc


def run_oracle(port_oracle, sc, first=0, count=None):
    """oracle/oracle_mmc.c on the material of a PortOracle; returns (hists, meta)"""
    L = oracle_lib()
    count = sc.n if count is None else count
    nt = len(sc.tallies)
    types = (C.c_int * nt)(*[TALLY_TYPES.index(t[0]) for t in sc.tallies])
    nbins = (C.c_int * nt)(*[t[1] for t in sc.tallies])
    lo = (C.c_double * nt)(*[t[2] for t in sc.tallies])
    hi = (C.c_double * nt)(*[t[3] for t in sc.tallies])
    out = np.zeros(sum(hist_doubles(t[1]) for t in sc.tallies))
    meta = np.zeros(5)
    cfg = sc.orc_cfg()
    rc = L.orc_minimc_run(port_oracle.h, C.byref(cfg), first, count, nt, types, nbins, lo, hi,
                          out.ctypes.data_as(C.POINTER(C.c_double)), meta.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0, "oracle transport error flags %d" % rc
    return split_device_layout(out, sc.tallies), dict(miss_count=meta[0], miss_weight=meta[1], tallied_count=meta[2],
                                                      tallied_weight=meta[3], steps=meta[4])


def hists_from_json(js, tallies):
    """result JSON (reference or product) -> same dict layout as split_device_layout, plus the totals"""
    d = json.loads(js) if isinstance(js, str) else js
    out = {}
    for name, nb, lo, hi in tallies:
        t = d["output"]["tally"][name]
        b = t["total"]["bindata"]
        assert b["nbins"] == nb and b["xmin"] == lo and b["xmax"] == hi
        tot_c = np.array([b["underflow"]] + b["content"] + [b["overflow"]])
        tot_e = np.array([b["underflow_errorsq"]] + b["errorsq"] + [b["overflow_errorsq"]])
        content, errsq = np.zeros((NCLASS, nb + 2)), np.zeros((NCLASS, nb + 2))
        for k, cn in enumerate(CLASS_NAMES):
            if cn in t.get("breakdown", {}):
                bb = t["breakdown"][cn]["bindata"]
                content[k] = [bb["underflow"]] + bb["content"] + [bb["overflow"]]
                errsq[k] = [bb["underflow_errorsq"]] + bb["errorsq"] + [bb["overflow_errorsq"]]
        out[name] = dict(content=content, errsq=errsq, total_content=tot_c, total_errsq=tot_e,
                         stats=t["total"]["stats"], nbins=nb, xmin=lo, xmax=hi)
    return out, d["output"]["metadata"]


def chi2_pvalue(c1, e1sq, c2, e2sq):
    """Hist1D.check_compat(force_norm=True): both normalised to unit integral (incl. flow bins), chi-square over bins
    filled in either, dof = number of such bins (flow bins count as two more)."""
    from scipy.stats import chi2
    c1, e1sq, c2, e2sq = [np.asarray(a, dtype=float) for a in (c1, e1sq, c2, e2sq)]
    s1, s2 = c1.sum(), c2.sum()
    c1, e1sq, c2, e2sq = c1 / s1, e1sq / s1 ** 2, c2 / s2, e2sq / s2 ** 2
    inner = slice(1, -1)
    mask = (e1sq[inner] > 0) | (e2sq[inner] > 0)
    chi = (((c1[inner] - c2[inner]) ** 2)[mask] / (e1sq[inner] + e2sq[inner])[mask]).sum()
    k = int(mask.sum()) + 2
    for j in (0, -1):
        if c1[j] > 0 or c2[j] > 0:
            chi += (c1[j] - c2[j]) ** 2 / (e1sq[j] + e2sq[j])
    k = max(1, k - 1)
    return float(chi2.sf(chi, k)), float(chi), k


def reference_minimc(cfgstr, sc, nthreads=2, n=None):
    """the reference's own MiniMC through ncrystal_jsonquery (only where oracle/_ref holds libNCrystal.so)"""
    p = os.path.join(ROOT, "oracle", "_ref", "lib", "libNCrystal.so")
    L = C.CDLL(p)
    L.ncrystal_jsonquery.restype = C.c_void_p
    L.ncrystal_jsonquery.argtypes = [C.c_char_p]
    L.ncrystal_dealloc_string.argtypes = [C.c_void_p]
    eng = "nthreads=%d;" % nthreads + sc.enginecfg()
    if sc.seed == 0:
        eng = eng.replace(";seed=0", "")
    q = "\x07".join(["mmc", "run", cfgstr, sc.geomcfg, sc.srccfg(n), eng]).encode()
    ptr = L.ncrystal_jsonquery(q)
    if not ptr:
        raise RuntimeError("reference MiniMC query failed")
    s = C.string_at(ptr).decode()
    L.ncrystal_dealloc_string(ptr)
    return s


def std_sphere_radius(macroxs_per_m):
    """minimc_unittest_stdsphere (tests/pypath/NCTestUtils/minimc_ref.py:30-62): diameter = 1/Sigma_scat"""
    return 0.5 * (1.0 / macroxs_per_m)


def scenarios(macroxs):
    """macroxs(material key, ekin) -> macroscopic scattering cross section [1/m] (numdens*xs*100), used like the
    reference's tests do to size the standard sphere (diameter = one scattering mean free path)."""
    r_al4 = std_sphere_radius(macroxs("Al", WL2EKIN / 16.0))
    r_al1 = std_sphere_radius(macroxs("Al", WL2EKIN / 1.0))
    r_h2o = std_sphere_radius(macroxs("H2O", WL2EKIN / 1.0))
    S = []
    # tests/scripts/mmc_al.py: Al_sg225 at 4.0 Aa / 1.0 Aa, pencil beam entering a sphere of diameter ~1/Sigma_s
    S.append(Scenario("al_4Aa", "Al", ("sphere", {"r": r_al4}), "constant", ("wl", 4.0), 100000,
                      pos=(0, 0, -r_al4 * (1 - 1e-13))))
    S.append(Scenario("al_1Aa", "Al", ("sphere", {"r": r_al1}), "constant", ("wl", 1.0), 100000,
                      pos=(0, 0, -r_al1 * (1 - 1e-13)),
                      tallies=(("theta", 90, 0.0, 180.0), ("mu", 200, -1.0, 1.0), ("nscat", 22, -1.5, 20.5),
                               ("e", 20, 0.0, 0.2), ("q", 100, 0.0, 15.0))))
    # tests/scripts/mmc_circ.py: water, uniformly illuminated sphere (circular beam of the sphere's radius)
    S.append(Scenario("circ_h2o", "H2O", ("sphere", {"r": r_h2o}), "circular", ("wl", 1.0), 100000,
                      pos=(0, 0, -r_h2o * (1 - 1e-13)), radius=r_h2o))
    # unbounded geometries, energy ranges, scattering limit, no absorption
    S.append(Scenario("slab_ch2", "CH2", ("slab", {"dz": 0.002}), "constant", ("ekin", (0.01, 0.1)), 50000,
                      pos=(0, 0, -0.01), tallies=(("theta", 90, 0.0, 180.0), ("de", 20, -0.1, 0.1), ("l", 25, 0.0, 5.0),
                                                  ("w", 50, 0.0, 1.0), ("nscat_uw", 22, -1.5, 20.5))))
    S.append(Scenario("box_yag", "YAG", ("box", {"dx": 0.01, "dy": 0.02, "dz": 0.005}), "circular", ("wl", (1.5, 2.5)), 50000,
                      pos=(0.001, -0.002, -0.05), radius=0.015, nscatlimit=3, absorption=False,
                      tallies=(("theta", 90, 0.0, 180.0), ("nscat", 22, -1.5, 20.5))))
    S.append(Scenario("cyl_al", "Al", ("cyl", {"r": 0.02, "dy": 0.03}), "constant", ("wl", 1.8), 50000,
                      pos=(0.005, 0.01, -0.1), direction=(0.0, 0.1, 1.0), roulette=(0.5, 1e-3, 3)))
    S.append(Scenario("cylinf_h2o", "H2O", ("cyl", {"r": 0.003}), "circular", ("ekin", 0.0253), 50000,
                      pos=(0, 0, -0.02), radius=0.004, ignore_miss=True))
    # tests/scripts/mmc_scge.py analogue on the oriented benchmark crystal (mosaic single crystal, 1 cm sphere)
    S.append(Scenario("scge", "Ge", ("sphere", {"r": 0.005}), "constant", ("wl", 3.2), 20000,
                      pos=(0, 0, -0.005 * (1 - 1e-13))))
    # isotropic point source inside / spherical-shell source around the sample, log-normal and Maxwell spectra
    # (tests/scripts/mmcenergy.py exercises these energy modes)
    S.append(Scenario("iso_al", "Al", ("sphere", {"r": 0.03}), "isotropic", ("ekin", (0.025, 0.003, "lognormal")), 50000,
                      pos=(0.005, 0.0, -0.01), tallies=(("theta", 90, 0.0, 180.0), ("mu", 100, -1.0, 1.0), ("e", 20, 0.0, 0.1))))
    S.append(Scenario("isoshell_ch2", "CH2", ("sphere", {"r": 0.004}), "isotropic",
                      ("wl", (1.8, 0.1, "lognormal")), 50000, pos=(0.0, 0.0, 0.0), radius=0.05,
                      tallies=(("theta", 90, 0.0, 180.0), ("mu", 100, -1.0, 1.0))))
    # isotropic point source OUTSIDE a box: most neutrons miss (tallied at theta = 0 w.r.t. their own direction)
    S.append(Scenario("isopoint_box_ch2", "CH2", ("box", {"dx": 0.004, "dy": 0.006, "dz": 0.002}), "isotropic",
                      ("wl", (1.8, 0.1, "lognormal")), 50000, pos=(0.001, 0.0, -0.006),
                      tallies=(("theta", 90, 0.0, 180.0), ("q", 50, 0.0, 10.0))))
    S.append(Scenario("thermal_h2o", "H2O", ("slab", {"dz": 0.001}), "constant", ("thermal", 300.0), 50000,
                      pos=(0, 0, -0.01), tallies=(("theta", 90, 0.0, 180.0), ("e", 20, 0.0, 0.3), ("de", 20, -0.3, 0.3))))
    return {s.key: s for s in S}


def port_oracle_for(key):
    """(PortOracle, header dict) of a benchmark material's compiled blob"""
    import sys
    sys.path.insert(0, ROOT)
    from __graft_entry__ import CONFIGS
    from _oracle_port import PortOracle
    from oracle_check import material_path
    from test_cpu_blob import parse_header
    blob = open(material_path(CONFIGS[key]), "rb").read()
    return PortOracle(blob), parse_header(blob)


_cache = {}


def cached_oracle(key):
    if key not in _cache:
        _cache[key] = port_oracle_for(key)
    return _cache[key]


def oracle_macroxs(key, ekin):
    o, h = cached_oracle(key)
    return 100.0 * h["numdens"] * float(o.xs_iso(np.array([ekin]))[0])


def all_scenarios():
    return scenarios(oracle_macroxs)


NO_REFERENCE = ()


def load_golden():
    return json.load(open(os.path.join(GOLDEN, "mmc_reference.json")))


# Tallies of quantities that survive elastic scatterings (energy, wavelength, weight) get several correlated records
# from one neutron history, so the per-bin errors (sum of squared weights) underestimate the spread and the
# chi-square of two independent runs is inflated -- the reference against itself gives chi2/dof ~ 1.5-2.5 there
# (p < 1e-3).  The reference's tests only use theta/q; for the others the criterion is chi2/dof < 4.
CORRELATED_TALLIES = ("e", "l", "de", "w", "nscat_uw")


def compatible(name, c1, e1, c2, e2, pmin=0.001):
    p, chi, k = chi2_pvalue(c1, e1, c2, e2)
    ok = (chi / k < 4.0) if name in CORRELATED_TALLIES else (p > pmin)
    return ok, "chi2=%.1f dof=%d p=%.2g" % (chi, k, p)


def result_dict_from_layout(h, meta, tallies, provided):
    """device/oracle tally layout -> the result-JSON layout (what Scatter.minimc returns), for host-logic tests"""
    def one(c, e, st):
        sw, swx, swx2 = st[..., 0].sum(), st[..., 1].sum(), st[..., 2].sum()
        stats = dict(integral=0.0, mean=None, rms=None, minfilled=None, maxfilled=None)
        if sw > 0:
            m = swx / sw
            stats = dict(integral=float(sw), mean=float(m), rms=float(max(0.0, swx2 / sw - m * m) ** 0.5),
                         minfilled=float(st[..., 3].min()), maxfilled=float(st[..., 4].max()))
        return dict(stats=stats, bindata=dict(nbins=len(c) - 2, content=c[1:-1].tolist(), errorsq=e[1:-1].tolist(),
                                              underflow=float(c[0]), overflow=float(c[-1]),
                                              underflow_errorsq=float(e[0]), overflow_errorsq=float(e[-1])))
    out = {}
    for name, nb, lo, hi in tallies:
        v = h[name]
        filled = v["stats"][:, 0] > 0
        st_tot = v["stats"][filled] if filled.any() else np.zeros((1, NSTAT))
        out[name] = dict(total=one(v["content"].sum(axis=0), v["errsq"].sum(axis=0), st_tot),
                         breakdown={cn: one(v["content"][k], v["errsq"][k],
                                            v["stats"][k:k + 1] if v["stats"][k, 0] > 0 else np.zeros((1, NSTAT)))
                                    for k, cn in enumerate(CLASS_NAMES)})
    md = dict(provided=dict(count=int(provided), weight=float(provided)),
              miss=dict(count=int(meta["miss_count"]), weight=float(meta["miss_weight"])),
              tallied=dict(count=int(meta["tallied_count"]), weight=float(meta["tallied_weight"])))
    return dict(output=dict(tally=out, metadata=md))


def run_hostsim(hostsim, hdr, sc, first=0, count=None):
    """the PRODUCT's NCB_HD transport physics compiled for the host (tests/hostsim), history by history"""
    L = hostsim.lib()
    count = sc.n if count is None else count
    nt = len(sc.tallies)
    types = (C.c_int * nt)(*[TALLY_TYPES.index(t[0]) for t in sc.tallies])
    nbins = (C.c_int * nt)(*[t[1] for t in sc.tallies])
    lo = (C.c_double * nt)(*[t[2] for t in sc.tallies])
    hi = (C.c_double * nt)(*[t[3] for t in sc.tallies])
    out = np.zeros(sum(hist_doubles(t[1]) for t in sc.tallies))
    meta = np.zeros(5)
    cfg = sc.orc_cfg()
    dp = C.POINTER(C.c_double)
    rc = L.hostsim_minimc_run(hostsim.h, C.byref(cfg), hdr["numdens"], hdr["abs_c"], first, count, nt, types, nbins, lo, hi,
                              out.ctypes.data_as(dp), meta.ctypes.data_as(dp))
    assert rc == 0, "hostsim transport error flags %d" % rc
    return split_device_layout(out, sc.tallies), dict(miss_count=meta[0], miss_weight=meta[1], tallied_count=meta[2],
                                                      tallied_weight=meta[3], steps=meta[4])
