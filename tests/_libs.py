"""ctypes wrappers for the TEST-ONLY libraries:

  * RefDrv  -- oracle/_ref/lib/libncb200_refdrv.so: the unmodified reference
               (material compiler, Philox-replay oracle, CPU baseline)
  * HostSim -- tests/hostsim/libncb200_hostsim.so: host compilation of the
               product's device functions (kernel logic without a GPU)

Neither is ever imported by the product package.
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDRV_PATH = os.path.join(ROOT, "oracle", "_ref", "lib", "libncb200_refdrv.so")
HOSTSIM_DIR = os.path.join(ROOT, "tests", "hostsim")
HOSTSIM_PATH = os.path.join(HOSTSIM_DIR, "libncb200_hostsim.so")

_dp = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)


def _d(a):
    return a.ctypes.data_as(_dp)


def have_refdrv():
    return os.path.exists(REFDRV_PATH)


class RefDrv:
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(REFDRV_PATH)
            L.refdrv_create.restype = C.c_void_p
            L.refdrv_create.argtypes = [C.c_char_p]
            L.refdrv_destroy.argtypes = [C.c_void_p]
            L.refdrv_lasterror.restype = C.c_char_p
            L.refdrv_compile.restype = C.c_void_p
            L.refdrv_compile.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
            L.refdrv_compile_ex.restype = C.c_void_p
            L.refdrv_compile_ex.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_uint]
            L.refdrv_free.argtypes = [C.c_void_p]
            L.refdrv_ncomp.argtypes = [C.c_void_p]
            L.refdrv_compname.restype = C.c_char_p
            L.refdrv_compname.argtypes = [C.c_void_p, C.c_int]
            L.refdrv_compscale.restype = C.c_double
            L.refdrv_compscale.argtypes = [C.c_void_p, C.c_int]
            L.refdrv_isoriented.argtypes = [C.c_void_p]
            L.refdrv_xs_iso.argtypes = [C.c_void_p, _dp, C.c_uint64, _dp]
            L.refdrv_xs_iso_components.argtypes = [C.c_void_p, _dp, C.c_uint64, _dp]
            L.refdrv_xs.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_uint64, _dp]
            L.refdrv_sample_iso.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, _dp, C.c_uint64, _dp, _dp, _u32p]
            L.refdrv_sample_iso_leaf.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, _dp, C.c_uint64, _dp, _dp, _u32p]
            L.refdrv_sample.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, _dp, _dp, _dp, _dp, C.c_uint64,
                                        _dp, _dp, _dp, _dp, _u32p]
            L.refdrv_sab_sampler_dump.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]
            L.refdrv_bench_capi.restype = C.c_double
            L.refdrv_bench_capi.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_uint64,
                                            _dp, _dp, _dp, _dp]
            cls._lib = L
        return cls._lib

    def __init__(self, cfg):
        L = self.lib()
        self.cfg = cfg
        self.h = L.refdrv_create(cfg.encode())
        if not self.h:
            raise RuntimeError("refdrv_create failed: %s" % L.refdrv_lasterror().decode())

    def __del__(self):
        if getattr(self, "h", None):
            self.lib().refdrv_destroy(self.h)
            self.h = None

    @property
    def ncomp(self):
        return self.lib().refdrv_ncomp(self.h)

    def compnames(self):
        return [self.lib().refdrv_compname(self.h, i).decode() for i in range(self.ncomp)]

    def compile(self, flags=0):
        """flags & 1: S(alpha,beta) leaves derived from a phonon density of states are emitted as that density."""
        n = C.c_uint64(0)
        p = self.lib().refdrv_compile_ex(self.h, C.byref(n), flags)
        if not p:
            raise RuntimeError("refdrv_compile failed: %s" % self.lib().refdrv_lasterror().decode())
        try:
            return C.string_at(p, n.value)
        finally:
            self.lib().refdrv_free(p)

    def xs_iso(self, ekin):
        ekin = np.ascontiguousarray(ekin, dtype=np.float64)
        out = np.empty_like(ekin)
        self.lib().refdrv_xs_iso(self.h, _d(ekin), ekin.size, _d(out))
        return out

    def xs_iso_components(self, ekin):
        ekin = np.ascontiguousarray(ekin, dtype=np.float64)
        out = np.empty((self.ncomp, ekin.size))
        self.lib().refdrv_xs_iso_components(self.h, _d(ekin), ekin.size, _d(out))
        return out

    def xs(self, ekin, ux, uy, uz):
        ekin, ux, uy, uz = [np.ascontiguousarray(a, dtype=np.float64) for a in (ekin, ux, uy, uz)]
        out = np.empty_like(ekin)
        self.lib().refdrv_xs(self.h, _d(ekin), _d(ux), _d(uy), _d(uz), ekin.size, _d(out))
        return out

    def sample_iso(self, ekin, seed=1, first_index=0, leaf=None):
        ekin = np.ascontiguousarray(ekin, dtype=np.float64)
        eo = np.empty_like(ekin)
        mu = np.empty_like(ekin)
        nd = np.zeros(ekin.size, dtype=np.uint32)
        if leaf is None:
            self.lib().refdrv_sample_iso(self.h, seed, first_index, _d(ekin), ekin.size, _d(eo), _d(mu),
                                         nd.ctypes.data_as(_u32p))
        else:
            self.lib().refdrv_sample_iso_leaf(self.h, leaf, seed, first_index, _d(ekin), ekin.size, _d(eo), _d(mu),
                                              nd.ctypes.data_as(_u32p))
        return eo, mu, nd

    def sample(self, ekin, ux, uy, uz, seed=1, first_index=0):
        ekin, ux, uy, uz = [np.ascontiguousarray(a, dtype=np.float64) for a in (ekin, ux, uy, uz)]
        eo, ox, oy, oz = [np.empty_like(ekin) for _ in range(4)]
        nd = np.zeros(ekin.size, dtype=np.uint32)
        self.lib().refdrv_sample(self.h, seed, first_index, _d(ekin), _d(ux), _d(uy), _d(uz), ekin.size,
                                 _d(eo), _d(ox), _d(oy), _d(oz), nd.ctypes.data_as(_u32p))
        return eo, ox, oy, oz, nd

    def sab_sampler_dump(self, c, iE, nbeta):
        return _sab_dump(self.lib().refdrv_sab_sampler_dump, self.h, c, iE, nbeta)

    def sab_auto_egrid(self, c, cap=1000):
        """(egrid, xs) of the reference's SABIntegrator on leaf c's kernel with a fully automatic energy grid."""
        eg, xs = np.zeros(cap), np.zeros(cap)
        f = self.lib().refdrv_sab_auto_egrid
        f.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int]
        n = f(self.h, c, _d(eg), _d(xs), cap)
        if n < 0:
            raise RuntimeError("refdrv_sab_auto_egrid failed (%d)" % n)
        return eg[:n], xs[:n]

    @classmethod
    def capi_sample_iso(cls, cfg, ekin, nthreads=None):
        """Independent-stream sampling through the reference's own C-API and builtin RNG
        (ncrystal_samplescatterisotropic_many on cloned handles, one per host thread)."""
        ekin = np.ascontiguousarray(ekin, dtype=np.float64)
        nthreads = nthreads or len(os.sched_getaffinity(0))
        eo, mu = np.empty_like(ekin), np.empty_like(ekin)
        null = C.cast(None, _dp)
        cls.lib().refdrv_bench_capi(cfg.encode(), 1, nthreads, 0, _d(ekin), null, null, null, ekin.size,
                                    _d(eo), _d(mu), null, null)
        return eo, mu


def _capi_sample_aniso(cls, cfg, ekin, ux, uy, uz, nthreads=None):
    """Independent-stream oriented sampling through the reference's per-neutron C-API (ncrystal_samplescatter on
    cloned handles, one per host thread)."""
    ekin, ux, uy, uz = [np.ascontiguousarray(a, dtype=np.float64) for a in (ekin, ux, uy, uz)]
    nthreads = nthreads or len(os.sched_getaffinity(0))
    eo, ox, oy, oz = [np.empty_like(ekin) for _ in range(4)]
    cls.lib().refdrv_bench_capi(cfg.encode(), 3, nthreads, 0, _d(ekin), _d(ux), _d(uy), _d(uz), ekin.size,
                                _d(eo), _d(ox), _d(oy), _d(oz))
    return eo, ox, oy, oz


RefDrv.capi_sample_aniso = classmethod(_capi_sample_aniso)


def _sab_dump(fn, h, c, iE, nbeta):
    x = np.zeros(nbeta + 1)
    pdf = np.zeros(nbeta + 1)
    cdf = np.zeros(nbeta + 1)
    infos = np.zeros((nbeta, 10))
    meta = np.zeros(2)
    n = fn(h, c, iE, _d(x), _d(pdf), _d(cdf), _d(infos), _d(meta))
    if n < 0:
        raise RuntimeError("sab dump failed")
    return dict(n=n, x=x[:n], pdf=pdf[:n], cdf=cdf[:n], infos=infos[:max(n - 1, 0)], ibeta_off=int(meta[0]),
                first_bin=meta[1])


def build_hostsim():
    subprocess.check_call(["make", "-s", "-C", HOSTSIM_DIR])


class HostSim:
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            build_hostsim()
            L = C.CDLL(HOSTSIM_PATH)
            L.hostsim_load.restype = C.c_void_p
            L.hostsim_load.argtypes = [C.c_char_p, C.c_uint64]
            L.hostsim_free.argtypes = [C.c_void_p]
            L.hostsim_lasterror.restype = C.c_char_p
            L.hostsim_ncomp.argtypes = [C.c_void_p]
            L.hostsim_component_kind.argtypes = [C.c_void_p, C.c_int]
            L.hostsim_xs_iso.argtypes = [C.c_void_p, _dp, C.c_uint64, _dp]
            L.hostsim_xs_iso_components.argtypes = [C.c_void_p, _dp, C.c_uint64, _dp]
            L.hostsim_sample_iso.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, _dp, C.c_uint64, _dp, _dp, _u32p, _i32p]
            L.hostsim_sample_iso_leaf.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, _dp, C.c_uint64, _dp, _dp,
                                                  _u32p, _i32p]
            L.hostsim_sample_sab_staged.argtypes = L.hostsim_sample_iso_leaf.argtypes
            L.hostsim_xs.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_uint64, _dp]
            L.hostsim_sample.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, _dp, _dp, _dp, _dp, C.c_uint64,
                                         _dp, _dp, _dp, _dp, _u32p, _i32p]
            L.hostsim_uniforms.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, _dp]
            L.hostsim_sab_sampler_dump.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]
            L.hostsim_sab_xscheck.argtypes = [C.c_void_p, C.c_int, _dp]
            L.hostsim_minimc_run.restype = C.c_int
            L.hostsim_minimc_run.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_uint64, C.c_uint64, C.c_int,
                                             C.POINTER(C.c_int), C.POINTER(C.c_int), _dp, _dp, _dp, _dp]
            cls._lib = L
        return cls._lib

    def __init__(self, blob):
        L = self.lib()
        self.h = L.hostsim_load(blob, len(blob))
        if not self.h:
            raise RuntimeError("hostsim_load failed: %s" % L.hostsim_lasterror().decode())

    def __del__(self):
        if getattr(self, "h", None):
            self.lib().hostsim_free(self.h)
            self.h = None

    @property
    def ncomp(self):
        return self.lib().hostsim_ncomp(self.h)

    def component_kind(self, c):
        return self.lib().hostsim_component_kind(self.h, c)

    def xs_iso(self, ekin):
        ekin = np.ascontiguousarray(ekin, dtype=np.float64)
        out = np.empty_like(ekin)
        self.lib().hostsim_xs_iso(self.h, _d(ekin), ekin.size, _d(out))
        return out

    def xs_iso_components(self, ekin):
        ekin = np.ascontiguousarray(ekin, dtype=np.float64)
        out = np.empty((self.ncomp, ekin.size))
        self.lib().hostsim_xs_iso_components(self.h, _d(ekin), ekin.size, _d(out))
        return out

    def sample_iso(self, ekin, seed=1, first_index=0, leaf=None):
        ekin = np.ascontiguousarray(ekin, dtype=np.float64)
        eo = np.empty_like(ekin)
        mu = np.empty_like(ekin)
        nd = np.zeros(ekin.size, dtype=np.uint32)
        er = np.zeros(ekin.size, dtype=np.int32)
        if leaf is None:
            self.lib().hostsim_sample_iso(self.h, seed, first_index, _d(ekin), ekin.size, _d(eo), _d(mu),
                                          nd.ctypes.data_as(_u32p), er.ctypes.data_as(_i32p))
        else:
            self.lib().hostsim_sample_iso_leaf(self.h, leaf, seed, first_index, _d(ekin), ekin.size, _d(eo), _d(mu),
                                               nd.ctypes.data_as(_u32p), er.ctypes.data_as(_i32p))
        return eo, mu, nd, er

    def sample_sab_staged(self, ekin, leaf, seed=1, first_index=0):
        ekin = np.ascontiguousarray(ekin, dtype=np.float64)
        eo = np.empty_like(ekin)
        mu = np.empty_like(ekin)
        nd = np.zeros(ekin.size, dtype=np.uint32)
        er = np.zeros(ekin.size, dtype=np.int32)
        self.lib().hostsim_sample_sab_staged(self.h, leaf, seed, first_index, _d(ekin), ekin.size, _d(eo), _d(mu),
                                             nd.ctypes.data_as(_u32p), er.ctypes.data_as(_i32p))
        return eo, mu, nd, er

    def xs(self, ekin, ux, uy, uz):
        ekin, ux, uy, uz = [np.ascontiguousarray(a, dtype=np.float64) for a in (ekin, ux, uy, uz)]
        out = np.empty_like(ekin)
        self.lib().hostsim_xs(self.h, _d(ekin), _d(ux), _d(uy), _d(uz), ekin.size, _d(out))
        return out

    def sample(self, ekin, ux, uy, uz, seed=1, first_index=0):
        ekin, ux, uy, uz = [np.ascontiguousarray(a, dtype=np.float64) for a in (ekin, ux, uy, uz)]
        eo, ox, oy, oz = [np.empty_like(ekin) for _ in range(4)]
        nd = np.zeros(ekin.size, dtype=np.uint32)
        er = np.zeros(ekin.size, dtype=np.int32)
        self.lib().hostsim_sample(self.h, seed, first_index, _d(ekin), _d(ux), _d(uy), _d(uz), ekin.size,
                                  _d(eo), _d(ox), _d(oy), _d(oz), nd.ctypes.data_as(_u32p), er.ctypes.data_as(_i32p))
        return eo, ox, oy, oz, nd, er

    def uniforms(self, seed, index, n):
        out = np.empty(n)
        self.lib().hostsim_uniforms(seed, index, n, _d(out))
        return out

    def sab_sampler_dump(self, c, iE, nbeta):
        return _sab_dump(self.lib().hostsim_sab_sampler_dump, self.h, c, iE, nbeta)

    def sab_egrid(self, c, negrid):
        out = np.zeros(negrid)
        self.lib().hostsim_sab_egrid.argtypes = [C.c_void_p, C.c_int, _dp]
        n = self.lib().hostsim_sab_egrid(self.h, c, _d(out))
        if n < 0:
            raise RuntimeError("egrid failed")
        return out[:n]

    def sab_xscheck(self, c, negrid):
        out = np.zeros(negrid)
        n = self.lib().hostsim_sab_xscheck(self.h, c, _d(out))
        if n < 0:
            raise RuntimeError("xscheck failed")
        return out[:n]


def isotropic_directions(n, seed=54321):
    rng = np.random.Generator(np.random.Philox(key=seed))
    z = 2.0 * rng.random(n) - 1.0
    phi = 2.0 * np.pi * rng.random(n)
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    return r * np.cos(phi), r * np.sin(phi), z


def loguniform_energies(n, seed=12345, lo=1e-5, hi=10.0):
    """ekin[i] = 10^(log10(lo) + (log10(hi)-log10(lo))*u_i), u from a fixed-seed counter generator."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    u = rng.random(n)
    return 10.0 ** (np.log10(lo) + (np.log10(hi) - np.log10(lo)) * u)


def strip_sab_energy_grids(blob, emax_request=0.0):
    """A copy of a compiled material in which every S(alpha,beta) leaf has lost what the reference's SABIntegrator
    derived for it -- the energy grid, the cross sections on it and the extension constants -- and asks the library to
    determine them itself (ncb_sab_t::auto_egrid = 1).  The kernel (SABData), its extender constants and the number of
    grid points stay.  emax_request: an Emax the material's file asks for ("egrid" line), 0 = automatic."""
    import struct
    b = bytearray(blob)
    ncomp = struct.unpack_from("<I", b, 12)[0]
    hdr_fixed = struct.calcsize("<QIIIIQdddddd176s")
    comp_sz = struct.calcsize("<IIdddQQ")
    for i in range(ncomp):
        kind, _r, _s, _lo, _hi, off, _n = struct.unpack_from("<IIdddQQ", b, hdr_fixed + i * comp_sz)
        if kind != 3:
            continue
        negrid = struct.unpack_from("<Q", b, off + 14 * 8)[0]
        struct.pack_into("<4d", b, off + 9 * 8, 0.0, 0.0, 0.0, 0.0)       # k_extension, xs_at_emax, k1, k2
        struct.pack_into("<Q", b, off + 14 * 8 + 3 * 8, 1)                   # auto_egrid
        base = off + 14 * 8 + 4 * 8
        b[base:base + 16 * negrid] = bytes(16 * negrid)                      # egrid[], xs[]
        struct.pack_into("<2d", b, base, 0.0, float(emax_request))           # requested (emin, emax)
    return bytes(b)
