"""Host-side mirror of the reference's Python Process/Scatter interface for the hot path
(ref: ncrystal_python/src/NCrystal/core.py:1226-1590): same method names, argument meaning
and error behaviour, over the C ABI of libncrystal_b200.so.

Two kinds of arguments are accepted by the batched methods:
  * numpy arrays / scalars (host memory)  -> the reference-facing `*_many` C entry points
    (host<->device copies inside the call), results are numpy arrays;
  * torch CUDA tensors (float64, contiguous) -> the `*_dev` entry points on the tensor's
    device and torch's current stream, results are torch tensors, no synchronisation.
"""
import ctypes as C

import numpy as np

from . import _lib

_dblp = C.POINTER(C.c_double)


class NCException(RuntimeError):
    pass


class NCBadInput(NCException):
    pass


class NCCalcError(NCException):
    pass


class NCLogicError(NCException):
    pass


class NCFileNotFound(NCException):
    pass


_ERRCLASSES = {"BadInput": NCBadInput, "CalcError": NCCalcError, "LogicError": NCLogicError,
               "FileNotFound": NCFileNotFound}


def _check_error():
    L = _lib.lib()
    if L.ncrystal_error():
        msg = (L.ncrystal_lasterror() or b"").decode()
        typ = (L.ncrystal_lasterrortype() or b"").decode()
        L.ncrystal_clearerror()
        raise _ERRCLASSES.get(typ, NCException)(msg)


def _is_torch_cuda(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "is_cuda") and x.is_cuda


def _np_d(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def _dev_check(*tensors):
    import torch
    dev = tensors[0].device
    for t in tensors:
        if t is None:
            continue
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.device == dev):
            raise NCBadInput("device arrays must be contiguous float64 CUDA tensors on one device")
    return dev


def _stream_ptr(dev):
    import torch
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class Process:
    """ref: core.py:1226 (class Process)"""

    def __init__(self, handle):
        self._h = handle  # ncrystal_scatter_t or ncrystal_absorption_t
        self._L = _lib.lib()
        if isinstance(handle, _lib.ncrystal_absorption_t):
            self._p = self._L.ncrystal_cast_abs2proc(handle)
        else:
            self._p = self._L.ncrystal_cast_scat2proc(handle)
        _check_error()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.internal:
            try:
                self._L.ncrystal_unref(C.byref(h))
            except Exception:
                pass

    def getName(self):
        n = self._L.ncrystal_name(self._p)
        _check_error()
        return n.decode()

    getCalcName = getName
    name = property(getName)

    def getUniqueID(self):
        """ref: core.py:1245 -- UID of the underlying (shared, immutable) process: equal for a handle and its clones"""
        p = self._L.ncrystal_process_uid(self._p)
        _check_error()
        v = int(C.cast(p, C.c_char_p).value)
        self._L.ncrystal_dealloc_string(p)
        return v
    uid = property(getUniqueID)

    def isNull(self):
        """ref: core.py:1261 -- a null process vanishes everywhere (domain with elow >= ehigh or elow = inf)"""
        elow, ehigh = self.domain()
        return elow == float("inf") or elow >= ehigh

    def crossSectionNonOriented(self, ekin, repeat=None):
        """ref: core.py:1289 -- deprecated alias of crossSectionIsotropic"""
        return self.crossSectionIsotropic(ekin, repeat)

    def domain(self):
        a, b = C.c_double(), C.c_double()
        self._L.ncrystal_domain(self._p, C.byref(a), C.byref(b))
        _check_error()
        return a.value, b.value

    def isNonOriented(self):
        return bool(self._L.ncrystal_isnonoriented(self._p))

    def isOriented(self):
        return not self.isNonOriented()

    def refCount(self):
        return self._L.ncrystal_refcount(C.byref(self._h))

    # -- introspection (extensions)
    def components(self):
        n = self._L.ncb200_ncomponents(self._p)
        return [(self._L.ncb200_component_kind(self._p, i), self._L.ncb200_component_scale(self._p, i))
                for i in range(n)]

    def tableBytes(self):
        return int(self._L.ncb200_table_bytes(self._p))

    # -- cross sections
    def crossSectionIsotropic(self, ekin, repeat=None):
        """ref: core.py:1277; C: ncrystal_crosssection_nonoriented[_many]"""
        if _is_torch_cuda(ekin):
            import torch
            dev = _dev_check(ekin)
            out = torch.empty_like(ekin)
            with torch.cuda.device(dev):
                self._L.ncb200_crosssection_nonoriented_many_dev(self._p, ekin.data_ptr(), ekin.numel(),
                                                                 out.data_ptr(), _stream_ptr(dev))
            _check_error()
            return out if repeat is None else out.repeat(repeat)
        if repeat is None and not hasattr(ekin, "__len__"):
            res = C.c_double()
            self._L.ncrystal_crosssection_nonoriented(self._p, float(ekin), C.byref(res))
            _check_error()
            return res.value
        e = _np_d(ekin if hasattr(ekin, "__len__") else [ekin])
        rep = 1 if repeat is None else int(repeat)
        out = np.empty(e.size * rep)
        if out.size:
            self._L.ncrystal_crosssection_nonoriented_many(self._p, e.ctypes.data_as(_dblp), e.size, rep,
                                                           out.ctypes.data_as(_dblp))
            _check_error()
        return out

    crossSectionNonOriented = crossSectionIsotropic

    def crossSection(self, ekin, direction):
        """ref: core.py:1274; C: ncrystal_crosssection / ncb200_crosssection_many[_dev].
        direction: (3,) for a scalar ekin, or (ux,uy,uz) arrays / (n,3) array for a batch."""
        if _is_torch_cuda(ekin):
            import torch
            ux, uy, uz = direction
            dev = _dev_check(ekin, ux, uy, uz)
            out = torch.empty_like(ekin)
            with torch.cuda.device(dev):
                self._L.ncb200_crosssection_many_dev(self._p, ekin.data_ptr(), ux.data_ptr(), uy.data_ptr(), uz.data_ptr(),
                                                     ekin.numel(), out.data_ptr(), _stream_ptr(dev))
            _check_error()
            return out
        if not hasattr(ekin, "__len__"):
            d = (C.c_double * 3)(*[float(x) for x in direction])
            res = C.c_double()
            self._L.ncrystal_crosssection(self._p, float(ekin), C.byref(d), C.byref(res))
            _check_error()
            return res.value
        e = _np_d(ekin)
        ux, uy, uz = _split_dirs(direction, e.size)
        out = np.empty(e.size)
        if e.size:
            self._L.ncb200_crosssection_many(self._p, e.ctypes.data_as(_dblp), ux.ctypes.data_as(_dblp),
                                             uy.ctypes.data_as(_dblp), uz.ctypes.data_as(_dblp), e.size,
                                             out.ctypes.data_as(_dblp))
            _check_error()
        return out

    def xsect(self, ekin=None, direction=None, wl=None, repeat=None):
        """ref: core.py:1297"""
        ekin = _parse_ekin(ekin, wl)
        if direction is None:
            return self.crossSectionIsotropic(ekin, repeat)
        return self.crossSection(ekin, direction)


def _parse_ekin(ekin, wl):
    if wl is not None:
        if ekin is not None:
            raise NCBadInput("Do not specify both ekin and wl")
        wl = np.asarray(wl, dtype=float) if hasattr(wl, "__len__") else float(wl)
        return 0.081804209605330899 / (wl * wl)  # wl2ekin, ref: NCDefs.hh:834-868
    if ekin is None:
        raise NCBadInput("Specify either ekin or wl")
    return ekin


def _split_dirs(direction, n):
    d = direction
    if isinstance(d, (tuple, list)) and len(d) == 3 and hasattr(d[0], "__len__"):
        ux, uy, uz = [_np_d(x) for x in d]
    else:
        a = np.asarray(d, dtype=np.float64)
        if a.ndim == 1:
            a = np.broadcast_to(a, (n, 3))
        ux, uy, uz = [np.ascontiguousarray(a[:, k]) for k in range(3)]
    if not (ux.size == uy.size == uz.size == n):
        raise NCBadInput("direction arrays must match ekin in length")
    return ux, uy, uz


class Scatter(Process):
    """ref: core.py:1421 (class Scatter)"""

    def __init__(self, cfgstr=None, seed=None, _handle=None):
        L = _lib.lib()
        if _handle is None:
            if seed is None:
                _handle = L.ncrystal_create_scatter(cfgstr.encode())
            else:
                _handle = L.ncrystal_create_scatter_builtinrng(cfgstr.encode(), int(seed))
            _check_error()
            if not _handle.internal:
                raise NCException("could not create scatter for cfg %r" % cfgstr)
        super().__init__(_handle)

    @classmethod
    def fromBlob(cls, blob, seed=0):
        L = _lib.lib()
        h = L.ncb200_create_scatter_from_blob(blob, len(blob), int(seed))
        _check_error()
        return cls(_handle=h)

    @classmethod
    def fromFile(cls, path, seed=0):
        L = _lib.lib()
        h = L.ncb200_create_scatter_from_file(str(path).encode(), int(seed))
        _check_error()
        return cls(_handle=h)

    def clone(self, rng_stream_index=None, for_current_thread=False):
        """ref: core.py:1441"""
        if rng_stream_index is not None:
            h = self._L.ncrystal_clone_scatter_rngbyidx(self._h, int(rng_stream_index))
        elif for_current_thread:
            h = self._L.ncrystal_clone_scatter_rngforcurrentthread(self._h)
        else:
            h = self._L.ncrystal_clone_scatter(self._h)
        _check_error()
        return Scatter(_handle=h)

    # -- RNG
    def setRNGStream(self, seed, stream_id=0, next_index=0):
        self._L.ncb200_set_rng_stream(self._h, int(seed), int(stream_id), int(next_index))
        _check_error()

    def getRNGStream(self):
        a, b, c = C.c_uint64(), C.c_uint32(), C.c_uint64()
        self._L.ncb200_get_rng_stream(self._h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def rngSupportsStateManipulation(self):
        return bool(self._L.ncrystal_rngsupportsstatemanip_ofscatter(self._h))

    def getRNGState(self):
        p = self._L.ncrystal_getrngstate_ofscatter(self._h)
        _check_error()
        s = C.string_at(p).decode()
        self._L.ncrystal_dealloc_string(p)
        return s

    def setRNGState(self, state):
        self._L.ncrystal_setrngstate_ofscatter(self._h, state.encode())
        _check_error()

    # -- sampling
    def sampleScatterIsotropic(self, ekin, repeat=None, with_xs=False):
        """ref: core.py:1492; C: ncrystal_samplescatterisotropic[_many].  Returns (ekin_final, mu)."""
        if _is_torch_cuda(ekin):
            import torch
            dev = _dev_check(ekin)
            eo = torch.empty_like(ekin)
            mu = torch.empty_like(ekin)
            xs = torch.empty_like(ekin) if with_xs else None
            with torch.cuda.device(dev):
                if with_xs:
                    self._L.ncb200_xs_and_samplescatterisotropic_many_dev(self._h, ekin.data_ptr(), ekin.numel(),
                                                                          xs.data_ptr(), eo.data_ptr(), mu.data_ptr(),
                                                                          _stream_ptr(dev))
                else:
                    self._L.ncb200_samplescatterisotropic_many_dev(self._h, ekin.data_ptr(), ekin.numel(),
                                                                   eo.data_ptr(), mu.data_ptr(), _stream_ptr(dev))
            _check_error()
            return (xs, eo, mu) if with_xs else (eo, mu)
        if repeat is None and not hasattr(ekin, "__len__"):
            a, b = C.c_double(), C.c_double()
            self._L.ncrystal_samplescatterisotropic(self._h, float(ekin), C.byref(a), C.byref(b))
            _check_error()
            return a.value, b.value
        e = _np_d(ekin if hasattr(ekin, "__len__") else [ekin])
        rep = 1 if repeat is None else int(repeat)
        if with_xs:
            # fused host-pointer call (one pass over the bus); repeat is not supported here
            if rep != 1:
                raise NCBadInput("with_xs=True does not take repeat")
            xs, eo, mu = np.empty(e.size), np.empty(e.size), np.empty(e.size)
            if e.size:
                self._L.ncb200_xs_and_samplescatterisotropic_many(self._h, e.ctypes.data_as(_dblp), e.size,
                                                                  xs.ctypes.data_as(_dblp), eo.ctypes.data_as(_dblp),
                                                                  mu.ctypes.data_as(_dblp))
                _check_error()
            return xs, eo, mu
        eo = np.empty(e.size * rep)
        mu = np.empty(e.size * rep)
        if eo.size:
            self._L.ncrystal_samplescatterisotropic_many(self._h, e.ctypes.data_as(_dblp), e.size, rep,
                                                         eo.ctypes.data_as(_dblp), mu.ctypes.data_as(_dblp))
            _check_error()
        return eo, mu

    def sampleScatter(self, ekin, direction, repeat=None):
        """ref: core.py:1478; C: ncrystal_samplescatter[_many] / ncb200_samplescatter_manydir[_dev].
        Returns (ekin_final, (ux,uy,uz))."""
        if _is_torch_cuda(ekin):
            import torch
            ux, uy, uz = direction
            dev = _dev_check(ekin, ux, uy, uz)
            eo, ox, oy, oz = [torch.empty_like(ekin) for _ in range(4)]
            with torch.cuda.device(dev):
                self._L.ncb200_samplescatter_manydir_dev(self._h, ekin.data_ptr(), ux.data_ptr(), uy.data_ptr(),
                                                         uz.data_ptr(), ekin.numel(), eo.data_ptr(), ox.data_ptr(),
                                                         oy.data_ptr(), oz.data_ptr(), _stream_ptr(dev))
            _check_error()
            return eo, (ox, oy, oz)
        if not hasattr(ekin, "__len__"):
            d = (C.c_double * 3)(*[float(x) for x in direction])
            if repeat is None:
                eo = C.c_double()
                do = (C.c_double * 3)()
                self._L.ncrystal_samplescatter(self._h, float(ekin), C.byref(d), C.byref(eo), C.byref(do))
                _check_error()
                return eo.value, (do[0], do[1], do[2])
            rep = int(repeat)
            eo, ox, oy, oz = [np.empty(rep) for _ in range(4)]
            if rep:
                self._L.ncrystal_samplescatter_many(self._h, float(ekin), C.byref(d), rep, eo.ctypes.data_as(_dblp),
                                                    ox.ctypes.data_as(_dblp), oy.ctypes.data_as(_dblp),
                                                    oz.ctypes.data_as(_dblp))
                _check_error()
            return eo, (ox, oy, oz)
        e = _np_d(ekin)
        ux, uy, uz = _split_dirs(direction, e.size)
        eo, ox, oy, oz = [np.empty(e.size) for _ in range(4)]
        if e.size:
            self._L.ncb200_samplescatter_manydir(self._h, e.ctypes.data_as(_dblp), ux.ctypes.data_as(_dblp),
                                                 uy.ctypes.data_as(_dblp), uz.ctypes.data_as(_dblp), e.size,
                                                 eo.ctypes.data_as(_dblp), ox.ctypes.data_as(_dblp),
                                                 oy.ctypes.data_as(_dblp), oz.ctypes.data_as(_dblp))
            _check_error()
        return eo, (ox, oy, oz)

    def generateScatteringNonOriented(self, ekin, repeat=None):
        """ref: core.py:1523 (deprecated spelling); C: ncrystal_genscatter_nonoriented[_many].
        Returns (scatter_angle [rad], delta_ekin)."""
        if repeat is None and not hasattr(ekin, "__len__"):
            a, b = C.c_double(), C.c_double()
            self._L.ncrystal_genscatter_nonoriented(self._h, float(ekin), C.byref(a), C.byref(b))
            _check_error()
            return a.value, b.value
        e = _np_d(ekin if hasattr(ekin, "__len__") else [ekin])
        rep = 1 if repeat is None else int(repeat)
        ang, de = np.empty(e.size * rep), np.empty(e.size * rep)
        if ang.size:
            self._L.ncrystal_genscatter_nonoriented_many(self._h, e.ctypes.data_as(_dblp), e.size, rep,
                                                         ang.ctypes.data_as(_dblp), de.ctypes.data_as(_dblp))
            _check_error()
        return ang, de

    def generateScattering(self, ekin, direction, repeat=None):
        """ref: core.py:1506 (deprecated spelling); C: ncrystal_genscatter[_many].
        Returns ((ux,uy,uz), delta_ekin) for one fixed (ekin, direction)."""
        d = (C.c_double * 3)(*[float(x) for x in direction])
        if repeat is None:
            de, do = C.c_double(), (C.c_double * 3)()
            self._L.ncrystal_genscatter(self._h, float(ekin), C.byref(d), C.byref(do), C.byref(de))
            _check_error()
            return (do[0], do[1], do[2]), de.value
        rep = int(repeat)
        ox, oy, oz, de = [np.empty(rep) for _ in range(4)]
        if rep:
            self._L.ncrystal_genscatter_many(self._h, float(ekin), C.byref(d), rep, ox.ctypes.data_as(_dblp),
                                             oy.ctypes.data_as(_dblp), oz.ctypes.data_as(_dblp), de.ctypes.data_as(_dblp))
            _check_error()
        return (ox, oy, oz), de

    def genscat(self, ekin=None, direction=None, wl=None, repeat=None):
        """ref: core.py:1554 (deprecated spelling of scatter)"""
        ekin = _parse_ekin(ekin, wl)
        if direction is None:
            return self.generateScatteringNonOriented(ekin, repeat)
        return self.generateScattering(ekin, direction, repeat)

    def scatter(self, ekin=None, direction=None, wl=None, repeat=None):
        """ref: core.py:1543"""
        ekin = _parse_ekin(ekin, wl)
        if direction is None:
            return self.sampleScatterIsotropic(ekin, repeat)
        return self.sampleScatter(ekin, direction, repeat)

    # -- device-resident transport step (the reference's MiniMC, ncrystal_python/src/NCrystal/minimc.py: run)
    def minimc(self, geomcfg, srccfg, enginecfg="", first=None, count=None):
        """Run source neutrons (optionally only the slice [first, first+count)) through a single volume of this
        material on the device; returns the decoded result dictionary ("NCrystalMiniMCResults_v1").
        ref: minimc.py run(cfgstr, geomcfg, srccfg, enginecfg) / ncrystal_jsonquery ["mmc","run",...]"""
        import json
        if first is None and count is None:
            p = self._L.ncb200_minimc_run(self._h, geomcfg.encode(), srccfg.encode(), enginecfg.encode())
        else:
            p = self._L.ncb200_minimc_run_slice(self._h, geomcfg.encode(), srccfg.encode(), enginecfg.encode(),
                                                int(first or 0), int(count if count is not None else 2 ** 63))
        _check_error()
        if not p:
            raise NCCalcError("ncb200_minimc_run failed")
        s = C.string_at(p).decode()
        self._L.ncrystal_dealloc_string(p)
        return json.loads(s)

    def materialBulk(self):
        """(number density [atoms/Aa^3], absorption constant xs*sqrt(E) [barn sqrt(eV)], temperature [K])"""
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._L.ncb200_material_bulk(self._p, C.byref(a), C.byref(b), C.byref(c))
        _check_error()
        return a.value, b.value, c.value

    def checkDeviceErrors(self, device=None):
        """Synchronise torch's current stream and raise if a kernel flagged an error."""
        import torch
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        flags = self._L.ncb200_check_device_errors(self._h, _stream_ptr(dev))
        _check_error()
        return flags


class Absorption(Process):
    """ref: core.py class Absorption -- here always the 1/v process of the compiled material (AbsOOV)."""

    def __init__(self, cfgstr=None, _handle=None):
        L = _lib.lib()
        if _handle is None:
            _handle = L.ncrystal_create_absorption(cfgstr.encode())
            _check_error()
        super().__init__(_handle)

    @classmethod
    def fromBlob(cls, blob):
        h = _lib.lib().ncb200_create_absorption_from_blob(blob, len(blob))
        _check_error()
        return cls(_handle=h)

    def clone(self):
        """ref: core.py:1410; C: ncrystal_clone_absorption"""
        h = self._L.ncrystal_clone_absorption(self._h)
        _check_error()
        return Absorption(_handle=h)


def createAbsorption(cfgstr):
    """ref: core.py createAbsorption / C: ncrystal_create_absorption"""
    return Absorption(cfgstr)


def createScatter(cfgstr, seed=None):
    """ref: core.py createScatter / C: ncrystal_create_scatter[_builtinrng]"""
    return Scatter(cfgstr, seed)


def generateSource(n, seed=12345, first_index=0, lo=1e-5, hi=10.0, directions=False, device=None):
    """Synthetic benchmark source on the device: log-uniform energies (+ isotropic directions)."""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    e = torch.empty(n, dtype=torch.float64, device=dev)
    dirs = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)] if directions else [None] * 3
    with torch.cuda.device(dev):
        _lib.lib().ncb200_generate_source_dev(seed, first_index, n, lo, hi, e.data_ptr(),
                                              *[d.data_ptr() if d is not None else None for d in dirs],
                                              _stream_ptr(dev))
    _check_error()
    return (e, tuple(dirs)) if directions else e


def tallyHist(values, lo, hi, nbins, weights=None, hist=None, sumw2=None):
    """Accumulate a weighted histogram (nbins+2 bins incl. under/overflow) on the device."""
    import torch
    dev = _dev_check(values, weights)
    if hist is None:
        hist = torch.zeros(nbins + 2, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.lib().ncb200_tally_hist_dev(values.data_ptr(), weights.data_ptr() if weights is not None else None,
                                         values.numel(), float(lo), float(hi), int(nbins), hist.data_ptr(),
                                         sumw2.data_ptr() if sumw2 is not None else None, _stream_ptr(dev))
    _check_error()
    return hist


def kernelLaunchCount():
    return int(_lib.lib().ncb200_kernel_launch_count())
