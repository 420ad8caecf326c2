"""CPU tests of the compiled-material format and of the C-ABI library's exported surface."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HDR_FMT = "<QIIIIQdd4d176s"
COMP_FMT = "<IIdddQQ"


def parse_header(blob):
    magic, version, ncomp, oriented, _r, nbytes, dlo, dhi, numdens, abs_c, temp, _r3, cfg = struct.unpack_from(HDR_FMT, blob, 0)
    off = struct.calcsize(HDR_FMT)
    comps = []
    for i in range(8):
        kind, _r2, scale, clo, chi, coff, cn = struct.unpack_from(COMP_FMT, blob, off + i * struct.calcsize(COMP_FMT))
        if i < ncomp:
            comps.append(dict(kind=kind, scale=scale, dom=(clo, chi), off=coff, nbytes=cn))
    return dict(magic=magic, version=version, ncomp=ncomp, oriented=oriented, nbytes=nbytes, dom=(dlo, dhi), numdens=numdens, abs_c=abs_c,
                temperature=temp,
                cfg=cfg.split(b"\0")[0].decode(), comps=comps)


def sab_grids(blob, icomp):
    h = parse_header(blob)
    c = h["comps"][icomp]
    assert c["kind"] == 3
    off = c["off"]
    vals = struct.unpack_from("<14d4Q", blob, off)
    negrid = vals[14]
    base = off + 14 * 8 + 4 * 8
    egrid = np.frombuffer(blob, dtype=np.float64, count=negrid, offset=base)
    xs = np.frombuffer(blob, dtype=np.float64, count=negrid, offset=base + 8 * negrid)
    return egrid, xs


def test_header_declares_only_exported_symbols():
    """Every function include/ncrystal_b200.h declares is exported by the built library
    (no compute calls here: there is no GPU on this box)."""
    hdr = open(os.path.join(ROOT, "include", "ncrystal_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(ncrystal_[a-z0-9_]+|ncb200_[a-z0-9_]+)\s*\(", hdr))
    names -= {"ncrystal_process_t", "ncrystal_scatter_t"}
    assert len(names) > 50
    from ncrystal_b200 import _lib
    L = _lib.lib()
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    assert set(_lib.SIGNATURES) == names, sorted(set(_lib.SIGNATURES) ^ names)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ncrystal_b200 as nc
    with pytest.raises(nc.NCException) as ei:
        nc.Scatter.fromBlob(b"\0" * 1024)
    assert "CUDA device" in str(ei.value) or "magic" in str(ei.value)


def test_error_state_machinery():
    from ncrystal_b200 import _lib
    L = _lib.lib()
    L.ncrystal_clearerror()
    assert L.ncrystal_error() == 0
    h = _lib.ncrystal_scatter_t(None)
    p = L.ncrystal_cast_scat2proc(h)   # invalid handle -> error state, null result
    assert L.ncrystal_error() == 1 and not p.internal
    assert b"Invalid" in L.ncrystal_lasterror()
    assert L.ncrystal_lasterrortype() == b"LogicError"
    L.ncrystal_clearerror()
    assert L.ncrystal_error() == 0 and L.ncrystal_lasterror() is None
    out = np.zeros(6)
    e = np.ones(3)
    dp = C.POINTER(C.c_double)
    L.ncrystal_crosssection_nonoriented_many(_lib.ncrystal_process_t(None), e.ctypes.data_as(dp), 3, 2, out.ctypes.data_as(dp))
    assert L.ncrystal_error() == 1 and np.all(out == -1.0)   # sentinel fill, ref: ncrystal.cc:1136-1140
    L.ncrystal_clearerror()
    eo = np.zeros(3)
    mu = np.zeros(3)
    L.ncrystal_samplescatterisotropic_many(_lib.ncrystal_scatter_t(None), e.ctypes.data_as(dp), 3, 1,
                                           eo.ctypes.data_as(dp), mu.ctypes.data_as(dp))
    assert np.all(eo == -1.0) and np.all(mu == -999.0)       # ref: ncrystal.cc:1239-1245
    L.ncrystal_clearerror()
    L.ncrystal_setrandgen(None)
    assert L.ncrystal_error() == 1
    L.ncrystal_clearerror()


def test_cfg_filestem():
    from ncrystal_b200 import _lib
    L = _lib.lib()
    buf = C.create_string_buffer(256)
    L.ncb200_cfg_to_filestem(b"Al_sg225.ncmat ; temp=293.15K", buf, 256)
    assert buf.value == b"Al_sg225.ncmat+temp=293.15K"


def test_blob_layout_al(configs):
    from oracle_check import material_path
    p = material_path(configs["Al"])
    if not os.path.exists(p):
        pytest.skip("compiled material not built")
    blob = open(p, "rb").read()
    h = parse_header(blob)
    assert h["magic"] == 0x0030303242434e and h["ncomp"] == 3 and h["oriented"] == 0
    assert [c["kind"] for c in h["comps"]] == [2, 1, 3]       # ElInc, PowderBragg, SAB (reference order)
    assert h["nbytes"] == len(blob)
    egrid, xs = sab_grids(blob, 2)
    assert egrid.size == 300 and np.all(np.diff(egrid) > 0) and abs(egrid[-1] - 5.0) < 1e-9


def _build_capi_caller(tmpdir):
    import subprocess
    exe = os.path.join(str(tmpdir), "capi_caller")
    libdir = os.path.join(ROOT, "ncrystal_b200", "lib")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "capi_caller.c"), "-o", exe, "-L", libdir, "-lncrystal_b200",
                           "-Wl,-rpath," + libdir])
    return exe


def test_plain_c_caller_compiles_and_links_against_the_header(tmp_path):
    # a C99 translation unit written against include/ncrystal_b200.h only (call sequence of the reference's
    # examples/ncrystal_example_c.c) builds warning-free and resolves every symbol from the library
    assert os.path.exists(_build_capi_caller(tmp_path))


def _build_capi_crng(tmpdir):
    import subprocess
    exe = os.path.join(str(tmpdir), "capi_crng")
    libdir = os.path.join(ROOT, "ncrystal_b200", "lib")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "capi_crng.c"), "-o", exe, "-L", libdir, "-lncrystal_b200",
                           "-Wl,-rpath," + libdir])
    return exe


def test_caller_rng_program_compiles_and_links(tmp_path):
    assert os.path.exists(_build_capi_crng(tmp_path))


def _build_cxx_caller(tmpdir):
    import subprocess
    exe = os.path.join(str(tmpdir), "cxx_caller")
    libdir = os.path.join(ROOT, "ncrystal_b200", "lib")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cxx_caller.cc"), "-o", exe, "-L", libdir, "-lncrystal_b200",
                           "-Wl,-rpath," + libdir])
    return exe


def test_cxx_mirror_compiles_and_links(tmp_path):
    assert os.path.exists(_build_cxx_caller(tmp_path))


def _build_virtapi_caller(tmpdir):
    import subprocess
    exe = os.path.join(str(tmpdir), "virtapi_caller")
    libdir = os.path.join(ROOT, "ncrystal_b200", "lib")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-pthread", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "virtapi_caller.cc"), "-o", exe, "-L", libdir, "-lncrystal_b200",
                           "-Wl,-rpath," + libdir])
    return exe


def test_virtual_api_client_compiles_and_links(tmp_path):
    # a client of the OpenMC boundary (ncrystal_access_virtual_api, VirtAPI_Type1_v1) builds against the header
    assert os.path.exists(_build_virtapi_caller(tmp_path))


def test_virtual_api_vtable_layout_matches_reference_header():
    # the abstract class must declare the same virtual methods in the same order as the reference's header
    # (include/NCrystal/virtualapi/NCVirtAPI_Type1_v1.hh:79-91): that order IS the binary interface
    import re
    src = open(os.path.join(ROOT, "include", "ncrystal_b200_virtapi.hh")).read()
    src = re.sub(r"//[^\n]*", "", src)
    names = re.findall(r"virtual\s+[^;(]*?(~?\w+)\s*\(", src)
    assert names == ["createScatter", "cloneScatter", "deallocateScatter", "crossSectionUncached",
                     "sampleScatterUncached", "~VirtAPI_Type1_v1"]
    assert "interface_id = 1001" in src


def test_cfg_spelling_variants_find_the_same_compiled_material(configs):
    # "stdlib::" prefix and another parameter order resolve to the same .ncb file; an unknown material does not.
    # Without a GPU the call then stops at the device (LogicError / CUDA message), never at FileNotFound.
    from ncrystal_b200 import _lib
    from oracle_check import material_path
    if not os.path.exists(material_path(configs["Ge"])):
        pytest.skip("compiled materials not built")
    L = _lib.lib()
    old = L.ncrystal_sethaltonerror(0)
    L.ncrystal_setquietonerror(1)
    try:
        variants = ["stdlib::" + configs["Al"],
                    "Ge_sg227.ncmat;dir2=@crys_hkl:0,-1,1@lab:0,1,0;mos=40arcsec;dir1=@crys_hkl:5,1,1@lab:0,0,1"]
        for cfg in variants + ["no_such_material.ncmat;temp=5K"]:
            L.ncrystal_clearerror()
            h = L.ncrystal_create_scatter(cfg.encode())
            typ = L.ncrystal_lasterrortype() if L.ncrystal_error() else b""
            if h.internal:
                L.ncrystal_unref(C.byref(h))
            if cfg in variants:
                assert typ != b"FileNotFound", (cfg, L.ncrystal_lasterror())
            else:
                assert typ == b"FileNotFound"
    finally:
        L.ncrystal_clearerror()
        L.ncrystal_sethaltonerror(old)
        L.ncrystal_setquietonerror(0)


# ---- untrusted buffers: the loader (shared by the product and the host build) must reject truncated or corrupt
# compiled materials before copying anything (counts inside the payloads were trusted)
def _patched(blob, off, fmt, *vals):
    b = bytearray(blob)
    struct.pack_into(fmt, b, off, *vals)
    return bytes(b)


@pytest.mark.parametrize("key", ["Al", "Ge", "H2O"])
def test_loader_rejects_corrupt_blobs(configs, key):
    from oracle_check import material_path
    from _libs import HostSim
    p = material_path(configs[key])
    if not os.path.exists(p):
        pytest.skip("compiled material not built")
    blob = open(p, "rb").read()
    h = parse_header(blob)
    L = HostSim.lib()
    nbytes_off = struct.calcsize("<QIIII")
    comp0 = struct.calcsize(HDR_FMT)
    csz = struct.calcsize(COMP_FMT)
    bad = []
    # (a) truncated buffer: header claims more than is there, or a component reaches past the end
    bad.append(("short buffer", blob[: len(blob) // 2]))
    for i, c in enumerate(h["comps"]):
        cut = c["off"] + c["nbytes"] // 2
        bad.append(("comp %d truncated" % i, _patched(blob[:cut], nbytes_off, "<Q", cut)))
        # (b) component offset/size fields: out of bounds, wrapping, misaligned
        o = comp0 + i * csz + struct.calcsize("<IIddd")
        bad.append(("comp %d off past end" % i, _patched(blob, o, "<Q", len(blob) + 16)))
        bad.append(("comp %d off+nbytes wraps" % i, _patched(blob, o, "<QQ", 2**64 - 8, 64)))
        bad.append(("comp %d misaligned" % i, _patched(blob, o, "<Q", c["off"] + 4)))
        bad.append(("comp %d inside header" % i, _patched(blob, o, "<Q", 8)))
        # (c) array counts inside the payload larger than the payload
        if c["kind"] == 1:
            bad.append(("nplanes huge", _patched(blob, c["off"], "<Q", 2**40)))
            bad.append(("nplanes +1", _patched(blob, c["off"], "<Q", (c["nbytes"] - 16) // 16 + 1)))
        if c["kind"] == 2:
            bad.append(("nelem 13", _patched(blob, c["off"], "<Q", 13)))
        if c["kind"] == 3:
            g = c["off"] + 14 * 8
            ne, na, nb = struct.unpack_from("<3Q", blob, g)
            bad.append(("negrid 0", _patched(blob, g, "<Q", 0)))
            bad.append(("nalpha huge", _patched(blob, g + 8, "<Q", 2**33)))
            bad.append(("nbeta x2", _patched(blob, g + 16, "<Q", 2 * nb)))
            bad.append(("product wraps", _patched(blob, g + 8, "<QQ", 2**32, 2**32)))
        if c["kind"] == 5:
            g = c["off"] + 20 * 8
            nf, nn = struct.unpack_from("<2Q", blob, g)
            bad.append(("nnormals x4", _patched(blob, g + 8, "<Q", 4 * nn)))
            bad.append(("nfam huge", _patched(blob, g, "<Q", 2**50)))
            bad.append(("family index beyond normals", _patched(blob, c["off"] + 24 * 8 + 8 * (2 * nf + 1), "<d", 1e9)))
    for what, b in bad:
        hnd = L.hostsim_load(b, len(b))
        assert not hnd, "%s: %s was accepted" % (key, what)
        assert b"compiled material" in L.hostsim_lasterror() or b"too many" in L.hostsim_lasterror(), what
