"""ctypes binding of libncrystal_b200.so (the CUDA library).

Mirrors how the reference's Python layer binds its C-API
(ref: ncrystal_python/src/NCrystal/_chooks.py:499-578): same C symbols, same
argument conventions.  There is no CPU fallback: if the CUDA library is not
built, importing anything that needs it raises immediately.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NCB200_LIB", os.path.join(_HERE, "lib", "libncrystal_b200.so"))
DATA_DIR = os.path.join(_HERE, "data")


class ncrystal_scatter_t(C.Structure):
    _fields_ = [("internal", C.c_void_p)]


class ncrystal_process_t(C.Structure):
    _fields_ = [("internal", C.c_void_p)]


class ncrystal_absorption_t(C.Structure):
    _fields_ = [("internal", C.c_void_p)]


_dblp = C.POINTER(C.c_double)
_u64 = C.c_uint64
_ulong = C.c_ulong
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/ncrystal_b200.h declares
SIGNATURES = {
    "ncrystal_refcount": (C.c_int, [_vp]),
    "ncrystal_ref": (None, [_vp]),
    "ncrystal_unref": (None, [_vp]),
    "ncrystal_valid": (C.c_int, [_vp]),
    "ncrystal_invalidate": (None, [_vp]),
    "ncrystal_cast_scat2proc": (ncrystal_process_t, [ncrystal_scatter_t]),
    "ncrystal_cast_proc2scat": (ncrystal_scatter_t, [ncrystal_process_t]),
    "ncrystal_create_scatter": (ncrystal_scatter_t, [C.c_char_p]),
    "ncrystal_create_scatter_builtinrng": (ncrystal_scatter_t, [C.c_char_p, _ulong]),
    "ncrystal_clone_scatter": (ncrystal_scatter_t, [ncrystal_scatter_t]),
    "ncrystal_clone_scatter_rngbyidx": (ncrystal_scatter_t, [ncrystal_scatter_t, _ulong]),
    "ncrystal_clone_scatter_rngforcurrentthread": (ncrystal_scatter_t, [ncrystal_scatter_t]),
    "ncrystal_name": (C.c_char_p, [ncrystal_process_t]),
    "ncrystal_isnonoriented": (C.c_int, [ncrystal_process_t]),
    "ncrystal_domain": (None, [ncrystal_process_t, _dblp, _dblp]),
    "ncrystal_crosssection_nonoriented": (None, [ncrystal_process_t, C.c_double, _dblp]),
    "ncrystal_crosssection": (None, [ncrystal_process_t, C.c_double, C.POINTER(C.c_double * 3), _dblp]),
    "ncrystal_samplescatterisotropic": (None, [ncrystal_scatter_t, C.c_double, _dblp, _dblp]),
    "ncrystal_samplescatter": (None, [ncrystal_scatter_t, C.c_double, C.POINTER(C.c_double * 3), _dblp,
                                      C.POINTER(C.c_double * 3)]),
    "ncrystal_genscatter_nonoriented": (None, [ncrystal_scatter_t, C.c_double, _dblp, _dblp]),
    "ncrystal_genscatter_nonoriented_many": (None, [ncrystal_scatter_t, _dblp, _ulong, _ulong, _dblp, _dblp]),
    "ncrystal_genscatter": (None, [ncrystal_scatter_t, C.c_double, C.POINTER(C.c_double * 3), C.POINTER(C.c_double * 3), _dblp]),
    "ncrystal_genscatter_many": (None, [ncrystal_scatter_t, C.c_double, C.POINTER(C.c_double * 3), _ulong,
                                        _dblp, _dblp, _dblp, _dblp]),
    "ncrystal_clone_absorption": (ncrystal_absorption_t, [ncrystal_absorption_t]),
    "ncrystal_process_uid": (C.c_void_p, [ncrystal_process_t]),
    "ncrystal_version": (C.c_int, []),
    "ncrystal_version_str": (C.c_char_p, []),
    "ncrystal_namespace": (C.c_char_p, []),
    "ncrystal_dealloc_doubleptr": (None, [_dblp]),
    "ncrystal_runmmcsim_stdengine": (None, [C.c_uint, C.c_uint, C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p]),
    "ncrystal_setbuiltinrandgen_withstate": (None, [C.c_char_p]),
    # caller-supplied generator double(*)(void*) + its state (ncrystal.h:792)
    "ncrystal_samplescatter_rs": (None, [C.CFUNCTYPE(C.c_double, C.c_void_p), C.c_void_p, ncrystal_scatter_t, C.c_double,
                                         C.POINTER(C.c_double * 3), _dblp, C.POINTER(C.c_double * 3)]),
    # OpenMC's C++ boundary (include/ncrystal_b200_virtapi.hh); listed so that the export is checked
    "ncrystal_access_virtual_api": (C.c_void_p, [C.c_uint]),
    "ncrystal_crosssection_nonoriented_many": (None, [ncrystal_process_t, _dblp, _ulong, _ulong, _dblp]),
    "ncrystal_samplescatterisotropic_many": (None, [ncrystal_scatter_t, _dblp, _ulong, _ulong, _dblp, _dblp]),
    "ncrystal_samplescatter_many": (None, [ncrystal_scatter_t, C.c_double, C.POINTER(C.c_double * 3), _ulong,
                                           _dblp, _dblp, _dblp, _dblp]),
    "ncrystal_error": (C.c_int, []),
    "ncrystal_lasterror": (C.c_char_p, []),
    "ncrystal_lasterrortype": (C.c_char_p, []),
    "ncrystal_clearerror": (None, []),
    "ncrystal_setquietonerror": (C.c_int, [C.c_int]),
    "ncrystal_sethaltonerror": (C.c_int, [C.c_int]),
    "ncrystal_seterrhandler": (None, [_vp]),
    "ncrystal_setmsghandler": (None, [_vp]),
    "ncb200_emit_message": (None, [C.c_char_p, C.c_uint]),
    "ncrystal_setrandgen": (None, [_vp]),
    "ncrystal_setbuiltinrandgen": (None, []),
    "ncrystal_setbuiltinrandgen_withseed": (None, [_ulong]),
    "ncrystal_rngsupportsstatemanip_ofscatter": (C.c_int, [ncrystal_scatter_t]),
    "ncrystal_getrngstate_ofscatter": (_vp, [ncrystal_scatter_t]),
    "ncrystal_setrngstate_ofscatter": (None, [ncrystal_scatter_t, C.c_char_p]),
    "ncrystal_dealloc_string": (None, [_vp]),
    "ncb200_create_scatter_from_blob": (ncrystal_scatter_t, [C.c_char_p, C.c_size_t, _ulong]),
    "ncb200_create_scatter_from_file": (ncrystal_scatter_t, [C.c_char_p, _ulong]),
    "ncb200_set_data_path": (None, [C.c_char_p]),
    "ncb200_cfg_to_filestem": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int]),
    "ncb200_set_rng_stream": (None, [ncrystal_scatter_t, _u64, C.c_uint32, _u64]),
    "ncb200_get_rng_stream": (None, [ncrystal_scatter_t, C.POINTER(_u64), C.POINTER(C.c_uint32), C.POINTER(_u64)]),
    "ncb200_crosssection_nonoriented_many_dev": (None, [ncrystal_process_t, _vp, _u64, _vp, _vp]),
    "ncb200_samplescatterisotropic_many_dev": (None, [ncrystal_scatter_t, _vp, _u64, _vp, _vp, _vp]),
    "ncb200_xs_and_samplescatterisotropic_many_dev": (None, [ncrystal_scatter_t, _vp, _u64, _vp, _vp, _vp, _vp]),
    "ncb200_crosssection_many": (None, [ncrystal_process_t, _dblp, _dblp, _dblp, _dblp, _u64, _dblp]),
    "ncb200_samplescatter_manydir": (None, [ncrystal_scatter_t, _dblp, _dblp, _dblp, _dblp, _u64,
                                            _dblp, _dblp, _dblp, _dblp]),
    "ncb200_crosssection_many_dev": (None, [ncrystal_process_t, _vp, _vp, _vp, _vp, _u64, _vp, _vp]),
    "ncb200_samplescatter_manydir_dev": (None, [ncrystal_scatter_t, _vp, _vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp]),
    "ncb200_check_device_errors": (C.c_int, [ncrystal_scatter_t, _vp]),
    "ncb200_set_diagnostics_dev": (None, [ncrystal_scatter_t, _vp, _vp]),
    "ncb200_generate_source_dev": (None, [_u64, _u64, _u64, C.c_double, C.c_double, _vp, _vp, _vp, _vp, _vp]),
    "ncb200_tally_hist_dev": (None, [_vp, _vp, _u64, C.c_double, C.c_double, C.c_uint32, _vp, _vp, _vp]),
    "ncb200_ncomponents": (C.c_int, [ncrystal_process_t]),
    "ncb200_component_kind": (C.c_int, [ncrystal_process_t, C.c_int]),
    "ncb200_component_scale": (C.c_double, [ncrystal_process_t, C.c_int]),
    "ncb200_kernel_launch_count": (_u64, []),
    "ncb200_table_bytes": (_u64, [ncrystal_process_t]),
    "ncrystal_wl2ekin": (C.c_double, [C.c_double]),
    "ncrystal_ekin2wl": (C.c_double, [C.c_double]),
    "ncrystal_cast_abs2proc": (ncrystal_process_t, [ncrystal_absorption_t]),
    "ncrystal_cast_proc2abs": (ncrystal_absorption_t, [ncrystal_process_t]),
    "ncrystal_create_absorption": (ncrystal_absorption_t, [C.c_char_p]),
    "ncb200_create_absorption_from_blob": (ncrystal_absorption_t, [_vp, C.c_size_t]),
    "ncb200_minimc_run": (C.c_void_p, [ncrystal_scatter_t, C.c_char_p, C.c_char_p, C.c_char_p]),
    "ncb200_minimc_run_slice": (C.c_void_p, [ncrystal_scatter_t, C.c_char_p, C.c_char_p, C.c_char_p, _u64, _u64]),
    "ncb200_material_bulk": (None, [ncrystal_process_t, _dblp, _dblp, _dblp]),
    "ncb200_fp64_fma_probe": (C.c_double, []),
    "ncb200_set_fg_staged_min": (None, [C.c_uint64]),
    "ncb200_set_mmc_tail_mode": (None, [C.c_int]),
    "ncb200_xs_and_samplescatterisotropic_many": (None, [ncrystal_scatter_t, _dblp, C.c_uint64, _dblp, _dblp, _dblp]),
    "ncb200_kernel_timing": (None, [C.c_int]),
    "ncb200_kernel_timing_report": (C.c_int, [C.c_char_p, C.c_int]),
    "ncb200_last_queue_counts": (C.c_int, [ncrystal_scatter_t, C.POINTER(C.c_uint32)]),
    "ncb200_version": (C.c_char_p, []),
    "ncb200_sab_xscheck": (C.c_int, [ncrystal_process_t, C.c_int, _dblp, C.c_int]),
    "ncb200_sab_energy_grid": (C.c_int, [ncrystal_process_t, C.c_int, _dblp, _dblp, C.c_int, _dblp]),
    "ncb200_sab_sampler_dump": (C.c_int, [ncrystal_process_t, C.c_int, C.c_int, _dblp, _dblp, _dblp, _dblp, _dblp]),
    "ncb200_sab_selfcheck": (C.c_long, [ncrystal_process_t, C.c_int]),
    "ncb200_pin_host_buffer": (C.c_int, [_vp, _u64]),
    "ncb200_unpin_host_buffer": (C.c_int, [_vp]),
    "ncb200_set_devices": (C.c_int, [C.c_int]),
    "ncb200_get_devices": (C.c_int, []),
    "ncb200_set_fanout_min": (None, [_u64]),
    "ncb200_tally_hist_many": (None, [_dblp, _dblp, _u64, C.c_double, C.c_double, C.c_uint32, _dblp, _dblp]),
    "ncrystal_raw_vdos2gn": (None, [_dblp, _dblp, C.c_uint, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_uint,
                                    _dblp, _dblp, C.POINTER(C.c_uint), C.POINTER(_dblp)]),
    "ncrystal_raw_vdos2kernel": (None, [_dblp, _dblp, C.c_uint, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_uint,
                                        _vp, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(_dblp),
                                        C.POINTER(_dblp), C.POINTER(_dblp), C.c_double, _dblp]),
    "ncrystal_raw_vdos2knl": (None, [_dblp, _dblp, C.c_uint, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_uint,
                                     _vp, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(_dblp),
                                     C.POINTER(_dblp), C.POINTER(_dblp)]),
    "ncb200_vdos_expansion_count": (C.c_ulong, []),
}

_lib = None


def lib():
    """Load the CUDA library (once).  Raises if it has not been built: there is no other backend."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "ncrystal_b200: CUDA library %s not found. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
                "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            f.restype = res
            f.argtypes = args
        L.ncb200_set_data_path(os.environ.get("NCB200_DATA_PATH", DATA_DIR).encode())
        # Python callers get exceptions instead of exit(1) (same as the reference's python layer)
        L.ncrystal_sethaltonerror(0)
        L.ncrystal_setquietonerror(1)
        _lib = L
    return _lib
