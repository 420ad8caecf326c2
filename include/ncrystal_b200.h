/* ncrystal_b200.h -- C ABI of libncrystal_b200.so: NCrystal's batched cross-section
 * evaluation + scatter sampling on B200 (sm_100a).
 *
 * Part 1 re-declares, with IDENTICAL names and signatures, the entry points of the
 * reference C-API that sit on the hot path (ref: ncrystal_core/include/NCrystal/
 * cinterface/ncrystal.h, line numbers given per function), so a caller written
 * against NCrystal's C-API (ctypes/_chooks.py, McStas NCrystal_sample, Geant4
 * bindings, examples/ncrystal_example_c.c) binds unchanged.
 *
 * Part 2 (prefix ncb200_) adds what the reference ABI lacks for an accelerator:
 * handles from a compiled material, device-pointer variants (inputs/outputs
 * resident in HBM, caller's CUDA stream), batched ORIENTED calls with
 * per-neutron (E,dir) (modelled on the reference's experimental batch ABI,
 * ProcImpl::Process::evalManyXS, NCProcImpl.hh:136-140), a fused xs+sample call
 * (NCABIUtils.hh:78-100), RNG stream control and the tally histogram kernel.
 *
 * Plain C: pointers and sizes only.  All arrays are fp64, caller-owned, no aliasing.
 */
#ifndef NCRYSTAL_B200_H
#define NCRYSTAL_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ===================== Part 1: NCrystal C-API (same symbols) ===================== */

/* ncrystal.h:666-670 */
typedef struct { void * internal; } ncrystal_process_t;
typedef struct { void * internal; } ncrystal_scatter_t;
typedef struct { void * internal; } ncrystal_absorption_t;

/* ncrystal.h:672-676 -- all take the ADDRESS of a handle */
int  ncrystal_refcount( void* object );
void ncrystal_ref( void* object );
void ncrystal_unref( void* object );
int  ncrystal_valid( void* object );
void ncrystal_invalidate( void* object );

/* ncrystal.h:680,682 */
ncrystal_process_t ncrystal_cast_scat2proc( ncrystal_scatter_t );
ncrystal_scatter_t ncrystal_cast_proc2scat( ncrystal_process_t );

/* ncrystal.h:681,684 / :703 -- absorption: the 1/v process (AbsOOV, src/absoov/NCAbsOOV.cc) of the compiled material;
 * accepted by the cross-section entry points after ncrystal_cast_abs2proc, never by the sampling ones */
ncrystal_process_t ncrystal_cast_abs2proc( ncrystal_absorption_t );
ncrystal_absorption_t ncrystal_cast_proc2abs( ncrystal_process_t );
ncrystal_absorption_t ncrystal_create_absorption( const char * cfgstr );

/* ncrystal.h:699,708 -- cfgstr is resolved to a compiled material (see
 * ncb200_create_scatter_from_blob and INTEGRATION.md) */
ncrystal_scatter_t ncrystal_create_scatter( const char * cfgstr );
ncrystal_scatter_t ncrystal_create_scatter_builtinrng( const char * cfgstr, unsigned long seed );

/* ncrystal.h:718-731 -- clones share the immutable device tables and get an
 * independent random stream */
ncrystal_scatter_t ncrystal_clone_scatter( ncrystal_scatter_t );
ncrystal_scatter_t ncrystal_clone_scatter_rngbyidx( ncrystal_scatter_t, unsigned long rngstreamidx );
ncrystal_scatter_t ncrystal_clone_scatter_rngforcurrentthread( ncrystal_scatter_t );

/* ncrystal.h:759-773 */
const char * ncrystal_name( ncrystal_process_t );
int  ncrystal_isnonoriented( ncrystal_process_t );
void ncrystal_domain( ncrystal_process_t, double* ekin_low, double* ekin_high );

/* ncrystal.h:765,768 */
void ncrystal_crosssection_nonoriented( ncrystal_process_t, double ekin, double* result );
void ncrystal_crosssection( ncrystal_process_t, double ekin, const double (*direction)[3], double* result );

/* ncrystal.h:777,782 */
void ncrystal_samplescatterisotropic( ncrystal_scatter_t, double ekin, double* ekin_final, double* cos_scat_angle );
void ncrystal_samplescatter( ncrystal_scatter_t, double ekin, const double (*direction)[3],
                             double* ekin_final, double (*direction_final)[3] );

/* ncrystal.h:792 -- caller-supplied generator rngfct(rngstate): two numbers are drawn from it per call and key the
 * neutron's device stream (a host callback cannot run on the device; see ncrystal_b200_virtapi.hh) */
void ncrystal_samplescatter_rs( double (*rngfct)(void*), void* rngstate, ncrystal_scatter_t, double ekin,
                                const double (*direction)[3], double* ekin_final, double (*direction_final)[3] );

/* NCVirtAPIFactory.hh:45-52 -- OpenMC's boundary; C++ class layout in ncrystal_b200_virtapi.hh.  Returns the address
 * of a static std::shared_ptr<const VirtAPI_Type1_v1> for interface_id 1001, else NULL. */
void * ncrystal_access_virtual_api( unsigned interface_id );

/* ncrystal.h:1305 -- results[r*n_ekin+i] */
void ncrystal_crosssection_nonoriented_many( ncrystal_process_t, const double * ekin, unsigned long n_ekin,
                                             unsigned long repeat, double* results );
/* ncrystal.h:1289 -- results[r*n_ekin+i] */
void ncrystal_samplescatterisotropic_many( ncrystal_scatter_t, const double * ekin, unsigned long n_ekin,
                                           unsigned long repeat, double* results_ekin, double* results_cos_scat_angle );
/* ncrystal.h:1296 -- fixed (ekin,direction), `repeat` samples */
void ncrystal_samplescatter_many( ncrystal_scatter_t, double ekin, const double (*direction)[3], unsigned long repeat,
                                  double* results_ekin, double * results_dirx, double * results_diry, double * results_dirz );

/* ncrystal.h:1340-1368 -- the reference's obsolete "genscatter" entry points (older McStas / Geant4 bindings):
 * the same sampling reported as (scattering angle [rad], energy transfer) / (direction, energy transfer) */
void ncrystal_genscatter_nonoriented( ncrystal_scatter_t, double ekin, double* result_angle, double* result_dekin );
void ncrystal_genscatter_nonoriented_many( ncrystal_scatter_t, const double * ekin, unsigned long n_ekin, unsigned long repeat,
                                           double* results_angle, double* results_dekin );
void ncrystal_genscatter( ncrystal_scatter_t, double ekin, const double (*direction)[3],
                          double (*result_direction)[3], double* result_deltaekin );
void ncrystal_genscatter_many( ncrystal_scatter_t, double ekin, const double (*direction)[3], unsigned long repeat,
                               double * results_dirx, double * results_diry, double * results_dirz, double * results_dekin );

/* ncrystal.h:719 */
ncrystal_absorption_t ncrystal_clone_absorption( ncrystal_absorption_t );
/* ncrystal.h:1164 -- id of the underlying immutable process (same for clones); free with ncrystal_dealloc_string */
char * ncrystal_process_uid( ncrystal_process_t );
/* ncrystal.h:1229-1234 -- version of the NCrystal release whose hot path this library restates (4.4.2) */
int ncrystal_version(void);
const char * ncrystal_version_str(void);
const char * ncrystal_namespace(void);
/* ncrystal.h:1220, :1389 (obsolete in the reference too: raises an error) */
void ncrystal_dealloc_doubleptr( double* );
void ncrystal_runmmcsim_stdengine( unsigned, unsigned, const char *, const char *, const char *, char **, unsigned *,
                                   double **, double ** );

/* ncrystal.h:1147-1148 -- unit conversions (Aa <-> eV) */
double ncrystal_wl2ekin( double wl );
double ncrystal_ekin2wl( double ekin );

/* ncrystal.h:1030-1045 -- error state: global, message printed unless quiet, process
 * exit(1) unless ncrystal_sethaltonerror(0); non-halting calls fill outputs with
 * -1.0 (xs, ekin) / -999 (mu) / 0-vector (direction). */
int  ncrystal_error(void);
const char * ncrystal_lasterror(void);
const char * ncrystal_lasterrortype(void);
void ncrystal_clearerror(void);
int  ncrystal_setquietonerror( int );
int  ncrystal_sethaltonerror( int );
void ncrystal_seterrhandler( void (*handler)(char*,char*) );
/* ncrystal.h:1051 -- library output (warnings raised on the device included) goes to stdout unless a handler is set;
 * second argument: 0 info, 1 warning, 2 raw output.  NULL restores the default. */
void ncrystal_setmsghandler( void (*handler)(const char*,unsigned) );
/* extension (test hook): send a message through the handler */
void ncb200_emit_message( const char* msg, unsigned msgtype );

/* ncrystal.h:1067-1087 -- RNG control.  Host callbacks cannot be honoured on the
 * device: ncrystal_setrandgen raises an error.  The ncrystal_setbuiltinrandgen* calls set the seed (and restart the
 * stream numbering) of the scatter handles that ncrystal_create_scatter makes afterwards.  State strings serialise
 * (seed, stream id, next neutron index). */
void ncrystal_setrandgen( double (*rg)(void) );
void ncrystal_setbuiltinrandgen(void);
void ncrystal_setbuiltinrandgen_withseed( unsigned long seed );
void ncrystal_setbuiltinrandgen_withstate( const char* state );   /* a string from ncrystal_getrngstate_ofscatter */
int  ncrystal_rngsupportsstatemanip_ofscatter( ncrystal_scatter_t );
char* ncrystal_getrngstate_ofscatter( ncrystal_scatter_t );  /* free with ncrystal_dealloc_string */
void ncrystal_setrngstate_ofscatter( ncrystal_scatter_t, const char* );
void ncrystal_dealloc_string( char* );

/* ===================== Part 2: B200 extensions ===================== */

/* Handle from a compiled material (ncrystal_b200/csrc/ncb_blob.h), the call the
 * reference-side binding makes after flattening its ProcComposition.  The tables
 * are uploaded to the CURRENT CUDA device; S(alpha,beta) sampler tables are built
 * there.  Returns {NULL} on error. */
ncrystal_scatter_t ncb200_create_scatter_from_blob( const void* blob, size_t nbytes, unsigned long seed );
ncrystal_scatter_t ncb200_create_scatter_from_file( const char* path, unsigned long seed );
ncrystal_absorption_t ncb200_create_absorption_from_blob( const void* blob, size_t nbytes );
/* Directory list (':'-separated) searched by ncrystal_create_scatter for
 * "<sanitised cfgstr>.ncb"; default: $NCB200_DATA_PATH then <libdir>/../data. */
void ncb200_set_data_path( const char* path );
/* writes the sanitised file stem for a cfg string into buf (returns needed length) */
int  ncb200_cfg_to_filestem( const char* cfgstr, char* buf, int buflen );

/* Random stream of a scatter handle: neutron j of the next sampling call draws
 * from Philox stream (seed, stream_id, next_index + j). */
void ncb200_set_rng_stream( ncrystal_scatter_t, uint64_t seed, uint32_t stream_id, uint64_t next_index );
void ncb200_get_rng_stream( ncrystal_scatter_t, uint64_t* seed, uint32_t* stream_id, uint64_t* next_index );

/* Device-pointer variants: all arrays are device pointers on the handle's device;
 * `stream` is a cudaStream_t (NULL = default stream).  Asynchronous: no host
 * synchronisation; errors raised by the device are collected with
 * ncb200_check_device_errors. */
void ncb200_crosssection_nonoriented_many_dev( ncrystal_process_t, const double* d_ekin, uint64_t n,
                                               double* d_results, void* stream );
void ncb200_samplescatterisotropic_many_dev( ncrystal_scatter_t, const double* d_ekin, uint64_t n,
                                             double* d_ekin_final, double* d_mu, void* stream );
/* fused: also returns the total cross section (d_xs may be NULL) */
void ncb200_xs_and_samplescatterisotropic_many_dev( ncrystal_scatter_t, const double* d_ekin, uint64_t n,
                                                    double* d_xs, double* d_ekin_final, double* d_mu, void* stream );

/* the fused call on HOST arrays: one pass, 8 B in + 24 B out per neutron over the bus (the two separate *_many calls
 * move 16 B in + 24 B out); same results as those two calls */
void ncb200_xs_and_samplescatterisotropic_many( ncrystal_scatter_t, const double* ekin, uint64_t n,
                                                double* results_xs, double* results_ekin, double* results_cos_scat_angle );

/* Batched oriented calls, per-neutron (E,dir), SoA (host pointers / device pointers). */
void ncb200_crosssection_many( ncrystal_process_t, const double* ekin, const double* ux, const double* uy, const double* uz,
                               uint64_t n, double* results );
void ncb200_samplescatter_manydir( ncrystal_scatter_t, const double* ekin, const double* ux, const double* uy, const double* uz,
                                   uint64_t n, double* ekin_final, double* ox, double* oy, double* oz );
void ncb200_crosssection_many_dev( ncrystal_process_t, const double* d_ekin, const double* d_ux, const double* d_uy,
                                   const double* d_uz, uint64_t n, double* d_results, void* stream );
void ncb200_samplescatter_manydir_dev( ncrystal_scatter_t, const double* d_ekin, const double* d_ux, const double* d_uy,
                                       const double* d_uz, uint64_t n, double* d_ekin_final,
                                       double* d_ox, double* d_oy, double* d_oz, void* stream );

/* Synchronises `stream`, then reports (through the ncrystal_error machinery) any
 * error raised on the device since the last check.  Returns the raw flag word
 * (0 = none; bits: ncb::SampleErr in csrc/ncb_phys_sab.cuh). */
int  ncb200_check_device_errors( ncrystal_scatter_t, void* stream );

/* Per-neutron diagnostics of the NEXT sampling call on this handle (device pointers,
 * may be NULL): number of uniforms consumed, chosen component index. */
void ncb200_set_diagnostics_dev( ncrystal_scatter_t, uint32_t* d_ndraws, int32_t* d_component );

/* Synthetic source used by the benchmarks: ekin[i] = 10^(log10(lo)+(log10(hi)-log10(lo))*u_i),
 * u_i = first uniform of Philox stream (seed, 0xE0, first_index+i); isotropic directions
 * z=2u-1, phi=2*pi*u' from the next two uniforms (d_u* may be NULL). */
void ncb200_generate_source_dev( uint64_t seed, uint64_t first_index, uint64_t n, double lo, double hi,
                                 double* d_ekin, double* d_ux, double* d_uy, double* d_uz, void* stream );

/* Tally: weighted 1D histogram with under/overflow bins (d_hist, d_sumw2: nbins+2 doubles,
 * ACCUMULATED into; d_weights/d_sumw2 may be NULL).  Analogue of the reference's
 * Hist1D filling in MiniMC tallies (NCHists.hh); merging across GPUs is an
 * element-wise sum (NCCL all-reduce), like Tally::merge (NCMMC_Tally.hh:40-62). */
void ncb200_tally_hist_dev( const double* d_values, const double* d_weights, uint64_t n,
                            double lo, double hi, uint32_t nbins, double* d_hist, double* d_sumw2, void* stream );

/* Host buffers.  The *_many entry points take caller-owned host arrays.  Page-locked arrays are copied directly;
 * pageable (malloc'd) arrays of calls with >= 2^16 neutrons go through a pinned bounce ring filled / drained by a
 * small pool of host threads (bounded by the host's memcpy bandwidth: measured 1.0e9 neutrons/s against 1.6e9 from
 * pinned arrays).  A caller that reuses its arrays can page-lock them once with ncb200_pin_host_buffer (about 20 ms
 * per 80 MB) and unpin them before freeing.  0 on success, -1 on error. */
int      ncb200_pin_host_buffer( void* p, uint64_t nbytes );
int      ncb200_unpin_host_buffer( void* p );

/* Several GPUs from ONE process (the single-process shape of an OpenMC / McStas caller).  After
 * ncb200_set_devices(n) (0 = all visible devices; 1 = off) the host-pointer batch entry points (ncrystal_*_many,
 * ncb200_crosssection_many, ncb200_samplescatter_manydir, ncb200_xs_and_samplescatterisotropic_many) split a batch
 * of at least ncb200_set_fanout_min() neutrons (default 2^20) into contiguous slices, one per device: tables are
 * replicated per device, streams are keyed by the global neutron index, results are bit-identical to the one-device
 * result.  ncb200_tally_hist_many histograms host arrays on the devices and merges with ncclAllReduce(sum, fp64) --
 * the analogue of the reference's worker threads + Tally::merge (src/minimc/NCMMC_SimMgr.cc:167-263,
 * NCMMC_Tally.hh:40-62).  hist[nbins+2] (underflow, bins, overflow) and sumw2 (may be null) are ADDED to.
 * Returns the number of devices now in use, -1 on error. */
int      ncb200_set_devices( int n );
int      ncb200_get_devices( void );
void     ncb200_set_fanout_min( uint64_t n );
void     ncb200_tally_hist_many( const double* values, const double* weights, uint64_t n,
                                 double lo, double hi, uint32_t nbins, double* hist, double* sumw2 );

/* Introspection */
int      ncb200_ncomponents( ncrystal_process_t );
int      ncb200_component_kind( ncrystal_process_t, int i ); /* enum ncb_kind */
double   ncb200_component_scale( ncrystal_process_t, int i );
uint64_t ncb200_kernel_launch_count(void);   /* kernels launched by this library so far */
uint64_t ncb200_table_bytes( ncrystal_process_t ); /* HBM footprint of the material tables */
/* Device-resident transport step ("MiniMC" on the device): the reference's ["mmc","run",CFGSTR,GEOMCFG,SRCCFG,
 * ENGINECFG] query of ncrystal_jsonquery (ncrystal.h:1253; NCMMC_Query.cc:38-56) for the material of a scatter handle.
 * Same cfg-string vocabulary and result JSON ("NCrystalMiniMCResults_v1"); supported subset documented in
 * csrc/ncb_lib_mmc.inc.  Returns NULL on error; free the string with ncrystal_dealloc_string.
 * _slice simulates source neutrons [first, first+count) only -- tallies of disjoint slices add up (multi-GPU). */
char*    ncb200_minimc_run( ncrystal_scatter_t, const char* geomcfg, const char* srccfg, const char* enginecfg );
char*    ncb200_minimc_run_slice( ncrystal_scatter_t, const char* geomcfg, const char* srccfg, const char* enginecfg,
                                  uint64_t first, uint64_t count );
/* bulk quantities of the compiled material: number density [atoms/Aa^3], AbsOOV constant xs_abs*sqrt(E)
 * [barn*sqrt(eV)] (NCAbsOOV.cc:33-45), temperature [K] */
void     ncb200_material_bulk( ncrystal_process_t, double* numdens, double* abs_c, double* temperature );
/* Per-kernel timing with CUDA events on the launching stream (off by default).  enable!=0 starts a fresh
 * recording; the report is a JSON object {"kernel": {"launches": n, "ms_avg": t}, ...} (returns its length). */
void     ncb200_kernel_timing( int enable );
int      ncb200_kernel_timing_report( char* buf, int buflen );
/* Batches of at least nmin neutrons sample their free-gas queue with the staged kernels (k_fg_prep ... k_fg_finish),
 * smaller ones with the single neutron-per-lane kernel; identical results.  Default 4e6 ($NCB200_FG_STAGED_MIN). */
void     ncb200_set_fg_staged_min( uint64_t nmin );
/* test hook: 1 (default) = the tail of a transport run (<= 32 Ki live neutrons) is finished by the one-launch
 * warp-per-history kernel, 0 = by groups of 16 steps of the multi-kernel sequence.  Same tallies either way. */
void     ncb200_set_mmc_tail_mode( int persistent );
/* sizes of the work queues of the most recent isotropic sampling launch: table path, free-gas path, table at Emax */
int      ncb200_last_queue_counts( ncrystal_scatter_t, uint32_t* out3 );
const char* ncb200_version(void);
/* Measured vector-FP64 FMA rate of the current device in TFLOP/s (8 independent DFMA chains per thread, all SMs): the
 * denominator for "FP64 pipe utilisation against peak" in profiles/ (not a product function). */
double   ncb200_fp64_fma_probe(void);
/* SAB table builder check: per-energy-point total xs recomputed on the device while
 * building the sampler tables (compare with the xs grid of the compiled material). */
int      ncb200_sab_xscheck( ncrystal_process_t, int component, double* out, int nmax );
/* energy grid / grid cross sections / {k_extension, k1, k2, egrid_margin} of a S(alpha,beta) leaf: SABXSProvider::m_egrid,
 * m_xs, m_kExtension and SABSampler::m_k1, m_k2, m_egridMargin (NCSABXSProvider.cc:35-52, NCSABSampler.cc:41-57), as
 * delivered in the compiled material or as determined by the library itself (ncb_blob.h: ncb_sab_t::auto_egrid,
 * restating NCSABIntegrator.cc:147-283).  Returns the number of grid points. */
int      ncb200_sab_energy_grid( ncrystal_process_t, int component, double* egrid, double* xs, int nmax, double* consts4 );
/* Copy of built sampler tables for one energy point (layout as oracle refdrv_sab_sampler_dump) */
int      ncb200_sab_sampler_dump( ncrystal_process_t, int component, int iE, double* x, double* pdf, double* cdf,
                                  double* infos, double* meta );
/* Consistency of the gather-friendly copies of the sampler tables and of the log guide with the tables they were
 * derived from, all read back from the device: number of violations (0 = consistent), -1 on error. */
long     ncb200_sab_selfcheck( ncrystal_process_t, int component );

/* ---- VDOS -> S(alpha,beta): the reference's own entry points for expanding a phonon density of states into a
 * scattering kernel with Sjolander's method (ref: include/NCrystal/cinterface/ncrystal.h:885-925,
 * src/cinterface/ncrystal.cc:755-883), same signatures and ownership (arrays are freed with
 * ncrystal_dealloc_doubleptr).  The convolutions that produce the phonon-order spectra G_n and the sum over orders run
 * on the device; the tables equal the reference's bit for bit. */
void ncrystal_raw_vdos2gn( const double* vdos_egrid, const double* vdos_density, unsigned vdos_egrid_npts,
                           unsigned vdos_density_npts, double scattering_xs, double mass_amu, double temperature,
                           unsigned nvalue, double* res_gn_xmin, double* res_gn_xmax, unsigned* res_gn_npts,
                           double** res_gn_vals );
void ncrystal_raw_vdos2kernel( const double* vdos_egrid, const double* vdos_density, unsigned vdos_egrid_npts,
                               unsigned vdos_density_npts, double scattering_xs, double mass_amu, double temperature,
                               unsigned vdoslux, double (*order_weight_fct)( unsigned order ),
                               unsigned* nalpha, unsigned* nbeta, double** alpha, double** beta, double** sab,
                               double target_emax, double* suggested_emax );
void ncrystal_raw_vdos2knl( const double* vdos_egrid, const double* vdos_density, unsigned vdos_egrid_npts,
                            unsigned vdos_density_npts, double scattering_xs, double mass_amu, double temperature,
                            unsigned vdoslux, double (*order_weight_fct)( unsigned order ),
                            unsigned* nalpha, unsigned* nbeta, double** alpha, double** beta, double** sab ); /* obsolete spelling */
/* number of expansions run so far (raw calls and VDOS leaves of compiled materials, ncb_blob.h: NCB_KIND_SABVDOS) */
unsigned long ncb200_vdos_expansion_count( void );

#ifdef __cplusplus
}
#endif
#endif
