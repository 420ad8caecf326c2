/* ncb_blob.h -- "compiled material" interchange format (plain C, no dependencies).
 *
 * A compiled material is the flat, POD image of what NCrystal's scatter factory
 * hands to the hot path: the ordered list of (scale, leaf process) pairs of a
 * ProcComposition (ref: ncrystal_core/include/NCrystal/interfaces/NCProcImpl.hh:310-373,
 * flattened by ProcComposition::consumeAndCombine, src/interfaces/NCProcImpl.cc:410-454)
 * with, per leaf, exactly the immutable tables its crossSection/sampleScatter
 * methods read.  Producing it (NCMAT parsing, HKL lists) is the reference's setup
 * pipeline and is out of scope here; the reference-side producer is
 * bridge/matcompile_impl.icc (the binding a maintainer would add, see
 * INTEGRATION.md).  Everything derived that the hot path needs beyond these
 * inputs is built natively on the device at load time: the S(alpha,beta) sampler
 * tables and energy grids (csrc/ncb_sabbuild.cuh, ncb_sabgrid.h) and, for leaves
 * delivered as a phonon density of states, S(alpha,beta) itself (csrc/ncb_vdos.h).
 *
 * Layout: one contiguous little-endian buffer.  ncb_header_t at offset 0, then
 * one payload per component at comp[i].off (byte offset from the start of the
 * buffer, 16-byte aligned).  Arrays follow their payload struct directly in the
 * order documented below; all arrays are fp64.
 */
#ifndef NCB_BLOB_H
#define NCB_BLOB_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NCB_MAGIC    0x0030303242434eULL /* "NCB200\0" */
#define NCB_VERSION  3u
#define NCB_MAXCOMP  8

enum ncb_kind {
  NCB_KIND_POWDERBRAGG = 1, /* ref: src/powderbragg/NCPowderBragg.cc */
  NCB_KIND_ELINC       = 2, /* ref: src/elincscatter/NCElIncScatter.cc + src/phys_utils/NCElIncXS.cc */
  NCB_KIND_SAB         = 3, /* ref: src/sabscatter/NCSABScatter.cc + src/sab */
  NCB_KIND_FREEGAS     = 4, /* ref: src/freegas/NCFreeGas.cc + src/phys_utils/NCFreeGasUtils.cc */
  NCB_KIND_SCBRAGG     = 5, /* ref: src/scbragg/NCSCBragg.cc + src/phys_utils/NCGaussMos.cc */
  NCB_KIND_LCBRAGG     = 7, /* ref: src/lcbragg/NCLCBragg.cc + src/extd_utils/NCLCUtils.cc (6 is taken by the
                               loader-internal absorption kind) */
  NCB_KIND_SABVDOS     = 8  /* a SABScatter leaf delivered as the phonon density of states it derives from; the
                               library runs the VDOS -> S(alpha,beta) expansion itself (ref: src/vdos/) and then
                               treats the leaf as NCB_KIND_SAB with auto_egrid = 1 */
};

typedef struct {
  uint32_t kind;       /* enum ncb_kind */
  uint32_t reserved;
  double   scale;      /* ProcComposition::Component::scale */
  double   dom_lo;     /* leaf Process::domain() (EnergyDomain, NCTypes.hh:427-439) */
  double   dom_hi;
  uint64_t off;        /* byte offset of the payload */
  uint64_t nbytes;     /* payload size incl. arrays */
} ncb_comp_t;

typedef struct {
  uint64_t   magic;
  uint32_t   version;
  uint32_t   ncomp;
  uint32_t   oriented; /* 1 if MaterialType::Anisotropic */
  uint32_t   reserved;
  uint64_t   nbytes;   /* total size of the buffer */
  double     dom_lo;   /* ProcComposition::domain() */
  double     dom_hi;
  /* bulk quantities the transport step needs next to the scatter process (MiniMC MatDef,
   * ref: include/NCrystal/internal/minimc/NCMMC_Defs.hh; absorption = AbsOOV, src/absoov/NCAbsOOV.cc:33-45) */
  double     numdens;     /* Info::getNumberDensity() [atoms/Aa^3] */
  double     abs_c;       /* AbsOOV::m_c = sigma_abs(2200m/s)*sqrt(E_2200): xs_abs(E) = abs_c/sqrt(E) [barn*sqrt(eV)];
                             0 = no absorption, <0 = absorption is not a 1/v process (unsupported) */
  double     temperature; /* Info::getTemperature() [K], -1 if not available */
  double     reserved2;
  char       cfg[176]; /* the cfg-string the material was compiled from (NUL terminated) */
  ncb_comp_t comp[NCB_MAXCOMP];
} ncb_header_t;

/* PowderBragg: followed by e2d[nplanes] (= m_2dE, ascending energies wl2ekin(2d)),
 * fdm[nplanes] (= m_fdm_commul).  NCPowderBragg.cc:68-108. */
typedef struct {
  uint64_t nplanes;
  double   threshold;  /* m_threshold = e2d[0] */
} ncb_powderbragg_t;

/* ElIncScatter: followed by msd[nelem], bixs[nelem] (= ElIncXS::m_elm_data
 * .first / .second, second already multiplied by the element scale;
 * NCElIncXS.cc:59-79). */
typedef struct {
  uint64_t nelem;
  uint64_t reserved;
} ncb_elinc_t;

/* FreeGas leaf (NCFreeGas.cc:28-41) = FreeGasXSProvider{m_sigmaFree,m_ca} +
 * temperature + target mass. */
typedef struct {
  double sigma_free;   /* FreeGasXSProvider::m_sigmaFree */
  double ca;           /* FreeGasXSProvider::m_ca = A/kT */
  double temperature;  /* kelvin */
  double mass_amu;     /* target mass */
} ncb_freegas_t;

/* SABScatter.  Followed by egrid[negrid], xs[negrid] (SABXSProvider::m_egrid/m_xs,
 * identical to SABSampler::m_egrid), alpha[nalpha], beta[nbeta],
 * sab[nbeta*nalpha] (SABData, row = beta index; NCSABData.hh). */
typedef struct {
  double   scale;          /* SABScatter::m_scale (NCSABScatter.cc:87) */
  double   temperature;    /* SABData::temperature() [K] */
  double   mass_amu;       /* SABData::elementMassAMU() */
  double   bound_xs;       /* SABData::boundXS() */
  double   suggested_emax; /* SABData::suggestedEmax() */
  double   ext_sigma_free; /* extender: SABFGExtender::m_xsprovider.m_sigmaFree */
  double   ext_ca;         /* extender: m_ca */
  double   ext_temperature;/* extender: m_t */
  double   ext_mass_amu;   /* extender: m_m */
  double   k_extension;    /* SABXSProvider::m_kExtension (NCSABXSProvider.cc:50) */
  double   xs_at_emax;     /* SABSampler::m_xsAtEmax */
  double   k1, k2;         /* SABSampler::m_k1, m_k2 (NCSABSampler.cc:51-52) */
  double   egrid_margin;   /* SABSampler::m_egridMargin (1.05) */
  uint64_t negrid, nalpha, nbeta;
  uint64_t auto_egrid;     /* 0: egrid[]/xs[] and k_extension, xs_at_emax, k1, k2 come from the reference.
                              1: they are placeholders -- the library determines Emin/Emax (SABIntegrator::
                              determineEMin/determineEMax, NCSABIntegrator.cc:147-200; egrid[0], egrid[1] hold a
                              requested emin, emax -- an NCMAT "egrid" line -- or 0 = automatic), lays out the negrid-point
                              geometric grid (:203-283) and integrates the cross sections itself */
} ncb_sab_t;

/* SABScatter leaf given by its VDOS (DI_VDOS / DI_VDOSDebye, NCDynamicInfo; expansion: createScatteringKernel,
 * src/vdos/NCVDOSToScatKnl.cc:787-944).  Followed by density[ndensity] (VDOSData::vdos_density(), regular grid
 * over [emin, emax]). */
typedef struct {
  double   scale;           /* SABScatter::m_scale */
  double   temperature;     /* VDOSData::temperature() [K] */
  double   mass_amu;        /* VDOSData::elementMassAMU() */
  double   bound_xs;        /* VDOSData::boundXS() */
  double   ext_sigma_free, ext_ca, ext_temperature, ext_mass_amu;   /* free-gas extender, as in ncb_sab_t */
  double   egrid_margin;    /* SABSampler::m_egridMargin (1.05) */
  double   emin, emax;      /* VDOSData::vdos_egrid() */
  double   target_emax;     /* Emax the expansion must reach (an "egrid" request of the material), 0 = by vdoslux */
  double   req_emin, req_emax; /* request for the energy grid of the integrated cross sections, 0 = automatic */
  uint64_t vdoslux;         /* 0..5 (for a Debye-model leaf: the reduced value the reference uses, max(0,vdoslux-3)) */
  uint64_t negrid;          /* number of energy grid points (SABIntegrator, 300 by default) */
  uint64_t ndensity;
  uint64_t reserved;
} ncb_sabvdos_t;

/* SCBragg (mosaic single crystal; NCSCBragg.cc:33-90 pimpl + GaussMos + GaussOnSphere).
 * Followed by, in order:
 *   fam_xsfact[nfam], fam_inv2d[nfam]   ReflectionFamily::xsfact / inv2d (families sorted by
 *                                       d-spacing descending, NCSCBragg.cc:50-56)
 *   fam_first[nfam+1]                   (as doubles) index of each family's first demi-normal
 *   normals[3*nnormals]                 demi-normals in the lab frame, xyz interleaved
 *   lut_sofcosd[2*(lut_sofcosd_n)]      CubicSpline::m_data pairs (value, second derivative)
 *   lut_evalcosx[2*(lut_evalcosx_n)]    of GaussOnSphere::m_lt_sofcosd / m_lt_evalcosx */
typedef struct {
  double   threshold_ekin;     /* SCBragg::pimpl::m_threshold_ekin */
  double   gos_cta;            /* GaussOnSphere::m_cta (NCGaussOnSphere.hh) */
  double   gos_sta;
  double   gos_circleint_k1, gos_circleint_k2;
  double   gos_norm, gos_expfact, gos_truncangle, gos_sigma;
  double   gos_numint_accuracy;
  double   gos_prec;
  double   sofcosd_a, sofcosd_invdelta;     /* SplinedLookupTable::m_a / m_invdelta */
  double   evalcosx_a, evalcosx_invdelta;
  double   mos_fwhm, mos_truncN;            /* GaussMos (informational) */
  double   reserved[3];
  uint64_t nfam, nnormals;
  uint64_t lut_sofcosd_n, lut_evalcosx_n;   /* number of (value,d2) pairs; CubicSpline::m_nm2 = n-2 */
} ncb_scbragg_t;

/* LCBragg (layered crystal, e.g. pyrolytic graphite; mode 0 = LCHelper, NCLCBragg.cc:52-69).  Followed by
 *   planesets[7*nplanesets]   LCPlaneSet members in declaration order (NCLCUtils.hh:35-50): twodsp, inv_twodsp,
 *                             cosalpha, sinalpha, cosalphaminus, cosalphaplus, fsq; sorted by d-spacing, largest first
 *   lut_sofcosd[2*n], lut_evalcosx[2*n]   spline tables of the GaussOnSphere inside LCStdFrame::m_gm (as for SCBragg) */
typedef struct {
  double   ekin_low;           /* LCBragg::pimpl::m_ekin_low */
  double   lcaxis_lab[3];      /* LCHelper::m_lcaxislab */
  double   xsfact;             /* LCHelper::m_xsfact = 1/(V0*natoms) */
  double   gos_cta, gos_sta;
  double   gos_circleint_k1, gos_circleint_k2;
  double   gos_numint_accuracy;
  double   gos_prec;           /* GaussMos::precision() (accuracy of the phi integration, NCLCUtils.cc:469) */
  double   gos_truncangle;
  double   sofcosd_a, sofcosd_invdelta;
  double   evalcosx_a, evalcosx_invdelta;
  double   mos_fwhm;
  double   reserved[3];
  uint64_t nplanesets;
  uint64_t lut_sofcosd_n, lut_evalcosx_n;
  uint64_t reserved2;
} ncb_lcbragg_t;

static inline uint64_t ncb_align16(uint64_t x) { return (x + 15u) & ~(uint64_t)15u; }

#ifdef __cplusplus
}
#endif
#endif
