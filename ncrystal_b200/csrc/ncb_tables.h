// ncb_tables.h -- resident (HBM) table layout of one compiled material.
//
// A Material is passed BY VALUE to every kernel (__grid_constant__, < 4 KB): it
// holds the flat (scale, leaf) list of the reference's ProcComposition
// (ref: NCProcImpl.hh:310-373) with, per leaf, scalars inline and pointers to the
// leaf's fp64 arrays in global memory.  All arrays are immutable after upload.
#pragma once
#include <cstdint>

namespace ncb {

  enum Kind : int { KIND_NONE = 0, KIND_POWDERBRAGG = 1, KIND_ELINC = 2, KIND_SAB = 3, KIND_FREEGAS = 4, KIND_SCBRAGG = 5,
                    KIND_ABSOOV = 6 /* 1/v absorption, ref: src/absoov/NCAbsOOV.cc:33-45; not a blob kind */,
                    KIND_LCBRAGG = 7 /* layered crystal, ref: src/lcbragg/NCLCBragg.cc, src/extd_utils/NCLCUtils.cc */ };

  constexpr int kMaxComp = 8;
  constexpr int kMaxPB = 4, kMaxSab = 8;   // PowderBragg / S(alpha,beta) leaves per material (multiphase mixes)
  constexpr int kMaxElIncElems = 12;

  // ref: NCPowderBragg.hh:91-93
  struct PowderBraggT {
    const double* e2d;   // m_2dE[n], ascending
    const double* fdm;   // m_fdm_commul[n]
    int n;
    double threshold;
    const uint16_t* lut; // energy-key table over e2d (KeyLut, ncb_common.cuh), or null
    int lut_key0, lut_shift, lut_nk;
  };

  // ref: NCElIncXS.hh (m_elm_data)
  struct ElIncT {
    int n;
    double msd[kMaxElIncElems];
    double bixs[kMaxElIncElems]; // bound incoherent xs * element scale
  };

  // ref: NCFreeGasUtils.hh:96-141 (FreeGasXSProvider + what FreeGasSampler's ctor needs)
  struct FreeGasT {
    double sigma_free; // m_sigmaFree
    double ca;         // m_ca = A/kT
    double kT;         // Temperature::kT()
    double mass_amu;
  };

  // Per-energy-point overlay sampler (SABSamplerAtE_Alg1, ref: NCSABSamplerModels.hh:30-80)
  struct SabEPoint {
    int32_t  npts;       // points of the beta sampler (0: SABSamplerAtE_NoScatter)
    int32_t  ibeta_off;  // m_ibetaOffset
    uint32_t off_b;      // offset of this point's (x,pdf,cdf) rows in the packed beta arrays
    uint32_t off_i;      // offset of this point's AlphaInfo entries (npts-1 of them)
    double   first_bin_endpoint; // m_firstBinKinematicEndpointValue
    const uint16_t* guide;       // this point's beta guide table (kSabGB+1 entries), or null
  };

  // AlphaSampleInfo, ref: NCSABSamplerModels.hh:47-57
  struct SabAlphaInfo {
    double f_alpha, f_sval, f_logsval;  // pt_front
    double b_alpha, b_sval, b_logsval;  // pt_back
    double prob_front, prob_notback;
    int32_t f_idx, b_idx;               // pt_front.alpha_idx, pt_back.alpha_idx
    double pad;                         // 80 bytes
  };

  // ---- gather-friendly copies of the sampler tables (layout only; same values as the arrays above them).  A table
  // attempt is a chain of dependent gathers from L2; these records put what one step of the chain needs into one
  // 32-byte sector (or two adjacent ones), and carry copies of the values the next step's address depends on.
  // Per (energy point, beta row): how SABSamplerAtE_Alg1::sampleAlpha chooses its case + what its "whole bins" case needs.
  struct SabHead {
    double prob_front, prob_notback;
    double clow, cupp;        // cumul[row][f_idx], cumul[row][b_idx]
    double inv_total;         // 1 / cumul[row][nalpha-1] (0 for an all-zero row): scales an area to the row's log-guide key
    uint32_t f_idx, b_idx;    // pt_front.alpha_idx, pt_back.alpha_idx
  };                          // 48 B
  // Per (energy point, beta row): pt_front and pt_back of AlphaSampleInfo (the two partial-bin "tails")
  struct SabTail { double alpha, sval, logsval, pad; };   // [2] per row: front, back
  // Per (beta row, alpha grid point): everything sampleAlpha gathers at one grid point
  struct SabPoint { double alpha, sab, logsab, cumul; };
  // Per (energy point, point of its beta distribution)
  struct SabBPoint { double x, pdf, cdf, pad; };

  // Log-spaced guide over a row of cumulative alpha integrals: key = exponent and top 4 mantissa bits of
  // area/total (16 buckets per octave, 32 octaves below 1).  The linear guide of r1 (256 buckets) left ~23-entry
  // searches for heavy scatterers (Al: the kinematic window usually sits in the lowest percent of the row's
  // integral); with this key the range is <= 2 entries for > 90 % of the lookups.
  constexpr int kSabGLOct = 32;
  constexpr int kSabGL = kSabGLOct*16 + 1;      // keys 0 .. kSabGL-1 (last: area/total >= 1)
  constexpr int kSabGLStride = 520;             // uint16 entries per row (kSabGL+1 used)
  inline
#if defined(__CUDACC__)
  __host__ __device__
#endif
  int sabLogKey( double x )
  {
    long long b;
#if defined(__CUDA_ARCH__)
    b = __double_as_longlong( x );
#else
    static_assert( sizeof(long long) == sizeof(double), "" );
    __builtin_memcpy( &b, &x, sizeof(b) );
#endif
    const long long k = ( b >> 48 ) - (long long)( ( 1023 - kSabGLOct ) << 4 );   // (x >= 0: sign bit clear)
    return k < 0 ? 0 : ( k > kSabGL-1 ? kSabGL-1 : (int)k );
  }

  struct SabT {
    // SABScatter / SABXSProvider / SABSampler scalars
    double scale;          // SABScatter::m_scale
    double kT;             // SABSampler::m_kT
    double k_extension;    // SABXSProvider::m_kExtension
    double k1, k2;         // SABSampler::m_k1/m_k2
    double egrid_margin;
    double egrid_log0, egrid_invdlog; // starting guess for searches in the (geometrically spaced) energy grid
    double bound_xs;       // SABData::boundXS (table builder only)
    FreeGasT ext;          // SABFGExtender
    int negrid, nalpha, nbeta;
    const double* egrid;   // [negrid]
    const double* xs;      // [negrid]
    const uint16_t* elut;  // energy-key table over egrid (KeyLut), or null
    int elut_key0, elut_shift, elut_nk;
    const double* alpha;   // [nalpha]
    const double* beta;    // [nbeta]
    const double* sab;     // [nbeta*nalpha]
    const double* logsab;  // [nbeta*nalpha]  (CommonCache::logsab)
    const double* cumul;   // [nbeta*nalpha]  (CommonCache::alphaintegrals_cumul)
    const SabEPoint* ep;   // [negrid]
    const double* bx;      // packed beta-sampler x
    const double* bpdf;    // packed normalised pdf
    const double* bcdf;    // packed cdf
    const SabAlphaInfo* ainfo; // packed
    // Guide tables (inverse-CDF bucket index -> first candidate position) that replace most steps of the
    // two binary searches of a sampling attempt; the search result is unchanged (see ncb_phys_sab.cuh).
    const uint16_t* bguide;    // [negrid][kSabGBStride] (kSabGB+1 used) over each energy point's beta CDF
    const uint16_t* aguide;    // [nbeta][kSabGA+1]   over each beta row of the cumulative alpha integrals
    const double* ascale;      // [nbeta]  kSabGA / cumul[row][nalpha-1]  (0 for an all-zero row)
    const SabHead* heads;      // [negrid*nbeta], indexed like ainfo
    const SabTail* tails;      // [negrid*nbeta][2]
    const SabPoint* pts;       // [nbeta*nalpha]
    const SabBPoint* bpts;     // [negrid*bstride], indexed like bx
    const uint16_t* lguide;    // [nbeta][kSabGLStride]
    int bstride;               // entries between the beta-sampler rows of consecutive energy points
  };
  constexpr int kSabGBStride = 1032;   // uint16 entries between the beta guides of consecutive energy points
  inline int sabBStride( int nbeta ) { return ( nbeta + 2 ) & ~1; }
  constexpr int kSabGB = 1024;
  constexpr int kSabGA = 256;

  // ref: NCSCBragg.cc:33-90 (pimpl), NCGaussOnSphere.hh (private members), NCSpline.hh:40-46
  struct SplineLutT {
    const double* data;  // (value, second derivative) pairs, CubicSpline::m_data
    int nm2;             // CubicSpline::m_nm2
    double a, invdelta;  // SplinedLookupTable::m_a / m_invdelta
  };
  struct ScBraggT {
    double threshold_ekin;
    double cta;                       // GaussOnSphere::m_cta
    double sta;                       // GaussOnSphere::m_sta (only for the scan's pre-filter window)
    double circleint_k1, circleint_k2;
    double numint_accuracy;
    int nfam, nnormals;
    const double* fam_xsfact;         // [nfam]
    const double* fam_inv2d;          // [nfam] ascending
    const int* fam_first;             // [nfam+1]
    const double* normals;            // [3*nnormals] lab frame
    const float* normals_f;           // single-precision copy for the pre-filter of k_sc_find: float4 records (x, y, z, family
                                      // index as integer bits), padded to a multiple of 128 records
    SplineLutT sofcosd, evalcosx;
  };

  // ref: NCLCUtils.hh:35-50 (LCPlaneSet), :184-246 (LCHelper); the GaussMos of its LCStdFrame lives in Material::sc
  // (nfam = 0: cta/sta, integration accuracy and the two spline tables only)
  constexpr int kLcPlaneStride = 7;   // twodsp, inv_twodsp, cosalpha, sinalpha, cosalphaminus, cosalphaplus, fsq
  struct LcBraggT {
    double ekin_low;       // LCBragg::pimpl::m_ekin_low
    double xsfact;         // LCHelper::m_xsfact
    double acc;            // LCStdFrameIntegrator::m_acc = ncclamp(precision,1e-7,1e-2)
    double ax, ay, az;     // LCHelper::m_lcaxislab
    int nplanes;
    const double* planes;  // [kLcPlaneStride*nplanes], d-spacing descending
  };

  struct Comp {
    int kind;
    int idx;       // index into the per-kind arrays of Material
    double scale;
    double dom_lo, dom_hi;
    double par;    // KIND_ABSOOV: AbsOOV::m_c
  };

  struct Material {
    int ncomp;
    int oriented;
    double dom_lo, dom_hi;
    Comp comp[kMaxComp];
    PowderBraggT pb[kMaxPB];
    ElIncT elinc[1];
    FreeGasT fg[kMaxComp];   // (gas mixtures: one free-gas leaf per element)
    SabT sab[kMaxSab];
    ScBraggT sc;   // at most one SCBragg component (oriented materials only)
    LcBraggT lc;   // at most one LCBragg component (then sc holds its mosaicity tables and has no families)
  };

}
