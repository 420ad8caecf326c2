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
                    KIND_ABSOOV = 6 /* 1/v absorption, ref: src/absoov/NCAbsOOV.cc:33-45; not a blob kind */ };

  constexpr int kMaxComp = 8;
  constexpr int kMaxPB = 4, kMaxSab = 8;   // PowderBragg / S(alpha,beta) leaves per material (multiphase mixes)
  constexpr int kMaxElIncElems = 12;

  // ref: NCPowderBragg.hh:91-93
  struct PowderBraggT {
    const double* e2d;   // m_2dE[n], ascending
    const double* fdm;   // m_fdm_commul[n]
    int n;
    double threshold;
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

  // Per (energy point, beta row): what SABSamplerAtE_Alg1::sampleAlpha needs to choose its case and to run the
  // "whole bins" case, packed so that one overlay sampler's entries can be staged in shared memory (48 B, a
  // multiple of 16 for the bulk copy).  clow/cupp/ascale are copies of cumul[row][f_idx], cumul[row][b_idx] and
  // ascale[row]: they take two dependent gathers out of every alpha sample.
  struct SabHead {
    double prob_front, prob_notback;
    double clow, cupp;
    double ascale;
    uint32_t f_idx, b_idx;
  };
  // Per (beta row, alpha grid point): everything sampleAlpha gathers at one grid point, in one 32-byte sector.
  struct SabPoint { double alpha, sab, logsab, cumul; };

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
    const SabPoint* pts;       // [nbeta*nalpha]
    int bstride;               // doubles between the beta-sampler rows of consecutive energy points (even: 16-byte rows)
  };
  constexpr int kSabGBStride = 1032;   // uint16 entries between the beta guides of consecutive energy points (16-byte rows)
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
    const float* normals_f;           // [3][nnormals] single-precision copy, SoA (pre-filter of k_sc_find only)
    SplineLutT sofcosd, evalcosx;
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
  };

}
