// ncb_proc.cuh -- the flattened ProcComposition: fused weighted sum over the
// components (= phases x leaf processes) and component choice for sampling.
// Restates ProcComposition::Impl::updateCacheIsotropic (ref: src/interfaces/NCProcImpl.cc:166-204),
// crossSectionIsotropic (:353-362) and sampleScatterIsotropic (:379-389).
#pragma once
#include "ncb_phys_sab.cuh"
#include "ncb_phys_scbragg.cuh"
#include "ncb_phys_lcbragg.cuh"

namespace ncb {

  // The small, hot lookup tables of the cross-section path.  The pointers are either
  // the material's global-memory arrays or their per-CTA copies staged into shared
  // memory by TMA bulk copies (ncb_kernels.cu: stageHotTabs).
  struct HotTabs {
    const double* pb_e2d[kMaxPB];
    const double* pb_fdm[kMaxPB];
    const double* sab_egrid[kMaxSab];
    const double* sab_xs[kMaxSab];
    const uint16_t* pb_lut[kMaxPB];      // energy-key tables (KeyLut) of the two searches, or null
    const uint16_t* sab_elut[kMaxSab];
    const ScBraggT* sc;   // the material's SCBragg tables, or `scv` (arrays staged in shared memory)
    ScBraggT scv;
  };
  NCB_HD void hotTabsFromMaterial( const Material& M, HotTabs& H )
  {
    for ( int i = 0; i < kMaxPB; ++i ) { H.pb_e2d[i] = M.pb[i].e2d; H.pb_fdm[i] = M.pb[i].fdm; H.pb_lut[i] = M.pb[i].lut; }
    for ( int i = 0; i < kMaxSab; ++i ) { H.sab_egrid[i] = M.sab[i].egrid; H.sab_xs[i] = M.sab[i].xs; H.sab_elut[i] = M.sab[i].elut; }
    H.sc = &M.sc;
  }

  // Unscaled isotropic xs of component i.  aux receives the PowderBragg plane index / the S(alpha,beta) energy-grid
  // position upper_bound(egrid,E).
  NCB_HD double compXSIso( const Material& M, const HotTabs& H, int i, double ekin, int& aux )
  {
    const Comp& c = M.comp[i];
    switch ( c.kind ) {
    case KIND_POWDERBRAGG: {
      const PowderBraggT& T = M.pb[c.idx];
      return pbXS( T, H.pb_e2d[c.idx], H.pb_fdm[c.idx], H.pb_lut[c.idx], ekin, aux );
    }
    case KIND_ELINC:
      return elincXS( M.elinc[c.idx], ekin, nullptr );
    case KIND_SAB: {
      const SabT& T = M.sab[c.idx];
      return sabXS( T, H.sab_egrid[c.idx], H.sab_xs[c.idx], H.sab_elut[c.idx], ekin, &aux );
    }
    case KIND_FREEGAS:
      return fgXS( M.fg[c.idx], ekin );
    case KIND_ABSOOV: {
      // AbsOOV::crossSectionIsotropic, ref: src/absoov/NCAbsOOV.cc:41-45
      const double sqrtE = sqrt( ekin );
      return sqrtE ? c.par / sqrtE : kInf;
    }
    default:
      return 0.0;
    }
  }

  // Total xs; fills cumul[0..ncomp) (componentXSectCommul) and aux[] when non-null.
  NCB_HD double matXSIso( const Material& M, const HotTabs& H, double ekin, double* cumul, int* aux )
  {
    if ( !domainContains( M.dom_lo, M.dom_hi, ekin ) )
      return 0.0;
    double tot = 0.0;
    for ( int i = 0; i < M.ncomp; ++i ) {
      const Comp& c = M.comp[i];
      int a = -1;
      const double xs = domainContains( c.dom_lo, c.dom_hi, ekin ) ? compXSIso( M, H, i, ekin, a ) : 0.0;
      tot += c.scale * xs;
      if ( cumul ) cumul[i] = tot;
      if ( aux ) aux[i] = a;
    }
    return tot;
  }

  // leaf sampleScatterIsotropic dispatch
  NCB_HD void compSampleIso( const Material& M, const HotTabs& H, int i, int aux, double ekin, Rng& rng,
                             double& ekin_out, double& mu, int& err )
  {
    const Comp& c = M.comp[i];
    switch ( c.kind ) {
    case KIND_POWDERBRAGG: {
      // PowderBragg::sampleScatterIsotropic, ref: NCPowderBragg.cc:202-216
      const PowderBraggT& T = M.pb[c.idx];
      ekin_out = ekin;
      if ( ekin < T.threshold || !isFinite(ekin) ) {
        mu = 1.0;
        return;
      }
      if ( aux < 0 )
        aux = pbLastValidPlane( T, H.pb_e2d[c.idx], H.pb_lut[c.idx], ekin );
      mu = pbSampleMu( H.pb_e2d[c.idx], H.pb_fdm[c.idx], aux, ekin, rng );
      return;
    }
    case KIND_ELINC:
      ekin_out = ekin;
      mu = elincSampleMu( M.elinc[c.idx], ekin, rng );
      return;
    case KIND_SAB:
      sabSampleScatter( M.sab[c.idx], ekin, rng, ekin_out, mu, err );
      return;
    case KIND_FREEGAS:
      fgSampleScatter( M.fg[c.idx], ekin, rng, ekin_out, mu, err );
      return;
    default:
      ekin_out = ekin; mu = 1.0;
      return;
    }
  }

  // ProcComposition::sampleScatterIsotropic.  Also returns the total xs (the
  // "evalXSAndSampleScatterIsotropic" fused entry of the reference's batch ABI,
  // ref: include/NCrystal/internal/extd_utils/NCABIUtils.hh:78-100).
  NCB_HD double matSampleIso( const Material& M, const HotTabs& H, double ekin, Rng& rng,
                              double& ekin_out, double& mu, int& err, int& ichoice_out )
  {
    ichoice_out = -1;
    if ( !domainContains( M.dom_lo, M.dom_hi, ekin ) ) {
      ekin_out = ekin; mu = 1.0;
      return 0.0;
    }
    double cumul[kMaxComp];
    int aux[kMaxComp];
    const double tot = matXSIso( M, H, ekin, cumul, aux );
    const int ichoice = ( M.ncomp == 1 ? 0 : pickIdxByWeight( rng.generate(), cumul, M.ncomp ) );
    ichoice_out = ichoice;
    compSampleIso( M, H, ichoice, aux[ichoice], ekin, rng, ekin_out, mu, err );
    return tot;
  }

  // ------------------------------------------------------------------ oriented API
  // ProcComposition::crossSection (ref: NCProcImpl.cc:340-351) with updateCacheAnisotropic
  // (:206-249): isotropic leaves answer crossSection(E,dir) with their isotropic value
  // (NCProcImpl.hh:463).  aux[i]: PowderBragg plane index, or for SCBragg the number of
  // contributing normals (entries of xs_commul) / for LCBragg the number of ROIs; sc_total: SCBragg's unscaled xs /
  // LCBragg's sum over the ROIs before the 1/(V0*natoms) factor.
  NCB_HD double matXS( const Material& M, const HotTabs& H, double ekin, const Vec3& dir,
                       double* cumul, int* aux, double* sc_total )
  {
    if ( !domainContains( M.dom_lo, M.dom_hi, ekin ) )
      return 0.0;
    double tot = 0.0;
    for ( int i = 0; i < M.ncomp; ++i ) {
      const Comp& c = M.comp[i];
      int a = -1;
      double xs = 0.0;
      if ( domainContains( c.dom_lo, c.dom_hi, ekin ) ) {
        if ( c.kind == KIND_SCBRAGG ) {
          xs = scXS( *H.sc, ekin, dir, a );
          if ( sc_total ) *sc_total = xs;
        } else if ( c.kind == KIND_LCBRAGG ) {
          double raw; int lcerr = 0;
          xs = lcXS( *H.sc, M.lc, ekin, dir, raw, a, lcerr );
          if ( sc_total ) *sc_total = raw;
        } else {
          xs = compXSIso( M, H, i, ekin, a );
        }
      }
      tot += c.scale * xs;
      if ( cumul ) cumul[i] = tot;
      if ( aux ) aux[i] = a;
    }
    return tot;
  }

  // ProcComposition::sampleScatter, ref: NCProcImpl.cc:364-377.  Returns the total xs.
  NCB_HD double matSample( const Material& M, const HotTabs& H, double ekin, const Vec3& dir, Rng& rng,
                           double& ekin_out, Vec3& outdir, int& err, int& ichoice_out )
  {
    ichoice_out = -1;
    ekin_out = ekin;
    outdir = dir;
    if ( !domainContains( M.dom_lo, M.dom_hi, ekin ) )
      return 0.0;
    double cumul[kMaxComp];
    int aux[kMaxComp];
    double sc_total = 0.0;
    const double tot = matXS( M, H, ekin, dir, cumul, aux, &sc_total );
    const int ichoice = ( M.ncomp == 1 ? 0 : pickIdxByWeight( rng.generate(), cumul, M.ncomp ) );
    ichoice_out = ichoice;
    if ( M.comp[ichoice].kind == KIND_SCBRAGG ) {
      // (when the component's domain excludes E its xs pass was skipped: aux = -1 -> no entries)
      scSampleScatter( *H.sc, ekin, dir, aux[ichoice] > 0 ? aux[ichoice] : 0, sc_total, rng, outdir );
    } else if ( M.comp[ichoice].kind == KIND_LCBRAGG ) {
      lcSampleScatter( *H.sc, M.lc, ekin, dir, sc_total, aux[ichoice] > 0 ? aux[ichoice] : 0, rng, outdir, err );
    } else {
      // ScatterIsotropicMat::sampleScatter, ref: NCProcImpl.cc:29-37
      double mu;
      compSampleIso( M, H, ichoice, aux[ichoice], ekin, rng, ekin_out, mu, err );
      if ( !( err & ( ERR_SAB_LOOP_INNER | ERR_SAB_LOOP_OUTER | ERR_SAB_DISCARD | ERR_KIN_DENOM ) ) )
        outdir = randDirectionGivenScatterMu( rng, mu, dir );
      else
        outdir = { 0.0, 0.0, 0.0 };
    }
    return tot;
  }

}
