// ncb_phys_basic.cuh -- PowderBragg, incoherent-elastic (Debye-Waller), free-gas
// cross section, S(alpha,beta) cross section, kinematics helpers.
// Every function cites the reference routine whose arithmetic it restates
// (paths relative to /root/reference/ncrystal_core).
#pragma once
#include "ncb_common.cuh"
#include "ncb_rng.cuh"
#include "ncb_tables.h"

namespace ncb {

  // ---------------------------------------------------------------- PowderBragg
  // findLastValidPlaneIdx, ref: src/powderbragg/NCPowderBragg.cc:152-163
  template <class Ptr>
  NCB_HD int pbLastValidPlane( Ptr e2d, int n, double ekin )
  {
    return upperBound( e2d, 1, n, ekin ) - 1;
  }
  // same through the table's energy-key lut (`lut`: the table's, possibly staged; may be null)
  template <class Ptr>
  NCB_HD int pbLastValidPlane( const PowderBraggT& T, Ptr e2d, const uint16_t* lut, double ekin )
  {
    const int u = upperBoundKeyed( e2d, T.n, ekin, lut, T.lut_key0, T.lut_shift, T.lut_nk );   // over [0,n)
    return ( u > 1 ? u : 1 ) - 1;                                                               // = upper_bound over [1,n) - 1
  }

  // crossSectionIsotropic, ref: NCPowderBragg.cc:166-176 (+ cache update :42-50: inv_ekin = 1/E)
  // `idx_out` receives the last valid plane (or -1).
  template <class Ptr>
  NCB_HD double pbXS( const PowderBraggT& T, Ptr e2d, Ptr fdm, const uint16_t* lut, double ekin, int& idx_out )
  {
    idx_out = -1;
    if ( ekin < T.threshold || !isFinite(ekin) )
      return 0.0;
    const int idx = pbLastValidPlane( T, e2d, lut, ekin );
    idx_out = idx;
    const double inv_ekin = 1.0 / ekin;
    return fdm[idx] * inv_ekin;
  }

  // genScatterMu, ref: NCPowderBragg.cc:178-200 (1 draw)
  template <class Ptr>
  NCB_HD double pbSampleMu( Ptr e2d, Ptr fdm, int last_valid_idx, double ekin, Rng& rng )
  {
    const double target = rng.generate() * fdm[last_valid_idx];
    const int j = lowerBound( fdm, 0, last_valid_idx, target ); // == last_valid_idx if none
    const double sin_theta_bragg_squared = e2d[j] / ekin;
    return 1.0 - 2.0 * sin_theta_bragg_squared;
  }

  // ------------------------------------------------------------------- ElIncXS
  // eval_1mexpmtdivt, ref: src/phys_utils/NCElIncXS.cc:34-51
  NCB_HD double elinc_1mexpmtdivt( double t )
  {
    if ( t < 0.01 )
      return ( 1 + t * ( -0.5 + t * 0.16666666666666666666666666666666666666666667 * ( 1. - 0.25*t ) ) );
    if ( t > 24.0 )
      return 1.0 / t;
    t = -t;
    return m_expm1(t) / t;
  }

  // evaluate / evalXSContribsCommul, ref: NCElIncXS.cc:117-141.  `contribs` may be null.
  NCB_HD double elincXS( const ElIncT& T, double ekin, double* contribs )
  {
    constexpr double kkk = 16.0 * kPiSq * kEkin2WlSqInv;
    const double e = kkk * ekin;
    double xs = 0.0;
    for ( int i = 0; i < T.n; ++i ) {
      xs += T.bixs[i] * elinc_1mexpmtdivt( T.msd[i] * e );
      if ( contribs ) contribs[i] = xs;
    }
    return xs;
  }

  // exp_smallarg_approx, ref: include/NCrystal/internal/utils/NCMath.hh:434-440
  NCB_HD double expSmallArg( double x )
  {
    return 1.0+x*(1+x*(0.5+x*(0.16666666666666666666666666666666666667+x*(0.04166666666666666666666666666666666667
           +x*(0.00833333333333333333333333333333333333+x*(0.00138888888888888888888888888888888889
           +x*0.00019841269841269841269841269841269841))))));
  }

  // sampleMuMonoAtomic, ref: NCElIncXS.cc:81-115
  NCB_HD_NOINLINE double elincSampleMuMono( Rng& rng, double ekin, double msd )
  {
    constexpr double kkk = 8.0 * kPiSq * kEkin2WlSqInv;
    const double twoksq = kkk * ekin;
    const double a = twoksq * msd;
    if ( a < 0.01 ) {
      const double maxval = expSmallArg( a );
      while ( true ) {
        const double mu = rng.generate()*2.0 - 1.0;
        if ( rng.generate()*maxval < expSmallArg( a*mu ) )
          return mu;
      }
    }
    return dclamp( m_log1p( rng.generate() * m_expm1( 2.0*a ) ) / a - 1.0, -1.0, 1.0 );
  }

  // EPointAnalysis::sampleMu, ref: NCElIncXS.cc:178-190 (+ ElIncScatter::sampleScatterIsotropic,
  // src/elincscatter/NCElIncScatter.cc:198-204)
  NCB_HD double elincSampleMu( const ElIncT& T, double ekin, Rng& rng )
  {
    if ( T.n == 1 )
      return elincSampleMuMono( rng, ekin, T.msd[0] );
    double contribs[kMaxElIncElems];
    elincXS( T, ekin, contribs );
    const int choice = pickIdxByWeight( rng.generate(), contribs, T.n );
    return elincSampleMuMono( rng, ekin, T.msd[choice] );
  }

  // ------------------------------------------------------------ free-gas xs
  // FreeGasXSProvider::evalXSShapeASq, ref: src/phys_utils/NCFreeGasUtils.cc:64-83
  NCB_HD_NOINLINE double fgXSShapeASq( double a_squared )
  {
    if ( a_squared > 36.0 )
      return 1.0 + 0.5 / a_squared;
    const double a = sqrt( a_squared );
    if ( a < 0.1 ) {
      if ( a == 0.0 )
        return kInf;
      constexpr double c1 = 2.0/3.0;
      constexpr double c2 = 1.0/15.0;
      constexpr double c3 = 1.0/105.0;
      constexpr double c4 = 1.0/756.0;
      constexpr double c5 = 1.0/5940.0;
      const double a2 = a_squared;
      return kInvSqrtPi * ( 2.0 / a + a *( c1- a2*(c2-a2*(c3-a2*(c4-a2*c5)))));
    }
    const double inva = 1.0 / a;
    return ( 1.0 + 0.5*inva*inva ) * m_erf(a) + kInvSqrtPi * m_exp(-a_squared)*inva;
  }

  // FreeGasXSProvider::crossSection, ref: include/NCrystal/internal/phys_utils/NCFreeGasUtils.hh:122-125
  NCB_HD double fgXS( const FreeGasT& T, double ekin )
  {
    return T.sigma_free * fgXSShapeASq( T.ca * ekin );
  }

  // --------------------------------------------------------------- SAB xs
  // SABXSProvider::crossSection, ref: src/sab/NCSABXSProvider.cc:54-95, times
  // SABScatter::m_scale (src/sabscatter/NCSABScatter.cc:87-90).
  // `iu_out` (optional) receives upper_bound(egrid, E), which the sampler's choice of overlay starts from.
  template <class Ptr>
  NCB_HD double sabXS( const SabT& T, Ptr egrid, Ptr xsv, const uint16_t* elut, double ekin, int* iu_out = nullptr )
  {
    const int n = T.negrid;
    const int iu = upperBoundKeyed( egrid, n, ekin, elut, T.elut_key0, T.elut_shift, T.elut_nk );
    if ( iu_out ) *iu_out = iu;
    double xs;
    if ( iu == n ) {
      xs = T.k_extension / ekin + fgXS( T.ext, ekin );
    } else if ( iu == 0 ) {
      xs = ekin > 0.0 ? sqrt( egrid[0] / ekin ) * xsv[0] : kInf;
    } else {
      const double e0 = egrid[iu-1], e1 = egrid[iu];
      const double x0 = xsv[iu-1], x1 = xsv[iu];
      const double dXS = x1 - x0;
      const double dEkin = e1 - e0;
      xs = x0 + dXS * ( ekin - e0 ) / dEkin;
    }
    return xs * T.scale;
  }

  // --------------------------------------------------------------- kinematics
  // getAlphaLimits, ref: include/NCrystal/internal/phys_utils/NCKinUtils.hh:85-124
  struct AlphaLimits { double first, second; };
  NCB_HD AlphaLimits getAlphaLimits( double ekin_div_kT, double beta )
  {
    const double kk = ekin_div_kT + beta;
    if ( !( kk >= 0.0 ) )
      return { 1.0, -1.0 };
    const double a = kk + ekin_div_kT;
    const double b = 2.0 * sqrt( ekin_div_kT * kk );
    double aminus;
    if ( fabs(beta) < 0.01*ekin_div_kT ) {
      // alphaMinusTaylor, NCKinUtils.hh:91-105
      const double x = beta / ekin_div_kT;
      constexpr double c9 = -715./32768.;
      constexpr double c8 = 429./16384.;
      constexpr double c7 = -33./1024.;
      constexpr double c6 = 21./512.;
      constexpr double c5 = -7./128.;
      constexpr double c4 = 5./64.;
      constexpr double c3 = - 1./8.;
      constexpr double c2 = 1./4.;
      aminus = beta*x*(c2+x*(c3+x*(c4+x*(c5+x*(c6+x*(c7+x*(c8+x*c9)))))));
    } else {
      aminus = dmax( 0.0, a - b );
    }
    return { aminus, a + b };
  }

  // muIsotropicAtBeta, ref: NCKinUtils.hh:64-70
  NCB_HD bool muIsotropicAtBeta( double beta, double ekin_div_kT )
  {
    constexpr double lim = -1.0 + 1e-14;
    return beta <= ekin_div_kT * lim;
  }

  // convertAlphaBetaToDeltaEMu, ref: src/phys_utils/NCKinUtils.cc:25-56.
  // Sets `err` when the reference would throw (denominator == 0).
  NCB_HD void alphaBetaToDeltaEMu( double alpha, double beta, double ekin, double kT,
                                   double& deltaE, double& mu, int& err )
  {
    deltaE = beta * kT;
    const double ekinfinal = ekin + deltaE;
    const double denom = 2.0 * sqrt( ekin * ekinfinal );
    if ( !denom ) {
      err = 1;
      mu = -999.0;
      return;
    }
    StableSum sum;
    sum.add( ekin );
    sum.add( ekinfinal );
    sum.add( -alpha*kT );
    mu = dclamp( sum.sum() / denom, -1.0, 1.0 );
  }

}
