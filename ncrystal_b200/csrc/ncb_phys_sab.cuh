// ncb_phys_sab.cuh -- S(alpha,beta) rejection sampling (Cai et al., JCP 2019, Alg. 1).
// Restates, per neutron,
//   SABSampler::sampleAlphaBeta / sampleHighE / sampleDeltaEMu   ref: src/sab/NCSABSampler.cc:59-236
//   SABSamplerAtE_Alg1::sampleAlphaBeta / sampleAlpha            ref: src/sab/NCSABSamplerModels.cc:48-233
//   PointwiseDist::percentileWithIndex                           ref: src/utils/NCPointwiseDist.cc:76-105
//   SABUtils::sampleLogLinDist_fast                              ref: include/NCrystal/internal/sab/NCSABUtils.hh:282-303
//   SABScatter::sampleScatterIsotropic                           ref: src/sabscatter/NCSABScatter.cc:93-100
#pragma once
#include "ncb_phys_freegas.cuh"

namespace ncb {

  enum SampleErr : int {
    ERR_NONE = 0,
    ERR_KIN_DENOM = 1,       // convertAlphaBetaToDeltaEMu: beta == -E/kT
    ERR_SAB_LOOP_OUTER = 2,  // SABSampler::sampleAlphaBeta: 100 tries (NCSABSampler.cc:226)
    ERR_SAB_LOOP_INNER = 4,  // SABSamplerAtE_Alg1: 100 tries (NCSABSamplerModels.cc:150)
    ERR_SAB_DISCARD = 8,     // sampleHighE: P_discardinside > 0.95 (NCSABSampler.cc:112)
    ERR_SAB_ISOFALLBACK = 16,// (warning only) isotropic fallback after 30 tries (NCSABSamplerModels.cc:99)
    ERR_SAB_ROUTING = 32,    // internal: E > Emax neutron reached the table-only kernel
    ERR_MMC_NOTERM = 128,    // transport: a history did not end within the step limit
    ERR_LC_ROMBERG = 64      // LCBragg: phi integration did not converge (Romberg::convergenceError, NCRomberg.cc:48-61)
  };

  // sampleLogLinDist_fast, ref: NCSABUtils.hh:282-303
  NCB_HD_NOINLINE double sampleLogLinDistFast( double a, double fa, double b, double fb, double rand, double logfa, double logfb )
  {
    double df = fb - fa;
    if ( fa*fb*df != 0.0 ) {
      const double a_sub_b = a - b;
      const double logfa_fb = logfb - logfa;
      if ( a_sub_b * logfa_fb != 0.0 )
        return a_sub_b * m_log( fa*m_exp( a*logfa_fb/a_sub_b ) / ( fa + rand*df ) ) / logfa_fb;
      df = 0.0;
    }
    if ( !df )
      return a + rand*(b-a);
    const double x = (b-a)*sqrt(rand);
    return ( fa ? b-x : a+x );
  }

  // PointwiseDist::percentileWithIndex, ref: NCPointwiseDist.cc:76-105
  // `guide` (optional): g[b] = lower_bound(cdf, b/kSabGB), b = 0..kSabGB.  For p in [b/G,(b+1)/G) the
  // lower_bound lies in [g[b], g[b+1]] (p*G and b/G are exact: G is a power of two), so only that
  // slice is searched -- identical result, ~2 instead of ~10 dependent loads.
  NCB_HD double pwdPercentileWithIndex( const double* x, const double* y, const double* cdf, int n, double p, int& idx,
                                        const uint16_t* guide = nullptr )
  {
    if ( p == 1. ) {
      idx = n-2;
      return ldTable( x + n-1 );
    }
    int i;
    if ( guide ) {
      const int b = (int)( p * (double)kSabGB );
      i = lowerBoundTable( cdf, (int)ldTable( guide + b ), (int)ldTable( guide + b+1 ), p );
    } else {
      i = lowerBoundTable( cdf, 0, n, p );
    }
    i = i < n-1 ? i : n-1;
    i = i > 1 ? i : 1;
    const double x0 = ldTable( x + i-1 ), x1 = ldTable( x + i );
    const double dx = x1 - x0;
    const double c = ( p - ldTable( cdf + i-1 ) );
    const double a = ldTable( y + i-1 );
    const double d = ldTable( y + i ) - a;
    double zdx;
    if ( !a ) {
      zdx = d > 0.0 ? sqrt( ( 2.0 * c * dx ) / d ) : 0.5*dx;
    } else {
      const double e = d * c / ( dx * a * a );
      if ( fabs(e) > 1e-7 )
        zdx = ( sqrt( 1.0 + 2.0 * e ) - 1.0 ) * dx * a / d;
      else
        zdx = ( 1 + 0.5 * e * ( e - 1.0 ) ) * c / a;
    }
    idx = i-1;
    return dclamp( x0 + zdx, x0, x1 );
  }

  // SABSamplerAtE_Alg1::sampleAlpha, ref: NCSABSamplerModels.cc:157-233.
  // The reference's three cases (front tail / whole bins / back tail) each end in a call of
  // sampleLogLinDist_fast; here the cases only select its arguments and all lanes then meet at
  // ONE call site, so a warp whose lanes fall into different cases evaluates the expensive
  // log/exp sequence once instead of up to three times.  Same arithmetic, same results.
  NCB_HD_NOINLINE double sabSampleAlpha( const SabT& T, const SabEPoint& ep, int ibeta, double rand_percentile )
  {
    const SabAlphaInfo* info = &T.ainfo[ ep.off_i + ( ibeta - ep.ibeta_off ) ];
    const int nalpha = T.nalpha;
    const double* cumul  = T.cumul  + (size_t)ibeta*nalpha;
    const double* sab    = T.sab    + (size_t)ibeta*nalpha;
    const double* logsab = T.logsab + (size_t)ibeta*nalpha;
    const double* agrid  = T.alpha;
    double a, fa, b, fb, r, la, lb;
    const double prob_front = ldTable( &info->prob_front ), prob_notback = ldTable( &info->prob_notback );

    if ( rand_percentile <= prob_front ) {
      const double f_alpha = ldTable( &info->f_alpha );
      if ( prob_front == 2.0 ) {
        const double da = ldTable( &info->b_alpha ) - f_alpha;
        return f_alpha + rand_percentile*da;
      } else if ( prob_front == 1.0 ) {
        a = f_alpha; fa = ldTable( &info->f_sval ); b = ldTable( &info->b_alpha ); fb = ldTable( &info->b_sval );
        r = rand_percentile; la = ldTable( &info->f_logsval ); lb = ldTable( &info->b_logsval );
      } else {
        const int fi = ldTable( &info->f_idx );
        a = f_alpha; fa = ldTable( &info->f_sval ); b = ldTable( agrid + fi ); fb = ldTable( sab + fi );
        r = dclamp( rand_percentile / prob_front, kDblMin, 1.0 );
        la = ldTable( &info->f_logsval ); lb = ldTable( logsab + fi );
      }
    } else if ( rand_percentile <= prob_notback ) {
      const double percentile2 = dclamp( ( rand_percentile - prob_front ) / ( prob_notback - prob_front ), 0.0, 1.0 );
      const int ilow = ldTable( &info->f_idx ), iupp = ldTable( &info->b_idx );
      const double clow = ldTable( cumul + ilow ), cupp = ldTable( cumul + iupp );
      const double selectedArea = clow + percentile2 * ( cupp - clow );
      int isel_upp;
      {
        // upper_bound( cumul[ilow..iupp], selectedArea ) = clamp( upper_bound over the whole row ) to that
        // range; the whole-row position is bracketed by the row's guide table and verified (the bucket
        // index involves a rounded product), with a full search as fall-back.
        const double sc = ldTable( T.ascale + ibeta );
        int bk = (int)( selectedArea * sc );
        bk = bk < 0 ? 0 : ( bk > kSabGA-1 ? kSabGA-1 : bk );
        const uint16_t* g = T.aguide + (size_t)ibeta*( kSabGA+1 ) + bk;
        int r0 = upperBoundTable( cumul, (int)ldTable( g ), (int)ldTable( g + 1 ), selectedArea );
        const bool ok = ( r0 == 0 || !( selectedArea < ldTable( cumul + r0-1 ) ) ) && ( r0 == nalpha || selectedArea < ldTable( cumul + r0 ) );
        if ( !ok )
          r0 = upperBoundTable( cumul, 0, nalpha, selectedArea );
        isel_upp = r0 < ilow ? ilow : ( r0 > iupp+1 ? iupp+1 : r0 );
      }
      if ( isel_upp > iupp )
        return ldTable( agrid + iupp );
      if ( isel_upp <= ilow )
        return ldTable( agrid + ilow );
      const int a0 = isel_upp - 1;
      const int a1 = isel_upp;
      const double c0 = ldTable( cumul + a0 ), c1 = ldTable( cumul + a1 );
      const double binArea = c1 - c0;
      r = dclamp( ( selectedArea - c0 ) / binArea, kDblMin, 1.0 );
      a = ldTable( agrid + a0 ); fa = ldTable( sab + a0 ); b = ldTable( agrid + a1 ); fb = ldTable( sab + a1 );
      la = ldTable( logsab + a0 ); lb = ldTable( logsab + a1 );
    } else {
      const int bi = ldTable( &info->b_idx );
      r = dclamp( ( rand_percentile - prob_notback ) / ( 1.0 - prob_notback ), kDblMin, 1.0 );
      a = ldTable( agrid + bi ); fa = ldTable( sab + bi ); b = ldTable( &info->b_alpha ); fb = ldTable( &info->b_sval );
      la = ldTable( logsab + bi ); lb = ldTable( &info->b_logsval );
    }
    return sampleLogLinDistFast( a, fa, b, fb, r, la, lb );
  }

  // One pass of the rejection loop body of SABSamplerAtE_Alg1::sampleAlphaBeta
  // (ref: NCSABSamplerModels.cc:62-148; requires ep.npts>0).  Returns true when a point was
  // accepted (the reference `return`s), false where it `continue`s.
  NCB_HD bool sabAttemptAtE( const SabT& T, const SabEPoint& ep, double ekin_div_kT, Rng& rng,
                             double& alpha_out, double& beta_out, int& err )
  {
    const double* bx = T.bx + ep.off_b;
    const double* betaGrid = T.beta;
    const double firstBin = ep.first_bin_endpoint;
    int ibetaSampled;
    double beta = pwdPercentileWithIndex( bx, T.bpdf + ep.off_b, T.bcdf + ep.off_b, ep.npts, rng.generate(), ibetaSampled, ep.guide );

    if ( ibetaSampled == 0 && firstBin <= 0.0 ) {
      const double b0 = firstBin;
      const double b1 = bx[1];
      if ( b1 < -ekin_div_kT )
        return false;
      const double delta_beta = b1 - b0;
      double alphaval = 0.0;
      constexpr int nsampletries = 30;
      for ( int iii = 0; iii < nsampletries; ++iii ) {
        beta = dmax( firstBin, b0 + delta_beta*rng.generate() );
        if ( beta < -ekin_div_kT )
          break;
        alphaval = sabSampleAlpha( T, ep, ep.ibeta_off, rng.generate() );
        AlphaLimits alims = getAlphaLimits( -firstBin, beta );
        if ( inInterval( alims.first, alims.second, alphaval ) )
          break;
        if ( iii == nsampletries-1 ) {
          err |= ERR_SAB_ISOFALLBACK;
          alphaval = 0.5*( alims.first + alims.second );
          break;
        }
      }
      if ( beta < -ekin_div_kT )
        return false;
      AlphaLimits alimits = getAlphaLimits( ekin_div_kT, beta );
      if ( inInterval( alimits.first, alimits.second, alphaval ) ) {
        alpha_out = alphaval; beta_out = beta;
        return true;
      }
      return false;
    }

    if ( beta <= dmax( -ekin_div_kT, betaGrid[0] ) )
      return false;

    const double rand_percentile = rng.generate();
    const int ibeta = ep.ibeta_off + ibetaSampled;
    const double bl = betaGrid[ibeta-1];
    const double alphal = sabSampleAlpha( T, ep, ibeta-1, rand_percentile );
    const double bh = betaGrid[ibeta];
    const double alphah = sabSampleAlpha( T, ep, ibeta, rand_percentile );
    const double alpha = alphal + (alphah-alphal) * (beta-bl)/(bh-bl);
    AlphaLimits alimits = getAlphaLimits( ekin_div_kT, beta );
    if ( inInterval( alimits.first, alimits.second, alpha ) ) {
      alpha_out = alpha; beta_out = beta;
      return true;
    }
    return false;
  }

  // ---------------------------------------------------------------------------------------------------------
  // Short-chain variants: the same arithmetic as pwdPercentileWithIndex / sabSampleAlpha / sabAttemptAtE above on
  // the gather-friendly copies of the tables (SabBPoint / SabHead / SabTail / SabPoint, ncb_tables.h).  A table
  // attempt is latency-bound on its chain of DEPENDENT gathers; here the chain is
  //     beta guide -> beta points | heads of both rows -> log guide -> alpha points        (~6 round trips)
  // instead of r1's ~20 (energy-point record, guide, CDF search, x/pdf/cdf rows; per row: info, cumul[ilow]/[iupp]
  // and scale, linear guide, a 5.8-step search for heavy scatterers, alpha/sab/logsab rows).
  // Measured and dropped (r2, Al 1e7: this version 0.873 ms): four independent probes instead of the bisection
  // behind the log guide (0.900), two for the beta CDF (0.882), both rows' chains run side by side (0.983: the
  // inlined double bookkeeping costs instruction-cache misses and issue slots) -- extra loads and code cost more
  // than the shorter chain saves once the chain is this short.
  NCB_HD SabPoint ldPoint( const SabPoint* p )
  {
#if defined(__CUDA_ARCH__)
    const double2 a = __ldg( reinterpret_cast<const double2*>( p ) );
    const double2 b = __ldg( reinterpret_cast<const double2*>( p ) + 1 );
    return SabPoint{ a.x, a.y, b.x, b.y };
#else
    return *p;
#endif
  }
  NCB_HD SabHead ldHead( const SabHead* p )
  {
#if defined(__CUDA_ARCH__)
    const double2 a = __ldg( reinterpret_cast<const double2*>( p ) );
    const double2 b = __ldg( reinterpret_cast<const double2*>( p ) + 1 );
    const uint4 c = __ldg( reinterpret_cast<const uint4*>( p ) + 2 );
    SabHead h;
    h.prob_front = a.x; h.prob_notback = a.y; h.clow = b.x; h.cupp = b.y;
    h.inv_total = __hiloint2double( (int)c.y, (int)c.x ); h.f_idx = c.z; h.b_idx = c.w;
    return h;
#else
    return *p;
#endif
  }
  NCB_HD void prefetchL1( const void* p )
  {
#if defined(__CUDA_ARCH__)
    asm volatile( "prefetch.global.L1 [%0];" :: "l"(p) );
#else
    (void)p;
#endif
  }
  NCB_HD int upperBoundPts( const SabPoint* a, int lo, int hi, double v )
  {
    while ( lo < hi ) {
      int mid = lo + ( ( hi - lo ) >> 1 );
      if ( !( v < ldTable( &a[mid].cumul ) ) ) lo = mid + 1; else hi = mid;
    }
    return lo;
  }

#if !defined(__CUDA_ARCH__) && defined(NCB_HOST_TRACE)
  inline void (*g_alpha_trace)( int, double, int, int, int, int, int, double ) = nullptr;   // tests/hostsim only
#endif

  // PointwiseDist::percentileWithIndex (ref: NCPointwiseDist.cc:76-105) over SabBPoint records
  NCB_HD double pwdPercentileFast( const SabBPoint* B, const uint16_t* guide, int n, double p, int& idx )
  {
    if ( p == 1. ) {
      idx = n-2;
      return ldTable( &B[n-1].x );
    }
    const int b = (int)( p * (double)kSabGB );
    int lo = (int)ldTable( guide + b ), hi = (int)ldTable( guide + b+1 );
    while ( lo < hi ) {
      const int mid = lo + ( ( hi - lo ) >> 1 );
      if ( ldTable( &B[mid].cdf ) < p ) lo = mid + 1; else hi = mid;
    }
    int i = lo < n-1 ? lo : n-1;
    i = i > 1 ? i : 1;
#if defined(__CUDA_ARCH__)
    const double2 q0 = __ldg( reinterpret_cast<const double2*>( B + i-1 ) );
    const double cdf0 = __ldg( &B[i-1].cdf );
    const double2 q1 = __ldg( reinterpret_cast<const double2*>( B + i ) );
    const double x0 = q0.x, a = q0.y, x1 = q1.x, y1 = q1.y;
#else
    const double x0 = B[i-1].x, a = B[i-1].pdf, cdf0 = B[i-1].cdf, x1 = B[i].x, y1 = B[i].pdf;
#endif
    const double dx = x1 - x0;
    const double c = ( p - cdf0 );
    const double d = y1 - a;
    double zdx;
    if ( !a ) {
      zdx = d > 0.0 ? sqrt( ( 2.0 * c * dx ) / d ) : 0.5*dx;
    } else {
      const double e = d * c / ( dx * a * a );
      if ( fabs(e) > 1e-7 )
        zdx = ( sqrt( 1.0 + 2.0 * e ) - 1.0 ) * dx * a / d;
      else
        zdx = ( 1 + 0.5 * e * ( e - 1.0 ) ) * c / a;
    }
    idx = i-1;
    return dclamp( x0 + zdx, x0, x1 );
  }

  // SABSamplerAtE_Alg1::sampleAlpha (ref: NCSABSamplerModels.cc:157-233).  `hr` = ie*nbeta + ibeta.
  // The reference's three cases each rescale the percentile with one division; the operands are selected first and
  // ONE division serves all lanes of a warp (same operands per case -> same quotient).
  NCB_HD_NOINLINE double sabSampleAlphaFast( const SabT& T, size_t hr, int ibeta, double rand_percentile )
  {
    const SabHead h = ldHead( T.heads + hr );
    const SabPoint* P = T.pts + (size_t)ibeta*T.nalpha;
    const double prob_front = h.prob_front, prob_notback = h.prob_notback;
    const bool front = ( rand_percentile <= prob_front );
    const bool middle = !front && ( rand_percentile <= prob_notback );
    if ( front && prob_front == 2.0 ) {
      const double f_alpha = ldTable( &T.tails[2*hr].alpha );
      const double da = ldTable( &T.tails[2*hr+1].alpha ) - f_alpha;
      return f_alpha + rand_percentile*da;
    }
    // (prob_front == 1.0, the single-bin case, uses the percentile as it is: x/1.0 == x)
    const double num = front ? rand_percentile : ( middle ? rand_percentile - prob_front : rand_percentile - prob_notback );
    const double den = front ? ( prob_front == 1.0 ? 1.0 : prob_front ) : ( middle ? prob_notback - prob_front : 1.0 - prob_notback );
    const double q = num / den;
    double a, fa, b, fb, r, la, lb;
    if ( middle ) {
      const double percentile2 = dclamp( q, 0.0, 1.0 );
      const int ilow = (int)h.f_idx, iupp = (int)h.b_idx;
      const double selectedArea = h.clow + percentile2 * ( h.cupp - h.clow );
      const uint16_t* g = T.lguide + (size_t)ibeta*kSabGLStride + sabLogKey( selectedArea * h.inv_total );
      // upper_bound( cumul[ilow..iupp], selectedArea ) = clamp( upper_bound over the whole row ) to that range; the
      // whole-row position lies in [g[0], g[1]] (sabLogGuideEntry)
      const int r0 = upperBoundPts( P, (int)ldTable( g ), (int)ldTable( g + 1 ), selectedArea );
#if !defined(__CUDA_ARCH__) && defined(NCB_HOST_TRACE)
      if ( g_alpha_trace ) g_alpha_trace( ibeta, selectedArea, ilow, iupp, (int)g[0], (int)g[1], r0, percentile2 );
#endif
      const int isel_upp = r0 < ilow ? ilow : ( r0 > iupp+1 ? iupp+1 : r0 );
      if ( isel_upp > iupp )
        return ldTable( &P[iupp].alpha );
      if ( isel_upp <= ilow )
        return ldTable( &P[ilow].alpha );
      const SabPoint p0 = ldPoint( P + isel_upp - 1 ), p1 = ldPoint( P + isel_upp );
      const double binArea = p1.cumul - p0.cumul;
      r = dclamp( ( selectedArea - p0.cumul ) / binArea, kDblMin, 1.0 );
      a = p0.alpha; fa = p0.sab; b = p1.alpha; fb = p1.sab; la = p0.logsab; lb = p1.logsab;
    } else if ( front ) {
      const SabTail* t = T.tails + 2*hr;
      a = ldTable( &t[0].alpha ); fa = ldTable( &t[0].sval ); la = ldTable( &t[0].logsval );
      if ( prob_front == 1.0 ) {
        b = ldTable( &t[1].alpha ); fb = ldTable( &t[1].sval ); lb = ldTable( &t[1].logsval );
        r = rand_percentile;
      } else {
        const SabPoint p = ldPoint( P + h.f_idx );
        b = p.alpha; fb = p.sab; lb = p.logsab;
        r = dclamp( q, kDblMin, 1.0 );
      }
    } else {
      const SabTail* t = T.tails + 2*hr + 1;
      const SabPoint p = ldPoint( P + h.b_idx );
      a = p.alpha; fa = p.sab; la = p.logsab;
      b = ldTable( &t->alpha ); fb = ldTable( &t->sval ); lb = ldTable( &t->logsval );
      r = dclamp( q, kDblMin, 1.0 );
    }
    return sampleLogLinDistFast( a, fa, b, fb, r, la, lb );
  }

  // One pass of the rejection loop of SABSamplerAtE_Alg1::sampleAlphaBeta (ref: NCSABSamplerModels.cc:62-148) for
  // energy point `ie` (requires ep.npts>0).  true: a point was accepted, false where the reference `continue`s.
  NCB_HD bool sabAttemptFast( const SabT& T, int ie, const SabEPoint& ep, double ekin_div_kT, Rng& rng,
                              double& alpha_out, double& beta_out, int& err )
  {
    const SabBPoint* B = T.bpts + ep.off_b;
    const double firstBin = ep.first_bin_endpoint;
    const size_t hr0 = (size_t)ie*T.nbeta;
    int ibetaSampled;
    double beta = pwdPercentileFast( B, T.bguide + (size_t)ie*kSabGBStride, ep.npts, rng.generate(), ibetaSampled );

    if ( ibetaSampled == 0 && firstBin <= 0.0 ) {
      const double b0 = firstBin;
      const double b1 = ldTable( &B[1].x );
      if ( b1 < -ekin_div_kT )
        return false;
      const double delta_beta = b1 - b0;
      double alphaval = 0.0;
      constexpr int nsampletries = 30;
      for ( int iii = 0; iii < nsampletries; ++iii ) {
        beta = dmax( firstBin, b0 + delta_beta*rng.generate() );
        if ( beta < -ekin_div_kT )
          break;
        alphaval = sabSampleAlphaFast( T, hr0 + ep.ibeta_off, ep.ibeta_off, rng.generate() );
        AlphaLimits alims = getAlphaLimits( -firstBin, beta );
        if ( inInterval( alims.first, alims.second, alphaval ) )
          break;
        if ( iii == nsampletries-1 ) {
          err |= ERR_SAB_ISOFALLBACK;
          alphaval = 0.5*( alims.first + alims.second );
          break;
        }
      }
      if ( beta < -ekin_div_kT )
        return false;
      AlphaLimits alimits = getAlphaLimits( ekin_div_kT, beta );
      if ( inInterval( alimits.first, alimits.second, alphaval ) ) {
        alpha_out = alphaval; beta_out = beta;
        return true;
      }
      return false;
    }

    if ( beta <= dmax( -ekin_div_kT, ldTable( T.beta ) ) )
      return false;

    const int ibeta = ep.ibeta_off + ibetaSampled;
    const double rand_percentile = rng.generate();
    const double bl = ldTable( T.beta + ibeta-1 );
    const double bh = ldTable( T.beta + ibeta );
    prefetchL1( T.heads + hr0 + ibeta );               // the second row's head, while the first row is processed
    const double alphal = sabSampleAlphaFast( T, hr0 + ibeta-1, ibeta-1, rand_percentile );
    const double alphah = sabSampleAlphaFast( T, hr0 + ibeta, ibeta, rand_percentile );
    const double alpha = alphal + (alphah-alphal) * (beta-bl)/(bh-bl);
    AlphaLimits alimits = getAlphaLimits( ekin_div_kT, beta );
    if ( inInterval( alimits.first, alimits.second, alpha ) ) {
      alpha_out = alpha; beta_out = beta;
      return true;
    }
    return false;
  }

  // SABSamplerAtE_Alg1::sampleAlphaBeta, ref: NCSABSamplerModels.cc:48-155
  // (npts==0 is SABSamplerAtE_NoScatter: returns (0,0), NCSABSamplerModels.hh:86)
  NCB_HD void sabSampleAtE( const SabT& T, const SabEPoint& ep, double ekin_div_kT, Rng& rng,
                            double& alpha_out, double& beta_out, int& err )
  {
    if ( ep.npts == 0 ) {
      alpha_out = 0.0; beta_out = 0.0;
      return;
    }
    for ( int iloop = 0; iloop < 100; ++iloop )
      if ( sabAttemptAtE( T, ep, ekin_div_kT, rng, alpha_out, beta_out, err ) )
        return;
    err |= ERR_SAB_LOOP_INNER;
    alpha_out = -1.0; beta_out = 0.0;
  }

  // Choice of the overlay sampler for an in-grid or below-grid energy
  // (SABSampler::sampleAlphaBeta, ref: NCSABSampler.cc:166-192; E < Emax only).
  // Same, starting from iu = upper_bound(egrid,E) when the caller already has it (the cross-section pass does).
  template <class Ptr>
  NCB_HD int sabPickSamplerFrom( const SabT& T, Ptr egrid, double ekin, int iu, bool& ultra_small_ekin_mode )
  {
    const int n = T.negrid;
    ultra_small_ekin_mode = false;
    if ( iu == 0 ) {
      ultra_small_ekin_mode = ( ekin < egrid[0] );
      return 0;
    }
    if ( T.egrid_margin > 1.0 ) {
      while ( iu+1 != n && ekin*T.egrid_margin > egrid[iu] )
        ++iu;
    }
    return iu;
  }
  NCB_HD int sabPickSampler( const SabT& T, const double* egrid, double ekin, bool& ultra_small_ekin_mode )
  {
    const int n = T.negrid;
    int iu = T.egrid_invdlog > 0.0 ? upperBoundLogGuess( egrid, n, ekin, T.egrid_log0, T.egrid_invdlog )
                                   : upperBound( egrid, 0, n, ekin );
    ultra_small_ekin_mode = false;
    if ( iu == 0 ) {
      ultra_small_ekin_mode = ( ekin < egrid[0] );
      return 0;
    }
    if ( T.egrid_margin > 1.0 ) {
      while ( iu+1 != n && ekin*T.egrid_margin > egrid[iu] )
        ++iu;
    }
    return iu;
  }

  // SABSampler::sampleHighE, ref: NCSABSampler.cc:59-156, in two pieces (the kernels schedule the free-gas sampling
  // between them attempt by attempt); sabSampleHighE below is their composition.
  enum { kHighEGoOn = 0, kHighEDiscard = 1, kHighEToEmax = 2, kHighEAccept = 3, kHighERedo = 4 };

  // Before the free-gas loop (:59-120): kHighEDiscard (error raised), kHighEToEmax (sample the table at E=Emax;
  // one uniform consumed) or kHighEGoOn (0 or 1 uniforms consumed).
  NCB_HD int sabHighEBegin( const SabT& T, double ekin, Rng& rng, double& P_discardinside, int& err )
  {
    const double extenderXSMultE = ekin * fgXS( T.ext, ekin );
    const double P_inside = T.k1 / ( (T.k1-T.k2) + extenderXSMultE );
    const double P_extender_inside = T.k2 / extenderXSMultE;
    P_discardinside = ( P_extender_inside >= P_inside ? (1.0-P_inside/P_extender_inside) : 0.0 );
    if ( P_discardinside > 0.95 ) {
      err |= ERR_SAB_DISCARD;
      return kHighEDiscard;
    }
    if ( P_extender_inside < P_inside ) {
      const double aa = 1.0 - P_extender_inside;
      const double P_extrainside = aa > 1e-10 ? (P_inside-P_extender_inside)/aa : 1.0;
      if ( rng.generate() < P_extrainside )
        return kHighEToEmax;
    }
    return kHighEGoOn;
  }

  // After one free-gas (alpha,beta) (:122-155): accept it, draw again, or sample the table at E=Emax.
  NCB_HD int sabHighECheck( const SabT& T, double alpha, double beta, double P_discardinside, Rng& rng )
  {
    const double emax_div_kt = T.egrid[T.negrid-1] / T.kT;
    if ( beta <= -emax_div_kt )
      return kHighEAccept;
    AlphaLimits alims = getAlphaLimits( emax_div_kt, beta );
    if ( !inInterval( alims.first, alims.second, alpha ) )
      return kHighEAccept;
    if ( P_discardinside && rng.generate() < P_discardinside )
      return kHighERedo;
    return kHighEToEmax;
  }

  // Returns true if (alpha,beta) was sampled with the free-gas extender; false => sample the table at E=Emax.
  NCB_HD bool sabSampleHighE( const SabT& T, double ekin, Rng& rng, double& alpha, double& beta, int& err )
  {
    double P_discardinside;
    const int b = sabHighEBegin( T, ekin, rng, P_discardinside, err );
    if ( b == kHighEDiscard ) {
      alpha = -1.0; beta = 0.0;
      return true;
    }
    if ( b == kHighEToEmax )
      return false;
    FreeGasSampler fgs( ekin, T.ext.kT, T.ext.mass_amu );
    while ( true ) {
      fgs.sampleAlphaBeta( rng, alpha, beta );
      const int c = sabHighECheck( T, alpha, beta, P_discardinside, rng );
      if ( c == kHighEAccept )
        return true;
      if ( c == kHighEToEmax )
        return false;
    }
  }

  // SABSampler::sampleAlphaBeta, ref: NCSABSampler.cc:158-227.
  // kHighE=false instantiates the tabulated-kernel path only (E <= Emax): the kernels route
  // E > Emax neutrons to a separate launch, which keeps the free-gas extender code (and its
  // register footprint) out of the table-sampling kernel.
  template <bool kHighE>
  NCB_HD void sabSampleAlphaBeta( const SabT& T, double ekin, Rng& rng, double& alpha, double& beta, int& err )
  {
    const int n = T.negrid;
    const double* egrid = T.egrid;
    int iu = upperBound( egrid, 0, n, ekin );
    int isampler;
    bool ultra_small_ekin_mode = false;
    const double ultra_small_ekin = egrid[0];
    if ( iu == n ) {
      if ( !kHighE ) {
        err |= ERR_SAB_ROUTING;
        alpha = -1.0; beta = 0.0;
        return;
      }
      if ( sabSampleHighE( T, ekin, rng, alpha, beta, err ) )
        return; // (the reference returns whenever alpha>=0; errors flagged separately)
      ekin = egrid[n-1];
      isampler = n-1;
    } else {
      isampler = sabPickSampler( T, egrid, ekin, ultra_small_ekin_mode );
    }
    const SabEPoint ep = T.ep[isampler];
    const double ekin_div_kT = ekin / T.kT;
    const double sampling_ekin_div_kT = ( ultra_small_ekin_mode ? ultra_small_ekin/T.kT : ekin_div_kT );
    for ( int loop = 0; loop < 100; ++loop ) {
      sabSampleAtE( T, ep, sampling_ekin_div_kT, rng, alpha, beta, err );
      if ( err & ERR_SAB_LOOP_INNER )
        return;
      if ( beta < -ekin_div_kT )
        continue;
      AlphaLimits alims = getAlphaLimits( ekin_div_kT, beta );
      if ( inInterval( alims.first, alims.second, alpha ) )
        return;
      if ( ultra_small_ekin_mode ) {
        alpha = alims.first + rng.generate()*( alims.second - alims.first );
        return;
      }
    }
    err |= ERR_SAB_LOOP_OUTER;
  }

  // Tail of SABSampler::sampleDeltaEMu (NCSABSampler.cc:229-236) + SABScatter::sampleScatterIsotropic
  // (NCSABScatter.cc:93-100): (alpha,beta) at the neutron's own energy -> (E_final, mu).
  NCB_HD void sabFinishScatter( const SabT& T, double ekin, double alpha, double beta, Rng& rng,
                                double& ekin_out, double& mu, int& err )
  {
    double deltaE;
    if ( muIsotropicAtBeta( beta, ekin/T.kT ) ) {
      deltaE = beta*T.kT;
      mu = rng.generate()*2.0 - 1.0;
    } else {
      alphaBetaToDeltaEMu( alpha, beta, ekin, T.kT, deltaE, mu, err );
      if ( err & ERR_KIN_DENOM ) {
        ekin_out = -1.0; mu = -999.0;
        return;
      }
    }
    ekin_out = dmax( 0.0, ekin + deltaE );
  }

  // Table path (E < Emax) of SABScatter::sampleScatterIsotropic through the short-chain functions: the sequence of
  // attempts k_sample_sab_refill runs for one neutron, as one loop (host build: tests/hostsim).
  NCB_HD void sabSampleScatterFast( const SabT& T, double ekin, Rng& rng, double& ekin_out, double& mu, int& err )
  {
    bool ultra = false;
    const int ie = sabPickSampler( T, T.egrid, ekin, ultra );
    const SabEPoint ep = T.ep[ie];
    const double ekin_div_kT = ekin / T.kT;
    const double sampling_ediv = ultra ? T.egrid[0] / T.kT : ekin_div_kT;
    int inner = 0, outer = 0;
    ekin_out = -1.0; mu = -999.0;
    while ( true ) {
      double alpha = 0.0, beta = 0.0;
      bool inner_ok = true;
      if ( ep.npts != 0 )
        inner_ok = sabAttemptFast( T, ie, ep, sampling_ediv, rng, alpha, beta, err );
      if ( !inner_ok ) {
        if ( ++inner == 100 ) { err |= ERR_SAB_LOOP_INNER; return; }
        continue;
      }
      inner = 0;
      bool acc = false;
      if ( !( beta < -ekin_div_kT ) ) {
        AlphaLimits al = getAlphaLimits( ekin_div_kT, beta );
        if ( inInterval( al.first, al.second, alpha ) ) {
          acc = true;
        } else if ( ultra ) {
          alpha = al.first + rng.generate()*( al.second - al.first );
          acc = true;
        }
      }
      if ( acc ) {
        sabFinishScatter( T, ekin, alpha, beta, rng, ekin_out, mu, err );
        return;
      }
      if ( ++outer == 100 ) { err |= ERR_SAB_LOOP_OUTER; return; }
    }
  }

  // Table sampling for a neutron above Emax whose high-E analysis (sabSampleHighE) asked for
  // the tabulated kernel at E=Emax (NCSABSampler.cc:173-178): (alpha,beta) from the last
  // overlay sampler with ekin:=Emax, then the outcome at the neutron's own energy.
  NCB_HD void sabSampleScatterAtEmax( const SabT& T, double ekin_orig, Rng& rng, double& ekin_out, double& mu, int& err )
  {
    double alpha = 0.0, beta = 0.0;
    const double emax = T.egrid[T.negrid-1];
    // upper_bound(egrid, Emax) == end, so run the in-grid branch by hand on the last sampler:
    const SabEPoint ep = T.ep[T.negrid-1];
    const double ekin_div_kT = emax / T.kT;
    bool ok = false;
    for ( int loop = 0; loop < 100; ++loop ) {
      sabSampleAtE( T, ep, ekin_div_kT, rng, alpha, beta, err );
      if ( err & ERR_SAB_LOOP_INNER )
        break;
      if ( beta < -ekin_div_kT )
        continue;
      AlphaLimits alims = getAlphaLimits( ekin_div_kT, beta );
      if ( inInterval( alims.first, alims.second, alpha ) ) { ok = true; break; }
    }
    if ( !ok ) {
      if ( !( err & ERR_SAB_LOOP_INNER ) ) err |= ERR_SAB_LOOP_OUTER;
      ekin_out = -1.0; mu = -999.0;
      return;
    }
    sabFinishScatter( T, ekin_orig, alpha, beta, rng, ekin_out, mu, err );
  }

  // SABSampler::sampleDeltaEMu (NCSABSampler.cc:229-236) + SABScatter::sampleScatterIsotropic
  // (NCSABScatter.cc:93-100)
  template <bool kHighE = true>
  NCB_HD void sabSampleScatter( const SabT& T, double ekin, Rng& rng, double& ekin_out, double& mu, int& err )
  {
    double alpha = 0.0, beta = 0.0;
    sabSampleAlphaBeta<kHighE>( T, ekin, rng, alpha, beta, err );
    if ( err & ( ERR_SAB_LOOP_INNER | ERR_SAB_LOOP_OUTER | ERR_SAB_DISCARD | ERR_SAB_ROUTING ) ) {
      ekin_out = -1.0; mu = -999.0;
      return;
    }
    double deltaE;
    if ( muIsotropicAtBeta( beta, ekin/T.kT ) ) {
      deltaE = beta*T.kT;
      mu = rng.generate()*2.0 - 1.0;
    } else {
      alphaBetaToDeltaEMu( alpha, beta, ekin, T.kT, deltaE, mu, err );
      if ( err & ERR_KIN_DENOM ) {
        ekin_out = -1.0; mu = -999.0;
        return;
      }
    }
    ekin_out = dmax( 0.0, ekin + deltaE );
  }

}
