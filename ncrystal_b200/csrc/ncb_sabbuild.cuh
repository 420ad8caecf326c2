// ncb_sabbuild.cuh -- native builder of the S(alpha,beta) sampler tables.
//
// The reference builds, per material and on one CPU thread, for each of the
// ~300 energy grid points a SABSamplerAtE_Alg1 (beta CDF + per-beta-row alpha
// sampling info) in SABIntegrator::Impl::analyseEnergyPoint
// (ref: src/sab/NCSABIntegrator.cc:352-559) on top of derived per-kernel data
// (log S and cumulative alpha integrals, :105-141).  These tables are private to
// the reference, so the product derives them itself from the SABData it is
// given -- on the device, at handle creation:
//   stage 0  (one thread per element / per beta row)  logsab, alphaintegrals_cumul
//   stage 1  (one thread per (energy point, beta row)) kinematic range, tailed
//            breakdown, xs at that beta, AlphaSampleInfo
//   stage 2  (one thread per energy point)  beta sampler (PointwiseDist ctor,
//            ref: src/utils/NCPointwiseDist.cc:32-74) and the total xs
// The row-parallel form relies on activeGridRanges' result being a pure function
// of (alow,aupp) per row (ref: src/sab/NCSABUtils.cc:460-536: its iterator
// "hinting" only speeds up the search).
#pragma once
#include "ncb_phys_basic.cuh"

namespace ncb {

  // integrateAlphaInterval_fast, ref: include/NCrystal/internal/sab/NCSABUtils.hh:236-258
  NCB_HD double integrateAlphaIntervalFast( double a1, double s1, double a2, double s2, double logs1, double logs2 )
  {
    const double da = a2 - a1;
    const double ps = s1 + s2;
    const double ds = s2 - s1;
    if ( dmin(s1,s2) < 1e-300 )
      return 0.5*da*ps;
    if ( fabs(ds) > 0.006*ps )
      return da*ds/(logs2-logs1);
    const double y = ds / ps;
    const double c1 = 0.166666666666666666666666666666666666666666667;
    const double c2 = 0.0444444444444444444444444444444444444444444444;
    const double c3 = 0.0232804232804232804232804232804232804232804233;
    const double ysq = y*y;
    return da*ps*(0.5-ysq*(c1+ysq*(c2+ysq*c3)));
  }

  // interpolate_loglin_fallbacklinlin_fast, ref: NCSABUtils.hh:181-200
  NCB_HD double interpLogLinFast( double a, double fa, double b, double fb, double x, double logfa, double logfb )
  {
    const double bma = b - a;
    const double midpoint = 0.5 * ( b + a );
    const bool linlin_mode = ( fa*fb == 0.0 );
    if ( x < midpoint ) {
      const double r = (x-a) / bma;
      return ( linlin_mode ? ( fa + (fb-fa)*r ) : m_exp( logfa+(logfb-logfa)*r ) );
    } else {
      const double s = (b-x) / bma;
      return ( linlin_mode ? ( fb + (fa-fb)*s ) : m_exp( logfb+(logfa-logfb)*s ) );
    }
  }

  // stage 0a: m_log(S), ref: NCSABIntegrator.cc:113-117
  NCB_HD double sabLogS( double s ) { return s > 0.0 ? m_log(s) : -kInf; }

  // stage 0b: cumulative alpha integrals of one beta row, ref: NCSABIntegrator.cc:119-133
  NCB_HD void sabCumulRow( const double* agrid, const double* sab_row, const double* logsab_row, int nalpha, double* cumul_row )
  {
    double cumul = 0.0;
    cumul_row[0] = 0.0;
    for ( int ai = 0; ai+1 < nalpha; ++ai ) {
      const double integ = integrateAlphaIntervalFast( agrid[ai], sab_row[ai], agrid[ai+1], sab_row[ai+1],
                                                       logsab_row[ai], logsab_row[ai+1] );
      cumul_row[ai+1] = ( cumul += integ );
    }
  }

  struct SabRow {          // stage-1 result for one (energy point, beta row)
    int has_range;         // row has a kinematically accessible alpha grid range
    double xs;             // xs_at_this_beta
  };

  // TailPoint setter, ref: NCSABUtils.cc:576-594
  NCB_HD void sabSetTailPoint( const double* agrid, const double* sab, const double* logsab,
                               int aidx, double alpha, double& t_alpha, double& t_sval, double& t_logsval )
  {
    t_alpha = alpha;
    t_sval = interpLogLinFast( agrid[aidx], sab[aidx], agrid[aidx+1], sab[aidx+1], alpha, logsab[aidx], logsab[aidx+1] );
    t_logsval = m_log( dmax( t_sval, kDblMin ) );
  }

  // Stage 1.  Per-row part of activeGridRanges (NCSABUtils.cc:460-536), of
  // createTailedBreakdown (:538-633) and of the loop body of analyseEnergyPoint
  // (NCSABIntegrator.cc:441-507).  `info` is fully written in all cases.
  NCB_HD SabRow sabAnalyseRow( const double* agrid, int nalpha, const double* betaGrid,
                               const double* sab_all, const double* logsab_all, const double* cumul_all,
                               double ekin_div_kT, int ibeta, SabAlphaInfo& info )
  {
    SabRow out; out.has_range = 0; out.xs = 0.0;
    info.f_alpha = info.f_sval = info.f_logsval = 0.0;
    info.b_alpha = info.b_sval = info.b_logsval = 0.0;
    info.prob_front = info.prob_notback = 0.0;
    info.f_idx = info.b_idx = 0;
    info.pad = 0.0;

    const double beta = betaGrid[ibeta];
    double alow(-1.0), aupp(-2.0);
    if ( beta > -ekin_div_kT ) {
      AlphaLimits al = getAlphaLimits( ekin_div_kT, beta );
      alow = al.first; aupp = al.second;
    }
    const double agrid_front = agrid[0], agrid_back = agrid[nalpha-1];
    int aidx_low = nalpha, aidx_upp = nalpha; // "empty" marker
    if ( !( agrid_back <= alow || agrid_front >= aupp || aupp < alow ) ) {
      out.has_range = 1;
      // largest i with agrid[i] <= alow (0 if none)
      int il = upperBound( agrid, 0, nalpha, alow ) - 1;
      aidx_low = il > 0 ? il : 0;
      // smallest i with agrid[i] >= aupp (last if none)
      int iu = lowerBound( agrid, 0, nalpha, aupp );
      aidx_upp = iu < nalpha-1 ? iu : nalpha-1;
      if ( aidx_upp < aidx_low ) aidx_upp = aidx_low;
    }
    if ( !( beta > -ekin_div_kT ) )
      return out; // below the kinematic end point: never part of the sampler

    // analyseEnergyPoint loop body (for rows >= ibeta_low; harmless otherwise)
    double xs_front = 0.0, xs_middle = 0.0, xs_back = 0.0;
    int imiddle_low = 0, imiddle_upp = 0;
    bool narrow = false;
    double f_alpha = 0, f_sval = 0, f_logsval = 0, b_alpha = 0, b_sval = 0, b_logsval = 0;
    if ( aidx_upp > aidx_low && aupp > alow ) {
      const double* sab = sab_all + (size_t)ibeta*nalpha;
      const double* logsab = logsab_all + (size_t)ibeta*nalpha;
      const double* cumul = cumul_all + (size_t)ibeta*nalpha;
      // createTailedBreakdown
      const double alpha_low = dclamp( alow, agrid_front, agrid_back );
      const double alpha_upp = dclamp( aupp, agrid_front, agrid_back );
      if ( !( aidx_low == aidx_upp || alpha_low == alpha_upp ) ) {
        if ( aidx_low + 1 == aidx_upp ) {
          narrow = true;
          sabSetTailPoint( agrid, sab, logsab, aidx_low, alpha_low, f_alpha, f_sval, f_logsval );
          sabSetTailPoint( agrid, sab, logsab, aidx_low, alpha_upp, b_alpha, b_sval, b_logsval );
          xs_front = integrateAlphaIntervalFast( f_alpha, f_sval, b_alpha, b_sval, f_logsval, b_logsval );
        } else {
          imiddle_low = aidx_low;
          imiddle_upp = aidx_upp;
          if ( alpha_low >= agrid[aidx_low] ) {
            sabSetTailPoint( agrid, sab, logsab, aidx_low, alpha_low, f_alpha, f_sval, f_logsval );
            xs_front = integrateAlphaIntervalFast( f_alpha, f_sval, agrid[aidx_low+1], sab[aidx_low+1],
                                                   f_logsval, logsab[aidx_low+1] );
            ++imiddle_low;
          }
          if ( alpha_upp <= agrid[aidx_upp] ) {
            sabSetTailPoint( agrid, sab, logsab, aidx_upp-1, alpha_upp, b_alpha, b_sval, b_logsval );
            xs_back = integrateAlphaIntervalFast( agrid[aidx_upp-1], sab[aidx_upp-1], b_alpha, b_sval,
                                                  logsab[aidx_upp-1], b_logsval );
            --imiddle_upp;
          }
          xs_middle = ( imiddle_upp > imiddle_low ? cumul[imiddle_upp] - cumul[imiddle_low] : 0.0 );
        }
      }
    }
    const double xs = xs_front + xs_back + xs_middle;
    out.xs = xs;
    if ( xs > 0.0 ) {
      info.f_alpha = f_alpha; info.f_sval = f_sval; info.f_logsval = f_logsval;
      info.b_alpha = b_alpha; info.b_sval = b_sval; info.b_logsval = b_logsval;
      if ( narrow ) {
        info.prob_front = 1.0;
      } else {
        info.prob_front = xs_front/xs;
        info.prob_notback = 1.0 - xs_back/xs;
        info.f_idx = imiddle_low;
        info.b_idx = imiddle_upp;
      }
    } else {
      info.prob_front = 2.0;
      info.f_alpha = alow;
      info.b_alpha = aupp;
    }
    return out;
  }

  // Stage 2: one energy point.  rows[] = stage-1 results of that point (nbeta).
  // Writes ep, the beta sampler rows bx/bpdf/bcdf[off_b .. off_b+npts) and returns the
  // total cross section at this energy (to be compared with the reference's xs grid).
  // err!=0: layout assumption violated (see comment in body).
  NCB_HD double sabAssembleEPoint( const double* betaGrid, int nbeta, double kT, double bound_xs, double ekin,
                                   const SabRow* rows, uint32_t off_b, uint32_t off_i_base,
                                   SabEPoint& ep, double* bx, double* bpdf, double* bcdf, int& err )
  {
    const double ekin_div_kT = ekin / kT;
    double beta_lower_limit = -ekin_div_kT;
    bool starts_at_kinematic_endpoint = true;
    if ( beta_lower_limit < betaGrid[0] ) {
      const double c1 = betaGrid[0] - ( betaGrid[1] - betaGrid[0] ) * 1e-6;
      const double c2 = betaGrid[0] - fabs( betaGrid[0] ) * 1e-13;
      const double c3 = nextafter( betaGrid[0], beta_lower_limit );
      beta_lower_limit = dmin( c1, dmin( c2, c3 ) );
      starts_at_kinematic_endpoint = false;
    }
    int ibeta_low = nbeta;
    for ( int i = 0; i < nbeta; ++i )
      if ( rows[i].has_range ) { ibeta_low = i; break; }

    ep.npts = 0; ep.ibeta_off = 0; ep.off_b = off_b; ep.off_i = off_i_base; ep.first_bin_endpoint = 1.0; ep.guide = nullptr;
    if ( ibeta_low >= nbeta )
      return 0.0; // SABSamplerAtE_NoScatter

    if ( ibeta_low > 0 && beta_lower_limit < betaGrid[ibeta_low-1] ) {
      beta_lower_limit = betaGrid[ibeta_low-1];
      starts_at_kinematic_endpoint = false;
    }
    double prev_b = beta_lower_limit, prev_xs = 0.0;
    int np = 0;
    bx[np] = prev_b; bpdf[np] = prev_xs; ++np;
    StableSum xs_total_stable;
    bool next_bin_has_kinematic_endpoint = starts_at_kinematic_endpoint;
    for ( int ib = ibeta_low; ib < nbeta; ++ib ) {
      const double beta = betaGrid[ib];
      if ( beta == beta_lower_limit ) {
        // The reference skips such a row (NCSABIntegrator.cc:442-443) which would
        // desynchronise its own ibetaOffset bookkeeping; cannot happen for rows
        // >= ibeta_low (see activeGridRanges) -- flag it rather than guess.
        err = 1;
        continue;
      }
      const double xs_at_this_beta = rows[ib].xs;
      if ( next_bin_has_kinematic_endpoint ) {
        next_bin_has_kinematic_endpoint = false;
        const double db_real = ( beta - prev_b );
        prev_b -= db_real * ( 1.0 / 3.0 );
        bx[0] = prev_b;
      }
      xs_total_stable.add( 0.5 * ( beta - prev_b ) * ( xs_at_this_beta + prev_xs ) );
      prev_b = beta; prev_xs = xs_at_this_beta;
      bx[np] = prev_b; bpdf[np] = prev_xs; ++np;
    }
    double xs_total = xs_total_stable.sum() * bound_xs / ( 4*ekin_div_kT );
    if ( !( xs_total >= 0.0 ) )
      xs_total = 0.0;
    if ( xs_total == 0.0 )
      return 0.0; // NoScatter

    // PointwiseDist ctor, ref: NCPointwiseDist.cc:32-74
    StableSum totalArea;
    bcdf[0] = 0.0;
    for ( int i = 1; i < np; ++i ) {
      const double area = ( bx[i]-bx[i-1] ) * 0.5 * ( bpdf[i]+bpdf[i-1] );
      totalArea.add( area );
      bcdf[i] = totalArea.sum();
    }
    const double totalAreaVal = totalArea.sum();
    if ( !( totalAreaVal > 0.0 ) ) { err = 2; return xs_total; }
    const double normfact = 1.0/totalAreaVal;
    for ( int i = 0; i < np; ++i ) {
      bcdf[i] *= normfact;
      bpdf[i] *= normfact;
    }
    bcdf[np-1] = 1.0;

    ep.npts = np;
    ep.ibeta_off = ibeta_low;
    // ainfo is stored with absolute beta-row indexing (off_i_base + ibeta); the sampler
    // addresses it as off_i + (ibeta - ibeta_off).
    ep.off_i = off_i_base + (uint32_t)ibeta_low;
    ep.first_bin_endpoint = ( starts_at_kinematic_endpoint ? beta_lower_limit : 1.0 );
    return xs_total;
  }

  // ---- stage 3: guide tables (an acceleration structure of the product, not part of the reference)
  // beta guide of one energy point: g[b] = lower_bound( cdf, b/kSabGB )
  NCB_HD uint16_t sabBetaGuideEntry( const double* cdf, int npts, int b )
  {
    if ( npts <= 0 ) return 0;
    return (uint16_t)lowerBound( cdf, 0, npts, (double)b / (double)kSabGB );
  }
  // alpha guide of one beta row: g[b] = upper_bound( cumul_row, b/scale ), scale = kSabGA/cumul_row[last]
  NCB_HD double sabAlphaScale( const double* cumul_row, int nalpha )
  {
    const double tot = cumul_row[nalpha-1];
    return ( tot > 0.0 && isFinite( (double)kSabGA / tot ) ) ? (double)kSabGA / tot : 0.0;
  }
  NCB_HD uint16_t sabAlphaGuideEntry( const double* cumul_row, int nalpha, double scale, int b )
  {
    if ( !( scale > 0.0 ) ) return (uint16_t)( b == 0 ? 0 : nalpha );
    return (uint16_t)upperBound( cumul_row, 0, nalpha, (double)b / scale );
  }

  // ---- stage 4: gather-friendly copies for the table-sampling kernel (layout only, same values) + log guide
  NCB_HD SabHead sabMakeHead( const SabAlphaInfo& info, const double* cumul_row, int nalpha )
  {
    SabHead h;
    h.prob_front = info.prob_front; h.prob_notback = info.prob_notback;
    h.f_idx = (uint32_t)info.f_idx; h.b_idx = (uint32_t)info.b_idx;
    h.clow = cumul_row[info.f_idx]; h.cupp = cumul_row[info.b_idx];
    const double tot = cumul_row[nalpha-1];
    h.inv_total = ( tot > 0.0 && isFinite( 1.0/tot ) ) ? 1.0/tot : 0.0;
    return h;
  }
  NCB_HD void sabMakeTails( const SabAlphaInfo& info, SabTail* t2 )
  {
    t2[0].alpha = info.f_alpha; t2[0].sval = info.f_sval; t2[0].logsval = info.f_logsval; t2[0].pad = 0.0;
    t2[1].alpha = info.b_alpha; t2[1].sval = info.b_sval; t2[1].logsval = info.b_logsval; t2[1].pad = 0.0;
  }
  NCB_HD SabPoint sabMakePoint( const double* agrid, const double* sab, const double* logsab, const double* cumul, int nalpha, size_t k )
  {
    SabPoint p;
    p.alpha = agrid[k % (size_t)nalpha]; p.sab = sab[k]; p.logsab = logsab[k]; p.cumul = cumul[k];
    return p;
  }
  // Log guide of one row: g[k] = number of grid points whose own key (computed exactly like the sampler computes the
  // key of an area: cumul*inv_total) is below k.  Rounded multiplication by a positive constant is monotone, so for
  // an area with key k every point counted by g[k] has cumul < area and every point from g[k+1] on has cumul > area:
  // upper_bound(row, area) lies in [g[k], g[k+1]] -- exactly, no verification needed.
  NCB_HD uint16_t sabLogGuideEntry( const double* cumul_row, int nalpha, double inv_total, int k )
  {
    int lo = 0, hi = nalpha;     // keys are non-decreasing along the row: binary search for the first key >= k
    while ( lo < hi ) {
      const int mid = lo + ( ( hi - lo ) >> 1 );
      if ( sabLogKey( cumul_row[mid]*inv_total ) < k ) lo = mid + 1; else hi = mid;
    }
    return (uint16_t)lo;
  }

}
