// ncb_phys_lcbragg.cuh -- layered crystals (LCBragg, lcmode 0: e.g. pyrolytic graphite).
//
// Restates, per neutron, what the reference keeps in LCHelper::Cache:
//   LCHelper::forceUpdateCache / crossSection / genScatter  src/extd_utils/NCLCUtils.cc:354-439,534-678
//   LCROIFinder::findROIs                                   src/extd_utils/NCLCUtils.cc:172-326
//   LCStdFrame::calcXS[_OnAxis] / calcXSIntegral / genScat  src/extd_utils/NCLCUtils.cc:684-772
//   LCStdFrameIntegrator (Romberg over the crystallite rotation phi)  :455-519
//   LCBragg::crossSection / sampleScatter                   src/lcbragg/NCLCBragg.cc:110-141
// The mosaicity model (GaussMos / GaussOnSphere) is the one of ncb_phys_scbragg.cuh; its tables sit in Material::sc.
//
// A "ROI" is a range of crystallite rotations phi (or one of two degenerate cases) for which one plane set
// contributes; the list of ROIs of a neutron is a pure function of (wavelength, |cos(angle to the layer axis)|),
// both discretised to 2^-40 as in the reference's cache signature.
#pragma once
#include "ncb_phys_scbragg.cuh"

namespace ncb {

  constexpr double kInvPi = 0.318309886183790671537767526745028724068919291;
  constexpr double kLcDiscrFact = 1099511627776.0;   // NCRYSTAL_LCUTILS_DISCRFACT = 2^40

  // LCdiscretizeValue / LCdediscretizeValue, ref: NCLCUtils.cc:44-52
  NCB_HD double lcDiscretise( double value )
  {
    const uint64_t d = (uint64_t)( value*kLcDiscrFact + 0.5 );
    return (double)d * ( 1.0 / kLcDiscrFact );
  }

  // Vector::unit, ref: NCVector.hh:172-181
  NCB_HD Vec3 vunit( const Vec3& v )
  {
    const double m2 = vmag2( v );
    if ( m2 == 1.0 ) return v;
    const double f = 1.0 / sqrt( m2 );
    return { v.x*f, v.y*f, v.z*f };
  }

  struct LcNeutron { double wl, c3, s3; };   // LCHelper::Cache::m_wl / m_c3 / m_s3

  // (wl, c3, s3) of the reference's cache for (ekin, unit direction); false: no scattering possible (wl not > 0)
  NCB_HD bool lcNeutronPars( const LcBraggT& L, double ekin, const Vec3& u, LcNeutron& N )
  {
    const double wl_raw = ekin ? sqrt( kWl2Ekin / ekin ) : kInf;    // NeutronWavelength{ekin}, NCDefs.hh:840-845
    if ( !( wl_raw > 0.0 ) || !( wl_raw < 1e7 ) )
      return false;
    const double c3_raw = L.ax*u.x + L.ay*u.y + L.az*u.z;
    N.wl = lcDiscretise( wl_raw );
    N.c3 = dmin( lcDiscretise( fabs( c3_raw ) ), 1.0 );
    N.s3 = sqrt( fabs( 1.0 - N.c3*N.c3 ) );
    return true;
  }

  // LCROI, ref: NCLCUtils.hh:52-77.  Degenerate cases have rotmin == rotmax: 0 = plane normal on the layer axis,
  // pi = neutron on the layer axis.
  struct LcRoi {
    double rotmin, rotmax;
    int ips;      // plane-set index
    int sign;     // +1 normal, -1 anti-normal
  };

  // LCROIFinder::findROIs for one plane set, ref: NCLCUtils.cc:172-326.  The reference first tests with an upper
  // estimate of sin(alpha2) and repeats the test with the exact value when that passes; testing with the exact value
  // directly selects the same plane sets.  Returns the number of ROIs written (0..2), normal before anti-normal.
  NCB_HD int lcFindROIs( const double* P, int ips, const LcNeutron& N, double cta, double sta, LcRoi* out )
  {
    const double c1minus = P[4], c1plus = P[5];
    const double c2 = P[1] * N.wl;
    const double c23 = N.c3*c2;
    const double s2 = sqrt( 1.0 - c2*c2 );
    const double s23 = N.s3*s2;
    const double lo = c23 - s23, hi = c23 + s23;
    if ( c1plus > hi || lo > c1minus )                   // intervalsDisjoint(lo,hi,c1plus,c1minus)
      return 0;
    const bool anti = ( -c1minus <= hi ) && ( lo <= -c1plus );   // intervalsOverlap(lo,hi,-c1minus,-c1plus)
    int n = 0;
    const double s1 = P[3];
    if ( !s1 ) {
      out[n++] = { 0.0, 0.0, ips, 1 };
      if ( anti ) out[n++] = { 0.0, 0.0, ips, -1 };
      return n;
    }
    if ( fabs( N.s3 ) < 1e-10 ) {
      out[n++] = { kPi, kPi, ips, 1 };
      if ( anti ) out[n++] = { kPi, kPi, ips, -1 };
      return n;
    }
    const double c2ta = c2 * cta;
    const double s2ta = s2 * sta;
    const double c2low = c2ta - s2ta;
    const double c2high = ( s2 < sta ? 1.0 : c2ta + s2ta );
    const double c1 = P[2];
    const double a = c1 * N.c3;
    const double invb = s1 * N.s3;
    const double b = 1.0 / invb;
    const double mab = -a*b;
    const double c2lowb = c2low*b;
    const double c2highb = c2high*b;
    double cosphi1 = dmax( -1.0, dmin( 1.0, mab + c2lowb ) );
    double cosphi2 = dmax( -1.0, dmin( 1.0, mab + c2highb ) );
    const double mindist = 1e-10;
    if ( fabs( cosphi1 - cosphi2 ) > mindist )
      out[n++] = { acos( dmax( cosphi1, cosphi2 ) ), acos( dmin( cosphi1, cosphi2 ) ), ips, 1 };
    if ( anti ) {
      cosphi1 = dmax( -1.0, dmin( 1.0, mab - c2lowb ) );
      cosphi2 = dmax( -1.0, dmin( 1.0, mab - c2highb ) );
      if ( fabs( cosphi1 - cosphi2 ) > mindist )
        out[n++] = { acos( dmax( cosphi1, cosphi2 ) ), acos( dmin( cosphi1, cosphi2 ) ), ips, -1 };
    }
    return n;
  }

  // LCStdFrameIntegrator, ref: NCLCUtils.cc:455-519
  struct LcPhiIntegrand {
    const ScBraggT& S;
    InteractionPars ip;
    double sn_s3, cn_c3, acc;
    NCB_HD void evalMany( double* fvals, unsigned n, double offset, double delta )
    {
      CosSinGridGen grid( n, offset, delta );
      unsigned i = 0;
      do {
        const double cosgamma = sn_s3 * grid.c + cn_c3;
        fvals[i++] = gmRawXS( S, ip, cosgamma );
      } while ( grid.step() );
    }
    NCB_HD double evalManySum( unsigned n, double offset, double delta )
    {
      CosSinGridGen grid( n, offset, delta );
      double sum = 0.;
      do {
        const double cosgamma = sn_s3 * grid.c + cn_c3;
        sum += gmRawXS( S, ip, cosgamma );
      } while ( grid.step() );
      return sum;
    }
    NCB_HD bool accept( unsigned, double prev_estimate, double estimate ) const
    {
      return fabs( estimate - prev_estimate ) <= acc*fabs( estimate );
    }
  };
  NCB_HD void lcIntegrandInit( LcPhiIntegrand& f, const LcBraggT& L, const double* P, int sign, const LcNeutron& N )
  {
    f.ip.set( N.wl, P[1], P[6] );
    f.sn_s3 = P[3]*N.s3*(double)sign;
    f.cn_c3 = P[2]*(double)sign*N.c3;
    f.acc = L.acc;
  }

  // LCStdFrame::calcXS, ref: NCLCUtils.cc:723-735
  NCB_HD double lcCalcXS( const ScBraggT& S, const double* P, int sign, const LcNeutron& N, double cosphi )
  {
    const double cosgamma = ( P[3] * N.s3 * cosphi + P[2] * N.c3 )*(double)sign;
    InteractionPars ip;
    ip.set( N.wl, P[1], P[6] );
    return gmRawXS( S, ip, cosgamma );
  }

  // cross-section contribution of one ROI, ref: NCLCUtils.cc:412-434.  err: ERR_LC_ROMBERG when the phi integration
  // does not converge (the reference throws CalcError, NCRomberg.cc:48-61).
  NCB_HD_NOINLINE double lcRoiXS( const ScBraggT& S, const LcBraggT& L, const LcNeutron& N, const LcRoi& roi, int& err )
  {
    const double* P = L.planes + kLcPlaneStride*roi.ips;
    if ( roi.rotmin == roi.rotmax ) {
      if ( roi.rotmax == 0.0 ) {
        // LCStdFrame::calcXS_OnAxis, ref: NCLCUtils.cc:689-696
        InteractionPars ip;
        ip.set( N.wl, P[1], P[6] );
        return gmRawXS( S, ip, (double)roi.sign * N.c3 );
      }
      return lcCalcXS( S, P, roi.sign, N, 0.0 );
    }
    LcPhiIntegrand f{ S };
    lcIntegrandInit( f, L, P, roi.sign, N );
    bool converged = true;
    const double r = rombergIntegrate( f, roi.rotmin, roi.rotmax, converged ) * kInvPi;
    if ( !converged ) err |= ERR_LC_ROMBERG;
    return r;
  }

  // Thread-level walk over the ROIs in the reference's order.  mode 0: total and count.  mode 1: stop at the ROI
  // selected by `choice` among the cumulative values (pickRandIdxByWeight: '>' for n<5, lower_bound otherwise; the
  // last ROI when none qualifies).
  struct LcWalk {
    double sum;
    int n;
    LcRoi chosen;
  };
  NCB_HD void lcWalk( const ScBraggT& S, const LcBraggT& L, const LcNeutron& N, int mode, bool linear, double choice,
                      LcWalk& W, int& err )
  {
    W.sum = 0.0; W.n = 0; W.chosen = { 0.0, 0.0, 0, 1 };
    for ( int ips = 0; ips < L.nplanes; ++ips ) {
      const double* P = L.planes + kLcPlaneStride*ips;
      if ( N.wl > P[0] )
        break;
      LcRoi r[2];
      const int nr = lcFindROIs( P, ips, N, S.cta, S.sta, r );
      for ( int k = 0; k < nr; ++k ) {
        W.sum += lcRoiXS( S, L, N, r[k], err );
        ++W.n;
        if ( mode ) {
          W.chosen = r[k];
          if ( linear ? ( W.sum > choice ) : !( W.sum < choice ) )
            return;
        }
      }
    }
  }

  // LCBragg::crossSection (lcmode 0), ref: NCLCBragg.cc:110-125 + LCHelper::crossSection NCLCUtils.cc:436-440.
  // Returns the leaf's cross section; raw_sum / n_roi are what the sampling needs (m_roixs_commul.back(), size()).
  NCB_HD double lcXS( const ScBraggT& S, const LcBraggT& L, double ekin, const Vec3& dir, double& raw_sum, int& n_roi, int& err )
  {
    raw_sum = 0.0; n_roi = 0;
    if ( ekin < L.ekin_low )
      return 0.0;
    LcNeutron N;
    if ( !lcNeutronPars( L, ekin, vunit( dir ), N ) )
      return 0.0;
    LcWalk W;
    lcWalk( S, L, N, 0, false, 0.0, W, err );
    raw_sum = W.sum; n_roi = W.n;
    return W.n ? L.xsfact * W.sum : 0.0;
  }

  // Scattering in one chosen ROI + rotation to the lab frame, ref: LCHelper::genScatter NCLCUtils.cc:556-677
  // (overlay construction :580-612, genPhiVal :521-528, LCStdFrame::genScat[_OnAxis] :698-721,752-772).
  NCB_HD_NOINLINE void lcGenScatterRoi( const ScBraggT& S, const LcBraggT& L, const LcNeutron& N, const LcRoi& roi,
                                        const Vec3& indir, Rng& rng, Vec3& outdir )
  {
    const double* P = L.planes + kLcPlaneStride*roi.ips;
    const double sgn = (double)roi.sign;
    const Vec3 indir_std = { -N.s3, 0., -N.c3 };
    if ( roi.rotmin == roi.rotmax && roi.rotmax == 0.0 ) {
      const Vec3 pn = { 0., 0., sgn };
      gmGenScat( S, rng, pn, P[1], N.wl, indir_std, outdir );
    } else {
      double phi, cosphi;
      if ( roi.rotmin == roi.rotmax ) {
        phi = rng.generate()*kPi;
        cosphi = cos_mpipi( phi );
      } else {
        // overlay: xs at the edges of 8 phi bins, per bin max of its two edges * 1.7 + 2 % of the largest value,
        // accumulated in single precision like the reference's float array
        constexpr int ndata = 8;
        const double length = roi.rotmax - roi.rotmin;
        double tmp[ndata+1];
        {
          LcPhiIntegrand f{ S };
          lcIntegrandInit( f, L, P, roi.sign, N );
          f.evalMany( tmp, ndata+1, roi.rotmin, length/ndata );
        }
        double maxval = 0.0;
        for ( int i = 0; i <= ndata; ++i ) maxval = dmax( maxval, tmp[i] );
        const double safety_offset = 0.02 * maxval;
        const double safety_factor = 1.7;
        float data[ndata];
        float sum = 0.0f;
        for ( int i = 0; i < ndata; ++i )
          data[i] = ( sum = (float)( (double)sum + ( dmax( tmp[i], tmp[i+1] ) * safety_factor + safety_offset ) ) );
        int triesleft = 1000;
        phi = roi.rotmin; cosphi = 1.0;
        while ( triesleft-- ) {
          // genPhiVal
          const double target = (double)data[ndata-1] * rng.generate();
          int ichoice = 0;
          while ( ichoice < ndata && (double)data[ichoice] < target ) ++ichoice;    // std::lower_bound
          if ( ichoice > ndata-1 ) ichoice = ndata-1;
          const double overlay_at_phi = ichoice ? (double)data[ichoice] - (double)data[ichoice-1] : (double)data[ichoice];
          const double rel_phi_pos = ( ichoice + rng.generate() )/ndata;
          phi = roi.rotmin + rel_phi_pos*length;
          cosphi = cos_mpipi( phi );
          const double xsphi = lcCalcXS( S, P, roi.sign, N, cosphi );
          if ( xsphi > overlay_at_phi * rng.generate() )
            break;
        }
      }
      const double sinphisign = ( rng.generate() > 0.5 ) ? 1.0 : -1.0;   // RNGStream::coinflip of a non-builtin stream
      const double sinphi = sinphisign*sqrt( 1.0 - cosphi*cosphi );
      // LCStdFrame::normalInStdFrame
      const double ns1 = sgn*P[3];
      const Vec3 pn = { ns1*cosphi, ns1*sinphi, sgn*P[2] };
      gmGenScat( S, rng, pn, P[1], N.wl, indir_std, outdir );
    }
    const double indirsign = ( L.ax*indir.x + L.ay*indir.y + L.az*indir.z ) >= 0.0 ? 1.0 : -1.0;
    const Vec3 axis = { L.ax*indirsign, L.ay*indirsign, L.az*indirsign };
    rotateToFrame( N.s3, N.c3, indir, axis, outdir, rng );
    outdir.x *= -1.0; outdir.y *= -1.0; outdir.z *= -1.0;
  }

  // LCBragg::sampleScatter (lcmode 0), ref: NCLCBragg.cc:127-141.  raw_sum / n_roi: results of lcXS for the same
  // (E, dir) (the reference's cache).  E is unchanged (elastic).
  NCB_HD void lcSampleScatter( const ScBraggT& S, const LcBraggT& L, double ekin, const Vec3& indir_raw, double raw_sum, int n_roi,
                               Rng& rng, Vec3& outdir, int& err )
  {
    outdir = indir_raw;
    if ( ekin < L.ekin_low )
      return;
    const Vec3 u = vunit( indir_raw );
    LcNeutron N;
    if ( !lcNeutronPars( L, ekin, u, N ) )
      return;
    outdir = u;
    if ( n_roi <= 0 || !raw_sum )
      return;
    LcWalk W;
    if ( n_roi == 1 )
      lcWalk( S, L, N, 1, true, -1.0, W, err );          // first ROI, no draw
    else
      lcWalk( S, L, N, 1, n_roi < 5, raw_sum * rng.generate(), W, err );
    lcGenScatterRoi( S, L, N, W.chosen, u, rng, outdir );
  }

}
