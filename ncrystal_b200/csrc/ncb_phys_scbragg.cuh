// ncb_phys_scbragg.cuh -- oriented mosaic single-crystal Bragg diffraction, per neutron (E, dir).
// Restates
//   SCBragg::pimpl::updateCache / genScat / crossSection / sampleScatter   ref: src/scbragg/NCSCBragg.cc:225-330
//   GaussMos::InteractionPars::set, calcRawCrossSectionValue[Init],
//     calcCrossSections, genScat                                            ref: src/phys_utils/NCGaussMos.cc:116-280, NCGaussMos.hh:248-258
//   GaussOnSphere::circleIntegral[Slow], genPointOnCircle, evalCosX*        ref: NCGaussOnSphere.hh:165-213, NCGaussOnSphere.cc:380-508
//   GOSCircleInt (Romberg integrand on the spline LUT)                      ref: NCGaussOnSphere.cc:64-143
//   Romberg::integrate                                                      ref: src/utils/NCRomberg.cc:62-146
//   CosSinGridGen, sincos_* , cos_mpipi                                     ref: NCMath.hh:164-190,333-375,489-525, NCMath.cc:96-175
//   CubicSpline::evalUnbounded / SplinedLookupTable::eval                   ref: NCSpline.hh:122-160
//   rotateToFrame, PhiRot                                                   ref: src/utils/NCRotMatrix.cc:84-147, NCRotMatrix.hh:172-191
//   randDirectionGivenScatterMu, randPointOnUnitCircle                      ref: src/utils/NCRandUtils.cc:51-112
// The reference keeps the per-(E,dir) list of contributing normals in a cache object; here the
// list is never materialised: the cross section is one pass over the reflection families, and
// sampling re-walks them up to the chosen entry.
#pragma once
#include "ncb_phys_basic.cuh"

namespace ncb {

  constexpr double kPiHalf = 1.5707963267948966192313216916397514420985847;
  constexpr double k2Pi    = 6.2831853071795864769252867665590057683943388;
  constexpr double kArcSec = 0.00000484813681109535993589914102357947975956353302;
  constexpr double kDblEps = 2.220446049250313e-16;

  struct Vec3 { double x, y, z; };
  NCB_HD double vdot( const Vec3& a, const Vec3& b ) { return a.x*b.x + a.y*b.y + a.z*b.z; }
  NCB_HD double vmag2( const Vec3& a ) { return a.x*a.x + a.y*a.y + a.z*a.z; }
  NCB_HD Vec3 vcross( const Vec3& a, const Vec3& o )   // Vector::operator*(Vector), NCVector.hh:141-146
  {
    return { a.y*o.z - a.z*o.y, a.z*o.x - a.x*o.z, a.x*o.y - a.y*o.x };
  }
  // Vector::normalise, ref: NCVector.hh:197-209 (null / infinite vectors: caller's responsibility)
  NCB_HD void vnormalise( Vec3& v )
  {
    const double themag2 = vmag2( v );
    constexpr double one_low = 1.0 - 2.0 * kDblEps;
    constexpr double one_high = 1.0 + 2.0 * kDblEps;
    if ( themag2 >= one_low && themag2 <= one_high )
      return;
    const double ff = 1.0 / sqrt( themag2 );
    v.x *= ff; v.y *= ff; v.z *= ff;
  }

  // ---- trig approximations (NCMath.cc:96-175, NCMath.hh:358-375)
  NCB_HD void sincos_mpi2pi2( double A, double& cosA, double& sinA )
  {
    double x = 0.5*A;
    double mx2 = -x*x;
    double s2 =  x*(1.0 + mx2 * ( 1.66666666666666666666666666666666666666666667e-1
                        + mx2 * ( 8.33333333333333333333333333333333333333333333e-3
                        + mx2 * ( 1.98412698412698412698412698412698412698412698e-4
                        + mx2 * ( 2.75573192239858906525573192239858906525573192e-6
                        + mx2 * ( 2.50521083854417187750521083854417187750521084e-8
                        + mx2 * ( 1.60590438368216145993923771701549479327257105e-10
                        + mx2 * ( 7.64716373181981647590113198578807044415510024e-13
                                  ))))))));
    double c2m1 = mx2 * ( 0.5
                + mx2 * ( 4.16666666666666666666666666666666666666666667e-2
                + mx2 * ( 1.38888888888888888888888888888888888888888889e-3
                + mx2 * ( 2.48015873015873015873015873015873015873015873e-5
                + mx2 * ( 2.75573192239858906525573192239858906525573192e-7
                + mx2 * ( 2.08767569878680989792100903212014323125434237e-9
                + mx2 * ( 1.14707455977297247138516979786821056662326504e-11
                + mx2 * ( 4.77947733238738529743820749111754402759693765e-14
                          ))))))));
    double k = 2.0*c2m1;
    sinA = (k+2.0)*s2;
    cosA = 1.0+k*(c2m1+2.0);
  }
  NCB_HD void sincos_mpi8pi8( double A, double& cosA, double& sinA )
  {
    double x = 0.5*A;
    double mx2 = -x*x;
    double s2 = x*(1.0 + mx2 * ( 1.66666666666666666666666666666666666666666667e-1
                       + mx2 * ( 8.33333333333333333333333333333333333333333333e-3
                       + mx2 * ( 1.98412698412698412698412698412698412698412698e-4
                       + mx2 * ( 2.75573192239858906525573192239858906525573192e-6
                       + mx2 * ( 2.50521083854417187750521083854417187750521084e-8
                                 ))))));
    double c2m1 = mx2 * ( 0.5
                + mx2 * ( 4.16666666666666666666666666666666666666666667e-2
                + mx2 * ( 1.38888888888888888888888888888888888888888889e-3
                + mx2 * ( 2.48015873015873015873015873015873015873015873e-5
                + mx2 * ( 2.75573192239858906525573192239858906525573192e-7
                          )))));
    double k = 2.0*c2m1;
    sinA = (k+2.0)*s2;
    cosA = 1.0+k*(c2m1+2.0);
  }
  NCB_HD void sincos_0pi32( double A, double& cosA, double& sinA )
  {
    sincos_mpi2pi2( dmin( A, kPi-A ), cosA, sinA );
    cosA = copysign( cosA, kPiHalf-A );
  }
  NCB_HD_NOINLINE double cos_mpipi( double A )
  {
    double Aabs = fabs(A);
    double x = dmin( Aabs, kPi-Aabs );
    double mx2 = -x*x;
    double c = 1.0 + mx2 * ( 0.5
               + mx2 * ( 4.16666666666666666666666666666666666666666667e-2
               + mx2 * ( 1.38888888888888888888888888888888888888888889e-3
               + mx2 * ( 2.48015873015873015873015873015873015873015873e-5
               + mx2 * ( 2.75573192239858906525573192239858906525573192e-7
               + mx2 * ( 2.08767569878680989792100903212014323125434237e-9
               + mx2 * ( 1.14707455977297247138516979786821056662326504e-11
               + mx2 * ( 4.77947733238738529743820749111754402759693765e-14
               + mx2 * ( 1.56192069685862264622163643500573334235194041e-16
               + mx2 * ( 4.1103176233121648584779906184361403746103695e-19
               + mx2 * ( 8.89679139245057328674889744250246834331248809e-22
                )))))))))));
    return copysign( c, kPiHalf-Aabs );
  }

  // ---- spline lookup table, ref: NCSpline.hh:122-135,158-160
  NCB_HD double splineEval( const SplineLutT& L, double xin )
  {
    const double x = ( xin - L.a ) * L.invdelta;
    // static_cast<std::size_t>(x): truncation; tiny negative -> 0 on both x86-64 and sm_100
    long long ll = (long long)x;
    if ( ll < 0 ) ll = 0;
    const int idx = ll < (long long)L.nm2 ? (int)ll : L.nm2;
    const double b = x - idx;
    const double a = 1.0 - b;
    const double* it = L.data + 2*idx;
    double tmp = a * it[0];
    double tmp2 = ( a*a*a - a ) * it[1];
    tmp += b * it[2];
    tmp2 += ( b*b*b - b ) * it[3];
    return tmp + 0.166666666666666666666666666666666666666666666666666667 * tmp2;
  }

  // GaussOnSphere::evalCosXInRange / evalCosX, ref: NCGaussOnSphere.hh:165-181
  NCB_HD double gosEvalCosXInRange( const ScBraggT& S, double cosx ) { return dmax( 0.0, splineEval( S.evalcosx, cosx ) ); }
  NCB_HD double gosEvalCosX( const ScBraggT& S, double cosx ) { return cosx >= S.evalcosx.a ? gosEvalCosXInRange( S, cosx ) : 0.0; }

  // CosSinGridGen, ref: NCMath.hh:164-190,489-525
  struct CosSinGridGen {
    double c, s, cd, sd, phimax, negdelta;
    unsigned left, recalc;
    NCB_HD CosSinGridGen( unsigned n, double offset, double delta )
    {
      left = n-1;
      recalc = ( ( 127u + ( n/128u )*128u ) - n );
      phimax = offset + (n-1)*delta;
      negdelta = -delta;
      sincos_0pi32( offset, c, s );
      sincos_mpi8pi8( delta, cd, sd );
    }
    NCB_HD bool step()
    {
      if ( !left ) return false;
      --left;
      if ( ( left + recalc ) % 128u ) {
        const double cc = c*cd - s*sd;
        s = c*sd + s*cd;
        c = cc;
      } else {
        const double v = phimax + negdelta*left;
        c = cos(v); s = sin(v);   // NCrystal::sincos = std::cos/std::sin (NCMath.hh:333-337)
      }
      return true;
    }
  };

  // GOSCircleInt::evalFuncMany / evalFuncManySum, ref: NCGaussOnSphere.cc:83-113
  NCB_HD void gosEvalMany( const ScBraggT& S, double sasg, double cacg, double* fvals, unsigned n, double offset, double delta )
  {
    CosSinGridGen grid( n, offset, delta );
    unsigned i = 0;
    do {
      const double cb = sasg * grid.c + cacg;
      fvals[i++] = gosEvalCosXInRange( S, cb );
    } while ( grid.step() );
  }
  NCB_HD double gosEvalManySum( const ScBraggT& S, double sasg, double cacg, unsigned n, double offset, double delta )
  {
    CosSinGridGen grid( n, offset, delta );
    double sum = 0.;
    do {
      const double cb = sasg * grid.c + cacg;
      sum += gosEvalCosXInRange( S, cb );
    } while ( grid.step() );
    return sum;
  }
  // GOSCircleInt::accept, ref: NCGaussOnSphere.cc:115-141 (the one-time warning is not reproduced)
  NCB_HD bool gosAccept( double acc, unsigned level, double prev_estimate, double estimate )
  {
    if ( fabs( prev_estimate - estimate ) <= acc*fabs( estimate ) )
      return true;
    if ( level < 11 )
      return false;
    return true;
  }

  // (rombergIntegrate: ncb_common.cuh)

  // GOSCircleInt as integrand, ref: NCGaussOnSphere.cc:64-143
  struct GosCircleIntegrand {
    const ScBraggT& S;
    double sasg, cacg, acc;
    NCB_HD void evalMany( double* fvals, unsigned n, double offset, double delta ) const { gosEvalMany( S, sasg, cacg, fvals, n, offset, delta ); }
    NCB_HD double evalManySum( unsigned n, double offset, double delta ) const { return gosEvalManySum( S, sasg, cacg, n, offset, delta ); }
    NCB_HD bool accept( unsigned level, double prev_estimate, double estimate ) const { return gosAccept( acc, level, prev_estimate, estimate ); }
  };
  NCB_HD_NOINLINE double gosRomberg( const ScBraggT& S, double sasg, double cacg, double acc, double a, double b )
  {
    GosCircleIntegrand f{ S, sasg, cacg, acc };
    bool converged = true;   // (gosAccept is always true for level >= 11)
    return rombergIntegrate( f, a, b, converged );
  }

  // GaussOnSphere::circleIntegralSlow, ref: NCGaussOnSphere.cc:380-433
  NCB_HD_NOINLINE double gosCircleIntegralSlow( const ScBraggT& S, double cg, double sg, double ca, double sa )
  {
    const double sasg = sa*sg;
    const double cacg = ca*cg;
    const double cd = cacg + sasg;
    if ( cd <= S.cta )
      return 0.0;
    if ( sasg < 1e-14 )
      return k2Pi*sa*gosEvalCosX( S, ca );
    const double cos_tmax = ( S.cta - cacg )/sasg;
    const double tmax = ( cos_tmax <= -1.0 ? kPi : acos( dmin( 1.0, cos_tmax ) ) );
    if ( tmax <= 1e-12 )
      return 0.0;
    double intacc = S.numint_accuracy;
    if ( tmax < 10*kArcSec ) {
      intacc = dmax( intacc, 1e-6 );
      if ( tmax < kArcSec ) {
        intacc = dmax( intacc, 1e-5 );
        if ( tmax < 0.1*kArcSec )
          intacc = dmax( intacc, 1e-4 );
      }
    }
    return 2.0*sa*gosRomberg( S, sasg, cacg, intacc, 0, tmax );
  }

  // GaussOnSphere::circleIntegral, ref: NCGaussOnSphere.hh:194-208
  NCB_HD double gosCircleIntegral( const ScBraggT& S, double cg, double sg, double ca, double sa )
  {
    const double sasg = sa*sg;
    const double cacg = ca*cg;
    const double cd = cacg + sasg;
    if ( cd > S.cta && sasg >= 1e-14 && S.circleint_k2 > S.circleint_k1*sasg + cacg )
      return splineEval( S.sofcosd, cd ) * sqrt( sa/sg );
    return gosCircleIntegralSlow( S, cg, sg, ca, sa );
  }

  // GaussMos_cacheRound / SCBragg_cacheRound, ref: NCGaussMos.cc:28-36, NCSCBragg.cc:224-230
  NCB_HD double gmCacheRound( double x ) { return floor( dmax( x, 1e-15 )*1e15 + 0.5 )*1e-15; }
  NCB_HD double scCacheRound( double x ) { return floor( x*1e15 + 0.5 )*1e-15; }

  // InteractionPars for one (wavelength, family); Q is evaluated lazily like the reference
  // (calcRawCrossSectionValueInit, NCGaussMos.cc:116-145) but without cross-family caching --
  // the cached values are pure functions of (wl, inv2dsp, xsfact).
  struct InteractionPars {
    double Q, sin_perfect_theta, cos_perfect_theta, cos_perfect_theta_sq, wl3, Qprime, xsfact, inv2dsp;
    // InteractionPars::set, ref: NCGaussMos.cc:252-280
    NCB_HD void set( double wl_raw, double inv2dsp_raw, double xsfact_ )
    {
      xsfact = xsfact_ * 0.5;
      const double wl = gmCacheRound( wl_raw );
      inv2dsp = gmCacheRound( inv2dsp_raw );
      wl3 = wl*wl*wl;
      sin_perfect_theta = wl * inv2dsp;
      cos_perfect_theta_sq = 1 - sin_perfect_theta*sin_perfect_theta;
      Q = Qprime = cos_perfect_theta = -1;
    }
  };

  // calcRawCrossSectionValue (+Init), ref: NCGaussMos.hh:248-258, NCGaussMos.cc:116-145
  NCB_HD_NOINLINE double gmRawXS( const ScBraggT& S, InteractionPars& ip, double cos_angle_indir_normal )
  {
    cos_angle_indir_normal = dclamp( cos_angle_indir_normal, -1.0, 1.0 );
    if ( !( ip.Q > 0. ) ) {
      if ( ip.Qprime == -1 ) {
        ip.cos_perfect_theta = sqrt( ip.cos_perfect_theta_sq );
        const double tmp2 = ip.cos_perfect_theta*ip.sin_perfect_theta;
        if ( tmp2 > 0 )
          ip.Qprime = ip.wl3 / tmp2;
        else
          ip.Qprime = ( ip.sin_perfect_theta > 0.5 && ip.xsfact ) ? -2.0 : 0.0;
      }
      if ( ip.Qprime > 0. ) {
        ip.Q = ip.Qprime * ip.xsfact;
      } else {
        return ip.Qprime ? kInf : 0.0;
      }
    }
    const double sin_angle_indir_normal = sqrt( 1.0 - cos_angle_indir_normal*cos_angle_indir_normal );
    return ip.Q * gosCircleIntegral( S, cos_angle_indir_normal, sin_angle_indir_normal, ip.sin_perfect_theta, ip.cos_perfect_theta );
  }

  // State of one walk over the contributing (signed) normals, in the reference's order
  // (SCBragg::pimpl::updateCache NCSCBragg.cc:233-275 + GaussMos::calcCrossSections
  // NCGaussMos.cc:147-194).  The running cumulative value reproduces xs_commul:
  // xsoffset + (xssum += xs) per family.  mode 0: total + number of entries;
  // mode 1: stop at the entry picked by pickRandIdxByWeight (linear rule '>' for n<5,
  // lower_bound '>=' otherwise; falls through to the last entry = clamp n-1).
  struct ScWalkState {
    int mode;
    bool linear;
    double choice;
    double commul_last, xsoffset, xssum;
    int n;
    Vec3 chosen_n;
    double chosen_inv2d;
  };

  // Slow path for one candidate normal (inside the truncation cone for +n or -n); returns true to stop.
  NCB_HD_NOINLINE bool scCandidate( const ScBraggT& S, InteractionPars& ip, ScWalkState& W,
                                    double nx, double ny, double nz, double dot, double sdotcptsq, double ds )
  {
    const double cta = S.cta;
    const double Am = dmax( 0.0, cta - ds );
    if ( sdotcptsq > Am*Am ) {
      const double xs = gmRawXS( S, ip, dot );
      if ( xs ) {
        W.commul_last = W.xsoffset + ( W.xssum += xs );
        ++W.n;
        if ( W.mode ) {
          W.chosen_n = Vec3{ -nx, -ny, -nz }; W.chosen_inv2d = ip.inv2dsp;
          if ( W.linear ? ( W.commul_last > W.choice ) : !( W.commul_last < W.choice ) ) return true;
        }
      }
    }
    const double Ap = dmax( 0.0, cta + ds );
    if ( sdotcptsq > Ap*Ap ) {
      const double xs = gmRawXS( S, ip, -dot );
      if ( xs ) {
        W.commul_last = W.xsoffset + ( W.xssum += xs );
        ++W.n;
        if ( W.mode ) {
          W.chosen_n = Vec3{ nx, ny, nz }; W.chosen_inv2d = ip.inv2dsp;
          if ( W.linear ? ( W.commul_last > W.choice ) : !( W.commul_last < W.choice ) ) return true;
        }
      }
    }
    return false;
  }

  // Per-family scan parameters of one neutron (InteractionPars::set, NCGaussMos.cc:252-280) and the
  // combined truncation test of GaussMos::calcCrossSections (NCGaussMos.cc:160-170).
  NCB_HD bool scIsCandidate( double cta, double cptsq, double spt, double dot, double& sdotcptsq, double& ds )
  {
    sdotcptsq = ( 1.0 - dot*dot )*cptsq;
    ds = dot * spt;
    const double A0 = dmax( 0.0, cta - fabs(ds) );
    return !( sdotcptsq <= A0*A0 );
  }

  // Sequential walk (one neutron per thread / host loop): candidates are evaluated where found.
  // The device kernels for oriented materials use the warp-cooperative form in ncb_kernels_sc.cuh.
  NCB_HD void scWalk( const ScBraggT& S, double ekin_raw, const Vec3& dir_norm, double& wl_out, ScWalkState& W )
  {
    const double ekin = scCacheRound( ekin_raw );
    const double wl = ekin ? sqrt( kWl2Ekin / ekin ) : kInf;   // ekin2wl, NCDefs.hh:840-845
    wl_out = wl;
    W.commul_last = 0.0; W.n = 0;
    if ( wl == 0 )
      return;
    const double inv2dcutoff = ( 1.0 - 2*kDblEps )/wl;
    const double cta = S.cta;
    const double dx = dir_norm.x, dy = dir_norm.y, dz = dir_norm.z;
    for ( int ifam = 0; ifam < S.nfam; ++ifam ) {
      const double inv2d = S.fam_inv2d[ifam];
      if ( inv2d >= inv2dcutoff )
        break;
      InteractionPars ip;
      ip.set( wl, inv2d, S.fam_xsfact[ifam] );
      W.xsoffset = W.commul_last;
      W.xssum = 0.0;
      const double cptsq = ip.cos_perfect_theta_sq;
      const double spt = ip.sin_perfect_theta;
      const int n0 = S.fam_first[ifam], n1 = S.fam_first[ifam+1];
      const double* nrm = S.normals + 3*n0;
      for ( int in = n0; in < n1; ++in, nrm += 3 ) {
        const double nx = nrm[0], ny = nrm[1], nz = nrm[2];
        const double dot = nx*dx + ny*dy + nz*dz;
        double sdotcptsq, ds;
        if ( !scIsCandidate( cta, cptsq, spt, dot, sdotcptsq, ds ) )
          continue;
        if ( scCandidate( S, ip, W, nx, ny, nz, dot, sdotcptsq, ds ) )
          return;
      }
    }
  }

  // SCBragg::crossSection, ref: NCSCBragg.cc:295-302.  n_out = number of entries of xs_commul.
  NCB_HD double scXS( const ScBraggT& S, double ekin, const Vec3& dir, int& n_out )
  {
    n_out = 0;
    if ( ekin <= S.threshold_ekin )
      return 0.0;
    Vec3 d = dir;
    vnormalise( d );
    double wl;
    ScWalkState W; W.mode = 0; W.linear = false; W.choice = 0.0;
    scWalk( S, ekin, d, wl, W );
    n_out = W.n;
    return W.commul_last;
  }

  // randPointOnUnitCircle, ref: NCRandUtils.cc:98-112
  NCB_HD void randPointOnUnitCircle( Rng& rng, double& x, double& y )
  {
    double a, b, m2;
    do {
      a = -1.0 + rng.generate()*2.0;
      b = -1.0 + rng.generate()*2.0;
      m2 = a*a + b*b;
    } while ( !inInterval( 0.001, 1.0, m2 ) );
    const double m = 1.0/sqrt( m2 );
    x = a*m; y = b*m;
  }

  // GaussOnSphere::genPointOnCircle, ref: NCGaussOnSphere.cc:435-508
  // (RNGStream::coinflip of a non-builtin stream is generate()>0.5, NCRNG.cc:35-38)
  NCB_HD bool gosGenPointOnCircle( const ScBraggT& S, Rng& rng, double cg, double sg, double ca, double sa, double& ct, double& st )
  {
    const double sasg = sa*sg;
    const double cacg = ca*cg;
    const double cd = cacg + sasg;
    if ( cd <= S.cta )
      return false;
    if ( sasg < 1e-14 ) {
      if ( sa < 1e-7 )
        return false;
      randPointOnUnitCircle( rng, ct, st );
      return true;
    }
    const double cos_tmax = ( S.cta - cacg )/sasg;
    if ( cos_tmax >= 1.0 )
      return false;
    const double tmax = ( cos_tmax <= -1.0 ? kPi : acos( cos_tmax ) );
    const double densitymax = gosEvalCosXInRange( S, cd )*1.00000001;
    int triesleft = 1001;
    while ( --triesleft ) {
      ct = cos_mpipi( rng.generate()*tmax );
      const double cd_at_t = sasg*ct + cacg;
      const double density_at_t = gosEvalCosXInRange( S, cd_at_t );
      if ( density_at_t > densitymax * rng.generate() )
        break;
    }
    if ( triesleft <= 0 )
      return false;
    st = sqrt( 1.0 - ct*ct );
    st = ( ( rng.generate() > 0.5 ) ? st : -st );
    return true;
  }

  // PhiRot::rotateVectorAroundAxis, ref: NCRotMatrix.hh:172-191
  NCB_HD Vec3 phiRotAroundAxis( double cosphi, double sinphi, const Vec3& v, const Vec3& axis )
  {
    const Vec3 axv = vcross( axis, v );
    const double adv = vdot( axis, v );
    Vec3 r = { v.x*cosphi, v.y*cosphi, v.z*cosphi };
    const double k = sinphi * 1.0;
    r.x += axv.x*k; r.y += axv.y*k; r.z += axv.z*k;
    const double k2 = adv*( 1.0 - cosphi );
    r.x += axis.x*k2; r.y += axis.y*k2; r.z += axis.z*k2;
    return r;
  }

  // rotateToFrame, ref: NCRotMatrix.cc:84-147
  NCB_HD void rotateToFrame( double sinab, double cosab, const Vec3& a, const Vec3& b, Vec3& v, Rng& rng )
  {
    if ( fabs(sinab) < 1e-10 ) {
      const double pcos = b.z, psin = -sqrt( 1.0 - b.z*b.z );
      Vec3 axis = { b.y, -b.x, 0. };
      const double m2 = vmag2( axis );
      if ( m2 > 1e-12 ) {
        const double f = 1.0/sqrt( m2 );
        axis.x *= f; axis.y *= f; axis.z *= f;
        v = phiRotAroundAxis( pcos, psin, v, axis );
      } else {
        if ( b.z < 0.0 )
          v.z *= -1.0;
      }
      double rc, rs;
      randPointOnUnitCircle( rng, rc, rs );
      v = phiRotAroundAxis( rc, rs, v, b );
      vnormalise( v );
      return;
    }
    const double s = 1.0/sinab;
    Vec3 col1 = { b.x*(-cosab), b.y*(-cosab), b.z*(-cosab) };
    col1.x += a.x; col1.y += a.y; col1.z += a.z;
    col1.x *= s; col1.y *= s; col1.z *= s;
    Vec3 col2 = vcross( b, a );
    col2.x *= s; col2.y *= s; col2.z *= s;
    const Vec3 r = { v.x*col1.x + v.y*col2.x + v.z*b.x,
                     v.x*col1.y + v.y*col2.y + v.z*b.y,
                     v.x*col1.z + v.y*col2.z + v.z*b.z };
    v = r;
    vnormalise( v );
  }

  // GaussMos::genScat, ref: NCGaussMos.cc:196-250
  NCB_HD void gmGenScat( const ScBraggT& S, Rng& rng, const Vec3& plane_normal, double plane_inv2d, double wl_raw,
                         const Vec3& indir, Vec3& outdir )
  {
    const double wl = gmCacheRound( wl_raw );
    const double inv2d = gmCacheRound( plane_inv2d );
    const double sinthetabragg = wl * inv2d;
    if ( sinthetabragg == 0. ) {
      outdir = indir;
      return;
    }
    const double ca = sinthetabragg;
    const double sa = sqrt( 1.0 - ca*ca );
    const double cg = dclamp( -vdot( indir, plane_normal ), -1.0, 1.0 );
    const double sg = sqrt( 1.0 - cg*cg );
    double ct, st;
    if ( !gosGenPointOnCircle( S, rng, cg, sg, ca, sa, ct, st ) ) {
      outdir = indir;
      return;
    }
    const double s2a = 2*sa*ca;
    const double c2a = ca*ca - sa*sa;
    outdir = { s2a*ct, s2a*st, c2a };
    const Vec3 mindir = { -indir.x, -indir.y, -indir.z };
    rotateToFrame( sg, cg, plane_normal, mindir, outdir, rng );
    vnormalise( outdir );
  }

  // SCBragg::sampleScatter, ref: NCSCBragg.cc:304-323 (+ genScat :277-288).  `n_entries`/`total`
  // are the results of the xs pass over the same (E,dir) (the reference's cache).
  NCB_HD void scSampleScatter( const ScBraggT& S, double ekin, const Vec3& indir_raw, int n_entries, double total,
                               Rng& rng, Vec3& outdir )
  {
    outdir = indir_raw;
    if ( ekin <= S.threshold_ekin )
      return;
    if ( n_entries == 0 || total <= 0.0 )
      return;
    Vec3 d = indir_raw;
    vnormalise( d );
    // pickRandIdxByWeight over xs_commul (NCRandUtils.hh:130, .cc:198-220), without storing the list
    double wl = 0.0;
    ScWalkState W; W.mode = 1; W.chosen_n = Vec3{ 0, 0, 1 }; W.chosen_inv2d = 0.0;
    if ( n_entries == 1 ) {
      W.linear = true; W.choice = -1.0;   // first entry is taken (no draw)
    } else {
      W.choice = total * rng.generate();
      W.linear = ( n_entries < 5 );
    }
    scWalk( S, ekin, d, wl, W );
    const Vec3 chosen_n = W.chosen_n;
    const double chosen_inv2d = W.chosen_inv2d;
    gmGenScat( S, rng, chosen_n, chosen_inv2d, wl, d, outdir );
  }

  // randDirectionGivenScatterMu, ref: NCRandUtils.cc:51-96 (ScatterIsotropicMat::sampleScatter,
  // src/interfaces/NCProcImpl.cc:29-37: isotropic leaf inside an oriented composition)
  NCB_HD Vec3 randDirectionGivenScatterMu( Rng& rng, double mu, const Vec3& indir )
  {
    const double m2 = vmag2( indir );
    const double invm = ( fabs( m2 - 1.0 ) < 1e-14 ? 1.0 : 1.0/sqrt( m2 ) );
    Vec3 u = { indir.x*invm, indir.y*invm, indir.z*invm };
    Vec3 tmpdir = { 0, 0, 0 };
    double tmpdir_mag2 = 0.0;
    do {
      const double x0 = 2.0*rng.generate() - 1.0;
      const double x1 = 2.0*rng.generate() - 1.0;
      const double s = x0*x0 + x1*x1;
      if ( s < 1.0 ) {
        const double t = 2.0*sqrt( 1.0 - s );
        tmpdir = { x0*t, x1*t, 1.0 - 2.0*s };
        tmpdir = vcross( tmpdir, u );
        tmpdir_mag2 = vmag2( tmpdir );
      }
    } while ( tmpdir_mag2 < 0.001 );
    u.x *= mu; u.y *= mu; u.z *= mu;
    const double f = sqrt( ( 1 - mu*mu )/tmpdir_mag2 );
    u.x += tmpdir.x*f; u.y += tmpdir.y*f; u.z += tmpdir.z*f;
    return u;
  }

}
