// ncb_mmc.cuh -- device-resident transport step ("MiniMC" on the device): sources, single-volume geometries,
// forced-collision stepping with russian roulette, standard exit tallies.  SURVEY.md 8(f) next-3.
//
// Restates the physics of the reference's MiniMC standard engine for ONE volume of ONE material:
//   step            src/minimc/NCMMC_SimEngine.cc:168-455 (StdSimEngine::processBasket)
//   geometry        src/minimc/NCMMC_{Sphere,Slab,Box,Cyl}.hh, NCMMC_Utils.cc:516-584 (slab distances)
//   transmission    src/minimc/NCMMC_Utils.cc (calcProbTransm, propagateAndAttenuate, sampleRandDists)
//   sources         src/minimc/NCMMC_Source.cc:545-640 (constant), :642-800 (circular), :205-300 (energies)
//   source entry    src/minimc/NCMMC_BasketSrcFiller.hh:84-160 (propagateToVolume / missed neutrons)
//   tallies         src/minimc/NCMMC_StdTallies.cc:240-640, include/NCrystal/internal/utils/NCHists.hh:378-396,613-657
// The reference moves 4096-neutron SoA baskets between worker threads; here the whole population is a set of SoA
// arrays in HBM, every step is a handful of launches over all live neutrons, survivors are compacted into the
// second buffer with block-aggregated atomics, and cross sections / scatterings come from the same batched
// kernels that serve the C-API (launchXSAniso / launchSampleAniso).
//
// Random numbers: neutron `id` (its index in the source) owns, for transport step s, the Philox streams
// (seed, id, kMmcSidBase+2s) [flight distance, roulette] and (seed, id, kMmcSidBase+2s+1) [scattering];
// the source uses (seed, id, kMmcSidSrc).  Results therefore do not depend on batching, ordering or GPU count.
#pragma once
#include "ncb_common.cuh"
#include "ncb_rng.cuh"

namespace ncb {

  constexpr uint32_t kMmcSidSrc  = 0x4D4D0000u;
  constexpr uint32_t kMmcSidBase = 0x4D4D0010u;

  enum { GEOM_SPHERE = 1, GEOM_SLAB = 2, GEOM_BOX = 3, GEOM_CYL = 4 };
  struct MmcGeom {
    int kind, unbounded;
    double a, b, c;    // sphere: a=r | slab: c=dz | box: a=dx,b=dy,c=dz | cyl: a=r, b=dy (0: infinitely long, axis = y)
    double rsq;
  };

  // Utils::distToSlabExit, NCMMC_Utils.cc:516-537
  NCB_HD double mmcSlabExit( double x, double ux, double d )
  {
    if ( ux > 0.0 ) return ( d - x )/ux;
    if ( ux < 0.0 ) return -( d + x )/ux;
    return kInf;
  }
  // Utils::distToSlabEntry, NCMMC_Utils.cc:539-584.  <0: miss
  NCB_HD double mmcSlabEntry( double x, double ux, double d )
  {
    const double a = fabs(x) - d;
    const double x_ux = x*ux;
    if ( a <= 0.0 ) {
      if ( a ) return 0.0;
      return x_ux > 0.0 ? -1.0 : 0.0;
    }
    if ( x_ux >= 0.0 ) return -1.0;
    return a / fabs(ux);
  }

  struct CylPars { double twoA, B, C, D; };
  // Cyl::calcCylIntersectionParams, NCMMC_Cyl.hh (axis along y)
  NCB_HD CylPars mmcCylPars( double x, double z, double ux, double uz, double rsq )
  {
    CylPars p;
    p.twoA = ux*ux; p.twoA += uz*uz; p.twoA *= 2;
    p.B = x*ux; p.B += z*uz; p.B *= 2;
    p.C = x*x; p.C += z*z; p.C -= rsq;
    p.D = p.twoA*p.C; p.D *= -2.0; p.D += p.B*p.B;
    return p;
  }

  // Geometry::distToVolumeExit for a point inside (or on the surface of) the volume.
  NCB_HD double mmcDistToExit( const MmcGeom& G, double x, double y, double z, double ux, double uy, double uz )
  {
    switch ( G.kind ) {
    case GEOM_SPHERE: {
      // Sphere::distToVolumeExitImpl: sqrt(max(0,(p.u)^2-(p^2-r^2))) - p.u, clipped at 0
      double t = -x*x; t -= y*y; t -= z*z; t += G.rsq;
      double pd = x*ux; pd += y*uy; pd += z*uz;
      t += pd*pd;
      t = sqrt( fmax( 0.0, t ) );
      t -= pd;
      return dmax( 0.0, t );
    }
    case GEOM_SLAB:
      return mmcSlabExit( z, uz, G.c );
    case GEOM_BOX: {
      double t = mmcSlabExit( x, ux, G.a );
      t = dmin( t, mmcSlabExit( y, uy, G.b ) );
      return dmin( t, mmcSlabExit( z, uz, G.c ) );
    }
    default: {
      // Cyl::distToVolumeExitUnboundedImpl (+ the y slab if the cylinder is finite)
      CylPars p = mmcCylPars( x, z, ux, uz, G.rsq );
      double t = sqrt( fabs( p.D ) );
      t -= p.B;
      t = dmax( 0.0, t );
      const double C = dmin( p.C, 0.0 );
      bool done = false;
      if ( p.twoA*C == 0.0 ) {
        if ( !p.twoA ) { t = kInf; done = true; }
        else if ( p.B >= 0.0 ) { t = 0.0; done = true; }
      }
      if ( !done ) t /= p.twoA;
      if ( G.b )
        t = dmin( t, mmcSlabExit( y, uy, G.b ) );
      return t;
    }
    }
  }

  // Geometry::distToVolumeEntry: distance to the volume for a source particle, 0 if already inside, <0 if it misses.
  NCB_HD double mmcDistToEntry( const MmcGeom& G, double x, double y, double z, double ux, double uy, double uz )
  {
    switch ( G.kind ) {
    case GEOM_SPHERE: {
      double pdotu = x*ux; pdotu += y*uy; pdotu += z*uz;
      double psq = x*x; psq += y*y; psq += z*z; psq -= G.rsq;
      if ( psq <= 0.0 )
        return dmin( psq, pdotu ) < 0.0 ? 0.0 : -1.0;
      const double D = pdotu*pdotu - psq;
      if ( D < 0 ) return -1.0;
      const double t = -( sqrt(D) + pdotu );
      return t >= 0.0 ? t : -1.0;
    }
    case GEOM_SLAB:
      return mmcSlabEntry( z, uz, G.c );
    case GEOM_BOX: {
      // Box::distToVolumeEntryImpl: intersect the three slab intervals
      double t1, t2;
      bool miss = false;
      const double pos[3] = { x, y, z }, dir[3] = { ux, uy, uz }, half[3] = { G.a, G.b, G.c };
      t1 = -kInf; t2 = kInf;
      for ( int k = 0; k < 3 && !miss; ++k ) {
        if ( dir[k] ) {
          const double inv = 1.0/dir[k];
          const double q1 = ( half[k] - pos[k] )*inv, q2 = ( -half[k] - pos[k] )*inv;
          t1 = dmax( t1, dmin( q1, q2 ) );
          t2 = dmin( t2, dmax( q1, q2 ) );
        } else if ( fabs( pos[k] ) > half[k] ) {
          miss = true;
        }
      }
      if ( miss || t1 >= t2 || t2 <= 0.0 ) return -1.0;
      return dmax( 0.0, t1 );
    }
    default: {
      CylPars p = mmcCylPars( x, z, ux, uz, G.rsq );
      if ( !G.b ) {
        // Cyl::distToVolumeEntryUnboundedImpl
        if ( p.twoA == 0.0 ) return p.C <= 0.0 ? 0.0 : -1.0;
        if ( p.C <= 0.0 ) return ( p.C == 0.0 && p.B >= 0.0 ) ? -1.0 : 0.0;
        if ( p.D <= 0 ) return -1.0;
        const double sqrtD = sqrt( p.D );
        const double tmin = -p.B - sqrtD, tmax = -p.B + sqrtD;
        if ( tmax < 0.0 ) return -1.0;
        if ( tmin > 0.0 ) return tmin / p.twoA;
        return p.B < 0.0 ? 0.0 : -1.0;
      }
      // Cyl::distToVolumeEntryBoundedImpl
      double tmin, tmax;
      if ( p.twoA == 0.0 ) {
        if ( p.C <= 0 ) { tmin = -kInf; tmax = kInf; }
        else return -1.0;
      } else {
        if ( p.D <= 0 ) return -1.0;
        const double sqrtD = sqrt( p.D ), inv = 1.0/p.twoA;
        tmin = dmax( 0.0, ( -p.B - sqrtD )*inv );
        tmax = dmax( 0.0, ( -p.B + sqrtD )*inv );
      }
      if ( fabs(y) == G.b && y*uy > 0.0 ) return -1.0;
      if ( uy == 0.0 ) {
        if ( fabs(y) > G.b ) return -1.0;
      } else {
        const double t1 = -( y + G.b )/uy, t2 = ( G.b - y )/uy;
        tmin = dmax( 0.0, dmax( dmin( t1, t2 ), tmin ) );
        tmax = dmax( 0.0, dmin( dmax( t1, t2 ), tmax ) );
      }
      if ( !( tmax > tmin ) ) return -1.0;
      return tmin;
    }
    }
  }

  // ------------------------------------------------------------------ sources
  enum { SRC_CONSTANT = 1, SRC_CIRCULAR = 2, SRC_ISOTROPIC = 3 };
  enum { SRCE_FIXED = 0, SRCE_UNIFORM_EKIN = 1, SRCE_UNIFORM_WL = 2, SRCE_LOGNORMAL_EKIN = 3, SRCE_LOGNORMAL_WL = 4,
         SRCE_MAXWELL = 5 };
  struct MmcSource {
    int kind, emode;
    double pos[3], dir[3];   // dir normalised
    double va[3], vb[3];     // circular: radius-scaled basis of the disk (NCMMC_Source.cc:664-680)
    double w, e0, e1;        // weight; fixed energy [eV] | range (eV or Aa) | log-normal (mu_n, sigma_n) | Maxwell (kT/2, -)
    double minus_r;          // isotropic: start at pos + dir*minus_r (NCMMC_Source.cc:433-543)
    int may_be_outside;
  };

  // randPointInUnitCircle, ref: src/utils/NCRandUtils.cc (rejection in the unit square)
  NCB_HD void mmcRandPointInUnitCircle( Rng& rng, double& a, double& b )
  {
    while ( true ) {
      a = 2.0*rng.generate() - 1.0;
      b = 2.0*rng.generate() - 1.0;
      if ( a*a + b*b <= 1.0 ) return;
    }
  }

  struct MmcNeutron { double x, y, z, ux, uy, uz, w, ekin; };

  // randNorm (ratio method), ref: src/utils/NCRandUtils.cc:113-137
  NCB_HD double mmcRandNorm( Rng& rng )
  {
    double g, g2, u, v, invu;
    while ( true ) {
      u = rng.generate();
      invu = 1.0/u;
      v = rng.generate();
      g = 1.71552776992141354 * ( v - 0.5 )*invu;
      g2 = g*g;
      if ( g2 <= 5.0 - 5.13610166675096558 * u )
        break;
      // (the reference's quick-reject test `if ( g2 >= 1.0369.../u ) continue;` sits in a do-while, where
      //  `continue` jumps to the loop condition below -- it never changes the outcome, so it is not restated)
      if ( !( g2 >= -4.0 * m_log( u ) ) )
        break;
    }
    return g;
  }

  NCB_HD MmcNeutron mmcGenerate( const MmcSource& S, Rng& rng )
  {
    MmcNeutron n;
    if ( S.kind == SRC_CIRCULAR && ( S.va[0] != 0.0 || S.va[1] != 0.0 || S.va[2] != 0.0 ) ) {
      double a, b;
      mmcRandPointInUnitCircle( rng, a, b );
      n.x = S.pos[0] + S.va[0]*a + S.vb[0]*b;
      n.y = S.pos[1] + S.va[1]*a + S.vb[1]*b;
      n.z = S.pos[2] + S.va[2]*a + S.vb[2]*b;
    } else {
      n.x = S.pos[0]; n.y = S.pos[1]; n.z = S.pos[2];
    }
    n.ux = S.dir[0]; n.uy = S.dir[1]; n.uz = S.dir[2];
    if ( S.kind == SRC_ISOTROPIC ) {
      // randIsotropicDirection (Marsaglia 1972), ref: src/utils/NCRandUtils.cc:25-49
      double x0, x1, s;
      do {
        x0 = 2.0*rng.generate() - 1.0;
        x1 = 2.0*rng.generate() - 1.0;
        s = x0*x0 + x1*x1;
      } while ( s >= 1.0 );
      const double t = 2.0*sqrt( 1.0 - s );
      n.ux = x0*t; n.uy = x1*t; n.uz = 1.0 - 2.0*s;
    }
    n.w = S.w;
    if ( S.emode == SRCE_FIXED ) {
      n.ekin = S.e0;
    } else if ( S.emode == SRCE_MAXWELL ) {
      // setEnergy_Maxwell, NCMMC_Source.cc:76-101: (G1^2+G2^2+G3^2)*kT/2
      double v = mmcRandNorm( rng );
      v *= v;
      const double g2 = mmcRandNorm( rng ); v += g2*g2;
      const double g3 = mmcRandNorm( rng ); v += g3*g3;
      n.ekin = v*S.e0;
    } else if ( S.emode == SRCE_LOGNORMAL_EKIN || S.emode == SRCE_LOGNORMAL_WL ) {
      // setEnergy_LogNormal, NCMMC_Source.cc:264-283
      double v = mmcRandNorm( rng );
      v *= S.e1; v += S.e0;
      v = m_exp( v );
      if ( S.emode == SRCE_LOGNORMAL_WL ) {
        v *= v;
        v = 1.0/dmax( 4.9406564584124654e-324, v );
        v *= kWl2Ekin;
      }
      n.ekin = v;
    } else {
      // setEnergy_UniformRange, NCMMC_Source.cc:214-240 (+ convertBufWl2E :104-126)
      double v = rng.generate()*( S.e1 - S.e0 );
      v += S.e0;
      v = dmin( S.e1, v );
      if ( S.emode == SRCE_UNIFORM_WL ) {
        v *= v;
        v = 1.0/dmax( 4.9406564584124654e-324, v );
        v *= kWl2Ekin;
      }
      n.ekin = v;
    }
    if ( S.kind == SRC_ISOTROPIC && S.minus_r != 0.0 ) {
      n.x += n.ux*S.minus_r; n.y += n.uy*S.minus_r; n.z += n.uz*S.minus_r;
    }
    return n;
  }

  // ------------------------------------------------------------------ the step
  struct MmcEngine {
    double macro_factor;          // Utils::macroXSFactor = 100*numdens: barn -> 1/m
    double abs_c;                 // AbsOOV constant, 0 = no absorption
    double roulette_psurv, roulette_wthr, roulette_boost;
    int roulette_nscat;
    int nscatlimit;               // -1: none
  };

  // Utils::calcProbTransm, NCMMC_Utils.cc
  NCB_HD double mmcProbTransm( bool has_xs, double xs, double dist, bool unbounded )
  {
    if ( !has_xs )
      return unbounded ? ( isFinite(dist) ? 1.0 : 0.0 ) : 1.0;
    double t = dmin( DBL_MAX, xs );
    if ( unbounded ) t *= ( t ? dist : 0.0 );
    else t *= dist;
    t = -t;
    t = m_exp( t );
    if ( unbounded ) t *= ( isFinite(dist) ? 1.0 : 0.0 );
    return t;
  }

  struct MmcStepOut {
    double wt;          // weight leaving the volume unscattered (tallied)
    bool survives;      // continues to a scattering at (x,y,z) with weight w
    double x, y, z, w;
  };

  // One neutron through StdSimEngine::processBasket up to (not including) the scattering itself.
  //   xs_scat_micro : scattering cross section at (ekin, dir) [barn]
  NCB_HD MmcStepOut mmcForward( const MmcGeom& G, const MmcEngine& E, Rng& rng,
                                double x, double y, double z, double ux, double uy, double uz,
                                double w, double ekin, int nscat, double xs_scat_micro )
  {
    MmcStepOut o;
    const bool unb = G.unbounded != 0;
    const double d_exit = mmcDistToExit( G, x, y, z, ux, uy, uz );
    const bool has_abs = E.abs_c > 0.0;
    double xs_a = 0.0;
    if ( has_abs ) {
      const double sqrtE = sqrt( ekin );
      xs_a = ( sqrtE ? E.abs_c/sqrtE : kInf )*E.macro_factor;     // AbsOOV::crossSectionIsotropic, NCAbsOOV.cc:41-45
    }
    double xs_s = E.macro_factor * xs_scat_micro;
    if ( E.nscatlimit >= 0 && nscat >= E.nscatlimit )
      xs_s = 0.0;
    const double ptransm = mmcProbTransm( true, xs_s, d_exit, unb );
    // sampleRandDists: one uniform per neutron, consumed whatever happens
    const double u = rng.generate();
    double d_scat;
    if ( !xs_s ) d_scat = kInf;
    else if ( !isFinite( d_exit ) ) d_scat = m_log( u )/( -xs_s );
    else {
      // RandExpIntervalSampler(0,d_exit,xs_s).sample(u), NCRandUtils.hh:180-215
      const double c1 = -1.0/xs_s;
      const double c2 = m_expm1( -xs_s*( d_exit - 0.0 ) );
      d_scat = 0.0 + c1*m_log( 1.0 + u*c2 );
    }
    // ---- scattered part
    o.survives = false;
    o.x = x; o.y = y; o.z = z; o.w = w;
    if ( w != 0.0 && isFinite( d_scat ) && xs_s > 0.0 ) {
      double rfact = 1.0;
      bool alive = true;
      if ( nscat >= E.roulette_nscat && w < E.roulette_wthr ) {
        if ( rng.generate() > E.roulette_psurv ) alive = false;
        else rfact = E.roulette_boost;
      }
      if ( alive ) {
        const double wred = m_exp( -xs_a*d_scat );
        if ( wred > 0.0 ) {
          double wn = w*rfact;
          o.x = x + d_scat*ux; o.y = y + d_scat*uy; o.z = z + d_scat*uz;
          wn *= wred;
          wn *= ( 1.0 - ptransm );
          o.w = wn;
          o.survives = true;
        }
      }
    }
    // ---- transmitted part (propagateAndAttenuate with the absorption xs, then the transmission probability)
    double wt = w;
    wt *= mmcProbTransm( has_abs, xs_a, d_exit, unb );
    wt *= ptransm;
    o.wt = wt;
    return o;
  }

  // ------------------------------------------------------------------ tallies
  enum { TALLY_THETA = 0, TALLY_MU = 1, TALLY_NSCAT = 2, TALLY_NSCAT_UW = 3, TALLY_W = 4, TALLY_E = 5, TALLY_L = 6,
         TALLY_DE = 7, TALLY_Q = 8, TALLY_NTYPES = 9 };
  constexpr int kMmcNClass = 5;      // NOSCAT, SINGLESCAT_ELAS, SINGLESCAT_INELAS, MULTISCAT_PUREELAS, MULTISCAT_OTHER
  constexpr int kMmcNStat = 5;       // sumw, sumwx, sumwx2, min, max   (per class)
  struct MmcHist { int type, nbins; double xmin, xmax, invdelta; uint32_t off; };   // off: doubles into the tally buffer
  struct MmcTally {
    int nh;
    MmcHist h[TALLY_NTYPES];
    double dir0[3];
    int dir0_is_z;
    int has_dir0_fixed;       // 0: per-neutron initial directions (isotropic source)
    double e0_fixed;
    int has_e0_fixed;
  };
  // storage of one histogram: [class][0..nbins+1] contents, then same for sum of squared weights, then
  // [class][kMmcNStat] running statistics.  (Totals = sums over classes, formed on the host.)
  NCB_HD uint32_t mmcHistDoubles( int nbins ) { return (uint32_t)( kMmcNClass*( 2*( nbins + 2 ) + kMmcNStat ) ); }

  // HistBinData1D::valueToBin with under/overflow bins, NCHists.hh:388-396
  NCB_HD int mmcValueToBin( const MmcHist& h, double val )
  {
    if ( val < h.xmin ) return 0;
    if ( val >= h.xmax ) return val == h.xmax ? h.nbins : h.nbins + 1;
    const unsigned long long k = (unsigned long long)( h.invdelta*( val - h.xmin ) );
    return 1 + (int)( k < (unsigned long long)h.nbins ? k : (unsigned long long)h.nbins );
  }

  // init_dethistidvect, NCMMC_StdTallies.cc
  NCB_HD int mmcClass( int nscat, int nscat_inelas )
  {
    if ( nscat > 1 ) return nscat_inelas ? 4 : 3;
    if ( nscat == 1 ) return nscat_inelas ? 2 : 1;
    return 0;
  }

  // value of tally `type` for an exiting neutron; `weighted` tells whether the fill carries the neutron weight
  NCB_HD double mmcTallyValue( const MmcTally& T, int type, double ux, double uy, double uz, double ekin, double w,
                               int nscat, double e_initial, double ux0, double uy0, double uz0, bool& weighted )
  {
    constexpr double kToDeg = 57.2957795130823208767981548141051703324054725;
    weighted = true;
    switch ( type ) {
    case TALLY_THETA: case TALLY_MU: case TALLY_Q: {
      double mu;
      if ( !T.has_dir0_fixed ) { mu = ux0*ux; mu += uy0*uy; mu += uz0*uz; }
      else if ( T.dir0_is_z ) mu = uz;
      else { mu = T.dir0[0]*ux; mu += T.dir0[1]*uy; mu += T.dir0[2]*uz; }
      mu = dclamp( mu, -1.0, 1.0 );
      if ( type == TALLY_MU ) return mu;
      if ( type == TALLY_THETA ) return acos( mu )*kToDeg;
      const double Ei = T.has_e0_fixed ? T.e0_fixed : e_initial;
      double q = ekin*Ei;
      q = sqrt( q ); q *= mu; q *= -2.0; q += Ei; q += ekin;
      q *= 39.4784176043574344753379639995046045412547976*kEkin2WlSqInv;     // ekin2ksq(1) = k4PiSq*ekin2wlsqinv(1)
      return sqrt( fmax( 0.0, q ) );
    }
    case TALLY_NSCAT: return (double)nscat;
    case TALLY_NSCAT_UW: weighted = false; return (double)nscat;
    case TALLY_W: weighted = false; return w;
    case TALLY_E: return ekin;
    case TALLY_L: {
      double t = 1.0/dmax( 4.9406564584124654e-324, ekin );
      t = sqrt( t );
      return t*0.2860143520967626;    // constexpr_ekin2wl(1.0) = sqrt(0.081804209605330899)
    }
    default: { // TALLY_DE
      const double Ei = T.has_e0_fixed ? T.e0_fixed : e_initial;
      return Ei - ekin;
    }
    }
  }

}
