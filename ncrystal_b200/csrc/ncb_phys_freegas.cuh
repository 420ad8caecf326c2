// ncb_phys_freegas.cuh -- exact quantum free-gas (alpha,beta) sampling by nested
// rejection with adaptive support shrinking.  Restates, per neutron,
//   FreeGasSampler            ref: src/phys_utils/NCFreeGasUtils.cc:492-935
//   FGEvalBetaDistHelper      ref: NCFreeGasUtils.cc:153-233
//   ErfcBounds                ref: NCFreeGasUtils.cc:92-150
//   randExpMInvXMCXDivSqrtX   ref: NCFreeGasUtils.cc:237-490
//   randExpDivSqrt            ref: src/utils/NCRandUtils.cc:224-380
//   RandExpIntervalSampler    ref: include/NCrystal/internal/utils/NCRandUtils.hh:180-215
//   erfcdiff / erfc_rescaled  ref: src/utils/NCMath.cc:338-415
// Branch structure and evaluation order follow the reference so that the same
// uniforms give the same accept/reject decisions.
#pragma once
#include "ncb_phys_basic.cuh"

namespace ncb {

  // ---- erfc lookup table (1103 entries; filled on the host at library load with
  // the same libm erfc the reference uses, uploaded once per device).
  constexpr int kErfcLutLen = 1103; // = 3+(9.0-(-2.0))/0.01
#if defined(__CUDACC__)
  // defined here: the library is a single translation unit (no relocatable device code)
  __device__ double g_erfc_lut_dev[kErfcLutLen];
#endif
  extern double g_erfc_lut_host[kErfcLutLen];

  NCB_HD double erfcLut( int i )
  {
#if defined(__CUDA_ARCH__)
    return g_erfc_lut_dev[i];
#else
    return g_erfc_lut_host[i];
#endif
  }

  // initCache, ref: NCFreeGasUtils.cc:107-121 (+ linspace, src/utils/NCMath.cc:68-81)
  inline void fillErfcLutHost( double* v )
  {
    constexpr double lo = -2.0, hi = 9.0;
    constexpr int nbinedges = kErfcLutLen - 2;
    const double interval = ( hi - lo ) / ( nbinedges - 1 );
    v[0] = 2.0;
    for ( int i = 0; i < nbinedges - 1; ++i )
      v[1+i] = std::erfc( lo + i*interval );
    v[nbinedges] = std::erfc( hi );
    v[kErfcLutLen-1] = 0.0;
  }

  struct PairDD { double first, second; };

  // erfcQuickBounds, ref: NCFreeGasUtils.cc:125-149
  NCB_HD PairDD erfcQuickBounds( double x )
  {
    constexpr double lookupTableLowerEdge = -2.0;
    constexpr double lookupTableUpperEdge = 9.0;
    constexpr double lookupTableBinWidth = (lookupTableUpperEdge-lookupTableLowerEdge)/(kErfcLutLen-3);
    constexpr double lookupTableInvBinWidth = (kErfcLutLen-3)/(lookupTableUpperEdge-lookupTableLowerEdge);
    constexpr int nbins_inclinf = kErfcLutLen-1;
    constexpr double verylow = lookupTableLowerEdge-0.5*lookupTableBinWidth;
    constexpr double veryhigh = lookupTableUpperEdge+0.5*lookupTableBinWidth;
    const double x_safe = dclamp( x, verylow, veryhigh );
    const double x_relpos = ( x_safe - lookupTableLowerEdge );
    int binidx = static_cast<int>( 1.0 + x_relpos * lookupTableInvBinWidth );
    binidx = binidx < nbins_inclinf ? binidx : nbins_inclinf;
    binidx = binidx > 0 ? binidx : 0;
    PairDD bounds;
    bounds.first = erfcLut( binidx+1 ) * 0.99999999;
    bounds.second = erfcLut( binidx ) * 1.00000001;
    return bounds;
  }

  // erfcdiff_notaylor, ref: NCMath.cc:337-361
  NCB_HD double erfcdiffNoTaylor( double a, double b )
  {
    if ( b < 0 ) {
      b = -b;
      a = -a;
      const double t = a; a = b; b = t;
    }
    const double erfca = a > 27.3 ? 0.0 : m_erfc(a);
    if ( b > a+4.0 && ( a >= 4 || ( a < 0.0 && b > 6.0 ) ) )
      return erfca;
    const double erfcb = b > 27.3 ? 0.0 : m_erfc(b);
    return erfca - erfcb;
  }

  // erfcdiff, ref: NCMath.cc:363-390
  NCB_HD_NOINLINE double erfcdiff( double a, double b )
  {
    if ( dmax( fabs(a), fabs(b) ) < 0.32 ) {
      constexpr double c1  = - 2.0 * kInvSqrtPi;
      constexpr double c3  =   2.0 * kInvSqrtPi / 3.0;
      constexpr double c5  = - 0.2 * kInvSqrtPi;
      constexpr double c7  =   kInvSqrtPi / 21.0;
      constexpr double c9  = - kInvSqrtPi / 108.0;
      constexpr double c11 =   kInvSqrtPi / 660.0;
      constexpr double c13 = - kInvSqrtPi / 4680.0;
      constexpr double c15 =   kInvSqrtPi / 37800.0;
      const double a2 = a*a;
      const double b2 = b*b;
      const double a3to11 = a * a2 * ( c3 + a2 * ( c5 + a2 * ( c7 + ( a2 * ( c9 + a2 * ( c11 + a2 * ( c13 + a2 * c15 ) ) ) ) ) ) );
      const double b3to11 = b * b2 * ( c3 + b2 * ( c5 + b2 * ( c7 + ( b2 * ( c9 + b2 * ( c11 + b2 * ( c13 + b2 * c15 ) ) ) ) ) ) );
      return c1*(a-b) + ( a3to11 - b3to11 );
    }
    return a > b ? -erfcdiffNoTaylor( b, a ) : erfcdiffNoTaylor( a, b );
  }

  // erfc_rescaled, ref: NCMath.cc:393-415
  NCB_HD_NOINLINE double erfcRescaled( double x, double b )
  {
    if ( b < -745.1 )
      return 0.0;
    if ( ( x < 23.0 && fabs(b) < 700 ) || x < 5 )
      return m_exp(b) * m_erfc(x);
    const double bxx = b - x*x;
    if ( bxx < -745.1 )
      return 0.0;
    const double c3  = -0.5;
    const double c5  =  0.75;
    const double c7  = -1.875;
    const double c9  =  6.5625;
    const double c11 = -29.53125;
    const double y = 1/x;
    const double y2 = y*y;
    return kInvSqrtPi*m_exp(bxx)*(y+y2*(c3+y2*(c5+y2*(c7+y2*(c9+y2*c11)))));
  }

  // RandExpIntervalSampler, ref: NCRandUtils.hh:180-215
  struct ExpIntervalSampler {
    double a = 0, c1 = 0, c2 = 0;
    NCB_HD void set( double a_, double b_, double c_ ) { a = a_; c1 = -1.0/c_; c2 = m_expm1( -c_*(b_-a_) ); }
    NCB_HD void invalidate() { a = c1 = c2 = 0.0; }
    NCB_HD bool isValid() const { return c1 < 0.0; }
    NCB_HD double sample( Rng& rng ) const { return a + c1 * m_log( 1.0 + rng.generate() * c2 ); }
  };

  // randExpDivSqrt, ref: NCRandUtils.cc:224-380.  Sample m_exp(-c*x)/m_sqrt(x) on [a,b].
  NCB_HD_NOINLINE double randExpDivSqrt( Rng& rng, double c, double a, double b )
  {
    const double A = c*a;
    const double large_A_threshold = 0.1;
    if ( A > large_A_threshold ) {
      const double U = c*(b-a);
      const double invA = 1.0/A;
      ExpIntervalSampler expsampler;
      expsampler.set( 0, U, 1.0 );
      while ( true ) {
        const double ugen = expsampler.sample( rng );
        const double R = rng.generate();
        if ( (1.0+ugen*invA)*R*R < 1.0 )
          return dclamp( (ugen+A)/c, a, b );
      }
    } else {
      const double Ulim = 16.1180956509583;
      const double U = dmin( c*(b-a), Ulim );
      const double B = U+A;
      if ( !(B>A) )
        return a;
      const double sqrtA = m_sqrt(A);
      const double sqrtB = m_sqrt(B);
      const double sqrtB_minus_sqrtA = sqrtB - sqrtA;
      const double twosqrtA = 2*sqrtA;
      double ugen;
      while ( true ) {
        const double T = rng.generate()*sqrtB_minus_sqrtA;
        ugen = T*(T+twosqrtA);
        const double Raccept = rng.generate();
        if ( ugen < 2.0 ) {
          constexpr double c1 = -1.0;
          constexpr double c2 =  1.0/2.0;
          constexpr double c3 = -1.0/6.0;
          constexpr double c4 =  1.0/24.0;
          constexpr double c5 = -1.0/120.0;
          constexpr double c6 =  1.0/720.0;
          const double taylor6 = 1.0+ugen*(c1+ugen*(c2+ugen*(c3+ugen*(c4+ugen*(c5+ugen*c6)))));
          if ( Raccept > taylor6 )
            continue;
          if ( Raccept+0.020221 < taylor6 )
            break;
        } else {
          if ( Raccept > 0.135335283236614 )
            continue;
          if ( ugen > 4.0 && Raccept > 0.0183156388887343 )
            continue;
        }
        if ( Raccept < m_exp(-ugen) )
          break;
      }
      return dclamp( (ugen+A)/c, a, b );
    }
  }

  // f_eval lambda of randExpMInvXMCXDivSqrtX, ref: NCFreeGasUtils.cc:314-322
  NCB_HD_NOINLINE double fgFEval( double xmax, double c, double x )
  {
    const double exparg = (x-xmax)/(x*xmax) - c*(x-xmax);
    if ( exparg >= 706.0 )
      return 1.0;
    return exparg < -745.1 ? 0.0 : m_exp(exparg)*m_sqrt(xmax/x);
  }

  // randExpMInvXMCXDivSqrtX, ref: NCFreeGasUtils.cc:237-490.
  // Sample f(x)=m_exp(-1/x-c*x)/m_sqrt(x) over [xm,xp].  In pieces, so that the kernels can schedule the rejection
  // loop attempt by attempt (k_fg_alpha_prep / k_fg_alpha): xsBegin is everything before the loop (no uniforms),
  // xsAttempt is ONE pass of it; randExpMInvXMCXDivSqrtX below is their composition.
  struct XSamplerState {
    double c, xm, xp, xmax, xswitch, probability_flat, area_right;
    bool always_left, always_right, single_side;
  };

  // true: the result is x_out already (the reference's early returns); false: run the loop on st.
  NCB_HD_NOINLINE bool xsBegin( XSamplerState& st, double c, double xm, double xp, double& x_out )
  {
    if ( xp == xm ) { x_out = xm; return true; }
    const double sqrtc = m_sqrt(c);
    const double invsqrtc = 1/sqrtc;
    const double xpeak = ( c > 1e-5
                           ? ( c > 1e200 ? invsqrtc : (m_sqrt(16.0*c+1.0)-1.0)/(4.0*c) )
                           : ( 2.0-c*(8.0-c*(64.0-c*(640.0-c*7168.0))) ) );
    if ( xpeak == 0.0 ) { x_out = xm > 0.0 ? xm : dmin( kDblMin, xp ); return true; }
    const double xmax = ( xm > xpeak ? xm : dmin( xp, xpeak ) );
    if ( !(xmax > 0.0) ) { x_out = xm; return true; }
    double xlarge = dmax( 5.0/m_sqrt(c), 2*xpeak );
    double xsmall = dmin( 0.2/m_sqrt(c), 0.5*xpeak );
    if ( xp > xlarge )
      xp = dmin( xp, dmax( xm, xlarge ) + 15.0/c );
    if ( xm < xsmall ) {
      double xsm = dmin( xp, xsmall );
      xm = dmax( xm, xsm / ( 1 + 30.0*xsm ) );
    }
    constexpr double fpmin = kDblMin;
    if ( ( xm = dmax( fpmin, dmax( fpmin/xp, xm ) ) ) >= xp ) { x_out = xp; return true; }
    constexpr double fcutoff_limit = 1e-9;
    if ( xp < xpeak ) {
      while ( true ) {
        double xm_new = xp - 0.01*(xp-xm);
        double fval = fgFEval( xmax, c, xm_new );
        if ( fval >= fcutoff_limit )
          break;
        xm = xm_new;
      }
    }
    double probability_flat(-1.0);
    double xswitch(-1.0);
    double area_right(-1.0);
    if ( xm >= xlarge ) {
      probability_flat = 0.0;
      xswitch = xm;
    } else if ( c > 25 || xp <= xlarge ) {
      probability_flat = 1.0;
      xswitch = xp;
    } else {
      xswitch = xlarge;
      const double area_left = (xswitch-xm);
      const double B = c*xmax+1/xmax-1/xp;
      area_right = ( erfcRescaled( sqrtc*m_sqrt(xswitch), B )
                     - erfcRescaled( sqrtc*m_sqrt(xp), B ) ) * m_sqrt( kPi*(xmax/c) );
      probability_flat = area_left/(area_left+area_right);
    }
    // eval_probability lambda, :399-407
    bool always_left = ( probability_flat > 1.0-fcutoff_limit );
    bool always_right = ( probability_flat < fcutoff_limit );
    bool single_side = ( always_left || always_right );
    if ( !single_side && fgFEval( xmax, c, xswitch ) < fcutoff_limit*1.1 ) {
      probability_flat = 1.0;
      area_right = 0.0;
      xp = xswitch;
      always_left = true;
      single_side = false;
    }
    st.c = c; st.xm = xm; st.xp = xp; st.xmax = xmax; st.xswitch = xswitch;
    st.probability_flat = probability_flat; st.area_right = area_right;
    st.always_left = always_left; st.always_right = always_right; st.single_side = single_side;
    return false;
  }

  // ONE pass of the rejection loop (:409-489): true = accepted (x_out), false = the reference's `continue`.
  NCB_HD_NOINLINE bool xsAttempt( XSamplerState& st, Rng& rng, double& x_out )
  {
    constexpr double fcutoff_limit = 1e-9;
    const bool do_flat = ( st.single_side ? st.always_left : ( rng.generate() < st.probability_flat ) );
    if ( do_flat ) {
      double dx = st.xswitch - st.xm;
      // NB: the reference's xthr_low/xthr_up are loop-local (reset every pass,
      // :434-441), so its cheap pre-rejection never fires; both uniforms are
      // still consumed in this order.
      const double xgen = st.xm + rng.generate()*dx;
      const double Raccept = rng.generate();
      constexpr double fthreshold = 0.05;
      if ( !inInterval( st.xm, st.xswitch, xgen ) && Raccept > fthreshold )
        return false;
      const double fval = fgFEval( st.xmax, st.c, xgen );
      if ( fval < fthreshold ) {
        if ( fval < fcutoff_limit ) {
          if ( xgen < st.xmax )
            st.xm = xgen;
          else
            st.xswitch = xgen;
          dx = st.xswitch - st.xm;
          if ( !st.single_side ) {
            const double area_left = dx;
            st.probability_flat = area_left/(area_left+st.area_right);
            st.always_left = ( st.probability_flat > 1.0-fcutoff_limit );
            st.always_right = ( st.probability_flat < fcutoff_limit );
            st.single_side = ( st.always_left || st.always_right );
          }
          return false;
        }
      }
      if ( Raccept <= fval ) {
        x_out = xgen;
        return true;
      }
    } else {
      const double xgen = randExpDivSqrt( rng, st.c, st.xswitch, st.xp );
      if ( rng.generate() < m_exp( (xgen-st.xp)/(xgen*st.xp) ) ) {
        x_out = xgen;
        return true;
      }
    }
    return false;
  }

  NCB_HD_NOINLINE double randExpMInvXMCXDivSqrtX( Rng& rng, double c, double xm, double xp )
  {
    XSamplerState st;
    double x;
    if ( xsBegin( st, c, xm, xp, x ) )
      return x;
    while ( !xsAttempt( st, rng, x ) ) {}
    return x;
  }

  // FGEvalBetaDistHelper, ref: NCFreeGasUtils.cc:153-233
  struct FGBetaDist {
    double beta, normfact, expmbeta;
    double k11, k12, k21, k22;
    NCB_HD FGBetaDist( double c, double invA, double sqrtAc, double beta_, double normfact_ )
    {
      init( c, invA, sqrtAc, beta_, normfact_ );
    }
    NCB_HD_NOINLINE void init( double c, double invA, double sqrtAc, double beta_, double normfact_ )
    {
      beta = beta_; normfact = normfact_; expmbeta = -1.0;
      const double eps = beta/c;
      const double sqrt1pluseps = m_sqrt(1+eps);
      const double S = ( beta < 0.0 ? -1.0 : 1.0 );
      const double sqrtepsprime = ( eps >= 0.0 ? 1.0 : sqrt1pluseps );
      const double sqrtgammaplus = m_sqrt( 2.0+eps+2.0*sqrt1pluseps );
      const double SP = 0.5*(S+invA);
      const double SM = 0.5*(S-invA);
      const double invA_sqrtepsprime = invA*sqrtepsprime;
      const double mS_sqrtepsprime = -S*sqrtepsprime;
      const double SPsgp = sqrtgammaplus*SP;
      const double SMsgp = sqrtgammaplus*SM;
      const double j11 = -invA_sqrtepsprime + SPsgp;
      const double j12 = mS_sqrtepsprime + SPsgp;
      const double j21 = mS_sqrtepsprime + SMsgp;
      const double j22 =  invA_sqrtepsprime + SMsgp;
      k11 = sqrtAc*j11;
      k12 = sqrtAc*j12;
      k21 = sqrtAc*j21;
      k22 = sqrtAc*j22;
    }
    NCB_HD void evalExpMBeta()
    {
      if ( expmbeta < 0 )
        expmbeta = beta < -700.0 ? 0.0 : m_exp(-beta);
    }
    NCB_HD_NOINLINE double evalExact()
    {
      double t1 = erfcdiff( k11, k12 );
      evalExpMBeta();
      if ( !expmbeta )
        return normfact*t1;
      double t2 = erfcdiff( k21, k22 );
      return normfact*( t1 + t2*expmbeta );
    }
    NCB_HD_NOINLINE PairDD evalQuickBounds()
    {
      PairDD e11 = erfcQuickBounds( k11 );
      PairDD e12 = erfcQuickBounds( k12 );
      PairDD t1{ e11.first - e12.second, e11.second - e12.first };
      PairDD e21 = erfcQuickBounds( k21 );
      PairDD e22 = erfcQuickBounds( k22 );
      PairDD t2{ e21.first - e22.second, e21.second - e22.first };
      if ( t2.second > 0.0 ) {
        evalExpMBeta();
        return { normfact*( t1.first + t2.first*expmbeta ), normfact*( t1.second + t2.second*expmbeta ) };
      }
      return { normfact*t1.first, normfact*t1.second };
    }
  };

  // FreeGasSampler, ref: NCFreeGasUtils.cc:492-515 (ctor), :530-849 (sampleBeta),
  // :851-935 (sampleAlpha); NCFreeGasUtils.hh:146-176 (sampleAlphaBeta/sampleDeltaEMu)
  struct FreeGasSampler {
    double m_c, m_kT, m_sqrtAc, m_invA, m_Adiv4, m_normfact, m_c_real;

    NCB_HD FreeGasSampler() {}
    NCB_HD FreeGasSampler( double ekin, double kT, double mass_amu )
    {
      m_c = dmin( 1e14, dmax( 1e-10, ekin/kT ) );
      m_kT = kT;
      m_sqrtAc = m_sqrt( mass_amu*m_c/kNeutronMassAmu );
      const double A = kInvNeutronMassAmu * mass_amu; // AtomMass::relativeToNeutronMass, NCTypes.hh:815
      m_invA = 1.0/A;
      m_Adiv4 = 0.25*A;
      m_normfact = 0.5/m_erf( m_sqrt( m_c*m_invA ) );
      m_c_real = ekin/kT;
    }

    // setAB lambda, :682-713
    struct Overlay {
      double a, b, prob_downscat, prob_notclosetail;
      ExpIntervalSampler expsampler;
      NCB_HD_NOINLINE void setAB( double aaa, double bbb )
      {
        constexpr double Tlim = 2.0;
        constexpr double Tlim_k1 = 0.135335283236612691893999494972484403407;
        constexpr double Tlim_k2 = 274./315.;
        a = aaa;
        b = bbb;
        double area_downscat(-aaa), area_fartail, area_closetail;
        if ( bbb <= Tlim ) {
          area_fartail = 0.0;
          constexpr double c2 = -1./2.;
          constexpr double c3 =  1./6.;
          constexpr double c4 = -1./24.;
          constexpr double c5 =  1./120.;
          constexpr double c6 = -1./720.;
          constexpr double c7 =  1./5040.;
          area_closetail = b * (1.0+b*(c2+b*(c3+b*(c4+b*(c5+b*(c6+b*c7))))));
        } else {
          area_fartail = Tlim_k1 - m_exp(-bbb);
          area_closetail = Tlim_k2;
        }
        const double invareatot = 1.0 / ( area_downscat + area_closetail + area_fartail );
        prob_downscat = area_downscat * invareatot;
        prob_notclosetail = ( area_fartail + area_downscat ) * invareatot;
        expsampler.invalidate();
      }
    };

    // sampleBeta, ref: NCFreeGasUtils.cc:530-849, in three pieces so that the kernels can schedule it attempt by
    // attempt (k_fg_prep / k_sample_fg_refill); sampleBeta() below is their plain composition.
    //
    // (1) betaSupport: the part before the rejection loop, which consumes no random numbers (:530-680).
    //     kBetaLoop: rejection loop over [aa,bb];  kBetaHighE: beta = aa*u (one draw; aa = -elossmax);
    //     kBetaFixed: beta = aa.
    enum { kBetaLoop = 0, kBetaHighE = 1, kBetaFixed = 2 };
    NCB_HD_NOINLINE int betaSupport( double& aa, double& bb ) const
    {
      bb = 13.815510557964274;
      if ( m_c_real > 1e4 ) {
        const double A = 1.0 / m_invA;
        const double A2 = A*A;
        const double c_highe_threshold = 1e4*dmin( 1000.0*A, A2*A2*A2 );
        if ( m_c_real > c_highe_threshold ) {
          double am1_div_ap1 = (1.0-m_invA)/(1.0+m_invA);
          double elossmax = m_c_real * (1.0-am1_div_ap1*am1_div_ap1);
          aa = -elossmax;
          return kBetaHighE;
        }
      }
      const double fcutoff_limit = 1e-6;
      aa = dmax( -m_c_real, -m_c );
      if ( m_invA <= 1.0/10.0 ) {
        if ( m_c > 10.1 ) {
          while ( true ) {
            double aa_new = aa*0.2;
            if ( aa_new > -1e-99 )
              break;
            double fval = FGBetaDist( m_c, m_invA, m_sqrtAc, aa_new, m_normfact ).evalExact();
            if ( fval > fcutoff_limit )
              break;
            aa = aa_new;
          }
        }
        {
          while ( true ) {
            double bb_new = bb*0.25;
            if ( bb_new < 1e-99 )
              break;
            double fval_upperbound = FGBetaDist( m_c, m_invA, m_sqrtAc, bb_new, m_normfact ).evalQuickBounds().second;
            if ( fval_upperbound > fcutoff_limit )
              break;
            bb = bb_new;
          }
        }
      }
      return ( bb > aa ) ? kBetaLoop : kBetaFixed;
    }

    // (2) state of the rejection loop: the overlay and the two thresholds that adapt while it runs (:682-720)
    struct BetaState {
      Overlay ov;
      double afthreshold, bfthreshold;
    };
    NCB_HD void betaBegin( BetaState& st, double aa, double bb ) const
    {
      st.ov.setAB( aa, bb );
      st.afthreshold = st.ov.a;
      st.bfthreshold = st.ov.b;
    }

    // (3) ONE pass of the rejection loop (:722-849), itself in two halves so that the kernel can run the expensive
    //     exact evaluation for many lanes at once: betaAttemptQuick draws the candidate and decides what the cheap
    //     erfc bounds can decide; betaAttemptExact is the exact evaluation for candidates they could not decide.
    enum { kBetaReject = 0, kBetaAccept = 1, kBetaNeedExact = 2 };
    // after a candidate was not accepted (:822-848): shrink the overlay / tighten the thresholds
    NCB_HD void betaNotAccepted( BetaState& st, double beta, double fval ) const
    {
      const double fcutoff_limit = 1e-6;
      constexpr double fthreshold = 0.1;
      if ( fval < fcutoff_limit ) {
        if ( beta < 0 )
          st.ov.setAB( beta, st.ov.b );
        else
          st.ov.setAB( st.ov.a, beta );
        return;
      }
      if ( fval < fthreshold ) {
        if ( beta < 0 )
          st.afthreshold = dmax( st.afthreshold, beta );
        else
          st.bfthreshold = dmin( st.bfthreshold, beta );
      }
    }
    NCB_HD_NOINLINE int betaAttemptQuick( BetaState& st, Rng& rng, double& beta_out, double& faccept_out ) const
    {
      constexpr double Tlim = 2.0;
      constexpr double fthreshold = 0.1;
      Overlay& ov = st.ov;
      double beta, foverlay;
      const double R_selectregion = rng.generate();
      if ( R_selectregion < ov.prob_downscat ) {
        beta = rng.generate()*ov.a;
        foverlay = 1.0;
      } else {
        if ( R_selectregion < ov.prob_notclosetail ) {
          if ( !ov.expsampler.isValid() )
            ov.expsampler.set( Tlim, ov.b, 1.0 );
          beta = ov.expsampler.sample( rng );
          foverlay = m_exp(-beta);
        } else {
          const double bmax = dmin( ov.b, Tlim );
          while ( true ) {
            beta = rng.generate()*bmax;
            const double Raccept0 = rng.generate();
            constexpr double kcheap = 19./45.;
            if ( Raccept0 > 1.0 - kcheap*beta )
              continue;
            constexpr double c1 = -1.;
            constexpr double c2 = 1./2.;
            constexpr double c3 = -1./6.;
            constexpr double c4 = 1./24.;
            constexpr double c5 = -1./120.;
            constexpr double c6 = 1./720.;
            foverlay = 1.0+beta*(c1+beta*(c2+beta*(c3+beta*(c4+beta*(c5+beta*c6)))));
            if ( Raccept0 < foverlay )
              break;
          }
        }
      }
      const double faccept = rng.generate()*foverlay;
      if ( faccept > fthreshold && !inInterval( st.afthreshold, st.bfthreshold, beta ) )
        return kBetaReject;
      beta_out = beta;
      faccept_out = faccept;
      if ( beta > 0 ) {
        FGBetaDist eval_helper( m_c, m_invA, m_sqrtAc, beta, m_normfact );
        PairDD bnd = eval_helper.evalQuickBounds();
        if ( faccept <= bnd.first )
          return kBetaAccept;
        if ( faccept > bnd.second ) {
          betaNotAccepted( st, beta, bnd.second );
          return kBetaReject;
        }
      }
      return kBetaNeedExact;
    }
    NCB_HD_NOINLINE bool betaAttemptExact( BetaState& st, double beta, double faccept ) const
    {
      const double fval = FGBetaDist( m_c, m_invA, m_sqrtAc, beta, m_normfact ).evalExact();
      if ( faccept < fval )
        return true;
      betaNotAccepted( st, beta, fval );
      return false;
    }
    // true = beta accepted, false = the reference's `continue`
    NCB_HD bool betaAttempt( BetaState& st, Rng& rng, double& beta_out ) const
    {
      double faccept;
      const int q = betaAttemptQuick( st, rng, beta_out, faccept );
      if ( q != kBetaNeedExact )
        return q == kBetaAccept;
      return betaAttemptExact( st, beta_out, faccept );
    }

    // beta from the outcome of betaSupport when no rejection loop is needed
    NCB_HD double betaDirect( int kind, double aa, Rng& rng ) const
    {
      return kind == kBetaHighE ? aa*rng.generate() : aa;
    }

    NCB_HD_NOINLINE double sampleBeta( Rng& rng ) const
    {
      double aa, bb;
      const int kind = betaSupport( aa, bb );
      if ( kind != kBetaLoop )
        return betaDirect( kind, aa, rng );
      BetaState st;
      betaBegin( st, aa, bb );
      double beta;
      while ( !betaAttempt( st, rng, beta ) ) {}
      return beta;
    }

    // sampleAlpha, ref: NCFreeGasUtils.cc:851-935, in pieces (see XSamplerState): alphaBegin does everything up to the
    // rejection loop of randExpMInvXMCXDivSqrtX (the cheap cases are finished there), alphaFromX maps the loop's
    // result back; sampleAlpha() is their composition.
    enum { kAlphaDone = 0, kAlphaLoop = 1 };
    NCB_HD_NOINLINE int alphaBegin( double beta, Rng& rng, XSamplerState& st, double& alpha_out ) const
    {
      if ( m_c_real < m_c || muIsotropicAtBeta( beta, m_c ) ) {
        AlphaLimits alim = getAlphaLimits( m_c_real, beta );
        double alpha = alim.first + rng.generate()*(alim.second-alim.first);
        alpha_out = dclamp( alpha, alim.first, alim.second );
        return kAlphaDone;
      }
      beta = dmax( -m_c, beta );
      AlphaLimits alims = getAlphaLimits( m_c, beta );
      const double am = alims.first;
      const double ap = alims.second;
      if ( am == ap ) {
        alpha_out = am;
        return kAlphaDone;
      }
      const double betasq = beta*beta;
      const double t = betasq * m_Adiv4;
      const double c = 0.0625 * betasq;
      if ( dmin(t,c) < 1e-5 ) {
        const double fourA = m_Adiv4*16.0;
        const double inv4A = 1.0/fourA;
        const double xxm( am*inv4A ), xxp( ap*inv4A );
        while ( true ) {
          const double xx = randExpDivSqrt( rng, 1.0, xxm, xxp );
          const double alpha = xx * fourA;
          if ( alpha < am || alpha > ap )
            continue;
          // randExp(rng) = -m_log(rng.generate()), NCRandUtils.hh:175-178
          if ( alpha*ap*( -m_log( rng.generate() ) ) >= t*(ap-alpha) ) {
            alpha_out = alpha;
            return kAlphaDone;
          }
        }
      } else {
        const double invt( 1.0/t );
        const double xm( am*invt ), xp( ap*invt );
        double x;
        if ( xsBegin( st, c, xm, xp, x ) ) {
          alpha_out = dclamp( x*t, am, ap );
          return kAlphaDone;
        }
        return kAlphaLoop;
      }
    }
    NCB_HD static double alphaFromX( double c_clamped, double Adiv4, double beta, double x )
    {
      beta = dmax( -c_clamped, beta );
      AlphaLimits alims = getAlphaLimits( c_clamped, beta );
      const double t = ( beta*beta ) * Adiv4;
      return dclamp( x*t, alims.first, alims.second );
    }
    NCB_HD_NOINLINE double sampleAlpha( double beta, Rng& rng ) const
    {
      XSamplerState st;
      double alpha;
      if ( alphaBegin( beta, rng, st, alpha ) == kAlphaDone )
        return alpha;
      double x;
      while ( !xsAttempt( st, rng, x ) ) {}
      return alphaFromX( m_c, m_Adiv4, beta, x );
    }

    // FreeGasSampler::sampleAlphaBeta, ref: NCFreeGasUtils.hh:146-158 (alphaGivenBeta: the part after sampleBeta)
    NCB_HD double alphaGivenBeta( double beta, Rng& rng ) const
    {
      if ( beta < -m_c || muIsotropicAtBeta( beta, m_c ) ) {
        AlphaLimits alim = getAlphaLimits( m_c_real, beta );
        const double a = alim.first + rng.generate()*(alim.second-alim.first);
        return dclamp( a, alim.first, alim.second );
      }
      return sampleAlpha( beta, rng );
    }
    NCB_HD void sampleAlphaBeta( Rng& rng, double& alpha, double& beta ) const
    {
      beta = sampleBeta( rng );
      alpha = alphaGivenBeta( beta, rng );
    }

    // FreeGasSampler::sampleDeltaEMu, ref: NCFreeGasUtils.hh:160-174 (deltaEMuGivenBeta: the part after sampleBeta)
    NCB_HD void deltaEMuGivenBeta( double beta, Rng& rng, double& deltaE, double& mu, int& err ) const
    {
      if ( beta <= -m_c || muIsotropicAtBeta( beta, m_c ) ) {
        deltaE = beta*m_kT;
        mu = rng.generate()*2.0 - 1.0;
        return;
      }
      const double alpha = sampleAlpha( beta, rng );
      alphaBetaToDeltaEMu( alpha, beta, m_c*m_kT, m_kT, deltaE, mu, err );
    }
    NCB_HD void sampleDeltaEMu( Rng& rng, double& deltaE, double& mu, int& err ) const
    {
      const double beta = sampleBeta( rng );
      deltaEMuGivenBeta( beta, rng, deltaE, mu, err );
    }
  };

  // FreeGas::sampleScatterIsotropic, ref: src/freegas/NCFreeGas.cc:70-75
  NCB_HD void fgSampleScatter( const FreeGasT& T, double ekin, Rng& rng, double& ekin_out, double& mu, int& err )
  {
    FreeGasSampler s( ekin, T.kT, T.mass_amu );
    double dE;
    s.sampleDeltaEMu( rng, dE, mu, err );
    ekin_out = dmax( 0.0, ekin + dE );
  }

}
