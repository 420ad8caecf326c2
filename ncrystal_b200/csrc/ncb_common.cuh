// ncb_common.cuh -- shared definitions for the sm_100a hot path.
//
// All per-neutron physics below is written as NCB_HD functions so that the very
// same source is (a) inlined into the CUDA kernels of ncb_kernels.cu (the
// product) and (b) compilable by a plain host compiler for the CPU-side unit
// tests in tests/hostsim (test-only; the product never dispatches to it).
#pragma once
#include <cstdint>
#include <cmath>
#include <cfloat>

#if defined(__CUDACC__)
#  define NCB_HD __host__ __device__ __forceinline__
#  define NCB_HD_NOINLINE __host__ __device__ __noinline__
#else
#  define NCB_HD inline
#  define NCB_HD_NOINLINE inline
#endif

namespace ncb {

  // Constants, ref: ncrystal_core/include/NCrystal/core/NCDefs.hh:79-120,834-868
  constexpr double kBoltzmann       = 8.6173303e-5;      // eV/K
  constexpr double kNeutronMassAmu  = 1.00866491588;
  constexpr double kInvNeutronMassAmu = 1.0/kNeutronMassAmu;
  constexpr double kPi              = 3.1415926535897932384626433832795028841971694;
  constexpr double kPiSq            = 9.86960440108935861883449099987615113531369941;
  constexpr double kInvSqrtPi       = 0.564189583547756286948079451560772585844050629;
  constexpr double kEkin2WlSqInv    = 12.22430978582345950656; // 1/0.081804209605330899
  constexpr double kWl2Ekin         = 0.081804209605330899;
  constexpr double kInf             = HUGE_VAL;
  constexpr double kDblMin          = 2.2250738585072014e-308; // numeric_limits<double>::min()

  // Out-of-line fp64 libm: CUDA inlines the (large) double-precision implementations of
  // exp/log/erf/... at every call site; with ~40 call sites on the sampling path that made
  // the kernels >64 KB of SASS and instruction-fetch bound (ncu: stall_no_instruction).
  // One shared copy per function keeps the hot loops inside the instruction cache.
#if defined(NCB_MATHFN)
   // (a translation unit that only borrows the small helpers of this header -- ncb_vdos.cu -- asks for inline copies)
#elif defined(__CUDACC__)
#  define NCB_MATHFN __host__ __device__ __noinline__
#else
#  define NCB_MATHFN inline
#endif
  NCB_MATHFN double m_exp( double x ) { return exp(x); }
  NCB_MATHFN double m_log( double x ) { return log(x); }
  NCB_MATHFN double m_expm1( double x ) { return expm1(x); }
  NCB_MATHFN double m_log1p( double x ) { return log1p(x); }
  NCB_MATHFN double m_erf( double x ) { return erf(x); }
  NCB_MATHFN double m_erfc( double x ) { return erfc(x); }
  NCB_MATHFN double m_sqrt( double x ) { return sqrt(x); }   // (used on the free-gas path only)
  NCB_MATHFN double m_div( double a, double b ) { return a / b; }

  NCB_HD double dmin( double a, double b ) { return a < b ? a : b; }       // ncmin
  NCB_HD double dmax( double a, double b ) { return a > b ? a : b; }       // ncmax
  NCB_HD double dclamp( double v, double lo, double hi ) { return dmin( dmax( v, lo ), hi ); } // ncclamp (NCMath.hh)
  NCB_HD bool isFinite( double x ) { return fabs(x) <= DBL_MAX; } // std::isfinite (false for NaN/inf)
  NCB_HD bool inInterval( double a, double b, double x ) { return ( a <= x ) & ( x <= b ); }   // valueInInterval

  // EnergyDomain::contains, ref: NCTypes.hh:435,833-842
  NCB_HD bool domainContains( double lo, double hi, double e )
  {
    const bool isnull = ( lo > DBL_MAX ) || ( lo == hi );
    return !isnull && e >= lo && e <= hi;
  }

  // Neumaier summation, ref: NCMath.hh:526-537 (StableSum)
  struct StableSum {
    double s = 0.0, c = 0.0;
    NCB_HD void add( double x )
    {
      double t = s + x;
      c += ( fabs(s) >= fabs(x) ) ? ( (s-t) + x ) : ( (x-t) + s );
      s = t;
    }
    NCB_HD double sum() const { return s + c; }
  };

  // std::upper_bound / std::lower_bound over a sorted fp64 array, returning indices.
  template <class Ptr>
  NCB_HD int upperBound( Ptr a, int lo, int hi, double v )
  {
    // first index i in [lo,hi) with a[i] > v  (hi if none)
    while ( lo < hi ) {
      int mid = lo + ( ( hi - lo ) >> 1 );
      if ( !( v < a[mid] ) ) lo = mid + 1; else hi = mid;
    }
    return lo;
  }
  // upperBound(a,0,n,v) for a grid that is (close to) geometrically spaced: the spacing only supplies the
  // starting point, the result is verified against the grid itself, so it is the exact upper bound for any
  // ascending grid (falls back to the binary search when the guess is more than 3 entries off).
  template <class Ptr>
  NCB_HD int upperBoundLogGuess( Ptr a, int n, double v, double log_a0, double inv_dlog )
  {
    double t = ( m_log( v ) - log_a0 )*inv_dlog + 1.0;
    int i = t > 0.0 ? ( t < (double)n ? (int)t : n ) : 0;     // (NaN -> 0)
    int steps = 0;
    while ( i < n && !( v < a[i] ) ) {
      ++i;
      if ( ++steps > 3 ) return upperBound( a, i, n, v );
    }
    while ( i > 0 && v < a[i-1] ) {
      --i;
      if ( ++steps > 3 ) return upperBound( a, 0, i, v );
    }
    return i;
  }

  // Energy-key lookup table over an ascending fp64 table a[0..n): key(v) = bits(v) >> shift (exponent and the top
  // 52-shift mantissa bits: a monotone function of v for v >= 0), lut[k] = number of table entries whose key is
  // below key0 + k.  For a value with (clamped) key k, upper_bound(a, v) lies in [lut[k], lut[k+1]]: entries with
  // a smaller key are < v, entries with a larger key are > v.  The search inside that range gives the exact
  // std::upper_bound for every input (negative, inf, NaN included: they land in the first / last bucket, whose
  // range ends at 0 / n).  Replaces the 8-13 step whole-table bisections of the cross-section path by 0-2 steps.
  struct KeyLut {
    const uint16_t* lut;   // nk+1 entries, or null: plain bisection
    int key0, shift, nk;
  };
  NCB_HD long long doubleBits( double v )
  {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong( v );
#else
    long long b; __builtin_memcpy( &b, &v, sizeof(b) ); return b;
#endif
  }
  template <class Ptr>
  NCB_HD int upperBoundKeyed( Ptr a, int n, double v, const uint16_t* lut, int key0, int shift, int nk )
  {
    if ( !lut )
      return upperBound( a, 0, n, v );
    long long k = ( doubleBits( v ) >> shift ) - (long long)key0;
    k = k < 0 ? 0 : ( k > (long long)( nk-1 ) ? (long long)( nk-1 ) : k );
    return upperBound( a, (int)lut[k], (int)lut[k+1], v );
  }

  // Loads from the immutable material tables in global memory: ld.global.nc instead of the generic-address loads the
  // compiler emits for pointers it only knows from a by-value struct (ncu: LD.E + two R2UR per search step).
  template <class T>
  NCB_HD T ldTable( const T* p )
  {
#if defined(__CUDA_ARCH__)
    return __ldg( p );
#else
    return *p;
#endif
  }
  // Streaming arrays (per-neutron inputs and outputs, each touched once per launch): evict-first loads / stores so
  // that they do not displace the material tables from L2 (NCB_STREAM=0 compiles plain accesses, for A/B runs).
#if !defined(NCB_STREAM)
#  define NCB_STREAM 1
#endif
  NCB_HD double ldStream( const double* p )
  {
#if defined(__CUDA_ARCH__) && NCB_STREAM
    return __ldcs( p );
#else
    return *p;
#endif
  }
  NCB_HD void stStream( double* p, double v )
  {
#if defined(__CUDA_ARCH__) && NCB_STREAM
    __stcs( p, v );
#else
    *p = v;
#endif
  }
  // upperBound / lowerBound over a table in global memory
  NCB_HD int upperBoundTable( const double* a, int lo, int hi, double v )
  {
    while ( lo < hi ) {
      int mid = lo + ( ( hi - lo ) >> 1 );
      if ( !( v < ldTable( a + mid ) ) ) lo = mid + 1; else hi = mid;
    }
    return lo;
  }
  NCB_HD int lowerBoundTable( const double* a, int lo, int hi, double v )
  {
    while ( lo < hi ) {
      int mid = lo + ( ( hi - lo ) >> 1 );
      if ( ldTable( a + mid ) < v ) lo = mid + 1; else hi = mid;
    }
    return lo;
  }

  template <class Ptr>
  NCB_HD int lowerBound( Ptr a, int lo, int hi, double v )
  {
    // first index i in [lo,hi) with !(a[i] < v)  (hi if none)
    while ( lo < hi ) {
      int mid = lo + ( ( hi - lo ) >> 1 );
      if ( a[mid] < v ) lo = mid + 1; else hi = mid;
    }
    return lo;
  }

  // pickRandIdxByWeight, ref: NCRandUtils.cc:198-220 (n>=2; caller handles n==1 w/o draw)
  template <class Ptr>
  NCB_HD int pickIdxByWeight( double rand01, Ptr cumul, int n )
  {
    if ( n < 5 ) {
      const double choice = cumul[n-1] * rand01;
      for ( int i = 0; i < n; ++i )
        if ( cumul[i] > choice )
          return i;
      return n-1;
    }
    int i = lowerBound( cumul, 0, n, cumul[n-1] * rand01 );
    return i < n-1 ? i : n-1;
  }


  // Romberg::integrate, ref: NCRomberg.cc:62-146, over an integrand object F with the three hooks of the reference's
  // class: evalMany(fvals,n,offset,delta), evalManySum(n,offset,delta), accept(level,prev_estimate,estimate).
  // `converged` is cleared when the last level is reached without acceptance (the reference's convergenceError).
  // (also instantiated with host-only integrands by the setup code of ncb_vdos.h: the host-call-from-host-device
  //  diagnostic is silenced for this template)
#ifdef __CUDACC__
#pragma nv_diag_suppress 20011
#endif
  template <class F>
  NCB_HD double rombergIntegrate( F& f, double a, double b, bool& converged )
  {
    double h = ( b - a );
    double fvals[17];
    f.evalMany( fvals, 17, a, h*0.0625 );
    h *= 0.5;
    const double R00 = (fvals[0] + fvals[16])*h;
    const double R10 = h*fvals[8] + 0.5*R00;
    const double R11 = (4./3.)*R10 + (-1./3.)*R00;
    h *= 0.5;
    const double R20 = h*(fvals[4]+fvals[12]) + 0.5*R10;
    const double R21 = (4./3.) * R20 + (-1./3.)* R10;
    const double R22 = (16./15.) * R21 + (-1./15.) * R11;
    h *= 0.5;
    const double R30 = h*((fvals[2]+fvals[6])+(fvals[10]+fvals[14])) + 0.5*R20;
    const double R31 = (4./3.) * R30 + (-1./3.)* R20;
    const double R32 = (16./15.) * R31 + (-1./15.) * R21;
    const double R33 = (64./63.) * R32 + (-1./63.) * R22;
    h *= 0.5;
    const double R40 = h*(((fvals[1]+fvals[3])+(fvals[5]+fvals[7]))+((fvals[9]+fvals[11])+(fvals[13]+fvals[15]))) + 0.5*R30;
    const double R41 = (4./3.) * R40 + (-1./3.)* R30;
    const double R42 = (16./15.) * R41 + (-1./15.) * R31;
    const double R43 = (64./63.) * R42 + (-1./63.) * R32;
    const double R44 = (256./255.) * R43 + (-1./255.) * R33;
    if ( f.accept( 4, R33, R44 ) )
      return R44;
    const double c5 = f.evalManySum( 16, a+h*0.5, h );
    h *= 0.5;
    const double R50 = h*c5 + 0.5*R40;
    const double R51 = (4./3.) * R50 + (-1./3.)* R40;
    const double R52 = (16./15.) * R51 + (-1./15.) * R41;
    const double R53 = (64./63.) * R52 + (-1./63.) * R42;
    const double R54 = (256./255.) * R53 + (-1./255.) * R43;
    const double R55 = (1024./1023.) * R54 + (-1./1023.) * R44;
    if ( f.accept( 5, R44, R55 ) )
      return R55;
    constexpr unsigned maxlevel = 16;
    double cache1[maxlevel], cache2[maxlevel];
    double *row_prev = &cache1[0], *row = &cache2[0];
    row_prev[0] = R50; row_prev[1] = R51; row_prev[2] = R52;
    row_prev[3] = R53; row_prev[4] = R54; row_prev[5] = R55;
    unsigned nj = 16;
    for ( unsigned i = 6; i < maxlevel; ++i ) {
      const double hh = h;
      h *= 0.5;
      nj *= 2;
      const double c = f.evalManySum( nj, a+h, hh );
      row[0] = h*c + 0.5*row_prev[0];
      double n_k = 1.;
      for ( unsigned j = 0; j < i; ++j ) {
        n_k *= 4.0;
        row[j+1] = ( n_k * row[j] - row_prev[j] ) / ( n_k - 1.0 );
      }
      if ( f.accept( i, row_prev[i-1], row[i] ) )
        return row[i];
      double* t = row_prev; row_prev = row; row = t;
    }
    converged = false;
    return row_prev[maxlevel-1];
  }
#ifdef __CUDACC__
#pragma nv_diag_default 20011
#endif

}
