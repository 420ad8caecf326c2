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
#if defined(__CUDACC__)
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

}
