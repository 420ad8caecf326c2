// ncb_vdos.h -- phonon density of states (VDOS) -> S(alpha,beta) scattering kernel, Sjolander's expansion
// (SURVEY §8f next-4).  Host orchestration; the O(N log N) and O(N^2) arithmetic -- the FFT convolutions that
// produce the phonon-order spectra G_n and the sum over orders that fills the table -- runs on a backend:
// the CUDA kernels of ncb_vdos_dev.cuh in the product (ncb_lib_vdos.inc), plain loops over the same NCB_HD
// functions in the CPU test build (tests/hostsim).
//
// Restates, with the reference's order of floating-point operations (the expansion reproduces the reference's
// tables bit for bit; every transcendental set-up value -- G_1, exp/log factors of the alpha dependence -- is
// computed here on the host with the same libm, the device does only +,-,*,/ and sqrt):
//   VDOSEval (normalisation, gamma0, mean-squared displacement, asymmetric G_1)   ref: src/vdos/NCVDOSEval.cc:139-446
//   regulariseVDOSGrid / checkIsRegularVDOSGrid                                    ref: src/vdos/NCVDOSEval.cc:466-719
//   VDOSGn (G_1 on its grid, G_n = G_{n-n/2} (x) G_{n/2}, truncation + thinning)   ref: src/vdos/NCVDOSGn.cc:62-105,135-210,259-289,372-492
//   createScatteringKernel, setupAlphaGrid, setupBetaGrid, fillSABFromVDOS[Concurrent],
//   rangeXNexpMX, findExtremeSABPointWithinAlphaPlusCurve, sabPointWithinAlphaPlusCurve
//                                                                                  ref: src/vdos/NCVDOSToScatKnl.cc:39-944
//   trimming of all-zero table edges (transformKernelToStdFormat)                  ref: src/sab/NCSABUtils.cc:30-146
//   reducePtsInDistribution, findRoot, linspace/geomspace, Romberg17/33            ref: src/utils/NCMath.cc:44-81,291-336,419-512
//                                                                                       include/NCrystal/internal/utils/NCMath.hh:580-603
// The reference grows the order one at a time (up to four concurrently in worker threads); here every order that
// the existing ones allow (n+1 .. 2n) goes into ONE batch of device work, and the reference's stopping rule is
// applied to the statistics that come back.
#pragma once
#include "ncb_common.cuh"   // rombergIntegrate, StableSum
#include "ncb_vdos_dev.cuh"
#include "ncb_blob.h"
#include <cstring>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <list>
#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#ifdef NCB_VDOS_TIMING   // (development aid: host-side stage times of expand() on stderr)
#include <chrono>
#include <cstdio>
#define NCB_VDOS_TICK(label) do { auto t_now = std::chrono::steady_clock::now(); std::fprintf( stderr, "vdos-timing %-28s %8.3f ms\n", label, std::chrono::duration<double,std::milli>( t_now - t_last ).count() ); t_last = t_now; } while (0)
#define NCB_VDOS_TICK_INIT auto t_last = std::chrono::steady_clock::now()
#else
#define NCB_VDOS_TICK(label) do {} while (0)
#define NCB_VDOS_TICK_INIT do {} while (0)
#endif

namespace ncb { namespace vdos {

  // physics constants, spelled like the reference so that the compile-time products round identically
  // (NCDefs.hh:79-80,120; NCMath.hh:34-48)
  constexpr double kBoltzmann = 8.6173303e-5;
  constexpr double kNeutronMassAmu = 1.00866491588;
  constexpr double kSpeedOfLight = 299792458e10;
  constexpr double kDalton2eVc2 = 931494095.17;
  constexpr double kPlanck = 4.135667662e-15;
  constexpr double kInv2Pi = 0.159154943091895335768883763372514362034459646;
  constexpr double kInvSqrt2Pi = 0.398942280401432677939946059934381868475858631;
  constexpr double kInvE = 0.367879441171442321595523770161460867445811131;
  constexpr double kNeutronMassEvc2 = kNeutronMassAmu * kDalton2eVc2 / ( kSpeedOfLight*kSpeedOfLight );
  constexpr double kHbar = kPlanck*kInv2Pi;

  struct Error : std::runtime_error {
    std::string type;
    Error( const char* t, const std::string& m ) : std::runtime_error( m ), type( t ) {}
  };

  typedef std::vector<double> VectD;
  typedef std::pair<double,double> PairDD;

  inline bool floatEq( double a, double b, double rtol = 1.0e-6, double atol = 1.0e-6 )
  {
    if ( std::isinf( a ) || std::isinf( b ) ) return a == b;
    return std::fabs( a - b ) <= ( 0.5*rtol )*( std::fabs( a ) + std::fabs( b ) ) + atol;
  }
  inline VectD linSpace( double start, double stop, unsigned num )
  {
    VectD v; v.reserve( num );
    const unsigned nm1 = num - 1;
    const double interval = ( stop - start )/nm1;
    for ( unsigned i = 0; i < nm1; ++i ) v.push_back( start + i*interval );
    v.push_back( stop );
    return v;
  }
  inline VectD geomSpace( double start, double stop, unsigned num )
  {
    VectD v( num );
    double s = std::log10( start );
    const double e = std::log10( stop ), interval = ( e - s )/( num - 1 );
    for ( double& x : v ) { x = std::pow( 10.0, s ); s += interval; }
    v.back() = std::pow( 10.0, e );
    v.front() = start; v.back() = stop;
    return v;
  }
  inline bool isGrid( const VectD& v )
  {
    if ( v.empty() || !( std::fabs( v[0] ) <= std::numeric_limits<double>::max() ) ) return false;
    for ( size_t i = 1; i < v.size(); ++i )
      if ( !( v[i] > v[i-1] ) || !( std::fabs( v[i] ) <= std::numeric_limits<double>::max() ) ) return false;
    return true;
  }

  // Romberg integration with a fixed number of levels: 17 points (accepted at level 4) or 33 (level 5)
  template <class Fn, unsigned kAcceptLevel>
  struct FixedRomberg {
    Fn fn;
    void evalMany( double* fvals, unsigned n, double offset, double delta ) const
    {
      const double nn = n;
      for ( double i = 0; i < nn; ++i ) *( fvals++ ) = fn( offset + delta*i );
    }
    double evalManySum( unsigned n, double offset, double delta ) const
    {
      double sum = 0.0;
      const double nn = n;
      for ( double i = 0; i < nn; ++i ) sum += fn( offset + delta*i );
      return sum;
    }
    bool accept( unsigned level, double, double ) const { return level >= kAcceptLevel; }
  };
  template <class Fn> inline double romberg17( Fn&& fn, double a, double b )
  {
    FixedRomberg<Fn&,4> r{ fn }; bool ok = true; return rombergIntegrate( r, a, b, ok );
  }
  template <class Fn> inline double romberg33( Fn&& fn, double a, double b )
  {
    FixedRomberg<Fn&,5> r{ fn }; bool ok = true; return rombergIntegrate( r, a, b, ok );
  }

  // bisection / false-position mix, ref: NCMath.cc:291-336
  template <class Fn> inline double findRoot( Fn&& f, double a, double b, double acc )
  {
    double fa = f( a ), fb = f( b );
    if ( !( b > a ) || !fa*fb < 0.0 )     // (sic: the reference's test, operator precedence included)
      throw Error( "CalcError", "root finding requires b>a and f(a)*f(b)<0." );
    acc *= 0.5;
    unsigned i = 60;
    while ( --i ) {
      const double dfba = fb - fa;
      double c = ( a*fb - b*fa )/dfba;
      if ( b - a < acc ) return c;
      const double k = 0.15*( b - a );
      c = dmax( a + k, dmin( b - k, c ) );
      const double fc = f( c );
      if ( !fc ) return c;
      if ( fa*fc < 0 ) { b = c; fb = fc; } else { a = c; fa = fc; }
    }
    throw Error( "CalcError", "Root search failed to converge!" );
  }

  // Thin a sampled curve to targetN points, removing the least important interior point first (importance =
  // area change x (change of the log-curve area)^2).  Among equal scores the one scored first goes first -- the
  // order a std::multimap gives the reference; here the ranking is a binary heap of (score, sequence number) keys
  // over an index-linked list, stale entries skipped when they surface.
  inline std::pair<VectD,VectD> reducePoints( const VectD& x, const VectD& y, size_t targetN )
  {
    const size_t n = x.size();
    if ( targetN >= n ) return { x, y };
    const double inv_ymax = 1.0 / *std::max_element( y.begin(), y.end() );
    VectD lny( n );
    for ( size_t i = 0; i < n; ++i ) lny[i] = std::log( std::max<double>( 1e-20, y[i]*inv_ymax ) );
    std::vector<uint32_t> prev( n ), next( n );
    for ( size_t i = 0; i < n; ++i ) { prev[i] = (uint32_t)( i - 1 ); next[i] = (uint32_t)( i + 1 ); }
    // indexed binary min-heap over the interior points, ordered by ( score, sequence number of the scoring ): a
    // point whose neighbour was removed is re-scored IN PLACE (sift up or down), so the heap never holds stale entries
    std::vector<double> sc( n, 0.0 );
    std::vector<uint64_t> sq( n, 0 );
    std::vector<uint32_t> heap, pos( n, 0 );
    heap.reserve( n );
    uint64_t seq = 0;
    auto score = [&]( uint32_t i ) {
      const uint32_t i0 = prev[i], i2 = next[i];
      const double area = std::fabs( x[i0]*( y[i] - y[i2] ) + x[i]*( y[i2] - y[i0] ) + x[i2]*( y[i0] - y[i] ) );
      const double larea = std::fabs( x[i0]*( lny[i] - lny[i2] ) + x[i]*( lny[i2] - lny[i0] ) + x[i2]*( lny[i0] - lny[i] ) );
      return area*larea*larea;
    };
    auto before = [&]( uint32_t a, uint32_t b ) { return sc[a] < sc[b] || ( sc[a] == sc[b] && sq[a] < sq[b] ); };
    auto siftUp = [&]( size_t k ) {
      const uint32_t v = heap[k];
      while ( k > 0 ) {
        const size_t parent = ( k - 1 )/2;
        if ( !before( v, heap[parent] ) ) break;
        heap[k] = heap[parent]; pos[heap[k]] = (uint32_t)k;
        k = parent;
      }
      heap[k] = v; pos[v] = (uint32_t)k;
    };
    auto siftDown = [&]( size_t k ) {
      const uint32_t v = heap[k];
      const size_t m = heap.size();
      while ( true ) {
        size_t c = 2*k + 1;
        if ( c >= m ) break;
        if ( c + 1 < m && before( heap[c+1], heap[c] ) ) ++c;
        if ( !before( heap[c], v ) ) break;
        heap[k] = heap[c]; pos[heap[k]] = (uint32_t)k;
        k = c;
      }
      heap[k] = v; pos[v] = (uint32_t)k;
    };
    for ( uint32_t i = 1; i + 1 < n; ++i ) {              // (end points are never candidates)
      sc[i] = score( i ); sq[i] = seq++;
      heap.push_back( i );
      siftUp( heap.size() - 1 );
    }
    auto rescore = [&]( uint32_t i ) {
      if ( i == 0 || i == n - 1 ) return;
      sc[i] = score( i ); sq[i] = seq++;
      const size_t k = pos[i];
      siftUp( k );
      if ( pos[i] == k ) siftDown( k );
    };
    size_t left = n;
    while ( left > targetN ) {
      const uint32_t idx = heap.front();
      heap.front() = heap.back(); pos[heap.front()] = 0;
      heap.pop_back();
      if ( !heap.empty() ) siftDown( 0 );
      const uint32_t pb = prev[idx], pa = next[idx];
      next[pb] = pa; prev[pa] = pb;
      --left;
      rescore( pb );
      rescore( pa );
    }
    VectD nx, ny;
    nx.reserve( left ); ny.reserve( left );
    for ( uint32_t i = 0; i < n; i = next[i] ) { nx.push_back( x[i] ); ny.push_back( y[i] ); }
    return { nx, ny };
  }

  // exp(i 2 pi k / 2^n), ref: NCFastConvolve.cc:466-564 (the reference's cached 21 high-precision values; products of them)
  inline PairDD calcPhase( unsigned k, unsigned n )
  {
    static const double cosvals[21] = { 1.0, -1.0, 0.0, 0.707106781186547524401, 0.923879532511286756128, 0.980785280403230449126,
      0.995184726672196886245, 0.998795456205172392715, 0.999698818696204220116, 0.999924701839144540922, 0.999981175282601142657,
      0.999995293809576171512, 0.999998823451701909929, 0.99999970586288221916, 0.999999926465717851145, 0.999999981616429293808,
      0.999999995404107312891, 0.999999998851026827563, 0.999999999712756706849, 0.99999999992818917671, 0.999999999982047294177 };
    static const double sinvals[21] = { 0.0, 0.0, 1.0, 0.707106781186547524401, 0.382683432365089771728, 0.195090322016128267848,
      0.0980171403295606019942, 0.049067674327418014255, 0.0245412285229122880317, 0.0122715382857199260794, 0.00613588464915447535964,
      0.00306795676296597627015, 0.0015339801862847656123, 0.000766990318742704526939, 0.000383495187571395589072, 0.00019174759731070330744,
      0.0000958737990959773458705, 0.000047936899603066884549, 0.0000239684498084182187292, 0.0000119842249050697064215, 0.00000599211245264242784288 };
    if ( k == 0 ) return PairDD( 1.0, 0.0 );
    while ( k % 2 == 0 ) { n -= 1; k /= 2; }
    if ( k == 1 ) {
      if ( n >= 21 ) throw Error( "CalcError", "VDOS spectrum too long for the FFT twiddle table (more than 2^20 bins)" );
      return PairDD( cosvals[n], sinvals[n] );
    }
    const PairDD a = calcPhase( 1, n ), b = calcPhase( k - 1, n );
    return PairDD( a.first*b.first - a.second*b.second, a.first*b.second + a.second*b.first );
  }
  // W[k] = exp(i 2 pi k / size), the reference's twiddle table (initWTable: even entries from calcPhase, odd entries
  // = W[1] * previous entry).  The entries do not depend on the table size they were generated for: calcPhase strips
  // common factors of two, so W_2N[2k] == W_N[k] exactly, and an odd entry is the same product in either
  // formulation.  Hence (a) one table of the largest size serves every transform length (stride access) and (b) the
  // table of size 2N follows from the table of size N in one pass -- which is how it is built here (the recursive
  // definition costs ~popcount(k) complex products per entry: 5 ms for 65536 entries, more than all the FFTs of an
  // expansion).  Process-wide cache, grown on demand.
  inline const std::vector<Cplx>& twiddles( unsigned log2size )
  {
    static std::mutex mtx;
    static std::vector<std::vector<Cplx>> tables;     // [L] = table of size 2^L (kept: references stay valid)
    std::lock_guard<std::mutex> g( mtx );
    if ( tables.empty() ) { tables.reserve( 32 ); tables.push_back( std::vector<Cplx>( 1, Cplx{ 1.0, 0.0 } ) ); }
    while ( tables.size() <= log2size ) {
      const unsigned L = (unsigned)tables.size();
      const std::vector<Cplx>& half = tables.back();
      const PairDD p1 = calcPhase( 1, L );
      std::vector<Cplx> w( (size_t)1 << L );
      for ( size_t i = 0; i < w.size(); i += 2 ) {
        w[i] = half[i/2];
        w[i+1] = Cplx{ p1.first*w[i].re - p1.second*w[i].im, p1.first*w[i].im + p1.second*w[i].re };
      }
      tables.push_back( std::move( w ) );
    }
    return tables[log2size];
  }
  // the same table by the reference's own recipe (kept for the unit test of the construction above)
  inline std::vector<Cplx> makeTwiddles( unsigned log2size )
  {
    const unsigned size = 1u << log2size;
    std::vector<Cplx> w( size );
    PairDD v( 1.0, 0.0 );
    const PairDD p1 = log2size ? calcPhase( 1, log2size ) : PairDD( 1.0, 0.0 );
    for ( unsigned i = 0; i < size; ++i ) {
      if ( i % 2 == 1 ) v = PairDD( p1.first*v.first - p1.second*v.second, p1.first*v.second + p1.second*v.first );
      else v = calcPhase( i, log2size );
      w[i] = Cplx{ v.first, v.second };
    }
    return w;
  }

  // ---------------------------------------------------------------------------------------------------------
  struct Input {
    double emin = 0.0, emax = 0.0;   // VDOSData::vdos_egrid()
    VectD density;                   // VDOSData::vdos_density() (regular grid over [emin,emax])
    double temperature = 0.0, mass_amu = 0.0, bound_xs = 0.0;
  };

  // checkIsRegularVDOSGrid: emax of the exactly regular grid, 0 if [0,emin] is not a whole number of bins
  inline double regularGridEmax( double emin, double emax, size_t npts, double tolerance = 1e-6 )
  {
    const double binwidth_approx = ( emax - emin )/( npts - 1 );
    const double nbelow = emin/binwidth_approx;
    if ( nbelow < 0.99 || std::fabs( nbelow - std::round( nbelow ) ) > tolerance ) return 0.0;
    const unsigned long nb = static_cast<unsigned long>( nbelow + 0.5 );
    const double binwidth = emin/nb;
    return emin + ( npts - 1 )*binwidth;
  }

  // regulariseVDOSGrid: re-sample a curve given on an arbitrary grid (2 end points or one energy per density value)
  // so that [0,emin] is a whole number of bins.
  inline void regularise( const VectD& egrid, const VectD& density, double& out_emin, double& out_emax, VectD& out_density )
  {
    if ( !( density.size() > 2 ) || !( egrid.size() == 2 || egrid.size() == density.size() ) || !isGrid( egrid ) || !( egrid.front() >= 0.0 ) )
      throw Error( "BadInput", "invalid VDOS input arrays" );
    if ( egrid.front() < 1e-5 )
      throw Error( "BadInput", "VDOS energy range can not be specified for values less than 1e-5eV = 0.01meV" );
    const double tolerance = 1e-6;
    bool linear = true;
    if ( egrid.size() > 2 ) {
      const double bw = ( egrid.back() - egrid.front() )/( egrid.size() - 1.0 ), eps = tolerance*bw;
      for ( size_t i = 0; i < egrid.size(); ++i )
        if ( std::fabs( ( egrid.front() + i*bw ) - egrid[i] ) > eps ) { linear = false; break; }
    }
    const double already = linear ? regularGridEmax( egrid.front(), egrid.back(), density.size(), tolerance ) : 0.0;
    if ( already ) { out_emin = egrid.front(); out_emax = already; out_density = density; return; }
    const double emin = egrid.front(), old_emax = egrid.back();
    const double span = old_emax - emin, span_div_emin = span/emin;
    double k = dmax( 1.0, std::round( 2000.0/span_div_emin ) );
    PairDD best( kInf, 0.0 );
    while ( true ) {
      const double m = std::floor( span_div_emin*k );
      if ( m < 1.0 ) { k += 1.0; continue; }
      const double binwidth = emin/k;
      const double eps = dmax( 0.0, span - ( m*binwidth ) );
      const double kbest = best.second;
      if ( eps == 0.0 || kbest == 0.0 || ( kbest*std::floor( span_div_emin*k ) > k*std::floor( span_div_emin*kbest ) ) )
        best = PairDD( eps, k );
      double tol = 1e-6;
      if ( m > 5000 ) { tol = 1e-5; if ( m > 10000 ) { tol = 1e-4; if ( m > 15000 ) tol = m > 19000 ? 1e-2 : 1e-3; } }
      if ( best.first < span*tol ) break;
      if ( m >= 20000 )
        throw Error( "BadInput", "Could not regularise input energy grid. Are the energy ranges highly unusual?" );
      k += 1.0;
    }
    const double new_binwidth = emin/best.second;
    const double mm = std::floor( span_div_emin*best.second );
    unsigned new_npts = static_cast<unsigned>( mm + 0.5 ) + 1;
    double new_emax = emin + new_binwidth*( new_npts - 1 );
    if ( new_emax < old_emax ) { ++new_npts; new_emax = emin + new_binwidth*( new_npts - 1 ); }
    double inv_bw_orig = 0.0;
    if ( egrid.size() == 2 ) inv_bw_orig = ( density.size() - 1.0 )/( egrid.back() - egrid.front() );
    const VectD eg = egrid.size() == 2 ? linSpace( egrid.front(), egrid.back(), (unsigned)density.size() ) : egrid;
    out_density.clear(); out_density.reserve( new_npts );
    size_t it = 0;
    const size_t last = eg.size() - 1;
    for ( unsigned ieval = 0; ieval < new_npts; ++ieval ) {
      const double ev = ( ieval + 1 == new_npts ) ? new_emax : emin + new_binwidth*ieval;
      while ( it != last && ev >= eg[it+1] ) ++it;
      if ( ev == eg[it] ) { out_density.push_back( density[it] ); continue; }
      if ( it == last ) { out_density.push_back( ev > eg[it] ? 0.0 : density.back() ); continue; }
      const double y0 = density[it], y1 = density[it+1], x0 = eg[it], x1 = eg[it+1];
      StableSum s;
      s.add( y0*x1 ); s.add( -y0*ev ); s.add( ev*y1 ); s.add( -x0*y1 );
      out_density.push_back( inv_bw_orig ? s.sum()*inv_bw_orig : s.sum()/( x1 - x0 ) );
    }
    out_emin = emin; out_emax = new_emax;
  }

  // ---------------------------------------------------------------------------------------------------------
  // VDOSEval: the density curve continued below emin by a parabola, its integrals and Sjolander's G_1.
  class Eval {
  public:
    explicit Eval( const Input& in )
      : m_density( in.density ), m_emin( in.emin ), m_emax( in.emax ), m_kT( kBoltzmann*in.temperature ),
        m_temperature( in.temperature ), m_mass( in.mass_amu )
    {
      if ( !( in.temperature > 0.0 ) || !( in.mass_amu > 0.5 && in.mass_amu < 2000.0 ) || m_density.size() < 2 || !( m_emax > m_emin ) )
        throw Error( "BadInput", "invalid VDOS data" );
      if ( m_emin < 1e-5 )
        throw Error( "BadInput", "VDOS energy range should not be specified for values less than 1e-5eV = 0.01meV" );
      if ( !regularGridEmax( m_emin, m_emax, m_density.size() ) )
        throw Error( "BadInput", "Received non-regularised VDOS. The VDOSEval class expects regularised equidistant grid which can be extended downwards and exactly coincide with 0." );
      // (the reference applies the corrected emax only in its verbose mode; the default keeps emax as given)
      const double binwidth = ( m_emax - m_emin )/( m_density.size() - 1 );
      const unsigned long nbelow = static_cast<unsigned long>( m_emin/binwidth + 0.5 );
      m_npts_extended = (unsigned)( m_density.size() + nbelow );
      m_k = m_density.front()/( m_emin*m_emin );
      m_binwidth = binwidth;
      m_invbinwidth = 1.0/m_binwidth;
      StableSum s;
      constexpr double onethird = 1.0/3.0;
      s.add( onethird*m_density.front()*m_emin );
      integrateBins( []( double ) { return 1.0; }, s );
      if ( !( s.sum() > 0.0 ) ) throw Error( "BadInput", "VDOS density integrates to zero" );
      const double scale = 1.0/s.sum();
      for ( double& e : m_density ) e *= scale;
      m_k *= scale;
    }
    double kT() const { return m_kT; }
    double temperature() const { return m_temperature; }
    double emax() const { return m_emax; }
    unsigned nptsExtended() const { return m_npts_extended; }

    double eval( double energy ) const
    {
      if ( energy <= m_emin ) return m_k*energy*energy;
      double relpos = dclamp( ( energy - m_emin )*m_invbinwidth, -0.5, m_density.size() + 0.5 );
      const int ibin = static_cast<int>( relpos );
      if ( ibin >= static_cast<int>( m_density.size() - 1 ) ) return 0.0;
      relpos = dclamp( relpos - ibin, 0.0, 1.0 );
      return ( 1.0 - relpos )*m_density[ibin] + relpos*m_density[ibin+1];
    }
    double calcGamma0() const
    {
      const double twokT = 2.0*m_kT, inv2kT = 1.0/twokT;
      StableSum s;
      s.add( m_k*( twokT*twokT )*romberg33( xcothx, 0.0, m_emin*inv2kT ) );
      integrateBins( [inv2kT]( double e ) { return 1.0/( e*std::tanh( e*inv2kT ) ); }, s );
      return m_emax*s.sum();
    }
    double getMSD( double gamma0 ) const
    {
      constexpr double convfact = 0.5*( kHbar*kHbar )*( kSpeedOfLight*kSpeedOfLight )/kDalton2eVc2;
      return convfact*gamma0/( m_mass*m_emax );
    }
    // G_1 at (-energy, +energy)
    PairDD g1AsymmetricPair( double energy, double gamma0 ) const
    {
      if ( energy < 200.0*m_kT ) {
        const double sym = g1Symmetric( energy, gamma0 );
        if ( !sym ) return PairDD( 0.0, 0.0 );
        const double dbfact = std::exp( energy/( 2*m_kT ) );
        return PairDD( sym*dbfact, sym/dbfact );
      }
      const double kkk = eval( energy )*m_emax/( energy*gamma0 );
      if ( !kkk ) return PairDD( 0.0, 0.0 );
      const double e_div_kT = energy/m_kT;
      return PairDD( -kkk/std::expm1( -e_div_kT ), kkk*recExpm1( e_div_kT ) );
    }
  private:
    VectD m_density;
    double m_emin, m_emax, m_kT, m_temperature, m_mass;
    double m_k = 0.0, m_binwidth = 0.0, m_invbinwidth = 0.0;
    unsigned m_npts_extended = 0;

    static double recExpm1( double x )
    {
      if ( x <= 700 ) return 1.0/std::expm1( x );
      const double emx = std::exp( -x );
      return emx/( 1.0 - emx );
    }
    static double xcothx( double x )
    {
      if ( x < 0.1 ) {
        constexpr double c0 = 1, c2 = 1./3, c4 = -1./45, c6 = 2./945, c8 = -1./4725, c10 = 2./93555, c12 = -1382./638512875., c14 = 4./18243225.;
        const double y = x*x;
        return c0+y*(c2+y*(c4+y*(c6+y*(c8+y*(c10+y*(c12+y*c14))))));
      }
      return x/std::tanh( x );
    }
    static double xdivsinhx( double x )
    {
      if ( x < 0.07 ) {
        constexpr double c2 = -1./6., c4 = 7./360., c6 = -31./15120., c8 = 127./604800.;
        const double y = x*x;
        return 1.0+y*(c2+y*(c4+y*(c6+y*c8)));
      }
      return x/std::sinh( x );
    }
    double g1Symmetric( double energy, double gamma0 ) const
    {
      const double twokT = 2*m_kT, u = energy/twokT;
      if ( energy <= m_emin ) return ( m_k*m_kT*m_emax/gamma0 )*xdivsinhx( u );
      return eval( energy )*m_emax/( energy*2.0*gamma0*std::sinh( u ) );
    }
    // integral of density(E)*f(E) over [emin,emax], bin by bin (17-point Romberg each)
    template <class Fn> void integrateBins( Fn f, StableSum& sum ) const
    {
      const unsigned nbins = (unsigned)m_density.size() - 1;
      auto binIntegral = [this,&f,nbins]( unsigned ibin ) {
        const double d0 = m_density[ibin], d1 = m_density[ibin+1];
        const double e0 = m_emin + m_binwidth*ibin;
        const double e1 = ( ibin + 1 == nbins ? m_emax : m_emin + m_binwidth*( ibin + 1 ) );
        const double A = ( d1 - d0 )*m_invbinwidth, B = d0 - e0*A;
        auto g = [&f,A,B]( double e ) { return f( e )*( A*e + B ); };
        return romberg17( g, e0, e1 );
      };
      const unsigned nthreads = nbins >= 2000 ? std::min<unsigned>( 8, std::max<unsigned>( 1, std::thread::hardware_concurrency() ) ) : 1;
      if ( nthreads <= 1 ) {
        for ( unsigned ibin = 0; ibin < nbins; ++ibin ) sum.add( binIntegral( ibin ) );
        return;
      }
      // finely binned curve: the bin integrals are independent (host threads); they are ADDED in bin order
      VectD contrib( nbins );
      std::vector<std::thread> pool;
      const unsigned per = ( nbins + nthreads - 1 )/nthreads;
      for ( unsigned t = 0; t < nthreads; ++t ) {
        const unsigned b0 = t*per, b1 = std::min( nbins, b0 + per );
        if ( b0 >= b1 ) break;
        pool.emplace_back( [&contrib,&binIntegral,b0,b1]() { for ( unsigned ibin = b0; ibin < b1; ++ibin ) contrib[ibin] = binIntegral( ibin ); } );
      }
      for ( auto& th : pool ) th.join();
      for ( unsigned ibin = 0; ibin < nbins; ++ibin ) sum.add( contrib[ibin] );
    }
  };

  // ---------------------------------------------------------------------------------------------------------
  // G_n ladder.  The spectra live with the backend; the host keeps their grids and statistics.
  struct GnMeta {
    double lower = 0.0, upper = 0.0, binwidth = 0.0;
    size_t n = 0;
    unsigned long thin = 1;        // binwidth = thin * binwidth of G_1
    double maxval = 0.0;
    long first_above = -1, last_above = -1;   // first / last bin with density >= relthr*maxval (relthr of the expansion)
  };
  struct ConvJob {
    unsigned order, o1, o2;        // G_order = G_o1 (x) G_o2
    unsigned stride1, stride2;     // on-demand thinning of an input (take every stride'th bin)
    size_t n1, n2;                 // input lengths after that thinning
    double dt;                     // common bin width of the inputs
    bool trunc_thin;               // truncation + thinning apply at this order
    double trunc_threshold; unsigned thin_nbins; bool gentle_thinning;   // TruncAndThinningParams; order <= 2*minOrder
    double relthr;
  };
  struct ConvResult {
    size_t ifront = 0;             // bins dropped at the low edge by the truncation
    size_t n = 0;                  // final length
    unsigned long extra_thin = 1;
    double maxval = 0.0;
    long first_above = -1, last_above = -1;
  };
  struct TruncThin { int min_order = 5; unsigned thin_nbins = 1000; double trunc_threshold = 1e-14; };

  // Backend concept:
  //   void setSpectrum( unsigned order, const VectD& spec );                       (order 1, already normalised)
  //   void convolve( const std::vector<ConvJob>&, std::vector<ConvResult>& );      (all jobs of a batch at once)
  //   VectD spectrum( unsigned order );                                            (read back)
  //   void fill( const FillPlan&, VectD& sab );
  struct FillPlan {
    unsigned norders = 0;
    std::vector<GnMeta> meta;                 // [order-1]
    VectD scale;                              // [order-1] contribution scale of the order
    VectD alpha_factor;                       // [(order-1)*nalpha + ia]
    std::vector<int> a_first, a_end;          // [order-1]: the contiguous run of positive alpha factors
    std::vector<unsigned char> skip;          // [order-1]: order contributes nothing (Stirling prefactor underflow)
    std::vector<PairDD> job_orders;           // (first, last) order of every summation group, in summation order
    VectD beta_nonpos, expbeta;               // the non-positive part of the beta grid and exp(beta) there
    size_t nalpha = 0, nbeta = 0, idx_zero = 0, idx_firstflip = 0;
    double kT = 0.0;
  };

  template <class Backend>
  class Ladder {
  public:
    // gamma0 = ev.calcGamma0() (passed in: the caller needs it as well and it is an integral over the whole curve)
    Ladder( const Eval& ev, double gamma0, Backend& be, TruncThin tt, double relthr ) : m_be( be ), m_tt( tt ), m_relthr( relthr ), m_kT( ev.kT() )
    {
      unsigned long nbins = ev.nptsExtended() - 1;
      constexpr unsigned long min_nbins = 400;
      const unsigned long thicken = static_cast<unsigned long>( std::ceil( double( min_nbins )/nbins ) );
      nbins *= thicken;
      if ( !( nbins < 10000000 ) ) throw Error( "CalcError", "VDOS grid too fine" );
      const VectD egrid = linSpace( 0.0, ev.emax(), (unsigned)( nbins + 1 ) );
      const double binwidth = egrid.back()/nbins;
      VectD g1( egrid.size()*2 - 1, 0.0 );
      for ( size_t i = 0; i < egrid.size(); ++i ) {
        const PairDD v = ev.g1AsymmetricPair( egrid[i], gamma0 );
        g1[nbins+i] = v.second;
        g1[nbins-i] = v.first;
      }
      if ( !( m_tt.trunc_threshold >= 0.0 && m_tt.trunc_threshold <= 0.1 ) || m_tt.min_order < -1 )
        throw Error( "BadInput", "invalid truncation/thinning parameters" );
      // at most one zero at each edge
      double lower = -ev.emax();
      size_t first = 0, last = g1.size() - 1;
      while ( first != last && !( g1[first] > 0.0 ) && !( g1[first+1] > 0.0 ) ) ++first;
      while ( last != 0 && !( g1[last] > 0.0 ) && !( g1[last-1] > 0.0 ) ) --last;
      if ( first >= last || last - first < 3 ) throw Error( "CalcError", "Too few non-zero pts in G1 spectrum." );
      if ( first != 0 || last != g1.size() - 1 ) {
        lower += first*binwidth;
        g1 = VectD( g1.begin() + first, g1.begin() + last + 1 );
      }
      GnMeta m;
      m.lower = lower; m.binwidth = binwidth; m.n = g1.size(); m.thin = 1;
      m.upper = m.lower + ( m.n - 1 )*m.binwidth;
      normalise( g1, binwidth );
      m.maxval = *std::max_element( g1.begin(), g1.end() );
      aboveRange( g1, m );
      m_g1 = g1;
      m_meta.push_back( m );
      m_be.setSpectrum( 1, g1 );
    }
    unsigned maxOrder() const { return (unsigned)m_meta.size(); }
    double kT() const { return m_kT; }
    const GnMeta& meta( unsigned n ) const { return m_meta.at( n - 1 ); }
    const std::vector<GnMeta>& allMeta() const { return m_meta; }
    const VectD& g1() const { return m_g1; }
    PairDD eRange( unsigned n ) const { const GnMeta& m = meta( n ); return PairDD( m.lower, m.upper ); }
    // energy range where the spectrum is at least relthr (the expansion's level) of its maximum
    PairDD eRangeAbove( unsigned n ) const
    {
      const GnMeta& m = meta( n );
      PairDD r( m.lower, m.upper );
      if ( m.first_above >= 0 ) r.first = m.lower + m.first_above*m.binwidth;
      if ( m.last_above >= 0 ) r.second = dmin( r.second, m.lower + m.last_above*m.binwidth );
      return r;
    }
    // make every order up to `target` available; orders are produced in batches of all that the existing ones allow
    void grow( unsigned target, unsigned lookahead = 64 )
    {
      while ( maxOrder() < target ) {
        const unsigned have = maxOrder();
        const unsigned upto = std::min<unsigned>( 2*have, std::max<unsigned>( target, have + lookahead ) );
        std::vector<ConvJob> jobs;
        for ( unsigned n = have + 1; n <= upto; ++n ) jobs.push_back( planJob( n ) );
        std::vector<ConvResult> res;
        m_be.convolve( jobs, res );
        for ( size_t j = 0; j < jobs.size(); ++j ) {
          const ConvJob& J = jobs[j]; const ConvResult& R = res[j];
          const GnMeta& p1 = meta( J.o1 ); const GnMeta& p2 = meta( J.o2 );
          GnMeta m;
          double start = p1.lower + p2.lower;
          double dt = J.dt;
          if ( J.trunc_thin && m_tt.trunc_threshold > 0 ) start += R.ifront*dt;
          if ( R.extra_thin > 1 ) dt *= R.extra_thin;
          m.lower = start; m.binwidth = dt; m.n = R.n;
          m.thin = p1.thin*J.stride1*R.extra_thin;
          m.upper = m.lower + ( m.n - 1 )*m.binwidth;
          m.maxval = R.maxval; m.first_above = R.first_above; m.last_above = R.last_above;
          // (orders produced ahead of need may degenerate harmlessly; one that is asked for must be a spectrum)
          if ( J.order <= target && ( !( m.maxval > 0.0 ) || !( m.n > 3 ) || !( m.maxval <= std::numeric_limits<double>::max() ) ) )
            throw Error( "CalcError", "VDOS expansion: degenerate phonon spectrum at order "+std::to_string( J.order ) );
          m_meta.push_back( m );
        }
      }
    }
  private:
    Backend& m_be;
    TruncThin m_tt;
    double m_relthr, m_kT;
    std::vector<GnMeta> m_meta;
    VectD m_g1;

    static void normalise( VectD& spec, double binwidth )
    {
      double area = 0.;
      for ( double v : spec ) area += v;
      area *= binwidth;
      const double inv = 1.0/area;
      for ( double& v : spec ) v *= inv;
    }
    void aboveRange( const VectD& spec, GnMeta& m ) const
    {
      const double thr = m_relthr*m.maxval;
      m.first_above = m.last_above = -1;
      for ( size_t i = 0; i < spec.size(); ++i ) if ( spec[i] >= thr ) { m.first_above = (long)i; break; }
      for ( size_t i = spec.size(); i > 0; --i ) if ( spec[i-1] >= thr ) { m.last_above = (long)( i-1 ); break; }
    }
    ConvJob planJob( unsigned order ) const
    {
      ConvJob J;
      J.order = order; J.o2 = order/2; J.o1 = order - J.o2;
      const GnMeta& p1 = meta( J.o1 ); const GnMeta& p2 = meta( J.o2 );
      J.stride1 = J.stride2 = 1;
      J.n1 = p1.n; J.n2 = p2.n;
      if ( p1.thin == p2.thin ) {
        J.dt = p1.binwidth;
      } else {
        J.dt = std::max<double>( p1.binwidth, p2.binwidth );
        if ( p1.thin > p2.thin ) {
          const unsigned long f = p1.thin/p2.thin;
          if ( p1.thin % p2.thin || !floatEq( J.dt, p2.binwidth*f ) ) throw Error( "CalcError", "incompatible thinning factors" );
          J.stride2 = (unsigned)f; J.n2 = ( p2.n + f - 1 )/f;
        } else {
          const unsigned long f = p2.thin/p1.thin;
          if ( p2.thin % p1.thin || !floatEq( J.dt, p1.binwidth*f ) ) throw Error( "CalcError", "incompatible thinning factors" );
          J.stride1 = (unsigned)f; J.n1 = ( p1.n + f - 1 )/f;
        }
      }
      J.trunc_thin = m_tt.min_order >= 0 && order >= static_cast<unsigned>( m_tt.min_order );
      J.trunc_threshold = m_tt.trunc_threshold; J.thin_nbins = m_tt.thin_nbins;
      J.gentle_thinning = order <= static_cast<unsigned>( m_tt.min_order*2 );
      J.relthr = m_relthr;
      return J;
    }
  };

  // ---------------------------------------------------------------------------------------------------------
  constexpr double alpha2xFactor( double kT, double msd ) { return ( 2.0*kNeutronMassEvc2/( kHbar*kHbar ) )*kT*msd; }

  // the two x with x^n e^-x = eps * (peak value)
  inline PairDD rangeXNexpMX( unsigned n, double eps, double accuracy = 1e-13 )
  {
    const double fn = static_cast<double>( n );
    const double k = kInvE*std::pow( eps, 1.0/fn );
    auto f = [k]( double y ) { return y*std::exp( -y ) - k; };
    return PairDD( fn*findRoot( f, 0.0, 1.0, accuracy ), fn*findRoot( f, 1.0, 700.0, accuracy ) );
  }
  inline bool withinAlphaPlusCurve( double c, double alpha, double beta )
  {
    const double cpb = c + beta;
    if ( cpb < 0.0 ) return false;
    const double t = 0.5*( alpha - beta ) - c;
    return t <= 0.0 || c*cpb >= t*t;
  }
  inline PairDD extremePointWithinAlphaPlusCurve( double c, PairDD arange, PairDD brange )
  {
    if ( brange.second <= -c ) return PairDD( -1.0, -1.0 );
    auto alphaPlus = [c]( double beta ) { return 2*c + beta + 2*std::sqrt( c*( c + beta ) ); };
    const double apb1 = alphaPlus( brange.second );
    if ( apb1 <= arange.first ) return PairDD( -1.0, -1.0 );
    brange.first = dmax( brange.first, -c );
    const double apb0 = alphaPlus( brange.first );
    if ( apb0 >= arange.second ) return PairDD( arange.second, brange.first );
    arange.second = dmin( arange.second, apb1 );
    if ( apb0 < arange.first ) brange.first = arange.first - 2.0*std::sqrt( c*arange.first );
    return PairDD( arange.second, brange.first );
  }

  inline VectD setupAlphaGrid( double kT, double msd, double alphaMax, unsigned npts )
  {
    if ( npts < 20 ) throw Error( "CalcError", "too few alpha points" );
    const double x2alpha = 1.0/alpha2xFactor( kT, msd );
    const double alphaMin = x2alpha*1e-50, alphaMin2 = x2alpha*1e-10, alphaG1Maxx = x2alpha*1.0, alphaG15Maxx = x2alpha*15.0;
    const unsigned n0 = static_cast<unsigned>( npts*0.15 + 0.5 );
    const unsigned n1 = static_cast<unsigned>( npts*0.29 + 0.5 );
    const unsigned n2 = static_cast<unsigned>( npts*0.23 + 0.5 );
    const unsigned n3 = npts - ( n1 + n2 + n0 );
    const unsigned npts_123 = n1 + n2 + n3;
    const double alpha_upscatmax = ( n0 < 10 ? 6.0 : ( n0 > 50 ? 14.0 : 10.0 ) );
    const VectD region0 = linSpace( dmin( 1e-3, alphaMax*0.01 ), dmin( alpha_upscatmax, alphaMax*0.99 ), n0 );
    // region-0 points are merged in and then moved half way between their neighbours
    auto finalise = [&region0,npts]( const VectD& grid ) {
      std::vector<std::pair<double,bool>> all;
      for ( double e : grid ) all.emplace_back( e, false );
      for ( double e : region0 ) all.emplace_back( e, true );
      if ( all.size() != npts ) throw Error( "CalcError", "alpha grid size mismatch" );
      std::stable_sort( all.begin(), all.end() );
      for ( size_t i = 1; i + 1 < all.size(); ++i )
        if ( all[i].second ) all[i].first = 0.5*( all[i-1].first + all[i+1].first );
      VectD out;
      for ( auto& e : all ) out.push_back( e.first );
      if ( !isGrid( out ) ) throw Error( "CalcError", "alpha grid not ascending" );
      return out;
    };
    if ( alphaMax <= alphaMin*100.0 )
      return finalise( linSpace( alphaMax*0.001, alphaMax, npts_123 ) );
    VectD grid;
    grid.push_back( alphaMin );
    auto append = [&grid]( const VectD& v, size_t skip_front = 0, size_t skip_back = 0 ) {
      grid.insert( grid.end(), v.begin() + skip_front, v.end() - skip_back );
    };
    if ( alphaMax <= alphaG1Maxx*10.0 ) { append( linSpace( alphaMin2, alphaMax, npts_123 - 1 ) ); return finalise( grid ); }
    append( linSpace( alphaMin2, alphaG1Maxx, n1 - 1 ) );
    if ( alphaMax < 2.0*alphaG15Maxx ) { append( linSpace( alphaG1Maxx, alphaMax, n2 + n3 + 2 ), 1, 1 ); return finalise( grid ); }
    append( linSpace( alphaG1Maxx, alphaG15Maxx, n2 + 2 ), 1, 1 );
    append( geomSpace( alphaG15Maxx, alphaMax, n3 ) );
    return finalise( grid );
  }

  template <class LadderT>
  inline VectD setupBetaGrid( const LadderT& Gn, double betaMax, unsigned vdoslux, unsigned override_nbins = 0 )
  {
    const double invkT = 1.0/Gn.kT();
    const double G1 = std::fabs( Gn.eRange( 1 ).first*invkT );
    double G3 = std::fabs( Gn.eRange( std::min<unsigned>( 3, Gn.maxOrder() ) ).first*invkT );
    if ( G1 == G3 ) G3 = G1*1.0001;
    betaMax = dmax( betaMax, G1*1.01 );
    G3 = dmin( G3, betaMax*0.9999 );
    const double D = dmin( betaMax*0.9999, dmax( G3, std::max<int>( 2, (int)vdoslux - 1 )*10.0 ) );
    const unsigned ntotal = override_nbins ? override_nbins : 100*( 1u << vdoslux );
    VectD grid; grid.reserve( ntotal );
    VectD spec = Gn.g1();
    PairDD er = Gn.eRange( 1 );
    er.first *= invkT; er.second *= invkT;
    VectD evals = linSpace( er.first, er.second, (unsigned)spec.size() );
    const double epsilon = -0.1*Gn.meta( 1 ).binwidth;
    if ( !( evals.front() < 0.0 ) || !( evals.back() > 0.0 ) || !( evals.front() < epsilon ) ) throw Error( "CalcError", "unexpected G1 range" );
    while ( evals.back() > epsilon ) evals.pop_back();
    spec.resize( evals.size() );
    const unsigned n1_max = static_cast<unsigned>( evals.size()*1.25 + 6.5 ) & ~1u;
    const unsigned n0 = 1;
    if ( !( D > G1 ) || !( betaMax > D ) ) throw Error( "CalcError", "beta grid regions degenerate" );
    double L1 = 2*G1, L2 = 2*( D - G1 ), L3 = betaMax - D;
    L1 *= 4; L2 *= 2;
    const double Lnorm = 1.0/( L1 + L2 + L3 );
    double f1 = dmax( 0.2, L1*Lnorm ), f2 = L2*Lnorm;
    if ( f1 + f2 > 0.99 ) { const double t = 0.99/( f1 + f2 ); f1 *= t; f2 *= t; }
    unsigned n1 = std::min<unsigned>( n1_max, static_cast<unsigned>( f1*ntotal + 0.5 )/2 );
    unsigned n2 = static_cast<unsigned>( f2*ntotal + 0.5 )/2;
    n1 = std::max<unsigned>( 15, n1 );
    n2 = std::max<unsigned>( 1, n2 );
    unsigned n012;
    while ( true ) {
      n012 = 2*( n1 + n2 ) + n0;
      if ( n012 >= ntotal - 1 ) { if ( n2 > n1 ) --n2; else --n1; } else break;
    }
    if ( n1 < 10 ) throw Error( "CalcError", "too few beta points for the one-phonon region" );
    const unsigned n3 = ntotal - n012;
    { const VectD v = linSpace( -betaMax, -D, n3 + 1 ); grid.insert( grid.begin(), v.begin(), v.end() - 1 ); }
    const size_t idx_r2start = grid.size();
    { const VectD v = linSpace( -D, -G1, n2 + 1 ); grid.insert( grid.end(), v.begin(), v.end() - 1 ); }
    {
      unsigned n1_near0 = std::max<unsigned>( 5, static_cast<unsigned>( n1*0.2 + 0.5 ) );
      unsigned n1_spectrum = n1 - n1_near0;
      if ( n1_spectrum >= evals.size() ) {
        n1_spectrum = (unsigned)evals.size();
        n1_near0 = n1 - n1_spectrum;
      } else {
        unsigned n_for_gaps = ( ( n1_spectrum > 30 && evals.size() - n1_spectrum > 10 )
                                ? std::max<unsigned>( 5, static_cast<unsigned>( n1_spectrum*0.1 + 0.5 ) ) : 0 );
        n1_spectrum -= n_for_gaps;
        const double g1bw = evals.at( 1 ) - evals.at( 0 );
        std::tie( evals, spec ) = reducePoints( evals, spec, n1_spectrum );
        if ( n_for_gaps > 0 ) {
          struct Gap {
            double b0, b1; unsigned n;
            bool operator<( const Gap& o ) const
            {
              const double a = ( b1 - b0 )/( n + 1 ), b = ( o.b1 - o.b0 )/( o.n + 1 );
              if ( floatEq( a, b, 1e-13, 1e-13 ) ) return b0 > o.b0;
              return a > b;
            }
          };
          std::vector<Gap> gaps;
          const double dmin_gap = 1.5*g1bw;
          for ( size_t i = 0; i + 1 < evals.size(); ++i )
            if ( evals[i+1] - evals[i] > dmin_gap ) gaps.push_back( Gap{ evals[i], evals[i+1], 0 } );
          while ( n_for_gaps > 0 && !gaps.empty() ) {
            std::stable_sort( gaps.begin(), gaps.end() );
            gaps.front().n += 1;
            --n_for_gaps;
          }
          for ( const Gap& g : gaps )
            if ( g.n > 0 ) {
              const double bw = ( g.b1 - g.b0 )/( g.n + 1.0 );
              for ( unsigned i = 0; i < g.n; ++i ) evals.push_back( g.b0 + ( i + 1 )*bw );
            }
          std::sort( evals.begin(), evals.end() );
        }
        if ( n_for_gaps > 0 ) n1_near0 += n_for_gaps;
      }
      for ( double e : evals ) grid.push_back( e );
      VectD v = geomSpace( dmin( 1e-50, -0.001*grid.back() ), -grid.back()*0.1, n1_near0 );
      std::reverse( v.begin(), v.end() );
      for ( double e : v ) grid.push_back( -e );
    }
    const size_t idx_r1back = grid.size() - 1;
    grid.push_back( 0.0 );
    for ( size_t i = idx_r1back + 1; i-- > idx_r2start; ) grid.push_back( -grid[i] );
    if ( grid.size() != ntotal || !isGrid( grid ) || grid.front() != -betaMax || grid.back() != D )
      throw Error( "CalcError", "beta grid construction failed" );
    return grid;
  }

  inline double stirlingSeries9( double inv_n )
  {
    constexpr double c1 = 1./12., c2 = 1./288., c3 = -139/51840., c4 = -571./2488320., c5 = 163879./209018880.,
      c6 = 5246819./75246796800., c7 = -534703531./902961561600., c8 = -4483131259./86684309913600.,
      c9 = 432261921612371./514904800886784000.;
    return 1.0 + inv_n*(c1+inv_n*(c2+inv_n*(c3+inv_n*(c4+inv_n*(c5+inv_n*(c6+inv_n*(c7+inv_n*(c8+inv_n*c9))))))));
  }
  inline size_t closestIndex( const VectD& v, double value )
  {
    auto it = std::lower_bound( v.begin(), v.end(), value );
    if ( it == v.begin() ) return 0;
    if ( it == v.end() ) return v.size() - 1;
    return ( std::fabs( *it - value ) < std::fabs( *std::prev( it ) - value ) ? it : std::prev( it ) ) - v.begin();
  }

  // Everything of fillSABFromVDOS[Concurrent] that is not the sum itself: the alpha dependence
  // f(x,n) = e^-x x^n / n! for every order (recursively below order 16, through Stirling's series above), the
  // detailed-balance mirror of the negative-beta rows, and the fixed grouping of the orders into partial sums.
  inline FillPlan planFill( const std::vector<GnMeta>& meta, double kT, double msd, const VectD& alphaGrid, const VectD& betaGrid,
                            const std::function<double(unsigned)>& scaleFct )
  {
    FillPlan P;
    const unsigned norders = (unsigned)meta.size();
    const size_t na = alphaGrid.size();
    P.norders = norders; P.meta = meta; P.nalpha = na; P.nbeta = betaGrid.size(); P.kT = kT;
    P.scale.resize( norders ); P.alpha_factor.assign( (size_t)norders*na, 0.0 );
    P.a_first.assign( norders, 0 ); P.a_end.assign( norders, 0 ); P.skip.assign( norders, 0 );
    constexpr unsigned stirling_threshold = 16;
    const double alpha2x = alpha2xFactor( kT, msd );
    VectD x( na ), expmhalfx( na ), logx;
    for ( size_t i = 0; i < na; ++i ) { x[i] = alphaGrid[i]*alpha2x; expmhalfx[i] = std::exp( -0.5*x[i] ); }
    if ( norders >= stirling_threshold ) { logx.resize( na ); for ( size_t i = 0; i < na; ++i ) logx[i] = std::log( x[i] ); }
    VectD fxn = expmhalfx;
    P.idx_zero = closestIndex( betaGrid, 0.0 );
    P.idx_firstflip = closestIndex( betaGrid, -betaGrid.back() );
    if ( betaGrid[P.idx_zero] != 0.0 || betaGrid[P.idx_firstflip] != -betaGrid.back() )
      throw Error( "CalcError", "beta grid lacks beta=0 or the mirror of its upper end" );
    P.beta_nonpos.assign( betaGrid.begin(), betaGrid.begin() + P.idx_zero + 1 );
    P.expbeta.resize( P.beta_nonpos.size() );
    for ( size_t i = 0; i < P.beta_nonpos.size(); ++i ) P.expbeta[i] = std::exp( P.beta_nonpos[i] );
    for ( unsigned n = 1; n <= norders; ++n ) {
      P.scale[n-1] = scaleFct ? scaleFct( n ) : 1.0;
      if ( !( P.scale[n-1] >= 0.0 ) ) throw Error( "BadInput", "order weight function must return non-negative values" );
    }
    // one row of the table: orders below 16 by the recursion (sequential), the others independently of each other
    auto positiveRun = [&P,na]( unsigned n ) {
      const double* af = &P.alpha_factor[(size_t)( n-1 )*na];
      size_t first = 0;
      while ( first != na && !( af[first] > 0.0 ) ) ++first;
      size_t end = first;
      while ( end != na && af[end] > 0.0 ) ++end;
      P.a_first[n-1] = (int)first; P.a_end[n-1] = (int)end;
    };
    for ( unsigned n = 1; n <= norders && n < stirling_threshold; ++n ) {
      double* af = &P.alpha_factor[(size_t)( n-1 )*na];
      const double invn = 1.0/n;
      for ( size_t i = 0; i < na; ++i ) fxn[i] *= x[i]*invn;
      for ( size_t i = 0; i < na; ++i ) af[i] = fxn[i]*expmhalfx[i]*kT;
      positiveRun( n );
    }
    auto stirlingRows = [&]( unsigned n0, unsigned n1 ) {
      for ( unsigned n = n0; n <= n1; ++n ) {
        double* af = &P.alpha_factor[(size_t)( n-1 )*na];
        const double invn = 1.0/n;
        const double gn = stirlingSeries9( invn );
        const double fact = kT*kInvSqrt2Pi/( std::sqrt( n )*gn );
        if ( !fact ) { P.skip[n-1] = 1; continue; }
        const double logn = std::log( n );
        for ( size_t i = 0; i < na; ++i ) {
          const double exparg = n*( logx[i] - logn + 1.0 ) - x[i];
          af[i] = fact*std::exp( exparg );
        }
        positiveRun( n );
      }
    };
    if ( norders >= stirling_threshold ) {
      // (a vdoslux-5 expansion has ~1e3 orders x 1600 alpha points: the exponentials are spread over host threads;
      //  every row is computed by one thread, so the values do not depend on the split)
      const size_t work = (size_t)( norders - stirling_threshold + 1 )*na;
      unsigned nthreads = work > 200000 ? std::min<unsigned>( 8, std::max<unsigned>( 1, std::thread::hardware_concurrency() ) ) : 1;
      if ( nthreads <= 1 ) {
        stirlingRows( stirling_threshold, norders );
      } else {
        std::vector<std::thread> pool;
        const unsigned total = norders - stirling_threshold + 1, per = ( total + nthreads - 1 )/nthreads;
        for ( unsigned t = 0; t < nthreads; ++t ) {
          const unsigned n0 = stirling_threshold + t*per, n1 = std::min<unsigned>( norders, n0 + per - 1 );
          if ( n0 > norders ) break;
          pool.emplace_back( stirlingRows, n0, n1 );
        }
        for ( auto& th : pool ) th.join();
      }
    }
    // fixed groups of >= 16 orders (independent of any thread count), summed group by group
    const unsigned njobs = ( norders <= 16 ? 1 : norders/16 );
    if ( njobs == 1 ) {
      P.job_orders.push_back( PairDD( 1, norders ) );
    } else {
      const unsigned per_job = norders/njobs;
      unsigned next = 1;
      for ( unsigned j = 0; j < njobs; ++j ) {
        const unsigned lo = next;
        next += per_job;
        const unsigned hi = std::min<unsigned>( next - 1, norders );
        P.job_orders.push_back( PairDD( lo, hi ) );
      }
      // (orders beyond njobs*per_job are left out by the reference's grouping as well)
    }
    return P;
  }

  struct Kernel {
    VectD alpha, beta, sab;     // sab[ibeta*nalpha + ialpha]
    double temperature = 0.0, bound_xs = 0.0, mass_amu = 0.0, suggested_emax = 0.0;
    unsigned max_order = 0;
    double gamma0 = 0.0, msd = 0.0;
    unsigned ntrimmed = 0;
  };

  // detail_trimZeroEdgesFromKernel (transformKernelToStdFormat of a kernel that is already in S(alpha,beta) form)
  inline unsigned trimZeroEdges( Kernel& K )
  {
    const size_t na = K.alpha.size(), nb = K.beta.size();
    auto rowZero = [&]( size_t ib ) { for ( size_t i = ib*na; i != ( ib + 1 )*na; ++i ) if ( K.sab[i] ) return false; return true; };
    auto colZero = [&]( size_t ia ) { for ( size_t i = ia; i < K.sab.size(); i += na ) if ( K.sab[i] ) return false; return true; };
    size_t tbu = 0, tbl = 0, tau = 0;
    for ( size_t i = 0; i < na; ++i ) { const size_t ia = na - i - 1; if ( K.alpha[ia] > 0.0 && colZero( ia ) ) ++tau; else break; }
    for ( size_t i = 0; i < nb; ++i ) { const size_t ib = nb - i - 1; if ( K.beta[ib] > 0.0 && rowZero( ib ) ) ++tbu; else break; }
    for ( size_t ib = 0; ib < nb; ++ib ) { if ( K.beta[ib] < 0.0 && rowZero( ib ) ) ++tbl; else break; }
    if ( tau >= na ) tau = tbu = tbl = 0;
    const size_t ntot = tbu + tbl + tau;
    if ( !ntot ) return 0;
    VectD s;
    s.reserve( ( na - tau )*( nb - tbu - tbl ) );
    for ( size_t ib = tbl; ib < nb - tbu; ++ib )
      for ( size_t ia = 0; ia < na - tau; ++ia ) s.push_back( K.sab[ia + na*ib] );
    K.sab.swap( s );
    K.alpha.resize( na - tau );
    K.beta = VectD( K.beta.begin() + tbl, K.beta.begin() + ( nb - tbu ) );
    return (unsigned)ntot;
  }

  // createScatteringKernel + transformKernelToStdFormat
  template <class Backend>
  inline Kernel expand( const Input& in, unsigned vdoslux, double targetEmax_requested, Backend& be,
                        const std::function<double(unsigned)>& scaleFct = nullptr, TruncThin tt = TruncThin() )
  {
    if ( vdoslux > 5 ) throw Error( "BadInput", "vdoslux must be in 0..5" );
    if ( !( targetEmax_requested >= 0.0 ) ) throw Error( "BadInput", "target Emax must be non-negative" );
    constexpr double lux2emax[6] = { 0.5, 1.0, 3.0, 5.0, 8.0, 12.0 };
    double targetEmax = targetEmax_requested > 0.0 ? targetEmax_requested : lux2emax[vdoslux];
    NCB_VDOS_TICK_INIT;
    Eval ev( in );
    NCB_VDOS_TICK( "Eval (normalisation)" );
    const double kT = ev.kT(), invkT = 1.0/kT;
    const double gamma0 = ev.calcGamma0();
    NCB_VDOS_TICK( "gamma0" );
    const double msd = ev.getMSD( gamma0 );
    double targetEmax_div_kT = targetEmax*invkT;
    unsigned max_order = 4;
    const double relcontriblvl = std::pow( 10.0, -( 3.0 + 2.0*vdoslux ) );
    Ladder<Backend> Gn( ev, gamma0, be, tt, relcontriblvl );
    NCB_VDOS_TICK( "G1" );
    Gn.grow( max_order );
    unsigned order_limit = 1000;
    if ( targetEmax_requested > 0.0 || vdoslux == 5 ) order_limit *= 10;
    if ( vdoslux == 0 ) order_limit /= 10;
    const double emax_lowest_allowed = ( targetEmax_requested > 0.0 ? targetEmax_requested : 1e-15 );
    const double x2alpha = 1.0/alpha2xFactor( kT, msd );
    auto rangesOfOrder = [&]( unsigned n, PairDD& arange, PairDD& brange ) {
      const PairDD e = Gn.eRangeAbove( n );
      brange = PairDD( e.first*invkT, e.second*invkT );
      const PairDD xr = rangeXNexpMX( n, relcontriblvl );
      arange = PairDD( xr.first*x2alpha, xr.second*x2alpha );
    };
    while ( true ) {
      Gn.grow( max_order );
      PairDD ar, br;
      rangesOfOrder( max_order, ar, br );
      if ( withinAlphaPlusCurve( targetEmax_div_kT, ar.first, br.second ) ) ++max_order; else break;
      if ( max_order > order_limit ) {
        double reduced = targetEmax;
        do {
          reduced *= 0.99;
          if ( reduced < emax_lowest_allowed )
            throw Error( "CalcError", "VDOS expansion too slow - can not reach the requested energy after "+std::to_string( order_limit )+" phonon convolutions (likely causes: either the target energy value is too high, vdoslux too low, the temperature too high, or the VDOS is very unusual)." );
        } while ( withinAlphaPlusCurve( reduced*invkT, ar.first, br.second ) );
        targetEmax_div_kT = reduced*invkT;
        targetEmax = reduced;
        break;
      }
    }
    Gn.grow( max_order );
    NCB_VDOS_TICK( "order loop" );
    double betaMin = 0.0, alphaMax = 0.0;
    for ( unsigned n = 1; n <= max_order; ++n ) {
      PairDD ar, br;
      rangesOfOrder( n, ar, br );
      const PairDD ep = extremePointWithinAlphaPlusCurve( targetEmax_div_kT, ar, br );
      alphaMax = dmax( alphaMax, ep.first );
      betaMin = dmin( betaMin, ep.second );
    }
    if ( !( betaMin < 0.0 && alphaMax > 0.0 ) ) throw Error( "CalcError", "VDOS expansion: empty kinematic region" );
    const double upper_beta = -betaMin*1.01, upper_alpha = alphaMax*1.01;
    NCB_VDOS_TICK( "alpha/beta reach" );
    // (orders beyond max_order may exist in the ladder -- produced ahead in a batch; the reference's maxOrder() is max_order)
    struct View {
      const Ladder<Backend>& L; unsigned n;
      double kT() const { return L.kT(); }
      unsigned maxOrder() const { return n; }
      PairDD eRange( unsigned k ) const { return L.eRange( k ); }
      const GnMeta& meta( unsigned k ) const { return L.meta( k ); }
      const VectD& g1() const { return L.g1(); }
    } view{ Gn, max_order };
    Kernel K;
    K.beta = setupBetaGrid( view, upper_beta, vdoslux );
    NCB_VDOS_TICK( "beta grid" );
    K.alpha = setupAlphaGrid( kT, msd, upper_alpha, (unsigned)( K.beta.size()/2 ) );
    std::vector<GnMeta> meta( Gn.allMeta().begin(), Gn.allMeta().begin() + max_order );
    const FillPlan P = planFill( meta, kT, msd, K.alpha, K.beta, scaleFct );
    NCB_VDOS_TICK( "alpha grid + fill plan" );
    be.fill( P, K.sab );
    NCB_VDOS_TICK( "fill" );
    K.suggested_emax = targetEmax;
    if ( scaleFct && scaleFct( max_order ) == 0.0 ) K.suggested_emax = 0.0;
    K.temperature = ev.temperature(); K.bound_xs = in.bound_xs; K.mass_amu = in.mass_amu;
    K.max_order = max_order; K.gamma0 = gamma0; K.msd = msd;
    K.ntrimmed = trimZeroEdges( K );
    NCB_VDOS_TICK( "trim" );
    return K;
  }

  // A compiled material (ncb_blob.h) whose S(alpha,beta) leaves are given as phonon densities of states
  // (NCB_KIND_SABVDOS) is re-assembled with ordinary NCB_KIND_SAB leaves: expandFn( Input, vdoslux, target_emax )
  // returns the expanded kernel, the leaf's energy grid is left to the library (auto_egrid = 1).  Returns false when
  // the material has no such leaf (or is not a readable material: the loader reports what is wrong with it).
  template <class ExpandFn, class WarnFn>
  inline bool rewriteVdosLeaves( const void* blob_, size_t nbytes, std::vector<unsigned char>& out, ExpandFn&& expandFn, WarnFn&& warn )
  {
    const unsigned char* blob = static_cast<const unsigned char*>( blob_ );
    if ( nbytes < sizeof(ncb_header_t) ) return false;
    ncb_header_t hdr;
    std::memcpy( &hdr, blob, sizeof(hdr) );
    if ( hdr.magic != NCB_MAGIC || hdr.version != NCB_VERSION || hdr.nbytes > nbytes || hdr.ncomp == 0 || hdr.ncomp > NCB_MAXCOMP )
      return false;
    bool any = false;
    for ( uint32_t i = 0; i < hdr.ncomp; ++i ) any = any || hdr.comp[i].kind == NCB_KIND_SABVDOS;
    if ( !any ) return false;
    out.assign( sizeof(ncb_header_t), 0 );
    ncb_header_t nh = hdr;
    auto append = [&out]( const void* p, size_t n ) {
      const size_t off = out.size();
      out.resize( off + n );
      std::memcpy( out.data() + off, p, n );
    };
    for ( uint32_t i = 0; i < hdr.ncomp; ++i ) {
      const ncb_comp_t& c = hdr.comp[i];
      if ( c.off > hdr.nbytes || c.nbytes > hdr.nbytes - c.off || c.off < sizeof(ncb_header_t) || c.off % 8 != 0 )
        throw Error( "BadInput", "compiled material: component out of bounds" );
      out.resize( ncb_align16( out.size() ), 0 );
      nh.comp[i].off = out.size();
      if ( c.kind != NCB_KIND_SABVDOS ) {
        append( blob + c.off, c.nbytes );
        nh.comp[i].nbytes = c.nbytes;
        continue;
      }
      if ( c.nbytes < sizeof(ncb_sabvdos_t) ) throw Error( "BadInput", "compiled material: truncated VDOS payload" );
      ncb_sabvdos_t v; std::memcpy( &v, blob + c.off, sizeof(v) );
      if ( v.ndensity > ( (uint64_t)1 << 31 ) || ( c.nbytes - sizeof(v) )/8 < v.ndensity ) throw Error( "BadInput", "compiled material: truncated VDOS payload" );
      if ( v.ndensity < 2 || v.negrid < 10 || v.negrid > 65535 || v.vdoslux > 5 ) throw Error( "BadInput", "compiled material: invalid VDOS leaf" );
      Input in;
      const double* dens = reinterpret_cast<const double*>( blob + c.off + sizeof(v) );
      in.emin = v.emin; in.emax = v.emax; in.density.assign( dens, dens + v.ndensity );
      in.temperature = v.temperature; in.mass_amu = v.mass_amu; in.bound_xs = v.bound_xs;
      const Kernel K = expandFn( in, (unsigned)v.vdoslux, v.target_emax );
      if ( K.ntrimmed )
        warn( "Discarding "+std::to_string( K.ntrimmed )+" edges of provided kernel data due to missing S values." );
      ncb_sab_t h; std::memset( &h, 0, sizeof(h) );
      h.scale = v.scale; h.temperature = K.temperature; h.mass_amu = K.mass_amu; h.bound_xs = K.bound_xs;
      h.suggested_emax = K.suggested_emax;
      h.ext_sigma_free = v.ext_sigma_free; h.ext_ca = v.ext_ca; h.ext_temperature = v.ext_temperature; h.ext_mass_amu = v.ext_mass_amu;
      h.egrid_margin = v.egrid_margin;
      h.negrid = v.negrid; h.nalpha = K.alpha.size(); h.nbeta = K.beta.size(); h.auto_egrid = 1;
      nh.comp[i].kind = NCB_KIND_SAB;
      append( &h, sizeof(h) );
      VectD placeholder( 2*(size_t)v.negrid, 0.0 );
      placeholder[0] = v.req_emin; placeholder[1] = v.req_emax;
      append( placeholder.data(), placeholder.size()*8 );
      append( K.alpha.data(), K.alpha.size()*8 );
      append( K.beta.data(), K.beta.size()*8 );
      append( K.sab.data(), K.sab.size()*8 );
      nh.comp[i].nbytes = out.size() - nh.comp[i].off;
    }
    nh.nbytes = out.size();
    std::memcpy( out.data(), &nh, sizeof(nh) );
    return true;
  }

} }
