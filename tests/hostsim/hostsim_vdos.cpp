// tests/hostsim/hostsim_vdos.cpp -- TEST-ONLY host build of the VDOS -> S(alpha,beta) expansion: the host
// orchestration of ncrystal_b200/csrc/ncb_vdos.h over a backend that runs the NCB_HD functions of
// ncb_vdos_dev.cuh (butterflies, interpolation, cell sums) in plain loops.  Never loaded by the package.
#include "ncb_vdos.h"
#include <cstring>
#include <string>

namespace {
  using namespace ncb::vdos;
  thread_local std::string g_verr;

  class HostBackend {
  public:
    void setSpectrum( unsigned order, const VectD& s ) { at( order ) = s; }
    VectD spectrum( unsigned order ) { return at( order ); }
    void convolve( const std::vector<ConvJob>& jobs, std::vector<ConvResult>& res )
    {
      res.assign( jobs.size(), ConvResult() );
      for ( size_t j = 0; j < jobs.size(); ++j ) one( jobs[j], res[j] );
    }
    void fill( const FillPlan& P, VectD& sab )
    {
      const size_t na = P.nalpha;
      sab.assign( na*P.nbeta, 0.0 );
      std::vector<GnDev> gn( P.norders );
      for ( unsigned n = 1; n <= P.norders; ++n ) {
        const GnMeta& m = P.meta[n-1];
        gn[n-1] = GnDev{ at( n ).data(), (unsigned long long)m.n, m.lower, m.upper, 1.0/m.binwidth };
      }
      FillDev F;
      F.gn = gn.data(); F.scale = P.scale.data(); F.afact = P.alpha_factor.data(); F.a_first = P.a_first.data(); F.a_end = P.a_end.data();
      F.skip = P.skip.data(); F.beta_nonpos = P.beta_nonpos.data(); F.expbeta = P.expbeta.data(); F.sab = sab.data();
      F.nalpha = (unsigned)na; F.idx_zero = (unsigned)P.idx_zero; F.idx_firstflip = (unsigned)P.idx_firstflip; F.order0 = 1; F.kT = P.kT;
      for ( size_t ib = 0; ib < P.beta_nonpos.size(); ++ib ) {
        const double beta = P.beta_nonpos[ib], energy = beta*P.kT;
        const bool flip = ib >= P.idx_firstflip && beta != 0.0;
        const double expMbeta = flip ? P.expbeta[ib] : 0.0;
        const size_t mirror = P.idx_zero + ( P.idx_zero - ib );
        for ( unsigned ia = 0; ia < na; ++ia )
          for ( auto& g : P.job_orders ) {
            double acc, accp;
            fillGroup( F, ia, energy, expMbeta, (unsigned)g.first, (unsigned)g.second, acc, accp );
            sab[ib*na+ia] += acc;
            if ( expMbeta ) sab[mirror*na+ia] += accp;
          }
      }
    }
  private:
    std::vector<VectD> m_spec;
    std::vector<Cplx> m_w; int m_wlog = -1;
    VectD& at( unsigned order ) { if ( m_spec.size() < order ) m_spec.resize( order ); return m_spec[order-1]; }

    void fft( std::vector<Cplx>& d, int logn, bool inverse )
    {
      const unsigned wsize = 1u << m_wlog;
      for ( int i = 0; i < logn; ++i )
        for ( unsigned t = 0; t < ( 1u << ( logn - 1 ) ); ++t ) {
          unsigned lo, off;
          butterflyIndex( t, i, lo, off );
          const Cplx w = m_w[(size_t)off*( wsize >> ( i + 1 ) )];
          butterfly( d[lo], d[lo + ( 1u << i )], w.re, inverse ? -w.im : w.im );
        }
    }
    void one( const ConvJob& J, ConvResult& R )
    {
      const size_t nout = J.n1 + J.n2 - 1;
      int logn = 0;
      while ( ( (size_t)1 << logn ) < nout ) ++logn;
      if ( logn > m_wlog ) { m_w = twiddles( (unsigned)logn ); m_wlog = logn; }
      const size_t N = (size_t)1 << logn;
      const VectD& a1 = at( J.o1 ); const VectD& a2 = at( J.o2 );
      std::vector<Cplx> b1( N ), b2( N ), bo( N );
      for ( unsigned j = 0; j < N; ++j ) {
        const unsigned r = bitReverse( j, logn );
        b1[j] = Cplx{ r < J.n1 ? a1[(size_t)r*J.stride1] : 0.0, 0.0 };
        b2[j] = Cplx{ r < J.n2 ? a2[(size_t)r*J.stride2] : 0.0, 0.0 };
      }
      fft( b1, logn, false ); fft( b2, logn, false );
      for ( unsigned j = 0; j < N; ++j ) { const unsigned r = bitReverse( j, logn ); bo[j] = cmul( b1[r], b2[r] ); }
      fft( bo, logn, true );
      const double k = J.dt/(double)N;
      VectD y( nout );
      for ( size_t i = 0; i < nout; ++i ) { double v = bo[i].re*bo[i].re + bo[i].im*bo[i].im; v = std::sqrt( v ); v *= k; y[i] = v; }
      double dt = J.dt;
      R.ifront = 0; R.extra_thin = 1;
      if ( J.trunc_thin ) {
        if ( J.trunc_threshold > 0 ) {
          const double cutoff = J.trunc_threshold * *std::max_element( y.begin(), y.end() );
          size_t ifront = 0, iback = y.size() - 1;
          for ( ; ifront < iback; ++ifront ) if ( y[ifront] > cutoff ) break;
          for ( ; iback > ifront; --iback ) if ( y[iback] > cutoff ) break;
          if ( iback > ifront ) y = VectD( y.begin() + ifront, y.begin() + iback + 1 );
          R.ifront = ifront;
        }
        if ( J.thin_nbins > 0 && y.size() > J.thin_nbins ) {
          unsigned long extra = 1;
          while ( y.size() > J.thin_nbins*extra ) extra *= 2;
          if ( extra >= 8 && J.gentle_thinning ) extra /= 2;
          VectD t;
          for ( size_t i = 0; i < y.size(); i += extra ) t.push_back( y[i] );
          y.swap( t );
          dt *= extra;
          R.extra_thin = extra;
        }
      }
      double area = 0.;
      for ( double v : y ) area += v;
      area *= dt;
      const double inv = 1.0/area;
      for ( double& v : y ) v *= inv;
      R.n = y.size();
      R.maxval = *std::max_element( y.begin(), y.end() );
      const double thr = J.relthr*R.maxval;
      R.first_above = R.last_above = -1;
      for ( size_t i = 0; i < y.size(); ++i ) if ( y[i] >= thr ) { R.first_above = (long)i; break; }
      for ( size_t i = y.size(); i > 0; --i ) if ( y[i-1] >= thr ) { R.last_above = (long)( i-1 ); break; }
      at( J.order ) = y;
    }
  };

  Input makeInput( const double* egrid, unsigned negrid, const double* density, unsigned ndensity, double temperature, double mass_amu, double bound_xs )
  {
    Input in;
    regularise( VectD( egrid, egrid + negrid ), VectD( density, density + ndensity ), in.emin, in.emax, in.density );
    in.temperature = temperature; in.mass_amu = mass_amu; in.bound_xs = bound_xs;
    return in;
  }
}

bool hostsim_expand_vdos_leaves( const void* blob, size_t nbytes, std::vector<unsigned char>& out )
{
  return rewriteVdosLeaves( blob, nbytes, out,
    []( const Input& in, unsigned vdoslux, double target_emax ) { HostBackend be; return expand( in, vdoslux, target_emax, be ); },
    []( const std::string& ) {} );
}

extern "C" {

  const char* hostsim_vdos_lasterror() { return g_verr.c_str(); }

  // meta = { suggestedEmax, gamma0, msd, max order, number of trimmed edges }
  int hostsim_vdos_expand( const double* egrid, unsigned negrid, const double* density, unsigned ndensity,
                           double temperature, double mass_amu, double bound_xs, unsigned vdoslux, double target_emax,
                           double* alpha, int* nalpha, double* beta, int* nbeta, double* sab, int cap_sab, double* meta5 )
  {
    try {
      HostBackend be;
      const Kernel K = expand( makeInput( egrid, negrid, density, ndensity, temperature, mass_amu, bound_xs ), vdoslux, target_emax, be );
      if ( (int)K.sab.size() > cap_sab ) return -2;
      *nalpha = (int)K.alpha.size(); *nbeta = (int)K.beta.size();
      std::memcpy( alpha, K.alpha.data(), K.alpha.size()*8 );
      std::memcpy( beta, K.beta.data(), K.beta.size()*8 );
      std::memcpy( sab, K.sab.data(), K.sab.size()*8 );
      meta5[0] = K.suggested_emax; meta5[1] = K.gamma0; meta5[2] = K.msd; meta5[3] = K.max_order; meta5[4] = K.ntrimmed;
      return 0;
    } catch ( std::exception& e ) { g_verr = e.what(); return -3; }
  }

  // meta = { lower edge, upper edge, bin width, max density }
  int hostsim_vdos_gn( const double* egrid, unsigned negrid, const double* density, unsigned ndensity,
                       double temperature, double mass_amu, double bound_xs, int order, double* spec, int cap, double* meta4 )
  {
    try {
      HostBackend be;
      Eval ev( makeInput( egrid, negrid, density, ndensity, temperature, mass_amu, bound_xs ) );
      Ladder<HostBackend> Gn( ev, ev.calcGamma0(), be, TruncThin(), 1e-9 );
      Gn.grow( (unsigned)order, 0 );
      const VectD s = be.spectrum( (unsigned)order );
      if ( (int)s.size() > cap ) return -2;
      std::memcpy( spec, s.data(), s.size()*8 );
      const GnMeta& m = Gn.meta( (unsigned)order );
      meta4[0] = m.lower; meta4[1] = m.upper; meta4[2] = m.binwidth; meta4[3] = m.maxval;
      return (int)s.size();
    } catch ( std::exception& e ) { g_verr = e.what(); return -3; }
  }

  // twiddle table built by doubling == table built by the reference's recipe (number of differing entries)
  int hostsim_vdos_twiddle_check( int log2size )
  {
    const std::vector<Cplx>& a = twiddles( (unsigned)log2size );
    const std::vector<Cplx> b = makeTwiddles( (unsigned)log2size );
    int bad = a.size() == b.size() ? 0 : 1;
    for ( size_t i = 0; i < a.size() && i < b.size(); ++i ) if ( a[i].re != b[i].re || a[i].im != b[i].im ) ++bad;
    return bad;
  }

  // regulariseVDOSGrid alone
  int hostsim_vdos_regularise( const double* egrid, unsigned negrid, const double* density, unsigned ndensity, double* emin_emax, double* out, int cap )
  {
    try {
      VectD d;
      regularise( VectD( egrid, egrid + negrid ), VectD( density, density + ndensity ), emin_emax[0], emin_emax[1], d );
      if ( (int)d.size() > cap ) return -2;
      std::memcpy( out, d.data(), d.size()*8 );
      return (int)d.size();
    } catch ( std::exception& e ) { g_verr = e.what(); return -3; }
  }
}
