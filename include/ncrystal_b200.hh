// ncrystal_b200.hh -- header-only C++ mirror of the reference's managed Scatter/Process objects
// (ref: ncrystal_core/include/NCrystal/interfaces/NCProc.hh:56-140) over the C ABI of
// libncrystal_b200.so (ncrystal_b200.h).  Same method names and value semantics; batched overloads
// take std::vector / raw pointers.  Errors follow the C-API convention (call
// ncrystal_sethaltonerror(0) to get exceptions of type NCrystalB200::Error instead of exit(1)).
#ifndef NCRYSTAL_B200_HH
#define NCRYSTAL_B200_HH
#include "ncrystal_b200.h"
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace NCrystalB200 {

  struct Error : public std::runtime_error { using std::runtime_error::runtime_error; };

  inline void checkError()
  {
    if ( ncrystal_error() ) {
      std::string msg = ncrystal_lasterror() ? ncrystal_lasterror() : "";
      std::string typ = ncrystal_lasterrortype() ? ncrystal_lasterrortype() : "";
      ncrystal_clearerror();
      throw Error( typ + ": " + msg );
    }
  }

  struct ScatterOutcomeIsotropic { double ekin, mu; };
  struct ScatterOutcome { double ekin, dir[3]; };

  class Scatter {
  public:
    explicit Scatter( const std::string& cfgstr ) : m_h( ncrystal_create_scatter( cfgstr.c_str() ) ) { checkError(); }
    Scatter( const std::string& cfgstr, unsigned long seed ) : m_h( ncrystal_create_scatter_builtinrng( cfgstr.c_str(), seed ) ) { checkError(); }
    Scatter( const void* blob, size_t nbytes, unsigned long seed ) : m_h( ncb200_create_scatter_from_blob( blob, nbytes, seed ) ) { checkError(); }
    ~Scatter() { if ( m_h.internal ) ncrystal_unref( &m_h ); }
    Scatter( Scatter&& o ) noexcept : m_h( o.m_h ) { o.m_h.internal = nullptr; }
    Scatter& operator=( Scatter&& o ) noexcept { std::swap( m_h, o.m_h ); return *this; }
    Scatter( const Scatter& ) = delete;
    Scatter& operator=( const Scatter& ) = delete;

    Scatter clone() const { Scatter s( ncrystal_clone_scatter( m_h ) ); checkError(); return s; }
    Scatter cloneByIdx( unsigned long idx ) const { Scatter s( ncrystal_clone_scatter_rngbyidx( m_h, idx ) ); checkError(); return s; }
    Scatter cloneForCurrentThread() const { Scatter s( ncrystal_clone_scatter_rngforcurrentthread( m_h ) ); checkError(); return s; }

    const char* name() const { return ncrystal_name( proc() ); }
    bool isOriented() const { return !ncrystal_isnonoriented( proc() ); }
    std::pair<double,double> domain() const { double a, b; ncrystal_domain( proc(), &a, &b ); return { a, b }; }
    // ProcImpl::Process::isNull / getUniqueID (NCProc.hh:86-92): id of the shared immutable process, equal for clones
    bool isNull() const { auto d = domain(); return !( d.first < d.second ); }
    unsigned long long getUniqueID() const
    {
      char* u = ncrystal_process_uid( proc() ); checkError();
      const unsigned long long v = std::stoull( u ); ncrystal_dealloc_string( u );
      return v;
    }

    // single-neutron calls (ref: NCProc.hh:65-66,112-113)
    double crossSectionIsotropic( double ekin ) const { double r; ncrystal_crosssection_nonoriented( proc(), ekin, &r ); checkError(); return r; }
    double crossSection( double ekin, const double (&dir)[3] ) const { double r; ncrystal_crosssection( proc(), ekin, &dir, &r ); checkError(); return r; }
    ScatterOutcomeIsotropic sampleScatterIsotropic( double ekin ) { ScatterOutcomeIsotropic o; ncrystal_samplescatterisotropic( m_h, ekin, &o.ekin, &o.mu ); checkError(); return o; }
    ScatterOutcome sampleScatter( double ekin, const double (&dir)[3] ) { ScatterOutcome o; ncrystal_samplescatter( m_h, ekin, &dir, &o.ekin, &o.dir ); checkError(); return o; }

    // batched calls, host memory (C-API *_many)
    std::vector<double> crossSectionIsotropic( const std::vector<double>& ekin ) const
    {
      std::vector<double> out( ekin.size() );
      ncrystal_crosssection_nonoriented_many( proc(), ekin.data(), ekin.size(), 1, out.data() ); checkError();
      return out;
    }
    void sampleScatterIsotropic( const std::vector<double>& ekin, std::vector<double>& ekin_final, std::vector<double>& mu )
    {
      ekin_final.resize( ekin.size() ); mu.resize( ekin.size() );
      ncrystal_samplescatterisotropic_many( m_h, ekin.data(), ekin.size(), 1, ekin_final.data(), mu.data() ); checkError();
    }
    std::vector<double> crossSection( const std::vector<double>& ekin, const std::vector<double>& ux,
                                      const std::vector<double>& uy, const std::vector<double>& uz ) const
    {
      std::vector<double> out( ekin.size() );
      ncb200_crosssection_many( proc(), ekin.data(), ux.data(), uy.data(), uz.data(), ekin.size(), out.data() ); checkError();
      return out;
    }
    // per-neutron (E, direction) batch, SoA in and out
    void sampleScatter( const std::vector<double>& ekin, const std::vector<double>& ux, const std::vector<double>& uy,
                        const std::vector<double>& uz, std::vector<double>& ekin_final, std::vector<double>& ox,
                        std::vector<double>& oy, std::vector<double>& oz )
    {
      const size_t n = ekin.size();
      ekin_final.resize( n ); ox.resize( n ); oy.resize( n ); oz.resize( n );
      ncb200_samplescatter_manydir( m_h, ekin.data(), ux.data(), uy.data(), uz.data(), n, ekin_final.data(),
                                    ox.data(), oy.data(), oz.data() ); checkError();
    }

    // device-resident batches (caller's stream; asynchronous)
    void crossSectionIsotropicDevice( const double* d_ekin, uint64_t n, double* d_xs, void* stream ) const
    { ncb200_crosssection_nonoriented_many_dev( proc(), d_ekin, n, d_xs, stream ); checkError(); }
    void sampleScatterIsotropicDevice( const double* d_ekin, uint64_t n, double* d_ekin_final, double* d_mu, void* stream )
    { ncb200_samplescatterisotropic_many_dev( m_h, d_ekin, n, d_ekin_final, d_mu, stream ); checkError(); }

    // device-resident transport step (NCrystal::MiniMC "mmc run" query); returns the result JSON
    std::string minimc( const std::string& geomcfg, const std::string& srccfg, const std::string& enginecfg = "" )
    {
      char* js = ncb200_minimc_run( m_h, geomcfg.c_str(), srccfg.c_str(), enginecfg.c_str() );
      checkError();
      std::string out( js ? js : "" );
      if ( js ) ncrystal_dealloc_string( js );
      return out;
    }
    void setRNGStream( uint64_t seed, uint32_t stream_id, uint64_t next_index ) { ncb200_set_rng_stream( m_h, seed, stream_id, next_index ); checkError(); }
    ncrystal_scatter_t handle() const { return m_h; }
  private:
    explicit Scatter( ncrystal_scatter_t h ) : m_h( h ) {}
    ncrystal_process_t proc() const { return ncrystal_cast_scat2proc( m_h ); }
    ncrystal_scatter_t m_h;
  };


  // NCrystal::Absorption (NCProc.hh): here always the 1/v process of the compiled material
  class Absorption {
  public:
    explicit Absorption( const std::string& cfgstr ) : m_h( ncrystal_create_absorption( cfgstr.c_str() ) ) { checkError(); }
    ~Absorption() { if ( m_h.internal ) ncrystal_unref( &m_h ); }
    Absorption( const Absorption& ) = delete;
    Absorption& operator=( const Absorption& ) = delete;
    Absorption( Absorption&& o ) noexcept : m_h( o.m_h ) { o.m_h.internal = nullptr; }
    Absorption clone() const { Absorption a( ncrystal_clone_absorption( m_h ) ); checkError(); return a; }
    double crossSectionIsotropic( double ekin ) const
    { double r; ncrystal_crosssection_nonoriented( ncrystal_cast_abs2proc( m_h ), ekin, &r ); checkError(); return r; }
    std::vector<double> crossSectionIsotropic( const std::vector<double>& ekin ) const
    {
      std::vector<double> out( ekin.size() );
      ncrystal_crosssection_nonoriented_many( ncrystal_cast_abs2proc( m_h ), ekin.data(), ekin.size(), 1, out.data() ); checkError();
      return out;
    }
  private:
    explicit Absorption( ncrystal_absorption_t h ) : m_h( h ) {}
    ncrystal_absorption_t m_h;
  };

}
#endif
