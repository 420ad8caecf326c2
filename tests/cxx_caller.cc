// tests/cxx_caller.cc -- a C++17 caller of the header-only mirror include/ncrystal_b200.hh (same method names as
// NCrystal::Scatter, ncrystal_core/include/NCrystal/interfaces/NCProc.hh:56-140).  Prints "key value" lines.
#include "ncrystal_b200.hh"
#include <cstdio>
#include <vector>

int main()
{
  namespace NB = NCrystalB200;
  ncrystal_sethaltonerror( 0 );      // errors become NB::Error exceptions instead of exit(1)
  ncrystal_setquietonerror( 1 );
  try {
    NB::Scatter al( "Al_sg225.ncmat;temp=293.15K", 123ul );
    const double ekin = 0.0253;
    std::printf( "al_xs %.17g\n", al.crossSectionIsotropic( ekin ) );
    auto o = al.sampleScatterIsotropic( ekin );
    std::printf( "al_sample_ok %d\n", ( o.ekin >= 0.0 && o.mu >= -1.0 && o.mu <= 1.0 ) ? 1 : 0 );
    std::vector<double> e( 1000 ), ef, mu;
    for ( size_t i = 0; i < e.size(); ++i ) e[i] = 1e-3*( 1 + i );
    auto xs = al.crossSectionIsotropic( e );
    al.sampleScatterIsotropic( e, ef, mu );
    std::printf( "al_batch %zu %zu %zu\n", xs.size(), ef.size(), mu.size() );
    NB::Scatter c = al.clone();
    std::printf( "clone_xs_equal %d\n", c.crossSectionIsotropic( ekin ) == al.crossSectionIsotropic( ekin ) ? 1 : 0 );
    std::printf( "clone_uid_equal %d\n", ( c.getUniqueID() == al.getUniqueID() && !al.isNull() ) ? 1 : 0 );
    {
      std::vector<double> ux( e.size(), 0.0 ), uy( e.size(), 0.0 ), uz( e.size(), 1.0 ), ox, oy, oz;
      al.sampleScatter( e, ux, uy, uz, ef, ox, oy, oz );
      double worst = 0.0;
      for ( size_t i = 0; i < e.size(); ++i ) {
        const double n2 = ox[i]*ox[i] + oy[i]*oy[i] + oz[i]*oz[i] - 1.0;
        worst = n2 < 0 ? ( -n2 > worst ? -n2 : worst ) : ( n2 > worst ? n2 : worst );
      }
      std::printf( "al_dir_batch_ok %d\n", worst < 1e-12 ? 1 : 0 );
    }
    NB::Absorption ab( "Al_sg225.ncmat;temp=293.15K" );
    std::printf( "abs_clone_equal %d\n", ab.clone().crossSectionIsotropic( 0.0253 ) == ab.crossSectionIsotropic( 0.0253 ) ? 1 : 0 );
    std::printf( "al_abs_xs_2200 %.17g\n", ab.crossSectionIsotropic( 0.02529886 ) );
    const std::string js = al.minimc( "sphere;r=0.01", "constant;ekin=0.0253;z=-0.01;n=10000", "tally=mu" );
    std::printf( "minimc_json_ok %d\n", js.find( "NCrystalMiniMCResults_v1" ) != std::string::npos ? 1 : 0 );
    try { NB::Scatter bad( "no_such_material.ncmat" ); std::printf( "bad_cfg_throws 0\n" ); }
    catch ( const NB::Error& ) { std::printf( "bad_cfg_throws 1\n" ); }
  } catch ( const std::exception& e ) {
    std::printf( "error %s\n", e.what() );
    return 1;
  }
  return 0;
}
