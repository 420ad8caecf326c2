// tests/virtapi_caller.cc -- a client of the OpenMC boundary (include/ncrystal_b200_virtapi.hh), following the
// reference's own test of that interface (tests/src/app_vapit1v1/main.cc:30-150): its twelve Al cross sections and
// two oriented Ge cross sections are the reference's golden values.  Prints "key value..." lines.
#include "ncrystal_b200_virtapi.hh"
#include <cmath>
#include <cstdio>
#include <thread>
#include <vector>

int main()
{
  using API = NCrystalVirtualAPI::VirtAPI_Type1_v1;
  auto api = ncrystal_b200::createVirtAPI<API>();
  if ( !api ) { std::printf( "error no_api\n" ); return 1; }
  std::printf( "bad_id_null %d\n", ncrystal_access_virtual_api( 999 ) == nullptr ? 1 : 0 );
  try {
    auto al = api->createScatter( "stdlib::Al_sg225.ncmat;temp=293.15K" );   // prefix as in the reference test
    auto ge = api->createScatter( "Ge_sg227.ncmat;mos=40arcsec;dir1=@crys_hkl:5,1,1@lab:0,0,1;dir2=@crys_hkl:0,-1,1@lab:0,1,0" );
    auto al2 = api->cloneScatter( al );
    auto wl2ekin = []( double wl ) { return 0.081804209605330899 / ( wl*wl ); };
    // app_vapit1v1/main.cc:45-56
    const double ref_al[12] = { 1.39245855, 1.37301271, 1.37152295, 1.29456141, 1.11097728, 1.05893956,
                                1.38597156, 1.76864975, 1.40305622, 0.143286738, 0.148731135, 0.154985793 };
    int nbad = 0;
    for ( int i = 0; i < 12; ++i ) {
      const double n[4] = { wl2ekin( 0.5*( i + 1 ) ), 1.0, 0.0, 0.0 };
      const double xs = api->crossSectionUncached( *( i%2 == 0 ? al : al2 ), n );
      if ( !( std::fabs( xs - ref_al[i] ) <= 0.5e-6*( std::fabs( xs ) + std::fabs( ref_al[i] ) ) + 1e-6 ) ) ++nbad;
    }
    std::printf( "al_xs_mismatches %d\n", nbad );
    {
      const double n1[4] = { wl2ekin( 1.54 ), 0.0, 1.0, 1.0 }, n2[4] = { wl2ekin( 1.54 ), 1.0, 1.0, 0.0 };
      std::printf( "ge_xs %.15g %.15g\n", api->crossSectionUncached( *ge, n1 ), api->crossSectionUncached( *ge, n2 ) );
    }
    // the reference test's generator (main.cc:109-116)
    auto run_ge = [&]( std::vector<double>& out ) {
      unsigned long state = 1789569706;
      std::function<double()> fakerng = [&state]() {
        state = ( 1103515245 * state + 12345 ) % 2147483648;
        return state * ( 1.0 / 2147483648 );
      };
      fakerng(); fakerng(); fakerng();
      double n[4] = { wl2ekin( 1.54 ), 0.0, 1.0, 1.0 };
      for ( int k = 0; k < 4; ++k ) {
        api->sampleScatterUncached( *ge, fakerng, n );
        for ( double v : n ) out.push_back( v );
      }
    };
    {
      // the sampling part of the reference's test on its exact cfg string, printed in its log format
      // (app_vapit1v1/main.cc:109-139; the lines are compared with tests/golden/ref_app_vapit1v1_test.log)
      auto ge_ref = api->createScatter( "stdlib::Ge_sg227.ncmat;dcutoff=0.5;mos=40.0arcsec;dir1=@crys_hkl:5,1,1@lab:0,0,1"
                                        ";dir2=@crys_hkl:0,-1,1@lab:0,1,0" );
      unsigned long state = 1789569706;
      std::function<double()> fakerng = [&state]() {
        state = ( 1103515245 * state + 12345 ) % 2147483648;
        return state * ( 1.0 / 2147483648 );
      };
      fakerng(); fakerng(); fakerng();
      double n[4] = { wl2ekin( 1.54 ), 0.0, 1.0, 1.0 };
      for ( int k = 0; k < 5; ++k ) {
        if ( k ) api->sampleScatterUncached( *ge_ref, fakerng, n );
        std::printf( "reflog Neutron state: (wl=%.5g u=(%.5g, %.5g, %.5g)\n", std::sqrt( 0.081804209605330899 / n[0] ), n[1], n[2], n[3] );
      }
      std::printf( "reflog_rng_state %lu\n", state );
      api->deallocateScatter( ge_ref );
    }
    std::vector<double> a, b;
    run_ge( a ); run_ge( b );
    std::printf( "ge_deterministic %d\n", a == b ? 1 : 0 );
    std::printf( "ge_first %.10g %.6f %.6f %.6f\n", std::sqrt( 0.081804209605330899 / a[0] ), a[1], a[2], a[3] );
    double worst = 0.0;
    for ( size_t k = 0; k < a.size(); k += 4 )
      worst = std::fmax( worst, std::fabs( a[k+1]*a[k+1] + a[k+2]*a[k+2] + a[k+3]*a[k+3] - 1.0 ) );
    std::printf( "ge_norm_dev %.3g\n", worst );
    // Al, thermal: mean scattering cosine over many calls, two client threads on the same ScatterProcess
    auto mean_mu = [&]( unsigned seed, int nn, double* res ) {
      unsigned long state = seed;
      std::function<double()> rng = [&state]() {
        state = ( 6364136223846793005ull * state + 1442695040888963407ull );
        return ( ( state >> 11 ) + 0.5 ) * ( 1.0 / 9007199254740992.0 );
      };
      double sum = 0.0;
      for ( int k = 0; k < nn; ++k ) {
        double n[4] = { 0.0253, 0.0, 0.0, 1.0 };
        api->sampleScatterUncached( *al, rng, n );
        sum += n[3];
      }
      *res = sum / nn;
    };
    double m1 = 0, m2 = 0;
    std::thread t1( mean_mu, 11u, 1500, &m1 ), t2( mean_mu, 22u, 1500, &m2 );
    t1.join(); t2.join();
    std::printf( "al_mean_mu %.6f %.6f\n", m1, m2 );
    api->deallocateScatter( al ); api->deallocateScatter( al2 ); api->deallocateScatter( ge );
    try { api->createScatter( "no_such_material.ncmat" ); std::printf( "bad_cfg_throws 0\n" ); }
    catch ( const std::exception& ) { std::printf( "bad_cfg_throws 1\n" ); }
  } catch ( const std::exception& e ) {
    std::printf( "error %s\n", e.what() );
    return 1;
  }
  return 0;
}
