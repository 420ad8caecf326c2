// Development aid (not a test, not part of the library): host-side stage times of one VDOS -> S(alpha,beta) expansion
// on the device.  Built on the GPU box with the timing macro of ncb_vdos.h switched on:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -DNCB_VDOS_TIMING -Incrystal_b200/csrc \
//        tests/tools/vdos_stage_times.cu ncrystal_b200/csrc/ncb_vdos.cu -o /tmp/vdos_stage_times
//   /tmp/vdos_stage_times curve.bin vdoslux      (curve.bin: emin, emax, sigma, mass, T, density[...] as raw doubles)
#include "ncb_vdos_api.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
int main( int argc, char** argv )
{
  if ( argc < 3 ) return 2;
  FILE* f = std::fopen( argv[1], "rb" );
  if ( !f ) return 3;
  std::vector<double> d; double x;
  while ( std::fread( &x, 8, 1, f ) == 1 ) d.push_back( x );
  std::fclose( f );
  ncb::vdos::Input in;
  in.emin = d[0]; in.emax = d[1]; in.bound_xs = d[2]; in.mass_amu = d[3]; in.temperature = d[4];
  in.density.assign( d.begin() + 5, d.end() );
  for ( int rep = 0; rep < 3; ++rep ) {
    unsigned launches = 0;
    const auto t0 = std::chrono::steady_clock::now();
    ncb::vdos::Kernel K = ncb::vdos::expandOnDevice( in, (unsigned)std::atoi( argv[2] ), 0.0, nullptr, &launches );
    const double ms = std::chrono::duration<double,std::milli>( std::chrono::steady_clock::now() - t0 ).count();
    std::fprintf( stderr, "== rep %d: %.3f ms, %u launches, %zux%zu, order %u\n", rep, ms, launches, K.alpha.size(), K.beta.size(), K.max_order );
  }
  return 0;
}
