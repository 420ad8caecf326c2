// ncb_replay.cu -- one neutron, sampled with the CALLER's random numbers.
//
// ncrystal_samplescatter_rs (ncrystal.h:792, src/cinterface/ncrystal.cc:1176-1196) and the OpenMC virtual API
// (sampleScatterUncached, src/virtualapi/NCVirtAPI_Type1_v1_impl.hh:64-92) hand the library a generator that the
// reference consumes draw by draw (the reference's own tests pin the consumption: tests/src/app_crng/test.log,
// app_vapit1v1/test.log).  This translation unit compiles the same per-neutron device functions as the batch
// kernels (ncb_proc.cuh) against the replay form of Rng (ncb_rng.cuh, NCB_RNG_REPLAY): the kernel takes the numbers
// drawn so far as a kernel argument and reports whether it ran past them.  The host side (ncb_lib_virtapi.inc) draws
// lazily: k numbers, run; "needs more" -> draw number k+1, run again.  Consumption is a deterministic function of
// the numbers, so the caller's generator advances exactly as under the reference, one launch per number consumed.
// (Everything is compiled into namespace ncb_rs so that the two Rng definitions never meet in one namespace.)
#define NCB_RNG_REPLAY 1
#define ncb ncb_rs
#include <cuda_runtime.h>
#include "ncb_proc.cuh"
#include "ncb_replay.h"
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace ncb {

  double g_erfc_lut_host[kErfcLutLen];

  struct ReplayNumbers { double u[kReplayMaxNumbers]; };

  __global__ void __launch_bounds__(32)
  k_sample_one_replay( const __grid_constant__ Material M, const __grid_constant__ ReplayNumbers U, uint32_t nu,
                       double ekin, double dx, double dy, double dz, double* out, uint32_t* info )
  {
    if ( threadIdx.x ) return;
    HotTabs H;
    hotTabsFromMaterial( M, H );
    Rng rng; rng.init( 0, 0 );
    rng.u = U.u; rng.nu = nu;
    double eo = ekin;
    Vec3 o = { dx, dy, dz };
    int err = 0, ich = -1;
    matSample( M, H, ekin, Vec3{ dx, dy, dz }, rng, eo, o, err, ich );
    out[0] = eo; out[1] = o.x; out[2] = o.y; out[3] = o.z;
    info[0] = rng.ndraws; info[1] = rng.overrun; info[2] = (uint32_t)err;
    __threadfence_system();
  }

  namespace {
    struct DevScratch { int device; double* h_out; uint32_t* h_info; double* d_out; uint32_t* d_info; };
    std::mutex g_mtx;
    std::vector<DevScratch> g_scratch;

    void check( cudaError_t e, const char* what )
    {
      if ( e != cudaSuccess )
        throw std::runtime_error( std::string("CUDA error in ")+what+": "+cudaGetErrorString(e) );
    }

    DevScratch& scratchFor( int device )
    {
      for ( auto& s : g_scratch ) if ( s.device == device ) return s;
      static bool lut_done = false;
      if ( !lut_done ) { fillErfcLutHost( g_erfc_lut_host ); lut_done = true; }
      check( cudaMemcpyToSymbol( g_erfc_lut_dev, g_erfc_lut_host, sizeof(g_erfc_lut_host) ), "erfc table upload" );
      DevScratch s; s.device = device;
      check( cudaHostAlloc( (void**)&s.h_out, 4*sizeof(double), cudaHostAllocMapped ), "cudaHostAlloc" );
      check( cudaHostAlloc( (void**)&s.h_info, 4*sizeof(uint32_t), cudaHostAllocMapped ), "cudaHostAlloc" );
      check( cudaHostGetDevicePointer( (void**)&s.d_out, s.h_out, 0 ), "cudaHostGetDevicePointer" );
      check( cudaHostGetDevicePointer( (void**)&s.d_info, s.h_info, 0 ), "cudaHostGetDevicePointer" );
      g_scratch.push_back( s );
      return g_scratch.back();
    }
  }
}

// One run of the neutron with the first `nu` numbers of `u`.  `material` is the device-resident ncb::Material of the
// calling translation unit (same layout: same headers).  The current device must be `device`.
void ncb_replay_sample_one( const void* material, size_t material_bytes, int device, const double* u, uint32_t nu,
                            double ekin, const double* dir, double* out4, uint32_t* ndraws, int* overrun, int* err,
                            void* stream )
{
  using namespace ncb;
  if ( material_bytes != sizeof(Material) )
    throw std::runtime_error( "ncb_replay: Material layout mismatch between translation units" );
  if ( nu > (uint32_t)kReplayMaxNumbers )
    throw std::runtime_error( "ncb_replay: too many random numbers for one scattering" );
  Material M;
  std::memcpy( &M, material, sizeof(M) );
  ReplayNumbers U;
  std::memcpy( U.u, u, nu*sizeof(double) );
  std::lock_guard<std::mutex> g( g_mtx );
  DevScratch& s = scratchFor( device );
  cudaStream_t st = static_cast<cudaStream_t>( stream );
  k_sample_one_replay<<< 1, 32, 0, st >>>( M, U, nu, ekin, dir[0], dir[1], dir[2], s.d_out, s.d_info );
  check( cudaGetLastError(), "k_sample_one_replay launch" );
  check( cudaStreamSynchronize( st ), "k_sample_one_replay" );
  for ( int k = 0; k < 4; ++k ) out4[k] = s.h_out[k];
  *ndraws = s.h_info[0]; *overrun = (int)s.h_info[1]; *err = (int)s.h_info[2];
}
