// ncb_kernels_lc.cuh -- warp-cooperative kernels for layered crystals (LCBragg, ncb_phys_lcbragg.cuh).
//
// Per neutron the reference rebuilds LCHelper::Cache (NCLCUtils.cc:354-439): a walk over all plane sets with
// 2d >= wavelength (pyrolytic graphite: 129), one or two ROIs for each set that the mosaic band can reach (60 on
// average, up to 250) and for every ROI a Romberg integral over the crystallite rotation of the mosaic cross
// section -- 70-80 us per neutron on one CPU core.  Here ONE WARP handles one neutron:
//   1. lanes stride the plane sets; the ROIs they find are compacted IN THE REFERENCE'S ORDER (plane set, normal
//      before anti-normal) into the warp's shared-memory list with shuffle prefix sums;
//   2. lanes stride that list and integrate their ROIs in parallel;
//   3. the running sum (m_roixs_commul) is formed in list order.
//   k_lc_scan            per neutron: sum over the ROIs + their number (what crossSection and the component pick need);
//                        in a sampling call also the ROI that pickRandIdxByWeight over the cumulative values would
//                        choose with the neutron's own uniform number
//   k_lc_sample_threads  per neutron whose chosen component is LCBragg (one per thread): the scattering in the
//                        recorded ROI (overlay rejection sampling of phi, GaussMos::genScat, rotation to the lab frame)
#pragma once
#include "ncb_kernels_sc.cuh"

namespace ncb {

  constexpr int kLcWarps = 4;   // warps per CTA
  inline
#if defined(__CUDACC__)
  __host__ __device__
#endif
  uint32_t lcRoiCap( int nplanes ) { return (uint32_t)( ( 2*nplanes + 3 ) & ~3 ); }
  inline uint32_t lcSmemBytes( int nplanes ) { return kLcWarps * lcRoiCap( nplanes ) * (uint32_t)( sizeof(LcRoi) + sizeof(double) ); }

  // Steps 1-3 for the neutron owned by this warp.  Returns the number of ROIs; rois[] / commul[] hold the list and
  // the cumulative cross sections.
  __device__ __forceinline__ int lcBuildWarp( const ScBraggT& S, const LcBraggT& L, const LcNeutron& N,
                                              LcRoi* rois, double* commul, int& err )
  {
    const int lane = threadIdx.x & 31;
    __syncwarp();   // the lists are reused from the previous neutron of this warp
    int nroi = 0;
    for ( int p0 = 0; p0 < L.nplanes; p0 += 32 ) {
      const int ips = p0 + lane;
      LcRoi r[2];
      int nr = 0;
      bool act = false;
      if ( ips < L.nplanes ) {
        const double* P = L.planes + kLcPlaneStride*ips;
        act = !( N.wl > P[0] );
        if ( act )
          nr = lcFindROIs( P, ips, N, S.cta, S.sta, r );
      }
      // exclusive prefix sum of nr over the lanes
      int incl = nr;
      #pragma unroll
      for ( int d = 1; d < 32; d <<= 1 ) {
        const int v = __shfl_up_sync( 0xffffffffu, incl, d );
        if ( lane >= d ) incl += v;
      }
      const int base = nroi + incl - nr;
      for ( int k = 0; k < nr; ++k ) rois[base+k] = r[k];
      nroi += __shfl_sync( 0xffffffffu, incl, 31 );
      // plane sets are sorted by d-spacing: once a lane sees wl > 2d no later set can contribute
      if ( __ballot_sync( 0xffffffffu, ( ips < L.nplanes ) && !act ) )
        break;
    }
    __syncwarp();
    for ( int k = lane; k < nroi; k += 32 )
      commul[k] = lcRoiXS( S, L, N, rois[k], err );
    __syncwarp();
    if ( lane == 0 ) {
      double sum = 0.0;
      for ( int k = 0; k < nroi; ++k ) commul[k] = ( sum += commul[k] );
    }
    __syncwarp();
    return nroi;
  }

  struct LcScanArgs {
    const double* ekin; const double* ux; const double* uy; const double* uz;
    uint64_t n;
    double* lc_sum;    // out: m_roixs_commul.back() (before the 1/(V0*natoms) factor), 0 without ROIs
    int32_t* lc_n;     // out: number of ROIs
    double dom_lo, dom_hi;
    int* err_flags;
    const uint32_t* n_dev = nullptr;
    // sampling calls: the scan also makes the choice a scattering on this leaf would make among its rotation ranges
    // (pickRandIdxByWeight with the neutron's uniform number `pick_draw` of its stream) and stores the chosen range,
    // so that the sampling pass does not have to build the list a second time (k_lc_sample_threads); null = off
    LcRoi* pick_roi = nullptr;
    uint64_t seed = 0, first_index = 0; uint32_t sid = 0, pick_draw = 0;
    const uint64_t* ids = nullptr;
  };

  __global__ void __launch_bounds__(32*kLcWarps)
  k_lc_scan( const __grid_constant__ Material M, const __grid_constant__ LcScanArgs A )
  {
    extern __shared__ __align__(128) unsigned char smem[];
    const ScBraggT& S = M.sc;
    const LcBraggT& L = M.lc;
    const uint32_t cap = lcRoiCap( L.nplanes );
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    LcRoi* rois = reinterpret_cast<LcRoi*>( smem ) + (size_t)w*cap;
    double* commul = reinterpret_cast<double*>( smem + (size_t)kLcWarps*cap*sizeof(LcRoi) ) + (size_t)w*cap;
    const uint64_t nwarps = (uint64_t)gridDim.x * kLcWarps;
    const uint64_t ntot = A.n_dev ? (uint64_t)min( (uint64_t)*A.n_dev, A.n ) : A.n;
    int err = 0;
    for ( uint64_t i = (uint64_t)blockIdx.x * kLcWarps + w; i < ntot; i += nwarps ) {
      const double ekin = A.ekin[i];
      double sum = 0.0; int nroi = 0;
      if ( domainContains( A.dom_lo, A.dom_hi, ekin ) && !( ekin < L.ekin_low ) ) {
        const Vec3 u = vunit( Vec3{ A.ux[i], A.uy[i], A.uz[i] } );
        LcNeutron N;
        if ( lcNeutronPars( L, ekin, u, N ) ) {
          nroi = lcBuildWarp( S, L, N, rois, commul, err );
          if ( nroi ) sum = commul[nroi-1];
        }
      }
      if ( A.pick_roi ) {
        LcRoi chosen; chosen.rotmin = chosen.rotmax = 0.0; chosen.ips = -1; chosen.sign = 1;
        if ( nroi > 0 && sum ) {
          int idx = 0;
          if ( nroi > 1 ) {
            Rng rng; rng.init( A.seed, A.ids ? A.ids[i] : A.first_index + i, A.sid );
            rng.seek( A.pick_draw );
            idx = pickIdxByWeight( rng.generate(), commul, nroi );
          }
          chosen = rois[idx];
        }
        if ( lane == 0 ) A.pick_roi[i] = chosen;
      }
      if ( lane == 0 ) { A.lc_sum[i] = sum; A.lc_n[i] = nroi; }
    }
    if ( err && A.err_flags )
      atomicOr( A.err_flags, err );
  }

  // One queued neutron per THREAD, for sampling calls whose scan recorded the chosen rotation range: the scattering
  // itself (overlay sampling of the crystallite rotation, GaussMos::genScat, rotation to the lab frame) is
  // thread-level work; building the list of ranges a second time, one warp per neutron (r2 first version), cost half
  // as much again as the cross-section scan.
  __global__ void __launch_bounds__(128)
  k_lc_sample_threads( const __grid_constant__ Material M, const __grid_constant__ SampleArgs A, const __grid_constant__ AnisoArgs X,
                       const LcRoi* __restrict__ pick_roi )
  {
    const ScBraggT& S = M.sc;
    const LcBraggT& L = M.lc;
    const uint32_t nq = *X.q_sc_count;
    const uint32_t stride = gridDim.x * blockDim.x;
    for ( uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nq; j += stride ) {
      const uint32_t i = X.q_sc[j] & kQueueIdxMask;
      const double ekin = A.ekin[i];
      const Vec3 u = vunit( Vec3{ X.D.ux[i], X.D.uy[i], X.D.uz[i] } );
      Vec3 o = u;
      const LcRoi roi = pick_roi[i];
      Rng rng; rng.init( A.seed, A.streamIndex( i ), A.sid );
      // (the pick among the ranges consumed one number when there was more than one range)
      rng.seek( ( M.ncomp > 1 ? 1u : 0u ) + ( roi.ips >= 0 && X.sc_n[i] > 1 ? 1u : 0u ) );
      LcNeutron N;
      if ( roi.ips >= 0 && lcNeutronPars( L, ekin, u, N ) )
        lcGenScatterRoi( S, L, N, roi, u, rng, o );
      A.ekin_out[i] = ekin;
      X.D.ox[i] = o.x; X.D.oy[i] = o.y; X.D.oz[i] = o.z;
      if ( A.ndraws ) A.ndraws[i] = rng.ndraws;
    }
  }


}
