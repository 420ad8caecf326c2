// ncb_kernels.cuh -- sm_100a kernels of the hot path (included by ncb_lib.cu only).
//
//   k_xs_iso        batched isotropic cross sections         (HBM-streaming: 16 B/neutron)
//   k_sample_iso    batched isotropic scatter sampling        (24 B/neutron; +8 with xs)
//   k_sab_*         S(alpha,beta) sampler table builder       (setup, per material)
//   k_gen_source    synthetic log-uniform source              (bench input)
//   k_tally_hist    smem-privatised fp64 histogram            (tally)
//
// Table residency: the small hot lookup tables (PowderBragg 2dE / cumulative
// F^2 table, SAB energy/xs grids) are staged once per CTA into shared memory
// with TMA bulk copies (cp.async.bulk + mbarrier) when they fit; the large
// S(alpha,beta) sampler tables stay in HBM/L2 (tens of MB << 126 MB L2).
#pragma once
#include <cuda_runtime.h>
#include "ncb_proc.cuh"
#include "ncb_sabbuild.cuh"

namespace ncb {

  constexpr int kHotSlots = 2*kMaxPB + 2*kMaxSab;

  // Host-computed staging plan: slot -> (source, bytes, smem offset). bytes==0: not staged.
  struct StagePlan {
    const double* src[kHotSlots];
    uint32_t nbytes[kHotSlots];   // multiple of 16
    uint32_t off[kHotSlots];      // 16-byte aligned offsets into dynamic smem
    uint32_t total;               // dynamic smem bytes
    uint32_t copy_bytes;          // sum of nbytes (the mbarrier's expected transaction count)
  };

  __device__ __forceinline__ uint32_t smemAddr( const void* p )
  {
    return static_cast<uint32_t>( __cvta_generic_to_shared( p ) );
  }

  // One elected thread issues all bulk copies against one mbarrier; everybody waits
  // on phase 0.  (TMA 1-D bulk copy: SASS UBLKCP.)
  __device__ __forceinline__ void stageHotTabs( const Material& M, const StagePlan& sp,
                                                unsigned char* smem, uint64_t* mbar, HotTabs& H )
  {
    hotTabsFromMaterial( M, H );
    if ( sp.total == 0 )
      return;
    const uint32_t mb = smemAddr( mbar );
    if ( threadIdx.x == 0 ) {
      asm volatile( "mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mb) );
      asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
    }
    __syncthreads();
    if ( threadIdx.x == 0 ) {
      asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"(sp.copy_bytes) : "memory" );
      for ( int s = 0; s < kHotSlots; ++s ) {
        if ( !sp.nbytes[s] ) continue;
        asm volatile( "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                      :: "r"( smemAddr( smem + sp.off[s] ) ), "l"( sp.src[s] ), "r"( sp.nbytes[s] ), "r"(mb) : "memory" );
      }
    }
    // wait for phase 0
    uint32_t done = 0;
    while ( !done ) {
      asm volatile( "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                    : "=r"(done) : "r"(mb) : "memory" );
    }
    for ( int s = 0; s < kHotSlots; ++s ) {
      if ( !sp.nbytes[s] ) continue;
      const double* p = reinterpret_cast<const double*>( smem + sp.off[s] );
      if ( s < kMaxPB ) H.pb_e2d[s] = p;
      else if ( s < 2*kMaxPB ) H.pb_fdm[s-kMaxPB] = p;
      else if ( s < 2*kMaxPB+kMaxSab ) H.sab_egrid[s-2*kMaxPB] = p;
      else H.sab_xs[s-2*kMaxPB-kMaxSab] = p;
    }
  }

  // ---------------------------------------------------------------- xs (isotropic)
  // One neutron per thread per grid-stride step.  ekin is read / xs written fully
  // coalesced (8 B per lane).  n_in: length of ekin (outputs index idx use ekin[idx % n_in]
  // only through the host wrapper's repeat handling; here n_in == n).
  __global__ void __launch_bounds__(256)
  k_xs_iso( const __grid_constant__ Material M, const __grid_constant__ StagePlan sp,
            const double* __restrict__ ekin, uint64_t n, double* __restrict__ out )
  {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar;
    HotTabs H;
    stageHotTabs( M, sp, smem, &mbar, H );
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for ( uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride )
      out[i] = matXSIso( M, H, ekin[i], nullptr, nullptr );
  }

  // ------------------------------------------------------- sampling (isotropic), v1
  // All-in-one: xs -> component pick -> leaf sampler, one neutron per thread.
  struct SampleArgs {
    const double* ekin;
    uint64_t n;
    uint64_t seed;
    uint64_t first_index;
    uint32_t sid;
    double* xs_out;      // may be null
    double* ekin_out;
    double* mu_out;
    uint32_t* ndraws;    // may be null
    int32_t* component;  // may be null
    int* err_flags;      // device word, atomicOr'ed
  };

  __global__ void __launch_bounds__(128)
  k_sample_iso( const __grid_constant__ Material M, const __grid_constant__ StagePlan sp,
                const __grid_constant__ SampleArgs A )
  {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar;
    HotTabs H;
    stageHotTabs( M, sp, smem, &mbar, H );
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    int errs = 0;
    for ( uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += stride ) {
      Rng rng; rng.init( A.seed, A.first_index + i, A.sid );
      double eout, mu;
      int err = 0, ich;
      const double xs = matSampleIso( M, H, A.ekin[i], rng, eout, mu, err, ich );
      A.ekin_out[i] = eout;
      A.mu_out[i] = mu;
      if ( A.xs_out ) A.xs_out[i] = xs;
      if ( A.ndraws ) A.ndraws[i] = rng.ndraws;
      if ( A.component ) A.component[i] = ich;
      errs |= err;
    }
    if ( errs )
      atomicOr( A.err_flags, errs );
  }

  // ------------------------------------------------------------ SAB table builder
  __global__ void k_sab_logs( const double* __restrict__ sab, double* __restrict__ logsab, size_t n )
  {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ( i < n ) logsab[i] = sabLogS( sab[i] );
  }

  __global__ void k_sab_cumul( const double* __restrict__ agrid, const double* __restrict__ sab,
                               const double* __restrict__ logsab, int nalpha, int nbeta, double* __restrict__ cumul )
  {
    const int ib = blockIdx.x * blockDim.x + threadIdx.x;
    if ( ib < nbeta )
      sabCumulRow( agrid, sab + (size_t)ib*nalpha, logsab + (size_t)ib*nalpha, nalpha, cumul + (size_t)ib*nalpha );
  }

  // grid: (ceil(nbeta/128), negrid)
  __global__ void k_sab_rows( SabT T, SabRow* __restrict__ rows, SabAlphaInfo* __restrict__ ainfo )
  {
    const int ib = blockIdx.x * blockDim.x + threadIdx.x;
    const int ie = blockIdx.y;
    if ( ib >= T.nbeta ) return;
    const double ekin_div_kT = T.egrid[ie] / T.kT;
    const size_t o = (size_t)ie*T.nbeta + ib;
    SabAlphaInfo info;
    rows[o] = sabAnalyseRow( T.alpha, T.nalpha, T.beta, T.sab, T.logsab, T.cumul, ekin_div_kT, ib, info );
    ainfo[o] = info;
  }

  __global__ void k_sab_epoints( SabT T, const SabRow* __restrict__ rows, SabEPoint* __restrict__ ep,
                                 double* __restrict__ bx, double* __restrict__ bpdf, double* __restrict__ bcdf,
                                 double* __restrict__ xscheck, int* __restrict__ errs )
  {
    const int ie = blockIdx.x * blockDim.x + threadIdx.x;
    if ( ie >= T.negrid ) return;
    const uint32_t off_b = (uint32_t)( (size_t)ie*( T.nbeta+1 ) );
    int err = 0;
    SabEPoint e;
    xscheck[ie] = sabAssembleEPoint( T.beta, T.nbeta, T.kT, T.bound_xs, T.egrid[ie], rows + (size_t)ie*T.nbeta,
                                     off_b, (uint32_t)( (size_t)ie*T.nbeta ), e, bx + off_b, bpdf + off_b, bcdf + off_b, err );
    ep[ie] = e;
    errs[ie] = err;
  }

  // ------------------------------------------------------------ synthetic source
  __global__ void k_gen_source( uint64_t seed, uint64_t first_index, uint64_t n, double loglo, double logspan,
                                double* __restrict__ ekin, double* __restrict__ ux, double* __restrict__ uy, double* __restrict__ uz )
  {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for ( uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride ) {
      Rng rng; rng.init( seed, first_index + i, 0xE0u );
      const double u = rng.generate();
      ekin[i] = exp10( loglo + logspan * u );
      if ( ux ) {
        const double z = 2.0*rng.generate() - 1.0;
        const double phi = 2.0*kPi*rng.generate();
        const double r = sqrt( dmax( 0.0, 1.0 - z*z ) );
        double s, c;
        sincos( phi, &s, &c );
        ux[i] = r*c; uy[i] = r*s; uz[i] = z;
      }
    }
  }

  // ------------------------------------------------------------ tally histogram
  // Bins: [0]=underflow, [1..nbins], [nbins+1]=overflow.  Block-private fp64 histogram in
  // shared memory (atomicAdd.f64 on smem), flushed once per CTA to global.
  __global__ void __launch_bounds__(256)
  k_tally_hist( const double* __restrict__ values, const double* __restrict__ weights, uint64_t n,
                double lo, double invbinw, uint32_t nbins, double* __restrict__ hist, double* __restrict__ sumw2 )
  {
    extern __shared__ __align__(128) unsigned char smem[];
    double* sh = reinterpret_cast<double*>( smem );
    double* sh2 = sh + ( nbins + 2 );
    const uint32_t ntot = nbins + 2;
    for ( uint32_t b = threadIdx.x; b < ntot * ( sumw2 ? 2u : 1u ); b += blockDim.x )
      sh[b] = 0.0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for ( uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride ) {
      const double v = values[i];
      const double w = weights ? weights[i] : 1.0;
      const double rel = ( v - lo ) * invbinw;
      uint32_t b;
      if ( !( rel >= 0.0 ) ) b = 0;
      else if ( rel >= (double)nbins ) b = nbins + 1;
      else b = 1u + (uint32_t)rel;
      atomicAdd( &sh[b], w );
      if ( sumw2 ) atomicAdd( &sh2[b], w*w );
    }
    __syncthreads();
    for ( uint32_t b = threadIdx.x; b < ntot; b += blockDim.x ) {
      if ( sh[b] != 0.0 ) atomicAdd( &hist[b], sh[b] );
      if ( sumw2 && sh2[b] != 0.0 ) atomicAdd( &sumw2[b], sh2[b] );
    }
  }

}
