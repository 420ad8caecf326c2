// ncb_kernels_mmc.cuh -- kernels of the device-resident transport step (see ncb_mmc.cuh for the physics and the
// reference line numbers).  Included by ncb_lib.cu only.
//
// Population layout: SoA arrays in HBM (MmcState), two buffers.  One step =
//   launchXSAniso(A)            scattering cross section at every live neutron's (E, dir)
//   k_mmc_forward(A -> B)       flight distances, transmission, roulette, weights; survivors are appended to B
//                               at their scattering point (block-aggregated atomic cursor), wt[] = tallied weight
//   k_mmc_tally(A, wt)          exit tallies, shared-memory privatised per CTA, one histogram at a time
//   launchSampleAniso(B)        scattering outcome for every survivor (same kernels as the C-API)
//   k_mmc_post(B)               new (E, dir), scattering counters
// then A <-> B.  All per-neutron accesses are coalesced 8-byte streams; the only scattered traffic is the
// compaction write (contiguous per CTA).
#pragma once
#include "ncb_mmc.cuh"

namespace ncb {

  struct MmcState {
    double *x, *y, *z, *ux, *uy, *uz, *w, *ekin, *e0;
    double *ux0, *uy0, *uz0;   // initial directions; null when the source direction is fixed
    int32_t *nscat, *ninel;
    uint64_t* id;
  };

  // CTA-wide stream compaction cursor: returns the output slot of every thread with pred (undefined otherwise).
  // One global atomic per CTA per call.  Must be called by all threads of the CTA.
  __device__ __forceinline__ uint32_t blockAppend( bool pred, uint32_t* counter, uint32_t* s_warp, uint32_t* s_base )
  {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t m = __ballot_sync( 0xffffffffu, pred );
    if ( lane == 0 ) s_warp[w] = __popc( m );
    __syncthreads();
    if ( threadIdx.x == 0 ) {
      uint32_t tot = 0;
      for ( int k = 0; k < nw; ++k ) { const uint32_t c = s_warp[k]; s_warp[k] = tot; tot += c; }
      *s_base = tot ? atomicAdd( counter, tot ) : 0u;
    }
    __syncthreads();
    const uint32_t pos = *s_base + s_warp[w] + __popc( m & ( ( 1u << lane ) - 1u ) );
    __syncthreads();
    return pos;
  }

  // counters: [0] live neutrons appended to A, [1] missed neutrons appended to Miss
  __global__ void __launch_bounds__(256)
  k_mmc_source( const __grid_constant__ MmcSource S, const __grid_constant__ MmcGeom G, uint64_t seed, uint64_t first_id,
                uint32_t n, MmcState A, MmcState Miss, uint32_t* counters, int record_miss )
  {
    __shared__ uint32_t s_warp[8], s_base;
    const uint32_t n_up = ( n + blockDim.x - 1 )/blockDim.x*blockDim.x;
    for ( uint32_t i = blockIdx.x*blockDim.x + threadIdx.x; i < n_up; i += gridDim.x*blockDim.x ) {
      bool live = false, miss = false;
      MmcNeutron nt{};
      if ( i < n ) {
        Rng rng; rng.init( seed, first_id + i, kMmcSidSrc );
        nt = mmcGenerate( S, rng );
        live = true;
        if ( S.may_be_outside ) {
          const double d = mmcDistToEntry( G, nt.x, nt.y, nt.z, nt.ux, nt.uy, nt.uz );
          if ( d < 0.0 ) { live = false; miss = true; }
          else { nt.x += d*nt.ux; nt.y += d*nt.uy; nt.z += d*nt.uz; }   // detail::propagateDistance
        }
      }
      const uint32_t pa = blockAppend( live, counters + 0, s_warp, &s_base );
      const uint32_t pm = blockAppend( miss, counters + 1, s_warp, &s_base );
      if ( live || ( miss && record_miss ) ) {
        const MmcState& T = live ? A : Miss;
        const uint32_t p = live ? pa : pm;
        T.x[p] = nt.x; T.y[p] = nt.y; T.z[p] = nt.z; T.ux[p] = nt.ux; T.uy[p] = nt.uy; T.uz[p] = nt.uz;
        T.w[p] = nt.w; T.ekin[p] = nt.ekin; T.e0[p] = nt.ekin;
        if ( T.ux0 ) { T.ux0[p] = nt.ux; T.uy0[p] = nt.uy; T.uz0[p] = nt.uz; }
        T.nscat[p] = live ? 0 : -1;       // markAsMissedTarget, NCMMC_Baskets.cc:81
        T.ninel[p] = 0;
        T.id[p] = first_id + i;
      }
    }
  }

  __global__ void __launch_bounds__(256)
  k_mmc_forward( const __grid_constant__ MmcGeom G, const __grid_constant__ MmcEngine E, uint64_t seed, uint32_t step,
                 uint32_t n, MmcState A, const double* __restrict__ xs, double* __restrict__ wt,
                 MmcState B, uint32_t* counter, const uint32_t* __restrict__ n_dev, uint32_t* __restrict__ n_log )
  {
    __shared__ uint32_t s_warp[8], s_base;
    if ( n_dev ) n = min( *n_dev, n );      // population count kept on the device (several steps per host sync)
    if ( n_log && blockIdx.x == 0 && threadIdx.x == 0 ) *n_log = n;
    const uint32_t n_up = ( n + blockDim.x - 1 )/blockDim.x*blockDim.x;
    for ( uint32_t i = blockIdx.x*blockDim.x + threadIdx.x; i < n_up; i += gridDim.x*blockDim.x ) {
      MmcStepOut o; o.survives = false;
      double ux = 0, uy = 0, uz = 0, ekin = 0, e0 = 0, ux0 = 0, uy0 = 0, uz0 = 0; int nscat = 0, ninel = 0; uint64_t id = 0;
      if ( i < n ) {
        id = A.id[i];
        if ( A.ux0 ) { ux0 = A.ux0[i]; uy0 = A.uy0[i]; uz0 = A.uz0[i]; }
        ux = A.ux[i]; uy = A.uy[i]; uz = A.uz[i]; ekin = A.ekin[i]; e0 = A.e0[i];
        nscat = A.nscat[i]; ninel = A.ninel[i];
        Rng rng; rng.init( seed, id, kMmcSidBase + 2u*step );
        o = mmcForward( G, E, rng, A.x[i], A.y[i], A.z[i], ux, uy, uz, A.w[i], ekin, nscat, xs[i] );
        wt[i] = o.wt;
      }
      const uint32_t p = blockAppend( o.survives, counter, s_warp, &s_base );
      if ( o.survives ) {
        B.x[p] = o.x; B.y[p] = o.y; B.z[p] = o.z; B.ux[p] = ux; B.uy[p] = uy; B.uz[p] = uz;
        B.w[p] = o.w; B.ekin[p] = ekin; B.e0[p] = e0; B.nscat[p] = nscat; B.ninel[p] = ninel; B.id[p] = id;
        if ( B.ux0 ) { B.ux0[p] = ux0; B.uy0[p] = uy0; B.uz0[p] = uz0; }
      }
    }
  }

  // after the scattering kernels: SimEngine.cc:392-404
  __global__ void __launch_bounds__(256)
  k_mmc_post( uint32_t n, MmcState B, const double* __restrict__ eout, const double* __restrict__ ox,
              const double* __restrict__ oy, const double* __restrict__ oz, const uint32_t* __restrict__ n_dev )
  {
    if ( n_dev ) n = min( *n_dev, n );
    for ( uint32_t i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x ) {
      const double e_new = eout[i];
      const bool was_elastic = ( B.ekin[i] == e_new );
      B.ux[i] = ox[i]; B.uy[i] = oy[i]; B.uz[i] = oz[i];
      B.ekin[i] = e_new;
      B.nscat[i] += 1;
      if ( !was_elastic ) B.ninel[i] += 1;
    }
  }

  // order-preserving map double -> uint64 (for atomicMin/atomicMax on the running min/max of filled values)
  __device__ __forceinline__ unsigned long long mmcOrdered( double d )
  {
    const unsigned long long b = (unsigned long long)__double_as_longlong( d );
    return ( b >> 63 ) ? ~b : ( b | 0x8000000000000000ull );
  }

  // Exit tallies.  wt == nullptr: the records are source neutrons that missed the volume (weight = S.w).
  // dynamic smem: mmcHistDoubles(max nbins) doubles.
  __global__ void __launch_bounds__(256)
  k_mmc_tally( const __grid_constant__ MmcTally T, uint32_t n, MmcState A, const double* __restrict__ wt,
               double* __restrict__ tally, const uint32_t* __restrict__ n_dev )
  {
    extern __shared__ __align__(16) double sh[];
    if ( n_dev ) n = min( *n_dev, n );
    const int lane = threadIdx.x & 31;
    const uint32_t n_up = ( n + 31u ) & ~31u;
    for ( int ih = 0; ih < T.nh; ++ih ) {
      const MmcHist h = T.h[ih];
      const int nb2 = h.nbins + 2;
      const uint32_t nd = mmcHistDoubles( h.nbins );
      double* s_cont = sh;                                  // [class][nb2]
      double* s_err = sh + kMmcNClass*nb2;                  // [class][nb2]
      double* s_stat = sh + 2*kMmcNClass*nb2;               // [class][kMmcNStat]
      unsigned long long* s_statu = reinterpret_cast<unsigned long long*>( s_stat );
      for ( uint32_t k = threadIdx.x; k < nd; k += blockDim.x ) sh[k] = 0.0;
      __syncthreads();
      if ( threadIdx.x < kMmcNClass ) {
        s_statu[threadIdx.x*kMmcNStat + 3] = ~0ull;         // min
        s_statu[threadIdx.x*kMmcNStat + 4] = 0ull;          // max
      }
      __syncthreads();
      for ( uint32_t i = blockIdx.x*blockDim.x + threadIdx.x; i < n_up; i += gridDim.x*blockDim.x ) {
        int cls = -1, key = -1;
        double val = 0.0, wgt = 0.0;
        if ( i < n ) {
          const double w = wt ? wt[i] : A.w[i];
          const int nscat = A.nscat[i];
          bool weighted;
          double ux0 = 0, uy0 = 0, uz0 = 0;
          if ( A.ux0 ) { ux0 = A.ux0[i]; uy0 = A.uy0[i]; uz0 = A.uz0[i]; }
          val = mmcTallyValue( T, h.type, A.ux[i], A.uy[i], A.uz[i], A.ekin[i], w, nscat, A.e0[i], ux0, uy0, uz0, weighted );
          wgt = weighted ? w : 1.0;
          if ( wgt > 0.0 ) {
            cls = mmcClass( nscat, A.ninel[i] );
            key = cls*nb2 + mmcValueToBin( h, val );
          }
        }
        // Lanes that fill the same (class, bin) combine first: one shared-memory atomic per distinct bin and warp.
        // (Unscattered neutrons of a pencil beam all land in ONE bin: without this the first step of a run spent
        //  1.1 ms per 4 Mi neutrons in 32-way serialised atomics.)
        {
          const unsigned peers = __match_any_sync( 0xffffffffu, key );
          const int maxcnt = __reduce_max_sync( 0xffffffffu, __popc( peers ) );
          unsigned m = peers;
          double s1 = 0.0, s2 = 0.0;
          const double w1 = key >= 0 ? wgt : 0.0, w2 = w1*w1;
          for ( int it = 0; it < maxcnt; ++it ) {
            const int j = m ? __ffs( m ) - 1 : lane;
            const double a = __shfl_sync( 0xffffffffu, w1, j ), b = __shfl_sync( 0xffffffffu, w2, j );
            if ( m ) { s1 += a; s2 += b; m &= m - 1u; }
          }
          if ( key >= 0 && lane == __ffs( peers ) - 1 ) {
            atomicAdd( &s_cont[key], s1 );
            atomicAdd( &s_err[key], s2 );
          }
        }
        // running statistics (RunningStats1D): warp-reduce per class, one shared atomic per warp and class
        for ( int c = 0; c < kMmcNClass; ++c ) {
          const uint32_t m = __ballot_sync( 0xffffffffu, cls == c );
          if ( !m ) continue;
          const bool mine = ( cls == c );
          double a0 = mine ? wgt : 0.0, a1 = mine ? wgt*val : 0.0, a2 = mine ? wgt*val*val : 0.0;
          unsigned long long lo = mine ? mmcOrdered( val ) : ~0ull, hi = mine ? mmcOrdered( val ) : 0ull;
          for ( int d = 16; d; d >>= 1 ) {
            a0 += __shfl_xor_sync( 0xffffffffu, a0, d );
            a1 += __shfl_xor_sync( 0xffffffffu, a1, d );
            a2 += __shfl_xor_sync( 0xffffffffu, a2, d );
            const unsigned long long l2 = __shfl_xor_sync( 0xffffffffu, lo, d ), h2 = __shfl_xor_sync( 0xffffffffu, hi, d );
            lo = l2 < lo ? l2 : lo; hi = h2 > hi ? h2 : hi;
          }
          if ( lane == 0 ) {
            atomicAdd( &s_stat[c*kMmcNStat + 0], a0 );
            atomicAdd( &s_stat[c*kMmcNStat + 1], a1 );
            atomicAdd( &s_stat[c*kMmcNStat + 2], a2 );
            atomicMin( &s_statu[c*kMmcNStat + 3], lo );
            atomicMax( &s_statu[c*kMmcNStat + 4], hi );
          }
        }
      }
      __syncthreads();
      double* g = tally + h.off;
      unsigned long long* gu = reinterpret_cast<unsigned long long*>( g );
      const uint32_t nbins_all = 2u*kMmcNClass*nb2;
      for ( uint32_t k = threadIdx.x; k < nbins_all; k += blockDim.x )
        if ( sh[k] != 0.0 ) atomicAdd( &g[k], sh[k] );
      if ( threadIdx.x < kMmcNClass ) {
        const uint32_t so = threadIdx.x*kMmcNStat;     // within s_stat / s_statu
        const uint32_t o = nbins_all + so;             // within the histogram's global block
        if ( s_statu[so+3] != ~0ull ) {
          atomicAdd( &g[o+0], s_stat[so+0] ); atomicAdd( &g[o+1], s_stat[so+1] ); atomicAdd( &g[o+2], s_stat[so+2] );
          atomicMin( &gu[o+3], s_statu[so+3] );
          atomicMax( &gu[o+4], s_statu[so+4] );
        }
      }
      __syncthreads();
    }
  }

  // sum of the tallied weights and record count (metadata "tallied"): block reduction
  __global__ void __launch_bounds__(256)
  k_mmc_sum_weights( uint32_t n, const double* __restrict__ w, double* __restrict__ out, const uint32_t* __restrict__ n_dev )
  {
    __shared__ double s[8];
    if ( n_dev ) n = min( *n_dev, n );
    double a = 0.0;
    for ( uint32_t i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x ) a += w[i];
    for ( int d = 16; d; d >>= 1 ) a += __shfl_xor_sync( 0xffffffffu, a, d );
    if ( ( threadIdx.x & 31 ) == 0 ) s[threadIdx.x >> 5] = a;
    __syncthreads();
    if ( threadIdx.x == 0 ) {
      double t = 0.0;
      for ( int k = 0; k < (int)( blockDim.x >> 5 ); ++k ) t += s[k];
      atomicAdd( out, t );
    }
  }

}

// ---------------------------------------------------------------------------------------------------------------
// Tail of a transport run.  Once few neutrons are left (<= 32 Ki) a step is a dozen launches on nearly empty grids --
// and a neutron caught between Bragg reflections of a single crystal can live for hundreds of steps (Ge sphere at
// 3.2 Aa: 851 steps, 11,000 launches, 0.1 ms each).  k_mmc_tail finishes such a population in ONE launch: one warp
// per history, looping over its steps until the history ends.  Per step it does what the launch sequence above does
// for that neutron, with the same per-(source id, step) random streams -- hence the same tallies:
//   cross section   isotropic leaves by every lane redundantly (same instruction stream, no extra issue cost),
//                   the single-crystal leaf by the warp-cooperative walk of k_sc_scan (scWalkWarp);
//   forward step    mmcForward;  exit tallies with global fp64 atomics, one lane per histogram;
//   scattering      component pick; single crystal: second walk (mode 1) + GaussMos::genScat as in k_sc_sample;
//                   other leaves: the thread-level samplers + randDirectionGivenScatterMu (ncb_proc.cuh: matSample);
//   bookkeeping     k_mmc_post.
// dynamic smem: [staged tables (sp.total)] [fam_of: nnormals bytes] [kMmcTailWarps x ScWarpScratch]
constexpr int kMmcTailWarps = 8;
struct MmcTailOut {
  unsigned long long records;   // tally records made (= live neutrons summed over the steps)
  unsigned int last_step;       // highest step index + 1 reached by any history
  unsigned int pad;
};

namespace ncb {

  // s_stat: the CTA's running statistics [histogram][class][kMmcNStat] in shared memory -- every record of a class
  // updates the same five words, so with global atomics all warps of the grid serialised on a handful of L2
  // addresses (Ge tail: 17 ms); the bins themselves are spread and stay global.
  __device__ __forceinline__ void mmcTallyAtomic( const MmcTally& T, double* __restrict__ tally, double* s_stat, int lane,
                                                  double ux, double uy, double uz, double ekin, double wt, int nscat, int ninel,
                                                  double e0, double ux0, double uy0, double uz0 )
  {
    if ( lane >= T.nh ) return;
    const MmcHist h = T.h[lane];
    bool weighted;
    const double val = mmcTallyValue( T, h.type, ux, uy, uz, ekin, wt, nscat, e0, ux0, uy0, uz0, weighted );
    const double wgt = weighted ? wt : 1.0;
    if ( !( wgt > 0.0 ) ) return;
    const int nb2 = h.nbins + 2;
    const int cls = mmcClass( nscat, ninel );
    const int key = cls*nb2 + mmcValueToBin( h, val );
    double* g = tally + h.off;
    atomicAdd( &g[key], wgt );
    atomicAdd( &g[kMmcNClass*nb2 + key], wgt*wgt );
    double* st = s_stat + ( lane*kMmcNClass + cls )*kMmcNStat;
    unsigned long long* stu = reinterpret_cast<unsigned long long*>( st );
    atomicAdd( &st[0], wgt ); atomicAdd( &st[1], wgt*val ); atomicAdd( &st[2], wgt*val*val );
    atomicMin( &stu[3], mmcOrdered( val ) );
    atomicMax( &stu[4], mmcOrdered( val ) );
  }

  template <int kMinBlocks>
  __global__ void __launch_bounds__(32*kMmcTailWarps, kMinBlocks)
  k_mmc_tail( const __grid_constant__ Material M, const __grid_constant__ StagePlan sp,
              const __grid_constant__ MmcGeom G, const __grid_constant__ MmcEngine E, const __grid_constant__ MmcTally T,
              uint64_t seed, uint32_t step0, uint32_t max_steps, uint32_t n, MmcState A,
              double* __restrict__ tally, double* __restrict__ meta, MmcTailOut* __restrict__ out, int* __restrict__ err_flags,
              uint32_t fam_of_off, uint32_t scratch_off, uint32_t* __restrict__ next_history )
  {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar;
    __shared__ double s_stat[TALLY_NTYPES*kMmcNClass*kMmcNStat];
    {
      unsigned long long* su = reinterpret_cast<unsigned long long*>( s_stat );
      for ( int k = threadIdx.x; k < TALLY_NTYPES*kMmcNClass*kMmcNStat; k += blockDim.x ) {
        const int stat = k % kMmcNStat;
        if ( stat == 3 ) su[k] = ~0ull; else if ( stat == 4 ) su[k] = 0ull; else s_stat[k] = 0.0;
      }
    }
    __syncthreads();
    HotTabs H;
    uint8_t* fam_of = smem + fam_of_off;
    int isc = -1;
    for ( int c = 0; c < M.ncomp; ++c ) if ( M.comp[c].kind == KIND_SCBRAGG ) isc = c;
    if ( isc >= 0 ) scBlockSetup( M, sp, smem, &mbar, H, fam_of );
    else stageHotTabs( M, sp, smem, &mbar, H );
    const ScBraggT& S = *H.sc;
    ScWarpScratch& ws = reinterpret_cast<ScWarpScratch*>( smem + scratch_off )[ threadIdx.x >> 5 ];
    const int lane = threadIdx.x & 31;
    unsigned long long records = 0;
    unsigned int last_step = 0;
    double sumw = 0.0;
    int errs = 0;
    // histories differ in length by three orders of magnitude: warps fetch them one at a time from a global counter
    while ( true ) {
      uint32_t ih = 0;
      if ( lane == 0 ) ih = atomicAdd( next_history, 1u );
      ih = __shfl_sync( 0xffffffffu, ih, 0 );
      if ( ih >= n ) break;
      double x = A.x[ih], y = A.y[ih], z = A.z[ih], ux = A.ux[ih], uy = A.uy[ih], uz = A.uz[ih];
      double w = A.w[ih], ekin = A.ekin[ih];
      const double e0 = A.e0[ih];
      double ux0 = 0, uy0 = 0, uz0 = 0;
      if ( A.ux0 ) { ux0 = A.ux0[ih]; uy0 = A.uy0[ih]; uz0 = A.uz0[ih]; }
      int nscat = A.nscat[ih], ninel = A.ninel[ih];
      const uint64_t id = A.id[ih];
      uint32_t step = step0;
      double iso_ekin = -1.0, iso_xs[kMaxComp]; int iso_aux[kMaxComp];
      while ( true ) {
        if ( step - step0 >= max_steps ) { errs |= ERR_MMC_NOTERM; break; }
        // ---- cross section at (E, dir): launchXSAniso
        double sc_xs = 0.0, sc_wl = 0.0; int sc_n = 0;
        const Vec3 dir = { ux, uy, uz };
        Vec3 dnorm = dir;
        if ( isc >= 0 && domainContains( M.comp[isc].dom_lo, M.comp[isc].dom_hi, ekin ) && !( ekin <= S.threshold_ekin ) ) {
          vnormalise( dnorm );
          ScAccum acc; double wl;
          scWalkWarp( S, ws, fam_of, ekin, dnorm, wl, acc, 2, false, 0.0 );     // (mode 2: the entries are recorded)
          sc_xs = acc.commul_last; sc_n = acc.n; sc_wl = wl;
        }
        // composition sum (matXSPre).  The isotropic leaves depend on the energy alone, and a neutron caught between
        // Bragg reflections keeps its energy for hundreds of steps: their values are kept while ekin does not change
        // (same values, same summation order).
        double cumul[kMaxComp]; int aux[kMaxComp];
        double xs = 0.0;
        if ( domainContains( M.dom_lo, M.dom_hi, ekin ) ) {
          const bool fresh = !( ekin == iso_ekin );
          for ( int i = 0; i < M.ncomp; ++i ) {
            const Comp& c = M.comp[i];
            int a = -1;
            double v = 0.0;
            if ( domainContains( c.dom_lo, c.dom_hi, ekin ) ) {
              if ( c.kind == KIND_SCBRAGG ) { v = sc_xs; a = sc_n; }
              else if ( c.kind == KIND_LCBRAGG ) { v = sc_n ? M.lc.xsfact * sc_xs : 0.0; a = sc_n; }
              else {
                if ( fresh ) { iso_xs[i] = compXSIso( M, H, i, ekin, a ); iso_aux[i] = a; }
                v = iso_xs[i]; a = iso_aux[i];
              }
            }
            xs += c.scale * v;
            cumul[i] = xs; aux[i] = a;
          }
          iso_ekin = ekin;
        }
        // ---- forward step + exit tallies: k_mmc_forward, k_mmc_tally, k_mmc_sum_weights
        Rng rng; rng.init( seed, id, kMmcSidBase + 2u*step );
        const MmcStepOut o = mmcForward( G, E, rng, x, y, z, ux, uy, uz, w, ekin, nscat, xs );
        mmcTallyAtomic( T, tally, s_stat, lane, ux, uy, uz, ekin, o.wt, nscat, ninel, e0, ux0, uy0, uz0 );
        sumw += o.wt; ++records;
        ++step;
        if ( !o.survives ) break;
        x = o.x; y = o.y; z = o.z; w = o.w;
        // ---- scattering: launchSampleAniso on the stream (id, step) with the odd stream id
        double eout = ekin; Vec3 od = dir;
        if ( domainContains( M.dom_lo, M.dom_hi, ekin ) ) {
          Rng r2; r2.init( seed, id, kMmcSidBase + 2u*( step - 1u ) + 1u );
          const int ich = ( M.ncomp == 1 ? 0 : pickIdxByWeight( r2.generate(), cumul, M.ncomp ) );
          const Comp& c = M.comp[ich];
          if ( c.kind == KIND_SCBRAGG ) {
            if ( !( ekin <= M.sc.threshold_ekin ) && aux[ich] > 0 && sc_xs > 0.0 ) {
              double choice = -1.0; bool linear = true;
              if ( sc_n > 1 ) { choice = sc_xs * r2.generate(); linear = ( sc_n < 5 ); }
              // the plane is selected among the entries the cross-section walk of this step recorded (same rule as
              // the second walk of k_sc_sample: linear search '>' below five entries, lower_bound '>=' above); only a
              // neutron with more contributing planes than the record holds is walked again
              int in, sgn;
              double wl = sc_wl;
              if ( sc_n <= kScRecCap ) {
                __syncwarp();
                int k = 0;
                for ( ; k < sc_n - 1; ++k ) {
                  const double cum = ws.rec_cumul[k];
                  if ( linear ? ( cum > choice ) : !( cum < choice ) ) break;
                }
                const int e = ws.rec_entry[k];
                in = e & 0x7fff; sgn = e >> 15;
              } else {
                ScAccum acc;
                scWalkWarp( S, ws, fam_of, ekin, dnorm, wl, acc, 1, linear, choice );
                in = acc.chosen_in; sgn = acc.chosen_sign;
              }
              const double sg = sgn ? 1.0 : -1.0;
              const Vec3 pn = { sg*S.normals[3*in], sg*S.normals[3*in+1], sg*S.normals[3*in+2] };
              const double inv2dsp = gmCacheRound( S.fam_inv2d[ fam_of[in] ] );
              gmGenScat( S, r2, pn, inv2dsp, wl, dnorm, od );
            }
          } else {
            double mu = 1.0; int err = 0;
            compSampleIso( M, H, ich, aux[ich], ekin, r2, eout, mu, err );
            errs |= err;
            if ( !( err & ( ERR_SAB_LOOP_INNER | ERR_SAB_LOOP_OUTER | ERR_SAB_DISCARD | ERR_KIN_DENOM ) ) )
              od = randDirectionGivenScatterMu( r2, mu, dir );
            else { od = { 0.0, 0.0, 0.0 }; eout = -1.0; }
          }
        }
        if ( !( eout >= 0.0 ) ) break;     // a sampler raised an error: the flags end the run on the host
        // ---- k_mmc_post
        const bool was_elastic = ( ekin == eout );
        ux = od.x; uy = od.y; uz = od.z; ekin = eout;
        ++nscat;
        if ( !was_elastic ) ++ninel;
      }
      if ( step > last_step ) last_step = step;
    }
    if ( lane == 0 ) {
      if ( records ) { atomicAdd( &out->records, records ); atomicAdd( &meta[0], sumw ); }
      atomicMax( &out->last_step, last_step );
    }
    if ( errs ) atomicOr( err_flags, errs );
    // the CTA's running statistics -> global (one set of atomics per histogram and class that was filled)
    __syncthreads();
    for ( int k = threadIdx.x; k < T.nh*kMmcNClass; k += blockDim.x ) {
      const int ih = k / kMmcNClass, cls = k % kMmcNClass;
      const double* st = s_stat + k*kMmcNStat;
      const unsigned long long* stu = reinterpret_cast<const unsigned long long*>( st );
      if ( stu[3] == ~0ull ) continue;
      const MmcHist h = T.h[ih];
      double* g = tally + h.off + 2*kMmcNClass*( h.nbins + 2 ) + cls*kMmcNStat;
      unsigned long long* gu = reinterpret_cast<unsigned long long*>( g );
      atomicAdd( &g[0], st[0] ); atomicAdd( &g[1], st[1] ); atomicAdd( &g[2], st[2] );
      atomicMin( &gu[3], stu[3] );
      atomicMax( &gu[4], stu[4] );
    }
  }

}
