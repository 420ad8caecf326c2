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

  // slots: PowderBragg 2dE | cumulative tables | SAB energy grids | xs grids | PowderBragg key luts | SAB key luts
  constexpr int kHotSlotsIso = 3*kMaxPB + 3*kMaxSab;
  // + SCBragg: demi-normals, family xsfact / inv2d / first-index arrays, the two spline LUTs
  constexpr int kHotSlots = kHotSlotsIso + 6;

  // Host-computed staging plan: slot -> (source, bytes, smem offset). bytes==0: not staged.
  struct StagePlan {
    const void* src[kHotSlots];
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
    for ( int s = 0; s < kHotSlotsIso; ++s ) {
      if ( !sp.nbytes[s] ) continue;
      const double* p = reinterpret_cast<const double*>( smem + sp.off[s] );
      if ( s < kMaxPB ) H.pb_e2d[s] = p;
      else if ( s < 2*kMaxPB ) H.pb_fdm[s-kMaxPB] = p;
      else if ( s < 2*kMaxPB+kMaxSab ) H.sab_egrid[s-2*kMaxPB] = p;
      else if ( s < 2*kMaxPB+2*kMaxSab ) H.sab_xs[s-2*kMaxPB-kMaxSab] = p;
      else if ( s < 3*kMaxPB+2*kMaxSab ) H.pb_lut[s-2*kMaxPB-2*kMaxSab] = reinterpret_cast<const uint16_t*>( p );
      else H.sab_elut[s-3*kMaxPB-2*kMaxSab] = reinterpret_cast<const uint16_t*>( p );
    }
    if ( sp.nbytes[kHotSlotsIso] ) {   // SCBragg tables are staged all-or-nothing
      H.scv = M.sc;
      H.scv.normals      = reinterpret_cast<const double*>( smem + sp.off[kHotSlotsIso+0] );
      H.scv.fam_xsfact   = reinterpret_cast<const double*>( smem + sp.off[kHotSlotsIso+1] );
      H.scv.fam_inv2d    = reinterpret_cast<const double*>( smem + sp.off[kHotSlotsIso+2] );
      H.scv.fam_first    = reinterpret_cast<const int*>( smem + sp.off[kHotSlotsIso+3] );
      H.scv.sofcosd.data = reinterpret_cast<const double*>( smem + sp.off[kHotSlotsIso+4] );
      H.scv.evalcosx.data= reinterpret_cast<const double*>( smem + sp.off[kHotSlotsIso+5] );
      H.sc = &H.scv;
    }
  }

  // ---------------------------------------------------------------- xs (isotropic)
  // One neutron per thread per grid-stride step.  ekin is read / xs written fully
  // coalesced (8 B per lane).  n_in: length of ekin (outputs index idx use ekin[idx % n_in]
  // only through the host wrapper's repeat handling; here n_in == n).
  // (register caps measured on B200, r2: 5 CTAs/SM (48 registers) is the best trade for the cross-section kernel --
  // Al 4.25e10 -> 4.44e10 xs/s against 4 CTAs, 3.9e10 at 6 --, 8 CTAs/SM (32 registers, 80 B of spills) for the
  // classify kernel)
  __global__ void __launch_bounds__(256, 5)
  k_xs_iso( const __grid_constant__ Material M, const __grid_constant__ StagePlan sp,
            const double* __restrict__ ekin, uint64_t n, double* __restrict__ out )
  {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar;
    HotTabs H;
    stageHotTabs( M, sp, smem, &mbar, H );
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double e = i < n ? ldStream( ekin + i ) : 0.0;
    while ( i < n ) {
      // the next step's energy is requested before this step's arithmetic (the load latency was the top stall)
      const uint64_t inext = i + stride;
      const double enext = inext < n ? ldStream( ekin + inext ) : 0.0;
      stStream( out + i, matXSIso( M, H, e, nullptr, nullptr ) );
      e = enext; i = inext;
    }
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
    const uint64_t* ids = nullptr;   // optional: random-stream index of neutron i (transport: source particle id);
                                     // default first_index + i
    NCB_HD uint64_t streamIndex( uint64_t i ) const { return ids ? ids[i] : first_index + i; }
    // optional: the number of neutrons lives on the device (<= n); used by the transport step to run several steps
    // without a host round trip (grids are sized for n, the kernels stop at *n_dev)
    const uint32_t* n_dev = nullptr;
    __device__ __forceinline__ uint64_t count() const { return n_dev ? (uint64_t)min( (uint64_t)*n_dev, n ) : n; }
  };

  // ------------------------------------------------------- sampling (isotropic)
  // A batch is sampled by a sequence of launches instead of one divergent kernel:
  //   k_sample_classify  xs + component pick; the cheap elastic leaves (PowderBragg, ElInc) are sampled in place;
  //                      S(alpha,beta) table neutrons and free-gas neutrons (FreeGas leaf, SAB above Emax) are
  //                      appended to two index queues
  //   S(alpha,beta) table queue: k_sample_sab_refill (attempt-level lane refill, short-chain table lookups)
  //   free-gas queue:    k_fg_* staged pipeline / k_sample_fg
  // Each queue holds homogeneous work, so warps do not serialise over unrelated code paths.  Because the random
  // streams are counter based, a later kernel resumes a neutron's stream by re-deriving one block (no RNG state
  // is stored).
  constexpr uint32_t kQueueIdxBits = 28;
  constexpr uint32_t kQueueIdxMask = ( 1u << kQueueIdxBits ) - 1u;   // <= 2^28 neutrons per launch

  constexpr int kSortBins = 1024;   // size of the scratch area behind the queue counters (class totals, cursors)
  struct QueueArgs {
    uint32_t* q_sab;    // S(alpha,beta) table path, E < Emax
    uint32_t* q_fg;     // free-gas leaf and S(alpha,beta) above Emax
    uint32_t* q_emax;   // pairs (entry, draws consumed): table sampling at E=Emax requested by the high-E analysis
    uint32_t* counts;   // [0] = #q_sab, [1] = #q_fg, [2] = #q_emax (pairs), [3],[4] refill cursors
    // elastic leaves (isotropic sampling): neutrons whose chosen component is PowderBragg / ElIncScatter, sampled by
    // k_sample_elastic; null = sampled in place by k_sample_classify
    uint32_t* q_pb = nullptr;  uint32_t* q_el = nullptr;
    uint32_t* counts_el = nullptr;   // [0] = #q_pb, [1] = #q_el
  };

  // Monotonic energy bin: exponent + top 4 mantissa bits of the double (16 bins per octave).
  __device__ __forceinline__ uint32_t sortKey( double ekin, int shift = 0 )
  {
    const int k = (int)( ( (unsigned long long)__double_as_longlong( ekin ) >> 48 ) & 0x7FFFull ) - ( 999 << 4 );
    return ( (uint32_t)( k < 0 ? 0 : ( k > kSortBins-1 ? kSortBins-1 : k ) ) >> shift ) << shift;
  }

  __device__ __forceinline__ void warpPush( bool pred, uint32_t* q, uint32_t* counter, uint32_t entry )
  {
    const uint32_t mask = __ballot_sync( 0xffffffffu, pred );
    if ( !mask ) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs( mask ) - 1;
    uint32_t base = 0;
    if ( lane == leader )
      base = atomicAdd( counter, (uint32_t)__popc( mask ) );
    base = __shfl_sync( 0xffffffffu, base, leader );
    if ( pred )
      q[ base + __popc( mask & ( ( 1u << lane ) - 1u ) ) ] = entry;
  }

  // Queue pushes of k_sample_classify: every warp collects its entries in a private shared-memory buffer and
  // writes them out in runs (one global atomic and a coalesced copy per ~100 entries).  No CTA barrier on the
  // path (r1: three __syncthreads per CTA iteration, `barrier` was the top stall reason of the kernel).
  constexpr int kWarpBuf = 128;            // entries per warp and queue; flushed when fewer than 32 slots are left
  template <int kQueues>
  struct WarpQueueBuf {
    uint32_t e[kQueues][kWarpBuf];     // S(alpha,beta) table, free gas [, PowderBragg, ElIncScatter]
  };
  __device__ __forceinline__ void warpBufFlush( const uint32_t* buf, uint32_t n, uint32_t* q, uint32_t* counter )
  {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    uint32_t base = 0;
    if ( lane == 0 ) base = atomicAdd( counter, n );
    base = __shfl_sync( 0xffffffffu, base, 0 );
    for ( uint32_t k = lane; k < n; k += 32 )
      q[base + k] = buf[k];
    __syncwarp();
  }

  template <bool kDeferElastic>
  __global__ void __launch_bounds__(256, 8)
  k_sample_classify( const __grid_constant__ Material M, const __grid_constant__ StagePlan sp,
                     const __grid_constant__ SampleArgs A, const __grid_constant__ QueueArgs Q )
  {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar;
    __shared__ WarpQueueBuf<kDeferElastic ? 4 : 2> s_wq[8];
    HotTabs H;
    stageHotTabs( M, sp, smem, &mbar, H );
    const int lane = threadIdx.x & 31;
    auto& wq = s_wq[threadIdx.x >> 5];
    uint32_t c1 = 0, c2 = 0, c3 = 0, c4 = 0; // entries in the warp's buffers (warp-uniform)
    constexpr bool defer_elastic = kDeferElastic;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n = A.n;
    double enext = ( (uint64_t)blockIdx.x * blockDim.x + threadIdx.x < n ) ? ldStream( A.ekin + ( (uint64_t)blockIdx.x * blockDim.x + threadIdx.x ) ) : 0.0;
    for ( uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < n; base += stride ) {
      const uint64_t i = base + threadIdx.x;
      int cls = 0;            // 0: done here, 1: SAB table queue, 2: free-gas queue, 3: PowderBragg queue, 4: ElInc queue
      uint32_t entry = 0;
      const double ekin = enext;
      enext = ( i + stride < n ) ? ldStream( A.ekin + i + stride ) : 0.0;     // requested one step ahead
      if ( i < n ) {
        double eout = ekin, mu = 1.0, tot = 0.0;
        int ich = -1;
        uint32_t nd = 0;
        if ( domainContains( M.dom_lo, M.dom_hi, ekin ) ) {
          double cumul[kMaxComp];
          int aux[kMaxComp];
          tot = matXSIso( M, H, ekin, cumul, aux );
          Rng rng; rng.init( A.seed, A.streamIndex( i ), A.sid );
          ich = ( M.ncomp == 1 ? 0 : pickIdxByWeight( rng.generate(), cumul, M.ncomp ) );
          const Comp& c = M.comp[ich];
          if ( c.kind == KIND_SAB ) {
            const SabT& T = M.sab[c.idx];
            // upper_bound(egrid,E)==end  <=>  !(E < egrid.back())  (NCSABSampler.cc:166-171)
            int iu = aux[ich];
            if ( iu < 0 ) iu = upperBound( H.sab_egrid[c.idx], 0, T.negrid, ekin );   // (E outside the leaf's domain)
            cls = ( iu < T.negrid ) ? 1 : 2;
          } else if ( c.kind == KIND_FREEGAS ) {
            cls = 2;
          } else if ( c.kind == KIND_POWDERBRAGG ) {
            // PowderBragg::sampleScatterIsotropic, ref: NCPowderBragg.cc:202-216
            // The elastic leaves are sampled by k_sample_elastic over their own queues (r2): in place, every warp
            // walked the PowderBragg path, the ElInc path and the queue pushes one after the other for a third of
            // its lanes (ncu: 19.9 of 32 lanes active).  A neutron below the Bragg threshold needs no sampling.
            const PowderBraggT& T = M.pb[c.idx];
            if ( !( ekin < T.threshold || !isFinite(ekin) ) ) {
              if ( defer_elastic ) cls = 3;
              else {
                const int iv = aux[ich] >= 0 ? aux[ich] : pbLastValidPlane( T, H.pb_e2d[c.idx], H.pb_lut[c.idx], ekin );
                mu = pbSampleMu( H.pb_e2d[c.idx], H.pb_fdm[c.idx], iv, ekin, rng );
              }
            }
            nd = rng.ndraws;
          } else if ( c.kind == KIND_ELINC ) {
            if ( defer_elastic ) cls = 4;
            else {
              mu = elincSampleMu( M.elinc[c.idx], ekin, rng );
              nd = rng.ndraws;
            }
          }
          entry = (uint32_t)i | ( (uint32_t)ich << kQueueIdxBits );
        }
        if ( A.xs_out ) stStream( A.xs_out + i, tot );
        if ( A.component ) A.component[i] = ich;
        if ( cls == 0 ) {
          A.ekin_out[i] = eout;
          A.mu_out[i] = mu;
          if ( A.ndraws ) A.ndraws[i] = nd;
        }
      }
      const uint32_t m1 = __ballot_sync( 0xffffffffu, cls == 1 );
      const uint32_t m2 = __ballot_sync( 0xffffffffu, cls == 2 );
      const uint32_t lt = ( 1u << lane ) - 1u;
      if ( cls == 1 ) wq.e[0][ c1 + __popc( m1 & lt ) ] = entry;
      if ( cls == 2 ) wq.e[1][ c2 + __popc( m2 & lt ) ] = entry;
      c1 += __popc( m1 ); c2 += __popc( m2 );
      if ( c1 > (uint32_t)( kWarpBuf - 32 ) ) { warpBufFlush( wq.e[0], c1, Q.q_sab, Q.counts + 0 ); c1 = 0; }
      if ( c2 > (uint32_t)( kWarpBuf - 32 ) ) { warpBufFlush( wq.e[1], c2, Q.q_fg, Q.counts + 1 ); c2 = 0; }
      if constexpr ( kDeferElastic ) {
        const uint32_t m3 = __ballot_sync( 0xffffffffu, cls == 3 );
        const uint32_t m4 = __ballot_sync( 0xffffffffu, cls == 4 );
        if ( cls == 3 ) wq.e[2][ c3 + __popc( m3 & lt ) ] = entry;
        if ( cls == 4 ) wq.e[3][ c4 + __popc( m4 & lt ) ] = entry;
        c3 += __popc( m3 ); c4 += __popc( m4 );
        if ( c3 > (uint32_t)( kWarpBuf - 32 ) ) { warpBufFlush( wq.e[2], c3, Q.q_pb, Q.counts_el + 0 ); c3 = 0; }
        if ( c4 > (uint32_t)( kWarpBuf - 32 ) ) { warpBufFlush( wq.e[3], c4, Q.q_el, Q.counts_el + 1 ); c4 = 0; }
      }
    }
    if ( c1 ) warpBufFlush( wq.e[0], c1, Q.q_sab, Q.counts + 0 );
    if ( c2 ) warpBufFlush( wq.e[1], c2, Q.q_fg, Q.counts + 1 );
    if constexpr ( kDeferElastic ) {
      if ( c3 ) warpBufFlush( wq.e[2], c3, Q.q_pb, Q.counts_el + 0 );
      if ( c4 ) warpBufFlush( wq.e[3], c4, Q.q_el, Q.counts_el + 1 );
    }
  }

  // Elastic leaves over their queues: blockIdx.y = 0 PowderBragg (genScatterMu: one draw, search in the cumulative
  // structure-factor table staged in shared memory), 1 = ElIncScatter.  One entry per thread, all lanes on one path.
  __global__ void __launch_bounds__(256)
  k_sample_elastic( const __grid_constant__ Material M, const __grid_constant__ StagePlan sp,
                    const __grid_constant__ SampleArgs A, const __grid_constant__ QueueArgs Q )
  {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar;
    const bool is_pb = blockIdx.y == 0;
    const uint32_t nq = Q.counts_el[is_pb ? 0 : 1];
    if ( blockIdx.x * blockDim.x >= nq ) return;       // (before any staging: surplus CTAs leave at once)
    HotTabs H;
    if ( is_pb ) stageHotTabs( M, sp, smem, &mbar, H );
    const uint32_t* q = is_pb ? Q.q_pb : Q.q_el;
    const uint32_t stride = gridDim.x * blockDim.x;
    for ( uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nq; j += stride ) {
      const uint32_t entry = q[j];
      const uint32_t i = entry & kQueueIdxMask;
      const Comp& c = M.comp[ entry >> kQueueIdxBits ];
      const double ekin = A.ekin[i];
      Rng rng; rng.init( A.seed, A.streamIndex( i ), A.sid );
      rng.seek( M.ncomp > 1 ? 1u : 0u );
      double mu;
      if ( is_pb ) {
        const PowderBraggT& T = M.pb[c.idx];
        const int iv = pbLastValidPlane( T, H.pb_e2d[c.idx], H.pb_lut[c.idx], ekin );
        mu = pbSampleMu( H.pb_e2d[c.idx], H.pb_fdm[c.idx], iv, ekin, rng );
      } else {
        mu = elincSampleMu( M.elinc[c.idx], ekin, rng );
      }
      A.ekin_out[i] = ekin;
      A.mu_out[i] = mu;
      if ( A.ndraws ) A.ndraws[i] = rng.ndraws;
    }
  }

  // S(alpha,beta) table path, attempt-level scheduling ("lane refill").  The reference's
  // sampler is two nested rejection loops (NCSABSampler.cc:203-225 around
  // NCSABSamplerModels.cc:62-148) with ~75% acceptance: run neutron-per-lane, a warp waits
  // for its unluckiest lane.  Here every pass of the loop below executes exactly ONE
  // rejection attempt for every lane; a lane whose neutron is done pulls the next queue
  // entry (warp-aggregated atomic on a global cursor) before the next pass, so lanes stay
  // converged on the same code and busy until the queue drains.  The arithmetic and the
  // order in which each neutron consumes its uniforms are unchanged.
  template <bool kAtEmax, int kMinBlocks>
  __global__ void __launch_bounds__(128, kMinBlocks)
  k_sample_sab_refill( const __grid_constant__ Material M, const __grid_constant__ SampleArgs A,
                       const uint32_t* __restrict__ queue, const uint32_t* __restrict__ count,
                       uint32_t* __restrict__ cursor )
  {
    const uint32_t nq = *count;
    const int lane = threadIdx.x & 31;
    int errs = 0;
    // per-lane work item
    bool have = false;
    uint32_t idx = 0;
    int isab = 0, isampler = 0, inner = 0, outer = 0;
    bool ultra = false;
    double ekin_orig = 0.0, ekin_div_kT = 0.0, sampling_ediv = 0.0;
    Rng rng; rng.init( A.seed, A.first_index, A.sid );
    while ( true ) {
      // ---- refill
      const uint32_t need = __ballot_sync( 0xffffffffu, !have );
      if ( need ) {
        uint32_t base = 0;
        const int leader = __ffs( need ) - 1;
        if ( lane == leader )
          base = atomicAdd( cursor, (uint32_t)__popc( need ) );
        base = __shfl_sync( 0xffffffffu, base, leader );
        if ( !have ) {
          const uint32_t j = base + __popc( need & ( ( 1u << lane ) - 1u ) );
          if ( j < nq ) {
            const uint32_t entry = kAtEmax ? queue[2*j] : queue[j];
            idx = entry & kQueueIdxMask;
            isab = M.comp[ entry >> kQueueIdxBits ].idx;
            const SabT& T = M.sab[isab];
            ekin_orig = A.ekin[idx];
            rng.init( A.seed, A.streamIndex( idx ), A.sid );
            rng.seek( kAtEmax ? queue[2*j+1] : ( M.ncomp > 1 ? 1u : 0u ) );
            double ekin_eff;
            if ( kAtEmax ) {
              ekin_eff = T.egrid[T.negrid-1];
              isampler = T.negrid-1;
              ultra = false;
            } else {
              ekin_eff = ekin_orig;
              isampler = sabPickSampler( T, T.egrid, ekin_orig, ultra );
            }
            ekin_div_kT = ekin_eff / T.kT;
            sampling_ediv = ultra ? T.egrid[0] / T.kT : ekin_div_kT;
            inner = outer = 0;
            have = true;
          }
        }
      }
      if ( !__ballot_sync( 0xffffffffu, have ) )
        break;
      // ---- one attempt
      if ( have ) {
        const SabT& T = M.sab[isab];
        const SabEPoint ep = T.ep[isampler];
        double alpha = 0.0, beta = 0.0;
        int err = 0;
        bool done = false, failed = false;
        bool inner_ok = true;
        if ( ep.npts != 0 )
          inner_ok = sabAttemptFast( T, isampler, ep, sampling_ediv, rng, alpha, beta, err );
        if ( !inner_ok ) {
          if ( ++inner == 100 ) { err |= ERR_SAB_LOOP_INNER; failed = true; }
        } else {
          inner = 0;
          // outer acceptance at the neutron's (effective) energy, NCSABSampler.cc:205-224
          bool acc = false;
          if ( !( beta < -ekin_div_kT ) ) {
            AlphaLimits al = getAlphaLimits( ekin_div_kT, beta );
            if ( inInterval( al.first, al.second, alpha ) ) {
              acc = true;
            } else if ( ultra ) {
              alpha = al.first + rng.generate()*( al.second - al.first );
              acc = true;
            }
          }
          if ( acc ) done = true;
          else if ( ++outer == 100 ) { err |= ERR_SAB_LOOP_OUTER; failed = true; }
        }
        if ( done || failed ) {
          double eout = -1.0, mu = -999.0;
          if ( done )
            sabFinishScatter( T, ekin_orig, alpha, beta, rng, eout, mu, err );
          A.ekin_out[idx] = eout;
          A.mu_out[idx] = mu;
          if ( A.ndraws ) A.ndraws[idx] = rng.ndraws;
          have = false;
        }
        errs |= err;
      }
    }
    if ( errs )
      atomicOr( A.err_flags, errs );
  }

  // Partition the free-gas queue by energy class (two octaves per class): the free-gas samplers branch on E/kT
  // regimes, and warps -- and, for the instruction cache, whole SMs -- whose neutrons sit in one regime diverge far
  // less (water: k_sample_fg 2.85 -> 1.8 ms per 1e7; finer classes give nothing more).  Two small kernels with
  // block-aggregated bookkeeping: class totals per 8192-entry chunk (k_fg_hist), then every chunk reserves one run
  // per class in the output queue with ONE global atomic per class (k_fg_partition).  The complete 1024-bin
  // counting sort (NCB200_SORT=1) paid 1.9 ms for per-entry global atomics.
  constexpr int kFgGroupChunk = 8192;
  constexpr int kFgGroupClasses = 16;
  __device__ __forceinline__ uint32_t fgClass( double ekin )
  {
    const uint32_t c = sortKey( ekin ) >> 5;
    return c < (uint32_t)kFgGroupClasses ? c : (uint32_t)kFgGroupClasses - 1u;
  }
  // cls[0..16) class totals, cls[16..32) running cursors
  __global__ void __launch_bounds__(256)
  k_fg_hist( const double* __restrict__ ekin, const uint32_t* __restrict__ q_fg, const uint32_t* __restrict__ count,
             uint32_t* __restrict__ cls )
  {
    __shared__ uint32_t s_cnt[kFgGroupClasses];
    const uint32_t nq = *count;
    if ( threadIdx.x < kFgGroupClasses ) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for ( uint32_t k = blockIdx.x*blockDim.x + threadIdx.x; k < nq; k += gridDim.x*blockDim.x )
      atomicAdd( &s_cnt[ fgClass( ekin[ q_fg[k] & kQueueIdxMask ] ) ], 1u );
    __syncthreads();
    if ( threadIdx.x < kFgGroupClasses && s_cnt[threadIdx.x] )
      atomicAdd( &cls[threadIdx.x], s_cnt[threadIdx.x] );
  }
  __global__ void __launch_bounds__(256)
  k_fg_partition( const double* __restrict__ ekin, const uint32_t* __restrict__ q_fg, const uint32_t* __restrict__ count,
                  uint32_t* __restrict__ cls, uint32_t* __restrict__ q_out )
  {
    __shared__ uint32_t s_in[kFgGroupChunk];
    __shared__ uint8_t s_cls[kFgGroupChunk];
    __shared__ uint32_t s_cnt[kFgGroupClasses], s_pos[kFgGroupClasses];
    const uint32_t nq = *count;
    for ( uint32_t base = blockIdx.x * kFgGroupChunk; base < nq; base += gridDim.x * kFgGroupChunk ) {
      const uint32_t m = min( (uint32_t)kFgGroupChunk, nq - base );
      if ( threadIdx.x < kFgGroupClasses ) s_cnt[threadIdx.x] = 0;
      __syncthreads();
      for ( uint32_t k = threadIdx.x; k < m; k += blockDim.x ) {
        const uint32_t entry = q_fg[base + k];
        s_in[k] = entry;
        const uint32_t c = fgClass( ekin[ entry & kQueueIdxMask ] );
        s_cls[k] = (uint8_t)c;
        atomicAdd( &s_cnt[c], 1u );
      }
      __syncthreads();
      if ( threadIdx.x < kFgGroupClasses ) {
        uint32_t start = 0;                                   // first slot of this class in the output queue
        for ( int c = 0; c < (int)threadIdx.x; ++c ) start += cls[c];
        const uint32_t n = s_cnt[threadIdx.x];
        s_pos[threadIdx.x] = start + ( n ? atomicAdd( &cls[kFgGroupClasses + threadIdx.x], n ) : 0u );
      }
      __syncthreads();
      for ( uint32_t k = threadIdx.x; k < m; k += blockDim.x )
        q_out[ atomicAdd( &s_pos[ s_cls[k] ], 1u ) ] = s_in[k];
      __syncthreads();
    }
  }

  // Free-gas samplers over q_fg: the FreeGas leaf, and for S(alpha,beta) above Emax the
  // high-E analysis (SABSampler::sampleHighE); neutrons it sends back to the tabulated kernel
  // are appended to q_emax together with their stream position.
  template <int kMinBlocks>
  __global__ void __launch_bounds__(128, kMinBlocks)
  k_sample_fg( const __grid_constant__ Material M, const __grid_constant__ SampleArgs A,
               const __grid_constant__ QueueArgs Q )
  {
    const uint32_t nq = Q.counts[1];
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t nq_up = ( nq + 31u ) & ~31u;   // warp-uniform trip count (warpPush uses full-mask ballots)
    int errs = 0;
    for ( uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nq_up; j += stride ) {
      bool to_emax = false;
      uint32_t entry = 0, nd = 0;
      if ( j < nq ) {
        entry = Q.q_fg[j];
        const uint32_t i = entry & kQueueIdxMask;
        const int ich = (int)( entry >> kQueueIdxBits );
        const double ekin = A.ekin[i];
        Rng rng; rng.init( A.seed, A.streamIndex( i ), A.sid );
        rng.seek( M.ncomp > 1 ? 1u : 0u );
        const Comp& c = M.comp[ich];
        double eout = -1.0, mu = -999.0;
        int err = 0;
        if ( c.kind == KIND_FREEGAS ) {
          fgSampleScatter( M.fg[c.idx], ekin, rng, eout, mu, err );
        } else {
          const SabT& T = M.sab[c.idx];
          double alpha, beta;
          if ( sabSampleHighE( T, ekin, rng, alpha, beta, err ) ) {
            if ( !( err & ERR_SAB_DISCARD ) )
              sabFinishScatter( T, ekin, alpha, beta, rng, eout, mu, err );
          } else {
            to_emax = true;
          }
        }
        nd = rng.ndraws;
        if ( !to_emax ) {
          A.ekin_out[i] = eout;
          A.mu_out[i] = mu;
          if ( A.ndraws ) A.ndraws[i] = nd;
        }
        errs |= err;
      }
      // paired push: entry and stream position
      const uint32_t mask = __ballot_sync( 0xffffffffu, to_emax );
      if ( mask ) {
        const int lane = threadIdx.x & 31;
        const int leader = __ffs( mask ) - 1;
        uint32_t base = 0;
        if ( lane == leader )
          base = atomicAdd( Q.counts + 2, (uint32_t)__popc( mask ) );
        base = __shfl_sync( 0xffffffffu, base, leader );
        if ( to_emax ) {
          const uint32_t pos = base + __popc( mask & ( ( 1u << lane ) - 1u ) );
          Q.q_emax[2*pos] = entry;
          Q.q_emax[2*pos+1] = nd;
        }
      }
    }
    if ( errs )
      atomicOr( A.err_flags, errs );
  }

  // Free-gas samplers as a staged pipeline over the free-gas queue.  Run neutron-per-lane (k_sample_fg above), a warp
  // waits for its unluckiest lane twice: the beta sampler is a rejection loop with a long tail (Al above Emax: 4.2
  // attempts on average, 15 for the worst of 32 lanes) and so is the alpha sampler (6.7 uniforms on average, 27 for the
  // worst of 32); ncu: 9.9 of 32 lanes active.  Here each stage is a kernel in which all lanes run the same code:
  //   k_fg_prep        per entry: everything before the beta loop -- the high-E analysis of an S(alpha,beta) leaf
  //                    (discard / table at E=Emax / go on; <= 1 uniform) and the support [aa,bb] of the beta
  //                    distribution (FreeGasSampler::betaSupport: erfc evaluations, no uniforms);
  //   k_fg_beta        attempt-level: every pass of its loop runs ONE attempt of the beta rejection loop per lane; a
  //                    lane whose beta was accepted stores it and pulls the next queue entry before the next pass;
  //   k_fg_alpha_prep  per entry: the cheap alpha cases are finished, for the others the set-up of the alpha
  //                    rejection loop (xsBegin) is stored;
  //   k_fg_alpha       attempt-level, like k_fg_beta, over the alpha rejection loop (xsAttempt);
  //   k_fg_finish      per entry: (alpha,beta) -> outcome; S(alpha,beta) leaves: accept / table at E=Emax / draw
  //                    again (rare: finished in place with the plain nested loops).
  // The state of a neutron between stages is a per-entry record (9 doubles + 1 word); its random stream is resumed
  // from the number of uniforms consumed so far.  Each neutron consumes its uniforms in the reference's order
  // (NCFreeGasUtils.cc:530-935, NCSABSampler.cc:59-156).
  struct FgPrep {
    double* r;        // records: slot k of entry j at r[k*cap + j]
    uint64_t cap;
    uint32_t* w;      // uniforms consumed so far (low 25 bits) | flags (bits 25-27) | stage (bits 28-31); kFgSkip: done
    uint32_t* cursor; // refill cursors: [0] k_fg_beta, [1] k_fg_alpha
    uint32_t epl;     // refill kernels: target number of entries per lane (surplus CTAs exit at once)
    uint32_t batch;   // k_fg_beta: lanes waiting for the exact evaluation before it is run
    __device__ __forceinline__ double& slot( int k, uint32_t j ) const { return r[ (uint64_t)k*cap + j ]; }
  };
  constexpr uint32_t kFgSkip = 0xFFFFFFFFu;
  constexpr uint32_t kFgNdMask = ( 1u << 25 ) - 1u;
  // stages: 0..2 = FreeGasSampler::kBetaLoop/kBetaHighE/kBetaFixed (after k_fg_prep), then:
  enum { kFgBetaDone = 3, kFgAlphaDone = 4, kFgAlphaLoop = 5, kFgXDone = 6 };
  // record slots
  enum { kSlotAA = 0, kSlotBB = 1, kSlotPdisc = 2, kSlotBeta = 3,               // prep / beta
         kSlotC = 0, kSlotXm = 1, kSlotXp = 4, kSlotXmax = 5, kSlotXswitch = 6,  // alpha loop state
         kSlotPflat = 7, kSlotAright = 8,
         kSlotAlpha = 0,                                                         // result of the alpha stage (alpha or x)
         kFgSlots = 9 };

  __device__ __forceinline__ void fgLeafPars( const Material& M, const Comp& c, double& kT, double& mass )
  {
    if ( c.kind == KIND_FREEGAS ) { kT = M.fg[c.idx].kT; mass = M.fg[c.idx].mass_amu; }
    else { kT = M.sab[c.idx].ext.kT; mass = M.sab[c.idx].ext.mass_amu; }
  }

  // paired push of (entry, stream position) to the E=Emax queue; called by all lanes of a warp
  __device__ __forceinline__ void pushEmax( bool pred, const QueueArgs& Q, uint32_t entry, uint32_t nd )
  {
    const uint32_t mask = __ballot_sync( 0xffffffffu, pred );
    if ( !mask ) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs( mask ) - 1;
    uint32_t base = 0;
    if ( lane == leader )
      base = atomicAdd( Q.counts + 2, (uint32_t)__popc( mask ) );
    base = __shfl_sync( 0xffffffffu, base, leader );
    if ( pred ) {
      const uint32_t pos = base + __popc( mask & ( ( 1u << lane ) - 1u ) );
      Q.q_emax[2*pos] = entry;
      Q.q_emax[2*pos+1] = nd;
    }
  }

  __global__ void __launch_bounds__(128)
  k_fg_prep( const __grid_constant__ Material M, const __grid_constant__ SampleArgs A,
             const __grid_constant__ QueueArgs Q, const __grid_constant__ FgPrep P )
  {
    const uint32_t nq = Q.counts[1];
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t nq_up = ( nq + 31u ) & ~31u;   // warp-uniform trip count (full-mask ballots in pushEmax)
    int errs = 0;
    for ( uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nq_up; j += stride ) {
      bool to_emax = false;
      uint32_t entry = 0, nd = 0;
      if ( j < nq ) {
        entry = Q.q_fg[j];
        const uint32_t i = entry & kQueueIdxMask;
        const Comp& c = M.comp[ entry >> kQueueIdxBits ];
        const double ekin = A.ekin[i];
        nd = M.ncomp > 1 ? 1u : 0u;
        double pdisc = 0.0;
        int begin = kHighEGoOn;
        if ( c.kind != KIND_FREEGAS ) {
          Rng rng; rng.init( A.seed, A.streamIndex( i ), A.sid );
          rng.seek( nd );
          int err = 0;
          begin = sabHighEBegin( M.sab[c.idx], ekin, rng, pdisc, err );
          nd = rng.ndraws;
          errs |= err;
        }
        if ( begin == kHighEDiscard ) {
          A.ekin_out[i] = -1.0;
          A.mu_out[i] = -999.0;
          if ( A.ndraws ) A.ndraws[i] = nd;
          P.w[j] = kFgSkip;
        } else if ( begin == kHighEToEmax ) {
          to_emax = true;
          P.w[j] = kFgSkip;
        } else {
          double kT, mass;
          fgLeafPars( M, c, kT, mass );
          FreeGasSampler s( ekin, kT, mass );
          double aa, bb;
          const int kind = s.betaSupport( aa, bb );
          P.slot( kSlotAA, j ) = aa; P.slot( kSlotBB, j ) = bb; P.slot( kSlotPdisc, j ) = pdisc;
          P.w[j] = nd | ( (uint32_t)kind << 28 );
        }
      }
      pushEmax( to_emax, Q, entry, nd );
    }
    if ( errs )
      atomicOr( A.err_flags, errs );
  }

  // Work distribution of the two attempt-level kernels: lanes pull queue positions from a global cursor
  // (warp-aggregated atomic).  Returns the position for this lane (>= nq: none) and whether the queue is drained.
  __device__ __forceinline__ uint32_t fgPull( uint32_t need, bool mine, uint32_t* cursor, uint32_t nq, bool& drained )
  {
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    const int leader = __ffs( need ) - 1;
    if ( lane == leader )
      base = atomicAdd( cursor, (uint32_t)__popc( need ) );
    base = __shfl_sync( 0xffffffffu, base, leader );
    drained = ( base >= nq || nq - base <= (uint32_t)__popc( need ) );
    return mine ? base + __popc( need & ( ( 1u << lane ) - 1u ) ) : 0xFFFFFFFFu;
  }

  template <int kMinBlocks>
  __global__ void __launch_bounds__(128, kMinBlocks)
  k_fg_beta( const __grid_constant__ Material M, const __grid_constant__ SampleArgs A,
             const __grid_constant__ QueueArgs Q, const __grid_constant__ FgPrep P )
  {
    const uint32_t nq = Q.counts[1];
    if ( (uint64_t)blockIdx.x * blockDim.x * P.epl >= nq && blockIdx.x )
      return;
    // An attempt has a cheap half (candidate + erfc bounds from the lookup table) and an expensive one (exact erfc
    // evaluation, needed by roughly every second candidate).  Lanes that need the exact evaluation wait while the
    // others go on with cheap halves (and refills), until P.batch lanes are waiting or nobody else can move: the
    // expensive code then runs with most lanes active instead of about a third of them.
    bool have = false, wait_exact = false, drained = false;
    uint32_t jrec = 0;
    double beta = 0.0, faccept = 0.0;
    FreeGasSampler s;
    FreeGasSampler::BetaState st;
    Rng rng; rng.init( A.seed, A.first_index, A.sid );
    while ( true ) {
      while ( true ) {
        const uint32_t need = __ballot_sync( 0xffffffffu, !have );
        if ( need && !drained ) {
          const uint32_t j = fgPull( need, !have, P.cursor + 0, nq, drained );
          if ( !have && j < nq ) {
            const uint32_t w = P.w[j];
            if ( w != kFgSkip ) {
              const uint32_t entry = Q.q_fg[j];
              const uint32_t i = entry & kQueueIdxMask;
              double kT, mass;
              fgLeafPars( M, M.comp[ entry >> kQueueIdxBits ], kT, mass );
              s = FreeGasSampler( A.ekin[i], kT, mass );
              rng.init( A.seed, A.streamIndex( i ), A.sid );
              rng.seek( w & kFgNdMask );
              const int kind = (int)( w >> 28 );
              const double aa = P.slot( kSlotAA, j );
              if ( kind != FreeGasSampler::kBetaLoop ) {
                P.slot( kSlotBeta, j ) = s.betaDirect( kind, aa, rng );
                P.w[j] = rng.ndraws | ( (uint32_t)kFgBetaDone << 28 );
              } else {
                s.betaBegin( st, aa, P.slot( kSlotBB, j ) );
                jrec = j;
                have = true;
                wait_exact = false;
              }
            }
          }
        }
        const bool run = have && !wait_exact;
        const uint32_t running = __ballot_sync( 0xffffffffu, run );
        if ( !running ) {
          // nobody can do a cheap half: leave unless lanes are idle only because they met finished records
          if ( drained || !__ballot_sync( 0xffffffffu, !have ) ) break;
          continue;
        }
        if ( run ) {
          const int q = s.betaAttemptQuick( st, rng, beta, faccept );
          if ( q == FreeGasSampler::kBetaAccept ) {
            P.slot( kSlotBeta, jrec ) = beta;
            P.w[jrec] = rng.ndraws | ( (uint32_t)kFgBetaDone << 28 );
            have = false;
          } else if ( q == FreeGasSampler::kBetaNeedExact ) {
            wait_exact = true;
          }
        }
        if ( (uint32_t)__popc( __ballot_sync( 0xffffffffu, have && wait_exact ) ) >= P.batch )
          break;
      }
      if ( !__ballot_sync( 0xffffffffu, have ) )
        break;      // the inner loop ends with nothing in flight only when the queue is drained
      if ( have && wait_exact ) {
        wait_exact = false;
        if ( s.betaAttemptExact( st, beta, faccept ) ) {
          P.slot( kSlotBeta, jrec ) = beta;
          P.w[jrec] = rng.ndraws | ( (uint32_t)kFgBetaDone << 28 );
          have = false;
        }
      }
    }
  }

  __global__ void __launch_bounds__(128)
  k_fg_alpha_prep( const __grid_constant__ Material M, const __grid_constant__ SampleArgs A,
                   const __grid_constant__ QueueArgs Q, const __grid_constant__ FgPrep P )
  {
    const uint32_t nq = Q.counts[1];
    const uint32_t stride = gridDim.x * blockDim.x;
    for ( uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nq; j += stride ) {
      const uint32_t w = P.w[j];
      if ( w == kFgSkip ) continue;
      const uint32_t entry = Q.q_fg[j];
      const uint32_t i = entry & kQueueIdxMask;
      const Comp& c = M.comp[ entry >> kQueueIdxBits ];
      const double ekin = A.ekin[i];
      double kT, mass;
      fgLeafPars( M, c, kT, mass );
      FreeGasSampler s( ekin, kT, mass );
      Rng rng; rng.init( A.seed, A.streamIndex( i ), A.sid );
      rng.seek( w & kFgNdMask );
      const double beta = P.slot( kSlotBeta, j );
      const bool iso = muIsotropicAtBeta( beta, s.m_c );
      if ( c.kind == KIND_FREEGAS && ( beta <= -s.m_c || iso ) ) {
        // FreeGasSampler::sampleDeltaEMu, isotropic branch (NCFreeGasUtils.hh:163-167)
        const double mu = rng.generate()*2.0 - 1.0;
        A.ekin_out[i] = dmax( 0.0, ekin + beta*s.m_kT );
        A.mu_out[i] = mu;
        if ( A.ndraws ) A.ndraws[i] = rng.ndraws;
        P.w[j] = kFgSkip;
        continue;
      }
      double alpha;
      int kind = FreeGasSampler::kAlphaDone;
      XSamplerState st;
      if ( c.kind != KIND_FREEGAS && ( beta < -s.m_c || iso ) ) {
        // FreeGasSampler::sampleAlphaBeta, flat branch (NCFreeGasUtils.hh:149-154)
        AlphaLimits alim = getAlphaLimits( s.m_c_real, beta );
        const double a = alim.first + rng.generate()*(alim.second-alim.first);
        alpha = dclamp( a, alim.first, alim.second );
      } else {
        kind = s.alphaBegin( beta, rng, st, alpha );
      }
      if ( kind == FreeGasSampler::kAlphaDone ) {
        P.slot( kSlotAlpha, j ) = alpha;
        P.w[j] = rng.ndraws | ( (uint32_t)kFgAlphaDone << 28 );
      } else {
        P.slot( kSlotC, j ) = st.c; P.slot( kSlotXm, j ) = st.xm; P.slot( kSlotXp, j ) = st.xp;
        P.slot( kSlotXmax, j ) = st.xmax; P.slot( kSlotXswitch, j ) = st.xswitch;
        P.slot( kSlotPflat, j ) = st.probability_flat; P.slot( kSlotAright, j ) = st.area_right;
        const uint32_t flags = ( st.always_left ? 1u : 0u ) | ( st.always_right ? 2u : 0u ) | ( st.single_side ? 4u : 0u );
        P.w[j] = rng.ndraws | ( flags << 25 ) | ( (uint32_t)kFgAlphaLoop << 28 );
      }
    }
  }

  template <int kMinBlocks>
  __global__ void __launch_bounds__(128, kMinBlocks)
  k_fg_alpha( const __grid_constant__ Material M, const __grid_constant__ SampleArgs A,
              const __grid_constant__ QueueArgs Q, const __grid_constant__ FgPrep P )
  {
    const uint32_t nq = Q.counts[1];
    if ( (uint64_t)blockIdx.x * blockDim.x * P.epl >= nq && blockIdx.x )
      return;
    bool have = false, drained = false;
    uint32_t jrec = 0;
    XSamplerState st;
    Rng rng; rng.init( A.seed, A.first_index, A.sid );
    while ( true ) {
      const uint32_t need = __ballot_sync( 0xffffffffu, !have );
      if ( need && !drained ) {
        const uint32_t j = fgPull( need, !have, P.cursor + 1, nq, drained );
        if ( !have && j < nq ) {
          const uint32_t w = P.w[j];
          if ( w != kFgSkip && ( w >> 28 ) == (uint32_t)kFgAlphaLoop ) {
            const uint32_t i = Q.q_fg[j] & kQueueIdxMask;
            rng.init( A.seed, A.streamIndex( i ), A.sid );
            rng.seek( w & kFgNdMask );
            st.c = P.slot( kSlotC, j ); st.xm = P.slot( kSlotXm, j ); st.xp = P.slot( kSlotXp, j );
            st.xmax = P.slot( kSlotXmax, j ); st.xswitch = P.slot( kSlotXswitch, j );
            st.probability_flat = P.slot( kSlotPflat, j ); st.area_right = P.slot( kSlotAright, j );
            const uint32_t flags = ( w >> 25 ) & 7u;
            st.always_left = flags & 1u; st.always_right = flags & 2u; st.single_side = flags & 4u;
            jrec = j;
            have = true;
          }
        }
      }
      if ( !__ballot_sync( 0xffffffffu, have ) ) {
        if ( drained ) break;
        continue;
      }
      if ( have ) {
        double x;
        if ( xsAttempt( st, rng, x ) ) {
          P.slot( kSlotAlpha, jrec ) = x;
          P.w[jrec] = rng.ndraws | ( (uint32_t)kFgXDone << 28 );
          have = false;
        }
      }
    }
  }

  __global__ void __launch_bounds__(128)
  k_fg_finish( const __grid_constant__ Material M, const __grid_constant__ SampleArgs A,
               const __grid_constant__ QueueArgs Q, const __grid_constant__ FgPrep P )
  {
    const uint32_t nq = Q.counts[1];
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t nq_up = ( nq + 31u ) & ~31u;   // warp-uniform trip count (full-mask ballots in pushEmax)
    int errs = 0;
    for ( uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nq_up; j += stride ) {
      bool to_emax = false;
      uint32_t entry = 0, nd = 0;
      const uint32_t w = j < nq ? P.w[j] : kFgSkip;
      if ( w != kFgSkip ) {
        entry = Q.q_fg[j];
        const uint32_t i = entry & kQueueIdxMask;
        const Comp& c = M.comp[ entry >> kQueueIdxBits ];
        const double ekin = A.ekin[i];
        double kT, mass;
        fgLeafPars( M, c, kT, mass );
        // FreeGasSampler ctor, the two members needed here (NCFreeGasUtils.cc:492-515)
        const double cc = dmin( 1e14, dmax( 1e-10, ekin/kT ) );
        const double Adiv4 = 0.25*( kInvNeutronMassAmu * mass );
        double beta = P.slot( kSlotBeta, j );
        double alpha = P.slot( kSlotAlpha, j );
        if ( ( w >> 28 ) == (uint32_t)kFgXDone )
          alpha = FreeGasSampler::alphaFromX( cc, Adiv4, beta, alpha );
        Rng rng; rng.init( A.seed, A.streamIndex( i ), A.sid );
        rng.seek( w & kFgNdMask );
        double eout = -1.0, mu = -999.0;
        int err = 0;
        if ( c.kind == KIND_FREEGAS ) {
          // tail of FreeGasSampler::sampleDeltaEMu + FreeGas::sampleScatterIsotropic (src/freegas/NCFreeGas.cc:70-75)
          double dE;
          alphaBetaToDeltaEMu( alpha, beta, cc*kT, kT, dE, mu, err );
          eout = dmax( 0.0, ekin + dE );
        } else {
          const SabT& T = M.sab[c.idx];
          const double pdisc = P.slot( kSlotPdisc, j );
          int chk = sabHighECheck( T, alpha, beta, pdisc, rng );
          if ( chk == kHighERedo ) {
            // rare: draw (alpha,beta) again with the plain nested loops
            FreeGasSampler s( ekin, kT, mass );
            do {
              s.sampleAlphaBeta( rng, alpha, beta );
              chk = sabHighECheck( T, alpha, beta, pdisc, rng );
            } while ( chk == kHighERedo );
          }
          if ( chk == kHighEAccept )
            sabFinishScatter( T, ekin, alpha, beta, rng, eout, mu, err );
          else
            to_emax = true;
        }
        nd = rng.ndraws;
        if ( !to_emax ) {
          A.ekin_out[i] = eout;
          A.mu_out[i] = mu;
          if ( A.ndraws ) A.ndraws[i] = nd;
        }
        errs |= err;
      }
      pushEmax( to_emax, Q, entry, nd );
    }
    if ( errs )
      atomicOr( A.err_flags, errs );
  }

  // FP64 FMA probe: the denominator for "FP64 pipe utilisation against B200 peak" (SURVEY 8d asks for a measured
  // figure: the vendor's vector-FP64 number is not in MEASURED_PEAKS.json).  8 independent DFMA chains per thread.
  __global__ void __launch_bounds__(256)
  k_fp64_fma_probe( double* __restrict__ out, int iters, double a, double b )
  {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for ( int i = 0; i < iters; ++i ) {
      x0 = fma( x0, a, b ); x1 = fma( x1, a, b ); x2 = fma( x2, a, b ); x3 = fma( x3, a, b );
      x4 = fma( x4, a, b ); x5 = fma( x5, a, b ); x6 = fma( x6, a, b ); x7 = fma( x7, a, b );
    }
    const double s = ( ( x0 + x1 ) + ( x2 + x3 ) ) + ( ( x4 + x5 ) + ( x6 + x7 ) );
    if ( s == 12345.678 ) out[0] = s;   // keeps the chains alive
  }

  // ------------------------------------------------------------ oriented (single crystal)
  // Per-neutron (E, direction), SoA.  One neutron per thread: the lanes of a warp walk the
  // reflection-family / demi-normal tables in lockstep (same addresses -> broadcast loads);
  // only normals inside the mosaic truncation cone (a handful per neutron) reach the
  // circle-integral code.
  struct DirArgs {
    const double* ux; const double* uy; const double* uz;   // in
    double* ox; double* oy; double* oz;                      // out (sampling only)
  };

  __global__ void __launch_bounds__(128)
  k_xs_aniso( const __grid_constant__ Material M, const __grid_constant__ StagePlan sp,
              const double* __restrict__ ekin, const __grid_constant__ DirArgs D, uint64_t n, double* __restrict__ out )
  {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar;
    HotTabs H;
    stageHotTabs( M, sp, smem, &mbar, H );
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for ( uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride )
      out[i] = matXS( M, H, ekin[i], Vec3{ D.ux[i], D.uy[i], D.uz[i] }, nullptr, nullptr, nullptr );
  }

  __global__ void __launch_bounds__(128)
  k_sample_aniso( const __grid_constant__ Material M, const __grid_constant__ StagePlan sp,
                  const __grid_constant__ SampleArgs A, const __grid_constant__ DirArgs D )
  {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar;
    HotTabs H;
    stageHotTabs( M, sp, smem, &mbar, H );
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    int errs = 0;
    for ( uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += stride ) {
      Rng rng; rng.init( A.seed, A.streamIndex( i ), A.sid );
      double eout;
      Vec3 o;
      int err = 0, ich;
      const double xs = matSample( M, H, A.ekin[i], Vec3{ D.ux[i], D.uy[i], D.uz[i] }, rng, eout, o, err, ich );
      if ( err & ( ERR_SAB_LOOP_INNER | ERR_SAB_LOOP_OUTER | ERR_SAB_DISCARD | ERR_KIN_DENOM ) ) {
        eout = -1.0; o = { 0.0, 0.0, 0.0 };
      }
      A.ekin_out[i] = eout;
      D.ox[i] = o.x; D.oy[i] = o.y; D.oz[i] = o.z;
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
    const uint32_t off_b = (uint32_t)( (size_t)ie*(size_t)T.bstride );
    int err = 0;
    SabEPoint e;
    xscheck[ie] = sabAssembleEPoint( T.beta, T.nbeta, T.kT, T.bound_xs, T.egrid[ie], rows + (size_t)ie*T.nbeta,
                                     off_b, (uint32_t)( (size_t)ie*T.nbeta ), e, bx + off_b, bpdf + off_b, bcdf + off_b, err );
    ep[ie] = e;
    errs[ie] = err;
  }

  // guide tables: blockIdx.y = energy point (beta guides) or beta row - negrid (alpha guides)
  __global__ void k_sab_guides( SabT T, SabEPoint* __restrict__ ep, uint16_t* __restrict__ bguide,
                                uint16_t* __restrict__ aguide, double* __restrict__ ascale )
  {
    const int y = blockIdx.y;
    if ( y < T.negrid ) {
      const SabEPoint e = ep[y];
      const double* cdf = T.bcdf + e.off_b;
      uint16_t* g = bguide + (size_t)y*kSabGBStride;
      for ( int b = blockIdx.x*blockDim.x + threadIdx.x; b <= kSabGB; b += gridDim.x*blockDim.x )
        g[b] = sabBetaGuideEntry( cdf, e.npts, b );
      if ( blockIdx.x == 0 && threadIdx.x == 0 )
        ep[y].guide = e.npts > 0 ? g : nullptr;
    } else {
      const int ib = y - T.negrid;
      const double* row = T.cumul + (size_t)ib*T.nalpha;
      const double sc = sabAlphaScale( row, T.nalpha );
      uint16_t* g = aguide + (size_t)ib*( kSabGA+1 );
      for ( int b = blockIdx.x*blockDim.x + threadIdx.x; b <= kSabGA; b += gridDim.x*blockDim.x )
        g[b] = sabAlphaGuideEntry( row, T.nalpha, sc, b );
      if ( blockIdx.x == 0 && threadIdx.x == 0 )
        ascale[ib] = sc;
    }
  }

  // stage 4: gather-friendly copies.  blockIdx.y = 0: heads + tails (one thread per (energy point, beta row));
  // 1: points (one per (beta row, alpha)); 2: beta-distribution points (one per (energy point, point));
  // 3: log guide (one thread per (beta row, key))
  __global__ void k_sab_gather_tabs( SabT T, SabHead* __restrict__ heads, SabTail* __restrict__ tails,
                                     SabPoint* __restrict__ pts, SabBPoint* __restrict__ bpts, uint16_t* __restrict__ lguide )
  {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ( blockIdx.y == 0 ) {
      if ( k < (size_t)T.negrid*T.nbeta ) {
        const int ib = (int)( k % (size_t)T.nbeta );
        const SabAlphaInfo info = T.ainfo[k];
        heads[k] = sabMakeHead( info, T.cumul + (size_t)ib*T.nalpha, T.nalpha );
        sabMakeTails( info, tails + 2*k );
      }
    } else if ( blockIdx.y == 1 ) {
      if ( k < (size_t)T.nbeta*T.nalpha )
        pts[k] = sabMakePoint( T.alpha, T.sab, T.logsab, T.cumul, T.nalpha, k );
    } else if ( blockIdx.y == 2 ) {
      if ( k < (size_t)T.negrid*T.bstride ) {
        SabBPoint b; b.x = T.bx[k]; b.pdf = T.bpdf[k]; b.cdf = T.bcdf[k]; b.pad = 0.0;
        bpts[k] = b;
      }
    } else if ( k < (size_t)T.nbeta*( kSabGL + 1 ) ) {
      const int ib = (int)( k / (size_t)( kSabGL + 1 ) ), key = (int)( k % (size_t)( kSabGL + 1 ) );
      const double* row = T.cumul + (size_t)ib*T.nalpha;
      const double tot = row[T.nalpha-1];
      const double inv = ( tot > 0.0 && isFinite( 1.0/tot ) ) ? 1.0/tot : 0.0;
      lguide[(size_t)ib*kSabGLStride + key] = sabLogGuideEntry( row, T.nalpha, inv, key );
    }
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
      const double v = ldStream( values + i );
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
