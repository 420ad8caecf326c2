// ncb_kernels_sc.cuh -- warp-cooperative kernels for oriented materials (SCBragg).
//
// For every neutron (E, direction) SCBragg tests ALL demi-normals of the reflection families with
// wl < 2d (Ge: up to 1119) against the mosaic truncation cone and integrates the few that pass
// (NCSCBragg.cc:233-275, NCGaussMos.cc:147-194).  One neutron per lane makes a warp (a) run as long
// as its highest-energy lane and (b) wait for every lane's rare circle-integral in turn.  Here ONE
// WARP handles one neutron: lane l tests normals l, l+32, ... (the family/normal tables are staged
// in shared memory by TMA bulk copies), candidates are compacted in order with ballots, their
// integrals are evaluated by different lanes in parallel, and the running cumulative sums
// ("xs_commul") are then formed in the reference's order.
//
//   k_sc_scan        per neutron: SCBragg cross section + number of contributing normals (small batches); large
//                    batches take the two-stage form:
//   k_sc_find        candidate search, a warp per neutron in batches of 32 consecutive neutrons (single-precision
//                    pre-filter on packed records, exact test on the survivors) -> work list + candidate lists
//   k_sc_eval_flat   evaluation of the recorded candidates, one candidate per lane, accumulated per neutron in
//                    plane order (k_sc_eval_groups: the eight-lanes-per-neutron form it replaced, tuning builds;
//                    k_sc_eval: warp per neutron, for the rare lists longer than the record)
//   k_classify_aniso per neutron (one per lane): composition sum with that result, component pick,
//                    PowderBragg/ElInc sampled in place (+direction), S(a,b)/free-gas -> queues,
//                    SCBragg-chosen -> queue
//   k_sc_sample      per queued neutron (one per warp): pick the normal, GaussMos::genScat
//   k_dir_from_mu    outgoing direction for the neutrons sampled by the isotropic queue kernels
//                    (ScatterIsotropicMat::sampleScatter, NCProcImpl.cc:29-37): uniform azimuth
#pragma once
#include "ncb_kernels.cuh"

namespace ncb {

  constexpr int kScMaxFam = 128;      // families whose per-neutron parameters fit the per-warp scratch
  constexpr int kScCandCap = 64;
  constexpr int kScRecCap = 32;
  constexpr int kScWarps = 8;         // warps per CTA (they share the staged tables)

  struct ScWarpScratch {
    // per-family window of |normal . direction| outside which a plane cannot pass the truncation test (a superset,
    // widened by 1e-6; the exact test of the reference is then run on the few survivors)
    float lo[kScMaxFam];
    float hi[kScMaxFam];
    double cptsq[kScMaxFam];          // InteractionPars of the neutron per family, for the exact test
    double spt[kScMaxFam];
    double vals[2*kScCandCap];        // raw xs of (-normal, +normal) per candidate
    uint16_t cand[kScCandCap];
    // record of a cross-section walk (mode 2): cumulative sum and (normal, sign) of every contributing entry, so that a
    // caller that goes on to scatter on this crystal selects among them without walking again (transport tail kernel)
    double rec_cumul[kScRecCap];
    uint16_t rec_entry[kScRecCap];    // normal index | sign << 15
  };

  // state of the ordered accumulation (warp-uniform)
  struct ScAccum {
    int cur_fam, n;
    double xsoffset, xssum, commul_last;
    // mode 1
    bool found;
    int chosen_in, chosen_sign;
  };

  // Evaluate the `count` recorded candidates (lanes in parallel), then accumulate in order.
  // mode 0: total/count.  mode 1: stop at the entry selected by `choice` (rule: linear '>' / lower_bound '>=').
  // mode 2: as 0, and the first kScRecCap contributing entries are recorded in the scratch (rec_cumul / rec_entry).
  __device__ __forceinline__ void scFlush( const ScBraggT& S, ScWarpScratch& ws, const uint8_t* fam_of,
                                           double wl, const Vec3& d, int count, ScAccum& acc,
                                           int mode, bool linear, double choice )
  {
    const int lane = threadIdx.x & 31;
    for ( int k0 = 0; k0 < count; k0 += 32 ) {
      const int k = k0 + lane;
      if ( k < count ) {
        const int in = ws.cand[k];
        const int f = fam_of[in];
        InteractionPars ip;
        ip.set( wl, S.fam_inv2d[f], S.fam_xsfact[f] );
        const double nx = S.normals[3*in], ny = S.normals[3*in+1], nz = S.normals[3*in+2];
        const double dot = nx*d.x + ny*d.y + nz*d.z;
        const double sdotcptsq = ( 1.0 - dot*dot )*ip.cos_perfect_theta_sq;
        const double ds = dot * ip.sin_perfect_theta;
        double xm = 0.0, xp = 0.0;
        const double Am = dmax( 0.0, S.cta - ds );
        if ( sdotcptsq > Am*Am ) xm = gmRawXS( S, ip, dot );     // anti-normal
        const double Ap = dmax( 0.0, S.cta + ds );
        if ( sdotcptsq > Ap*Ap ) xp = gmRawXS( S, ip, -dot );    // normal
        ws.vals[2*k] = xm; ws.vals[2*k+1] = xp;
      }
    }
    __syncwarp();
    // ordered accumulation, all lanes redundantly (<= 128 values)
    for ( int k = 0; k < count && !( mode == 1 && acc.found ); ++k ) {
      const int in = ws.cand[k];
      const int f = fam_of[in];
      if ( f != acc.cur_fam ) { acc.cur_fam = f; acc.xsoffset = acc.commul_last; acc.xssum = 0.0; }
      for ( int sgn = 0; sgn < 2; ++sgn ) {
        const double xs = ws.vals[2*k+sgn];
        if ( xs ) {
          acc.commul_last = acc.xsoffset + ( acc.xssum += xs );
          ++acc.n;
          if ( mode == 2 ) {
            if ( acc.n <= kScRecCap && lane == 0 ) { ws.rec_cumul[acc.n-1] = acc.commul_last; ws.rec_entry[acc.n-1] = (uint16_t)( in | ( sgn << 15 ) ); }
          } else if ( mode ) {
            acc.chosen_in = in; acc.chosen_sign = sgn;
            if ( linear ? ( acc.commul_last > choice ) : !( acc.commul_last < choice ) ) { acc.found = true; break; }
          }
        }
      }
    }
    __syncwarp();
  }

  // One walk for the neutron owned by this warp.  Returns through acc.
  __device__ __forceinline__ void scWalkWarp( const ScBraggT& S, ScWarpScratch& ws, const uint8_t* fam_of,
                                              double ekin_raw, const Vec3& d, double& wl_out, ScAccum& acc,
                                              int mode, bool linear, double choice )
  {
    const int lane = threadIdx.x & 31;
    __syncwarp();   // the scratch is reused from the previous neutron of this warp: its last reads come first
    acc.cur_fam = -1; acc.n = 0; acc.xsoffset = acc.xssum = acc.commul_last = 0.0;
    acc.found = false; acc.chosen_in = 0; acc.chosen_sign = 1;
    const double ekin = scCacheRound( ekin_raw );
    const double wl = ekin ? sqrt( kWl2Ekin / ekin ) : kInf;
    wl_out = wl;
    if ( wl == 0 ) return;
    const double inv2dcutoff = ( 1.0 - 2*kDblEps )/wl;
    // number of active families (sorted by inv2d ascending) and their scan parameters
    const double cta = S.cta;
    int nfam_act = 0;
    for ( int f0 = 0; f0 < S.nfam; f0 += 32 ) {
      const int f = f0 + lane;
      const bool act = ( f < S.nfam ) && ( S.fam_inv2d[f] < inv2dcutoff );
      if ( act ) {
        InteractionPars ip;
        ip.set( wl, S.fam_inv2d[f], S.fam_xsfact[f] );
        // The truncation test of scIsCandidate, (1-x^2) cos^2(thB) > max(0, cos(tau) - x sin(thB))^2 with
        // x = |n.d| = sin(phi), holds iff |thB - phi| < tau, i.e. x in ( sin(thB-tau), sin(thB+tau) )
        // (no upper limit once thB+tau >= 90 deg, no lower limit once thB-tau <= 0).
        const double spt = ip.sin_perfect_theta, cpt = sqrt( ip.cos_perfect_theta_sq );
        const double slo = spt*cta - cpt*S.sta, shi = spt*cta + cpt*S.sta;
        const bool open_hi = !( cpt*cta - spt*S.sta > 1e-6 );
        ws.lo[f] = (float)( slo - 1e-6 );
        ws.hi[f] = open_hi ? 2.0f : (float)( shi + 1e-6 );
        ws.cptsq[f] = ip.cos_perfect_theta_sq;
        ws.spt[f] = spt;
      }
      const uint32_t m = __ballot_sync( 0xffffffffu, act );
      nfam_act += __popc( m );
      if ( m != 0xffffffffu ) break;
    }
    __syncwarp();
    if ( nfam_act == 0 ) return;
    const int n_act = S.fam_first[nfam_act];
    int count = 0;
    // two independent 32-normal slices per iteration (instruction-level parallelism: the kernel runs at
    // 2 CTAs/SM and was latency-bound on its own dependent fp64 chain)
    for ( int base = 0; base < n_act; base += 64 ) {
      bool c[2];
      int inn[2];
#pragma unroll
      for ( int u = 0; u < 2; ++u ) {
        const int in = base + 32*u + lane;
        inn[u] = in;
        bool pre = false;
        double dot = 0.0;
        int f = 0;
        if ( in < n_act ) {
          f = fam_of[in];
          dot = S.normals[3*in]*d.x + S.normals[3*in+1]*d.y + S.normals[3*in+2]*d.z;
          const double x = fabs( dot );
          pre = ( x > (double)ws.lo[f] ) & ( x < (double)ws.hi[f] );
        }
        c[u] = false;
        if ( pre ) {
          double sd, ds;
          c[u] = scIsCandidate( cta, ws.cptsq[f], ws.spt[f], dot, sd, ds );
        }
      }
#pragma unroll
      for ( int u = 0; u < 2; ++u ) {
        const uint32_t m = __ballot_sync( 0xffffffffu, c[u] );
        if ( m ) {
          if ( c[u] )
            ws.cand[ count + __popc( m & ( ( 1u << lane ) - 1u ) ) ] = (uint16_t)inn[u];
          count += __popc( m );
          __syncwarp();
          if ( count > kScCandCap - 32 ) {
            scFlush( S, ws, fam_of, wl, d, count, acc, mode, linear, choice );
            count = 0;
            if ( mode == 1 && acc.found ) return;
          }
        }
      }
    }
    if ( count )
      scFlush( S, ws, fam_of, wl, d, count, acc, mode, linear, choice );
  }

  // shared set-up of the SC kernels: staged tables + normal->family map
  __device__ __forceinline__ void scBlockSetup( const Material& M, const StagePlan& sp, unsigned char* smem, uint64_t* mbar,
                                                HotTabs& H, uint8_t* fam_of )
  {
    stageHotTabs( M, sp, smem, mbar, H );
    const ScBraggT& S = *H.sc;
    for ( int f = threadIdx.x; f < S.nfam; f += blockDim.x )
      for ( int in = S.fam_first[f]; in < S.fam_first[f+1]; ++in )
        fam_of[in] = (uint8_t)f;
    __syncthreads();
  }

  struct ScScanArgs {
    const double* ekin; const double* ux; const double* uy; const double* uz;
    uint64_t n;
    double* sc_xs;     // out: unscaled SCBragg xs
    int32_t* sc_n;     // out: entries of xs_commul
    double dom_lo, dom_hi;   // the SCBragg component's domain
    const uint32_t* n_dev = nullptr;   // see SampleArgs::n_dev
  };

  // dynamic smem layout: [staged tables (sp.total)] [fam_of: nnormals bytes, 16-aligned] [kScWarps x ScWarpScratch]
  __global__ void __launch_bounds__(32*kScWarps, 2)
  k_sc_scan( const __grid_constant__ Material M, const __grid_constant__ StagePlan sp,
             const __grid_constant__ ScScanArgs A, uint32_t fam_of_off, uint32_t scratch_off )
  {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar;
    HotTabs H;
    uint8_t* fam_of = smem + fam_of_off;
    scBlockSetup( M, sp, smem, &mbar, H, fam_of );
    const ScBraggT& S = *H.sc;
    ScWarpScratch& ws = reinterpret_cast<ScWarpScratch*>( smem + scratch_off )[ threadIdx.x >> 5 ];
    const int lane = threadIdx.x & 31;
    const uint64_t nwarps = (uint64_t)gridDim.x * kScWarps;
    const uint64_t ntot = A.n_dev ? (uint64_t)min( (uint64_t)*A.n_dev, A.n ) : A.n;
    for ( uint64_t i = (uint64_t)blockIdx.x * kScWarps + ( threadIdx.x >> 5 ); i < ntot; i += nwarps ) {
      const double ekin = A.ekin[i];
      double xs = 0.0; int nent = 0;
      if ( domainContains( A.dom_lo, A.dom_hi, ekin ) && !( ekin <= S.threshold_ekin ) ) {
        Vec3 d = { A.ux[i], A.uy[i], A.uz[i] };
        vnormalise( d );
        ScAccum acc; double wl;
        scWalkWarp( S, ws, fam_of, ekin, d, wl, acc, 0, false, 0.0 );
        xs = acc.commul_last; nent = acc.n;
      }
      if ( lane == 0 ) { A.sc_xs[i] = xs; A.sc_n[i] = nent; }
    }
  }

  // ---- two-kernel form of the scan (default): k_sc_find walks the normals with a lean register footprint
  // (no evaluation code inlined: 6 instead of 2 CTAs per SM) and records, per neutron with at least one candidate
  // plane, the candidate list; k_sc_eval then evaluates and accumulates them with the code of scFlush.  Neutrons
  // with more than kScFindCap candidates are walked again by k_sc_eval with the combined scWalkWarp.
  constexpr int kScFindCap = 32;
#ifndef NCB_SC_FIND_WARPS
#  define NCB_SC_FIND_WARPS 20
#endif
  constexpr int kScFindWarps = NCB_SC_FIND_WARPS;   // more warps per CTA share one copy of the staged tables
  struct ScFindScratch {
    float2 lohi[kScMaxFam];           // window of |normal . direction| per family (x: lower, y: upper limit)
    double cptsq[kScMaxFam];
    double spt[kScMaxFam];
    uint16_t cand[kScFindCap];
  };
  struct ScFindArgs {
    const double* ekin; const double* ux; const double* uy; const double* uz;
    uint64_t n;
    double dom_lo, dom_hi;
    double* sc_xs; int32_t* sc_n;      // zeroed here for neutrons without candidates
    uint32_t* work; uint32_t* work_count; uint8_t* ncand; uint16_t* cand;   // work list (bit 31: overflow)
    uint32_t* overflow_count;          // number of work items flagged "overflow"
    int32_t* wpos;                     // per neutron: its work-list position, -1 none, -2 overflow (for k_sc_sample)
  };

  __global__ void __launch_bounds__(32*kScFindWarps, 2)
  k_sc_find( const __grid_constant__ Material M, const __grid_constant__ StagePlan sp,
             const __grid_constant__ ScFindArgs A, uint32_t fam_of_off, uint32_t scratch_off )
  {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar;
    // staged: the single-precision normal records (TMA bulk copy); everything else of the crystal is read through L1
    HotTabs H;
    stageHotTabs( M, sp, smem, &mbar, H );
    const ScBraggT& S = M.sc;
    const float4* nrec = reinterpret_cast<const float4*>( smem + sp.off[kHotSlotsIso] );
    ScFindScratch& ws = reinterpret_cast<ScFindScratch*>( smem + scratch_off )[ threadIdx.x >> 5 ];
    const int lane = threadIdx.x & 31;
    const uint64_t nwarps = (uint64_t)gridDim.x * kScFindWarps;
    const double cta = S.cta;
    double scratch_wl = -1.0;       // wavelength the windows in the scratch were computed for (warp-uniform)
    int nfam_act = 0;
    // The warp takes its neutrons 32 at a time: the part of a neutron's set-up that depends on its energy alone
    // (wavelength, d-spacing cut: a division, a square root and another division in double precision) is done by
    // ONE LANE PER NEUTRON for the whole batch and handed out with shuffles -- with the whole warp repeating it for
    // each neutron it was a fifth of the kernel's instructions.
    // Batches are 32 CONSECUTIVE neutrons (coalesced reads of the energies, neighbouring direction reads hit the
    // same sectors), dealt out to the warps round-robin.
    for ( uint64_t ib = 32*( (uint64_t)blockIdx.x * kScFindWarps + ( threadIdx.x >> 5 ) ); ib < A.n; ib += 32*nwarps ) {
      double wl_mine = 0.0, cut_mine = 0.0;      // wl == 0: nothing to search for this neutron
      int count_mine = 0;                        // lane k: number of candidates of the batch's k-th neutron
      const uint64_t im = ib + lane;
      {
        if ( im < A.n ) {
          const double ekin_raw = A.ekin[im];
          if ( domainContains( A.dom_lo, A.dom_hi, ekin_raw ) && !( ekin_raw <= S.threshold_ekin ) ) {
            const double ekin = scCacheRound( ekin_raw );
            wl_mine = ekin ? sqrt( kWl2Ekin / ekin ) : kInf;
            if ( wl_mine != 0 ) cut_mine = ( 1.0 - 2*kDblEps )/wl_mine;
          }
        }
      }
    for ( int kb = 0; kb < 32; ++kb ) {
      const uint64_t i = ib + kb;
      if ( i >= A.n ) break;
      const double wl = __shfl_sync( 0xffffffffu, wl_mine, kb );
      int count = 0;
      bool overflow = false;
      {
        if ( wl != 0 ) {
          Vec3 d = { A.ux[i], A.uy[i], A.uz[i] };
          vnormalise( d );
          // per-family windows: functions of the wavelength alone, kept in the warp's scratch while consecutive neutrons
          // have the same energy (a mono-energetic beam in a transport run: every neutron of the first steps)
          if ( !( wl == scratch_wl ) ) {
            const double inv2dcutoff = __shfl_sync( 0xffffffffu, cut_mine, kb );
            nfam_act = 0;
            for ( int f0 = 0; f0 < S.nfam; f0 += 32 ) {
              const int f = f0 + lane;
              const bool act = ( f < S.nfam ) && ( S.fam_inv2d[f] < inv2dcutoff );
              if ( act ) {
                InteractionPars ip;
                ip.set( wl, S.fam_inv2d[f], S.fam_xsfact[f] );
                const double spt = ip.sin_perfect_theta, cpt = sqrt( ip.cos_perfect_theta_sq );
                const double slo = spt*cta - cpt*S.sta, shi = spt*cta + cpt*S.sta;
                const bool open_hi = !( cpt*cta - spt*S.sta > 1e-6 );
                // (the window is compared with a float dot product: widened by 4e-6)
                ws.lohi[f] = make_float2( (float)( slo - 4e-6 ), open_hi ? 2.0f : (float)( shi + 4e-6 ) );
                ws.cptsq[f] = ip.cos_perfect_theta_sq;
                ws.spt[f] = spt;
              }
              const uint32_t m = __ballot_sync( 0xffffffffu, act );
              nfam_act += __popc( m );
              if ( m != 0xffffffffu ) break;
            }
            scratch_wl = wl;
            __syncwarp();
          }
          const int n_act = nfam_act ? S.fam_first[nfam_act] : 0;
          const float dxf = (float)d.x, dyf = (float)d.y, dzf = (float)d.z;
          // Four 32-normal slices per pass.  The pre-filter runs in single precision on the packed records (one
          // 16-byte and one 8-byte shared-memory load, three multiply-adds and two compares per normal; it only has
          // to be a superset of the exact test).  Planes pass it rarely, so the pass votes ONCE on "any lane has a
          // survivor"; only then the reference's test runs in double precision from the fp64 normals and the
          // candidates are compacted in index order.  (r2: 32 -> ~9 warp instructions per normal tested.)  The last
          // pass may run over slots beyond n_act (families that are not active, whose windows in the scratch are stale,
          // or padding): the pre-filter does not test for that -- four compares per pass -- the exact test does.
          for ( int base = 0; base < n_act; base += 128 ) {
            bool pre[4];
            int fam[4];
#pragma unroll
            for ( int u = 0; u < 4; ++u ) {
              const int in = base + 32*u + lane;       // (records are padded to a multiple of 128: always readable)
              const float4 r = nrec[in];
              fam[u] = __float_as_int( r.w );
              const float2 w = ws.lohi[fam[u]];
              const float xf = fabsf( __fmaf_rn( r.x, dxf, __fmaf_rn( r.y, dyf, r.z*dzf ) ) );
              pre[u] = ( xf > w.x ) & ( xf < w.y );      // (slots beyond n_act in the last pass: sorted out below)
            }
            if ( !__any_sync( 0xffffffffu, pre[0] | pre[1] | pre[2] | pre[3] ) ) continue;
#pragma unroll
            for ( int u = 0; u < 4; ++u ) {
              const int in = base + 32*u + lane;
              bool c = false;
              if ( pre[u] && in < n_act ) {
                const int f = fam[u];
                const double dot = S.normals[3*in]*d.x + S.normals[3*in+1]*d.y + S.normals[3*in+2]*d.z;
                double sd, ds;
                c = scIsCandidate( cta, ws.cptsq[f], ws.spt[f], dot, sd, ds );
              }
              const uint32_t m = __ballot_sync( 0xffffffffu, c );
              if ( m ) {
                const int pos = count + __popc( m & ( ( 1u << lane ) - 1u ) );
                if ( c && pos < kScFindCap ) ws.cand[pos] = (uint16_t)in;
                count += __popc( m );
              }
            }
            if ( count > kScFindCap ) { overflow = true; break; }
          }
          __syncwarp();
        }
      }
      if ( lane == kb ) count_mine = count;
      if ( count != 0 ) {
        uint32_t pos = 0;
        if ( lane == 0 ) pos = atomicAdd( A.work_count, 1u );
        pos = __shfl_sync( 0xffffffffu, pos, 0 );
        if ( lane == 0 ) {
          A.work[pos] = (uint32_t)i | ( overflow ? 0x80000000u : 0u );
          if ( overflow ) atomicAdd( A.overflow_count, 1u );
          A.ncand[pos] = (uint8_t)( overflow ? 0 : count );
          A.wpos[i] = overflow ? -2 : (int32_t)pos;
        }
        if ( !overflow && lane < count )
          A.cand[(size_t)pos*kScFindCap + lane] = ws.cand[lane];
      }
      __syncwarp();
    }
      // neutrons without candidates: zeroed by the lane that owns them, one coalesced store per array and batch
      if ( im < A.n && count_mine == 0 ) { A.sc_xs[im] = 0.0; A.sc_n[im] = 0; A.wpos[im] = -1; }
    }
  }

  // Evaluation of the recorded candidates, EIGHT LANES PER NEUTRON (four neutrons per warp): a neutron has 1-3
  // candidate planes on average, so with a warp per neutron 29 of 32 lanes idled through the circle integrals (ncu:
  // 13 of 32 lanes active, 24 % occupancy at 128 registers).  The lanes of a group evaluate the group's candidates
  // in parallel (raw cross sections of -normal / +normal), then the group accumulates them in plane order exactly as
  // scFlush does.  Work items flagged "overflow" (more than kScFindCap candidates) are left to k_sc_eval below.
  // dynamic smem: [staged tables (sp.total)] [fam_of: nnormals bytes] [kScWarps x 4 x ScGroupScratch]
  struct ScGroupScratch { double vals[2*kScFindCap]; };
  __global__ void __launch_bounds__(32*kScWarps, 2)
  k_sc_eval_groups( const __grid_constant__ Material M, const __grid_constant__ StagePlan sp,
                    const __grid_constant__ ScFindArgs A, uint32_t fam_of_off, uint32_t scratch_off )
  {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar;
    HotTabs H;
    uint8_t* fam_of = smem + fam_of_off;
    scBlockSetup( M, sp, smem, &mbar, H, fam_of );
    const ScBraggT& S = *H.sc;
    const int lane = threadIdx.x & 31, grp = lane >> 3, sl = lane & 7;
    ScGroupScratch& gs = reinterpret_cast<ScGroupScratch*>( smem + scratch_off )[ ( threadIdx.x >> 5 )*4 + grp ];
    const uint32_t nwork = *A.work_count;
    const uint32_t nitems = gridDim.x * kScWarps * 4;
    for ( uint32_t w0 = ( blockIdx.x * kScWarps + ( threadIdx.x >> 5 ) )*4; w0 < nwork; w0 += nitems ) {
      const uint32_t w = w0 + grp;
      uint32_t entry = 0x80000000u;
      int count = 0;
      if ( w < nwork ) { entry = A.work[w]; if ( !( entry & 0x80000000u ) ) count = A.ncand[w]; }
      const uint32_t i = entry & 0x7fffffffu;
      Vec3 d = { 0, 0, 1 };
      double wl = 0.0;
      if ( count ) {
        d = Vec3{ A.ux[i], A.uy[i], A.uz[i] };
        vnormalise( d );
        const double ekin = scCacheRound( A.ekin[i] );
        wl = ekin ? sqrt( kWl2Ekin / ekin ) : kInf;
      }
      const int maxcount = __reduce_max_sync( 0xffffffffu, count );
      const uint16_t* cand = A.cand + (size_t)w*kScFindCap;
      for ( int k0 = 0; k0 < maxcount; k0 += 8 ) {
        const int k = k0 + sl;
        if ( k < count ) {
          const int in = cand[k];
          const int f = fam_of[in];
          InteractionPars ip;
          ip.set( wl, S.fam_inv2d[f], S.fam_xsfact[f] );
          const double nx = S.normals[3*in], ny = S.normals[3*in+1], nz = S.normals[3*in+2];
          const double dot = nx*d.x + ny*d.y + nz*d.z;
          const double sdotcptsq = ( 1.0 - dot*dot )*ip.cos_perfect_theta_sq;
          const double ds = dot * ip.sin_perfect_theta;
          double xm = 0.0, xp = 0.0;
          const double Am = dmax( 0.0, S.cta - ds );
          if ( sdotcptsq > Am*Am ) xm = gmRawXS( S, ip, dot );     // anti-normal
          const double Ap = dmax( 0.0, S.cta + ds );
          if ( sdotcptsq > Ap*Ap ) xp = gmRawXS( S, ip, -dot );    // normal
          gs.vals[2*k] = xm; gs.vals[2*k+1] = xp;
        }
      }
      __syncwarp();
      if ( count && sl == 0 ) {
        // ordered accumulation (scFlush, mode 0)
        int cur_fam = -1, n = 0;
        double xsoffset = 0.0, xssum = 0.0, commul_last = 0.0;
        for ( int k = 0; k < count; ++k ) {
          const int f = fam_of[ cand[k] ];
          if ( f != cur_fam ) { cur_fam = f; xsoffset = commul_last; xssum = 0.0; }
          for ( int sgn = 0; sgn < 2; ++sgn ) {
            const double xs = gs.vals[2*k+sgn];
            if ( xs ) { commul_last = xsoffset + ( xssum += xs ); ++n; }
          }
        }
        A.sc_xs[i] = commul_last; A.sc_n[i] = n;
      }
      __syncwarp();
    }
  }

  // Evaluation of the recorded candidates, ONE CANDIDATE PER LANE.  With eight lanes per neutron (k_sc_eval_groups
  // above) the raw cross sections -- the bulk of the kernel -- still ran with 2.8 of 32 lanes active (source-level
  // profile: a neutron has 1-3 candidates, seven of the eight lanes of its group idled).  Here a warp takes 32 work
  // items at a time: lane j prepares neutron j (direction, wavelength) into the warp's scratch, the candidates of
  // all 32 neutrons are numbered consecutively (prefix sum of the counts) and evaluated 32 at a time, every lane
  // its own candidate of whatever neutron; then lane j accumulates neutron j's values in plane order exactly as
  // scFlush does.  A round holds kScFlatVals candidates; a batch with more (rare) takes several rounds.
  // dynamic smem: [staged tables (sp.total)] [fam_of: nnormals bytes] [kScWarps x ScFlatScratch]
  constexpr int kScFlatVals = 256;
  struct ScFlatScratch {
    double par[32][4];                 // per work item of the batch: normalised direction, wavelength
    double vals[2*kScFlatVals];        // raw cross sections (anti-normal, normal) per candidate slot of the round
    int off[32];                       // first slot of each item of the round
  };
  __global__ void __launch_bounds__(32*kScWarps, 2)
  k_sc_eval_flat( const __grid_constant__ Material M, const __grid_constant__ StagePlan sp,
                  const __grid_constant__ ScFindArgs A, uint32_t fam_of_off, uint32_t scratch_off )
  {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar;
    HotTabs H;
    uint8_t* fam_of = smem + fam_of_off;
    scBlockSetup( M, sp, smem, &mbar, H, fam_of );
    const ScBraggT& S = *H.sc;
    const int lane = threadIdx.x & 31;
    ScFlatScratch& fs = reinterpret_cast<ScFlatScratch*>( smem + scratch_off )[ threadIdx.x >> 5 ];
    const uint32_t nwork = *A.work_count;
    const uint32_t stride = gridDim.x * kScWarps * 32;
    for ( uint32_t w0 = ( blockIdx.x * kScWarps + ( threadIdx.x >> 5 ) )*32; w0 < nwork; w0 += stride ) {
      const uint32_t w = w0 + lane;
      uint32_t entry = 0x80000000u;
      int count = 0;
      if ( w < nwork ) { entry = A.work[w]; if ( !( entry & 0x80000000u ) ) count = A.ncand[w]; }
      const uint32_t i = entry & 0x7fffffffu;
      __syncwarp();     // (the scratch is reused from the previous batch)
      if ( count ) {
        Vec3 d = { A.ux[i], A.uy[i], A.uz[i] };
        vnormalise( d );
        const double ekin = scCacheRound( A.ekin[i] );
        fs.par[lane][0] = d.x; fs.par[lane][1] = d.y; fs.par[lane][2] = d.z;
        fs.par[lane][3] = ekin ? sqrt( kWl2Ekin / ekin ) : kInf;
      }
      int incl = count;      // inclusive prefix sum of the counts over the lanes
      #pragma unroll
      for ( int dd = 1; dd < 32; dd <<= 1 ) {
        const int v = __shfl_up_sync( 0xffffffffu, incl, dd );
        if ( lane >= dd ) incl += v;
      }
      int start = 0, base_slots = 0;
      while ( start < 32 ) {
        // items [start, end) of the batch: as many as fit the slot table (the prefix sum does not decrease, so the
        // lanes that fit form one run from `start`; a single item always fits: count <= kScFindCap)
        const uint32_t fits = __ballot_sync( 0xffffffffu, lane >= start && incl - base_slots <= kScFlatVals );
        const int end = start + __popc( fits );
        const int total = __shfl_sync( 0xffffffffu, incl, end - 1 ) - base_slots;
        const bool mine = lane >= start && lane < end;
        if ( mine ) fs.off[lane] = incl - count - base_slots;
        __syncwarp();
        for ( int c0 = 0; c0 < total; c0 += 32 ) {
          const int c = c0 + lane;
          if ( c < total ) {
            // owner: the last item of the round whose first slot is <= c (items without candidates share the
            // first slot of their successor and are passed over)
            int lo = start, hi = end - 1;
            while ( lo < hi ) {
              const int mid = ( lo + hi + 1 ) >> 1;
              if ( fs.off[mid] <= c ) lo = mid; else hi = mid - 1;
            }
            const int k = c - fs.off[lo];
            const int in = A.cand[ (size_t)( w0 + lo )*kScFindCap + k ];
            const int f = fam_of[in];
            const double* P = fs.par[lo];
            InteractionPars ip;
            ip.set( P[3], S.fam_inv2d[f], S.fam_xsfact[f] );
            const double nx = S.normals[3*in], ny = S.normals[3*in+1], nz = S.normals[3*in+2];
            const double dot = nx*P[0] + ny*P[1] + nz*P[2];
            const double sdotcptsq = ( 1.0 - dot*dot )*ip.cos_perfect_theta_sq;
            const double ds = dot * ip.sin_perfect_theta;
            double xm = 0.0, xp = 0.0;
            const double Am = dmax( 0.0, S.cta - ds );
            const double Ap = dmax( 0.0, S.cta + ds );
            const bool pm = sdotcptsq > Am*Am;      // anti-normal contributes
            const bool pp = sdotcptsq > Ap*Ap;      // normal contributes
            // (usually one of the two: the lanes make their first -- mostly only -- evaluation together, whichever
            //  side it is for; the value does not depend on which side is evaluated first)
            if ( pm || pp ) {
              const double x1 = gmRawXS( S, ip, pm ? dot : -dot );
              if ( pm ) xm = x1; else xp = x1;
            }
            if ( pm && pp ) xp = gmRawXS( S, ip, -dot );
            fs.vals[2*c] = xm; fs.vals[2*c+1] = xp;
          }
        }
        __syncwarp();
        if ( mine && count ) {
          // ordered accumulation (scFlush, mode 0)
          const uint16_t* cand = A.cand + (size_t)w*kScFindCap;
          const double* v = fs.vals + 2*fs.off[lane];
          int cur_fam = -1, n = 0;
          double xsoffset = 0.0, xssum = 0.0, commul_last = 0.0;
          for ( int k = 0; k < count; ++k ) {
            const int f = fam_of[ cand[k] ];
            if ( f != cur_fam ) { cur_fam = f; xsoffset = commul_last; xssum = 0.0; }
            for ( int sgn = 0; sgn < 2; ++sgn ) {
              const double xs = v[2*k+sgn];
              if ( xs ) { commul_last = xsoffset + ( xssum += xs ); ++n; }
            }
          }
          A.sc_xs[i] = commul_last; A.sc_n[i] = n;
        }
        base_slots += total;
        start = end;
        __syncwarp();
      }
    }
  }

  // warp-per-neutron evaluation: all work items (only_overflow = 0), or just the ones k_sc_eval_groups leaves
  __global__ void __launch_bounds__(32*kScWarps, 2)
  k_sc_eval( const __grid_constant__ Material M, const __grid_constant__ StagePlan sp,
             const __grid_constant__ ScFindArgs A, uint32_t fam_of_off, uint32_t scratch_off, int only_overflow )
  {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar;
    if ( only_overflow && *A.overflow_count == 0 ) return;
    HotTabs H;
    uint8_t* fam_of = smem + fam_of_off;
    scBlockSetup( M, sp, smem, &mbar, H, fam_of );
    const ScBraggT& S = *H.sc;
    ScWarpScratch& ws = reinterpret_cast<ScWarpScratch*>( smem + scratch_off )[ threadIdx.x >> 5 ];
    const int lane = threadIdx.x & 31;
    const uint32_t nwork = *A.work_count;
    const uint32_t nwarps = gridDim.x * kScWarps;
    for ( uint32_t w = blockIdx.x * kScWarps + ( threadIdx.x >> 5 ); w < nwork; w += nwarps ) {
      const uint32_t entry = A.work[w];
      if ( only_overflow && !( entry & 0x80000000u ) ) continue;
      const uint32_t i = entry & 0x7fffffffu;
      Vec3 d = { A.ux[i], A.uy[i], A.uz[i] };
      vnormalise( d );
      ScAccum acc; double wl;
      if ( entry & 0x80000000u ) {
        scWalkWarp( S, ws, fam_of, A.ekin[i], d, wl, acc, 0, false, 0.0 );
      } else {
        acc.cur_fam = -1; acc.n = 0; acc.xsoffset = acc.xssum = acc.commul_last = 0.0;
        acc.found = false; acc.chosen_in = 0; acc.chosen_sign = 1;
        const double ekin = scCacheRound( A.ekin[i] );
        wl = ekin ? sqrt( kWl2Ekin / ekin ) : kInf;
        const int count = A.ncand[w];
        if ( lane < count ) ws.cand[lane] = A.cand[(size_t)w*kScFindCap + lane];
        __syncwarp();
        scFlush( S, ws, fam_of, wl, d, count, acc, 0, false, 0.0 );
      }
      if ( lane == 0 ) { A.sc_xs[i] = acc.commul_last; A.sc_n[i] = acc.n; }
      __syncwarp();
    }
  }

  // ---- thread-per-neutron kernels that consume the scan result
  struct AnisoArgs {
    DirArgs D;
    const double* sc_xs; const int32_t* sc_n;   // from k_sc_scan (null for materials without SCBragg)
    double* mu_tmp;                              // scratch: mu of the isotropic queue kernels
    uint32_t* nd_tmp;                            // scratch: stream position after isotropic sampling
    uint32_t* q_sc; uint32_t* q_sc_count;        // neutrons whose chosen component is SCBragg
    // candidate lists recorded by k_sc_find (null: walk the normals again)
    const int32_t* sc_wpos = nullptr; const uint8_t* sc_ncand = nullptr; const uint16_t* sc_cand = nullptr;
  };

  // total xs with precomputed SCBragg part (matXS, ncb_proc.cuh)
  __device__ __forceinline__ double matXSPre( const Material& M, const HotTabs& H, double ekin, double sc_xs, int sc_n,
                                              double* cumul, int* aux )
  {
    if ( !domainContains( M.dom_lo, M.dom_hi, ekin ) )
      return 0.0;
    double tot = 0.0;
    for ( int i = 0; i < M.ncomp; ++i ) {
      const Comp& c = M.comp[i];
      int a = -1;
      double xs = 0.0;
      if ( domainContains( c.dom_lo, c.dom_hi, ekin ) ) {
        if ( c.kind == KIND_SCBRAGG ) { xs = sc_xs; a = sc_n; }
        else if ( c.kind == KIND_LCBRAGG ) { xs = sc_n ? M.lc.xsfact * sc_xs : 0.0; a = sc_n; }   // (sc_xs: sum over the ROIs, k_lc_scan)
        else xs = compXSIso( M, H, i, ekin, a );
      }
      tot += c.scale * xs;
      if ( cumul ) cumul[i] = tot;
      if ( aux ) aux[i] = a;
    }
    return tot;
  }

  __global__ void __launch_bounds__(256)
  k_xs_aniso_pre( const __grid_constant__ Material M, const __grid_constant__ StagePlan sp,
                  const double* __restrict__ ekin, const double* __restrict__ sc_xs, const int32_t* __restrict__ sc_n,
                  uint64_t n, double* __restrict__ out, const uint32_t* __restrict__ n_dev )
  {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar;
    HotTabs H;
    stageHotTabs( M, sp, smem, &mbar, H );
    if ( n_dev ) n = min( (uint64_t)*n_dev, n );
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for ( uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride )
      out[i] = matXSPre( M, H, ekin[i], sc_xs ? sc_xs[i] : 0.0, sc_n ? sc_n[i] : 0, nullptr, nullptr );
  }

  __global__ void __launch_bounds__(256)
  k_classify_aniso( const __grid_constant__ Material M, const __grid_constant__ StagePlan sp,
                    const __grid_constant__ SampleArgs A, const __grid_constant__ QueueArgs Q,
                    const __grid_constant__ AnisoArgs X )
  {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar;
    HotTabs H;
    stageHotTabs( M, sp, smem, &mbar, H );
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    int errs = 0;
    const uint64_t ntot = A.count();
    for ( uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < ntot; base += stride ) {
      const uint64_t i = base + threadIdx.x;
      int cls = 0;            // 0: done here, 1: SAB table queue, 2: free-gas queue, 3: SCBragg queue
      uint32_t entry = 0;
      if ( i < ntot ) {
        const double ekin = A.ekin[i];
        const Vec3 dir = { X.D.ux[i], X.D.uy[i], X.D.uz[i] };
        double eout = ekin, tot = 0.0;
        Vec3 o = dir;
        int ich = -1;
        uint32_t nd = 0;
        if ( domainContains( M.dom_lo, M.dom_hi, ekin ) ) {
          double cumul[kMaxComp];
          int aux[kMaxComp];
          tot = matXSPre( M, H, ekin, X.sc_xs ? X.sc_xs[i] : 0.0, X.sc_n ? X.sc_n[i] : 0, cumul, aux );
          Rng rng; rng.init( A.seed, A.streamIndex( i ), A.sid );
          ich = ( M.ncomp == 1 ? 0 : pickIdxByWeight( rng.generate(), cumul, M.ncomp ) );
          const Comp& c = M.comp[ich];
          if ( c.kind == KIND_SCBRAGG ) {
            // no-scatter cases of SCBragg::sampleScatter (NCSCBragg.cc:306-318) finish here
            if ( !( ekin <= M.sc.threshold_ekin ) && aux[ich] > 0 && X.sc_xs[i] > 0.0 )
              cls = 3;
          } else if ( c.kind == KIND_LCBRAGG ) {
            // no-scatter cases of LCBragg::sampleScatter / LCHelper::genScatter (NCLCBragg.cc:129-136,
            // NCLCUtils.cc:539-544) finish here: below the Bragg threshold the direction is returned as given,
            // without ROIs it is returned normalised
            LcNeutron N;
            const Vec3 u = vunit( dir );
            if ( !( ekin < M.lc.ekin_low ) && lcNeutronPars( M.lc, ekin, u, N ) ) {
              o = u;
              if ( aux[ich] > 0 && X.sc_xs[i] != 0.0 )
                cls = 3;
            }
          } else if ( c.kind == KIND_SAB ) {
            const SabT& T = M.sab[c.idx];
            cls = ( ekin < H.sab_egrid[c.idx][T.negrid-1] ) ? 1 : 2;
          } else if ( c.kind == KIND_FREEGAS ) {
            cls = 2;
          } else {
            double mu = 1.0;
            if ( c.kind == KIND_POWDERBRAGG ) {
              const PowderBraggT& T = M.pb[c.idx];
              if ( !( ekin < T.threshold || !isFinite(ekin) ) ) {
                const int iv = aux[ich] >= 0 ? aux[ich] : pbLastValidPlane( T, H.pb_e2d[c.idx], H.pb_lut[c.idx], ekin );
                mu = pbSampleMu( H.pb_e2d[c.idx], H.pb_fdm[c.idx], iv, ekin, rng );
              }
            } else if ( c.kind == KIND_ELINC ) {
              mu = elincSampleMu( M.elinc[c.idx], ekin, rng );
            }
            o = randDirectionGivenScatterMu( rng, mu, dir );
          }
          nd = rng.ndraws;
          entry = (uint32_t)i | ( (uint32_t)ich << kQueueIdxBits );
        }
        if ( A.xs_out ) A.xs_out[i] = tot;
        if ( A.component ) A.component[i] = ich;
        if ( cls == 0 ) {
          A.ekin_out[i] = eout;
          X.D.ox[i] = o.x; X.D.oy[i] = o.y; X.D.oz[i] = o.z;
          if ( A.ndraws ) A.ndraws[i] = nd;
        }
      }
      warpPush( cls == 1, Q.q_sab, Q.counts + 0, entry );
      warpPush( cls == 2, Q.q_fg, Q.counts + 1, entry );
      warpPush( cls == 3, X.q_sc, X.q_sc_count, entry );
    }
    if ( errs )
      atomicOr( A.err_flags, errs );
  }

  // One queued neutron per THREAD: the planes that can contribute were recorded by k_sc_find (1-3 on average), the
  // thread evaluates them in order until the entry picked by pickRandIdxByWeight is reached (same values, same order
  // and stop rule as scFlush in mode 1) and generates the scattering (GaussMos::genScat).  With a warp per neutron
  // (k_sc_sample below, r1) all 32 lanes repeated the thread-level part.  Neutrons without a recorded list (more
  // candidates than the record holds) are left to the warp-per-neutron kernel.
  __global__ void __launch_bounds__(32*kScWarps, 2)
  k_sc_sample_threads( const __grid_constant__ Material M, const __grid_constant__ StagePlan sp,
                       const __grid_constant__ SampleArgs A, const __grid_constant__ AnisoArgs X,
                       uint32_t fam_of_off, uint32_t* __restrict__ n_left )
  {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar;
    HotTabs H;
    uint8_t* fam_of = smem + fam_of_off;
    scBlockSetup( M, sp, smem, &mbar, H, fam_of );
    const ScBraggT& S = *H.sc;
    const uint32_t nq = *X.q_sc_count;
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t left = 0;
    for ( uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nq; j += stride ) {
      const uint32_t i = X.q_sc[j] & kQueueIdxMask;
      const int32_t wp = X.sc_wpos ? X.sc_wpos[i] : -2;
      if ( wp < 0 ) { ++left; continue; }
      const double ekin = A.ekin[i];
      Vec3 d = { X.D.ux[i], X.D.uy[i], X.D.uz[i] };
      vnormalise( d );
      Rng rng; rng.init( A.seed, A.streamIndex( i ), A.sid );
      rng.seek( M.ncomp > 1 ? 1u : 0u );
      const int nent = X.sc_n[i];
      const double total = X.sc_xs[i];
      double choice = -1.0; bool linear = true;
      if ( nent > 1 ) { choice = total * rng.generate(); linear = ( nent < 5 ); }
      const double ekr = scCacheRound( ekin );
      const double wl = ekr ? sqrt( kWl2Ekin / ekr ) : kInf;
      const int count = X.sc_ncand[wp];
      const uint16_t* cand = X.sc_cand + (size_t)wp*kScFindCap;
      int cur_fam = -1, chosen_in = 0, chosen_sign = 1;
      double xsoffset = 0.0, xssum = 0.0, commul_last = 0.0;
      bool found = false;
      for ( int k = 0; k < count && !found; ++k ) {
        const int in = cand[k];
        const int f = fam_of[in];
        InteractionPars ip;
        ip.set( wl, S.fam_inv2d[f], S.fam_xsfact[f] );
        const double dot = S.normals[3*in]*d.x + S.normals[3*in+1]*d.y + S.normals[3*in+2]*d.z;
        const double sdotcptsq = ( 1.0 - dot*dot )*ip.cos_perfect_theta_sq;
        const double ds = dot * ip.sin_perfect_theta;
        if ( f != cur_fam ) { cur_fam = f; xsoffset = commul_last; xssum = 0.0; }
        for ( int sgn = 0; sgn < 2 && !found; ++sgn ) {
          const double Aa = dmax( 0.0, sgn ? S.cta + ds : S.cta - ds );
          double xs = 0.0;
          if ( sdotcptsq > Aa*Aa ) xs = gmRawXS( S, ip, sgn ? -dot : dot );     // sgn 0: anti-normal, 1: normal
          if ( xs ) {
            commul_last = xsoffset + ( xssum += xs );
            chosen_in = in; chosen_sign = sgn;
            if ( linear ? ( commul_last > choice ) : !( commul_last < choice ) ) found = true;
          }
        }
      }
      const double sg = chosen_sign ? 1.0 : -1.0;
      const Vec3 pn = { sg*S.normals[3*chosen_in], sg*S.normals[3*chosen_in+1], sg*S.normals[3*chosen_in+2] };
      const double inv2dsp = gmCacheRound( S.fam_inv2d[ fam_of[chosen_in] ] );
      Vec3 o;
      gmGenScat( S, rng, pn, inv2dsp, wl, d, o );
      A.ekin_out[i] = ekin;
      X.D.ox[i] = o.x; X.D.oy[i] = o.y; X.D.oz[i] = o.z;
      if ( A.ndraws ) A.ndraws[i] = rng.ndraws;
    }
    if ( left ) atomicAdd( n_left, left );
  }

  // one queued neutron per warp: choose the normal (second walk) and generate the scattering.  only_unlisted: just the
  // neutrons k_sc_sample_threads left (no recorded candidate list); *n_left == 0: nothing to do
  __global__ void __launch_bounds__(32*kScWarps, 2)
  k_sc_sample( const __grid_constant__ Material M, const __grid_constant__ StagePlan sp,
               const __grid_constant__ SampleArgs A, const __grid_constant__ AnisoArgs X,
               uint32_t fam_of_off, uint32_t scratch_off, int only_unlisted, const uint32_t* __restrict__ n_left )
  {
    if ( only_unlisted && *n_left == 0 ) return;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar;
    HotTabs H;
    uint8_t* fam_of = smem + fam_of_off;
    scBlockSetup( M, sp, smem, &mbar, H, fam_of );
    const ScBraggT& S = *H.sc;
    ScWarpScratch& ws = reinterpret_cast<ScWarpScratch*>( smem + scratch_off )[ threadIdx.x >> 5 ];
    const int lane = threadIdx.x & 31;
    const uint32_t nq = *X.q_sc_count;
    const uint32_t nwarps = gridDim.x * kScWarps;
    for ( uint32_t j = blockIdx.x * kScWarps + ( threadIdx.x >> 5 ); j < nq; j += nwarps ) {
      const uint32_t i = X.q_sc[j] & kQueueIdxMask;
      if ( only_unlisted && X.sc_wpos && X.sc_wpos[i] >= 0 ) continue;
      const double ekin = A.ekin[i];
      Vec3 d = { X.D.ux[i], X.D.uy[i], X.D.uz[i] };
      vnormalise( d );
      Rng rng; rng.init( A.seed, A.streamIndex( i ), A.sid );
      rng.seek( M.ncomp > 1 ? 1u : 0u );
      const int nent = X.sc_n[i];
      const double total = X.sc_xs[i];
      // pickRandIdxByWeight over xs_commul (NCSCBragg.cc:283; NCRandUtils.cc:198-220)
      double choice = -1.0; bool linear = true;
      if ( nent > 1 ) { choice = total * rng.generate(); linear = ( nent < 5 ); }
      ScAccum acc; double wl;
      const int32_t wp = X.sc_wpos ? X.sc_wpos[i] : -2;
      if ( wp >= 0 ) {
        // the planes that can contribute were recorded by k_sc_find: select among them (same order, same values)
        acc.cur_fam = -1; acc.n = 0; acc.xsoffset = acc.xssum = acc.commul_last = 0.0;
        acc.found = false; acc.chosen_in = 0; acc.chosen_sign = 1;
        const double ekr = scCacheRound( ekin );
        wl = ekr ? sqrt( kWl2Ekin / ekr ) : kInf;
        const int count = X.sc_ncand[wp];
        if ( lane < count ) ws.cand[lane] = X.sc_cand[(size_t)wp*kScFindCap + lane];
        __syncwarp();
        scFlush( S, ws, fam_of, wl, d, count, acc, 1, linear, choice );
      } else {
        scWalkWarp( S, ws, fam_of, ekin, d, wl, acc, 1, linear, choice );
      }
      const int in = acc.chosen_in;
      const double sg = acc.chosen_sign ? 1.0 : -1.0;   // vals[2k] = anti-normal, vals[2k+1] = normal
      const Vec3 pn = { sg*S.normals[3*in], sg*S.normals[3*in+1], sg*S.normals[3*in+2] };
      const double inv2dsp = gmCacheRound( S.fam_inv2d[ fam_of[in] ] );   // ip.m_inv2dsp (NCGaussMos.cc:258)
      Vec3 o;
      gmGenScat( S, rng, pn, inv2dsp, wl, d, o );
      if ( lane == 0 ) {
        A.ekin_out[i] = ekin;
        X.D.ox[i] = o.x; X.D.oy[i] = o.y; X.D.oz[i] = o.z;
        if ( A.ndraws ) A.ndraws[i] = rng.ndraws;
      }
    }
  }

  // direction for neutrons sampled by the isotropic queue kernels (entries of q_sab and q_fg)
  __global__ void __launch_bounds__(256)
  k_dir_from_mu( const __grid_constant__ SampleArgs A, const __grid_constant__ QueueArgs Q, const __grid_constant__ AnisoArgs X )
  {
    const uint32_t n0 = Q.counts[0], n1 = Q.counts[1];
    const uint32_t stride = gridDim.x * blockDim.x;
    for ( uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n0 + n1; j += stride ) {
      const uint32_t i = ( j < n0 ? Q.q_sab[j] : Q.q_fg[j - n0] ) & kQueueIdxMask;
      Vec3 o = { 0.0, 0.0, 0.0 };
      if ( A.ekin_out[i] >= 0.0 ) {     // (-1: the sampler raised an error for this neutron)
        Rng rng; rng.init( A.seed, A.streamIndex( i ), A.sid );
        rng.seek( X.nd_tmp[i] );
        o = randDirectionGivenScatterMu( rng, X.mu_tmp[i], Vec3{ X.D.ux[i], X.D.uy[i], X.D.uz[i] } );
        if ( A.ndraws ) A.ndraws[i] = rng.ndraws;
      }
      X.D.ox[i] = o.x; X.D.oy[i] = o.y; X.D.oz[i] = o.z;
    }
  }

}
