// ncb_rng.cuh -- per-neutron counter-based random streams (Philox4x32-10).
//
// The reference draws from ONE sequential xoroshiro128+ stream per handle
// (ref: NCRandUtils.hh:208-245, NCDefs.hh:1276), with data-dependent draw
// counts -- not reproducible in parallel.  Here every neutron owns a stream
//     key = 64-bit seed, counter = (global neutron index, block number, stream id)
// and its k'th uniform is word (k&1) of Philox block (k>>1), mapped to (0,1]
// exactly like the reference maps its 64 random bits (randUInt64ToFP01,
// NCDefs.hh:1308-1330).  Results are therefore independent of launch shape and
// GPU count, the state is (seed, index, k), and the oracle replays the same
// numbers through the reference's RNG hook (oracle/philox_ref.h).
#pragma once
#include "ncb_common.cuh"

namespace ncb {

  NCB_HD uint32_t mulhi32( uint32_t a, uint32_t b )
  {
#if defined(__CUDA_ARCH__)
    return __umulhi( a, b );
#else
    return (uint32_t)( ( (uint64_t)a * b ) >> 32 );
#endif
  }

#if defined(NCB_RNG_REPLAY)
  // Caller-supplied numbers instead of a Philox stream (ncb_replay.cu: ncrystal_samplescatter_rs and the virtual
  // API hand the library a generator that must be consumed draw by draw).  The first `nu` numbers come from `u`;
  // asking for more sets `overrun` -- the host then fetches one more number from the caller and repeats the
  // neutron -- and returns throw-away values from a small generator so that rejection loops still terminate.
  struct Rng {
    const double* u;
    uint32_t nu;
    uint32_t ndraws;
    uint32_t overrun;
    uint32_t lcg;
    NCB_HD void init( uint64_t, uint64_t, uint32_t = 0 ) { u = nullptr; nu = 0; ndraws = 0; overrun = 0; lcg = 12345u; }
    NCB_HD void seek( uint32_t k ) { ndraws = k; }
    NCB_HD double generate()
    {
      const uint32_t k = ndraws++;
      if ( k < nu )
        return u[k];
      overrun = 1;
      lcg = lcg*1664525u + 1013904223u;
      return ( (double)( lcg >> 8 ) + 0.5 ) * ( 1.0/16777216.0 );
    }
  };
#else
  struct Rng {
    uint32_t k0, k1;   // key  (seed)
    uint32_t c0, c1;   // counter words 0,1 (neutron index)
    uint32_t sid;      // counter word 3: stream id (0 for a fresh handle, k for its k'th clone)
    uint32_t ndraws;   // uniforms consumed
    uint32_t b2, b3;   // second half of the current block

    NCB_HD void init( uint64_t seed, uint64_t index, uint32_t stream_id = 0 )
    {
      sid = stream_id;
      k0 = (uint32_t)seed; k1 = (uint32_t)( seed >> 32 );
      c0 = (uint32_t)index; c1 = (uint32_t)( index >> 32 );
      ndraws = 0; b2 = b3 = 0;
    }

    // Position the stream so that the next generate() returns uniform number k.
    NCB_HD void seek( uint32_t k )
    {
      if ( k & 1u ) { ndraws = k - 1u; generate(); }  // refill the second half of block k>>1
      else ndraws = k;
    }

    NCB_HD static double toFP01( uint32_t lo, uint32_t hi )
    {
      // x = hi:lo ; r1 = (x>>11)*2^-53 ; r2 = (x&0x7FF)*2^-64 ; (1-r1)-r2
      const uint64_t x = ( (uint64_t)hi << 32 ) | lo;
      const double r1 = (double)( x >> 11 ) * 0x1.0p-53;
      const double r2 = (double)( lo & 0x7FFu ) * 0x1.0p-64;
      return ( 1.0 - r1 ) - r2;
    }

    // One Philox4x32-10 block (counter = (c0, c1, block, sid)); kept out of line: generate() has
    // ~40 call sites on the sampling path and the 10 inlined rounds per site made those kernels
    // instruction-fetch bound (ncu: ~30% of the free-gas kernel's samples sat in inlined rounds).
    NCB_HD_NOINLINE double refill( uint32_t block )
    {
      uint32_t x0 = c0, x1 = c1, x2 = block, x3 = sid;
      uint32_t ka = k0, kb = k1;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for ( int r = 0; r < 10; ++r ) {
        const uint32_t hi0 = mulhi32( 0xD2511F53u, x0 ), lo0 = 0xD2511F53u * x0;
        const uint32_t hi1 = mulhi32( 0xCD9E8D57u, x2 ), lo1 = 0xCD9E8D57u * x2;
        const uint32_t n0 = hi1 ^ x1 ^ ka;
        const uint32_t n2 = hi0 ^ x3 ^ kb;
        x0 = n0; x1 = lo1; x2 = n2; x3 = lo0;
        ka += 0x9E3779B9u; kb += 0xBB67AE85u;
      }
      b2 = x2; b3 = x3;
      return toFP01( x0, x1 );
    }

    NCB_HD double generate()
    {
      const uint32_t k = ndraws++;
      if ( k & 1u )
        return toFP01( b2, b3 );
      return refill( k >> 1 );
    }
  };
#endif

}
