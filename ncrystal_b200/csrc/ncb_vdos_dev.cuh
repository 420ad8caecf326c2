// ncb_vdos_dev.cuh -- device side of the VDOS -> S(alpha,beta) expansion (host orchestration: ncb_vdos.h).
//
//   * FFT convolution of two phonon spectra, G_n = G_{n1} (x) G_{n2}: radix-2 decimation-in-time transform with
//     the reference's butterfly arithmetic and twiddle table (ref: src/utils/NCFastConvolve.cc:183-268,325-459,
//     466-564), so that every output bin equals the reference's bit for bit -- a stage's butterflies are
//     independent, so they run one per thread.  The first 12 stages of a transform work on 4096-point chunks held in
//     shared memory (one CTA per chunk, 64 KB), later stages (inputs longer than 4096 bins, i.e. orders 2-5 of a
//     finely binned VDOS) one launch per stage over global memory.  All orders of a batch run in the same launches
//     (blockIdx.y = job).
//   * k_vdos_finish: |.|*dt/N, truncation of the noise floor, thinning, normalisation to unit area and the statistics
//     the host's stopping rule needs (ref: src/vdos/NCVDOSGn.cc:62-90,437-477), one CTA per order.
//   * k_vdos_fill: S(alpha,beta) = sum over phonon orders of f(x_alpha,n) * G_n(beta kT), in the reference's
//     summation order (ref: src/vdos/NCVDOSToScatKnl.cc:83-325), one thread per (alpha, non-positive beta) cell, which
//     also owns the mirrored positive-beta cell.
#pragma once
#include "ncb_common.cuh"
#include <cstdint>

namespace ncb { namespace vdos {

  struct Cplx { double re, im; };

  // (a+ib)(c+id), evaluated like the compiler evaluates std::complex<double>::operator*= for finite operands
  NCB_HD Cplx cmul( Cplx x, Cplx y ) { return Cplx{ x.re*y.re - x.im*y.im, x.re*y.im + x.im*y.re }; }

  // one butterfly: `hi` is the element whose index has the stage bit set; (wr,wi) = twiddle (wi negated for the inverse)
  NCB_HD void butterfly( Cplx& lo, Cplx& hi, double wr, double wi )
  {
    const double jr = hi.re*wr - hi.im*wi;
    const double ji = hi.re*wi + hi.im*wr;
    hi.re = lo.re - jr; hi.im = lo.im - ji;
    lo.re += jr; lo.im += ji;
  }
  NCB_HD unsigned bitReverse( unsigned j, int logn )
  {
#ifdef __CUDA_ARCH__
    return __brev( j ) >> ( 32 - logn );
#else
    unsigned r = 0;
    for ( int k = 0; k < logn; ++k ) { r = ( r << 1 ) | ( j & 1u ); j >>= 1; }
    return r;
#endif
  }
  // position of butterfly t of stage i: indices (lo, lo + 2^i) and its offset inside the group (twiddle = off * Wsize/2^(i+1))
  NCB_HD void butterflyIndex( unsigned t, int i, unsigned& lo, unsigned& off )
  {
    const unsigned i1 = 1u << i;
    off = t & ( i1 - 1 );
    lo = ( ( t >> i ) << ( i + 1 ) ) + off;
  }

  // VDOSGnData::interpolateDensity, ref: NCVDOSGn.cc:92-105
  struct GnDev { const double* spec; unsigned long long n; double lower, upper, invbinwidth; };
  NCB_HD double gnInterpolate( const GnDev& g, double energy )
  {
    if ( !inInterval( g.lower, g.upper, energy ) ) return 0.0;
    const double a = ( energy - g.lower )*g.invbinwidth;
    const double floor_a = floor( a );
    unsigned long long index = (unsigned long long)floor_a;
    if ( index > g.n - 2 ) index = g.n - 2;
    const double f = a - floor_a;
    const double* p = g.spec + index;
    return p[0]*( 1.0 - f ) + f*p[1];
  }

  struct JobDev {
    const double* in1; const double* in2;
    unsigned n1, n2, stride1, stride2;
    Cplx* b1; Cplx* b2; Cplx* bo;       // transforms of the inputs, inverse transform of their product (N each)
    int logn; int same;                 // N = 2^logn >= nout; same: both inputs are the same spectrum
    unsigned nout;                      // n1 + n2 - 1
    double k;                           // dt / N
    double dt, trunc_threshold, relthr;
    unsigned thin_nbins; int trunc_thin, gentle;
    double* ytmp;                       // nout scratch
    double* out;                        // result spectrum (capacity nout)
  };
  struct JobStat { unsigned long long ifront, n; unsigned long long extra_thin; double maxval; long long first_above, last_above; };

  // cell sum over one group of orders; returns the contributions to the (beta<=0) cell and, through accp, to its mirror
  struct FillDev {
    const GnDev* gn;              // [order-1]
    const double* scale;          // [order-1]
    const double* afact;          // [(order-order0)*nalpha + ia], orders order0.. of the current window
    const int* a_first; const int* a_end; const unsigned char* skip;
    const double* beta_nonpos; const double* expbeta;
    double* sab;
    unsigned nalpha, idx_zero, idx_firstflip;
    unsigned order0;
    double kT;
  };
  NCB_HD void fillGroup( const FillDev& F, unsigned ia, double energy, double expMbeta, unsigned lo, unsigned hi, double& acc, double& accp )
  {
    acc = 0.0; accp = 0.0;
    for ( unsigned n = lo; n <= hi; ++n ) {
      if ( F.skip[n-1] ) continue;
      const GnDev g = F.gn[n-1];
      if ( !inInterval( g.lower, g.upper, energy ) ) continue;
      const double G = F.scale[n-1]*gnInterpolate( g, energy );
      if ( !( G > 0.0 ) ) continue;
      if ( (int)ia < F.a_first[n-1] || (int)ia >= F.a_end[n-1] ) continue;
      const double c = F.afact[(size_t)( n - F.order0 )*F.nalpha + ia]*G;
      acc += c;
      if ( expMbeta ) accp += c*expMbeta;
    }
  }

#if defined(__CUDACC__) && defined(NCB_VDOS_KERNELS)   // (the kernels belong to ncb_vdos.cu alone)
  constexpr int kFftLocalLog = 12;                 // stages done in shared memory
  constexpr int kFftThreads = 512;
  constexpr unsigned kSumChunk = 4096;             // bins staged per round of the serial area sum

  // blockIdx = (chunk, job, which input).  INV: transform of b1*b2 into bo with conjugated twiddles.
  template <bool INV>
  __global__ void __launch_bounds__( kFftThreads ) k_vdos_fft_local( const JobDev* __restrict__ jobs, const Cplx* __restrict__ W, unsigned wsize )
  {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cplx* s = reinterpret_cast<Cplx*>( smem_raw );
    const JobDev J = jobs[blockIdx.y];
    const int which = INV ? 0 : (int)blockIdx.z;
    if ( !INV && which == 1 && J.same ) return;
    const int llog = J.logn < kFftLocalLog ? J.logn : kFftLocalLog;
    const unsigned chunk = 1u << llog;
    const unsigned base = blockIdx.x*chunk;
    if ( base >= ( 1u << J.logn ) ) return;
    if ( INV ) {
      const Cplx* b2 = J.same ? J.b1 : J.b2;
      for ( unsigned j = threadIdx.x; j < chunk; j += kFftThreads ) {
        const unsigned r = bitReverse( base + j, J.logn );
        s[j] = cmul( J.b1[r], b2[r] );
      }
    } else {
      const double* in = which ? J.in2 : J.in1;
      const unsigned n = which ? J.n2 : J.n1, stride = which ? J.stride2 : J.stride1;
      for ( unsigned j = threadIdx.x; j < chunk; j += kFftThreads ) {
        const unsigned r = bitReverse( base + j, J.logn );
        s[j] = Cplx{ r < n ? in[(size_t)r*stride] : 0.0, 0.0 };
      }
    }
    __syncthreads();
    for ( int i = 0; i < llog; ++i ) {
      const unsigned wstep = wsize >> ( i + 1 );
      for ( unsigned t = threadIdx.x; t < chunk/2; t += kFftThreads ) {
        unsigned lo, off;
        butterflyIndex( t, i, lo, off );
        const Cplx w = W[(size_t)off*wstep];
        butterfly( s[lo], s[lo + ( 1u << i )], w.re, INV ? -w.im : w.im );
      }
      __syncthreads();
    }
    Cplx* dst = INV ? J.bo : ( which ? J.b2 : J.b1 );
    for ( unsigned j = threadIdx.x; j < chunk; j += kFftThreads ) dst[base + j] = s[j];
  }

  // one stage i >= kFftLocalLog over global memory; blockIdx = (butterfly block, job, which input)
  template <bool INV>
  __global__ void k_vdos_fft_stage( const JobDev* __restrict__ jobs, const Cplx* __restrict__ W, unsigned wsize, int i )
  {
    const JobDev J = jobs[blockIdx.y];
    const int which = INV ? 0 : (int)blockIdx.z;
    if ( J.logn <= i || ( !INV && which == 1 && J.same ) ) return;
    const unsigned t = blockIdx.x*blockDim.x + threadIdx.x;
    if ( t >= ( 1u << ( J.logn - 1 ) ) ) return;
    Cplx* d = INV ? J.bo : ( which ? J.b2 : J.b1 );
    unsigned lo, off;
    butterflyIndex( t, i, lo, off );
    const Cplx w = W[(size_t)off*( wsize >> ( i + 1 ) )];
    Cplx a = d[lo], b = d[lo + ( 1u << i )];
    butterfly( a, b, w.re, INV ? -w.im : w.im );
    d[lo] = a; d[lo + ( 1u << i )] = b;
  }

  template <class T, class Op>
  __device__ T blockReduce( T v, T* scratch, Op op )
  {
    for ( int o = 16; o > 0; o >>= 1 ) v = op( v, __shfl_xor_sync( 0xffffffffu, v, o ) );
    __syncthreads();
    if ( ( threadIdx.x & 31 ) == 0 ) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    T r = scratch[0];
    for ( unsigned w = 1; w < ( blockDim.x >> 5 ); ++w ) r = op( r, scratch[w] );
    return r;
  }

  __global__ void __launch_bounds__( 512 ) k_vdos_finish( const JobDev* __restrict__ jobs, JobStat* __restrict__ stats )
  {
    __shared__ double sd[16];
    __shared__ long long sl[16];
    __shared__ double s_inv;
    __shared__ double s_chunk[kSumChunk];
    const JobDev J = jobs[blockIdx.x];
    const unsigned nout = J.nout;
    auto fmaxop = []( double a, double b ) { return a > b ? a : b; };
    auto lminop = []( long long a, long long b ) { return a < b ? a : b; };
    auto lmaxop = []( long long a, long long b ) { return a > b ? a : b; };
    double mx = -1.0;
    for ( unsigned i = threadIdx.x; i < nout; i += blockDim.x ) {
      const Cplx c = J.bo[i];
      double y = c.re*c.re + c.im*c.im;
      y = sqrt( y );
      y *= J.k;
      J.ytmp[i] = y;
      mx = fmaxop( mx, y );
    }
    const double spec_max = blockReduce( mx, sd, fmaxop );
    unsigned long long ifront = 0, lo = 0, len = nout;
    if ( J.trunc_thin && J.trunc_threshold > 0 ) {
      const double cutoff = J.trunc_threshold*spec_max;
      long long f = (long long)nout - 1;
      for ( unsigned i = threadIdx.x; i + 1 < nout; i += blockDim.x )
        if ( J.ytmp[i] > cutoff ) { f = i; break; }
      f = blockReduce( f, sl, lminop );
      long long b = f;
      for ( long long i = (long long)nout - 1 - threadIdx.x; i > f; i -= blockDim.x )
        if ( J.ytmp[i] > cutoff ) { b = i; break; }
      b = blockReduce( b, sl, lmaxop );
      ifront = (unsigned long long)f;
      if ( b > f ) { lo = (unsigned long long)f; len = (unsigned long long)( b - f + 1 ); }
    }
    unsigned long long extra = 1;
    double dt = J.dt;
    if ( J.trunc_thin && J.thin_nbins > 0 && len > J.thin_nbins ) {
      while ( len > (unsigned long long)J.thin_nbins*extra ) extra *= 2;
      if ( extra >= 8 && J.gentle ) extra /= 2;
      len = ( len + extra - 1 )/extra;
      dt *= (double)extra;
    }
    // unit area: the bins are summed in index order by ONE thread, as the reference sums them (a tree reduction would
    // round differently); the CTA stages the bins through shared memory so that the serial chain runs at the latency
    // of a shared-memory load + one DADD per bin instead of an L2 round trip
    {
      double area = 0.;
      for ( unsigned long long m0 = 0; m0 < len; m0 += kSumChunk ) {
        const unsigned cnt = (unsigned)( len - m0 < kSumChunk ? len - m0 : kSumChunk );
        for ( unsigned m = threadIdx.x; m < cnt; m += blockDim.x ) s_chunk[m] = J.ytmp[lo + ( m0 + m )*extra];
        __syncthreads();
        if ( threadIdx.x == 0 ) {
          unsigned m = 0;
          for ( ; m + 8 <= cnt; m += 8 ) {
            const double v0 = s_chunk[m], v1 = s_chunk[m+1], v2 = s_chunk[m+2], v3 = s_chunk[m+3],
                         v4 = s_chunk[m+4], v5 = s_chunk[m+5], v6 = s_chunk[m+6], v7 = s_chunk[m+7];
            area += v0; area += v1; area += v2; area += v3; area += v4; area += v5; area += v6; area += v7;
          }
          for ( ; m < cnt; ++m ) area += s_chunk[m];
        }
        __syncthreads();
      }
      if ( threadIdx.x == 0 ) { area *= dt; s_inv = 1.0/area; }
    }
    __syncthreads();
    const double inv = s_inv;
    mx = -1.0;
    for ( unsigned long long m = threadIdx.x; m < len; m += blockDim.x ) {
      const double v = J.ytmp[lo + m*extra]*inv;
      J.out[m] = v;
      mx = fmaxop( mx, v );
    }
    const double maxval = blockReduce( mx, sd, fmaxop );
    const double thr = J.relthr*maxval;
    long long fa = -1, la = -1;
    {
      long long f = (long long)len;
      for ( unsigned long long m = threadIdx.x; m < len; m += blockDim.x )
        if ( J.ytmp[lo + m*extra]*inv >= thr ) { f = (long long)m; break; }
      f = blockReduce( f, sl, lminop );
      long long l = -1;
      for ( long long m = (long long)len - 1 - threadIdx.x; m >= 0; m -= blockDim.x )
        if ( J.ytmp[lo + (unsigned long long)m*extra]*inv >= thr ) { l = m; break; }
      l = blockReduce( l, sl, lmaxop );
      fa = f < (long long)len ? f : -1; la = l;
    }
    if ( threadIdx.x == 0 ) {
      JobStat S;
      S.ifront = ifront; S.n = len; S.extra_thin = extra; S.maxval = maxval; S.first_above = fa; S.last_above = la;
      stats[blockIdx.x] = S;
    }
  }

  // groups[2*g], groups[2*g+1] = first / last order of summation group g
  __global__ void __launch_bounds__( 128 ) k_vdos_fill( FillDev F, const unsigned* __restrict__ groups, unsigned ngroups )
  {
    const unsigned ia = blockIdx.x*blockDim.x + threadIdx.x, ib = blockIdx.y;
    if ( ia >= F.nalpha ) return;
    const double beta = F.beta_nonpos[ib];
    const double energy = beta*F.kT;
    const bool flip = ib >= F.idx_firstflip && beta != 0.0;
    const double expMbeta = flip ? F.expbeta[ib] : 0.0;
    const size_t cell = (size_t)ib*F.nalpha + ia, mirror = (size_t)( F.idx_zero + ( F.idx_zero - ib ) )*F.nalpha + ia;
    double S = F.sab[cell], Sp = expMbeta ? F.sab[mirror] : 0.0;
    for ( unsigned g = 0; g < ngroups; ++g ) {
      double acc, accp;
      fillGroup( F, ia, energy, expMbeta, groups[2*g], groups[2*g+1], acc, accp );
      S += acc; Sp += accp;
    }
    F.sab[cell] = S;
    if ( expMbeta ) F.sab[mirror] = Sp;
  }
#endif

} }
