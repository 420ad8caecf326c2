// ncb_vdos.cu -- CUDA backend of the VDOS -> S(alpha,beta) expansion (ncb_vdos.h drives it, ncb_vdos_dev.cuh holds
// the kernels).  Own translation unit of libncrystal_b200.so; entry points in ncb_vdos_api.h, C ABI in ncb_lib.cu
// (ncrystal_raw_vdos2kernel / ncrystal_raw_vdos2gn, ref: include/NCrystal/cinterface/ncrystal.h:885-925).
#define NCB_VDOS_KERNELS 1
#define NCB_MATHFN __host__ __device__ __forceinline__   // (ncb_common.cuh: no second out-of-line copy of its libm wrappers)
#include "ncb_vdos_dev.cuh"
#include "ncb_vdos.h"
#include "ncb_vdos_api.h"
#include <cuda_runtime.h>
#include <map>
#include <mutex>

namespace ncb { namespace vdos {

  namespace {

    void cudaOk( cudaError_t e, const char* what )
    {
      if ( e != cudaSuccess ) throw Error( "CalcError", std::string( "CUDA failure in " ) + what + ": " + cudaGetErrorString( e ) );
    }
#define VDOS_CUDA_OK(x) cudaOk( (x), #x )

    // device scratch comes from the stream-ordered pool of the device (cudaMallocAsync): after the first expansion
    // the blocks are recycled, an expansion then costs no cudaMalloc / cudaFree round trips
    template <class T> struct DevBuf {
      T* p = nullptr; size_t cap = 0; cudaStream_t st = 0;
      ~DevBuf() { if ( p ) cudaFreeAsync( p, st ); }
      void need( size_t n )
      {
        if ( n <= cap ) return;
        if ( p ) { cudaFreeAsync( p, st ); p = nullptr; cap = 0; }
        size_t c = std::max<size_t>( n, 1024 );
        VDOS_CUDA_OK( cudaMallocAsync( &p, c*sizeof(T), st ) );
        cap = c;
      }
      explicit DevBuf( cudaStream_t s ) : st( s ) {}
      DevBuf( const DevBuf& ) = delete; DevBuf& operator=( const DevBuf& ) = delete;
    };

    class CudaBackend {
    public:
      explicit CudaBackend( cudaStream_t st ) : m_st( st ), m_work( st ), m_ytmp( st ), m_jobs( st ), m_stats( st )
      {
        static std::mutex mtx; static std::map<int,bool> done;
        int dev = 0; VDOS_CUDA_OK( cudaGetDevice( &dev ) );
        std::lock_guard<std::mutex> g( mtx );
        if ( !done[dev] ) {
          const int smem = (int)( sizeof(Cplx) << kFftLocalLog );
          VDOS_CUDA_OK( cudaFuncSetAttribute( k_vdos_fft_local<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem ) );
          VDOS_CUDA_OK( cudaFuncSetAttribute( k_vdos_fft_local<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem ) );
          cudaMemPool_t pool;
          VDOS_CUDA_OK( cudaDeviceGetDefaultMemPool( &pool, dev ) );
          unsigned long long keep = (unsigned long long)256 << 20;    // freed scratch stays with the pool up to 256 MB
          VDOS_CUDA_OK( cudaMemPoolSetAttribute( pool, cudaMemPoolAttrReleaseThreshold, &keep ) );
          done[dev] = true;
        }
      }
      ~CudaBackend() { for ( double* p : m_pools ) cudaFreeAsync( p, m_st ); }
      unsigned launches = 0;

      void setSpectrum( unsigned order, const VectD& spec )
      {
        slot( order, spec.size() );
        VDOS_CUDA_OK( cudaMemcpyAsync( m_spec[order-1], spec.data(), spec.size()*8, cudaMemcpyHostToDevice, m_st ) );
        VDOS_CUDA_OK( cudaStreamSynchronize( m_st ) );
        m_len[order-1] = spec.size();
      }
      VectD spectrum( unsigned order )
      {
        VectD v( m_len.at( order-1 ) );
        VDOS_CUDA_OK( cudaMemcpyAsync( v.data(), m_spec[order-1], v.size()*8, cudaMemcpyDeviceToHost, m_st ) );
        VDOS_CUDA_OK( cudaStreamSynchronize( m_st ) );
        return v;
      }
      void convolve( const std::vector<ConvJob>& jobs, std::vector<ConvResult>& res )
      {
        const size_t nj = jobs.size();
        res.assign( nj, ConvResult() );
        if ( !nj ) return;
        std::vector<JobDev> hj( nj );
        size_t work = 0, ytot = 0;
        int maxlog = 0;
        for ( size_t j = 0; j < nj; ++j ) {
          const ConvJob& J = jobs[j];
          const size_t nout = J.n1 + J.n2 - 1;
          int logn = 0;
          while ( ( (size_t)1 << logn ) < nout ) ++logn;
          if ( logn > 20 ) throw Error( "CalcError", "VDOS spectrum too long for the FFT twiddle table (more than 2^20 bins)" );
          hj[j].logn = logn; hj[j].nout = (unsigned)nout;
          work += (size_t)3 << logn; ytot += nout;
          maxlog = std::max( maxlog, logn );
        }
        needTwiddles( maxlog );
        reservePool( ytot + 32*nj );
        m_work.need( work ); m_ytmp.need( ytot ); m_jobs.need( nj ); m_stats.need( nj );
        size_t woff = 0, yoff = 0;
        for ( size_t j = 0; j < nj; ++j ) {
          const ConvJob& J = jobs[j];
          JobDev& D = hj[j];
          const size_t N = (size_t)1 << D.logn;
          D.in1 = m_spec.at( J.o1-1 ); D.in2 = m_spec.at( J.o2-1 );
          D.n1 = (unsigned)J.n1; D.n2 = (unsigned)J.n2; D.stride1 = J.stride1; D.stride2 = J.stride2;
          D.same = ( J.o1 == J.o2 && J.stride1 == J.stride2 ) ? 1 : 0;
          D.b1 = m_work.p + woff; D.b2 = D.b1 + N; D.bo = D.b2 + N; woff += 3*N;
          D.k = J.dt/(double)N;
          D.dt = J.dt; D.trunc_threshold = J.trunc_threshold; D.relthr = J.relthr;
          D.thin_nbins = J.thin_nbins; D.trunc_thin = J.trunc_thin ? 1 : 0; D.gentle = J.gentle_thinning ? 1 : 0;
          D.ytmp = m_ytmp.p + yoff; yoff += D.nout;
          D.out = slot( J.order, D.nout );
        }
        VDOS_CUDA_OK( cudaMemcpyAsync( m_jobs.p, hj.data(), nj*sizeof(JobDev), cudaMemcpyHostToDevice, m_st ) );
        const int llog = std::min( maxlog, kFftLocalLog );
        const unsigned nchunks = 1u << ( maxlog - llog );
        const size_t smem = sizeof(Cplx) << llog;
        const unsigned wsize = 1u << m_wlog;
        k_vdos_fft_local<false><<< dim3( nchunks, (unsigned)nj, 2 ), kFftThreads, smem, m_st >>>( m_jobs.p, m_wdev, wsize );
        ++launches;
        const unsigned bfblocks = maxlog > 0 ? ( ( 1u << ( maxlog - 1 ) ) + 255 )/256 : 1;
        for ( int i = kFftLocalLog; i < maxlog; ++i ) {
          k_vdos_fft_stage<false><<< dim3( bfblocks, (unsigned)nj, 2 ), 256, 0, m_st >>>( m_jobs.p, m_wdev, wsize, i );
          ++launches;
        }
        k_vdos_fft_local<true><<< dim3( nchunks, (unsigned)nj, 1 ), kFftThreads, smem, m_st >>>( m_jobs.p, m_wdev, wsize );
        ++launches;
        for ( int i = kFftLocalLog; i < maxlog; ++i ) {
          k_vdos_fft_stage<true><<< dim3( bfblocks, (unsigned)nj, 1 ), 256, 0, m_st >>>( m_jobs.p, m_wdev, wsize, i );
          ++launches;
        }
        k_vdos_finish<<< (unsigned)nj, 512, 0, m_st >>>( m_jobs.p, m_stats.p );
        ++launches;
        VDOS_CUDA_OK( cudaGetLastError() );
        std::vector<JobStat> hs( nj );
        VDOS_CUDA_OK( cudaMemcpyAsync( hs.data(), m_stats.p, nj*sizeof(JobStat), cudaMemcpyDeviceToHost, m_st ) );
        VDOS_CUDA_OK( cudaStreamSynchronize( m_st ) );
        for ( size_t j = 0; j < nj; ++j ) {
          ConvResult& R = res[j];
          R.ifront = (size_t)hs[j].ifront; R.n = (size_t)hs[j].n; R.extra_thin = (unsigned long)hs[j].extra_thin;
          R.maxval = hs[j].maxval; R.first_above = (long)hs[j].first_above; R.last_above = (long)hs[j].last_above;
          m_len[jobs[j].order-1] = R.n;
        }
      }
      void fill( const FillPlan& P, VectD& sab )
      {
        const size_t na = P.nalpha, nb = P.nbeta, nrows = P.beta_nonpos.size();
        const unsigned norders = P.norders;
        std::vector<GnDev> gn( norders );
        for ( unsigned n = 1; n <= norders; ++n ) {
          const GnMeta& m = P.meta[n-1];
          gn[n-1] = GnDev{ m_spec.at( n-1 ), (unsigned long long)m.n, m.lower, m.upper, 1.0/m.binwidth };
        }
        DevBuf<GnDev> d_gn( m_st ); DevBuf<double> d_scale( m_st ), d_af( m_st ), d_beta( m_st ), d_expb( m_st ), d_sab( m_st );
        DevBuf<int> d_first( m_st ), d_end( m_st ); DevBuf<unsigned char> d_skip( m_st ); DevBuf<unsigned> d_groups( m_st );
        d_gn.need( norders ); d_scale.need( norders ); d_first.need( norders ); d_end.need( norders ); d_skip.need( norders );
        d_beta.need( nrows ); d_expb.need( nrows ); d_sab.need( na*nb );
        auto up = [&]( void* d, const void* h, size_t n ) { VDOS_CUDA_OK( cudaMemcpyAsync( d, h, n, cudaMemcpyHostToDevice, m_st ) ); };
        up( d_gn.p, gn.data(), norders*sizeof(GnDev) ); up( d_scale.p, P.scale.data(), norders*8 );
        up( d_first.p, P.a_first.data(), norders*4 ); up( d_end.p, P.a_end.data(), norders*4 ); up( d_skip.p, P.skip.data(), norders );
        up( d_beta.p, P.beta_nonpos.data(), nrows*8 ); up( d_expb.p, P.expbeta.data(), nrows*8 );
        VDOS_CUDA_OK( cudaMemsetAsync( d_sab.p, 0, na*nb*8, m_st ) );
        // windows of summation groups whose alpha-factor rows fit the staging buffer
        const size_t max_rows = std::max<size_t>( 64, ( (size_t)256 << 20 )/( na*8 ) );
        size_t g0 = 0;
        while ( g0 < P.job_orders.size() ) {
          size_t g1 = g0;
          const unsigned first_order = (unsigned)P.job_orders[g0].first;
          unsigned last_order = first_order;
          while ( g1 < P.job_orders.size() && ( g1 == g0 || (unsigned)P.job_orders[g1].second - first_order + 1 <= max_rows ) ) {
            last_order = (unsigned)P.job_orders[g1].second; ++g1;
          }
          const size_t rows = last_order - first_order + 1;
          d_af.need( rows*na );
          up( d_af.p, &P.alpha_factor[(size_t)( first_order-1 )*na], rows*na*8 );
          std::vector<unsigned> groups;
          for ( size_t g = g0; g < g1; ++g ) { groups.push_back( (unsigned)P.job_orders[g].first ); groups.push_back( (unsigned)P.job_orders[g].second ); }
          d_groups.need( groups.size() );
          up( d_groups.p, groups.data(), groups.size()*4 );
          FillDev F;
          F.gn = d_gn.p; F.scale = d_scale.p; F.afact = d_af.p; F.a_first = d_first.p; F.a_end = d_end.p; F.skip = d_skip.p;
          F.beta_nonpos = d_beta.p; F.expbeta = d_expb.p; F.sab = d_sab.p;
          F.nalpha = (unsigned)na; F.idx_zero = (unsigned)P.idx_zero; F.idx_firstflip = (unsigned)P.idx_firstflip;
          F.order0 = first_order; F.kT = P.kT;
          k_vdos_fill<<< dim3( (unsigned)( ( na + 127 )/128 ), (unsigned)nrows ), 128, 0, m_st >>>( F, d_groups.p, (unsigned)( g1 - g0 ) );
          ++launches;
          VDOS_CUDA_OK( cudaGetLastError() );
          VDOS_CUDA_OK( cudaStreamSynchronize( m_st ) );   // (the host vectors of this window are reused)
          g0 = g1;
        }
        sab.resize( na*nb );
        VDOS_CUDA_OK( cudaMemcpyAsync( sab.data(), d_sab.p, na*nb*8, cudaMemcpyDeviceToHost, m_st ) );
        VDOS_CUDA_OK( cudaStreamSynchronize( m_st ) );
      }
    private:
      cudaStream_t m_st;
      std::vector<double*> m_spec;      // [order-1]
      std::vector<double*> m_pools;
      double* m_pool_cur = nullptr; size_t m_pool_left = 0, m_next_pool = (size_t)1 << 16;
      std::vector<size_t> m_len;
      DevBuf<Cplx> m_work;
      const Cplx* m_wdev = nullptr;
      DevBuf<double> m_ytmp;
      DevBuf<JobDev> m_jobs;
      DevBuf<JobStat> m_stats;
      int m_wlog = -1;

      // spectra are carved out of pools (one per batch of orders), never moved afterwards
      double* slot( unsigned order, size_t capacity )
      {
        if ( m_spec.size() < order ) { m_spec.resize( order, nullptr ); m_len.resize( order, 0 ); }
        capacity = ( std::max<size_t>( capacity, 4 ) + 31 ) & ~(size_t)31;
        if ( m_pool_left < capacity ) {
          const size_t sz = std::max<size_t>( capacity, m_next_pool );
          double* p = nullptr;
          VDOS_CUDA_OK( cudaMallocAsync( &p, sz*8, m_st ) );
          m_pools.push_back( p ); m_pool_cur = p; m_pool_left = sz;
        }
        m_spec[order-1] = m_pool_cur; m_pool_cur += capacity; m_pool_left -= capacity;
        return m_spec[order-1];
      }
      void reservePool( size_t ndoubles ) { m_next_pool = std::max<size_t>( ndoubles, (size_t)1 << 16 ); m_pool_left = 0; }
      // device copy of the twiddle table: one per device for the life of the process, grown on demand
      void needTwiddles( int logn )
      {
        if ( logn <= m_wlog ) return;
        static std::mutex mtx;
        struct DevTable { Cplx* p = nullptr; int log = -1; };
        static std::map<int,DevTable> cache;
        int dev = 0; VDOS_CUDA_OK( cudaGetDevice( &dev ) );
        std::lock_guard<std::mutex> g( mtx );
        DevTable& t = cache[dev];
        if ( t.log < logn ) {
          const int l = std::max( logn, 13 );
          const std::vector<Cplx>& w = twiddles( (unsigned)l );
          Cplx* p = nullptr;
          VDOS_CUDA_OK( cudaMalloc( &p, w.size()*sizeof(Cplx) ) );
          VDOS_CUDA_OK( cudaMemcpy( p, w.data(), w.size()*sizeof(Cplx), cudaMemcpyHostToDevice ) );
          // (a smaller table stays allocated: an expansion on another thread may still be reading it)
          t.p = p; t.log = l;
        }
        m_wdev = t.p; m_wlog = t.log;
      }
    };
  }

  namespace {
    // the expansion runs on a stream of its own; the guard waits for the stream-ordered frees before destroying it
    struct StreamGuard {
      cudaStream_t st = nullptr;
      StreamGuard()
      {
        int ndev = 0;
        cudaError_t ce = cudaGetDeviceCount( &ndev );
        if ( ce != cudaSuccess || ndev <= 0 )
          throw Error( "CalcError", std::string( "ncrystal_b200 requires a CUDA device (no CPU fallback): " )
                       + ( ce != cudaSuccess ? cudaGetErrorString( ce ) : "no devices found" ) );
        VDOS_CUDA_OK( cudaStreamCreateWithFlags( &st, cudaStreamNonBlocking ) );
      }
      ~StreamGuard() { if ( st ) { cudaStreamSynchronize( st ); cudaStreamDestroy( st ); } }
    };
  }

  Kernel expandOnDevice( const Input& in, unsigned vdoslux, double target_emax, const std::function<double(unsigned)>& scaleFct, unsigned* launches )
  {
    StreamGuard sg;
    Kernel K;
    {
      CudaBackend be( sg.st );
      K = expand( in, vdoslux, target_emax, be, scaleFct );
      if ( launches ) *launches = be.launches;
    }
    return K;
  }

  VectD gnOnDevice( const Input& in, unsigned order, double& xmin, double& xmax )
  {
    if ( order < 1 || order >= 100000 ) throw Error( "BadInput", "invalid phonon order" );
    StreamGuard sg;
    CudaBackend be( sg.st );
    Eval ev( in );
    Ladder<CudaBackend> Gn( ev, ev.calcGamma0(), be, TruncThin(), 1e-9 );
    Gn.grow( order, 0 );
    const PairDD r = Gn.eRange( order );
    xmin = r.first; xmax = r.second;
    return be.spectrum( order );
  }

} }
