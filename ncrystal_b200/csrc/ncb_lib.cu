// ncb_lib.cu -- libncrystal_b200.so: host runtime + C ABI (include/ncrystal_b200.h)
// over the sm_100a kernels of ncb_kernels.cuh.
//
// Host side mirrors the reference's C interface layer (ref: ncrystal_core/src/
// cinterface/ncrystal.cc): ref-counted opaque handles with a 32-bit type tag as
// first member (:138-191), global error state with halt/quiet switches (:280-306),
// sentinel outputs on error (:1095,:1136-1140,:1239-1245).
//
// There is NO CPU fallback: every compute entry point launches CUDA kernels and
// raises an error when no usable device is present.
#include "ncb_kernels.cuh"
#include "ncb_kernels_sc.cuh"
#include "ncb_kernels_lc.cuh"
#include "ncb_kernels_mmc.cuh"
#include "ncb_loader.h"
#include "ncb_loader_sc.h"
#include "ncb_sabgrid.h"
#include "ncb_vdos_api.h"
#include "../../include/ncrystal_b200.h"

#include <atomic>
#include <condition_variable>
#include <deque>
#include <cstdio>
#include <dirent.h>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

namespace ncb {
  double g_erfc_lut_host[kErfcLutLen];
}

namespace {

  using namespace ncb;

  // ------------------------------------------------------------------ error state
  // ref: ncrystal.cc:280-306
  int g_quietonerror = 0;
  int g_haltonerror = 1;
  int g_waserror = 0;
  char g_errmsg[512];
  char g_errtype[64];
  // Kernel-tuning switches (chunk sizes, CTAs per SM, alternative kernel orderings: the A/B experiments recorded in
  // DESIGN.md) are compiled out of the product: they are read from the environment only in a -DNCB200_TUNING=1 build.
  // Run-time configuration that stays: NCB200_DATA_PATH (compiled materials), NCB200_COPY_THREADS (host copy pool).
  inline const char* tuneEnv( const char* name )
  {
#if defined(NCB200_TUNING) && NCB200_TUNING
    return std::getenv( name );
#else
    (void)name;
    return nullptr;
#endif
  }
  void (*g_custom_error_handler)(char*,char*) = nullptr;
  // message handler, ref: ncrystal.h:1051, ncrystal.cc:2428-2446, src/utils/NCMsg.cc:44-82 (0: info, 1: warning, 2: raw)
  std::mutex g_msg_mtx;
  void (*g_msg_handler)(const char*,unsigned) = nullptr;
  void emitMsg( const char* msg, unsigned type )
  {
    std::lock_guard<std::mutex> g( g_msg_mtx );
    if ( g_msg_handler ) { g_msg_handler( msg, type ); return; }
    if ( type == 2 ) std::fputs( msg, stdout );
    else std::printf( "%s%s\n", type == 1 ? "NCrystal WARNING: " : "NCrystal: ", msg );
    std::fflush( stdout );
  }

  void setError( const char* msg, const char* etype = nullptr ) noexcept
  {
    if ( !etype ) etype = "ncrystal_c-interface";
    std::strncpy( g_errmsg, msg, sizeof(g_errmsg)-1 );
    std::strncpy( g_errtype, etype, sizeof(g_errtype)-1 );
    g_errmsg[sizeof(g_errmsg)-1] = '\0';
    g_errtype[sizeof(g_errtype)-1] = '\0';
    if ( g_custom_error_handler )
      (*g_custom_error_handler)( g_errtype, g_errmsg );
    g_waserror = 1;
    if ( !g_quietonerror )
      std::fprintf( stdout, "NCrystal ERROR [%s]: %s\n", g_errtype, g_errmsg );
    if ( g_haltonerror ) {
      std::fprintf( stdout, "NCrystal terminating due to ERROR\n" );
      std::fflush( stdout );
      std::exit(1);
    }
  }

  struct Err : public std::runtime_error {
    std::string type;
    Err( const std::string& t, const std::string& m ) : std::runtime_error(m), type(t) {}
  };
  void handleError( const std::exception& e ) noexcept
  {
    if ( auto x = dynamic_cast<const Err*>( &e ) ) setError( x->what(), x->type.c_str() );
    else if ( dynamic_cast<const std::runtime_error*>( &e ) ) setError( e.what(), "std::runtime_error" );
    else setError( "<unknown>", "std::exception" );
  }
#define NCBCATCH catch ( std::exception& e ) { handleError(e); }

  void cudaCheck( cudaError_t e, const char* what )
  {
    if ( e != cudaSuccess )
      throw Err( "CalcError", std::string("CUDA failure in ")+what+": "+cudaGetErrorString(e) );
  }
#define CUDA_OK(x) cudaCheck( (x), #x )

  std::atomic<uint64_t> g_launches{0};

  // ---- optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg).
  // Off by default; when on, every timed launch is bracketed by two events.
  struct KernelTimer {
    struct Rec { const char* name; cudaEvent_t a, b; };
    bool enabled = false;
    std::vector<Rec> recs;
    std::mutex mtx;
    void begin( const char* name, cudaStream_t st )
    {
      if ( !enabled ) return;
      Rec r; r.name = name;
      cudaEventCreate( &r.a ); cudaEventCreate( &r.b );
      cudaEventRecord( r.a, st );
      std::lock_guard<std::mutex> g( mtx );
      recs.push_back( r );
    }
    void end( cudaStream_t st )
    {
      if ( !enabled ) return;
      std::lock_guard<std::mutex> g( mtx );
      cudaEventRecord( recs.back().b, st );
    }
  } g_ktimer;
  struct TimedLaunch {
    cudaStream_t st;
    TimedLaunch( const char* name, cudaStream_t s ) : st(s) { g_ktimer.begin( name, s ); }
    ~TimedLaunch() { g_ktimer.end( st ); }
  };

  // ------------------------------------------------------------------ material on device
  std::atomic<uint64_t> g_material_uid_counter{0};
  struct DeviceMaterial {
    int device = 0;
    void* d_arena = nullptr;
    size_t arena_bytes = 0;
    Material mat;            // device pointers
    StagePlan sp;            // smem staging plan for the hot tables
    StagePlan sp_sc;         // staging plan of the warp-cooperative SCBragg kernels (SCBragg tables only)
    StagePlan sp_find;       // k_sc_find: single-precision normals only
    StagePlan sp_iso;        // hot tables of the isotropic leaves only
    bool sc_warp_ok = false; // SCBragg tables fit the warp-cooperative kernels
    bool has_fg_leaf = false; // a FreeGas leaf: its queue spans all energies (-> k_fg_group)
    uint32_t sc_famof_off = 0, sc_scratch_off = 0, sc_smem = 0, sc_find_smem = 0, sc_find_famof_off = 0, sc_find_scratch_off = 0;
    std::string cfg;
    double numdens = 0.0, abs_c = 0.0, temperature = -1.0;
    std::vector<SabBuildPlan> sabplans;
    std::atomic<uint32_t> clone_counter{0};
    uint64_t uid = 0;        // ncrystal_process_uid
    // multi-device fan-out (ncb_lib_multigpu.inc): the compiled material it was made from, its copies on other devices
    std::shared_ptr<const std::vector<unsigned char>> blob;
    std::map<int, std::shared_ptr<DeviceMaterial>> peers;
    std::mutex peers_mtx;
    ~DeviceMaterial() { if ( d_arena ) cudaFree( d_arena ); }
  };

  std::once_flag g_lut_once;
  std::mutex g_lutdev_mtx;
  std::vector<int> g_lut_devices;

  void ensureErfcLut( int device )
  {
    std::call_once( g_lut_once, [](){ fillErfcLutHost( g_erfc_lut_host ); } );
    std::lock_guard<std::mutex> g( g_lutdev_mtx );
    for ( int d : g_lut_devices ) if ( d == device ) return;
    CUDA_OK( cudaMemcpyToSymbol( g_erfc_lut_dev, g_erfc_lut_host, sizeof(g_erfc_lut_host) ) );
    g_lut_devices.push_back( device );
  }

  void buildStagePlan( DeviceMaterial& dm )
  {
    StagePlan& sp = dm.sp;
    std::memset( &sp, 0, sizeof(sp) );
    const Material& M = dm.mat;
    // Budget: what still lets 8 CTAs of 256 threads share an SM (227 KB usable, the classify kernel has 8 KB of
    // its own).  With the energy-key luts a search touches 0-2 table entries, so large Bragg tables (YAG: 2 x 48 KB)
    // are better left to L1/L2 than staged at the price of occupancy (measured, YAG cross sections: 1.6e10/s with
    // the tables in global memory and 8 CTAs/SM, 1.25e10/s with 2dE staged and 2 CTAs/SM).
    uint32_t budget = 20u*1024u;
    uint32_t off = 0;
    auto add = [&]( int slot, const void* p, int n, int elem = 8 ) {
      if ( !p || n <= 0 ) return;
      const uint32_t nb = (uint32_t)( ( (size_t)n*elem + 15 ) & ~(size_t)15 ); // arena sections are 256B padded
      if ( off + nb > budget ) return;
      sp.src[slot] = p; sp.nbytes[slot] = nb; sp.off[slot] = off;
      off += ( nb + 127u ) & ~127u;
      sp.copy_bytes += nb;
    };
    int npb = 0, nsab = 0;
    for ( int i = 0; i < M.ncomp; ++i ) {
      if ( M.comp[i].kind == KIND_POWDERBRAGG ) ++npb;
      if ( M.comp[i].kind == KIND_SAB ) ++nsab;
    }
    // small tables first (SAB grids, the energy-key luts), then the Bragg tables: 2dE (searched) before the
    // cumulative table (one or two reads per neutron); whatever does not fit is read through L1
    for ( int i = 0; i < nsab; ++i ) {
      add( 2*kMaxPB + i, M.sab[i].egrid, M.sab[i].negrid );
      add( 2*kMaxPB + kMaxSab + i, M.sab[i].xs, M.sab[i].negrid );
      if ( M.sab[i].elut ) add( 3*kMaxPB + 2*kMaxSab + i, M.sab[i].elut, M.sab[i].elut_nk + 1, 2 );
    }
    for ( int i = 0; i < npb; ++i )
      if ( M.pb[i].lut ) add( 2*kMaxPB + 2*kMaxSab + i, M.pb[i].lut, M.pb[i].lut_nk + 1, 2 );
    for ( int i = 0; i < npb; ++i ) add( i, M.pb[i].e2d, M.pb[i].n );
    for ( int i = 0; i < npb; ++i ) add( kMaxPB + i, M.pb[i].fdm, M.pb[i].n );
    if ( M.sc.nfam ) {
      // single-crystal tables: staged whole by the warp-per-neutron kernels (their own plans below, 1-3 CTAs/SM)
      budget = off + 100u*1024u;
      const ScBraggT& S = M.sc;
      auto al = []( size_t b ) { return (uint32_t)( ( b + 127 ) & ~(size_t)127 ); };
      const uint32_t need = al( (size_t)S.nnormals*24 ) + 2*al( (size_t)S.nfam*8 ) + al( (size_t)( S.nfam+1 )*4 )
                          + al( (size_t)( S.sofcosd.nm2+2 )*16 ) + al( (size_t)( S.evalcosx.nm2+2 )*16 );
      if ( off + need <= budget ) {
        add( kHotSlotsIso+0, S.normals, 3*S.nnormals );
        add( kHotSlotsIso+1, S.fam_xsfact, S.nfam );
        add( kHotSlotsIso+2, S.fam_inv2d, S.nfam );
        add( kHotSlotsIso+3, S.fam_first, S.nfam+1, 4 );
        add( kHotSlotsIso+4, S.sofcosd.data, 2*( S.sofcosd.nm2+2 ) );
        add( kHotSlotsIso+5, S.evalcosx.data, 2*( S.evalcosx.nm2+2 ) );
      }
    }
    sp.total = off;
    // derived plans
    dm.sp_iso = sp; dm.sp_sc = sp;
    std::memset( &dm.sp_sc, 0, sizeof(StagePlan) );
    if ( sp.nbytes[kHotSlotsIso] ) {
      uint32_t o2 = 0;
      for ( int sl = kHotSlotsIso; sl < kHotSlots; ++sl ) {
        dm.sp_sc.src[sl] = sp.src[sl]; dm.sp_sc.nbytes[sl] = sp.nbytes[sl]; dm.sp_sc.off[sl] = o2;
        o2 += ( sp.nbytes[sl] + 127u ) & ~127u;
        dm.sp_sc.copy_bytes += sp.nbytes[sl];
        // the isotropic-only plan drops the SCBragg slots (they were appended last)
        dm.sp_iso.copy_bytes -= sp.nbytes[sl];
        dm.sp_iso.nbytes[sl] = 0;
      }
      dm.sp_iso.total = sp.off[kHotSlotsIso];
      dm.sp_sc.total = o2;
      dm.sc_warp_ok = ( M.sc.nfam <= kScMaxFam && M.sc.nfam <= 255 && M.sc.nnormals <= 65535 );
      dm.sc_famof_off = o2;
      dm.sc_scratch_off = ( o2 + (uint32_t)M.sc.nnormals + 127u ) & ~127u;
      dm.sc_smem = dm.sc_scratch_off + (uint32_t)( kScWarps*sizeof(ScWarpScratch) );
      // k_sc_find stages the float normals (slot of the double normals) and nothing else
      std::memset( &dm.sp_find, 0, sizeof(StagePlan) );
      {
        const uint32_t nnp = ( (uint32_t)M.sc.nnormals + 127u ) & ~127u;
        const uint32_t nb = 4u*nnp*(uint32_t)sizeof(float);
        dm.sp_find.src[kHotSlotsIso] = M.sc.normals_f; dm.sp_find.nbytes[kHotSlotsIso] = nb; dm.sp_find.off[kHotSlotsIso] = 0;
        dm.sp_find.copy_bytes = nb;
        dm.sp_find.total = ( nb + 127u ) & ~127u;
        dm.sc_find_famof_off = dm.sp_find.total;     // (no family map: the records carry the family index)
        dm.sc_find_scratch_off = dm.sp_find.total;
        dm.sc_find_smem = dm.sc_find_scratch_off + (uint32_t)( kScFindWarps*sizeof(ScFindScratch) );
      }
    }
  }

  // The dynamic shared-memory limit is an attribute of the KERNEL (per device), not of a material: it is raised once
  // per device to the opt-in maximum for every kernel that stages tables, so that handles of several materials can
  // be alive together whatever order they were created in (r1 set it to the size of the material loaded last).
  constexpr int kSmemOptIn = 227*1024;
  template <class K>
  void setSmemAttr( K kernel )
  {
    cudaFuncAttributes fa;
    CUDA_OK( cudaFuncGetAttributes( &fa, kernel ) );   // static shared memory counts against the same per-CTA limit
    const int dyn = kSmemOptIn - (int)( ( fa.sharedSizeBytes + 1023 ) & ~(size_t)1023 );
    CUDA_OK( cudaFuncSetAttribute( kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn ) );
  }
  void ensureKernelAttrs( int device );   // after the kernels' declarations, below

  int numSMs( int device )
  {
    static std::mutex m; static std::vector<std::pair<int,int>> cache;
    std::lock_guard<std::mutex> g(m);
    for ( auto& e : cache ) if ( e.first == device ) return e.second;
    int n = 0;
    CUDA_OK( cudaDeviceGetAttribute( &n, cudaDevAttrMultiProcessorCount, device ) );
    cache.emplace_back( device, n );
    return n;
  }

  void buildSabTablesOnDevice( DeviceMaterial& dm, cudaStream_t st )
  {
    unsigned char* base = static_cast<unsigned char*>( dm.d_arena );
    for ( auto& pl : dm.sabplans ) {
      const SabT& T = dm.mat.sab[pl.sab_index];   // (reference into dm.mat: sees the updates of the auto-grid branch)
      const int na = T.nalpha, nb = T.nbeta, ne = T.negrid;
      double* logsab = reinterpret_cast<double*>( base + pl.off_logsab );
      double* cumul = reinterpret_cast<double*>( base + pl.off_cumul );
      SabRow* rows = reinterpret_cast<SabRow*>( base + pl.off_rows );
      SabAlphaInfo* ainfo = reinterpret_cast<SabAlphaInfo*>( base + pl.off_ainfo );
      SabEPoint* ep = reinterpret_cast<SabEPoint*>( base + pl.off_ep );
      double* bx = reinterpret_cast<double*>( base + pl.off_bx );
      double* bpdf = reinterpret_cast<double*>( base + pl.off_bpdf );
      double* bcdf = reinterpret_cast<double*>( base + pl.off_bcdf );
      double* xscheck = reinterpret_cast<double*>( base + pl.off_xscheck );
      int* errs = reinterpret_cast<int*>( base + pl.off_xscheck + (size_t)ne*8 );
      const size_t ntot = (size_t)na*nb;
      k_sab_logs<<< (unsigned)( ( ntot + 255 )/256 ), 256, 0, st >>>( T.sab, logsab, ntot );
      k_sab_cumul<<< ( nb + 63 )/64, 64, 0, st >>>( T.alpha, T.sab, logsab, na, nb, cumul );
      if ( pl.auto_egrid ) {
        // ---- energy grid from the scattering kernel alone (ncb_sabgrid.h): every point of a probe sequence is
        // integrated in ONE pass of the table-build kernels (they run per energy point anyway)
        SabT& Tm = dm.mat.sab[pl.sab_index];
        double* d_egrid = reinterpret_cast<double*>( base + pl.off_egrid );
        auto sigmaAt = [&]( const std::vector<double>& e ) {
          if ( e.empty() || (int)e.size() > ne )
            throw Err( "CalcError", "SAB energy grid: probe sequence does not fit the table scratch" );
          CUDA_OK( cudaMemcpyAsync( d_egrid, e.data(), e.size()*8, cudaMemcpyHostToDevice, st ) );
          SabT Tp = Tm; Tp.negrid = (int)e.size();
          k_sab_rows<<< dim3( ( nb + 127 )/128, (unsigned)e.size() ), 128, 0, st >>>( Tp, rows, ainfo );
          k_sab_epoints<<< ( (int)e.size() + 31 )/32, 32, 0, st >>>( Tp, rows, ep, bx, bpdf, bcdf, xscheck, errs );
          g_launches += 2;
          std::vector<double> xs( e.size() );
          std::vector<int> er( e.size() );
          CUDA_OK( cudaMemcpyAsync( xs.data(), xscheck, e.size()*8, cudaMemcpyDeviceToHost, st ) );
          CUDA_OK( cudaMemcpyAsync( er.data(), errs, e.size()*4, cudaMemcpyDeviceToHost, st ) );
          CUDA_OK( cudaStreamSynchronize( st ) );
          for ( int c : er ) if ( c ) throw Err( "CalcError", "S(alpha,beta) energy-point analysis failed on device (code "+std::to_string(c)+")" );
          if ( tuneEnv( "NCB200_DEBUG_EGRID" ) )
            for ( size_t k = 0; k < e.size(); ++k ) std::fprintf( stderr, "egrid-probe %zu %.17g %.17g\n", k, e[k], xs[k] );
          return xs;
        };
        // (alpha, beta grid ends for the kinematic limit of the table)
        double bmin = 0.0, amax = 0.0;
        CUDA_OK( cudaMemcpy( &bmin, T.beta, 8, cudaMemcpyDeviceToHost ) );
        CUDA_OK( cudaMemcpy( &amax, T.alpha + ( na-1 ), 8, cudaMemcpyDeviceToHost ) );
        std::vector<double> egrid;
        try {
          egrid = sabDetermineEnergyGrid( ne, T.kT, bmin, amax, pl.suggested_emax, pl.req_emin, pl.req_emax, T.ext, sigmaAt, []( const char* m ) { emitMsg( m, 1 ); } );
        } catch ( std::runtime_error& e ) {
          throw Err( "BadInput", e.what() );
        }
        CUDA_OK( cudaMemcpyAsync( d_egrid, egrid.data(), (size_t)ne*8, cudaMemcpyHostToDevice, st ) );
        CUDA_OK( cudaStreamSynchronize( st ) );
        // derived search aids of the grid (what the loader makes for a grid that comes with the blob)
        {
          const double l0 = std::log( egrid[0] ), l1 = std::log( egrid[ne-1] );
          Tm.egrid_log0 = l0;
          Tm.egrid_invdlog = l1 > l0 ? ( (double)ne - 1.0 )/( l1 - l0 ) : 0.0;
          int key0 = 0, shift = 0, nk = 0;
          const std::vector<uint16_t> lut = makeKeyLut( egrid.data(), (size_t)ne, key0, shift, nk );
          if ( !lut.empty() && lut.size() <= kKeyLutMaxEntries ) {
            CUDA_OK( cudaMemcpy( base + pl.off_elut, lut.data(), lut.size()*sizeof(uint16_t), cudaMemcpyHostToDevice ) );
            Tm.elut = reinterpret_cast<const uint16_t*>( base + pl.off_elut );
            Tm.elut_key0 = key0; Tm.elut_shift = shift; Tm.elut_nk = nk;
          }
        }
      }
      k_sab_rows<<< dim3( ( nb + 127 )/128, ne ), 128, 0, st >>>( T, rows, ainfo );
      k_sab_epoints<<< ( ne + 31 )/32, 32, 0, st >>>( T, rows, ep, bx, bpdf, bcdf, xscheck, errs );
      k_sab_guides<<< dim3( 4, ne + nb ), 256, 0, st >>>( T, ep, reinterpret_cast<uint16_t*>( base + pl.off_bguide ),
                                                       reinterpret_cast<uint16_t*>( base + pl.off_aguide ),
                                                       reinterpret_cast<double*>( base + pl.off_ascale ) );
      {
        const size_t nmax = std::max( std::max( (size_t)ne*nb, ntot ), std::max( (size_t)ne*T.bstride, (size_t)nb*( kSabGL + 1 ) ) );
        k_sab_gather_tabs<<< dim3( (unsigned)( ( nmax + 255 )/256 ), 4 ), 256, 0, st >>>(
          T, reinterpret_cast<SabHead*>( base + pl.off_heads ), reinterpret_cast<SabTail*>( base + pl.off_tails ),
          reinterpret_cast<SabPoint*>( base + pl.off_pts ), reinterpret_cast<SabBPoint*>( base + pl.off_bpts ),
          reinterpret_cast<uint16_t*>( base + pl.off_lguide ) );
      }
      g_launches += 6;
      CUDA_OK( cudaGetLastError() );
      std::vector<int> herrs( ne );
      CUDA_OK( cudaMemcpyAsync( herrs.data(), errs, (size_t)ne*4, cudaMemcpyDeviceToHost, st ) );
      CUDA_OK( cudaStreamSynchronize( st ) );
      for ( int e : herrs )
        if ( e )
          throw Err( "CalcError", "S(alpha,beta) sampler table build failed on device (code "+std::to_string(e)+")" );
      if ( pl.auto_egrid ) {
        // cross sections of the grid = the integrals of the table build; constants of SABXSProvider::setData
        // (NCSABXSProvider.cc:35-52) and SABSampler::setData (NCSABSampler.cc:41-57)
        SabT& Tm = dm.mat.sab[pl.sab_index];
        CUDA_OK( cudaMemcpy( base + pl.off_xs, xscheck, (size_t)ne*8, cudaMemcpyDeviceToDevice ) );
        double emax = 0.0, xs_emax = 0.0;
        CUDA_OK( cudaMemcpy( &emax, base + pl.off_egrid + (size_t)( ne-1 )*8, 8, cudaMemcpyDeviceToHost ) );
        CUDA_OK( cudaMemcpy( &xs_emax, xscheck + ( ne-1 ), 8, cudaMemcpyDeviceToHost ) );
        const double ext_emax = fgXS( Tm.ext, emax );
        Tm.k_extension = ( xs_emax - ext_emax )*emax;
        Tm.k1 = xs_emax*emax;
        Tm.k2 = ext_emax*emax;
      }
    }
  }

#define NCB_LIB_VDOS_HELPERS
#include "ncb_lib_vdos.inc"
#undef NCB_LIB_VDOS_HELPERS

  std::shared_ptr<DeviceMaterial> uploadMaterial( const void* blob, size_t nbytes )
  {
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount( &ndev );
    if ( ce != cudaSuccess || ndev <= 0 )
      throw Err( "CalcError", std::string("ncrystal_b200 requires a CUDA device (no CPU fallback): ")
                 + ( ce != cudaSuccess ? cudaGetErrorString(ce) : "no devices found" ) );
    // leaves delivered as a phonon density of states are expanded to S(alpha,beta) on the device first
    std::vector<unsigned char> expanded;
    if ( expandVdosLeaves( blob, nbytes, expanded ) ) { blob = expanded.data(); nbytes = expanded.size(); }
    LoadedMaterial lm;
    try {
      loadBlob( blob, nbytes, lm );
    } catch ( std::runtime_error& e ) {
      throw Err( "BadInput", e.what() );
    }
    auto dm = std::make_shared<DeviceMaterial>();
    dm->uid = ++g_material_uid_counter;
    CUDA_OK( cudaGetDevice( &dm->device ) );
    ensureErfcLut( dm->device );
    dm->arena_bytes = lm.arena.size();
    CUDA_OK( cudaMalloc( &dm->d_arena, dm->arena_bytes ) );
    CUDA_OK( cudaMemcpy( dm->d_arena, lm.arena.data(), dm->arena_bytes, cudaMemcpyHostToDevice ) );
    dm->mat = relocated( lm, dm->d_arena );
    dm->cfg = lm.cfg;
    for ( int i = 0; i < lm.mat.ncomp; ++i ) if ( lm.mat.comp[i].kind == KIND_FREEGAS ) dm->has_fg_leaf = true;
    dm->numdens = lm.numdens; dm->abs_c = lm.abs_c; dm->temperature = lm.temperature;
    dm->sabplans = lm.sabplans;
    dm->blob = std::make_shared<const std::vector<unsigned char>>( static_cast<const unsigned char*>( blob ),
                                                                   static_cast<const unsigned char*>( blob ) + nbytes );
    ensureKernelAttrs( dm->device );
    buildSabTablesOnDevice( *dm, 0 );
    buildStagePlan( *dm );     // (after the build: an energy grid determined by the library has its key lut only now)
    return dm;
  }

  void ensureKernelAttrs( int device )
  {
    static std::mutex m; static std::vector<int> done;
    std::lock_guard<std::mutex> g( m );
    for ( int d : done ) if ( d == device ) return;
    setSmemAttr( k_xs_iso );
    setSmemAttr( k_sample_classify<true> );
    setSmemAttr( k_sample_classify<false> );
    setSmemAttr( k_sample_elastic );
    setSmemAttr( k_xs_aniso );
    setSmemAttr( k_sample_aniso );
    setSmemAttr( k_xs_aniso_pre );
    setSmemAttr( k_classify_aniso );
    setSmemAttr( k_sc_scan );
    setSmemAttr( k_sc_sample );
    setSmemAttr( k_sc_sample_threads );
    setSmemAttr( k_sc_eval );
    setSmemAttr( k_sc_eval_groups );
    setSmemAttr( k_sc_eval_flat );
    setSmemAttr( k_sc_find );
    setSmemAttr( k_mmc_tail<1> );
    setSmemAttr( k_mmc_tail<2> );
    setSmemAttr( k_lc_scan );
    setSmemAttr( k_tally_hist );
    done.push_back( device );
  }

  // ------------------------------------------------------------------ handles
  constexpr uint32_t kScatterTag = 0x7d6b0637u; // same tag value as the reference's Scatter wrapper (ncrystal.cc:228)
  constexpr uint32_t kAbsorptionTag = 0xede2eb9du; // ... and its Absorption wrapper (ncrystal.cc:243)

  struct Scatter;
  struct FingerPrint { uint32_t tag; Scatter* self; };

  constexpr size_t kChunkMax = (size_t)1 << 22; // staging capacity (neutrons) per pipeline slot of the host-pointer path
  size_t chunkSize()
  {
    static const size_t c = []{ const char* e = tuneEnv( "NCB200_CHUNK" ); size_t v = e ? (size_t)std::atoll(e) : ( (size_t)1 << 20 ); return std::min( std::max<size_t>( v, 1024 ), kChunkMax ); }();
    return c;
  }
  constexpr int kSlots = 3;
  constexpr uint64_t kWindowMax = (uint64_t)1 << 24;   // neutrons staged on the device per window of the host-pointer path
  struct PipeSchedule { uint64_t first, max; double growth; };
  PipeSchedule pipeSchedule()
  {
    static const PipeSchedule ps = []{
      PipeSchedule p;
      const char* e0 = tuneEnv( "NCB200_CHUNK0" );
      const char* eg = tuneEnv( "NCB200_CHUNK_GROWTH" );
      p.max = chunkSize();
      p.first = std::min<uint64_t>( p.max, std::max<uint64_t>( 4096, e0 ? (uint64_t)std::atoll(e0) : ( (uint64_t)1 << 18 ) ) );
      p.growth = std::max( 1.0, eg ? std::atof(eg) : 2.0 );
      return p;
    }();
    return ps;
  }

  struct Scatter {
    FingerPrint fp;
    std::atomic<int> refcount{1};
    std::shared_ptr<DeviceMaterial> dm;
    uint64_t seed = 0;
    uint32_t sid = 0;
    uint64_t next_index = 0;
    int* d_err = nullptr;
    const uint32_t* n_dev_override = nullptr; // transport: device-side neutron count for the next launches
    const uint64_t* ids_override = nullptr; // transport: device array of random-stream indices for the next sampling launch
    uint32_t* last_counts_ptr = nullptr;   // queue counters of the most recent split-path launch (diagnostics)
    uint32_t* d_diag_ndraws = nullptr;
    int32_t* d_diag_comp = nullptr;
    // host-pointer pipeline resources (lazily created)
    cudaStream_t streams[kSlots] = {};      // kernel streams, one per slot
    cudaStream_t st_h2d = nullptr, st_d2h = nullptr;
    double* d_win = nullptr;                // staging window: (nin+nout) arrays of win_doubles/(nin+nout) neutrons
    size_t win_doubles = 0;
    std::vector<cudaEvent_t> ev_h, ev_c, ev_d;   // per chunk: copy-in done, kernels done, copy-out done
    double* h_ring = nullptr; size_t ring_doubles = 0;   // pinned bounce ring for pageable caller arrays
    // work queues of the split sampling path: one context per pipeline slot + one for the
    // device-pointer entry points (a handle has at most one launch sequence in flight per context)
    struct QueueCtx {
      uint32_t* q = nullptr; uint32_t* counts = nullptr; size_t cap = 0;
      // oriented path scratch (per neutron): SCBragg scan results, mu / stream position of the isotropic samplers
      double* sc_xs = nullptr; int32_t* sc_n = nullptr; double* mu_tmp = nullptr; uint32_t* nd_tmp = nullptr;
      uint32_t* q_sc = nullptr; size_t acap = 0;
      uint32_t* sc_work = nullptr; uint8_t* sc_ncand = nullptr; uint16_t* sc_cand = nullptr;   // k_sc_find -> k_sc_eval
      int32_t* sc_wpos = nullptr; bool sc_lists_valid = false;
      double* fg_prep = nullptr; uint32_t* fg_nd = nullptr; size_t fcap = 0;   // k_fg_prep records, one per queue slot
      cudaStream_t side = nullptr;              // free-gas kernels run here, concurrently with the table kernel
      cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    };
    QueueCtx qctx[kSlots+1];
    std::vector<std::unique_ptr<Scatter>> peer_handles;   // this handle's twins on other devices (multi-device fan-out)

    Scatter() { fp.tag = kScatterTag; fp.self = this; }
    ~Scatter()
    {
      if ( d_err ) cudaFree( d_err );
      for ( auto& c : qctx ) {
        if ( c.q ) cudaFree( c.q );
        if ( c.counts ) cudaFree( c.counts );
        if ( c.fg_prep ) { cudaFree( c.fg_prep ); cudaFree( c.fg_nd ); }
        if ( c.sc_xs ) { cudaFree( c.sc_xs ); cudaFree( c.sc_n ); cudaFree( c.mu_tmp ); cudaFree( c.nd_tmp ); cudaFree( c.q_sc );
                         cudaFree( c.sc_work ); cudaFree( c.sc_ncand ); cudaFree( c.sc_cand ); cudaFree( c.sc_wpos ); }
        if ( c.side ) cudaStreamDestroy( c.side );
        if ( c.ev_fork ) cudaEventDestroy( c.ev_fork );
        if ( c.ev_join ) cudaEventDestroy( c.ev_join );
      }
      for ( int s = 0; s < kSlots; ++s )
        if ( streams[s] ) cudaStreamDestroy( streams[s] );
      if ( st_h2d ) cudaStreamDestroy( st_h2d );
      if ( st_d2h ) cudaStreamDestroy( st_d2h );
      if ( d_win ) cudaFree( d_win );
      for ( auto e : ev_h ) cudaEventDestroy( e );
      for ( auto e : ev_c ) cudaEventDestroy( e );
      for ( auto e : ev_d ) cudaEventDestroy( e );
      if ( h_ring ) cudaFreeHost( h_ring );
    }
    void ensureErrWord()
    {
      if ( !d_err ) {
        CUDA_OK( cudaMalloc( &d_err, sizeof(int) ) );
        CUDA_OK( cudaMemset( d_err, 0, sizeof(int) ) );
      }
    }
    QueueCtx& ensureQueues( int ictx, size_t n )
    {
      QueueCtx& c = qctx[ictx];
      if ( !c.counts ) CUDA_OK( cudaMalloc( &c.counts, ( 8 + 2*kSortBins )*sizeof(uint32_t) ) );
      if ( !c.side ) {
        CUDA_OK( cudaStreamCreateWithFlags( &c.side, cudaStreamNonBlocking ) );
        CUDA_OK( cudaEventCreateWithFlags( &c.ev_fork, cudaEventDisableTiming ) );
        CUDA_OK( cudaEventCreateWithFlags( &c.ev_join, cudaEventDisableTiming ) );
      }
      if ( c.cap < n ) {
        if ( c.q ) { CUDA_OK( cudaDeviceSynchronize() ); CUDA_OK( cudaFree( c.q ) ); c.q = nullptr; }
        c.cap = n + n/8 + 1024;
        CUDA_OK( cudaMalloc( &c.q, 8*c.cap*sizeof(uint32_t) ) );     // (q_sab, q_fg, q_emax x2, -, fg partition, q_pb, q_el)
      }
      return c;
    }
    // per-entry records of the staged free-gas kernels (kFgSlots doubles + 1 word per queue slot); after ensureQueues
    void ensureFgPrep( QueueCtx& c )
    {
      if ( c.fcap >= c.cap ) return;
      if ( c.fg_prep ) { CUDA_OK( cudaDeviceSynchronize() ); cudaFree( c.fg_prep ); cudaFree( c.fg_nd ); }
      c.fcap = c.cap;
      CUDA_OK( cudaMalloc( &c.fg_prep, kFgSlots*c.fcap*sizeof(double) ) );
      CUDA_OK( cudaMalloc( &c.fg_nd, c.fcap*sizeof(uint32_t) ) );
    }
    QueueCtx& ensureAnisoBuffers( int ictx, size_t n )
    {
      QueueCtx& c = ensureQueues( ictx, n );
      if ( c.acap < n ) {
        if ( c.sc_xs ) {
          CUDA_OK( cudaDeviceSynchronize() );
          cudaFree( c.sc_xs ); cudaFree( c.sc_n ); cudaFree( c.mu_tmp ); cudaFree( c.nd_tmp ); cudaFree( c.q_sc );
          cudaFree( c.sc_work ); cudaFree( c.sc_ncand ); cudaFree( c.sc_cand ); cudaFree( c.sc_wpos );
        }
        c.acap = n + n/8 + 1024;
        CUDA_OK( cudaMalloc( &c.sc_xs, c.acap*sizeof(double) ) );
        CUDA_OK( cudaMalloc( &c.sc_n, c.acap*sizeof(int32_t) ) );
        CUDA_OK( cudaMalloc( &c.mu_tmp, c.acap*sizeof(double) ) );
        CUDA_OK( cudaMalloc( &c.nd_tmp, c.acap*sizeof(uint32_t) ) );
        CUDA_OK( cudaMalloc( &c.q_sc, c.acap*sizeof(uint32_t) ) );
        CUDA_OK( cudaMalloc( &c.sc_work, c.acap*sizeof(uint32_t) ) );
        CUDA_OK( cudaMalloc( &c.sc_ncand, c.acap ) );
        CUDA_OK( cudaMalloc( &c.sc_cand, c.acap*kScFindCap*sizeof(uint16_t) ) );
        CUDA_OK( cudaMalloc( &c.sc_wpos, c.acap*sizeof(int32_t) ) );
      }
      return c;
    }
    void ensurePipeline( size_t ndoubles )
    {
      for ( int s = 0; s < kSlots; ++s )
        if ( !streams[s] ) CUDA_OK( cudaStreamCreateWithFlags( &streams[s], cudaStreamNonBlocking ) );
      if ( !st_h2d ) CUDA_OK( cudaStreamCreateWithFlags( &st_h2d, cudaStreamNonBlocking ) );
      if ( !st_d2h ) CUDA_OK( cudaStreamCreateWithFlags( &st_d2h, cudaStreamNonBlocking ) );
      if ( ndoubles > win_doubles ) {
        if ( d_win ) { CUDA_OK( cudaDeviceSynchronize() ); CUDA_OK( cudaFree( d_win ) ); d_win = nullptr; }
        const size_t want = std::max( ndoubles, (size_t)3 << 20 );
        CUDA_OK( cudaMalloc( &d_win, want*sizeof(double) ) );
        win_doubles = want;
      }
    }
    void ensureChunkEvents( size_t k )
    {
      while ( ev_h.size() < k ) {
        cudaEvent_t a, b, c;
        CUDA_OK( cudaEventCreateWithFlags( &a, cudaEventDisableTiming ) );
        CUDA_OK( cudaEventCreateWithFlags( &b, cudaEventDisableTiming ) );
        CUDA_OK( cudaEventCreateWithFlags( &c, cudaEventDisableTiming ) );
        ev_h.push_back( a ); ev_c.push_back( b ); ev_d.push_back( c );
      }
    }
    void ensureBounce( size_t ndoubles )
    {
      if ( ndoubles <= ring_doubles ) return;
      if ( h_ring ) { CUDA_OK( cudaDeviceSynchronize() ); cudaFreeHost( h_ring ); h_ring = nullptr; }
      CUDA_OK( cudaHostAlloc( (void**)&h_ring, ndoubles*sizeof(double), cudaHostAllocDefault ) );
      ring_doubles = ndoubles;
    }
  };

  Scatter* fromInternal( void* internal, const char* fct )
  {
    if ( !internal )
      throw Err( "LogicError", std::string("Invalid (null) handle passed to ")+fct );
    auto fp = static_cast<FingerPrint*>( internal );
    if ( fp->tag != kScatterTag && fp->tag != kAbsorptionTag )
      throw Err( "LogicError", std::string("Invalid object handle type passed to ")+fct );
    return fp->self;
  }
  // for entry points that need a scattering process (sampling, transport)
  void requireScatter( const FingerPrint& fp, const char* fct )
  {
    if ( fp.tag != kScatterTag )
      throw Err( "LogicError", std::string("Invalid object handle type passed to ")+fct+" (absorption processes cannot be sampled)" );
  }

  struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard( int dev ) { cudaGetDevice( &prev ); if ( prev != dev ) CUDA_OK( cudaSetDevice( dev ) ); else prev = -1; }
    ~DeviceGuard() { if ( prev >= 0 ) cudaSetDevice( prev ); }
  };

  std::atomic<uint32_t> g_default_stream_counter{0};
  constexpr uint64_t kDefaultSeed = 0x4e4372797374616cULL; // "NCrystal"
  std::atomic<uint64_t> g_default_seed{ kDefaultSeed };    // ncrystal_setbuiltinrandgen[_withseed|_withstate]

  ncrystal_scatter_t newHandle( std::shared_ptr<DeviceMaterial> dm, uint64_t seed, uint32_t sid )
  {
    auto s = new Scatter;
    s->dm = std::move( dm );
    s->seed = seed;
    s->sid = sid;
    return { &s->fp };
  }

  // ------------------------------------------------------------------ material lookup
  std::mutex g_path_mtx;
  std::string g_data_path;

  std::string libDir()
  {
    Dl_info info;
    if ( dladdr( (void*)&libDir, &info ) && info.dli_fname ) {
      std::string p( info.dli_fname );
      auto pos = p.rfind( '/' );
      if ( pos != std::string::npos ) return p.substr( 0, pos );
    }
    return ".";
  }

  std::string cfgToStem( const char* cfg )
  {
    std::string out;
    for ( const char* c = cfg; *c; ++c ) {
      const char ch = *c;
      if ( ch == ' ' || ch == '\t' ) continue;
      const bool ok = ( ch >= 'a' && ch <= 'z' ) || ( ch >= 'A' && ch <= 'Z' ) || ( ch >= '0' && ch <= '9' )
                      || ch == '.' || ch == '_' || ch == '-' || ch == '=' || ch == ',' || ch == '@';
      out += ok ? ch : ( ch == ';' ? '+' : '_' );
    }
    return out;
  }

  std::vector<unsigned char> readFile( const std::string& path )
  {
    std::vector<unsigned char> d;
    FILE* f = std::fopen( path.c_str(), "rb" );
    if ( !f ) return d;
    std::fseek( f, 0, SEEK_END );
    long n = std::ftell( f );
    std::fseek( f, 0, SEEK_SET );
    if ( n > 0 ) {
      d.resize( (size_t)n );
      if ( std::fread( d.data(), 1, (size_t)n, f ) != (size_t)n ) d.clear();
    }
    std::fclose( f );
    return d;
  }

  // Order- and prefix-insensitive key of a cfg string / file stem: "stdlib::" dropped, parameters (after the data
  // name) sorted.  The reference normalises cfg strings completely (units, defaults: ncrystal_normalisecfg); material
  // setup is not restated here, so only these spelling variants of one cfg find the same compiled material.
  std::string looseCfgKey( std::string stem )
  {
    if ( stem.rfind( "stdlib__", 0 ) == 0 ) stem = stem.substr( 8 );
    std::vector<std::string> parts;
    std::stringstream ss( stem );
    std::string item;
    while ( std::getline( ss, item, '+' ) ) if ( !item.empty() ) parts.push_back( item );
    if ( parts.size() > 2 ) std::sort( parts.begin() + 1, parts.end() );
    std::string out;
    for ( auto& q : parts ) { if ( !out.empty() ) out += '+'; out += q; }
    return out;
  }

  std::vector<unsigned char> findCompiledMaterial( const char* cfg )
  {
    std::vector<std::string> dirs;
    {
      std::lock_guard<std::mutex> g( g_path_mtx );
      std::string p = g_data_path;
      if ( p.empty() ) { const char* e = std::getenv( "NCB200_DATA_PATH" ); if ( e ) p = e; }
      std::stringstream ss( p );
      std::string item;
      while ( std::getline( ss, item, ':' ) ) if ( !item.empty() ) dirs.push_back( item );
    }
    dirs.push_back( libDir() + "/../data" );
    dirs.push_back( libDir() + "/data" );
    const std::string stem = cfgToStem( cfg );
    std::string tried;
    // <stem>.ncb, else <stem>.vdos.ncb: the same material with its S(alpha,beta) leaves delivered as phonon
    // densities of states (a fraction of the size; expanded on the device when the material is created)
    for ( const char* ext : { ".ncb", ".vdos.ncb" } )
      for ( auto& d : dirs ) {
        const std::string path = d + "/" + stem + ext;
        auto data = readFile( path );
        if ( !data.empty() ) return data;
        if ( ext[1] == 'n' ) tried += " " + path;
      }
    // second chance: same cfg spelled with another parameter order or the "stdlib::" prefix
    const std::string key = looseCfgKey( stem );
    for ( auto& d : dirs ) {
      DIR* dir = opendir( d.c_str() );
      if ( !dir ) continue;
      std::string hit;
      while ( dirent* ent = readdir( dir ) ) {
        const std::string fn = ent->d_name;
        if ( fn.size() > 4 && fn.compare( fn.size()-4, 4, ".ncb" ) == 0 && looseCfgKey( fn.substr( 0, fn.size()-4 ) ) == key ) {
          hit = d + "/" + fn;
          break;
        }
      }
      closedir( dir );
      if ( !hit.empty() ) {
        auto data = readFile( hit );
        if ( !data.empty() ) return data;
      }
    }
    throw Err( "FileNotFound", std::string("No compiled material for cfg \"")+cfg+"\" (looked for:"+tried
               +"). Material setup is done by the NCrystal reference: compile it with oracle/_ref/bin/ncb200_matcompile"
               " or hand the tables over with ncb200_create_scatter_from_blob (see INTEGRATION.md)." );
  }

  // ------------------------------------------------------------------ launches
  unsigned gridFor( uint64_t n, int threads, int device, int ctas_per_sm )
  {
    const uint64_t need = ( n + threads - 1 ) / threads;
    const uint64_t cap = (uint64_t)numSMs( device ) * ctas_per_sm;
    return (unsigned)( need < cap ? ( need ? need : 1 ) : cap );
  }

  void launchXSIso( Scatter* s, const double* d_ekin, uint64_t n, double* d_out, cudaStream_t st )
  {
    if ( !n ) return;
    const DeviceMaterial& dm = *s->dm;
    if ( dm.mat.oriented )
      throw Err( "LogicError", "Process::crossSectionIsotropic can only be called for isotropic materials." );
    const int threads = 256;
    const int ctas = dm.sp.total > 56u*1024u ? 2 : 8;
    { TimedLaunch tl( "k_xs_iso", st );
      k_xs_iso<<< gridFor( n, threads, dm.device, ctas ), threads, dm.sp.total, st >>>( dm.mat, dm.sp, d_ekin, n, d_out ); }
    ++g_launches;
    CUDA_OK( cudaGetLastError() );
  }

  // Materials with a FreeGas leaf: reorder the free-gas queue by energy class (k_fg_hist / k_fg_partition);
  // afterwards Q.q_fg points at the partitioned copy.  NCB200_FG_GROUP=0 switches it off.
  void partitionFgQueue( const DeviceMaterial& dm, Scatter::QueueCtx& qc, QueueArgs& Q, const double* d_ekin, uint64_t m,
                         cudaStream_t st )
  {
    static const bool group = []{ const char* e = tuneEnv( "NCB200_FG_GROUP" ); return !e || std::atoi(e) != 0; }();
    if ( !group || !dm.has_fg_leaf ) return;
    TimedLaunch tl( "k_fg_partition", st );
    uint32_t* cls = qc.counts + 8;                  // [16] class totals + [16] cursors (the sort's histogram area)
    uint32_t* q_out = qc.q + 5*qc.cap;
    CUDA_OK( cudaMemsetAsync( cls, 0, 2*kFgGroupClasses*sizeof(uint32_t), st ) );
    k_fg_hist<<< gridFor( m, 256, dm.device, 8 ), 256, 0, st >>>( d_ekin, Q.q_fg, Q.counts + 1, cls );
    k_fg_partition<<< gridFor( ( m + 31 )/32, 256, dm.device, 4 ), 256, 0, st >>>( d_ekin, Q.q_fg, Q.counts + 1, cls, q_out );
    g_launches += 2;
    Q.q_fg = q_out;
  }

  // Free-gas queue: the staged pipeline k_fg_prep / k_fg_beta / k_fg_alpha_prep / k_fg_alpha / k_fg_finish for large
  // batches (>= NCB200_FG_STAGED_MIN neutrons, default 4e6), or the neutron-per-lane k_sample_fg (four launches and
  // two kernel tails less: measured better for the 1 Mi-neutron chunks of the host-pointer pipeline and for the
  // shrinking populations of a transport run; NCB200_FG_MODE=0 forces it).  The two refill cursors live in the (otherwise unused) sort-histogram area of the
  // counters, zeroed with them at the start of the launch sequence.
  std::atomic<uint64_t> g_fg_staged_min{ []{ const char* e = tuneEnv( "NCB200_FG_STAGED_MIN" );
                                             return e ? (uint64_t)std::atoll(e) : (uint64_t)4000000; }() };
  void launchFgSampling( Scatter* s, const DeviceMaterial& dm, Scatter::QueueCtx& qc, const SampleArgs& A, const QueueArgs& Q,
                         uint64_t m, cudaStream_t st, bool timed )
  {
    static const int mode = []{ const char* e = tuneEnv( "NCB200_FG_MODE" ); return e ? std::atoi(e) : 1; }();
    static const int fgctas = []{ const char* e = tuneEnv( "NCB200_FG_CTAS" ); return e ? std::atoi(e) : 16; }();
    static const int fgminb = []{ const char* e = tuneEnv( "NCB200_FG_MINB" ); return e ? std::atoi(e) : 8; }();
    static const int epl = []{ const char* e = tuneEnv( "NCB200_FG_EPL" ); return e ? std::atoi(e) : 4; }();
    const uint64_t minbatch = g_fg_staged_min.load();
    const unsigned nsm = (unsigned)numSMs( dm.device );
    auto timer = [&]( const char* name ) { return std::unique_ptr<TimedLaunch>( timed ? new TimedLaunch( name, st ) : nullptr ); };
    if ( mode == 0 || m < minbatch ) {
      const unsigned g = (unsigned)std::min<uint64_t>( ( m + 127 )/128, (uint64_t)nsm*fgctas );
      auto tl = timer( "k_sample_fg" );
#if defined(NCB200_TUNING) && NCB200_TUNING
      if ( fgminb < 8 ) k_sample_fg<4><<< g, 128, 0, st >>>( dm.mat, A, Q );
      else
#endif
      k_sample_fg<8><<< g, 128, 0, st >>>( dm.mat, A, Q );
      ++g_launches;
      return;
    }
    s->ensureFgPrep( qc );
    FgPrep P;
    P.r = qc.fg_prep; P.cap = qc.fcap; P.w = qc.fg_nd;
    P.cursor = Q.counts + 8 + 48;
    P.epl = (uint32_t)std::max( 1, epl );
    static const int batch = []{ const char* e = tuneEnv( "NCB200_FG_BATCH" ); return e ? std::atoi(e) : 20; }();
    P.batch = (uint32_t)std::min( 32, std::max( 1, batch ) );
    const unsigned gflat = (unsigned)std::min<uint64_t>( ( m + 127 )/128, (uint64_t)nsm*16 );
    const unsigned grefill = (unsigned)std::min<uint64_t>( ( m + 127 )/128, (uint64_t)nsm*( fgminb >= 8 ? 8 : 6 ) );
    { auto tl = timer( "k_fg_prep" );
      k_fg_prep<<< gflat, 128, 0, st >>>( dm.mat, A, Q, P ); }
    { auto tl = timer( "k_fg_beta" );
#if defined(NCB200_TUNING) && NCB200_TUNING
      if ( fgminb < 8 ) k_fg_beta<6><<< grefill, 128, 0, st >>>( dm.mat, A, Q, P );
      else
#endif
      k_fg_beta<8><<< grefill, 128, 0, st >>>( dm.mat, A, Q, P ); }
    { auto tl = timer( "k_fg_alpha_prep" );
      k_fg_alpha_prep<<< gflat, 128, 0, st >>>( dm.mat, A, Q, P ); }
    { auto tl = timer( "k_fg_alpha" );
#if defined(NCB200_TUNING) && NCB200_TUNING
      if ( fgminb < 8 ) k_fg_alpha<6><<< grefill, 128, 0, st >>>( dm.mat, A, Q, P );
      else
#endif
      k_fg_alpha<8><<< grefill, 128, 0, st >>>( dm.mat, A, Q, P ); }
    { auto tl = timer( "k_fg_finish" );
      k_fg_finish<<< gflat, 128, 0, st >>>( dm.mat, A, Q, P ); }
    g_launches += 5;
  }

  // S(alpha,beta) table queue of a launch sequence: attempt-level lane refill over the queue in arrival order.
  // (r2 experiment, dropped: partitioning the queue by overlay sampler and staging each sampler's beta tables in
  // shared memory with TMA -- the un-sorted reads/writes of the neutron arrays cost more DRAM traffic (986 MB vs
  // 371 MB per 1e7 batch) and the per-class CTA tails more time than the staged lookups saved; see DESIGN.md.)
  void launchSabQueue( Scatter* s, const DeviceMaterial& dm, Scatter::QueueCtx& qc, const SampleArgs& A, const QueueArgs& Q,
                       uint64_t m, cudaStream_t st, bool timed )
  {
    const unsigned nsm = (unsigned)numSMs( dm.device );
    static const int minb = []{ const char* e = tuneEnv( "NCB200_SAB_MINB" ); return e ? std::atoi(e) : 8; }();
    const unsigned gr = (unsigned)std::min<uint64_t>( ( m + 127 )/128, (uint64_t)nsm*minb );
    std::unique_ptr<TimedLaunch> tl( timed ? new TimedLaunch( "k_sample_sab_refill", st ) : nullptr );
#if defined(NCB200_TUNING) && NCB200_TUNING
    if ( minb == 6 ) k_sample_sab_refill<false,6><<< gr, 128, 0, st >>>( dm.mat, A, Q.q_sab, Q.counts + 0, Q.counts + 3 );
    else if ( minb == 7 ) k_sample_sab_refill<false,7><<< gr, 128, 0, st >>>( dm.mat, A, Q.q_sab, Q.counts + 0, Q.counts + 3 );
    else
#endif
    k_sample_sab_refill<false,8><<< gr, 128, 0, st >>>( dm.mat, A, Q.q_sab, Q.counts + 0, Q.counts + 3 );
    ++g_launches;
  }

  void launchSampleIso( Scatter* s, const double* d_ekin, uint64_t n, double* d_xs, double* d_eout, double* d_mu,
                        cudaStream_t st, int ictx = kSlots )
  {
    if ( !n ) return;
    requireScatter( s->fp, "sampleScatterIsotropic" );
    const DeviceMaterial& dm = *s->dm;
    if ( dm.mat.oriented )
      throw Err( "LogicError", "Process::sampleScatterIsotropic can only be called for isotropic materials." );
    s->ensureErrWord();
    uint32_t* diag_nd = s->d_diag_ndraws;
    int32_t* diag_comp = s->d_diag_comp;
    s->d_diag_ndraws = nullptr; s->d_diag_comp = nullptr;
    // Sub-launches of at most 2^26 neutrons (queue entries hold 28 index bits): bounds the scratch (index queues
    // 26 B and free-gas stage records 76 B per neutron of a sub-launch, ~7 GB) however large the batch is; results do
    // not depend on the split (streams are keyed by the global neutron index).  NCB200_SUBLAUNCH overrides (tests).
    static const uint64_t sub = []{ const char* e = tuneEnv( "NCB200_SUBLAUNCH" );
                                    const uint64_t v = e ? (uint64_t)std::atoll(e) : ( (uint64_t)1 << 26 );
                                    return std::min<uint64_t>( std::max<uint64_t>( v, 1024 ), (uint64_t)1 << kQueueIdxBits ); }();
    for ( uint64_t done = 0; done < n; done += sub ) {
      const uint64_t m = std::min<uint64_t>( sub, n - done );
      SampleArgs A;
      A.ekin = d_ekin + done; A.n = m; A.seed = s->seed; A.first_index = s->next_index + done; A.sid = s->sid;
      A.xs_out = d_xs ? d_xs + done : nullptr; A.ekin_out = d_eout + done; A.mu_out = d_mu + done;
      A.ndraws = diag_nd ? diag_nd + done : nullptr; A.component = diag_comp ? diag_comp + done : nullptr;
      A.err_flags = s->d_err;
      A.ids = s->ids_override ? s->ids_override + done : nullptr;
      const int ctas = dm.sp.total > 56u*1024u ? 2 : 8;
      Scatter::QueueCtx& qc = s->ensureQueues( ictx, m );
      QueueArgs Q;
      Q.q_sab = qc.q; Q.q_fg = qc.q + qc.cap; Q.q_emax = qc.q + 2*qc.cap; Q.counts = qc.counts;
      CUDA_OK( cudaMemsetAsync( qc.counts, 0, ( 8 + 2*kSortBins )*sizeof(uint32_t), st ) );
      // Elastic leaves: sampled in place by the classify kernel, or -- for materials whose incoherent-elastic leaf
      // has several elements (its sampler then re-evaluates the per-element contributions: the in-place code costs
      // every warp that path) -- over their own queues by k_sample_elastic.  Measured per 1e7 neutrons (classify +
      // elastic kernel against classify with in-place sampling): polyethylene 0.47 + 0.17 against 0.86 ms, YAG
      // 0.87 + 0.14 against 1.01 ms, Al (one element, 171 planes) 0.39 + 0.10 against 0.47 ms: the scattered
      // re-read / write-back of the queue kernel costs what the converged lanes save unless the paths are heavy.
      bool multi_elinc = false;
      for ( int ic = 0; ic < dm.mat.ncomp; ++ic )
        if ( dm.mat.comp[ic].kind == KIND_ELINC && dm.mat.elinc[dm.mat.comp[ic].idx].n >= 2 ) multi_elinc = true;
      const bool defer_elastic = multi_elinc && m >= ( (uint64_t)1 << 18 );
      if ( defer_elastic ) { Q.q_pb = qc.q + 6*qc.cap; Q.q_el = qc.q + 7*qc.cap; Q.counts_el = qc.counts + 8 + 64; }
      { TimedLaunch tl( "k_sample_classify", st );
        if ( defer_elastic ) k_sample_classify<true><<< gridFor( m, 256, dm.device, ctas ), 256, dm.sp.total, st >>>( dm.mat, dm.sp, A, Q );
        else k_sample_classify<false><<< gridFor( m, 256, dm.device, ctas ), 256, dm.sp.total, st >>>( dm.mat, dm.sp, A, Q ); }
      ++g_launches;
      if ( defer_elastic ) {
        TimedLaunch tl( "k_sample_elastic", st );
        k_sample_elastic<<< dim3( gridFor( m, 256, dm.device, ctas ), 2 ), 256, dm.sp.total, st >>>( dm.mat, dm.sp, A, Q );
        ++g_launches;
      }
      launchSabQueue( s, dm, qc, A, Q, m, st, true );
      partitionFgQueue( dm, qc, Q, A.ekin, m, st );
      launchFgSampling( s, dm, qc, A, Q, m, st, true );
      { TimedLaunch tl( "k_sample_sab_refill_emax", st );
        const unsigned gfg = (unsigned)std::min<uint64_t>( ( m + 127 )/128, (uint64_t)numSMs( dm.device )*4 );
        k_sample_sab_refill<true,5><<< gfg, 128, 0, st >>>( dm.mat, A, Q.q_emax, Q.counts + 2, Q.counts + 4 ); }
      ++g_launches;
      s->last_counts_ptr = Q.counts;
      CUDA_OK( cudaGetLastError() );
    }
    s->next_index += n;
  }

  bool useAnisoV1()
  {
    static const bool v1 = []{ const char* e = tuneEnv( "NCB200_ANISO_V1" ); return e && *e && *e != '0'; }();
    return v1;
  }

  int scCompIndex( const Material& M )
  {
    for ( int i = 0; i < M.ncomp; ++i ) if ( M.comp[i].kind == KIND_SCBRAGG ) return i;
    return -1;
  }

  int lcCompIndex( const Material& M )
  {
    for ( int i = 0; i < M.ncomp; ++i ) if ( M.comp[i].kind == KIND_LCBRAGG ) return i;
    return -1;
  }
  // layered-crystal plane sets whose per-warp ROI lists fit shared memory (else: thread-per-neutron kernels)
  bool lcWarpOk( const Material& M ) { return lcSmemBytes( M.lc.nplanes ) <= 160u*1024u; }

  // Batches below this size (the long tail of a transport run) are launch-latency bound: the all-in-one
  // thread-per-neutron kernels (1 launch instead of 2-8) are used for them.  NCB200_SMALL_V1 overrides (0 = never).
  uint64_t smallBatchV1()
  {
    static const uint64_t v = []{ const char* e = tuneEnv( "NCB200_SMALL_V1" ); return e ? (uint64_t)std::atoll(e) : (uint64_t)0; }();
    return v;
  }

  // SCBragg scan (one warp per neutron) -> sc_xs / sc_n.  Default: lean candidate search (k_sc_find) + evaluation of
  // the neutrons that have candidates (k_sc_eval); NCB200_SC_ONEKERNEL=1 selects the combined k_sc_scan.
  void launchScScan( const DeviceMaterial& dm, Scatter::QueueCtx& qc, const double* d_ekin, const double* ux,
                     const double* uy, const double* uz, uint64_t n, cudaStream_t st, const uint32_t* n_dev = nullptr )
  {
    static const bool onekernel = []{ const char* e = tuneEnv( "NCB200_SC_ONEKERNEL" ); return e && std::atoi(e) != 0; }();
    const int isc = scCompIndex( dm.mat );
    const uint64_t need = ( n + kScWarps - 1 ) / kScWarps;
    const unsigned nsm = (unsigned)numSMs( dm.device );
    // (small batches -- the tail of a transport run -- are launch-latency bound: one kernel instead of three launches)
    qc.sc_lists_valid = false;
    if ( onekernel || n < 32768 || n_dev || dm.sc_find_smem > 220u*1024u ) {
      ScScanArgs SA;
      SA.ekin = d_ekin; SA.ux = ux; SA.uy = uy; SA.uz = uz; SA.n = n; SA.sc_xs = qc.sc_xs; SA.sc_n = qc.sc_n;
      SA.dom_lo = dm.mat.comp[isc].dom_lo; SA.dom_hi = dm.mat.comp[isc].dom_hi;
      SA.n_dev = n_dev;
      const int ctas = std::max( 1, (int)( ( 200u*1024u ) / std::max( dm.sc_smem, 1u ) ) );
      const unsigned grid = (unsigned)std::min<uint64_t>( need, (uint64_t)nsm*std::min( ctas, 8 ) );
      { TimedLaunch tl( "k_sc_scan", st );
        k_sc_scan<<< grid, 32*kScWarps, dm.sc_smem, st >>>( dm.mat, dm.sp_sc, SA, dm.sc_famof_off, dm.sc_scratch_off ); }
      ++g_launches;
      CUDA_OK( cudaGetLastError() );
      return;
    }
    ScFindArgs FA;
    FA.ekin = d_ekin; FA.ux = ux; FA.uy = uy; FA.uz = uz; FA.n = n;
    FA.dom_lo = dm.mat.comp[isc].dom_lo; FA.dom_hi = dm.mat.comp[isc].dom_hi;
    FA.sc_xs = qc.sc_xs; FA.sc_n = qc.sc_n;
    FA.work = qc.sc_work; FA.work_count = qc.counts + 6; FA.overflow_count = qc.counts + 7; FA.ncand = qc.sc_ncand; FA.cand = qc.sc_cand;
    FA.wpos = qc.sc_wpos;
    qc.sc_lists_valid = true;
    CUDA_OK( cudaMemsetAsync( qc.counts + 6, 0, 2*sizeof(uint32_t), st ) );
    const int cf = std::min( 3, std::max( 1, (int)( ( 220u*1024u ) / std::max( dm.sc_find_smem, 1u ) ) ) );
    const uint64_t need_f = ( n + kScFindWarps - 1 ) / kScFindWarps;
    const int ce = std::min( 2, std::max( 1, (int)( ( 200u*1024u ) / std::max( dm.sc_smem, 1u ) ) ) );
    { TimedLaunch tl( "k_sc_find", st );
      k_sc_find<<< (unsigned)std::min<uint64_t>( need_f, (uint64_t)nsm*cf ), 32*kScFindWarps, dm.sc_find_smem, st >>>(
        dm.mat, dm.sp_find, FA, dm.sc_find_famof_off, dm.sc_find_scratch_off ); }
    // evaluation: eight lanes per neutron; the (rare) neutrons with more candidates than the record holds are
    // flagged in the work list (the find kernel counts them) and walked again by the warp-per-neutron kernel
    { TimedLaunch tl( "k_sc_eval", st );
      static const bool groups = []{ const char* e = tuneEnv( "NCB200_SC_EVAL_GROUPS" ); return e && std::atoi(e) != 0; }();
      if ( groups ) {       // (r2 first form: eight lanes per neutron)
        const uint32_t gsmem = dm.sc_scratch_off + (uint32_t)( kScWarps*4*sizeof(ScGroupScratch) );
        const int cg = std::min( 2, std::max( 1, (int)( ( 200u*1024u ) / std::max( gsmem, 1u ) ) ) );
        k_sc_eval_groups<<< (unsigned)std::min<uint64_t>( ( need + 3 )/4, (uint64_t)nsm*cg ), 32*kScWarps, gsmem, st >>>(
          dm.mat, dm.sp_sc, FA, dm.sc_famof_off, dm.sc_scratch_off );
      } else {              // one candidate per lane
        const uint32_t fsmem = dm.sc_scratch_off + (uint32_t)( kScWarps*sizeof(ScFlatScratch) );
        const int cg = std::min( 2, std::max( 1, (int)( ( 200u*1024u ) / std::max( fsmem, 1u ) ) ) );
        k_sc_eval_flat<<< (unsigned)std::min<uint64_t>( ( need + 31 )/32, (uint64_t)nsm*cg ), 32*kScWarps, fsmem, st >>>(
          dm.mat, dm.sp_sc, FA, dm.sc_famof_off, dm.sc_scratch_off );
      } }
    { TimedLaunch tl( "k_sc_eval_overflow", st );
      k_sc_eval<<< (unsigned)std::min<uint64_t>( need, (uint64_t)nsm*ce ), 32*kScWarps, dm.sc_smem, st >>>(
        dm.mat, dm.sp_sc, FA, dm.sc_famof_off, dm.sc_scratch_off, 1 ); }
    g_launches += 3;
    CUDA_OK( cudaGetLastError() );
  }

  // LCBragg scan (one warp per neutron) -> sc_xs (sum over the ROIs) / sc_n (number of ROIs)
  void launchLcScan( Scatter* s, const DeviceMaterial& dm, Scatter::QueueCtx& qc, const double* d_ekin, const double* ux,
                     const double* uy, const double* uz, uint64_t n, cudaStream_t st, const uint32_t* n_dev = nullptr,
                     const SampleArgs* pick = nullptr )
  {
    const int ilc = lcCompIndex( dm.mat );
    s->ensureErrWord();
    LcScanArgs LA;
    LA.ekin = d_ekin; LA.ux = ux; LA.uy = uy; LA.uz = uz; LA.n = n; LA.lc_sum = qc.sc_xs; LA.lc_n = qc.sc_n;
    LA.dom_lo = dm.mat.comp[ilc].dom_lo; LA.dom_hi = dm.mat.comp[ilc].dom_hi;
    LA.err_flags = s->d_err; LA.n_dev = n_dev;
    if ( pick ) {
      // (a material has either a single-crystal or a layered-crystal leaf: the candidate buffer of the former, 64
      //  bytes per neutron, holds the chosen ranges of the latter)
      static_assert( sizeof(LcRoi) <= kScFindCap*sizeof(uint16_t), "chosen rotation ranges must fit the candidate buffer" );
      LA.pick_roi = reinterpret_cast<LcRoi*>( qc.sc_cand );
      LA.seed = pick->seed; LA.first_index = pick->first_index; LA.sid = pick->sid; LA.ids = pick->ids;
      LA.pick_draw = dm.mat.ncomp > 1 ? 1u : 0u;
    }
    qc.sc_lists_valid = false;
    const unsigned nsm = (unsigned)numSMs( dm.device );
    const uint64_t need = ( n + kLcWarps - 1 ) / kLcWarps;
    const unsigned grid = (unsigned)std::min<uint64_t>( need, (uint64_t)nsm*16 );
    { TimedLaunch tl( "k_lc_scan", st );
      k_lc_scan<<< grid, 32*kLcWarps, lcSmemBytes( dm.mat.lc.nplanes ), st >>>( dm.mat, LA ); }
    ++g_launches;
    CUDA_OK( cudaGetLastError() );
  }

  void launchXSAniso( Scatter* s, const double* d_ekin, const double* ux, const double* uy, const double* uz,
                      uint64_t n, double* d_out, cudaStream_t st, int ictx = kSlots )
  {
    if ( !n ) return;
    const DeviceMaterial& dm = *s->dm;
    const bool has_sc = scCompIndex( dm.mat ) >= 0;
    const bool has_lc = lcCompIndex( dm.mat ) >= 0;
    if ( useAnisoV1() || ( has_sc && !dm.sc_warp_ok ) || ( has_lc && !lcWarpOk( dm.mat ) ) || n < smallBatchV1() ) {
      DirArgs D; D.ux = ux; D.uy = uy; D.uz = uz; D.ox = D.oy = D.oz = nullptr;
      const int ctas = dm.sp.total > 56u*1024u ? 2 : 8;
      k_xs_aniso<<< gridFor( n, 128, dm.device, ctas ), 128, dm.sp.total, st >>>( dm.mat, dm.sp, d_ekin, D, n, d_out );
      ++g_launches;
      CUDA_OK( cudaGetLastError() );
      return;
    }
    const double* sc_xs = nullptr; const int32_t* sc_n = nullptr;
    if ( has_sc || has_lc ) {
      Scatter::QueueCtx& qc = s->ensureAnisoBuffers( ictx, n );
      if ( has_sc )
        launchScScan( dm, qc, d_ekin, ux, uy, uz, n, st, s->n_dev_override );
      else
        launchLcScan( s, dm, qc, d_ekin, ux, uy, uz, n, st, s->n_dev_override );
      sc_xs = qc.sc_xs; sc_n = qc.sc_n;
    }
    const int ctas = dm.sp_iso.total > 56u*1024u ? 2 : 8;
    { TimedLaunch tl( "k_xs_aniso_pre", st );
      k_xs_aniso_pre<<< gridFor( n, 256, dm.device, ctas ), 256, dm.sp_iso.total, st >>>( dm.mat, dm.sp_iso, d_ekin, sc_xs, sc_n, n, d_out, s->n_dev_override ); }
    ++g_launches;
    CUDA_OK( cudaGetLastError() );
  }

  void launchSampleAniso( Scatter* s, const double* d_ekin, const double* ux, const double* uy, const double* uz,
                          uint64_t n, double* d_eout, double* ox, double* oy, double* oz, cudaStream_t st, int ictx = kSlots )
  {
    if ( !n ) return;
    requireScatter( s->fp, "sampleScatter" );
    const DeviceMaterial& dm = *s->dm;
    s->ensureErrWord();
    uint32_t* diag_nd = s->d_diag_ndraws;
    int32_t* diag_comp = s->d_diag_comp;
    s->d_diag_ndraws = nullptr; s->d_diag_comp = nullptr;
    const bool has_sc = scCompIndex( dm.mat ) >= 0;
    const bool has_lc = lcCompIndex( dm.mat ) >= 0;
    const bool v1 = useAnisoV1() || ( has_sc && !dm.sc_warp_ok ) || ( has_lc && !lcWarpOk( dm.mat ) ) || n < smallBatchV1();
    const uint64_t maxn = v1 ? n : ( (uint64_t)1 << kQueueIdxBits );
    for ( uint64_t done = 0; done < n; done += maxn ) {
      const uint64_t m = std::min<uint64_t>( maxn, n - done );
      SampleArgs A;
      A.ekin = d_ekin + done; A.n = m; A.seed = s->seed; A.first_index = s->next_index + done; A.sid = s->sid;
      A.xs_out = nullptr; A.ekin_out = d_eout + done; A.mu_out = nullptr;
      A.ndraws = diag_nd ? diag_nd + done : nullptr; A.component = diag_comp ? diag_comp + done : nullptr;
      A.err_flags = s->d_err;
      A.ids = s->ids_override ? s->ids_override + done : nullptr;
      A.n_dev = s->n_dev_override;
      DirArgs D; D.ux = ux + done; D.uy = uy + done; D.uz = uz + done; D.ox = ox + done; D.oy = oy + done; D.oz = oz + done;
      if ( v1 ) {
        const int ctas = dm.sp.total > 56u*1024u ? 2 : 4;
        k_sample_aniso<<< gridFor( m, 128, dm.device, ctas ), 128, dm.sp.total, st >>>( dm.mat, dm.sp, A, D );
        ++g_launches;
        CUDA_OK( cudaGetLastError() );
        continue;
      }
      Scatter::QueueCtx& qc = s->ensureAnisoBuffers( ictx, m );
      if ( has_sc )
        launchScScan( dm, qc, A.ekin, D.ux, D.uy, D.uz, m, st, s->n_dev_override );
      if ( has_lc )
        launchLcScan( s, dm, qc, A.ekin, D.ux, D.uy, D.uz, m, st, s->n_dev_override, &A );
      QueueArgs Q;
      Q.q_sab = qc.q; Q.q_fg = qc.q + qc.cap; Q.q_emax = qc.q + 2*qc.cap; Q.counts = qc.counts;
      AnisoArgs X;
      X.D = D; X.sc_xs = ( has_sc || has_lc ) ? qc.sc_xs : nullptr; X.sc_n = ( has_sc || has_lc ) ? qc.sc_n : nullptr;
      X.mu_tmp = qc.mu_tmp; X.nd_tmp = qc.nd_tmp; X.q_sc = qc.q_sc; X.q_sc_count = qc.counts + 5;
      if ( has_sc && qc.sc_lists_valid ) { X.sc_wpos = qc.sc_wpos; X.sc_ncand = qc.sc_ncand; X.sc_cand = qc.sc_cand; }
      CUDA_OK( cudaMemsetAsync( qc.counts, 0, ( 8 + 2*kSortBins )*sizeof(uint32_t), st ) );
      const int ctas = dm.sp_iso.total > 56u*1024u ? 2 : 8;
      { TimedLaunch tl( "k_classify_aniso", st );
        k_classify_aniso<<< gridFor( m, 256, dm.device, ctas ), 256, dm.sp_iso.total, st >>>( dm.mat, dm.sp_iso, A, Q, X ); }
      // isotropic leaves: same queue kernels as the isotropic path; they leave (E', mu, stream position)
      SampleArgs Ai = A;
      Ai.mu_out = qc.mu_tmp; Ai.ndraws = qc.nd_tmp; Ai.component = nullptr;
      const unsigned nsm = (unsigned)numSMs( dm.device );
      const unsigned gq = (unsigned)std::min<uint64_t>( ( m + 127 )/128, (uint64_t)nsm*8 );
      launchSabQueue( s, dm, qc, Ai, Q, m, st, true );
      if ( m >= 65536 ) partitionFgQueue( dm, qc, Q, A.ekin, m, st );
      launchFgSampling( s, dm, qc, Ai, Q, m, st, true );
      { TimedLaunch tl( "k_sample_sab_refill_emax", st );
        k_sample_sab_refill<true,5><<< gq, 128, 0, st >>>( dm.mat, Ai, Q.q_emax, Q.counts + 2, Q.counts + 4 ); }
      { TimedLaunch tl( "k_dir_from_mu", st );
        k_dir_from_mu<<< gridFor( m, 256, dm.device, 8 ), 256, 0, st >>>( A, Q, X ); }
      g_launches += 3;
      if ( has_sc ) {
        const unsigned gs = (unsigned)std::min<uint64_t>( ( m + kScWarps - 1 )/kScWarps, (uint64_t)nsm*3 );
        // neutrons with a recorded candidate list: one per thread; the others (none, normally): one per warp
        uint32_t* n_left = qc.counts + 7;
        CUDA_OK( cudaMemsetAsync( n_left, 0, sizeof(uint32_t), st ) );
        if ( X.sc_wpos ) {
          const uint32_t tsmem = dm.sc_famof_off + ( ( (uint32_t)dm.mat.sc.nnormals + 127u ) & ~127u );
          const unsigned gt = (unsigned)std::min<uint64_t>( ( m + 32*kScWarps - 1 )/( 32*kScWarps ), (uint64_t)nsm*2 );
          { TimedLaunch tl( "k_sc_sample", st );
            k_sc_sample_threads<<< gt, 32*kScWarps, tsmem, st >>>( dm.mat, dm.sp_sc, A, X, dm.sc_famof_off, n_left ); }
          { TimedLaunch tl( "k_sc_sample_unlisted", st );
            k_sc_sample<<< gs, 32*kScWarps, dm.sc_smem, st >>>( dm.mat, dm.sp_sc, A, X, dm.sc_famof_off, dm.sc_scratch_off, 1, n_left ); }
          g_launches += 2;
        } else {
          { TimedLaunch tl( "k_sc_sample", st );
            k_sc_sample<<< gs, 32*kScWarps, dm.sc_smem, st >>>( dm.mat, dm.sp_sc, A, X, dm.sc_famof_off, dm.sc_scratch_off, 0, n_left ); }
          ++g_launches;
        }
      }
      if ( has_lc ) {
        { TimedLaunch tl( "k_lc_sample", st );
          k_lc_sample_threads<<< gridFor( m, 128, dm.device, 8 ), 128, 0, st >>>( dm.mat, A, X, reinterpret_cast<const LcRoi*>( qc.sc_cand ) ); }
        ++g_launches;
      }
      CUDA_OK( cudaGetLastError() );
    }
    s->next_index += n;
  }

  int fetchDeviceErrors( Scatter* s, cudaStream_t st )
  {
    CUDA_OK( cudaStreamSynchronize( st ) );
    if ( !s->d_err ) return 0;
    int flags = 0;
    CUDA_OK( cudaMemcpy( &flags, s->d_err, sizeof(int), cudaMemcpyDeviceToHost ) );
    if ( flags )
      CUDA_OK( cudaMemset( s->d_err, 0, sizeof(int) ) );
    if ( flags & ERR_SAB_ISOFALLBACK ) {
      // the reference's warning (NCSABSamplerModels.cc:96-105), through the message handler; at most 20 times.
      // (One message per call in which the fallback happened: the device reports a flag, not a count.)
      static std::atomic<unsigned> s_nfail{0};
      const unsigned nfail = ++s_nfail;
      if ( nfail <= 20 )
        emitMsg( nfail == 20 ? "SABSampler reverts to isotropic model after 30 rejected attempts (suppressing further warnings of this type)"
                             : "SABSampler reverts to isotropic model after 30 rejected attempts", 1 );
    }
    return flags;
  }

  void raiseDeviceErrors( int flags )
  {
    // same conditions under which the reference throws CalcError / BadInput
    if ( flags & ERR_SAB_DISCARD )
      throw Err( "BadInput", "Scattering Kernel does not appear to match up very well with the chosen extrapolation model at Emax." );
    if ( flags & ERR_SAB_LOOP_INNER )
      throw Err( "CalcError", "Rejection method failed to sample kinematically valid (alpha,beta) point after 100 attempts." );
    if ( flags & ERR_SAB_LOOP_OUTER )
      throw Err( "CalcError", "Infinite looping in sampleAlphaBeta" );
    if ( flags & ERR_KIN_DENOM )
      throw Err( "CalcError", "convertAlphaBetaToDeltaEMu invalid for beta=-E/kT" );
    if ( flags & ERR_MMC_NOTERM )
      throw Err( "CalcError", "transport did not terminate" );
    if ( flags & ERR_LC_ROMBERG )
      throw Err( "CalcError", "Romberg integration did not converge." );
  }

  // ---- pageable caller buffers.  Real callers of the *_many entry points (OpenMC, McStas, the reference's Python
  // layer) pass malloc'd arrays; cudaMemcpyAsync from pageable memory is staged by the driver on one thread and
  // serialises the pipeline (r1: measured 0.33e9 neutrons/s against 1.6e9 from pinned buffers).  Arrays that are
  // not page-locked therefore go through a pinned bounce ring of this handle: a small pool of host threads copies
  // chunk k+1 into the ring while chunk k is on the bus, and a drain thread copies finished chunks out.
  class CopyPool {
  public:
    static CopyPool& instance() { static CopyPool p; return p; }
    // copy with all pool threads + the caller; returns when done
    void copy( void* dst, const void* src, size_t bytes )
    {
      const size_t piece = (size_t)1 << 20;
      if ( bytes <= piece || workers_.empty() ) { std::memcpy( dst, src, bytes ); return; }
      const size_t nparts = std::min<size_t>( workers_.size() + 1, ( bytes + piece - 1 )/piece );
      const size_t per = ( ( bytes + nparts - 1 )/nparts + 63 ) & ~(size_t)63;
      Batch bt; bt.left = 0;
      {
        std::lock_guard<std::mutex> g( m_ );
        for ( size_t o = per; o < bytes; o += per ) {
          tasks_.push_back( Task{ static_cast<char*>( dst ) + o, static_cast<const char*>( src ) + o, std::min( per, bytes - o ), &bt } );
          ++bt.left;
        }
      }
      cv_.notify_all();
      std::memcpy( dst, src, std::min( per, bytes ) );
      std::unique_lock<std::mutex> lk( m_ );
      while ( bt.left ) {
        if ( !tasks_.empty() ) {            // help instead of waiting
          Task t = tasks_.front(); tasks_.pop_front();
          lk.unlock(); std::memcpy( t.dst, t.src, t.n ); lk.lock();
          if ( --t.batch->left == 0 ) done_.notify_all();
        } else {
          done_.wait( lk );
        }
      }
    }
  private:
    struct Batch { size_t left; };
    struct Task { char* dst; const char* src; size_t n; Batch* batch; };
    CopyPool()
    {
      unsigned hw = std::thread::hardware_concurrency();
      if ( const char* e = std::getenv( "NCB200_COPY_THREADS" ) ) hw = (unsigned)std::max( 0, std::atoi( e ) ) + 2u;
      // measured on the 16-core GPU box: memcpy pageable->pinned 14.5 GB/s with 1 thread, 46 with 8, 50 with 16 (the
      // ceiling of the path); end to end 7.7e8 / 1.0e9 / 8.1e8 neutrons/s with 3 / 7 / 15 pool threads
      const unsigned nw = std::min( 7u, hw > 3 ? hw/2 - 1 : 0u );
      for ( unsigned i = 0; i < nw; ++i )
        workers_.emplace_back( [this]{ run(); } );
    }
    ~CopyPool()
    {
      { std::lock_guard<std::mutex> g( m_ ); stop_ = true; }
      cv_.notify_all();
      for ( auto& t : workers_ ) t.join();
    }
    void run()
    {
      std::unique_lock<std::mutex> lk( m_ );
      while ( true ) {
        cv_.wait( lk, [this]{ return stop_ || !tasks_.empty(); } );
        if ( stop_ ) return;
        Task t = tasks_.front(); tasks_.pop_front();
        lk.unlock(); std::memcpy( t.dst, t.src, t.n ); lk.lock();
        if ( --t.batch->left == 0 ) done_.notify_all();
      }
    }
    std::mutex m_;
    std::condition_variable cv_, done_;
    std::deque<Task> tasks_;
    std::vector<std::thread> workers_;
    bool stop_ = false;
  };

  bool isPageLocked( const void* p, size_t bytes )
  {
    auto one = []( const void* q ) {
      cudaPointerAttributes a;
      if ( cudaPointerGetAttributes( &a, q ) != cudaSuccess ) { cudaGetLastError(); return false; }
      return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
    };
    return one( p ) && one( static_cast<const char*>( p ) + ( bytes ? bytes - 1 : 0 ) );
  }
  constexpr uint64_t kBounceMin = (uint64_t)1 << 16;   // shorter calls: the driver's own staging is as good
  constexpr int kRingSlots = 6;                        // bounce-ring depth (chunks): the host copies run ahead of / behind the bus

  // Host-pointer pipeline.  The call's arrays are staged in one device window (<= kWindowMax neutrons; longer
  // calls run window after window).  Three kinds of streams: one for H2D copies, kSlots for the kernels
  // (round robin, so the tail of one chunk's rejection kernels overlaps the next chunk), one for D2H copies;
  // events order chunk k's copy-in -> kernels -> copy-out, the host only blocks at the end of a window (and, for
  // pageable arrays, on the bounce ring).  Chunks start small (the first D2H starts early) and grow geometrically
  // up to chunkSize() (launch efficiency).
  // `launch(chunk_n, in_dev[], out_dev[], stream, slot)` enqueues the kernel(s).
  void runHostPipeline( Scatter* s, uint64_t n, int nin, const double* const* in, int nout, double* const* out,
                        const std::function<void(uint64_t,double* const*,double* const*,cudaStream_t,int)>& launch )
  {
    if ( !n ) return;
    const uint64_t W = std::min<uint64_t>( n, kWindowMax );
    s->ensurePipeline( (size_t)W * (size_t)( nin + nout ) );
    const PipeSchedule ps = pipeSchedule();
    // which arrays need the bounce ring
    bool bin[4] = {}, bout[4] = {};
    bool any_in = false, any_out = false;
    if ( n >= kBounceMin ) {
      for ( int a = 0; a < nin; ++a ) { bin[a] = !isPageLocked( in[a], n*sizeof(double) ); any_in |= bin[a]; }
      for ( int a = 0; a < nout; ++a ) { bout[a] = !isPageLocked( out[a], n*sizeof(double) ); any_out |= bout[a]; }
    }
    const size_t cmax = (size_t)ps.max;
    if ( any_in || any_out ) s->ensureBounce( (size_t)kRingSlots*( nin + nout )*cmax );
    auto ringIn = [&]( int slot, int a ) { return s->h_ring + ( (size_t)slot*( nin + nout ) + a )*cmax; };
    auto ringOut = [&]( int slot, int a ) { return s->h_ring + ( (size_t)slot*( nin + nout ) + nin + a )*cmax; };

    // drain thread: chunk by chunk, waits for the D2H copies into the ring and copies them out to the caller
    struct OutJob { size_t k; int slot; uint64_t off, m; };
    std::mutex dm; std::condition_variable dcv;
    std::deque<OutJob> djobs; bool dstop = false; size_t ddone = 0;   // ddone: chunks copied out so far
    std::thread drainer;
    const int device = s->dm->device;
    if ( any_out )
      drainer = std::thread( [&]{
        cudaSetDevice( device );
        while ( true ) {
          OutJob j;
          {
            std::unique_lock<std::mutex> lk( dm );
            dcv.wait( lk, [&]{ return dstop || !djobs.empty(); } );
            if ( djobs.empty() ) return;
            j = djobs.front(); djobs.pop_front();
          }
          cudaEventSynchronize( s->ev_d[j.k] );
          for ( int a = 0; a < nout; ++a )
            if ( bout[a] ) CopyPool::instance().copy( out[a] + j.off, ringOut( j.slot, a ), j.m*sizeof(double) );
          { std::lock_guard<std::mutex> g( dm ); ddone = j.k + 1; }
          dcv.notify_all();
        }
      } );
    // A failure while work is queued must not return to the caller (who then fills the output arrays with the
    // error sentinels) before the copies already queued into those arrays have drained.
    struct Drain {
      Scatter* s; std::thread* t; std::mutex* m; std::condition_variable* cv; bool* stop; bool armed = true;
      void finish()
      {
        if ( t->joinable() ) {
          { std::lock_guard<std::mutex> g( *m ); *stop = true; }
          cv->notify_all();
          t->join();
        }
      }
      ~Drain() {
        if ( armed ) {
          cudaStreamSynchronize( s->st_h2d );
          for ( int c = 0; c < kSlots; ++c ) cudaStreamSynchronize( s->streams[c] );
          cudaStreamSynchronize( s->st_d2h );
        }
        finish();
      }
    } drain{ s, &drainer, &dm, &dcv, &dstop };
    for ( uint64_t w0 = 0; w0 < n; w0 += W ) {
      const uint64_t wn = std::min<uint64_t>( W, n - w0 );
      uint64_t done = 0;
      size_t k = 0;
      double chunk = (double)ps.first;
      while ( done < wn ) {
        uint64_t m = std::min<uint64_t>( (uint64_t)chunk, wn - done );
        if ( wn - done - m < ps.first/4 && wn - done <= cmax ) m = wn - done;      // no tiny last chunk
        chunk = std::min<double>( chunk*ps.growth, (double)ps.max );
        s->ensureChunkEvents( k + 1 );
        const int slot = (int)( k % kSlots );
        const int rslot = (int)( k % kRingSlots );
        cudaStream_t cs = s->streams[slot];
        double* din[4]; double* dout[4];
        if ( k >= (size_t)kRingSlots ) {
          // ring slot reuse: its previous H2D must be done, its previous chunk copied out
          if ( any_in ) CUDA_OK( cudaEventSynchronize( s->ev_h[k - kRingSlots] ) );
          if ( any_out ) { std::unique_lock<std::mutex> lk( dm ); dcv.wait( lk, [&]{ return ddone >= k - kRingSlots + 1; } ); }
        }
        for ( int a = 0; a < nin; ++a ) {
          din[a] = s->d_win + (size_t)a*W + done;
          const double* src = in[a] + w0 + done;
          if ( bin[a] ) { CopyPool::instance().copy( ringIn( rslot, a ), src, m*sizeof(double) ); src = ringIn( rslot, a ); }
          CUDA_OK( cudaMemcpyAsync( din[a], src, m*sizeof(double), cudaMemcpyHostToDevice, s->st_h2d ) );
        }
        CUDA_OK( cudaEventRecord( s->ev_h[k], s->st_h2d ) );
        for ( int a = 0; a < nout; ++a )
          dout[a] = s->d_win + (size_t)(nin+a)*W + done;
        CUDA_OK( cudaStreamWaitEvent( cs, s->ev_h[k], 0 ) );
        launch( m, din, dout, cs, slot );
        CUDA_OK( cudaEventRecord( s->ev_c[k], cs ) );
        CUDA_OK( cudaStreamWaitEvent( s->st_d2h, s->ev_c[k], 0 ) );
        for ( int a = 0; a < nout; ++a )
          CUDA_OK( cudaMemcpyAsync( bout[a] ? ringOut( rslot, a ) : out[a] + w0 + done, dout[a], m*sizeof(double),
                                    cudaMemcpyDeviceToHost, s->st_d2h ) );
        if ( any_out ) {
          CUDA_OK( cudaEventRecord( s->ev_d[k], s->st_d2h ) );
          { std::lock_guard<std::mutex> g( dm ); djobs.push_back( OutJob{ k, rslot, w0 + done, m } ); }
          dcv.notify_all();
        }
        done += m;
        ++k;
      }
      CUDA_OK( cudaStreamSynchronize( s->st_d2h ) );
      for ( int c = 0; c < kSlots; ++c )
        CUDA_OK( cudaStreamSynchronize( s->streams[c] ) );
      if ( any_out ) {
        std::unique_lock<std::mutex> lk( dm );
        dcv.wait( lk, [&]{ return ddone >= k; } );
        ddone = 0;      // chunk numbering restarts with the next window
      }
    }
    drain.armed = false;
  }

#include "ncb_lib_multigpu.inc"

  void xsIsoHost( Scatter* s, const double* ekin, uint64_t n, uint64_t repeat, double* results )
  {
    if ( !n || !repeat ) return;
    const auto devs = fanDevices( s, n );
    if ( !devs.empty() ) {
      fanOut( s, devs, n, false, [&]( Scatter* h, uint64_t b, uint64_t m ) { xsIsoHost( h, ekin + b, m, 1, results + b ); } );
      for ( uint64_t r = 1; r < repeat; ++r )
        std::memcpy( results + r*n, results, n*sizeof(double) );
      return;
    }
    DeviceGuard dg( s->dm->device );
    const double* in[1] = { ekin };
    double* out[1] = { results };
    runHostPipeline( s, n, 1, in, 1, out, [s]( uint64_t m, double* const* di, double* const* dout, cudaStream_t st, int ) {
      launchXSIso( s, di[0], m, dout[0], st );
    } );
    // deterministic: further repeats are copies (ref loop order: results[r*n+i], ncrystal.cc:1125-1133)
    for ( uint64_t r = 1; r < repeat; ++r )
      std::memcpy( results + r*n, results, n*sizeof(double) );
  }

  void sampleIsoHost( Scatter* s, const double* ekin, uint64_t n, uint64_t repeat, double* eout, double* mu )
  {
    if ( !n || !repeat ) return;
    const auto devs = fanDevices( s, n );
    if ( !devs.empty() ) {
      for ( uint64_t r = 0; r < repeat; ++r )
        fanOut( s, devs, n, true, [&]( Scatter* h, uint64_t b, uint64_t m ) { sampleIsoHost( h, ekin + b, m, 1, eout + r*n + b, mu + r*n + b ); } );
      return;
    }
    DeviceGuard dg( s->dm->device );
    for ( uint64_t r = 0; r < repeat; ++r ) {
      const double* in[1] = { ekin };
      double* out[2] = { eout + r*n, mu + r*n };
      runHostPipeline( s, n, 1, in, 2, out, [s]( uint64_t m, double* const* di, double* const* dout, cudaStream_t st, int slot ) {
        launchSampleIso( s, di[0], m, nullptr, dout[0], dout[1], st, slot );
      } );
    }
    raiseDeviceErrors( fetchDeviceErrors( s, s->streams[0] ) );
  }

  // fused: cross section + sampled outcome per neutron in one pass over the host arrays (32 B per neutron over the bus
  // instead of the 40 B of the two separate calls)
  void xsAndSampleIsoHost( Scatter* s, const double* ekin, uint64_t n, double* xs, double* eout, double* mu )
  {
    if ( !n ) return;
    const auto devs = fanDevices( s, n );
    if ( !devs.empty() ) {
      fanOut( s, devs, n, true, [&]( Scatter* h, uint64_t b, uint64_t m ) { xsAndSampleIsoHost( h, ekin + b, m, xs + b, eout + b, mu + b ); } );
      return;
    }
    DeviceGuard dg( s->dm->device );
    const double* in[1] = { ekin };
    double* out[3] = { xs, eout, mu };
    runHostPipeline( s, n, 1, in, 3, out, [s]( uint64_t m, double* const* di, double* const* dout, cudaStream_t st, int slot ) {
      launchSampleIso( s, di[0], m, dout[0], dout[1], dout[2], st, slot );
    } );
    raiseDeviceErrors( fetchDeviceErrors( s, s->streams[0] ) );
  }

  void xsAnisoHost( Scatter* s, const double* ekin, const double* ux, const double* uy, const double* uz,
                    uint64_t n, double* results )
  {
    if ( !n ) return;
    const auto devs = fanDevices( s, n );
    if ( !devs.empty() ) {
      fanOut( s, devs, n, false, [&]( Scatter* h, uint64_t b, uint64_t m ) { xsAnisoHost( h, ekin + b, ux + b, uy + b, uz + b, m, results + b ); } );
      return;
    }
    DeviceGuard dg( s->dm->device );
    const double* in[4] = { ekin, ux, uy, uz };
    double* out[1] = { results };
    runHostPipeline( s, n, 4, in, 1, out, [s]( uint64_t m, double* const* di, double* const* dout, cudaStream_t st, int slot ) {
      launchXSAniso( s, di[0], di[1], di[2], di[3], m, dout[0], st, slot );
    } );
  }

  void sampleAnisoHost( Scatter* s, const double* ekin, const double* ux, const double* uy, const double* uz,
                        uint64_t n, double* eout, double* ox, double* oy, double* oz )
  {
    if ( !n ) return;
    const auto devs = fanDevices( s, n );
    if ( !devs.empty() ) {
      fanOut( s, devs, n, true, [&]( Scatter* h, uint64_t b, uint64_t m ) {
        sampleAnisoHost( h, ekin + b, ux + b, uy + b, uz + b, m, eout + b, ox + b, oy + b, oz + b ); } );
      return;
    }
    DeviceGuard dg( s->dm->device );
    const double* in[4] = { ekin, ux, uy, uz };
    double* out[4] = { eout, ox, oy, oz };
    runHostPipeline( s, n, 4, in, 4, out, [s]( uint64_t m, double* const* di, double* const* dout, cudaStream_t st, int slot ) {
      launchSampleAniso( s, di[0], di[1], di[2], di[3], m, dout[0], dout[1], dout[2], dout[3], st, slot );
    } );
    raiseDeviceErrors( fetchDeviceErrors( s, s->streams[0] ) );
  }

#include "ncb_lib_mmc.inc"

}

// =============================================================================
//                                  C ABI
// =============================================================================
extern "C" {

  // ---- error API (ref: ncrystal.cc:396-440)
  void ncrystal_seterrhandler( void (*handler)(char*,char*) ) { g_custom_error_handler = handler; }
  void ncrystal_setmsghandler( void (*handler)(const char*,unsigned) ) { std::lock_guard<std::mutex> g( g_msg_mtx ); g_msg_handler = handler; }
  void ncb200_emit_message( const char* msg, unsigned msgtype ) { if ( msg && msgtype <= 2 ) emitMsg( msg, msgtype ); }
  int ncrystal_error(void) { return g_waserror; }
  const char* ncrystal_lasterror(void) { return g_waserror ? g_errmsg : nullptr; }
  const char* ncrystal_lasterrortype(void) { return g_waserror ? g_errtype : nullptr; }
  void ncrystal_clearerror(void) { g_waserror = 0; }
  int ncrystal_setquietonerror( int q ) { int old = g_quietonerror; g_quietonerror = q; return old; }
  int ncrystal_sethaltonerror( int h ) { int old = g_haltonerror; g_haltonerror = h; return old; }

  // ---- handle management (ref: ncrystal.cc:442-545)
  int ncrystal_valid( void* object )
  {
    if ( !object ) return 0;
    return *reinterpret_cast<void**>( object ) ? 1 : 0;
  }
  int ncrystal_refcount( void* object )
  {
    try { return fromInternal( *reinterpret_cast<void**>( object ), "ncrystal_refcount" )->refcount.load(); } NCBCATCH;
    return -999;
  }
  void ncrystal_ref( void* object )
  {
    try { ++fromInternal( *reinterpret_cast<void**>( object ), "ncrystal_ref" )->refcount; } NCBCATCH;
  }
  void ncrystal_unref( void* object )
  {
    try {
      void*& internal = *reinterpret_cast<void**>( object );
      Scatter* s = fromInternal( internal, "ncrystal_unref" );
      if ( s->refcount.fetch_sub(1) == 1 ) {
        { DeviceGuard dg( s->dm->device ); mmcReleaseBuffers( s ); delete s; }
        internal = nullptr;
      }
    } NCBCATCH;
  }
  void ncrystal_invalidate( void* object )
  {
    if ( !ncrystal_valid( object ) ) return;
    *reinterpret_cast<void**>( object ) = nullptr;
  }
  ncrystal_process_t ncrystal_cast_scat2proc( ncrystal_scatter_t s )
  {
    try { fromInternal( s.internal, "ncrystal_cast_scat2proc" ); return { s.internal }; } NCBCATCH;
    return { nullptr };
  }
  ncrystal_scatter_t ncrystal_cast_proc2scat( ncrystal_process_t p )
  {
    if ( p.internal && static_cast<FingerPrint*>( p.internal )->tag == kAbsorptionTag ) return { nullptr };
    try { fromInternal( p.internal, "ncrystal_cast_proc2scat" ); return { p.internal }; } NCBCATCH;
    return { nullptr };
  }

  // ---- absorption: 1/v process from the compiled material's header (ref: ncrystal.h:703 ncrystal_create_absorption,
  // :681-684 casts; NCAbsOOV.cc).  Same handle machinery; only the cross-section entry points accept it.
  ncrystal_absorption_t ncb200_create_absorption_from_blob( const void* blob, size_t nbytes )
  {
    try {
      if ( nbytes < sizeof(ncb_header_t) ) throw Err( "BadInput", "compiled material: buffer too small" );
      ncb_header_t hdr; std::memcpy( &hdr, blob, sizeof(hdr) );
      if ( hdr.magic != NCB_MAGIC || hdr.version != NCB_VERSION ) throw Err( "BadInput", "compiled material: bad magic or version" );
      if ( hdr.nbytes > nbytes ) throw Err( "BadInput", "compiled material: inconsistent header" );
      if ( hdr.abs_c < 0.0 ) throw Err( "BadInput", "the material's absorption process is not of the 1/v type" );
      auto dm = std::make_shared<DeviceMaterial>();
      dm->uid = ++g_material_uid_counter;
      CUDA_OK( cudaGetDevice( &dm->device ) );
      std::memset( &dm->mat, 0, sizeof(dm->mat) );
      std::memset( &dm->sp, 0, sizeof(dm->sp) ); std::memset( &dm->sp_sc, 0, sizeof(dm->sp_sc) ); std::memset( &dm->sp_iso, 0, sizeof(dm->sp_iso) );
      Material& m = dm->mat;
      m.ncomp = 1; m.oriented = 0;
      m.dom_lo = 0.0; m.dom_hi = hdr.abs_c > 0.0 ? kInf : 0.0;      // AbsOOV::m_domain, NCAbsOOV.cc:35-37
      m.comp[0].kind = KIND_ABSOOV; m.comp[0].scale = 1.0; m.comp[0].par = hdr.abs_c;
      m.comp[0].dom_lo = m.dom_lo; m.comp[0].dom_hi = m.dom_hi;
      hdr.cfg[sizeof(hdr.cfg)-1] = 0;
      dm->cfg = hdr.cfg; dm->numdens = hdr.numdens; dm->abs_c = hdr.abs_c; dm->temperature = hdr.temperature;
      ncrystal_scatter_t h = newHandle( dm, 0, 0 );
      static_cast<FingerPrint*>( h.internal )->tag = kAbsorptionTag;
      return { h.internal };
    } NCBCATCH;
    return { nullptr };
  }
  ncrystal_absorption_t ncrystal_create_absorption( const char* cfgstr )
  {
    try {
      auto d = findCompiledMaterial( cfgstr );
      return ncb200_create_absorption_from_blob( d.data(), d.size() );
    } NCBCATCH;
    return { nullptr };
  }
  ncrystal_process_t ncrystal_cast_abs2proc( ncrystal_absorption_t a )
  {
    try { fromInternal( a.internal, "ncrystal_cast_abs2proc" ); return { a.internal }; } NCBCATCH;
    return { nullptr };
  }
  ncrystal_absorption_t ncrystal_cast_proc2abs( ncrystal_process_t p )
  {
    // like the reference: a null handle (no error) when the process is not an absorption process
    if ( p.internal && static_cast<FingerPrint*>( p.internal )->tag == kAbsorptionTag ) return { p.internal };
    return { nullptr };
  }

  // ref: ncrystal.cc:1518-1525 -- absorption handles carry no state: the clone shares the material
  ncrystal_absorption_t ncrystal_clone_absorption( ncrystal_absorption_t a )
  {
    try {
      Scatter* s = fromInternal( a.internal, "ncrystal_clone_absorption" );
      if ( s->fp.tag != kAbsorptionTag ) throw Err( "LogicError", "ncrystal_clone_absorption: not an absorption handle" );
      ncrystal_scatter_t h = newHandle( s->dm, 0, 0 );
      static_cast<FingerPrint*>( h.internal )->tag = kAbsorptionTag;
      return { h.internal };
    } NCBCATCH;
    return { nullptr };
  }
  // ref: ncrystal.cc:2287-2295 -- id of the underlying (shared, immutable) process: equal for a handle and its clones
  char* ncrystal_process_uid( ncrystal_process_t p )
  {
    try {
      Scatter* s = fromInternal( p.internal, "ncrystal_process_uid" );
      const std::string u = std::to_string( s->dm->uid );
      char* out = static_cast<char*>( std::malloc( u.size()+1 ) );
      std::strcpy( out, u.c_str() );
      return out;
    } NCBCATCH;
    return nullptr;
  }
  // ref: ncrystal.h:1229-1234 -- the NCrystal release whose hot path this library restates
  int ncrystal_version(void) { return 4004002; }
  const char* ncrystal_version_str(void) { return "4.4.2"; }
  const char* ncrystal_namespace(void) { return ""; }
  void ncrystal_dealloc_doubleptr( double* p ) { std::free( p ); }
  // ref: ncrystal.h:1387-1393 -- obsolete in the reference as well ("Calling it will result in an error")
  void ncrystal_runmmcsim_stdengine( unsigned, unsigned, const char*, const char*, const char*, char**, unsigned*, double**, double** )
  {
    try {
      throw Err( "LogicError", "The ncrystal_runmmcsim_stdengine function is obsolete; use ncb200_minimc_run "
                 "(the [\"mmc\",\"run\",...] query of ncrystal_jsonquery)." );
    } NCBCATCH;
  }

  // ---- creation
  ncrystal_scatter_t ncb200_create_scatter_from_blob( const void* blob, size_t nbytes, unsigned long seed )
  {
    try { return newHandle( uploadMaterial( blob, nbytes ), seed, 0 ); } NCBCATCH;
    return { nullptr };
  }
  ncrystal_scatter_t ncb200_create_scatter_from_file( const char* path, unsigned long seed )
  {
    try {
      auto d = readFile( path );
      if ( d.empty() ) throw Err( "FileNotFound", std::string("Could not read compiled material file ")+path );
      return newHandle( uploadMaterial( d.data(), d.size() ), seed, 0 );
    } NCBCATCH;
    return { nullptr };
  }
  void ncb200_set_data_path( const char* path )
  {
    std::lock_guard<std::mutex> g( g_path_mtx );
    g_data_path = path ? path : "";
  }
  int ncb200_cfg_to_filestem( const char* cfgstr, char* buf, int buflen )
  {
    const std::string s = cfgToStem( cfgstr );
    if ( buf && buflen > 0 ) std::snprintf( buf, (size_t)buflen, "%s", s.c_str() );
    return (int)s.size();
  }
  ncrystal_scatter_t ncrystal_create_scatter( const char* cfgstr )
  {
    try {
      auto d = findCompiledMaterial( cfgstr );
      // every handle created without explicit seed gets its own stream, like the reference's
      // default RNG producer (NCFact.cc:28-35)
      return newHandle( uploadMaterial( d.data(), d.size() ), g_default_seed.load(), 0x10000000u + g_default_stream_counter++ );
    } NCBCATCH;
    return { nullptr };
  }
  ncrystal_scatter_t ncrystal_create_scatter_builtinrng( const char* cfgstr, unsigned long seed )
  {
    try {
      auto d = findCompiledMaterial( cfgstr );
      return newHandle( uploadMaterial( d.data(), d.size() ), seed, 0 );
    } NCBCATCH;
    return { nullptr };
  }
  ncrystal_scatter_t ncrystal_clone_scatter( ncrystal_scatter_t o )
  {
    try {
      Scatter* s = fromInternal( o.internal, "ncrystal_clone_scatter" );
      return newHandle( s->dm, s->seed, 0x20000000u + ( ++s->dm->clone_counter ) );
    } NCBCATCH;
    return { nullptr };
  }
  ncrystal_scatter_t ncrystal_clone_scatter_rngbyidx( ncrystal_scatter_t o, unsigned long idx )
  {
    try {
      Scatter* s = fromInternal( o.internal, "ncrystal_clone_scatter_rngbyidx" );
      return newHandle( s->dm, s->seed, 0x40000000u + (uint32_t)( idx & 0x3fffffffu ) );
    } NCBCATCH;
    return { nullptr };
  }
  ncrystal_scatter_t ncrystal_clone_scatter_rngforcurrentthread( ncrystal_scatter_t o )
  {
    try {
      Scatter* s = fromInternal( o.internal, "ncrystal_clone_scatter_rngforcurrentthread" );
      const size_t h = std::hash<std::thread::id>()( std::this_thread::get_id() );
      return newHandle( s->dm, s->seed, 0x80000000u + (uint32_t)( h & 0x7fffffffu ) );
    } NCBCATCH;
    return { nullptr };
  }

  // ---- queries (ref: ncrystal.cc:1040-1087)
  const char* ncrystal_name( ncrystal_process_t p )
  {
    try {
      Scatter* s = fromInternal( p.internal, "ncrystal_name" );
      if ( s->dm->mat.ncomp > 1 ) return "ProcComposition";
      switch ( s->dm->mat.comp[0].kind ) {
      case KIND_POWDERBRAGG: return "PowderBragg";
      case KIND_ELINC: return "ElIncScatter";
      case KIND_SAB: return "SABScatter";
      case KIND_FREEGAS: return "FreeGas";
      case KIND_SCBRAGG: return "SCBragg";
      case KIND_LCBRAGG: return "LCBragg";
      case KIND_ABSOOV: return "AbsOOV";
      default: return "Process";
      }
    } NCBCATCH;
    return nullptr;
  }
  int ncrystal_isnonoriented( ncrystal_process_t p )
  {
    try { return fromInternal( p.internal, "ncrystal_isnonoriented" )->dm->mat.oriented ? 0 : 1; } NCBCATCH;
    return -1;
  }
  void ncrystal_domain( ncrystal_process_t p, double* ekin_low, double* ekin_high )
  {
    try {
      Scatter* s = fromInternal( p.internal, "ncrystal_domain" );
      *ekin_low = s->dm->mat.dom_lo;
      *ekin_high = s->dm->mat.dom_hi;
      return;
    } NCBCATCH;
    *ekin_low = *ekin_high = -1.0;
  }

  // ---- cross sections
  void ncrystal_crosssection_nonoriented_many( ncrystal_process_t o, const double* ekin, unsigned long n_ekin,
                                               unsigned long repeat, double* results )
  {
    try {
      xsIsoHost( fromInternal( o.internal, "ncrystal_crosssection_nonoriented_many" ), ekin, n_ekin, repeat, results );
      return;
    } NCBCATCH;
    for ( unsigned long i = 0; i < n_ekin*repeat; ++i ) results[i] = -1.0;
  }
  void ncrystal_crosssection_nonoriented( ncrystal_process_t o, double ekin, double* result )
  {
    try {
      xsIsoHost( fromInternal( o.internal, "ncrystal_crosssection_nonoriented" ), &ekin, 1, 1, result );
      return;
    } NCBCATCH;
    *result = -1.0;
  }

  // ---- sampling
  void ncrystal_samplescatterisotropic_many( ncrystal_scatter_t o, const double* ekin, unsigned long n_ekin,
                                             unsigned long repeat, double* results_ekin, double* results_cos_scat_angle )
  {
    try {
      sampleIsoHost( fromInternal( o.internal, "ncrystal_samplescatterisotropic_many" ), ekin, n_ekin, repeat,
                     results_ekin, results_cos_scat_angle );
      return;
    } NCBCATCH;
    for ( unsigned long i = 0; i < n_ekin*repeat; ++i ) { results_ekin[i] = -1.0; results_cos_scat_angle[i] = -999.0; }
  }
  // host-pointer variant of ncb200_xs_and_samplescatterisotropic_many_dev (the reference's fused batch entry
  // evalXSAndSampleScatterIsotropic, NCABIUtils.hh:78-100): results as the two separate *_many calls give them
  void ncb200_xs_and_samplescatterisotropic_many( ncrystal_scatter_t o, const double* ekin, uint64_t n,
                                                  double* results_xs, double* results_ekin, double* results_cos_scat_angle )
  {
    try {
      xsAndSampleIsoHost( fromInternal( o.internal, "ncb200_xs_and_samplescatterisotropic_many" ), ekin, n,
                          results_xs, results_ekin, results_cos_scat_angle );
      return;
    } NCBCATCH;
    for ( uint64_t i = 0; i < n; ++i ) { results_xs[i] = -1.0; results_ekin[i] = -1.0; results_cos_scat_angle[i] = -999.0; }
  }
  void ncrystal_samplescatterisotropic( ncrystal_scatter_t o, double ekin, double* ekin_final, double* cos_scat_angle )
  {
    try {
      sampleIsoHost( fromInternal( o.internal, "ncrystal_samplescatterisotropic" ), &ekin, 1, 1, ekin_final, cos_scat_angle );
      return;
    } NCBCATCH;
    *ekin_final = -1.0;
    *cos_scat_angle = -999;
  }

  // ---- device-pointer variants
  void ncb200_crosssection_nonoriented_many_dev( ncrystal_process_t o, const double* d_ekin, uint64_t n,
                                                 double* d_results, void* stream )
  {
    try {
      Scatter* s = fromInternal( o.internal, "ncb200_crosssection_nonoriented_many_dev" );
      DeviceGuard dg( s->dm->device );
      launchXSIso( s, d_ekin, n, d_results, static_cast<cudaStream_t>( stream ) );
    } NCBCATCH;
  }
  void ncb200_samplescatterisotropic_many_dev( ncrystal_scatter_t o, const double* d_ekin, uint64_t n,
                                               double* d_ekin_final, double* d_mu, void* stream )
  {
    try {
      Scatter* s = fromInternal( o.internal, "ncb200_samplescatterisotropic_many_dev" );
      DeviceGuard dg( s->dm->device );
      launchSampleIso( s, d_ekin, n, nullptr, d_ekin_final, d_mu, static_cast<cudaStream_t>( stream ) );
    } NCBCATCH;
  }
  void ncb200_xs_and_samplescatterisotropic_many_dev( ncrystal_scatter_t o, const double* d_ekin, uint64_t n,
                                                      double* d_xs, double* d_ekin_final, double* d_mu, void* stream )
  {
    try {
      Scatter* s = fromInternal( o.internal, "ncb200_xs_and_samplescatterisotropic_many_dev" );
      DeviceGuard dg( s->dm->device );
      launchSampleIso( s, d_ekin, n, d_xs, d_ekin_final, d_mu, static_cast<cudaStream_t>( stream ) );
    } NCBCATCH;
  }
  int ncb200_check_device_errors( ncrystal_scatter_t o, void* stream )
  {
    int flags = 0;
    try {
      Scatter* s = fromInternal( o.internal, "ncb200_check_device_errors" );
      DeviceGuard dg( s->dm->device );
      flags = fetchDeviceErrors( s, static_cast<cudaStream_t>( stream ) );
      raiseDeviceErrors( flags );
    } NCBCATCH;
    return flags;
  }
  void ncb200_set_diagnostics_dev( ncrystal_scatter_t o, uint32_t* d_ndraws, int32_t* d_component )
  {
    try {
      Scatter* s = fromInternal( o.internal, "ncb200_set_diagnostics_dev" );
      s->d_diag_ndraws = d_ndraws;
      s->d_diag_comp = d_component;
    } NCBCATCH;
  }

  // ---- RNG control
  void ncb200_set_rng_stream( ncrystal_scatter_t o, uint64_t seed, uint32_t stream_id, uint64_t next_index )
  {
    try {
      Scatter* s = fromInternal( o.internal, "ncb200_set_rng_stream" );
      s->seed = seed; s->sid = stream_id; s->next_index = next_index;
    } NCBCATCH;
  }
  void ncb200_set_fg_staged_min( uint64_t nmin ) { g_fg_staged_min.store( nmin ); }
  void ncb200_set_mmc_tail_mode( int persistent ) { g_mmc_persistent_tail.store( persistent != 0 ); if ( persistent > 0 ) g_mmc_tail_ctas.store( persistent >= 2 ? 2 : 1 ); }
  void ncb200_get_rng_stream( ncrystal_scatter_t o, uint64_t* seed, uint32_t* stream_id, uint64_t* next_index )
  {
    try {
      Scatter* s = fromInternal( o.internal, "ncb200_get_rng_stream" );
      if ( seed ) *seed = s->seed;
      if ( stream_id ) *stream_id = s->sid;
      if ( next_index ) *next_index = s->next_index;
    } NCBCATCH;
  }
  void ncrystal_setrandgen( double (*)(void) )
  {
    try {
      throw Err( "LogicError", "ncrystal_setrandgen: host RNG callbacks cannot be evaluated on the GPU; "
                 "ncrystal_b200 uses per-neutron counter-based streams (see ncb200_set_rng_stream)" );
    } NCBCATCH;
  }
  // ref: ncrystal.cc:1996-2025.  The "default generator" is the seed (and stream numbering) that scatter handles
  // created afterwards by ncrystal_create_scatter start from: re-seeding makes their outcomes reproducible.
  void ncrystal_setbuiltinrandgen(void) { g_default_seed.store( kDefaultSeed ); g_default_stream_counter.store( 0 ); }
  void ncrystal_setbuiltinrandgen_withseed( unsigned long seed ) { g_default_seed.store( seed ); g_default_stream_counter.store( 0 ); }
  void ncrystal_setbuiltinrandgen_withstate( const char* st )
  {
    try {
      unsigned long long seed, idx; unsigned sid;
      if ( !st || std::strlen(st) != 48 || std::strcmp( st+40, "b2005eed" ) != 0
           || std::sscanf( st, "%16llx%8x%16llx", &seed, &sid, &idx ) != 3 )
        throw Err( "BadInput", std::string("ncrystal_setbuiltinrandgen_withstate got state which is not from this library's RNG: ")
                   + ( st ? st : "<null>" ) );
      g_default_seed.store( seed ); g_default_stream_counter.store( 0 );
    } NCBCATCH;
  }
  int ncrystal_rngsupportsstatemanip_ofscatter( ncrystal_scatter_t ) { return 1; }
  char* ncrystal_getrngstate_ofscatter( ncrystal_scatter_t o )
  {
    try {
      Scatter* s = fromInternal( o.internal, "ncrystal_getrngstate_ofscatter" );
      char buf[96];
      // 64-bit seed, 32-bit stream id, 64-bit next index, then a 4-byte type id (cf. the
      // reference's hex state + type UID, NCRNG.cc:89,165-186)
      std::snprintf( buf, sizeof(buf), "%016llx%08x%016llxb2005eed", (unsigned long long)s->seed, s->sid,
                     (unsigned long long)s->next_index );
      char* out = static_cast<char*>( std::malloc( std::strlen(buf)+1 ) );
      std::strcpy( out, buf );
      return out;
    } NCBCATCH;
    return nullptr;
  }
  void ncrystal_setrngstate_ofscatter( ncrystal_scatter_t o, const char* st )
  {
    try {
      Scatter* s = fromInternal( o.internal, "ncrystal_setrngstate_ofscatter" );
      unsigned long long seed, idx; unsigned sid;
      if ( !st || std::strlen(st) != 48 || std::strcmp( st+40, "b2005eed" ) != 0
           || std::sscanf( st, "%16llx%8x%16llx", &seed, &sid, &idx ) != 3 )
        throw Err( "BadInput", "ncrystal_setrngstate_ofscatter: not a state string of this RNG type" );
      s->seed = seed; s->sid = sid; s->next_index = idx;
    } NCBCATCH;
  }
  void ncrystal_dealloc_string( char* p ) { std::free( p ); }

  // ---- source + tally
  void ncb200_generate_source_dev( uint64_t seed, uint64_t first_index, uint64_t n, double lo, double hi,
                                   double* d_ekin, double* d_ux, double* d_uy, double* d_uz, void* stream )
  {
    try {
      if ( !n ) return;
      int dev = 0; CUDA_OK( cudaGetDevice( &dev ) );
      const double loglo = std::log10( lo ), logspan = std::log10( hi ) - std::log10( lo );
      k_gen_source<<< gridFor( n, 256, dev, 8 ), 256, 0, static_cast<cudaStream_t>( stream ) >>>(
        seed, first_index, n, loglo, logspan, d_ekin, d_ux, d_uy, d_uz );
      ++g_launches;
      CUDA_OK( cudaGetLastError() );
    } NCBCATCH;
  }
  void ncb200_tally_hist_dev( const double* d_values, const double* d_weights, uint64_t n,
                              double lo, double hi, uint32_t nbins, double* d_hist, double* d_sumw2, void* stream )
  {
    try {
      if ( !n ) return;
      if ( !( hi > lo ) || nbins == 0 || nbins > 12000 )
        throw Err( "BadInput", "ncb200_tally_hist_dev: need hi>lo and 1<=nbins<=12000" );
      int dev = 0; CUDA_OK( cudaGetDevice( &dev ) );
      const uint32_t smem = ( nbins + 2 ) * 8u * ( d_sumw2 ? 2u : 1u );
      ensureKernelAttrs( dev );
      k_tally_hist<<< gridFor( n, 256, dev, 4 ), 256, smem, static_cast<cudaStream_t>( stream ) >>>(
        d_values, d_weights, n, lo, (double)nbins/( hi - lo ), nbins, d_hist, d_sumw2 );
      ++g_launches;
      CUDA_OK( cudaGetLastError() );
    } NCBCATCH;
  }

  // ---- page-locking of caller buffers: a caller that reuses its arrays pins them once (about 20 ms per 80 MB) and
  // every later *_many call on them runs at the pinned-buffer rate instead of through the bounce ring
  int ncb200_pin_host_buffer( void* p, uint64_t nbytes )
  {
    try { CUDA_OK( cudaHostRegister( p, nbytes, cudaHostRegisterPortable ) ); return 0; } NCBCATCH;
    return -1;
  }
  int ncb200_unpin_host_buffer( void* p )
  {
    try { CUDA_OK( cudaHostUnregister( p ) ); return 0; } NCBCATCH;
    return -1;
  }

  // ---- several devices from one process (ncb_lib_multigpu.inc)
  int ncb200_set_devices( int n )
  {
    try {
      int ndev = 0;
      CUDA_OK( cudaGetDeviceCount( &ndev ) );
      if ( n < 0 || n > ndev ) throw Err( "BadInput", "ncb200_set_devices: "+std::to_string(n)+" devices requested, "+std::to_string(ndev)+" visible" );
      if ( n == 0 ) n = ndev;
      std::lock_guard<std::mutex> g( g_dev_mtx );
      g_devices.clear();
      if ( n > 1 ) for ( int d = 0; d < n; ++d ) g_devices.push_back( d );
      return n;
    } NCBCATCH;
    return -1;
  }
  int ncb200_get_devices(void) { std::lock_guard<std::mutex> g( g_dev_mtx ); return g_devices.empty() ? 1 : (int)g_devices.size(); }
  void ncb200_set_fanout_min( uint64_t n ) { std::lock_guard<std::mutex> g( g_dev_mtx ); g_fanout_min = n ? n : 1; }
  void ncb200_tally_hist_many( const double* values, const double* weights, uint64_t n, double lo, double hi, uint32_t nbins,
                               double* hist, double* sumw2 )
  {
    try { if ( n ) tallyHistHost( values, weights, n, lo, hi, nbins, hist, sumw2 ); } NCBCATCH;
  }

  // ---- introspection
  int ncb200_ncomponents( ncrystal_process_t p )
  {
    try { return fromInternal( p.internal, "ncb200_ncomponents" )->dm->mat.ncomp; } NCBCATCH;
    return -1;
  }
  int ncb200_component_kind( ncrystal_process_t p, int i )
  {
    try {
      Scatter* s = fromInternal( p.internal, "ncb200_component_kind" );
      if ( i < 0 || i >= s->dm->mat.ncomp ) throw Err( "BadInput", "component index out of range" );
      return s->dm->mat.comp[i].kind;
    } NCBCATCH;
    return -1;
  }
  double ncb200_component_scale( ncrystal_process_t p, int i )
  {
    try {
      Scatter* s = fromInternal( p.internal, "ncb200_component_scale" );
      if ( i < 0 || i >= s->dm->mat.ncomp ) throw Err( "BadInput", "component index out of range" );
      return s->dm->mat.comp[i].scale;
    } NCBCATCH;
    return -1.0;
  }
  // ref: ncrystal.cc:1051-1059 (NCDefs.hh:834-845)
  double ncrystal_wl2ekin( double wl ) { const double w2 = wl*wl; return w2 ? kWl2Ekin / w2 : kInf; }
  double ncrystal_ekin2wl( double ekin ) { return ekin ? std::sqrt( kWl2Ekin / ekin ) : kInf; }

  uint64_t ncb200_kernel_launch_count(void) { return g_launches.load(); }

  void ncb200_kernel_timing( int enable )
  {
    std::lock_guard<std::mutex> g( g_ktimer.mtx );
    for ( auto& r : g_ktimer.recs ) { cudaEventDestroy( r.a ); cudaEventDestroy( r.b ); }
    g_ktimer.recs.clear();
    g_ktimer.enabled = enable != 0;
  }
  int ncb200_kernel_timing_report( char* buf, int buflen )
  {
    try {
      CUDA_OK( cudaDeviceSynchronize() );
      std::lock_guard<std::mutex> g( g_ktimer.mtx );
      struct Acc { double ms = 0; int n = 0; };
      std::vector<std::pair<std::string,Acc>> acc;
      for ( auto& r : g_ktimer.recs ) {
        float ms = 0.f;
        if ( cudaEventElapsedTime( &ms, r.a, r.b ) != cudaSuccess ) continue;
        auto it = acc.begin();
        for ( ; it != acc.end(); ++it ) if ( it->first == r.name ) break;
        if ( it == acc.end() ) { acc.emplace_back( r.name, Acc() ); it = acc.end() - 1; }
        it->second.ms += ms; it->second.n += 1;
      }
      std::ostringstream ss;
      ss << "{";
      bool first = true;
      for ( auto& e : acc ) {
        ss << ( first ? "" : ", " ) << "\"" << e.first << "\": {\"launches\": " << e.second.n << ", \"ms_avg\": "
           << ( e.second.n ? e.second.ms/e.second.n : 0.0 ) << "}";
        first = false;
      }
      ss << "}";
      const std::string out = ss.str();
      if ( buf && buflen > 0 ) std::snprintf( buf, (size_t)buflen, "%s", out.c_str() );
      return (int)out.size();
    } NCBCATCH;
    return -1;
  }
  // queue sizes of the most recent isotropic sampling launch: [table path, free-gas path, table-at-Emax]
  int ncb200_last_queue_counts( ncrystal_scatter_t o, uint32_t* out3 )
  {
    try {
      Scatter* s = fromInternal( o.internal, "ncb200_last_queue_counts" );
      if ( !s->last_counts_ptr ) return -1;
      DeviceGuard dg( s->dm->device );
      CUDA_OK( cudaDeviceSynchronize() );
      CUDA_OK( cudaMemcpy( out3, s->last_counts_ptr, 3*sizeof(uint32_t), cudaMemcpyDeviceToHost ) );
      return 0;
    } NCBCATCH;
    return -1;
  }
  uint64_t ncb200_table_bytes( ncrystal_process_t p )
  {
    try { return fromInternal( p.internal, "ncb200_table_bytes" )->dm->arena_bytes; } NCBCATCH;
    return 0;
  }
  // measured vector-FP64 FMA rate of the current device [TFLOP/s] (best of 5 launches, CUDA events)
  double ncb200_fp64_fma_probe(void)
  {
    try {
      int dev = 0; CUDA_OK( cudaGetDevice( &dev ) );
      double* d_out = nullptr; CUDA_OK( cudaMalloc( &d_out, sizeof(double) ) );
      cudaEvent_t e0, e1; CUDA_OK( cudaEventCreate( &e0 ) ); CUDA_OK( cudaEventCreate( &e1 ) );
      const int iters = 1 << 15, threads = 256;
      const unsigned grid = (unsigned)numSMs( dev ) * 8;
      double best = 0.0;
      for ( int rep = 0; rep < 6; ++rep ) {
        CUDA_OK( cudaEventRecord( e0, nullptr ) );
        k_fp64_fma_probe<<< grid, threads >>>( d_out, iters, 0.999999, 1e-9 );
        CUDA_OK( cudaEventRecord( e1, nullptr ) );
        CUDA_OK( cudaEventSynchronize( e1 ) );
        float ms = 0.f; CUDA_OK( cudaEventElapsedTime( &ms, e0, e1 ) );
        const double tflops = 2.0*8.0*(double)iters*(double)threads*(double)grid / ( (double)ms*1e-3 ) / 1e12;
        if ( rep && tflops > best ) best = tflops;     // first launch: warm-up
      }
      ++g_launches;
      cudaEventDestroy( e0 ); cudaEventDestroy( e1 ); cudaFree( d_out );
      return best;
    } NCBCATCH;
    return -1.0;
  }
  const char* ncb200_version(void) { return "ncrystal_b200 0.1 (sm_100a; hot path of NCrystal 4.4.2)"; }

  int ncb200_sab_xscheck( ncrystal_process_t p, int component, double* out, int nmax )
  {
    try {
      Scatter* s = fromInternal( p.internal, "ncb200_sab_xscheck" );
      const Material& M = s->dm->mat;
      if ( component < 0 || component >= M.ncomp || M.comp[component].kind != KIND_SAB ) return -1;
      const int isab = M.comp[component].idx;
      for ( auto& pl : s->dm->sabplans )
        if ( pl.sab_index == isab ) {
          const int ne = M.sab[isab].negrid;
          const int m = ne < nmax ? ne : nmax;
          DeviceGuard dg( s->dm->device );
          CUDA_OK( cudaMemcpy( out, static_cast<unsigned char*>( s->dm->d_arena ) + pl.off_xscheck, (size_t)m*8, cudaMemcpyDeviceToHost ) );
          return ne;
        }
    } NCBCATCH;
    return -1;
  }

  // energy grid, grid cross sections and extension constants {k_extension, k1, k2, egrid_margin} of a S(alpha,beta) leaf
  // (as delivered in the compiled material, or as determined by the library: ncb_sab_t::auto_egrid)
  int ncb200_sab_energy_grid( ncrystal_process_t p, int component, double* egrid, double* xs, int nmax, double* consts4 )
  {
    try {
      Scatter* s = fromInternal( p.internal, "ncb200_sab_energy_grid" );
      const Material& M = s->dm->mat;
      if ( component < 0 || component >= M.ncomp || M.comp[component].kind != KIND_SAB ) return -1;
      const SabT& T = M.sab[M.comp[component].idx];
      const int m = T.negrid < nmax ? T.negrid : nmax;
      DeviceGuard dg( s->dm->device );
      if ( egrid ) CUDA_OK( cudaMemcpy( egrid, T.egrid, (size_t)m*8, cudaMemcpyDeviceToHost ) );
      if ( xs ) CUDA_OK( cudaMemcpy( xs, T.xs, (size_t)m*8, cudaMemcpyDeviceToHost ) );
      if ( consts4 ) { consts4[0] = T.k_extension; consts4[1] = T.k1; consts4[2] = T.k2; consts4[3] = T.egrid_margin; }
      return T.negrid;
    } NCBCATCH;
    return -1;
  }

  int ncb200_sab_sampler_dump( ncrystal_process_t p, int component, int iE, double* x, double* pdf, double* cdf,
                               double* infos, double* meta )
  {
    try {
      Scatter* s = fromInternal( p.internal, "ncb200_sab_sampler_dump" );
      const Material& M = s->dm->mat;
      if ( component < 0 || component >= M.ncomp || M.comp[component].kind != KIND_SAB ) return -1;
      const SabT& T = M.sab[M.comp[component].idx];
      if ( iE < 0 || iE >= T.negrid ) return -1;
      DeviceGuard dg( s->dm->device );
      SabEPoint ep;
      CUDA_OK( cudaMemcpy( &ep, T.ep + iE, sizeof(ep), cudaMemcpyDeviceToHost ) );
      const int n = ep.npts;
      if ( n <= 0 ) return 0;
      if ( x ) CUDA_OK( cudaMemcpy( x, T.bx + ep.off_b, (size_t)n*8, cudaMemcpyDeviceToHost ) );
      if ( pdf ) CUDA_OK( cudaMemcpy( pdf, T.bpdf + ep.off_b, (size_t)n*8, cudaMemcpyDeviceToHost ) );
      if ( cdf ) CUDA_OK( cudaMemcpy( cdf, T.bcdf + ep.off_b, (size_t)n*8, cudaMemcpyDeviceToHost ) );
      if ( infos ) {
        std::vector<SabAlphaInfo> v( n-1 );
        CUDA_OK( cudaMemcpy( v.data(), T.ainfo + ep.off_i, (size_t)(n-1)*sizeof(SabAlphaInfo), cudaMemcpyDeviceToHost ) );
        for ( int i = 0; i+1 < n; ++i ) {
          const SabAlphaInfo& f = v[i];
          double* o = infos + 10*i;
          o[0]=f.f_alpha; o[1]=f.f_sval; o[2]=f.f_logsval; o[3]=f.f_idx;
          o[4]=f.b_alpha; o[5]=f.b_sval; o[6]=f.b_logsval; o[7]=f.b_idx;
          o[8]=f.prob_front; o[9]=f.prob_notback;
        }
      }
      if ( meta ) { meta[0] = ep.ibeta_off; meta[1] = ep.first_bin_endpoint; }
      return n;
    } NCBCATCH;
    return -1;
  }

  // Consistency of the gather-friendly table copies and guides with the sampler tables they were derived from, all
  // read back from the device: heads / tails against AlphaSampleInfo + cumulative rows, points against the alpha /
  // S / log S / cumulative arrays, beta points against the (x,pdf,cdf) rows, and for every row and key of the log
  // guide the bracketing property the sampler relies on.  Returns the number of violations (0 = consistent), -1 on error.
  long ncb200_sab_selfcheck( ncrystal_process_t p, int component )
  {
    try {
      Scatter* s = fromInternal( p.internal, "ncb200_sab_selfcheck" );
      const Material& M = s->dm->mat;
      if ( component < 0 || component >= M.ncomp || M.comp[component].kind != KIND_SAB ) return -1;
      const SabT& T = M.sab[M.comp[component].idx];
      DeviceGuard dg( s->dm->device );
      CUDA_OK( cudaDeviceSynchronize() );
      const size_t ne = T.negrid, na = T.nalpha, nb = T.nbeta, bst = T.bstride;
      auto fetch = [&]( auto* dst, const void* src, size_t count ) {
        CUDA_OK( cudaMemcpy( dst, src, count*sizeof(*dst), cudaMemcpyDeviceToHost ) );
      };
      std::vector<double> alpha( na ), sab( na*nb ), logsab( na*nb ), cumul( na*nb ), bx( ne*bst ), bpdf( ne*bst ), bcdf( ne*bst );
      std::vector<SabAlphaInfo> ainfo( ne*nb );
      std::vector<SabHead> heads( ne*nb );
      std::vector<SabTail> tails( 2*ne*nb );
      std::vector<SabPoint> pts( na*nb );
      std::vector<SabBPoint> bpts( ne*bst );
      std::vector<uint16_t> lg( nb*(size_t)kSabGLStride );
      std::vector<SabEPoint> ep( ne );
      fetch( alpha.data(), T.alpha, na ); fetch( sab.data(), T.sab, na*nb ); fetch( logsab.data(), T.logsab, na*nb );
      fetch( cumul.data(), T.cumul, na*nb ); fetch( bx.data(), T.bx, ne*bst ); fetch( bpdf.data(), T.bpdf, ne*bst );
      fetch( bcdf.data(), T.bcdf, ne*bst ); fetch( ainfo.data(), T.ainfo, ne*nb ); fetch( heads.data(), T.heads, ne*nb );
      fetch( tails.data(), T.tails, 2*ne*nb ); fetch( pts.data(), T.pts, na*nb ); fetch( bpts.data(), T.bpts, ne*bst );
      fetch( lg.data(), T.lguide, nb*(size_t)kSabGLStride ); fetch( ep.data(), T.ep, ne );
      auto same = []( double a, double b ) { return std::memcmp( &a, &b, sizeof(double) ) == 0; };
      long bad = 0;
      for ( size_t k = 0; k < ne*nb; ++k ) {
        const SabAlphaInfo& f = ainfo[k]; const SabHead& h = heads[k];
        const double* row = cumul.data() + ( k % nb )*na;
        bad += !( same( h.prob_front, f.prob_front ) && same( h.prob_notback, f.prob_notback ) && (int)h.f_idx == f.f_idx
                  && (int)h.b_idx == f.b_idx && same( h.clow, row[f.f_idx] ) && same( h.cupp, row[f.b_idx] ) );
        bad += !( same( tails[2*k].alpha, f.f_alpha ) && same( tails[2*k].sval, f.f_sval ) && same( tails[2*k].logsval, f.f_logsval )
                  && same( tails[2*k+1].alpha, f.b_alpha ) && same( tails[2*k+1].sval, f.b_sval ) && same( tails[2*k+1].logsval, f.b_logsval ) );
      }
      for ( size_t k = 0; k < na*nb; ++k )
        bad += !( same( pts[k].alpha, alpha[k % na] ) && same( pts[k].sab, sab[k] ) && same( pts[k].logsab, logsab[k] ) && same( pts[k].cumul, cumul[k] ) );
      for ( size_t ie = 0; ie < ne; ++ie )
        for ( int j = 0; j < ep[ie].npts; ++j ) {
          const size_t k = ep[ie].off_b + j;
          bad += !( same( bpts[k].x, bx[k] ) && same( bpts[k].pdf, bpdf[k] ) && same( bpts[k].cdf, bcdf[k] ) );
        }
      for ( size_t ib = 0; ib < nb; ++ib ) {
        const double* row = cumul.data() + ib*na;
        const uint16_t* g = lg.data() + ib*(size_t)kSabGLStride;
        const double inv = heads[ib].inv_total;
        bad += !( g[0] == 0 && g[kSabGL] == (uint16_t)na );
        for ( int key = 0; key < kSabGL; ++key ) {
          bad += !( g[key] <= g[key+1] );
          // every point in [g[key], g[key+1]) has exactly this key; the points before / after have a smaller / larger one
          for ( int i = g[key]; i < g[key+1]; ++i ) bad += ( sabLogKey( row[i]*inv ) != key );
        }
      }
      return bad;
    } NCBCATCH;
    return -1;
  }

  // ---- oriented entry points: see ncb_lib_oriented.inc
#include "ncb_lib_oriented.inc"

  // ---- device-resident transport step: see ncb_lib_mmc.inc
#define NCB_MMC_CAPI
#include "ncb_lib_mmc.inc"
#undef NCB_MMC_CAPI

  // ---- VDOS -> S(alpha,beta) expansion on the device (ncrystal_raw_vdos2kernel / ncrystal_raw_vdos2gn)
#define NCB_LIB_VDOS_ENTRYPOINTS
#include "ncb_lib_vdos.inc"
#undef NCB_LIB_VDOS_ENTRYPOINTS

}

// ---- per-neutron boundaries with a caller-supplied generator (OpenMC virtual API, ncrystal_samplescatter_rs)
#include "ncb_lib_virtapi.inc"
