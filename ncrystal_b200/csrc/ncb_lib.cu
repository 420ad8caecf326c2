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
#include "ncb_kernels_cls.cuh"
#include "ncb_kernels_sc.cuh"
#include "ncb_kernels_mmc.cuh"
#include "ncb_loader.h"
#include "ncb_loader_sc.h"
#include "../../include/ncrystal_b200.h"

#include <atomic>
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
  void (*g_custom_error_handler)(char*,char*) = nullptr;

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
    // class-staged S(alpha,beta) sampling (ncb_kernels_cls.cuh): shared-memory plan, classes per leaf
    ClassSmem cls_smem = {};
    uint32_t cls_base[kMaxSab+1] = {};
    uint32_t ncls = 0;        // 0: the class path is not available for this material
    int cls_ctas = 0;         // resident CTAs per SM of k_sab_classes
    uint32_t sc_famof_off = 0, sc_scratch_off = 0, sc_smem = 0, sc_find_smem = 0, sc_find_famof_off = 0, sc_find_scratch_off = 0;
    std::string cfg;
    double numdens = 0.0, abs_c = 0.0, temperature = -1.0;
    std::vector<SabBuildPlan> sabplans;
    std::atomic<uint32_t> clone_counter{0};
    uint64_t uid = 0;        // ncrystal_process_uid
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
    // Budget: keep >= 2 CTAs/SM worth of shared memory (227 KB per SM usable).
    const uint32_t budget = 100u*1024u;
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
    // small SAB grids first, then the (possibly large) Bragg tables
    for ( int i = 0; i < nsab; ++i ) {
      add( 2*kMaxPB + i, M.sab[i].egrid, M.sab[i].negrid );
      add( 2*kMaxPB + kMaxSab + i, M.sab[i].xs, M.sab[i].negrid );
    }
    for ( int i = 0; i < npb; ++i ) {
      // both or none, so a lookup never mixes memory spaces
      const uint32_t need = 2u*( ( (uint32_t)M.pb[i].n*8u + 127u ) & ~127u );
      if ( off + need <= budget ) {
        add( i, M.pb[i].e2d, M.pb[i].n );
        add( kMaxPB + i, M.pb[i].fdm, M.pb[i].n );
      }
    }
    if ( M.sc.nfam ) {
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
    // class-staged sampling plan
    {
      int nbmax = 0, bsmax = 0;
      uint32_t nc = 0;
      for ( int i = 0; i < nsab; ++i ) {
        dm.cls_base[i] = nc;
        nc += (uint32_t)M.sab[i].negrid;
        nbmax = std::max( nbmax, M.sab[i].nbeta );
        bsmax = std::max( bsmax, M.sab[i].bstride );
      }
      for ( int i = nsab; i <= kMaxSab; ++i ) dm.cls_base[i] = nc;
      auto al = []( size_t b ) { return (uint32_t)( ( b + 127 ) & ~(size_t)127 ); };
      ClassSmem& S = dm.cls_smem;
      uint32_t o = 0;
      S.off_bx = o; o += al( (size_t)bsmax*8 );
      S.off_bpdf = o; o += al( (size_t)bsmax*8 );
      S.off_bcdf = o; o += al( (size_t)bsmax*8 );
      S.off_guide = o; o += al( (size_t)kSabGBStride*2 );
      S.off_heads = o; o += al( (size_t)nbmax*sizeof(SabHead) );
      S.off_beta = o; o += al( (size_t)nbmax*8 + 16 );
      S.total = o;
      dm.cls_ctas = nsab ? std::min( 4, (int)( ( 226u*1024u ) / ( S.total + 1024u ) ) ) : 0;
      dm.ncls = ( nsab && nc <= (uint32_t)kClsMax && dm.cls_ctas >= 1 ) ? nc : 0;
    }
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
        const uint32_t nn4 = ( (uint32_t)M.sc.nnormals + 3u ) & ~3u;
        const uint32_t nb = 3u*nn4*(uint32_t)sizeof(float);
        dm.sp_find.src[kHotSlotsIso] = M.sc.normals_f; dm.sp_find.nbytes[kHotSlotsIso] = nb; dm.sp_find.off[kHotSlotsIso] = 0;
        dm.sp_find.copy_bytes = nb;
        dm.sp_find.total = ( nb + 127u ) & ~127u;
        dm.sc_find_famof_off = dm.sp_find.total;
        dm.sc_find_scratch_off = ( dm.sc_find_famof_off + (uint32_t)M.sc.nnormals + 127u ) & ~127u;
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
      const SabT& T = dm.mat.sab[pl.sab_index];
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
      k_sab_rows<<< dim3( ( nb + 127 )/128, ne ), 128, 0, st >>>( T, rows, ainfo );
      k_sab_epoints<<< ( ne + 31 )/32, 32, 0, st >>>( T, rows, ep, bx, bpdf, bcdf, xscheck, errs );
      k_sab_guides<<< dim3( 4, ne + nb ), 256, 0, st >>>( T, ep, reinterpret_cast<uint16_t*>( base + pl.off_bguide ),
                                                       reinterpret_cast<uint16_t*>( base + pl.off_aguide ),
                                                       reinterpret_cast<double*>( base + pl.off_ascale ) );
      {
        const size_t nmax = std::max( (size_t)ne*nb, ntot );
        k_sab_gather_tabs<<< dim3( (unsigned)( ( nmax + 255 )/256 ), 2 ), 256, 0, st >>>(
          T, reinterpret_cast<SabHead*>( base + pl.off_heads ), reinterpret_cast<SabPoint*>( base + pl.off_pts ) );
      }
      g_launches += 6;
      CUDA_OK( cudaGetLastError() );
      std::vector<int> herrs( ne );
      CUDA_OK( cudaMemcpyAsync( herrs.data(), errs, (size_t)ne*4, cudaMemcpyDeviceToHost, st ) );
      CUDA_OK( cudaStreamSynchronize( st ) );
      for ( int e : herrs )
        if ( e )
          throw Err( "CalcError", "S(alpha,beta) sampler table build failed on device (code "+std::to_string(e)+")" );
    }
  }

  std::shared_ptr<DeviceMaterial> uploadMaterial( const void* blob, size_t nbytes )
  {
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount( &ndev );
    if ( ce != cudaSuccess || ndev <= 0 )
      throw Err( "CalcError", std::string("ncrystal_b200 requires a CUDA device (no CPU fallback): ")
                 + ( ce != cudaSuccess ? cudaGetErrorString(ce) : "no devices found" ) );
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
    buildStagePlan( *dm );
    ensureKernelAttrs( dm->device );
    buildSabTablesOnDevice( *dm, 0 );
    return dm;
  }

  void ensureKernelAttrs( int device )
  {
    static std::mutex m; static std::vector<int> done;
    std::lock_guard<std::mutex> g( m );
    for ( int d : done ) if ( d == device ) return;
    setSmemAttr( k_xs_iso );
    setSmemAttr( k_sample_classify );
    setSmemAttr( k_xs_aniso );
    setSmemAttr( k_sample_aniso );
    setSmemAttr( k_xs_aniso_pre );
    setSmemAttr( k_classify_aniso );
    setSmemAttr( k_sc_scan );
    setSmemAttr( k_sc_sample );
    setSmemAttr( k_sc_eval );
    setSmemAttr( k_sc_find );
    setSmemAttr( k_tally_hist );
    setSmemAttr( k_sab_classes<256,4> );
    setSmemAttr( k_sab_classes<320,3> );
    setSmemAttr( k_sab_classes<512,2> );
    setSmemAttr( k_sab_classes<1024,1> );
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
    static const size_t c = []{ const char* e = std::getenv( "NCB200_CHUNK" ); size_t v = e ? (size_t)std::atoll(e) : ( (size_t)1 << 20 ); return std::min( std::max<size_t>( v, 1024 ), kChunkMax ); }();
    return c;
  }
  constexpr int kSlots = 3;
  constexpr uint64_t kWindowMax = (uint64_t)1 << 24;   // neutrons staged on the device per window of the host-pointer path
  struct PipeSchedule { uint64_t first, max; double growth; double tail; };
  PipeSchedule pipeSchedule()
  {
    static const PipeSchedule ps = []{
      PipeSchedule p;
      const char* e0 = std::getenv( "NCB200_CHUNK0" );
      const char* eg = std::getenv( "NCB200_CHUNK_GROWTH" );
      p.max = chunkSize();
      p.first = std::min<uint64_t>( p.max, std::max<uint64_t>( 4096, e0 ? (uint64_t)std::atoll(e0) : ( (uint64_t)1 << 18 ) ) );
      p.growth = std::max( 1.0, eg ? std::atof(eg) : 2.0 );
      // ramp-down: a chunk is at most this fraction of what is left (0 = off), so the last copies out are short
      const char* et = std::getenv( "NCB200_CHUNK_TAIL" );
      p.tail = std::min( 1.0, std::max( 0.0, et ? std::atof(et) : 0.0 ) );   // measured: no gain (4.11 vs 4.15 ms per 1e7 samples) -> off
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
    std::vector<cudaEvent_t> ev_h, ev_c;    // per chunk: copy-in done, kernels done
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
      uint16_t* q_cls = nullptr; size_t ccap = 0;       // class of every q_sab entry
      uint32_t* cls_words = nullptr;                    // hist | start | fill | cursor | nticket(2) | tickets (uint16)
      cudaStream_t side = nullptr;              // free-gas kernels run here, concurrently with the table kernel
      cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    };
    QueueCtx qctx[kSlots+1];

    Scatter() { fp.tag = kScatterTag; fp.self = this; }
    ~Scatter()
    {
      if ( d_err ) cudaFree( d_err );
      for ( auto& c : qctx ) {
        if ( c.q ) cudaFree( c.q );
        if ( c.counts ) cudaFree( c.counts );
        if ( c.fg_prep ) { cudaFree( c.fg_prep ); cudaFree( c.fg_nd ); }
        if ( c.q_cls ) cudaFree( c.q_cls );
        if ( c.cls_words ) cudaFree( c.cls_words );
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
        CUDA_OK( cudaMalloc( &c.q, 6*c.cap*sizeof(uint32_t) ) );
      }
      return c;
    }
    // scratch of the class partition (after ensureQueues)
    void ensureClassScratch( QueueCtx& c )
    {
      if ( !c.cls_words )
        CUDA_OK( cudaMalloc( &c.cls_words, ( 4*( kClsMax + 1 ) + 2 )*sizeof(uint32_t) + kClsTicketsMax*sizeof(uint16_t) ) );
      if ( c.ccap >= c.cap ) return;
      if ( c.q_cls ) { CUDA_OK( cudaDeviceSynchronize() ); cudaFree( c.q_cls ); }
      c.ccap = c.cap;
      CUDA_OK( cudaMalloc( &c.q_cls, c.ccap*sizeof(uint16_t) ) );
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
        cudaEvent_t a, b;
        CUDA_OK( cudaEventCreateWithFlags( &a, cudaEventDisableTiming ) );
        CUDA_OK( cudaEventCreateWithFlags( &b, cudaEventDisableTiming ) );
        ev_h.push_back( a ); ev_c.push_back( b );
      }
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
    for ( auto& d : dirs ) {
      const std::string path = d + "/" + stem + ".ncb";
      auto data = readFile( path );
      if ( !data.empty() ) return data;
      tried += " " + path;
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
    static const bool group = []{ const char* e = std::getenv( "NCB200_FG_GROUP" ); return !e || std::atoi(e) != 0; }();
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
  std::atomic<uint64_t> g_fg_staged_min{ []{ const char* e = std::getenv( "NCB200_FG_STAGED_MIN" );
                                             return e ? (uint64_t)std::atoll(e) : (uint64_t)4000000; }() };
  void launchFgSampling( Scatter* s, const DeviceMaterial& dm, Scatter::QueueCtx& qc, const SampleArgs& A, const QueueArgs& Q,
                         uint64_t m, cudaStream_t st, bool timed )
  {
    static const int mode = []{ const char* e = std::getenv( "NCB200_FG_MODE" ); return e ? std::atoi(e) : 1; }();
    static const int fgctas = []{ const char* e = std::getenv( "NCB200_FG_CTAS" ); return e ? std::atoi(e) : 16; }();
    static const int fgminb = []{ const char* e = std::getenv( "NCB200_FG_MINB" ); return e ? std::atoi(e) : 8; }();
    static const int epl = []{ const char* e = std::getenv( "NCB200_FG_EPL" ); return e ? std::atoi(e) : 4; }();
    const uint64_t minbatch = g_fg_staged_min.load();
    const unsigned nsm = (unsigned)numSMs( dm.device );
    auto timer = [&]( const char* name ) { return std::unique_ptr<TimedLaunch>( timed ? new TimedLaunch( name, st ) : nullptr ); };
    if ( mode == 0 || m < minbatch ) {
      const unsigned g = (unsigned)std::min<uint64_t>( ( m + 127 )/128, (uint64_t)nsm*fgctas );
      auto tl = timer( "k_sample_fg" );
      if ( fgminb >= 8 ) k_sample_fg<8><<< g, 128, 0, st >>>( dm.mat, A, Q );
      else k_sample_fg<4><<< g, 128, 0, st >>>( dm.mat, A, Q );
      ++g_launches;
      return;
    }
    s->ensureFgPrep( qc );
    FgPrep P;
    P.r = qc.fg_prep; P.cap = qc.fcap; P.w = qc.fg_nd;
    P.cursor = Q.counts + 8 + 48;
    P.epl = (uint32_t)std::max( 1, epl );
    static const int batch = []{ const char* e = std::getenv( "NCB200_FG_BATCH" ); return e ? std::atoi(e) : 20; }();
    P.batch = (uint32_t)std::min( 32, std::max( 1, batch ) );
    const unsigned gflat = (unsigned)std::min<uint64_t>( ( m + 127 )/128, (uint64_t)nsm*16 );
    const unsigned grefill = (unsigned)std::min<uint64_t>( ( m + 127 )/128, (uint64_t)nsm*( fgminb >= 8 ? 8 : 6 ) );
    { auto tl = timer( "k_fg_prep" );
      k_fg_prep<<< gflat, 128, 0, st >>>( dm.mat, A, Q, P ); }
    { auto tl = timer( "k_fg_beta" );
      if ( fgminb >= 8 ) k_fg_beta<8><<< grefill, 128, 0, st >>>( dm.mat, A, Q, P );
      else k_fg_beta<6><<< grefill, 128, 0, st >>>( dm.mat, A, Q, P ); }
    { auto tl = timer( "k_fg_alpha_prep" );
      k_fg_alpha_prep<<< gflat, 128, 0, st >>>( dm.mat, A, Q, P ); }
    { auto tl = timer( "k_fg_alpha" );
      if ( fgminb >= 8 ) k_fg_alpha<8><<< grefill, 128, 0, st >>>( dm.mat, A, Q, P );
      else k_fg_alpha<6><<< grefill, 128, 0, st >>>( dm.mat, A, Q, P ); }
    { auto tl = timer( "k_fg_finish" );
      k_fg_finish<<< gflat, 128, 0, st >>>( dm.mat, A, Q, P ); }
    g_launches += 5;
  }

  // S(alpha,beta) table queue of a launch sequence.  Large batches: partition by overlay sampler and sample class
  // by class with the class tables staged in shared memory (ncb_kernels_cls.cuh); small batches (the chunks of a
  // short host call, transport steps, materials whose tables do not fit the plan): the lane-refill kernel on the
  // queue in arrival order.  NCB200_CLS_MIN overrides the threshold (0 = never use the class path).
  std::atomic<uint64_t> g_cls_min{ []{ const char* e = std::getenv( "NCB200_CLS_MIN" );
                                       return e ? (uint64_t)std::atoll(e) : (uint64_t)( 1u << 19 ); }() };
  bool useClassPath( const DeviceMaterial& dm, uint64_t m )
  {
    const uint64_t mn = g_cls_min.load();
    return dm.ncls && mn && m >= mn;
  }
  void launchSabQueue( Scatter* s, const DeviceMaterial& dm, Scatter::QueueCtx& qc, const SampleArgs& A, const QueueArgs& Q,
                       uint64_t m, cudaStream_t st, bool timed )
  {
    const unsigned nsm = (unsigned)numSMs( dm.device );
    auto timer = [&]( const char* name ) { return std::unique_ptr<TimedLaunch>( timed ? new TimedLaunch( name, st ) : nullptr ); };
    if ( !Q.q_cls ) {
      const unsigned gr = (unsigned)std::min<uint64_t>( ( m + 127 )/128, (uint64_t)nsm*8 );
      auto tl = timer( "k_sample_sab_refill" );
      k_sample_sab_refill<false,8><<< gr, 128, 0, st >>>( dm.mat, A, Q.q_sab, Q.counts + 0, Q.counts + 3 );
      ++g_launches;
      return;
    }
    ClassArgs K;
    uint32_t* w = qc.cls_words;
    K.q_in = Q.q_sab; K.q_cls = Q.q_cls; K.count = Q.counts + 0; K.q_out = qc.q + 4*qc.cap;
    K.hist = w; K.start = w + kClsMax; K.fill = w + 2*kClsMax + 1; K.cursor = w + 3*kClsMax + 1;
    K.nticket = w + 4*kClsMax + 1;
    K.tickets = reinterpret_cast<uint16_t*>( w + 4*( kClsMax + 1 ) + 2 );
    K.ncls = dm.ncls;
    static const int tmul = []{ const char* e = std::getenv( "NCB200_CLS_TICKETS" ); return e ? std::atoi(e) : 5; }();   // tickets per 2 resident CTAs
    static const int umin = []{ const char* e = std::getenv( "NCB200_CLS_UNIT" ); return e ? std::atoi(e) : 2048; }();
    K.tickets_target = std::max( 1u, nsm*(unsigned)dm.cls_ctas*(unsigned)tmul/2u );
    K.unit_min = (uint32_t)std::max( 256, umin );
    CUDA_OK( cudaMemsetAsync( K.hist, 0, dm.ncls*sizeof(uint32_t), st ) );
    { auto tl = timer( "k_cls_partition" );
      k_cls_hist<<< gridFor( m, 256, dm.device, 8 ), 256, 0, st >>>( K );
      k_cls_scan<<< 1, 1024, 0, st >>>( K );
      k_cls_partition<<< gridFor( ( m + 15 )/16, 256, dm.device, 4 ), 256, 0, st >>>( K ); }
    { auto tl = timer( "k_sab_classes" );
      const unsigned grid = nsm*(unsigned)dm.cls_ctas;
      const uint32_t sm = dm.cls_smem.total;
      switch ( dm.cls_ctas ) {
      case 4: k_sab_classes<256,4><<< grid, 256, sm, st >>>( dm.mat, A, K, Q, dm.cls_smem ); break;
      case 3: k_sab_classes<320,3><<< grid, 320, sm, st >>>( dm.mat, A, K, Q, dm.cls_smem ); break;
      case 2: k_sab_classes<512,2><<< grid, 512, sm, st >>>( dm.mat, A, K, Q, dm.cls_smem ); break;
      default: k_sab_classes<1024,1><<< grid, 1024, sm, st >>>( dm.mat, A, K, Q, dm.cls_smem ); break;
      } }
    g_launches += 4;
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
    static const uint64_t sub = []{ const char* e = std::getenv( "NCB200_SUBLAUNCH" );
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
      if ( useClassPath( dm, m ) ) {
        s->ensureClassScratch( qc );
        Q.q_cls = qc.q_cls;
        for ( int k = 0; k <= kMaxSab; ++k ) Q.cls_base[k] = dm.cls_base[k];
      }
      CUDA_OK( cudaMemsetAsync( qc.counts, 0, ( 8 + 2*kSortBins )*sizeof(uint32_t), st ) );
      { TimedLaunch tl( "k_sample_classify", st );
        k_sample_classify<<< gridFor( m, 256, dm.device, ctas ), 256, dm.sp.total, st >>>( dm.mat, dm.sp, A, Q ); }
      ++g_launches;
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
    static const bool v1 = []{ const char* e = std::getenv( "NCB200_ANISO_V1" ); return e && *e && *e != '0'; }();
    return v1;
  }

  int scCompIndex( const Material& M )
  {
    for ( int i = 0; i < M.ncomp; ++i ) if ( M.comp[i].kind == KIND_SCBRAGG ) return i;
    return -1;
  }

  // Batches below this size (the long tail of a transport run) are launch-latency bound: the all-in-one
  // thread-per-neutron kernels (1 launch instead of 2-8) are used for them.  NCB200_SMALL_V1 overrides (0 = never).
  uint64_t smallBatchV1()
  {
    static const uint64_t v = []{ const char* e = std::getenv( "NCB200_SMALL_V1" ); return e ? (uint64_t)std::atoll(e) : (uint64_t)0; }();
    return v;
  }

  // SCBragg scan (one warp per neutron) -> sc_xs / sc_n.  Default: lean candidate search (k_sc_find) + evaluation of
  // the neutrons that have candidates (k_sc_eval); NCB200_SC_ONEKERNEL=1 selects the combined k_sc_scan.
  void launchScScan( const DeviceMaterial& dm, Scatter::QueueCtx& qc, const double* d_ekin, const double* ux,
                     const double* uy, const double* uz, uint64_t n, cudaStream_t st, const uint32_t* n_dev = nullptr )
  {
    static const bool onekernel = []{ const char* e = std::getenv( "NCB200_SC_ONEKERNEL" ); return e && std::atoi(e) != 0; }();
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
    FA.work = qc.sc_work; FA.work_count = qc.counts + 6; FA.ncand = qc.sc_ncand; FA.cand = qc.sc_cand;
    FA.wpos = qc.sc_wpos;
    qc.sc_lists_valid = true;
    CUDA_OK( cudaMemsetAsync( qc.counts + 6, 0, sizeof(uint32_t), st ) );
    const int cf = std::min( 3, std::max( 1, (int)( ( 220u*1024u ) / std::max( dm.sc_find_smem, 1u ) ) ) );
    const uint64_t need_f = ( n + kScFindWarps - 1 ) / kScFindWarps;
    const int ce = std::min( 2, std::max( 1, (int)( ( 200u*1024u ) / std::max( dm.sc_smem, 1u ) ) ) );
    { TimedLaunch tl( "k_sc_find", st );
      k_sc_find<<< (unsigned)std::min<uint64_t>( need_f, (uint64_t)nsm*cf ), 32*kScFindWarps, dm.sc_find_smem, st >>>(
        dm.mat, dm.sp_find, FA, dm.sc_find_famof_off, dm.sc_find_scratch_off ); }
    { TimedLaunch tl( "k_sc_eval", st );
      k_sc_eval<<< (unsigned)std::min<uint64_t>( need, (uint64_t)nsm*ce ), 32*kScWarps, dm.sc_smem, st >>>(
        dm.mat, dm.sp_sc, FA, dm.sc_famof_off, dm.sc_scratch_off ); }
    g_launches += 2;
    CUDA_OK( cudaGetLastError() );
  }

  void launchXSAniso( Scatter* s, const double* d_ekin, const double* ux, const double* uy, const double* uz,
                      uint64_t n, double* d_out, cudaStream_t st, int ictx = kSlots )
  {
    if ( !n ) return;
    const DeviceMaterial& dm = *s->dm;
    const bool has_sc = scCompIndex( dm.mat ) >= 0;
    if ( useAnisoV1() || ( has_sc && !dm.sc_warp_ok ) || n < smallBatchV1() ) {
      DirArgs D; D.ux = ux; D.uy = uy; D.uz = uz; D.ox = D.oy = D.oz = nullptr;
      const int ctas = dm.sp.total > 56u*1024u ? 2 : 8;
      k_xs_aniso<<< gridFor( n, 128, dm.device, ctas ), 128, dm.sp.total, st >>>( dm.mat, dm.sp, d_ekin, D, n, d_out );
      ++g_launches;
      CUDA_OK( cudaGetLastError() );
      return;
    }
    const double* sc_xs = nullptr; const int32_t* sc_n = nullptr;
    if ( has_sc ) {
      Scatter::QueueCtx& qc = s->ensureAnisoBuffers( ictx, n );
      launchScScan( dm, qc, d_ekin, ux, uy, uz, n, st, s->n_dev_override );
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
    const bool v1 = useAnisoV1() || ( has_sc && !dm.sc_warp_ok ) || n < smallBatchV1();
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
      QueueArgs Q;
      Q.q_sab = qc.q; Q.q_fg = qc.q + qc.cap; Q.q_emax = qc.q + 2*qc.cap; Q.counts = qc.counts;
      AnisoArgs X;
      X.D = D; X.sc_xs = has_sc ? qc.sc_xs : nullptr; X.sc_n = has_sc ? qc.sc_n : nullptr;
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
        { TimedLaunch tl( "k_sc_sample", st );
          k_sc_sample<<< gs, 32*kScWarps, dm.sc_smem, st >>>( dm.mat, dm.sp_sc, A, X, dm.sc_famof_off, dm.sc_scratch_off ); }
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
    if ( flags & ~ERR_SAB_ISOFALLBACK )
      CUDA_OK( cudaMemset( s->d_err, 0, sizeof(int) ) );
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
  }

  // Host-pointer pipeline.  The call's arrays are staged in one device window (<= kWindowMax neutrons; longer
  // calls run window after window).  Three kinds of streams: one for H2D copies, kSlots for the kernels
  // (round robin, so the tail of one chunk's rejection kernels overlaps the next chunk), one for D2H copies;
  // events order chunk k's copy-in -> kernels -> copy-out, the host only blocks at the end of a window.
  // Chunks start small (the first D2H starts early) and grow geometrically up to chunkSize() (launch efficiency);
  // optionally (NCB200_CHUNK_TAIL) they shrink again towards the end of the call.
  // `launch(chunk_n, in_dev[], out_dev[], stream, slot)` enqueues the kernel(s).
  void runHostPipeline( Scatter* s, uint64_t n, int nin, const double* const* in, int nout, double* const* out,
                        const std::function<void(uint64_t,double* const*,double* const*,cudaStream_t,int)>& launch )
  {
    if ( !n ) return;
    const uint64_t W = std::min<uint64_t>( n, kWindowMax );
    s->ensurePipeline( (size_t)W * (size_t)( nin + nout ) );
    const PipeSchedule ps = pipeSchedule();
    // A failure while work is queued must not return to the caller (who then fills the output arrays with the
    // error sentinels) before the copies already queued into those arrays have drained.
    struct Drain {
      Scatter* s; bool armed = true;
      ~Drain() {
        if ( !armed ) return;
        cudaStreamSynchronize( s->st_h2d );
        for ( int c = 0; c < kSlots; ++c ) cudaStreamSynchronize( s->streams[c] );
        cudaStreamSynchronize( s->st_d2h );
      }
    } drain{ s };
    for ( uint64_t w0 = 0; w0 < n; w0 += W ) {
      const uint64_t wn = std::min<uint64_t>( W, n - w0 );
      uint64_t done = 0;
      size_t k = 0;
      double chunk = (double)ps.first;
      while ( done < wn ) {
        uint64_t m = std::min<uint64_t>( (uint64_t)chunk, wn - done );
        if ( ps.tail > 0.0 && w0 + wn == n )                    // (last window only)
          m = std::min<uint64_t>( m, std::max<uint64_t>( ps.first/2, (uint64_t)( ps.tail*(double)( wn - done ) ) ) );
        if ( wn - done - m < ps.first/4 ) m = wn - done;      // no tiny last chunk
        chunk = std::min<double>( chunk*ps.growth, (double)ps.max );
        s->ensureChunkEvents( k + 1 );
        const int slot = (int)( k % kSlots );
        cudaStream_t cs = s->streams[slot];
        double* din[4]; double* dout[4];
        for ( int a = 0; a < nin; ++a ) {
          din[a] = s->d_win + (size_t)a*W + done;
          CUDA_OK( cudaMemcpyAsync( din[a], in[a] + w0 + done, m*sizeof(double), cudaMemcpyHostToDevice, s->st_h2d ) );
        }
        CUDA_OK( cudaEventRecord( s->ev_h[k], s->st_h2d ) );
        for ( int a = 0; a < nout; ++a )
          dout[a] = s->d_win + (size_t)(nin+a)*W + done;
        CUDA_OK( cudaStreamWaitEvent( cs, s->ev_h[k], 0 ) );
        launch( m, din, dout, cs, slot );
        CUDA_OK( cudaEventRecord( s->ev_c[k], cs ) );
        CUDA_OK( cudaStreamWaitEvent( s->st_d2h, s->ev_c[k], 0 ) );
        for ( int a = 0; a < nout; ++a )
          CUDA_OK( cudaMemcpyAsync( out[a] + w0 + done, dout[a], m*sizeof(double), cudaMemcpyDeviceToHost, s->st_d2h ) );
        done += m;
        ++k;
      }
      CUDA_OK( cudaStreamSynchronize( s->st_d2h ) );
      for ( int c = 0; c < kSlots; ++c )
        CUDA_OK( cudaStreamSynchronize( s->streams[c] ) );
    }
    drain.armed = false;
  }

  void xsIsoHost( Scatter* s, const double* ekin, uint64_t n, uint64_t repeat, double* results )
  {
    if ( !n || !repeat ) return;
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

  // ---- oriented entry points: see ncb_lib_oriented.inc
#include "ncb_lib_oriented.inc"

  // ---- device-resident transport step: see ncb_lib_mmc.inc
#define NCB_MMC_CAPI
#include "ncb_lib_mmc.inc"
#undef NCB_MMC_CAPI

}

// ---- per-neutron boundaries with a caller-supplied generator (OpenMC virtual API, ncrystal_samplescatter_rs)
#include "ncb_lib_virtapi.inc"
