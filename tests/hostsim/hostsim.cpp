// tests/hostsim/hostsim.cpp -- TEST-ONLY host compilation of the device code.
//
// The per-neutron physics of the product lives in NCB_HD (__host__ __device__)
// functions under ncrystal_b200/csrc/*.cuh.  This file compiles those very
// functions with the host compiler and drives them with plain loops, so that the
// kernel logic can be unit-tested against the oracle on machines without a GPU
// (pytest -m "not gpu").  It is NOT part of the product: the package never loads
// this library and has no CPU fallback (ncrystal_b200/_lib.py fails loudly when
// the CUDA library is missing).
#define NCB_HOST_TRACE 1
#include "ncb_proc.cuh"
#include "ncb_sabbuild.cuh"
#include "ncb_loader.h"
#include "ncb_loader_sc.h"
#include "ncb_sabgrid.h"
#include "ncb_mmc.cuh"
#include <memory>
#include <cstdio>
#include <cstdlib>

namespace ncb {
  double g_erfc_lut_host[kErfcLutLen];
}
bool hostsim_expand_vdos_leaves( const void* blob, size_t nbytes, std::vector<unsigned char>& out );   // hostsim_vdos.cpp

namespace {
  struct Handle {
    ncb::LoadedMaterial lm;
    ncb::Material mat;
    ncb::HotTabs H;
  };
  struct LutInit { LutInit() { ncb::fillErfcLutHost( ncb::g_erfc_lut_host ); } } s_lutinit;
  thread_local std::string g_err;

  void buildSabHost( Handle& h )
  {
    using namespace ncb;
    unsigned char* base = h.lm.arena.data();
    for ( auto& pl : h.lm.sabplans ) {
      const SabT& T = h.mat.sab[pl.sab_index];
      const int na = T.nalpha, nb = T.nbeta, ne = T.negrid;
      double* logsab = reinterpret_cast<double*>( base + pl.off_logsab );
      double* cumul = reinterpret_cast<double*>( base + pl.off_cumul );
      for ( size_t i = 0; i < (size_t)na*nb; ++i )
        logsab[i] = sabLogS( T.sab[i] );
      for ( int ib = 0; ib < nb; ++ib )
        sabCumulRow( T.alpha, T.sab + (size_t)ib*na, logsab + (size_t)ib*na, na, cumul + (size_t)ib*na );
      SabRow* rows = reinterpret_cast<SabRow*>( base + pl.off_rows );
      SabAlphaInfo* ainfo = reinterpret_cast<SabAlphaInfo*>( base + pl.off_ainfo );
      SabEPoint* ep = reinterpret_cast<SabEPoint*>( base + pl.off_ep );
      double* bx = reinterpret_cast<double*>( base + pl.off_bx );
      double* bpdf = reinterpret_cast<double*>( base + pl.off_bpdf );
      double* bcdf = reinterpret_cast<double*>( base + pl.off_bcdf );
      double* xscheck = reinterpret_cast<double*>( base + pl.off_xscheck );
      if ( pl.auto_egrid ) {
        // energy grid determined from the kernel alone (ncb_sabgrid.h), as the product does on the device
        SabT& Tm = h.mat.sab[pl.sab_index];
        auto sigmaAt = [&]( const std::vector<double>& e ) {
          std::vector<double> xs( e.size() );
          std::vector<SabRow> r( nb );
          std::vector<SabAlphaInfo> inf( nb );
          std::vector<double> tx( T.bstride ), tp( T.bstride ), tc( T.bstride );
          for ( size_t k = 0; k < e.size(); ++k ) {
            for ( int ib = 0; ib < nb; ++ib )
              r[ib] = sabAnalyseRow( T.alpha, na, T.beta, T.sab, logsab, cumul, e[k]/T.kT, ib, inf[ib] );
            int err = 0; SabEPoint tmp;
            xs[k] = sabAssembleEPoint( T.beta, nb, T.kT, T.bound_xs, e[k], r.data(), 0, 0, tmp, tx.data(), tp.data(), tc.data(), err );
            if ( err ) throw std::runtime_error( "SAB energy-point analysis failed" );
          }
          if ( std::getenv( "NCB200_DEBUG_EGRID" ) )
            for ( size_t k = 0; k < e.size(); ++k ) std::fprintf( stderr, "egrid-probe %zu %.17g %.17g\n", k, e[k], xs[k] );
          return xs;
        };
        const std::vector<double> eg = sabDetermineEnergyGrid( ne, T.kT, T.beta[0], T.alpha[na-1], pl.suggested_emax, pl.req_emin, pl.req_emax, T.ext, sigmaAt, []( const char* ) {} );
        std::memcpy( base + pl.off_egrid, eg.data(), (size_t)ne*8 );
        const double l0 = std::log( eg[0] ), l1 = std::log( eg[ne-1] );
        Tm.egrid_log0 = l0; Tm.egrid_invdlog = l1 > l0 ? ( (double)ne - 1.0 )/( l1 - l0 ) : 0.0;
        int key0 = 0, shift = 0, nk = 0;
        const std::vector<uint16_t> lut = makeKeyLut( eg.data(), (size_t)ne, key0, shift, nk );
        if ( !lut.empty() && lut.size() <= kKeyLutMaxEntries ) {
          std::memcpy( base + pl.off_elut, lut.data(), lut.size()*sizeof(uint16_t) );
          Tm.elut = reinterpret_cast<const uint16_t*>( base + pl.off_elut );
          Tm.elut_key0 = key0; Tm.elut_shift = shift; Tm.elut_nk = nk;
        }
      }
      for ( int ie = 0; ie < ne; ++ie ) {
        const double ekin_div_kT = T.egrid[ie] / T.kT;
        for ( int ib = 0; ib < nb; ++ib )
          rows[(size_t)ie*nb+ib] = sabAnalyseRow( T.alpha, na, T.beta, T.sab, logsab, cumul, ekin_div_kT, ib, ainfo[(size_t)ie*nb+ib] );
        int err = 0;
        const uint32_t off_b = (uint32_t)( (size_t)ie*(size_t)T.bstride );
        xscheck[ie] = sabAssembleEPoint( T.beta, nb, T.kT, T.bound_xs, T.egrid[ie], rows + (size_t)ie*nb,
                                         off_b, (uint32_t)( (size_t)ie*nb ), ep[ie], bx + off_b, bpdf + off_b, bcdf + off_b, err );
        if ( err )
          throw std::runtime_error( "SAB table build failed (err="+std::to_string(err)+")" );
      }
      if ( pl.auto_egrid ) {
        SabT& Tm = h.mat.sab[pl.sab_index];
        std::memcpy( base + pl.off_xs, xscheck, (size_t)ne*8 );
        const double emax = T.egrid[ne-1], xs_emax = xscheck[ne-1], ext_emax = fgXS( Tm.ext, emax );
        Tm.k_extension = ( xs_emax - ext_emax )*emax;
        Tm.k1 = xs_emax*emax; Tm.k2 = ext_emax*emax;
      }
      // stage 3: guide tables
      uint16_t* bguide = reinterpret_cast<uint16_t*>( base + pl.off_bguide );
      uint16_t* aguide = reinterpret_cast<uint16_t*>( base + pl.off_aguide );
      double* ascale = reinterpret_cast<double*>( base + pl.off_ascale );
      for ( int ie = 0; ie < ne; ++ie ) {
        uint16_t* g = bguide + (size_t)ie*kSabGBStride;
        for ( int b = 0; b <= kSabGB; ++b )
          g[b] = sabBetaGuideEntry( bcdf + ep[ie].off_b, ep[ie].npts, b );
        ep[ie].guide = ep[ie].npts > 0 ? g : nullptr;
      }
      for ( int ib = 0; ib < nb; ++ib ) {
        const double* row = cumul + (size_t)ib*na;
        ascale[ib] = sabAlphaScale( row, na );
        for ( int b = 0; b <= kSabGA; ++b )
          aguide[(size_t)ib*( kSabGA+1 ) + b] = sabAlphaGuideEntry( row, na, ascale[ib], b );
      }
      // stage 4: gather-friendly copies + log guide
      SabHead* heads = reinterpret_cast<SabHead*>( base + pl.off_heads );
      SabTail* tails = reinterpret_cast<SabTail*>( base + pl.off_tails );
      SabPoint* pts = reinterpret_cast<SabPoint*>( base + pl.off_pts );
      SabBPoint* bpts = reinterpret_cast<SabBPoint*>( base + pl.off_bpts );
      uint16_t* lguide = reinterpret_cast<uint16_t*>( base + pl.off_lguide );
      for ( size_t k = 0; k < (size_t)ne*nb; ++k ) {
        heads[k] = sabMakeHead( ainfo[k], cumul + ( k % (size_t)nb )*na, na );
        sabMakeTails( ainfo[k], tails + 2*k );
      }
      for ( size_t k = 0; k < (size_t)nb*na; ++k )
        pts[k] = sabMakePoint( T.alpha, T.sab, logsab, cumul, na, k );
      for ( size_t k = 0; k < (size_t)ne*T.bstride; ++k ) { bpts[k].x = bx[k]; bpts[k].pdf = bpdf[k]; bpts[k].cdf = bcdf[k]; bpts[k].pad = 0.0; }
      for ( int ib = 0; ib < nb; ++ib ) {
        const double* row = cumul + (size_t)ib*na;
        const double inv = heads[ib].inv_total;      // (same for every energy point)
        for ( int key = 0; key <= kSabGL; ++key )
          lguide[(size_t)ib*kSabGLStride + key] = sabLogGuideEntry( row, na, inv, key );
      }
    }
  }
}

extern "C" {

  const char* hostsim_lasterror() { return g_err.c_str(); }

  void* hostsim_load( const void* blob, uint64_t nbytes )
  {
    try {
      auto h = std::make_unique<Handle>();
      std::vector<unsigned char> expanded;   // leaves given as a phonon density of states: expanded first (hostsim_vdos.cpp)
      if ( hostsim_expand_vdos_leaves( blob, nbytes, expanded ) ) { blob = expanded.data(); nbytes = expanded.size(); }
      ncb::loadBlob( blob, nbytes, h->lm );
      h->mat = ncb::relocated( h->lm, h->lm.arena.data() );
      ncb::hotTabsFromMaterial( h->mat, h->H );
      buildSabHost( *h );
      ncb::hotTabsFromMaterial( h->mat, h->H );   // (again: an automatic energy grid sets its key lut during the build)
      return h.release();
    } catch ( std::exception& e ) {
      g_err = e.what();
      return nullptr;
    }
  }
  void hostsim_free( void* h ) { delete static_cast<Handle*>(h); }

  int hostsim_ncomp( void* vh ) { return static_cast<Handle*>(vh)->mat.ncomp; }
  int hostsim_component_kind( void* vh, int c ) { return static_cast<Handle*>(vh)->mat.comp[c].kind; }

  void hostsim_xs_iso( void* vh, const double* ekin, uint64_t n, double* out )
  {
    auto& M = static_cast<Handle*>(vh)->mat;
    auto& H = static_cast<Handle*>(vh)->H;
    for ( uint64_t i = 0; i < n; ++i )
      out[i] = ncb::matXSIso( M, H, ekin[i], nullptr, nullptr );
  }

  // per-component unscaled xs out[c*n+i]
  void hostsim_xs_iso_components( void* vh, const double* ekin, uint64_t n, double* out )
  {
    auto& M = static_cast<Handle*>(vh)->mat;
    auto& H = static_cast<Handle*>(vh)->H;
    for ( int c = 0; c < M.ncomp; ++c )
      for ( uint64_t i = 0; i < n; ++i ) {
        int aux;
        out[c*n+i] = ncb::domainContains( M.comp[c].dom_lo, M.comp[c].dom_hi, ekin[i] ) ? ncb::compXSIso( M, H, c, ekin[i], aux ) : 0.0;
      }
  }

  void hostsim_sample_iso( void* vh, uint64_t seed, uint64_t first_index, const double* ekin, uint64_t n,
                           double* ekin_out, double* mu_out, uint32_t* ndraws, int32_t* errs )
  {
    auto& M = static_cast<Handle*>(vh)->mat;
    auto& H = static_cast<Handle*>(vh)->H;
    for ( uint64_t i = 0; i < n; ++i ) {
      ncb::Rng rng; rng.init( seed, first_index + i );
      int err = 0, ich;
      ncb::matSampleIso( M, H, ekin[i], rng, ekin_out[i], mu_out[i], err, ich );
      if ( ndraws ) ndraws[i] = rng.ndraws;
      if ( errs ) errs[i] = err;
    }
  }

  void hostsim_sample_iso_leaf( void* vh, int c, uint64_t seed, uint64_t first_index, const double* ekin, uint64_t n,
                                double* ekin_out, double* mu_out, uint32_t* ndraws, int32_t* errs )
  {
    auto& M = static_cast<Handle*>(vh)->mat;
    auto& H = static_cast<Handle*>(vh)->H;
    for ( uint64_t i = 0; i < n; ++i ) {
      ncb::Rng rng; rng.init( seed, first_index + i );
      int err = 0;
      ncb::compSampleIso( M, H, c, -1, ekin[i], rng, ekin_out[i], mu_out[i], err );
      if ( ndraws ) ndraws[i] = rng.ndraws;
      if ( errs ) errs[i] = err;
    }
  }

  // S(alpha,beta) leaf `c` through the short-chain variants of the table sampler (E below the table's Emax; above
  // it the plain path): must reproduce hostsim_sample_iso_leaf bit for bit
  void hostsim_sample_sab_staged( void* vh, int c, uint64_t seed, uint64_t first_index, const double* ekin, uint64_t n,
                                  double* ekin_out, double* mu_out, uint32_t* ndraws, int32_t* errs )
  {
    auto& M = static_cast<Handle*>(vh)->mat;
    const ncb::SabT& T = M.sab[M.comp[c].idx];
    for ( uint64_t i = 0; i < n; ++i ) {
      ncb::Rng rng; rng.init( seed, first_index + i );
      int err = 0;
      if ( ekin[i] < T.egrid[T.negrid-1] )
        ncb::sabSampleScatterFast( T, ekin[i], rng, ekin_out[i], mu_out[i], err );
      else
        ncb::sabSampleScatter( T, ekin[i], rng, ekin_out[i], mu_out[i], err );
      if ( ndraws ) ndraws[i] = rng.ndraws;
      if ( errs ) errs[i] = err;
    }
  }

  // design aid: trace of the "whole bins" alpha searches of the staged sampler (rows: ibeta, area, ilow, iupp, ga, gb, r0, pct)
  static std::vector<double>* s_trace = nullptr;
  static void traceHook( int ibeta, double area, int ilow, int iupp, int ga, int gb, int r0, double pct )
  {
    if ( s_trace ) { double v[8] = { (double)ibeta, area, (double)ilow, (double)iupp, (double)ga, (double)gb, (double)r0, pct }; s_trace->insert( s_trace->end(), v, v+8 ); }
  }
  uint64_t hostsim_alpha_trace( void* vh, int c, uint64_t seed, const double* ekin, uint64_t n, double* out, uint64_t maxrows )
  {
    auto& M = static_cast<Handle*>(vh)->mat;
    const ncb::SabT& T = M.sab[M.comp[c].idx];
    std::vector<double> tr; s_trace = &tr; ncb::g_alpha_trace = traceHook;
    for ( uint64_t i = 0; i < n; ++i ) {
      ncb::Rng rng; rng.init( seed, i );
      int err = 0; double eo, mu;
      if ( ekin[i] < T.egrid[T.negrid-1] ) ncb::sabSampleScatterFast( T, ekin[i], rng, eo, mu, err );
    }
    ncb::g_alpha_trace = nullptr; s_trace = nullptr;
    const uint64_t rows = std::min<uint64_t>( tr.size()/8, maxrows );
    std::memcpy( out, tr.data(), rows*8*sizeof(double) );
    return rows;
  }
  void hostsim_sab_dims( void* vh, int c, int* dims ) { auto& M = static_cast<Handle*>(vh)->mat; const ncb::SabT& T = M.sab[M.comp[c].idx]; dims[0]=T.negrid; dims[1]=T.nalpha; dims[2]=T.nbeta; }
  void hostsim_sab_cumul( void* vh, int c, double* out ) { auto& M = static_cast<Handle*>(vh)->mat; const ncb::SabT& T = M.sab[M.comp[c].idx]; std::memcpy( out, T.cumul, (size_t)T.nalpha*T.nbeta*8 ); }

  // energy-key lut (KeyLut) of component c's searched table: keyed vs plain upper_bound for arbitrary values;
  // returns the number of lut buckets (0: the table has no lut)
  int hostsim_keyed_upper_bound( void* vh, int c, const double* v, uint64_t n, int32_t* keyed, int32_t* plain )
  {
    auto& M = static_cast<Handle*>(vh)->mat;
    const ncb::Comp& cc = M.comp[c];
    if ( cc.kind == ncb::KIND_POWDERBRAGG ) {
      const ncb::PowderBraggT& T = M.pb[cc.idx];
      for ( uint64_t i = 0; i < n; ++i ) {
        keyed[i] = ncb::upperBoundKeyed( T.e2d, T.n, v[i], T.lut, T.lut_key0, T.lut_shift, T.lut_nk );
        plain[i] = ncb::upperBound( T.e2d, 0, T.n, v[i] );
      }
      return T.lut ? T.lut_nk : 0;
    }
    if ( cc.kind == ncb::KIND_SAB ) {
      const ncb::SabT& T = M.sab[cc.idx];
      for ( uint64_t i = 0; i < n; ++i ) {
        keyed[i] = ncb::upperBoundKeyed( T.egrid, T.negrid, v[i], T.elut, T.elut_key0, T.elut_shift, T.elut_nk );
        plain[i] = ncb::upperBound( T.egrid, 0, T.negrid, v[i] );
      }
      return T.elut ? T.elut_nk : 0;
    }
    return -1;
  }
  int hostsim_table_values( void* vh, int c, double* out, int nmax )
  {
    auto& M = static_cast<Handle*>(vh)->mat;
    const ncb::Comp& cc = M.comp[c];
    const double* a = nullptr; int n = 0;
    if ( cc.kind == ncb::KIND_POWDERBRAGG ) { a = M.pb[cc.idx].e2d; n = M.pb[cc.idx].n; }
    else if ( cc.kind == ncb::KIND_SAB ) { a = M.sab[cc.idx].egrid; n = M.sab[cc.idx].negrid; }
    for ( int i = 0; i < n && i < nmax; ++i ) out[i] = a[i];
    return n;
  }

  void hostsim_xs( void* vh, const double* ekin, const double* ux, const double* uy, const double* uz, uint64_t n, double* out )
  {
    auto& M = static_cast<Handle*>(vh)->mat;
    auto& H = static_cast<Handle*>(vh)->H;
    for ( uint64_t i = 0; i < n; ++i )
      out[i] = ncb::matXS( M, H, ekin[i], ncb::Vec3{ ux[i], uy[i], uz[i] }, nullptr, nullptr, nullptr );
  }

  void hostsim_sample( void* vh, uint64_t seed, uint64_t first_index, const double* ekin,
                       const double* ux, const double* uy, const double* uz, uint64_t n,
                       double* ekin_out, double* ox, double* oy, double* oz, uint32_t* ndraws, int32_t* errs )
  {
    auto& M = static_cast<Handle*>(vh)->mat;
    auto& H = static_cast<Handle*>(vh)->H;
    for ( uint64_t i = 0; i < n; ++i ) {
      ncb::Rng rng; rng.init( seed, first_index + i );
      int err = 0, ich;
      ncb::Vec3 o;
      ncb::matSample( M, H, ekin[i], ncb::Vec3{ ux[i], uy[i], uz[i] }, rng, ekin_out[i], o, err, ich );
      ox[i] = o.x; oy[i] = o.y; oz[i] = o.z;
      if ( ndraws ) ndraws[i] = rng.ndraws;
      if ( errs ) errs[i] = err;
    }
  }

  void hostsim_uniforms( uint64_t seed, uint64_t index, uint32_t n, double* out )
  {
    ncb::Rng rng; rng.init( seed, index );
    for ( uint32_t i = 0; i < n; ++i ) out[i] = rng.generate();
  }

  // same layout as refdrv_sab_sampler_dump; c = component index
  int hostsim_sab_sampler_dump( void* vh, int c, int iE, double* x, double* pdf, double* cdf, double* infos, double* meta )
  {
    auto& M = static_cast<Handle*>(vh)->mat;
    if ( c < 0 || c >= M.ncomp || M.comp[c].kind != ncb::KIND_SAB ) return -1;
    const ncb::SabT& T = M.sab[M.comp[c].idx];
    if ( iE < 0 || iE >= T.negrid ) return -1;
    const ncb::SabEPoint& ep = T.ep[iE];
    const int n = ep.npts;
    for ( int i = 0; i < n; ++i ) {
      if (x) x[i] = T.bx[ep.off_b+i];
      if (pdf) pdf[i] = T.bpdf[ep.off_b+i];
      if (cdf) cdf[i] = T.bcdf[ep.off_b+i];
    }
    if ( infos )
      for ( int i = 0; i+1 < n; ++i ) {
        const ncb::SabAlphaInfo& f = T.ainfo[ep.off_i+i];
        double* o = infos + 10*i;
        o[0]=f.f_alpha; o[1]=f.f_sval; o[2]=f.f_logsval; o[3]=f.f_idx;
        o[4]=f.b_alpha; o[5]=f.b_sval; o[6]=f.b_logsval; o[7]=f.b_idx;
        o[8]=f.prob_front; o[9]=f.prob_notback;
      }
    if ( meta ) { meta[0] = ep.ibeta_off; meta[1] = ep.first_bin_endpoint; }
    return n;
  }

  // energy grid of a S(alpha,beta) leaf (as loaded, or as determined by ncb_sabgrid.h)
  int hostsim_sab_egrid( void* vh, int c, double* out )
  {
    auto* h = static_cast<Handle*>(vh);
    auto& M = h->mat;
    if ( c < 0 || c >= M.ncomp || M.comp[c].kind != ncb::KIND_SAB ) return -1;
    const ncb::SabT& T = M.sab[M.comp[c].idx];
    for ( int i = 0; i < T.negrid; ++i ) out[i] = T.egrid[i];
    return T.negrid;
  }

  // total xs per energy point as recomputed by the native table builder
  int hostsim_sab_xscheck( void* vh, int c, double* out )
  {
    auto* h = static_cast<Handle*>(vh);
    auto& M = h->mat;
    if ( c < 0 || c >= M.ncomp || M.comp[c].kind != ncb::KIND_SAB ) return -1;
    const int isab = M.comp[c].idx;
    for ( auto& pl : h->lm.sabplans )
      if ( pl.sab_index == isab ) {
        const double* xc = reinterpret_cast<const double*>( h->lm.arena.data() + pl.off_xscheck );
        for ( int i = 0; i < M.sab[isab].negrid; ++i ) out[i] = xc[i];
        return M.sab[isab].negrid;
      }
    return -1;
  }

  // ---- transport step: the product's NCB_HD physics (ncb_mmc.cuh: source, geometry, forward step, tally values and
  // binning) driven history by history on the host; same argument layout as oracle_mmc.c::orc_minimc_run.
  struct hs_mmc_cfg {
    int geom_kind; double ga, gb, gc;
    int src_kind; double pos[3], dir[3]; double radius;
    int emode; double e0, e1; double weight;
    double roul_psurv, roul_wthr; int roul_nscat;
    int nscatlimit; int ignore_miss, include_abs; uint64_t seed;
  };
  int hostsim_minimc_run( void* vh, const hs_mmc_cfg* c, double numdens, double abs_c_mat, uint64_t first, uint64_t count,
                          int ntally, const int* types, const int* nbins, const double* xmin, const double* xmax,
                          double* out, double* meta )
  {
    using namespace ncb;
    auto& M = static_cast<Handle*>(vh)->mat;
    auto& H = static_cast<Handle*>(vh)->H;
    MmcGeom G; std::memset( &G, 0, sizeof(G) );
    G.kind = c->geom_kind;
    if ( G.kind == GEOM_SPHERE ) { G.a = c->ga; G.rsq = G.a*G.a; }
    else if ( G.kind == GEOM_SLAB ) { G.c = c->gc; G.unbounded = 1; }
    else if ( G.kind == GEOM_BOX ) { G.a = c->ga; G.b = c->gb; G.c = c->gc; }
    else { G.a = c->ga; G.rsq = G.a*G.a; G.b = c->gb; G.unbounded = ( G.b == 0.0 ); }
    MmcSource S; std::memset( &S, 0, sizeof(S) );
    S.kind = c->src_kind; S.w = c->weight;
    const double m2 = c->dir[0]*c->dir[0] + c->dir[1]*c->dir[1] + c->dir[2]*c->dir[2];
    const double fn = 1.0/std::sqrt( m2 );
    for ( int k = 0; k < 3; ++k ) { S.pos[k] = c->pos[k]; S.dir[k] = c->dir[k]*fn; }
    if ( S.kind == SRC_CIRCULAR && c->radius > 0.0 ) {
      double a[3] = { 1, 0, 0 };
      auto dot = [&]( const double* q ) { return q[0]*S.dir[0] + q[1]*S.dir[1] + q[2]*S.dir[2]; };
      if ( dot( a ) > 0.8 ) { a[0] = 0; a[1] = 1; a[2] = 0; }
      if ( dot( a ) > 0.8 ) { a[0] = 0; a[1] = 0; a[2] = 1; }
      auto cross_unit = [&]( const double* p, const double* q, double* o ) {
        const double c0 = p[1]*q[2] - p[2]*q[1], c1 = p[2]*q[0] - p[0]*q[2], c2 = p[0]*q[1] - p[1]*q[0];
        const double g = 1.0/std::sqrt( c0*c0 + c1*c1 + c2*c2 );
        o[0] = c0*g; o[1] = c1*g; o[2] = c2*g;
      };
      double va[3], vb[3];
      cross_unit( S.dir, a, va ); cross_unit( S.dir, va, vb );
      for ( int k = 0; k < 3; ++k ) { S.va[k] = va[k]*c->radius; S.vb[k] = vb[k]*c->radius; }
    }
    if ( S.kind == SRC_ISOTROPIC ) S.minus_r = c->radius ? -c->radius : 0.0;
    const int emap[6] = { SRCE_FIXED, SRCE_UNIFORM_EKIN, SRCE_UNIFORM_WL, SRCE_LOGNORMAL_EKIN, SRCE_LOGNORMAL_WL, SRCE_MAXWELL };
    S.emode = emap[c->emode]; S.e0 = c->e0; S.e1 = c->e1;
    if ( c->emode == 3 || c->emode == 4 ) {
      const double tmp = 1 + ( c->e1*c->e1 )/( c->e0*c->e0 );
      S.e0 = std::log( c->e0 / std::sqrt( tmp ) ); S.e1 = std::sqrt( std::log( tmp ) );
    }
    if ( c->emode == 5 ) { S.e0 = 0.5*kBoltzmann*c->e0; S.e1 = 0.0; }
    auto inside = [&]() {
      const double x = S.pos[0], y = S.pos[1], z = S.pos[2];
      switch ( G.kind ) {
      case GEOM_SPHERE: return x*x + y*y + z*z <= G.rsq;
      case GEOM_SLAB: return std::fabs(z) <= G.c;
      case GEOM_BOX: return std::fabs(x) <= G.a && std::fabs(y) <= G.b && std::fabs(z) <= G.c;
      default: return x*x + z*z <= G.rsq && ( G.b == 0 || std::fabs(y) <= G.b );
      }
    };
    S.may_be_outside = ( S.kind == SRC_CIRCULAR && c->radius > 0.0 ) ? 1 : ( inside() ? 0 : 1 );
    MmcEngine E;
    E.macro_factor = 100.0*numdens;
    E.abs_c = c->include_abs ? std::max( 0.0, abs_c_mat ) : 0.0;
    E.roulette_psurv = c->roul_psurv; E.roulette_wthr = c->roul_wthr; E.roulette_nscat = c->roul_nscat;
    E.roulette_boost = 1.0/c->roul_psurv; E.nscatlimit = c->nscatlimit;
    MmcTally T; std::memset( &T, 0, sizeof(T) );
    T.nh = ntally;
    uint32_t off = 0;
    for ( int i = 0; i < ntally; ++i ) {
      MmcHist& h = T.h[i];
      h.type = types[i]; h.nbins = nbins[i]; h.xmin = xmin[i]; h.xmax = xmax[i];
      h.invdelta = 1.0/( ( xmax[i] - xmin[i] )/nbins[i] ); h.off = off;
      const uint32_t nd = mmcHistDoubles( nbins[i] );
      std::fill( out + off, out + off + nd, 0.0 );
      for ( int k = 0; k < kMmcNClass; ++k ) {
        out[off + 2*kMmcNClass*( nbins[i] + 2 ) + k*kMmcNStat + 3] = kInf;
        out[off + 2*kMmcNClass*( nbins[i] + 2 ) + k*kMmcNStat + 4] = -kInf;
      }
      off += nd;
    }
    for ( int k = 0; k < 3; ++k ) T.dir0[k] = S.dir[k];
    T.dir0_is_z = ( S.dir[2] == 1.0 ); T.has_dir0_fixed = ( S.kind != SRC_ISOTROPIC );
    T.has_e0_fixed = ( S.emode == SRCE_FIXED ); T.e0_fixed = S.e0;
    double miss_n = 0, miss_w = 0, tall_n = 0, tall_w = 0, nsteps = 0;
    int errs = 0;
    auto record = [&]( double ux, double uy, double uz, double ekin, double w, int nscat, int ninel, double e_init,
                       double ux0, double uy0, double uz0 ) {
      tall_n += 1; tall_w += w;
      for ( int ih = 0; ih < T.nh; ++ih ) {
        const MmcHist& h = T.h[ih];
        bool weighted;
        const double v = mmcTallyValue( T, h.type, ux, uy, uz, ekin, w, nscat, e_init, ux0, uy0, uz0, weighted );
        const double wgt = weighted ? w : 1.0;
        if ( !( wgt > 0.0 ) ) continue;
        const int cls = mmcClass( nscat, ninel ), nb2 = h.nbins + 2, bin = mmcValueToBin( h, v );
        double* g = out + h.off;
        g[cls*nb2 + bin] += wgt;
        g[kMmcNClass*nb2 + cls*nb2 + bin] += wgt*wgt;
        double* st = g + 2*kMmcNClass*nb2 + cls*kMmcNStat;
        st[0] += wgt; st[1] += wgt*v; st[2] += wgt*v*v;
        if ( v < st[3] ) st[3] = v;
        if ( v > st[4] ) st[4] = v;
      }
    };
    for ( uint64_t id = first; id < first + count; ++id ) {
      Rng rs; rs.init( c->seed, id, kMmcSidSrc );
      MmcNeutron nt = mmcGenerate( S, rs );
      const double e_init = nt.ekin, ux0 = nt.ux, uy0 = nt.uy, uz0 = nt.uz;
      if ( S.may_be_outside ) {
        const double d = mmcDistToEntry( G, nt.x, nt.y, nt.z, nt.ux, nt.uy, nt.uz );
        if ( d < 0.0 ) {
          miss_n += 1; miss_w += nt.w;
          if ( !c->ignore_miss ) record( nt.ux, nt.uy, nt.uz, nt.ekin, nt.w, -1, 0, e_init, ux0, uy0, uz0 );
          continue;
        }
        nt.x += d*nt.ux; nt.y += d*nt.uy; nt.z += d*nt.uz;
      }
      int nscat = 0, ninel = 0;
      for ( uint32_t step = 0; ; ++step ) {
        nsteps += 1;
        Rng rng; rng.init( c->seed, id, kMmcSidBase + 2u*step );
        const double xs = matXS( M, H, nt.ekin, Vec3{ nt.ux, nt.uy, nt.uz }, nullptr, nullptr, nullptr );
        MmcStepOut o = mmcForward( G, E, rng, nt.x, nt.y, nt.z, nt.ux, nt.uy, nt.uz, nt.w, nt.ekin, nscat, xs );
        record( nt.ux, nt.uy, nt.uz, nt.ekin, o.wt, nscat, ninel, e_init, ux0, uy0, uz0 );
        if ( !o.survives ) break;
        nt.x = o.x; nt.y = o.y; nt.z = o.z; nt.w = o.w;
        Rng rq; rq.init( c->seed, id, kMmcSidBase + 2u*step + 1u );
        double eout; Vec3 od; int err = 0, ich;
        matSample( M, H, nt.ekin, Vec3{ nt.ux, nt.uy, nt.uz }, rq, eout, od, err, ich );
        errs |= err;
        const bool was_elastic = ( nt.ekin == eout );
        nt.ux = od.x; nt.uy = od.y; nt.uz = od.z; nt.ekin = eout;
        ++nscat; if ( !was_elastic ) ++ninel;
        if ( step > 100000u ) return -1;
      }
    }
    meta[0] = miss_n; meta[1] = miss_w; meta[2] = tall_n; meta[3] = tall_w; meta[4] = nsteps;
    return errs;
  }

}
