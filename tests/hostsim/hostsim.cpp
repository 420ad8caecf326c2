// tests/hostsim/hostsim.cpp -- TEST-ONLY host compilation of the device code.
//
// The per-neutron physics of the product lives in NCB_HD (__host__ __device__)
// functions under ncrystal_b200/csrc/*.cuh.  This file compiles those very
// functions with the host compiler and drives them with plain loops, so that the
// kernel logic can be unit-tested against the oracle on machines without a GPU
// (pytest -m "not gpu").  It is NOT part of the product: the package never loads
// this library and has no CPU fallback (ncrystal_b200/_lib.py fails loudly when
// the CUDA library is missing).
#include "ncb_proc.cuh"
#include "ncb_sabbuild.cuh"
#include "ncb_loader.h"
#include "ncb_loader_sc.h"
#include <memory>

namespace ncb {
  double g_erfc_lut_host[kErfcLutLen];
}

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
      for ( int ie = 0; ie < ne; ++ie ) {
        const double ekin_div_kT = T.egrid[ie] / T.kT;
        for ( int ib = 0; ib < nb; ++ib )
          rows[(size_t)ie*nb+ib] = sabAnalyseRow( T.alpha, na, T.beta, T.sab, logsab, cumul, ekin_div_kT, ib, ainfo[(size_t)ie*nb+ib] );
        int err = 0;
        const uint32_t off_b = (uint32_t)( (size_t)ie*(nb+1) );
        xscheck[ie] = sabAssembleEPoint( T.beta, nb, T.kT, T.bound_xs, T.egrid[ie], rows + (size_t)ie*nb,
                                         off_b, (uint32_t)( (size_t)ie*nb ), ep[ie], bx + off_b, bpdf + off_b, bcdf + off_b, err );
        if ( err )
          throw std::runtime_error( "SAB table build failed (err="+std::to_string(err)+")" );
      }
      // stage 3: guide tables
      uint16_t* bguide = reinterpret_cast<uint16_t*>( base + pl.off_bguide );
      uint16_t* aguide = reinterpret_cast<uint16_t*>( base + pl.off_aguide );
      double* ascale = reinterpret_cast<double*>( base + pl.off_ascale );
      for ( int ie = 0; ie < ne; ++ie ) {
        uint16_t* g = bguide + (size_t)ie*( kSabGB+1 );
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
    }
  }
}

extern "C" {

  const char* hostsim_lasterror() { return g_err.c_str(); }

  void* hostsim_load( const void* blob, uint64_t nbytes )
  {
    try {
      auto h = std::make_unique<Handle>();
      ncb::loadBlob( blob, nbytes, h->lm );
      h->mat = ncb::relocated( h->lm, h->lm.arena.data() );
      ncb::hotTabsFromMaterial( h->mat, h->H );
      buildSabHost( *h );
      return h.release();
    } catch ( std::exception& e ) {
      g_err = e.what();
      return nullptr;
    }
  }
  void hostsim_free( void* h ) { delete static_cast<Handle*>(h); }

  int hostsim_ncomp( void* vh ) { return static_cast<Handle*>(vh)->mat.ncomp; }

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
}
