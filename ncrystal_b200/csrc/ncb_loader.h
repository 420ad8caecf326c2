// ncb_loader.h -- host side: compiled-material blob (ncb_blob.h) -> arena image +
// Material descriptor (ncb_tables.h).  Header-only, no CUDA: the C-ABI library
// uploads the arena to HBM and runs the table-build kernels there; the CPU-side
// unit tests (tests/hostsim) run the same image in host memory.
//
// The arena is ONE contiguous allocation holding every array of the material
// (inputs copied from the blob, derived tables zero-initialised and filled by
// the SAB build stages), 256-byte aligned sections.
#pragma once
#include "ncb_blob.h"
#include "ncb_tables.h"
#include <algorithm>
#include <cstring>
#include <initializer_list>
#include <stdexcept>
#include <string>
#include <vector>

namespace ncb {

  struct SabBuildPlan {       // what the build stages need for one SAB leaf
    int sab_index;            // index into Material::sab
    size_t off_logsab, off_cumul, off_ep, off_bx, off_bpdf, off_bcdf, off_ainfo, off_rows, off_xscheck;
    size_t off_bguide, off_aguide, off_ascale, off_heads, off_pts, off_tails, off_bpts, off_lguide;
    // energy grid determined by the library (ncb_sab_t::auto_egrid): where egrid / xs / the key lut live in the arena
    bool auto_egrid = false;
    double suggested_emax = 0.0, req_emin = 0.0, req_emax = 0.0;   // (request: first two entries of the placeholder egrid[])
    size_t off_egrid = 0, off_xs = 0, off_elut = 0;
  };
  constexpr size_t kKeyLutMaxEntries = 4100;   // uint16 entries reserved for a key lut that is built after the upload

  struct LoadedMaterial {
    Material mat;                       // pointers are OFFSETS into the arena until relocate()
    std::vector<unsigned char> arena;   // host image
    std::vector<SabBuildPlan> sabplans;
    std::string cfg;
    double numdens = 0.0, abs_c = 0.0, temperature = -1.0;   // bulk quantities for the transport step
    bool relocated = false;

    size_t reserve( size_t nbytes )
    {
      size_t off = ( arena.size() + 255u ) & ~(size_t)255u;
      arena.resize( off + nbytes, 0 );
      return off;
    }
    size_t put( const void* src, size_t nbytes )
    {
      size_t off = reserve( nbytes );
      std::memcpy( arena.data() + off, src, nbytes );
      return off;
    }
  };

  template <class T> inline const T* offAsPtr( size_t off ) { return reinterpret_cast<const T*>( off ); }
  template <class T> inline void relocPtr( const T*& p, const unsigned char* base )
  {
    p = reinterpret_cast<const T*>( base + reinterpret_cast<size_t>( p ) );
  }

  // Convert offsets in lm.mat to pointers based at `base` (device or host address of the arena copy).
  inline Material relocated( const LoadedMaterial& lm, const void* base_ )
  {
    const unsigned char* base = static_cast<const unsigned char*>( base_ );
    Material m = lm.mat;
    int npb = 0, nsab = 0;
    for ( int i = 0; i < m.ncomp; ++i ) {
      if ( m.comp[i].kind == KIND_POWDERBRAGG ) ++npb;
      if ( m.comp[i].kind == KIND_SAB ) ++nsab;
    }
    for ( int i = 0; i < npb; ++i ) {
      relocPtr( m.pb[i].e2d, base );
      relocPtr( m.pb[i].fdm, base );
      if ( m.pb[i].lut ) relocPtr( m.pb[i].lut, base );
    }
    for ( int i = 0; i < nsab; ++i ) {
      SabT& s = m.sab[i];
      if ( s.elut ) relocPtr( s.elut, base );
      relocPtr( s.egrid, base ); relocPtr( s.xs, base ); relocPtr( s.alpha, base ); relocPtr( s.beta, base );
      relocPtr( s.sab, base ); relocPtr( s.logsab, base ); relocPtr( s.cumul, base );
      relocPtr( s.ep, base ); relocPtr( s.bx, base ); relocPtr( s.bpdf, base ); relocPtr( s.bcdf, base );
      relocPtr( s.ainfo, base );
      relocPtr( s.bguide, base ); relocPtr( s.aguide, base ); relocPtr( s.ascale, base );
      relocPtr( s.heads, base ); relocPtr( s.pts, base ); relocPtr( s.tails, base ); relocPtr( s.bpts, base ); relocPtr( s.lguide, base );
    }
    if ( m.lc.nplanes ) {
      relocPtr( m.lc.planes, base ); relocPtr( m.sc.sofcosd.data, base ); relocPtr( m.sc.evalcosx.data, base );
    }
    if ( m.sc.nfam ) {
      relocPtr( m.sc.fam_xsfact, base ); relocPtr( m.sc.fam_inv2d, base ); relocPtr( m.sc.fam_first, base );
      relocPtr( m.sc.normals, base ); relocPtr( m.sc.normals_f, base ); relocPtr( m.sc.sofcosd.data, base ); relocPtr( m.sc.evalcosx.data, base );
    }
    return m;
  }

  // Energy-key table (KeyLut) over an ascending table of positive energies; appended to the arena.  The number of
  // mantissa bits in the key is the largest (<= 8) that keeps the table within ~2 entries per point and 4096 entries.
  inline size_t buildKeyLut( LoadedMaterial& lm, const double* a, size_t n, int& key0, int& shift, int& nk )
  {
    key0 = shift = nk = 0;
    if ( n < 8 || !( a[0] > 0.0 ) || n > 65535 )
      return 0;
    auto bits = []( double v ) { long long b; std::memcpy( &b, &v, sizeof(b) ); return b; };
    int m = 8;
    long long span = 0;
    for ( ; m >= 0; --m ) {
      span = ( bits( a[n-1] ) >> ( 52 - m ) ) - ( bits( a[0] ) >> ( 52 - m ) ) + 1;
      if ( span + 1 <= 4096 && span <= (long long)std::max<size_t>( 2*n, 64 ) ) break;
    }
    if ( m < 0 )
      return 0;
    shift = 52 - m;
    key0 = (int)( bits( a[0] ) >> shift );
    nk = (int)span;
    std::vector<uint16_t> lut( (size_t)nk + 1 );
    size_t i = 0;
    for ( int k = 0; k <= nk; ++k ) {
      while ( i < n && ( bits( a[i] ) >> shift ) - key0 < k ) ++i;
      lut[k] = (uint16_t)i;
    }
    return lm.put( lut.data(), lut.size()*sizeof(uint16_t) );
  }

  // Same table as buildKeyLut, into a vector (for a grid that only exists after the upload)
  inline std::vector<uint16_t> makeKeyLut( const double* a, size_t n, int& key0, int& shift, int& nk )
  {
    LoadedMaterial tmp;
    tmp.reserve( 256 );
    const size_t off = buildKeyLut( tmp, a, n, key0, shift, nk );
    std::vector<uint16_t> lut;
    if ( off ) {
      lut.resize( (size_t)nk + 1 );
      std::memcpy( lut.data(), tmp.arena.data() + off, lut.size()*sizeof(uint16_t) );
    }
    return lut;
  }

  void loadScBragg( LoadedMaterial& lm, const unsigned char* blob, const ncb_comp_t& c ); // ncb_loader_sc.h
  void loadLcBragg( LoadedMaterial& lm, const unsigned char* blob, const ncb_comp_t& c ); // ncb_loader_sc.h

  // Payload validation: the array counts inside a component payload come from the (untrusted) buffer; before any
  // copy the payload must hold its header plus `ndoubles` fp64 values.  Counts are bounded first so that the sums
  // and products below cannot wrap.
  constexpr uint64_t kMaxBlobCount = (uint64_t)1 << 31;
  inline void checkPayload( const ncb_comp_t& c, size_t header_bytes, std::initializer_list<uint64_t> counts,
                            uint64_t ndoubles, const char* what )
  {
    for ( uint64_t n : counts )
      if ( n > kMaxBlobCount )
        throw std::runtime_error( std::string("compiled material: implausible array length in ")+what+" payload" );
    if ( ndoubles > ( (uint64_t)1 << 40 ) || c.nbytes < header_bytes || ( c.nbytes - header_bytes ) / 8 < ndoubles )
      throw std::runtime_error( std::string("compiled material: truncated ")+what+" payload" );
  }

  inline void loadBlob( const void* blob_, size_t nbytes, LoadedMaterial& lm )
  {
    const unsigned char* blob = static_cast<const unsigned char*>( blob_ );
    if ( nbytes < sizeof(ncb_header_t) )
      throw std::runtime_error( "compiled material: buffer too small" );
    ncb_header_t hdr;
    std::memcpy( &hdr, blob, sizeof(hdr) );
    if ( hdr.magic != NCB_MAGIC )
      throw std::runtime_error( "compiled material: bad magic" );
    if ( hdr.version != NCB_VERSION )
      throw std::runtime_error( "compiled material: unsupported version" );
    if ( hdr.nbytes > nbytes || hdr.ncomp == 0 || hdr.ncomp > (uint32_t)kMaxComp )
      throw std::runtime_error( "compiled material: inconsistent header" );
    hdr.cfg[sizeof(hdr.cfg)-1] = 0;
    lm.cfg = hdr.cfg;
    lm.numdens = hdr.numdens; lm.abs_c = hdr.abs_c; lm.temperature = hdr.temperature;
    Material& m = lm.mat;
    std::memset( &m, 0, sizeof(m) );
    m.ncomp = (int)hdr.ncomp;
    m.oriented = (int)hdr.oriented;
    m.dom_lo = hdr.dom_lo;
    m.dom_hi = hdr.dom_hi;
    lm.reserve( 256 ); // offset 0 is never a valid array => null stays null
    int npb = 0, nel = 0, nfg = 0, nsab = 0;
    for ( int i = 0; i < m.ncomp; ++i ) {
      const ncb_comp_t& c = hdr.comp[i];
      if ( c.off > hdr.nbytes || c.nbytes > hdr.nbytes - c.off || c.off < sizeof(ncb_header_t) )
        throw std::runtime_error( "compiled material: component out of bounds" );
      if ( c.off % 8 != 0 )
        throw std::runtime_error( "compiled material: misaligned component payload" );
      Comp& k = m.comp[i];
      k.kind = (int)c.kind;
      k.scale = c.scale;
      k.dom_lo = c.dom_lo;
      k.dom_hi = c.dom_hi;
      const unsigned char* p = blob + c.off;
      switch ( c.kind ) {
      case NCB_KIND_POWDERBRAGG: {
        if ( npb >= (int)( sizeof(m.pb)/sizeof(m.pb[0]) ) ) throw std::runtime_error( "too many PowderBragg components" );
        checkPayload( c, sizeof(ncb_powderbragg_t), {}, 0, "PowderBragg" );
        ncb_powderbragg_t h; std::memcpy( &h, p, sizeof(h) );
        checkPayload( c, sizeof(h), { h.nplanes }, 2*h.nplanes, "PowderBragg" );
        if ( h.nplanes == 0 ) throw std::runtime_error( "compiled material: PowderBragg without planes" );
        const double* arr = reinterpret_cast<const double*>( p + sizeof(h) );
        PowderBraggT& T = m.pb[npb];
        T.n = (int)h.nplanes;
        T.threshold = h.threshold;
        T.e2d = offAsPtr<double>( lm.put( arr, h.nplanes*8 ) );
        T.fdm = offAsPtr<double>( lm.put( arr + h.nplanes, h.nplanes*8 ) );
        T.lut = offAsPtr<uint16_t>( buildKeyLut( lm, arr, h.nplanes, T.lut_key0, T.lut_shift, T.lut_nk ) );
        k.idx = npb++;
        break;
      }
      case NCB_KIND_ELINC: {
        if ( nel >= 1 ) throw std::runtime_error( "too many ElIncScatter components" );
        checkPayload( c, sizeof(ncb_elinc_t), {}, 0, "ElIncScatter" );
        ncb_elinc_t h; std::memcpy( &h, p, sizeof(h) );
        if ( h.nelem > (uint64_t)kMaxElIncElems ) throw std::runtime_error( "too many ElInc elements" );
        checkPayload( c, sizeof(h), { h.nelem }, 2*h.nelem, "ElIncScatter" );
        const double* arr = reinterpret_cast<const double*>( p + sizeof(h) );
        ElIncT& T = m.elinc[nel];
        T.n = (int)h.nelem;
        for ( int j = 0; j < T.n; ++j ) { T.msd[j] = arr[j]; T.bixs[j] = arr[h.nelem+j]; }
        k.idx = nel++;
        break;
      }
      case NCB_KIND_FREEGAS: {
        if ( nfg >= (int)( sizeof(m.fg)/sizeof(m.fg[0]) ) ) throw std::runtime_error( "too many FreeGas components" );
        checkPayload( c, sizeof(ncb_freegas_t), {}, 0, "FreeGas" );
        ncb_freegas_t h; std::memcpy( &h, p, sizeof(h) );
        FreeGasT& T = m.fg[nfg];
        T.sigma_free = h.sigma_free;
        T.ca = h.ca;
        T.kT = 8.6173303e-5 * h.temperature; // Temperature::kT(), NCTypes.hh:754
        T.mass_amu = h.mass_amu;
        k.idx = nfg++;
        break;
      }
      case NCB_KIND_SAB: {
        if ( nsab >= (int)( sizeof(m.sab)/sizeof(m.sab[0]) ) ) throw std::runtime_error( "too many SABScatter components" );
        checkPayload( c, sizeof(ncb_sab_t), {}, 0, "SABScatter" );
        ncb_sab_t h; std::memcpy( &h, p, sizeof(h) );
        if ( h.negrid < 2 || h.nalpha < 2 || h.nbeta < 2 ) throw std::runtime_error( "compiled material: degenerate SAB grids" );
        if ( h.nbeta + 1 > 65535 || h.nalpha > 65535 || h.negrid > 65535 )
          throw std::runtime_error( "compiled material: SAB grids too large for 16-bit guide tables" );
        checkPayload( c, sizeof(h), { h.negrid, h.nalpha, h.nbeta }, 2*h.negrid + h.nalpha + h.nbeta + h.nalpha*h.nbeta, "SABScatter" );
        const double* arr = reinterpret_cast<const double*>( p + sizeof(h) );
        SabT& T = m.sab[nsab];
        T.scale = h.scale;
        T.kT = 8.6173303e-5 * h.temperature;
        T.k_extension = h.k_extension;
        T.k1 = h.k1; T.k2 = h.k2;
        T.egrid_margin = h.egrid_margin;
        if ( !h.auto_egrid ) {
          const double* eg = reinterpret_cast<const double*>( p + sizeof(h) );
          const double l0 = std::log( eg[0] ), l1 = std::log( eg[h.negrid-1] );
          T.egrid_log0 = l0;
          T.egrid_invdlog = ( h.negrid > 1 && l1 > l0 ) ? ( (double)h.negrid - 1.0 )/( l1 - l0 ) : 0.0;
        }
        T.bound_xs = h.bound_xs;
        T.ext.sigma_free = h.ext_sigma_free;
        T.ext.ca = h.ext_ca;
        T.ext.kT = 8.6173303e-5 * h.ext_temperature;
        T.ext.mass_amu = h.ext_mass_amu;
        T.negrid = (int)h.negrid; T.nalpha = (int)h.nalpha; T.nbeta = (int)h.nbeta;
        const size_t ne = h.negrid, na = h.nalpha, nb = h.nbeta;
        const size_t off_egrid = lm.put( arr, ne*8 ), off_xs = lm.put( arr + ne, ne*8 );
        T.egrid = offAsPtr<double>( off_egrid );
        T.xs    = offAsPtr<double>( off_xs );
        size_t off_elut = 0;
        if ( h.auto_egrid ) {
          if ( ne < 10 ) throw std::runtime_error( "compiled material: automatic SAB energy grid needs at least 10 points" );
          off_elut = lm.reserve( kKeyLutMaxEntries*sizeof(uint16_t) );
          T.elut = nullptr; T.elut_key0 = T.elut_shift = T.elut_nk = 0;
          T.egrid_log0 = 0.0; T.egrid_invdlog = 0.0;
        } else {
          T.elut  = offAsPtr<uint16_t>( buildKeyLut( lm, arr, ne, T.elut_key0, T.elut_shift, T.elut_nk ) );
        }
        T.alpha = offAsPtr<double>( lm.put( arr + 2*ne, na*8 ) );
        T.beta  = offAsPtr<double>( lm.put( arr + 2*ne + na, nb*8 ) );
        T.sab   = offAsPtr<double>( lm.put( arr + 2*ne + na + nb, na*nb*8 ) );
        SabBuildPlan pl;
        pl.sab_index = nsab;
        pl.auto_egrid = h.auto_egrid != 0; pl.suggested_emax = h.suggested_emax;
        if ( pl.auto_egrid ) { pl.req_emin = arr[0]; pl.req_emax = arr[1]; }
        pl.off_egrid = off_egrid; pl.off_xs = off_xs; pl.off_elut = off_elut;
        pl.off_logsab = lm.reserve( na*nb*8 );
        pl.off_cumul  = lm.reserve( na*nb*8 );
        pl.off_ep     = lm.reserve( ne*sizeof(SabEPoint) );
        T.bstride = sabBStride( (int)nb );
        const size_t bst = (size_t)T.bstride;
        pl.off_bx     = lm.reserve( ne*bst*8 );
        pl.off_bpdf   = lm.reserve( ne*bst*8 );
        pl.off_bcdf   = lm.reserve( ne*bst*8 );
        pl.off_ainfo  = lm.reserve( ne*nb*sizeof(SabAlphaInfo) );
        pl.off_rows   = lm.reserve( ne*nb*16 );
        pl.off_xscheck= lm.reserve( ne*8 + ne*4 );
        pl.off_bguide = lm.reserve( ne*(size_t)kSabGBStride*sizeof(uint16_t) );
        pl.off_aguide = lm.reserve( nb*( kSabGA+1 )*sizeof(uint16_t) );
        pl.off_ascale = lm.reserve( nb*8 );
        pl.off_heads  = lm.reserve( ne*nb*sizeof(SabHead) );
        pl.off_pts    = lm.reserve( na*nb*sizeof(SabPoint) );
        pl.off_tails  = lm.reserve( ne*nb*2*sizeof(SabTail) );
        pl.off_bpts   = lm.reserve( ne*bst*sizeof(SabBPoint) );
        pl.off_lguide = lm.reserve( nb*(size_t)kSabGLStride*sizeof(uint16_t) );
        T.logsab = offAsPtr<double>( pl.off_logsab );
        T.cumul  = offAsPtr<double>( pl.off_cumul );
        T.ep     = offAsPtr<SabEPoint>( pl.off_ep );
        T.bx     = offAsPtr<double>( pl.off_bx );
        T.bpdf   = offAsPtr<double>( pl.off_bpdf );
        T.bcdf   = offAsPtr<double>( pl.off_bcdf );
        T.ainfo  = offAsPtr<SabAlphaInfo>( pl.off_ainfo );
        T.bguide = offAsPtr<uint16_t>( pl.off_bguide );
        T.aguide = offAsPtr<uint16_t>( pl.off_aguide );
        T.ascale = offAsPtr<double>( pl.off_ascale );
        T.heads  = offAsPtr<SabHead>( pl.off_heads );
        T.pts    = offAsPtr<SabPoint>( pl.off_pts );
        T.tails  = offAsPtr<SabTail>( pl.off_tails );
        T.bpts   = offAsPtr<SabBPoint>( pl.off_bpts );
        T.lguide = offAsPtr<uint16_t>( pl.off_lguide );
        lm.sabplans.push_back( pl );
        k.idx = nsab++;
        break;
      }
      case NCB_KIND_SCBRAGG:
        loadScBragg( lm, blob, c );
        k.idx = 0;
        break;
      case NCB_KIND_LCBRAGG:
        loadLcBragg( lm, blob, c );
        k.idx = 0;
        break;
      default:
        throw std::runtime_error( "compiled material: unknown component kind" );
      }
    }
    lm.reserve( 0 );
  }

}
