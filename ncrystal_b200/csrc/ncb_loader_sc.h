// ncb_loader_sc.h -- SCBragg part of the blob loader (oriented path).
#pragma once
#include "ncb_loader.h"
namespace ncb {
  inline void loadScBragg( LoadedMaterial& lm, const unsigned char* blob, const ncb_comp_t& c )
  {
    checkPayload( c, sizeof(ncb_scbragg_t), {}, 0, "SCBragg" );
    ncb_scbragg_t h; std::memcpy( &h, blob + c.off, sizeof(h) );
    checkPayload( c, sizeof(h), { h.nfam, h.nnormals, h.lut_sofcosd_n, h.lut_evalcosx_n },
                  3*h.nfam + 1 + 3*h.nnormals + 2*h.lut_sofcosd_n + 2*h.lut_evalcosx_n, "SCBragg" );
    const double* arr = reinterpret_cast<const double*>( blob + c.off + sizeof(h) );
    ScBraggT& S = lm.mat.sc;
    if ( S.nfam != 0 )
      throw std::runtime_error( "compiled material: more than one SCBragg component" );
    const size_t nf = h.nfam, nn = h.nnormals;
    if ( nf == 0 || nn == 0 || h.lut_sofcosd_n < 4 || h.lut_evalcosx_n < 4 )
      throw std::runtime_error( "compiled material: degenerate SCBragg tables" );
    S.threshold_ekin = h.threshold_ekin;
    S.cta = h.gos_cta;
    S.sta = h.gos_sta;
    S.circleint_k1 = h.gos_circleint_k1; S.circleint_k2 = h.gos_circleint_k2;
    S.numint_accuracy = h.gos_numint_accuracy;
    S.nfam = (int)nf; S.nnormals = (int)nn;
    S.fam_xsfact = offAsPtr<double>( lm.put( arr, nf*8 ) );
    S.fam_inv2d = offAsPtr<double>( lm.put( arr + nf, nf*8 ) );
    std::vector<int> first( nf+1 );
    for ( size_t i = 0; i <= nf; ++i ) {
      const double v = arr[2*nf+i];
      if ( !( v >= 0.0 && v <= (double)nn ) || ( i && (int)v < first[i-1] ) )
        throw std::runtime_error( "compiled material: inconsistent SCBragg family index" );
      first[i] = (int)v;
    }
    if ( first[nf] != (int)nn )
      throw std::runtime_error( "compiled material: inconsistent SCBragg family index" );
    S.fam_first = offAsPtr<int>( lm.put( first.data(), (nf+1)*sizeof(int) ) );
    const double* pn = arr + 2*nf + nf + 1;
    S.normals = offAsPtr<double>( lm.put( pn, 3*nn*8 ) );
    {
      // single-precision copy for the pre-filter of k_sc_find: one 16-byte record per normal, (x, y, z, family index
      // as integer bits), padded to a multiple of 128 records with normals that can never pass (x=y=z=0)
      const size_t nnp = ( nn + 127 ) & ~(size_t)127;
      std::vector<float> nf32( 4*nnp, 0.0f );
      size_t f = 0;
      for ( size_t i = 0; i < nn; ++i ) {
        while ( f + 1 < nf && (int)i >= first[f+1] ) ++f;
        for ( int k = 0; k < 3; ++k ) nf32[4*i + k] = (float)pn[3*i+k];
        const uint32_t fi = (uint32_t)f;
        std::memcpy( &nf32[4*i + 3], &fi, 4 );
      }
      S.normals_f = offAsPtr<float>( lm.put( nf32.data(), nf32.size()*sizeof(float) ) );
    }
    const double* l1 = pn + 3*nn;
    const double* l2 = l1 + 2*h.lut_sofcosd_n;
    S.sofcosd.data = offAsPtr<double>( lm.put( l1, 2*h.lut_sofcosd_n*8 ) );
    S.sofcosd.nm2 = (int)h.lut_sofcosd_n - 2;
    S.sofcosd.a = h.sofcosd_a; S.sofcosd.invdelta = h.sofcosd_invdelta;
    S.evalcosx.data = offAsPtr<double>( lm.put( l2, 2*h.lut_evalcosx_n*8 ) );
    S.evalcosx.nm2 = (int)h.lut_evalcosx_n - 2;
    S.evalcosx.a = h.evalcosx_a; S.evalcosx.invdelta = h.evalcosx_invdelta;
  }

  // LCBragg: plane sets + the mosaicity tables of its GaussMos (stored in Material::sc, which then has no families)
  inline void loadLcBragg( LoadedMaterial& lm, const unsigned char* blob, const ncb_comp_t& c )
  {
    checkPayload( c, sizeof(ncb_lcbragg_t), {}, 0, "LCBragg" );
    ncb_lcbragg_t h; std::memcpy( &h, blob + c.off, sizeof(h) );
    checkPayload( c, sizeof(h), { h.nplanesets, h.lut_sofcosd_n, h.lut_evalcosx_n },
                  kLcPlaneStride*h.nplanesets + 2*h.lut_sofcosd_n + 2*h.lut_evalcosx_n, "LCBragg" );
    const double* arr = reinterpret_cast<const double*>( blob + c.off + sizeof(h) );
    ScBraggT& S = lm.mat.sc;
    LcBraggT& L = lm.mat.lc;
    if ( S.nfam != 0 || L.nplanes != 0 )
      throw std::runtime_error( "compiled material: more than one SCBragg/LCBragg component" );
    if ( h.nplanesets == 0 || h.nplanesets > 32767 || h.lut_sofcosd_n < 4 || h.lut_evalcosx_n < 4 )
      throw std::runtime_error( "compiled material: degenerate LCBragg tables" );
    const size_t np = h.nplanesets;
    for ( size_t i = 0; i < np; ++i ) {
      const double* p = arr + kLcPlaneStride*i;
      if ( !( p[0] > 0.0 ) || !( p[1] > 0.0 ) || ( i && p[0] > arr[kLcPlaneStride*(i-1)] ) )
        throw std::runtime_error( "compiled material: LCBragg plane sets not sorted by d-spacing" );
    }
    L.ekin_low = h.ekin_low;
    L.xsfact = h.xsfact;
    L.acc = std::min( std::max( h.gos_prec, 1e-7 ), 1e-2 );
    L.ax = h.lcaxis_lab[0]; L.ay = h.lcaxis_lab[1]; L.az = h.lcaxis_lab[2];
    L.nplanes = (int)np;
    L.planes = offAsPtr<double>( lm.put( arr, kLcPlaneStride*np*8 ) );
    S.threshold_ekin = h.ekin_low;
    S.cta = h.gos_cta;
    S.sta = h.gos_sta;
    S.circleint_k1 = h.gos_circleint_k1; S.circleint_k2 = h.gos_circleint_k2;
    S.numint_accuracy = h.gos_numint_accuracy;
    S.nfam = 0; S.nnormals = 0;
    const double* l1 = arr + kLcPlaneStride*np;
    const double* l2 = l1 + 2*h.lut_sofcosd_n;
    S.sofcosd.data = offAsPtr<double>( lm.put( l1, 2*h.lut_sofcosd_n*8 ) );
    S.sofcosd.nm2 = (int)h.lut_sofcosd_n - 2;
    S.sofcosd.a = h.sofcosd_a; S.sofcosd.invdelta = h.sofcosd_invdelta;
    S.evalcosx.data = offAsPtr<double>( lm.put( l2, 2*h.lut_evalcosx_n*8 ) );
    S.evalcosx.nm2 = (int)h.lut_evalcosx_n - 2;
    S.evalcosx.a = h.evalcosx_a; S.evalcosx.invdelta = h.evalcosx_invdelta;
  }
}
