// oracle/refdrv.cc -- TEST INFRASTRUCTURE.  Links the UNMODIFIED reference library
// (oracle/_ref/lib/libNCrystal.so).  Never linked or loaded by the product (ncrystal_b200/).
//
// Built on top of the reference-side binding (the material compiler, bridge/matcompile_impl.icc,
// whose handle it shares) it adds what only the tests and the bench need:
//  (2) replay oracle: evaluate crossSection / sampleScatter of the reference with
//      the per-neutron Philox streams of philox_ref.h plugged in through the
//      reference's own RNG interface, so device outputs can be compared 1:1.
//  (3) CPU baseline: time the reference's C-API *_many calls on all host cores
//      (one cloned handle per thread, ncrystal.h:711-731).
//
// Build: oracle/Makefile target "tools" (g++ -fno-access-control).

#include "../bridge/matcompile_impl.icc"
#include "philox_ref.h"
#include "NCrystal/internal/vdos/NCVDOSEval.hh"
#include "NCrystal/internal/vdos/NCVDOSGn.hh"
#include "NCrystal/internal/vdos/NCVDOSToScatKnl.hh"
#include "NCrystal/internal/sab/NCSABUtils.hh"
#include <algorithm>

namespace {
  class PhiloxStream final : public NC::RNGStream {
    // Replays the device's per-neutron stream through the reference's RNG hook.
    // coinflip() etc. keep the RNGStream defaults (coinflip = generate()>0.5,
    // NCRNG.cc:35-38), as for any non-builtin stream.
  public:
    ncb_stream_t st;
    void reset( uint64_t seed, uint64_t index ) { ncb_stream_init(&st,seed,index); }
  protected:
    double actualGenerate() override { return ncb_stream_next(&st); }
  };

}

extern "C" {

  // ---- (2) replay oracle -----------------------------------------------------
  // Cross sections through the ProcImpl interface (fresh cache each call; the
  // CachePtr is a pure CPU optimisation, NCProcImpl.hh:50-64).
  void refdrv_xs_iso( void* vh, const double* ekin, uint64_t n, double* out )
  {
    auto h = static_cast<Handle*>(vh);
    NC::CachePtr cp;
    for ( uint64_t i = 0; i < n; ++i )
      out[i] = h->proc->crossSectionIsotropic( cp, NC::NeutronEnergy{ekin[i]} ).dbl();
  }

  // per-component *unscaled* xs: out[c*n+i]
  void refdrv_xs_iso_components( void* vh, const double* ekin, uint64_t n, double* out )
  {
    auto h = static_cast<Handle*>(vh);
    for ( size_t c = 0; c < h->leaves.size(); ++c ) {
      NC::CachePtr cp;
      auto& p = *h->leaves[c].proc;
      auto dom = p.domain();
      for ( uint64_t i = 0; i < n; ++i ) {
        NC::NeutronEnergy e{ekin[i]};
        out[c*n+i] = dom.contains(e) ? p.crossSectionIsotropic( cp, e ).dbl() : 0.0;
      }
    }
  }

  void refdrv_xs( void* vh, const double* ekin, const double* ux, const double* uy, const double* uz,
                  uint64_t n, double* out )
  {
    auto h = static_cast<Handle*>(vh);
    NC::CachePtr cp;
    for ( uint64_t i = 0; i < n; ++i )
      out[i] = h->proc->crossSection( cp, NC::NeutronEnergy{ekin[i]}, NC::NeutronDirection{ux[i],uy[i],uz[i]} ).dbl();
  }

  // Neutron i consumes the stream (seed, first_index+i).  ndraws (optional) receives
  // the number of uniforms consumed.
  void refdrv_sample_iso( void* vh, uint64_t seed, uint64_t first_index, const double* ekin, uint64_t n,
                          double* ekin_out, double* mu_out, uint32_t* ndraws )
  {
    auto h = static_cast<Handle*>(vh);
    PhiloxStream rng;
    NC::CachePtr cp;
    for ( uint64_t i = 0; i < n; ++i ) {
      rng.reset( seed, first_index + i );
      try {
        auto o = h->proc->sampleScatterIsotropic( cp, rng, NC::NeutronEnergy{ekin[i]} );
        ekin_out[i] = o.ekin.dbl();
        mu_out[i] = o.mu.dbl();
      } catch ( std::exception& e ) {
        g_err = e.what();
        ekin_out[i] = -1.0; mu_out[i] = -999.0;
        cp = nullptr;
      }
      if (ndraws) ndraws[i] = rng.st.ndraws;
    }
  }

  void refdrv_sample( void* vh, uint64_t seed, uint64_t first_index, const double* ekin,
                      const double* ux, const double* uy, const double* uz, uint64_t n,
                      double* ekin_out, double* ox, double* oy, double* oz, uint32_t* ndraws )
  {
    auto h = static_cast<Handle*>(vh);
    PhiloxStream rng;
    NC::CachePtr cp;
    for ( uint64_t i = 0; i < n; ++i ) {
      rng.reset( seed, first_index + i );
      try {
        auto o = h->proc->sampleScatter( cp, rng, NC::NeutronEnergy{ekin[i]}, NC::NeutronDirection{ux[i],uy[i],uz[i]} );
        ekin_out[i] = o.ekin.dbl();
        ox[i] = o.direction[0]; oy[i] = o.direction[1]; oz[i] = o.direction[2];
      } catch ( std::exception& e ) {
        g_err = e.what();
        ekin_out[i] = -1.0; ox[i] = oy[i] = oz[i] = 0.0;
        cp = nullptr;
      }
      if (ndraws) ndraws[i] = rng.st.ndraws;
    }
  }

  // sample with a single chosen leaf (component index c), for leaf-level parity tests
  void refdrv_sample_iso_leaf( void* vh, int c, uint64_t seed, uint64_t first_index, const double* ekin, uint64_t n,
                               double* ekin_out, double* mu_out, uint32_t* ndraws )
  {
    auto h = static_cast<Handle*>(vh);
    auto& p = *h->leaves.at(c).proc;
    PhiloxStream rng;
    NC::CachePtr cp;
    for ( uint64_t i = 0; i < n; ++i ) {
      rng.reset( seed, first_index + i );
      try {
        auto o = p.sampleScatterIsotropic( cp, rng, NC::NeutronEnergy{ekin[i]} );
        ekin_out[i] = o.ekin.dbl();
        mu_out[i] = o.mu.dbl();
      } catch ( std::exception& e ) {
        g_err = e.what();
        ekin_out[i] = -1.0; mu_out[i] = -999.0;
        cp = nullptr;
      }
      if (ndraws) ndraws[i] = rng.st.ndraws;
    }
  }

  // ---- dumps of reference-internal SAB sampler tables (test-only cross-check of
  //      the product's native table builder, csrc/sab_build.cpp) ----------------
  // Returns number of beta-sampler points of energy point iE of SAB component c
  // (0 => NoScatter sampler), fills (if non-null) x/pdf/cdf [npts], infos [10*(npts-1)]
  // as {front.alpha,front.sval,front.logsval,front.idx,back.alpha,back.sval,back.logsval,back.idx,prob_front,prob_notback},
  // meta = {ibetaOffset, firstBinKinematicEndpointValue}.
  int refdrv_sab_sampler_dump( void* vh, int c, int iE, double* x, double* pdf, double* cdf, double* infos, double* meta )
  {
    auto h = static_cast<Handle*>(vh);
    auto sab = dynamic_cast<const NC::SABScatter*>( h->leaves.at(c).proc.get() );
    if (!sab) return -1;
    auto& smp = sab->m_sh->sampler;
    if ( iE < 0 || iE >= (int)smp.m_samplers.size() ) return -1;
    auto a = dynamic_cast<const NC::SAB::SABSamplerAtE_Alg1*>( smp.m_samplers[iE].get() );
    if (!a) return 0;
    const auto& xs = a->m_betaSampler.getXVals();
    int n = (int)xs.size();
    if (x) for (int i=0;i<n;++i) x[i] = xs[i];
    if (pdf) for (int i=0;i<n;++i) pdf[i] = a->m_betaSampler.m_y[i];
    if (cdf) for (int i=0;i<n;++i) cdf[i] = a->m_betaSampler.m_cdf[i];
    if (infos) {
      for (int i=0;i<n-1;++i) {
        auto& f = a->m_alphaSamplerInfos[i];
        double* o = infos + 10*i;
        o[0]=f.pt_front.alpha; o[1]=f.pt_front.sval; o[2]=f.pt_front.logsval; o[3]=f.pt_front.alpha_idx;
        o[4]=f.pt_back.alpha;  o[5]=f.pt_back.sval;  o[6]=f.pt_back.logsval;  o[7]=f.pt_back.alpha_idx;
        o[8]=f.prob_front; o[9]=f.prob_notback;
      }
    }
    if (meta) { meta[0] = (double)a->m_ibetaOffset; meta[1] = a->m_firstBinKinematicEndpointValue; }
    return n;
  }

  // The reference's SABIntegrator run on a leaf's scattering kernel with a FULLY automatic energy grid (no "egrid"
  // request; SABData::suggestedEmax still applies): what ncb_sabgrid.h restates.  egrid/xs: npts values each (300).
  int refdrv_sab_auto_egrid( void* vh, int c, double* egrid, double* xs, int cap )
  {
    auto h = static_cast<Handle*>(vh);
    auto sab = dynamic_cast<const NC::SABScatter*>( h->leaves.at(c).proc.get() );
    if (!sab) return -1;
    try {
      auto alg1 = firstAlg1( sab->m_sh->sampler );
      if (!alg1) return -1;
      NC::shared_obj<const NC::SABData> data = alg1->m_common->data;
      NC::SAB::SABIntegrator si( data );
      NC::SABXSProvider xp = si.createXSProvider();
      const int n = (int)xp.m_egrid.size();
      if ( n > cap ) return -2;
      for ( int i = 0; i < n; ++i ) { egrid[i] = xp.m_egrid[i]; xs[i] = xp.m_xs[i]; }
      return n;
    } catch ( std::exception& e ) {
      g_err = e.what();
      return -3;
    }
  }

  // ---- VDOS -> S(alpha,beta) expansion of the reference, stage by stage (what csrc/ncb_vdos.h restates) -----
  // The k'th DI_VDOS entry of the material's dynamic-info list.
  static const NC::DI_VDOS* vdosEntry( const NC::Info& info, int k )
  {
    int i = 0;
    for ( auto& di : info.getDynamicInfoList() )
      if ( auto v = dynamic_cast<const NC::DI_VDOS*>( di.get() ) ) { if ( i++ == k ) return v; }
    return nullptr;
  }
  // meta = { emin, emax, temperature, mass_amu, bound_xs }; returns the number of density points (or -1)
  int refdrv_vdos_data( const char* cfg, int k, double* meta5, double* density, int cap )
  {
    try {
      auto info = NC::createInfo( cfg );
      auto v = vdosEntry( *info, k );
      if ( !v ) return -1;
      const NC::VDOSData& vd = v->vdosData();
      meta5[0] = vd.vdos_egrid().first; meta5[1] = vd.vdos_egrid().second; meta5[2] = vd.temperature().dbl();
      meta5[3] = vd.elementMassAMU().dbl(); meta5[4] = vd.boundXS().dbl();
      const int n = (int)vd.vdos_density().size();
      if ( n > cap ) return -2;
      for ( int i = 0; i < n; ++i ) density[i] = vd.vdos_density()[i];
      return n;
    } catch ( std::exception& e ) { g_err = e.what(); return -3; }
  }
  // createScatteringKernel + transformKernelToStdFormat.  meta = { suggestedEmax, gamma0, msd, max order reached }.
  int refdrv_vdos_expand( const char* cfg, int k, int vdoslux, double* alpha, int* nalpha, double* beta, int* nbeta,
                          double* sab, int cap_sab, double* meta4 )
  {
    try {
      auto info = NC::createInfo( cfg );
      auto v = vdosEntry( *info, k );
      if ( !v ) return -1;
      const NC::VDOSData& vd = v->vdosData();
      NC::VDOSEval ve( vd );
      meta4[1] = ve.calcGamma0(); meta4[2] = ve.getMSD( meta4[1] );
      auto sd = NC::SABUtils::transformKernelToStdFormat( NC::createScatteringKernel( vd, (unsigned)vdoslux ) );
      meta4[0] = sd.suggestedEmax(); meta4[3] = 0.0;
      const int na = (int)sd.alphaGrid().size(), nb = (int)sd.betaGrid().size();
      if ( na*nb > cap_sab ) return -2;
      *nalpha = na; *nbeta = nb;
      for ( int i = 0; i < na; ++i ) alpha[i] = sd.alphaGrid()[i];
      for ( int i = 0; i < nb; ++i ) beta[i] = sd.betaGrid()[i];
      for ( int i = 0; i < na*nb; ++i ) sab[i] = sd.sab()[i];
      return 0;
    } catch ( std::exception& e ) { g_err = e.what(); return -3; }
  }
  // G_n spectrum of the reference (default truncation/thinning).  meta = { lower edge, bin width, max density }
  int refdrv_vdos_gn( const char* cfg, int k, int order, double* spec, int cap, double* meta3 )
  {
    try {
      auto info = NC::createInfo( cfg );
      auto v = vdosEntry( *info, k );
      if ( !v ) return -1;
      NC::VDOSEval ve( v->vdosData() );
      NC::VDOSGn gn( ve );
      gn.growMaxOrder( (unsigned)order );
      const auto& sp = gn.getRawSpectrum( (unsigned)order );
      if ( (int)sp.size() > cap ) return -2;
      for ( size_t i = 0; i < sp.size(); ++i ) spec[i] = sp[i];
      meta3[0] = gn.eRange( (unsigned)order ).first; meta3[1] = gn.binWidth( (unsigned)order );
      meta3[2] = *std::max_element( sp.begin(), sp.end() );
      return (int)sp.size();
    } catch ( std::exception& e ) { g_err = e.what(); return -3; }
  }

  // ---- (3) CPU baseline through the reference's own C-API --------------------
  // mode 0: ncrystal_crosssection_nonoriented_many; mode 1: ncrystal_samplescatterisotropic_many;
  // mode 2: per-neutron ncrystal_crosssection; mode 3: per-neutron ncrystal_samplescatter.
  // Returns seconds of the best of `nrep` timed passes after one warm-up pass.
  double refdrv_bench_capi( const char* cfg, int mode, int nthreads, int nrep,
                            const double* ekin, const double* ux, const double* uy, const double* uz,
                            uint64_t n, double* out0, double* out1, double* out2, double* out3 )
  {
    ncrystal_scatter_t sc0 = ncrystal_create_scatter_builtinrng( cfg, 12345 );
    std::vector<ncrystal_scatter_t> sc( nthreads );
    sc[0] = sc0;
    for ( int t = 1; t < nthreads; ++t )
      sc[t] = ncrystal_clone_scatter( sc0 );
    auto work = [&]( int t ) {
      uint64_t b = n*(uint64_t)t/nthreads, e = n*(uint64_t)(t+1)/nthreads;
      uint64_t m = e-b;
      if (!m) return;
      ncrystal_process_t pr = ncrystal_cast_scat2proc( sc[t] );
      if ( mode == 0 ) {
        ncrystal_crosssection_nonoriented_many( pr, ekin+b, m, 1, out0+b );
      } else if ( mode == 1 ) {
        ncrystal_samplescatterisotropic_many( sc[t], ekin+b, m, 1, out0+b, out1+b );
      } else if ( mode == 2 ) {
        for ( uint64_t i = b; i < e; ++i ) {
          double d[3] = { ux[i], uy[i], uz[i] };
          ncrystal_crosssection( pr, ekin[i], &d, out0+i );
        }
      } else {
        for ( uint64_t i = b; i < e; ++i ) {
          double d[3] = { ux[i], uy[i], uz[i] };
          double o[3];
          ncrystal_samplescatter( sc[t], ekin[i], &d, out0+i, &o );
          out1[i] = o[0]; out2[i] = o[1]; out3[i] = o[2];
        }
      }
    };
    double best = 1e99;
    for ( int rep = -1; rep < nrep; ++rep ) {
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> th;
      for ( int t = 1; t < nthreads; ++t )
        th.emplace_back( work, t );
      work(0);
      for ( auto& x : th ) x.join();
      double dt = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
      if ( rep >= 0 && dt < best ) best = dt;
    }
    for ( int t = 0; t < nthreads; ++t )
      ncrystal_unref( &sc[t] );
    return best;
  }
}
