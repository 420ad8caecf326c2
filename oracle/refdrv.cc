// oracle/refdrv.cc -- TEST INFRASTRUCTURE + reference-side binding.  Links the
// UNMODIFIED reference library (oracle/_ref/lib/libNCrystal.so).  Never linked or
// loaded by the product (ncrystal_b200/).
//
// Three jobs:
//  (1) "material compiler": walk the reference's ProcComposition for a cfg string
//      and flatten the immutable leaf tables into the POD blob of
//      ncrystal_b200/csrc/ncb_blob.h.  This is the reference-side binding a
//      maintainer would add next to the C-API; private
//      members are read with -fno-access-control instead of patched-in accessors.
//  (2) replay oracle: evaluate crossSection / sampleScatter of the reference with
//      the per-neutron Philox streams of philox_ref.h plugged in through the
//      reference's own RNG interface, so device outputs can be compared 1:1.
//  (3) CPU baseline: time the reference's C-API *_many calls on all host cores
//      (one cloned handle per thread, ncrystal.h:711-731).
//
// Build: oracle/Makefile target "tools" (g++ -fno-access-control).

#include "NCrystal/NCrystal.hh"
#include "NCrystal/ncrystal.h"
#include "NCrystal/internal/powderbragg/NCPowderBragg.hh"
#include "NCrystal/internal/elincscatter/NCElIncScatter.hh"
#include "NCrystal/internal/phys_utils/NCElIncXS.hh"
#include "NCrystal/internal/sabscatter/NCSABScatter.hh"
#include "NCrystal/internal/sab/NCSABScatterHelper.hh"
#include "NCrystal/internal/sab/NCSABSamplerModels.hh"
#include "NCrystal/internal/sab/NCSABExtender.hh"
#include "NCrystal/internal/sab/NCSABIntegrator.hh"
#include "NCrystal/internal/freegas/NCFreeGas.hh"
#include "NCrystal/internal/phys_utils/NCFreeGasUtils.hh"
#include "NCrystal/internal/utils/NCPointwiseDist.hh"

#include "ncb_blob.h"
#include "philox_ref.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace NC = NCrystal;
namespace NCPI = NCrystal::ProcImpl;

// FreeGas keeps its state behind a pimpl whose definition lives in NCFreeGas.cc:28-41.
// Restating the identical definition here makes the members reachable.
struct NC::FreeGas::Impl {
  Impl( Temperature t, AtomMass target_mass_amu, SigmaFree sigma )
    : m_xsprovider(t, target_mass_amu, sigma),
      m_temperature(DoValidate,t),
      m_target_mass_amu(DoValidate,target_mass_amu) {}
  FreeGasXSProvider m_xsprovider;
  Temperature m_temperature;
  AtomMass m_target_mass_amu;
};

#include "refdrv_scbragg.icc"

namespace {

  struct Leaf { double scale; NCPI::ProcPtr proc; };

  std::vector<Leaf> flatten( const NCPI::ProcPtr& top )
  {
    std::vector<Leaf> out;
    auto pc = dynamic_cast<const NCPI::ProcComposition*>( top.get() );
    if ( pc ) {
      for ( auto& c : pc->components() )
        out.push_back( { c.scale, c.process } );
    } else {
      out.push_back( { 1.0, top } );
    }
    return out;
  }

  struct Buf {
    std::vector<unsigned char> d;
    uint64_t reserve( uint64_t n ) { uint64_t off = ncb_align16(d.size()); d.resize(off+n,0); return off; }
    void put( uint64_t off, const void* p, uint64_t n ) { std::memcpy(&d[off],p,n); }
    uint64_t append( const void* p, uint64_t n ) { uint64_t off = d.size(); d.resize(off+n); std::memcpy(&d[off],p,n); return off; }
    uint64_t appendv( const std::vector<double>& v ) { return append(v.data(),v.size()*sizeof(double)); }
  };

  bool refdrv_compile_scbragg( const NCPI::Process* p, ncb_comp_t& comp, Buf& buf, std::string& err )
  {
    auto sc = dynamic_cast<const NC::SCBragg*>( p );
    if ( !sc )
      return false;
    static_assert( sizeof(SCBraggMirrorFamily) == sizeof(std::vector<NC::Vector>) + 2*sizeof(double), "layout" );
    auto pm = reinterpret_cast<const SCBraggMirrorPimpl*>( sc->m_pimpl.get() );
    const NC::GaussMos& gm = pm->m_gm;
    const NC::GaussOnSphere& gos = gm.m_gos;
    if ( pm->m_threshold_ekin != p->domain().elow.dbl() ) { err = "SCBragg pimpl mirror mismatch"; return true; }
    comp.kind = NCB_KIND_SCBRAGG;
    ncb_scbragg_t h; std::memset(&h,0,sizeof(h));
    h.threshold_ekin = pm->m_threshold_ekin;
    h.gos_cta = gos.m_cta; h.gos_sta = gos.m_sta;
    h.gos_circleint_k1 = gos.m_circleint_k1; h.gos_circleint_k2 = gos.m_circleint_k2;
    h.gos_norm = gos.m_norm; h.gos_expfact = gos.m_expfact; h.gos_truncangle = gos.m_truncangle; h.gos_sigma = gos.m_sigma;
    h.gos_numint_accuracy = gos.m_numint_accuracy;
    h.gos_prec = gos.m_prec;
    h.sofcosd_a = gos.m_lt_sofcosd.m_a; h.sofcosd_invdelta = gos.m_lt_sofcosd.m_invdelta;
    h.evalcosx_a = gos.m_lt_evalcosx.m_a; h.evalcosx_invdelta = gos.m_lt_evalcosx.m_invdelta;
    h.mos_fwhm = gm.m_mos_fwhm.dbl(); h.mos_truncN = gm.m_mos_truncN;
    h.nfam = pm->m_reflfamilies.size();
    std::vector<double> xsfact, inv2d, first, normals;
    uint64_t nn = 0;
    for ( auto& f : pm->m_reflfamilies ) {
      xsfact.push_back( f.xsfact ); inv2d.push_back( f.inv2d ); first.push_back( (double)nn );
      for ( auto& v : f.deminormals ) { normals.push_back(v.x()); normals.push_back(v.y()); normals.push_back(v.z()); ++nn; }
    }
    first.push_back( (double)nn );
    h.nnormals = nn;
    auto lutdata = []( const NC::SplinedLookupTable& L ) {
      std::vector<double> d;
      for ( auto& e : L.m_spline.m_data ) { d.push_back(e.first); d.push_back(e.second); }
      return d;
    };
    auto d1 = lutdata( gos.m_lt_sofcosd ), d2 = lutdata( gos.m_lt_evalcosx );
    h.lut_sofcosd_n = d1.size()/2; h.lut_evalcosx_n = d2.size()/2;
    if ( gos.m_lt_sofcosd.m_spline.m_nm2 + 2 != h.lut_sofcosd_n || gos.m_lt_evalcosx.m_spline.m_nm2 + 2 != h.lut_evalcosx_n ) {
      err = "unexpected spline layout"; return true;
    }
    comp.off = buf.reserve(sizeof(h));
    buf.put(comp.off,&h,sizeof(h));
    buf.appendv(xsfact); buf.appendv(inv2d); buf.appendv(first); buf.appendv(normals);
    buf.appendv(d1); buf.appendv(d2);
    return true;
  }

  bool refdrv_compile_lcbragg( const NCPI::Process* p, ncb_comp_t& comp, Buf& buf, std::string& err )
  {
    auto lc = dynamic_cast<const NC::LCBragg*>( p );
    if ( !lc )
      return false;
    auto pm = reinterpret_cast<const LCBraggMirrorPimpl*>( lc->m_pimpl.get() );
    if ( pm->m_ekin_low != p->domain().elow.dbl() ) { err = "LCBragg pimpl mirror mismatch"; return true; }
    if ( !pm->m_lchelper || pm->m_scmodel != nullptr ) { err = "LCBragg with lcmode!=0 (reference models built on SCBragg) is not supported"; return true; }
    const NC::LCHelper& H = *pm->m_lchelper;
    const NC::GaussMos& gm = H.m_lcstdframe.m_gm;
    const NC::GaussOnSphere& gos = gm.m_gos;
    comp.kind = NCB_KIND_LCBRAGG;
    ncb_lcbragg_t h; std::memset(&h,0,sizeof(h));
    h.ekin_low = pm->m_ekin_low;
    h.lcaxis_lab[0] = H.m_lcaxislab.x(); h.lcaxis_lab[1] = H.m_lcaxislab.y(); h.lcaxis_lab[2] = H.m_lcaxislab.z();
    h.xsfact = H.m_xsfact;
    h.gos_cta = gos.m_cta; h.gos_sta = gos.m_sta;
    h.gos_circleint_k1 = gos.m_circleint_k1; h.gos_circleint_k2 = gos.m_circleint_k2;
    h.gos_numint_accuracy = gos.m_numint_accuracy;
    h.gos_prec = gm.precision();
    h.gos_truncangle = gos.m_truncangle;
    h.sofcosd_a = gos.m_lt_sofcosd.m_a; h.sofcosd_invdelta = gos.m_lt_sofcosd.m_invdelta;
    h.evalcosx_a = gos.m_lt_evalcosx.m_a; h.evalcosx_invdelta = gos.m_lt_evalcosx.m_invdelta;
    h.mos_fwhm = gm.m_mos_fwhm.dbl();
    h.nplanesets = H.m_planes.size();
    std::vector<double> ps;
    for ( auto& e : H.m_planes ) {
      ps.push_back( e.twodsp ); ps.push_back( e.inv_twodsp ); ps.push_back( e.cosalpha ); ps.push_back( e.sinalpha );
      ps.push_back( e.cosalphaminus ); ps.push_back( e.cosalphaplus ); ps.push_back( e.fsq );
    }
    auto lutdata = []( const NC::SplinedLookupTable& L ) {
      std::vector<double> d;
      for ( auto& e : L.m_spline.m_data ) { d.push_back(e.first); d.push_back(e.second); }
      return d;
    };
    auto d1 = lutdata( gos.m_lt_sofcosd ), d2 = lutdata( gos.m_lt_evalcosx );
    h.lut_sofcosd_n = d1.size()/2; h.lut_evalcosx_n = d2.size()/2;
    comp.off = buf.reserve(sizeof(h));
    buf.put(comp.off,&h,sizeof(h));
    buf.appendv(ps); buf.appendv(d1); buf.appendv(d2);
    return true;
  }

  const NC::SAB::SABSamplerAtE_Alg1* firstAlg1( const NC::SABSampler& s )
  {
    for ( auto& up : s.m_samplers ) {
      auto p = dynamic_cast<const NC::SAB::SABSamplerAtE_Alg1*>(up.get());
      if (p) return p;
    }
    return nullptr;
  }

  bool compileLeaf( const Leaf& leaf, ncb_comp_t& comp, Buf& buf, std::string& err )
  {
    const NCPI::Process* p = leaf.proc.get();
    comp.scale = leaf.scale;
    auto dom = p->domain();
    comp.dom_lo = dom.elow.dbl();
    comp.dom_hi = dom.ehigh.dbl();
    if ( auto pb = dynamic_cast<const NC::PowderBragg*>(p) ) {
      comp.kind = NCB_KIND_POWDERBRAGG;
      ncb_powderbragg_t h; std::memset(&h,0,sizeof(h));
      h.nplanes = pb->m_2dE.size();
      h.threshold = pb->m_threshold.dbl();
      comp.off = buf.reserve(sizeof(h));
      buf.put(comp.off,&h,sizeof(h));
      buf.appendv(pb->m_2dE);
      buf.appendv(pb->m_fdm_commul);
    } else if ( auto ei = dynamic_cast<const NC::ElIncScatter*>(p) ) {
      comp.kind = NCB_KIND_ELINC;
      ncb_elinc_t h; std::memset(&h,0,sizeof(h));
      auto& ed = ei->m_elincxs->m_elm_data;
      h.nelem = ed.size();
      comp.off = buf.reserve(sizeof(h));
      buf.put(comp.off,&h,sizeof(h));
      std::vector<double> msd, bixs;
      for ( auto& e : ed ) { msd.push_back(e.first); bixs.push_back(e.second); }
      buf.appendv(msd);
      buf.appendv(bixs);
    } else if ( auto fg = dynamic_cast<const NC::FreeGas*>(p) ) {
      comp.kind = NCB_KIND_FREEGAS;
      ncb_freegas_t h; std::memset(&h,0,sizeof(h));
      h.sigma_free = fg->m_impl->m_xsprovider.m_sigmaFree;
      h.ca = fg->m_impl->m_xsprovider.m_ca;
      h.temperature = fg->m_impl->m_temperature.dbl();
      h.mass_amu = fg->m_impl->m_target_mass_amu.dbl();
      comp.off = buf.reserve(sizeof(h));
      buf.put(comp.off,&h,sizeof(h));
    } else if ( auto sab = dynamic_cast<const NC::SABScatter*>(p) ) {
      comp.kind = NCB_KIND_SAB;
      const auto& sh = *sab->m_sh;
      auto alg1 = firstAlg1( sh.sampler );
      if (!alg1) { err = "SABScatter without Alg1 samplers"; return false; }
      const NC::SABData& sd = *alg1->m_common->data;
      auto ext = dynamic_cast<const NC::SAB::SABFGExtender*>( sh.xsprovider.m_extender.get() );
      if (!ext) { err = "SABScatter with unsupported extender type"; return false; }
      if ( sh.sampler.m_egrid != sh.xsprovider.m_egrid ) { err = "sampler/xsprovider egrid mismatch"; return false; }
      ncb_sab_t h; std::memset(&h,0,sizeof(h));
      h.scale = sab->m_scale;
      h.temperature = sd.temperature().dbl();
      h.mass_amu = sd.elementMassAMU().dbl();
      h.bound_xs = sd.boundXS().dbl();
      h.suggested_emax = sd.suggestedEmax();
      h.ext_sigma_free = ext->m_xsprovider.m_sigmaFree;
      h.ext_ca = ext->m_xsprovider.m_ca;
      h.ext_temperature = ext->m_t.dbl();
      h.ext_mass_amu = ext->m_m.dbl();
      h.k_extension = sh.xsprovider.m_kExtension;
      h.xs_at_emax = sh.sampler.m_xsAtEmax;
      h.k1 = sh.sampler.m_k1;
      h.k2 = sh.sampler.m_k2;
      h.egrid_margin = sh.sampler.m_egridMargin.value;
      h.negrid = sh.xsprovider.m_egrid.size();
      h.nalpha = sd.alphaGrid().size();
      h.nbeta = sd.betaGrid().size();
      comp.off = buf.reserve(sizeof(h));
      buf.put(comp.off,&h,sizeof(h));
      buf.appendv(sh.xsprovider.m_egrid);
      buf.appendv(sh.xsprovider.m_xs);
      buf.appendv(sd.alphaGrid());
      buf.appendv(sd.betaGrid());
      buf.appendv(sd.sab());
    } else if ( refdrv_compile_scbragg( p, comp, buf, err ) ) {
      if ( !err.empty() )
        return false;
    } else if ( refdrv_compile_lcbragg( p, comp, buf, err ) ) {
      if ( !err.empty() )
        return false;
    } else {
      if (err.empty())
        err = std::string("unsupported leaf process type: ")+p->name();
      return false;
    }
    comp.nbytes = buf.d.size() - comp.off;
    return true;
  }

  struct Handle {
    Handle( NCPI::ProcPtr p, const char* c ) : proc(std::move(p)), leaves(flatten(proc)), cfg(c) {}
    NCPI::ProcPtr proc;
    std::vector<Leaf> leaves;
    std::string cfg;
  };

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

  thread_local std::string g_err;
}

extern "C" {

  const char* refdrv_lasterror() { return g_err.c_str(); }

  void* refdrv_create( const char* cfg )
  {
    try {
      auto sc = NC::createScatter( cfg );
      return new Handle( sc.underlyingPtr(), cfg );
    } catch ( std::exception& e ) {
      g_err = e.what();
      return nullptr;
    }
  }

  void refdrv_destroy( void* vh ) { delete static_cast<Handle*>(vh); }

  int refdrv_ncomp( void* vh ) { return (int)static_cast<Handle*>(vh)->leaves.size(); }
  const char* refdrv_compname( void* vh, int i ) { return static_cast<Handle*>(vh)->leaves.at(i).proc->name(); }
  double refdrv_compscale( void* vh, int i ) { return static_cast<Handle*>(vh)->leaves.at(i).scale; }
  int refdrv_isoriented( void* vh ) { return static_cast<Handle*>(vh)->proc->materialType() == NC::MaterialType::Anisotropic; }

  // ---- (1) material compiler -------------------------------------------------
  // Returns malloc'd blob (free with refdrv_free) or NULL.
  void* refdrv_compile( void* vh, uint64_t* nbytes )
  {
    auto h = static_cast<Handle*>(vh);
    try {
      Buf buf;
      ncb_header_t hdr; std::memset(&hdr,0,sizeof(hdr));
      hdr.magic = NCB_MAGIC;
      hdr.version = NCB_VERSION;
      hdr.ncomp = (uint32_t)h->leaves.size();
      if ( hdr.ncomp > NCB_MAXCOMP ) { g_err = "too many components"; return nullptr; }
      hdr.oriented = h->proc->materialType() == NC::MaterialType::Anisotropic ? 1 : 0;
      auto dom = h->proc->domain();
      hdr.dom_lo = dom.elow.dbl();
      hdr.dom_hi = dom.ehigh.dbl();
      std::snprintf( hdr.cfg, sizeof(hdr.cfg), "%s", h->cfg.c_str() );
      {
        // bulk quantities for the transport step (what MiniMC's MatDef holds next to the scatter process)
        auto info = NC::createInfo( h->cfg.c_str() );
        hdr.numdens = info->getNumberDensity().dbl();
        hdr.temperature = info->hasTemperature() ? info->getTemperature().dbl() : -1.0;
        auto absn = NC::createAbsorption( h->cfg.c_str() );
        if ( absn.isNull() ) {
          hdr.abs_c = 0.0;
        } else {
          // AbsOOV: xs = c/sqrt(E) (NCAbsOOV.cc:41-45).  c = xs(1 eV); anything that is not 1/v is flagged.
          const double c = absn.crossSectionIsotropic( NC::NeutronEnergy{1.0} ).dbl();
          const double x2 = absn.crossSectionIsotropic( NC::NeutronEnergy{0.04} ).dbl();
          const bool oov = !absn.isOriented() && std::fabs( x2*0.2 - c ) <= 1e-12*std::fabs(c);
          hdr.abs_c = oov ? c : -1.0;
        }
      }
      buf.reserve( sizeof(hdr) );
      for ( unsigned i = 0; i < hdr.ncomp; ++i ) {
        std::string err;
        if ( !compileLeaf( h->leaves[i], hdr.comp[i], buf, err ) ) { g_err = err; return nullptr; }
      }
      hdr.nbytes = ncb_align16( buf.d.size() );
      buf.d.resize( hdr.nbytes, 0 );
      buf.put( 0, &hdr, sizeof(hdr) );
      void* out = std::malloc( buf.d.size() );
      std::memcpy( out, buf.d.data(), buf.d.size() );
      *nbytes = buf.d.size();
      return out;
    } catch ( std::exception& e ) {
      g_err = e.what();
      return nullptr;
    }
  }
  void refdrv_free( void* p ) { std::free(p); }

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
