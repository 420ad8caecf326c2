// ncrystal_b200_virtapi.hh -- the per-neutron "virtual API" boundary that OpenMC binds.
//
// libncrystal_b200.so exports   extern "C" void* ncrystal_access_virtual_api( unsigned interface_id )
// (ref: ncrystal_core/include/NCrystal/virtualapi/NCVirtAPIFactory.hh:45-52, src/virtualapi/NCVirtAPIFactory.cc:24-33).
// For interface_id 1001 it returns the address of a static std::shared_ptr<const VirtAPI_Type1_v1>; the class below
// has the layout of the reference's abstract class of that name (ref: include/NCrystal/virtualapi/
// NCVirtAPI_Type1_v1.hh:60-95 -- five virtual methods in this order, then the virtual destructor), so a client that
// was compiled against the reference's header (OpenMC's src/ncrystal_load.cpp dlopen's the library, asks for id 1001
// and copies the shared_ptr) can be pointed at this library instead.
//
// Semantics kept: units (eV, barn/atom), neutron = {ekin, ux, uy, uz} modified in place by sampleScatterUncached, no
// client-side cache, calls on one ScatterProcess may come from several threads (serialised inside).
// The client's rng is consumed draw by draw exactly as by the reference (the device function runs with the numbers
// drawn so far and asks for one more when it runs past them, csrc/ncb_replay.cu), so a client generator advances
// identically under both libraries: tests/vapi_caller.cc reproduces the reference's tests/src/app_vapit1v1/test.log.
// One neutron per call leaves the GPU idle: callers that can batch should use ncb200_crosssection_many /
// ncb200_samplescatter_manydir (include/ncrystal_b200.h) -- this boundary exists so that they do not HAVE to.
#ifndef NCRYSTAL_B200_VIRTAPI_HH
#define NCRYSTAL_B200_VIRTAPI_HH
#include <functional>
#include <memory>

namespace NCrystalVirtualAPI {

  class VirtAPI_Type1_v1 {
  public:
    class ScatterProcess;   // opaque
    virtual const ScatterProcess * createScatter( const char * cfgstr ) const = 0;
    virtual const ScatterProcess * cloneScatter( const ScatterProcess * ) const = 0;
    virtual void deallocateScatter( const ScatterProcess * ) const = 0;
    virtual double crossSectionUncached( const ScatterProcess&, const double* neutron ) const = 0;
    virtual void sampleScatterUncached( const ScatterProcess&, std::function<double()>& rng, double* neutron ) const = 0;
    static constexpr unsigned interface_id = 1001;
    virtual ~VirtAPI_Type1_v1() = default;
    VirtAPI_Type1_v1() = default;
    VirtAPI_Type1_v1( const VirtAPI_Type1_v1& ) = delete;
    VirtAPI_Type1_v1& operator=( const VirtAPI_Type1_v1& ) = delete;
  };

}

extern "C" void * ncrystal_access_virtual_api( unsigned interface_id );

namespace ncrystal_b200 {
  // what the reference's NCrystal::createVirtAPI<T>() does (NCVirtAPIFactory.hh:57-70)
  template<class TVirtAPI>
  inline std::shared_ptr<const TVirtAPI> createVirtAPI()
  {
    void * o = ncrystal_access_virtual_api( TVirtAPI::interface_id );
    return o ? *reinterpret_cast<std::shared_ptr<const TVirtAPI>*>( o ) : nullptr;
  }
}
#endif
