// ncb_loader_sc.h -- SCBragg part of the blob loader (oriented path).
#pragma once
#include "ncb_loader.h"
namespace ncb {
  inline void loadScBragg( LoadedMaterial&, const unsigned char*, const ncb_comp_t& )
  {
    throw std::runtime_error( "compiled material: SCBragg components not supported yet" );
  }
}
