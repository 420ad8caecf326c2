// ncb_vdos_api.h -- what ncb_vdos.cu offers to the rest of the library (C ABI and blob loader in ncb_lib.cu).
#pragma once
#include "ncb_vdos.h"

namespace ncb { namespace vdos {
  // createScatteringKernel + transformKernelToStdFormat on the current CUDA device (throws vdos::Error)
  Kernel expandOnDevice( const Input& in, unsigned vdoslux, double target_emax,
                         const std::function<double(unsigned)>& scaleFct, unsigned* launches );
  // Sjolander's G_order on its grid [xmin,xmax]
  VectD gnOnDevice( const Input& in, unsigned order, double& xmin, double& xmax );
} }
