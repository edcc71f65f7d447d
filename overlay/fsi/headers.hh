// overlay/fsi/headers.hh -- see overlay/fsi/headers.h. Reference template definitions first, then the bindings' bodies.
#ifndef BIBLBM_OVERLAY_FSI_HEADERS_HH_
#define BIBLBM_OVERLAY_FSI_HEADERS_HH_
// see overlay/fsi/headers.h: fsi/Potentials.h, fsi/CollisionForces.h ... are first included from the .hh files
#define private public
#define protected public
#include_next "fsi/headers.hh"
#include "fsi/GuoBoundaryCondition.hh"
#undef private
#undef protected
#include "biblbm_overlay/impl.hh"
#endif
