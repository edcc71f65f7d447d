// overlay/fsi/headers.hh -- see overlay/fsi/headers.h. Reference template definitions first, then the bindings' bodies.
#ifndef BIBLBM_OVERLAY_FSI_HEADERS_HH_
#define BIBLBM_OVERLAY_FSI_HEADERS_HH_
#include_next "fsi/headers.hh"
#include "biblbm_overlay/impl.hh"
#endif
