// overlay/fsi/headers.h -- placed BEFORE the reference's src/ on the include path (-Ioverlay -I$REF/src -I$REF/progs).
// The reference drivers include "fsi/headers.h" then "fsi/headers.hh" (progs/rbc_in_shear/rbc_in_shear.cpp:6-7); this file
// first pulls in the reference's own declarations unchanged and then declares the B200 bindings (class/member
// specialisations) before the driver's main() instantiates anything. No reference source is copied here.
#ifndef BIBLBM_OVERLAY_FSI_HEADERS_H_
#define BIBLBM_OVERLAY_FSI_HEADERS_H_
#include_next "fsi/headers.h"
#include "biblbm_overlay/decl.h"
#endif
