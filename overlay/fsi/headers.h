// overlay/fsi/headers.h -- placed BEFORE the reference's src/ on the include path (-Ioverlay -I$REF/src -I$REF/progs).
// The reference drivers include "fsi/headers.h" then "fsi/headers.hh" (progs/rbc_in_shear/rbc_in_shear.cpp:6-7); this file
// first pulls in the reference's own declarations unchanged and then declares the B200 bindings (class/member
// specialisations) before the driver's main() instantiates anything. No reference source is copied here.
#ifndef BIBLBM_OVERLAY_FSI_HEADERS_H_
#define BIBLBM_OVERLAY_FSI_HEADERS_H_
// A handful of fsi value classes keep their parameters private without accessors (MorsePotential, SpringPotential,
// WallInteraction, PPPotentialForce, ParallelPlatesBoundary); the binding has to hand exactly those numbers to the device.
// Access control is lifted for the reference's fsi declarations only (no layout or code change; every standard header they
// use has already been included by palabos3D.h at this point).
#include <algorithm>
#include <cmath>
#include <fstream>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>
#define private public
#define protected public
#include_next "fsi/headers.h"
#include "fsi/GuoBoundaryCondition.h"      // not part of fsi/headers.h; rbc_in_pipe includes it itself (include-guarded)
#include "common/shear_flow_setup.h"       // CouetteVelocityExtrapolator keeps geoNy / margin private; drivers include it later (include-guarded)
#undef private
#undef protected
#include "biblbm_overlay/decl.h"
#endif
