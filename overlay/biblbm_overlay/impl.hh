// biblbm_overlay/impl.hh -- bodies of the reference-side binding declared in decl.h. Header-only, compiled into the
// driver's translation unit like all of Palabos / fsi. C++98.
#ifndef BIBLBM_OVERLAY_IMPL_HH_
#define BIBLBM_OVERLAY_IMPL_HH_

#include <cstdlib>
#include <cstring>
#include <typeinfo>

namespace plb {

namespace bibcommon {

// ------------------------------------------------------------------------------------------------ small helpers
inline void die(const char* what) {
    const char* e = bib_last_error();
    pcerr << "biblbm_b200 binding: " << what << ((e && *e) ? ": " : "") << ((e && *e) ? e : "") << std::endl;
    exit(-1);
}
#define BIBGLUE_OK(call) do { if (!(call)) ::plb::bibcommon::die(#call); } while (0)

inline int& hooks_suspended() { static int n = 0; return n; }
inline bool env_flag(const char* name) { const char* v = std::getenv(name); return v && *v && std::strcmp(v, "0") != 0; }
inline void box_to_int6(const Box3D& b, int out[6]) {
    out[0] = (int)b.x0; out[1] = (int)b.x1; out[2] = (int)b.y0; out[3] = (int)b.y1; out[4] = (int)b.z0; out[5] = (int)b.z1;
}
enum { BIB_PARTICLE_RIGID_ = 100 };      // binding-internal tag: RigidParticle3D groups go through bib_ib_rigid_particles_add

}  // namespace bibcommon

#define BIB_DESC ForcedPhaseD3Q19Descriptor
#define BIB_GLUE bibglue
#define BIB_N_EXT 4
#include "biblbm_overlay/impl_desc.inc"
#undef BIB_DESC
#undef BIB_GLUE
#undef BIB_N_EXT

#define BIB_DESC ForcedD3Q19Descriptor
#define BIB_GLUE bibglue3
#define BIB_N_EXT 3
#include "biblbm_overlay/impl_desc.inc"
#undef BIB_DESC
#undef BIB_GLUE
#undef BIB_N_EXT

// ================================================================================================ CouetteVelocityExtrapolator
inline void applyProcessingFunctional(::CouetteVelocityExtrapolator<double>* functional, Box3D domain, MultiTensorField3D<double, 3>& velocity) {
    if (bibglue::couette_on_device(*functional, domain, velocity) || bibglue3::couette_on_device(*functional, domain, velocity)) {
        delete functional;
        return;
    }
    // not a device-resident field: the reference's own path (host loop; the coherence hook hands an edited field back to the device)
    applyProcessingFunctional(static_cast<BoxProcessingFunctional3D_T<double, 3>*>(functional), domain, velocity);
}

// ================================================================================================ velocity field hook
template<>
inline TensorField3D<double, 3>& MultiTensorField3D<double, 3>::getComponent(plint blockId) {
    bibglue::before_host_access(*this, true);
    bibglue3::before_host_access(*this, true);
    BlockMap::iterator it = fields.find(blockId);
    PLB_ASSERT(it != fields.end());
    return *it->second;
}
template<>
inline TensorField3D<double, 3> const& MultiTensorField3D<double, 3>::getComponent(plint blockId) const {
    bibglue::before_host_access(*this, false);
    bibglue3::before_host_access(*this, false);
    BlockMap::const_iterator it = fields.find(blockId);
    PLB_ASSERT(it != fields.end());
    return *it->second;
}

}  // namespace plb

#endif
