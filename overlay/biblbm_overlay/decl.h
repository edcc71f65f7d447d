// biblbm_overlay/decl.h -- declarations of the reference-side binding to libbiblbm_b200.so (C-ABI: include/biblbm_b200.h).
//
// Host code stays the reference's C++: the unchanged drivers in progs/* keep calling
//     computeVelocity(lattice, velocity, domain)                       src/dataProcessors/dataAnalysisWrapper3D.hh
//     setExternalVector(lattice, domain, forceBeginsAt, F)             src/dataProcessors/dataInitializerWrapper3D.hh
//     applyProcessingFunctional(wrap_ibm_dynamics3D(fsi), ...)         src/fsi/DataProcessors.hh:94-108
//     lattice.collideAndStream()                                       src/multiBlock/multiBlockLattice3D.hh:306-348
// and this header re-routes exactly those four calls (for T = double and the two descriptors the drivers use:
// ForcedPhaseD3Q19Descriptor and ForcedD3Q19Descriptor, see decl_desc.inc) to the GPU-resident twin of the lattice. Everything else in Palabos / fsi keeps working on the host objects; coherence is kept with two
// dirty bits per object and one universal hook: MultiBlockLattice3D::getComponent / MultiTensorField3D::getComponent,
// through which every host-side reader or writer of the data (VTK, checkpoints, data processors, ...) has to pass.
//
// Mechanisms (all standard C++98, no reference source is copied or modified):
//   * explicit specialisation of members of MultiBlockLattice3D<double,ForcedPhaseD3Q19Descriptor> and
//     MultiTensorField3D<double,3> (collideAndStream, getComponent);
//   * non-template overloads of computeVelocity / setExternalVector (preferred over the templates by overload resolution);
//   * a partial specialisation of fsi::ImmersedBoundaryDynamics3D for <double, ForcedPhaseD3Q19Descriptor,
//     Periodicity3D<double,px,py,pz>> that owns the reference's own implementation (instantiated for a tagged periodicity
//     type, so it is the unmodified primary template) as its host mirror and forwards the complete public interface.
#ifndef BIBLBM_OVERLAY_DECL_H_
#define BIBLBM_OVERLAY_DECL_H_

#include <map>
#include <string>
#include <vector>

extern "C" {
#include "biblbm_b200.h"
}
// force decorator known to the binding (progs/common is on the include path of every driver build; include-guarded)
#include "common/particle_position_regulator.h"

namespace plb {

// ---- descriptor-independent pieces
namespace bibcommon {

typedef MultiTensorField3D<double, 3> VelocityField;
typedef TensorField3D<double, 3> VelocityBlock;

// what the particle side (the ImmersedBoundaryDynamics3D binding) exposes to the lattice twin
struct ParticleClient {
    virtual ~ParticleClient() { }
    virtual void on_twin_destroyed() = 0;
    virtual void make_resident() = 0;        // particles on the device and current
    virtual bool has_particles() const = 0;
    virtual void multi_direct_forcing(int relax_iterations) = 0;
};

void die(const char* what);                      // reference error convention: pcerr + exit(-1) (src/fsi/DataProcessors.hh:99-103)
int& hooks_suspended();
struct SuspendHooks { SuspendHooks() { ++hooks_suspended(); } ~SuspendHooks() { --hooks_suspended(); } };

}  // namespace bibcommon

namespace fsi {
// tagged periodicity: same arithmetic, distinct type, so that ImmersedBoundaryDynamics3D<..., BibHostPeriodicity<P>> is the
// reference's unmodified primary template (used as the host mirror) while <..., P> is the binding
template<class P> struct BibHostPeriodicity : public P { };
}

// ---- the typed binding, once per descriptor
#define BIB_DESC ForcedPhaseD3Q19Descriptor
#define BIB_GLUE bibglue
#define BIB_N_EXT 4
#include "biblbm_overlay/decl_desc.inc"
#undef BIB_DESC
#undef BIB_GLUE
#undef BIB_N_EXT

#define BIB_DESC ForcedD3Q19Descriptor
#define BIB_GLUE bibglue3
#define BIB_N_EXT 3
#include "biblbm_overlay/decl_desc.inc"
#undef BIB_DESC
#undef BIB_GLUE
#undef BIB_N_EXT

// ---- applyProcessingFunctional(new CouetteVelocityExtrapolator<double>(geoNy, margin_y, uMax), domain, velocity)
// (progs/rbcs_in_shear/rbcs_in_shear.cpp:184): exact-match overload; the two wall planes of the velocity field are extrapolated on the
// device (bib_velocity_extrapolate_plane) instead of a download / host loop / upload of the whole field every step
void applyProcessingFunctional(::CouetteVelocityExtrapolator<double>* functional, Box3D domain, MultiTensorField3D<double, 3>& velocity);

// ---- the velocity field's coherence hook (shared by both descriptors)
template<> TensorField3D<double, 3>& MultiTensorField3D<double, 3>::getComponent(plint blockId);
template<> TensorField3D<double, 3> const& MultiTensorField3D<double, 3>::getComponent(plint blockId) const;

}  // namespace plb

#endif
