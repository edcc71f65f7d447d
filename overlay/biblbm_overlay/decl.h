// biblbm_overlay/decl.h -- declarations of the reference-side binding to libbiblbm_b200.so (C-ABI: include/biblbm_b200.h).
//
// Host code stays the reference's C++: the unchanged drivers in progs/* keep calling
//     computeVelocity(lattice, velocity, domain)                       src/dataProcessors/dataAnalysisWrapper3D.hh
//     setExternalVector(lattice, domain, forceBeginsAt, F)             src/dataProcessors/dataInitializerWrapper3D.hh
//     applyProcessingFunctional(wrap_ibm_dynamics3D(fsi), ...)         src/fsi/DataProcessors.hh:94-108
//     lattice.collideAndStream()                                       src/multiBlock/multiBlockLattice3D.hh:306-348
// and this header re-routes exactly those four calls (for T = double, ForcedPhaseD3Q19Descriptor) to the GPU-resident
// twin of the lattice. Everything else in Palabos / fsi keeps working on the host objects; coherence is kept with two
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

namespace bibglue {

typedef descriptors::ForcedPhaseD3Q19Descriptor<double> DescriptorD;
typedef MultiBlockLattice3D<double, descriptors::ForcedPhaseD3Q19Descriptor> Lattice;
typedef BlockLattice3D<double, descriptors::ForcedPhaseD3Q19Descriptor> Block;
typedef MultiTensorField3D<double, 3> VelocityField;
typedef TensorField3D<double, 3> VelocityBlock;

// what the particle side (the ImmersedBoundaryDynamics3D binding) exposes to the lattice twin
struct ParticleClient {
    virtual ~ParticleClient() { }
    virtual void on_twin_destroyed() = 0;
    virtual void make_resident() = 0;        // particles on the device and current
    virtual bool has_particles() const = 0;
};

// the device twin of one MultiBlockLattice3D
struct Twin {
    Lattice* lattice;
    id_t lattice_id;            // MultiBlock3D::getId(): detects a new lattice object at a recycled address
    plint block_id;
    Block* block;
    bib_domain* d;
    int nx, ny, nz, env;
    plint ox, oy, oz;           // origin of the lattice's bounding box (rbc_in_pipe centres its pipe at y = z = 0): device index = global - origin
    plint ex, ey, ez;           // global index - this = index inside the atomic block (block location incl. envelope)
    bool device_valid;          // the device holds the current populations / externals
    bool host_valid;            // the host lattice holds the current populations / externals
    VelocityField* velocity;    // the driver's velocity field (bound by computeVelocity)
    id_t velocity_id;
    Box3D velocity_domain;
    bool velocity_pending;      // computeVelocity was called; evaluated lazily (band-limited) by the IB interpolation
    bool velocity_host_valid;   // the host velocity field holds what computeVelocity promised
    bool velocity_evaluable;    // populations and force field are still those computeVelocity referred to
    bool interpolated;          // this step's IB interpolation already ran (before the force reset)
    bool velocity_host_modified;// host code wrote to the velocity field after computeVelocity (e.g. CouetteVelocityExtrapolator)
    ParticleClient* particles;  // IB binding attached to this twin (or 0)
    long long steps;
    Twin() : lattice(0), lattice_id(0), block_id(0), block(0), d(0), nx(0), ny(0), nz(0), env(0), ox(0), oy(0), oz(0), ex(0), ey(0), ez(0), device_valid(false), host_valid(true),
             velocity(0), velocity_id(0), velocity_pending(false), velocity_host_valid(true), velocity_evaluable(false),
             interpolated(false), velocity_host_modified(false), particles(0), steps(0) { }
};

void die(const char* what);                      // reference error convention: pcerr + exit(-1) (src/fsi/DataProcessors.hh:99-103)
Twin* find_twin(const Lattice& lattice);         // 0 if the lattice has no (live) twin
Twin* find_twin_of_block(const Block& block);
Twin& acquire(Lattice& lattice);                 // twin with device_valid == true (created / uploaded on demand)
void sync_to_host(Twin& t);                      // host lattice := device state (no-op if host_valid)
void velocity_to_host(Twin& t);                  // host velocity field := computeVelocity result
void before_host_access(const Lattice& lattice, bool may_write);
void before_host_access(const VelocityField& field, bool may_write);
void release(Lattice& lattice);                  // drop the twin (device memory) of a lattice; its host state is synchronised first
int& hooks_suspended();
struct SuspendHooks { SuspendHooks() { ++hooks_suspended(); } ~SuspendHooks() { --hooks_suspended(); } };
void interpolate_if_pending(Twin& t);
void register_guo_wall(const Lattice& lattice, void* guo_boundary);

}  // namespace bibglue

// ---- (4) lattice.collideAndStream() and the universal coherence hook
template<> void MultiBlockLattice3D<double, descriptors::ForcedPhaseD3Q19Descriptor>::collideAndStream();
template<> BlockLattice3D<double, descriptors::ForcedPhaseD3Q19Descriptor>&
MultiBlockLattice3D<double, descriptors::ForcedPhaseD3Q19Descriptor>::getComponent(plint blockId);
template<> BlockLattice3D<double, descriptors::ForcedPhaseD3Q19Descriptor> const&
MultiBlockLattice3D<double, descriptors::ForcedPhaseD3Q19Descriptor>::getComponent(plint blockId) const;
template<> TensorField3D<double, 3>& MultiTensorField3D<double, 3>::getComponent(plint blockId);
template<> TensorField3D<double, 3> const& MultiTensorField3D<double, 3>::getComponent(plint blockId) const;

// ---- (1) computeVelocity, (2) setExternalVector: non-template overloads win over the Palabos templates
void computeVelocity(bibglue::Lattice& lattice, bibglue::VelocityField& velocity, Box3D domain);
void setExternalVector(bibglue::Lattice& lattice, Box3D domain, int vectorStartsAt, Array<double, 3> externalVector);

namespace fsi {

// GuoRigidWallBoundary::insert() (src/fsi/GuoBoundaryCondition.hh:25-29) additionally announces the wall-node list to the binding
template<> void GuoRigidWallBoundary<double, descriptors::ForcedPhaseD3Q19Descriptor>::insert();

// tagged periodicity: same arithmetic, distinct type, so that ImmersedBoundaryDynamics3D<..., BibHostPeriodicity<P>> is the
// reference's unmodified primary template (used as the host mirror) while <..., P> is the binding below
template<class P> struct BibHostPeriodicity : public P { };

// ---- (3) the immersed-boundary module. Public interface = src/fsi/ImmersedBoundaryDynamics.h:57-127, same names/semantics.
template<bool pX, bool pY, bool pZ>
class ImmersedBoundaryDynamics3D<double, descriptors::ForcedPhaseD3Q19Descriptor, Periodicity3D<double, pX, pY, pZ> >
    : public bibglue::ParticleClient {
public:
    typedef double T;
    typedef Periodicity3D<double, pX, pY, pZ> Periodicity;
    typedef ImmersedBoundaryDynamics3D<double, descriptors::ForcedPhaseD3Q19Descriptor, BibHostPeriodicity<Periodicity> > HostType;
    typedef ParticleBase3D<T> ParticleType;
    typedef typename std::map<plint, ParticleType*> ObjMapType;
    typedef typename ObjMapType::iterator ObjMapIterator;
    typedef typename ObjMapType::const_iterator ObjMapConstIterator;

    ImmersedBoundaryDynamics3D(const bibglue::Lattice& lattice, const ParticleShapeLibrary<T>& shape_library);
    virtual ~ImmersedBoundaryDynamics3D();

    void init(const Box3D& dom, bibglue::Block& lattice);
    void init();
    bool is_init() const { return host_.is_init(); }
    void synchronize_particle_states();
    void synchronize_particle_states_and_voxelization(const Box3D& dom, bibglue::Block& lattice);
    void clear();

    void add_force_decorator(ForceDecorator<T>* force_decorator);
    void add_pp_force(PPInteractionForce<T>* pp_force);

    void set_forces_to_zero();
    void interpolate_velocity(const Box3D& domain, TensorField3D<T, 3>& velocity);
    void compute_and_spread_forces(const Box3D& domain, bibglue::Block& lattice);
    void move_vertices_and_revoxelize(const Box3D& domain, bibglue::Block& lattice);
    void move_vertices();
    template<class InteractionFunctional> void compute_collision_forces(const InteractionFunctional& f, const Boundary<T>* b = 0) {
        pull(); host_.compute_collision_forces(f, b); host_modified_ = true;
    }

    void add_particle(const ParticleType* particle);
    void set_boundary(Boundary<T>* boundary);

    ObjMapIterator particles_begin() { pull(); host_modified_ = true; return host_.particles_begin(); }
    ObjMapIterator particles_end() { return host_.particles_end(); }
    pluint count_particles() const { return host_.count_particles(); }
    bool get_particle(plint i, ParticleType*& particle) { pull(); host_modified_ = true; return host_.get_particle(i, particle); }
    bool get_particle(plint i, const ParticleType*& particle) const { pull(); return host_.get_particle(i, particle); }

    plint count_total_particles() const { return host_.count_total_particles(); }
    T compute_total_volume() const { pull(); return host_.compute_total_volume(); }
    void get_all_particles(std::vector<const ParticleType*>& all) const { pull(); host_.get_all_particles(all); }

    void write_particles_as_vtk(pluint it) { pull(); host_.write_particles_as_vtk(it); }
    void write_lightweight_particle_data(pluint it) { pull(); host_.write_lightweight_particle_data(it); }
    void save_checkpoint(FileName fname) const { pull(); host_.save_checkpoint(fname); }
    void load_checkpoint(FileName fname) { host_.load_checkpoint(fname); drop_device_particles(); }

    // ParticleClient
    virtual void on_twin_destroyed() { twin_ = 0; uploaded_ = false; }
    virtual void make_resident() { device(); }
    virtual bool has_particles() const { return host_.count_particles() > 0; }

private:
    ImmersedBoundaryDynamics3D(const ImmersedBoundaryDynamics3D&);
    ImmersedBoundaryDynamics3D& operator=(const ImmersedBoundaryDynamics3D&);

    bibglue::Twin& device();                     // twin of lattice_ with the particles resident and current
    void push();                                 // host particles -> device (first use, or after the host modified them)
    void pull() const;                           // device vertices -> host particles (+ update()) if the device is ahead
    void drop_device_particles();
    void configure_decorators_and_boundary(bibglue::Twin& t);

    bibglue::Lattice* lattice_;
    mutable HostType host_;
    mutable bibglue::Twin* twin_;
    bool uploaded_;                              // the device holds this module's particles
    mutable bool device_ahead_;                  // the device advanced the particles since the last pull
    mutable bool host_modified_;                 // host code may have modified the particles since the last push
    pluint n_uploaded_;
    std::vector<ForceDecorator<T>*> decorators_;
    std::vector<PPInteractionForce<T>*> pp_forces_;
    Boundary<T>* boundary_;
    std::map<const ParticleShape<T>*, int> shape_ids_;
};

// applyProcessingFunctional(wrap_ibm_dynamics3D(fsi), domain, lattice, velocity): exact-match overload, so the three IB
// phases run on the device twin without touching the host blocks
template<bool pX, bool pY, bool pZ>
void bib_apply_ib(ImmersedBoundaryWrapperFunctional3D<double, descriptors::ForcedPhaseD3Q19Descriptor, Periodicity3D<double, pX, pY, pZ> >* functional,
                  Box3D domain, bibglue::Lattice& lattice, bibglue::VelocityField& velocity);

}  // namespace fsi

template<bool pX, bool pY, bool pZ>
inline void applyProcessingFunctional(
        fsi::ImmersedBoundaryWrapperFunctional3D<double, descriptors::ForcedPhaseD3Q19Descriptor, fsi::Periodicity3D<double, pX, pY, pZ> >* functional,
        Box3D domain, bibglue::Lattice& lattice, bibglue::VelocityField& velocity) {
    fsi::bib_apply_ib(functional, domain, lattice, velocity);
}

}  // namespace plb

#endif
