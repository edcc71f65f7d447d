// biblbm_overlay/impl.hh -- bodies of the reference-side binding declared in decl.h. Header-only, compiled into the
// driver's translation unit like all of Palabos / fsi. C++98.
#ifndef BIBLBM_OVERLAY_IMPL_HH_
#define BIBLBM_OVERLAY_IMPL_HH_

#include <cstdlib>
#include <cstring>
#include <typeinfo>

namespace plb {

namespace bibglue {

// ------------------------------------------------------------------------------------------------ small helpers
inline void die(const char* what) {
    const char* e = bib_last_error();
    pcerr << "biblbm_b200 binding: " << what << ((e && *e) ? ": " : "") << ((e && *e) ? e : "") << std::endl;
    exit(-1);
}
#define BIBGLUE_OK(call) do { if (!(call)) ::plb::bibglue::die(#call); } while (0)

inline int& hooks_suspended() { static int n = 0; return n; }
inline std::map<const void*, Twin*>& twins() { static std::map<const void*, Twin*> m; return m; }
inline bool env_flag(const char* name) { const char* v = std::getenv(name); return v && *v && std::strcmp(v, "0") != 0; }
inline void box_to_int6(const Box3D& b, int out[6]) {
    out[0] = (int)b.x0; out[1] = (int)b.x1; out[2] = (int)b.y0; out[3] = (int)b.y1; out[4] = (int)b.z0; out[5] = (int)b.z1;
}
inline Box3D full_box(const Twin& t) { return Box3D(0, t.nx - 1, 0, t.ny - 1, 0, t.nz - 1); }      // device index space
inline Box3D to_device_box(const Twin& t, const Box3D& global) { return global.shift(-t.ox, -t.oy, -t.oz); }
inline Cell<double, descriptors::ForcedPhaseD3Q19Descriptor>& host_cell(const Twin& t, plint x, plint y, plint z) {   // device index -> host cell
    return t.block->get(x + t.ox - t.ex, y + t.oy - t.ey, z + t.oz - t.ez);
}
inline std::map<const void*, void*>& guo_walls() { static std::map<const void*, void*> m; return m; }
inline void register_guo_wall(const Lattice& lattice, void* guo_boundary) { guo_walls()[&lattice] = guo_boundary; }

inline void destroy_twin(Twin* t) {
    if (t->particles) t->particles->on_twin_destroyed();
    if (t->d) bib_domain_destroy(t->d);
    delete t;
}

inline Twin* find_twin(const Lattice& lattice) {
    std::map<const void*, Twin*>::iterator it = twins().find(&lattice);
    if (it == twins().end()) return 0;
    if (it->second->lattice_id != lattice.getId()) {      // a different lattice object now lives at this address
        destroy_twin(it->second);
        twins().erase(it);
        return 0;
    }
    return it->second;
}
inline Twin* find_twin_of_block(const Block& block) {
    for (std::map<const void*, Twin*>::iterator it = twins().begin(); it != twins().end(); ++it)
        if (it->second->block == &block && multiBlockRegistration3D().find(it->second->lattice_id)) return it->second;
    return 0;
}
inline void collect_dead_twins() {
    std::vector<const void*> dead;
    for (std::map<const void*, Twin*>::iterator it = twins().begin(); it != twins().end(); ++it)
        if (!multiBlockRegistration3D().find(it->second->lattice_id)) dead.push_back(it->first);
    for (pluint i = 0; i < dead.size(); ++i) { destroy_twin(twins()[dead[i]]); twins().erase(dead[i]); }
}

// ------------------------------------------------------------------------------------------------ lattice -> device
// kind of a host cell's dynamics object: 0 bulk, 1 NoDynamics, 2 BounceBack, 16 + 2*dir + (orient>0) velocity wall plane
inline int classify_dynamics(const Dynamics<double, descriptors::ForcedPhaseD3Q19Descriptor>& dyn) {
    using namespace descriptors;
    if (dynamic_cast<const GuoExternalForceVOFDynamics<double, ForcedPhaseD3Q19Descriptor>*>(&dyn)) return BIB_CELL_BULK;
    if (dynamic_cast<const NoDynamics<double, ForcedPhaseD3Q19Descriptor>*>(&dyn)) return BIB_CELL_NODYNAMICS;
    if (dynamic_cast<const BounceBack<double, ForcedPhaseD3Q19Descriptor>*>(&dyn)) return BIB_CELL_BOUNCEBACK;
    if (dynamic_cast<const VelocityDirichletBoundaryDynamics<double, ForcedPhaseD3Q19Descriptor, 0, -1>*>(&dyn)) return 16;
    if (dynamic_cast<const VelocityDirichletBoundaryDynamics<double, ForcedPhaseD3Q19Descriptor, 0, +1>*>(&dyn)) return 17;
    if (dynamic_cast<const VelocityDirichletBoundaryDynamics<double, ForcedPhaseD3Q19Descriptor, 1, -1>*>(&dyn)) return 18;
    if (dynamic_cast<const VelocityDirichletBoundaryDynamics<double, ForcedPhaseD3Q19Descriptor, 1, +1>*>(&dyn)) return 19;
    if (dynamic_cast<const VelocityDirichletBoundaryDynamics<double, ForcedPhaseD3Q19Descriptor, 2, -1>*>(&dyn)) return 20;
    if (dynamic_cast<const VelocityDirichletBoundaryDynamics<double, ForcedPhaseD3Q19Descriptor, 2, +1>*>(&dyn)) return 21;
    return -1;
}

inline void read_omegas(const Dynamics<double, descriptors::ForcedPhaseD3Q19Descriptor>& dyn, double& omega, double& omega_inner) {
    const GuoExternalForceVOFDynamics<double, descriptors::ForcedPhaseD3Q19Descriptor>* vof =
        dynamic_cast<const GuoExternalForceVOFDynamics<double, descriptors::ForcedPhaseD3Q19Descriptor>*>(&dyn);
    if (!vof) die("the background dynamics must be GuoExternalForceVOFDynamics (src/fsi/VariableViscosityDynamics.h:45-77)");
    omega = vof->getOmega();
    omega_inner = vof->omega2_;           // no accessor in the reference; see overlay/fsi/headers.h
}

inline void upload_state(Twin& t) {
    std::vector<double> cells((size_t)t.nx * t.ny * t.nz * 23);
    size_t k = 0;
    for (plint x = 0; x < t.nx; ++x) for (plint y = 0; y < t.ny; ++y) for (plint z = 0; z < t.nz; ++z) {
        Cell<double, descriptors::ForcedPhaseD3Q19Descriptor>& c = host_cell(t, x, y, z);
        for (int i = 0; i < 19; ++i) cells[k++] = c[i];
        for (int i = 0; i < 4; ++i) cells[k++] = *c.getExternal(i);
    }
    int box[6]; box_to_int6(full_box(t), box);
    BIBGLUE_OK(bib_domain_upload(t.d, box, &cells[0]));
    t.device_valid = true;
    t.velocity_evaluable = false;
}

inline void configure_dynamics(Twin& t) {
    std::map<int, int> kind_of_id;                 // Dynamics::getId() -> kind
    Box3D wall_box[6]; long long wall_cells[6] = {0, 0, 0, 0, 0, 0};
    int box[6];
    for (plint x = 0; x < t.nx; ++x) for (plint y = 0; y < t.ny; ++y) {
        plint run_start = 0; int run_kind = BIB_CELL_BULK;
        for (plint z = 0; z <= t.nz; ++z) {
            int kind = BIB_CELL_BULK;
            if (z < t.nz) {
                const Dynamics<double, descriptors::ForcedPhaseD3Q19Descriptor>& dyn = host_cell(t, x, y, z).getDynamics();
                std::map<int, int>::iterator it = kind_of_id.find(dyn.getId());
                if (it == kind_of_id.end()) {
                    kind = classify_dynamics(dyn);
                    if (kind < 0) {
                        pcerr << "biblbm_b200 binding: unsupported dynamics object (" << typeid(dyn).name() << ") at cell ("
                              << x << "," << y << "," << z << ")" << std::endl;
                        exit(-1);
                    }
                    kind_of_id[dyn.getId()] = kind;
                } else kind = it->second;
                if (kind >= 16) {
                    const int w = kind - 16;
                    if (wall_cells[w] == 0) wall_box[w] = Box3D(x, x, y, y, z, z);
                    else {
                        Box3D& b = wall_box[w];
                        b.x0 = std::min(b.x0, x); b.x1 = std::max(b.x1, x); b.y0 = std::min(b.y0, y); b.y1 = std::max(b.y1, y);
                        b.z0 = std::min(b.z0, z); b.z1 = std::max(b.z1, z);
                    }
                    ++wall_cells[w];
                    kind = BIB_CELL_BULK;           // wall-plane cells collide with the bulk dynamics (CompositeDynamics::collide)
                }
            }
            if (z == t.nz || kind != run_kind) {
                if (run_kind != BIB_CELL_BULK && z > run_start) {
                    box[0] = box[1] = (int)x; box[2] = box[3] = (int)y; box[4] = (int)run_start; box[5] = (int)z - 1;
                    BIBGLUE_OK(bib_domain_set_celltype(t.d, box, run_kind));
                }
                run_start = z; run_kind = kind;
            }
        }
    }
    // velocity wall planes: addVelocityBoundary{0,1,2}{N,P} + setBoundaryVelocity of the driver's set-up
    for (int w = 0; w < 6; ++w) {
        if (wall_cells[w] == 0) continue;
        const Box3D& b = wall_box[w];
        if ((long long)b.nCells() != wall_cells[w]) die("velocity boundary cells of one direction/orientation do not form a box");
        const int dir = w / 2, orient = (w % 2) ? +1 : -1;
        box_to_int6(b, box);
        BIBGLUE_OK(bib_add_velocity_boundary(t.d, box, dir, orient));
        std::vector<double> vel((size_t)b.nCells() * 3);
        size_t k = 0;
        for (plint x = b.x0; x <= b.x1; ++x) for (plint y = b.y0; y <= b.y1; ++y) for (plint z = b.z0; z <= b.z1; ++z) {
            Cell<double, descriptors::ForcedPhaseD3Q19Descriptor>& c = host_cell(t, x, y, z);
            Array<double, 3> u;
            c.getDynamics().computeVelocity(c, u);          // StoreVelocityDynamics: the stored boundary velocity
            vel[k++] = u[0]; vel[k++] = u[1]; vel[k++] = u[2];
        }
        BIBGLUE_OK(bib_set_boundary_velocity(t.d, box, &vel[0]));
    }
}

// Guo off-lattice wall nodes announced by GuoRigidWallBoundary::insert() (see the specialisation below)
inline void configure_guo_wall(Twin& t) {
    std::map<const void*, void*>::iterator it = guo_walls().find(t.lattice);
    if (it == guo_walls().end()) return;
    typedef fsi::GuoRigidWallBoundary<double, descriptors::ForcedPhaseD3Q19Descriptor> Guo;
    Guo* guo = static_cast<Guo*>(it->second);
    std::vector<int> ld; std::vector<double> delta, vel;
    for (Guo::iterator n = guo->begin(); n != guo->end(); ++n) {
        ld.push_back((int)(n->location.x - t.ox)); ld.push_back((int)(n->location.y - t.oy)); ld.push_back((int)(n->location.z - t.oz));
        ld.push_back((int)n->direction.x); ld.push_back((int)n->direction.y); ld.push_back((int)n->direction.z);
        delta.push_back(n->delta);
        for (int c = 0; c < 3; ++c) vel.push_back(n->velocity[c]);
    }
    if (!delta.empty()) BIBGLUE_OK(bib_set_guo_wall_nodes(t.d, (int)delta.size(), &ld[0], &delta[0], &vel[0]));
}

inline Twin* create_twin(Lattice& lattice) {
    collect_dead_twins();
    const MultiBlockManagement3D& management = lattice.getMultiBlockManagement();
    std::vector<plint> const& blocks = lattice.getLocalInfo().getBlocks();
    if (blocks.size() != 1 || management.getSparseBlockStructure().getNumBlocks() != 1)
        die("the lattice must consist of one block per process (one process per GPU slab)");
    Twin* t = new Twin;
    t->lattice = &lattice; t->lattice_id = lattice.getId();
    t->block_id = blocks[0];
    t->block = lattice.getBlockLattices().find(t->block_id)->second;
    t->env = (int)management.getEnvelopeWidth();
    const Box3D bb = lattice.getBoundingBox();
    t->ox = bb.x0; t->oy = bb.y0; t->oz = bb.z0;
    const Dot3D loc = t->block->getLocation();
    t->ex = loc.x; t->ey = loc.y; t->ez = loc.z;
    t->nx = (int)bb.getNx(); t->ny = (int)bb.getNy(); t->nz = (int)bb.getNz();
    double omega = 1., omega_inner = 1.;
    read_omegas(lattice.getBackgroundDynamics(), omega, omega_inner);
    const int per[3] = { lattice.periodicity().get(0) ? 1 : 0, lattice.periodicity().get(1) ? 1 : 0, lattice.periodicity().get(2) ? 1 : 0 };
    const char* dev = std::getenv("BIB_DEVICE");
    if (!bib_domain_create(t->nx, t->ny, t->nz, 4, omega, omega_inner, per, dev ? std::atoi(dev) : 0, &t->d)) { delete t; die("bib_domain_create"); }
    if (env_flag("BIB_DETERMINISTIC_SPREAD")) BIBGLUE_OK(bib_domain_set_option(t->d, BIB_OPT_DETERMINISTIC_SPREAD, 1.0));
    configure_dynamics(*t);
    configure_guo_wall(*t);
    upload_state(*t);
    t->host_valid = true;
    twins()[&lattice] = t;
    return t;
}

inline Twin& acquire(Lattice& lattice) {
    Twin* t = find_twin(lattice);
    if (!t) t = create_twin(lattice);
    else if (!t->device_valid) upload_state(*t);
    return *t;
}

// ------------------------------------------------------------------------------------------------ device -> host
inline void sync_to_host(Twin& t) {
    if (t.host_valid) return;
    if (!t.device_valid) die("internal: neither host nor device copy of the lattice is valid");
    std::vector<double> cells((size_t)t.nx * t.ny * t.nz * 23);
    int box[6]; box_to_int6(full_box(t), box);
    BIBGLUE_OK(bib_domain_download(t.d, box, &cells[0]));
    size_t k = 0;
    for (plint x = 0; x < t.nx; ++x) for (plint y = 0; y < t.ny; ++y) for (plint z = 0; z < t.nz; ++z) {
        Cell<double, descriptors::ForcedPhaseD3Q19Descriptor>& c = host_cell(t, x, y, z);
        for (int i = 0; i < 19; ++i) c[i] = cells[k++];
        for (int i = 0; i < 4; ++i) *c.getExternal(i) = cells[k++];
    }
    t.host_valid = true;
    SuspendHooks guard;
    t.lattice->duplicateOverlaps(modif::staticVariables);     // refresh the periodic envelope like every Palabos writer does
}

inline void velocity_to_host(Twin& t) {
    if (t.velocity_host_valid || !t.velocity) return;
    if (!t.velocity_evaluable)
        die("the velocity field was read on the host after the force field or the populations changed; "
            "read it between computeVelocity and setExternalVector (as the reference drivers do) or set BIB_EAGER_VELOCITY=1");
    const Box3D b = to_device_box(t, t.velocity_domain);
    int box[6]; box_to_int6(b, box);
    if (!env_flag("BIB_EAGER_VELOCITY")) BIBGLUE_OK(bib_compute_velocity(t.d, box));
    std::vector<double> vel((size_t)b.nCells() * 3);
    BIBGLUE_OK(bib_velocity_download(t.d, box, &vel[0]));
    SuspendHooks guard;
    VelocityBlock& vb = t.velocity->getComponent(t.block_id);
    const Dot3D loc = vb.getLocation();
    size_t k = 0;
    for (plint x = b.x0; x <= b.x1; ++x) for (plint y = b.y0; y <= b.y1; ++y) for (plint z = b.z0; z <= b.z1; ++z) {
        Array<double, 3>& u = vb.get(x + t.ox - loc.x, y + t.oy - loc.y, z + t.oz - loc.z);
        u[0] = vel[k++]; u[1] = vel[k++]; u[2] = vel[k++];
    }
    t.velocity_host_valid = true;
    t.velocity->duplicateOverlaps(modif::staticVariables);
}

inline void velocity_to_device(Twin& t) {
    // the host edited the velocity field after computeVelocity (e.g. CouetteVelocityExtrapolator, progs/common/shear_flow_setup.h:121-183)
    const Box3D b = full_box(t);
    std::vector<double> vel((size_t)b.nCells() * 3);
    SuspendHooks guard;
    VelocityBlock& vb = t.velocity->getComponent(t.block_id);
    const Dot3D loc = vb.getLocation();
    size_t k = 0;
    for (plint x = b.x0; x <= b.x1; ++x) for (plint y = b.y0; y <= b.y1; ++y) for (plint z = b.z0; z <= b.z1; ++z) {
        const Array<double, 3>& u = vb.get(x + t.ox - loc.x, y + t.oy - loc.y, z + t.oz - loc.z);
        vel[k++] = u[0]; vel[k++] = u[1]; vel[k++] = u[2];
    }
    int box[6]; box_to_int6(b, box);
    BIBGLUE_OK(bib_velocity_upload(t.d, box, &vel[0]));
}

inline void before_host_access(const Lattice& lattice, bool may_write) {
    if (hooks_suspended()) return;
    Twin* t = find_twin(lattice);
    if (!t) return;
    sync_to_host(*t);
    if (may_write) {
        // the host may change populations / externals: materialise a pending computeVelocity result first, it refers to the state as it is now
        if (t->velocity && t->velocity_pending && !t->velocity_host_valid && t->velocity_evaluable
            && multiBlockRegistration3D().find(t->velocity_id) == t->velocity) velocity_to_host(*t);
        t->device_valid = false; t->velocity_evaluable = false;
    }
}

inline void before_host_access(const VelocityField& field, bool may_write) {
    if (hooks_suspended()) return;
    for (std::map<const void*, Twin*>::iterator it = twins().begin(); it != twins().end(); ++it) {
        Twin& t = *it->second;
        if (t.velocity != &field || t.velocity_id != field.getId()) continue;
        if (!multiBlockRegistration3D().find(t.lattice_id)) continue;
        velocity_to_host(t);
        if (may_write && t.velocity_pending && !t.interpolated) t.velocity_host_modified = true;
    }
}

inline void release(Lattice& lattice) {
    Twin* t = find_twin(lattice);
    if (!t) return;
    sync_to_host(*t);
    twins().erase(&lattice);
    destroy_twin(t);
}

inline void interpolate_if_pending(Twin& t) {
    if (!t.velocity_pending || t.interpolated || !t.particles) return;
    t.particles->make_resident();
    if (t.velocity_host_modified) { velocity_to_device(t); t.velocity_host_modified = false; }
    BIBGLUE_OK(bib_ib_interpolate_velocity(t.d));
    t.interpolated = true;
}

}  // namespace bibglue

// ================================================================================================ lattice members
template<>
inline void MultiBlockLattice3D<double, descriptors::ForcedPhaseD3Q19Descriptor>::collideAndStream() {
    bibglue::Twin& t = bibglue::acquire(*this);
    if (t.velocity_pending && !t.interpolated && t.particles && t.particles->has_particles())
        bibglue::die("collideAndStream() between computeVelocity and the immersed-boundary step: this call order is not supported "
                     "(supported: computeVelocity -> setExternalVector -> applyProcessingFunctional(wrap_ibm_dynamics3D) -> collideAndStream)");
    BIBGLUE_OK(bib_collide_and_stream(t.d));
    t.host_valid = false;
    t.velocity_pending = false; t.velocity_evaluable = false; t.interpolated = false; t.velocity_host_modified = false;
    ++t.steps;
    this->incrementTime();
}

template<>
inline BlockLattice3D<double, descriptors::ForcedPhaseD3Q19Descriptor>&
MultiBlockLattice3D<double, descriptors::ForcedPhaseD3Q19Descriptor>::getComponent(plint blockId) {
    bibglue::before_host_access(*this, true);
    BlockMap::iterator it = blockLattices.find(blockId);
    PLB_ASSERT(it != blockLattices.end());
    return *it->second;
}
template<>
inline BlockLattice3D<double, descriptors::ForcedPhaseD3Q19Descriptor> const&
MultiBlockLattice3D<double, descriptors::ForcedPhaseD3Q19Descriptor>::getComponent(plint blockId) const {
    bibglue::before_host_access(*this, false);
    BlockMap::const_iterator it = blockLattices.find(blockId);
    PLB_ASSERT(it != blockLattices.end());
    return *it->second;
}
template<>
inline TensorField3D<double, 3>& MultiTensorField3D<double, 3>::getComponent(plint blockId) {
    bibglue::before_host_access(*this, true);
    BlockMap::iterator it = fields.find(blockId);
    PLB_ASSERT(it != fields.end());
    return *it->second;
}
template<>
inline TensorField3D<double, 3> const& MultiTensorField3D<double, 3>::getComponent(plint blockId) const {
    bibglue::before_host_access(*this, false);
    BlockMap::const_iterator it = fields.find(blockId);
    PLB_ASSERT(it != fields.end());
    return *it->second;
}

// ================================================================================================ (1) computeVelocity
inline void computeVelocity(bibglue::Lattice& lattice, bibglue::VelocityField& velocity, Box3D domain) {
    bibglue::Twin& t = bibglue::acquire(lattice);
    t.velocity = &velocity; t.velocity_id = velocity.getId(); t.velocity_domain = domain;
    t.velocity_pending = true; t.velocity_host_valid = false; t.velocity_evaluable = true; t.interpolated = false;
    t.velocity_host_modified = false;
    if (bibglue::env_flag("BIB_EAGER_VELOCITY")) {      // dense pass like the reference; default: band-limited, on demand
        int box[6]; bibglue::box_to_int6(bibglue::to_device_box(t, domain), box);
        BIBGLUE_OK(bib_compute_velocity(t.d, box));
    }
}

// ================================================================================================ (2) setExternalVector
inline void setExternalVector(bibglue::Lattice& lattice, Box3D domain, int vectorStartsAt, Array<double, 3> externalVector) {
    bibglue::Twin* t = bibglue::find_twin(lattice);
    if (!t || !t->device_valid || vectorStartsAt != bibglue::DescriptorD::ExternalField::forceBeginsAt) {
        // set-up phase (no device twin yet) or a non-force vector: the reference's own host implementation
        setExternalVector<double, descriptors::ForcedPhaseD3Q19Descriptor>(lattice, domain, vectorStartsAt, externalVector);
        return;
    }
    bibglue::interpolate_if_pending(*t);       // the IB sees u = j/rho + F/2 with the force field as it is BEFORE this reset
    int box[6]; bibglue::box_to_int6(bibglue::to_device_box(*t, domain), box);
    BIBGLUE_OK(bib_set_external_vector(t->d, box, externalVector[0], externalVector[1], externalVector[2]));
    t->host_valid = false; t->velocity_evaluable = false;
}

// ================================================================================================ (3) immersed boundary
namespace fsi {

template<>
inline void GuoRigidWallBoundary<double, descriptors::ForcedPhaseD3Q19Descriptor>::insert() {
    // the reference's own body (the host lattice keeps its level-1 processor: lattice.initialize() runs it once at set-up) ...
    integrateProcessingFunctional(new GuoRigidWallBoundaryFunctional<double, descriptors::ForcedPhaseD3Q19Descriptor>(*this), lattice_.getBoundingBox(), lattice_, 1);
    // ... plus: the device twin applies the same wall-node list after every collideAndStream
    bibglue::register_guo_wall(lattice_, this);
}

#define BIB_IBD_TEMPLATE template<bool pX, bool pY, bool pZ>
#define BIB_IBD ImmersedBoundaryDynamics3D<double, descriptors::ForcedPhaseD3Q19Descriptor, Periodicity3D<double, pX, pY, pZ> >

BIB_IBD_TEMPLATE
BIB_IBD::ImmersedBoundaryDynamics3D(const bibglue::Lattice& lattice, const ParticleShapeLibrary<T>& shape_library)
    : lattice_(const_cast<bibglue::Lattice*>(&lattice)), host_(lattice, shape_library), twin_(0), uploaded_(false),
      device_ahead_(false), host_modified_(false), n_uploaded_(0), boundary_(0) { }

BIB_IBD_TEMPLATE
BIB_IBD::~ImmersedBoundaryDynamics3D() {
    if (twin_ && twin_->particles == this) {
        // The device domain holds this module's particles: drop the twin (a lattice that outlives its fsi module gets a fresh one on
        // its next device call). The lattice may already be gone (rbc_in_pipe deletes it first): it is only touched if still registered.
        twin_->particles = 0;
        std::map<const void*, bibglue::Twin*>::iterator it = bibglue::twins().find(lattice_);
        if (it != bibglue::twins().end() && it->second == twin_) {
            if (multiBlockRegistration3D().find(twin_->lattice_id)) bibglue::sync_to_host(*twin_);
            bibglue::twins().erase(it);
            bibglue::destroy_twin(twin_);
        }
    }
}

BIB_IBD_TEMPLATE void BIB_IBD::init(const Box3D& dom, bibglue::Block& lattice) { pull(); host_.init(dom, lattice); host_modified_ = true; }
BIB_IBD_TEMPLATE void BIB_IBD::init() { pull(); host_.init(); host_modified_ = true; }
BIB_IBD_TEMPLATE void BIB_IBD::synchronize_particle_states() { pull(); host_.synchronize_particle_states(); host_modified_ = true; }
BIB_IBD_TEMPLATE void BIB_IBD::synchronize_particle_states_and_voxelization(const Box3D& dom, bibglue::Block& lattice) {
    pull(); host_.synchronize_particle_states_and_voxelization(dom, lattice); host_modified_ = true;
}
BIB_IBD_TEMPLATE void BIB_IBD::clear() { host_.clear(); drop_device_particles(); }
BIB_IBD_TEMPLATE void BIB_IBD::add_force_decorator(ForceDecorator<T>* f) { decorators_.push_back(f); host_.add_force_decorator(f); }
BIB_IBD_TEMPLATE void BIB_IBD::add_pp_force(PPInteractionForce<T>* f) { pp_forces_.push_back(f); host_.add_pp_force(f); }
BIB_IBD_TEMPLATE void BIB_IBD::set_boundary(Boundary<T>* b) { boundary_ = b; host_.set_boundary(b); }
BIB_IBD_TEMPLATE void BIB_IBD::set_forces_to_zero() { pull(); host_.set_forces_to_zero(); host_modified_ = true; }
BIB_IBD_TEMPLATE void BIB_IBD::move_vertices() { pull(); host_.move_vertices(); host_modified_ = true; }
BIB_IBD_TEMPLATE void BIB_IBD::add_particle(const ParticleType* particle) { pull(); host_.add_particle(particle); host_modified_ = true; }

BIB_IBD_TEMPLATE void BIB_IBD::drop_device_particles() {
    if (uploaded_) bibglue::die("removing or reloading particles after they were moved to the device is not supported yet");
}

// the Boundary a WallInteraction decorator refers to (it need not be the one given to fsi.set_boundary)
inline void set_wall_boundary(bibglue::Twin& t, const Boundary<double>* b) {
    if (const ParallelPlatesBoundary<double>* plates = dynamic_cast<const ParallelPlatesBoundary<double>*>(b)) {
        BIBGLUE_OK(bib_ib_set_wall_interaction_boundary(t.d, 1, plates->y0 - (double)t.oy, plates->y1 - (double)t.oy, 0.));
    } else if (const PipeBoundary<double>* pipe = dynamic_cast<const PipeBoundary<double>*>(b)) {
        BIBGLUE_OK(bib_ib_set_wall_interaction_boundary(t.d, 2, pipe->y0 - (double)t.oy, pipe->z0 - (double)t.oz, pipe->radius));
    } else bibglue::die("WallInteraction: unsupported Boundary type");
}

BIB_IBD_TEMPLATE void BIB_IBD::configure_decorators_and_boundary(bibglue::Twin& t) {
    for (pluint i = 0; i < decorators_.size(); ++i) {
        ForceDecorator<T>* dec = decorators_[i];
        if (ParticlePositionRegulator<T>* r = dynamic_cast<ParticlePositionRegulator<T>*>(dec)) {
            BIBGLUE_OK(bib_ib_set_regulator(t.d, r->p, r->position[0] - (T)t.ox, r->position[1] - (T)t.oy, r->position[2] - (T)t.oz));
        } else if (WallInteraction<T, MorsePotential<T> >* w = dynamic_cast<WallInteraction<T, MorsePotential<T> >*>(dec)) {
            set_wall_boundary(t, w->boundary_);
            BIBGLUE_OK(bib_ib_set_wall_interaction(t.d, BIB_POTENTIAL_MORSE, w->interaction_.r_min, w->interaction_.beta, w->interaction_.strength));
        } else if (WallInteraction<T, SpringPotential<T> >* w2 = dynamic_cast<WallInteraction<T, SpringPotential<T> >*>(dec)) {
            set_wall_boundary(t, w2->boundary_);
            BIBGLUE_OK(bib_ib_set_wall_interaction(t.d, BIB_POTENTIAL_SPRING, w2->interaction_.min_dist, 0., w2->interaction_.strength));
        } else {
            pcerr << "biblbm_b200 binding: unsupported ForceDecorator " << typeid(*dec).name() << std::endl;
            exit(-1);
        }
    }
    if (pp_forces_.size() > 1) bibglue::die("only one particle-particle interaction force is supported");
    if (pp_forces_.size() == 1) {
        if (PPPotentialForce<T, MorsePotential<T> >* m = dynamic_cast<PPPotentialForce<T, MorsePotential<T> >*>(pp_forces_[0])) {
            BIBGLUE_OK(bib_ib_set_pp_interaction(t.d, BIB_POTENTIAL_MORSE, m->potential_.r_min, m->potential_.beta, m->potential_.strength));
        } else if (PPPotentialForce<T, SpringPotential<T> >* sp = dynamic_cast<PPPotentialForce<T, SpringPotential<T> >*>(pp_forces_[0])) {
            BIBGLUE_OK(bib_ib_set_pp_interaction(t.d, BIB_POTENTIAL_SPRING, sp->potential_.min_dist, 0., sp->potential_.strength));
        } else {
            pcerr << "biblbm_b200 binding: unsupported PPInteractionForce " << typeid(*pp_forces_[0]).name() << std::endl;
            exit(-1);
        }
    }
    if (boundary_) {
        if (ParallelPlatesBoundary<T>* plates = dynamic_cast<ParallelPlatesBoundary<T>*>(boundary_)) {
            BIBGLUE_OK(bib_ib_set_parallel_plates(t.d, 1, plates->y0 - (T)t.oy, plates->y1 - (T)t.oy));
        } else if (PipeBoundary<T>* pipe = dynamic_cast<PipeBoundary<T>*>(boundary_)) {
            BIBGLUE_OK(bib_ib_set_pipe(t.d, 1, pipe->y0 - (T)t.oy, pipe->z0 - (T)t.oz, pipe->radius));
        } else {
            pcerr << "biblbm_b200 binding: unsupported Boundary " << typeid(*boundary_).name() << std::endl;
            exit(-1);
        }
    }
}

BIB_IBD_TEMPLATE void BIB_IBD::push() {
    bibglue::Twin& t = *twin_;
    const T org[3] = { (T)t.ox, (T)t.oy, (T)t.oz };          // device coordinates = global - origin (exact no-op for origin 0)
    if (!uploaded_) {
        // particles in id order (std::map), consecutive particles with equal (type, shape, parameters) form one device group
        std::vector<double> positions;
        int cur_type = -1, cur_shape = -1, cur_n = 0; double cur_params[8]; std::memset(cur_params, 0, sizeof(cur_params));
        plint expect_id = 0;
        for (ObjMapIterator it = host_.particles_begin(); ; ++it) {
            int type = -1, shape_id = -1; double params[8]; std::memset(params, 0, sizeof(params));
            const bool end = it == host_.particles_end();
            if (!end) {
                ParticleType* p = it->second;
                if (it->first != expect_id++) bibglue::die("particle ids must be 0..n-1");
                if (RBCParticle<T>* rbc = dynamic_cast<RBCParticle<T>*>(p)) {
                    type = BIB_PARTICLE_RBC;
                    const RBCParameters<T>& q = rbc->params();
                    params[0] = q.shear_modulus; params[1] = q.L0; params[2] = q.k_area_global; params[3] = q.k_area_local;
                    params[4] = q.k_volume; params[5] = q.vol_desired; params[6] = q.k_bend; params[7] = q.theta0;
                } else if (DeformableCapsuleParticle<T>* cap = dynamic_cast<DeformableCapsuleParticle<T>*>(p)) {
                    type = BIB_PARTICLE_CAPSULE;
                    params[0] = cap->params().G; params[1] = cap->params().C; params[2] = cap->params().nu_s; params[3] = cap->params().kb;
                } else if (SemiRigidParticle3D<T>* sr = dynamic_cast<SemiRigidParticle3D<T>*>(p)) {
                    type = BIB_PARTICLE_SEMIRIGID;
                    params[0] = sr->params().l0; params[1] = sr->params().k_in_plane; params[2] = sr->params().k_out_of_plane; params[3] = sr->params().k_bend;
                } else {
                    pcerr << "biblbm_b200 binding: unsupported particle type " << typeid(*p).name() << std::endl;
                    exit(-1);
                }
                const ParticleShape<T>* shape = p->shape();
                typename std::map<const ParticleShape<T>*, int>::iterator sit = shape_ids_.find(shape);
                if (sit == shape_ids_.end()) {
                    const int nv = (int)shape->count_vertices(), nt = (int)shape->count_triangles();
                    std::vector<double> v((size_t)nv * 3); std::vector<int> tri((size_t)nt * 3);
                    for (int i = 0; i < nv; ++i) for (int c = 0; c < 3; ++c) v[(size_t)i * 3 + c] = shape->get_vertex(i)[c];
                    for (int i = 0; i < nt; ++i) {
                        tri[(size_t)i * 3] = (int)shape->get_triangle(i).i0; tri[(size_t)i * 3 + 1] = (int)shape->get_triangle(i).i1;
                        tri[(size_t)i * 3 + 2] = (int)shape->get_triangle(i).i2;
                    }
                    BIBGLUE_OK(bib_ib_shape_register(t.d, nv, nt, &v[0], &tri[0], &shape_id));
                    shape_ids_[shape] = shape_id;
                } else shape_id = sit->second;
            }
            if (cur_n > 0 && (end || type != cur_type || shape_id != cur_shape || std::memcmp(params, cur_params, sizeof(params)) != 0)) {
                BIBGLUE_OK(bib_ib_particles_add(t.d, cur_type, cur_shape, cur_params, cur_n, &positions[0]));
                positions.clear(); cur_n = 0;
            }
            if (end) break;
            cur_type = type; cur_shape = shape_id; std::memcpy(cur_params, params, sizeof(params)); ++cur_n;
            ParticleType* p = it->second;
            for (plint i = 0; i < p->count_nodes(); ++i) for (int c = 0; c < 3; ++c) positions.push_back(p->get_node(i).pos[c] - org[c]);
        }
        configure_decorators_and_boundary(t);
        uploaded_ = true;
        n_uploaded_ = host_.count_particles();
    } else {
        if (host_.count_particles() != n_uploaded_) bibglue::die("adding or removing particles after the first device step is not supported yet");
        std::vector<double> v9;
        for (ObjMapIterator it = host_.particles_begin(); it != host_.particles_end(); ++it) {
            ParticleType* p = it->second;
            for (plint i = 0; i < p->count_nodes(); ++i) {
                const Vertex<T>& n = p->get_node(i);
                for (int c = 0; c < 3; ++c) v9.push_back(n.pos[c] - org[c]);
                for (int c = 0; c < 3; ++c) v9.push_back(n.vel[c]);
                for (int c = 0; c < 3; ++c) v9.push_back(n.force[c]);
            }
        }
        if (!v9.empty()) BIBGLUE_OK(bib_ib_vertices_upload(t.d, 0, (long long)(v9.size() / 9), &v9[0]));
    }
    // the cached per-particle scalars continue from the host's values (they are refreshed by update() only, i.e. after a move)
    std::vector<double> s11;
    for (ObjMapIterator it = host_.particles_begin(); it != host_.particles_end(); ++it) {
        const ParticleType* p = it->second;
        const geo::Rect<T> bb = p->bounding_box();
        s11.push_back(p->area()); s11.push_back(p->volume());
        for (int c = 0; c < 3; ++c) s11.push_back(p->center_of_mass()[c] - org[c]);
        s11.push_back(bb.x0 - org[0]); s11.push_back(bb.x1 - org[0]); s11.push_back(bb.y0 - org[1]); s11.push_back(bb.y1 - org[1]);
        s11.push_back(bb.z0 - org[2]); s11.push_back(bb.z1 - org[2]);
    }
    if (!s11.empty()) BIBGLUE_OK(bib_ib_particle_scalars_upload(t.d, 0, (int)(s11.size() / 11), &s11[0]));
    host_modified_ = false; device_ahead_ = false;
}

BIB_IBD_TEMPLATE void BIB_IBD::pull() const {
    if (!twin_ || !uploaded_ || !device_ahead_) return;
    long long nv = 0; int np = 0;
    BIBGLUE_OK(bib_ib_count(twin_->d, &np, &nv));
    std::vector<double> v9((size_t)nv * 9);
    if (nv > 0) BIBGLUE_OK(bib_ib_vertices_download(twin_->d, 0, nv, &v9[0]));
    const T org[3] = { (T)twin_->ox, (T)twin_->oy, (T)twin_->oz };
    size_t k = 0;
    for (ObjMapIterator it = host_.particles_begin(); it != host_.particles_end(); ++it) {
        ParticleType* p = it->second;
        for (plint i = 0; i < p->count_nodes(); ++i) {
            Vertex<T>& n = p->get_node(i);
            for (int c = 0; c < 3; ++c) n.pos[c] = v9[k++] + org[c];
            for (int c = 0; c < 3; ++c) n.vel[c] = v9[k++];
            for (int c = 0; c < 3; ++c) n.force[c] = v9[k++];
        }
        p->update();        // area, volume, centre of mass, bounding box from the positions (ParticleBase3D::update)
    }
    device_ahead_ = false;
}

BIB_IBD_TEMPLATE bibglue::Twin& BIB_IBD::device() {
    bibglue::Twin& t = bibglue::acquire(*lattice_);
    if (twin_ != &t) {
        if (t.particles && t.particles != this) bibglue::die("only one ImmersedBoundaryDynamics3D per lattice can be resident on the device");
        twin_ = &t; t.particles = this; uploaded_ = false; shape_ids_.clear();
    }
    if (!uploaded_ || host_modified_) push();
    return t;
}

BIB_IBD_TEMPLATE void BIB_IBD::interpolate_velocity(const Box3D&, TensorField3D<T, 3>&) {
    bibglue::Twin& t = device();
    if (t.interpolated) return;                  // already done ahead of the force reset (see setExternalVector above)
    if (!t.velocity_pending)
        bibglue::die("interpolate_velocity without a preceding computeVelocity(lattice, velocity, domain) in this time step");
    bibglue::interpolate_if_pending(t);
}
BIB_IBD_TEMPLATE void BIB_IBD::compute_and_spread_forces(const Box3D&, bibglue::Block&) {
    bibglue::Twin& t = device();
    BIBGLUE_OK(bib_ib_compute_and_spread_forces(t.d));
    t.host_valid = false; t.velocity_evaluable = false; device_ahead_ = true;
}
BIB_IBD_TEMPLATE void BIB_IBD::move_vertices_and_revoxelize(const Box3D&, bibglue::Block&) {
    bibglue::Twin& t = device();
    BIBGLUE_OK(bib_ib_move_vertices_and_revoxelize(t.d));
    t.host_valid = false; device_ahead_ = true;
}

BIB_IBD_TEMPLATE
void bib_apply_ib(ImmersedBoundaryWrapperFunctional3D<double, descriptors::ForcedPhaseD3Q19Descriptor, Periodicity3D<double, pX, pY, pZ> >* functional,
                  Box3D domain, bibglue::Lattice& lattice, bibglue::VelocityField& velocity) {
    bibglue::Twin& t = bibglue::acquire(lattice);
    bibglue::SuspendHooks guard;
    bibglue::Block& block = *t.block;
    bibglue::VelocityBlock& vb = velocity.getComponent(t.block_id);
    const Dot3D loc = block.getLocation();
    functional->process(domain.shift(-loc.x, -loc.y, -loc.z), block, vb);    // global -> block-local box, like the multi-block machinery; -> the three methods above
    delete functional;
}

#undef BIB_IBD_TEMPLATE
#undef BIB_IBD

}  // namespace fsi

}  // namespace plb

#endif
