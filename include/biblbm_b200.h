/* biblbm_b200.h -- C-ABI of the B200-native hot path of niclasberg/biblbm.
 *
 * One handle = one x-slab of the global lattice resident on one GPU (a single GPU holds the whole lattice).
 * Every entry point returns 1 on success and 0 on failure, like the reference's own accelerator seam
 * plb::CoProcessor3D<T> (reference src/coProcessors/coProcessor3D.h:57-115, "return 1 = success, 0 = failure" :74,79,95);
 * bib_last_error() describes the last failure of the calling thread. Nothing here throws, owns host memory passed in,
 * or falls back to a CPU path: without a CUDA device bib_domain_create fails.
 *
 * Host data layouts are the reference's own:
 *   cells   : Palabos serialised cells, C order [x][y][z], per cell 19 fBar populations (f_i - t_i) in D3Q19Constants
 *             order followed by the descriptor's external scalars {Fx,Fy,Fz,phase}
 *             (reference src/atomicBlock/blockLattice3D.hh:671-689, src/core/cell.hh:103-121,
 *              src/fsi/VariableViscosityDynamics.h:25-42) = what BlockLatticeDataTransfer3D::send(staticVariables) and
 *             CoProcessor3D::send/receive carry (src/coProcessors/coProcessorCommunication3D.hh:46-50,84-90);
 *   velocity: TensorField3D<T,3>, [x][y][z][3] (reference src/atomicBlock/dataField3D.hh:272-281);
 *   vertices: fsi::Vertex<T> = {pos[3], vel[3], force[3]} (reference src/fsi/ParticleBase.h:32-37).
 * Boxes are inclusive {x0,x1,y0,y1,z0,z1} like plb::Box3D, in slab-local coordinates (x0 = 0 is the slab's first plane).
 */
#ifndef BIBLBM_B200_H
#define BIBLBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bib_domain bib_domain;   /* opaque */

/* cell kinds for bib_domain_set_celltype (reference: background dynamics / plb::NoDynamics / plb::BounceBack,
 * src/core/dynamics.hh:763-770,940-943) */
enum { BIB_CELL_BULK = 0, BIB_CELL_NODYNAMICS = 1, BIB_CELL_BOUNCEBACK = 2 };

/* particle types for bib_ib_particles_add (reference: fsi::RBCParticle, src/fsi/RBCParticle.h:68-100) */
enum { BIB_PARTICLE_RBC = 0,        /* params: RBCParameters, see bib_ib_particles_add */
       BIB_PARTICLE_CAPSULE = 1,  /* fsi::DeformableCapsuleParticle (src/fsi/DeformableCapsule.hh:47-150, neo-Hookean FEM; the shape's
                                       vertices are the stress-free state); params: {G, C, nu_s, kb, 0, 0, 0, 0} (CapsuleParameters) */
       BIB_PARTICLE_SEMIRIGID = 2 };/* fsi::SemiRigidParticle3D (src/fsi/SemiRigidParticle.hh:56-96; the platelets of rbc_in_pipe): radial springs to the
                                       centre of mass + link springs; params: {l0, k_in_plane, k_out_of_plane, k_bend, 0, 0, 0, 0} */

/* bib_domain_set_option keys */
enum {
    BIB_OPT_ENVELOPE_SWAP = 1,       /* 1 (default): reproduce the reference's periodic-envelope (omega,omega_inner) swap, see DESIGN.md */
    BIB_OPT_DETERMINISTIC_SPREAD = 2,/* 1: spread forces in the reference's vertex order (slow, bit-exact); 0 (default): fp64 atomics */
    BIB_OPT_REGULATOR = 3,           /* enable ParticlePositionRegulator (value = p, then target x,y,z via bib_ib_set_regulator) */
    BIB_OPT_SPREAD_MODE = 4,         /* 0 (default): per-particle shared-memory tile staging, one global fp64 reduction per touched cell;
                                        1: one global fp64 atomic per stencil point (81 per vertex) */
    BIB_OPT_STEP_GRAPH = 6,          /* 1 (default): bib_step replays a CUDA graph of two consecutive steps on single-slab domains (small lattices are launch-bound) */
    BIB_OPT_FUSED_FACE_SWAP = 7,     /* envelope swap of the periodic faces: 1 (default) fused into the collide-and-stream kernel, 2 fused for the x / y faces with
                                        a separate pass over the z faces, 0 separate pass over all faces */
    BIB_OPT_OVERLAP_FORCES = 8,      /* 1 (default): bib_step evaluates the vertex forces on a second stream beside the velocity interpolation (same results) */
    BIB_OPT_PREFETCH_FORCES = 9,     /* 1 (default): bib_step evaluates the vertex forces of step t+1 on the second stream while the collide-and-stream of
                                        step t runs (they depend on the advected positions only; same results; needs BIB_OPT_OVERLAP_FORCES) */
    BIB_OPT_FUSED_INTERPOLATION = 10,/* 1 (default): the band velocities are evaluated into a per-particle shared-memory tile and interpolated from there (one kernel,
                                        no dense velocity field); 0: k_band_velocity into the dense field + k_interpolate (same results, bit for bit) */
    BIB_OPT_FUSE_CALLS = 11,         /* 1 (default): on a single slab the call sequence of one driver step - bib_ib_interpolate_velocity, bib_set_external_vector(whole lattice),
                                        bib_ib_compute_and_spread_forces, bib_ib_move_vertices_and_revoxelize, bib_collide_and_stream - is deferred and executed as one
                                        bib_step (CUDA-graph replay); any other call first executes what is pending, so results are unchanged */
    BIB_OPT_INTERP_COMPACT = 13,     /* 1 (default): the fused interpolation kernel evaluates the band velocities slot by slot (only the touched cells, every lane busy);
                                        0: it walks the particle's box and skips the untouched cells; same results bit for bit */
    BIB_OPT_LBM_PREFETCH = 12,       /* L2 prefetch distance of the collide-and-stream kernel, in cells ahead in memory order (default 32768; 0 = off); results unaffected */
    BIB_OPT_SPARSE_FORCE = 5         /* 1 (default): the lattice kernel skips the force read of every cell that still holds the uniform body
                                        force (1 flag bit per cell, set by the spread, cleared by the band reset); 0: always read it */
};

const char* bib_last_error(void);
int bib_version(void);
int bib_device_count(int* count);

/* ---- lattice (replaces CoProcessor3D::addDomain / send / receive / collideAndStream, coProcessor3D.h:71-109, and
 *      MultiBlockLattice3D<double,ForcedPhaseD3Q19Descriptor>::collideAndStream, src/multiBlock/multiBlockLattice3D.hh:306-348)
 * nx,ny,nz: size of THIS slab (x is the slab direction); periodic[3]: global periodicity (x periodicity of a multi-slab
 * domain is realised by the halo exchange, see bib_comm_*). n_ext = 4: ForcedPhaseD3Q19Descriptor cells (19 + {Fx,Fy,Fz,phase}),
 * GuoExternalForceVOFDynamics with two relaxation rates; n_ext = 3: ForcedD3Q19Descriptor cells (19 + {Fx,Fy,Fz}), the single-rate
 * GuoExternalForceBGKdynamics (omega_inner is ignored; 22 doubles per cell in upload / download). */
int bib_domain_create(int nx, int ny, int nz, int n_ext, double omega_outer, double omega_inner,
                      const int periodic[3], int device, bib_domain** out);
/* The same lattice as x-slabs on SEVERAL GPUs owned by ONE host process: what `mpirun -np N ./driver` is to the reference
 * (one block per rank: src/fsi/ImmersedBoundaryDynamics.hh:49-54, src/multiBlock/multiBlockLattice3D.hh:306-348, progs/rbc_in_pipe/batch.sh).
 * nx is the GLOBAL size (slab widths differ by at most one plane: slab r starts at plane floor(r nx / n_slabs)); devices[n_slabs] = CUDA ordinal of every slab (NULL: 0..n_slabs-1; slabs may share
 * a device). The returned handle is used like any other: boxes, particle ids and vertex ranges are global, every entry point is split
 * over / broadcast to the slabs, collective work runs on one worker thread per slab with peer-to-peer (NVLink) halo and particle
 * exchanges. Not available on such a handle: rigid particles, velocity edges, vertex_forces_preload, multi_direct_forcing, the
 * asynchronous results read-back (bib_domain_slab gives the per-slab handles for that), planes normal to x. n_slabs == 1 is bib_domain_create. */
int bib_domain_create_multi(int nx, int ny, int nz, int n_ext, double omega_outer, double omega_inner, const int periodic[3],
                            int n_slabs, const int* devices, bib_domain** out);
/* In-place streaming (the reference streams in place too: BlockLattice3D::bulkStream swaps populations with the neighbours inside one
 * array, src/atomicBlock/blockLattice3D.hh:500-551). Population storage of the domains created AFTER this call:
 *   0  two buffers A/B (304 B of HBM per cell for the populations),
 *  >=2 ONE buffer of nx + 2 + shift planes (152 B per cell + shift planes): a step pushes into the same memory displaced by `shift`
 *      planes along x, in launches of shift - 1 planes ordered so that nothing unread is overwritten; results are bit-identical,
 *  -1  (default) A/B when it fits the device, in place (shift 17) when it does not.
 * The environment variable BIB_INPLACE gives the initial value. bib_domain_inplace_shift reports what a domain uses (0 = A/B). */
int bib_set_default_inplace(int shift_planes);
int bib_domain_inplace_shift(bib_domain* d, int* shift_planes);
/* The launches of ONE in-place step of an nx-plane slab whose current state is frame `frame` (0 / 1), in issue order: launch i covers the planes
 * [first_plane[i], first_plane[i] + n_planes[i]). Host-side arithmetic only (no device needed): what the library itself iterates over; the CPU
 * tests check on it that no launch overwrites a plane that it or a later launch still reads (tests/test_inplace_model.py). */
int bib_inplace_schedule(int nx, int shift_planes, int frame, int capacity, int* first_plane, int* n_planes, int* n_launches);
int bib_domain_slab_count(bib_domain* d, int* n_slabs);
int bib_domain_slab(bib_domain* d, int slab, bib_domain** out);      /* the handle of one slab (owned by the multi-slab handle) */
int bib_domain_destroy(bib_domain* d);
int bib_domain_set_option(bib_domain* d, int key, double value);
int bib_domain_upload(bib_domain* d, const int box[6], const double* cells);
int bib_domain_download(bib_domain* d, const int box[6], double* cells);
int bib_domain_set_celltype(bib_domain* d, const int box[6], int kind);
/* initializeAtEquilibrium(lattice, domain, rho, velocity): constant density/velocity equilibrium fill on the device
 * (reference src/dataProcessors/dataInitializerWrapper3D.hh:327-329, dataInitializerFunctional3D.hh:847-866) */
int bib_domain_init_equilibrium(bib_domain* d, const int box[6], double rho, double ux, double uy, double uz);

/* Velocity wall planes = OnLatticeBoundaryCondition3D::addVelocityBoundary{0,1,2}{N,P} of createInterpBoundaryCondition3D
 * (reference src/boundaryCondition/boundaryInstantiator3D.h:1017-1069: VelocityDirichletBoundaryDynamics + the level-0
 * PlaneFdBoundaryFunctional3D executed after streaming, src/boundaryCondition/finiteDifferenceBoundaryProcessor3D.hh:52-115)
 * and setBoundaryVelocity (src/dataProcessors/dataInitializerWrapper3D.hh:317-319). box must be one plane of thickness 1
 * normal to `direction` (0,1,2), orientation = -1 / +1 (outward normal). velocities: [cells of box in x,y,z order][3], or NULL
 * to keep the current values (initially 0). The two in-plane directions must be periodic. */
int bib_add_velocity_boundary(bib_domain* d, const int box[6], int direction, int orientation);
int bib_set_boundary_velocity(bib_domain* d, const int box[6], const double* velocities);
/* External velocity edge = boundaryCondition.addExternalVelocityEdge{axis}{N|P}{N|P}(box, lattice) with the "interp" boundary condition
 * (src/boundaryCondition/boundaryInstantiator3D.h:423-469: StoreVelocityDynamics + OuterVelocityEdgeFunctional3D<axis,normal1,normal2>,
 * finiteDifferenceBoundaryProcessor3D.hh:162-243), as progs/common/extensional_flow_setup.h:84-88 uses it: box is a line of cells along
 * `axis` on the lattice border, normal1 / normal2 = outward normals (+-1) along (axis+1)%3 / (axis+2)%3. The velocity planes of the two
 * faces it joins must have been added before; edges are applied after the planes (their order of insertion in the reference set-up).
 * bib_set_boundary_velocity stores velocities for every boundary cell (plane or edge) inside its box. */
int bib_add_velocity_edge(bib_domain* d, const int box[6], int axis, int normal1, int normal2);

/* Off-lattice (curved) no-slip / moving wall = fsi::GuoRigidWallBoundary (src/fsi/GuoBoundaryCondition.hh:25-198; rbc_in_pipe.cpp:303-306):
 * a list of wall nodes {location x,y,z (global lattice coordinates), inward lattice direction dx,dy,dz}, the fraction `delta` of the link
 * between the wall and the first fluid node, and the wall velocity per node (NULL = at rest). Applied after every
 * bib_collide_and_stream like the reference's level-1 internal processor. A second call replaces the list; n = 0 removes it.
 * (The list itself is geometry: GuoRigidWallBoundaryInstantiator::process builds it on the host at set-up time.) */
int bib_set_guo_wall_nodes(bib_domain* d, int n, const int* location_direction6, const double* delta, const double* wall_velocity3);

/* F7: setExternalVector(lattice, domain, forceBeginsAt, F)   (src/dataProcessors/dataInitializerFunctional3D.hh:1144-1154) */
int bib_set_external_vector(bib_domain* d, const int box[6], double fx, double fy, double fz);
/* F6: computeVelocity(lattice, velocity, domain)             (src/dataProcessors/dataAnalysisFunctional3D.hh:772-784) */
int bib_compute_velocity(bib_domain* d, const int box[6]);
int bib_velocity_download(bib_domain* d, const int box[6], double* velocity);
/* the IB interpolates from the TensorField3D the driver passes (ImmersedBoundaryDynamics3D::interpolate_velocity(domain, velocity),
 * src/fsi/ImmersedBoundaryDynamics_interpolation.hh:236-276); a driver that edits the field on the host after computeVelocity
 * (CouetteVelocityExtrapolator, progs/common/shear_flow_setup.h:121-183) hands the edited field back with this call */
int bib_velocity_upload(bib_domain* d, const int box[6], const double* velocity);
/* CouetteVelocityExtrapolator::process (progs/common/shear_flow_setup.h:121-183; rbcs_in_shear.cpp:184) on the device's velocity field:
 * velocity[plane `coord` normal to `direction`] := 2 velocity[coord + inward] - velocity[coord + 2 inward], inward = +-1. */
int bib_velocity_extrapolate_plane(bib_domain* d, int direction, int coord, int inward);
/* F1-F5: one lattice.collideAndStream() */
int bib_collide_and_stream(bib_domain* d);

/* ---- immersed boundary (replaces ImmersedBoundaryDynamics3D::{add_particle, init, interpolate_velocity,
 *      compute_and_spread_forces, move_vertices_and_revoxelize}, src/fsi/ImmersedBoundaryDynamics.h:83-85, driven through
 *      ImmersedBoundaryWrapperFunctional3D::process, src/fsi/DataProcessors.hh:94-108) */
/* builds links / rest lengths / adjacent-triangle pairs / rest angles exactly like ParticleShape (src/fsi/ParticleShape.hh:22-184) */
int bib_ib_shape_register(bib_domain* d, int nv, int nt, const double* vertices, const int* triangles, int* shape_id);
int bib_ib_shape_info(bib_domain* d, int shape_id, int* n_links, int* n_pairs, double* area, double* volume);
int bib_ib_shape_connectivity(bib_domain* d, int shape_id, int* links, double* link_len, int* pairs, double* pair_cos_sin, double* tri_area);
/* params for BIB_PARTICLE_RBC: {shear_modulus, L0, k_area_global, k_area_local, k_volume, vol_desired, k_bend, theta0}
 * (RBCParameters, src/fsi/RBCParticle.h:23-60). positions: [n][nv][3]. Particles get ids in order of addition. */
int bib_ib_particles_add(bib_domain* d, int type, int shape_id, const double* params, int n, const double* positions);
/* fsi::RigidParticle3D (src/fsi/RigidParticle.hh:31-120): direct-forcing coupling f_i += relaxation * M_i * (U_rigid(x_i) - u_i) and
 * Newton-Euler integration with a quaternion exponential map. inertia9 / shape_volume: ParticleShape::get_moment_of_inertia() (row
 * major) and get_volume() of the unscaled shape. state17 per particle: centre of mass[3], velocity[3], angular velocity[3] (world
 * frame), orientation quaternion[4], density, damping, relaxation_factor, scale. Single-slab domains (round 1). */
int bib_ib_rigid_particles_add(bib_domain* d, int shape_id, const double* inertia9, double shape_volume, int n, const double* state17);
/* state19 per particle: centre of mass, velocity, angular velocity, quaternion, total force, total torque */
int bib_ib_rigid_state_download(bib_domain* d, int first_particle, int count, double* state19);
/* same layout; a host mirror that handed RigidParticle3D objects to its caller (fsi.get_particle) gives their state back with this call
 * (node records: bib_ib_vertices_upload) */
int bib_ib_rigid_state_upload(bib_domain* d, int first_particle, int count, const double* state19);
int bib_ib_set_regulator(bib_domain* d, double p, double tx, double ty, double tz);
/* fsi.set_boundary(&ParallelPlatesBoundary(y0, y1)) (src/fsi/Boundary.h:300-365, progs/rbcs_in_shear/rbcs_in_shear.cpp:122-123): vertices
 * within 2.5 cells of a plate interpolate / spread with the Roma kernel blended into the 2-point kernel
 * (src/fsi/ImmersedBoundaryDynamics_interpolation.hh:181-230, _forces.hh:192-237) and are reflected specularly at the plates when
 * advected (src/fsi/DeformableParticle3D.hh:33-50). */
int bib_ib_set_parallel_plates(bib_domain* d, int enabled, double y0, double y1);
/* fsi.set_boundary(&PipeBoundary(y0, z0, radius, length)) (src/fsi/Boundary.h:131-253; progs/rbc_in_pipe/rbc_in_pipe.cpp:42,733): cylinder
 * along x. Same near-wall blending / reflection rules as the plates, with the pipe's distance, normal and ray-trace functions. */
int bib_ib_set_pipe(bib_domain* d, int enabled, double y0, double z0, double radius);
/* interaction potentials of src/fsi/Potentials.h:29-71 (cut-off = r_min for both) */
enum { BIB_POTENTIAL_NONE = 0, BIB_POTENTIAL_MORSE = 1 /* (r_min, beta, strength) */, BIB_POTENTIAL_SPRING = 2 /* (r_min, -, strength) */ };
/* fsi.add_force_decorator(new WallInteraction<T, Potential>(boundary, potential)) (src/fsi/ForceDecorator.hh:25-33): every vertex closer
 * than r_min to the boundary set above gets f -= potential(dist) * normal, before the membrane forces are added. */
int bib_ib_set_wall_interaction(bib_domain* d, int type, double r_min, double beta, double strength);
/* The WallInteraction decorator carries its own Boundary reference, independent of fsi.set_boundary() (rbc_in_pipe uses the wall
 * force WITHOUT the near-wall kernel blending): kind -1 = the boundary of bib_ib_set_parallel_plates / bib_ib_set_pipe (default),
 * 1 = plates (p0 = y0, p1 = y1), 2 = pipe (p0 = y0, p1 = z0, p2 = radius). */
int bib_ib_set_wall_interaction_boundary(bib_domain* d, int kind, double p0, double p1, double p2);
/* fsi.add_pp_force(new PPPotentialForce<T, Potential>(potential)) (src/fsi/ParticleParticleInteraction.h:28-44): short-range repulsion
 * between vertices of DIFFERENT particles, evaluated over the reference's collision grid (src/fsi/CollisionForces.h:62-217; r_min <= 1). */
int bib_ib_set_pp_interaction(bib_domain* d, int type, double r_min, double beta, double strength);
int bib_ib_count(bib_domain* d, int* n_particles, long long* n_vertices);
/* Multi-slab domains keep particles in slots (owned + ghost copies of neighbours' particles): global id and role
 * (0 dead, 1 owned, 2/3 ghost of the left/right neighbour) of each slot; vertex/scalar downloads index by slot. On a single
 * slab slot == particle index, all owned. bib_ib_particles_add must then be given the SAME global particle list on every
 * rank (after bib_comm_init); each rank keeps what it owns, like the reference's add_particle. */
int bib_ib_local_slots(bib_domain* d, int* global_ids, int* roles, int capacity, int* n_slots);
/* the three phases of the wrapper functional, in its order, and all three at once */
int bib_ib_interpolate_velocity(bib_domain* d);
int bib_ib_compute_and_spread_forces(bib_domain* d);
int bib_ib_move_vertices_and_revoxelize(bib_domain* d);
int bib_ib_step(bib_domain* d);
/* MultiDirectForcingFunctional3D::process of progs/sphere_in_ext_flow/sphere_in_ext_flow.cpp:26-98 (rigid particles, direct forcing with
 * sub-iterations): F := 0 on the whole lattice; relax_iterations times { velocity of every cell from the current populations and the
 * force accumulated so far (boundary cells report their stored velocity), interpolate_velocity, compute_and_spread_forces, sum the node
 * forces }; node force := -(sum); move_vertices() (no re-voxelization). Followed by bib_collide_and_stream in the driver's loop. */
int bib_ib_multi_direct_forcing(bib_domain* d, int relax_iterations);
int bib_ib_vertices_download(bib_domain* d, long long first_vertex, long long count, double* vertices9);
/* (one process per GPU: collective - every rank must call it, with count = 0 if it has nothing to upload, because the next interpolation
 * refreshes the neighbours' ghost copies with a neighbour exchange) */
int bib_ib_vertices_upload(bib_domain* d, long long first_vertex, long long count, const double* vertices9);
/* ForceDecorators the library does not implement itself (any user class derived from fsi::ForceDecorator, src/fsi/ForceDecorator.h:11-13,
 * e.g. OpticalTweezerForce of progs/deformation_test/deformation_test.cpp:24-39): the host evaluates them and hands the per-vertex result
 * ([count][3]) over; the NEXT force evaluation (bib_ib_compute_and_spread_forces / bib_ib_step) starts from these values instead of zero,
 * exactly where the reference applies its decorators (ImmersedBoundaryDynamics_forces.hh:293-297). Vertices outside the range start from 0.
 * One-shot: call it before every force evaluation that needs it. Single-slab domains. */
int bib_ib_vertex_forces_preload(bib_domain* d, long long first_vertex, long long count, const double* forces3);
/* per particle {area, volume, com[3], bbox x0,x1,y0,y1,z0,z1} (ParticleBase3D::area/volume/center_of_mass/bounding_box) */
int bib_ib_particle_scalars_download(bib_domain* d, int first_particle, int count, double* scalars11);
/* Non-blocking read-back of ALL vertex records and particle scalars of this slab (what write_lightweight_particle_data /
 * write_particles_as_vtk / the volume diagnostics of the drivers consume) into page-locked host buffers: the device->host copy
 * runs on its own stream while later steps execute. Up to two read-backs may be in flight (different host buffers); a third call
 * first completes the oldest. bib_ib_results_wait() completes the OLDEST read-back in flight; its buffers may be read afterwards. */
int bib_ib_results_download_async(bib_domain* d, double* vertices9, double* scalars11);
/* The record write_lightweight_particle_data writes per particle (src/fsi/DeformableParticle3D.hh:318-333): records11 = {global particle
 * id, centre of mass[3], tensor of gyration xx xy xz yy yz zz (:110-125), volume}, evaluated on the device for every particle this slab
 * owns, delivered through the same FIFO as bib_ib_results_download_async (bib_ib_results_wait / _count apply): 88 B per particle instead
 * of the 56 kB of its vertex records -- what a driver needs every output interval; full vertex records are needed for VTK output and
 * check-points only (progs/rbc_in_shear/rbc_in_shear.cpp:157-181). */
int bib_ib_lightweight_download_async(bib_domain* d, double* records11);
int bib_ib_results_wait(bib_domain* d);
/* On a multi-slab domain the slot arrays are sized for the worst case; the asynchronous read-back therefore delivers only the particles
 * this slab OWNS (what a rank of the reference writes), compacted on the device in ascending slot order (ids: bib_ib_local_slots, role 1),
 * at the start of both buffers. The asynchronous copy covers a prefix sized from the previous count plus slack; if more particles are
 * owned by now, bib_ib_results_wait fetches the remainder before it returns. n_particles = how many the last completed read-back
 * delivered (single slab: all particles). Call after the wait. */
int bib_ib_results_count(bib_domain* d, int* n_particles);
/* The reference caches these per particle and refreshes them only in update() (after a move); a particle that was translated
 * or rotated on the host keeps the values computed before (ParticleBase3D::set_center_of_mass, src/fsi/ParticleBase.hh). A host
 * mirror hands its cached values over with this call so that the first force evaluation sees exactly what the reference sees;
 * bib_ib_particles_add / bib_ib_vertices_upload recompute them from the positions. Single-slab domains. */
int bib_ib_particle_scalars_upload(bib_domain* d, int first_particle, int count, const double* scalars11);

/* ---- one whole driver time step, in the order every reference driver issues it
 *      (progs/rbc_in_shear/rbc_in_shear.cpp:154-186): computeVelocity -> setExternalVector(body) -> IB wrapper ->
 *      collideAndStream, repeated n_steps times without host synchronisation. */
int bib_step(bib_domain* d, double body_fx, double body_fy, double body_fz, int n_steps);

/* ---- multi-GPU, one process per GPU: x-slab ring over NCCL (replaces ParallelBlockCommunicator3D::duplicateOverlaps,
 *      src/parallelism/parallelBlockCommunicator3D.cpp:144-233). unique_id: the 128 bytes of an ncclUniqueId.
 *      NCCL carries set-up data and 1-element synchronisation tokens; the payloads (the 5 populations crossing each slab face, the
 *      particle records) are read by this library's kernels from the neighbours' buffers over NVLink peer memory (cudaIpc mappings made at
 *      the first exchange). Environment: BIB_HALO=nccl moves the population planes with ncclSend / ncclRecv instead, BIB_IB_EXCHANGE=copy
 *      the particle buffers (same value on every rank). */
int bib_comm_unique_id(void* unique_id_128);
int bib_comm_init(bib_domain* d, int rank, int n_ranks, const void* unique_id_128);

/* ---- stream / timing */
int bib_sync(bib_domain* d);
int bib_time_begin(bib_domain* d);                 /* records a CUDA event on the domain's stream */
int bib_time_end(bib_domain* d, float* milliseconds);  /* records a second event, waits, returns elapsed device time */
/* Per-phase device times of the steps issued since the last call (or since bib_phase_timing(d, 1)): milliseconds[14], summed over those
 * steps, indexed by the BIB_PHASE_* ids. While enabled every phase runs on the compute stream (no force prefetch / overlap, no step graph),
 * so the sum is an upper bound of the production step; the halo exchange itself (BIB_PHASE_HALO) runs beside BIB_PHASE_LATTICE on the
 * communication stream, BIB_PHASE_HALO_WAIT is what the compute stream still had to wait for it. */
enum { BIB_PHASE_INTERPOLATE = 0, BIB_PHASE_EXCHANGE_VEL = 1, BIB_PHASE_FORCE_RESET = 2, BIB_PHASE_FORCES = 3, BIB_PHASE_EXCHANGE_FORCE = 4,
       BIB_PHASE_SPREAD = 5, BIB_PHASE_MOVE = 6, BIB_PHASE_EXCHANGE_POS = 7, BIB_PHASE_VOXELIZE = 8, BIB_PHASE_LATTICE_BOUNDARY = 9,
       BIB_PHASE_LATTICE = 10, BIB_PHASE_HALO_WAIT = 11, BIB_PHASE_HALO = 12, BIB_PHASE_WALLS = 13, BIB_PHASE_COUNT = 14 };
int bib_phase_timing(bib_domain* d, int enable);
int bib_phase_times(bib_domain* d, double* milliseconds, int n);
int bib_kernel_launch_count(bib_domain* d, long long* launches);   /* kernels launched by this library on this domain so far */
int bib_device_memory_bytes(bib_domain* d, long long* bytes);

#ifdef __cplusplus
}
#endif
#endif
