// domain.cuh -- device-resident state of one x-slab and shared device helpers.
//
// HBM layout (all fp64 unless stated; z fastest, then y, then x; x carries one ghost plane on each side so that a slab's
// x-faces are contiguous ny*nz planes that can be sent/received without packing):
//   f[2][19][nx+2][ny][nz]   population buffers A/B holding PRE-collision fBar (= what the reference's cells hold)
//                            (in-place mode: one allocation [19][nx+2+S][ny][nz] seen through two frames S planes apart, see bib_domain)
//   force[3][nx+2][ny][nz]   total external force (body + spread IB force), the reference's external scalars 0..2
//   phase[nx+2][ny][nz] u8   bit0 = inside/outside phase (reference external scalar 3, exactly 0.0/1.0), bit1 = voxelizer toggle mark
//   kind[nx+2][ny][nz]  u8   optional cell kinds (allocated on first bib_domain_set_celltype)
//   vel[3][nx+2][ny][nz]     optional dense velocity field (allocated on first bib_compute_velocity)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#define BIB_Q 19

struct ShapeDev {
    int nv = 0, nt = 0, nl = 0, np = 0;
    double area = 0, volume = 0;
    // connectivity (device)
    int* tris = nullptr;          // [nt][3]
    double* tri_area = nullptr;   // [nt]
    double* verts = nullptr;      // [nv][3] rest configuration (capsule FEM reference state)
    double* rest_radial = nullptr; // [nv] |vertex - shape centre| (semi-rigid particles)
    int* links = nullptr;         // [nl][2]
    double* link_len = nullptr;   // [nl]
    int* pairs = nullptr;         // [np][4]
    double* pair_cs = nullptr;    // [np][2]
    // per-vertex incidence lists in the reference's accumulation order (CSR)
    int* v_link_off = nullptr; int* v_link = nullptr;      // entry = link*2 + end
    int* v_tri_off = nullptr;  int* v_tri = nullptr;       // entry = tri*3 + corner
    int* v_pair_off = nullptr; int* v_pair = nullptr;      // entry = pair*4 + corner
    int* spread_perm = nullptr;   // [nv] vertices along a Morton curve through the rest shape (spread: neighbouring lanes = neighbouring vertices)
    // host copies
    std::vector<int> h_tris, h_links, h_pairs;
    std::vector<double> h_tri_area, h_link_len, h_pair_cs, h_verts;
};

struct ParticleGroup {            // particles of one (type, shape, params) added by one bib_ib_particles_add call
    int type = 0, shape = 0;
    int first_particle = 0, count = 0;
    long long first_vertex = 0;
    double params[8] = {0};
    double* link_ks = nullptr;    // [nl] per-link WLC spring constant (host-computed with libm pow, like the reference)
    double* link_kp = nullptr;    // [nl]
    double* rigid = nullptr;      // type 3 (RigidParticle3D): [count][24] com, vel, angvel, quat, force, torque, density, damping, relax, scale
    double inertia[9] = {0}; double shape_volume = 0;   // body-frame moment of inertia and volume of the (unscaled) shape
};

// Immersed boundaries on several slabs: every rank keeps a slot array holding the particles it OWNS (centre of mass inside
// its x-range, the reference's ownership rule, src/fsi/ImmersedBoundaryDynamics.hh:232-246) plus GHOST copies of neighbours'
// particles that reach into its slab. Per step three fixed-capacity neighbour exchanges keyed by global particle id:
// positions (owner -> ghost holders, carries ownership hand-off), partial interpolated velocities (ghost holder -> owner,
// summed there) and vertex forces (owner -> ghost holders, each slab spreads into its own cells only).
struct MultiSlabIB {                        // one per particle group of a multi-slab domain
    int nv = 0, cap = 0, n_global = 0, xcap = 0;
    int first_slot = 0;                    // the group's slots are [first_slot, first_slot + cap) of the domain-wide slot arrays
    long long first_vertex = 0;            // slot s holds vertices [first_vertex + (s - first_slot) * nv, + nv)
    int gid_base = 0;                      // global particle id of the group's particle 0 (ids are global over all groups, in order of addition)
    size_t entry = 0;                      // doubles per exchanged particle: gid, flag, 3*nv payload
    int* slot_of = nullptr;                // [n_global] -1 when not resident (slot index is domain-wide)
    int* free_stack = nullptr;             // [cap]
    int* counters = nullptr;               // [0] free_top, [1] overflow flag
    // Exchange buffers [channel][parity][side] (side 0 = packed for the left neighbour, 1 = for the right), [1 + xcap*entry] doubles, element 0
    // holds the entry count (uint64). Channel 0 carries POS / FORCE / PPF, channel 1 VEL (so that the VEL entries can stay packed while the
    // forces are evaluated and ONE neighbour sync serves VEL + FORCE). Consecutive exchanges of a channel alternate the parity, so a neighbour
    // may still be reading exchange n while exchange n+1 is packed (a later sync orders pack n+2 after its unpack n).
    double* send[2][2][2] = {{{nullptr, nullptr}, {nullptr, nullptr}}, {{nullptr, nullptr}, {nullptr, nullptr}}};
    // PULL transport: the neighbours' send buffers mapped into this process / device (in-process group: plain pointers; one process per
    // GPU: cudaIpcOpenMemHandle). peer[channel][parity][side] = what the neighbour on `side` packed for me. NULL: copy transport (recv buffers).
    const double* peer[2][2][2] = {{{nullptr, nullptr}, {nullptr, nullptr}}, {{nullptr, nullptr}, {nullptr, nullptr}}};
    bool pull = false;
    unsigned xseq[2] = {0, 0};             // exchanges issued so far per channel (parity = xseq & 1)
    double* recv[2] = {nullptr, nullptr};  // copy transport only
};

struct PotentialDev { int type = 0; double r_min = 0, beta = 0, strength = 0; };   // fsi/Potentials.h: 1 Morse, 2 spring

struct WallPlane {                 // one velocity boundary plane (VelocityDirichlet + PlaneFd)
    int dir = 1, orient = -1, coord = 0;
    int na = 0, nb = 0;            // size of the whole lattice face along the two in-plane axes (x,y,z order)
    int gx = 0;                    // multi-slab: ghost columns on either side along x (the stored velocities of the neighbouring slabs' wall cells); rows = na + 2 gx
    int a0 = 0, a1 = -1, b0 = 0, b1 = -1;   // the cells of the face this boundary covers (extensional_flow_setup.h: planes stop short of the edges)
    double* vel = nullptr;         // device [na][nb][3] stored boundary velocities of the WHOLE face (edge cells included: in-plane neighbours)
    std::vector<double> h_vel;     // host mirror
};
struct WallEdge {                  // one external velocity edge (StoreVelocityDynamics + OuterVelocityEdgeFunctional3D<axis, n1, n2>)
    int axis = 2, n1 = -1, n2 = -1;   // edge along `axis`; outward normals along direction1 = (axis+1)%3 and direction2 = (axis+2)%3
    int c1 = 0, c2 = 0;            // position along direction1 / direction2
    int e0 = 0, e1 = -1, n = 0;    // covered cells along the axis, lattice size along the axis
    int plane1 = -1, plane2 = -1;  // walls[] index of the face normal to direction2 (holds the direction1 neighbours) / normal to direction1
    double* vel = nullptr;         // device [n][3]
    std::vector<double> h_vel;
};

// one captured driver step and the host-side state it starts from
struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    int cur = 0, band_cur = 0, spread_band = -1; const double* vfrc = nullptr; bool forces_ready = false;
    long long launches = 0;
};

// one asynchronous read-back (bib_ib_results_download_async): staging buffer, events, the caller's buffers
#define RES_DEPTH 2
struct ResSlot {
    void* dev = nullptr; size_t bytes = 0, vb = 0;
    cudaEvent_t packed = nullptr, copied = nullptr;
    double* host_v = nullptr; double* host_s = nullptr;
    int* count_host = nullptr;            // page-locked: number of owned particles (multi-slab)
    bool compact = false, light = false; int bound_used = 0; long long bound_used_v = 0;
};

// Several slabs owned by ONE host process (bib_domain_create_multi): the slabs are stepped by one worker thread each, SPMD like the
// ranks of a multi-process run, and the neighbour exchanges use peer memory (cudaMemcpyPeer / kernels reading the neighbour's
// buffers over NVLink) ordered by CUDA events + a host barrier instead of NCCL. Defined in group.cu.
struct bib_group;

struct bib_domain {
    int device = 0;
    int n_sms = 148;                      // multiprocessors of the device (kernel-variant heuristics)
    bib_group* group = nullptr;           // non-null: slab `rank` of an in-process slab group (n_ranks slabs)
    bool is_head = false;                 // the handle bib_domain_create_multi returned: every entry point dispatches to the group's slabs
    cudaEvent_t ev_x_ready = nullptr, ev_x_done = nullptr;      // in-process exchanges: my buffers are packed / I have pulled my neighbours' buffers
    cudaEvent_t ev_h_ready = nullptr, ev_h_done = nullptr;      // same for the population halo (runs on the communication stream)
    const double* x_send[2] = {nullptr, nullptr};               // buffers published for the neighbours of the exchange in flight
    int nx = 0, ny = 0, nz = 0;           // slab size (without ghost planes)
    int per[3] = {0, 0, 0};               // global periodicity
    double omega0 = 1, omega1 = 1;        // outer, inner (omega1 == 0: single relaxation rate, see n_ext)
    int n_ext = 4;                        // external scalars per cell: 4 = ForcedPhaseD3Q19Descriptor (GuoExternalForceVOFDynamics), 3 = ForcedD3Q19Descriptor (GuoExternalForceBGKdynamics)
    size_t plane = 0;                     // ny*nz
    size_t vol = 0;                       // (nx+2)*ny*nz  (storage incl. ghost planes)
    int cur = 0;                          // which population buffer holds the current state
    double* f[2] = {nullptr, nullptr};
    // In-place streaming (BIB_INPLACE / bib_set_default_inplace): ONE population allocation of nx+2+shift planes per
    // direction; the two "buffers" are the same memory seen through two frames shifted by `inplace_shift` planes along x
    // (f[1] = f_base, f[0] = f_base + inplace_shift*plane) and `vol` (the component stride of every field) is (nx+2+shift)*plane.
    // A step pushes from one frame into the other in launches of at most shift-1 planes, ordered so that no launch writes a plane
    // a later launch still has to read. 0 = two separate buffers (A/B).
    int inplace_shift = 0;
    double* f_base = nullptr;
    double* force = nullptr;
    uint8_t* phase = nullptr;
    uint8_t* kind = nullptr;
    double* vel = nullptr;
    bool vel_full_valid = false;          // vel holds the full field of the CURRENT populations (bib_compute_velocity)
    unsigned* band_mask = nullptr;        // [2][band_cap][512] touched-cell bitmasks of the particle bands (double buffered)
    int* band_box = nullptr;              // [2][band_cap][6] box origin / extent the masks refer to
    int band_cap = 0, band_cur = 0;
    bool band_valid[2] = {false, false};
    bool force_is_body = false;           // every cell outside the particle bands holds force_body
    unsigned* fflag = nullptr;            // 1 bit per cell (storage index): 0 => force == force_body exactly; only meaningful while fflag_valid && force_is_body
    bool fflag_valid = false;
    bool flags_by_masks = false;          // the last k_interp_tile raised the force flags of every cell its masks cover (the spread of the same positions need not)
    int sparse_force = 1;                 // BIB_OPT_SPARSE_FORCE
    // Which mask buffer describes the cells the last NOT YET RESET spread wrote: 0 / 1 = that buffer (the masks ib_interpolate took at
    // the positions the spread used); -1 = nothing to reset (no spread since the force field was last set); -2 = unknown (a spread
    // without matching masks, or two spreads since the last reset) -> the next reset runs over the whole lattice.
    int spread_band = -1;
    bool band_pos_valid = false;          // the masks in buffer band_cur were taken at the CURRENT vertex positions
    double force_body[3] = {0, 0, 0};
    uint8_t* col_flags = nullptr;         // [nx+2][ny] voxelizer per-column flags
    cudaStream_t stream = nullptr;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_halo = nullptr, ev_compute = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_forces = nullptr;   // bib_step: force evaluation on comm_stream beside the interpolation
    int overlap_forces = 1;               // BIB_OPT_OVERLAP_FORCES
    int prefetch_forces = 1;              // BIB_OPT_PREFETCH_FORCES: bib_step evaluates the forces of step t+1 beside the collide-and-stream of step t
    bool forces_ready = false;            // vfrc_next holds the forces of the current vertex positions (cleared by every mutating entry point)
    double* vfrc_next = nullptr; long long vfrc_next_cap = 0;
    bool vfrc_preloaded = false;          // bib_ib_vertex_forces_preload: the next force evaluation adds to vfrc instead of starting from zero
    long long launches = 0;
    int fuse_calls = 1, fuse_stage = 0; double fuse_body[3] = {0, 0, 0};     // call fusion (api.cu): deferred interpolate / reset / forces+spread / move
    long long wall_calls = 0; double wall_t100 = 0, wall_tprev = 0;      // BIB_STEP_WALL_REPORT (api.cu)
    // per-phase device timing (bib_phase_timing): CUDA events recorded at the END of each phase of a step; read back by bib_phase_times
    bool pt_on = false;
    std::vector<std::pair<int, cudaEvent_t>> pt_marks, pt_comm_marks;      // compute stream / communication stream
    std::vector<cudaEvent_t> pt_pool;
    // CUDA graph of two consecutive whole steps (bib_step; the A/B population buffers and the band masks alternate with period 2)
    StepGraph sg[2];                      // the two steps of a period, one CUDA graph each (bib_step)
    bool graph_dirty = true; int use_graph = 1; double graph_body[3] = {0, 0, 0};
    double last_body[3] = {0, 0, 0};
    int plain_steps_since_dirty = 0;      // steps issued with plain launches since the configuration last changed (one-step-per-call callers)
    long long bytes = 0;
    // options
    int envelope_swap = 1;
    int fused_face_swap = 1;              // BIB_OPT_FUSED_FACE_SWAP: 0 separate k_face_swap pass, 1 fused for all faces, 2 fused for x/y faces only
    int deterministic_spread = 0;
    int spread_mode = 0;                  // 0 shared-memory tile staging (default), 1 one global atomic per stencil point
    long long lbm_prefetch = 32768;       // BIB_OPT_LBM_PREFETCH: L2 prefetch distance of k_collide_stream in cells (0 = off)
    int spread_sorted = 1;                // BIB_SPREAD_SORTED (experiments): spread the vertices in space-filling-curve order
    int fused_interp = 1;                 // BIB_OPT_FUSED_INTERPOLATION: band velocity + interpolation in one per-particle tile kernel
    int interp_compact = 1;               // BIB_OPT_INTERP_COMPACT (default on): k_interp_tile evaluates the band velocities slot by slot (touched cells only) instead of cell by cell
    double regulator_p = 0, regulator_target[3] = {0, 0, 0};
    int plates = 0; double plate_y0 = 0, plate_y1 = 0;      // ImmersedBoundaryDynamics3D::set_boundary: 1 ParallelPlatesBoundary(y0, y1),
    double pipe_radius = 0;                                 //   2 PipeBoundary (axis y = plate_y0, z = plate_y1, radius)
    PotentialDev wall_pot, pp_pot;                          // WallInteraction decorator / PPPotentialForce
    int wall_bkind = -1; double wall_b[3] = {0, 0, 0};      // boundary the WallInteraction refers to (-1: the one of set_boundary)
    int* pp_head = nullptr; int* pp_next = nullptr; int* pp_vobj = nullptr; long long pp_cap = 0; int pp_groups = 0;
    // multi-GPU ring
    int rank = 0, n_ranks = 1;
    void* nccl_comm = nullptr;
    double* token = nullptr;              // 4 doubles: neighbour-sync tokens (comm_neighbour_sync)
    double* halo_stage = nullptr;         // [10][ny*nz] received ghost planes when y or z has walls
    // Population halo as a PULL (comm.cu): one kernel reads the neighbours' ghost planes over peer memory (NVLink) straight into this slab's
    // boundary planes. One process per GPU: the neighbours' population allocations are mapped with cudaIpcOpenMemHandle at the first exchange.
    bool halo_peer_ok = true;             // in-process slab group: both neighbours' memory can be read by this device's kernels
    int halo_pull = -1;                   // -1 not decided yet, 0 ncclSend / ncclRecv of the planes (BIB_HALO=nccl), 1 pull kernel
    const double* peer_f[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [side 0 left / 1 right][frame]: the neighbour's f[frame]
    long long peer_vol[2] = {0, 0}; int peer_nx[2] = {0, 0};
    void* peer_map[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};         // what cudaIpcOpenMemHandle returned (closed by comm_destroy)
    int x_offset = 0;                     // global x of local plane 0
    int gnx = 0;                          // global nx
    // IB
    std::vector<WallPlane> walls;
    std::vector<WallEdge> edges;
    std::vector<int> guo_levels;          // node ranges [guo_levels[l], guo_levels[l+1]) that can be updated concurrently (see lbm_set_guo_nodes)
    int n_guo = 0; int* guo_nodes = nullptr; double* guo_delta = nullptr; double* guo_vel = nullptr;   // Guo off-lattice wall nodes
    std::vector<ShapeDev> shapes;
    std::vector<ParticleGroup> groups;
    int n_particles = 0;
    long long n_vertices = 0;
    long long cap_vertices = 0;
    int cap_particles = 0;
    double* vpos = nullptr;   // [3][cap_vertices]  SoA
    double* vvel = nullptr;
    double* vfrc = nullptr;
    double* vfrc_sum = nullptr; long long vfrc_sum_cap = 0;   // multi-direct forcing: node forces summed over the sub-iterations
    double* pstate = nullptr; // [cap_particles][11] area, volume, com[3], bbox[6]
    int* p_first_vertex = nullptr;   // [cap_particles]
    void* scratch = nullptr; size_t scratch_bytes = 0;
    void* pinned = nullptr; size_t pinned_bytes = 0;      // pinned host staging for device->host result reads
    // asynchronous result read-back (bib_ib_results_download_async): packed on the compute stream, copied on copy_stream
    cudaStream_t copy_stream = nullptr;
    ResSlot res[RES_DEPTH];               // asynchronous read-backs in flight (FIFO)
    int res_head = 0, res_tail = 0, res_n = 0;
    int* res_list = nullptr;              // per slot [1 + cap]: number and slot indices of the owned particles (multi-slab read-back)
    int res_bound = 0;                    // particles covered by the asynchronous prefix copy (last count + slack)
    long long res_bound_v = 0; int res_list_cap = 0;
    int res_last_count = -1;              // particles delivered by the last completed read-back
    // multi-slab domains: domain-wide slot arrays (one slot per resident particle, owned or ghost) + one exchange context per particle group
    MultiSlabIB* ms = nullptr;       // non-null on multi-slab domains with particles (= msv[0])
    std::vector<MultiSlabIB*> msv;   // per particle group (same index as groups[])
    int* role_all = nullptr;         // [cap_particles] 0 dead, 1 owned, 2 ghost (owner = left neighbour), 3 ghost (owner = right)
    int* gid_all = nullptr;          // [cap_particles] global particle id of the slot
    uint8_t* seen_all = nullptr;     // [2][cap_particles] refreshed-by-this-exchange marks, owner snapshot
    int* vslot = nullptr;            // [cap_vertices] slot of each vertex
    int* slot_nv = nullptr;          // [cap_particles] vertices per slot
    long long* slot_fv = nullptr;    // [cap_particles] first vertex of the slot
    bool defer_vel = false;          // bib_step: the VEL entries stay packed after the interpolation, their sync is shared with the FORCE exchange
    bool vel_packed = false;         // ... and are waiting to be unpacked
    bool ghosts_dirty = true;        // ghost copies do not reflect the owners' positions (particles added / uploaded): POS exchange before the next use
};

void bib_set_error(const std::string& msg);
// phase ids of bib_phase_times (include/biblbm_b200.h)
enum { PT_BEGIN = -1, PT_INTERPOLATE = 0, PT_X_VEL = 1, PT_RESET = 2, PT_FORCES = 3, PT_X_FORCE = 4, PT_SPREAD = 5, PT_MOVE = 6, PT_X_POS = 7, PT_VOXELIZE = 8,
       PT_LBM_BOUNDARY = 9, PT_LBM = 10, PT_HALO_WAIT = 11, PT_HALO = 12, PT_WALLS = 13, PT_COUNT = 14 };
void pt_mark(bib_domain* d, int phase, cudaStream_t st = nullptr);
bool bib_cuda_ok(cudaError_t e, const char* what);

#define BIB_CHECK(call) do { if (!bib_cuda_ok((call), #call)) return 0; } while (0)

// ---- kernels / host launchers implemented in the .cu files
int lbm_launch_collide_stream(bib_domain* d);
void lbm_inplace_schedule(int nx, int shift, int cur, std::vector<std::pair<int, int>>& launches);
int lbm_launch_upload(bib_domain* d, const int box[6], const double* cells);
int lbm_launch_download(bib_domain* d, const int box[6], double* cells);
int lbm_launch_set_external(bib_domain* d, const int box[6], double fx, double fy, double fz);
int lbm_launch_velocity(bib_domain* d, const int box[6]);
int lbm_velocity_download(bib_domain* d, const int box[6], double* velocity);
int lbm_velocity_extrapolate(bib_domain* d, int dir, int coord, int inward);
int lbm_velocity_upload(bib_domain* d, const int box[6], const double* velocity);
int lbm_set_celltype(bib_domain* d, const int box[6], int kind);
int lbm_add_wall(bib_domain* d, const int box[6], int dir, int orient);
int lbm_add_edge(bib_domain* d, const int box[6], int axis, int normal1, int normal2);
int lbm_set_wall_velocity(bib_domain* d, const int box[6], const double* vel);
int lbm_apply_walls(bib_domain* d);
int lbm_set_guo_nodes(bib_domain* d, int n, const int* loc_dir, const double* delta, const double* wall_velocity);
int lbm_apply_guo(bib_domain* d);
int lbm_wall_velocity_fill(bib_domain* d);
int lbm_init_equilibrium(bib_domain* d, const int box[6], double rho, double ux, double uy, double uz);

int ib_shape_register(bib_domain* d, int nv, int nt, const double* verts, const int* tris, int* shape_id);
int ib_particles_add(bib_domain* d, int type, int shape_id, const double* params, int n, const double* positions);
int ib_rigid_add(bib_domain* d, int shape_id, const double* inertia9, double shape_volume, int n, const double* state17);
int ib_rigid_download(bib_domain* d, int first, int count, double* state19);
int ib_rigid_upload(bib_domain* d, int first, int count, const double* state19);
int ib_results_count(bib_domain* d, int* n);
int ib_forces_preload(bib_domain* d, long long first, long long count, const double* f3);
int ib_interpolate(bib_domain* d);
int ib_band_reset(bib_domain* d, double fx, double fy, double fz);
int ib_set_kernel_attributes();     // per device: opt-in dynamic shared memory sizes (called by bib_domain_create after cudaSetDevice)
int ib_forces_and_spread(bib_domain* d);
int ib_forces(bib_domain* d, cudaStream_t st);
int ib_spread(bib_domain* d);
bool ib_forces_independent_of_interpolation(const bib_domain* d);
int ib_move_and_voxelize(bib_domain* d);
int ib_vertices_download(bib_domain* d, long long first, long long count, double* v9);
int ib_vertices_upload(bib_domain* d, long long first, long long count, const double* v9);
int ib_update_all(bib_domain* d, int do_move);
int ib_force_accumulate(bib_domain* d, int mode);   // 0: sum := 0, 1: sum += vertex forces, 2: vertex forces := -sum
int ib_scalars_download(bib_domain* d, int first, int count, double* s11);
int ib_results_download_async(bib_domain* d, double* v9, double* s11);
int ib_results_wait(bib_domain* d);
int ib_scalars_upload(bib_domain* d, int first, int count, const double* s11);

int comm_unique_id(void* out128);
int comm_init(bib_domain* d, int rank, int n_ranks, const void* id128);
int comm_halo_exchange(bib_domain* d, cudaStream_t st);
void comm_destroy(bib_domain* d);
int comm_exchange_pair(bib_domain* d, const double* send_left, const double* send_right, double* recv_left, double* recv_right, size_t n_doubles);
int comm_neighbour_sync(bib_domain* d);        // stream-ordered: returns (on d->stream) after both neighbours' streams reached the same call
int comm_map_peer_buffers(bib_domain* d, double* mine_for_left, double* mine_for_right, const double** from_left, const double** from_right);
// in-process slab groups (group.cu)
bib_domain* group_slab(bib_group* g, int rank);
void group_barrier(bib_group* g);
inline bool bib_multi(const bib_domain* d) { return d->n_ranks > 1; }
int ib_local_slots(bib_domain* d, int* gids, int* roles, int capacity, int* n_slots);

// ---- device helpers shared by the LBM and IB kernels
#ifdef __CUDACC__
// Every arithmetic expression below mirrors the reference's evaluation order; the library is compiled with -fmad=false
// so that no multiply-add is contracted and results are bit-identical to the reference's x86-64 (SSE2) build.

// momentTemplates3D.h:351-373,411-423 of the reference (partial_rho + get_rhoBar_j)
__device__ __forceinline__ void d_rhoBar_j(const double* f, double& rhoBar, double& jx, double& jy, double& jz) {
    double surfX_M1 = f[1] + f[4] + f[5] + f[6] + f[7];
    double surfX_P1 = f[10] + f[13] + f[14] + f[15] + f[16];
    double surfY_M1 = f[2] + f[4] + f[8] + f[9] + f[14];
    double surfY_P1 = f[5] + f[11] + f[13] + f[17] + f[18];
    double surfZ_M1 = f[3] + f[6] + f[8] + f[16] + f[18];
    double surfZ_P1 = f[7] + f[9] + f[12] + f[15] + f[17];
    double surfX_0 = f[0] + f[2] + f[3] + f[8] + f[9] + f[11] + f[12] + f[17] + f[18];
    rhoBar = surfX_M1 + surfX_0 + surfX_P1;
    jx = surfX_P1 - surfX_M1;
    jy = surfY_P1 - surfY_M1;
    jz = surfZ_P1 - surfZ_M1;
}
#endif
