// api.cu -- the C-ABI of include/biblbm_b200.h: handle life cycle, argument checking, step sequencing.
#include "../../include/biblbm_b200.h"
#include "domain.cuh"

#include <algorithm>
#include <cstdio>
#include <ctime>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>
#include <vector>

static thread_local std::string g_last_error;
static int d_spread_sorted_env = -1;

void bib_set_error(const std::string& msg) { g_last_error = msg; }
bool bib_cuda_ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}


void ms_destroy(bib_domain* d);
int ms_check_overflow(bib_domain* d);

// Per-phase timing: an event at the END of every phase; the time since the previous mark on the same stream belongs to that phase.
void pt_mark(bib_domain* d, int phase, cudaStream_t st) {
    if (!d->pt_on) return;
    cudaEvent_t e = nullptr;
    if (!d->pt_pool.empty()) { e = d->pt_pool.back(); d->pt_pool.pop_back(); }
    else if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return; }
    const bool comm = st != nullptr && st == d->comm_stream;
    cudaEventRecord(e, comm ? d->comm_stream : d->stream);
    (comm ? d->pt_comm_marks : d->pt_marks).emplace_back(phase, e);
}

// in-process slab groups (group.cu): the head handle's version of each entry point
int group_create(int nx, int ny, int nz, int n_ext, double omega_outer, double omega_inner, const int periodic[3], int n_slabs, const int* devices, bib_domain** out);
int group_destroy(bib_domain* head);
int group_set_option(bib_domain* h, int key, double value);
int group_upload(bib_domain* h, const int box[6], const double* cells);
int group_download(bib_domain* h, const int box[6], double* cells);
int group_set_celltype(bib_domain* h, const int box[6], int kind);
int group_init_equilibrium(bib_domain* h, const int box[6], double rho, double ux, double uy, double uz);
int group_add_velocity_boundary(bib_domain* h, const int box[6], int direction, int orientation);
int group_set_boundary_velocity(bib_domain* h, const int box[6], const double* velocities);
int group_set_guo_wall_nodes(bib_domain* h, int n, const int* location_direction6, const double* delta, const double* wall_velocity3);
int group_set_external_vector(bib_domain* h, const int box[6], double fx, double fy, double fz);
int group_compute_velocity(bib_domain* h, const int box[6]);
int group_velocity_download(bib_domain* h, const int box[6], double* velocity);
int group_velocity_upload(bib_domain* h, const int box[6], const double* velocity);
int group_velocity_extrapolate_plane(bib_domain* h, int direction, int coord, int inward);
int group_collide_and_stream(bib_domain* h);
int group_shape_register(bib_domain* h, int nv, int nt, const double* vertices, const int* triangles, int* shape_id);
int group_particles_add(bib_domain* h, int type, int shape_id, const double* params, int n, const double* positions);
int group_count(bib_domain* h, int* n_particles, long long* n_vertices);
int group_vertices_download(bib_domain* h, long long first, long long count, double* v9);
int group_vertices_upload(bib_domain* h, long long first, long long count, const double* v9);
int group_scalars_download(bib_domain* h, int first, int count, double* s11);
int group_scalars_upload(bib_domain* h, int first, int count, const double* s11);
int group_local_slots(bib_domain* h, int* global_ids, int* roles, int capacity, int* n_slots);
int group_set_regulator(bib_domain* h, double p, double tx, double ty, double tz);
int group_set_parallel_plates(bib_domain* h, int enabled, double y0, double y1);
int group_set_pipe(bib_domain* h, int enabled, double y0, double z0, double radius);
int group_set_wall_interaction(bib_domain* h, int type, double r_min, double beta, double strength);
int group_set_wall_interaction_boundary(bib_domain* h, int kind, double p0, double p1, double p2);
int group_set_pp_interaction(bib_domain* h, int type, double r_min, double beta, double strength);
int group_ib_interpolate(bib_domain* h);
int group_ib_forces_spread(bib_domain* h);
int group_ib_move(bib_domain* h);
int group_ib_step(bib_domain* h);
int group_step(bib_domain* h, double fx, double fy, double fz, int n_steps);
int group_sync(bib_domain* h);
int group_time_begin(bib_domain* h);
int group_time_end(bib_domain* h, float* ms);
int group_launch_count(bib_domain* h, long long* launches);
int group_memory_bytes(bib_domain* h, long long* bytes);
bib_domain* group_slab_of(bib_domain* h, int r);
int group_size(bib_domain* h);
int group_phase_timing(bib_domain* h, int enable);
int group_phase_times(bib_domain* h, double* ms, int n);
#define BIB_HEAD(d, call) do { if ((d) && (d)->is_head) return (call); } while (0)
#define BIB_HEAD_UNSUPPORTED(d, what) do { if ((d) && (d)->is_head) { bib_set_error(what ": not available on a multi-slab (bib_domain_create_multi) handle"); return 0; } } while (0)

extern "C" {

const char* bib_last_error(void) { return g_last_error.c_str(); }
int bib_version(void) { return 100; }

int bib_device_count(int* count) {
    if (!count) { bib_set_error("device_count: null argument"); return 0; }
    BIB_CHECK(cudaGetDeviceCount(count));
    return 1;
}

// Population storage of the domains created from now on: 0 = two buffers (A/B), > 0 = in-place streaming with that many planes of
// shift (section 4a of DESIGN.md), -1 (default) = A/B when it fits the device, in-place otherwise. BIB_INPLACE in the environment sets
// the initial value (tests run the whole parity suite in both layouts that way).
static int g_default_inplace = -2;
static int default_inplace() {
    if (g_default_inplace == -2) {
        g_default_inplace = -1;
        if (const char* e = std::getenv("BIB_INPLACE")) g_default_inplace = std::atoi(e);
    }
    return g_default_inplace;
}
int bib_set_default_inplace(int shift_planes) {
    if (shift_planes < -1 || shift_planes == 1) { bib_set_error("set_default_inplace: -1 (automatic), 0 (two buffers) or a shift of >= 2 planes"); return 0; }
    g_default_inplace = shift_planes;
    return 1;
}
int bib_inplace_schedule(int nx, int shift_planes, int frame, int capacity, int* first_plane, int* n_planes, int* n_launches) {
    if (!first_plane || !n_planes || !n_launches) { bib_set_error("inplace_schedule: null argument"); return 0; }
    if (nx < 1 || shift_planes < 2 || (frame != 0 && frame != 1)) { bib_set_error("inplace_schedule: nx >= 1, shift >= 2, frame 0 or 1"); return 0; }
    std::vector<std::pair<int, int>> launches;
    lbm_inplace_schedule(nx, shift_planes, frame, launches);
    *n_launches = (int)launches.size();
    if ((int)launches.size() > capacity) { bib_set_error("inplace_schedule: capacity too small"); return 0; }
    for (size_t i = 0; i < launches.size(); ++i) { first_plane[i] = launches[i].first; n_planes[i] = launches[i].second; }
    return 1;
}
int bib_domain_inplace_shift(bib_domain* d, int* shift_planes) {
    if (!d) { bib_set_error("null domain handle"); return 0; }
    if (!shift_planes) { bib_set_error("domain_inplace_shift: null argument"); return 0; }
    *shift_planes = d->is_head ? group_slab_of(d, 0)->inplace_shift : d->inplace_shift;
    return 1;
}

// The halo exchange is issued while the interior-plane launch (a million CTAs) is still queued: without a higher priority the
// block scheduler would drain that grid first and the exchange would run after it instead of underneath it.
static cudaError_t create_comm_stream(cudaStream_t* st) {
    int lo = 0, hi = 0;
    cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);      // hi = numerically lowest = greatest priority
    if (e != cudaSuccess) return e;
    return cudaStreamCreateWithPriority(st, cudaStreamNonBlocking, hi);
}

int bib_domain_create(int nx, int ny, int nz, int n_ext, double omega_outer, double omega_inner, const int periodic[3], int device,
                      bib_domain** out) {
    if (!out || !periodic) { bib_set_error("domain_create: null argument"); return 0; }
    *out = nullptr;
    if (nx < 1 || ny < 1 || nz < 1) { bib_set_error("domain_create: empty lattice"); return 0; }
    if (n_ext != 4 && n_ext != 3) { bib_set_error("domain_create: n_ext must be 4 (ForcedPhaseD3Q19Descriptor) or 3 (ForcedD3Q19Descriptor)"); return 0; }
    if (n_ext == 3) omega_inner = 0.0;      // GuoExternalForceBGKdynamics: one relaxation rate (the kernels read omega1 == 0 as "use omega0 as it is")
    if (!(omega_outer > 0.0 && omega_outer < 2.0) || (n_ext == 4 && !(omega_inner > 0.0 && omega_inner < 2.0))) { bib_set_error("domain_create: omega must lie in (0,2)"); return 0; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        bib_set_error("domain_create: no CUDA device available (this library has no CPU fallback)");
        return 0;
    }
    if (device < 0 || device >= ndev) { bib_set_error("domain_create: device index out of range"); return 0; }
    BIB_CHECK(cudaSetDevice(device));
    if (!ib_set_kernel_attributes()) return 0;          // function attributes are per device: set for every device a domain is created on
    if (const char* so = std::getenv("BIB_SPREAD_SORTED")) d_spread_sorted_env = std::atoi(so);
    if (const char* g = std::getenv("BIB_L2_FETCH")) { cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)std::atoi(g)); cudaGetLastError(); }   // experiments: 32 / 64 / 128
    bib_domain* d = new (std::nothrow) bib_domain();
    if (!d) { bib_set_error("domain_create: out of host memory"); return 0; }
    { int sms = 0; if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) d->n_sms = sms; }
    if (d_spread_sorted_env >= 0) d->spread_sorted = d_spread_sorted_env;
    if (const char* ic = std::getenv("BIB_INTERP_COMPACT")) d->interp_compact = std::atoi(ic) != 0;      // experiments / tests: initial value of BIB_OPT_INTERP_COMPACT
    d->device = device; d->nx = nx; d->ny = ny; d->nz = nz; d->gnx = nx; d->x_offset = 0;
    for (int k = 0; k < 3; ++k) d->per[k] = periodic[k] ? 1 : 0;
    d->omega0 = omega_outer; d->omega1 = omega_inner; d->n_ext = n_ext;
    d->plane = (size_t)ny * nz;
    d->vol = (size_t)(nx + 2) * d->plane;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    size_t need = d->vol * (2 * BIB_Q + 3) * sizeof(double) + d->vol + (size_t)(nx + 2) * ny + (64 << 20);
    // In-place streaming: forced (shift > 0), or chosen when the two-buffer layout does not fit this device (auto)
    const int inplace_req = default_inplace();
    if (inplace_req > 0 || (inplace_req < 0 && need > free_b)) {
        d->inplace_shift = inplace_req > 0 ? std::max(2, inplace_req) : 17;
        d->vol = (size_t)(nx + 2 + d->inplace_shift) * d->plane;
        need = d->vol * (BIB_Q + 3) * sizeof(double) + d->vol + (size_t)(nx + 2) * ny + (64 << 20);
    }
    if (need > free_b) {
        char buf[256];
        std::snprintf(buf, sizeof buf, "domain_create: lattice needs %.1f GB of HBM%s, %.1f GB free", need / 1e9,
                      d->inplace_shift ? " (in-place streaming)" : "", free_b / 1e9);
        bib_set_error(buf); delete d; return 0;
    }
    bool ok = bib_cuda_ok(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking), "cudaStreamCreate") &&
              bib_cuda_ok(create_comm_stream(&d->comm_stream), "cudaStreamCreate") &&
              bib_cuda_ok(cudaEventCreate(&d->ev0), "cudaEventCreate") && bib_cuda_ok(cudaEventCreate(&d->ev1), "cudaEventCreate") &&
              bib_cuda_ok(cudaEventCreateWithFlags(&d->ev_halo, cudaEventDisableTiming), "cudaEventCreate") &&
              bib_cuda_ok(cudaEventCreateWithFlags(&d->ev_compute, cudaEventDisableTiming), "cudaEventCreate") &&
              bib_cuda_ok(cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming), "cudaEventCreate") &&
              bib_cuda_ok(cudaEventCreateWithFlags(&d->ev_forces, cudaEventDisableTiming), "cudaEventCreate");
    if (d->inplace_shift) {
        ok = ok && bib_cuda_ok(cudaMalloc(&d->f_base, d->vol * BIB_Q * sizeof(double)), "cudaMalloc(populations)") &&
             bib_cuda_ok(cudaMemsetAsync(d->f_base, 0, d->vol * BIB_Q * sizeof(double), d->stream), "cudaMemset");
        if (ok) { d->f[1] = d->f_base; d->f[0] = d->f_base + (size_t)d->inplace_shift * d->plane; }
    } else {
        for (int b = 0; ok && b < 2; ++b) {
            ok = bib_cuda_ok(cudaMalloc(&d->f[b], d->vol * BIB_Q * sizeof(double)), "cudaMalloc(populations)") &&
                 bib_cuda_ok(cudaMemsetAsync(d->f[b], 0, d->vol * BIB_Q * sizeof(double), d->stream), "cudaMemset");
        }
    }
    ok = ok && bib_cuda_ok(cudaMalloc(&d->force, d->vol * 3 * sizeof(double)), "cudaMalloc(force)") &&
         bib_cuda_ok(cudaMemsetAsync(d->force, 0, d->vol * 3 * sizeof(double), d->stream), "cudaMemset") &&
         bib_cuda_ok(cudaMalloc(&d->phase, d->vol), "cudaMalloc(phase)") && bib_cuda_ok(cudaMemsetAsync(d->phase, 0, d->vol, d->stream), "cudaMemset") &&
         bib_cuda_ok(cudaMalloc(&d->col_flags, (size_t)(nx + 2) * ny + 4), "cudaMalloc(col_flags)") &&
         bib_cuda_ok(cudaMemsetAsync(d->col_flags, 0, (size_t)(nx + 2) * ny + 4, d->stream), "cudaMemset") &&
         bib_cuda_ok(cudaStreamSynchronize(d->stream), "cudaStreamSynchronize");
    if (!ok) { bib_domain_destroy(d); return 0; }
    d->bytes = (long long)(d->vol * ((d->inplace_shift ? 1 : 2) * BIB_Q + 3) * sizeof(double) + d->vol);
    *out = d;
    return 1;
}

int bib_domain_create_multi(int nx, int ny, int nz, int n_ext, double omega_outer, double omega_inner, const int periodic[3],
                            int n_slabs, const int* devices, bib_domain** out) {
    if (!out || !periodic) { bib_set_error("domain_create_multi: null argument"); return 0; }
    *out = nullptr;
    if (n_slabs == 1) return bib_domain_create(nx, ny, nz, n_ext, omega_outer, omega_inner, periodic, devices ? devices[0] : 0, out);
    return group_create(nx, ny, nz, n_ext, omega_outer, omega_inner, periodic, n_slabs, devices, out);
}
int bib_domain_slab_count(bib_domain* d, int* n_slabs) {
    if (!d) { bib_set_error("null domain handle"); return 0; }
    if (!n_slabs) { bib_set_error("slab_count: null argument"); return 0; }
    *n_slabs = d->is_head ? group_size(d) : 1;
    return 1;
}
int bib_domain_slab(bib_domain* d, int slab, bib_domain** out) {
    if (!d) { bib_set_error("null domain handle"); return 0; }
    if (!out) { bib_set_error("domain_slab: null argument"); return 0; }
    *out = d->is_head ? group_slab_of(d, slab) : (slab == 0 ? d : nullptr);
    if (!*out) { bib_set_error("domain_slab: slab index out of range"); return 0; }
    return 1;
}

int bib_domain_destroy(bib_domain* d) {
    if (!d) return 1;
    if (d->is_head) return group_destroy(d);
    if (d->group) { bib_set_error("domain_destroy: this slab belongs to a multi-slab handle; destroy that handle"); return 0; }
    cudaSetDevice(d->device);
    if (d->stream) cudaStreamSynchronize(d->stream);
    comm_destroy(d);
    ms_destroy(d);
    if (d->f_base) cudaFree(d->f_base);
    else for (int b = 0; b < 2; ++b) if (d->f[b]) cudaFree(d->f[b]);
    if (d->force) cudaFree(d->force);
    if (d->phase) cudaFree(d->phase);
    if (d->kind) cudaFree(d->kind);
    if (d->vel) cudaFree(d->vel);
    if (d->col_flags) cudaFree(d->col_flags);
    if (d->halo_stage) cudaFree(d->halo_stage);
    if (d->band_mask) cudaFree(d->band_mask);
    if (d->band_box) cudaFree(d->band_box);
    if (d->scratch) cudaFree(d->scratch);
    if (d->pinned) cudaFreeHost(d->pinned);
    if (d->copy_stream) { cudaStreamSynchronize(d->copy_stream); cudaStreamDestroy(d->copy_stream); }
    for (ResSlot& r : d->res) {
        if (r.dev) cudaFree(r.dev);
        if (r.packed) cudaEventDestroy(r.packed);
        if (r.copied) cudaEventDestroy(r.copied);
        if (r.count_host) cudaFreeHost(r.count_host);
    }
    if (d->res_list) cudaFree(d->res_list);
    for (StepGraph& g : d->sg) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (d->fflag) cudaFree(d->fflag);
    if (d->guo_nodes) cudaFree(d->guo_nodes);
    if (d->guo_delta) cudaFree(d->guo_delta);
    if (d->guo_vel) cudaFree(d->guo_vel);
    if (d->pp_head) cudaFree(d->pp_head);
    if (d->pp_next) cudaFree(d->pp_next);
    if (d->pp_vobj) cudaFree(d->pp_vobj);
    if (d->vpos) cudaFree(d->vpos);
    if (d->vvel) cudaFree(d->vvel);
    if (d->vfrc) cudaFree(d->vfrc);
    if (d->vfrc_sum) cudaFree(d->vfrc_sum);
    if (d->pstate) cudaFree(d->pstate);
    for (void* q : {(void*)d->role_all, (void*)d->gid_all, (void*)d->seen_all, (void*)d->vslot, (void*)d->slot_nv, (void*)d->slot_fv}) if (q) cudaFree(q);
    for (ShapeDev& s : d->shapes) {
        void* ptrs[] = {s.tris, s.tri_area, s.links, s.link_len, s.pairs, s.pair_cs, s.v_link_off, s.v_link, s.v_tri_off, s.v_tri, s.v_pair_off, s.v_pair, s.verts, s.rest_radial, s.spread_perm};
        for (void* p : ptrs) if (p) cudaFree(p);
    }
    for (WallPlane& w : d->walls) if (w.vel) cudaFree(w.vel);
    for (WallEdge& e : d->edges) if (e.vel) cudaFree(e.vel);
    for (ParticleGroup& g : d->groups) { if (g.link_ks) cudaFree(g.link_ks); if (g.link_kp) cudaFree(g.link_kp); if (g.rigid) cudaFree(g.rigid); }
    if (d->ev0) cudaEventDestroy(d->ev0);
    if (d->ev1) cudaEventDestroy(d->ev1);
    if (d->ev_halo) cudaEventDestroy(d->ev_halo);
    if (d->ev_compute) cudaEventDestroy(d->ev_compute);
    if (d->ev_fork) cudaEventDestroy(d->ev_fork);
    if (d->vfrc_next) cudaFree(d->vfrc_next);
    if (d->token) cudaFree(d->token);
    if (d->ev_forces) cudaEventDestroy(d->ev_forces);
    if (d->stream) cudaStreamDestroy(d->stream);
    if (d->comm_stream) cudaStreamDestroy(d->comm_stream);
    delete d;
    return 1;
}

// BIB_ENTER: entry points that may change the configuration or the state outside the stepping loop (they invalidate the captured
// step graph); BIB_ENTER_RO: stepping itself, synchronisation, timing and read-only downloads.
// Call fusion (BIB_OPT_FUSE_CALLS): a driver issues the IB phases, the force reset and collideAndStream as separate calls
// (progs/rbc_in_shear/rbc_in_shear.cpp:154-186 through the binding: interpolate, setExternalVector, forces + spread, move, collideAndStream).
// On a single slab these five calls are DEFERRED: when the sequence completes in exactly that order it is executed as one bib_step, i.e. by
// replaying the captured step graph (one launch instead of ~16); any other entry point first executes what is pending, call by call, so
// nothing observable changes. Errors of deferred calls surface at the call that executes them.
static int fuse_flush(bib_domain* d);
#define BIB_ENTER_RO(d) do { if (!(d)) { bib_set_error("null domain handle"); return 0; } BIB_CHECK(cudaSetDevice((d)->device)); \
                             if ((d)->fuse_stage && !fuse_flush(d)) return 0; } while (0)
#define BIB_ENTER(d) do { BIB_ENTER_RO(d); (d)->graph_dirty = true; (d)->forces_ready = false; (d)->plain_steps_since_dirty = 0; } while (0)

static int set_external_full_or_band(bib_domain* d, const int box[6], double fx, double fy, double fz);
static int step_n(bib_domain* d, double body_fx, double body_fy, double body_fz, int n_steps);
static bool fusable(const bib_domain* d) {
    return d->fuse_calls && d->use_graph && d->n_ranks == 1 && !d->is_head && !d->pt_on && !d->vel_full_valid && d->n_vertices > 0 && !d->vfrc_preloaded;
}
static int fuse_flush(bib_domain* d) {
    const int stage = d->fuse_stage;
    d->fuse_stage = 0;
    d->graph_dirty = true; d->forces_ready = false; d->plain_steps_since_dirty = 0;      // what BIB_ENTER of the deferred calls would have done
    const int box[6] = {0, d->nx - 1, 0, d->ny - 1, 0, d->nz - 1};
    if (stage >= 1 && !ib_interpolate(d)) return 0;
    if (stage >= 2 && !set_external_full_or_band(d, box, d->fuse_body[0], d->fuse_body[1], d->fuse_body[2])) return 0;
    if (stage >= 3 && !ib_forces_and_spread(d)) return 0;
    if (stage >= 4 && !ib_move_and_voxelize(d)) return 0;
    return 1;
}

int bib_domain_set_option(bib_domain* d, int key, double value) {
    BIB_HEAD(d, group_set_option(d, key, value));
    BIB_ENTER(d);
    switch (key) {
        case BIB_OPT_ENVELOPE_SWAP: d->envelope_swap = value != 0.0; return 1;
        case BIB_OPT_DETERMINISTIC_SPREAD: d->deterministic_spread = value != 0.0; return 1;
        case BIB_OPT_REGULATOR: d->regulator_p = value; return 1;
        case BIB_OPT_SPREAD_MODE: d->spread_mode = value != 0.0 ? 1 : 0; return 1;
        case BIB_OPT_STEP_GRAPH: d->use_graph = value != 0.0 ? 1 : 0; return 1;
        case BIB_OPT_FUSED_FACE_SWAP: d->fused_face_swap = value == 2.0 ? 2 : (value != 0.0 ? 1 : 0); return 1;
        case BIB_OPT_OVERLAP_FORCES: d->overlap_forces = value != 0.0 ? 1 : 0; return 1;
        case BIB_OPT_PREFETCH_FORCES: d->prefetch_forces = value != 0.0 ? 1 : 0; return 1;
        case BIB_OPT_LBM_PREFETCH: d->lbm_prefetch = value > 0.0 ? (long long)value : 0; return 1;
        case BIB_OPT_FUSE_CALLS: d->fuse_calls = value != 0.0 ? 1 : 0; return 1;
        case BIB_OPT_FUSED_INTERPOLATION: d->fused_interp = value != 0.0 ? 1 : 0; return 1;
        case BIB_OPT_INTERP_COMPACT: d->interp_compact = value != 0.0 ? 1 : 0; return 1;
        case BIB_OPT_SPARSE_FORCE: d->sparse_force = value != 0.0 ? 1 : 0; d->fflag_valid = false; return 1;
    }
    bib_set_error("set_option: unknown key");
    return 0;
}

int bib_domain_upload(bib_domain* d, const int box[6], const double* cells) {
    if (d && d->is_head && box && cells) return group_upload(d, box, cells);
    BIB_ENTER(d);
    if (!box || !cells) { bib_set_error("upload: null argument"); return 0; }
    d->vel_full_valid = false; d->force_is_body = false; d->fflag_valid = false;
    if (!lbm_launch_upload(d, box, cells)) return 0;
    return 1;
}
int bib_domain_download(bib_domain* d, const int box[6], double* cells) {
    if (d && d->is_head && box && cells) return group_download(d, box, cells);
    BIB_ENTER_RO(d);
    if (!box || !cells) { bib_set_error("download: null argument"); return 0; }
    return lbm_launch_download(d, box, cells);
}
int bib_domain_set_celltype(bib_domain* d, const int box[6], int kind) {
    if (d && d->is_head && box) return group_set_celltype(d, box, kind);
    BIB_ENTER(d);
    if (!box) { bib_set_error("set_celltype: null argument"); return 0; }
    return lbm_set_celltype(d, box, kind);
}
int bib_domain_init_equilibrium(bib_domain* d, const int box[6], double rho, double ux, double uy, double uz) {
    if (d && d->is_head && box) return group_init_equilibrium(d, box, rho, ux, uy, uz);
    BIB_ENTER(d);
    if (!box) { bib_set_error("init_equilibrium: null argument"); return 0; }
    if (!(rho > 0.0)) { bib_set_error("init_equilibrium: density must be positive"); return 0; }
    return lbm_init_equilibrium(d, box, rho, ux, uy, uz);
}
int bib_add_velocity_boundary(bib_domain* d, const int box[6], int direction, int orientation) {
    if (d && d->is_head && box) return group_add_velocity_boundary(d, box, direction, orientation);
    BIB_ENTER(d);
    if (!box) { bib_set_error("add_velocity_boundary: null argument"); return 0; }
    return lbm_add_wall(d, box, direction, orientation);
}
int bib_set_boundary_velocity(bib_domain* d, const int box[6], const double* velocities) {
    if (d && d->is_head && box && velocities) return group_set_boundary_velocity(d, box, velocities);
    BIB_ENTER(d);
    if (!box || !velocities) { bib_set_error("set_boundary_velocity: null argument"); return 0; }
    return lbm_set_wall_velocity(d, box, velocities);
}
int bib_add_velocity_edge(bib_domain* d, const int box[6], int axis, int normal1, int normal2) {
    BIB_HEAD_UNSUPPORTED(d, "add_velocity_edge");
    BIB_ENTER(d);
    if (!box) { bib_set_error("add_velocity_edge: null argument"); return 0; }
    return lbm_add_edge(d, box, axis, normal1, normal2);
}
int bib_set_guo_wall_nodes(bib_domain* d, int n, const int* location_direction6, const double* delta, const double* wall_velocity3) {
    BIB_HEAD(d, group_set_guo_wall_nodes(d, n, location_direction6, delta, wall_velocity3));
    BIB_ENTER(d);
    if (n > 0 && (!location_direction6 || !delta)) { bib_set_error("set_guo_wall_nodes: null argument"); return 0; }
    return lbm_set_guo_nodes(d, n, location_direction6, delta, wall_velocity3);
}
int bib_set_external_vector(bib_domain* d, const int box[6], double fx, double fy, double fz) {
    if (d && d->is_head && box) return group_set_external_vector(d, box, fx, fy, fz);
    if (d && box && d->fuse_stage == 1 && box[0] == 0 && box[1] == d->nx - 1 && box[2] == 0 && box[3] == d->ny - 1 && box[4] == 0 && box[5] == d->nz - 1) {
        d->fuse_stage = 2; d->fuse_body[0] = fx; d->fuse_body[1] = fy; d->fuse_body[2] = fz;      // deferred (call fusion)
        return 1;
    }
    BIB_ENTER(d);
    if (!box) { bib_set_error("set_external_vector: null argument"); return 0; }
    return set_external_full_or_band(d, box, fx, fy, fz);
}
static int set_external_full_or_band(bib_domain* d, const int box[6], double fx, double fy, double fz) {
    const bool full = box[0] == 0 && box[1] == d->nx - 1 && box[2] == 0 && box[3] == d->ny - 1 && box[4] == 0 && box[5] == d->nz - 1;
    // Every cell outside the bands the last spread touched still holds the same body force: reset those bands only (ib_band_reset
    // knows which mask buffer the spread used and falls back to the whole lattice when it does not).
    if (full && d->force_is_body && d->force_body[0] == fx && d->force_body[1] == fy && d->force_body[2] == fz)
        return ib_band_reset(d, fx, fy, fz);
    d->force_is_body = full;
    if (full) { d->force_body[0] = fx; d->force_body[1] = fy; d->force_body[2] = fz; }
    return lbm_launch_set_external(d, box, fx, fy, fz);
}
int bib_compute_velocity(bib_domain* d, const int box[6]) {
    if (d && d->is_head && box) return group_compute_velocity(d, box);
    BIB_ENTER(d);
    if (!box) { bib_set_error("compute_velocity: null argument"); return 0; }
    if (!lbm_launch_velocity(d, box)) return 0;
    d->vel_full_valid = box[0] == 0 && box[1] == d->nx - 1 && box[2] == 0 && box[3] == d->ny - 1 && box[4] == 0 && box[5] == d->nz - 1;
    return 1;
}
int bib_velocity_download(bib_domain* d, const int box[6], double* velocity) {
    if (d && d->is_head && box && velocity) return group_velocity_download(d, box, velocity);
    BIB_ENTER(d);
    if (!box || !velocity) { bib_set_error("velocity_download: null argument"); return 0; }
    return lbm_velocity_download(d, box, velocity);
}
int bib_velocity_upload(bib_domain* d, const int box[6], const double* velocity) {
    if (d && d->is_head && box && velocity) return group_velocity_upload(d, box, velocity);
    BIB_ENTER(d);
    if (!box || !velocity) { bib_set_error("velocity_upload: null argument"); return 0; }
    if (!lbm_velocity_upload(d, box, velocity)) return 0;
    d->vel_full_valid = box[0] == 0 && box[1] == d->nx - 1 && box[2] == 0 && box[3] == d->ny - 1 && box[4] == 0 && box[5] == d->nz - 1;
    return 1;
}
int bib_velocity_extrapolate_plane(bib_domain* d, int direction, int coord, int inward) {
    BIB_HEAD(d, group_velocity_extrapolate_plane(d, direction, coord, inward));
    BIB_ENTER(d);
    return lbm_velocity_extrapolate(d, direction, coord, inward);
}
// BIB_STEP_WALL_REPORT=<file>: wall-clock time between successive bib_collide_and_stream calls (= one driver step when a driver's time
// loop calls it once per step), written every 500 calls as "<calls> <mean ms per step since call 100> <median ms per step>". The mean
// includes whatever the driver does between steps (its output cadence: VTK / particle dumps every 500 - 1000 steps); the median is the
// cost of an ordinary step. Measures a driver run without differencing whole-process run times (start-up jitter of +-1 s).
static void step_wall_report(bib_domain* d) {
    static const char* path = std::getenv("BIB_STEP_WALL_REPORT");
    if (!path) return;
    static std::vector<double> deltas;
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    const double now = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    d->wall_calls++;
    if (d->wall_calls == 100) d->wall_t100 = now;
    if (d->wall_calls > 100) deltas.push_back(now - d->wall_tprev);
    d->wall_tprev = now;
    if (d->wall_calls % 500 == 0 && d->wall_calls >= 1000) {
        std::vector<double> sorted(deltas);
        std::nth_element(sorted.begin(), sorted.begin() + sorted.size() / 2, sorted.end());
        if (FILE* f = std::fopen(path, "w")) { std::fprintf(f, "%lld %.6f %.6f\n", d->wall_calls, (now - d->wall_t100) / (double)(d->wall_calls - 100), sorted[sorted.size() / 2]); std::fclose(f); }
    }
}
int bib_collide_and_stream(bib_domain* d) {
    BIB_HEAD(d, group_collide_and_stream(d));
    if (d && d->fuse_stage == 4) {            // the deferred calls of one driver step are complete: execute them as ONE step (graph replay)
        BIB_CHECK(cudaSetDevice(d->device));
        d->fuse_stage = 0;
        step_wall_report(d);
        return step_n(d, d->fuse_body[0], d->fuse_body[1], d->fuse_body[2], 1);
    }
    BIB_ENTER(d);
    d->vel_full_valid = false;
    step_wall_report(d);
    return lbm_launch_collide_stream(d);
}

int bib_ib_shape_register(bib_domain* d, int nv, int nt, const double* vertices, const int* triangles, int* shape_id) {
    if (d && d->is_head && shape_id) return group_shape_register(d, nv, nt, vertices, triangles, shape_id);
    BIB_ENTER(d);
    if (!shape_id) { bib_set_error("shape_register: null argument"); return 0; }
    return ib_shape_register(d, nv, nt, vertices, triangles, shape_id);
}
int bib_ib_shape_info(bib_domain* d, int shape_id, int* n_links, int* n_pairs, double* area, double* volume) {
    if (d && d->is_head) return bib_ib_shape_info(group_slab_of(d, 0), shape_id, n_links, n_pairs, area, volume);
    BIB_ENTER(d);
    if (shape_id < 0 || shape_id >= (int)d->shapes.size()) { bib_set_error("shape_info: unknown shape"); return 0; }
    const ShapeDev& s = d->shapes[shape_id];
    if (n_links) *n_links = s.nl;
    if (n_pairs) *n_pairs = s.np;
    if (area) *area = s.area;
    if (volume) *volume = s.volume;
    return 1;
}
int bib_ib_shape_connectivity(bib_domain* d, int shape_id, int* links, double* link_len, int* pairs, double* pair_cos_sin, double* tri_area) {
    if (d && d->is_head) return bib_ib_shape_connectivity(group_slab_of(d, 0), shape_id, links, link_len, pairs, pair_cos_sin, tri_area);
    BIB_ENTER(d);
    if (shape_id < 0 || shape_id >= (int)d->shapes.size()) { bib_set_error("shape_connectivity: unknown shape"); return 0; }
    const ShapeDev& s = d->shapes[shape_id];
    if (links) std::memcpy(links, s.h_links.data(), s.h_links.size() * sizeof(int));
    if (link_len) std::memcpy(link_len, s.h_link_len.data(), s.h_link_len.size() * sizeof(double));
    if (pairs) std::memcpy(pairs, s.h_pairs.data(), s.h_pairs.size() * sizeof(int));
    if (pair_cos_sin) std::memcpy(pair_cos_sin, s.h_pair_cs.data(), s.h_pair_cs.size() * sizeof(double));
    if (tri_area) std::memcpy(tri_area, s.h_tri_area.data(), s.h_tri_area.size() * sizeof(double));
    return 1;
}
int bib_ib_particles_add(bib_domain* d, int type, int shape_id, const double* params, int n, const double* positions) {
    if (d && d->is_head && params && (n <= 0 || positions)) return group_particles_add(d, type, shape_id, params, n, positions);
    BIB_ENTER(d);
    if (!params || (n > 0 && !positions)) { bib_set_error("particles_add: null argument"); return 0; }
    return ib_particles_add(d, type, shape_id, params, n, positions);
}
int bib_ib_rigid_particles_add(bib_domain* d, int shape_id, const double* inertia9, double shape_volume, int n, const double* state17) {
    BIB_HEAD_UNSUPPORTED(d, "rigid_particles_add");
    BIB_ENTER(d);
    if (!inertia9 || (n > 0 && !state17)) { bib_set_error("rigid_particles_add: null argument"); return 0; }
    return ib_rigid_add(d, shape_id, inertia9, shape_volume, n, state17);
}
int bib_ib_rigid_state_download(bib_domain* d, int first_particle, int count, double* state19) {
    BIB_HEAD_UNSUPPORTED(d, "rigid_state_download");
    BIB_ENTER_RO(d);
    if (count > 0 && !state19) { bib_set_error("rigid_state_download: null argument"); return 0; }
    return ib_rigid_download(d, first_particle, count, state19);
}
int bib_ib_rigid_state_upload(bib_domain* d, int first_particle, int count, const double* state19) {
    BIB_HEAD_UNSUPPORTED(d, "rigid_state_upload");
    BIB_ENTER(d);
    if (count > 0 && !state19) { bib_set_error("rigid_state_upload: null argument"); return 0; }
    return ib_rigid_upload(d, first_particle, count, state19);
}
int bib_ib_set_regulator(bib_domain* d, double p, double tx, double ty, double tz) {
    BIB_HEAD(d, group_set_regulator(d, p, tx, ty, tz));
    BIB_ENTER(d);
    d->regulator_p = p; d->regulator_target[0] = tx; d->regulator_target[1] = ty; d->regulator_target[2] = tz;
    return 1;
}
int bib_ib_set_parallel_plates(bib_domain* d, int enabled, double y0, double y1) {
    BIB_HEAD(d, group_set_parallel_plates(d, enabled, y0, y1));
    BIB_ENTER(d);
    if (enabled && !(y1 > y0)) { bib_set_error("set_parallel_plates: need y1 > y0"); return 0; }
    d->plates = enabled ? 1 : 0; d->plate_y0 = y0; d->plate_y1 = y1;
    return 1;
}
int bib_ib_set_pipe(bib_domain* d, int enabled, double y0, double z0, double radius) {
    BIB_HEAD(d, group_set_pipe(d, enabled, y0, z0, radius));
    BIB_ENTER(d);
    if (enabled && !(radius > 0)) { bib_set_error("set_pipe: need radius > 0"); return 0; }
    d->plates = enabled ? 2 : 0; d->plate_y0 = y0; d->plate_y1 = z0; d->pipe_radius = radius;
    return 1;
}
static int set_potential(PotentialDev& q, int type, double r_min, double beta, double strength, const char* who) {
    if (type != BIB_POTENTIAL_NONE && type != BIB_POTENTIAL_MORSE && type != BIB_POTENTIAL_SPRING) { bib_set_error(std::string(who) + ": unknown potential type"); return 0; }
    q.type = type; q.r_min = r_min; q.beta = beta; q.strength = strength;
    return 1;
}
int bib_ib_set_wall_interaction(bib_domain* d, int type, double r_min, double beta, double strength) {
    BIB_HEAD(d, group_set_wall_interaction(d, type, r_min, beta, strength));
    BIB_ENTER(d);
    return set_potential(d->wall_pot, type, r_min, beta, strength, "set_wall_interaction");
}
int bib_ib_set_wall_interaction_boundary(bib_domain* d, int kind, double p0, double p1, double p2) {
    BIB_HEAD(d, group_set_wall_interaction_boundary(d, kind, p0, p1, p2));
    BIB_ENTER(d);
    if (kind < -1 || kind > 2) { bib_set_error("set_wall_interaction_boundary: kind must be -1 (boundary of set_boundary), 0, 1 (plates y0,y1) or 2 (pipe y0,z0,radius)"); return 0; }
    d->wall_bkind = kind; d->wall_b[0] = p0; d->wall_b[1] = p1; d->wall_b[2] = p2;
    return 1;
}
int bib_ib_set_pp_interaction(bib_domain* d, int type, double r_min, double beta, double strength) {
    BIB_HEAD(d, group_set_pp_interaction(d, type, r_min, beta, strength));
    BIB_ENTER(d);
    return set_potential(d->pp_pot, type, r_min, beta, strength, "set_pp_interaction");
}
int bib_ib_count(bib_domain* d, int* n_particles, long long* n_vertices) {
    BIB_HEAD(d, group_count(d, n_particles, n_vertices));
    BIB_ENTER_RO(d);
    if (n_particles) *n_particles = d->n_particles;
    if (n_vertices) *n_vertices = d->n_vertices;
    return 1;
}
int bib_ib_interpolate_velocity(bib_domain* d) {
    BIB_HEAD(d, group_ib_interpolate(d));
    if (d && d->fuse_stage == 0 && fusable(d)) { d->fuse_stage = 1; return 1; }       // deferred (call fusion)
    BIB_ENTER(d);
    // the reference interpolates from the TensorField the driver computed (and may have edited); use it when it exists
    return ib_interpolate(d);
}
int bib_ib_compute_and_spread_forces(bib_domain* d) {
    BIB_HEAD(d, group_ib_forces_spread(d));
    if (d && d->fuse_stage == 2) { d->fuse_stage = 3; return 1; }
    BIB_ENTER(d);
    return ib_forces_and_spread(d);
}
int bib_ib_move_vertices_and_revoxelize(bib_domain* d) {
    BIB_HEAD(d, group_ib_move(d));
    if (d && d->fuse_stage == 3) { d->fuse_stage = 4; return 1; }
    BIB_ENTER(d);
    return ib_move_and_voxelize(d);
}
int bib_ib_step(bib_domain* d) {
    BIB_HEAD(d, group_ib_step(d));
    BIB_ENTER(d);
    return ib_interpolate(d) && ib_forces_and_spread(d) && ib_move_and_voxelize(d);
}
// MultiDirectForcingFunctional3D::process (progs/sphere_in_ext_flow/sphere_in_ext_flow.cpp:26-98; oracle/biblbm_oracle.c orc_mdf_step)
int bib_ib_multi_direct_forcing(bib_domain* d, int relax_iterations) {
    BIB_HEAD_UNSUPPORTED(d, "multi_direct_forcing");
    BIB_ENTER(d);
    if (relax_iterations < 0) { bib_set_error("multi_direct_forcing: negative iteration count"); return 0; }
    if (d->n_ranks > 1) { bib_set_error("multi_direct_forcing: single-slab domains only"); return 0; }
    const int box[6] = {0, d->nx - 1, 0, d->ny - 1, 0, d->nz - 1};
    if (!lbm_launch_set_external(d, box, 0.0, 0.0, 0.0)) return 0;
    d->force_is_body = true; d->force_body[0] = d->force_body[1] = d->force_body[2] = 0.0;
    if (!ib_force_accumulate(d, 0)) return 0;
    for (int it = 0; it < relax_iterations; ++it) {
        if (!lbm_launch_velocity(d, box)) return 0;          // every cell, boundary cells report their stored velocity
        d->vel_full_valid = true;
        if (!ib_interpolate(d) || !ib_forces_and_spread(d) || !ib_force_accumulate(d, 1)) return 0;
    }
    d->vel_full_valid = false;
    return ib_force_accumulate(d, 2) && ib_update_all(d, 1);   // node force := -sum, then move_vertices()
}
int bib_ib_vertices_download(bib_domain* d, long long first_vertex, long long count, double* vertices9) {
    if (d && d->is_head && (count <= 0 || vertices9)) return group_vertices_download(d, first_vertex, count, vertices9);
    BIB_ENTER_RO(d);
    if (count > 0 && !vertices9) { bib_set_error("vertices_download: null argument"); return 0; }
    return ib_vertices_download(d, first_vertex, count, vertices9);
}
int bib_ib_vertices_upload(bib_domain* d, long long first_vertex, long long count, const double* vertices9) {
    if (d && d->is_head && (count <= 0 || vertices9)) return group_vertices_upload(d, first_vertex, count, vertices9);
    BIB_ENTER(d);
    if (count > 0 && !vertices9) { bib_set_error("vertices_upload: null argument"); return 0; }
    return ib_vertices_upload(d, first_vertex, count, vertices9);
}
int bib_ib_vertex_forces_preload(bib_domain* d, long long first_vertex, long long count, const double* forces3) {
    BIB_HEAD_UNSUPPORTED(d, "vertex_forces_preload");
    BIB_ENTER(d);
    if (count > 0 && !forces3) { bib_set_error("vertex_forces_preload: null argument"); return 0; }
    if (d->n_ranks > 1) { bib_set_error("vertex_forces_preload: single-slab domains only"); return 0; }
    return ib_forces_preload(d, first_vertex, count, forces3);
}
int bib_ib_local_slots(bib_domain* d, int* global_ids, int* roles, int capacity, int* n_slots) {
    if (d && d->is_head && global_ids && roles && n_slots) return group_local_slots(d, global_ids, roles, capacity, n_slots);
    BIB_ENTER_RO(d);
    if (!global_ids || !roles || !n_slots) { bib_set_error("local_slots: null argument"); return 0; }
    return ib_local_slots(d, global_ids, roles, capacity, n_slots);
}
int bib_ib_particle_scalars_download(bib_domain* d, int first_particle, int count, double* scalars11) {
    if (d && d->is_head && (count <= 0 || scalars11)) return group_scalars_download(d, first_particle, count, scalars11);
    BIB_ENTER_RO(d);
    if (count > 0 && !scalars11) { bib_set_error("particle_scalars_download: null argument"); return 0; }
    return ib_scalars_download(d, first_particle, count, scalars11);
}
int bib_ib_particle_scalars_upload(bib_domain* d, int first_particle, int count, const double* scalars11) {
    if (d && d->is_head && (count <= 0 || scalars11)) return group_scalars_upload(d, first_particle, count, scalars11);
    BIB_ENTER(d);
    if (count > 0 && !scalars11) { bib_set_error("particle_scalars_upload: null argument"); return 0; }
    if (d->n_ranks > 1 && !d->group) { bib_set_error("particle_scalars_upload: single-slab domains only"); return 0; }
    return ib_scalars_upload(d, first_particle, count, scalars11);
}
int bib_ib_results_download_async(bib_domain* d, double* vertices9, double* scalars11) {
    BIB_HEAD_UNSUPPORTED(d, "results_download_async (use bib_ib_vertices_download / bib_ib_particle_scalars_download, or the per-slab handles of bib_domain_slab)");
    BIB_ENTER_RO(d);
    if (!vertices9 || !scalars11) { bib_set_error("results_download_async: null argument"); return 0; }
    return ib_results_download_async(d, vertices9, scalars11);
}
int bib_ib_lightweight_download_async(bib_domain* d, double* records11) {
    BIB_HEAD_UNSUPPORTED(d, "lightweight_download_async (use the per-slab handles of bib_domain_slab)");
    BIB_ENTER_RO(d);
    if (!records11) { bib_set_error("lightweight_download_async: null argument"); return 0; }
    return ib_results_download_async(d, nullptr, records11);
}
int bib_ib_results_wait(bib_domain* d) { BIB_HEAD_UNSUPPORTED(d, "results_wait"); BIB_ENTER_RO(d); return ib_results_wait(d); }
int bib_ib_results_count(bib_domain* d, int* n_particles) {
    BIB_HEAD_UNSUPPORTED(d, "results_count");
    BIB_ENTER_RO(d);
    if (!n_particles) { bib_set_error("results_count: null argument"); return 0; }
    return ib_results_count(d, n_particles);
}

// One driver step: the velocity the IB interpolates is u = j(f^t)/rho + F^{t-1}/2 evaluated on demand from the current
// populations and the force field still holding last step's total force, so the dense
// computeVelocity pass of the reference is not needed; then F := body, IB, collide-and-stream.
static int step_once(bib_domain* d, double body_fx, double body_fy, double body_fz) {
    const int box[6] = {0, d->nx - 1, 0, d->ny - 1, 0, d->nz - 1};
    // The vertex forces need the positions only. (a) BIB_OPT_PREFETCH_FORCES: those of step t+1 are evaluated on the second stream
    // (FP64-bound) while the collide-and-stream of step t (HBM-bound) runs, into a second force array, so that vfrc keeps holding
    // "the forces of the last completed step" for every reader. (b) Otherwise / for the first step: on the second stream beside the
    // band velocity and the interpolation of the same step, joined before the spread.
    const bool independent = ib_forces_independent_of_interpolation(d) && !d->pt_on;      // phase timing: every phase on the compute stream
    const bool prefetch = d->prefetch_forces && d->overlap_forces && independent;
    pt_mark(d, PT_BEGIN);
    const bool have = prefetch && d->forces_ready && d->vfrc_next && d->vfrc_next_cap == d->cap_vertices;
    d->forces_ready = false;
    if (have) std::swap(d->vfrc, d->vfrc_next);
    const bool fork = !have && d->overlap_forces && independent;
    if (fork) {
        BIB_CHECK(cudaEventRecord(d->ev_fork, d->stream));
        BIB_CHECK(cudaStreamWaitEvent(d->comm_stream, d->ev_fork, 0));
        if (!ib_forces(d, d->comm_stream)) return 0;
        BIB_CHECK(cudaEventRecord(d->ev_forces, d->comm_stream));
    }
    d->defer_vel = d->ms != nullptr;           // multi-slab: one neighbour sync for VEL + FORCE (the forces in between do not read the velocities)
    const int ok_interp = ib_interpolate(d);
    d->defer_vel = false;
    if (!ok_interp) return 0;
    pt_mark(d, PT_X_VEL);
    // setExternalVector(body): the only cells that differ from the body force are those the last spread touched
    if (d->force_is_body && d->force_body[0] == body_fx && d->force_body[1] == body_fy && d->force_body[2] == body_fz) {
        if (!ib_band_reset(d, body_fx, body_fy, body_fz)) return 0;
    } else {
        if (!lbm_launch_set_external(d, box, body_fx, body_fy, body_fz)) return 0;
        d->force_is_body = true; d->force_body[0] = body_fx; d->force_body[1] = body_fy; d->force_body[2] = body_fz;
    }
    pt_mark(d, PT_RESET);
    if (have) {
        if (!ib_spread(d)) return 0;
    } else if (fork) {
        BIB_CHECK(cudaStreamWaitEvent(d->stream, d->ev_forces, 0));
        if (!ib_spread(d)) return 0;
    } else if (!ib_forces_and_spread(d)) return 0;
    if (!ib_move_and_voxelize(d)) return 0;
    pt_mark(d, PT_VOXELIZE);
    d->vel_full_valid = false;
    if (prefetch) {
        cudaStreamCaptureStatus cap_status = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(d->stream, &cap_status);
        if (d->vfrc_next_cap != d->cap_vertices && cap_status == cudaStreamCaptureStatusNone) {      // never allocate inside a capture
            if (d->vfrc_next) { cudaFree(d->vfrc_next); d->vfrc_next = nullptr; d->vfrc_next_cap = 0; }
            if (cudaMalloc(&d->vfrc_next, (size_t)d->cap_vertices * 3 * sizeof(double)) == cudaSuccess) d->vfrc_next_cap = d->cap_vertices;
            else { cudaGetLastError(); d->vfrc_next = nullptr; }
        }
        if (d->vfrc_next && d->vfrc_next_cap == d->cap_vertices) {
            BIB_CHECK(cudaEventRecord(d->ev_fork, d->stream));
            BIB_CHECK(cudaStreamWaitEvent(d->comm_stream, d->ev_fork, 0));
            std::swap(d->vfrc, d->vfrc_next);                 // ib_forces writes d->vfrc: aim it at the second array for this call
            const int ok = ib_forces(d, d->comm_stream);
            std::swap(d->vfrc, d->vfrc_next);
            if (!ok) return 0;
            BIB_CHECK(cudaEventRecord(d->ev_forces, d->comm_stream));
            if (!lbm_launch_collide_stream(d)) return 0;
            BIB_CHECK(cudaStreamWaitEvent(d->stream, d->ev_forces, 0));
            d->forces_ready = true;
            return 1;
        }
    }
    return lbm_launch_collide_stream(d);
}

// Small lattices (the reference's own driver configurations: 61^3 ... 97^3 cells, one to a few dozen particles) are launch-bound:
// ~15 kernels of a few microseconds each per step. In steady state the kernel sequence of a step is fixed and all its arguments
// repeat with period 2 (A/B population buffers, double-buffered band masks, the two vertex-force arrays of the prefetch), so the
// two steps of a period are captured once, EACH into its own CUDA graph, and replayed alternately: a caller that issues one step
// per call (a driver's time loop, bench.py's e2e leg) replays graphs as well. Any configuration-changing entry point invalidates
// the graphs (BIB_ENTER). Single-slab domains only (the multi-slab step interleaves NCCL calls on two streams).
static void step_graphs_destroy(bib_domain* d) {
    for (StepGraph& g : d->sg) { if (g.exec) cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
}
static bool step_graph_matches(const bib_domain* d, const StepGraph& g) {
    return g.exec && d->cur == g.cur && d->band_cur == g.band_cur && d->spread_band == g.spread_band && d->vfrc == g.vfrc && d->forces_ready == g.forces_ready;
}
int bib_step(bib_domain* d, double body_fx, double body_fy, double body_fz, int n_steps) {
    BIB_HEAD(d, group_step(d, body_fx, body_fy, body_fz, n_steps));
    BIB_ENTER_RO(d);
    return step_n(d, body_fx, body_fy, body_fz, n_steps);
}
static int step_n(bib_domain* d, double body_fx, double body_fy, double body_fz, int n_steps) {
    int s = 0;
    const bool same_body = d->graph_body[0] == body_fx && d->graph_body[1] == body_fy && d->graph_body[2] == body_fz;
    const bool have_graphs = d->sg[0].exec && d->sg[1].exec && !d->graph_dirty && same_body;
    if (d->last_body[0] != body_fx || d->last_body[1] != body_fy || d->last_body[2] != body_fz) {
        d->plain_steps_since_dirty = 0;
        d->last_body[0] = body_fx; d->last_body[1] = body_fy; d->last_body[2] = body_fz;
    }
    // capture when this call is long enough to pay for it, or when a one-step-per-call caller has kept the configuration unchanged for a while
    if (d->use_graph && !d->pt_on && d->n_ranks == 1 && !have_graphs && (n_steps >= 6 || d->plain_steps_since_dirty >= 4)) {
        step_graphs_destroy(d);
        for (; d->plain_steps_since_dirty < 2 && s < n_steps; ++s) {          // reach the steady state first (allocations, masks, flags)
            if (!step_once(d, body_fx, body_fy, body_fz)) return 0;
            d->plain_steps_since_dirty++;
        }
        if (d->plain_steps_since_dirty >= 2 && s < n_steps) {
            // host-side state the captured (not executed) steps advance: restored if the capture does not succeed
            struct { int cur, band_cur, spread_band; bool band_valid[2], band_pos_valid, forces_ready, vel_full_valid, force_is_body, fflag_valid; double* vfrc; double* vfrc_next; long long launches; } snap =
                {d->cur, d->band_cur, d->spread_band, {d->band_valid[0], d->band_valid[1]}, d->band_pos_valid, d->forces_ready, d->vel_full_valid, d->force_is_body, d->fflag_valid, d->vfrc, d->vfrc_next, d->launches};
            bool ok = true;
            for (int k = 0; k < 2 && ok; ++k) {
                StepGraph& g = d->sg[k];
                g.cur = d->cur; g.band_cur = d->band_cur; g.spread_band = d->spread_band; g.vfrc = d->vfrc; g.forces_ready = d->forces_ready;   // the state this step starts from
                const long long l0 = d->launches;
                cudaGraph_t graph = nullptr;
                ok = cudaStreamBeginCapture(d->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
                if (ok) {
                    const bool stepped = step_once(d, body_fx, body_fy, body_fz) != 0;
                    const cudaError_t e = cudaStreamEndCapture(d->stream, &graph);
                    g.launches = d->launches - l0;
                    d->launches = l0;                              // captured, not executed
                    ok = stepped && e == cudaSuccess && graph && cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess;
                }
                if (graph) cudaGraphDestroy(graph);
            }
            // period 2: after the two captured steps the host-side state must be the one the first graph starts from
            if (ok && step_graph_matches(d, d->sg[0])) {
                d->graph_dirty = false;
                d->graph_body[0] = body_fx; d->graph_body[1] = body_fy; d->graph_body[2] = body_fz;
            } else {
                step_graphs_destroy(d);
                d->use_graph = 0;                                  // something in this configuration is not capturable: plain launches from now on
                d->cur = snap.cur; d->band_cur = snap.band_cur; d->spread_band = snap.spread_band; d->band_valid[0] = snap.band_valid[0]; d->band_valid[1] = snap.band_valid[1];
                d->band_pos_valid = snap.band_pos_valid; d->forces_ready = snap.forces_ready; d->vel_full_valid = snap.vel_full_valid;
                d->force_is_body = snap.force_is_body; d->fflag_valid = snap.fflag_valid; d->vfrc = snap.vfrc; d->vfrc_next = snap.vfrc_next; d->launches = snap.launches;
            }
            cudaGetLastError();
        }
    }
    if (d->use_graph && !d->pt_on && d->n_ranks == 1 && d->sg[0].exec && d->sg[1].exec && !d->graph_dirty &&
        d->graph_body[0] == body_fx && d->graph_body[1] == body_fy && d->graph_body[2] == body_fz) {
        for (; s < n_steps; ++s) {
            const int k = step_graph_matches(d, d->sg[0]) ? 0 : (step_graph_matches(d, d->sg[1]) ? 1 : -1);
            if (k < 0) break;
            BIB_CHECK(cudaGraphLaunch(d->sg[k].exec, d->stream));
            d->launches += d->sg[k].launches;
            // the host-side state one step later = the state the other graph starts from
            const StepGraph& nx = d->sg[1 - k];
            d->cur = nx.cur; d->band_cur = nx.band_cur; d->spread_band = nx.spread_band; d->forces_ready = nx.forces_ready;
            if (d->vfrc != nx.vfrc) std::swap(d->vfrc, d->vfrc_next);
        }
    }
    for (; s < n_steps; ++s) { if (!step_once(d, body_fx, body_fy, body_fz)) return 0; d->plain_steps_since_dirty++; }
    return 1;
}

int bib_comm_unique_id(void* unique_id_128) {
    if (!unique_id_128) { bib_set_error("comm_unique_id: null argument"); return 0; }
    return comm_unique_id(unique_id_128);
}
int bib_comm_init(bib_domain* d, int rank, int n_ranks, const void* unique_id_128) {
    BIB_HEAD_UNSUPPORTED(d, "comm_init");
    BIB_ENTER(d);
    if (!unique_id_128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) { bib_set_error("comm_init: bad arguments"); return 0; }
    if (d->n_particles > 0) { bib_set_error("comm_init: call bib_comm_init before adding particles"); return 0; }
    return comm_init(d, rank, n_ranks, unique_id_128);
}

int bib_sync(bib_domain* d) {
    BIB_HEAD(d, group_sync(d));
    BIB_ENTER_RO(d);
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    BIB_CHECK(cudaStreamSynchronize(d->comm_stream));
    return ms_check_overflow(d);
}
int bib_time_begin(bib_domain* d) { BIB_HEAD(d, group_time_begin(d)); BIB_ENTER_RO(d); BIB_CHECK(cudaEventRecord(d->ev0, d->stream)); return 1; }
int bib_time_end(bib_domain* d, float* milliseconds) {
    if (d && d->is_head && milliseconds) return group_time_end(d, milliseconds);
    BIB_ENTER_RO(d);
    if (!milliseconds) { bib_set_error("time_end: null argument"); return 0; }
    BIB_CHECK(cudaEventRecord(d->ev1, d->stream));
    BIB_CHECK(cudaEventSynchronize(d->ev1));
    BIB_CHECK(cudaEventElapsedTime(milliseconds, d->ev0, d->ev1));
    return 1;
}
int bib_phase_timing(bib_domain* d, int enable) {
    if (d && d->is_head) return group_phase_timing(d, enable);
    BIB_ENTER(d);
    d->pt_on = enable != 0;
    for (auto* v : {&d->pt_marks, &d->pt_comm_marks}) { for (auto& m : *v) d->pt_pool.push_back(m.second); v->clear(); }
    return 1;
}
int bib_phase_times(bib_domain* d, double* milliseconds, int n) {
    if (d && d->is_head && milliseconds) return group_phase_times(d, milliseconds, n);
    BIB_ENTER_RO(d);
    if (!milliseconds || n < PT_COUNT) { bib_set_error("phase_times: need room for 14 doubles"); return 0; }
    for (int k = 0; k < n; ++k) milliseconds[k] = 0.0;
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    BIB_CHECK(cudaStreamSynchronize(d->comm_stream));
    for (auto* v : {&d->pt_marks, &d->pt_comm_marks}) {
        for (size_t k = 1; k < v->size(); ++k) {
            const int ph = (*v)[k].first;
            if (ph < 0) continue;                       // a BEGIN mark: nothing is attributed to the gap before it
            float t = 0.f;
            if (cudaEventElapsedTime(&t, (*v)[k - 1].second, (*v)[k].second) == cudaSuccess) milliseconds[ph] += t;
        }
        for (auto& m : *v) d->pt_pool.push_back(m.second);
        v->clear();
    }
    cudaGetLastError();
    return 1;
}
int bib_kernel_launch_count(bib_domain* d, long long* launches) {
    if (d && d->is_head && launches) return group_launch_count(d, launches);
    BIB_ENTER_RO(d);
    if (!launches) { bib_set_error("kernel_launch_count: null argument"); return 0; }
    *launches = d->launches;
    return 1;
}
int bib_device_memory_bytes(bib_domain* d, long long* bytes) {
    if (d && d->is_head && bytes) return group_memory_bytes(d, bytes);
    BIB_ENTER_RO(d);
    if (!bytes) { bib_set_error("device_memory_bytes: null argument"); return 0; }
    *bytes = d->bytes;
    return 1;
}

}  // extern "C"
