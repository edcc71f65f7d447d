// group.cu -- one host process owning several x-slabs (bib_domain_create_multi): the in-process counterpart of the reference's
// `mpirun -np N` run (one block per rank, src/fsi/ImmersedBoundaryDynamics.hh:49-54; progs/rbc_in_pipe/batch.sh).
//
// The handle returned to the caller (the "head") looks like any other bib_domain: boxes, particle ids and vertex ranges are GLOBAL,
// every entry point of the C-ABI is split over / broadcast to the slabs here. Collective work (stepping, the IB phases, the halo
// exchange) runs SPMD on one worker thread per slab, exactly the code path of a multi-process run; only the transport differs
// (comm.cu: peer copies ordered by CUDA events + the host barrier below instead of NCCL). Slabs may share a device (tests on a
// one-GPU box exercise the whole multi-slab protocol that way).
#include "../../include/biblbm_b200.h"
#include "domain.cuh"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

struct bib_group {
    int n = 0;
    std::vector<bib_domain*> slab;
    int gnx = 0, ny = 0, nz = 0, n_ext = 4;
    std::vector<int> x0;                   // [n + 1] first global plane of every slab (widths differ by at most one plane)
    bool same_device = true;
    // worker threads (one per slab) + job hand-over
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::function<int(int)> job;
    unsigned long long job_id = 0;
    int pending = 0;
    std::vector<int> result;
    std::vector<std::string> error;
    bool quit = false;
    // sense-reversing barrier for the exchanges (called from the worker threads)
    std::atomic<int> bar_count{0};
    std::atomic<int> bar_sense{0};
    // global particle bookkeeping of the head
    int n_particles = 0; long long n_vertices = 0;
    std::vector<int> group_nv;             // vertices per particle of each added particle group
    std::vector<int> group_first, group_count;
    std::vector<long long> group_first_vertex;
};

bib_domain* group_slab(bib_group* g, int rank) { return g->slab[rank]; }

void group_barrier(bib_group* g) {
    const int sense = g->bar_sense.load(std::memory_order_acquire);
    if (g->bar_count.fetch_add(1, std::memory_order_acq_rel) == g->n - 1) {
        g->bar_count.store(0, std::memory_order_relaxed);
        g->bar_sense.store(1 - sense, std::memory_order_release);
    } else {
        int spins = 0;
        while (g->bar_sense.load(std::memory_order_acquire) == sense) { if (++spins > 2000) std::this_thread::yield(); }
    }
}

namespace {

void worker_main(bib_group* g, int r) {
    unsigned long long seen = 0;
    for (;;) {
        std::function<int(int)> job;
        {
            std::unique_lock<std::mutex> lk(g->mu);
            g->cv_job.wait(lk, [&] { return g->quit || g->job_id != seen; });
            if (g->quit) return;
            seen = g->job_id;
            job = g->job;
        }
        const int ok = job(r);
        const std::string err = ok ? std::string() : std::string(bib_last_error());
        {
            std::lock_guard<std::mutex> lk(g->mu);
            g->result[r] = ok; g->error[r] = err;
            if (--g->pending == 0) g->cv_done.notify_all();
        }
    }
}

// run fn(rank) on every slab's worker thread at once (SPMD); 1 if all succeeded, else 0 with the first failing slab's message
int group_run(bib_group* g, const std::function<int(int)>& fn) {
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->job = fn; g->pending = g->n; ++g->job_id;
    }
    g->cv_job.notify_all();
    std::unique_lock<std::mutex> lk(g->mu);
    g->cv_done.wait(lk, [&] { return g->pending == 0; });
    for (int r = 0; r < g->n; ++r) if (!g->result[r]) { bib_set_error("slab " + std::to_string(r) + ": " + g->error[r]); return 0; }
    return 1;
}
// run fn(rank) slab after slab on the calling thread (no exchange inside)
int group_each(bib_group* g, const std::function<int(int)>& fn) {
    for (int r = 0; r < g->n; ++r) if (!fn(r)) return 0;
    return 1;
}

// the part of a global inclusive box that lies in slab r, in slab-local coordinates; false if empty
bool split_box(const bib_group* g, const int box[6], int r, int local[6], int* x_skip) {
    const int lo = g->x0[r], hi = g->x0[r + 1] - 1;
    const int x0 = std::max(box[0], lo), x1 = std::min(box[1], hi);
    if (x0 > x1) return false;
    local[0] = x0 - lo; local[1] = x1 - lo;
    for (int k = 2; k < 6; ++k) local[k] = box[k];
    *x_skip = x0 - box[0];
    return true;
}
bool box_in_lattice(const bib_group* g, const int box[6]) {
    return box[0] >= 0 && box[1] < g->gnx && box[0] <= box[1] && box[2] >= 0 && box[3] < g->ny && box[2] <= box[3] && box[4] >= 0 && box[5] < g->nz && box[4] <= box[5];
}
size_t yz_cells(const int box[6]) { return (size_t)(box[3] - box[2] + 1) * (box[5] - box[4] + 1); }

}  // namespace

// ---------------------------------------------------------------------------------------------------------------- life cycle
int group_create(int nx, int ny, int nz, int n_ext, double omega_outer, double omega_inner, const int periodic[3], int n_slabs, const int* devices,
                 bib_domain** out) {
    *out = nullptr;
    if (n_slabs < 1 || n_slabs > 64) { bib_set_error("domain_create_multi: 1 .. 64 slabs"); return 0; }
    if (nx / n_slabs < 3) { bib_set_error("domain_create_multi: fewer than 3 planes per slab"); return 0; }
    bib_group* g = new bib_group();
    g->n = n_slabs; g->gnx = nx; g->ny = ny; g->nz = nz; g->n_ext = n_ext;
    for (int r = 0; r <= n_slabs; ++r) g->x0.push_back((int)((long long)r * nx / n_slabs));
    g->slab.assign(n_slabs, nullptr); g->result.assign(n_slabs, 0); g->error.assign(n_slabs, std::string());
    bool ok = true;
    for (int r = 0; ok && r < n_slabs; ++r) {
        const int dev = devices ? devices[r] : r;
        if (dev != (devices ? devices[0] : 0)) g->same_device = false;
        ok = bib_domain_create(g->x0[r + 1] - g->x0[r], ny, nz, n_ext, omega_outer, omega_inner, periodic, dev, &g->slab[r]) != 0;
        if (!ok) break;
        bib_domain* d = g->slab[r];
        d->group = g; d->rank = r; d->n_ranks = n_slabs; d->gnx = nx; d->x_offset = g->x0[r];
        ok = bib_cuda_ok(cudaEventCreateWithFlags(&d->ev_x_ready, cudaEventDisableTiming), "cudaEventCreate") &&
             bib_cuda_ok(cudaEventCreateWithFlags(&d->ev_x_done, cudaEventDisableTiming), "cudaEventCreate") &&
             bib_cuda_ok(cudaEventCreateWithFlags(&d->ev_h_ready, cudaEventDisableTiming), "cudaEventCreate") &&
             bib_cuda_ok(cudaEventCreateWithFlags(&d->ev_h_done, cudaEventDisableTiming), "cudaEventCreate");
    }
    // peer access between neighbouring devices (NVLink): needed by cudaMemcpyPeerAsync's direct path and by kernels reading peer buffers
    for (int r = 0; ok && r < n_slabs; ++r) {
        const int a = g->slab[r]->device;
        for (int nb : {(r + 1) % n_slabs, (r + n_slabs - 1) % n_slabs}) {
            const int b = g->slab[nb]->device;
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, a, b);
            if (can) { cudaSetDevice(a); const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = bib_cuda_ok(e, "cudaDeviceEnablePeerAccess"); cudaGetLastError(); }
            else g->slab[r]->halo_peer_ok = false;      // the population halo falls back to cudaMemcpyPeerAsync (staged through the host by the driver)
        }
    }
    if (!ok) {
        const std::string err = bib_last_error();
        for (bib_domain* d : g->slab) if (d) { d->group = nullptr; d->n_ranks = 1; bib_domain_destroy(d); }
        delete g;
        bib_set_error(err);
        return 0;
    }
    for (int r = 0; r < n_slabs; ++r) g->workers.emplace_back(worker_main, g, r);
    bib_domain* head = new bib_domain();
    head->is_head = true; head->group = g; head->device = g->slab[0]->device;
    head->nx = nx; head->ny = ny; head->nz = nz; head->gnx = nx; head->n_ext = n_ext; head->n_ranks = n_slabs;
    for (int k = 0; k < 3; ++k) head->per[k] = periodic[k] ? 1 : 0;
    head->omega0 = omega_outer; head->omega1 = n_ext == 3 ? 0.0 : omega_inner;
    *out = head;
    return 1;
}

int group_destroy(bib_domain* head) {
    bib_group* g = head->group;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->quit = true;
    }
    g->cv_job.notify_all();
    for (std::thread& t : g->workers) t.join();
    for (bib_domain* d : g->slab) {
        if (!d) continue;
        cudaSetDevice(d->device);
        for (cudaEvent_t e : {d->ev_x_ready, d->ev_x_done, d->ev_h_ready, d->ev_h_done}) if (e) cudaEventDestroy(e);
        d->group = nullptr;
        bib_domain_destroy(d);
    }
    delete g;
    delete head;
    return 1;
}

// ---------------------------------------------------------------------------------------------------------------- dispatch
// Every function below is the head's version of the C-ABI entry point of the same name (api.cu forwards when d->is_head).
#define G(head) bib_group* g = (head)->group

int group_set_option(bib_domain* h, int key, double value) { G(h); return group_each(g, [&](int r) { return bib_domain_set_option(g->slab[r], key, value); }); }

int group_upload(bib_domain* h, const int box[6], const double* cells) {
    G(h);
    if (!box_in_lattice(g, box)) { bib_set_error("upload: box outside the lattice"); return 0; }
    const size_t per_x = yz_cells(box) * (19 + g->n_ext);
    return group_each(g, [&](int r) { int lb[6], skip; if (!split_box(g, box, r, lb, &skip)) return 1; return bib_domain_upload(g->slab[r], lb, cells + (size_t)skip * per_x); });
}
int group_download(bib_domain* h, const int box[6], double* cells) {
    G(h);
    if (!box_in_lattice(g, box)) { bib_set_error("download: box outside the lattice"); return 0; }
    const size_t per_x = yz_cells(box) * (19 + g->n_ext);
    return group_each(g, [&](int r) { int lb[6], skip; if (!split_box(g, box, r, lb, &skip)) return 1; return bib_domain_download(g->slab[r], lb, cells + (size_t)skip * per_x); });
}
int group_set_celltype(bib_domain* h, const int box[6], int kind) {
    G(h);
    if (!box_in_lattice(g, box)) { bib_set_error("set_celltype: box outside the lattice"); return 0; }
    return group_each(g, [&](int r) { int lb[6], skip; if (!split_box(g, box, r, lb, &skip)) return 1; return bib_domain_set_celltype(g->slab[r], lb, kind); });
}
int group_init_equilibrium(bib_domain* h, const int box[6], double rho, double ux, double uy, double uz) {
    G(h);
    if (!box_in_lattice(g, box)) { bib_set_error("init_equilibrium: box outside the lattice"); return 0; }
    return group_each(g, [&](int r) { int lb[6], skip; if (!split_box(g, box, r, lb, &skip)) return 1; return bib_domain_init_equilibrium(g->slab[r], lb, rho, ux, uy, uz); });
}
int group_add_velocity_boundary(bib_domain* h, const int box[6], int direction, int orientation) {
    G(h);
    if (!box_in_lattice(g, box)) { bib_set_error("add_velocity_boundary: box outside the lattice"); return 0; }
    if (direction == 0) { bib_set_error("add_velocity_boundary: on multi-slab domains only planes normal to y / z"); return 0; }
    if (box[0] != 0 || box[1] != g->gnx - 1) { bib_set_error("add_velocity_boundary: the plane must span x on a multi-slab domain"); return 0; }
    return group_each(g, [&](int r) { int lb[6], skip; if (!split_box(g, box, r, lb, &skip)) return 1; return bib_add_velocity_boundary(g->slab[r], lb, direction, orientation); });
}
int lbm_set_wall_velocity_ext(bib_domain* d, const int box[6], const double* vel, int gx0, int gnx_wrap);
int group_set_boundary_velocity(bib_domain* h, const int box[6], const double* velocities) {
    G(h);
    if (!box_in_lattice(g, box)) { bib_set_error("set_boundary_velocity: box outside the lattice"); return 0; }
    // every slab also stores the velocities of the wall cells just outside its x-range (ghost columns): the plane functional
    // differentiates the stored velocities along x (PlaneFdBoundaryFunctional3D, finiteDifferenceBoundaryProcessor3D.hh:52-115)
    return group_each(g, [&](int r) { return lbm_set_wall_velocity_ext(g->slab[r], box, velocities, g->x0[r], h->per[0] ? g->gnx : 0); });
}
int group_set_guo_wall_nodes(bib_domain* h, int n, const int* location_direction6, const double* delta, const double* wall_velocity3) {
    G(h);
    return group_each(g, [&](int r) { return bib_set_guo_wall_nodes(g->slab[r], n, location_direction6, delta, wall_velocity3); });
}
int group_set_external_vector(bib_domain* h, const int box[6], double fx, double fy, double fz) {
    G(h);
    if (!box_in_lattice(g, box)) { bib_set_error("set_external_vector: box outside the lattice"); return 0; }
    return group_each(g, [&](int r) { int lb[6], skip; if (!split_box(g, box, r, lb, &skip)) return 1; return bib_set_external_vector(g->slab[r], lb, fx, fy, fz); });
}
int group_compute_velocity(bib_domain* h, const int box[6]) {
    G(h);
    if (!box_in_lattice(g, box)) { bib_set_error("compute_velocity: box outside the lattice"); return 0; }
    return group_each(g, [&](int r) { int lb[6], skip; if (!split_box(g, box, r, lb, &skip)) return 1; return bib_compute_velocity(g->slab[r], lb); });
}
int group_velocity_download(bib_domain* h, const int box[6], double* velocity) {
    G(h);
    if (!box_in_lattice(g, box)) { bib_set_error("velocity_download: box outside the lattice"); return 0; }
    const size_t per_x = yz_cells(box) * 3;
    return group_each(g, [&](int r) { int lb[6], skip; if (!split_box(g, box, r, lb, &skip)) return 1; return bib_velocity_download(g->slab[r], lb, velocity + (size_t)skip * per_x); });
}
int group_velocity_upload(bib_domain* h, const int box[6], const double* velocity) {
    G(h);
    if (!box_in_lattice(g, box)) { bib_set_error("velocity_upload: box outside the lattice"); return 0; }
    const size_t per_x = yz_cells(box) * 3;
    return group_each(g, [&](int r) { int lb[6], skip; if (!split_box(g, box, r, lb, &skip)) return 1; return bib_velocity_upload(g->slab[r], lb, velocity + (size_t)skip * per_x); });
}
int group_velocity_extrapolate_plane(bib_domain* h, int direction, int coord, int inward) {
    G(h);
    if (direction == 0) { bib_set_error("velocity_extrapolate_plane: planes normal to x on single-slab domains only"); return 0; }
    return group_each(g, [&](int r) { return bib_velocity_extrapolate_plane(g->slab[r], direction, coord, inward); });
}
int group_collide_and_stream(bib_domain* h) { G(h); return group_run(g, [&](int r) { return bib_collide_and_stream(g->slab[r]); }); }

int group_shape_register(bib_domain* h, int nv, int nt, const double* vertices, const int* triangles, int* shape_id) {
    G(h);
    int id0 = -1;
    const int ok = group_each(g, [&](int r) { int id = -1; if (!bib_ib_shape_register(g->slab[r], nv, nt, vertices, triangles, &id)) return 0; if (r == 0) id0 = id; return id == id0 ? 1 : 0; });
    *shape_id = id0;
    return ok;
}
int group_particles_add(bib_domain* h, int type, int shape_id, const double* params, int n, const double* positions) {
    G(h);
    if (n <= 0) return 1;
    // every slab is given the same global list and keeps what it owns (+ ghosts), like the ranks of the reference (add_particle)
    if (!group_each(g, [&](int r) { return bib_ib_particles_add(g->slab[r], type, shape_id, params, n, positions); })) return 0;
    const ShapeDev& s = g->slab[0]->shapes[shape_id];
    g->group_nv.push_back(s.nv); g->group_first.push_back(g->n_particles); g->group_count.push_back(n); g->group_first_vertex.push_back(g->n_vertices);
    g->n_particles += n; g->n_vertices += (long long)n * s.nv;
    return 1;
}
int group_count(bib_domain* h, int* n_particles, long long* n_vertices) {
    G(h);
    if (n_particles) *n_particles = g->n_particles;
    if (n_vertices) *n_vertices = g->n_vertices;
    return 1;
}

// ---- vertex records / particle scalars in GLOBAL particle order: gathered from (scattered to) the slabs that own each particle
int ms_slot_table(bib_domain* d, std::vector<int>& gid, std::vector<int>& role, std::vector<long long>& first_vertex, std::vector<int>& nv);
namespace {
struct GlobalParticle { int nv; long long first_vertex; };
bool locate_particle(const bib_group* g, long long vertex, int* particle, int* offset) {
    for (size_t k = 0; k < g->group_nv.size(); ++k) {
        const long long v0 = g->group_first_vertex[k], v1 = v0 + (long long)g->group_count[k] * g->group_nv[k];
        if (vertex >= v0 && vertex < v1) { *particle = g->group_first[k] + (int)((vertex - v0) / g->group_nv[k]); *offset = (int)((vertex - v0) % g->group_nv[k]); return true; }
    }
    return false;
}
GlobalParticle particle_info(const bib_group* g, int p) {
    for (size_t k = 0; k < g->group_nv.size(); ++k)
        if (p >= g->group_first[k] && p < g->group_first[k] + g->group_count[k]) return {g->group_nv[k], g->group_first_vertex[k] + (long long)(p - g->group_first[k]) * g->group_nv[k]};
    return {0, 0};
}
}  // namespace

// which = 0: vertex records (9 doubles per vertex, range in vertices), 1: particle scalars (11 per particle, range in particles)
static int group_particle_io(bib_domain* h, int which, bool upload, long long first, long long count, double* data) {
    G(h);
    const long long total = which == 0 ? g->n_vertices : g->n_particles;
    if (first < 0 || count < 0 || first + count > total) { bib_set_error("particle data: range outside the global particle list"); return 0; }
    if (count == 0) return 1;
    int p0, p1, off;
    if (which == 0) { if (!locate_particle(g, first, &p0, &off) || !locate_particle(g, first + count - 1, &p1, &off)) { bib_set_error("particle data: bad range"); return 0; } }
    else { p0 = (int)first; p1 = (int)(first + count - 1); }
    std::vector<char> done(p1 - p0 + 1, 0);
    for (int r = 0; r < g->n; ++r) {
        std::vector<int> gid, role, nv; std::vector<long long> fv;
        if (!ms_slot_table(g->slab[r], gid, role, fv, nv)) return 0;
        for (size_t s = 0; s < gid.size(); ++s) {
            // downloads come from the owner; uploads go to every resident copy (owner and ghosts hold the same positions)
            if (role[s] == 0 || (!upload && role[s] != 1) || gid[s] < p0 || gid[s] > p1) continue;
            const GlobalParticle gp = particle_info(g, gid[s]);
            if (which == 0) {
                const long long a = std::max(first, gp.first_vertex), b = std::min(first + count, gp.first_vertex + gp.nv);
                if (a >= b) continue;
                double* host = data + (size_t)(a - first) * 9;
                const long long lv = fv[s] + (a - gp.first_vertex);
                if (!(upload ? bib_ib_vertices_upload(g->slab[r], lv, b - a, host) : bib_ib_vertices_download(g->slab[r], lv, b - a, host))) return 0;
            } else {
                double* host = data + (size_t)(gid[s] - first) * 11;
                if (!(upload ? bib_ib_particle_scalars_upload(g->slab[r], (int)s, 1, host) : bib_ib_particle_scalars_download(g->slab[r], (int)s, 1, host))) return 0;
            }
            if (role[s] == 1) done[gid[s] - p0] = 1;
        }
    }
    for (char c : done) if (!c) { bib_set_error("particle data: a particle of the range is owned by no slab"); return 0; }
    return 1;
}
int group_vertices_download(bib_domain* h, long long first, long long count, double* v9) { return group_particle_io(h, 0, false, first, count, v9); }
int group_vertices_upload(bib_domain* h, long long first, long long count, const double* v9) {
    if (!group_particle_io(h, 0, true, first, count, const_cast<double*>(v9))) return 0;
    // the POS exchange that refreshes the ghost copies is collective: every slab must take part, not only those that hold the particle
    for (bib_domain* d : h->group->slab) d->ghosts_dirty = true;
    return 1;
}
int group_scalars_download(bib_domain* h, int first, int count, double* s11) { return group_particle_io(h, 1, false, first, count, s11); }
int group_scalars_upload(bib_domain* h, int first, int count, const double* s11) { return group_particle_io(h, 1, true, first, count, const_cast<double*>(s11)); }

int group_local_slots(bib_domain* h, int* global_ids, int* roles, int capacity, int* n_slots) {
    G(h);
    for (int i = 0; i < std::min(capacity, g->n_particles); ++i) { global_ids[i] = i; roles[i] = 1; }
    *n_slots = g->n_particles;
    return 1;
}

int group_set_regulator(bib_domain* h, double p, double tx, double ty, double tz) { G(h); return group_each(g, [&](int r) { return bib_ib_set_regulator(g->slab[r], p, tx, ty, tz); }); }
int group_set_parallel_plates(bib_domain* h, int enabled, double y0, double y1) { G(h); return group_each(g, [&](int r) { return bib_ib_set_parallel_plates(g->slab[r], enabled, y0, y1); }); }
int group_set_pipe(bib_domain* h, int enabled, double y0, double z0, double radius) { G(h); return group_each(g, [&](int r) { return bib_ib_set_pipe(g->slab[r], enabled, y0, z0, radius); }); }
int group_set_wall_interaction(bib_domain* h, int type, double r_min, double beta, double strength) { G(h); return group_each(g, [&](int r) { return bib_ib_set_wall_interaction(g->slab[r], type, r_min, beta, strength); }); }
int group_set_wall_interaction_boundary(bib_domain* h, int kind, double p0, double p1, double p2) { G(h); return group_each(g, [&](int r) { return bib_ib_set_wall_interaction_boundary(g->slab[r], kind, p0, p1, p2); }); }
int group_set_pp_interaction(bib_domain* h, int type, double r_min, double beta, double strength) { G(h); return group_each(g, [&](int r) { return bib_ib_set_pp_interaction(g->slab[r], type, r_min, beta, strength); }); }

int group_ib_interpolate(bib_domain* h) { G(h); return group_run(g, [&](int r) { return bib_ib_interpolate_velocity(g->slab[r]); }); }
int group_ib_forces_spread(bib_domain* h) { G(h); return group_run(g, [&](int r) { return bib_ib_compute_and_spread_forces(g->slab[r]); }); }
int group_ib_move(bib_domain* h) { G(h); return group_run(g, [&](int r) { return bib_ib_move_vertices_and_revoxelize(g->slab[r]); }); }
int group_ib_step(bib_domain* h) { G(h); return group_run(g, [&](int r) { return bib_ib_step(g->slab[r]); }); }
int group_step(bib_domain* h, double fx, double fy, double fz, int n_steps) { G(h); return group_run(g, [&](int r) { return bib_step(g->slab[r], fx, fy, fz, n_steps); }); }
int group_sync(bib_domain* h) { G(h); return group_each(g, [&](int r) { return bib_sync(g->slab[r]); }); }
int group_time_begin(bib_domain* h) { G(h); return group_each(g, [&](int r) { return bib_time_begin(g->slab[r]); }); }
int group_time_end(bib_domain* h, float* ms) {
    G(h);
    float mx = 0.f;
    const int ok = group_each(g, [&](int r) { float t = 0.f; if (!bib_time_end(g->slab[r], &t)) return 0; mx = std::max(mx, t); return 1; });
    *ms = mx;                                  // max over the slabs, like the max over ranks of a multi-process run
    return ok;
}
int group_launch_count(bib_domain* h, long long* launches) {
    G(h);
    long long sum = 0;
    const int ok = group_each(g, [&](int r) { long long l = 0; if (!bib_kernel_launch_count(g->slab[r], &l)) return 0; sum += l; return 1; });
    *launches = sum;
    return ok;
}
int group_memory_bytes(bib_domain* h, long long* bytes) {
    G(h);
    long long sum = 0;
    const int ok = group_each(g, [&](int r) { long long b = 0; if (!bib_device_memory_bytes(g->slab[r], &b)) return 0; sum += b; return 1; });
    *bytes = sum;
    return ok;
}
bib_domain* group_slab_of(bib_domain* h, int r) { G(h); return (r >= 0 && r < g->n) ? g->slab[r] : nullptr; }
int group_size(bib_domain* h) { return h->group->n; }
int group_phase_timing(bib_domain* h, int enable) { G(h); return group_each(g, [&](int r) { return bib_phase_timing(g->slab[r], enable); }); }
int group_phase_times(bib_domain* h, double* ms, int n) {      // max over the slabs, phase by phase
    G(h);
    if (n < PT_COUNT) { bib_set_error("phase_times: need room for 14 doubles"); return 0; }
    for (int k = 0; k < n; ++k) ms[k] = 0.0;
    return group_each(g, [&](int r) { double t[PT_COUNT]; if (!bib_phase_times(g->slab[r], t, PT_COUNT)) return 0; for (int k = 0; k < PT_COUNT; ++k) ms[k] = std::max(ms[k], t[k]); return 1; });
}
