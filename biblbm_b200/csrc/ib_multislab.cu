// ib_multislab.cu -- immersed boundaries across x-slabs: ghost particles, partial-sum exchange, ownership hand-off.
//
// Replaces the reference's ghost-vertex machinery (NonLocalNode lists, three CommunicationBuffer exchanges per step and
// particle hand-off: src/fsi/ImmersedBoundaryDynamics.hh:163-246,300-330,416-450; _interpolation.hh:279-326;
// _forces.hh:378-456) with whole-particle ghosts in fixed-capacity slot arrays, so that every per-step exchange is one
// neighbour exchange of a fixed-size buffer per particle group and no host round trip is needed to learn message sizes.
// Protocol per step:
//   POS   after the owner moved a particle: positions -> both neighbours whose slab its box (+3 cells) reaches; the entry
//         carries a hand-off flag when the centre of mass left the owner's x-range (the owner alone decides);
//   VEL   after interpolation: a ghost holder returns the partial velocity sums of its cells, the owner adds them;
//   PPF   (only with a particle-particle force) a ghost holder returns the pair forces of the ghost's vertices that lie in
//         ITS x-range (every partner vertex within the cut-off is resident there), the owner adds them
//         (reference: ghost forces incl. pp, src/fsi/ImmersedBoundaryDynamics_forces.hh:378-456);
//   FORCE after the owner evaluated membrane forces: forces -> the same neighbours, each slab spreads into its own cells.
// Several particle groups (shapes / parameter sets / particle types) each have their own exchange context and slot range.
#include "domain.cuh"

#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

namespace {

struct MsGeo {
    int nv, cap, xcap, first_slot;
    long long first_vertex;
    size_t entry;
    double my_lo, my_hi, center, GX;    // owned x-range [my_lo, my_hi) in global lattice coordinates
    int per_x, has_left, has_right, two_periodic;
};

__device__ __forceinline__ int d_direction(const MsGeo& g, double com_x) {
    if (com_x >= g.my_lo && com_x < g.my_hi) return 0;
    double dlt = com_x - g.center;
    if (g.per_x) dlt = dlt - floor(dlt / g.GX + 0.5) * g.GX;
    return dlt < 0 ? -1 : 1;
}

enum { MS_POS = 0, MS_VEL = 1, MS_FORCE = 2, MS_PPF = 3 };

// Does the particle's box (+3 cells) reach into the neighbouring slab on `side` (0 = left, 1 = right)? Periodic images are tested: after
// a periodic remap of the centre of mass (all vertices shifted by the lattice length) the unwrapped box lies on the far side of the seam.
__device__ __forceinline__ bool d_reaches(const MsGeo& g, double blo, double bhi, int side) {
    const double W = g.my_hi - g.my_lo;
    const double t_lo = side == 0 ? g.my_lo - W : g.my_hi, t_hi = side == 0 ? g.my_lo : g.my_hi + W;
    for (int k = -1; k <= 1; ++k) {
        if (k != 0 && !g.per_x) continue;
        const double a = blo - 3.0 + k * g.GX, b = bhi + 3.0 + k * g.GX;
        if (b > t_lo && a < t_hi) return true;
    }
    return false;
}

// one CTA per slot of the group: append {gid, flag, payload[3][nv]} to the buffer of `side` (0 = left neighbour, 1 = right)
__global__ void __launch_bounds__(128) k_ms_pack(MsGeo g, int kind, const int* __restrict__ role, const int* __restrict__ gid_of,
                                                 const double* __restrict__ pstate, const double* __restrict__ src, long long cap_v,
                                                 double* buf0, double* buf1, int* counters) {
    const int side = blockIdx.y;
    double* buf = side == 0 ? buf0 : buf1;
    const int s = g.first_slot + blockIdx.x;
    const int r = role[s];
    __shared__ int s_idx;
    bool take = false;
    double flag = 0.0;
    if (kind == MS_VEL || kind == MS_PPF) {
        take = (r == 2 && side == 0) || (r == 3 && side == 1);
    } else if (r == 1) {
        const double* st = pstate + (long long)s * 11;
        const int dirn = d_direction(g, st[2]);
        const bool handoff = (side == 0 && dirn < 0) || (side == 1 && dirn > 0);
        take = handoff || d_reaches(g, st[5], st[6], side);
        flag = handoff ? 1.0 : 0.0;
        // Two slabs on a periodic axis: the left and the right neighbour are the SAME slab. One entry per particle, not one per side (two
        // entries of one particle would be unpacked concurrently there: an ownership hand-off could be overwritten by the plain ghost entry).
        if (g.two_periodic) {
            const bool other_handoff = (side == 0 && dirn > 0) || (side == 1 && dirn < 0);
            const bool other_take = other_handoff || d_reaches(g, st[5], st[6], 1 - side);
            if (other_take && (other_handoff || (!handoff && side == 0))) take = false;      // the hand-off side wins, else side 1
        }
    }
    if (side == 0 ? !g.has_left : !g.has_right) take = false;
    if (!take) return;
    if (threadIdx.x == 0) {
        const unsigned long long i = atomicAdd((unsigned long long*)buf, 1ull);
        if (i >= (unsigned long long)g.xcap) { counters[1] = 1; s_idx = -1; } else s_idx = (int)i;
    }
    __syncthreads();
    if (s_idx < 0) return;
    double* e = buf + 1 + (size_t)s_idx * g.entry;
    if (threadIdx.x == 0) { e[0] = (double)gid_of[s]; e[1] = flag; }
    const long long vbase = g.first_vertex + (long long)blockIdx.x * g.nv;
    for (int t = threadIdx.x; t < 3 * g.nv; t += blockDim.x) {
        const int c = t / g.nv, v = t % g.nv;
        e[2 + t] = src[(long long)c * cap_v + vbase + v];
    }
}
__global__ void k_ms_clear_counts(double* buf0, double* buf1) { *(unsigned long long*)buf0 = 0ull; *(unsigned long long*)buf1 = 0ull; }

// one CTA per received entry (blockIdx.y = side the buffer came from); reads at most `count` entries, wherever the buffer lives
__global__ void __launch_bounds__(128) k_ms_unpack(MsGeo g, int kind, int* role, int* gid_of, int* slot_of, int* free_stack,
                                                   int* counters, uint8_t* seen, const double* __restrict__ buf0, const double* __restrict__ buf1,
                                                   double* dst, long long cap_v, int gid_base) {
    const int side = blockIdx.y;
    const double* buf = side == 0 ? buf0 : buf1;
    if (!buf) return;
    const unsigned long long count = *(const unsigned long long*)buf;
    if (blockIdx.x >= count || blockIdx.x >= (unsigned)g.xcap) return;
    const double* e = buf + 1 + (size_t)blockIdx.x * g.entry;
    const int gid = (int)e[0];
    const int lid = gid - gid_base;
    __shared__ int s_slot;
    if (threadIdx.x == 0) {
        int s = slot_of[lid];
        if (kind == MS_POS) {
            if (s < 0) {
                const int top = atomicSub(&counters[0], 1) - 1;
                if (top < 0) { atomicAdd(&counters[0], 1); counters[1] = 2; s = -1; }
                else { s = free_stack[top]; slot_of[lid] = s; gid_of[s] = gid; role[s] = 0; }
            }
            if (s >= 0) {
                if (e[1] != 0.0) role[s] = 1;                    // ownership handed to this slab
                else if (role[s] != 1) role[s] = side == 0 ? 2 : 3;
                seen[s] = 1;
            }
        } else if (kind == MS_VEL || kind == MS_PPF) {
            if (s >= 0 && role[s] != 1) s = -1;                  // partial sums are only added by the owner
        } else {
            if (s >= 0 && role[s] < 2) s = -1;                   // forces are only taken by ghost copies
        }
        s_slot = s;
    }
    __syncthreads();
    const int s = s_slot;
    if (s < 0) return;
    const long long vbase = g.first_vertex + (long long)(s - g.first_slot) * g.nv;
    for (int t = threadIdx.x; t < 3 * g.nv; t += blockDim.x) {
        const int c = t / g.nv, v = t % g.nv;
        const long long i = (long long)c * cap_v + vbase + v;
        if (kind == MS_VEL || kind == MS_PPF) dst[i] += e[2 + t]; else dst[i] = e[2 + t];
    }
}

// after a POS exchange: owners whose centre of mass left become ghosts of the side it left to (the neighbour was told in the
// same exchange); ghost copies that were not refreshed are dropped.
__global__ void k_ms_housekeeping(MsGeo g, int* role, const int* gid_of, int* slot_of, int* free_stack, int* counters, uint8_t* seen,
                                  const double* __restrict__ pstate, const uint8_t* __restrict__ was_owner, int gid_base) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= g.cap) return;
    const int s = g.first_slot + k;
    const int r = role[s];
    if (r == 0) return;
    if (was_owner[s]) {
        const int dirn = d_direction(g, pstate[(long long)s * 11 + 2]);
        if (dirn != 0) role[s] = dirn < 0 ? 2 : 3;
    } else if (!seen[s]) {
        role[s] = 0;
        slot_of[gid_of[s] - gid_base] = -1;
        free_stack[atomicAdd(&counters[0], 1)] = s;
    }
    seen[s] = 0;
}
__global__ void k_ms_snapshot_owner(int n, const int* __restrict__ role, uint8_t* was_owner) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) was_owner[s] = role[s] == 1;
}

MsGeo make_ms_geo(const bib_domain* d, const MultiSlabIB* m) {
    MsGeo g;
    g.nv = m->nv; g.cap = m->cap; g.xcap = m->xcap; g.entry = m->entry; g.first_slot = m->first_slot; g.first_vertex = m->first_vertex;
    g.my_lo = (double)d->x_offset - 0.5; g.my_hi = (double)(d->x_offset + d->nx) - 0.5;
    g.center = 0.5 * (g.my_lo + g.my_hi); g.GX = (double)d->gnx; g.per_x = d->per[0];
    g.has_left = d->per[0] || d->rank > 0; g.has_right = d->per[0] || d->rank < d->n_ranks - 1;
    g.two_periodic = d->per[0] && d->n_ranks == 2;
    return g;
}

// One neighbour exchange of every particle group. PULL transport (default): pack into this slab's send buffers (counter reset + one
// launch for both sides), one stream-ordered neighbour sync for all groups, then the unpack kernel READS THE NEIGHBOURS' BUFFERS over
// NVLink peer memory (both sides in one launch): only the entries actually packed cross the link, no staging copy, no NCCL data
// transfer. COPY transport (BIB_IB_EXCHANGE=copy): the whole fixed-capacity buffers travel by ncclSend/ncclRecv / cudaMemcpyPeer.
int channel_of(int kind) { return kind == MS_VEL ? 1 : 0; }
int ms_pack(bib_domain* d, int kind, const double* src) {
    const int ch = channel_of(kind);
    for (MultiSlabIB* m : d->msv) {
        if (!m) continue;
        const MsGeo g = make_ms_geo(d, m);
        const int par = m->xseq[ch] & 1;
        k_ms_clear_counts<<<1, 1, 0, d->stream>>>(m->send[ch][par][0], m->send[ch][par][1]);
        k_ms_pack<<<dim3(m->cap, 2), 128, 0, d->stream>>>(g, kind, d->role_all, d->gid_all, d->pstate, src, d->cap_vertices, m->send[ch][par][0], m->send[ch][par][1], m->counters);
        d->launches += 2;
    }
    BIB_CHECK(cudaGetLastError());
    return 1;
}
// the transfer step: copy transport moves the buffers, pull transport only orders the streams
int ms_transfer(bib_domain* d, int kind) {
    const int ch = channel_of(kind);
    bool any_pull = false;
    for (MultiSlabIB* m : d->msv) {
        if (!m) continue;
        if (m->pull) { any_pull = true; continue; }
        const size_t nd = 1 + (size_t)m->xcap * m->entry;
        const int par = m->xseq[ch] & 1;
        if (!comm_exchange_pair(d, m->send[ch][par][0], m->send[ch][par][1], m->recv[0], m->recv[1], nd)) return 0;
    }
    if (any_pull && !comm_neighbour_sync(d)) return 0;
    return 1;
}
int ms_unpack(bib_domain* d, int kind, double* dst) {
    const int ch = channel_of(kind);
    const bool has_left = d->per[0] || d->rank > 0, has_right = d->per[0] || d->rank < d->n_ranks - 1;
    for (size_t gi = 0; gi < d->msv.size(); ++gi) {
        MultiSlabIB* m = d->msv[gi];
        if (!m) continue;
        const MsGeo g = make_ms_geo(d, m);
        const int par = m->xseq[ch] & 1;
        const double *b0 = m->recv[0], *b1 = m->recv[1];
        if (m->pull) {
            if (d->group) {           // in-process: the neighbours' buffers are plain pointers (same device or peer access enabled)
                const int right = (d->rank + 1) % d->n_ranks, left = (d->rank + d->n_ranks - 1) % d->n_ranks;
                b0 = group_slab(d->group, left)->msv[gi]->send[ch][par][1];
                b1 = group_slab(d->group, right)->msv[gi]->send[ch][par][0];
            } else { b0 = m->peer[ch][par][0]; b1 = m->peer[ch][par][1]; }
            if (!has_left) b0 = nullptr;
            if (!has_right) b1 = nullptr;
        }
        k_ms_unpack<<<dim3(m->xcap, 2), 128, 0, d->stream>>>(g, kind, d->role_all, d->gid_all, m->slot_of, m->free_stack, m->counters, d->seen_all,
                                                            b0, b1, dst, d->cap_vertices, m->gid_base);
        d->launches++;
        m->xseq[ch]++;
    }
    BIB_CHECK(cudaGetLastError());
    return 1;
}
int ms_exchange(bib_domain* d, int kind, const double* src, double* dst) {
    return ms_pack(d, kind, src) && ms_transfer(d, kind) && ms_unpack(d, kind, dst);
}
bool all_pull(const bib_domain* d) { for (const MultiSlabIB* m : d->msv) if (m && !m->pull) return false; return true; }

}  // namespace

int ib_update_all(bib_domain* d, int do_move);

// exchange context + slots of one more particle group (ParticleGroup index = d->msv.size())
int ms_create(bib_domain* d, int nv, int n_global, int cap_hint, int xcap_hint, int first_slot, long long first_vertex, int gid_base, MultiSlabIB** out) {
    MultiSlabIB* m = new MultiSlabIB();
    m->nv = nv; m->n_global = n_global; m->first_slot = first_slot; m->first_vertex = first_vertex; m->gid_base = gid_base;
    // small problems: every slab can hold (and exchange) everything; otherwise the geometric estimates of the caller
    m->cap = n_global <= 1024 ? n_global : std::min(n_global, std::max(cap_hint, 256));
    m->xcap = n_global <= 1024 ? m->cap : std::min(m->cap, std::max(xcap_hint, 64));
    m->entry = 2 + 3 * (size_t)nv;
    const size_t nd = 1 + (size_t)m->xcap * m->entry;
    const char* mode = getenv("BIB_IB_EXCHANGE");
    m->pull = !(mode && std::string(mode) == "copy");
    bool ok = bib_cuda_ok(cudaMalloc(&m->slot_of, std::max(1, n_global) * sizeof(int)), "cudaMalloc") &&
              bib_cuda_ok(cudaMalloc(&m->free_stack, std::max(1, m->cap) * sizeof(int)), "cudaMalloc") && bib_cuda_ok(cudaMalloc(&m->counters, 2 * sizeof(int)), "cudaMalloc");
    for (int ch = 0; ok && ch < 2; ++ch) for (int par = 0; ok && par < 2; ++par) for (int s = 0; ok && s < 2; ++s)
        ok = bib_cuda_ok(cudaMalloc(&m->send[ch][par][s], nd * sizeof(double)), "cudaMalloc") && bib_cuda_ok(cudaMemset(m->send[ch][par][s], 0, sizeof(double)), "cudaMemset");
    for (int s = 0; ok && !m->pull && s < 2; ++s)
        ok = bib_cuda_ok(cudaMalloc(&m->recv[s], nd * sizeof(double)), "cudaMalloc") && bib_cuda_ok(cudaMemset(m->recv[s], 0, sizeof(double)), "cudaMemset");
    // pull transport, one process per GPU: map the neighbours' send buffers (collective: every rank adds its particle groups in the same order)
    for (int ch = 0; ok && m->pull && !d->group && ch < 2; ++ch) for (int par = 0; ok && par < 2; ++par)
        ok = comm_map_peer_buffers(d, m->send[ch][par][0], m->send[ch][par][1], &m->peer[ch][par][0], &m->peer[ch][par][1]) != 0;
    if (!ok) { delete m; return 0; }
    d->bytes += (long long)(8 * nd * sizeof(double));
    *out = m;
    return 1;
}

// roles / ids of the group's initially owned slots [first_slot, first_slot + n_owned); the rest of its slots are free
int ms_set_initial(bib_domain* d, MultiSlabIB* m, const std::vector<int>& owned_lids) {
    std::vector<int> role(m->cap, 0), gid_of(m->cap, -1), slot_of(std::max(1, m->n_global), -1), free_stack(std::max(1, m->cap), 0), nvs(m->cap, m->nv);
    std::vector<long long> fvs(m->cap);
    const int n_owned = (int)owned_lids.size();
    for (int k = 0; k < n_owned; ++k) { role[k] = 1; gid_of[k] = m->gid_base + owned_lids[k]; slot_of[owned_lids[k]] = m->first_slot + k; }
    for (int k = 0; k < m->cap; ++k) fvs[k] = m->first_vertex + (long long)k * m->nv;
    int top = 0;
    for (int k = m->cap - 1; k >= n_owned; --k) free_stack[top++] = m->first_slot + k;
    const int counters[2] = {top, 0};
    BIB_CHECK(cudaMemcpy(d->role_all + m->first_slot, role.data(), m->cap * sizeof(int), cudaMemcpyHostToDevice));
    BIB_CHECK(cudaMemcpy(d->gid_all + m->first_slot, gid_of.data(), m->cap * sizeof(int), cudaMemcpyHostToDevice));
    BIB_CHECK(cudaMemcpy(d->slot_nv + m->first_slot, nvs.data(), m->cap * sizeof(int), cudaMemcpyHostToDevice));
    BIB_CHECK(cudaMemcpy(d->slot_fv + m->first_slot, fvs.data(), m->cap * sizeof(long long), cudaMemcpyHostToDevice));
    BIB_CHECK(cudaMemcpy(m->slot_of, slot_of.data(), slot_of.size() * sizeof(int), cudaMemcpyHostToDevice));
    BIB_CHECK(cudaMemcpy(m->free_stack, free_stack.data(), free_stack.size() * sizeof(int), cudaMemcpyHostToDevice));
    BIB_CHECK(cudaMemcpy(m->counters, counters, sizeof counters, cudaMemcpyHostToDevice));
    d->ghosts_dirty = true;
    return 1;
}

void ms_destroy(bib_domain* d) {
    for (MultiSlabIB* m : d->msv) {
        if (!m) continue;
        for (int ch = 0; ch < 2; ++ch) for (int par = 0; par < 2; ++par) for (int s = 0; s < 2; ++s) {
            if (m->peer[ch][par][s]) cudaIpcCloseMemHandle((void*)m->peer[ch][par][s]);
            if (m->send[ch][par][s]) cudaFree(m->send[ch][par][s]);
        }
        void* ptrs[] = {m->slot_of, m->free_stack, m->counters, m->recv[0], m->recv[1]};
        for (void* p : ptrs) if (p) cudaFree(p);
        delete m;
    }
    d->msv.clear();
    d->ms = nullptr;
}

// POS: refresh ghost copies, hand particles over, recompute the cached scalars (area, volume, COM, box) of ghosts
int ms_exchange_positions(bib_domain* d) {
    uint8_t* was_owner = d->seen_all + d->cap_particles;
    k_ms_snapshot_owner<<<(d->n_particles + 255) / 256, 256, 0, d->stream>>>(d->n_particles, d->role_all, was_owner);
    d->launches++;
    if (!ms_exchange(d, MS_POS, d->vpos, d->vpos)) return 0;
    for (MultiSlabIB* m : d->msv) {
        if (!m) continue;
        k_ms_housekeeping<<<(m->cap + 255) / 256, 256, 0, d->stream>>>(make_ms_geo(d, m), d->role_all, d->gid_all, m->slot_of, m->free_stack, m->counters, d->seen_all,
                                                                     d->pstate, was_owner, m->gid_base);
        d->launches++;
    }
    BIB_CHECK(cudaGetLastError());
    d->ghosts_dirty = false;
    return ib_update_all(d, 0);      // recompute the cached scalars of ghost / newly owned slots (owners are masked out), nothing moves
}
// VEL: bib_step defers the sync (d->defer_vel): the entries are packed right after the interpolation and unpacked together with the FORCE
// entries, behind ONE neighbour sync (the membrane forces in between do not read the vertex velocities). Pull transport only.
int ms_exchange_velocities(bib_domain* d) {
    if (d->defer_vel && all_pull(d)) { d->vel_packed = true; return ms_pack(d, MS_VEL, d->vvel); }
    return ms_exchange(d, MS_VEL, d->vvel, d->vvel);
}
int ms_finish_velocities(bib_domain* d) {          // stand-alone completion (no FORCE exchange follows)
    if (!d->vel_packed) return 1;
    d->vel_packed = false;
    return ms_transfer(d, MS_VEL) && ms_unpack(d, MS_VEL, d->vvel);
}
int ms_exchange_forces(bib_domain* d) {
    if (!d->vel_packed) return ms_exchange(d, MS_FORCE, d->vfrc, d->vfrc);
    d->vel_packed = false;
    return ms_pack(d, MS_FORCE, d->vfrc) && ms_transfer(d, MS_FORCE) && ms_unpack(d, MS_VEL, d->vvel) && ms_unpack(d, MS_FORCE, d->vfrc);
}
int ms_exchange_pair_forces(bib_domain* d) { return ms_exchange(d, MS_PPF, d->vfrc, d->vfrc); }

int ms_check_overflow(bib_domain* d) {
    for (MultiSlabIB* m : d->msv) {
        if (!m) continue;
        int c[2] = {0, 0};
        BIB_CHECK(cudaMemcpyAsync(c, m->counters, sizeof c, cudaMemcpyDeviceToHost, d->stream));
        BIB_CHECK(cudaStreamSynchronize(d->stream));
        if (c[1] == 1) { bib_set_error("multi-slab IB: neighbour exchange buffer overflow (too many particles near one slab face)"); return 0; }
        if (c[1] == 2) { bib_set_error("multi-slab IB: no free particle slot for an incoming ghost"); return 0; }
    }
    return 1;
}

int ib_local_slots(bib_domain* d, int* gids, int* roles, int capacity, int* n_slots) {
    if (!d->ms) {
        const int n = std::min(capacity, d->n_particles);
        for (int i = 0; i < n; ++i) { gids[i] = i; roles[i] = 1; }
        *n_slots = d->n_particles;
        return 1;
    }
    const int n = std::min(capacity, d->n_particles);
    BIB_CHECK(cudaMemcpyAsync(gids, d->gid_all, n * sizeof(int), cudaMemcpyDeviceToHost, d->stream));
    BIB_CHECK(cudaMemcpyAsync(roles, d->role_all, n * sizeof(int), cudaMemcpyDeviceToHost, d->stream));
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    *n_slots = d->n_particles;
    return 1;
}

// host copy of the slot table (group.cu gathers / scatters particle data in global order with it)
int ms_slot_table(bib_domain* d, std::vector<int>& gid, std::vector<int>& role, std::vector<long long>& first_vertex, std::vector<int>& nv) {
    const int n = d->n_particles;
    gid.assign(n, 0); role.assign(n, 1); first_vertex.assign(n, 0); nv.assign(n, 0);
    if (n == 0) return 1;
    if (!d->ms) {
        for (const ParticleGroup& gr : d->groups)
            for (int k = 0; k < gr.count; ++k) { const int s = gr.first_particle + k; gid[s] = s; nv[s] = d->shapes[gr.shape].nv; first_vertex[s] = gr.first_vertex + (long long)k * nv[s]; }
        return 1;
    }
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    BIB_CHECK(cudaMemcpy(gid.data(), d->gid_all, n * sizeof(int), cudaMemcpyDeviceToHost));
    BIB_CHECK(cudaMemcpy(role.data(), d->role_all, n * sizeof(int), cudaMemcpyDeviceToHost));
    BIB_CHECK(cudaMemcpy(nv.data(), d->slot_nv, n * sizeof(int), cudaMemcpyDeviceToHost));
    BIB_CHECK(cudaMemcpy(first_vertex.data(), d->slot_fv, n * sizeof(long long), cudaMemcpyDeviceToHost));
    return 1;
}
