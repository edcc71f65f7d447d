// ib_multislab.cu -- immersed boundaries across x-slabs: ghost particles, partial-sum exchange, ownership hand-off.
//
// Replaces the reference's ghost-vertex machinery (NonLocalNode lists, three CommunicationBuffer exchanges per step and
// particle hand-off: src/fsi/ImmersedBoundaryDynamics.hh:163-246,300-330,416-450; _interpolation.hh:279-326;
// _forces.hh:378-456) with whole-particle ghosts in fixed-capacity slot arrays, so that every per-step exchange is one
// grouped ncclSend/ncclRecv of a fixed-size buffer per neighbour and no host round trip is needed to learn message sizes.
// Protocol per step:
//   POS   after the owner moved a particle: positions -> both neighbours whose slab its box (+3 cells) reaches; the entry
//         carries a hand-off flag when the centre of mass left the owner's x-range (the owner alone decides);
//   VEL   after interpolation: a ghost holder returns the partial velocity sums of its cells, the owner adds them;
//   FORCE after the owner evaluated membrane forces: forces -> the same neighbours, each slab spreads into its own cells.
#include "domain.cuh"

#include <algorithm>
#include <vector>

namespace {

struct MsGeo {
    int nv, cap, xcap;
    size_t entry;
    double my_lo, my_hi, center, GX;    // owned x-range [my_lo, my_hi) in global lattice coordinates
    int per_x, has_left, has_right;
};

__device__ __forceinline__ int d_direction(const MsGeo& g, double com_x) {
    if (com_x >= g.my_lo && com_x < g.my_hi) return 0;
    double dlt = com_x - g.center;
    if (g.per_x) dlt = dlt - floor(dlt / g.GX + 0.5) * g.GX;
    return dlt < 0 ? -1 : 1;
}

enum { MS_POS = 0, MS_VEL = 1, MS_FORCE = 2 };

// one CTA per slot: append {gid, flag, payload[3][nv]} to the buffer of `side` (0 = left neighbour, 1 = right)
__global__ void __launch_bounds__(128) k_ms_pack(MsGeo g, int kind, int side, const int* __restrict__ role, const int* __restrict__ gid_of,
                                                 const double* __restrict__ pstate, const double* __restrict__ src, long long cap_v,
                                                 double* buf, int* counters) {
    const int s = blockIdx.x;
    const int r = role[s];
    __shared__ int s_idx;
    bool take = false;
    double flag = 0.0;
    if (kind == MS_VEL) {
        take = (r == 2 && side == 0) || (r == 3 && side == 1);
    } else if (r == 1) {
        const double* st = pstate + (long long)s * 11;
        const int dirn = d_direction(g, st[2]);
        const bool reach = side == 0 ? (st[5] - 3.0 < g.my_lo) : (st[6] + 3.0 > g.my_hi);
        const bool handoff = (side == 0 && dirn < 0) || (side == 1 && dirn > 0);
        take = reach || handoff;
        flag = handoff ? 1.0 : 0.0;
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
    for (int t = threadIdx.x; t < 3 * g.nv; t += blockDim.x) {
        const int c = t / g.nv, v = t % g.nv;
        e[2 + t] = src[(long long)c * cap_v + (long long)s * g.nv + v];
    }
}

// one CTA per received entry
__global__ void __launch_bounds__(128) k_ms_unpack(MsGeo g, int kind, int side, int* role, int* gid_of, int* slot_of, int* free_stack,
                                                   int* counters, uint8_t* seen, const double* __restrict__ buf, double* dst, long long cap_v) {
    const unsigned long long count = *(const unsigned long long*)buf;
    if (blockIdx.x >= count || blockIdx.x >= (unsigned)g.xcap) return;
    const double* e = buf + 1 + (size_t)blockIdx.x * g.entry;
    const int gid = (int)e[0];
    __shared__ int s_slot;
    if (threadIdx.x == 0) {
        int s = slot_of[gid];
        if (kind == MS_POS) {
            if (s < 0) {
                const int top = atomicSub(&counters[0], 1) - 1;
                if (top < 0) { atomicAdd(&counters[0], 1); counters[1] = 2; s = -1; }
                else { s = free_stack[top]; slot_of[gid] = s; gid_of[s] = gid; role[s] = 0; }
            }
            if (s >= 0) {
                if (e[1] != 0.0) role[s] = 1;                    // ownership handed to this slab
                else if (role[s] != 1) role[s] = side == 0 ? 2 : 3;
                seen[s] = 1;
            }
        } else if (kind == MS_VEL) {
            if (s >= 0 && role[s] != 1) s = -1;                  // partial sums are only added by the owner
        } else {
            if (s >= 0 && role[s] < 2) s = -1;                   // forces are only taken by ghost copies
        }
        s_slot = s;
    }
    __syncthreads();
    const int s = s_slot;
    if (s < 0) return;
    for (int t = threadIdx.x; t < 3 * g.nv; t += blockDim.x) {
        const int c = t / g.nv, v = t % g.nv;
        const long long i = (long long)c * cap_v + (long long)s * g.nv + v;
        if (kind == MS_VEL) dst[i] += e[2 + t]; else dst[i] = e[2 + t];
    }
}

// after a POS exchange: owners whose centre of mass left become ghosts of the side it left to (the neighbour was told in the
// same exchange); ghost copies that were not refreshed are dropped.
__global__ void k_ms_housekeeping(MsGeo g, int* role, const int* gid_of, int* slot_of, int* free_stack, int* counters, uint8_t* seen,
                                  const double* __restrict__ pstate, const uint8_t* __restrict__ was_owner) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= g.cap) return;
    const int r = role[s];
    if (r == 0) return;
    if (was_owner[s]) {
        const int dirn = d_direction(g, pstate[(long long)s * 11 + 2]);
        if (dirn != 0) role[s] = dirn < 0 ? 2 : 3;
    } else if (!seen[s]) {
        role[s] = 0;
        slot_of[gid_of[s]] = -1;
        free_stack[atomicAdd(&counters[0], 1)] = s;
    }
    seen[s] = 0;
}
__global__ void k_ms_snapshot_owner(int cap, const int* __restrict__ role, uint8_t* was_owner) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap) was_owner[s] = role[s] == 1;
}

MsGeo make_ms_geo(const bib_domain* d) {
    const MultiSlabIB* m = d->ms;
    MsGeo g;
    g.nv = m->nv; g.cap = m->cap; g.xcap = m->xcap; g.entry = m->entry;
    g.my_lo = (double)d->x_offset - 0.5; g.my_hi = (double)(d->x_offset + d->nx) - 0.5;
    g.center = 0.5 * (g.my_lo + g.my_hi); g.GX = (double)d->gnx; g.per_x = d->per[0];
    g.has_left = d->per[0] || d->rank > 0; g.has_right = d->per[0] || d->rank < d->n_ranks - 1;
    return g;
}

int ms_exchange(bib_domain* d, int kind, const double* src, double* dst) {
    MultiSlabIB* m = d->ms;
    const MsGeo g = make_ms_geo(d);
    const size_t nd = 1 + (size_t)m->xcap * m->entry;
    for (int side = 0; side < 2; ++side) {
        BIB_CHECK(cudaMemsetAsync(m->send[side], 0, sizeof(double), d->stream));
        k_ms_pack<<<m->cap, 128, 0, d->stream>>>(g, kind, side, m->role, m->gid_of, d->pstate, src, d->cap_vertices, m->send[side], m->counters);
        d->launches++;
    }
    if (!comm_exchange_pair(d, m->send[0], m->send[1], m->recv[0], m->recv[1], nd)) return 0;
    for (int side = 0; side < 2; ++side) {
        k_ms_unpack<<<m->xcap, 128, 0, d->stream>>>(g, kind, side, m->role, m->gid_of, m->slot_of, m->free_stack, m->counters, m->seen,
                                                   m->recv[side], dst, d->cap_vertices);
        d->launches++;
    }
    BIB_CHECK(cudaGetLastError());
    return 1;
}

}  // namespace

int ib_update_all(bib_domain* d, int do_move);

int ms_create(bib_domain* d, int nv, int n_global, int cap_hint, int xcap_hint) {
    MultiSlabIB* m = new MultiSlabIB();
    m->nv = nv; m->n_global = n_global;
    // small problems: every slab can hold (and exchange) everything; otherwise the geometric estimates of the caller
    m->cap = n_global <= 1024 ? n_global : std::min(n_global, std::max(cap_hint, 256));
    m->xcap = n_global <= 1024 ? m->cap : std::min(m->cap, std::max(xcap_hint, 64));
    m->entry = 2 + 3 * (size_t)nv;
    const size_t nd = 1 + (size_t)m->xcap * m->entry;
    bool ok = bib_cuda_ok(cudaMalloc(&m->role, m->cap * sizeof(int)), "cudaMalloc") && bib_cuda_ok(cudaMalloc(&m->gid_of, m->cap * sizeof(int)), "cudaMalloc") &&
              bib_cuda_ok(cudaMalloc(&m->slot_of, std::max(1, n_global) * sizeof(int)), "cudaMalloc") &&
              bib_cuda_ok(cudaMalloc(&m->free_stack, m->cap * sizeof(int)), "cudaMalloc") && bib_cuda_ok(cudaMalloc(&m->counters, 2 * sizeof(int)), "cudaMalloc") &&
              bib_cuda_ok(cudaMalloc(&m->seen, 2 * (size_t)m->cap), "cudaMalloc");
    for (int s = 0; ok && s < 2; ++s)
        ok = bib_cuda_ok(cudaMalloc(&m->send[s], nd * sizeof(double)), "cudaMalloc") && bib_cuda_ok(cudaMalloc(&m->recv[s], nd * sizeof(double)), "cudaMalloc") &&
             bib_cuda_ok(cudaMemset(m->recv[s], 0, sizeof(double)), "cudaMemset");
    if (!ok) return 0;
    d->bytes += (long long)(4 * nd * sizeof(double));
    d->ms = m;
    return 1;
}

// roles / ids of the initially owned slots [0, n_owned); the rest of the slots are free
int ms_set_initial(bib_domain* d, const std::vector<int>& owned_gids) {
    MultiSlabIB* m = d->ms;
    std::vector<int> role(m->cap, 0), gid_of(m->cap, -1), slot_of(std::max(1, m->n_global), -1), free_stack(m->cap, 0);
    const int n_owned = (int)owned_gids.size();
    for (int s = 0; s < n_owned; ++s) { role[s] = 1; gid_of[s] = owned_gids[s]; slot_of[owned_gids[s]] = s; }
    int top = 0;
    for (int s = m->cap - 1; s >= n_owned; --s) free_stack[top++] = s;
    const int counters[2] = {top, 0};
    BIB_CHECK(cudaMemcpy(m->role, role.data(), m->cap * sizeof(int), cudaMemcpyHostToDevice));
    BIB_CHECK(cudaMemcpy(m->gid_of, gid_of.data(), m->cap * sizeof(int), cudaMemcpyHostToDevice));
    BIB_CHECK(cudaMemcpy(m->slot_of, slot_of.data(), slot_of.size() * sizeof(int), cudaMemcpyHostToDevice));
    BIB_CHECK(cudaMemcpy(m->free_stack, free_stack.data(), m->cap * sizeof(int), cudaMemcpyHostToDevice));
    BIB_CHECK(cudaMemcpy(m->counters, counters, sizeof counters, cudaMemcpyHostToDevice));
    BIB_CHECK(cudaMemset(m->seen, 0, 2 * (size_t)m->cap));
    m->ghosts_dirty = true;
    return 1;
}

void ms_destroy(bib_domain* d) {
    MultiSlabIB* m = d->ms;
    if (!m) return;
    void* ptrs[] = {m->role, m->gid_of, m->slot_of, m->free_stack, m->counters, m->seen, m->send[0], m->send[1], m->recv[0], m->recv[1]};
    for (void* p : ptrs) if (p) cudaFree(p);
    delete m;
    d->ms = nullptr;
}

// POS: refresh ghost copies, hand particles over, recompute the cached scalars (area, volume, COM, box) of ghosts
int ms_exchange_positions(bib_domain* d) {
    MultiSlabIB* m = d->ms;
    uint8_t* was_owner = m->seen + m->cap;
    k_ms_snapshot_owner<<<(m->cap + 255) / 256, 256, 0, d->stream>>>(m->cap, m->role, was_owner);
    d->launches++;
    if (!ms_exchange(d, MS_POS, d->vpos, d->vpos)) return 0;
    k_ms_housekeeping<<<(m->cap + 255) / 256, 256, 0, d->stream>>>(make_ms_geo(d), m->role, m->gid_of, m->slot_of, m->free_stack, m->counters, m->seen,
                                                                 d->pstate, was_owner);
    d->launches++;
    BIB_CHECK(cudaGetLastError());
    m->ghosts_dirty = false;
    return ib_update_all(d, 0);      // recompute the cached scalars of ghost / newly owned slots (owners are masked out), nothing moves
}
int ms_exchange_velocities(bib_domain* d) { return ms_exchange(d, MS_VEL, d->vvel, d->vvel); }
int ms_exchange_forces(bib_domain* d) { return ms_exchange(d, MS_FORCE, d->vfrc, d->vfrc); }

int ms_check_overflow(bib_domain* d) {
    if (!d->ms) return 1;
    int c[2] = {0, 0};
    BIB_CHECK(cudaMemcpyAsync(c, d->ms->counters, sizeof c, cudaMemcpyDeviceToHost, d->stream));
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    if (c[1] == 1) { bib_set_error("multi-slab IB: neighbour exchange buffer overflow (too many particles near one slab face)"); return 0; }
    if (c[1] == 2) { bib_set_error("multi-slab IB: no free particle slot for an incoming ghost"); return 0; }
    return 1;
}

int ib_local_slots(bib_domain* d, int* gids, int* roles, int capacity, int* n_slots) {
    if (!d->ms) {
        const int n = std::min(capacity, d->n_particles);
        for (int i = 0; i < n; ++i) { gids[i] = i; roles[i] = 1; }
        *n_slots = d->n_particles;
        return 1;
    }
    MultiSlabIB* m = d->ms;
    const int n = std::min(capacity, m->cap);
    BIB_CHECK(cudaMemcpyAsync(gids, m->gid_of, n * sizeof(int), cudaMemcpyDeviceToHost, d->stream));
    BIB_CHECK(cudaMemcpyAsync(roles, m->role, n * sizeof(int), cudaMemcpyDeviceToHost, d->stream));
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    *n_slots = m->cap;
    return 1;
}
