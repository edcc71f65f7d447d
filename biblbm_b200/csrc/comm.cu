// comm.cu -- x-slab ring halo exchange over NCCL (NVLink 5 / NVSwitch).
//
// Replaces the reference's envelope exchange (src/parallelism/parallelBlockCommunicator3D.cpp:144-233, whole cells of
// 184 B through MPI point-to-point) by the minimum a D3Q19 slab face needs: the 5 populations that cross it, 40 B per
// face cell and direction. The push kernel deposits them in the slab's ghost planes; because every population is its own
// [x][y][z] array, a ghost plane and its destination plane in the neighbour are contiguous ny*nz runs, so ncclSend /
// ncclRecv work in place on the population buffers: no pack or unpack kernels.
//
// NCCL is resolved at run time (dlopen) so that the library has no link-time dependency on a particular libnccl and
// shares the copy an embedding process (e.g. PyTorch) has already loaded.
#include "domain.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

bool load_nccl() {
    if (g_nccl.handle) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.handle) break; }
    if (!g_nccl.handle) { bib_set_error(std::string("cannot load libnccl: ") + dlerror()); return false; }
#define LOAD(sym) g_nccl.sym = (decltype(g_nccl.sym))dlsym(g_nccl.handle, "nccl" #sym); if (!g_nccl.sym) { bib_set_error("libnccl lacks nccl" #sym); return false; }
    LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommDestroy) LOAD(GroupStart) LOAD(GroupEnd) LOAD(Send) LOAD(Recv) LOAD(GetErrorString)
#undef LOAD
    return true;
}
bool nccl_ok(ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return true;
    bib_set_error(std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error"));
    return false;
}
#define NCCL_CHECK(call) do { if (!nccl_ok((call), #call)) return 0; } while (0)

const int kPlusX[5] = {10, 13, 14, 15, 16};    // c_x = +1 (src/latticeBoltzmann/nearestNeighborLattices3D.hh:167-177)
const int kMinusX[5] = {1, 4, 5, 6, 7};        // c_x = -1
const int kCy[19] = {0, 0, -1, 0, -1, 1, 0, 0, -1, -1, 0, 1, 0, 1, -1, 0, 0, 1, 1};
const int kCz[19] = {0, 0, 0, -1, 0, 0, -1, 1, -1, 1, 0, 0, 1, 0, 0, 1, -1, 1, -1};

// With a non-periodic y or z face the neighbour's ghost plane has holes: an entry whose source row lies outside the
// lattice was never pushed by the sender, and the receiving cell has already filled that slot itself with its bounced
// population. Received planes are therefore staged and merged, skipping those entries.
__global__ void k_halo_merge(double* __restrict__ dst, const double* __restrict__ src, int ny, int nz, int cy, int cz, int per_y, int per_z) {
    const long long n = (long long)ny * nz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(t % nz), y = (int)(t / nz);
        const int sy = y - cy, sz = z - cz;
        if ((!per_y && (sy < 0 || sy >= ny)) || (!per_z && (sz < 0 || sz >= nz))) continue;
        dst[t] = src[t];
    }
}

// The whole exchange of a slab as ONE kernel: thread t handles population k = t / plane (k < 5: c_x = +1, from the LEFT neighbour's ghost plane
// nx_l + 1 into my plane 1; k >= 5: c_x = -1, from the RIGHT neighbour's ghost plane 0 into my plane nx) at cell (y, z) of the face. fl / fr
// point at the neighbours' CURRENT population frames (peer memory: another device of this process, or an IPC mapping of another process),
// NULL where there is no neighbour. The holes of a ghost plane behind a y / z wall are skipped (see k_halo_merge). Four independent loads
// per thread and iteration keep enough bytes in flight for the NVLink round trip.
__global__ void __launch_bounds__(256) k_halo_pull(double* __restrict__ f, long long vol, long long plane, int nx, int ny, int nz, int per_y, int per_z,
                                                   const double* __restrict__ fl, long long vol_l, int nx_l, const double* __restrict__ fr, long long vol_r) {
    const int plus_x[5] = {10, 13, 14, 15, 16}, minus_x[5] = {1, 4, 5, 6, 7};
    const int cy[19] = {0, 0, -1, 0, -1, 1, 0, 0, -1, -1, 0, 1, 0, 1, -1, 0, 0, 1, 1};
    const int cz[19] = {0, 0, 0, -1, 0, 0, -1, 1, -1, 1, 0, 0, 1, 0, 0, 1, -1, 1, -1};
    const long long n = 10 * plane, stride = (long long)gridDim.x * blockDim.x;
    for (long long t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; t0 < n; t0 += 4 * stride) {
        double v[4]; double* dst[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long t = t0 + u * stride;
            dst[u] = nullptr;
            if (t >= n) continue;
            const int k = (int)(t / plane);
            const long long c = t - (long long)k * plane;
            const int y = (int)(c / nz), z = (int)(c - (long long)y * nz);
            const int i = k < 5 ? plus_x[k] : minus_x[k - 5];
            const int sy = y - cy[i], sz = z - cz[i];
            if ((!per_y && (sy < 0 || sy >= ny)) || (!per_z && (sz < 0 || sz >= nz))) continue;
            if (k < 5) { if (fl) { v[u] = fl[(long long)i * vol_l + (long long)(nx_l + 1) * plane + c]; dst[u] = f + (long long)i * vol + plane + c; } }
            else if (fr) { v[u] = fr[(long long)i * vol_r + c]; dst[u] = f + (long long)i * vol + (long long)nx * plane + c; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) if (dst[u]) *dst[u] = v[u];
    }
}

}  // namespace

static int launch_halo_pull(bib_domain* d, cudaStream_t st, const double* fl, long long vol_l, int nx_l, const double* fr, long long vol_r) {
    const long long n = 10 * (long long)d->plane;
    const int grid = (int)std::min<long long>((n + 4 * 256 - 1) / (4 * 256), (long long)d->n_sms * 8);
    k_halo_pull<<<std::max(grid, 1), 256, 0, st>>>(d->f[d->cur], (long long)d->vol, (long long)d->plane, d->nx, d->ny, d->nz, d->per[1], d->per[2],
                                                  fl, vol_l, nx_l, fr, vol_r);
    d->launches++;
    BIB_CHECK(cudaGetLastError());
    return 1;
}

int comm_unique_id(void* out128) {
    if (!load_nccl()) return 0;
    ncclUniqueId id;
    NCCL_CHECK(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
    return 1;
}

int comm_init(bib_domain* d, int rank, int n_ranks, const void* id128) {
    d->rank = rank; d->n_ranks = n_ranks;
    d->gnx = d->nx * n_ranks;           // equal slabs
    d->x_offset = rank * d->nx;
    if (n_ranks == 1) return 1;
    if (!load_nccl()) return 0;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    NCCL_CHECK(g_nccl.CommInitRank(&comm, n_ranks, id, rank));
    d->nccl_comm = comm;
    return 1;
}

void comm_destroy(bib_domain* d) {
    for (int s = 0; s < 2; ++s) for (int k = 0; k < 2; ++k) {
        if (d->peer_map[s][k] && !(s == 1 && d->peer_map[1][k] == d->peer_map[0][k])) cudaIpcCloseMemHandle(d->peer_map[s][k]);
        d->peer_map[s][k] = nullptr;
    }
    if (d->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)d->nccl_comm);
    d->nccl_comm = nullptr;
}

// Runs on the compute stream right after the push kernel: ghost plane nx+1 (populations with c_x=+1) goes to the right
// neighbour's plane 1, ghost plane 0 (c_x=-1) to the left neighbour's plane nx.
// In-process slab group: the same exchange as a PULL over peer memory. Order: (1) my ghost planes are written (event on st),
// host barrier, (2) wait for the neighbours' events and copy their ghost planes into my boundary planes, (3) event "pulled",
// host barrier, the neighbours' compute streams wait for it before they overwrite their ghost planes in the next step.
static int local_halo_exchange(bib_domain* d, cudaStream_t st) {
    bib_group* g = d->group;
    double* f = d->f[d->cur];
    const size_t n = d->plane;
    const int right = (d->rank + 1) % d->n_ranks, left = (d->rank + d->n_ranks - 1) % d->n_ranks;
    const bool has_right = d->per[0] || d->rank < d->n_ranks - 1;
    const bool has_left = d->per[0] || d->rank > 0;
    const bool staged = !d->per[1] || !d->per[2];
    BIB_CHECK(cudaEventRecord(d->ev_h_ready, st));
    group_barrier(g);
    bib_domain* L = group_slab(g, left); bib_domain* R = group_slab(g, right);
    const double* fl = L->f[L->cur]; const double* fr = R->f[R->cur];      // every slab flipped `cur` before the exchange
    if (d->halo_pull < 0) {
        const char* mode = std::getenv("BIB_HALO");
        d->halo_pull = (d->halo_peer_ok && !(mode && std::string(mode) == "copy")) ? 1 : 0;
    }
    if (d->halo_pull == 1) {                // one kernel reading the neighbours' ghost planes (same device or NVLink peer access)
        if (has_left) BIB_CHECK(cudaStreamWaitEvent(st, L->ev_h_ready, 0));
        if (has_right) BIB_CHECK(cudaStreamWaitEvent(st, R->ev_h_ready, 0));
        if (!launch_halo_pull(d, st, has_left ? fl : nullptr, (long long)L->vol, L->nx, has_right ? fr : nullptr, (long long)R->vol)) return 0;
        BIB_CHECK(cudaEventRecord(d->ev_h_done, st));
        group_barrier(g);
        if (has_left) BIB_CHECK(cudaStreamWaitEvent(d->stream, L->ev_h_done, 0));
        if (has_right) BIB_CHECK(cudaStreamWaitEvent(d->stream, R->ev_h_done, 0));
        return 1;
    }
    if (staged && !d->halo_stage) {
        BIB_CHECK(cudaMalloc(&d->halo_stage, 10 * n * sizeof(double)));
        d->bytes += (long long)(10 * n * sizeof(double));
    }
    if (has_left) {
        BIB_CHECK(cudaStreamWaitEvent(st, L->ev_h_ready, 0));
        for (int k = 0; k < 5; ++k)
            BIB_CHECK(cudaMemcpyPeerAsync(staged ? d->halo_stage + (size_t)k * n : f + (size_t)kPlusX[k] * d->vol + (size_t)1 * n, d->device,
                                          fl + (size_t)kPlusX[k] * L->vol + (size_t)(L->nx + 1) * n, L->device, n * sizeof(double), st));
    }
    if (has_right) {
        BIB_CHECK(cudaStreamWaitEvent(st, R->ev_h_ready, 0));
        for (int k = 0; k < 5; ++k)
            BIB_CHECK(cudaMemcpyPeerAsync(staged ? d->halo_stage + (size_t)(5 + k) * n : f + (size_t)kMinusX[k] * d->vol + (size_t)d->nx * n, d->device,
                                          fr + (size_t)kMinusX[k] * R->vol + (size_t)0 * n, R->device, n * sizeof(double), st));
    }
    BIB_CHECK(cudaEventRecord(d->ev_h_done, st));
    d->launches++;
    group_barrier(g);
    // my ghost planes may be overwritten (next step's boundary launch on the compute stream) only after the neighbours pulled them
    if (has_left) BIB_CHECK(cudaStreamWaitEvent(d->stream, L->ev_h_done, 0));
    if (has_right) BIB_CHECK(cudaStreamWaitEvent(d->stream, R->ev_h_done, 0));
    if (staged) {
        const int grid = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
        for (int k = 0; k < 5; ++k) {
            if (has_left) { k_halo_merge<<<grid, 256, 0, st>>>(f + (size_t)kPlusX[k] * d->vol + (size_t)1 * n, d->halo_stage + (size_t)k * n, d->ny, d->nz, kCy[kPlusX[k]], kCz[kPlusX[k]], d->per[1], d->per[2]); d->launches++; }
            if (has_right) { k_halo_merge<<<grid, 256, 0, st>>>(f + (size_t)kMinusX[k] * d->vol + (size_t)d->nx * n, d->halo_stage + (size_t)(5 + k) * n, d->ny, d->nz, kCy[kMinusX[k]], kCz[kMinusX[k]], d->per[1], d->per[2]); d->launches++; }
        }
        BIB_CHECK(cudaGetLastError());
    }
    return 1;
}

// One process per GPU: a grouped 1-element ncclSend / ncclRecv with each neighbour on `st` = work issued on st afterwards runs after
// everything the neighbours issued before THEIR matching call on their stream.
static int nccl_token_sync(bib_domain* d, cudaStream_t st) {
    ncclComm_t comm = (ncclComm_t)d->nccl_comm;
    const int right = (d->rank + 1) % d->n_ranks, left = (d->rank + d->n_ranks - 1) % d->n_ranks;
    const bool has_right = d->per[0] || d->rank < d->n_ranks - 1;
    const bool has_left = d->per[0] || d->rank > 0;
    if (!d->token) { BIB_CHECK(cudaMalloc(&d->token, 4 * sizeof(double))); BIB_CHECK(cudaMemset(d->token, 0, 4 * sizeof(double))); }
    NCCL_CHECK(g_nccl.GroupStart());
    if (has_left) { NCCL_CHECK(g_nccl.Send(d->token, 1, ncclDouble, left, comm, st)); NCCL_CHECK(g_nccl.Recv(d->token + 2, 1, ncclDouble, left, comm, st)); }
    if (has_right) { NCCL_CHECK(g_nccl.Send(d->token + 1, 1, ncclDouble, right, comm, st)); NCCL_CHECK(g_nccl.Recv(d->token + 3, 1, ncclDouble, right, comm, st)); }
    NCCL_CHECK(g_nccl.GroupEnd());
    d->launches++;
    return 1;
}

// One process per GPU, first exchange (COLLECTIVE): every rank tells its neighbours how its populations are stored (component stride, planes,
// in-place shift) and exports the allocation(s) with cudaIpcGetMemHandle; the neighbours map them, so that k_halo_pull can read the ghost
// planes over NVLink. BIB_HALO=nccl (same value on every rank) keeps the ncclSend / ncclRecv transport of the planes.
struct HaloDesc { long long vol; int nx, shift; cudaIpcMemHandle_t h[2]; };
static int setup_halo_pull(bib_domain* d) {
    const char* mode = std::getenv("BIB_HALO");
    if (mode && std::string(mode) == "nccl") { d->halo_pull = 0; return 1; }
    ncclComm_t comm = (ncclComm_t)d->nccl_comm;
    const int right = (d->rank + 1) % d->n_ranks, left = (d->rank + d->n_ranks - 1) % d->n_ranks;
    const bool has_right = d->per[0] || d->rank < d->n_ranks - 1;
    const bool has_left = d->per[0] || d->rank > 0;
    HaloDesc desc[3];                          // [0] mine, [1] the left neighbour's, [2] the right neighbour's
    memset(desc, 0, sizeof desc);
    desc[0].vol = (long long)d->vol; desc[0].nx = d->nx; desc[0].shift = d->inplace_shift;
    BIB_CHECK(cudaIpcGetMemHandle(&desc[0].h[0], d->inplace_shift ? d->f_base : d->f[0]));
    if (!d->inplace_shift) BIB_CHECK(cudaIpcGetMemHandle(&desc[0].h[1], d->f[1]));
    char* dev = nullptr;
    BIB_CHECK(cudaMalloc(&dev, sizeof desc));
    BIB_CHECK(cudaMemcpy(dev, desc, sizeof desc, cudaMemcpyHostToDevice));
    const size_t nb = sizeof(HaloDesc);
    bool ok = nccl_ok(g_nccl.GroupStart(), "ncclGroupStart");
    if (ok && has_left) ok = nccl_ok(g_nccl.Send(dev, nb, ncclChar, left, comm, d->stream), "ncclSend") && nccl_ok(g_nccl.Recv(dev + nb, nb, ncclChar, left, comm, d->stream), "ncclRecv");
    if (ok && has_right) ok = nccl_ok(g_nccl.Send(dev, nb, ncclChar, right, comm, d->stream), "ncclSend") && nccl_ok(g_nccl.Recv(dev + 2 * nb, nb, ncclChar, right, comm, d->stream), "ncclRecv");
    ok = nccl_ok(g_nccl.GroupEnd(), "ncclGroupEnd") && ok;
    ok = ok && bib_cuda_ok(cudaStreamSynchronize(d->stream), "cudaStreamSynchronize") && bib_cuda_ok(cudaMemcpy(desc, dev, sizeof desc, cudaMemcpyDeviceToHost), "cudaMemcpy");
    cudaFree(dev);
    if (!ok) return 0;
    for (int s = 0; s < 2; ++s) {
        if (!(s == 0 ? has_left : has_right)) continue;
        const HaloDesc& q = desc[1 + s];
        d->peer_vol[s] = q.vol; d->peer_nx[s] = q.nx;
        if (s == 1 && has_left && left == right) {                 // two slabs on a periodic axis: both neighbours are the same process, map it once
            for (int k = 0; k < 2; ++k) { d->peer_map[1][k] = d->peer_map[0][k]; d->peer_f[1][k] = d->peer_f[0][k]; }
            continue;
        }
        for (int k = 0; k < (q.shift ? 1 : 2); ++k) {
            void* p = nullptr;
            if (!bib_cuda_ok(cudaIpcOpenMemHandle(&p, q.h[k], cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle(neighbour's populations; BIB_HALO=nccl selects the ncclSend / ncclRecv halo)")) return 0;
            d->peer_map[s][k] = p;
        }
        if (q.shift) {                                             // f[1] = base, f[0] = base + shift planes (api.cu, bib_domain_create)
            d->peer_f[s][1] = (const double*)d->peer_map[s][0];
            d->peer_f[s][0] = (const double*)d->peer_map[s][0] + (size_t)q.shift * d->plane;
        } else { d->peer_f[s][0] = (const double*)d->peer_map[s][0]; d->peer_f[s][1] = (const double*)d->peer_map[s][1]; }
    }
    d->halo_pull = 1;
    return 1;
}

int comm_halo_exchange(bib_domain* d, cudaStream_t st) {
    if (d->n_ranks == 1) return 1;
    if (d->group) return local_halo_exchange(d, st);
    ncclComm_t comm = (ncclComm_t)d->nccl_comm;
    if (!comm) { bib_set_error("halo exchange: bib_comm_init has not been called"); return 0; }
    if (d->halo_pull < 0) {
        // (the set-up synchronises d->stream with the host; every rank reaches its first exchange at the same point of the program)
        if (!setup_halo_pull(d)) return 0;
    }
    if (d->halo_pull == 1) {
        // "my ghost planes are written" -> pull -> "I have pulled" (in place the storage of a frame's ghost planes is overwritten by the
        // launches of the very next step; with two buffers only by step t + 2, but the layouts of the ranks may differ - each picks its own
        // when the choice is automatic -, so the second token is always exchanged: ~10 us on the stream of the exchange)
        const bool has_right = d->per[0] || d->rank < d->n_ranks - 1;
        const bool has_left = d->per[0] || d->rank > 0;
        if (!nccl_token_sync(d, st)) return 0;
        if (!launch_halo_pull(d, st, has_left ? d->peer_f[0][d->cur] : nullptr, d->peer_vol[0], d->peer_nx[0], has_right ? d->peer_f[1][d->cur] : nullptr, d->peer_vol[1])) return 0;
        if (!nccl_token_sync(d, st)) return 0;
        return 1;
    }
    double* f = d->f[d->cur];           // the buffer just written
    const size_t n = d->plane;
    const int right = (d->rank + 1) % d->n_ranks, left = (d->rank + d->n_ranks - 1) % d->n_ranks;
    const bool has_right = d->per[0] || d->rank < d->n_ranks - 1;
    const bool has_left = d->per[0] || d->rank > 0;
    const bool staged = !d->per[1] || !d->per[2];
    if (staged && !d->halo_stage) {
        BIB_CHECK(cudaMalloc(&d->halo_stage, 10 * n * sizeof(double)));
        d->bytes += (long long)(10 * n * sizeof(double));
    }
    NCCL_CHECK(g_nccl.GroupStart());
    for (int k = 0; k < 5; ++k) {
        double* pop = f + (size_t)kPlusX[k] * d->vol;
        if (has_right) NCCL_CHECK(g_nccl.Send(pop + (size_t)(d->nx + 1) * n, n, ncclDouble, right, comm, st));
        if (has_left) NCCL_CHECK(g_nccl.Recv(staged ? d->halo_stage + (size_t)k * n : pop + (size_t)1 * n, n, ncclDouble, left, comm, st));
    }
    for (int k = 0; k < 5; ++k) {
        double* pop = f + (size_t)kMinusX[k] * d->vol;
        if (has_left) NCCL_CHECK(g_nccl.Send(pop + (size_t)0 * n, n, ncclDouble, left, comm, st));
        if (has_right) NCCL_CHECK(g_nccl.Recv(staged ? d->halo_stage + (size_t)(5 + k) * n : pop + (size_t)d->nx * n, n, ncclDouble, right, comm, st));
    }
    NCCL_CHECK(g_nccl.GroupEnd());
    d->launches++;
    if (staged) {
        const int grid = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
        for (int k = 0; k < 5; ++k) {
            if (has_left) { k_halo_merge<<<grid, 256, 0, st>>>(f + (size_t)kPlusX[k] * d->vol + (size_t)1 * n, d->halo_stage + (size_t)k * n, d->ny, d->nz, kCy[kPlusX[k]], kCz[kPlusX[k]], d->per[1], d->per[2]); d->launches++; }
            if (has_right) { k_halo_merge<<<grid, 256, 0, st>>>(f + (size_t)kMinusX[k] * d->vol + (size_t)d->nx * n, d->halo_stage + (size_t)(5 + k) * n, d->ny, d->nz, kCy[kMinusX[k]], kCz[kMinusX[k]], d->per[1], d->per[2]); d->launches++; }
        }
        BIB_CHECK(cudaGetLastError());
    }
    return 1;
}

// Symmetric neighbour exchange of two fixed-size buffers (used by the multi-slab immersed-boundary protocol).
int comm_exchange_pair(bib_domain* d, const double* send_left, const double* send_right, double* recv_left, double* recv_right, size_t n_doubles) {
    if (d->group) {                       // in-process slab group: pull the neighbours' send buffers over peer memory
        bib_group* g = d->group;
        const int right = (d->rank + 1) % d->n_ranks, left = (d->rank + d->n_ranks - 1) % d->n_ranks;
        const bool has_right = d->per[0] || d->rank < d->n_ranks - 1;
        const bool has_left = d->per[0] || d->rank > 0;
        d->x_send[0] = send_left; d->x_send[1] = send_right;
        BIB_CHECK(cudaEventRecord(d->ev_x_ready, d->stream));
        group_barrier(g);
        bib_domain* L = group_slab(g, left); bib_domain* R = group_slab(g, right);
        if (has_left) {
            BIB_CHECK(cudaStreamWaitEvent(d->stream, L->ev_x_ready, 0));
            BIB_CHECK(cudaMemcpyPeerAsync(recv_left, d->device, L->x_send[1], L->device, n_doubles * sizeof(double), d->stream));
        } else BIB_CHECK(cudaMemsetAsync(recv_left, 0, sizeof(double), d->stream));
        if (has_right) {
            BIB_CHECK(cudaStreamWaitEvent(d->stream, R->ev_x_ready, 0));
            BIB_CHECK(cudaMemcpyPeerAsync(recv_right, d->device, R->x_send[0], R->device, n_doubles * sizeof(double), d->stream));
        } else BIB_CHECK(cudaMemsetAsync(recv_right, 0, sizeof(double), d->stream));
        BIB_CHECK(cudaEventRecord(d->ev_x_done, d->stream));
        d->launches++;
        group_barrier(g);
        if (has_left) BIB_CHECK(cudaStreamWaitEvent(d->stream, L->ev_x_done, 0));      // my send buffers are free again once both neighbours pulled them
        if (has_right) BIB_CHECK(cudaStreamWaitEvent(d->stream, R->ev_x_done, 0));
        return 1;
    }
    ncclComm_t comm = (ncclComm_t)d->nccl_comm;
    if (!comm) { bib_set_error("exchange: bib_comm_init has not been called"); return 0; }
    const int right = (d->rank + 1) % d->n_ranks, left = (d->rank + d->n_ranks - 1) % d->n_ranks;
    const bool has_right = d->per[0] || d->rank < d->n_ranks - 1;
    const bool has_left = d->per[0] || d->rank > 0;
    if (!has_left) BIB_CHECK(cudaMemsetAsync(recv_left, 0, sizeof(double), d->stream));
    if (!has_right) BIB_CHECK(cudaMemsetAsync(recv_right, 0, sizeof(double), d->stream));
    NCCL_CHECK(g_nccl.GroupStart());
    if (has_left) { NCCL_CHECK(g_nccl.Send(send_left, n_doubles, ncclDouble, left, comm, d->stream)); NCCL_CHECK(g_nccl.Recv(recv_left, n_doubles, ncclDouble, left, comm, d->stream)); }
    if (has_right) { NCCL_CHECK(g_nccl.Send(send_right, n_doubles, ncclDouble, right, comm, d->stream)); NCCL_CHECK(g_nccl.Recv(recv_right, n_doubles, ncclDouble, right, comm, d->stream)); }
    NCCL_CHECK(g_nccl.GroupEnd());
    d->launches++;
    return 1;
}

// Stream-ordered neighbour synchronisation: work issued on d->stream after this call runs after everything the two neighbours issued on
// their streams before THEIR matching call. In-process group: CUDA events + the host barrier; one process per GPU: a grouped
// ncclSend / ncclRecv of one element with each neighbour (NCCL's own flag exchange over NVLink).
int comm_neighbour_sync(bib_domain* d) {
    const int right = (d->rank + 1) % d->n_ranks, left = (d->rank + d->n_ranks - 1) % d->n_ranks;
    const bool has_right = d->per[0] || d->rank < d->n_ranks - 1;
    const bool has_left = d->per[0] || d->rank > 0;
    if (d->group) {
        bib_group* g = d->group;
        BIB_CHECK(cudaEventRecord(d->ev_x_ready, d->stream));
        group_barrier(g);
        if (has_left) BIB_CHECK(cudaStreamWaitEvent(d->stream, group_slab(g, left)->ev_x_ready, 0));
        if (has_right) BIB_CHECK(cudaStreamWaitEvent(d->stream, group_slab(g, right)->ev_x_ready, 0));
        group_barrier(g);                      // nobody re-records its event before every neighbour has enqueued its wait
        return 1;
    }
    ncclComm_t comm = (ncclComm_t)d->nccl_comm;
    if (!comm) { bib_set_error("neighbour sync: bib_comm_init has not been called"); return 0; }
    if (!d->token) { BIB_CHECK(cudaMalloc(&d->token, 4 * sizeof(double))); BIB_CHECK(cudaMemset(d->token, 0, 4 * sizeof(double))); }
    NCCL_CHECK(g_nccl.GroupStart());
    if (has_left) { NCCL_CHECK(g_nccl.Send(d->token, 1, ncclDouble, left, comm, d->stream)); NCCL_CHECK(g_nccl.Recv(d->token + 2, 1, ncclDouble, left, comm, d->stream)); }
    if (has_right) { NCCL_CHECK(g_nccl.Send(d->token + 1, 1, ncclDouble, right, comm, d->stream)); NCCL_CHECK(g_nccl.Recv(d->token + 3, 1, ncclDouble, right, comm, d->stream)); }
    NCCL_CHECK(g_nccl.GroupEnd());
    d->launches++;
    return 1;
}

// Make the neighbours' exchange buffers readable from this slab's kernels (NVLink peer loads). One process per GPU: every rank exports
// its two buffers with cudaIpcGetMemHandle, the 64-byte handles travel over NCCL, cudaIpcOpenMemHandle maps them. COLLECTIVE.
int comm_map_peer_buffers(bib_domain* d, double* mine_for_left, double* mine_for_right, const double** from_left, const double** from_right) {
    const int right = (d->rank + 1) % d->n_ranks, left = (d->rank + d->n_ranks - 1) % d->n_ranks;
    const bool has_right = d->per[0] || d->rank < d->n_ranks - 1;
    const bool has_left = d->per[0] || d->rank > 0;
    *from_left = nullptr; *from_right = nullptr;
    if (d->group) return 1;                    // in-process: the caller reads the neighbour's pointers directly
    ncclComm_t comm = (ncclComm_t)d->nccl_comm;
    if (!comm) { bib_set_error("map peer buffers: bib_comm_init has not been called"); return 0; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h[4];                   // [0] mine for left, [1] mine for right, [2] left's for me, [3] right's for me
    memset(h, 0, sizeof h);
    BIB_CHECK(cudaIpcGetMemHandle(&h[0], mine_for_left));
    BIB_CHECK(cudaIpcGetMemHandle(&h[1], mine_for_right));
    char* dev = nullptr;
    BIB_CHECK(cudaMalloc(&dev, sizeof h));
    BIB_CHECK(cudaMemcpy(dev, h, sizeof h, cudaMemcpyHostToDevice));
    NCCL_CHECK(g_nccl.GroupStart());
    if (has_left) { NCCL_CHECK(g_nccl.Send(dev, 64, ncclChar, left, comm, d->stream)); NCCL_CHECK(g_nccl.Recv(dev + 128, 64, ncclChar, left, comm, d->stream)); }
    if (has_right) { NCCL_CHECK(g_nccl.Send(dev + 64, 64, ncclChar, right, comm, d->stream)); NCCL_CHECK(g_nccl.Recv(dev + 192, 64, ncclChar, right, comm, d->stream)); }
    NCCL_CHECK(g_nccl.GroupEnd());
    const bool copied = bib_cuda_ok(cudaStreamSynchronize(d->stream), "cudaStreamSynchronize") && bib_cuda_ok(cudaMemcpy(h, dev, sizeof h, cudaMemcpyDeviceToHost), "cudaMemcpy");
    cudaFree(dev);
    if (!copied) return 0;
    void* p = nullptr;
    if (has_left) { BIB_CHECK(cudaIpcOpenMemHandle(&p, h[2], cudaIpcMemLazyEnablePeerAccess)); *from_left = (const double*)p; }
    if (has_right) { BIB_CHECK(cudaIpcOpenMemHandle(&p, h[3], cudaIpcMemLazyEnablePeerAccess)); *from_right = (const double*)p; }
    return 1;
}
