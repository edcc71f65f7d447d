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
#include <cstring>
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

}  // namespace

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
    if (staged && !d->halo_stage) {
        BIB_CHECK(cudaMalloc(&d->halo_stage, 10 * n * sizeof(double)));
        d->bytes += (long long)(10 * n * sizeof(double));
    }
    BIB_CHECK(cudaEventRecord(d->ev_h_ready, st));
    group_barrier(g);
    bib_domain* L = group_slab(g, left); bib_domain* R = group_slab(g, right);
    const double* fl = L->f[L->cur]; const double* fr = R->f[R->cur];      // every slab flipped `cur` before the exchange
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

int comm_halo_exchange(bib_domain* d, cudaStream_t st) {
    if (d->n_ranks == 1) return 1;
    if (d->group) return local_halo_exchange(d, st);
    ncclComm_t comm = (ncclComm_t)d->nccl_comm;
    if (!comm) { bib_set_error("halo exchange: bib_comm_init has not been called"); return 0; }
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
