/* Single-rank MPI shim -- TEST INFRASTRUCTURE ONLY (oracle build).
 *
 * The reference's fsi package cannot be compiled in Palabos' serial mode (it uses
 * MPI_Request / MPI_Waitall / MPI_Wtime directly: src/fsi/CommunicationBuffer.h:75-76,
 * src/fsi/Profile.h:52), so the oracle is built with -DPLB_MPI_PARALLEL against this
 * one-rank stand-in. With one rank Palabos routes every overlap through local copies
 * (src/parallelism/parallelBlockCommunicator3D.cpp:207-218) and fsi::CommunicationBuffer
 * has an empty neighbour list (src/fsi/CommunicationBuffer.hh:22-23), so no point-to-point
 * call is ever reached; they abort() to make that assumption loud.
 */
#ifndef BIBLBM_ORACLE_MPI_SHIM_H
#define BIBLBM_ORACLE_MPI_SHIM_H
#include <cstring>
#include <ctime>
#include <cstdlib>
#include <cstdio>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Info;
typedef long long MPI_Offset;
typedef FILE* MPI_File;
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; };

#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF 1
#define MPI_SUCCESS 0
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_INFO_NULL 0
#define MPI_MODE_CREATE 1
#define MPI_MODE_WRONLY 2
#define MPI_MODE_RDONLY 4
#define MPI_SEEK_SET SEEK_SET

enum { MPI_CHAR = 1, MPI_BYTE = 2, MPI_INT = 4, MPI_LONG = 8, MPI_UNSIGNED_LONG = 9, MPI_FLOAT = 104, MPI_DOUBLE = 108,
       MPI_LONG_DOUBLE = 116, MPI_LONG_LONG = 208, MPI_UNSIGNED_LONG_LONG = 308, MPI_BOOL_SHIM = 401 };
enum { MPI_SUM = 1, MPI_MAX, MPI_MIN, MPI_LAND, MPI_PROD };

static inline int mpi_shim_size(MPI_Datatype t) {
    switch (t) {
        case MPI_CHAR: case MPI_BYTE: return 1;
        case MPI_INT: return sizeof(int);
        case MPI_LONG: case MPI_UNSIGNED_LONG: return sizeof(long);
        case MPI_FLOAT: return 4;
        case MPI_DOUBLE: return 8;
        case MPI_LONG_DOUBLE: return sizeof(long double);
        case MPI_LONG_LONG: case MPI_UNSIGNED_LONG_LONG: return 8;
    }
    return 1;
}
static inline int mpi_shim_unreachable(const char* what) {
    std::fprintf(stderr, "mpi shim: %s reached with a single rank\n", what);
    std::abort();
    return 1;
}

static inline int MPI_Init(int*, char***) { return 0; }
static inline int MPI_Finalize() { return 0; }
static inline int MPI_Abort(MPI_Comm, int c) { std::exit(c); return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int* s) { *s = 1; return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline double MPI_Wtime() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static inline int MPI_Send(void*, int, MPI_Datatype, int, int, MPI_Comm) { return mpi_shim_unreachable("MPI_Send"); }
static inline int MPI_Rsend(void*, int, MPI_Datatype, int, int, MPI_Comm) { return mpi_shim_unreachable("MPI_Rsend"); }
static inline int MPI_Recv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*) { return mpi_shim_unreachable("MPI_Recv"); }
static inline int MPI_Isend(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*) { return mpi_shim_unreachable("MPI_Isend"); }
static inline int MPI_Irecv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*) { return mpi_shim_unreachable("MPI_Irecv"); }
static inline int MPI_Sendrecv(void*, int, MPI_Datatype, int, int, void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*) { return mpi_shim_unreachable("MPI_Sendrecv"); }
static inline int MPI_Request_free(MPI_Request*) { return 0; }
static inline int MPI_Wait(MPI_Request*, MPI_Status*) { return 0; }
static inline int MPI_Waitall(int, MPI_Request*, MPI_Status*) { return 0; }
static inline int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return 0; }
static inline int MPI_Reduce(void* s, void* r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm) { std::memcpy(r, s, (size_t)n * mpi_shim_size(t)); return 0; }
static inline int MPI_Allreduce(void* s, void* r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) { std::memcpy(r, s, (size_t)n * mpi_shim_size(t)); return 0; }
static inline int MPI_Scatterv(void* s, int*, int* d, MPI_Datatype t, void* r, int n, MPI_Datatype, int, MPI_Comm) { std::memcpy(r, (char*)s + (size_t)d[0] * mpi_shim_size(t), (size_t)n * mpi_shim_size(t)); return 0; }
static inline int MPI_Gatherv(void* s, int n, MPI_Datatype t, void* r, int*, int* d, MPI_Datatype, int, MPI_Comm) { std::memcpy((char*)r + (size_t)d[0] * mpi_shim_size(t), s, (size_t)n * mpi_shim_size(t)); return 0; }
static inline int MPI_File_open(MPI_Comm, char* n, int m, MPI_Info, MPI_File* f) { *f = std::fopen(n, (m & MPI_MODE_RDONLY) ? "rb" : "wb"); return *f ? 0 : 1; }
static inline int MPI_File_seek(MPI_File f, MPI_Offset o, int w) { return std::fseek(f, (long)o, w); }
static inline int MPI_File_write(MPI_File f, void* b, int n, MPI_Datatype t, MPI_Status*) { std::fwrite(b, mpi_shim_size(t), n, f); return 0; }
static inline int MPI_File_read(MPI_File f, void* b, int n, MPI_Datatype t, MPI_Status*) { size_t r = std::fread(b, mpi_shim_size(t), n, f); (void)r; return 0; }
static inline int MPI_File_close(MPI_File* f) { std::fclose(*f); return 0; }
#endif
