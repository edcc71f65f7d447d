// lbm_kernels.cu -- fused D3Q19 collide-and-stream (one-step push, A/B buffers, SoA fp64) and the lattice side passes.
//
// Replaces (reference file:line):
//   GuoExternalForceVOFDynamics::collide            src/fsi/VariableViscosityDynamics.hh:151-172
//   bgk_ma2_collision_base (D3Q19)                  src/latticeBoltzmann/dynamicsTemplates3D.h:759-838
//   vofTemplatesImpl::addGuoForce / computeOmega    src/fsi/VariableViscosityDynamics.hh:47-141
//   BlockLattice3D::collideAndStream + periodic envelope copy
//                                                   src/atomicBlock/blockLattice3D.hh:252-312,402-427,500-551
//   BoxVelocityFunctional3D / SetExternalVectorFunctional3D
//                                                   src/dataProcessors/dataAnalysisFunctional3D.hh:772-784,
//                                                   src/dataProcessors/dataInitializerFunctional3D.hh:1144-1154
//
// Kernel design: one thread per cell, z (the contiguous axis) along threadIdx.x, so the 19 loads of
// a warp are 256-byte contiguous and aligned; the thread collides in registers and PUSHES the 19 post-collision values
// to its neighbours in the other buffer (stores shifted by 0/+-1 element in z, +-1 row in y, +-1 plane in x). The buffers
// hold pre-collision populations, exactly the state the reference keeps, so upload/download/velocity need no gather.
// Bound: HBM. Algorithmic traffic 19*8*2 = 304 B/cell (+24 B force +1 B phase actually read).
#include "domain.cuh"

#include <algorithm>
#include <unordered_map>
#include <vector>

namespace {

struct LbmParams {
    const double* __restrict__ fin;
    double* __restrict__ fout;
    const double* __restrict__ force;
    const uint8_t* __restrict__ phase;
    const uint8_t* __restrict__ kind;
    // sparse force field: when fflag != NULL a cell whose bit is 0 carries exactly the uniform body force (bfx, bfy, bfz) and its
    // three force components are NOT read (24 of the 329 compulsory bytes per cell; only the cells the IB spread touched are flagged)
    const unsigned* __restrict__ fflag;
    double bfx, bfy, bfz;
    int nx, ny, nz;
    long long plane, vol;
    double omega0, omega1;
    int per_y, per_z;
    // x handling: 0 = write into the ghost plane (neighbour slab or periodic image handled by the halo exchange),
    //             1 = wrap inside this slab (single slab, periodic x), 2 = wall face (non-periodic: free bounce-back)
    int x_lo_mode, x_hi_mode;
    // planes handled by this launch: x = x_first + blockIdx.z * x_stride (boundary launch: planes 0 and nx-1; interior: 1..nx-2)
    int x_first, x_stride;
    // fused envelope quirk (SWAP != 0): is this slab's low / high x face a GLOBAL periodic face
    int swap_x_lo, swap_x_hi;
    long long prefetch_cells;     // L2 prefetch distance of the main body (cells ahead in memory order; 0 = off)
    long long pf_end;             // first cell index the prefetch must not reach (the high ghost plane)
    int flat;                     // thread index runs over the flattened (y, z) plane (grid.y == 1)
};

// Cache-policy variants of the streaming accesses of k_collide_stream (experiments: build with -DBIB_LBM_LD=1/2, -DBIB_LBM_ST=1/2):
// loads  0 = ld.global.nc (read-only path, default), 1 = ld.global.cs (streaming, evict first), 2 = ld.global.lu (last use)
// stores 0 = st.global (default), 1 = st.global.cs (streaming), 2 = st.global.wt (write through)
#ifndef BIB_LBM_LD
#define BIB_LBM_LD 0
#endif
#ifndef BIB_LBM_ST
#define BIB_LBM_ST 0
#endif

__device__ __forceinline__ double d_ld_pop(const double* p) {
#if BIB_LBM_LD == 1
    return __ldcs(p);
#elif BIB_LBM_LD == 2
    return __ldlu(p);
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ void d_st_pop(double* p, double v) {
#if BIB_LBM_ST == 1
    __stcs(p, v);
#elif BIB_LBM_ST == 2
    __stwt(p, v);
#else
    *p = v;
#endif
}
__constant__ int c_cx[BIB_Q] = {0, -1, 0, 0, -1, -1, -1, -1, 0, 0, 1, 0, 0, 1, 1, 1, 1, 0, 0};
__constant__ int c_cy[BIB_Q] = {0, 0, -1, 0, -1, 1, 0, 0, -1, -1, 0, 1, 0, 1, -1, 0, 0, 1, 1};
__constant__ int c_cz[BIB_Q] = {0, 0, 0, -1, 0, 0, -1, 1, -1, 1, 0, 0, 1, 0, 0, 1, -1, 1, -1};

__device__ __forceinline__ int d_opp(int i) { return i == 0 ? 0 : (i <= 9 ? i + 9 : i - 9); }

// GuoExternalForceVOFDynamics::collide, operation order as in the reference (see oracle/biblbm_oracle.c for the
// line-by-line citations). f: 19 fBar values in/out; F: force; c: phase; (w0, w1) = (omega, omega2_) of the dynamics object.
__device__ __forceinline__ void d_collide(double* f, double Fx, double Fy, double Fz, double c, double w0, double w1) {
    // computeRhoBar: plain left-to-right sum (momentTemplates3D.h:375-382)
    double rhoBar = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8] + f[9] + f[10] + f[11] + f[12] + f[13] + f[14] +
                    f[15] + f[16] + f[17] + f[18];
    // computeVelocity: grouped sums, u = j/rho + F/2 (externalForceDynamics.hh:59-74)
    double rhoBar2, jx, jy, jz;
    d_rhoBar_j(f, rhoBar2, jx, jy, jz);
    const double invRho2 = 1.0 / (rhoBar2 + 1.0);
    const double ux = jx * invRho2 + Fx / 2.0;
    const double uy = jy * invRho2 + Fy / 2.0;
    const double uz = jz * invRho2 + Fz / 2.0;
    const double rho = rhoBar + 1.0;
    jx = rho * ux; jy = rho * uy; jz = rho * uz;
    // w1 == 0: GuoExternalForceBGKdynamics (one relaxation rate, used as it is; basicDynamics/externalForceDynamics.hh:215-234)
    const double omega = w1 == 0.0 ? w0 : w0 * w1 / (c * w0 + (1. - c) * w1);

    // bgk_ma2_collision_base
    const double invRho = 1.0 / (rhoBar + 1.0);
    const double one_m_omega = 1.0 - omega;
    const double t0_omega = (1.0 / 3.0) * omega;
    const double t1_omega = (1.0 / 18.0) * omega;
    const double t4_omega = (1.0 / 36.0) * omega;
    const double jSqr = jx * jx + jy * jy + jz * jz;
    const double kx = 3.0 * jx, ky = 3.0 * jy, kz = 3.0 * jz;
    const double kxSqr_ = invRho / 2.0 * kx * kx;
    const double kySqr_ = invRho / 2.0 * ky * ky;
    const double kzSqr_ = invRho / 2.0 * kz * kz;
    const double kxky_ = invRho * kx * ky;
    const double kxkz_ = invRho * kx * kz;
    const double kykz_ = invRho * ky * kz;
    const double C1 = rhoBar + invRho * 3.0 * jSqr;
    double C2, C3;
    C3 = -kxSqr_ - kySqr_ - kzSqr_;
    f[0] *= one_m_omega; f[0] += t0_omega * (C1 + C3);
    C2 = -kx; C3 = -kySqr_ - kzSqr_;
    f[1] *= one_m_omega; f[1] += t1_omega * (C1 + C2 + C3);
    f[10] *= one_m_omega; f[10] += t1_omega * (C1 - C2 + C3);
    C2 = -ky; C3 = -kxSqr_ - kzSqr_;
    f[2] *= one_m_omega; f[2] += t1_omega * (C1 + C2 + C3);
    f[11] *= one_m_omega; f[11] += t1_omega * (C1 - C2 + C3);
    C2 = -kz; C3 = -kxSqr_ - kySqr_;
    f[3] *= one_m_omega; f[3] += t1_omega * (C1 + C2 + C3);
    f[12] *= one_m_omega; f[12] += t1_omega * (C1 - C2 + C3);
    C2 = -kx - ky; C3 = kxky_ - kzSqr_;
    f[4] *= one_m_omega; f[4] += t4_omega * (C1 + C2 + C3);
    f[13] *= one_m_omega; f[13] += t4_omega * (C1 - C2 + C3);
    C2 = -kx + ky; C3 = -kxky_ - kzSqr_;
    f[5] *= one_m_omega; f[5] += t4_omega * (C1 + C2 + C3);
    f[14] *= one_m_omega; f[14] += t4_omega * (C1 - C2 + C3);
    C2 = -kx - kz; C3 = kxkz_ - kySqr_;
    f[6] *= one_m_omega; f[6] += t4_omega * (C1 + C2 + C3);
    f[15] *= one_m_omega; f[15] += t4_omega * (C1 - C2 + C3);
    C2 = -kx + kz; C3 = -kxkz_ - kySqr_;
    f[7] *= one_m_omega; f[7] += t4_omega * (C1 + C2 + C3);
    f[16] *= one_m_omega; f[16] += t4_omega * (C1 - C2 + C3);
    C2 = -ky - kz; C3 = kykz_ - kxSqr_;
    f[8] *= one_m_omega; f[8] += t4_omega * (C1 + C2 + C3);
    f[17] *= one_m_omega; f[17] += t4_omega * (C1 - C2 + C3);
    C2 = -ky + kz; C3 = -kykz_ - kxSqr_;
    f[9] *= one_m_omega; f[9] += t4_omega * (C1 + C2 + C3);
    f[18] *= one_m_omega; f[18] += t4_omega * (C1 - C2 + C3);

    // addGuoForce (amplitude 1)
    const double mu = 1.0 * (1.0 - omega / 2.0);
    const double oneOver6 = 1.0 / 6.0;
    const double oneOver12 = 1.0 / 12.0;
    f[0] += -mu * (Fx * ux + Fy * uy + Fz * uz);
    f[1] += oneOver6 * mu * (Fx * (-1.0 + 2 * ux) - Fy * uy - Fz * uz);
    f[2] += -oneOver6 * mu * (Fx * ux + Fy * (1.0 - 2 * uy) + Fz * uz);
    f[3] += -oneOver6 * mu * (Fx * ux + Fy * uy + Fz * (1.0 - 2 * uz));
    f[4] += oneOver12 * mu * (Fx * (-1.0 + 2 * ux + 3 * uy) + Fy * (-1.0 + 2 * uy + 3 * ux) - Fz * uz);
    f[5] += oneOver12 * mu * (Fx * (-1.0 + 2 * ux - 3 * uy) + Fy * (1.0 + 2 * uy - 3 * ux) - Fz * uz);
    f[6] += oneOver12 * mu * (Fx * (-1.0 + 2 * ux + 3 * uz) - Fy * uy + Fz * (-1.0 + 2 * uz + 3 * ux));
    f[7] += oneOver12 * mu * (Fx * (-1.0 + 2 * ux - 3 * uz) - Fy * uy + Fz * (1.0 + 2 * uz - 3 * ux));
    f[8] += -oneOver12 * mu * (Fx * ux + Fy * (1.0 - 2 * uy - 3 * uz) + Fz * (1.0 - 2 * uz - 3 * uy));
    f[9] += -oneOver12 * mu * (Fx * ux + Fy * (1.0 - 2 * uy + 3 * uz) + Fz * (-1.0 - 2 * uz + 3 * uy));
    f[10] += oneOver6 * mu * (Fx * (1.0 + 2 * ux) - Fy * uy - Fz * uz);
    f[11] += -oneOver6 * mu * (Fx * ux + Fy * (-1.0 - 2 * uy) + Fz * uz);
    f[12] += -oneOver6 * mu * (Fx * ux + Fy * uy + Fz * (-1.0 - 2 * uz));
    f[13] += oneOver12 * mu * (Fx * (1.0 + 2 * ux + 3 * uy) + Fy * (1.0 + 2 * uy + 3 * ux) - Fz * uz);
    f[14] += oneOver12 * mu * (Fx * (1.0 + 2 * ux - 3 * uy) + Fy * (-1.0 + 2 * uy - 3 * ux) - Fz * uz);
    f[15] += oneOver12 * mu * (Fx * (1.0 + 2 * ux + 3 * uz) - Fy * uy + Fz * (1.0 + 2 * uz + 3 * ux));
    f[16] += oneOver12 * mu * (Fx * (1.0 + 2 * ux - 3 * uz) - Fy * uy + Fz * (-1.0 + 2 * uz - 3 * ux));
    f[17] += -oneOver12 * mu * (Fx * ux + Fy * (-1.0 - 2 * uy - 3 * uz) + Fz * (-1.0 - 2 * uz - 3 * uy));
    f[18] += -oneOver12 * mu * (Fx * ux + Fy * (-1.0 - 2 * uy + 3 * uz) + Fz * (1.0 - 2 * uz + 3 * uy));
}

__device__ __forceinline__ void d_apply_kind(double* f, int kind, double Fx, double Fy, double Fz, double c, double w0, double w1) {
    if (kind == 0) {
        d_collide(f, Fx, Fy, Fz, c, w0, w1);
    } else if (kind == 2) {   // BounceBack::collide = swap i <-> i+9 (core/dynamics.hh:763-770)
#pragma unroll
        for (int i = 1; i <= 9; ++i) { double t = f[i]; f[i] = f[i + 9]; f[i + 9] = t; }
    }                         // NoDynamics: populations are streamed untouched (core/dynamics.hh:940-943)
}

// Destination of population i pushed from (x,y,z) [x,y,z slab-local, storage plane = x+1]; returns false when the
// population leaves through a non-periodic face (it then stays in the cell's own opposite slot), sets `crossed` when it
// goes through a periodic face handled inside this slab.
__device__ __forceinline__ bool d_dest(const LbmParams& p, int i, int x, int y, int z, long long& dst) {
    int tx = x + c_cx[i], ty = y + c_cy[i], tz = z + c_cz[i];
    if (tx < 0) { if (p.x_lo_mode == 1) tx = p.nx - 1; else if (p.x_lo_mode == 2) return false; }
    else if (tx >= p.nx) { if (p.x_hi_mode == 1) tx = 0; else if (p.x_hi_mode == 2) return false; }
    if (ty < 0) { if (p.per_y) ty = p.ny - 1; else return false; }
    else if (ty >= p.ny) { if (p.per_y) ty = 0; else return false; }
    if (tz < 0) { if (p.per_z) tz = p.nz - 1; else return false; }
    else if (tz >= p.nz) { if (p.per_z) tz = 0; else return false; }
    dst = (long long)(tx + 1) * p.plane + (long long)ty * p.nz + tz;
    return true;
}

// Per-thread neighbour offsets (relative to the cell's own index, in elements) with periodic wrap applied, plus wall
// flags for faces through which nothing streams (non-periodic block faces: the population stays in the cell's opposite slot).
struct Nbr {
    int oxm, oxp, oym, oyp, ozm, ozp;
    bool wxm, wxp, wym, wyp, wzm, wzp;
};
template <bool HAS_WALLS>
__device__ __forceinline__ Nbr d_neighbours(const LbmParams& p, int x, int y, int z) {
    Nbr n;
    const int plane = (int)p.plane;
    n.oxm = -plane; n.oxp = plane; n.oym = -p.nz; n.oyp = p.nz; n.ozm = -1; n.ozp = 1;
    n.wxm = n.wxp = n.wym = n.wyp = n.wzm = n.wzp = false;
    if (x == 0) { if (p.x_lo_mode == 1) n.oxm = (p.nx - 1) * plane; else if (HAS_WALLS && p.x_lo_mode == 2) n.wxm = true; }
    if (x == p.nx - 1) { if (p.x_hi_mode == 1) n.oxp = -(p.nx - 1) * plane; else if (HAS_WALLS && p.x_hi_mode == 2) n.wxp = true; }
    if (y == 0) { if (p.per_y) n.oym = (p.ny - 1) * p.nz; else if (HAS_WALLS) n.wym = true; }
    if (y == p.ny - 1) { if (p.per_y) n.oyp = -(p.ny - 1) * p.nz; else if (HAS_WALLS) n.wyp = true; }
    if (z == 0) { if (p.per_z) n.ozm = p.nz - 1; else if (HAS_WALLS) n.wzm = true; }
    if (z == p.nz - 1) { if (p.per_z) n.ozp = -(p.nz - 1); else if (HAS_WALLS) n.wzp = true; }
    return n;
}

// push population I (compile-time direction) of the cell at idx
template <int I, bool HAS_WALLS>
__device__ __forceinline__ void d_push(const LbmParams& p, const Nbr& n, long long idx, double v) {
    constexpr int cx = (I == 1 || I == 4 || I == 5 || I == 6 || I == 7) ? -1 : ((I == 10 || I == 13 || I == 14 || I == 15 || I == 16) ? 1 : 0);
    constexpr int cy = (I == 2 || I == 4 || I == 8 || I == 9 || I == 14) ? -1 : ((I == 5 || I == 11 || I == 13 || I == 17 || I == 18) ? 1 : 0);
    constexpr int cz = (I == 3 || I == 6 || I == 8 || I == 16 || I == 18) ? -1 : ((I == 7 || I == 9 || I == 12 || I == 15 || I == 17) ? 1 : 0);
    constexpr int OPP = I == 0 ? 0 : (I <= 9 ? I + 9 : I - 9);
    if (HAS_WALLS) {
        const bool blocked = (cx < 0 && n.wxm) || (cx > 0 && n.wxp) || (cy < 0 && n.wym) || (cy > 0 && n.wyp) || (cz < 0 && n.wzm) || (cz > 0 && n.wzp);
        if (blocked) { d_st_pop(p.fout + (long long)OPP * p.vol + idx, v); return; }
    }
    const int off = (cx < 0 ? n.oxm : (cx > 0 ? n.oxp : 0)) + (cy < 0 ? n.oym : (cy > 0 ? n.oyp : 0)) + (cz < 0 ? n.ozm : (cz > 0 ? n.ozp : 0));
    d_st_pop(p.fout + (long long)I * p.vol + idx + off, v);
}

// re-push of population I after the swapped-omega collision: only when it crosses a periodic face (a population that also
// leaves through a non-periodic face stays bounced from the bulk collision)
template <int I, bool HAS_WALLS, bool WITH_Z>
__device__ __forceinline__ void d_push_crossing(const LbmParams& p, const Nbr& n, long long idx, double v, bool sxm, bool sxp, bool sym, bool syp, bool szm, bool szp) {
    constexpr int cx = (I == 1 || I == 4 || I == 5 || I == 6 || I == 7) ? -1 : ((I == 10 || I == 13 || I == 14 || I == 15 || I == 16) ? 1 : 0);
    constexpr int cy = (I == 2 || I == 4 || I == 8 || I == 9 || I == 14) ? -1 : ((I == 5 || I == 11 || I == 13 || I == 17 || I == 18) ? 1 : 0);
    constexpr int cz = (I == 3 || I == 6 || I == 8 || I == 16 || I == 18) ? -1 : ((I == 7 || I == 9 || I == 12 || I == 15 || I == 17) ? 1 : 0);
    const bool cross = (cx < 0 && sxm) || (cx > 0 && sxp) || (cy < 0 && sym) || (cy > 0 && syp) || (WITH_Z && ((cz < 0 && szm) || (cz > 0 && szp)));
    if (!cross) return;
    if (HAS_WALLS) {
        const bool blocked = (cx < 0 && n.wxm) || (cx > 0 && n.wxp) || (cy < 0 && n.wym) || (cy > 0 && n.wyp) || (cz < 0 && n.wzm) || (cz > 0 && n.wzp);
        if (blocked) return;
    }
    const int off = (cx < 0 ? n.oxm : (cx > 0 ? n.oxp : 0)) + (cy < 0 ? n.oym : (cy > 0 ? n.oyp : 0)) + (cz < 0 ? n.ozm : (cz > 0 ? n.ozp : 0));
    p.fout[(long long)I * p.vol + idx + off] = v;
}

// SWAP: 0 = the periodic-face fix-up runs as a separate kernel (k_face_swap); 1 = fused for all faces; 2 = fused for the cells of the
// x and y faces (whole planes / rows of threads take the tail; they also re-push what crosses z), the cells that lie on a z face only
// (one lane of two warps per row) stay in k_face_swap.
template <bool HAS_KIND, bool HAS_WALLS, int SWAP>
#ifndef BIB_LBM_MINB
#define BIB_LBM_MINB 5
#endif
__global__ void __launch_bounds__(128, (HAS_KIND && HAS_WALLS) ? 4 : BIB_LBM_MINB) k_collide_stream(const LbmParams p) {
    // threads along z (the contiguous axis), one row per blockIdx.y; p.flat: threads along the flattened (y, z) plane instead, so that a
    // row length that is not a multiple of the CTA size (the reference's 97 / 61 / 121-cell lattices) leaves no idle lanes
    int y, z;
    if (p.flat) {
        const int t = blockIdx.x * blockDim.x + threadIdx.x;
        if (t >= (int)p.plane) return;
        y = t / p.nz; z = t - y * p.nz;
    } else {
        z = blockIdx.x * blockDim.x + threadIdx.x;
        if (z >= p.nz) return;
        y = blockIdx.y;
    }
    const int x = p.x_first + (int)blockIdx.z * p.x_stride;
    const long long idx = (long long)(x + 1) * p.plane + (long long)y * p.nz + z;

    if (p.prefetch_cells > 0) {
        // Pull the populations that CTAs a fraction of a wave later will load into L2 (prefetch_cells ahead in memory order: 32 rows of 1024
        // cells = 256 KB per population stream; the hint holds no register): the kernel is latency-bound at 5 CTAs / SM, an L2 hit shortens
        // the wait of those warps. Measured on the bench slab: 6.82 -> 6.66 ms (16 / 64 rows: 6.70; 128 rows: 7.04; 256 rows: 8.67, evicted before use).
        const long long pf = idx + p.prefetch_cells;
        if (pf < p.pf_end) {
#pragma unroll
            for (int i = 0; i < BIB_Q; ++i) asm volatile("prefetch.global.L2 [%0];" :: "l"(p.fin + (long long)i * p.vol + pf));
        }
    }
    double f[BIB_Q];
#pragma unroll
    for (int i = 0; i < BIB_Q; ++i) f[i] = d_ld_pop(p.fin + (long long)i * p.vol + idx);
    double Fx = p.bfx, Fy = p.bfy, Fz = p.bfz;
    if (!p.fflag || ((__ldg(p.fflag + (idx >> 5)) >> (idx & 31)) & 1u)) { Fx = __ldg(p.force + idx); Fy = __ldg(p.force + p.vol + idx); Fz = __ldg(p.force + 2 * p.vol + idx); }
    const double c = (double)(__ldg(p.phase + idx) & 1);
    const int kind = HAS_KIND ? (int)__ldg(p.kind + idx) : 0;
    const bool on_face = SWAP != 0 && (((x == 0) && p.swap_x_lo) || ((x == p.nx - 1) && p.swap_x_hi) || ((y == 0 || y == p.ny - 1) && p.per_y) ||
                                       (SWAP == 1 && (z == 0 || z == p.nz - 1) && p.per_z));
    d_apply_kind(f, kind, Fx, Fy, Fz, c, p.omega0, p.omega1);

    const Nbr n = d_neighbours<HAS_WALLS>(p, x, y, z);
    d_push<0, HAS_WALLS>(p, n, idx, f[0]);
    d_push<1, HAS_WALLS>(p, n, idx, f[1]);   d_push<10, HAS_WALLS>(p, n, idx, f[10]);
    d_push<2, HAS_WALLS>(p, n, idx, f[2]);   d_push<11, HAS_WALLS>(p, n, idx, f[11]);
    d_push<3, HAS_WALLS>(p, n, idx, f[3]);   d_push<12, HAS_WALLS>(p, n, idx, f[12]);
    d_push<4, HAS_WALLS>(p, n, idx, f[4]);   d_push<13, HAS_WALLS>(p, n, idx, f[13]);
    d_push<5, HAS_WALLS>(p, n, idx, f[5]);   d_push<14, HAS_WALLS>(p, n, idx, f[14]);
    d_push<6, HAS_WALLS>(p, n, idx, f[6]);   d_push<15, HAS_WALLS>(p, n, idx, f[15]);
    d_push<7, HAS_WALLS>(p, n, idx, f[7]);   d_push<16, HAS_WALLS>(p, n, idx, f[16]);
    d_push<8, HAS_WALLS>(p, n, idx, f[8]);   d_push<17, HAS_WALLS>(p, n, idx, f[17]);
    d_push<9, HAS_WALLS>(p, n, idx, f[9]);   d_push<18, HAS_WALLS>(p, n, idx, f[18]);

    if (SWAP != 0) {
        // Envelope quirk fused (see k_face_swap below): a cell on a periodic face collides a second time with (omega, omega_inner)
        // swapped and re-pushes only the populations that cross such a face. Same thread, same addresses as the pushes above, so
        // program order makes the second value the one that stays.
        // Everything in the tail is re-derived from re-read thread coordinates (loads hit L1/L2): keeping the pre-collision values, the
        // force, the coordinates or the neighbour offsets alive across the pushes would cost EVERY thread registers for the sake of
        // the few face cells.
        if (on_face) {
            unsigned tid, bx, by, bz;
            asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
            asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bx));
            asm volatile("mov.u32 %0, %%ctaid.y;" : "=r"(by));
            asm volatile("mov.u32 %0, %%ctaid.z;" : "=r"(bz));
            int z2 = (int)(bx * blockDim.x + tid), y2 = (int)by;
            if (p.flat) { y2 = z2 / p.nz; z2 -= y2 * p.nz; }
            const int x2 = p.x_first + (int)bz * p.x_stride;
            const bool sxm = (x2 == 0) && p.swap_x_lo, sxp = (x2 == p.nx - 1) && p.swap_x_hi;
            const bool sym = (y2 == 0) && p.per_y, syp = (y2 == p.ny - 1) && p.per_y;
            const bool szm = (z2 == 0) && p.per_z, szp = (z2 == p.nz - 1) && p.per_z;
            const long long idx2 = (long long)(x2 + 1) * p.plane + (long long)y2 * p.nz + z2;
            double g[BIB_Q];
#pragma unroll
            for (int i = 0; i < BIB_Q; ++i) g[i] = __ldcg(p.fin + (long long)i * p.vol + idx2);
            double Gx = p.bfx, Gy = p.bfy, Gz = p.bfz;
            if (!p.fflag || ((__ldcg(p.fflag + (idx2 >> 5)) >> (idx2 & 31)) & 1u)) { Gx = __ldcg(p.force + idx2); Gy = __ldcg(p.force + p.vol + idx2); Gz = __ldcg(p.force + 2 * p.vol + idx2); }
            const double c2 = (double)(__ldcg(p.phase + idx2) & 1);
            const int kind2 = HAS_KIND ? (int)__ldcg(p.kind + idx2) : 0;
            const Nbr m = d_neighbours<HAS_WALLS>(p, x2, y2, z2);
            if (SWAP == 1 && !(sxm || sxp || sym || syp)) {
                // z faces only (one lane of the first / last warp of every row): just the five populations that cross are needed, the
                // rest of the second collision is dead code in these two branches
                if (szm) {
                    d_apply_kind(g, kind2, Gx, Gy, Gz, c2, p.omega1, p.omega0);
                    d_push_crossing<3, HAS_WALLS, true>(p, m, idx2, g[3], false, false, false, false, true, false);
                    d_push_crossing<6, HAS_WALLS, true>(p, m, idx2, g[6], false, false, false, false, true, false);
                    d_push_crossing<8, HAS_WALLS, true>(p, m, idx2, g[8], false, false, false, false, true, false);
                    d_push_crossing<16, HAS_WALLS, true>(p, m, idx2, g[16], false, false, false, false, true, false);
                    d_push_crossing<18, HAS_WALLS, true>(p, m, idx2, g[18], false, false, false, false, true, false);
                } else {
                    d_apply_kind(g, kind2, Gx, Gy, Gz, c2, p.omega1, p.omega0);
                    d_push_crossing<7, HAS_WALLS, true>(p, m, idx2, g[7], false, false, false, false, false, true);
                    d_push_crossing<9, HAS_WALLS, true>(p, m, idx2, g[9], false, false, false, false, false, true);
                    d_push_crossing<12, HAS_WALLS, true>(p, m, idx2, g[12], false, false, false, false, false, true);
                    d_push_crossing<15, HAS_WALLS, true>(p, m, idx2, g[15], false, false, false, false, false, true);
                    d_push_crossing<17, HAS_WALLS, true>(p, m, idx2, g[17], false, false, false, false, false, true);
                }
                return;
            }
            d_apply_kind(g, kind2, Gx, Gy, Gz, c2, p.omega1, p.omega0);
            d_push_crossing<1, HAS_WALLS, true>(p, m, idx2, g[1], sxm, sxp, sym, syp, szm, szp);   d_push_crossing<10, HAS_WALLS, true>(p, m, idx2, g[10], sxm, sxp, sym, syp, szm, szp);
            d_push_crossing<2, HAS_WALLS, true>(p, m, idx2, g[2], sxm, sxp, sym, syp, szm, szp);   d_push_crossing<11, HAS_WALLS, true>(p, m, idx2, g[11], sxm, sxp, sym, syp, szm, szp);
            d_push_crossing<3, HAS_WALLS, true>(p, m, idx2, g[3], sxm, sxp, sym, syp, szm, szp);   d_push_crossing<12, HAS_WALLS, true>(p, m, idx2, g[12], sxm, sxp, sym, syp, szm, szp);
            d_push_crossing<4, HAS_WALLS, true>(p, m, idx2, g[4], sxm, sxp, sym, syp, szm, szp);   d_push_crossing<13, HAS_WALLS, true>(p, m, idx2, g[13], sxm, sxp, sym, syp, szm, szp);
            d_push_crossing<5, HAS_WALLS, true>(p, m, idx2, g[5], sxm, sxp, sym, syp, szm, szp);   d_push_crossing<14, HAS_WALLS, true>(p, m, idx2, g[14], sxm, sxp, sym, syp, szm, szp);
            d_push_crossing<6, HAS_WALLS, true>(p, m, idx2, g[6], sxm, sxp, sym, syp, szm, szp);   d_push_crossing<15, HAS_WALLS, true>(p, m, idx2, g[15], sxm, sxp, sym, syp, szm, szp);
            d_push_crossing<7, HAS_WALLS, true>(p, m, idx2, g[7], sxm, sxp, sym, syp, szm, szp);   d_push_crossing<16, HAS_WALLS, true>(p, m, idx2, g[16], sxm, sxp, sym, syp, szm, szp);
            d_push_crossing<8, HAS_WALLS, true>(p, m, idx2, g[8], sxm, sxp, sym, syp, szm, szp);   d_push_crossing<17, HAS_WALLS, true>(p, m, idx2, g[17], sxm, sxp, sym, syp, szm, szp);
            d_push_crossing<9, HAS_WALLS, true>(p, m, idx2, g[9], sxm, sxp, sym, syp, szm, szp);   d_push_crossing<18, HAS_WALLS, true>(p, m, idx2, g[18], sxm, sxp, sym, syp, szm, szp);
        }
    }
}

// Envelope quirk: populations that cross a periodic face come from a collision with (omega, omega_inner) swapped
// (DESIGN.md "reference quirks"; core/dynamicsIdentifiers.h:131-141 of the reference). Runs after k_collide_stream over the
// cells of the periodic faces only and overwrites just the crossing populations.
struct FaceParams {
    LbmParams p;
    int face_dim;     // 0,1,2
    int face_side;    // 0 = low, 1 = high
    int swap_x_lo, swap_x_hi;   // is this slab's x face a GLOBAL periodic face
    int face_mask;              // bit (2*dim + side): faces handled by this launch
    int x_select;               // 0: every plane, 1: only the slab's boundary planes x = 0 / nx-1, 2: only the interior planes
};

template <bool HAS_KIND>
__global__ void __launch_bounds__(128) k_face_swap(const FaceParams q) {
    const LbmParams& p = q.p;
    // all six faces in one launch: blockIdx.z = 2*dim + side; (a, b) enumerate the two other dimensions
    FaceParams qq = q;
    qq.face_dim = blockIdx.z >> 1; qq.face_side = blockIdx.z & 1;
    {
        const bool active = qq.face_dim == 0 ? (qq.face_side == 0 ? q.swap_x_lo : q.swap_x_hi) : (qq.face_dim == 1 ? p.per_y : p.per_z);
        if (!active || !((q.face_mask >> blockIdx.z) & 1)) return;
    }
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (b >= (qq.face_dim == 0 ? p.ny : p.nx)) return;
    int x, y, z;
    if (q.face_dim == -1) return;
    if (qq.face_dim == 0) { x = qq.face_side ? p.nx - 1 : 0; y = b; z = a; if (a >= p.nz) return; }
    else if (qq.face_dim == 1) { y = qq.face_side ? p.ny - 1 : 0; x = b; z = a; if (a >= p.nz) return; }
    else { z = qq.face_side ? p.nz - 1 : 0; x = b; y = a; if (a >= p.ny) return; }
    if (q.x_select == 1 && x != 0 && x != p.nx - 1) return;
    if (q.x_select == 2 && (x == 0 || x == p.nx - 1)) return;
    // a cell on several periodic faces is handled once, by the first (dim, side) that contains it
    const bool on_xlo = (x == 0) && q.swap_x_lo, on_xhi = (x == p.nx - 1) && q.swap_x_hi;
    const bool on_ylo = (y == 0) && p.per_y, on_yhi = (y == p.ny - 1) && p.per_y;
    const bool on_zlo = (z == 0) && p.per_z, on_zhi = (z == p.nz - 1) && p.per_z;
    if (qq.face_dim == 0 && qq.face_side == 1 && on_xlo) return;
    if (qq.face_dim == 1 && (on_xlo || on_xhi || (qq.face_side == 1 && on_ylo))) return;
    if (qq.face_dim == 2 && (on_xlo || on_xhi || on_ylo || on_yhi || (qq.face_side == 1 && on_zlo))) return;

    const long long idx = (long long)(x + 1) * p.plane + (long long)y * p.nz + z;
    double f[BIB_Q];
#pragma unroll
    for (int i = 0; i < BIB_Q; ++i) f[i] = __ldg(p.fin + (long long)i * p.vol + idx);
    double Fx = p.bfx, Fy = p.bfy, Fz = p.bfz;
    if (!p.fflag || ((__ldg(p.fflag + (idx >> 5)) >> (idx & 31)) & 1u)) { Fx = __ldg(p.force + idx); Fy = __ldg(p.force + p.vol + idx); Fz = __ldg(p.force + 2 * p.vol + idx); }
    const double c = (double)(__ldg(p.phase + idx) & 1);
    const int kind = HAS_KIND ? (int)__ldg(p.kind + idx) : 0;
    d_apply_kind(f, kind, Fx, Fy, Fz, c, p.omega1, p.omega0);   // swapped
#pragma unroll
    for (int i = 1; i < BIB_Q; ++i) {
        const int tx = x + c_cx[i], ty = y + c_cy[i], tz = z + c_cz[i];
        const bool cross = (tx < 0 && q.swap_x_lo) || (tx >= p.nx && q.swap_x_hi) ||
                           ((ty < 0 || ty >= p.ny) && p.per_y) || ((tz < 0 || tz >= p.nz) && p.per_z);
        if (!cross) continue;
        long long dst;
        if (d_dest(p, i, x, y, z, dst)) p.fout[(long long)i * p.vol + dst] = f[i];
        // (a population that also leaves through a non-periodic face is bounced from the bulk collision: untouched)
    }
}

// ---------------------------------------------------------------------------------------------- transposes & side passes
// host AoS [x][y][z][23] (staged in device memory) <-> device SoA. One thread per (cell, component).
__global__ void k_unpack_cells(const double* __restrict__ aos, double* f, double* force, uint8_t* phase, int* bad_phase,
                               int bx0, int by0, int bz0, int bnx, int bny, int bnz, long long plane, long long vol, int nz, int ncomp) {
    const long long n = (long long)bnx * bny * bnz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(t % bnz), y = (int)((t / bnz) % bny), x = (int)(t / ((long long)bnz * bny));
        const long long idx = (long long)(x + bx0 + 1) * plane + (long long)(y + by0) * nz + (z + bz0);
        const double* c = aos + t * ncomp;
        for (int i = 0; i < BIB_Q; ++i) f[(long long)i * vol + idx] = c[i];
        for (int k = 0; k < 3; ++k) force[(long long)k * vol + idx] = c[BIB_Q + k];
        const double ph = ncomp > BIB_Q + 3 ? c[BIB_Q + 3] : 0.0;
        if (ph != 0.0 && ph != 1.0) atomicAdd(bad_phase, 1);
        phase[idx] = (ph != 0.0) ? 1 : 0;
    }
}
__global__ void k_pack_cells(double* __restrict__ aos, const double* f, const double* force, const uint8_t* phase,
                             int bx0, int by0, int bz0, int bnx, int bny, int bnz, long long plane, long long vol, int nz, int ncomp) {
    const long long n = (long long)bnx * bny * bnz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(t % bnz), y = (int)((t / bnz) % bny), x = (int)(t / ((long long)bnz * bny));
        const long long idx = (long long)(x + bx0 + 1) * plane + (long long)(y + by0) * nz + (z + bz0);
        double* c = aos + t * ncomp;
        for (int i = 0; i < BIB_Q; ++i) c[i] = f[(long long)i * vol + idx];
        for (int k = 0; k < 3; ++k) c[BIB_Q + k] = force[(long long)k * vol + idx];
        if (ncomp > BIB_Q + 3) c[BIB_Q + 3] = (double)(phase[idx] & 1);
    }
}
__global__ void k_set_external(double* force, int bx0, int by0, int bz0, int bnx, int bny, int bnz, long long plane, long long vol, int nz,
                               double fx, double fy, double fz) {
    const long long n = (long long)bnx * bny * bnz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(t % bnz), y = (int)((t / bnz) % bny), x = (int)(t / ((long long)bnz * bny));
        const long long idx = (long long)(x + bx0 + 1) * plane + (long long)(y + by0) * nz + (z + bz0);
        force[idx] = fx; force[vol + idx] = fy; force[2 * vol + idx] = fz;
    }
}
// NoDynamics / BounceBack cells report zero velocity (their computeVelocity, src/core/dynamics.hh:763-770,940-943)
__global__ void k_velocity(const double* __restrict__ f, const double* __restrict__ force, double* vel, const uint8_t* __restrict__ kind,
                           int bx0, int by0, int bz0, int bnx, int bny, int bnz, long long plane, long long vol, int nz) {
    const long long n = (long long)bnx * bny * bnz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(t % bnz), y = (int)((t / bnz) % bny), x = (int)(t / ((long long)bnz * bny));
        const long long idx = (long long)(x + bx0 + 1) * plane + (long long)(y + by0) * nz + (z + bz0);
        if (kind && kind[idx] != 0) { vel[idx] = 0.0; vel[vol + idx] = 0.0; vel[2 * vol + idx] = 0.0; continue; }
        double c[BIB_Q];
#pragma unroll
        for (int i = 0; i < BIB_Q; ++i) c[i] = f[(long long)i * vol + idx];
        double rhoBar, jx, jy, jz;
        d_rhoBar_j(c, rhoBar, jx, jy, jz);
        const double invRho = 1.0 / (rhoBar + 1.0);
        vel[idx] = jx * invRho + force[idx] / 2.0;
        vel[vol + idx] = jy * invRho + force[vol + idx] / 2.0;
        vel[2 * vol + idx] = jz * invRho + force[2 * vol + idx] / 2.0;
    }
}
__global__ void k_pack_velocity(double* __restrict__ out, const double* vel, int bx0, int by0, int bz0, int bnx, int bny, int bnz,
                                long long plane, long long vol, int nz) {
    const long long n = (long long)bnx * bny * bnz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(t % bnz), y = (int)((t / bnz) % bny), x = (int)(t / ((long long)bnz * bny));
        const long long idx = (long long)(x + bx0 + 1) * plane + (long long)(y + by0) * nz + (z + bz0);
        out[t * 3] = vel[idx]; out[t * 3 + 1] = vel[vol + idx]; out[t * 3 + 2] = vel[2 * vol + idx];
    }
}
__global__ void k_unpack_velocity(const double* __restrict__ in, double* vel, int bx0, int by0, int bz0, int bnx, int bny, int bnz,
                                  long long plane, long long vol, int nz) {
    const long long n = (long long)bnx * bny * bnz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(t % bnz), y = (int)((t / bnz) % bny), x = (int)(t / ((long long)bnz * bny));
        const long long idx = (long long)(x + bx0 + 1) * plane + (long long)(y + by0) * nz + (z + bz0);
        vel[idx] = in[t * 3]; vel[vol + idx] = in[t * 3 + 1]; vel[2 * vol + idx] = in[t * 3 + 2];
    }
}
// CouetteVelocityExtrapolator (progs/common/shear_flow_setup.h:121-183): u[plane] = 2 u[plane + inward] - u[plane + 2 inward], per component
__global__ void k_extrapolate_plane(double* vel, int dir, int coord, int inward, int nx, int ny, int nz, long long plane, long long vol) {
    const int na = dir == 0 ? ny : nx, nb = dir == 2 ? ny : nz;
    const long long n = (long long)na * nb;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(t / nb), b = (int)(t % nb);
        int c[3], c1[3], c2[3];
        c[dir] = coord; c1[dir] = coord + inward; c2[dir] = coord + 2 * inward;
        const int ta = dir == 0 ? 1 : 0, tb = dir == 2 ? 1 : 2;
        c[ta] = c1[ta] = c2[ta] = a; c[tb] = c1[tb] = c2[tb] = b;
        const long long i0 = (long long)(c[0] + 1) * plane + (long long)c[1] * nz + c[2];
        const long long i1 = (long long)(c1[0] + 1) * plane + (long long)c1[1] * nz + c1[2];
        const long long i2 = (long long)(c2[0] + 1) * plane + (long long)c2[1] * nz + c2[2];
#pragma unroll
        for (int k = 0; k < 3; ++k) vel[k * vol + i0] = 2.0 * vel[k * vol + i1] - vel[k * vol + i2];
    }
}
__global__ void k_set_kind(uint8_t* kind, int bx0, int by0, int bz0, int bnx, int bny, int bnz, long long plane, int nz, uint8_t v) {
    const long long n = (long long)bnx * bny * bnz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(t % bnz), y = (int)((t / bnz) % bny), x = (int)(t / ((long long)bnz * bny));
        kind[(long long)(x + bx0 + 1) * plane + (long long)(y + by0) * nz + (z + bz0)] = v;
    }
}

struct Eq19 { double v[BIB_Q]; };
__global__ void k_fill_populations(double* f, Eq19 eq, int bx0, int by0, int bz0, int bnx, int bny, int bnz, long long plane, long long vol, int nz) {
    const long long n = (long long)bnx * bny * bnz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(t % bnz), y = (int)((t / bnz) % bny), x = (int)(t / ((long long)bnz * bny));
        const long long idx = (long long)(x + bx0 + 1) * plane + (long long)(y + by0) * nz + (z + bz0);
#pragma unroll
        for (int i = 0; i < BIB_Q; ++i) f[(long long)i * vol + idx] = eq.v[i];
    }
}

// PlaneFdBoundaryFunctional3D::processCell + VelocityDirichletBoundaryDynamics::computeRhoBar on one plane (citations and the
// operation order: oracle/biblbm_oracle.c, orc_wall_plane). One thread per plane cell; reads the post-stream populations of
// the wall cell and of its fluid neighbour, writes the regularised populations of the wall cell.
struct WallParams {
    double* f; const double* force; const double* wallvel;
    int nx, ny, nz; long long plane, vol;
    int dir, orient, coord, na, nb, gx;
    int a0, b0, ca, cb;          // covered cells of the face: a = a0 .. a0+ca-1, b = b0 .. b0+cb-1
    double omega;
};
__global__ void __launch_bounds__(128) k_wall_plane(const WallParams w) {
    const int bb = blockIdx.x * blockDim.x + threadIdx.x, a = w.a0 + (int)blockIdx.y;
    if (bb >= w.cb) return;
    const int b = w.b0 + bb;
    const int ta = w.dir == 0 ? 1 : 0, tb = w.dir == 2 ? 1 : 2;
    int c[3]; c[w.dir] = w.coord; c[ta] = a; c[tb] = b;
    const long long idx = (long long)(c[0] + 1) * w.plane + (long long)c[1] * w.nz + c[2];
    double f[BIB_Q];
#pragma unroll
    for (int i = 0; i < BIB_Q; ++i) f[i] = w.f[(long long)i * w.vol + idx];
    const double* vel = w.wallvel + ((long long)(a + w.gx) * w.nb + b) * 3;
    const double v[3] = {vel[0], vel[1], vel[2]};
    const double Fd = w.force[(long long)w.dir * w.vol + idx];
    double rhoOnWall = 0.0, rhoNormal = 0.0;
#pragma unroll
    for (int i = 0; i < BIB_Q; ++i) { const int ci = w.dir == 0 ? c_cx[i] : (w.dir == 1 ? c_cy[i] : c_cz[i]); if (ci == 0) rhoOnWall += f[i]; }
#pragma unroll
    for (int i = 0; i < BIB_Q; ++i) { const int ci = w.dir == 0 ? c_cx[i] : (w.dir == 1 ? c_cy[i] : c_cz[i]); if (ci == w.orient) rhoNormal += f[i]; }
    const double velNormal = (double)w.orient * (v[w.dir] - 0.5 * Fd);
    const double rhoBarW = (2.0 * rhoNormal + rhoOnWall - 1 * velNormal) / (1.0 + velNormal);
    const double rho = rhoBarW + 1.0;
    const int n[3] = {w.nx, w.ny, w.nz};
    double grad[3][3];
#pragma unroll
    for (int dd = 0; dd < 3; ++dd) {
        if (dd == w.dir) {
            int m[3] = {c[0], c[1], c[2]}; m[w.dir] = w.coord - w.orient;
            const long long mi = (long long)(m[0] + 1) * w.plane + (long long)m[1] * w.nz + m[2];
            double g[BIB_Q];
#pragma unroll
            for (int i = 0; i < BIB_Q; ++i) g[i] = w.f[(long long)i * w.vol + mi];
            double rb, jx, jy, jz;
            d_rhoBar_j(g, rb, jx, jy, jz);
            const double invRho = 1.0 / (rb + 1.0);
            const double u1[3] = {jx * invRho + w.force[mi] / 2.0, jy * invRho + w.force[w.vol + mi] / 2.0, jz * invRho + w.force[2 * w.vol + mi] / 2.0};
            for (int k = 0; k < 3; ++k) grad[dd][k] = -w.orient * (-v[k] + u1[k]);
        } else {
            int p[3] = {c[0], c[1], c[2]}, q[3] = {c[0], c[1], c[2]};
            if (dd == 0 && w.gx) { p[0] = c[0] + 1; q[0] = c[0] - 1; }      // multi-slab: the neighbouring slabs' columns are stored as ghost columns
            else { p[dd] = (c[dd] + 1) % n[dd]; q[dd] = (c[dd] + n[dd] - 1) % n[dd]; }
            const double* up = w.wallvel + ((long long)(p[ta] + w.gx) * w.nb + p[tb]) * 3;
            const double* um = w.wallvel + ((long long)(q[ta] + w.gx) * w.nb + q[tb]) * 3;
            for (int k = 0; k < 3; ++k) grad[dd][k] = (up[k] - um[k]) / 2.0;
        }
    }
    const double sToPi = -rho / 3.0 / w.omega;
    double pi[6];
    pi[0] = 2.0 * grad[0][0] * sToPi; pi[3] = 2.0 * grad[1][1] * sToPi; pi[5] = 2.0 * grad[2][2] * sToPi;
    pi[1] = (grad[0][1] + grad[1][0]) * sToPi; pi[2] = (grad[0][2] + grad[2][0]) * sToPi; pi[4] = (grad[1][2] + grad[2][1]) * sToPi;
    const double j[3] = {rho * v[0], rho * v[1], rho * v[2]};
    const double jSqr = j[0] * j[0] + j[1] * j[1] + j[2] * j[2];
    const double rhoBar = rho - 1.0;
    const double invRho = 1.0 / (rhoBar + 1.0);
#pragma unroll
    for (int i = 0; i < BIB_Q; ++i) {
        const int cc[3] = {c_cx[i], c_cy[i], c_cz[i]};
        const double ti = i == 0 ? 1. / 3 : ((i >= 1 && i <= 3) || (i >= 10 && i <= 12) ? 1. / 18 : 1. / 36);
        const double c_j = cc[0] * j[0] + cc[1] * j[1] + cc[2] * j[2];
        const double eq = ti * (rhoBar + 3. * c_j + invRho * (4.5 * c_j * c_j - 1.5 * jSqr));
        double fNeq = 0.0; int iPi = 0;
#pragma unroll
        for (int al = 0; al < 3; ++al) {
            fNeq += pi[iPi] * (cc[al] * cc[al] - 1.0 / 3.0);
            ++iPi;
#pragma unroll
            for (int be = al + 1; be < 3; ++be) { fNeq += pi[iPi] * 2.0 * cc[al] * cc[be]; ++iPi; }
        }
        fNeq *= ti * 3.0 * 3.0 / 2.0;
        w.f[(long long)i * w.vol + idx] = eq + fNeq;
    }
}
// computeVelocity on a wall cell returns the stored boundary velocity (StoreVelocityDynamics::computeVelocity)
__global__ void k_wall_velocity_fill(double* vel, const double* __restrict__ wallvel, int dir, int coord, int na, int nb, int a0, int b0, int cb,
                                     long long plane, long long vol, int nz, int gx) {
    const int bb = blockIdx.x * blockDim.x + threadIdx.x, a = a0 + (int)blockIdx.y;
    if (bb >= cb) return;
    const int b = b0 + bb;
    const int ta = dir == 0 ? 1 : 0, tb = dir == 2 ? 1 : 2;
    int c[3]; c[dir] = coord; c[ta] = a; c[tb] = b;
    const long long idx = (long long)(c[0] + 1) * plane + (long long)c[1] * nz + c[2];
    const double* v = wallvel + ((long long)(a + gx) * nb + b) * 3;
    vel[idx] = v[0]; vel[vol + idx] = v[1]; vel[2 * vol + idx] = v[2];
}

// VelocityDirichletBoundaryDynamics<dir,orient>::computeRhoBar of a wall-plane cell (the formula k_wall_plane evaluates for its own cell)
__device__ __forceinline__ double d_wall_rhoBar(const double* f, double Fd, double vd, int dir, int orient) {
    double rhoOnWall = 0.0, rhoNormal = 0.0;
#pragma unroll
    for (int i = 0; i < BIB_Q; ++i) { const int ci = dir == 0 ? c_cx[i] : (dir == 1 ? c_cy[i] : c_cz[i]); if (ci == 0) rhoOnWall += f[i]; }
#pragma unroll
    for (int i = 0; i < BIB_Q; ++i) { const int ci = dir == 0 ? c_cx[i] : (dir == 1 ? c_cy[i] : c_cz[i]); if (ci == orient) rhoNormal += f[i]; }
    const double velNormal = (double)orient * (vd - 0.5 * Fd);
    return (2.0 * rhoNormal + rhoOnWall - 1 * velNormal) / (1.0 + velNormal);
}
// OuterVelocityEdgeFunctional3D<axis,n1,n2>::processCell (citations and operation order: oracle/biblbm_oracle.c, orc_wall_edge). One thread
// per edge cell. The density is extrapolated from the two wall cells next to the edge in each of the two directions (whose density is the
// velocity-Dirichlet formula of the face they lie on), the velocity gradient comes from stored boundary velocities only.
struct EdgeParams {
    double* f; const double* force; const double* vel_e; const double* vel_p1; const double* vel_p2;   // edge [n][3]; faces normal to direction2 / direction1
    int nx, ny, nz; long long plane, vol;
    int axis, n1, n2, c1, c2, e0, count, n, per_axis;
    int p1_na, p1_nb, p2_na, p2_nb;
    double omega;
};
__device__ __forceinline__ const double* d_face_vel(const double* vel, int dir, int nb, const int* c) {
    const int ta = dir == 0 ? 1 : 0, tb = dir == 2 ? 1 : 2;
    return vel + ((long long)c[ta] * nb + c[tb]) * 3;
}
__global__ void __launch_bounds__(128) k_wall_edge(const EdgeParams w) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= w.count) return;
    const int d1 = (w.axis + 1) % 3, d2 = (w.axis + 2) % 3;
    int c[3]; c[w.axis] = w.e0 + t; c[d1] = w.c1; c[d2] = w.c2;
    double rho_n[2][2];
    for (int which = 0; which < 2; ++which) for (int step = 1; step <= 2; ++step) {
        int m[3] = {c[0], c[1], c[2]};
        if (which == 0) m[d1] += -w.n1 * step; else m[d2] += -w.n2 * step;
        const long long mi = (long long)(m[0] + 1) * w.plane + (long long)m[1] * w.nz + m[2];
        double g[BIB_Q];
#pragma unroll
        for (int i = 0; i < BIB_Q; ++i) g[i] = w.f[(long long)i * w.vol + mi];
        const int fdir = which == 0 ? d2 : d1, forient = which == 0 ? w.n2 : w.n1;
        const double* fv = which == 0 ? d_face_vel(w.vel_p1, d2, w.p1_nb, m) : d_face_vel(w.vel_p2, d1, w.p2_nb, m);
        rho_n[which][step - 1] = d_wall_rhoBar(g, w.force[(long long)fdir * w.vol + mi], fv[fdir], fdir, forient) + 1.0;
    }
    const double rho10 = rho_n[0][0], rho01 = rho_n[1][0], rho20 = rho_n[0][1], rho02 = rho_n[1][1];
    const double rho = 2.0 / 3.0 * (rho01 + rho10) - 1.0 / 6.0 * (rho02 + rho20);
    const double* ve = w.vel_e + (long long)c[w.axis] * 3;
    const double v[3] = {ve[0], ve[1], ve[2]};
    double grad[3][3];
    {
        int ip = c[w.axis] + 1, im = c[w.axis] - 1;
        if (w.per_axis) { ip = (ip + w.n) % w.n; im = (im + w.n) % w.n; }
        const double* up = w.vel_e + (long long)ip * 3; const double* um = w.vel_e + (long long)im * 3;
        for (int k = 0; k < 3; ++k) grad[w.axis][k] = (up[k] - um[k]) / 2.0;
    }
    for (int which = 0; which < 2; ++which) {
        const int dd = which == 0 ? d1 : d2, orient = which == 0 ? w.n1 : w.n2;
        int m[3] = {c[0], c[1], c[2]}; m[dd] -= orient;
        const double* u1 = which == 0 ? d_face_vel(w.vel_p1, d2, w.p1_nb, m) : d_face_vel(w.vel_p2, d1, w.p2_nb, m);
        for (int k = 0; k < 3; ++k) grad[dd][k] = -orient * (-v[k] + u1[k]);
    }
    const double sToPi = -rho / 3.0 / w.omega;
    double pi[6];
    pi[0] = 2.0 * grad[0][0] * sToPi; pi[3] = 2.0 * grad[1][1] * sToPi; pi[5] = 2.0 * grad[2][2] * sToPi;
    pi[1] = (grad[0][1] + grad[1][0]) * sToPi; pi[2] = (grad[0][2] + grad[2][0]) * sToPi; pi[4] = (grad[1][2] + grad[2][1]) * sToPi;
    const double j[3] = {rho * v[0], rho * v[1], rho * v[2]};
    const double jSqr = j[0] * j[0] + j[1] * j[1] + j[2] * j[2];
    const double rhoBar = rho - 1.0;
    const double invRho = 1.0 / (rhoBar + 1.0);
    const long long idx = (long long)(c[0] + 1) * w.plane + (long long)c[1] * w.nz + c[2];
#pragma unroll
    for (int i = 0; i < BIB_Q; ++i) {
        const int cc[3] = {c_cx[i], c_cy[i], c_cz[i]};
        const double ti = i == 0 ? 1. / 3 : ((i >= 1 && i <= 3) || (i >= 10 && i <= 12) ? 1. / 18 : 1. / 36);
        const double c_j = cc[0] * j[0] + cc[1] * j[1] + cc[2] * j[2];
        const double eq = ti * (rhoBar + 3. * c_j + invRho * (4.5 * c_j * c_j - 1.5 * jSqr));
        double fNeq = 0.0; int iPi = 0;
#pragma unroll
        for (int al = 0; al < 3; ++al) {
            fNeq += pi[iPi] * (cc[al] * cc[al] - 1.0 / 3.0);
            ++iPi;
#pragma unroll
            for (int be = al + 1; be < 3; ++be) { fNeq += pi[iPi] * 2.0 * cc[al] * cc[be]; ++iPi; }
        }
        fNeq *= ti * 3.0 * 3.0 / 2.0;
        w.f[(long long)i * w.vol + idx] = eq + fNeq;
    }
}
__global__ void k_edge_velocity_fill(double* vel, const double* __restrict__ vel_e, int axis, int c1, int c2, int e0, int count, long long plane, long long vol, int nz) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int d1 = (axis + 1) % 3, d2 = (axis + 2) % 3;
    int c[3]; c[axis] = e0 + t; c[d1] = c1; c[d2] = c2;
    const long long idx = (long long)(c[0] + 1) * plane + (long long)c[1] * nz + c[2];
    const double* v = vel_e + (long long)c[axis] * 3;
    vel[idx] = v[0]; vel[vol + idx] = v[1]; vel[2 * vol + idx] = v[2];
}

// In-place streaming on one slab with periodic x: the populations pushed into the ghost planes re-enter on the other side. A slot whose
// upstream cell lies beyond a y / z wall was filled by the local bounce and is left alone (same rule as k_halo_merge between slabs).
__global__ void k_ghost_wrap(double* f, long long vol, long long plane, int nx, int ny, int nz, int per_y, int per_z) {
    const int plus_x[5] = {10, 13, 14, 15, 16}, minus_x[5] = {1, 4, 5, 6, 7};
    const long long n = 10 * plane;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t / plane);
        const long long c = t - (long long)k * plane;
        const int y = (int)(c / nz), z = (int)(c - (long long)y * nz);
        const int i = k < 5 ? plus_x[k] : minus_x[k - 5];
        const int sy = y - c_cy[i], sz = z - c_cz[i];
        if ((!per_y && (sy < 0 || sy >= ny)) || (!per_z && (sz < 0 || sz >= nz))) continue;
        double* pop = f + (long long)i * vol;
        if (k < 5) pop[plane + c] = pop[(long long)(nx + 1) * plane + c];
        else pop[(long long)nx * plane + c] = pop[c];
    }
}

bool box_ok(const bib_domain* d, const int box[6]) {
    return box[0] >= 0 && box[1] < d->nx && box[0] <= box[1] && box[2] >= 0 && box[3] < d->ny && box[2] <= box[3] &&
           box[4] >= 0 && box[5] < d->nz && box[4] <= box[5];
}
int grid_for(long long n, int bs) {
    long long g = (n + bs - 1) / bs;
    const long long cap = 148LL * 16;   // grid-stride loops: a few waves over the 148 SMs
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}
bool ensure_scratch(bib_domain* d, size_t bytes) {
    if (d->scratch_bytes >= bytes) return true;
    if (d->scratch) cudaFree(d->scratch);
    d->scratch = nullptr; d->scratch_bytes = 0;
    if (!bib_cuda_ok(cudaMalloc(&d->scratch, bytes), "cudaMalloc(scratch)")) return false;
    d->scratch_bytes = bytes;
    return true;
}

}  // namespace

// --------------------------------------------------------------------------------------------------- host launchers
static LbmParams make_params(bib_domain* d) {
    LbmParams p;
    p.fin = d->f[d->cur]; p.fout = d->f[1 - d->cur];
    p.force = d->force; p.phase = d->phase; p.kind = d->kind;
    const bool sparse = d->sparse_force && d->fflag && d->fflag_valid && d->force_is_body;
    p.fflag = sparse ? d->fflag : nullptr;
    p.bfx = d->force_body[0]; p.bfy = d->force_body[1]; p.bfz = d->force_body[2];
    p.nx = d->nx; p.ny = d->ny; p.nz = d->nz;
    p.plane = (long long)d->plane; p.vol = (long long)d->vol;
    p.omega0 = d->omega0; p.omega1 = d->omega1;
    p.per_y = d->per[1]; p.per_z = d->per[2];
    const bool first = d->rank == 0, last = d->rank == d->n_ranks - 1;
    // (in place: a periodic x of a single slab goes through the ghost planes like a slab face - a push wrapped inside the array would land
    // on planes the later launches of the step still have to read - and lbm_ghost_wrap carries the ghost planes over afterwards)
    if (d->n_ranks == 1) { p.x_lo_mode = p.x_hi_mode = d->per[0] ? (d->inplace_shift ? 0 : 1) : 2; }
    else {
        p.x_lo_mode = (first && !d->per[0]) ? 2 : 0;
        p.x_hi_mode = (last && !d->per[0]) ? 2 : 0;
    }
    p.x_first = 0; p.x_stride = 1;
    p.swap_x_lo = d->per[0] && first; p.swap_x_hi = d->per[0] && last;
    // only worth it when the lattice does not fit the 126 MB L2 anyway (A + B populations of 400 k cells)
    p.prefetch_cells = (d->lbm_prefetch > 0 && d->vol > (size_t)400000) ? d->lbm_prefetch : 0;
    p.pf_end = (long long)(d->nx + 1) * (long long)d->plane;
    p.flat = (d->nz % 128 != 0 && d->plane < ((size_t)1 << 30)) ? 1 : 0;
    return p;
}

// 0: separate k_face_swap launches; 1: fused into k_collide_stream for all faces; 2: x and y faces fused, z faces separate
static int fused_swap_mode(const bib_domain* d) {
    if (!d->envelope_swap || d->omega0 == d->omega1 || d->omega1 == 0.0) return 0;      // equal rates / one rate: the swap changes nothing
    if (!(d->per[0] || d->per[1] || d->per[2])) return 0;
    if (d->inplace_shift) return 1;      // a separate fix-up pass would re-read source planes the step has already overwritten
    return d->fused_face_swap;
}
static void launch_main(bib_domain* d, const LbmParams& p, int n_planes, cudaStream_t st) {
    if (n_planes <= 0) return;
    int bs = d->nz >= 128 ? 128 : ((d->nz + 31) / 32) * 32;
    dim3 grid((d->nz + bs - 1) / bs, d->ny, n_planes);
    if (p.flat) { bs = 128; grid = dim3((unsigned)((d->plane + bs - 1) / bs), 1, n_planes); }
    const bool walls = p.x_lo_mode == 2 || p.x_hi_mode == 2 || !p.per_y || !p.per_z;
    const int swap = fused_swap_mode(d);
#define BIB_LAUNCH_CS(K, W) do { if (swap == 1) k_collide_stream<K, W, 1><<<grid, bs, 0, st>>>(p); else if (swap == 2) k_collide_stream<K, W, 2><<<grid, bs, 0, st>>>(p); \
                                 else k_collide_stream<K, W, 0><<<grid, bs, 0, st>>>(p); } while (0)
    if (d->kind) { if (walls) BIB_LAUNCH_CS(true, true); else BIB_LAUNCH_CS(true, false); }
    else { if (walls) BIB_LAUNCH_CS(false, true); else BIB_LAUNCH_CS(false, false); }
#undef BIB_LAUNCH_CS
    d->launches++;
}
static void launch_faces(bib_domain* d, const LbmParams& p, int face_mask, int x_select, cudaStream_t st) {
    if (!d->envelope_swap || d->omega0 == d->omega1 || d->omega1 == 0.0) return;        // equal rates / one rate: the swapped collision is the same collision
    const int fused = fused_swap_mode(d);
    if (fused == 1) return;
    if (fused == 2) face_mask &= 48;                                // the z faces only
    FaceParams q; q.p = p; q.x_select = x_select;
    const bool first = d->rank == 0, last = d->rank == d->n_ranks - 1;
    q.swap_x_lo = d->per[0] && first; q.swap_x_hi = d->per[0] && last;
    int active = (q.swap_x_lo ? 1 : 0) | (q.swap_x_hi ? 2 : 0) | (d->per[1] ? 12 : 0) | (d->per[2] ? 48 : 0);
    q.face_mask = face_mask & active;
    if (!q.face_mask) return;
    q.face_dim = 0; q.face_side = 0;
    const int na = std::max(d->ny, d->nz), nb = std::max(d->nx, d->ny);
    dim3 g((na + 127) / 128, nb, 6);
    if (d->kind) k_face_swap<true><<<g, 128, 0, st>>>(q);
    else k_face_swap<false><<<g, 128, 0, st>>>(q);
    d->launches++;
}

// In-place streaming. The source frame f[cur] and the destination frame f[1 - cur] are the same allocation
// displaced by S = inplace_shift planes: storage plane of (frame 0, x) = x + 1 + S, of (frame 1, x) = x + 1. A launch over the planes
// [a, a + C) reads storage planes a + 1 + S .. a + C + S and pushes into a .. a + C + 1 (frame 0 -> 1; mirrored for 1 -> 0), so with
// C <= S - 1 it never writes a plane it reads, and taking the launches in ascending x (frame 0 -> 1; descending for 1 -> 0) every plane a
// launch overwrites has been read by an earlier launch of the same stream. Every destination slot is still written exactly once.
// (host only; exported as bib_inplace_schedule so that the CPU tests check the invariants on the launch list the library really uses)
void lbm_inplace_schedule(int nx, int shift, int cur, std::vector<std::pair<int, int>>& launches) {
    launches.clear();
    const int C = shift - 1;
    const bool ascending = cur == 0;
    for (int done = 0; done < nx; done += C) {
        const int n = std::min(C, nx - done);
        launches.emplace_back(ascending ? done : nx - done - n, n);      // (first plane, planes)
    }
}
static int launch_inplace(bib_domain* d, const LbmParams& p) {
    std::vector<std::pair<int, int>> launches;
    lbm_inplace_schedule(d->nx, d->inplace_shift, d->cur, launches);
    for (const std::pair<int, int>& l : launches) {
        LbmParams pc = p;
        pc.x_first = l.first;
        pc.x_stride = 1;
        launch_main(d, pc, l.second, d->stream);
    }
    BIB_CHECK(cudaGetLastError());
    return 1;
}
// single slab, periodic x, in place: what the halo exchange does between slabs (ghost plane nx+1 -> plane 1, ghost plane 0 -> plane nx)
static int lbm_ghost_wrap(bib_domain* d) {
    const long long n = 10 * (long long)d->plane;
    k_ghost_wrap<<<grid_for(n, 256), 256, 0, d->stream>>>(d->f[d->cur], (long long)d->vol, (long long)d->plane, d->nx, d->ny, d->nz, d->per[1], d->per[2]);
    d->launches++;
    BIB_CHECK(cudaGetLastError());
    return 1;
}

// One lattice step. Single slab: one launch over all planes (+ the periodic-face fix-up). Several slabs: the two boundary
// planes go first, their ghost planes are exchanged over NCCL on the communication stream WHILE the interior planes are
// processed on the compute stream (the exchange writes populations c_x=+-1 of planes 1 / nx, the interior launch writes
// the other populations of those planes: disjoint arrays).
int lbm_launch_collide_stream(bib_domain* d) {
    LbmParams p = make_params(d);
    if (d->inplace_shift) {
        // the halo of several slabs is not hidden in this mode (the plane that feeds one of the two ghost planes is processed last); it is
        // meant for slabs too large for two buffers, where the exchange is ~1 % of the step
        if (!launch_inplace(d, p)) return 0;
        pt_mark(d, PT_LBM);
        d->cur = 1 - d->cur;
        if (d->n_ranks > 1) { if (!comm_halo_exchange(d, d->stream)) return 0; pt_mark(d, PT_HALO_WAIT); }
        else if (d->per[0] && !lbm_ghost_wrap(d)) return 0;
        const int okw = lbm_apply_walls(d) && lbm_apply_guo(d);
        pt_mark(d, PT_WALLS);
        return okw;
    }
    if (d->n_ranks == 1 || d->nx < 3) {
        launch_main(d, p, d->nx, d->stream);
        launch_faces(d, p, 63, 0, d->stream);
        BIB_CHECK(cudaGetLastError());
        pt_mark(d, PT_LBM);
        d->cur = 1 - d->cur;
        if (d->n_ranks > 1 && !comm_halo_exchange(d, d->stream)) return 0;
        if (d->n_ranks > 1) pt_mark(d, PT_HALO_WAIT);
        const int okw = lbm_apply_walls(d) && lbm_apply_guo(d);
        pt_mark(d, PT_WALLS);
        return okw;      // level-0, then level-1 internal processors run after streaming (and after the halo is in place)
    }
    LbmParams pb = p; pb.x_first = 0; pb.x_stride = d->nx - 1;
    launch_main(d, pb, 2, d->stream);
    launch_faces(d, p, 63, 1, d->stream);                   // every periodic-face fix-up of the boundary planes precedes the exchange
    pt_mark(d, PT_LBM_BOUNDARY);
    BIB_CHECK(cudaEventRecord(d->ev_compute, d->stream));
    BIB_CHECK(cudaStreamWaitEvent(d->comm_stream, d->ev_compute, 0));
    d->cur = 1 - d->cur;
    pt_mark(d, PT_BEGIN, d->comm_stream);
    if (!comm_halo_exchange(d, d->comm_stream)) return 0;
    pt_mark(d, PT_HALO, d->comm_stream);
    BIB_CHECK(cudaEventRecord(d->ev_halo, d->comm_stream));
    d->cur = 1 - d->cur;
    LbmParams pi = p; pi.x_first = 1; pi.x_stride = 1;
    launch_main(d, pi, d->nx - 2, d->stream);
    launch_faces(d, p, 60, 2, d->stream);                   // y and z faces of the interior planes
    BIB_CHECK(cudaGetLastError());
    pt_mark(d, PT_LBM);
    d->cur = 1 - d->cur;
    BIB_CHECK(cudaStreamWaitEvent(d->stream, d->ev_halo, 0));
    pt_mark(d, PT_HALO_WAIT);
    const int okw = lbm_apply_walls(d) && lbm_apply_guo(d);
    pt_mark(d, PT_WALLS);
    return okw;
}

static bool wall_box(const bib_domain* d, const int box[6], int dir) {
    return box_ok(d, box) && dir >= 0 && dir <= 2 && box[2 * dir] == box[2 * dir + 1];
}
int lbm_add_wall(bib_domain* d, const int box[6], int dir, int orient) {
    if (!wall_box(d, box, dir) || (orient != 1 && orient != -1)) { bib_set_error("add_velocity_boundary: box must be one plane normal to direction; orientation +-1"); return 0; }
    const int ta = dir == 0 ? 1 : 0, tb = dir == 2 ? 1 : 2;
    const int n[3] = {d->nx, d->ny, d->nz};
    // along a non-periodic in-plane axis the plane must stop short of the lattice border: its border cells differentiate the stored
    // velocities of their in-plane neighbours (edge cells, extensional_flow_setup.h:66-88); along a periodic axis it spans the lattice
    for (int k = 0; k < 2; ++k) {
        const int t = k == 0 ? ta : tb;
        if (d->per[t]) { if (box[2 * t] != 0 || box[2 * t + 1] != n[t] - 1) { bib_set_error("add_velocity_boundary: the plane must span every periodic in-plane direction"); return 0; } }
        else if (box[2 * t] < 1 || box[2 * t + 1] > n[t] - 2) { bib_set_error("add_velocity_boundary: along a non-periodic in-plane direction the plane must leave the border cells to edge conditions"); return 0; }
    }
    if (d->n_ranks > 1 && (dir == 0 || !d->per[0])) { bib_set_error("add_velocity_boundary: on multi-slab domains only planes normal to y / z with periodic x"); return 0; }
    const int gxc = d->n_ranks > 1 ? 1 : 0;
    for (const WallPlane& o : d->walls) if (o.dir == dir && o.coord == box[2 * dir]) { bib_set_error("add_velocity_boundary: a velocity boundary already exists on this plane"); return 0; }
    WallPlane w; w.dir = dir; w.orient = orient; w.coord = box[2 * dir]; w.na = n[ta]; w.nb = n[tb];
    w.a0 = box[2 * ta]; w.a1 = box[2 * ta + 1]; w.b0 = box[2 * tb]; w.b1 = box[2 * tb + 1];
    w.gx = gxc;
    w.h_vel.assign((size_t)(w.na + 2 * w.gx) * w.nb * 3, 0.0);
    BIB_CHECK(cudaMalloc(&w.vel, w.h_vel.size() * sizeof(double)));
    BIB_CHECK(cudaMemsetAsync(w.vel, 0, w.h_vel.size() * sizeof(double), d->stream));
    d->walls.push_back(w);
    return 1;
}
// addExternalVelocityEdge{axis}{N|P}{N|P} (boundaryCondition/boundaryInstantiator3D.h:423-469 with the interp manager): the two faces the
// edge joins must already carry velocity planes (their cells next to the edge provide density and velocity neighbours)
int lbm_add_edge(bib_domain* d, const int box[6], int axis, int normal1, int normal2) {
    if (!box_ok(d, box) || axis < 0 || axis > 2 || (normal1 != 1 && normal1 != -1) || (normal2 != 1 && normal2 != -1)) { bib_set_error("add_velocity_edge: bad arguments"); return 0; }
    const int d1 = (axis + 1) % 3, d2 = (axis + 2) % 3;
    const int n[3] = {d->nx, d->ny, d->nz};
    if (box[2 * d1] != box[2 * d1 + 1] || box[2 * d2] != box[2 * d2 + 1]) { bib_set_error("add_velocity_edge: the box must be a line along the axis"); return 0; }
    if (d->n_ranks > 1) { bib_set_error("add_velocity_edge: single-slab domains only"); return 0; }
    WallEdge e; e.axis = axis; e.n1 = normal1; e.n2 = normal2; e.c1 = box[2 * d1]; e.c2 = box[2 * d2]; e.e0 = box[2 * axis]; e.e1 = box[2 * axis + 1]; e.n = n[axis];
    if (e.c1 != (normal1 < 0 ? 0 : n[d1] - 1) || e.c2 != (normal2 < 0 ? 0 : n[d2] - 1)) { bib_set_error("add_velocity_edge: the edge must lie on the lattice border its normals point out of"); return 0; }
    if (d->per[axis] ? (e.e0 != 0 || e.e1 != n[axis] - 1) : (e.e0 < 1 || e.e1 > n[axis] - 2)) { bib_set_error("add_velocity_edge: the edge must span a periodic axis / stop short of the corners of a non-periodic one"); return 0; }
    for (size_t i = 0; i < d->walls.size(); ++i) {
        const WallPlane& w = d->walls[i];
        if (w.dir == d2 && w.coord == e.c2 && w.orient == normal2) e.plane1 = (int)i;
        if (w.dir == d1 && w.coord == e.c1 && w.orient == normal1) e.plane2 = (int)i;
    }
    if (e.plane1 < 0 || e.plane2 < 0 || n[d1] < 4 || n[d2] < 4) { bib_set_error("add_velocity_edge: add the velocity planes of the two faces the edge joins first"); return 0; }
    e.h_vel.assign((size_t)e.n * 3, 0.0);
    BIB_CHECK(cudaMalloc(&e.vel, (size_t)e.n * 3 * sizeof(double)));
    BIB_CHECK(cudaMemsetAsync(e.vel, 0, (size_t)e.n * 3 * sizeof(double), d->stream));
    d->edges.push_back(e);
    return 1;
}
// setBoundaryVelocity(lattice, box, u): every boundary cell inside the box stores its velocity; a face array also keeps the velocities
// of the edge cells on its border (they are the in-plane neighbours of the plane's outermost cells)
// `box` / `vel` in coordinates whose x = 0 is global plane gx0 of this slab's frame shifted by -gx0 (gx0 = 0: slab-local coordinates).
// On a multi-slab domain the box may reach one column beyond either x-end of the slab (x = -1, x = nx): the stored velocities of the
// neighbouring slabs' wall cells, which the plane functional differentiates along x; columns never supplied copy the nearest own column
// (exact for a wall velocity that does not depend on x). gnx_wrap != 0: x of the box is global and periodic with that period.
int lbm_set_wall_velocity_ext(bib_domain* d, const int box[6], const double* vel, int gx0, int gnx_wrap) {
    const int by = box[3] - box[2] + 1, bz = box[5] - box[4] + 1;
    if (box[0] > box[1] || box[2] < 0 || box[3] >= d->ny || box[2] > box[3] || box[4] < 0 || box[5] >= d->nz || box[4] > box[5]) { bib_set_error("set_boundary_velocity: box outside the lattice"); return 0; }
    bool any = false;
    for (WallPlane& w : d->walls) {
        if (w.coord < box[2 * w.dir] || w.coord > box[2 * w.dir + 1]) continue;
        const int ta = w.dir == 0 ? 1 : 0, tb = w.dir == 2 ? 1 : 2;
        std::vector<char> have(w.na + 2 * w.gx, 0);
        const int alo = ta == 0 ? -w.gx : 0, ahi = ta == 0 ? w.na - 1 + w.gx : w.na - 1;
        for (int a = alo; a <= ahi; ++a) {
            int ga = a;                                    // coordinate of row a in the box's frame
            if (ta == 0) {
                ga = a + gx0;
                if (gnx_wrap) ga = ((ga % gnx_wrap) + gnx_wrap) % gnx_wrap;
            }
            if (ga < box[2 * ta] || ga > box[2 * ta + 1]) continue;
            for (int b = std::max(0, box[2 * tb]); b <= std::min(w.nb - 1, box[2 * tb + 1]); ++b) {
                int c[3]; c[w.dir] = w.coord; c[ta] = ga; c[tb] = b;
                const double* src = vel + (((size_t)(c[0] - box[0]) * by + (c[1] - box[2])) * bz + (c[2] - box[4])) * 3;
                double* dst = &w.h_vel[((size_t)(a + (ta == 0 ? w.gx : 0)) * w.nb + b) * 3];
                dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
            }
            if (ta == 0) have[a + w.gx] = 1;
            any = true;
        }
        if (w.gx) {
            if (!have[0]) std::copy(w.h_vel.begin() + (size_t)1 * w.nb * 3, w.h_vel.begin() + (size_t)2 * w.nb * 3, w.h_vel.begin());
            if (!have[w.na + 1]) std::copy(w.h_vel.begin() + (size_t)w.na * w.nb * 3, w.h_vel.begin() + (size_t)(w.na + 1) * w.nb * 3, w.h_vel.begin() + (size_t)(w.na + 1) * w.nb * 3);
        }
        BIB_CHECK(cudaMemcpyAsync(w.vel, w.h_vel.data(), w.h_vel.size() * sizeof(double), cudaMemcpyHostToDevice, d->stream));
    }
    if (gx0 != 0 || gnx_wrap != 0) {        // called by the slab group with the global box: edges do not exist on multi-slab domains
        BIB_CHECK(cudaStreamSynchronize(d->stream));
        if (!any) { bib_set_error("set_boundary_velocity: no velocity boundary registered inside this box"); return 0; }
        return 1;
    }
    for (WallEdge& e : d->edges) {
        const int d1 = (e.axis + 1) % 3, d2 = (e.axis + 2) % 3;
        if (e.c1 < box[2 * d1] || e.c1 > box[2 * d1 + 1] || e.c2 < box[2 * d2] || e.c2 > box[2 * d2 + 1]) continue;
        for (int t = std::max(0, box[2 * e.axis]); t <= std::min(e.n - 1, box[2 * e.axis + 1]); ++t) {
            int c[3]; c[e.axis] = t; c[d1] = e.c1; c[d2] = e.c2;
            const double* src = vel + (((size_t)(c[0] - box[0]) * by + (c[1] - box[2])) * bz + (c[2] - box[4])) * 3;
            e.h_vel[(size_t)t * 3] = src[0]; e.h_vel[(size_t)t * 3 + 1] = src[1]; e.h_vel[(size_t)t * 3 + 2] = src[2];
        }
        BIB_CHECK(cudaMemcpyAsync(e.vel, e.h_vel.data(), e.h_vel.size() * sizeof(double), cudaMemcpyHostToDevice, d->stream));
        any = true;
    }
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    if (!any) { bib_set_error("set_boundary_velocity: no velocity boundary registered inside this box"); return 0; }
    return 1;
}
int lbm_set_wall_velocity(bib_domain* d, const int box[6], const double* vel) {
    const int xlo = d->n_ranks > 1 ? -1 : 0, xhi = d->n_ranks > 1 ? d->nx : d->nx - 1;
    if (box[0] < xlo || box[1] > xhi) { bib_set_error("set_boundary_velocity: box outside the slab"); return 0; }
    return lbm_set_wall_velocity_ext(d, box, vel, 0, 0);
}
int lbm_apply_walls(bib_domain* d) {
    for (const WallPlane& wp : d->walls) {
        WallParams w;
        w.f = d->f[d->cur]; w.force = d->force; w.wallvel = wp.vel; w.nx = d->nx; w.ny = d->ny; w.nz = d->nz;
        w.plane = (long long)d->plane; w.vol = (long long)d->vol; w.dir = wp.dir; w.orient = wp.orient; w.coord = wp.coord; w.na = wp.na; w.nb = wp.nb; w.gx = wp.gx;
        w.a0 = wp.a0; w.b0 = wp.b0; w.ca = wp.a1 - wp.a0 + 1; w.cb = wp.b1 - wp.b0 + 1;
        w.omega = d->omega0;
        k_wall_plane<<<dim3((w.cb + 127) / 128, w.ca), 128, 0, d->stream>>>(w);
        d->launches++;
    }
    for (const WallEdge& e : d->edges) {     // after the planes: order of insertion in every shipped set-up (extensional_flow_setup.h:78-88)
        const WallPlane& p1 = d->walls[e.plane1]; const WallPlane& p2 = d->walls[e.plane2];
        EdgeParams w;
        w.f = d->f[d->cur]; w.force = d->force; w.vel_e = e.vel; w.vel_p1 = p1.vel; w.vel_p2 = p2.vel;
        w.nx = d->nx; w.ny = d->ny; w.nz = d->nz; w.plane = (long long)d->plane; w.vol = (long long)d->vol;
        w.axis = e.axis; w.n1 = e.n1; w.n2 = e.n2; w.c1 = e.c1; w.c2 = e.c2; w.e0 = e.e0; w.count = e.e1 - e.e0 + 1; w.n = e.n; w.per_axis = d->per[e.axis];
        w.p1_na = p1.na; w.p1_nb = p1.nb; w.p2_na = p2.na; w.p2_nb = p2.nb;
        w.omega = d->omega0;
        k_wall_edge<<<(w.count + 127) / 128, 128, 0, d->stream>>>(w);
        d->launches++;
    }
    if (!d->walls.empty()) BIB_CHECK(cudaGetLastError());
    return 1;
}
// GuoRigidWallBoundaryFunctional::process (src/fsi/GuoBoundaryCondition.hh:133-198, depth 2): one thread per wall node. The node's
// 19 populations are rebuilt from the equilibrium at the extrapolated wall momentum plus the extrapolated non-equilibrium part
// of the two fluid nodes behind it (loc + dir, loc + 2 dir). Nodes whose own dynamics is NoDynamics / BounceBack get no
// equilibrium part (their computeEquilibrium returns 0; see oracle/biblbm_oracle.c orc_guo_apply for where that happens).
__device__ __forceinline__ double d_ma2_equilibrium(int i, double rhoBar, double invRho, const double* j, double jSqr) {
    const double t = i == 0 ? 1. / 3 : ((i >= 1 && i <= 3) || (i >= 10 && i <= 12) ? 1. / 18 : 1. / 36);
    const double c_j = c_cx[i] * j[0] + c_cy[i] * j[1] + c_cz[i] * j[2];
    return t * (rhoBar + 3. * c_j + invRho * (4.5 * c_j * c_j - 1.5 * jSqr));
}
__global__ void __launch_bounds__(128) k_guo_wall(double* __restrict__ f, const uint8_t* __restrict__ kind, const int* __restrict__ nodes,
                                                  const double* __restrict__ delta_arr, const double* __restrict__ wall_vel, int first_node, int n_nodes,
                                                  int nxl, int gnx, int ny, int nz, int x_offset, int per_x, int per_y, int per_z,
                                                  long long plane, long long vol) {
    const int w = first_node + blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= first_node + n_nodes) return;
    const int* nd = nodes + 6 * w;
    const int lx = nd[0] - x_offset;
    if (lx < 0 || lx >= nxl) return;                       // another slab owns this node
    int c1[3] = {nd[0] + nd[3], nd[1] + nd[4], nd[2] + nd[5]}, c2[3] = {nd[0] + 2 * nd[3], nd[1] + 2 * nd[4], nd[2] + 2 * nd[5]};
    if (per_x) { c1[0] = (c1[0] + gnx) % gnx; c2[0] = (c2[0] + 2 * gnx) % gnx; }
    if (per_y) { c1[1] = (c1[1] + ny) % ny; c2[1] = (c2[1] + 2 * ny) % ny; }
    if (per_z) { c1[2] = (c1[2] + nz) % nz; c2[2] = (c2[2] + 2 * nz) % nz; }
    const long long i0 = (long long)(lx + 1) * plane + (long long)nd[1] * nz + nd[2];
    const long long i1 = (long long)(c1[0] - x_offset + 1) * plane + (long long)c1[1] * nz + c1[2];
    const long long i2 = (long long)(c2[0] - x_offset + 1) * plane + (long long)c2[1] * nz + c2[2];
    double f1[BIB_Q], f2[BIB_Q];
#pragma unroll
    for (int i = 0; i < BIB_Q; ++i) { f1[i] = f[(long long)i * vol + i1]; f2[i] = f[(long long)i * vol + i2]; }
    double rhoBar1, rhoBar2, j1[3], j2[3];
    d_rhoBar_j(f1, rhoBar1, j1[0], j1[1], j1[2]);
    d_rhoBar_j(f2, rhoBar2, j2[0], j2[1], j2[2]);
    // a neighbour that is itself NoDynamics / BounceBack answers through ITS dynamics: NoDynamics: rhoBar = 0 (rho = 1), j = 0, equilibrium = 0
    // (src/core/dynamics.hh:773-777,852-857,946-950,1027-1032)
    const bool dead1 = kind && kind[i1] != 0, dead2 = kind && kind[i2] != 0;
    if (dead1) { rhoBar1 = kind[i1] == 2 ? -1.0 : 0.0; j1[0] = j1[1] = j1[2] = 0.0; }      // BounceBack(rho = T()) reports rhoBar(0) = -1
    if (dead2) { rhoBar2 = kind[i2] == 2 ? -1.0 : 0.0; j2[0] = j2[1] = j2[2] = 0.0; }
    const double j1sqr = j1[0] * j1[0] + j1[1] * j1[1] + j1[2] * j1[2];
    const double j2sqr = j2[0] * j2[0] + j2[1] * j2[1] + j2[2] * j2[2];
    const double inv1 = 1. / (rhoBar1 + 1.), inv2 = 1. / (rhoBar2 + 1.);
    const double delta = delta_arr[w];
    const double rhoBar_w = rhoBar1;
    double wall_j[3], j_w[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) wall_j[d] = (rhoBar_w + 1.) * (wall_vel ? wall_vel[3 * w + d] : 0.0);
    const bool near = delta < 0.75;
#pragma unroll
    for (int d = 0; d < 3; ++d)
        j_w[d] = near ? (wall_j[d] + (delta - 1.) * j1[d]) + ((1. - delta) / (1. + delta)) * (2. * wall_j[d] + (delta - 1.) * j2[d])
                      : (1. / delta) * (wall_j[d] + (delta - 1.) * j1[d]);
    const double jSqr_w = j_w[0] * j_w[0] + j_w[1] * j_w[1] + j_w[2] * j_w[2];
    const bool no_eq = kind && kind[i0] != 0;
#pragma unroll
    for (int i = 0; i < BIB_Q; ++i) {
        const double n1 = f1[i] - (dead1 ? 0.0 : d_ma2_equilibrium(i, rhoBar1, inv1, j1, j1sqr));
        const double n2 = f2[i] - (dead2 ? 0.0 : d_ma2_equilibrium(i, rhoBar2, inv2, j2, j2sqr));
        const double neq = near ? delta * n1 + (1. - delta) * n2 : n1;
        f[(long long)i * vol + i0] = (no_eq ? 0.0 : d_ma2_equilibrium(i, rhoBar_w, 1. / (rhoBar_w + 1.), j_w, jSqr_w)) + neq;
    }
}
int lbm_set_guo_nodes(bib_domain* d, int n, const int* loc_dir, const double* delta, const double* wall_velocity) {
    if (d->guo_nodes) { cudaFree(d->guo_nodes); cudaFree(d->guo_delta); if (d->guo_vel) cudaFree(d->guo_vel); d->guo_nodes = nullptr; d->guo_delta = nullptr; d->guo_vel = nullptr; }
    d->n_guo = 0;
    if (n <= 0) return 1;
    for (int w = 0; w < n; ++w) {
        const int* q = loc_dir + 6 * w;
        bool ok = q[0] >= 0 && q[0] < d->gnx && q[1] >= 0 && q[1] < d->ny && q[2] >= 0 && q[2] < d->nz;
        const int nn[3] = {d->gnx, d->ny, d->nz};
        for (int k = 0; ok && k < 3; ++k) {
            const int far = q[k] + 2 * q[3 + k];
            if (!d->per[k] && (far < 0 || far >= nn[k])) ok = false;
        }
        if (!ok) { bib_set_error("set_guo_wall_nodes: node or its two fluid neighbours outside the lattice"); return 0; }
        if (d->n_ranks > 1 && q[3] != 0) { bib_set_error("set_guo_wall_nodes: on multi-slab domains the node direction must not have an x component"); return 0; }
    }
    // The reference updates the nodes one after the other, in place (GuoRigidWallBoundaryFunctional::process): a node whose first or second
    // neighbour is itself a wall node (narrow gaps, staircase corners) reads that node's NEW value if it comes earlier in the list and its
    // OLD value if it comes later. One thread per node reproduces that when the list is processed in LEVELS: level(w) = 1 + the highest
    // level among the earlier nodes w depends on (read-after-write) or that depend on w's old value (write-after-read); every such
    // dependence points from a lower to a higher list index, so one forward pass assigns the levels. Typical lists have one level.
    std::vector<int> level(n, 0);
    {
        std::unordered_map<long long, int> at;      // cell -> node index
        const auto key = [&](long long x, long long y, long long z) { return (x * d->ny + y) * d->nz + z; };
        for (int w = 0; w < n; ++w) at[key(loc_dir[6 * w], loc_dir[6 * w + 1], loc_dir[6 * w + 2])] = w;
        const int nn[3] = {d->gnx, d->ny, d->nz};
        std::vector<std::vector<int>> readers(n);   // readers[v] = nodes that read cell v
        for (int w = 0; w < n; ++w)
            for (int m = 1; m <= 2; ++m) {
                long long c[3];
                for (int k = 0; k < 3; ++k) { c[k] = loc_dir[6 * w + k] + (long long)m * loc_dir[6 * w + 3 + k]; if (d->per[k]) c[k] = ((c[k] % nn[k]) + nn[k]) % nn[k]; }
                const auto it = at.find(key(c[0], c[1], c[2]));
                if (it != at.end() && it->second != w) readers[it->second].push_back(w);
            }
        for (int v = 0; v < n; ++v)
            for (int w : readers[v]) {
                if (v < w) level[w] = std::max(level[w], level[v] + 1);          // w reads v's new value: after v
            }
        // write-after-read: v (later in the list) must not overwrite its cell before the earlier reader w has run; forward pass in index order
        bool changed = true;
        for (int pass = 0; changed && pass < 64; ++pass) {
            changed = false;
            for (int v = 0; v < n; ++v)
                for (int w : readers[v]) {
                    if (w < v && level[v] < level[w] + 1) { level[v] = level[w] + 1; changed = true; }
                    if (v < w && level[w] < level[v] + 1) { level[w] = level[v] + 1; changed = true; }
                }
        }
    }
    std::vector<int> order(n);
    for (int w = 0; w < n; ++w) order[w] = w;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return level[a] < level[b]; });
    std::vector<int> ld((size_t)n * 6); std::vector<double> dl(n), wv(wall_velocity ? (size_t)n * 3 : 0);
    d->guo_levels.clear();
    for (int k = 0; k < n; ++k) {
        const int w = order[k];
        for (int c = 0; c < 6; ++c) ld[(size_t)k * 6 + c] = loc_dir[6 * w + c];
        dl[k] = delta[w];
        if (wall_velocity) for (int c = 0; c < 3; ++c) wv[(size_t)k * 3 + c] = wall_velocity[3 * w + c];
        if (k == 0 || level[w] != level[order[k - 1]]) d->guo_levels.push_back(k);
    }
    d->guo_levels.push_back(n);
    BIB_CHECK(cudaMalloc(&d->guo_nodes, (size_t)n * 6 * sizeof(int)));
    BIB_CHECK(cudaMalloc(&d->guo_delta, (size_t)n * sizeof(double)));
    BIB_CHECK(cudaMemcpy(d->guo_nodes, ld.data(), (size_t)n * 6 * sizeof(int), cudaMemcpyHostToDevice));
    BIB_CHECK(cudaMemcpy(d->guo_delta, dl.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    if (wall_velocity) {
        BIB_CHECK(cudaMalloc(&d->guo_vel, (size_t)n * 3 * sizeof(double)));
        BIB_CHECK(cudaMemcpy(d->guo_vel, wv.data(), (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice));
    }
    d->n_guo = n;
    return 1;
}
int lbm_apply_guo(bib_domain* d) {
    if (d->n_guo == 0) return 1;
    for (size_t l = 0; l + 1 < d->guo_levels.size(); ++l) {          // one launch per dependence level (normally one)
        const int first = d->guo_levels[l], cnt = d->guo_levels[l + 1] - first;
        k_guo_wall<<<(cnt + 127) / 128, 128, 0, d->stream>>>(d->f[d->cur], d->kind, d->guo_nodes, d->guo_delta, d->guo_vel, first, cnt, d->nx, d->gnx, d->ny, d->nz,
                                                          d->x_offset, d->n_ranks == 1 ? d->per[0] : 0, d->per[1], d->per[2], (long long)d->plane, (long long)d->vol);
        d->launches++;
    }
    BIB_CHECK(cudaGetLastError());
    return 1;
}
int lbm_wall_velocity_fill(bib_domain* d) {
    if (!d->vel) return 1;
    for (const WallPlane& wp : d->walls) {
        const int ca = wp.a1 - wp.a0 + 1, cb = wp.b1 - wp.b0 + 1;
        k_wall_velocity_fill<<<dim3((cb + 127) / 128, ca), 128, 0, d->stream>>>(d->vel, wp.vel, wp.dir, wp.coord, wp.na, wp.nb, wp.a0, wp.b0, cb,
                                                                              (long long)d->plane, (long long)d->vol, d->nz, wp.gx);
        d->launches++;
    }
    for (const WallEdge& e : d->edges) {
        const int count = e.e1 - e.e0 + 1;
        k_edge_velocity_fill<<<(count + 127) / 128, 128, 0, d->stream>>>(d->vel, e.vel, e.axis, e.c1, e.c2, e.e0, count, (long long)d->plane, (long long)d->vol, d->nz);
        d->launches++;
    }
    return 1;
}

int lbm_launch_upload(bib_domain* d, const int box[6], const double* cells) {
    if (!box_ok(d, box)) { bib_set_error("upload: box outside the slab"); return 0; }
    const int bnx = box[1] - box[0] + 1, bny = box[3] - box[2] + 1, bnz = box[5] - box[4] + 1;
    const long long n = (long long)bnx * bny * bnz;
    // staged in chunks of x-planes so that very large slabs do not need a second full copy in HBM
    const long long per_plane = (long long)bny * bnz;
    const int ncomp = BIB_Q + d->n_ext;
    const int chunk = (int)std::max<long long>(1, std::min<long long>(bnx, (256LL << 20) / (per_plane * ncomp * 8) + 1));
    if (!ensure_scratch(d, (size_t)chunk * per_plane * ncomp * 8 + 16)) return 0;
    int* bad = (int*)((char*)d->scratch + (size_t)chunk * per_plane * ncomp * 8);
    BIB_CHECK(cudaMemsetAsync(bad, 0, sizeof(int), d->stream));
    for (int x0 = 0; x0 < bnx; x0 += chunk) {
        const int cn = std::min(chunk, bnx - x0);
        BIB_CHECK(cudaMemcpyAsync(d->scratch, cells + (size_t)x0 * per_plane * ncomp, (size_t)cn * per_plane * ncomp * 8, cudaMemcpyHostToDevice, d->stream));
        k_unpack_cells<<<grid_for(cn * per_plane, 256), 256, 0, d->stream>>>((const double*)d->scratch, d->f[d->cur], d->force, d->phase, bad,
                                                                          box[0] + x0, box[2], box[4], cn, bny, bnz, (long long)d->plane, (long long)d->vol, d->nz, ncomp);
        d->launches++;
    }
    int h_bad = 0;
    BIB_CHECK(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, d->stream));
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    (void)n;
    if (h_bad) { bib_set_error("upload: phase scalar must be exactly 0.0 or 1.0 (it is stored as one byte)"); return 0; }
    return 1;
}

int lbm_launch_download(bib_domain* d, const int box[6], double* cells) {
    if (!box_ok(d, box)) { bib_set_error("download: box outside the slab"); return 0; }
    const int bnx = box[1] - box[0] + 1, bny = box[3] - box[2] + 1, bnz = box[5] - box[4] + 1;
    const long long per_plane = (long long)bny * bnz;
    const int ncomp = BIB_Q + d->n_ext;
    const int chunk = (int)std::max<long long>(1, std::min<long long>(bnx, (256LL << 20) / (per_plane * ncomp * 8) + 1));
    if (!ensure_scratch(d, (size_t)chunk * per_plane * ncomp * 8 + 16)) return 0;
    for (int x0 = 0; x0 < bnx; x0 += chunk) {
        const int cn = std::min(chunk, bnx - x0);
        k_pack_cells<<<grid_for(cn * per_plane, 256), 256, 0, d->stream>>>((double*)d->scratch, d->f[d->cur], d->force, d->phase,
                                                                        box[0] + x0, box[2], box[4], cn, bny, bnz, (long long)d->plane, (long long)d->vol, d->nz, ncomp);
        d->launches++;
        BIB_CHECK(cudaMemcpyAsync(cells + (size_t)x0 * per_plane * ncomp, d->scratch, (size_t)cn * per_plane * ncomp * 8, cudaMemcpyDeviceToHost, d->stream));
    }
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    return 1;
}

int lbm_launch_set_external(bib_domain* d, const int box[6], double fx, double fy, double fz) {
    if (!box_ok(d, box)) { bib_set_error("set_external_vector: box outside the slab"); return 0; }
    const int bnx = box[1] - box[0] + 1, bny = box[3] - box[2] + 1, bnz = box[5] - box[4] + 1;
    k_set_external<<<grid_for((long long)bnx * bny * bnz, 256), 256, 0, d->stream>>>(d->force, box[0], box[2], box[4], bnx, bny, bnz,
                                                                                  (long long)d->plane, (long long)d->vol, d->nz, fx, fy, fz);
    d->launches++;
    BIB_CHECK(cudaGetLastError());
    const bool full = box[0] == 0 && box[1] == d->nx - 1 && box[2] == 0 && box[3] == d->ny - 1 && box[4] == 0 && box[5] == d->nz - 1;
    if (full) d->spread_band = -1;          // nothing left to reset
    d->flags_by_masks = false;              // (a full reset clears the force flags an earlier interpolation may have raised)
    if (full && d->sparse_force) {          // every cell now holds the body force: all flags off
        const size_t words = (d->vol + 31) / 32;
        if (!d->fflag) { BIB_CHECK(cudaMalloc(&d->fflag, words * sizeof(unsigned))); d->bytes += (long long)(words * sizeof(unsigned)); }
        BIB_CHECK(cudaMemsetAsync(d->fflag, 0, words * sizeof(unsigned), d->stream));
        d->fflag_valid = true;
    } else d->fflag_valid = false;
    return 1;
}

int lbm_launch_velocity(bib_domain* d, const int box[6]) {
    if (!box_ok(d, box)) { bib_set_error("compute_velocity: box outside the slab"); return 0; }
    if (!d->vel) {
        BIB_CHECK(cudaMalloc(&d->vel, d->vol * 3 * sizeof(double)));
        BIB_CHECK(cudaMemsetAsync(d->vel, 0, d->vol * 3 * sizeof(double), d->stream));
        d->bytes += (long long)d->vol * 3 * 8;
    }
    const int bnx = box[1] - box[0] + 1, bny = box[3] - box[2] + 1, bnz = box[5] - box[4] + 1;
    k_velocity<<<grid_for((long long)bnx * bny * bnz, 256), 256, 0, d->stream>>>(d->f[d->cur], d->force, d->vel, d->kind, box[0], box[2], box[4], bnx, bny, bnz,
                                                                              (long long)d->plane, (long long)d->vol, d->nz);
    d->launches++;
    BIB_CHECK(cudaGetLastError());
    return lbm_wall_velocity_fill(d);
}

int lbm_velocity_download(bib_domain* d, const int box[6], double* velocity) {
    if (!box_ok(d, box)) { bib_set_error("velocity_download: box outside the slab"); return 0; }
    if (!d->vel) { bib_set_error("velocity_download: call bib_compute_velocity first"); return 0; }
    const int bnx = box[1] - box[0] + 1, bny = box[3] - box[2] + 1, bnz = box[5] - box[4] + 1;
    const long long n = (long long)bnx * bny * bnz;
    if (!ensure_scratch(d, (size_t)n * 3 * 8)) return 0;
    k_pack_velocity<<<grid_for(n, 256), 256, 0, d->stream>>>((double*)d->scratch, d->vel, box[0], box[2], box[4], bnx, bny, bnz,
                                                          (long long)d->plane, (long long)d->vol, d->nz);
    d->launches++;
    BIB_CHECK(cudaMemcpyAsync(velocity, d->scratch, (size_t)n * 3 * 8, cudaMemcpyDeviceToHost, d->stream));
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    return 1;
}

int lbm_velocity_extrapolate(bib_domain* d, int dir, int coord, int inward) {
    const int n[3] = {d->nx, d->ny, d->nz};
    if (dir < 0 || dir > 2 || (inward != 1 && inward != -1) || coord < 0 || coord >= n[dir] || coord + 2 * inward < 0 || coord + 2 * inward >= n[dir]) {
        bib_set_error("velocity_extrapolate_plane: bad arguments"); return 0;
    }
    if (!d->vel) { bib_set_error("velocity_extrapolate_plane: call bib_compute_velocity / bib_velocity_upload first"); return 0; }
    if (d->n_ranks > 1 && dir == 0) { bib_set_error("velocity_extrapolate_plane: planes normal to x on single-slab domains only"); return 0; }
    const long long cells = (long long)(dir == 0 ? d->ny : d->nx) * (dir == 2 ? d->ny : d->nz);
    k_extrapolate_plane<<<grid_for(cells, 256), 256, 0, d->stream>>>(d->vel, dir, coord, inward, d->nx, d->ny, d->nz, (long long)d->plane, (long long)d->vol);
    d->launches++;
    BIB_CHECK(cudaGetLastError());
    return 1;
}

// the driver's (possibly host-edited) TensorField3D becomes the field the IB interpolates from
int lbm_velocity_upload(bib_domain* d, const int box[6], const double* velocity) {
    if (!box_ok(d, box)) { bib_set_error("velocity_upload: box outside the slab"); return 0; }
    if (!d->vel) {
        BIB_CHECK(cudaMalloc(&d->vel, d->vol * 3 * sizeof(double)));
        BIB_CHECK(cudaMemsetAsync(d->vel, 0, d->vol * 3 * sizeof(double), d->stream));
        d->bytes += (long long)d->vol * 3 * 8;
    }
    const int bnx = box[1] - box[0] + 1, bny = box[3] - box[2] + 1, bnz = box[5] - box[4] + 1;
    const long long n = (long long)bnx * bny * bnz;
    if (!ensure_scratch(d, (size_t)n * 3 * 8)) return 0;
    BIB_CHECK(cudaMemcpyAsync(d->scratch, velocity, (size_t)n * 3 * 8, cudaMemcpyHostToDevice, d->stream));
    k_unpack_velocity<<<grid_for(n, 256), 256, 0, d->stream>>>((const double*)d->scratch, d->vel, box[0], box[2], box[4], bnx, bny, bnz,
                                                            (long long)d->plane, (long long)d->vol, d->nz);
    d->launches++;
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    return 1;
}

int lbm_set_celltype(bib_domain* d, const int box[6], int kind) {
    if (!box_ok(d, box)) { bib_set_error("set_celltype: box outside the slab"); return 0; }
    if (kind < 0 || kind > 2) { bib_set_error("set_celltype: unknown kind"); return 0; }
    if (!d->kind) {
        BIB_CHECK(cudaMalloc(&d->kind, d->vol));
        BIB_CHECK(cudaMemsetAsync(d->kind, 0, d->vol, d->stream));
        d->bytes += (long long)d->vol;
    }
    const int bnx = box[1] - box[0] + 1, bny = box[3] - box[2] + 1, bnz = box[5] - box[4] + 1;
    k_set_kind<<<grid_for((long long)bnx * bny * bnz, 256), 256, 0, d->stream>>>(d->kind, box[0], box[2], box[4], bnx, bny, bnz, (long long)d->plane, d->nz, (uint8_t)kind);
    d->launches++;
    BIB_CHECK(cudaGetLastError());
    return 1;
}

// bgk_ma2_equilibrium (src/latticeBoltzmann/dynamicsTemplates3D.h:675-678) for constant (rho, u), evaluated once on the host
int lbm_init_equilibrium(bib_domain* d, const int box[6], double rho, double ux, double uy, double uz) {
    if (!box_ok(d, box)) { bib_set_error("init_equilibrium: box outside the slab"); return 0; }
    static const int c[BIB_Q][3] = {{0, 0, 0}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {-1, -1, 0}, {-1, 1, 0}, {-1, 0, -1}, {-1, 0, 1}, {0, -1, -1}, {0, -1, 1},
                                    {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {1, -1, 0}, {1, 0, 1}, {1, 0, -1}, {0, 1, 1}, {0, 1, -1}};
    static const double t[BIB_Q] = {1. / 3, 1. / 18, 1. / 18, 1. / 18, 1. / 36, 1. / 36, 1. / 36, 1. / 36, 1. / 36, 1. / 36,
                                    1. / 18, 1. / 18, 1. / 18, 1. / 36, 1. / 36, 1. / 36, 1. / 36, 1. / 36, 1. / 36};
    const double j[3] = {rho * ux, rho * uy, rho * uz};
    const double jSqr = j[0] * j[0] + j[1] * j[1] + j[2] * j[2];
    const double rhoBar = rho - 1.0, invRho = 1.0 / (rhoBar + 1.0);
    Eq19 eq;
    for (int i = 0; i < BIB_Q; ++i) {
        const double c_j = c[i][0] * j[0] + c[i][1] * j[1] + c[i][2] * j[2];
        eq.v[i] = t[i] * (rhoBar + 3. * c_j + invRho * (4.5 * c_j * c_j - 1.5 * jSqr));
    }
    const int bnx = box[1] - box[0] + 1, bny = box[3] - box[2] + 1, bnz = box[5] - box[4] + 1;
    k_fill_populations<<<grid_for((long long)bnx * bny * bnz, 256), 256, 0, d->stream>>>(d->f[d->cur], eq, box[0], box[2], box[4], bnx, bny, bnz,
                                                                                       (long long)d->plane, (long long)d->vol, d->nz);
    d->launches++;
    BIB_CHECK(cudaGetLastError());
    return 1;
}
