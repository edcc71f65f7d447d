// ib_kernels.cu -- immersed-boundary coupling and membrane forces on SoA vertex arrays.
//
// Replaces (reference file:line):
//   RomaDirac<T,3>::eval, get_dirac_compact_support_box      src/fsi/Dirac.h:73-83,131-139
//   interpolate_bulk / interpolate (BulkInterpolator, GeneralInterpolator)
//                                                            src/fsi/ImmersedBoundaryDynamics_interpolation.hh:43-153,236-276
//   spread_forces_bulk / spread_forces                       src/fsi/ImmersedBoundaryDynamics_forces.hh:52-163,243-284
//   RBCParticle::compute_forces, in_plane_ks_kp              src/fsi/RBCParticle.hh:27-44,140-247
//   tri::triangle_area_and_gradient, grad_signed_volume, grad_angle_between_pair
//                                                            src/fsi/TriangleUtils.h:57-164
//   ParticleShape (connectivity)                             src/fsi/ParticleShape.hh:22-184
//   DeformableParticle3D::move_vertices / update, map_center_of_mass_to_periodic_grid
//                                                            src/fsi/DeformableParticle3D.hh:26-82, src/fsi/ParticleBase.hh:52-69,113-138
//   RayTracer::find_intersections_z, ZBuffer, apply_voxelization
//                                                            src/fsi/RayTracer.hh:19-108, src/fsi/ZBuffer.h:39-79,
//                                                            src/fsi/ImmersedBoundaryDynamics_voxelization.hh:17-29,121-165
//   ParticlePositionRegulator                                progs/common/particle_position_regulator.h:5-18
//
// All kernels are gathers over vertex/triangle arrays except the spread, which scatters with fp64 atomics (or, in
// deterministic mode, in the reference's exact vertex order). Arithmetic mirrors the reference operation by operation
// (library built with -fmad=false), so everything except the atomic spread is bit-identical to the reference.
#include "domain.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace {

// ------------------------------------------------------------------------------------------------ geometry helpers
struct Geo {
    int nx, ny, nz;          // GLOBAL lattice size (the reference's single block)
    int per[3];
    long long plane, vol;
    int x_offset;            // global x of storage plane 1
    int nxl;                 // planes owned by this slab (== nx on a single slab)
    const int* role;         // per particle slot: 0 dead, 1 owned, 2/3 ghost of the left/right neighbour; NULL = all owned
    const int* vslot;        // slot of each vertex (used when role != NULL)
    const uint8_t* kind;     // cell kinds (NULL = all bulk)
    unsigned* fflag;         // force-differs-from-body flags (1 bit per cell), set by every atomic spread write; NULL = not maintained
    int plates;              // fsi::Boundary kind: 0 none, 1 ParallelPlatesBoundary (walls at y = py0, py1; src/fsi/Boundary.h:300-365),
    double py0, py1, prad;   //   2 PipeBoundary along x with axis (y, z) = (py0, py1) and radius prad (src/fsi/Boundary.h:131-253)
};
// slab-local plane index of global x, or -1 when another slab owns it
__device__ __forceinline__ bool d_in_slab(const Geo& g, int gx) { const int lx = gx - g.x_offset; return lx >= 0 && lx < g.nxl; }

__device__ __forceinline__ double d_roma(double x) {
    const double dx = fabs(x);
    return (1.0 / 6.0) * (dx > 0.5 ? (5.0 - 3 * dx - sqrt(-3 * (1 - dx) * (1 - dx) + 1)) : 2.0 * (1 + sqrt(1 - 3 * dx * dx)));
}
__device__ __forceinline__ double d_min_mod(double x, double N, double N_inv) { return x - floor(x * N_inv + 0.5) * N; }

// Rect::contains(p, arithmetic), src/fsi/geometry.h:95-100, with Arithmetic3D::dist_* (src/fsi/Periodicity.h:24-32,99-102)
__device__ __forceinline__ bool d_rect_contains(const Geo& g, double lo_enl, double hi_enl, const double* p) {
    const int n[3] = {g.nx, g.ny, g.nz};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double lo = 0.0, hi = (double)(n[d] - 1);
        lo -= lo_enl; hi += hi_enl;
        const double mid = 0.5 * (lo + hi);
        const double L = (double)n[d];
        const double dist = g.per[d] ? d_min_mod(mid - p[d], L, 1. / L) : mid - p[d];
        if (!(fabs(dist) < 0.5 * (hi - lo))) return false;
    }
    return true;
}
// Boundary::distance_to_boundary_less_than / distance_to_boundary / get_normal / trace_ray of the two boundary classes
__device__ __forceinline__ double d_sqr(double x) { return x * x; }
__device__ __forceinline__ bool d_boundary_closer_than(const Geo& g, const double* p, double dist) {
    if (g.plates == 2) return (d_sqr(p[1] - g.py0) + d_sqr(p[2] - g.py1)) > d_sqr(g.prad - dist);
    return (p[1] + dist) > g.py1 || (p[1] - dist) < g.py0;
}
__device__ __forceinline__ double d_plates_distance(const Geo& g, const double* r) {
    if (g.plates == 2) return g.prad - sqrt(d_sqr(r[1] - g.py0) + d_sqr(r[2] - g.py1));
    return (r[1] < 0.5 * (g.py0 + g.py1)) ? r[1] - g.py0 : g.py1 - r[1];
}
__device__ __forceinline__ void d_boundary_normal(const Geo& g, const double* r, double* n) {
    if (g.plates == 2) {
        const double dy = r[1] - g.py0, dz = r[2] - g.py1;
        const double nrm = sqrt((0.0 * 0.0 + dy * dy) + dz * dz);
        n[0] = 0.0 / nrm; n[1] = dy / nrm; n[2] = dz / nrm;
    } else { n[0] = 0; n[1] = (r[1] < 0.5 * (g.py0 + g.py1)) ? -1.0 : 1.0; n[2] = 0; }
}
__device__ __forceinline__ bool d_boundary_trace_ray(const Geo& g, const double* x0, const double* x1, double& t) {
    if (g.plates == 2) {
        const double y1 = x0[1] - g.py0, y2 = x1[1] - g.py0, z1 = x0[2] - g.py1, z2 = x1[2] - g.py1;
        const double a = (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
        const double b = y1 * (y2 - y1) + z1 * (z2 - z1);
        const double c = y1 * y1 + z1 * z1 - g.prad * g.prad;
        const double disc = b * b - c * a;
        if (disc < 0. || fabs(a) < 10 * 2.220446049250313e-16) return false;
        const double bOverA = -b / a;
        const double sqrtDisc = sqrt(disc) / a;
        t = bOverA - sqrtDisc;
        if (t >= 0. && t <= 1.) return true;
        t = bOverA + sqrtDisc;
        return t >= 0. && t <= 1.;
    }
    t = (g.py1 - x0[1]) / (x1[1] - x0[1]);
    if (t > 0 && t < 1) return true;
    t = (g.py0 - x0[1]) / (x1[1] - x0[1]);
    return t > 0 && t < 1;
}
// 0 = outside domain_with_fsi_envelope, 1 = bulk list, 2 = envelope list (src/fsi/ImmersedBoundaryDynamics.hh:46-65,346-358)
__device__ __forceinline__ int d_classify(const Geo& g, const double* p) {
    const double eps = 2.220446049250313e-16;
    if (!d_rect_contains(g, 1.5, 1.5, p)) return 0;
    if (g.plates && d_boundary_closer_than(g, p, 1.5 + 1.)) return 3;   // near-boundary list (ImmersedBoundaryDynamics.hh:349-350)
    const double d_bulk = 0.5 - 1.5 - eps;
    if (d_rect_contains(g, d_bulk, d_bulk, p)) return 1;
    return 2;
}
// local stencil index (block frame, envelope width 1) -> global lattice index
__device__ __forceinline__ int d_remap(const Geo& g, int d, int i_local) {
    const int n = d == 0 ? g.nx : (d == 1 ? g.ny : g.nz);
    const int gi = i_local - 1;
    if (!g.per[d]) return gi;
    return (gi + 100 * n) % n;
}
__device__ __forceinline__ long long d_cell(const Geo& g, int gx, int gy, int gz) {
    return (long long)(gx - g.x_offset + 1) * g.plane + (long long)gy * g.nz + gz;
}

// u = j/rho + F/2 from the current populations (ExternalForceDynamics::computeVelocity, externalForceDynamics.hh:59-74)
__device__ __forceinline__ void d_cell_velocity(const double* __restrict__ f, const double* __restrict__ force, long long vol,
                                                long long idx, double* u) {
    double c[BIB_Q];
#pragma unroll
    for (int i = 0; i < BIB_Q; ++i) c[i] = __ldg(f + (long long)i * vol + idx);
    double rhoBar, jx, jy, jz;
    d_rhoBar_j(c, rhoBar, jx, jy, jz);
    const double invRho = 1.0 / (rhoBar + 1.0);
    u[0] = jx * invRho + __ldg(force + idx) / 2.0;
    u[1] = jy * invRho + __ldg(force + vol + idx) / 2.0;
    u[2] = jz * invRho + __ldg(force + 2 * vol + idx) / 2.0;
}

// ------------------------------------------------------------------------------------------------ I2: interpolation
__device__ __forceinline__ double d_tophat(double x) { return 1. - fabs(x); }        // TopHatDirac, src/fsi/Dirac.h:55-58
// Where the interpolation reads cell velocities from. Cells are named twice: (i, j, k) = unwrapped stencil indices in the block frame
// (envelope width 1, as the reference computes them) and c0 = storage index of the (wrapped) lattice cell.
struct SrcDense {                 // the dense velocity field (bib_compute_velocity / the caller's own field)
    const double* __restrict__ vel; long long vol;
    __device__ __forceinline__ void load(int, int, int, long long c0, double* u) const { u[0] = vel[c0]; u[1] = vel[vol + c0]; u[2] = vel[2 * vol + c0]; }
};
struct WallsDev { int n; int dir[6], coord[6], nb[6], gx[6], a0[6], a1[6], b0[6], b1[6]; const double* vel[6]; };
// stored boundary velocity of a wall-plane cell (slab-local x, global y / z); false if the cell lies on no wall plane
__device__ __forceinline__ bool d_wall_velocity(const WallsDev& W, int lx, int y, int z, double* u) {
    const int c[3] = {lx, y, z};
    for (int q = 0; q < W.n; ++q) {
        const int dir = W.dir[q], ta = dir == 0 ? 1 : 0, tb = dir == 2 ? 1 : 2;
        if (c[dir] == W.coord[q] && c[ta] >= W.a0[q] && c[ta] <= W.a1[q] && c[tb] >= W.b0[q] && c[tb] <= W.b1[q]) {
            const double* wv = W.vel[q] + ((long long)(c[ta] + W.gx[q]) * W.nb[q] + c[tb]) * 3;
            u[0] = wv[0]; u[1] = wv[1]; u[2] = wv[2];
            return true;
        }
    }
    return false;
}
struct SrcPopW {                  // from the populations, honouring wall planes (stored velocity) and cell kinds (zero): the few-particle path
    const double* __restrict__ f; const double* __restrict__ force; const uint8_t* __restrict__ kind; long long vol, plane; int nz; WallsDev W;
    __device__ __forceinline__ void load(int, int, int, long long c0, double* u) const {
        if (W.n > 0) {
            const int lx = (int)(c0 / plane) - 1; const long long r = c0 % plane;
            if (d_wall_velocity(W, lx, (int)(r / nz), (int)(r % nz), u)) return;
        }
        if (kind && kind[c0] != 0) { u[0] = u[1] = u[2] = 0.0; return; }
        d_cell_velocity(f, force, vol, c0, u);
    }
};
struct SrcPop {                   // straight from the populations: u = j/rho + F/2
    const double* __restrict__ f; const double* __restrict__ force; long long vol;
    __device__ __forceinline__ void load(int, int, int, long long c0, double* u) const { d_cell_velocity(f, force, vol, c0, u); }
};
#define BAND_WORDS 512
struct SrcTile {                  // the particle's shared-memory tile of band velocities (k_interp_tile): compact slots of the touched cells
    const unsigned* mask; const int* pre; const double* u; int slots; int lo[3], n[3];
    __device__ __forceinline__ void load(int i, int j, int k, long long, double* o) const {
        const int t = ((i - 1 - lo[0]) * n[1] + (j - 1 - lo[1])) * n[2] + (k - 1 - lo[2]);
        const int s = pre[t >> 5] + __popc(mask[t >> 5] & ((1u << (t & 31)) - 1u));
        o[0] = u[s]; o[1] = u[slots + s]; o[2] = u[2 * slots + s];
    }
};

// GeneralInterpolator::interpolate_node for the Roma (3-point) or top-hat (2-point) kernel
template <class SRC>
__device__ __forceinline__ void d_general_interp(const Geo& g, const SRC& src, const double* p, bool th, double* o) {
    const int sup = th ? 2 : 3;
    const double hs = th ? 1.0 : 1.5;
    const int x0 = (int)ceil(p[0] - hs), y0 = (int)ceil(p[1] - hs), z0 = (int)ceil(p[2] - hs);
    for (int i = x0; i <= x0 + sup - 1; ++i) {
        const int i2 = d_remap(g, 0, i);
        if (i2 < 0 || i2 > g.nx - 1 || !d_in_slab(g, i2)) continue;
        const double dirac_x = th ? d_tophat((double)i - p[0]) : d_roma((double)i - p[0]);
        for (int j = y0; j <= y0 + sup - 1; ++j) {
            const int j2 = d_remap(g, 1, j);
            if (j2 < 0 || j2 > g.ny - 1) continue;
            const double dirac_xy = dirac_x * (th ? d_tophat((double)j - p[1]) : d_roma((double)j - p[1]));
            for (int k = z0; k <= z0 + sup - 1; ++k) {
                const int k2 = d_remap(g, 2, k);
                if (k2 < 0 || k2 > g.nz - 1) continue;
                const double dirac_xyz = dirac_xy * (th ? d_tophat((double)k - p[2]) : d_roma((double)k - p[2]));
                double u[3];
                src.load(i, j, k, d_cell(g, i2, j2, k2), u);
#pragma unroll
                for (int c = 0; c < 3; ++c) o[c] += dirac_xyz * u[c];
            }
        }
    }
}
// interpolate_velocity for one vertex (bulk / envelope / near-boundary lists of the reference: _interpolation.hh:43-78,112-153,181-230)
template <class SRC, bool ROLLED = false>
__device__ __forceinline__ void d_interp_vertex(const Geo& g, const SRC& src, const double* pp, double* o) {
    o[0] = o[1] = o[2] = 0.0;
    int cls = d_classify(g, pp);
    if (cls == 0) return;
    const double p[3] = {pp[0] - (-1.0), pp[1] - (-1.0), pp[2] - (-1.0)};   // pos - getLocation()
    const int x0 = (int)ceil(p[0] - 1.5), y0 = (int)ceil(p[1] - 1.5), z0 = (int)ceil(p[2] - 1.5);
    // a stencil that reaches into a neighbouring slab is summed cell by cell; the neighbour adds its part (multi-slab only)
    if (cls == 1 && !(d_in_slab(g, d_remap(g, 0, x0)) && d_in_slab(g, d_remap(g, 0, x0) + 2))) cls = 2;
    if (cls == 3) {
        // interpolate_near_boundary (_interpolation.hh:181-230): blend to the 2-point kernel within 1.5..2.5 of the wall
        const double weight = d_plates_distance(g, pp) - 1.5;
        if (weight >= 1) d_general_interp(g, src, p, false, o);
        else if (weight <= 0) d_general_interp(g, src, p, true, o);
        else {
            double sup[3] = {0, 0, 0}, th[3] = {0, 0, 0};
            d_general_interp(g, src, p, false, sup);
            d_general_interp(g, src, p, true, th);
#pragma unroll
            for (int c = 0; c < 3; ++c) o[c] = (1. - weight) * th[c] + weight * sup[c];
        }
    } else if (cls == 1) {
        const double dx[3] = {d_roma((double)x0 - p[0]), d_roma((double)x0 + 1. - p[0]), d_roma((double)(x0 + 2) - p[0])};
        const double dy[3] = {d_roma((double)y0 - p[1]), d_roma((double)y0 + 1. - p[1]), d_roma((double)(y0 + 2) - p[1])};
        const double dz[3] = {d_roma((double)z0 - p[2]), d_roma((double)z0 + 1. - p[2]), d_roma((double)(z0 + 2) - p[2])};
        const int i0 = d_remap(g, 0, x0), j0 = d_remap(g, 1, y0), k0 = d_remap(g, 2, z0);
        double ax[3][3];
#pragma unroll(ROLLED ? 1 : 3)
        for (int i = 0; i < 3; ++i) {
            double ay[3][3];
#pragma unroll(ROLLED ? 1 : 3)
            for (int j = 0; j < 3; ++j) {
                double u0[3], u1[3], u2[3];
                const long long c0 = d_cell(g, i0 + i, j0 + j, k0);
                src.load(x0 + i, y0 + j, z0, c0, u0);
                src.load(x0 + i, y0 + j, z0 + 1, c0 + 1, u1);
                src.load(x0 + i, y0 + j, z0 + 2, c0 + 2, u2);
#pragma unroll
                for (int c = 0; c < 3; ++c) ay[j][c] = dy[j] * ((dz[0] * u0[c] + dz[1] * u1[c]) + dz[2] * u2[c]);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) ax[i][c] = dx[i] * ((ay[0][c] + ay[1][c]) + ay[2][c]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c] = (ax[0][c] + ax[1][c]) + ax[2][c];
    } else d_general_interp(g, src, p, false, o);
}
template <bool DENSE>
__global__ void __launch_bounds__(128) k_interpolate(Geo g, const double* __restrict__ f, const double* __restrict__ force,
                                                     const double* __restrict__ vel, long long nvert, long long cap,
                                                     const double* __restrict__ vpos, double* __restrict__ vvel) {
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= nvert) return;
    if (g.role && g.role[g.vslot[v]] == 0) return;
    const double pp[3] = {vpos[v], vpos[cap + v], vpos[2 * cap + v]};
    double o[3];
    if (DENSE) { const SrcDense src = {vel, g.vol}; d_interp_vertex(g, src, pp, o); }
    else { const SrcPop src = {f, force, g.vol}; d_interp_vertex(g, src, pp, o); }
    vvel[v] = o[0]; vvel[cap + v] = o[1]; vvel[2 * cap + v] = o[2];
}

// Few particles: one WARP per vertex. Lane l < 27 evaluates the velocity of stencil cell l = (i*3 + j)*3 + k straight from the populations
// (22 independent loads per lane, all in flight at once); the interpolation sum is then formed by every lane in lockstep, in the
// reference's order, fetching cell values with warp shuffles (same bits as the tile / dense paths).
struct SrcWarp {
    double u[3]; int x0, y0, z0;
    __device__ __forceinline__ void load(int i, int j, int k, long long, double* o) const {
        const int m = ((i - x0) * 3 + (j - y0)) * 3 + (k - z0);
        o[0] = __shfl_sync(0xffffffffu, u[0], m); o[1] = __shfl_sync(0xffffffffu, u[1], m); o[2] = __shfl_sync(0xffffffffu, u[2], m);
    }
};
__global__ void __launch_bounds__(128) k_interpolate_warp(Geo g, WallsDev W, const double* __restrict__ f, const double* __restrict__ force,
                                                          long long nvert, long long cap, const double* __restrict__ vpos, double* __restrict__ vvel) {
    const long long v = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= nvert) return;                                   // whole warps leave together
    const double pp[3] = {vpos[v], vpos[cap + v], vpos[2 * cap + v]};
    SrcWarp src;
    src.x0 = (int)ceil((pp[0] - (-1.0)) - 1.5); src.y0 = (int)ceil((pp[1] - (-1.0)) - 1.5); src.z0 = (int)ceil((pp[2] - (-1.0)) - 1.5);
    src.u[0] = src.u[1] = src.u[2] = 0.0;
    if (lane < 27) {
        const int i2 = d_remap(g, 0, src.x0 + lane / 9), j2 = d_remap(g, 1, src.y0 + (lane / 3) % 3), k2 = d_remap(g, 2, src.z0 + lane % 3);
        if (i2 >= 0 && i2 <= g.nx - 1 && d_in_slab(g, i2) && j2 >= 0 && j2 <= g.ny - 1 && k2 >= 0 && k2 <= g.nz - 1) {
            const SrcPopW cell = {f, force, g.kind, g.vol, g.plane, g.nz, W};
            cell.load(0, 0, 0, d_cell(g, i2, j2, k2), src.u);
        }
    }
    double o[3];
    d_interp_vertex<SrcWarp, true>(g, src, pp, o);
    if (lane == 0) { vvel[v] = o[0]; vvel[cap + v] = o[1]; vvel[2 * cap + v] = o[2]; }
}

// ------------------------------------------------------------------------------------------------ I3: spreading
__device__ __forceinline__ void d_general_spread(const Geo& g, double* force, const double* p, bool th, const double* F, bool atomic) {
    const int sup = th ? 2 : 3;
    const double hs = th ? 1.0 : 1.5;
    const int x0 = (int)ceil(p[0] - hs), y0 = (int)ceil(p[1] - hs), z0 = (int)ceil(p[2] - hs);
    for (int i = x0; i <= x0 + sup - 1; ++i) {
        const int i2 = d_remap(g, 0, i);
        if (i2 < 0 || i2 > g.nx - 1 || !d_in_slab(g, i2)) continue;
        const double dirac_x = th ? d_tophat((double)i - p[0]) : d_roma((double)i - p[0]);
        for (int j = y0; j <= y0 + sup - 1; ++j) {
            const int j2 = d_remap(g, 1, j);
            if (j2 < 0 || j2 > g.ny - 1) continue;
            const double dirac_xy = dirac_x * (th ? d_tophat((double)j - p[1]) : d_roma((double)j - p[1]));
            for (int k = z0; k <= z0 + sup - 1; ++k) {
                const int k2 = d_remap(g, 2, k);
                if (k2 < 0 || k2 > g.nz - 1) continue;
                const double dirac_xyz = dirac_xy * (th ? d_tophat((double)k - p[2]) : d_roma((double)k - p[2]));
                const long long c0 = d_cell(g, i2, j2, k2);
                for (int c = 0; c < 3; ++c) {
                    const double val = dirac_xyz * F[c];
                    if (atomic) atomicAdd(force + c * g.vol + c0, val); else force[c * g.vol + c0] += val;
                }
                if (atomic && g.fflag) atomicOr(&g.fflag[c0 >> 5], 1u << (c0 & 31));
            }
        }
    }
}
__device__ __forceinline__ void d_spread_vertex(const Geo& g, double* force, const double* pp, const double* F, int cls, bool atomic) {
    // (multi-slab: cells owned by a neighbouring slab are skipped here and spread there from the ghost copy of the vertex)
    const double p[3] = {pp[0] - (-1.0), pp[1] - (-1.0), pp[2] - (-1.0)};
    const int x0 = (int)ceil(p[0] - 1.5), y0 = (int)ceil(p[1] - 1.5), z0 = (int)ceil(p[2] - 1.5);
    if (cls == 1 && !(d_in_slab(g, d_remap(g, 0, x0)) && d_in_slab(g, d_remap(g, 0, x0) + 2))) cls = 2;
    if (cls == 3) {
        // spread_forces_near_boundary (_forces.hh:192-237)
        const double weight = d_plates_distance(g, pp) - 1.5;
        if (weight >= 1) d_general_spread(g, force, p, false, F, atomic);
        else if (weight <= 0) d_general_spread(g, force, p, true, F, atomic);
        else {
            const double Fw[3] = {F[0] * weight, F[1] * weight, F[2] * weight};
            const double Ft[3] = {F[0] * (1. - weight), F[1] * (1. - weight), F[2] * (1. - weight)};
            d_general_spread(g, force, p, false, Fw, atomic);
            d_general_spread(g, force, p, true, Ft, atomic);
        }
        return;
    }
    if (cls == 1) {
        const double dx[3] = {d_roma((double)x0 - p[0]), d_roma((double)x0 + 1. - p[0]), d_roma((double)(x0 + 2) - p[0])};
        const double dy[3] = {d_roma((double)y0 - p[1]), d_roma((double)y0 + 1. - p[1]), d_roma((double)(y0 + 2) - p[1])};
        const double dz[3] = {d_roma((double)z0 - p[2]), d_roma((double)z0 + 1. - p[2]), d_roma((double)(z0 + 2) - p[2])};
        const int i0 = d_remap(g, 0, x0), j0 = d_remap(g, 1, y0), k0 = d_remap(g, 2, z0);
        for (int x = 0; x < 3; ++x) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) {
            const long long c0 = d_cell(g, i0 + x, j0 + j, k0 + k);
            const double w = dx[x] * dy[j] * dz[k];
            for (int c = 0; c < 3; ++c) {
                const double val = w * F[c];
                if (atomic) atomicAdd(force + c * g.vol + c0, val); else force[c * g.vol + c0] += val;
            }
            if (atomic && g.fflag) atomicOr(&g.fflag[c0 >> 5], 1u << (c0 & 31));
        }
    } else {
        for (int i = x0; i <= x0 + 2; ++i) {
            const int i2 = d_remap(g, 0, i);
            if (i2 < 0 || i2 > g.nx - 1 || !d_in_slab(g, i2)) continue;
            const double dirac_x = d_roma((double)i - p[0]);
            for (int j = y0; j <= y0 + 2; ++j) {
                const int j2 = d_remap(g, 1, j);
                if (j2 < 0 || j2 > g.ny - 1) continue;
                const double dirac_xy = dirac_x * d_roma((double)j - p[1]);
                for (int k = z0; k <= z0 + 2; ++k) {
                    const int k2 = d_remap(g, 2, k);
                    if (k2 < 0 || k2 > g.nz - 1) continue;
                    const double dirac_xyz = dirac_xy * d_roma((double)k - p[2]);
                    const long long c0 = d_cell(g, i2, j2, k2);
                    for (int c = 0; c < 3; ++c) {
                        const double val = dirac_xyz * F[c];
                        if (atomic) atomicAdd(force + c * g.vol + c0, val); else force[c * g.vol + c0] += val;
                    }
                    if (atomic && g.fflag) atomicOr(&g.fflag[c0 >> 5], 1u << (c0 & 31));
                }
            }
        }
    }
}

__global__ void __launch_bounds__(128) k_spread_atomic(Geo g, double* force, long long nvert, long long cap,
                                                       const double* __restrict__ vpos, const double* __restrict__ vfrc) {
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= nvert) return;
    if (g.role && g.role[g.vslot[v]] == 0) return;
    const double pp[3] = {vpos[v], vpos[cap + v], vpos[2 * cap + v]};
    const double F[3] = {vfrc[v], vfrc[cap + v], vfrc[2 * cap + v]};
    const int cls = d_classify(g, pp);
    if (cls != 0) d_spread_vertex(g, force, pp, F, cls, true);
}
// Deterministic mode: the reference's own order (all bulk-list vertices, then all envelope-list vertices), one thread.
__global__ void k_spread_sequential(Geo g, double* force, long long nvert, long long cap, const double* vpos, const double* vfrc) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int pass = 1; pass <= 3; ++pass)
        for (long long v = 0; v < nvert; ++v) {
            const double pp[3] = {vpos[v], vpos[cap + v], vpos[2 * cap + v]};
            const double F[3] = {vfrc[v], vfrc[cap + v], vfrc[2 * cap + v]};
            if (g.role && g.role[g.vslot[v]] == 0) continue;
            if (d_classify(g, pp) == pass) d_spread_vertex(g, force, pp, F, pass, false);
        }
}

// ------------------------------------------------------------------------------------------------ M1: RBC forces
struct V3 { double x, y, z; };
__device__ __forceinline__ V3 ld3(const double* a, long long cap, long long i) { return {a[i], a[cap + i], a[2 * cap + i]}; }
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 crs(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double nrm(V3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
__device__ __forceinline__ V3 divs(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }

// tri::triangle_area_and_gradient (TriangleUtils.h:57-103); returns the gradient w.r.t. corner `k`
__device__ __forceinline__ V3 d_area_grad(V3 v0, V3 v1, V3 v2, int k, double& area) {
    const double v01v11 = v0.y - v1.y, v02v22 = v0.z - v2.z, v02v12 = v0.z - v1.z;
    const double v01v21 = v0.y - v2.y, v00v20 = v0.x - v2.x, v00v10 = v0.x - v1.x;
    const double a = v01v11 * v02v22 - v02v12 * v01v21;
    const double b = v02v12 * v00v20 - (v0.x - v1.x) * v02v22;
    const double c = v00v10 * v01v21 - v01v11 * v00v20;
    area = 0.5 * sqrt(a * a + b * b + c * c);
    const double denom = 1. / (4. * area);
    V3 g;
    if (k == 0) {
        g.x = (b * (v2.z - v1.z) + c * (v1.y - v2.y)) * denom;
        g.y = (a * (v1.z - v2.z) + c * (v2.x - v1.x)) * denom;
        g.z = (a * (v2.y - v1.y) + b * (v1.x - v2.x)) * denom;
    } else if (k == 1) {
        g.x = (b * v02v22 + c * (v2.y - v0.y)) * denom;
        g.y = (a * (v2.z - v0.z) + c * v00v20) * denom;
        g.z = (a * v01v21 + b * (v2.x - v0.x)) * denom;
    } else {
        g.x = (b * (v1.z - v0.z) + c * v01v11) * denom;
        g.y = (a * (v02v12) + c * (v1.x - v0.x)) * denom;
        g.z = (a * (v1.y - v0.y) + b * v00v10) * denom;
    }
    return g;
}

// tri::grad_angle_between_pair (TriangleUtils.h:138-164); gradient w.r.t. corner k
__device__ __forceinline__ V3 d_angle_grad(V3 v0, V3 v1, V3 v2, V3 v3, int k, double& cs, double& sn) {
    const V3 e0 = sub(v1, v0), e1 = sub(v2, v0), e2 = sub(v3, v0);
    const double e1_norm = nrm(e1);
    const V3 e1u = divs(e1, e1_norm);
    const V3 n1 = crs(e0, e1), n2 = crs(e1, e2);
    const double n1_norm = nrm(n1), n2_norm = nrm(n2);
    const V3 n1u = divs(n1, n1_norm), n2u = divs(n2, n2_norm);
    cs = dot(n1u, n2u);
    sn = ((dot(sub(n1, n2), sub(v1, v3)) > 0) ? 1. : -1.) * nrm(crs(n1u, n2u));
    V3 g;
    if (k == 0) {
        const double a = dot(e1u, sub(e1, e2)) / n2_norm, b = dot(e1u, sub(e1, e0)) / n1_norm;
        g = {a * n2u.x + b * n1u.x, a * n2u.y + b * n1u.y, a * n2u.z + b * n1u.z};
    } else if (k == 1) {
        const double a = -e1_norm / n1_norm;
        g = {a * n1u.x, a * n1u.y, a * n1u.z};
    } else if (k == 2) {
        const double a = dot(e1u, e2) / n2_norm, b = dot(e1u, e0) / n1_norm;
        g = {a * n2u.x + b * n1u.x, a * n2u.y + b * n1u.y, a * n2u.z + b * n1u.z};
    } else {
        const double a = -e1_norm / n2_norm;
        g = {a * n2u.x, a * n2u.y, a * n2u.z};
    }
    return g;
}

struct RbcParams {
    int nv;
    long long first_vertex, cap;
    int first_particle, count;
    const int* tris; const double* tri_area;
    const int* links; const double* link_len; const double* link_ks; const double* link_kp;
    const int* pairs; const double* pair_cs;
    const int* v_link_off; const int* v_link;
    const int* v_tri_off; const int* v_tri;
    const int* v_pair_off; const int* v_pair;
    double shape_area, k_ag, k_al, k_v, vol_desired, k_b;
    double reg_p, reg_t[3];
    const int* role;
    const double* f_init;     // non-NULL: vertex forces start from the values a decorator pass left in vfrc (wall interaction)
};

// Force on vertex v of one particle: gathers the contributions of its incident links, triangles and pairs in the order in
// which RBCParticle::compute_forces adds them to that vertex (decorator, links, triangles [area, volume], pairs).
// P/capP/baseP address the particle's vertex positions (global SoA arrays or a shared-memory copy).
__device__ __forceinline__ V3 d_rbc_vertex_force(const RbcParams& q, const double* P, long long capP, long long baseP, int v, const double* st,
                                                  long long gv) {
    const double area = st[0], volume = st[1];
    double fx = 0.0, fy = 0.0, fz = 0.0;
    if (q.f_init) { fx = q.f_init[gv]; fy = q.f_init[q.cap + gv]; fz = q.f_init[2 * q.cap + gv]; }
    if (q.reg_p != 0.0) {
        fx += (q.reg_p * (q.reg_t[0] - st[2])) / (double)q.nv;
        fy += (q.reg_p * (q.reg_t[1] - st[3])) / (double)q.nv;
        fz += (q.reg_p * (q.reg_t[2] - st[4])) / (double)q.nv;
    }
    for (int e = q.v_link_off[v]; e < q.v_link_off[v + 1]; ++e) {
        const int code = q.v_link[e], l = code >> 1, end = code & 1;
        const V3 a = ld3(P, capP, baseP + q.links[2 * l]), b = ld3(P, capP, baseP + q.links[2 * l + 1]);
        const V3 dv = sub(a, b);
        const double L = nrm(dv);
        const double x = L / (q.link_len[l] * 2.2);
        const double df = -0.25 * q.link_ks[l] * (6. * x + (3. * x * x - 2. * x * x * x) / ((1 - x) * (1 - x))) + q.link_kp[l] / (L * L);
        const double tx = df * dv.x / L, ty = df * dv.y / L, tz = df * dv.z / L;
        if (end == 0) { fx += tx; fy += ty; fz += tz; } else { fx -= tx; fy -= ty; fz -= tz; }
    }
    const double df_vol = q.k_v * (q.vol_desired - volume) / q.vol_desired;
    for (int e = q.v_tri_off[v]; e < q.v_tri_off[v + 1]; ++e) {
        const int code = q.v_tri[e], tr = code / 3, k = code % 3;
        const V3 v0 = ld3(P, capP, baseP + q.tris[3 * tr]), v1 = ld3(P, capP, baseP + q.tris[3 * tr + 1]), v2 = ld3(P, capP, baseP + q.tris[3 * tr + 2]);
        double local_area;
        const V3 g = d_area_grad(v0, v1, v2, k, local_area);
        double df_area = 0.;
        df_area += q.k_ag * (q.shape_area - area) / q.shape_area;
        df_area += q.k_al * (q.tri_area[tr] - local_area) / q.tri_area[tr];
        fx += df_area * g.x; fy += df_area * g.y; fz += df_area * g.z;
        const V3 gv = k == 0 ? crs(v1, v2) : (k == 1 ? crs(v2, v0) : crs(v0, v1));   // grad_signed_volume, TriangleUtils.h:112-120
        fx += df_vol * (gv.x / 6.); fy += df_vol * (gv.y / 6.); fz += df_vol * (gv.z / 6.);
    }
    for (int e = q.v_pair_off[v]; e < q.v_pair_off[v + 1]; ++e) {
        const int code = q.v_pair[e], pr = code >> 2, k = code & 3;
        const V3 v0 = ld3(P, capP, baseP + q.pairs[4 * pr]), v1 = ld3(P, capP, baseP + q.pairs[4 * pr + 1]);
        const V3 v2 = ld3(P, capP, baseP + q.pairs[4 * pr + 2]), v3 = ld3(P, capP, baseP + q.pairs[4 * pr + 3]);
        double cs, sn;
        const V3 g = d_angle_grad(v0, v1, v2, v3, k, cs, sn);
        const double prefactor = q.k_b * (-q.pair_cs[2 * pr] * sn + q.pair_cs[2 * pr + 1] * cs);
        fx += prefactor * g.x; fy += prefactor * g.y; fz += prefactor * g.z;
    }
    return {fx, fy, fz};
}

// One CTA per particle: the particle's vertex positions are staged in shared memory once (nv*24 B), every thread then
// gathers from there; the connectivity arrays are shared by all particles of the shape and stay in L1/L2.
__global__ void __launch_bounds__(256) k_rbc_forces_cta(RbcParams q, const double* __restrict__ vpos, double* __restrict__ vfrc,
                                                        const double* __restrict__ pstate) {
    extern __shared__ double s_pos[];
    const int pl = blockIdx.x;
    if (q.role && q.role[q.first_particle + pl] != 1) return;      // only the owner evaluates membrane forces
    const long long base = q.first_vertex + (long long)pl * q.nv;
    for (int v = threadIdx.x; v < q.nv; v += blockDim.x) {
        s_pos[v] = vpos[base + v]; s_pos[q.nv + v] = vpos[q.cap + base + v]; s_pos[2 * q.nv + v] = vpos[2 * q.cap + base + v];
    }
    __syncthreads();
    const double* st = pstate + (long long)(q.first_particle + pl) * 11;
    for (int v = threadIdx.x; v < q.nv; v += blockDim.x) {
        const V3 f = d_rbc_vertex_force(q, s_pos, q.nv, 0, v, st, base + v);
        vfrc[base + v] = f.x; vfrc[q.cap + base + v] = f.y; vfrc[2 * q.cap + base + v] = f.z;
    }
}
// Element-parallel version: every link / triangle / adjacent-triangle pair is evaluated ONCE per particle (instead of once
// per incident vertex) into a shared-memory chunk, and each thread then adds the chunk's contributions to its own
// vertices in ascending element order -- the same per-vertex summation order as RBCParticle::compute_forces, hence the
// same bits as the gather version, at ~1/3 of the fp64 divide/sqrt work. One CTA per particle, nv <= RF_THREADS*RF_VPT (vertices per
// thread: a template parameter so that the per-thread accumulators cost no more registers than the mesh needs).
#define RF_THREADS 256
#define RF_VPT_MAX 4
#define RF_CHUNK 7488          // doubles per chunk (58.5 KB): 2496 links x3, 416 triangles x18, 624 pairs x12
template <int RF_VPT>
__global__ void __launch_bounds__(RF_THREADS) k_rbc_forces_elem(RbcParams q, const double* __restrict__ vpos, double* __restrict__ vfrc,
                                                               const double* __restrict__ pstate, int nt, int nl, int np) {
    extern __shared__ double s_mem[];
    double* s_pos = s_mem;
    double* s_buf = s_mem + 3 * q.nv;
    const int pl = blockIdx.x;
    if (q.role && q.role[q.first_particle + pl] != 1) return;
    const long long base = q.first_vertex + (long long)pl * q.nv;
    for (int v = threadIdx.x; v < q.nv; v += RF_THREADS) {
        s_pos[v] = vpos[base + v]; s_pos[q.nv + v] = vpos[q.cap + base + v]; s_pos[2 * q.nv + v] = vpos[2 * q.cap + base + v];
    }
    const double* st = pstate + (long long)(q.first_particle + pl) * 11;
    const double area = st[0], volume = st[1];
    double fx[RF_VPT], fy[RF_VPT], fz[RF_VPT];
    int cur[RF_VPT], end[RF_VPT];
#pragma unroll
    for (int k = 0; k < RF_VPT; ++k) {
        fx[k] = fy[k] = fz[k] = 0.0;
        const int v = threadIdx.x + k * RF_THREADS;
        if (v < q.nv) {
            if (q.f_init) { fx[k] = q.f_init[base + v]; fy[k] = q.f_init[q.cap + base + v]; fz[k] = q.f_init[2 * q.cap + base + v]; }
            if (q.reg_p != 0.0) {
                fx[k] += (q.reg_p * (q.reg_t[0] - st[2])) / (double)q.nv;
                fy[k] += (q.reg_p * (q.reg_t[1] - st[3])) / (double)q.nv;
                fz[k] += (q.reg_p * (q.reg_t[2] - st[4])) / (double)q.nv;
            }
            cur[k] = q.v_link_off[v]; end[k] = q.v_link_off[v + 1];
        } else { cur[k] = end[k] = 0; }
    }
    __syncthreads();
    // ---- links
    const int CL = RF_CHUNK / 3;
    for (int l0 = 0; l0 < nl; l0 += CL) {
        const int ln = min(CL, nl - l0);
        for (int i = threadIdx.x; i < ln; i += RF_THREADS) {
            const int l = l0 + i;
            const V3 a = ld3(s_pos, q.nv, q.links[2 * l]), b = ld3(s_pos, q.nv, q.links[2 * l + 1]);
            const V3 dv = sub(a, b);
            const double L = nrm(dv);
            const double x = L / (q.link_len[l] * 2.2);
            const double df = -0.25 * q.link_ks[l] * (6. * x + (3. * x * x - 2. * x * x * x) / ((1 - x) * (1 - x))) + q.link_kp[l] / (L * L);
            s_buf[3 * i] = df * dv.x / L; s_buf[3 * i + 1] = df * dv.y / L; s_buf[3 * i + 2] = df * dv.z / L;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < RF_VPT; ++k) {
            while (cur[k] < end[k]) {
                const int code = q.v_link[cur[k]], l = code >> 1;
                if (l >= l0 + ln) break;
                const double* t = s_buf + 3 * (l - l0);
                if ((code & 1) == 0) { fx[k] += t[0]; fy[k] += t[1]; fz[k] += t[2]; } else { fx[k] -= t[0]; fy[k] -= t[1]; fz[k] -= t[2]; }
                ++cur[k];
            }
        }
        __syncthreads();
    }
    // ---- triangles: per corner {df_area*grad_A (3), df_vol*grad_V (3)}
#pragma unroll
    for (int k = 0; k < RF_VPT; ++k) { const int v = threadIdx.x + k * RF_THREADS; if (v < q.nv) { cur[k] = q.v_tri_off[v]; end[k] = q.v_tri_off[v + 1]; } }
    const double df_vol = q.k_v * (q.vol_desired - volume) / q.vol_desired;
    const int CT = RF_CHUNK / 18;
    for (int t0 = 0; t0 < nt; t0 += CT) {
        const int tn = min(CT, nt - t0);
        for (int i = threadIdx.x; i < tn; i += RF_THREADS) {
            const int tr = t0 + i;
            const V3 v0 = ld3(s_pos, q.nv, q.tris[3 * tr]), v1 = ld3(s_pos, q.nv, q.tris[3 * tr + 1]), v2 = ld3(s_pos, q.nv, q.tris[3 * tr + 2]);
            double local_area;
            const V3 g0 = d_area_grad(v0, v1, v2, 0, local_area), g1 = d_area_grad(v0, v1, v2, 1, local_area), g2 = d_area_grad(v0, v1, v2, 2, local_area);
            double df_area = 0.;
            df_area += q.k_ag * (q.shape_area - area) / q.shape_area;
            df_area += q.k_al * (q.tri_area[tr] - local_area) / q.tri_area[tr];
            const V3 w0 = crs(v1, v2), w1 = crs(v2, v0), w2 = crs(v0, v1);
            double* o = s_buf + 18 * i;
            o[0] = df_area * g0.x; o[1] = df_area * g0.y; o[2] = df_area * g0.z; o[3] = df_vol * (w0.x / 6.); o[4] = df_vol * (w0.y / 6.); o[5] = df_vol * (w0.z / 6.);
            o[6] = df_area * g1.x; o[7] = df_area * g1.y; o[8] = df_area * g1.z; o[9] = df_vol * (w1.x / 6.); o[10] = df_vol * (w1.y / 6.); o[11] = df_vol * (w1.z / 6.);
            o[12] = df_area * g2.x; o[13] = df_area * g2.y; o[14] = df_area * g2.z; o[15] = df_vol * (w2.x / 6.); o[16] = df_vol * (w2.y / 6.); o[17] = df_vol * (w2.z / 6.);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < RF_VPT; ++k) {
            while (cur[k] < end[k]) {
                const int code = q.v_tri[cur[k]], tr = code / 3;
                if (tr >= t0 + tn) break;
                const double* o = s_buf + 18 * (tr - t0) + 6 * (code % 3);
                fx[k] += o[0]; fy[k] += o[1]; fz[k] += o[2];
                fx[k] += o[3]; fy[k] += o[4]; fz[k] += o[5];
                ++cur[k];
            }
        }
        __syncthreads();
    }
    // ---- pairs: per corner prefactor*grad_theta (3)
#pragma unroll
    for (int k = 0; k < RF_VPT; ++k) { const int v = threadIdx.x + k * RF_THREADS; if (v < q.nv) { cur[k] = q.v_pair_off[v]; end[k] = q.v_pair_off[v + 1]; } }
    const int CP = RF_CHUNK / 12;
    for (int p0 = 0; p0 < np; p0 += CP) {
        const int pn = min(CP, np - p0);
        for (int i = threadIdx.x; i < pn; i += RF_THREADS) {
            const int pr = p0 + i;
            const V3 v0 = ld3(s_pos, q.nv, q.pairs[4 * pr]), v1 = ld3(s_pos, q.nv, q.pairs[4 * pr + 1]);
            const V3 v2 = ld3(s_pos, q.nv, q.pairs[4 * pr + 2]), v3 = ld3(s_pos, q.nv, q.pairs[4 * pr + 3]);
            // tri::grad_angle_between_pair (TriangleUtils.h:138-164), all four corners from the shared intermediate values
            const V3 e0 = sub(v1, v0), e1 = sub(v2, v0), e2 = sub(v3, v0);
            const double e1_norm = nrm(e1);
            const V3 e1u = divs(e1, e1_norm);
            const V3 n1 = crs(e0, e1), n2 = crs(e1, e2);
            const double n1_norm = nrm(n1), n2_norm = nrm(n2);
            const V3 n1u = divs(n1, n1_norm), n2u = divs(n2, n2_norm);
            const double cs = dot(n1u, n2u);
            const double sn = ((dot(sub(n1, n2), sub(v1, v3)) > 0) ? 1. : -1.) * nrm(crs(n1u, n2u));
            const double prefactor = q.k_b * (-q.pair_cs[2 * pr] * sn + q.pair_cs[2 * pr + 1] * cs);
            const double a0 = dot(e1u, sub(e1, e2)) / n2_norm, b0 = dot(e1u, sub(e1, e0)) / n1_norm;
            const double a1 = -e1_norm / n1_norm;
            const double a2 = dot(e1u, e2) / n2_norm, b2 = dot(e1u, e0) / n1_norm;
            const double a3 = -e1_norm / n2_norm;
            double* o = s_buf + 12 * i;
            o[0] = prefactor * (a0 * n2u.x + b0 * n1u.x); o[1] = prefactor * (a0 * n2u.y + b0 * n1u.y); o[2] = prefactor * (a0 * n2u.z + b0 * n1u.z);
            o[3] = prefactor * (a1 * n1u.x); o[4] = prefactor * (a1 * n1u.y); o[5] = prefactor * (a1 * n1u.z);
            o[6] = prefactor * (a2 * n2u.x + b2 * n1u.x); o[7] = prefactor * (a2 * n2u.y + b2 * n1u.y); o[8] = prefactor * (a2 * n2u.z + b2 * n1u.z);
            o[9] = prefactor * (a3 * n2u.x); o[10] = prefactor * (a3 * n2u.y); o[11] = prefactor * (a3 * n2u.z);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < RF_VPT; ++k) {
            while (cur[k] < end[k]) {
                const int code = q.v_pair[cur[k]], pr = code >> 2;
                if (pr >= p0 + pn) break;
                const double* o = s_buf + 12 * (pr - p0) + 3 * (code & 3);
                fx[k] += o[0]; fy[k] += o[1]; fz[k] += o[2];
                ++cur[k];
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < RF_VPT; ++k) {
        const int v = threadIdx.x + k * RF_THREADS;
        if (v < q.nv) { vfrc[base + v] = fx[k]; vfrc[q.cap + base + v] = fy[k]; vfrc[2 * q.cap + base + v] = fz[k]; }
    }
}

// Fallback for meshes whose vertices do not fit in shared memory: one thread per vertex, positions gathered from HBM/L2.
__global__ void __launch_bounds__(128) k_rbc_forces(RbcParams q, const double* __restrict__ vpos, double* __restrict__ vfrc,
                                                    const double* __restrict__ pstate) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)q.count * q.nv) return;
    const int pl = (int)(t / q.nv), v = (int)(t % q.nv);
    if (q.role && q.role[q.first_particle + pl] != 1) return;
    const long long base = q.first_vertex + (long long)pl * q.nv;
    const V3 f = d_rbc_vertex_force(q, vpos, q.cap, base, v, pstate + (long long)(q.first_particle + pl) * 11, base + v);
    vfrc[base + v] = f.x; vfrc[q.cap + base + v] = f.y; vfrc[2 * q.cap + base + v] = f.z;
}

// ------------------------------------------------------------------------------------------------ M2: capsule forces
// DeformableCapsuleParticle::compute_forces (src/fsi/DeformableCapsule.hh:60-150), neo-Hookean membrane FEM; see
// oracle/biblbm_oracle.c orc_capsule_forces for the line-by-line restatement (incl. the reference's `c` quirk). Returns
// the contribution of triangle (v0,v1,v2) with rest corners (r0,r1,r2) and rest area A0 to its corner `i`.
__device__ __forceinline__ V3 d_capsule_corner(V3 v0, V3 v1, V3 v2, V3 r0, V3 r1, V3 r2, double A0, double G, int i) {
    const V3 e0 = sub(v1, v0), e1 = sub(v2, v0), e0_0 = sub(r1, r0), e1_0 = sub(r2, r0);
    const double l = nrm(e0), lprime = nrm(e1), l0 = nrm(e0_0), l0prime = nrm(e1_0);
    const double e0_dot_e1 = dot(e0, e1);
    const double sin_phi = nrm(crs(e0, e1)) / (l * lprime);
    const double cos_phi = e0_dot_e1 / (l * lprime);
    const double sin_phi_0 = nrm(crs(e0_0, e1_0)) / (l0 * l0prime);
    const double cos_phi_0 = dot(e0_0, e1_0) / (l0 * l0prime);
    const double a = l / l0;
    const double b = (lprime * cos_phi / l0prime - l * cos_phi_0 / l0) / sin_phi_0;
    const double c = lprime * sin_phi / (lprime * sin_phi_0);
    const double D11 = (a * a - 1.) / 2.;
    const double D22 = (b * b + c * c - 1) / 2.;
    const double D12 = a * b / 2.;
    const double I1 = 2. * (D11 + D22);
    const double I2 = I1 + 4 * (D11 * D22 - D12 * D12);
    const double E0[3] = {e0.x, e0.y, e0.z}, E1[3] = {e1.x, e1.y, e1.z};
    double out[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double gl = i == 0 ? E0[d] / (-l) : (i == 1 ? E0[d] / l : 0.0);
        const double glp = i == 0 ? E1[d] / (-lprime) : (i == 1 ? 0.0 : E1[d] / lprime);
        double gc;
        if (i == 0) gc = (-1. / (l * lprime) + e0_dot_e1 / (l * l * l * lprime)) * E0[d] + (-1. / (l * lprime) + e0_dot_e1 / (l * lprime * lprime * lprime)) * E1[d];
        else if (i == 1) gc = (1. / (l * lprime)) * E1[d] - (e0_dot_e1 / (l * l * l * lprime)) * E0[d];
        else gc = (1. / (l * lprime)) * E0[d] - (e0_dot_e1 / (l * lprime * lprime * lprime)) * E1[d];
        const double grad_D11 = (l / (l0 * l0)) * gl;
        const double grad_D22 = (1. / (sin_phi_0 * sin_phi_0)) * (
                (lprime / (l0prime * l0prime) - l * cos_phi * cos_phi_0 / (l0 * l0prime)) * glp +
                (l * (cos_phi_0 * cos_phi_0) / (l0 * l0) - lprime * cos_phi * cos_phi_0 / (l0prime * l0)) * gl -
                (l * lprime * cos_phi_0 / (l0 * l0prime)) * gc);
        const double grad_D12 = (1. / (2. * l0prime * sin_phi_0)) * (
                (lprime * cos_phi / l0prime - 2. * l * cos_phi_0 / l0) * gl + (l * cos_phi / l0prime) * glp + (l * lprime / l0prime) * gc);
        const double grad_I1 = 2. * (grad_D11 + grad_D22);
        const double grad_I2 = (2. + 4. * D22) * grad_D11 + (2. + 4. * D11) * grad_D22 - 8. * D12 * grad_D12;
        out[d] = -G / 2. * A0 * (grad_I1 - (1. / ((1 + I2) * (1 + I2))) * grad_I2);
    }
    return {out[0], out[1], out[2]};
}
struct CapParams {
    int nv; long long first_vertex, cap; int first_particle, count;
    const int* tris; const double* tri_area; const double* rest;   // rest: [nv][3]
    const int* v_tri_off; const int* v_tri;
    double G; const int* role; int use_smem; int f_init;
};
// One CTA per particle, per-vertex gather over incident triangles in ascending order (= the reference's accumulation order).
__global__ void __launch_bounds__(256) k_capsule_forces(CapParams q, const double* __restrict__ vpos, double* __restrict__ vfrc) {
    extern __shared__ double s_cp[];
    const int pl = blockIdx.x;
    if (q.role && q.role[q.first_particle + pl] != 1) return;
    const long long base = q.first_vertex + (long long)pl * q.nv;
    const double* P = vpos; long long capP = q.cap, baseP = base;
    if (q.use_smem) {
        for (int v = threadIdx.x; v < q.nv; v += blockDim.x) { s_cp[v] = vpos[base + v]; s_cp[q.nv + v] = vpos[q.cap + base + v]; s_cp[2 * q.nv + v] = vpos[2 * q.cap + base + v]; }
        __syncthreads();
        P = s_cp; capP = q.nv; baseP = 0;
    }
    for (int v = threadIdx.x; v < q.nv; v += blockDim.x) {
        double fx = 0.0, fy = 0.0, fz = 0.0;
        if (q.f_init) { fx = vfrc[base + v]; fy = vfrc[q.cap + base + v]; fz = vfrc[2 * q.cap + base + v]; }
        for (int e = q.v_tri_off[v]; e < q.v_tri_off[v + 1]; ++e) {
            const int code = q.v_tri[e], tr = code / 3, k = code % 3;
            const int i0 = q.tris[3 * tr], i1 = q.tris[3 * tr + 1], i2 = q.tris[3 * tr + 2];
            const V3 r0 = {q.rest[3 * i0], q.rest[3 * i0 + 1], q.rest[3 * i0 + 2]}, r1 = {q.rest[3 * i1], q.rest[3 * i1 + 1], q.rest[3 * i1 + 2]};
            const V3 r2 = {q.rest[3 * i2], q.rest[3 * i2 + 1], q.rest[3 * i2 + 2]};
            const V3 f = d_capsule_corner(ld3(P, capP, baseP + i0), ld3(P, capP, baseP + i1), ld3(P, capP, baseP + i2), r0, r1, r2, q.tri_area[tr], q.G, k);
            fx += f.x; fy += f.y; fz += f.z;
        }
        vfrc[base + v] = fx; vfrc[q.cap + base + v] = fy; vfrc[2 * q.cap + base + v] = fz;
    }
}

// ------------------------------------------------------------------------------------------------ platelets: SemiRigidParticle3D
// SemiRigidParticle3D::compute_forces (src/fsi/SemiRigidParticle.hh:56-96). One CTA per particle; the radial spring terms are
// computed in parallel, their reaction is accumulated by one thread in vertex order (the reference's sequential sum), and each
// thread then adds radial term, reaction share and its incident link springs in link order.
struct SemiParams {
    int nv; long long first_vertex, cap; int first_particle;
    const int* links; const double* link_len; const int* v_link_off; const int* v_link;
    const double* rest_radial;      // [nv] |shape vertex - shape centre|
    double k_in, k_out; const int* role; int f_init;
};
__global__ void __launch_bounds__(256) k_semirigid_forces(SemiParams q, const double* __restrict__ vpos, double* __restrict__ vfrc,
                                                          const double* __restrict__ pstate) {
    extern __shared__ double s_sr[];
    double* s_pos = s_sr;               // [3][nv]
    double* s_df = s_sr + 3 * q.nv;     // [3][nv]
    __shared__ double s_com[3];
    const int pl = blockIdx.x;
    if (q.role && q.role[q.first_particle + pl] != 1) return;
    const long long base = q.first_vertex + (long long)pl * q.nv;
    const double* st = pstate + (long long)(q.first_particle + pl) * 11;
    for (int v = threadIdx.x; v < q.nv; v += blockDim.x) {
        const double x = vpos[base + v], y = vpos[q.cap + base + v], z = vpos[2 * q.cap + base + v];
        s_pos[v] = x; s_pos[q.nv + v] = y; s_pos[2 * q.nv + v] = z;
        const double dx = x - st[2], dy = y - st[3], dz = z - st[4];
        const double l = sqrt(dx * dx + dy * dy + dz * dz);
        const double l0 = q.rest_radial[v];
        const double k = -q.k_out * (l - l0) / (l0 * l);
        s_df[v] = k * dx; s_df[q.nv + v] = k * dy; s_df[2 * q.nv + v] = k * dz;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double acc = 0.0;
        for (int v = 0; v < q.nv; ++v) acc -= s_df[threadIdx.x * q.nv + v];
        s_com[threadIdx.x] = acc / (double)q.nv;
    }
    __syncthreads();
    for (int v = threadIdx.x; v < q.nv; v += blockDim.x) {
        double f[3] = {0.0, 0.0, 0.0};
        if (q.f_init) { f[0] = vfrc[base + v]; f[1] = vfrc[q.cap + base + v]; f[2] = vfrc[2 * q.cap + base + v]; }
#pragma unroll
        for (int d = 0; d < 3; ++d) { f[d] += s_df[d * q.nv + v]; f[d] += s_com[d]; }
        for (int e = q.v_link_off[v]; e < q.v_link_off[v + 1]; ++e) {
            const int code = q.v_link[e], l = code >> 1, end = code & 1;
            const int i0 = q.links[2 * l], i1 = q.links[2 * l + 1];
            const double dv[3] = {s_pos[i0] - s_pos[i1], s_pos[q.nv + i0] - s_pos[q.nv + i1], s_pos[2 * q.nv + i0] - s_pos[2 * q.nv + i1]};
            const double L = sqrt(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2]);
            const double df = -q.k_in * (L - q.link_len[l]) / (L * q.link_len[l]);
#pragma unroll
            for (int d = 0; d < 3; ++d) { if (end == 0) f[d] += df * dv[d]; else f[d] -= df * dv[d]; }
        }
        vfrc[base + v] = f[0]; vfrc[q.cap + base + v] = f[1]; vfrc[2 * q.cap + base + v] = f[2];
    }
}

// ------------------------------------------------------------------------------------------------ M3: RigidParticle3D
// src/fsi/RigidParticle.hh:31-120 (update, move_vertices, compute_forces), Quaternion.h, Matrix3.h:129-144. One CTA per particle;
// rigid state RS = 24 doubles: com[3], velocity[3], angular velocity[3], quaternion[4], force[3], torque[3], density, damping,
// relaxation factor, scale, pad. Sums over vertices run sequentially in one thread (the reference's order).
#define RS 24
__device__ __forceinline__ void d_quat_rot(const double* q, const double* v, double* r) {
    r[0] = (1 - 2 * (q[2] * q[2] + q[3] * q[3])) * v[0] + 2 * (q[1] * q[2] - q[0] * q[3]) * v[1] + 2 * (q[0] * q[2] + q[1] * q[3]) * v[2];
    r[1] = 2 * (q[1] * q[2] + q[0] * q[3]) * v[0] + (1 - 2 * (q[1] * q[1] + q[3] * q[3])) * v[1] + 2 * (q[2] * q[3] - q[0] * q[1]) * v[2];
    r[2] = 2 * (q[1] * q[3] - q[0] * q[2]) * v[0] + 2 * (q[0] * q[1] + q[2] * q[3]) * v[1] + (1 - 2 * (q[1] * q[1] + q[2] * q[2])) * v[2];
}
__device__ __forceinline__ void d_quat_inv_rot(const double* q, const double* v, double* r) {
    r[0] = (1 - 2 * (q[2] * q[2] + q[3] * q[3])) * v[0] + 2 * (q[1] * q[2] + q[0] * q[3]) * v[1] + 2 * (q[1] * q[3] - q[0] * q[2]) * v[2];
    r[1] = 2 * (q[1] * q[2] - q[0] * q[3]) * v[0] + (1 - 2 * (q[1] * q[1] + q[3] * q[3])) * v[1] + 2 * (q[0] * q[1] + q[2] * q[3]) * v[2];
    r[2] = 2 * (q[0] * q[2] + q[1] * q[3]) * v[0] + 2 * (q[2] * q[3] - q[0] * q[1]) * v[1] + (1 - 2 * (q[1] * q[1] + q[2] * q[2])) * v[2];
}
__device__ __forceinline__ void d_mat_vec(const double* m, const double* v, double* r) {
    r[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2]; r[1] = m[3] * v[0] + m[4] * v[1] + m[5] * v[2]; r[2] = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
}
struct RigidParams {
    int nv, nt; long long first_vertex, cap; int first_particle;
    const int* tris; const int* v_tri_off; const int* v_tri; const double* rest;   // rest: [nv][3] shape vertices
    double* rigid; double inertia[9]; double shape_volume; const int* role; int f_init;
};
// compute_forces: f_i += relaxation * M_i * (V + W x (x_i - X) - u_i), M_i = sum of a third of the incident triangle areas
__global__ void __launch_bounds__(256) k_rigid_forces(RigidParams q, const double* __restrict__ vpos, const double* __restrict__ vvel, double* __restrict__ vfrc) {
    const int pl = blockIdx.x;
    if (q.role && q.role[q.first_particle + pl] != 1) return;
    const long long base = q.first_vertex + (long long)pl * q.nv;
    const double* rs = q.rigid + (long long)pl * RS;
    for (int v = threadIdx.x; v < q.nv; v += blockDim.x) {
        double M = 0.0;
        for (int e = q.v_tri_off[v]; e < q.v_tri_off[v + 1]; ++e) {
            const int tr = q.v_tri[e] / 3;
            const V3 v0 = ld3(vpos, q.cap, base + q.tris[3 * tr]), v1 = ld3(vpos, q.cap, base + q.tris[3 * tr + 1]), v2 = ld3(vpos, q.cap, base + q.tris[3 * tr + 2]);
            const double local_area = 0.5 * nrm(crs(sub(v1, v0), sub(v2, v0)));
            M += local_area / 3.;
        }
        const double p[3] = {vpos[base + v], vpos[q.cap + base + v], vpos[2 * q.cap + base + v]};
        const double r[3] = {p[0] - rs[0], p[1] - rs[1], p[2] - rs[2]};
        const double w[3] = {rs[7] * r[2] - rs[8] * r[1], rs[8] * r[0] - rs[6] * r[2], rs[6] * r[1] - rs[7] * r[0]};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double f = q.f_init ? vfrc[d * q.cap + base + v] : 0.0;
            f += rs[21] * M * ((rs[3 + d] + w[d]) - vvel[d * q.cap + base + v]);
            vfrc[d * q.cap + base + v] = f;
        }
    }
}
// move_vertices + update + map_center_of_mass_to_periodic_grid (do_move = 0: only place the nodes and refresh the bounding box)
__global__ void __launch_bounds__(256) k_rigid_move(Geo g, RigidParams q, double* __restrict__ vpos, const double* __restrict__ vfrc,
                                                    double* __restrict__ pstate, int do_move) {
    __shared__ double s_rs[RS];
    __shared__ double s_R[9];
    __shared__ double s_red[6][8];
    const int pl = blockIdx.x;
    if (q.role && q.role[q.first_particle + pl] == 0) return;
    const long long base = q.first_vertex + (long long)pl * q.nv;
    double* rsg = q.rigid + (long long)pl * RS;
    if (threadIdx.x < RS) s_rs[threadIdx.x] = rsg[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        double* com = s_rs; double* vel = s_rs + 3; double* om = s_rs + 6; double* qt = s_rs + 9; double* F = s_rs + 13; double* Tq = s_rs + 16;
        if (do_move) {
            const double density = s_rs[19], damping = s_rs[20], scale = s_rs[22];
            for (int d = 0; d < 3; ++d) F[d] = Tq[d] = 0.0;
            for (int i = 0; i < q.nv; ++i) {
                const double f[3] = {vfrc[base + i], vfrc[q.cap + base + i], vfrc[2 * q.cap + base + i]};
                const double r[3] = {vpos[base + i] - com[0], vpos[q.cap + base + i] - com[1], vpos[2 * q.cap + base + i] - com[2]};
                for (int d = 0; d < 3; ++d) F[d] += f[d];
                Tq[0] += r[1] * f[2] - r[2] * f[1]; Tq[1] += r[2] * f[0] - r[0] * f[2]; Tq[2] += r[0] * f[1] - r[1] * f[0];
            }
            double om_body[3]; d_quat_inv_rot(qt, om, om_body);
            const double sI = density * ((((scale * scale) * scale) * scale) * scale);      // std::pow(scale_, 5) as the reference build evaluates it
            double I[9];
            for (int k = 0; k < 9; ++k) I[k] = sI * q.inertia[k];
            const double mass = (scale * scale * scale) * q.shape_volume * density;
            double v_dot[3], tq_body[3], Iw[3], rhs[3], sol[3], rot[3];
            for (int d = 0; d < 3; ++d) v_dot[d] = F[d] / mass - damping * vel[d];
            d_quat_inv_rot(qt, Tq, tq_body);
            d_mat_vec(I, om_body, Iw);
            const double cr[3] = {om_body[1] * Iw[2] - om_body[2] * Iw[1], om_body[2] * Iw[0] - om_body[0] * Iw[2], om_body[0] * Iw[1] - om_body[1] * Iw[0]};
            for (int d = 0; d < 3; ++d) rhs[d] = tq_body[d] - cr[d];
            const double* a = I;
            const double det = a[0] * a[4] * a[8] - a[0] * a[5] * a[7] - a[1] * a[3] * a[8] + a[1] * a[5] * a[6] + a[2] * a[3] * a[7] - a[2] * a[4] * a[6];
            sol[0] = (rhs[2] * (a[1] * a[5] - a[2] * a[4]) + rhs[1] * (a[2] * a[7] - a[1] * a[8]) + rhs[0] * (a[4] * a[8] - a[5] * a[7])) / det;
            sol[1] = (rhs[2] * (a[2] * a[3] - a[0] * a[5]) + rhs[1] * (a[0] * a[8] - a[2] * a[6]) + rhs[0] * (a[5] * a[6] - a[3] * a[8])) / det;
            sol[2] = (rhs[2] * (a[0] * a[4] - a[1] * a[3]) + rhs[1] * (a[1] * a[6] - a[0] * a[7]) + rhs[0] * (a[3] * a[7] - a[4] * a[6])) / det;
            d_quat_rot(qt, sol, rot);
            for (int d = 0; d < 3; ++d) { vel[d] += v_dot[d]; om[d] += rot[d] - damping * om[d]; }
            for (int d = 0; d < 3; ++d) com[d] += vel[d];
            const double om_norm = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
            if (om_norm >= 10 * 2.220446049250313e-16) {        // exponential map (sin / cos: last-bit differences vs the host libm are possible)
                const double ax[3] = {om[0] / om_norm, om[1] / om_norm, om[2] / om_norm};
                const double sh = sin(om_norm / 2);
                const double p[4] = {cos(om_norm / 2), sh * ax[0], sh * ax[1], sh * ax[2]};
                const double r4[4] = {p[0] * qt[0] - (p[1] * qt[1] + p[2] * qt[2] + p[3] * qt[3]),
                                      p[0] * qt[1] + qt[0] * p[1] + p[2] * qt[3] - p[3] * qt[2],
                                      p[0] * qt[2] + qt[0] * p[2] + p[3] * qt[1] - p[1] * qt[3],
                                      p[0] * qt[3] + qt[0] * p[3] + p[1] * qt[2] - p[2] * qt[1]};
                for (int k = 0; k < 4; ++k) qt[k] = r4[k];
            }
        }
        s_R[0] = 1 - 2 * (qt[2] * qt[2] + qt[3] * qt[3]); s_R[1] = 2 * (qt[1] * qt[2] - qt[0] * qt[3]); s_R[2] = 2 * (qt[0] * qt[2] + qt[1] * qt[3]);
        s_R[3] = 2 * (qt[1] * qt[2] + qt[0] * qt[3]); s_R[4] = 1 - 2 * (qt[1] * qt[1] + qt[3] * qt[3]); s_R[5] = 2 * (qt[2] * qt[3] - qt[0] * qt[1]);
        s_R[6] = 2 * (qt[1] * qt[3] - qt[0] * qt[2]); s_R[7] = 2 * (qt[0] * qt[1] + qt[2] * qt[3]); s_R[8] = 1 - 2 * (qt[1] * qt[1] + qt[2] * qt[2]);
    }
    __syncthreads();
    // node positions and bounding box, then the periodic remap of the centre of mass
    double delta[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int n = d == 0 ? g.nx : (d == 1 ? g.ny : g.nz);
        double x = s_rs[d];
        if (g.per[d]) { const double L = (double)n; x = x - L * floor((x - (0.0 - 0.5)) * (1. / L)); }
        delta[d] = x - s_rs[d];
    }
    double mn[3] = {1.7976931348623157e308, 1.7976931348623157e308, 1.7976931348623157e308};
    double mx[3] = {-1.7976931348623157e308, -1.7976931348623157e308, -1.7976931348623157e308};
    const double scale = s_rs[22];
    for (int v = threadIdx.x; v < q.nv; v += blockDim.x) {
        const double sv[3] = {scale * q.rest[3 * v], scale * q.rest[3 * v + 1], scale * q.rest[3 * v + 2]};
        double r[3]; d_mat_vec(s_R, sv, r);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double x = s_rs[d] + r[d];
            mn[d] = fmin(mn[d], x); mx[d] = fmax(mx[d], x);
            vpos[d * q.cap + base + v] = x + delta[d];
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        for (int o = 16; o > 0; o >>= 1) { mn[d] = fmin(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o)); mx[d] = fmax(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o)); }
        if ((threadIdx.x & 31) == 0) { s_red[2 * d][threadIdx.x >> 5] = mn[d]; s_red[2 * d + 1][threadIdx.x >> 5] = mx[d]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double* st = pstate + (long long)(q.first_particle + pl) * 11;
        for (int d = 0; d < 3; ++d) {
            double lo = s_red[2 * d][0], hi = s_red[2 * d + 1][0];
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = fmin(lo, s_red[2 * d][w]); hi = fmax(hi, s_red[2 * d + 1][w]); }
            s_rs[d] = s_rs[d] + delta[d];
            st[2 + d] = s_rs[d]; st[5 + 2 * d] = lo + delta[d]; st[5 + 2 * d + 1] = hi + delta[d];
        }
        for (int k = 0; k < RS; ++k) rsg[k] = s_rs[k];
    }
}

// ------------------------------------------------------------------------------------------------ M4: wall / particle-particle forces
// fsi/Potentials.h:29-71: type 1 = MorsePotential(r_min, beta, strength), 2 = SpringPotential(min_dist, strength); cutoff = r_min
__device__ __forceinline__ double d_potential(const PotentialDev& q, double dist) {
    if (q.type == 1) {
        if (dist < 0) dist = 0;
        const double x = exp(q.beta * (q.r_min - dist));
        return ((2 * q.beta) * q.strength) * x * (x - 1);
    }
    return (q.r_min - dist) * q.strength;
}
// WallInteraction::apply_force (src/fsi/ForceDecorator.hh:25-33), run BEFORE the membrane force kernels, which then start
// from these values (decorators act first: src/fsi/ImmersedBoundaryDynamics_forces.hh:293-297). One thread per vertex.
__global__ void __launch_bounds__(256) k_wall_interaction(Geo g, PotentialDev pot, long long nvert, long long cap, const double* __restrict__ vpos,
                                                          double* __restrict__ vfrc) {
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= nvert) return;
    const double p[3] = {vpos[v], vpos[cap + v], vpos[2 * cap + v]};
    double f[3] = {0.0, 0.0, 0.0};
    const double dist = d_plates_distance(g, p);
    if (dist <= pot.r_min) {
        double n[3]; d_boundary_normal(g, p, n);
        const double I = d_potential(pot, dist);
#pragma unroll
        for (int d = 0; d < 3; ++d) f[d] -= I * n[d];
    }
    vfrc[v] = f[0]; vfrc[cap + v] = f[1]; vfrc[2 * cap + v] = f[2];
}

// Particle-particle interaction (compute_particle_particle_interaction_impl, src/fsi/ImmersedBoundaryDynamics_forces.hh:336-381;
// collision::CollisionHandler, src/fsi/CollisionForces.h:62-217; Grid, src/fsi/Grid.hh:36-103). The reference bins the
// vertices into a uniform grid and walks it once, applying +df / -df to both members of each pair. Here every vertex
// GATHERS its own force from exactly the same set of pairs (no atomics on forces): the cells the reference would have
// paired with the vertex's cell, including what the reference leaves out -- only the 7 neighbours with non-negative
// offsets are visited from a bulk cell (so a pair is seen iff its cell offset is one of those 7 or their negatives), and
// periodic images are only paired between directly opposite boundary cells.
struct PPGrid {
    int nc[3]; double g0[3], cs[3], lo[3], hi[3]; int per[3]; double L[3];
    // multi-slab: dead slots are not binned; a slab evaluates the vertices that lie in ITS x-range [x_lo, x_hi) (positions clamped into
    // the lattice), whoever owns them -- every partner within the cut-off is resident there -- and ghost holders return their part to the owner
    const int* role; const int* vslot; double x_lo, x_hi, x_min, x_max;
};
__device__ __forceinline__ bool d_pp_live(const PPGrid& G, long long v) { return !G.role || G.role[G.vslot[v]] != 0; }
__device__ __forceinline__ bool d_pp_cell(const PPGrid& G, const double* p, int* c) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (!(p[d] > G.lo[d] && p[d] < G.hi[d])) return false;          // coll_domain.contains(pos)
        c[d] = (int)floor((p[d] - G.g0[d]) / G.cs[d]);
    }
    return true;
}
__global__ void __launch_bounds__(256) k_pp_build(PPGrid G, long long nvert, long long cap, const double* __restrict__ vpos, int* head, int* next) {
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= nvert) return;
    const double p[3] = {vpos[v], vpos[cap + v], vpos[2 * cap + v]};
    int c[3];
    if (!d_pp_live(G, v) || !d_pp_cell(G, p, c)) { next[v] = -2; return; }
    const long long cell = (long long)c[2] + (long long)G.nc[2] * ((long long)c[1] + (long long)G.nc[1] * c[0]);
    next[v] = atomicExch(&head[cell], (int)v);
}
__global__ void __launch_bounds__(256) k_pp_clear(PPGrid G, long long nvert, long long cap, const double* __restrict__ vpos, int* head) {
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= nvert) return;
    const double p[3] = {vpos[v], vpos[cap + v], vpos[2 * cap + v]};
    int c[3];
    if (!d_pp_live(G, v) || !d_pp_cell(G, p, c)) return;
    head[(long long)c[2] + (long long)G.nc[2] * ((long long)c[1] + (long long)G.nc[1] * c[0])] = -1;
}
__device__ __forceinline__ void d_pp_cell_sum(const PPGrid& G, const PotentialDev& pot, double cut_sqr, const int* __restrict__ head, const int* __restrict__ next,
                                              const double* __restrict__ vpos, long long cap, const int* __restrict__ vobj, int ci, int cj, int ck,
                                              long long self, int obj, const double* p, double* f) {
    for (int b = head[(long long)ck + (long long)G.nc[2] * ((long long)cj + (long long)G.nc[1] * ci)]; b >= 0; b = next[b]) {
        if (b == self || vobj[b] == obj) continue;
        double dr[3] = {vpos[b] - p[0], vpos[cap + b] - p[1], vpos[2 * cap + b] - p[2]};
#pragma unroll
        for (int d = 0; d < 3; ++d) if (G.per[d]) dr[d] = d_min_mod(dr[d], G.L[d], 1. / G.L[d]);
        const double dist_sqr = dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2];
        if (dist_sqr <= cut_sqr) {
            const double dist = sqrt(dist_sqr);
            const double k = -d_potential(pot, dist) / dist;
#pragma unroll
            for (int d = 0; d < 3; ++d) f[d] += k * dr[d];
        }
    }
}
__global__ void __launch_bounds__(128) k_pp_forces(PPGrid G, PotentialDev pot, long long nvert, long long cap, const double* __restrict__ vpos,
                                                   const int* __restrict__ vobj, const int* __restrict__ head, const int* __restrict__ next,
                                                   double* __restrict__ vfrc) {
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= nvert || next[v] == -2) return;
    const double p[3] = {vpos[v], vpos[cap + v], vpos[2 * cap + v]};
    if (G.role) { const double xc = fmin(fmax(p[0], G.x_min), G.x_max); if (!(xc >= G.x_lo && xc < G.x_hi)) return; }
    int c[3];
    d_pp_cell(G, p, c);
    const int obj = vobj[v];
    const double cut_sqr = pot.r_min * pot.r_min;
    const int lo[3] = {1, 1, 1}, hi[3] = {G.nc[0] - 2, G.nc[1] - 2, G.nc[2] - 2};
    const bool bulk = c[0] >= lo[0] && c[0] <= hi[0] && c[1] >= lo[1] && c[1] <= hi[1] && c[2] >= lo[2] && c[2] <= hi[2];
    double f[3] = {0.0, 0.0, 0.0};
    const int off[7][3] = {{0, 1, 0}, {0, 1, 1}, {0, 0, 1}, {1, 0, 0}, {1, 1, 0}, {1, 1, 1}, {1, 0, 1}};
    if (bulk) {
        d_pp_cell_sum(G, pot, cut_sqr, head, next, vpos, cap, vobj, c[0], c[1], c[2], v, obj, p, f);            // same cell
        for (int o = 0; o < 7; ++o)                                                                               // this vertex is the pair's first member
            d_pp_cell_sum(G, pot, cut_sqr, head, next, vpos, cap, vobj, c[0] + off[o][0], c[1] + off[o][1], c[2] + off[o][2], v, obj, p, f);
    }
    for (int o = 0; o < 7; ++o) {                                                                                 // ... or its second member
        const int b[3] = {c[0] - off[o][0], c[1] - off[o][1], c[2] - off[o][2]};
        if (b[0] >= lo[0] && b[0] <= hi[0] && b[1] >= lo[1] && b[1] <= hi[1] && b[2] >= lo[2] && b[2] <= hi[2])
            d_pp_cell_sum(G, pot, cut_sqr, head, next, vpos, cap, vobj, b[0], b[1], b[2], v, obj, p, f);
    }
    if (bulk) {
        // periodic faces / edges / corner: the cell at the low end of every direction in D is paired with the cell at the high end of all of them
        for (int mask = 1; mask < 8; ++mask) {
            bool ok = true, all_lo = true, all_hi = true;
#pragma unroll
            for (int d = 0; d < 3; ++d) if (mask & (1 << d)) {
                if (!G.per[d]) ok = false;
                if (c[d] != lo[d]) all_lo = false;
                if (c[d] != hi[d]) all_hi = false;
            }
            if (!ok) continue;
            int q[3] = {c[0], c[1], c[2]};
            if (all_lo) {
                for (int d = 0; d < 3; ++d) if (mask & (1 << d)) q[d] = hi[d];
                d_pp_cell_sum(G, pot, cut_sqr, head, next, vpos, cap, vobj, q[0], q[1], q[2], v, obj, p, f);
            }
            if (all_hi) {
                for (int d = 0; d < 3; ++d) if (mask & (1 << d)) q[d] = lo[d];
                d_pp_cell_sum(G, pot, cut_sqr, head, next, vpos, cap, vobj, q[0], q[1], q[2], v, obj, p, f);
            }
        }
    }
    vfrc[v] += f[0]; vfrc[cap + v] += f[1]; vfrc[2 * cap + v] += f[2];
}
// multi-slab: the vertex forces of ghost copies start from zero before the pair forces are gathered (they only carry the partial pair
// force of the vertices inside this slab back to the owner)
__global__ void k_zero_ghost_forces(const int* __restrict__ role, const int* __restrict__ vslot, long long nvert, long long cap, double* __restrict__ vfrc) {
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= nvert || role[vslot[v]] == 1) return;
    vfrc[v] = 0.0; vfrc[cap + v] = 0.0; vfrc[2 * cap + v] = 0.0;
}
__global__ void k_fill_slots(int first_particle, int count, int nv, long long first_vertex, int* role, int* gid, int* slot_nv, long long* slot_fv) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int s = first_particle + k;
    role[s] = 1; gid[s] = s; slot_nv[s] = nv; slot_fv[s] = first_vertex + (long long)k * nv;
}
__global__ void k_fill_vobj(int* vobj, long long first_vertex, int nv, int first_particle, long long n) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t < n) vobj[first_vertex + t] = first_particle + (int)(t / nv);
}

// ------------------------------------------------------------------------------------------------ I4: move + update
// One CTA per particle. Triangle terms are computed in parallel in chunks and accumulated by one thread in triangle
// order, so area / volume / centre of mass carry exactly the reference's sequential rounding.
#define UPD_CHUNK 512
__global__ void __launch_bounds__(256) k_move_update(Geo g, int nv, int nt, const int* __restrict__ tris, long long first_vertex, long long cap,
                                                     int first_particle, double* vpos, const double* __restrict__ vvel, double* pstate, int do_move,
                                                     const uint8_t* __restrict__ skip_mask) {
    __shared__ double s_dA[UPD_CHUNK], s_vol[UPD_CHUNK], s_cx[UPD_CHUNK], s_cy[UPD_CHUNK], s_cz[UPD_CHUNK];
    __shared__ double s_acc[5];
    __shared__ double s_red[6][8];
    __shared__ double s_delta[3];
    const int pl = blockIdx.x;
    if (skip_mask && skip_mask[first_particle + pl]) return;
    if (g.role) {
        const int r = g.role[first_particle + pl];
        if (r == 0) return;
        if (r != 1) do_move = 0;            // ghost copies are advected by their owner
    }
    const long long base = first_vertex + (long long)pl * nv;
    double* st = pstate + (long long)(first_particle + pl) * 11;
    double mn[3] = {1.7976931348623157e308, 1.7976931348623157e308, 1.7976931348623157e308};
    double mx[3] = {-1.7976931348623157e308, -1.7976931348623157e308, -1.7976931348623157e308};
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
#pragma unroll
        double x[3] = {vpos[base + v], vpos[cap + base + v], vpos[2 * cap + base + v]};
        if (do_move) {
            const double u[3] = {vvel[base + v], vvel[cap + base + v], vvel[2 * cap + base + v]};
            const double old[3] = {x[0], x[1], x[2]};
#pragma unroll
            for (int d = 0; d < 3; ++d) x[d] += u[d];
            if (g.plates) {   // specular reflection (DeformableParticle3D.hh:33-50; Boundary::trace_ray / get_normal)
                double t;
                if (d_boundary_trace_ray(g, old, x, t)) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) x[d] = old[d] + t * u[d];
                    double n[3]; d_boundary_normal(g, x, n);
                    const double un = u[0] * n[0] + u[1] * n[1] + u[2] * n[2];
#pragma unroll
                    for (int d = 0; d < 3; ++d) { const double vpc = u[d] - 2. * un * n[d]; x[d] += vpc * (1. - t); }
                }
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) vpos[d * cap + base + v] = x[d];
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) { mn[d] = fmin(mn[d], x[d]); mx[d] = fmax(mx[d], x[d]); }
    }
    // bounding box: order-independent min/max
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        for (int o = 16; o > 0; o >>= 1) { mn[d] = fmin(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o)); mx[d] = fmax(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o)); }
        if ((threadIdx.x & 31) == 0) { s_red[2 * d][threadIdx.x >> 5] = mn[d]; s_red[2 * d + 1][threadIdx.x >> 5] = mx[d]; }
    }
    if (threadIdx.x == 0) { for (int k = 0; k < 5; ++k) s_acc[k] = 0.0; }
    __syncthreads();
    for (int t0 = 0; t0 < nt; t0 += UPD_CHUNK) {
        const int cn = min(UPD_CHUNK, nt - t0);
        for (int k = threadIdx.x; k < cn; k += blockDim.x) {
            const int t = t0 + k;
            const V3 v0 = ld3(vpos, cap, base + tris[3 * t]), v1 = ld3(vpos, cap, base + tris[3 * t + 1]), v2 = ld3(vpos, cap, base + tris[3 * t + 2]);
            const V3 centroid = {(v0.x + v1.x + v2.x) / 3.0, (v0.y + v1.y + v2.y) / 3.0, (v0.z + v1.z + v2.z) / 3.0};
            const V3 n = crs(sub(v1, v0), sub(v2, v0));
            double dA = nrm(n);
            dA *= 0.5;                                            // tri::normal_and_area, TriangleUtils.h:27-34
            s_dA[k] = dA;
            s_vol[k] = dot(v0, crs(v1, v2)) / 6.;                 // tri::signed_volume, TriangleUtils.h:105-110
            s_cx[k] = centroid.x * dA; s_cy[k] = centroid.y * dA; s_cz[k] = centroid.z * dA;
        }
        __syncthreads();
        if (threadIdx.x < 5) {
            const double* src = threadIdx.x == 0 ? s_dA : threadIdx.x == 1 ? s_vol : threadIdx.x == 2 ? s_cx : threadIdx.x == 3 ? s_cy : s_cz;
            double acc = s_acc[threadIdx.x];
            for (int k = 0; k < cn; ++k) acc += src[k];
            s_acc[threadIdx.x] = acc;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) / 32;
        double bb[6];
        for (int k = 0; k < 6; ++k) {
            double r = s_red[k][0];
            for (int w = 1; w < nw; ++w) r = (k & 1) ? fmax(r, s_red[k][w]) : fmin(r, s_red[k][w]);
            bb[k] = r;
        }
        const double area = s_acc[0];
        double com[3] = {s_acc[2] / area, s_acc[3] / area, s_acc[4] / area};
        // map_center_of_mass_to_periodic_grid: remap_pos (Periodicity.h:83-87) then set_center_of_mass (ParticleBase.hh:122-138)
        const int n[3] = {g.nx, g.ny, g.nz};
        for (int d = 0; d < 3; ++d) {
            double x = com[d];
            if (g.per[d]) { const double L = (double)n[d], Linv = 1. / L; x = x - L * floor((x - (0.0 - 0.5)) * Linv); }
            s_delta[d] = x - com[d];
            com[d] = x;
            bb[2 * d] += s_delta[d]; bb[2 * d + 1] += s_delta[d];
        }
        st[0] = area; st[1] = s_acc[1]; st[2] = com[0]; st[3] = com[1]; st[4] = com[2];
        for (int k = 0; k < 6; ++k) st[5 + k] = bb[k];
    }
    __syncthreads();
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) vpos[d * cap + base + v] += s_delta[d];
    }
}

// ------------------------------------------------------------------------------------------------ V1: voxelization
__device__ __forceinline__ void d_toggle_mark(uint8_t* phase, long long idx) {
    const unsigned long long addr = (unsigned long long)(phase + idx);
    unsigned int* word = (unsigned int*)(addr & ~3ull);
    atomicXor(word, 2u << (8 * (unsigned)(addr & 3ull)));
}
// One thread per (particle, triangle): Moeller-Trumbore against the +z rays through the integer (x,y) of the triangle's
// bounding box (RayTracer.hh:19-39,80-108). A hit at height h toggles the mark bit of cell ceil(h) in column (x,y).
__global__ void __launch_bounds__(128) k_raycast(Geo g, int nv, int nt, const int* __restrict__ tris, long long first_vertex, long long cap,
                                                 int count, int first_particle, const double* __restrict__ vpos, uint8_t* phase, uint8_t* col_flags) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)count * nt) return;
    const int pl = (int)(t / nt), tr = (int)(t % nt);
    if (g.role && g.role[first_particle + pl] == 0) return;
    const long long base = first_vertex + (long long)pl * nv;
    const V3 p0 = ld3(vpos, cap, base + tris[3 * tr]), p1 = ld3(vpos, cap, base + tris[3 * tr + 1]), p2 = ld3(vpos, cap, base + tris[3 * tr + 2]);
    const int iXmin = (int)ceil(fmin(p0.x, fmin(p1.x, p2.x))), iXmax = (int)floor(fmax(p0.x, fmax(p1.x, p2.x)));
    const int iYmin = (int)ceil(fmin(p0.y, fmin(p1.y, p2.y))), iYmax = (int)floor(fmax(p0.y, fmax(p1.y, p2.y)));
    const V3 dir = {0., 0., 1.};
    const V3 e1 = sub(p1, p0), e2 = sub(p2, p0);
    const V3 q = crs(dir, e2);
    const double a = dot(e1, q);
    if (fabs(a) < 2.220446049250313e-16) return;
    const double f = 1 / a;
    const int z1 = g.nz - 1;
    for (int i = iXmin; i <= iXmax; ++i) for (int j = iYmin; j <= iYmax; ++j) {
        const V3 pos = {(double)i, (double)j, 0.};
        const V3 s = sub(pos, p0);
        const double u = f * dot(s, q);
        if (u < 0.0) continue;
        const V3 r = crs(s, e1);
        const double v = f * dot(dir, r);
        if (v < 0.0 || u + v > 1.0) continue;
        double z = f * dot(e2, r);
        if (g.per[2]) { const double L = (double)g.nz, Linv = 1. / L; z = z - L * floor((z - (-0.5)) * Linv); }   // ZBuffer::put, ZBuffer.h:39-50
        const int ci = g.per[0] ? ((i + 100 * g.nx) % g.nx) : i;
        const int cj = g.per[1] ? ((j + 100 * g.ny) % g.ny) : j;
        if (ci < 0 || ci >= g.nx || cj < 0 || cj >= g.ny) continue;
        const int lx = ci - g.x_offset;
        if (lx < 0 || lx >= g.nxl) continue;                               // column belongs to another slab
        const double zc_d = ceil(z);
        uint8_t* flag = col_flags + (long long)(lx + 1) * g.ny + cj;
        unsigned bits = 0;
        if (zc_d > (double)z1) bits = 4;
        else {
            int zc = (int)zc_d; if (zc < 0) zc = 0;
            d_toggle_mark(phase, (long long)(lx + 1) * g.plane + (long long)cj * g.nz + zc);
            if (zc == z1) bits = (z == (double)z1) ? 2 : 1;
        }
        if (bits) {
            const unsigned long long addr = (unsigned long long)flag;
            atomicOr((unsigned int*)(addr & ~3ull), bits << (8 * (unsigned)(addr & 3ull)));
        }
    }
}

// One warp per (x,y) column: phase(z) = parity of the number of hits <= z = prefix XOR of the toggle marks
// (apply_voxelization, ImmersedBoundaryDynamics_voxelization.hh:121-165, incl. its z1 corner case).
__global__ void __launch_bounds__(256) k_phase_scan(int nx, int ny, int nz, long long plane, uint8_t* phase, uint8_t* col_flags) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nx * ny) return;
    const int x = warp / ny, y = warp % ny;
    uint8_t* col = phase + (long long)(x + 1) * plane + (long long)y * nz;
    const int chunk = (nz + 31) / 32;
    const int zb = lane * chunk, ze = min(nz, zb + chunk);
    unsigned par = 0;
    for (int z = zb; z < ze; ++z) par ^= (col[z] >> 1) & 1u;
    unsigned incl = par;
    for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl ^= t; }
    unsigned run = incl ^ par;   // exclusive prefix
    const unsigned flags = col_flags[(long long)(x + 1) * ny + y];
    for (int z = zb; z < ze; ++z) {
        const uint8_t b = col[z];
        run ^= (b >> 1) & 1u;
        unsigned val = run;
        if (z == nz - 1 && !(flags & 4u) && (flags & 3u)) val = (flags & 2u) ? (run ^ 1u) : (b & 1u);
        col[z] = (uint8_t)val;
    }
    __syncwarp();
    if (lane == 0 && flags) col_flags[(long long)(x + 1) * ny + y] = 0;
}

// Word-parallel version of k_phase_scan for nz % (4*W) == 0: a lane handles W consecutive 32-bit words (4 cells each),
// so a warp reads and writes 128*W contiguous bytes per iteration.
template <int W>
__global__ void __launch_bounds__(256) k_phase_scan_words(int nx, int ny, int nz, long long plane, uint8_t* phase, uint8_t* col_flags) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nx * ny) return;
    const int x = warp / ny, y = warp % ny;
    uint32_t* col = (uint32_t*)(phase + (long long)(x + 1) * plane + (long long)y * nz);
    const int nw = nz / 4;
    const unsigned flags = col_flags[(long long)(x + 1) * ny + y];
    unsigned carry = 0;
    for (int w0 = 0; w0 < nw; w0 += 32 * W) {
        const int w = w0 + lane * W;
        uint32_t v[W], p[W];
        if (W == 4) { if (w < nw) { const uint4 q = *(const uint4*)(col + w); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; } else { v[0] = v[1] = v[2] = v[3] = 0; } }
        else { v[0] = (w < nw) ? col[w] : 0u; }
        unsigned lane_par = 0;
#pragma unroll
        for (int k = 0; k < W; ++k) {
            const uint32_t m = (v[k] >> 1) & 0x01010101u;
            uint32_t q = m ^ (m << 8); q ^= (q << 16);           // inclusive prefix parity per byte within the word
            p[k] = q ^ (lane_par ? 0x01010101u : 0u);
            lane_par ^= (q >> 24) & 1u;
        }
        unsigned incl = lane_par;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl ^= t; }
        const unsigned before = incl ^ lane_par ^ carry;
        if (w < nw) {
#pragma unroll
            for (int k = 0; k < W; ++k) {
                uint32_t out = p[k] ^ (before ? 0x01010101u : 0u);
                if (w + k == nw - 1 && !(flags & 4u) && (flags & 3u)) {
                    const unsigned val = (flags & 2u) ? (((out >> 24) & 1u) ^ 1u) : ((v[k] >> 24) & 1u);
                    out = (out & 0x00ffffffu) | (val << 24);
                }
                v[k] = out;
            }
            if (W == 4) *(uint4*)(col + w) = make_uint4(v[0], v[1], v[2], v[3]); else col[w] = v[0];
        }
        carry ^= __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0 && flags) col_flags[(long long)(x + 1) * ny + y] = 0;
}

// ------------------------------------------------------------------------------------------------ band passes
// The reference recomputes the velocity field and resets the force field over the WHOLE lattice every step
// (computeVelocity / setExternalVector, 48 B/cell of extra traffic). Only cells inside some vertex stencil are ever read
// by the interpolation or changed by the spread, so both passes run over the particles' bounding boxes only:
// one CTA per particle, cells of bbox +- margin, periodic wrap per cell. Overlapping boxes write identical values.
struct Band { int lo[3], n[3]; };
__device__ __forceinline__ Band d_band(const double* st, int margin) {
    Band b;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        b.lo[d] = (int)floor(st[5 + 2 * d]) - margin;
        b.n[d] = (int)ceil(st[5 + 2 * d + 1]) + margin - b.lo[d] + 1;
    }
    return b;
}
__device__ __forceinline__ bool d_band_cell(const Geo& g, const Band& b, int t, long long& idx) {
    const int kz = t % b.n[2], ky = (t / b.n[2]) % b.n[1], kx = t / (b.n[2] * b.n[1]);
    int c[3] = {b.lo[0] + kx, b.lo[1] + ky, b.lo[2] + kz};
    const int n[3] = {g.nx, g.ny, g.nz};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (c[d] < 0 || c[d] >= n[d]) {
            if (!g.per[d]) return false;
            c[d] = ((c[d] % n[d]) + n[d]) % n[d];
        }
    }
    const int lx = c[0] - g.x_offset;
    if (lx < 0 || lx >= g.nxl) return false;
    idx = (long long)(lx + 1) * g.plane + (long long)c[1] * g.nz + c[2];
    return true;
}
// Cells actually touched by the particle's vertex stencils are marked in a shared-memory bitmask over the band box
// (a deflated RBC touches ~1.6k of the ~6k cells of its box); only those get a velocity, and the mask is kept (double
// buffered) so that the NEXT step resets exactly the cells this step's spread wrote.
__global__ void __launch_bounds__(256) k_band_velocity(Geo g, const double* __restrict__ f, const double* __restrict__ force, double* vel,
                                                       const double* __restrict__ pstate, const double* __restrict__ vpos, long long cap_v,
                                                       int nv, long long first_vertex, int first_particle, unsigned* mask_store, int* band_store) {
    __shared__ unsigned s_mask[BAND_WORDS];
    const int p = first_particle + blockIdx.x;
    int* bs = band_store + (long long)p * 6;
    if (g.role && g.role[p] == 0) { if (threadIdx.x == 0) bs[3] = 0; return; }
    const Band b = d_band(pstate + (long long)p * 11, 2);
    const int total = b.n[0] * b.n[1] * b.n[2];
    const bool masked = total <= BAND_WORDS * 32;
    if (masked) {
        for (int w = threadIdx.x; w < BAND_WORDS; w += blockDim.x) s_mask[w] = 0u;
        __syncthreads();
        const long long base = first_vertex + (long long)blockIdx.x * nv;
        for (int v = threadIdx.x; v < nv; v += blockDim.x) {
            // stencil origin in global (unwrapped) cell coordinates: local index ceil(p + 1 - 1.5) minus the envelope width
            const int rx = (int)ceil((vpos[base + v] - (-1.0)) - 1.5) - 1 - b.lo[0];
            const int ry = (int)ceil((vpos[cap_v + base + v] - (-1.0)) - 1.5) - 1 - b.lo[1];
            const int rz = (int)ceil((vpos[2 * cap_v + base + v] - (-1.0)) - 1.5) - 1 - b.lo[2];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
                const int row = ((rx + i) * b.n[1] + (ry + j)) * b.n[2] + rz;
#pragma unroll
                for (int k = 0; k < 3; ++k) atomicOr(&s_mask[(row + k) >> 5], 1u << ((row + k) & 31));
            }
        }
        __syncthreads();
        for (int w = threadIdx.x; w < BAND_WORDS; w += blockDim.x) mask_store[(long long)p * BAND_WORDS + w] = s_mask[w];
    }
    if (threadIdx.x == 0) { bs[0] = b.lo[0]; bs[1] = b.lo[1]; bs[2] = b.lo[2]; bs[3] = masked ? b.n[0] : -b.n[0]; bs[4] = b.n[1]; bs[5] = b.n[2]; }
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        if (masked && !((s_mask[t >> 5] >> (t & 31)) & 1u)) continue;
        long long idx;
        if (!d_band_cell(g, b, t, idx)) continue;
        double u[3];
        if (g.kind && g.kind[idx] != 0) u[0] = u[1] = u[2] = 0.0;       // NoDynamics / BounceBack: computeVelocity gives 0
        else d_cell_velocity(f, force, g.vol, idx, u);
        vel[idx] = u[0]; vel[g.vol + idx] = u[1]; vel[2 * g.vol + idx] = u[2];
    }
}
// Fused band velocity + interpolation (one CTA per particle): the velocities of the cells the particle's stencils touch are evaluated
// ONCE from the populations into a shared-memory tile (compact: one slot per touched cell, addressed through the touched-cell bit mask and
// a per-word prefix count) and every vertex interpolates from the tile. Compared with k_band_velocity + k_interpolate this removes the
// 24 B/cell round trip through a dense velocity field (and the field itself: 3.3 GB on the bench slab) and the 81 L2 gathers per vertex.
// Same arithmetic in the same order as the reference (_interpolation.hh:43-153): bit-identical to the two-kernel path.
// Wall-plane cells report their stored boundary velocity (StoreVelocityDynamics), NoDynamics / BounceBack cells zero.
#define TILE_SLOTS 3072
#ifndef BIB_INTERP_MINB
#define BIB_INTERP_MINB 2
#endif
__global__ void __launch_bounds__(256, BIB_INTERP_MINB) k_interp_tile(Geo g, WallsDev W, const double* __restrict__ f, const double* __restrict__ force,
                                                     const double* __restrict__ pstate, const double* __restrict__ vpos, double* __restrict__ vvel,
                                                     long long cap_v, int nv, long long first_vertex, int first_particle, unsigned* mask_store, int* band_store,
                                                     unsigned* fflag, int compact) {
    __shared__ unsigned s_mask[BAND_WORDS];
    __shared__ int s_pre[BAND_WORDS];
    __shared__ int s_wsum[8];
    __shared__ int s_total;
    extern __shared__ double s_u[];                   // [3][TILE_SLOTS]
    const int p = first_particle + blockIdx.x;
    int* bs = band_store + (long long)p * 6;
    if (g.role && g.role[p] == 0) { if (threadIdx.x == 0) bs[3] = 0; return; }
    const Band b = d_band(pstate + (long long)p * 11, 2);
    const long long total_ll = (long long)b.n[0] * b.n[1] * b.n[2];
    const bool masked = total_ll > 0 && total_ll <= BAND_WORDS * 32;
    const int total = masked ? (int)total_ll : 0;
    const long long base = first_vertex + (long long)blockIdx.x * nv;
    if (threadIdx.x == 0) { bs[0] = b.lo[0]; bs[1] = b.lo[1]; bs[2] = b.lo[2]; bs[3] = masked ? b.n[0] : -b.n[0]; bs[4] = b.n[1]; bs[5] = b.n[2]; }
    bool tiled = masked;
    if (masked) {
        for (int w = threadIdx.x; w < BAND_WORDS; w += blockDim.x) s_mask[w] = 0u;
        __syncthreads();
        for (int v = threadIdx.x; v < nv; v += blockDim.x) {
            const int rx = (int)ceil((vpos[base + v] - (-1.0)) - 1.5) - 1 - b.lo[0];
            const int ry = (int)ceil((vpos[cap_v + base + v] - (-1.0)) - 1.5) - 1 - b.lo[1];
            const int rz = (int)ceil((vpos[2 * cap_v + base + v] - (-1.0)) - 1.5) - 1 - b.lo[2];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
                const int row = ((rx + i) * b.n[1] + (ry + j)) * b.n[2] + rz;
#pragma unroll
                for (int k = 0; k < 3; ++k) atomicOr(&s_mask[(row + k) >> 5], 1u << ((row + k) & 31));
            }
        }
        __syncthreads();
        // exclusive prefix of the per-word population counts (BAND_WORDS = 2 * blockDim.x words)
        const int w0 = 2 * threadIdx.x;
        const unsigned m0 = s_mask[w0], m1 = s_mask[w0 + 1];
        mask_store[(long long)p * BAND_WORDS + w0] = m0; mask_store[(long long)p * BAND_WORDS + w0 + 1] = m1;
        const int c0 = __popc(m0), c1 = __popc(m1);
        int incl = c0 + c1;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_wsum[wid] = incl;
        __syncthreads();
        int before = 0;
        for (int k = 0; k < wid; ++k) before += s_wsum[k];
        s_pre[w0] = before + incl - c0 - c1; s_pre[w0 + 1] = before + incl - c1;
        if (threadIdx.x == blockDim.x - 1) s_total = before + incl;
        __syncthreads();
        tiled = s_total <= TILE_SLOTS;
    }
    if (tiled) {
        // velocities of the touched cells. compact == 0: the threads walk the cells of the box and skip the untouched ones (40 - 60 % of the
        // lanes idle while the others wait for their 22 loads); compact != 0: the threads walk the SLOTS, i.e. the touched cells only
        // (slot -> cell: binary search of the word in the prefix counts, then the (slot - prefix)-th set bit of its mask) - every lane has a
        // cell, more loads in flight per warp, fewer trips. Same cells, same arithmetic per cell.
        const int n_iter = compact ? s_total : total;
        for (int it = threadIdx.x; it < n_iter; it += blockDim.x) {
            int t, slot;
            if (compact) {
                slot = it;
                int lo = 0, hi = BAND_WORDS - 1;                       // last word w with s_pre[w] <= slot (it holds the slot: empty words repeat the prefix of their successor)
                while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_pre[mid] <= slot) lo = mid; else hi = mid - 1; }
                t = (lo << 5) + (int)__fns(s_mask[lo], 0u, slot - s_pre[lo] + 1);
            } else {
                t = it;
                const unsigned mw0 = s_mask[t >> 5];
                if (!((mw0 >> (t & 31)) & 1u)) continue;
                slot = s_pre[t >> 5] + __popc(mw0 & ((1u << (t & 31)) - 1u));
            }
            double u[3] = {0.0, 0.0, 0.0};
            long long idx;
            if (d_band_cell(g, b, t, idx)) {
                // the touched cells are exactly the cells the spread of these positions will write: raise their "force differs from the body
                // force" flags here, once per cell, instead of once per stencil point in the spread (27 atomics per vertex less)
                if (fflag) atomicOr(&fflag[idx >> 5], 1u << (idx & 31));
                bool wall = false;
                if (W.n > 0) {
                    const int kz = t % b.n[2], ky = (t / b.n[2]) % b.n[1], kx = t / (b.n[2] * b.n[1]);
                    int c[3] = {b.lo[0] + kx, b.lo[1] + ky, b.lo[2] + kz};
                    const int nn[3] = {g.nx, g.ny, g.nz};
#pragma unroll
                    for (int d = 0; d < 3; ++d) if (g.per[d]) c[d] = ((c[d] % nn[d]) + nn[d]) % nn[d];
                    wall = d_wall_velocity(W, c[0] - g.x_offset, c[1], c[2], u);          // wall arrays are slab-local along x
                }
                if (!wall && !(g.kind && g.kind[idx] != 0)) d_cell_velocity(f, force, g.vol, idx, u);
            }
            s_u[slot] = u[0]; s_u[TILE_SLOTS + slot] = u[1]; s_u[2 * TILE_SLOTS + slot] = u[2];
        }
        __syncthreads();
        SrcTile src;
        src.mask = s_mask; src.pre = s_pre; src.u = s_u; src.slots = TILE_SLOTS;
        src.lo[0] = b.lo[0]; src.lo[1] = b.lo[1]; src.lo[2] = b.lo[2]; src.n[0] = b.n[0]; src.n[1] = b.n[1]; src.n[2] = b.n[2];
        for (int v = threadIdx.x; v < nv; v += blockDim.x) {
            const double pp[3] = {vpos[base + v], vpos[cap_v + base + v], vpos[2 * cap_v + base + v]};
            double o[3];
            d_interp_vertex(g, src, pp, o);
            vvel[base + v] = o[0]; vvel[cap_v + base + v] = o[1]; vvel[2 * cap_v + base + v] = o[2];
        }
    } else {
        // box or touched-cell count beyond the tile: straight from the populations, cell by cell (only without wall planes / cell kinds, see the launcher)
        const SrcPop src = {f, force, g.vol};
        for (int v = threadIdx.x; v < nv; v += blockDim.x) {
            const double pp[3] = {vpos[base + v], vpos[cap_v + base + v], vpos[2 * cap_v + base + v]};
            double o[3];
            if (fflag && d_classify(g, pp) != 0) {          // force flags of this vertex's stencil cells (the tile path sets them cell by cell)
                const int x0 = (int)ceil((pp[0] - (-1.0)) - 1.5), y0 = (int)ceil((pp[1] - (-1.0)) - 1.5), z0 = (int)ceil((pp[2] - (-1.0)) - 1.5);
                for (int i = x0; i <= x0 + 2; ++i) for (int j = y0; j <= y0 + 2; ++j) for (int k = z0; k <= z0 + 2; ++k) {
                    const int i2 = d_remap(g, 0, i), j2 = d_remap(g, 1, j), k2 = d_remap(g, 2, k);
                    if (i2 < 0 || i2 > g.nx - 1 || !d_in_slab(g, i2) || j2 < 0 || j2 > g.ny - 1 || k2 < 0 || k2 > g.nz - 1) continue;
                    const long long c0 = d_cell(g, i2, j2, k2);
                    atomicOr(&fflag[c0 >> 5], 1u << (c0 & 31));
                }
            }
            d_interp_vertex<SrcPop, true>(g, src, pp, o);
            vvel[base + v] = o[0]; vvel[cap_v + base + v] = o[1]; vvel[2 * cap_v + base + v] = o[2];
        }
    }
}
// F := body on the cells the previous step's spread touched (stored mask), or on the whole box +-3 when no mask exists
__global__ void __launch_bounds__(256) k_band_reset(Geo g, double* force, const double* __restrict__ pstate, double fx, double fy, double fz,
                                                    const unsigned* __restrict__ mask_store, const int* __restrict__ band_store) {
    const int p = blockIdx.x;
    Band b;
    bool masked = false;
    if (band_store) {
        const int* bs = band_store + (long long)p * 6;
        if (bs[3] == 0) return;                       // slot was dead when the mask was taken
        b.lo[0] = bs[0]; b.lo[1] = bs[1]; b.lo[2] = bs[2]; b.n[0] = abs(bs[3]); b.n[1] = bs[4]; b.n[2] = bs[5];
        masked = bs[3] > 0;
    } else {
        if (g.role && g.role[p] == 0) return;
        b = d_band(pstate + (long long)p * 11, 3);
    }
    const unsigned* m = mask_store + (long long)p * BAND_WORDS;
    const int total = b.n[0] * b.n[1] * b.n[2];
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        if (masked && !((m[t >> 5] >> (t & 31)) & 1u)) continue;
        long long idx;
        if (!d_band_cell(g, b, t, idx)) continue;
        force[idx] = fx; force[g.vol + idx] = fy; force[2 * g.vol + idx] = fz;
    }
}

// Production spread (north-star: shared-memory tile staging): one CTA per particle. The particle's vertex forces are first
// accumulated into a shared-memory tile covering its bounding box + 2 cells (the box k_band_velocity uses; a deflated RBC
// touches ~1.6k cells, every one of them ~30 times), then each touched cell is added to the lattice force field ONCE with a
// global fp64 reduction: 81 global atomics per vertex become ~8.
// Measured alternatives (782k vertices, B200; profiles/r1_ib_kernels.md): global fp64 atomics only 0.45 ms; this kernel (fp64
// shared-memory atomicAdd = ATOMS.CAST.SPIN loops, sm_100 has no native 64-bit shared atomic) 0.38 ms and the best whole-IB time;
// fixed-point limbs on native 32-bit ATOMS.ADD and an atomic-free binned gather were both slower end to end (0.38-0.44 ms).
// Cells are addressed unwrapped inside the tile; periodic wrap, walls and slab ownership are applied at flush time
// (d_band_cell), which is exactly where the reference's general kernel drops or wraps them. Boxes larger than the tile, or
// stencil points outside it, fall back to direct global atomics.
#define TILE_CELLS 3072
__device__ __forceinline__ void d_tile_pass(const Geo& g, const Band& b, int total, double* s_tile, double* force, const double* p, bool th, const double* F) {
    const int sup = th ? 2 : 3;
    const double hs = th ? 1.0 : 1.5;
    const int x0 = (int)ceil(p[0] - hs), y0 = (int)ceil(p[1] - hs), z0 = (int)ceil(p[2] - hs);
    for (int i = x0; i <= x0 + sup - 1; ++i) {
        const double wx = th ? d_tophat((double)i - p[0]) : d_roma((double)i - p[0]);
        const int tx = i - 1 - b.lo[0];
        for (int j = y0; j <= y0 + sup - 1; ++j) {
            const double wxy = wx * (th ? d_tophat((double)j - p[1]) : d_roma((double)j - p[1]));
            const int ty = j - 1 - b.lo[1];
            for (int k = z0; k <= z0 + sup - 1; ++k) {
                const double w = wxy * (th ? d_tophat((double)k - p[2]) : d_roma((double)k - p[2]));
                const int tz = k - 1 - b.lo[2];
                if (tx >= 0 && tx < b.n[0] && ty >= 0 && ty < b.n[1] && tz >= 0 && tz < b.n[2]) {
                    const int t = (tx * b.n[1] + ty) * b.n[2] + tz;
#pragma unroll
                    for (int c = 0; c < 3; ++c) atomicAdd(&s_tile[c * total + t], w * F[c]);
                } else {
                    const int i2 = d_remap(g, 0, i), j2 = d_remap(g, 1, j), k2 = d_remap(g, 2, k);
                    if (i2 < 0 || i2 > g.nx - 1 || !d_in_slab(g, i2) || j2 < 0 || j2 > g.ny - 1 || k2 < 0 || k2 > g.nz - 1) continue;
                    const long long c0 = d_cell(g, i2, j2, k2);
#pragma unroll
                    for (int c = 0; c < 3; ++c) atomicAdd(force + c * g.vol + c0, w * F[c]);
                    if (g.fflag) atomicOr(&g.fflag[c0 >> 5], 1u << (c0 & 31));
                }
            }
        }
    }
}
__global__ void __launch_bounds__(256) k_spread_tile(Geo g, double* force, const double* __restrict__ pstate, const double* __restrict__ vpos,
                                                     const double* __restrict__ vfrc, long long cap, int nv, long long first_vertex, int first_particle,
                                                     const int* __restrict__ perm) {
    extern __shared__ double s_tile[];
    const int p = first_particle + blockIdx.x;
    if (g.role && g.role[p] == 0) return;
    const Band b = d_band(pstate + (long long)p * 11, 2);
    const long long total_ll = (long long)b.n[0] * b.n[1] * b.n[2];
    const bool tiled = total_ll > 0 && total_ll <= TILE_CELLS;
    const int total = tiled ? (int)total_ll : 0;
    for (int t = threadIdx.x; t < 3 * total; t += blockDim.x) s_tile[t] = 0.0;
    __syncthreads();
    const long long base = first_vertex + (long long)blockIdx.x * nv;
    for (int kk = threadIdx.x; kk < nv; kk += blockDim.x) {
        // lanes of a warp take vertices that are NEIGHBOURS ON THE SURFACE (perm: the shape's vertices along a space-filling curve), so that the
        // 32 reductions of one instruction fall into few 32-byte sectors of the force field
        const int v = perm ? perm[kk] : kk;
        const double pp[3] = {vpos[base + v], vpos[cap + base + v], vpos[2 * cap + base + v]};
        const double F[3] = {vfrc[base + v], vfrc[cap + base + v], vfrc[2 * cap + base + v]};
        const int cls = d_classify(g, pp);
        if (cls == 0) continue;
        if (!tiled) { d_spread_vertex(g, force, pp, F, cls, true); continue; }
        const double pr[3] = {pp[0] - (-1.0), pp[1] - (-1.0), pp[2] - (-1.0)};
        if (cls == 3) {
            const double weight = d_plates_distance(g, pp) - 1.5;
            if (weight >= 1) d_tile_pass(g, b, total, s_tile, force, pr, false, F);
            else if (weight <= 0) d_tile_pass(g, b, total, s_tile, force, pr, true, F);
            else {
                const double Fw[3] = {F[0] * weight, F[1] * weight, F[2] * weight};
                const double Ft[3] = {F[0] * (1. - weight), F[1] * (1. - weight), F[2] * (1. - weight)};
                d_tile_pass(g, b, total, s_tile, force, pr, false, Fw);
                d_tile_pass(g, b, total, s_tile, force, pr, true, Ft);
            }
        } else d_tile_pass(g, b, total, s_tile, force, pr, false, F);
    }
    __syncthreads();
    // flush: one global reduction per touched cell; the force flags of cells sharing a 32-cell word are OR-ed inside the warp first
    for (int t0 = 0; t0 < total; t0 += blockDim.x) {
        const int t = t0 + threadIdx.x;
        long long idx = -1;
        if (t < total) {
            const double fx = s_tile[t], fy = s_tile[total + t], fz = s_tile[2 * total + t];
            long long i2;
            if (!(fx == 0.0 && fy == 0.0 && fz == 0.0) && d_band_cell(g, b, t, i2)) {
                idx = i2;
                atomicAdd(force + idx, fx); atomicAdd(force + g.vol + idx, fy); atomicAdd(force + 2 * g.vol + idx, fz);
            }
        }
        if (g.fflag) {
            const long long word = idx >= 0 ? (idx >> 5) : -1 - (long long)(threadIdx.x & 31);      // invalid lanes get unique keys
            const unsigned grp = __match_any_sync(0xffffffffu, word);
            const unsigned bits = __reduce_or_sync(grp, idx >= 0 ? (1u << (idx & 31)) : 0u);
            if (idx >= 0 && (threadIdx.x & 31) == (unsigned)(__ffs(grp) - 1)) atomicOr(&g.fflag[word], bits);
        }
    }
}

Geo make_geo(const bib_domain* d) {
    Geo g;
    g.nx = d->gnx; g.ny = d->ny; g.nz = d->nz;
    g.per[0] = d->per[0]; g.per[1] = d->per[1]; g.per[2] = d->per[2];
    g.plane = (long long)d->plane; g.vol = (long long)d->vol; g.x_offset = d->x_offset;
    g.nxl = d->nx; g.role = d->ms ? d->role_all : nullptr; g.vslot = d->vslot;
    g.plates = d->plates; g.py0 = d->plate_y0; g.py1 = d->plate_y1; g.prad = d->pipe_radius; g.kind = d->kind;
    g.fflag = (d->sparse_force && d->fflag_valid && !d->deterministic_spread) ? d->fflag : nullptr;
    return g;
}

template <class T>
bool to_device(T** dst, const std::vector<T>& src) {
    if (!bib_cuda_ok(cudaMalloc((void**)dst, std::max<size_t>(1, src.size()) * sizeof(T)), "cudaMalloc(shape)")) return false;
    if (!src.empty() && !bib_cuda_ok(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice), "cudaMemcpy(shape)")) return false;
    return true;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host: shapes
// ParticleShape constructor + compute_connectivity (src/fsi/ParticleShape.hh:22-57,90-184): links in std::set<(i0,i1)>
// order, t1 = first triangle naming the edge, t2 = last other triangle; one pair per interior link, in link order.
int ib_shape_register(bib_domain* d, int nv, int nt, const double* verts, const int* tris, int* shape_id) {
    if (nv <= 0 || nt <= 0 || !verts || !tris) { bib_set_error("shape_register: empty mesh"); return 0; }
    for (int i = 0; i < 3 * nt; ++i) if (tris[i] < 0 || tris[i] >= nv) { bib_set_error("shape_register: triangle index out of range"); return 0; }
    ShapeDev s;
    s.nv = nv; s.nt = nt;
    s.h_verts.assign(verts, verts + 3 * (size_t)nv);
    s.h_tris.assign(tris, tris + 3 * (size_t)nt);
    auto V = [&](int i) { return verts + 3 * (size_t)i; };
    auto cross3 = [](const double* a, const double* b, double* r) { r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0]; };
    auto norm3 = [](const double* a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); };
    s.h_tri_area.resize(nt);
    double area = 0, volume = 0;
    for (int t = 0; t < nt; ++t) {
        const double *v0 = V(tris[3 * t]), *v1 = V(tris[3 * t + 1]), *v2 = V(tris[3 * t + 2]);
        double e1[3], e2[3], n[3], c[3];
        for (int k = 0; k < 3; ++k) { e1[k] = v1[k] - v0[k]; e2[k] = v2[k] - v0[k]; }
        cross3(e1, e2, n);
        double a = norm3(n);
        a *= 0.5;
        s.h_tri_area[t] = a;
        area += a;
        cross3(v1, v2, c);
        volume += (v0[0] * c[0] + v0[1] * c[1] + v0[2] * c[2]) / 6.;
    }
    s.area = area; s.volume = volume;
    {   // ParticleShape centre (src/fsi/ParticleShape.hh:36-50) and the rest length of the semi-rigid model's radial springs
        double center[3] = {0, 0, 0};
        for (int t = 0; t < nt; ++t) {
            const double *v0 = V(tris[3 * t]), *v1 = V(tris[3 * t + 1]), *v2 = V(tris[3 * t + 2]);
            for (int k = 0; k < 3; ++k) center[k] += ((v0[k] + v1[k] + v2[k]) / 3.0) * s.h_tri_area[t];
        }
        for (int k = 0; k < 3; ++k) center[k] /= area;
        std::vector<double> rr(nv);
        for (int i = 0; i < nv; ++i) {
            const double r0[3] = {V(i)[0] - center[0], V(i)[1] - center[1], V(i)[2] - center[2]};
            rr[i] = std::sqrt(r0[0] * r0[0] + r0[1] * r0[1] + r0[2] * r0[2]);
        }
        if (!to_device(&s.rest_radial, rr)) return 0;
    }
    struct Key { int i0, i1, order; };
    std::vector<Key> keys(3 * (size_t)nt);
    for (int t = 0; t < nt; ++t) for (int j = 0; j < 3; ++j) {
        const int a = tris[3 * t + j], b = tris[3 * t + (j + 1) % 3];
        keys[3 * t + j] = {std::min(a, b), std::max(a, b), 3 * t + j};
    }
    std::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) { return a.i0 != b.i0 ? a.i0 < b.i0 : (a.i1 != b.i1 ? a.i1 < b.i1 : a.order < b.order); });
    std::vector<int> link_tri;
    for (size_t i = 0; i < keys.size();) {
        size_t j = i;
        const int t1 = keys[i].order / 3; int t2 = -1;
        while (j < keys.size() && keys[j].i0 == keys[i].i0 && keys[j].i1 == keys[i].i1) { const int t = keys[j].order / 3; if (t != t1) t2 = t; ++j; }
        s.h_links.push_back(keys[i].i0); s.h_links.push_back(keys[i].i1);
        link_tri.push_back(t1); link_tri.push_back(t2);
        const double* a = V(keys[i].i0); const double* b = V(keys[i].i1);
        const double dd[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
        s.h_link_len.push_back(norm3(dd));
        i = j;
    }
    s.nl = (int)s.h_link_len.size();
    auto unit_normal = [&](const double* v0, const double* v1, const double* v2, double* n) {
        double e1[3], e2[3];
        for (int k = 0; k < 3; ++k) { e1[k] = v1[k] - v0[k]; e2[k] = v2[k] - v0[k]; }
        cross3(e1, e2, n);
        const double nn = norm3(n);
        n[0] /= nn; n[1] /= nn; n[2] /= nn;
    };
    for (int l = 0; l < s.nl; ++l) {
        if (link_tri[2 * l] == -1 || link_tri[2 * l + 1] == -1) continue;
        int a[3] = {tris[3 * link_tri[2 * l]], tris[3 * link_tri[2 * l] + 1], tris[3 * link_tri[2 * l] + 2]};
        int b[3] = {tris[3 * link_tri[2 * l + 1]], tris[3 * link_tri[2 * l + 1] + 1], tris[3 * link_tri[2 * l + 1] + 2]};
        const int i0 = s.h_links[2 * l], i1 = s.h_links[2 * l + 1];
        while (!((a[0] == i0 && a[2] == i1) || (a[0] == i1 && a[2] == i0))) { const int tmp = a[0]; a[0] = a[2]; a[2] = a[1]; a[1] = tmp; }
        while (!((b[0] == i0 && b[1] == i1) || (b[0] == i1 && b[1] == i0))) { const int tmp = b[0]; b[0] = b[2]; b[2] = b[1]; b[1] = tmp; }
        const int q[4] = {a[0], a[1], a[2], b[2]};
        for (int k = 0; k < 4; ++k) s.h_pairs.push_back(q[k]);
        // tri::cos_sin_of_angle_between_pair on the REST mesh (TriangleUtils.h:124-136)
        double n1[3], n2[3], dn[3], dv[3], cr[3];
        unit_normal(V(q[0]), V(q[1]), V(q[2]), n1);
        unit_normal(V(q[0]), V(q[2]), V(q[3]), n2);
        const double cs = n1[0] * n2[0] + n1[1] * n2[1] + n1[2] * n2[2];
        for (int k = 0; k < 3; ++k) { dn[k] = n1[k] - n2[k]; dv[k] = V(q[1])[k] - V(q[3])[k]; }
        cross3(n1, n2, cr);
        const double sn = ((dn[0] * dv[0] + dn[1] * dv[1] + dn[2] * dv[2] > 0) ? 1. : -1.) * norm3(cr);
        s.h_pair_cs.push_back(cs); s.h_pair_cs.push_back(sn);
    }
    s.np = (int)s.h_pairs.size() / 4;
    // per-vertex incidence lists, element index ascending
    auto build_csr = [&](int n_elem, int arity, const std::vector<int>& elems, std::vector<int>& off, std::vector<int>& lst, int mul) {
        off.assign(nv + 1, 0);
        for (int e = 0; e < n_elem; ++e) for (int k = 0; k < arity; ++k) off[elems[arity * e + k] + 1]++;
        for (int v = 0; v < nv; ++v) off[v + 1] += off[v];
        lst.assign(off[nv], 0);
        std::vector<int> cur(off.begin(), off.end() - 1);
        for (int e = 0; e < n_elem; ++e) for (int k = 0; k < arity; ++k) lst[cur[elems[arity * e + k]]++] = e * mul + k;
    };
    std::vector<int> lo, ll, to, tl, po, pl;
    build_csr(s.nl, 2, s.h_links, lo, ll, 2);
    build_csr(s.nt, 3, s.h_tris, to, tl, 3);
    build_csr(s.np, 4, s.h_pairs, po, pl, 4);
    BIB_CHECK(cudaSetDevice(d->device));
    if (!to_device(&s.tris, s.h_tris) || !to_device(&s.tri_area, s.h_tri_area) || !to_device(&s.links, s.h_links) ||
        !to_device(&s.link_len, s.h_link_len) || !to_device(&s.pairs, s.h_pairs) || !to_device(&s.pair_cs, s.h_pair_cs) ||
        !to_device(&s.v_link_off, lo) || !to_device(&s.v_link, ll) || !to_device(&s.v_tri_off, to) || !to_device(&s.v_tri, tl) ||
        !to_device(&s.v_pair_off, po) || !to_device(&s.v_pair, pl) || !to_device(&s.verts, s.h_verts)) return 0;
    {   // vertices along a Morton curve through the rest shape: consecutive entries are neighbours on the surface (k_spread_tile)
        double lo[3] = {verts[0], verts[1], verts[2]}, hi[3] = {verts[0], verts[1], verts[2]};
        for (int v = 0; v < nv; ++v) for (int c = 0; c < 3; ++c) { lo[c] = std::min(lo[c], verts[3 * v + c]); hi[c] = std::max(hi[c], verts[3 * v + c]); }
        std::vector<std::pair<unsigned long long, int>> key(nv);
        for (int v = 0; v < nv; ++v) {
            unsigned long long m = 0;
            unsigned q[3];
            for (int c = 0; c < 3; ++c) q[c] = (unsigned)(1023.0 * (verts[3 * v + c] - lo[c]) / std::max(hi[c] - lo[c], 1e-300));
            for (int b = 9; b >= 0; --b) for (int c = 0; c < 3; ++c) m = (m << 1) | ((q[c] >> b) & 1u);
            key[v] = std::make_pair(m, v);
        }
        std::sort(key.begin(), key.end());
        std::vector<int> perm(nv);
        for (int v = 0; v < nv; ++v) perm[v] = key[v].second;
        if (!to_device(&s.spread_perm, perm)) return 0;
    }
    d->shapes.push_back(s);
    *shape_id = (int)d->shapes.size() - 1;
    return 1;
}

// RBCParameters::in_plane_ks_kp (src/fsi/RBCParticle.hh:27-44), evaluated on the host with libm pow like the reference
static void rbc_ks_kp(double l0, double G, double& ks, double& kp) {
    const double x0 = 1. / 2.2, m = 2;
    const double ks_coeff1 = 0.25 * (6. * x0 + (3 * x0 * x0 - 2 * x0 * x0 * x0) / ((1 - x0) * (1 - x0)));
    const double ks_coeff2 = x0 / (2. * (1 - x0) * (1 - x0) * (1 - x0)) + 0.25 * (1. - 1. / ((1 - x0) * (1 - x0)));
    // a genuine libm pow() call, as in the reference (m is a run-time argument there): glibc's pow(x, 2.0) differs from x*x
    // by one ulp for a few rest lengths, so the exponent is hidden from the compiler's pow(x,2) -> x*x folding
    static volatile double m_runtime = 2.0;
    const double p_l0 = std::pow(l0, (double)m_runtime);
    const double kp_coeff1 = -1. / p_l0;
    const double kp_coeff2 = (m + 1) / p_l0;
    const double G_lhs = 4. * l0 * G / std::sqrt(3);
    ks = -kp_coeff1 * G_lhs / (ks_coeff1 * kp_coeff2 - ks_coeff2 * kp_coeff1);
    kp = ks_coeff1 * G_lhs / (ks_coeff1 * kp_coeff2 - ks_coeff2 * kp_coeff1);
}

static int grow_particles(bib_domain* d, long long need_v, int need_p) {
    if (need_v > d->cap_vertices) {
        long long cap = std::max<long long>(need_v, d->cap_vertices * 2);
        double* np[3];
        double** old[3] = {&d->vpos, &d->vvel, &d->vfrc};
        for (int k = 0; k < 3; ++k) {
            BIB_CHECK(cudaMalloc(&np[k], (size_t)cap * 3 * sizeof(double)));
            BIB_CHECK(cudaMemsetAsync(np[k], 0, (size_t)cap * 3 * sizeof(double), d->stream));
            if (*old[k]) {
                for (int c = 0; c < 3; ++c)
                    BIB_CHECK(cudaMemcpyAsync(np[k] + c * cap, *old[k] + c * d->cap_vertices, (size_t)d->n_vertices * sizeof(double), cudaMemcpyDeviceToDevice, d->stream));
            }
        }
        BIB_CHECK(cudaStreamSynchronize(d->stream));
        for (int k = 0; k < 3; ++k) { if (*old[k]) cudaFree(*old[k]); *old[k] = np[k]; }
        if (d->n_ranks > 1) {
            int* nvs;
            BIB_CHECK(cudaMalloc(&nvs, (size_t)cap * sizeof(int)));
            BIB_CHECK(cudaMemset(nvs, 0, (size_t)cap * sizeof(int)));
            if (d->vslot) { BIB_CHECK(cudaMemcpy(nvs, d->vslot, (size_t)d->n_vertices * sizeof(int), cudaMemcpyDeviceToDevice)); cudaFree(d->vslot); }
            d->vslot = nvs;
        }
        d->bytes += (cap - d->cap_vertices) * 9 * 8;
        d->cap_vertices = cap;
    }
    if (need_p > d->cap_particles) {
        int cap = std::max(need_p, d->cap_particles * 2);
        double* ns;
        BIB_CHECK(cudaMalloc(&ns, (size_t)cap * 11 * sizeof(double)));
        BIB_CHECK(cudaMemsetAsync(ns, 0, (size_t)cap * 11 * sizeof(double), d->stream));
        if (d->pstate) BIB_CHECK(cudaMemcpyAsync(ns, d->pstate, (size_t)d->n_particles * 11 * sizeof(double), cudaMemcpyDeviceToDevice, d->stream));
        BIB_CHECK(cudaStreamSynchronize(d->stream));
        if (d->pstate) cudaFree(d->pstate);
        {                              // slot tables (one slot per resident particle; on a single slab slot == particle, all owned)
            int *nr, *ng, *nn; long long* nf; uint8_t* nsn;
            BIB_CHECK(cudaMalloc(&nr, (size_t)cap * sizeof(int))); BIB_CHECK(cudaMalloc(&ng, (size_t)cap * sizeof(int))); BIB_CHECK(cudaMalloc(&nn, (size_t)cap * sizeof(int)));
            BIB_CHECK(cudaMalloc(&nf, (size_t)cap * sizeof(long long))); BIB_CHECK(cudaMalloc(&nsn, 2 * (size_t)cap));
            BIB_CHECK(cudaMemset(nr, 0, (size_t)cap * sizeof(int))); BIB_CHECK(cudaMemset(ng, 0xff, (size_t)cap * sizeof(int))); BIB_CHECK(cudaMemset(nn, 0, (size_t)cap * sizeof(int)));
            BIB_CHECK(cudaMemset(nf, 0, (size_t)cap * sizeof(long long))); BIB_CHECK(cudaMemset(nsn, 0, 2 * (size_t)cap));
            if (d->role_all) {
                const size_t n = (size_t)d->n_particles;
                BIB_CHECK(cudaMemcpy(nr, d->role_all, n * sizeof(int), cudaMemcpyDeviceToDevice)); BIB_CHECK(cudaMemcpy(ng, d->gid_all, n * sizeof(int), cudaMemcpyDeviceToDevice));
                BIB_CHECK(cudaMemcpy(nn, d->slot_nv, n * sizeof(int), cudaMemcpyDeviceToDevice)); BIB_CHECK(cudaMemcpy(nf, d->slot_fv, n * sizeof(long long), cudaMemcpyDeviceToDevice));
                cudaFree(d->role_all); cudaFree(d->gid_all); cudaFree(d->slot_nv); cudaFree(d->slot_fv); cudaFree(d->seen_all);
            }
            d->role_all = nr; d->gid_all = ng; d->slot_nv = nn; d->slot_fv = nf; d->seen_all = nsn;
        }
        d->pstate = ns; d->cap_particles = cap;
    }
    return 1;
}

static int launch_update(bib_domain* d, const ParticleGroup& gr, int do_move, const uint8_t* skip_mask = nullptr) {
    const ShapeDev& s = d->shapes[gr.shape];
    if (do_move) d->band_pos_valid = false;      // the band masks were taken at the positions before this move
    if (gr.type == 3) {
        RigidParams c;
        c.nv = s.nv; c.nt = s.nt; c.first_vertex = gr.first_vertex; c.cap = d->cap_vertices; c.first_particle = gr.first_particle;
        c.tris = s.tris; c.v_tri_off = s.v_tri_off; c.v_tri = s.v_tri; c.rest = s.verts; c.rigid = gr.rigid;
        for (int k = 0; k < 9; ++k) c.inertia[k] = gr.inertia[k];
        c.shape_volume = gr.shape_volume; c.role = d->ms ? d->role_all : nullptr; c.f_init = 0;
        k_rigid_move<<<gr.count, 256, 0, d->stream>>>(make_geo(d), c, d->vpos, d->vfrc, d->pstate, do_move);
        d->launches++;
        BIB_CHECK(cudaGetLastError());
        return 1;
    }
    k_move_update<<<gr.count, 256, 0, d->stream>>>(make_geo(d), s.nv, s.nt, s.tris, gr.first_vertex, d->cap_vertices, gr.first_particle,
                                                 d->vpos, d->vvel, d->pstate, do_move, skip_mask);
    d->launches++;
    BIB_CHECK(cudaGetLastError());
    return 1;
}

int ms_create(bib_domain* d, int nv, int n_global, int cap_hint, int xcap_hint, int first_slot, long long first_vertex, int gid_base, MultiSlabIB** out);
int ms_set_initial(bib_domain* d, MultiSlabIB* m, const std::vector<int>& owned_lids);
int ms_exchange_positions(bib_domain* d);
int ms_exchange_velocities(bib_domain* d);
int ms_exchange_forces(bib_domain* d);
int ms_exchange_pair_forces(bib_domain* d);
int ms_finish_velocities(bib_domain* d);

// update (and optionally advect) every live particle slot; skip_mask[slot] != 0 leaves a slot alone
int ib_update_all(bib_domain* d, int do_move) {
    const uint8_t* skip = (d->ms && do_move == 0) ? d->seen_all + d->cap_particles : nullptr;   // POS exchange: owners keep their cached scalars
    for (const ParticleGroup& gr : d->groups) if (!launch_update(d, gr, do_move, skip)) return 0;
    return 1;
}

// Multi-slab domains: every rank is given the same global particle list (like the reference's add_particle, which keeps
// only what it owns, src/fsi/ImmersedBoundaryDynamics.hh:500-525); ownership = slab containing the centre of mass.
// Every call adds one particle group with its own slot range and exchange context; global particle ids continue across groups.
static int ib_particles_add_multislab(bib_domain* d, int type, int shape_id, const double* params, int n, const double* positions) {
    const ShapeDev& s = d->shapes[shape_id];
    // centre of mass exactly as DeformableParticle3D::update computes it (sequential, area weighted), then the periodic remap
    std::vector<int> owned;
    double max_extent = 0.0;
    for (int p = 0; p < n; ++p) {
        const double* P = positions + (size_t)p * s.nv * 3;
        double area = 0, cx = 0, lo = P[0], hi = P[0];
        for (int v = 0; v < s.nv; ++v) { lo = std::min(lo, P[3 * v]); hi = std::max(hi, P[3 * v]); }
        max_extent = std::max(max_extent, hi - lo);
        for (int t = 0; t < s.nt; ++t) {
            const double *v0 = P + 3 * s.h_tris[3 * t], *v1 = P + 3 * s.h_tris[3 * t + 1], *v2 = P + 3 * s.h_tris[3 * t + 2];
            const double e1[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]}, e2[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
            const double nrm[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
            double dA = std::sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
            dA *= 0.5;
            area += dA;
            cx += ((v0[0] + v1[0] + v2[0]) / 3.0) * dA;
        }
        double x = cx / area;
        if (d->per[0]) { const double L = (double)d->gnx; x = x - L * std::floor((x - (0.0 - 0.5)) * (1. / L)); }
        if (x >= (double)d->x_offset - 0.5 && x < (double)(d->x_offset + d->nx) - 0.5) owned.push_back(p);
    }
    const double slab_w = (double)(d->gnx / d->n_ranks);            // narrowest slab: a global quantity (capacities must agree on every rank)
    if (max_extent + 6.0 >= slab_w) { bib_set_error("particles_add: the slab width must exceed the particle extent + 6 cells (a particle may only reach the two neighbouring slabs)"); return 0; }
    // slot / exchange capacities from the geometry: a particle is mirrored on a neighbour while its box (+3 cells) reaches over the
    // face, i.e. a fraction (extent + 6) / slab width of the owned ones per side; generous safety factors, overflow is detected
    const double frac = std::min(1.0, (max_extent + 8.0) / slab_w);
    // (the exchange capacity fixes the message size, so it must come out identical on every rank: global quantities only)
    const double base = 1.25 * (double)n / d->n_ranks;
    const int xcap_hint = (int)(base * frac * 2.0) + 32;
    const int cap_hint = std::max((int)(base * (1.0 + 2.0 * frac) * 1.3) + 64, (int)owned.size() + 2 * xcap_hint);
    int gid_base = 0;
    for (MultiSlabIB* m : d->msv) if (m) gid_base += m->n_global;
    MultiSlabIB* m = nullptr;
    if (!ms_create(d, s.nv, n, cap_hint, xcap_hint, d->n_particles, d->n_vertices, gid_base, &m)) return 0;
    if ((int)owned.size() > m->cap) { bib_set_error("particles_add: more owned particles than slots"); return 0; }
    if (!grow_particles(d, d->n_vertices + (long long)m->cap * s.nv, d->n_particles + m->cap)) return 0;
    const long long capv = d->cap_vertices;
    std::vector<double> soa((size_t)owned.size() * s.nv * 3);
    const size_t no = owned.size() * (size_t)s.nv;
    for (size_t k = 0; k < owned.size(); ++k) for (int v = 0; v < s.nv; ++v) for (int c = 0; c < 3; ++c)
        soa[(size_t)c * no + k * s.nv + v] = positions[((size_t)owned[k] * s.nv + v) * 3 + c];
    for (int c = 0; c < 3 && no > 0; ++c)
        BIB_CHECK(cudaMemcpyAsync(d->vpos + (size_t)c * capv + m->first_vertex, soa.data() + (size_t)c * no, no * sizeof(double), cudaMemcpyHostToDevice, d->stream));
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    if (!ms_set_initial(d, m, owned)) return 0;
    ParticleGroup gr;
    gr.type = type; gr.shape = shape_id; gr.first_particle = m->first_slot; gr.count = m->cap; gr.first_vertex = m->first_vertex;
    std::memcpy(gr.params, params, 8 * sizeof(double));
    std::vector<double> ks(s.nl), kp(s.nl);
    for (int l = 0; l < s.nl; ++l) rbc_ks_kp(s.h_link_len[l], params[0], ks[l], kp[l]);
    if (!to_device(&gr.link_ks, ks) || !to_device(&gr.link_kp, kp)) return 0;
    d->groups.push_back(gr);
    d->msv.push_back(m);
    d->ms = d->msv[0];
    d->n_particles += m->cap; d->n_vertices += (long long)m->cap * s.nv;
    d->band_pos_valid = false;
    const long long nvg = (long long)m->cap * s.nv;
    k_fill_vobj<<<(unsigned)((nvg + 255) / 256), 256, 0, d->stream>>>(d->vslot, gr.first_vertex, s.nv, gr.first_particle, nvg);
    d->launches++;
    return launch_update(d, gr, 0);
}

int ib_particles_add(bib_domain* d, int type, int shape_id, const double* params, int n, const double* positions) {
    if (type < 0 || type > 2) { bib_set_error("particles_add: unknown particle type (BIB_PARTICLE_RBC, BIB_PARTICLE_CAPSULE, BIB_PARTICLE_SEMIRIGID; rigid particles: bib_ib_rigid_particles_add)"); return 0; }
    if (shape_id < 0 || shape_id >= (int)d->shapes.size()) { bib_set_error("particles_add: unknown shape"); return 0; }
    if (n <= 0) return 1;
    if (d->n_ranks > 1) return ib_particles_add_multislab(d, type, shape_id, params, n, positions);
    const ShapeDev& s = d->shapes[shape_id];
    const long long nvtot = (long long)n * s.nv;
    if (!grow_particles(d, d->n_vertices + nvtot, d->n_particles + n)) return 0;
    ParticleGroup gr;
    gr.type = type; gr.shape = shape_id; gr.first_particle = d->n_particles; gr.count = n; gr.first_vertex = d->n_vertices;
    std::memcpy(gr.params, params, 8 * sizeof(double));
    std::vector<double> ks(s.nl), kp(s.nl);
    for (int l = 0; l < s.nl; ++l) rbc_ks_kp(s.h_link_len[l], params[0], ks[l], kp[l]);
    if (!to_device(&gr.link_ks, ks) || !to_device(&gr.link_kp, kp)) return 0;
    // AoS [n][nv][3] -> SoA
    std::vector<double> soa((size_t)nvtot * 3);
    for (long long v = 0; v < nvtot; ++v) for (int c = 0; c < 3; ++c) soa[(size_t)c * nvtot + v] = positions[(size_t)v * 3 + c];
    for (int c = 0; c < 3; ++c)
        BIB_CHECK(cudaMemcpyAsync(d->vpos + (size_t)c * d->cap_vertices + gr.first_vertex, soa.data() + (size_t)c * nvtot, (size_t)nvtot * sizeof(double), cudaMemcpyHostToDevice, d->stream));
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    d->groups.push_back(gr);
    d->n_particles += n; d->n_vertices += nvtot;
    d->band_pos_valid = false;
    k_fill_slots<<<(n + 255) / 256, 256, 0, d->stream>>>(gr.first_particle, n, s.nv, gr.first_vertex, d->role_all, d->gid_all, d->slot_nv, d->slot_fv);
    d->launches++;
    // ParticleBase3D::update() at construction + add_particle's map_center_of_mass_to_periodic_grid (ImmersedBoundaryDynamics.hh:500-525)
    return launch_update(d, gr, 0);
}

// RigidParticle3D particles: state17 per particle = com[3], velocity[3], angular velocity[3], quaternion[4], density, damping,
// relaxation factor, scale. Node positions follow from the state (RigidParticle3D::update).
int ib_rigid_add(bib_domain* d, int shape_id, const double* inertia9, double shape_volume, int n, const double* state17) {
    if (shape_id < 0 || shape_id >= (int)d->shapes.size()) { bib_set_error("rigid_particles_add: unknown shape"); return 0; }
    if (d->n_ranks > 1) { bib_set_error("rigid particles: single-slab domains only (round 1)"); return 0; }
    if (n <= 0) return 1;
    const ShapeDev& s = d->shapes[shape_id];
    const long long nvtot = (long long)n * s.nv;
    if (!grow_particles(d, d->n_vertices + nvtot, d->n_particles + n)) return 0;
    ParticleGroup gr;
    gr.type = 3; gr.shape = shape_id; gr.first_particle = d->n_particles; gr.count = n; gr.first_vertex = d->n_vertices;
    for (int k = 0; k < 9; ++k) gr.inertia[k] = inertia9[k];
    gr.shape_volume = shape_volume;
    std::vector<double> rs((size_t)n * RS, 0.0), st((size_t)n * 11, 0.0);
    for (int p = 0; p < n; ++p) {
        const double* in = state17 + (size_t)p * 17;
        for (int k = 0; k < 13; ++k) rs[(size_t)p * RS + k] = in[k];
        for (int k = 0; k < 4; ++k) rs[(size_t)p * RS + 19 + k] = in[13 + k];
        st[(size_t)p * 11] = s.area; st[(size_t)p * 11 + 1] = s.volume;       // ParticleBase3D constructor: the shape's area and volume
    }
    if (!to_device(&gr.rigid, rs)) return 0;
    BIB_CHECK(cudaMemcpyAsync(d->pstate + (size_t)gr.first_particle * 11, st.data(), st.size() * sizeof(double), cudaMemcpyHostToDevice, d->stream));
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    d->groups.push_back(gr);
    d->n_particles += n; d->n_vertices += nvtot;
    k_fill_slots<<<(n + 255) / 256, 256, 0, d->stream>>>(gr.first_particle, n, s.nv, gr.first_vertex, d->role_all, d->gid_all, d->slot_nv, d->slot_fv);
    d->launches++;
    return launch_update(d, gr, 0);
}
int ib_rigid_download(bib_domain* d, int first, int count, double* state19) {
    if (first < 0 || count < 0 || first + count > d->n_particles) { bib_set_error("rigid_state_download: range"); return 0; }
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    for (int p = first; p < first + count; ++p) {
        const ParticleGroup* g = nullptr;
        for (const ParticleGroup& gr : d->groups) if (p >= gr.first_particle && p < gr.first_particle + gr.count) g = &gr;
        if (!g || g->type != 3) { bib_set_error("rigid_state_download: not a rigid particle"); return 0; }
        BIB_CHECK(cudaMemcpy(state19 + (size_t)(p - first) * 19, g->rigid + (size_t)(p - g->first_particle) * RS, 19 * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return 1;
}

// Host-evaluated ForceDecorators (src/fsi/ForceDecorator.h:11-13; ImmersedBoundaryDynamics_forces.hh:293-297 applies them after reset_forces()
// and before compute_forces()): the next force evaluation starts from these per-vertex values instead of zero.
int ib_forces_preload(bib_domain* d, long long first, long long count, const double* f3) {
    if (first < 0 || count < 0 || first + count > d->n_vertices) { bib_set_error("vertex_forces_preload: range"); return 0; }
    if (d->n_vertices == 0) return 1;
    if (!d->vfrc_preloaded) BIB_CHECK(cudaMemsetAsync(d->vfrc, 0, (size_t)d->cap_vertices * 3 * sizeof(double), d->stream));
    if (count > 0) {
        std::vector<double> tmp((size_t)count * 3);
        for (long long v = 0; v < count; ++v) for (int c = 0; c < 3; ++c) tmp[(size_t)c * count + v] = f3[(size_t)v * 3 + c];
        for (int c = 0; c < 3; ++c)
            BIB_CHECK(cudaMemcpyAsync(d->vfrc + (size_t)c * d->cap_vertices + first, tmp.data() + (size_t)c * count, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, d->stream));
        BIB_CHECK(cudaStreamSynchronize(d->stream));
    }
    d->vfrc_preloaded = true;
    return 1;
}

// the host mirror hands back the rigid-body state it holds (the node records travel through ib_vertices_upload)
int ib_rigid_upload(bib_domain* d, int first, int count, const double* state19) {
    if (first < 0 || count < 0 || first + count > d->n_particles) { bib_set_error("rigid_state_upload: range"); return 0; }
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    for (int p = first; p < first + count; ++p) {
        const ParticleGroup* g = nullptr;
        for (const ParticleGroup& gr : d->groups) if (p >= gr.first_particle && p < gr.first_particle + gr.count) g = &gr;
        if (!g || g->type != 3) { bib_set_error("rigid_state_upload: not a rigid particle"); return 0; }
        BIB_CHECK(cudaMemcpy(g->rigid + (size_t)(p - g->first_particle) * RS, state19 + (size_t)(p - first) * 19, 19 * sizeof(double), cudaMemcpyHostToDevice));
    }
    return 1;
}

// ------------------------------------------------------------------------------------------------ host: IB phases
static int ensure_band_masks(bib_domain* d) {
    if (d->band_mask && d->band_cap >= d->n_particles) return 1;
    if (d->band_mask) { cudaFree(d->band_mask); cudaFree(d->band_box); d->band_mask = nullptr; d->band_box = nullptr; }
    d->band_cap = std::max(d->n_particles, 1);
    BIB_CHECK(cudaMalloc(&d->band_mask, 2 * (size_t)d->band_cap * BAND_WORDS * sizeof(unsigned)));
    BIB_CHECK(cudaMalloc(&d->band_box, 2 * (size_t)d->band_cap * 6 * sizeof(int)));
    BIB_CHECK(cudaMemsetAsync(d->band_box, 0, 2 * (size_t)d->band_cap * 6 * sizeof(int), d->stream));
    d->band_valid[0] = d->band_valid[1] = false;
    d->bytes += 2LL * d->band_cap * (BAND_WORDS * 4 + 24);
    return 1;
}
static int ensure_velocity_field(bib_domain* d) {
    if (d->vel) return 1;
    BIB_CHECK(cudaMalloc(&d->vel, d->vol * 3 * sizeof(double)));
    BIB_CHECK(cudaMemsetAsync(d->vel, 0, d->vol * 3 * sizeof(double), d->stream));
    d->bytes += (long long)d->vol * 3 * 8;
    return 1;
}

// interpolate_velocity: velocity u = j/rho + F/2 is evaluated in the particles' bands unless the caller has just computed
// the full field with bib_compute_velocity (the reference's calling sequence), then gathered with the Roma kernel.
int ib_interpolate(bib_domain* d) {
    if (d->n_vertices == 0) return 1;
    d->flags_by_masks = false;
    if (d->ms && d->ghosts_dirty && !ms_exchange_positions(d)) return 0;
    if (!d->vel_full_valid) {
        if (!ensure_band_masks(d)) return 0;
        // never overwrite the masks of a spread that has not been reset yet (ib_band_reset reads them)
        d->band_cur = d->spread_band >= 0 ? 1 - d->spread_band : 1 - d->band_cur;
        unsigned* ms = d->band_mask + (size_t)d->band_cur * d->band_cap * BAND_WORDS;
        int* bs = d->band_box + (size_t)d->band_cur * d->band_cap * 6;
        // fused tile kernel unless the set-up has boundary types it does not know (velocity edges) or asks for the two-kernel path
        bool fused = d->fused_interp && d->edges.empty() && d->walls.size() <= 6;
        // particles larger than the tile read the populations directly: not valid next to wall planes / cell kinds (they need the stored velocities)
        if (fused && (!d->walls.empty() || d->kind)) for (const ParticleGroup& gr : d->groups) if (d->shapes[gr.shape].nv > 2048) fused = false;
        WallsDev W; W.n = fused ? (int)d->walls.size() : 0;
        for (int q = 0; q < W.n; ++q) {
            const WallPlane& w = d->walls[q];
            W.dir[q] = w.dir; W.coord[q] = w.coord; W.nb[q] = w.nb; W.gx[q] = w.gx; W.a0[q] = w.a0; W.a1[q] = w.a1; W.b0[q] = w.b0; W.b1[q] = w.b1; W.vel[q] = w.vel;
        }
        // A handful of particles on a small lattice (the reference's own driver set-ups): one CTA per particle leaves the GPU idle. One warp per
        // vertex reads the populations directly (k_interpolate_warp); no band masks are taken, so the next force reset runs over the whole (small) lattice.
        if (fused && !d->ms && d->n_particles * 4 < d->n_sms && d->n_vertices <= 8192 && d->vol <= (size_t)4 << 20) {      // (27 RBCs: the tile kernel is faster again)
            k_interpolate_warp<<<(unsigned)((d->n_vertices * 32 + 127) / 128), 128, 0, d->stream>>>(make_geo(d), W, d->f[d->cur], d->force, d->n_vertices, d->cap_vertices, d->vpos, d->vvel);
            d->launches++;
            d->band_valid[d->band_cur] = false;     // no masks were written into this buffer
            d->band_pos_valid = false;
            BIB_CHECK(cudaGetLastError());
            pt_mark(d, PT_INTERPOLATE);
            return 1;
        }
        if (fused) {
            const size_t smem = (size_t)3 * TILE_SLOTS * sizeof(double);
            unsigned* flags = make_geo(d).fflag;          // non-null when the sparse force flags are maintained
            for (const ParticleGroup& gr : d->groups) {
                const ShapeDev& s = d->shapes[gr.shape];
                k_interp_tile<<<gr.count, 256, smem, d->stream>>>(make_geo(d), W, d->f[d->cur], d->force, d->pstate, d->vpos, d->vvel, d->cap_vertices,
                                                                  s.nv, gr.first_vertex, gr.first_particle, ms, bs, flags, d->interp_compact);
                d->launches++;
            }
            d->flags_by_masks = flags != nullptr;
            d->band_valid[d->band_cur] = true;
            d->band_pos_valid = true;
            BIB_CHECK(cudaGetLastError());
            pt_mark(d, PT_INTERPOLATE);
            if (d->ms && !ms_exchange_velocities(d)) return 0;     // owners add the partial sums of the neighbouring slabs
            return 1;
        }
        if (!ensure_velocity_field(d)) return 0;
        for (const ParticleGroup& gr : d->groups) {
            const ShapeDev& s = d->shapes[gr.shape];
            k_band_velocity<<<gr.count, 256, 0, d->stream>>>(make_geo(d), d->f[d->cur], d->force, d->vel, d->pstate, d->vpos, d->cap_vertices,
                                                           s.nv, gr.first_vertex, gr.first_particle, ms, bs);
            d->launches++;
        }
        d->band_valid[d->band_cur] = true;
        d->band_pos_valid = true;
        if (!lbm_wall_velocity_fill(d)) return 0;      // wall cells report their stored boundary velocity
    } else {
        if (!ensure_velocity_field(d)) return 0;
        d->band_pos_valid = false;                     // the caller's full velocity field is used: no masks were taken at these positions
    }
    const int bs = 128;
    const unsigned grid = (unsigned)((d->n_vertices + bs - 1) / bs);
    k_interpolate<true><<<grid, bs, 0, d->stream>>>(make_geo(d), d->f[d->cur], d->force, d->vel, d->n_vertices, d->cap_vertices, d->vpos, d->vvel);
    d->launches++;
    BIB_CHECK(cudaGetLastError());
    pt_mark(d, PT_INTERPOLATE);
    if (d->ms && !ms_exchange_velocities(d)) return 0;     // owners add the partial sums of the neighbouring slabs
    return 1;
}

// F := body inside the particles' bands (everywhere else the force already equals the body force)
int ib_band_reset(bib_domain* d, double fx, double fy, double fz) {
    if (d->n_particles == 0 || d->spread_band == -1) return 1;        // no spread has written since the force field was last set
    // the masks ib_interpolate took at the positions the last spread used are in buffer spread_band (recorded by ib_spread; any
    // number of interpolations may have run since, they write the other buffer). Unknown -> the whole lattice.
    const int prev = d->spread_band;
    if (prev < 0 || !d->band_mask || !d->band_valid[prev] || d->band_cap < d->n_particles) {
        const int box[6] = {0, d->nx - 1, 0, d->ny - 1, 0, d->nz - 1};
        return lbm_launch_set_external(d, box, fx, fy, fz);
    }
    k_band_reset<<<d->n_particles, 256, 0, d->stream>>>(make_geo(d), d->force, d->pstate, fx, fy, fz,
                                                      d->band_mask + (size_t)prev * d->band_cap * BAND_WORDS, d->band_box + (size_t)prev * d->band_cap * 6);
    d->launches++;
    BIB_CHECK(cudaGetLastError());
    d->spread_band = -1;
    return 1;
}

static PPGrid make_pp_grid(const bib_domain* d) {
    PPGrid G;
    const double eps = 2.220446049250313e-16;
    const int n[3] = {d->gnx, d->ny, d->nz};
    for (int k = 0; k < 3; ++k) {
        const double dom0 = 0 - 0.5, dom1 = (double)(n[k] - 1) + 0.5 - eps;
        G.nc[k] = (int)std::floor((dom1 - dom0) / 1.);
        G.cs[k] = (dom1 - dom0) / (double)G.nc[k];
        G.nc[k] += 2;
        G.g0[k] = dom0 - G.cs[k];
        G.lo[k] = dom0 - d->pp_pot.r_min; G.hi[k] = dom1 + d->pp_pot.r_min;
        G.per[k] = d->per[k]; G.L[k] = (double)n[k];
    }
    G.role = d->ms ? d->role_all : nullptr; G.vslot = d->vslot;
    G.x_lo = (double)d->x_offset - 0.5; G.x_hi = (double)(d->x_offset + d->nx) - 0.5;
    G.x_min = -0.5; G.x_max = std::nextafter((double)d->gnx - 0.5, 0.0);
    return G;
}
static int ib_pp_forces(bib_domain* d, cudaStream_t st) {
    if (d->pp_pot.r_min > 1.0) { bib_set_error("The interaction distance cannot be greater than the lattice spacing (1)"); return 0; }   // _forces.hh:347-350
    const PPGrid G = make_pp_grid(d);
    const size_t ncell = (size_t)G.nc[0] * G.nc[1] * G.nc[2];
    if (!d->pp_head) {
        BIB_CHECK(cudaMalloc(&d->pp_head, ncell * sizeof(int)));
        BIB_CHECK(cudaMemsetAsync(d->pp_head, 0xff, ncell * sizeof(int), st));
        d->bytes += (long long)(ncell * sizeof(int));
    }
    if (d->pp_cap < d->cap_vertices) {
        if (d->pp_next) cudaFree(d->pp_next);
        if (d->pp_vobj) cudaFree(d->pp_vobj);
        BIB_CHECK(cudaMalloc(&d->pp_next, (size_t)d->cap_vertices * sizeof(int)));
        BIB_CHECK(cudaMalloc(&d->pp_vobj, (size_t)d->cap_vertices * sizeof(int)));
        d->pp_cap = d->cap_vertices; d->pp_groups = 0;
    }
    if (d->ms) {          // ghost copies carry only their partial pair force
        k_zero_ghost_forces<<<(unsigned)((d->n_vertices + 255) / 256), 256, 0, st>>>(d->role_all, d->vslot, d->n_vertices, d->cap_vertices, d->vfrc);
        d->launches++;
    }
    if (d->pp_groups != (int)d->groups.size()) {
        for (const ParticleGroup& gr : d->groups) {
            const long long n = (long long)gr.count * d->shapes[gr.shape].nv;
            k_fill_vobj<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d->pp_vobj, gr.first_vertex, d->shapes[gr.shape].nv, gr.first_particle, n);
            d->launches++;
        }
        d->pp_groups = (int)d->groups.size();
    }
    const unsigned grid = (unsigned)((d->n_vertices + 255) / 256);
    k_pp_build<<<grid, 256, 0, st>>>(G, d->n_vertices, d->cap_vertices, d->vpos, d->pp_head, d->pp_next);
    k_pp_forces<<<(unsigned)((d->n_vertices + 127) / 128), 128, 0, st>>>(G, d->pp_pot, d->n_vertices, d->cap_vertices, d->vpos, d->pp_vobj,
                                                                            d->pp_head, d->pp_next, d->vfrc);
    k_pp_clear<<<grid, 256, 0, st>>>(G, d->n_vertices, d->cap_vertices, d->vpos, d->pp_head);
    d->launches += 3;
    BIB_CHECK(cudaGetLastError());
    if (d->ms && !ms_exchange_pair_forces(d)) return 0;      // ghost holders -> owner (added there): src/fsi/ImmersedBoundaryDynamics_forces.hh:378-456
    return 1;
}

// Vertex forces of every particle group (decorators first, then the particle models, then the particle-particle force), launched on `st`.
// They depend on the vertex positions (and, for rigid particles, on the interpolated velocities) only, not on the lattice.
int ib_forces(bib_domain* d, cudaStream_t st) {
    if (d->n_vertices == 0) return 1;
    Geo wg = make_geo(d);
    if (d->wall_bkind >= 0) { wg.plates = d->wall_bkind; wg.py0 = d->wall_b[0]; wg.py1 = d->wall_b[1]; wg.prad = d->wall_b[2]; }
    const bool wall = d->wall_pot.type != 0 && wg.plates != 0;
    if (wall && d->vfrc_preloaded) { bib_set_error("vertex_forces_preload cannot be combined with a device-side wall interaction"); return 0; }
    const bool init = wall || d->vfrc_preloaded;       // vfrc already holds the decorators' contribution
    d->vfrc_preloaded = false;
    if (wall) {
        k_wall_interaction<<<(unsigned)((d->n_vertices + 255) / 256), 256, 0, st>>>(wg, d->wall_pot, d->n_vertices, d->cap_vertices, d->vpos, d->vfrc);
        d->launches++;
    }
    for (const ParticleGroup& gr : d->groups) {
        const ShapeDev& s = d->shapes[gr.shape];
        if (gr.type == 3) {
            RigidParams c;
            c.nv = s.nv; c.nt = s.nt; c.first_vertex = gr.first_vertex; c.cap = d->cap_vertices; c.first_particle = gr.first_particle;
            c.tris = s.tris; c.v_tri_off = s.v_tri_off; c.v_tri = s.v_tri; c.rest = s.verts; c.rigid = gr.rigid;
            for (int k = 0; k < 9; ++k) c.inertia[k] = gr.inertia[k];
            c.shape_volume = gr.shape_volume; c.role = d->ms ? d->role_all : nullptr; c.f_init = init ? 1 : 0;
            k_rigid_forces<<<gr.count, 256, 0, st>>>(c, d->vpos, d->vvel, d->vfrc);
            d->launches++;
            continue;
        }
        if (gr.type == 2) {
            SemiParams c;
            c.nv = s.nv; c.first_vertex = gr.first_vertex; c.cap = d->cap_vertices; c.first_particle = gr.first_particle;
            c.links = s.links; c.link_len = s.link_len; c.v_link_off = s.v_link_off; c.v_link = s.v_link; c.rest_radial = s.rest_radial;
            c.k_in = gr.params[1]; c.k_out = gr.params[2]; c.role = d->ms ? d->role_all : nullptr; c.f_init = init ? 1 : 0;
            const size_t smem = (size_t)s.nv * 6 * sizeof(double);
            if (smem > 200 * 1024) { bib_set_error("semi-rigid particles: mesh too large for the shared-memory kernel"); return 0; }
            if (smem > 48 * 1024) cudaFuncSetAttribute(k_semirigid_forces, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_semirigid_forces<<<gr.count, 256, smem, st>>>(c, d->vpos, d->vfrc, d->pstate);
            d->launches++;
            continue;
        }
        if (gr.type == 1) {      // BIB_PARTICLE_CAPSULE
            CapParams c;
            c.nv = s.nv; c.first_vertex = gr.first_vertex; c.cap = d->cap_vertices; c.first_particle = gr.first_particle; c.count = gr.count;
            c.tris = s.tris; c.tri_area = s.tri_area; c.rest = s.verts; c.v_tri_off = s.v_tri_off; c.v_tri = s.v_tri;
            c.G = gr.params[0]; c.role = d->ms ? d->role_all : nullptr; c.f_init = init ? 1 : 0;
            const size_t smem = (size_t)s.nv * 3 * sizeof(double);
            c.use_smem = smem <= 200 * 1024;
            if (c.use_smem && smem > 48 * 1024) cudaFuncSetAttribute(k_capsule_forces, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_capsule_forces<<<gr.count, 256, c.use_smem ? smem : 0, st>>>(c, d->vpos, d->vfrc);
            d->launches++;
            continue;
        }
        RbcParams q;
        q.nv = s.nv; q.first_vertex = gr.first_vertex; q.cap = d->cap_vertices; q.first_particle = gr.first_particle; q.count = gr.count;
        q.tris = s.tris; q.tri_area = s.tri_area; q.links = s.links; q.link_len = s.link_len; q.link_ks = gr.link_ks; q.link_kp = gr.link_kp;
        q.pairs = s.pairs; q.pair_cs = s.pair_cs;
        q.v_link_off = s.v_link_off; q.v_link = s.v_link; q.v_tri_off = s.v_tri_off; q.v_tri = s.v_tri; q.v_pair_off = s.v_pair_off; q.v_pair = s.v_pair;
        q.shape_area = s.area; q.k_ag = gr.params[2]; q.k_al = gr.params[3]; q.k_v = gr.params[4]; q.vol_desired = gr.params[5]; q.k_b = gr.params[6];
        q.reg_p = d->regulator_p; for (int k = 0; k < 3; ++k) q.reg_t[k] = d->regulator_target[k];
        q.role = d->ms ? d->role_all : nullptr;
        q.f_init = init ? d->vfrc : nullptr;
        const size_t smem = (size_t)s.nv * 3 * sizeof(double);
        // A handful of particles (the reference's own driver set-ups: 1 .. 28 RBCs) cannot fill 148 SMs with one CTA per particle: there the
        // vertex-parallel gather kernel runs (one thread per vertex, 3x the arithmetic, nv / 128 CTAs per particle; the same bits).
        const bool few = gr.count * 4 < d->n_sms && !d->ms;
        if (few) {
            const long long n = (long long)gr.count * s.nv;
            k_rbc_forces<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(q, d->vpos, d->vfrc, d->pstate);
        } else if (s.nv <= RF_THREADS * RF_VPT_MAX) {
            const size_t sm2 = smem + RF_CHUNK * sizeof(double);
            const int vpt = (s.nv + RF_THREADS - 1) / RF_THREADS;        // vertices per thread (measured: 3 instead of 4 for the 626-vertex RBC = -0.05 ms)
#define BIB_RF_LAUNCH(V) k_rbc_forces_elem<V><<<gr.count, RF_THREADS, sm2, st>>>(q, d->vpos, d->vfrc, d->pstate, s.nt, s.nl, s.np)      /* shared-memory opt-in: ib_set_kernel_attributes */
            if (vpt <= 1) BIB_RF_LAUNCH(1); else if (vpt == 2) BIB_RF_LAUNCH(2); else if (vpt == 3) BIB_RF_LAUNCH(3); else BIB_RF_LAUNCH(4);
#undef BIB_RF_LAUNCH
        } else if (smem <= 200 * 1024) {
            if (smem > 48 * 1024) cudaFuncSetAttribute(k_rbc_forces_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_rbc_forces_cta<<<gr.count, 256, smem, st>>>(q, d->vpos, d->vfrc, d->pstate);
        } else {
            const long long n = (long long)gr.count * s.nv;
            k_rbc_forces<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(q, d->vpos, d->vfrc, d->pstate);
        }
        d->launches++;
    }
    if (d->pp_pot.type != 0 && !ib_pp_forces(d, st)) return 0;   // after every particle's own forces (_forces.hh:300-302)
    BIB_CHECK(cudaGetLastError());
    return 1;
}

// Can the force evaluation run beside the velocity interpolation of the same step (bib_step)? Not with rigid particles (their direct
// forcing reads the interpolated velocities) and not on several slabs (the force exchange is ordered on the compute stream).
bool ib_forces_independent_of_interpolation(const bib_domain* d) {
    if (d->n_vertices == 0 || d->ms || d->n_ranks > 1) return false;
    for (const ParticleGroup& gr : d->groups) if (gr.type == 3) return false;
    return true;
}

int ib_spread(bib_domain* d) {
    if (d->n_vertices == 0) return 1;
    // remember which masks describe the cells written now; a second spread before any reset, or one without masks of the current positions: unknown
    d->spread_band = (d->spread_band == -1 && d->band_pos_valid && d->band_mask && d->band_valid[d->band_cur]) ? d->band_cur : -2;
    // the interpolation of these very positions has raised the force flags of every cell written below (k_interp_tile): no per-point flag atomics
    const bool flags_done = d->flags_by_masks && d->band_pos_valid && d->band_valid[d->band_cur];
    if (d->ms && !ms_exchange_forces(d)) return 0;         // ghost copies receive the owner's vertex forces
    pt_mark(d, PT_X_FORCE);
    if (d->deterministic_spread) d->fflag_valid = false;
    if (d->deterministic_spread) k_spread_sequential<<<1, 32, 0, d->stream>>>(make_geo(d), d->force, d->n_vertices, d->cap_vertices, d->vpos, d->vfrc);
    else if (d->spread_mode == 1 || (d->n_particles * 4 < d->n_sms && !d->ms)) {      // few particles: one thread per vertex instead of one CTA per particle
        k_spread_atomic<<<(unsigned)((d->n_vertices + 127) / 128), 128, 0, d->stream>>>(make_geo(d), d->force, d->n_vertices, d->cap_vertices, d->vpos, d->vfrc);
    } else {
        const size_t smem = (size_t)TILE_CELLS * 3 * sizeof(double);
        for (const ParticleGroup& gr : d->groups) {
            Geo sg = make_geo(d);
            if (flags_done) sg.fflag = nullptr;
            k_spread_tile<<<gr.count, 256, smem, d->stream>>>(sg, d->force, d->pstate, d->vpos, d->vfrc, d->cap_vertices, d->shapes[gr.shape].nv,
                                                         gr.first_vertex, gr.first_particle, d->spread_sorted ? d->shapes[gr.shape].spread_perm : nullptr);
            d->launches++;
        }
        d->launches--;
    }
    d->launches++;
    BIB_CHECK(cudaGetLastError());
    pt_mark(d, PT_SPREAD);
    return 1;
}

int ib_forces_and_spread(bib_domain* d) {
    if (!ib_forces(d, d->stream)) return 0;
    pt_mark(d, PT_FORCES);
    return ib_spread(d);
}

// Function attributes belong to the CURRENT device: every device a domain lives on needs the opt-in above 48 KB of dynamic shared
// memory (k_spread_tile: 72 KB tile; k_rbc_forces_elem: vertex positions of up to RF_THREADS*RF_VPT_MAX vertices + one element chunk).
int ib_set_kernel_attributes() {
    const int tile = (int)((size_t)TILE_CELLS * 3 * sizeof(double));
    const int rf = (int)((size_t)RF_THREADS * RF_VPT_MAX * 3 * sizeof(double) + RF_CHUNK * sizeof(double));
    BIB_CHECK(cudaFuncSetAttribute(k_spread_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, tile));
    BIB_CHECK(cudaFuncSetAttribute(k_interp_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)3 * TILE_SLOTS * sizeof(double))));
    BIB_CHECK(cudaFuncSetAttribute(k_rbc_forces_elem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, rf));
    BIB_CHECK(cudaFuncSetAttribute(k_rbc_forces_elem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, rf));
    BIB_CHECK(cudaFuncSetAttribute(k_rbc_forces_elem<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, rf));
    BIB_CHECK(cudaFuncSetAttribute(k_rbc_forces_elem<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, rf));
    return 1;
}

__global__ void k_force_accumulate(double* __restrict__ sum, double* __restrict__ vfrc, long long n, int mode) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        if (mode == 0) sum[t] = 0.0;
        else if (mode == 1) sum[t] += vfrc[t];
        else vfrc[t] = -sum[t];
    }
}
int ib_force_accumulate(bib_domain* d, int mode) {
    if (d->n_vertices == 0) return 1;
    if (d->vfrc_sum_cap < d->cap_vertices) {
        if (d->vfrc_sum) cudaFree(d->vfrc_sum);
        BIB_CHECK(cudaMalloc(&d->vfrc_sum, (size_t)d->cap_vertices * 3 * sizeof(double)));
        d->vfrc_sum_cap = d->cap_vertices;
    }
    const long long n = d->cap_vertices * 3;
    k_force_accumulate<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, d->stream>>>(d->vfrc_sum, d->vfrc, n, mode);
    d->launches++;
    BIB_CHECK(cudaGetLastError());
    return 1;
}

int ib_move_and_voxelize(bib_domain* d) {
    if (d->ms && !ms_finish_velocities(d)) return 0;       // (a VEL exchange whose sync was deferred and that no FORCE exchange completed)
    for (const ParticleGroup& gr : d->groups) if (!launch_update(d, gr, 1)) return 0;
    pt_mark(d, PT_MOVE);
    if (d->ms && !ms_exchange_positions(d)) return 0;      // refresh ghosts / hand particles over before voxelizing
    pt_mark(d, PT_X_POS);
    for (const ParticleGroup& gr : d->groups) {
        const ShapeDev& s = d->shapes[gr.shape];
        const long long n = (long long)gr.count * s.nt;
        k_raycast<<<(unsigned)((n + 127) / 128), 128, 0, d->stream>>>(make_geo(d), s.nv, s.nt, s.tris, gr.first_vertex, d->cap_vertices, gr.count, gr.first_particle, d->vpos, d->phase, d->col_flags);
        d->launches++;
    }
    // apply_voxelization rewrites the phase of every cell of the block each step, also without any particle
    const long long warps = (long long)d->nx * d->ny;
    const unsigned scan_grid = (unsigned)((warps * 32 + 255) / 256);
    if (d->nz % 16 == 0) k_phase_scan_words<4><<<scan_grid, 256, 0, d->stream>>>(d->nx, d->ny, d->nz, (long long)d->plane, d->phase, d->col_flags);
    else if (d->nz % 4 == 0) k_phase_scan_words<1><<<scan_grid, 256, 0, d->stream>>>(d->nx, d->ny, d->nz, (long long)d->plane, d->phase, d->col_flags);
    else k_phase_scan<<<scan_grid, 256, 0, d->stream>>>(d->nx, d->ny, d->nz, (long long)d->plane, d->phase, d->col_flags);
    d->launches++;
    BIB_CHECK(cudaGetLastError());
    return 1;
}

// SoA device arrays -> fsi::Vertex AoS {pos, vel, force}: packed on the device, one D2H copy into a pinned staging buffer
__global__ void k_pack_vertices(double* __restrict__ out, const double* __restrict__ vpos, const double* __restrict__ vvel,
                                const double* __restrict__ vfrc, long long cap, long long first, long long count) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < count * 9; t += (long long)gridDim.x * blockDim.x) {
        const long long v = t / 9; const int k = (int)(t % 9);
        const double* src = k < 3 ? vpos : (k < 6 ? vvel : vfrc);
        out[t] = src[(long long)(k % 3) * cap + first + v];
    }
}
static bool ensure_pinned(bib_domain* d, size_t bytes) {
    if (d->pinned_bytes >= bytes) return true;
    if (d->pinned) cudaFreeHost(d->pinned);
    d->pinned = nullptr; d->pinned_bytes = 0;
    if (!bib_cuda_ok(cudaHostAlloc(&d->pinned, bytes, cudaHostAllocDefault), "cudaHostAlloc")) return false;
    d->pinned_bytes = bytes;
    return true;
}
static bool ensure_ib_scratch(bib_domain* d, size_t bytes) {
    if (d->scratch_bytes >= bytes) return true;
    if (d->scratch) cudaFree(d->scratch);
    d->scratch = nullptr; d->scratch_bytes = 0;
    if (!bib_cuda_ok(cudaMalloc(&d->scratch, bytes), "cudaMalloc(scratch)")) return false;
    d->scratch_bytes = bytes;
    return true;
}
int ib_vertices_download(bib_domain* d, long long first, long long count, double* v9) {
    if (first < 0 || count < 0 || first + count > d->n_vertices) { bib_set_error("vertices_download: range"); return 0; }
    if (count == 0) return 1;
    const size_t bytes = (size_t)count * 9 * sizeof(double);
    if (!ensure_ib_scratch(d, bytes)) return 0;
    const int grid = (int)std::min<long long>((count * 9 + 255) / 256, 148 * 16);
    k_pack_vertices<<<grid, 256, 0, d->stream>>>((double*)d->scratch, d->vpos, d->vvel, d->vfrc, d->cap_vertices, first, count);
    d->launches++;
    // a caller buffer that is already page-locked (cudaHostAlloc / cudaHostRegister) receives the DMA directly
    cudaPointerAttributes attr;
    const bool user_pinned = cudaPointerGetAttributes(&attr, v9) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (user_pinned) {
        BIB_CHECK(cudaMemcpyAsync(v9, d->scratch, bytes, cudaMemcpyDeviceToHost, d->stream));
        BIB_CHECK(cudaStreamSynchronize(d->stream));
        return 1;
    }
    if (!ensure_pinned(d, bytes)) return 0;
    BIB_CHECK(cudaMemcpyAsync(d->pinned, d->scratch, bytes, cudaMemcpyDeviceToHost, d->stream));
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    std::memcpy(v9, d->pinned, bytes);
    return 1;
}
// Multi-slab domains deliver the particles this slab OWNS, compacted in slot order (what the reference's per-rank writers see,
// ImmersedBoundaryDynamics.hh:232-246): the slot arrays are sized for the worst case, and a DMA of all of them (113 MB per step and rank
// in the bench) does not fit under one step once the lattice kernel saturates HBM. The number of owned particles lives on the device;
// the host copies a prefix sized by the last count it saw plus slack, and ib_results_wait fetches the (rare) remainder synchronously.
// (A copy-stream kernel storing straight into mapped host memory was measured too: it only ran after the following step had finished.)
__global__ void __launch_bounds__(1024) k_results_list(const int* __restrict__ role, const int* __restrict__ slot_nv, int cap, int* __restrict__ list,
                                                       int* __restrict__ voff, int* __restrict__ count) {
    __shared__ int s_warp[32], s_wv[32];
    __shared__ int s_base, s_vbase;
    if (threadIdx.x == 0) { s_base = 0; s_vbase = 0; }
    __syncthreads();
    for (int s0 = 0; s0 < cap; s0 += 1024) {
        const int s = s0 + threadIdx.x;
        const bool take = s < cap && role[s] == 1;
        const int nv = take ? slot_nv[s] : 0;
        const unsigned m = __ballot_sync(0xffffffffu, take);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        int incl = nv;                                    // inclusive warp scan of the vertex counts
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_wv[w] = incl;
        if (lane == 0) s_warp[w] = __popc(m);
        __syncthreads();
        int before = s_base, vbefore = s_vbase;
        for (int k = 0; k < w; ++k) { before += s_warp[k]; vbefore += s_wv[k]; }
        if (take) { const int j = before + __popc(m & ((1u << lane) - 1u)); list[j] = s; voff[j] = vbefore + incl - nv; }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0, tv = 0; for (int k = 0; k < 32; ++k) { t += s_warp[k]; tv += s_wv[k]; } s_base += t; s_vbase += tv; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { count[0] = s_base; count[1] = s_vbase; }
}
__global__ void __launch_bounds__(128) k_pack_slots(double* __restrict__ out_v, double* __restrict__ out_s, const int* __restrict__ list, const int* __restrict__ voff,
                                                    const int* __restrict__ count, const int* __restrict__ slot_nv, const long long* __restrict__ slot_fv,
                                                    const double* __restrict__ vpos, const double* __restrict__ vvel, const double* __restrict__ vfrc,
                                                    const double* __restrict__ pstate, long long cap_v) {
    const int j = blockIdx.x;
    if (j >= *count) return;
    const int s = list[j];
    const int nv = slot_nv[s];
    const long long fv = slot_fv[s];
    for (int t = threadIdx.x; t < nv * 9; t += blockDim.x) {
        const int v = t / 9, k = t % 9;
        const double* src = k < 3 ? vpos : (k < 6 ? vvel : vfrc);
        out_v[(long long)voff[j] * 9 + t] = src[(long long)(k % 3) * cap_v + fv + v];
    }
    if (threadIdx.x < 11) out_s[(long long)j * 11 + threadIdx.x] = pstate[(long long)s * 11 + threadIdx.x];
}
// write_lightweight_particle_data (src/fsi/DeformableParticle3D.hh:318-333, compute_tensor_of_gyration :110-125): per particle
// {global id, centre of mass[3], gyration tensor xx xy xz yy yz zz, volume}. One warp per listed slot: every lane sums a strided
// subset of the vertices, the lanes are combined by a shuffle tree (the reference sums sequentially: equal to ~1e-16 relative).
__global__ void __launch_bounds__(128) k_lightweight(double* __restrict__ out, const int* __restrict__ list, const int* __restrict__ count, int n_slots,
                                                     const int* __restrict__ gid, const int* __restrict__ slot_nv, const long long* __restrict__ slot_fv,
                                                     const double* __restrict__ vpos, const double* __restrict__ pstate, long long cap_v) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= (count ? *count : n_slots)) return;
    const int s = list ? list[j] : j;
    const double* st = pstate + (long long)s * 11;
    const int nv = slot_nv[s];
    const long long fv = slot_fv[s];
    double A[6] = {0, 0, 0, 0, 0, 0};
    for (int v = lane; v < nv; v += 32) {
        const double r[3] = {vpos[fv + v] - st[2], vpos[cap_v + fv + v] - st[3], vpos[2 * cap_v + fv + v] - st[4]};
        A[0] += r[0] * r[0]; A[1] += r[0] * r[1]; A[2] += r[0] * r[2]; A[3] += r[1] * r[1]; A[4] += r[1] * r[2]; A[5] += r[2] * r[2];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) for (int o = 16; o > 0; o >>= 1) A[k] += __shfl_xor_sync(0xffffffffu, A[k], o);
    if (lane == 0) {
        double* o = out + (long long)j * 11;
        o[0] = (double)gid[s]; o[1] = st[2]; o[2] = st[3]; o[3] = st[4];
        for (int k = 0; k < 6; ++k) o[4 + k] = A[k] / (double)nv;
        o[10] = st[1];
    }
}
// Read-back that overlaps the following steps: the fsi::Vertex records and the particle scalars are packed on the compute stream into
// a staging buffer; a second stream copies them to the caller's page-locked buffers while the compute stream carries on. Up to
// RES_DEPTH read-backs are in flight (one staging buffer each); ib_results_wait completes the OLDEST one. Two in flight let a caller
// keep one whole step queued ahead of the copy it is waiting for (bench.py e2e): the D2H DMA of 56 MB takes about as long as a step
// while the lattice kernel saturates HBM.
static int res_complete_oldest(bib_domain* d) {
    if (d->res_n == 0) return 1;
    ResSlot& r = d->res[d->res_tail];
    BIB_CHECK(cudaEventSynchronize(r.copied));
    int n = d->n_particles;
    if (r.compact && d->ms) {
        n = r.count_host[0];
        const long long nvert = r.count_host[1];
        const int got = r.bound_used; const long long got_v = r.bound_used_v;
        if (n > got || (!r.light && nvert > got_v)) {       // more particles are owned than the prefix covered (ownership moved towards this slab): fetch the rest now
            const double* stage_v = (const double*)r.dev; const double* stage_s = (const double*)((const char*)r.dev + r.vb);
            if (!r.light && nvert > got_v) BIB_CHECK(cudaMemcpyAsync(r.host_v + (size_t)got_v * 9, stage_v + (size_t)got_v * 9, (size_t)(nvert - got_v) * 9 * sizeof(double), cudaMemcpyDeviceToHost, d->copy_stream));
            if (n > got) BIB_CHECK(cudaMemcpyAsync(r.host_s + (size_t)got * 11, stage_s + (size_t)got * 11, (size_t)(n - got) * 11 * sizeof(double), cudaMemcpyDeviceToHost, d->copy_stream));
            BIB_CHECK(cudaStreamSynchronize(d->copy_stream));
        }
        d->res_bound = std::min(d->n_particles, n + std::max(8, n / 64));     // ownership changes by a few particles per step at most
        if (!r.light) d->res_bound_v = std::min(d->n_vertices, nvert + std::max<long long>(8 * 1024, nvert / 64));
    }
    d->res_last_count = n;
    d->res_tail = (d->res_tail + 1) % RES_DEPTH; d->res_n--;
    return 1;
}
// v9 == NULL: the lightweight records only (s11 then receives {id, com, gyration, volume} per owned particle instead of the cached scalars)
int ib_results_download_async(bib_domain* d, double* v9, double* s11) {
    const bool light = v9 == nullptr;
    cudaPointerAttributes a0, a1;
    const bool pinned = (light || (cudaPointerGetAttributes(&a0, v9) == cudaSuccess && a0.type == cudaMemoryTypeHost)) &&
                        cudaPointerGetAttributes(&a1, s11) == cudaSuccess && a1.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (!pinned) { bib_set_error("results_download_async: the host buffers must be page-locked (cudaHostAlloc / cudaHostRegister)"); return 0; }
    if (!d->copy_stream) BIB_CHECK(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));   // (a high-priority stream was measured: no difference)
    const size_t vb = light ? 0 : (size_t)d->n_vertices * 9 * sizeof(double), sb = (size_t)d->n_particles * 11 * sizeof(double);
    if (vb + sb == 0) return 1;
    if (d->res_n == RES_DEPTH && !res_complete_oldest(d)) return 0;        // every staging buffer is busy: finish the oldest read-back first
    ResSlot& r = d->res[d->res_head];
    if (!r.packed) {
        BIB_CHECK(cudaEventCreateWithFlags(&r.packed, cudaEventDisableTiming));
        BIB_CHECK(cudaEventCreateWithFlags(&r.copied, cudaEventDisableTiming));
        BIB_CHECK(cudaHostAlloc((void**)&r.count_host, 64, cudaHostAllocDefault));
        *r.count_host = 0;
    }
    if (r.bytes < vb + sb) {
        if (r.dev) { cudaFree(r.dev); d->bytes -= (long long)r.bytes; }
        r.dev = nullptr; r.bytes = 0;
        BIB_CHECK(cudaMalloc(&r.dev, vb + sb));
        r.bytes = vb + sb; d->bytes += (long long)(vb + sb);
    }
    r.vb = vb; r.host_v = v9; r.host_s = s11; r.compact = false; r.light = light; r.bound_used = 0; r.bound_used_v = 0;
    double* stage_v = (double*)r.dev; double* stage_s = (double*)((char*)r.dev + vb);
    if (d->ms) {
        // multi-slab: the owned particles, compacted in slot order (vertex records back to back, whatever the particles' vertex counts)
        const int cap = d->n_particles;
        if (!d->res_list || d->res_list_cap < cap) {
            if (d->res_list) cudaFree(d->res_list);
            BIB_CHECK(cudaMalloc(&d->res_list, (size_t)RES_DEPTH * (2 * (size_t)cap + 2) * sizeof(int)));
            d->res_list_cap = cap;
            d->res_bound = cap; d->res_bound_v = d->n_vertices;    // first read-back: everything; afterwards the last count + slack
        }
        int* count = d->res_list + (size_t)d->res_head * (2 * (size_t)cap + 2); int* list = count + 2; int* voff = list + cap;
        k_results_list<<<1, 1024, 0, d->stream>>>(d->role_all, d->slot_nv, cap, list, voff, count);
        if (light) k_lightweight<<<(cap + 3) / 4, 128, 0, d->stream>>>(stage_s, list, count, cap, d->gid_all, d->slot_nv, d->slot_fv, d->vpos, d->pstate, d->cap_vertices);
        else k_pack_slots<<<cap, 128, 0, d->stream>>>(stage_v, stage_s, list, voff, count, d->slot_nv, d->slot_fv, d->vpos, d->vvel, d->vfrc, d->pstate, d->cap_vertices);
        d->launches += 2;
        BIB_CHECK(cudaGetLastError());
        BIB_CHECK(cudaEventRecord(r.packed, d->stream));
        BIB_CHECK(cudaStreamWaitEvent(d->copy_stream, r.packed, 0));
        size_t nb = (size_t)std::min(d->res_bound, cap);
        size_t nbv = light ? 0 : (size_t)std::min(d->res_bound_v, d->n_vertices);
        if (const char* e = getenv("BIB_RESULTS_BOUND")) {     // tests: force the remainder path
            nb = (size_t)std::max(0, std::min(atoi(e), cap));
            nbv = std::min<size_t>(nbv, nb * (size_t)(d->msv[0] ? d->msv[0]->nv : 0));
        }
        BIB_CHECK(cudaMemcpyAsync(r.count_host, count, 2 * sizeof(int), cudaMemcpyDeviceToHost, d->copy_stream));
        if (nbv) BIB_CHECK(cudaMemcpyAsync(v9, stage_v, nbv * 9 * sizeof(double), cudaMemcpyDeviceToHost, d->copy_stream));
        if (nb) BIB_CHECK(cudaMemcpyAsync(s11, stage_s, nb * 11 * sizeof(double), cudaMemcpyDeviceToHost, d->copy_stream));
        r.compact = true; r.bound_used = (int)nb; r.bound_used_v = (long long)nbv;
    } else {
        if (vb) {
            const int grid = (int)std::min<long long>((d->n_vertices * 9 + 255) / 256, 148 * 16);
            k_pack_vertices<<<grid, 256, 0, d->stream>>>(stage_v, d->vpos, d->vvel, d->vfrc, d->cap_vertices, 0, d->n_vertices);
            d->launches++;
        }
        if (sb && light) { k_lightweight<<<(d->n_particles + 3) / 4, 128, 0, d->stream>>>(stage_s, nullptr, nullptr, d->n_particles, d->gid_all, d->slot_nv, d->slot_fv, d->vpos, d->pstate, d->cap_vertices); d->launches++; }
        else if (sb) BIB_CHECK(cudaMemcpyAsync(stage_s, d->pstate, sb, cudaMemcpyDeviceToDevice, d->stream));
        BIB_CHECK(cudaEventRecord(r.packed, d->stream));
        BIB_CHECK(cudaStreamWaitEvent(d->copy_stream, r.packed, 0));
        if (vb) BIB_CHECK(cudaMemcpyAsync(v9, stage_v, vb, cudaMemcpyDeviceToHost, d->copy_stream));
        if (sb) BIB_CHECK(cudaMemcpyAsync(s11, stage_s, sb, cudaMemcpyDeviceToHost, d->copy_stream));
    }
    BIB_CHECK(cudaEventRecord(r.copied, d->copy_stream));
    d->res_head = (d->res_head + 1) % RES_DEPTH; d->res_n++;
    return 1;
}
int ib_results_wait(bib_domain* d) { return res_complete_oldest(d); }
// number of particles the last COMPLETED read-back delivered (single slab: all of them)
int ib_results_count(bib_domain* d, int* n) {
    *n = d->res_last_count >= 0 ? d->res_last_count : d->n_particles;
    return 1;
}
int ib_vertices_upload(bib_domain* d, long long first, long long count, const double* v9) {
    d->band_pos_valid = false;
    if (first < 0 || count < 0 || first + count > d->n_vertices) { bib_set_error("vertices_upload: range"); return 0; }
    if (count == 0) { if (d->ms) d->ghosts_dirty = true; return 1; }
    std::vector<double> tmp((size_t)count * 9);
    for (long long v = 0; v < count; ++v) for (int k = 0; k < 9; ++k) tmp[(size_t)k * count + v] = v9[(size_t)v * 9 + k];
    double* dst[3] = {d->vpos, d->vvel, d->vfrc};
    for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c)
        BIB_CHECK(cudaMemcpyAsync(dst[a] + (size_t)c * d->cap_vertices + first, tmp.data() + (size_t)(a * 3 + c) * count, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, d->stream));
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    if (d->ms) d->ghosts_dirty = true;
    return ib_update_all(d, 0);        // area, volume, centre of mass, bounding box follow the new positions (ParticleBase3D::update)
}
int ib_scalars_upload(bib_domain* d, int first, int count, const double* s11) {
    if (first < 0 || count < 0 || first + count > d->n_particles) { bib_set_error("particle_scalars_upload: range"); return 0; }
    if (count == 0) return 1;
    BIB_CHECK(cudaMemcpyAsync(d->pstate + (size_t)first * 11, s11, (size_t)count * 11 * sizeof(double), cudaMemcpyHostToDevice, d->stream));
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    return 1;
}
int ib_scalars_download(bib_domain* d, int first, int count, double* s11) {
    if (first < 0 || count < 0 || first + count > d->n_particles) { bib_set_error("particle_scalars_download: range"); return 0; }
    if (count == 0) return 1;
    BIB_CHECK(cudaMemcpyAsync(s11, d->pstate + (size_t)first * 11, (size_t)count * 11 * sizeof(double), cudaMemcpyDeviceToHost, d->stream));
    BIB_CHECK(cudaStreamSynchronize(d->stream));
    return 1;
}
