"""Host-side mirror of the reference interface over the C-ABI library (ctypes; no torch types cross the boundary).

`Lattice` plays the role of the reference's ``MultiBlockLattice3D<double, ForcedPhaseD3Q19Descriptor>`` +
``ImmersedBoundaryDynamics3D`` pair for the hot path (reference src/multiBlock/multiBlockLattice3D.hh:306-348,
src/fsi/ImmersedBoundaryDynamics.h:50-217): the method names follow the reference's
(``collideAndStream``, ``computeVelocity``, ``setExternalVector``, ``add_particle`` ...). Every call goes to
libbiblbm_b200.so; there is no CPU fallback: on a machine without a CUDA device construction raises.
"""
from __future__ import annotations

import ctypes as C
import pathlib

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
import os as _os
LIB_PATH = pathlib.Path(_os.environ["BIB_LIB"]) if _os.environ.get("BIB_LIB") else _HERE / "libbiblbm_b200.so"      # BIB_LIB: experiments with alternative builds

# every symbol include/biblbm_b200.h declares (tests check that the built library exports all of them)
SYMBOLS = [
    "bib_last_error", "bib_version", "bib_device_count", "bib_domain_create", "bib_domain_create_multi", "bib_domain_slab_count", "bib_domain_slab", "bib_domain_destroy", "bib_domain_set_option",
    "bib_domain_upload", "bib_domain_download", "bib_domain_set_celltype", "bib_domain_init_equilibrium", "bib_add_velocity_boundary", "bib_set_boundary_velocity", "bib_set_guo_wall_nodes", "bib_set_external_vector", "bib_compute_velocity",
    "bib_velocity_download", "bib_velocity_upload", "bib_velocity_extrapolate_plane", "bib_collide_and_stream", "bib_ib_shape_register", "bib_ib_shape_info", "bib_ib_shape_connectivity",
    "bib_ib_particles_add", "bib_ib_rigid_particles_add", "bib_ib_rigid_state_download", "bib_ib_rigid_state_upload", "bib_ib_set_regulator", "bib_ib_set_parallel_plates", "bib_ib_set_pipe", "bib_ib_set_wall_interaction", "bib_ib_set_wall_interaction_boundary", "bib_ib_set_pp_interaction", "bib_ib_count", "bib_ib_local_slots", "bib_ib_interpolate_velocity", "bib_ib_compute_and_spread_forces",
    "bib_ib_move_vertices_and_revoxelize", "bib_ib_step", "bib_ib_multi_direct_forcing", "bib_add_velocity_edge", "bib_ib_vertices_download", "bib_ib_vertices_upload", "bib_ib_vertex_forces_preload",
    "bib_ib_particle_scalars_download", "bib_ib_particle_scalars_upload", "bib_ib_results_download_async", "bib_ib_lightweight_download_async", "bib_ib_results_wait", "bib_ib_results_count", "bib_step", "bib_comm_unique_id", "bib_comm_init", "bib_sync", "bib_time_begin", "bib_time_end",
    "bib_kernel_launch_count", "bib_device_memory_bytes", "bib_phase_timing", "bib_phase_times", "bib_set_default_inplace", "bib_domain_inplace_shift", "bib_inplace_schedule",
]

OPT_ENVELOPE_SWAP, OPT_DETERMINISTIC_SPREAD, OPT_REGULATOR, OPT_SPREAD_MODE, OPT_SPARSE_FORCE, OPT_STEP_GRAPH, OPT_FUSED_FACE_SWAP, OPT_OVERLAP_FORCES, OPT_PREFETCH_FORCES = 1, 2, 3, 4, 5, 6, 7, 8, 9
OPT_FUSED_INTERPOLATION = 10
OPT_FUSE_CALLS = 11
OPT_LBM_PREFETCH = 12
OPT_INTERP_COMPACT = 13
CELL_BULK, CELL_NODYNAMICS, CELL_BOUNCEBACK = 0, 1, 2
PARTICLE_RBC, PARTICLE_CAPSULE, PARTICLE_SEMIRIGID = 0, 1, 2
POTENTIAL_NONE, POTENTIAL_MORSE, POTENTIAL_SPRING = 0, 1, 2

_lib = None


class BibError(RuntimeError):
    pass


def load_library():
    """dlopen the in-tree library (building it first if the sources are newer). Raises if it cannot be produced."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    if not _os.environ.get("BIB_LIB"):
        _build.build()
    lib = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
    lib.bib_last_error.restype = C.c_char_p
    for name in SYMBOLS:
        getattr(lib, name)  # AttributeError if the library does not export a declared symbol
    _lib = lib
    return lib


def _check(ok):
    if not ok:
        raise BibError(load_library().bib_last_error().decode())


def _box(b):
    return (C.c_int * 6)(*[int(v) for v in b])


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def device_count() -> int:
    n = C.c_int(0)
    ok = load_library().bib_device_count(C.byref(n))
    return n.value if ok else 0


def set_default_inplace(shift_planes: int) -> None:
    """Population storage of the lattices created from now on: 0 two buffers, >= 2 in place with that shift, -1 automatic."""
    _check(load_library().bib_set_default_inplace(int(shift_planes)))


def inplace_schedule(nx: int, shift_planes: int, frame: int):
    """[(first plane, planes)] of the launches of one in-place step, as the library issues them (host-side; no device needed)."""
    cap = int(nx) + 1
    first, count, n = (C.c_int * cap)(), (C.c_int * cap)(), C.c_int(0)
    _check(load_library().bib_inplace_schedule(int(nx), int(shift_planes), int(frame), cap, first, count, C.byref(n)))
    return [(first[i], count[i]) for i in range(n.value)]


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(load_library().bib_comm_unique_id(buf))
    return buf.raw


class Lattice:
    """One x-slab of a D3Q19 ForcedPhase lattice + its immersed particles, resident on one GPU."""

    def __init__(self, nx, ny, nz, omega, omega_inner, periodic=(True, True, True), device=0, slabs=1, devices=None):
        """omega_inner=None: ForcedD3Q19Descriptor cells (22 scalars) with the single-rate GuoExternalForceBGKdynamics(omega).
        slabs > 1: the lattice is split into equal x-slabs owned by this process (bib_domain_create_multi), on `devices` (default:
        every slab on `device`); the object is then used exactly like a single-slab one (global boxes and particle ids)."""
        self.lib = load_library()
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.h = C.c_void_p()
        per = (C.c_int * 3)(*[int(bool(p)) for p in periodic])
        self.n_ext = 3 if omega_inner is None else 4
        self.ncomp = 19 + self.n_ext
        self.slabs = int(slabs)
        if self.slabs > 1:
            devs = (C.c_int * self.slabs)(*[int(v) for v in (devices if devices is not None else [device] * self.slabs)])
            _check(self.lib.bib_domain_create_multi(self.nx, self.ny, self.nz, self.n_ext, C.c_double(omega), C.c_double(omega_inner or 0.0), per, self.slabs, devs, C.byref(self.h)))
        else:
            _check(self.lib.bib_domain_create(self.nx, self.ny, self.nz, self.n_ext, C.c_double(omega), C.c_double(omega_inner or 0.0), per, int(device), C.byref(self.h)))
        self.full = (0, self.nx - 1, 0, self.ny - 1, 0, self.nz - 1)
        self.shape_nv = {}

    def inplace_shift(self) -> int:
        n = C.c_int(0)
        _check(self.lib.bib_domain_inplace_shift(self.h, C.byref(n)))
        return n.value

    def close(self):
        if self.h:
            self.lib.bib_domain_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- options
    def set_option(self, key, value):
        _check(self.lib.bib_domain_set_option(self.h, int(key), C.c_double(value)))

    # ---- lattice data (Palabos serialised cells [x][y][z][23])
    def upload(self, cells, box=None):
        box = box or self.full
        cells = np.ascontiguousarray(cells, dtype=np.float64)
        assert cells.size == (box[1] - box[0] + 1) * (box[3] - box[2] + 1) * (box[5] - box[4] + 1) * self.ncomp
        _check(self.lib.bib_domain_upload(self.h, _box(box), _dptr(cells)))

    def download(self, box=None, out=None):
        box = box or self.full
        shape = (box[1] - box[0] + 1, box[3] - box[2] + 1, box[5] - box[4] + 1, self.ncomp)
        if out is None:
            out = np.empty(shape, dtype=np.float64)
        _check(self.lib.bib_domain_download(self.h, _box(box), _dptr(out)))
        return out

    def defineDynamics(self, box, kind):
        _check(self.lib.bib_domain_set_celltype(self.h, _box(box), int(kind)))

    def initializeAtEquilibrium(self, rho, u, box=None):
        _check(self.lib.bib_domain_init_equilibrium(self.h, _box(box or self.full), C.c_double(rho), C.c_double(u[0]), C.c_double(u[1]), C.c_double(u[2])))

    def addVelocityBoundary(self, box, direction, orientation, velocities=None):
        """boundaryCondition.addVelocityBoundary{dir}{N|P}(box, lattice) [+ setBoundaryVelocity(lattice, box, velocities)]."""
        _check(self.lib.bib_add_velocity_boundary(self.h, _box(box), int(direction), int(orientation)))
        if velocities is not None:
            self.setBoundaryVelocity(box, velocities)

    def addExternalVelocityEdge(self, box, axis, normal1, normal2):
        """boundaryCondition.addExternalVelocityEdge{axis}{N|P}{N|P}(box, lattice)."""
        _check(self.lib.bib_add_velocity_edge(self.h, _box(box), int(axis), int(normal1), int(normal2)))

    def setBoundaryVelocity(self, box, velocities):
        v = np.ascontiguousarray(velocities, dtype=np.float64)
        _check(self.lib.bib_set_boundary_velocity(self.h, _box(box), _dptr(v)))

    def set_guo_wall_nodes(self, nodes, wall_velocity=None):
        """GuoRigidWallBoundary: nodes [n][7] = {x, y, z, dir_x, dir_y, dir_z, delta} (kit.Oracle.guo_build / the host instantiator)."""
        nodes = np.asarray(nodes, dtype=np.float64).reshape(-1, 7)
        ld = np.ascontiguousarray(nodes[:, :6], dtype=np.int32)
        delta = np.ascontiguousarray(nodes[:, 6], dtype=np.float64)
        wv = None if wall_velocity is None else np.ascontiguousarray(wall_velocity, dtype=np.float64)
        _check(self.lib.bib_set_guo_wall_nodes(self.h, ld.shape[0], ld.ctypes.data_as(C.POINTER(C.c_int)), _dptr(delta), None if wv is None else _dptr(wv)))

    def setExternalVector(self, F, box=None):
        _check(self.lib.bib_set_external_vector(self.h, _box(box or self.full), C.c_double(F[0]), C.c_double(F[1]), C.c_double(F[2])))

    def computeVelocity(self, box=None, download=True):
        box = box or self.full
        _check(self.lib.bib_compute_velocity(self.h, _box(box)))
        if not download:
            return None
        out = np.empty((box[1] - box[0] + 1, box[3] - box[2] + 1, box[5] - box[4] + 1, 3), dtype=np.float64)
        _check(self.lib.bib_velocity_download(self.h, _box(box), _dptr(out)))
        return out

    def collideAndStream(self, n=1):
        for _ in range(n):
            _check(self.lib.bib_collide_and_stream(self.h))

    # ---- immersed boundary
    def register_shape(self, verts, tris) -> int:
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        tris = np.ascontiguousarray(tris, dtype=np.int32)
        sid = C.c_int(-1)
        _check(self.lib.bib_ib_shape_register(self.h, verts.shape[0], tris.shape[0], _dptr(verts), tris.ctypes.data_as(C.POINTER(C.c_int)), C.byref(sid)))
        self.shape_nv[sid.value] = (verts.shape[0], tris.shape[0])
        return sid.value

    def shape_connectivity(self, sid):
        nl, npair = C.c_int(0), C.c_int(0)
        area, vol = C.c_double(0), C.c_double(0)
        _check(self.lib.bib_ib_shape_info(self.h, sid, C.byref(nl), C.byref(npair), C.byref(area), C.byref(vol)))
        nt = self.shape_nv[sid][1]
        links = np.zeros((nl.value, 2), dtype=np.int32)
        link_len = np.zeros(nl.value)
        pairs = np.zeros((npair.value, 4), dtype=np.int32)
        pair_cs = np.zeros((npair.value, 2))
        tri_area = np.zeros(nt)
        ip = C.POINTER(C.c_int)
        _check(self.lib.bib_ib_shape_connectivity(self.h, sid, links.ctypes.data_as(ip), _dptr(link_len), pairs.ctypes.data_as(ip), _dptr(pair_cs), _dptr(tri_area)))
        return dict(links=links, link_len=link_len, pairs=pairs, pair_cs=pair_cs, tri_area=tri_area, area=area.value, volume=vol.value)

    def add_particles(self, sid, params, positions, kind=PARTICLE_RBC):
        nv = self.shape_nv[sid][0]
        positions = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, nv, 3)
        params = np.ascontiguousarray(params, dtype=np.float64)
        assert params.size == 8
        _check(self.lib.bib_ib_particles_add(self.h, int(kind), int(sid), _dptr(params), positions.shape[0], _dptr(positions)))

    def add_rigid_particles(self, sid, inertia, shape_volume, state17):
        """RigidParticle3D: state17 [n][17] = com, velocity, angular velocity, quaternion, density, damping, relaxation, scale."""
        st = np.ascontiguousarray(state17, dtype=np.float64).reshape(-1, 17)
        I = np.ascontiguousarray(inertia, dtype=np.float64).reshape(9)
        _check(self.lib.bib_ib_rigid_particles_add(self.h, int(sid), _dptr(I), C.c_double(shape_volume), st.shape[0], _dptr(st)))

    def rigid_state(self, first=0, count=None):
        if count is None:
            count = self.count()[0] - first
        out = np.empty((count, 19), dtype=np.float64)
        _check(self.lib.bib_ib_rigid_state_download(self.h, int(first), int(count), _dptr(out)))
        return out

    def rigid_state_upload(self, state19, first=0):
        st = np.ascontiguousarray(state19, dtype=np.float64).reshape(-1, 19)
        _check(self.lib.bib_ib_rigid_state_upload(self.h, int(first), st.shape[0], _dptr(st)))

    def set_regulator(self, p, target):
        _check(self.lib.bib_ib_set_regulator(self.h, C.c_double(p), C.c_double(target[0]), C.c_double(target[1]), C.c_double(target[2])))

    def set_boundary(self, y0, y1, enabled=True):
        """fsi.set_boundary(&ParallelPlatesBoundary(y0, y1))"""
        _check(self.lib.bib_ib_set_parallel_plates(self.h, int(enabled), C.c_double(y0), C.c_double(y1)))

    def set_pipe(self, y0, z0, radius, enabled=True):
        """fsi.set_boundary(&PipeBoundary(y0, z0, radius, length))"""
        _check(self.lib.bib_ib_set_pipe(self.h, int(enabled), C.c_double(y0), C.c_double(z0), C.c_double(radius)))

    def set_wall_interaction(self, kind, r_min=0.0, beta=0.0, strength=0.0):
        _check(self.lib.bib_ib_set_wall_interaction(self.h, int(kind), C.c_double(r_min), C.c_double(beta), C.c_double(strength)))

    def set_pp_interaction(self, kind, r_min=0.0, beta=0.0, strength=0.0):
        _check(self.lib.bib_ib_set_pp_interaction(self.h, int(kind), C.c_double(r_min), C.c_double(beta), C.c_double(strength)))

    def count(self):
        n_p, n_v = C.c_int(0), C.c_longlong(0)
        _check(self.lib.bib_ib_count(self.h, C.byref(n_p), C.byref(n_v)))
        return n_p.value, n_v.value

    def local_slots(self):
        """(global ids, roles) of the particle slots on this slab (roles: 0 dead, 1 owned, 2/3 ghost)."""
        cap = max(1, self.count()[0])
        gids = np.zeros(cap, dtype=np.int32); roles = np.zeros(cap, dtype=np.int32)
        n = C.c_int(0)
        ip = C.POINTER(C.c_int)
        _check(self.lib.bib_ib_local_slots(self.h, gids.ctypes.data_as(ip), roles.ctypes.data_as(ip), cap, C.byref(n)))
        return gids[:n.value], roles[:n.value]

    def velocity_extrapolate_plane(self, direction, coord, inward):
        """CouetteVelocityExtrapolator on the device's velocity field: u[coord] = 2 u[coord + inward] - u[coord + 2 inward]."""
        _check(self.lib.bib_velocity_extrapolate_plane(self.h, int(direction), int(coord), int(inward)))

    def velocity(self, box=None):
        box = box or self.full
        out = np.empty((box[1] - box[0] + 1, box[3] - box[2] + 1, box[5] - box[4] + 1, 3), dtype=np.float64)
        _check(self.lib.bib_velocity_download(self.h, _box(box), _dptr(out)))
        return out

    def interpolate_velocity(self):
        _check(self.lib.bib_ib_interpolate_velocity(self.h))

    def compute_and_spread_forces(self):
        _check(self.lib.bib_ib_compute_and_spread_forces(self.h))

    def move_vertices_and_revoxelize(self):
        _check(self.lib.bib_ib_move_vertices_and_revoxelize(self.h))

    def multi_direct_forcing(self, relax_iterations):
        _check(self.lib.bib_ib_multi_direct_forcing(self.h, int(relax_iterations)))

    def ib_step(self):
        _check(self.lib.bib_ib_step(self.h))

    def vertices(self, first=0, count=None, out=None):
        """fsi::Vertex records {pos, vel, force}; `out` may be a preallocated (ideally page-locked) float64 array."""
        if count is None:
            count = self.count()[1] - first
        if out is None:
            out = np.empty((count, 9), dtype=np.float64)
        assert out.dtype == np.float64 and out.size >= count * 9 and out.flags["C_CONTIGUOUS"]
        _check(self.lib.bib_ib_vertices_download(self.h, C.c_longlong(first), C.c_longlong(count), _dptr(out)))
        return out

    def upload_vertices(self, vertices9, first=0):
        v = np.ascontiguousarray(vertices9, dtype=np.float64).reshape(-1, 9)
        _check(self.lib.bib_ib_vertices_upload(self.h, C.c_longlong(first), C.c_longlong(v.shape[0]), _dptr(v)))

    def preload_vertex_forces(self, forces3, first=0):
        f = np.ascontiguousarray(forces3, dtype=np.float64).reshape(-1, 3)
        _check(self.lib.bib_ib_vertex_forces_preload(self.h, C.c_longlong(first), C.c_longlong(f.shape[0]), _dptr(f)))

    def results_download_async(self, out_vertices, out_scalars):
        """Non-blocking read-back of all vertex records and particle scalars into page-locked arrays; see results_wait()."""
        _check(self.lib.bib_ib_results_download_async(self.h, _dptr(out_vertices), _dptr(out_scalars)))

    def lightweight_download_async(self, out_records):
        """Non-blocking read-back of the lightweight record {id, com, gyration tensor, volume} of every owned particle ([n][11], page-locked)."""
        _check(self.lib.bib_ib_lightweight_download_async(self.h, _dptr(out_records)))

    def results_wait(self):
        _check(self.lib.bib_ib_results_wait(self.h))

    def results_count(self):
        """Particles delivered by the last completed results_download_async (multi-slab: the owned ones, compacted)."""
        n = C.c_int(0)
        _check(self.lib.bib_ib_results_count(self.h, C.byref(n)))
        return n.value

    def particle_scalars(self, first=0, count=None, out=None):
        if count is None:
            count = self.count()[0] - first
        if out is None:
            out = np.empty((count, 11), dtype=np.float64)
        _check(self.lib.bib_ib_particle_scalars_download(self.h, int(first), int(count), _dptr(out)))
        return out

    # ---- whole driver steps
    def step(self, body=(0.0, 0.0, 0.0), n=1):
        _check(self.lib.bib_step(self.h, C.c_double(body[0]), C.c_double(body[1]), C.c_double(body[2]), int(n)))

    # ---- multi-GPU ring
    def comm_init(self, rank, n_ranks, unique_id: bytes):
        _check(self.lib.bib_comm_init(self.h, int(rank), int(n_ranks), C.c_char_p(unique_id)))

    # ---- stream / timing
    def sync(self):
        _check(self.lib.bib_sync(self.h))

    def time_begin(self):
        _check(self.lib.bib_time_begin(self.h))

    def time_end(self) -> float:
        ms = C.c_float(0)
        _check(self.lib.bib_time_end(self.h, C.byref(ms)))
        return ms.value

    PHASES = ["interpolate", "exchange_vel", "force_reset", "forces", "exchange_force", "spread", "move", "exchange_pos", "voxelize",
              "lattice_boundary", "lattice", "halo_wait", "halo", "walls"]

    def phase_timing(self, enable=True):
        _check(self.lib.bib_phase_timing(self.h, int(bool(enable))))

    def phase_times(self) -> dict:
        """Device milliseconds per phase, summed over the steps since the last call (see include/biblbm_b200.h, BIB_PHASE_*)."""
        ms = (C.c_double * 14)()
        _check(self.lib.bib_phase_times(self.h, ms, 14))
        return dict(zip(self.PHASES, [float(v) for v in ms]))

    def launches(self) -> int:
        n = C.c_longlong(0)
        _check(self.lib.bib_kernel_launch_count(self.h, C.byref(n)))
        return n.value

    def device_bytes(self) -> int:
        n = C.c_longlong(0)
        _check(self.lib.bib_device_memory_bytes(self.h, C.byref(n)))
        return n.value
