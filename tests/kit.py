"""Shared test infrastructure: ctypes binding of the CPU oracle (oracle/libbiblbm_oracle.so), a runner for the
reference harness (oracle/_ref/ref_harness, the UNMODIFIED reference built by oracle/build_ref.sh) and loaders for
the committed golden fixtures. Nothing here is imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib
import subprocess
import tempfile

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
GOLDEN = ROOT / "tests" / "golden"
REF_HARNESS = ORACLE_DIR / "_ref" / "ref_harness"

Q = 19
CELL = 23
# D3Q19 directions, latticeBoltzmann/nearestNeighborLattices3D.hh:167-177 of the reference
C19 = np.array([
    [0, 0, 0],
    [-1, 0, 0], [0, -1, 0], [0, 0, -1], [-1, -1, 0], [-1, 1, 0], [-1, 0, -1], [-1, 0, 1], [0, -1, -1], [0, -1, 1],
    [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, -1, 0], [1, 0, 1], [1, 0, -1], [0, 1, 1], [0, 1, -1]], dtype=np.int64)
T19 = np.array([1 / 3] + [1 / 18] * 3 + [1 / 36] * 6 + [1 / 18] * 3 + [1 / 36] * 6)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_u8p = C.POINTER(C.c_uint8)


def _d(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def _i(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_ip)


def build_oracle() -> pathlib.Path:
    so = ORACLE_DIR / "libbiblbm_oracle.so"
    src = ORACLE_DIR / "biblbm_oracle.c"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(ORACLE_DIR), "libbiblbm_oracle.so"], check=True, capture_output=True)
    return so


class Shape:
    """Connectivity of one particle shape (what ParticleShape holds in the reference)."""

    def __init__(self, verts, tris, tri_area, links, link_len, link_tri, pairs, pair_cs, scalars):
        self.verts = np.ascontiguousarray(verts, dtype=np.float64)
        self.tris = np.ascontiguousarray(tris, dtype=np.int32)
        self.tri_area = np.ascontiguousarray(tri_area, dtype=np.float64)
        self.links = np.ascontiguousarray(links, dtype=np.int32)
        self.link_len = np.ascontiguousarray(link_len, dtype=np.float64)
        self.link_tri = np.ascontiguousarray(link_tri, dtype=np.int32)
        self.pairs = np.ascontiguousarray(pairs, dtype=np.int32)
        self.pair_cs = np.ascontiguousarray(pair_cs, dtype=np.float64)
        self.area, self.volume = float(scalars[0]), float(scalars[1])
        self.center = np.array(scalars[2:5], dtype=np.float64)

    nv = property(lambda s: s.verts.shape[0])
    nt = property(lambda s: s.tris.shape[0])
    nl = property(lambda s: s.links.shape[0])
    npairs = property(lambda s: s.pairs.shape[0])


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(str(build_oracle()))
        L = self.lib
        L.orc_roma.restype = C.c_double
        L.orc_roma.argtypes = [C.c_double]
        L.orc_shape_build.restype = C.c_int
        L.orc_classify_vertex.restype = C.c_int

    # ---- lattice
    def collide_and_stream(self, cells, per, omega0, omega1, kind=None, envelope_swap=True):
        nx, ny, nz, _ = cells.shape
        out = np.empty_like(cells)
        p = np.asarray(per, dtype=np.int32)
        k = None if kind is None else np.ascontiguousarray(kind, dtype=np.uint8)
        self.lib.orc_collide_and_stream(_d(cells), _d(out), C.c_int(nx), C.c_int(ny), C.c_int(nz), _i(p),
                                        C.c_double(omega0), C.c_double(omega1),
                                        None if k is None else k.ctypes.data_as(_u8p), C.c_int(int(envelope_swap)))
        return out

    def collide_cell(self, cell, omega0, omega1):
        c = np.array(cell, dtype=np.float64)
        self.lib.orc_collide_cell(_d(c), C.c_double(omega0), C.c_double(omega1))
        return c

    def compute_velocity(self, cells):
        vel = np.empty(cells.shape[:3] + (3,), dtype=np.float64)
        self.lib.orc_compute_velocity(_d(cells), _d(vel), C.c_long(cells.size // CELL))
        return vel

    def set_external_vector(self, cells, F):
        f = np.asarray(F, dtype=np.float64)
        self.lib.orc_set_external_vector(_d(cells), C.c_long(cells.size // CELL), _d(f))

    # ---- IB
    def roma(self, x):
        return self.lib.orc_roma(float(x))

    def classify(self, pos, dims, per):
        p = np.asarray(pos, dtype=np.float64)
        pr = np.asarray(per, dtype=np.int32)
        return self.lib.orc_classify_vertex(_d(p), C.c_int(dims[0]), C.c_int(dims[1]), C.c_int(dims[2]), _i(pr))

    def interpolate(self, vel, per, pos):
        nx, ny, nz, _ = vel.shape
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        out = np.zeros_like(pos)
        pr = np.asarray(per, dtype=np.int32)
        self.lib.orc_interpolate(_d(vel), C.c_int(nx), C.c_int(ny), C.c_int(nz), _i(pr), C.c_long(pos.size // 3), _d(pos), _d(out))
        return out

    def spread(self, cells, per, pos, force):
        nx, ny, nz, _ = cells.shape
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        force = np.ascontiguousarray(force, dtype=np.float64)
        pr = np.asarray(per, dtype=np.int32)
        self.lib.orc_spread(_d(cells), C.c_int(nx), C.c_int(ny), C.c_int(nz), _i(pr), C.c_long(pos.size // 3), _d(pos), _d(force))

    def shape_build(self, verts, tris) -> Shape:
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        tris = np.ascontiguousarray(tris, dtype=np.int32)
        nv, nt = verts.shape[0], tris.shape[0]
        tri_area = np.zeros(nt)
        links = np.zeros((3 * nt, 2), dtype=np.int32)
        link_len = np.zeros(3 * nt)
        link_tri = np.zeros((3 * nt, 2), dtype=np.int32)
        pairs = np.zeros((3 * nt, 4), dtype=np.int32)
        pair_cs = np.zeros((3 * nt, 2))
        scalars = np.zeros(5)
        npairs = C.c_int(0)
        nl = self.lib.orc_shape_build(C.c_int(nv), C.c_int(nt), _d(verts), _i(tris), _d(tri_area), _i(links), _d(link_len),
                                      _i(link_tri), _i(pairs), _d(pair_cs), _d(scalars), C.byref(npairs))
        n_p = npairs.value
        return Shape(verts, tris, tri_area, links[:nl], link_len[:nl], link_tri[:nl], pairs[:n_p], pair_cs[:n_p], scalars)

    def particle_update(self, shape: Shape, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        out = np.zeros(11)
        self.lib.orc_particle_update(C.c_int(shape.nv), C.c_int(shape.nt), _i(shape.tris), _d(pos), _d(out))
        return out

    def rbc_forces(self, shape: Shape, params, pos, area, volume, force=None):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        f = np.zeros_like(pos) if force is None else force
        params = np.ascontiguousarray(params, dtype=np.float64)
        self.lib.orc_rbc_forces(C.c_int(shape.nv), C.c_int(shape.nt), C.c_int(shape.nl), C.c_int(shape.npairs),
                                _i(shape.tris), _d(shape.tri_area), _i(shape.links), _d(shape.link_len),
                                _i(shape.pairs), _d(shape.pair_cs), C.c_double(shape.area), _d(params),
                                _d(pos), C.c_double(area), C.c_double(volume), _d(f))
        return f

    def voxelize(self, cells, per, shape: Shape, pos):
        nx, ny, nz, _ = cells.shape
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, shape.nv, 3)
        pr = np.asarray(per, dtype=np.int32)
        self.lib.orc_voxelize(_d(cells), C.c_int(nx), C.c_int(ny), C.c_int(nz), _i(pr), C.c_int(pos.shape[0]),
                              C.c_int(shape.nv), C.c_int(shape.nt), _i(shape.tris), _d(pos))

    def set_particle_model(self, ptype, rest_mesh=None):
        """0 = RBC (default), 1 = neo-Hookean capsule / 2 = semi-rigid particle whose rest configuration is rest_mesh [nv][3]."""
        self._rest = None if rest_mesh is None else np.ascontiguousarray(rest_mesh, dtype=np.float64)
        self.lib.orc_set_particle_model(C.c_int(ptype), None if self._rest is None else _d(self._rest))

    def set_parallel_plates(self, on, y0=0.0, y1=0.0):
        self.lib.orc_set_parallel_plates(C.c_int(int(on)), C.c_double(y0), C.c_double(y1))

    def set_pipe(self, on, y0=0.0, z0=0.0, radius=0.0):
        """fsi.set_boundary(&PipeBoundary(y0, z0, radius, length))"""
        self.lib.orc_set_pipe(C.c_int(int(on)), C.c_double(y0), C.c_double(z0), C.c_double(radius))

    def set_wall_interaction(self, kind, r_min=0.0, beta=0.0, strength=0.0):
        """WallInteraction force decorator: kind 0 none, 1 MorsePotential(r_min, beta, strength), 2 SpringPotential(r_min, strength)"""
        self.lib.orc_set_wall_interaction(C.c_int(kind), C.c_double(r_min), C.c_double(beta), C.c_double(strength))

    def set_pp_interaction(self, kind, r_min=0.0, beta=0.0, strength=0.0):
        """fsi.add_pp_force(PPPotentialForce<...>): same kinds as set_wall_interaction"""
        self.lib.orc_set_pp_interaction(C.c_int(kind), C.c_double(r_min), C.c_double(beta), C.c_double(strength))

    def pp_forces(self, dims, per, pos, obj):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        obj = np.ascontiguousarray(obj, dtype=np.int32)
        f = np.zeros_like(pos)
        pr = np.asarray(per, dtype=np.int32)
        self.lib.orc_pp_forces(C.c_int(dims[0]), C.c_int(dims[1]), C.c_int(dims[2]), _i(pr), C.c_long(pos.shape[0]), _d(pos), _i(obj), _d(f))
        return f

    def guo_build(self, dims):
        """Wall nodes of the Guo off-lattice condition for the boundary set with set_pipe / set_parallel_plates: [n][7]
        {x, y, z, dir_x, dir_y, dir_z, delta}."""
        self.lib.orc_guo_build.restype = C.c_long
        n = self.lib.orc_guo_build(C.c_int(dims[0]), C.c_int(dims[1]), C.c_int(dims[2]), None, C.c_long(0))
        nodes = np.zeros((n, 7))
        self.lib.orc_guo_build(C.c_int(dims[0]), C.c_int(dims[1]), C.c_int(dims[2]), _d(nodes), C.c_long(n))
        return nodes

    def guo_apply(self, cells, per, nodes, wall_velocity=None, kind=None):
        nx, ny, nz, _ = cells.shape
        pr = np.asarray(per, dtype=np.int32)
        wv = None if wall_velocity is None else np.ascontiguousarray(wall_velocity, dtype=np.float64)
        self.lib.orc_guo_apply(_d(cells), C.c_int(nx), C.c_int(ny), C.c_int(nz), _i(pr), C.c_long(nodes.shape[0]), _d(nodes),
                               None if wv is None else _d(wv),
                               None if kind is None else np.ascontiguousarray(kind, dtype=np.uint8).ctypes.data_as(_u8p))

    def set_rigid(self, rigid_state, inertia, shape_volume):
        """RigidParticle3D model (particle model 3): rigid_state [np][24] is updated in place by ib_step / full_step."""
        self._rs = rigid_state
        self._inertia = np.ascontiguousarray(inertia, dtype=np.float64)
        self.lib.orc_set_rigid(_d(self._rs), _d(self._inertia), C.c_double(shape_volume))

    def rigid_place(self, rest_verts, rs):
        rest = np.ascontiguousarray(rest_verts, dtype=np.float64)
        pos = np.zeros_like(rest)
        self.lib.orc_rigid_place(C.c_int(rest.shape[0]), _d(rest), _d(np.ascontiguousarray(rs, dtype=np.float64)), _d(pos))
        return pos

    def reset_modes(self):
        self.set_parallel_plates(0); self.set_wall_interaction(0); self.set_pp_interaction(0); self.set_particle_model(0)

    def capsule_forces(self, shape: Shape, cparams, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        f = np.zeros_like(pos)
        cp = np.ascontiguousarray(cparams, dtype=np.float64)
        self.lib.orc_capsule_forces(C.c_int(shape.nv), C.c_int(shape.nt), _i(shape.tris), _d(shape.tri_area), _d(shape.verts), _d(cp), _d(pos), _d(f))
        return f

    def wall_plane(self, cells, direction, orient, coord, wallvel, omega):
        nx, ny, nz, _ = cells.shape
        wallvel = np.ascontiguousarray(wallvel, dtype=np.float64)
        self.lib.orc_wall_plane(_d(cells), C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(direction), C.c_int(orient), C.c_int(coord),
                                _d(wallvel), C.c_double(omega))

    # ---- sphere_in_ext_flow: partial velocity planes + z edges + multi-direct-forcing (cells carry a zero phase: 23 scalars per cell)
    def wall_plane_box(self, cells, per, direction, orient, box, bvel, omega):
        nx, ny, nz, _ = cells.shape
        self.lib.orc_wall_plane_box(_d(cells), C.c_int(nx), C.c_int(ny), C.c_int(nz), _i(np.asarray(per, dtype=np.int32)), C.c_int(direction), C.c_int(orient),
                                    _i(np.asarray(box, dtype=np.int32)), _d(bvel), C.c_double(omega))

    def wall_edge(self, cells, per, plane, normal1, normal2, box, bvel, omega):
        nx, ny, nz, _ = cells.shape
        self.lib.orc_wall_edge(_d(cells), C.c_int(nx), C.c_int(ny), C.c_int(nz), _i(np.asarray(per, dtype=np.int32)), C.c_int(plane), C.c_int(normal1),
                               C.c_int(normal2), _i(np.asarray(box, dtype=np.int32)), _d(bvel), C.c_double(omega))

    def mdf_step(self, cells, per, bvel, bmask, n_iter, shape: Shape, vert, state):
        """MultiDirectForcingFunctional3D::process for one rigid particle (set_rigid / set_particle_model(3, rest) beforehand)."""
        nx, ny, nz, _ = cells.shape
        vel = np.empty((nx, ny, nz, 3))
        bm = np.ascontiguousarray(bmask, dtype=np.uint8)
        self.lib.orc_mdf_step(_d(cells), _d(vel), C.c_int(nx), C.c_int(ny), C.c_int(nz), _i(np.asarray(per, dtype=np.int32)), _d(bvel), bm.ctypes.data_as(_u8p),
                              C.c_int(n_iter), C.c_int(shape.nv), C.c_int(shape.nt), _i(shape.tris), _d(vert), _d(state))

    def ext_flow_step(self, cells, bvel, omega, n_iter, shape, vert, state):
        """One time step of progs/sphere_in_ext_flow (sphere_in_ext_flow.cpp:285-292) on the extensionalFlowSetup lattice: multi-direct
        forcing, collideAndStream (x, y walls; z periodic; single relaxation rate), then the level-0 boundary functionals in their
        order of insertion (extensional_flow_setup.h:78-88): planes 0N, 0P, 1N, 1P, edges 2NN, 2PN, 2NP, 2PP. Returns the new cells."""
        per = (0, 0, 1)
        planes, edges, bmask = ext_flow_boundaries(cells.shape[:3])
        self.mdf_step(cells, per, bvel, bmask, n_iter, shape, vert, state)
        out = self.collide_and_stream(cells, per, omega, 0.0, envelope_swap=False)
        for (d, o, box) in planes:
            self.wall_plane_box(out, per, d, o, box, bvel, omega)
        for (n1, n2, box) in edges:
            self.wall_edge(out, per, 2, n1, n2, box, bvel, omega)
        return out

    def wall_velocity_fill(self, vel, direction, coord, wallvel):
        nx, ny, nz, _ = vel.shape
        wallvel = np.ascontiguousarray(wallvel, dtype=np.float64)
        self.lib.orc_wall_velocity_fill(_d(vel), C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(direction), C.c_int(coord), _d(wallvel))

    def ib_step(self, cells, vel, per, shape: Shape, params, vert, state, regulator_p=0.0, regulator_target=(0, 0, 0)):
        nx, ny, nz, _ = cells.shape
        pr = np.asarray(per, dtype=np.int32)
        tgt = np.asarray(regulator_target, dtype=np.float64)
        params = np.ascontiguousarray(params, dtype=np.float64)
        self.lib.orc_ib_step(_d(cells), _d(vel), C.c_int(nx), C.c_int(ny), C.c_int(nz), _i(pr), C.c_int(vert.shape[0]), C.c_int(shape.nv),
                             C.c_int(shape.nt), C.c_int(shape.nl), C.c_int(shape.npairs), _i(shape.tris), _d(shape.tri_area), _i(shape.links),
                             _d(shape.link_len), _i(shape.pairs), _d(shape.pair_cs), C.c_double(shape.area), _d(params), _d(vert), _d(state),
                             C.c_double(regulator_p), _d(tgt))

    def shear_step(self, cells, omega0, omega1, body, u_wall_lo, u_wall_hi, shape=None, params=None, vert=None, state=None,
                   regulator_p=0.0, regulator_target=(0, 0, 0)):
        """One driver step of a Couette cell (x,z periodic, velocity planes at y=0 and y=ny-1: progs/common/shear_flow_setup.h:50-97),
        in the reference's order incl. the level-0 wall functionals after streaming. Returns (new cells, velocity field)."""
        nx, ny, nz, _ = cells.shape
        per = (1, 0, 1)
        vel = self.compute_velocity(cells)
        self.wall_velocity_fill(vel, 1, 0, u_wall_lo)
        self.wall_velocity_fill(vel, 1, ny - 1, u_wall_hi)
        self.set_external_vector(cells, body)
        if vert is not None and vert.shape[0] > 0:
            self.ib_step(cells, vel, per, shape, params, vert, state, regulator_p, regulator_target)
        else:
            cells[..., 22] = 0.0
        out = self.collide_and_stream(cells, per, omega0, omega1)
        self.wall_plane(out, 1, -1, 0, u_wall_lo, omega0)
        self.wall_plane(out, 1, 1, ny - 1, u_wall_hi, omega0)
        return out, vel

    def full_step(self, cells, per, omega0, omega1, body, shape: Shape | None, params, vert, state,
                  regulator_p=0.0, regulator_target=(0, 0, 0), kind=None, envelope_swap=True):
        """One driver step in place. vert: [np][nv][9], state: [np][11]. Returns the velocity field used."""
        nx, ny, nz, _ = cells.shape
        scratch = np.empty_like(cells)
        vel = np.empty((nx, ny, nz, 3))
        pr = np.asarray(per, dtype=np.int32)
        body = np.asarray(body, dtype=np.float64)
        tgt = np.asarray(regulator_target, dtype=np.float64)
        k = None if kind is None else np.ascontiguousarray(kind, dtype=np.uint8)
        n_p = 0 if vert is None else vert.shape[0]
        if n_p == 0:
            z = np.zeros(1)
            zi = np.zeros(4, dtype=np.int32)
            self.lib.orc_full_step(_d(cells), _d(scratch), _d(vel), C.c_int(nx), C.c_int(ny), C.c_int(nz), _i(pr),
                                   C.c_double(omega0), C.c_double(omega1), _d(body),
                                   None if k is None else k.ctypes.data_as(_u8p),
                                   C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0),
                                   _i(zi), _d(z), _i(zi), _d(z), _i(zi), _d(z), C.c_double(0.0), _d(z), _d(z), _d(z),
                                   C.c_double(0.0), _d(tgt), C.c_int(int(envelope_swap)))
        else:
            params = np.ascontiguousarray(params, dtype=np.float64)
            self.lib.orc_full_step(_d(cells), _d(scratch), _d(vel), C.c_int(nx), C.c_int(ny), C.c_int(nz), _i(pr),
                                   C.c_double(omega0), C.c_double(omega1), _d(body),
                                   None if k is None else k.ctypes.data_as(_u8p),
                                   C.c_int(n_p), C.c_int(shape.nv), C.c_int(shape.nt), C.c_int(shape.nl), C.c_int(shape.npairs),
                                   _i(shape.tris), _d(shape.tri_area), _i(shape.links), _d(shape.link_len),
                                   _i(shape.pairs), _d(shape.pair_cs), C.c_double(shape.area), _d(params),
                                   _d(vert), _d(state), C.c_double(regulator_p), _d(tgt), C.c_int(int(envelope_swap)))
        return vel


_ORACLE = None


def oracle() -> Oracle:
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE


# ------------------------------------------------------------------------------------------- reference harness
def have_reference() -> bool:
    return REF_HARNESS.exists() and os.access(REF_HARNESS, os.X_OK)


def write_mesh_bin(path, verts, tris):
    with open(path, "wb") as f:
        np.array([verts.shape[0], tris.shape[0]], dtype=np.int64).tofile(f)
        np.ascontiguousarray(verts, dtype=np.float64).tofile(f)
        np.ascontiguousarray(tris, dtype=np.int64).tofile(f)


def run_reference(cfg: dict, workdir=None, verts=None, tris=None, params=None, particle_pos=None,
                  externals=None, populations=None, timeout=600, rigid_state=None):
    """Run oracle/_ref/ref_harness `run` and load its dumps. Returns dict name -> ndarray (+ 'stdout')."""
    assert have_reference(), "oracle/_ref/ref_harness missing: run oracle/build_ref.sh (needs /root/reference)"
    tmp = None
    if workdir is None:
        tmp = tempfile.TemporaryDirectory()
        workdir = tmp.name
    wd = pathlib.Path(workdir)
    wd.mkdir(parents=True, exist_ok=True)
    cfg = dict(cfg)
    cfg["out"] = str(wd)
    n_p = 0
    nv = 0
    if verts is not None:
        write_mesh_bin(wd / "shape.mesh", verts, tris)
        cfg["mesh"] = str(wd / "shape.mesh")
        nv = verts.shape[0]
    if particle_pos is not None:
        pp = np.ascontiguousarray(particle_pos, dtype=np.float64).reshape(-1, nv, 3)
        n_p = pp.shape[0]
        np.concatenate([np.asarray(params, dtype=np.float64), pp.ravel()]).tofile(wd / "particles.f64")
        cfg["particles"] = str(wd / "particles.f64")
    if rigid_state is not None:        # RigidParticle3D: params = {density, damping, relaxation, scale}; state [np][13] = com, vel, angvel, quaternion
        rs = np.ascontiguousarray(rigid_state, dtype=np.float64).reshape(-1, 13)
        n_p = rs.shape[0]
        np.asarray(params, dtype=np.float64).tofile(wd / "particles.f64")
        rs.tofile(wd / "rigid_state_in.f64")
        cfg["particles"] = str(wd / "particles.f64")
        cfg["rigid_state"] = str(wd / "rigid_state_in.f64")
        cfg["particle_type"] = "rigid"
    if externals is not None:
        np.ascontiguousarray(externals, dtype=np.float64).tofile(wd / "externals.f64")
        cfg["externals"] = str(wd / "externals.f64")
    if populations is not None:
        np.ascontiguousarray(populations, dtype=np.float64).tofile(wd / "populations.f64")
        cfg["populations"] = str(wd / "populations.f64")
    with open(wd / "config.txt", "w") as f:
        for k, v in cfg.items():
            f.write(f"{k}={v}\n")
    r = subprocess.run([str(REF_HARNESS), "run", str(wd / "config.txt")], capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"ref_harness failed: {r.stdout}\n{r.stderr}")
    nx, ny, nz = int(cfg["nx"]), int(cfg["ny"]), int(cfg["nz"])
    out = {"stdout": r.stdout}
    for p in sorted(wd.glob("*.f64")):
        name = p.stem
        if name in ("particles", "externals", "populations", "rigid_state_in"):
            continue
        a = np.fromfile(p)
        if name.startswith("pop_"):
            a = a.reshape(nx, ny, nz, CELL)
        elif name.startswith("vel_"):
            a = a.reshape(nx, ny, nz, 3)
        elif name.startswith("vert_"):
            a = a.reshape(n_p, nv, 9)
        elif name.startswith("part_"):
            a = a.reshape(n_p, 11)
        elif name.startswith("rigid_") and name != "rigid_shape":
            a = a.reshape(n_p, 19)
        out[name] = a
    if tmp is not None:
        tmp.cleanup()
    return out


REF_SPHERE_HARNESS = REF_HARNESS.parent / "ref_sphere_harness"


def ext_flow_boundaries(dims):
    """Boundary boxes of progs/common/extensional_flow_setup.h:66-88 for an (nx, ny, nz) lattice: [(dir, orient, box)] planes in their
    order of insertion, [(normal1, normal2, box)] z-edges likewise, and the mask of boundary cells."""
    nx, ny, nz = dims
    gx, gy, gz = nx - 1, ny - 1, nz - 1
    planes = [(0, -1, (0, 0, 1, gy - 1, 0, gz)), (0, 1, (gx, gx, 1, gy - 1, 0, gz)), (1, -1, (1, gx - 1, 0, 0, 0, gz)), (1, 1, (1, gx - 1, gy, gy, 0, gz))]
    edges = [(-1, -1, (0, 0, 0, 0, 0, gz)), (1, -1, (gx, gx, 0, 0, 0, gz)), (-1, 1, (0, 0, gy, gy, 0, gz)), (1, 1, (gx, gx, gy, gy, 0, gz))]
    bmask = np.zeros(dims, dtype=np.uint8)
    bmask[0, :, :] = 1; bmask[gx, :, :] = 1; bmask[:, 0, :] = 1; bmask[:, gy, :] = 1
    return planes, edges, bmask


def run_reference_sphere(cfg: dict, verts, tris, timeout=600):
    """Run oracle/_ref/ref_sphere_harness (miniature of progs/sphere_in_ext_flow on the unmodified reference) and load its dumps.
    cfg: geo_nx, geo_ny, geo_nz, tau, gamma, relaxation_factor, relax_iterations, density, damping, scale, pos_x, pos_y, pos_z, steps."""
    assert REF_SPHERE_HARNESS.exists(), "oracle/_ref/ref_sphere_harness missing: run oracle/build_ref.sh (needs /root/reference)"
    with tempfile.TemporaryDirectory() as tmp:
        wd = pathlib.Path(tmp)
        cfg = dict(cfg, out=str(wd), mesh=str(wd / "shape.mesh"))
        write_mesh_bin(wd / "shape.mesh", verts, tris)
        with open(wd / "config.txt", "w") as f:
            for k, v in cfg.items():
                f.write(f"{k}={v}\n")
        r = subprocess.run([str(REF_SPHERE_HARNESS), str(wd / "config.txt")], capture_output=True, text=True, timeout=timeout)
        if r.returncode != 0:
            raise RuntimeError(f"ref_sphere_harness failed: {r.stdout}\n{r.stderr}")
        nx, ny, nz = int(cfg["geo_nx"]) + 1, int(cfg["geo_ny"]) + 1, int(cfg["geo_nz"]) + 1
        out = {"stdout": r.stdout}
        for p in sorted(wd.glob("*.f64")):
            a = np.fromfile(p)
            name = p.stem
            if name.startswith("pop_") or name.startswith("mid_"):
                a = a.reshape(nx, ny, nz, 22)
            elif name == "bvel":
                a = a.reshape(nx, ny, nz, 3)
            elif name.startswith("vert_"):
                a = a.reshape(1, verts.shape[0], 9)
            elif name.startswith("part_"):
                a = a.reshape(1, 11)
            elif name.startswith("rigid_") and name != "rigid_shape":
                a = a.reshape(1, 19)
            out[name] = a
        return out


# ------------------------------------------------------------------------------------------- golden fixtures
def load_rbc_fixture():
    """The reference's 626-vertex sphere mesh with the connectivity its ParticleShape builds, plus the biconcave RBC
    obtained by the reference's shrink_rbc_volume (tests/golden/make_golden.py)."""
    z = np.load(GOLDEN / "rbc626.npz")
    shape = Shape(z["verts"], z["tris"], z["tri_area"], z["links"], z["link_len"], z["link_tri"], z["pairs"], z["pair_cs"],
                  np.concatenate([z["scalars"][:2], z["scalars"][3:6]]))
    return shape, z["rbc_params"].copy(), z["rbc_pos"].copy(), z


def rotation_matrix(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def place_particles(rbc_pos, centers, seed=0):
    """Randomly oriented copies of one particle at the given centres: [np][nv][3]."""
    rng = np.random.default_rng(seed)
    base = rbc_pos - rbc_pos.mean(0)
    out = []
    for c in centers:
        R = rotation_matrix(rng.normal(size=3), rng.uniform(0, np.pi))
        out.append(base @ R.T + np.asarray(c))
    return np.array(out)


def synthetic_externals(nx, ny, nz, seed=0, amp=1e-5, phase=True):
    """Frozen smooth force field + spherical phase blobs, [x][y][z][4]."""
    x, y, z = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ext = np.zeros((nx, ny, nz, 4))
    ext[..., 0] = amp * np.sin(2 * np.pi * (y / ny + 0.1))
    ext[..., 1] = amp * np.cos(2 * np.pi * (z / nz + 0.2))
    ext[..., 2] = amp * np.sin(2 * np.pi * (x / nx + 0.3)) * np.cos(2 * np.pi * y / ny)
    if phase:
        rng = np.random.default_rng(seed)
        for _ in range(3):
            c = rng.uniform(0, 1, 3) * [nx, ny, nz]
            r = min(nx, ny, nz) * 0.2
            ext[..., 3] = np.maximum(ext[..., 3], ((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2 < r * r) * 1.0)
    return ext


def rel_err(a, b):
    """max |a-b| / max |b| (the 'relative error' of BASELINE.json's tolerances, scale = field magnitude)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    s = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (s if s > 0 else 1.0))


def pipe_outside_kind(dims, y0, z0, radius):
    """Cell kinds of rbc_in_pipe.cpp:314-318: NoDynamics where BoundaryOutside(pipe, std::sqrt(2)) holds -- its margin is truncated
    to plint 1 (reference src/fsi/Boundary.h:113-128), i.e. where (y-y0)^2 + (z-z0)^2 >= (radius + 1)^2."""
    nx, ny, nz = dims
    yy, zz = np.meshgrid(np.arange(ny, dtype=np.float64), np.arange(nz, dtype=np.float64), indexing="ij")
    inside = ((yy - y0) * (yy - y0) + (zz - z0) * (zz - z0)) < (radius - (-1.0)) * (radius - (-1.0))
    return np.ascontiguousarray(np.where(inside, 0, 1).astype(np.uint8)[None].repeat(nx, 0))


def write_gmsh(path, verts, tris):
    """gmsh 2.1 ASCII triangle mesh, the format ParticleShapeLibrary::read_and_store_mesh reads (reference
    src/fsi/ParticleShapeFactory.hh:105-117); %.17g keeps the doubles exact."""
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.1 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % len(verts))
        for i, v in enumerate(verts):
            f.write("%d %.17g %.17g %.17g\n" % (i + 1, v[0], v[1], v[2]))
        f.write("$EndNodes\n$Elements\n%d\n" % len(tris))
        for i, t in enumerate(tris):
            f.write("%d 2 3 0 1 0 %d %d %d\n" % (i + 1, t[0] + 1, t[1] + 1, t[2] + 1))
        f.write("$EndElements\n")
