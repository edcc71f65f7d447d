#!/usr/bin/env python
"""Generates the committed golden fixtures by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref/ref_harness, built by
oracle/build_ref.sh from /root/reference). Run it in the build container (it needs /root/reference):

    oracle/build_ref.sh && python tests/golden/make_golden.py

Fixtures (all float64 unless noted, produced by the reference's own code paths):
  rbc626.npz        the reference's sphere mesh (progs/rbc_in_shear/sphere_TRImesh_626n_1248el.msh) as arrays, the
                    connectivity its ParticleShape builds (links, rest lengths, rest areas, pairs, rest angles), and the
                    biconcave RBC produced by shrink_rbc_volume(0.59 V, 10000) with create_rbc_params (parameters as
                    progs/rbcs_in_shear/rbcs_in_shear.cpp:44-71).
  lbm_periodic.npz  12x10x14 fully periodic box, frozen synthetic force/phase field: populations after 0, 1, 10
                    collideAndStream() calls.
  lbm_walls.npz     same with y non-periodic (free bounce-back at the block faces).
  ib_rbc1.npz       24x22x20 periodic box with one RBC + position regulator: vertex state after steps 1..5, and the
                    lattice / velocity field of step 5 sampled at every 2nd cell (full driver sequence).
  ib_rbc2_pfp.npz   36x22x20 box periodic in x,z, 2 RBCs one of which straddles the periodic x face, 3 steps.
"""
import pathlib
import subprocess
import sys
import tempfile

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import kit  # noqa: E402

MSH = "/root/reference/progs/rbc_in_shear/sphere_TRImesh_626n_1248el.msh"


def harness(*args):
    r = subprocess.run([str(kit.REF_HARNESS), *map(str, args)], capture_output=True, text=True, check=True)
    return r.stdout


def make_rbc_fixture():
    with tempfile.TemporaryDirectory() as d:
        d = pathlib.Path(d)
        harness("mesh2bin", MSH, d / "s.mesh")
        harness("shape", d / "s.mesh", d / "s")
        raw = np.fromfile(d / "s.mesh", dtype=np.int64, count=2)
        nv, nt = int(raw[0]), int(raw[1])
        blob = open(d / "s.mesh", "rb").read()
        verts = np.frombuffer(blob, dtype=np.float64, count=nv * 3, offset=16).reshape(nv, 3)
        tris = np.frombuffer(blob, dtype=np.int64, count=nt * 3, offset=16 + nv * 24).reshape(nt, 3)
        li = np.fromfile(d / "s.links.i64", dtype=np.int64).reshape(-1, 4)
        lf = np.fromfile(d / "s.links.f64")
        ti = np.fromfile(d / "s.tris.i64", dtype=np.int64).reshape(-1, 3)
        tf = np.fromfile(d / "s.tris.f64").reshape(-1, 2)
        pi = np.fromfile(d / "s.pairs.i64", dtype=np.int64).reshape(-1, 4)
        pf = np.fromfile(d / "s.pairs.f64").reshape(-1, 2)
        sc = np.fromfile(d / "s.scalars.f64")  # area, volume, radius, cx, cy, cz
        assert np.array_equal(ti, tris)
        # membrane parameters exactly as progs/rbcs_in_shear/rbcs_in_shear.cpp:44-71
        a_lb, area = sc[2], sc[0]
        tau, Re_p, Ca, k_bend_rel, poisson = 1.0, 0.1, 0.1, 0.006, 0.99
        nu = (tau - 0.5) / 3.0
        gamma = Re_p * nu / (a_lb * a_lb)
        G = a_lb * gamma * nu / Ca
        K_area = (poisson + 1) / (1 - poisson) * G
        k_bend = k_bend_rel * G * area
        harness("deflate", d / "s.mesh", d / "rbc.f64", repr(float(G)), repr(float(K_area)), repr(float(k_bend)), 0.59, 10000)
        r = np.fromfile(d / "rbc.f64")
        assert r[0] > 0 and r[8 + nv * 3 + 1] < 0.7 * sc[1], "deflation did not run with the intended membrane parameters"
        np.savez_compressed(HERE / "rbc626.npz", verts=verts, tris=tris.astype(np.int32), tri_area=tf[:, 0], tri_lmean=tf[:, 1],
                            links=li[:, :2].astype(np.int32), link_tri=li[:, 2:].astype(np.int32), link_len=lf,
                            pairs=pi.astype(np.int32), pair_cs=pf, scalars=sc,
                            rbc_params=r[:8], rbc_pos=r[8:8 + nv * 3].reshape(nv, 3), rbc_area_volume=r[8 + nv * 3:],
                            membrane_inputs=np.array([G, K_area, k_bend]))
        print("rbc626.npz: nv", nv, "nt", nt, "links", len(lf), "pairs", len(pf))


def make_mesh_fixture(msh, out):
    """A reference mesh file as arrays (vertices, triangles), read by the reference's own gmsh reader."""
    with tempfile.TemporaryDirectory() as d:
        d = pathlib.Path(d)
        harness("mesh2bin", msh, d / "s.mesh")
        blob = open(d / "s.mesh", "rb").read()
        nv, nt = (int(v) for v in np.frombuffer(blob, dtype=np.int64, count=2))
        verts = np.frombuffer(blob, dtype=np.float64, count=nv * 3, offset=16).reshape(nv, 3)
        tris = np.frombuffer(blob, dtype=np.int64, count=nt * 3, offset=16 + nv * 24).reshape(nt, 3)
        np.savez_compressed(HERE / out, verts=verts, tris=tris.astype(np.int32))
        print(out, "nv", nv, "nt", nt)


def make_lbm(name, per):
    nx, ny, nz = 12, 10, 14
    ext = kit.synthetic_externals(nx, ny, nz, seed=3)
    out = kit.run_reference(dict(nx=nx, ny=ny, nz=nz, steps=10, dump_every=1, lbm_only=1, periodicity=per,
                                 omega=1.0 / 0.9, omega_inner=1.0 / 2.5), externals=ext)
    np.savez_compressed(HERE / name, pop_init=out["pop_init"], pop_1=out["pop_1"], pop_2=out["pop_2"], pop_10=out["pop_10"],
                        omega=1.0 / 0.9, omega_inner=1.0 / 2.5, per=np.array([c == "p" for c in per], dtype=np.int32))
    print(name, "ok")


def make_ib(name, dims, per, centers, steps, seed, regulator_p=0.0):
    shape, params, rbc_pos, _ = kit.load_rbc_fixture()
    nx, ny, nz = dims
    pos = kit.place_particles(rbc_pos, centers, seed=seed)
    cfg = dict(nx=nx, ny=ny, nz=nz, steps=steps, dump_every=1, periodicity=per, omega=1.0, omega_inner=1.0 / 3.0,
               body_x=1e-6, body_y=0.0, body_z=-2e-6, u0=0.02)
    if regulator_p:
        cfg.update(regulator_p=regulator_p, regulator_x=centers[0][0], regulator_y=centers[0][1], regulator_z=centers[0][2])
    out = kit.run_reference(cfg, verts=shape.verts, tris=shape.tris, params=params, particle_pos=pos)
    keep = dict(pop_init=out["pop_init"], vert_init=out["vert_init"], part_init=out["part_init"],
                pop_final_sample=out[f"pop_{steps}"][::2, ::2, ::2].copy(), vel_final_sample=out[f"vel_{steps}"][::2, ::2, ::2].copy(),
                per=np.array([c == "p" for c in per], dtype=np.int32), body=np.array([1e-6, 0.0, -2e-6]),
                omega=1.0, omega_inner=1.0 / 3.0, steps=steps, regulator_p=regulator_p, regulator_target=np.array(centers[0], dtype=np.float64))
    for s in range(1, steps + 1):
        keep[f"vert_{s}"] = out[f"vert_{s}"]
        keep[f"part_{s}"] = out[f"part_{s}"]
    np.savez_compressed(HERE / name, **keep)
    print(name, "ok")


def make_shear():
    """Couette cell with velocity planes (VelocityDirichlet + PlaneFd, the rbc_in_shear set-up) : LBM only with a frozen
    random force/phase field, and a miniature rbc_in_shear (one RBC + position regulator)."""
    nx, ny, nz, uw = 10, 12, 9, 0.03
    rng = np.random.default_rng(17)
    ext = np.zeros((nx, ny, nz, 4))
    ext[..., :3] = 1e-5 * rng.normal(size=(nx, ny, nz, 3))
    ext[..., 3] = rng.integers(0, 2, size=(nx, ny, nz))
    out = kit.run_reference(dict(nx=nx, ny=ny, nz=nz, steps=10, dump_every=1, lbm_only=1, walls="shear", u_wall=uw, omega=1 / 0.8, omega_inner=0.5),
                            externals=ext)
    np.savez_compressed(HERE / "lbm_shear.npz", pop_init=out["pop_init"], pop_1=out["pop_1"], pop_2=out["pop_2"], pop_10=out["pop_10"],
                        omega=1 / 0.8, omega_inner=0.5, u_wall=uw)
    shape, params, rbc_pos, _ = kit.load_rbc_fixture()
    nx, ny, nz, uw, steps = 26, 24, 22, 0.02, 4
    c = (0.5 * (nx - 1), 0.5 * (ny - 1), 0.5 * (nz - 1))
    pos = kit.place_particles(rbc_pos, [c], seed=3)
    out = kit.run_reference(dict(nx=nx, ny=ny, nz=nz, steps=steps, dump_every=1, walls="shear", u_wall=uw, omega=1.0, omega_inner=1 / (0.5 + 8.48 * 0.5),
                                 regulator_p=1e-2, regulator_x=c[0], regulator_y=c[1], regulator_z=c[2]),
                            verts=shape.verts, tris=shape.tris, params=params, particle_pos=pos)
    keep = dict(pop_init=out["pop_init"], vert_init=out["vert_init"], part_init=out["part_init"], pop_final_sample=out[f"pop_{steps}"][::2, ::2, ::2].copy(),
                vel_final_sample=out[f"vel_{steps}"][::2, ::2, ::2].copy(), omega=1.0, omega_inner=1 / (0.5 + 8.48 * 0.5), u_wall=uw, steps=steps,
                regulator_p=1e-2, regulator_target=np.array(c))
    for s in range(1, steps + 1):
        keep[f"vert_{s}"] = out[f"vert_{s}"]; keep[f"part_{s}"] = out[f"part_{s}"]
    np.savez_compressed(HERE / "ib_rbc_in_shear.npz", **keep)
    print("shear goldens ok")


def make_capsule():
    """Miniature progs/capsule_in_shear: Couette walls + one neo-Hookean capsule (DeformableCapsuleParticle) whose stress-free
    state is the sphere mesh, started from an affinely pre-deformed shape so that the membrane forces are non-zero at once."""
    shape, _, _, _ = kit.load_rbc_fixture()
    nx, ny, nz, uw, steps = 26, 24, 22, 0.02, 4
    c = np.array([0.5 * (nx - 1) + 0.3, 0.5 * (ny - 1) - 0.2, 0.5 * (nz - 1) + 0.1])
    pos = (shape.verts * np.array([1.12, 0.93, 0.97]))[None] + c
    cp = np.array([0.004, 1.0, 0.5, 0.0, 0, 0, 0, 0])           # G, C, nu_s, kb
    out = kit.run_reference(dict(nx=nx, ny=ny, nz=nz, steps=steps, dump_every=1, walls="shear", u_wall=uw, omega=1.0, omega_inner=0.5,
                                 particle_type="capsule"), verts=shape.verts, tris=shape.tris, params=cp, particle_pos=pos)
    keep = dict(pop_init=out["pop_init"], vert_init=out["vert_init"], part_init=out["part_init"], pop_final_sample=out[f"pop_{steps}"][::2, ::2, ::2].copy(),
                omega=1.0, omega_inner=0.5, u_wall=uw, steps=steps, cparams=cp)
    for s in range(1, steps + 1):
        keep[f"vert_{s}"] = out[f"vert_{s}"]; keep[f"part_{s}"] = out[f"part_{s}"]
    np.savez_compressed(HERE / "ib_capsule_in_shear.npz", **keep)
    print("capsule golden ok")


def make_plates():
    """RBC pressed against the lower plate of a Couette cell by a body force, fsi.set_boundary(ParallelPlatesBoundary) active
    (rbcs_in_shear set-up): exercises the near-wall vertex list, the Roma/top-hat blending and the wall reflection."""
    shape, params, rbc_pos, _ = kit.load_rbc_fixture()
    nx, ny, nz, uw, steps = 26, 24, 22, 0.02, 40
    base = rbc_pos - rbc_pos.mean(0)
    flat = np.stack([-base[:, 1], base[:, 0], base[:, 2]], 1)   # proper rotation (a plain axis swap would mirror the mesh)
    pos = (flat + np.array([12.3, -flat[:, 1].min() + 0.004, 10.8]))[None]
    body = (0.0, -1e-4, 0.0)
    out = kit.run_reference(dict(nx=nx, ny=ny, nz=nz, steps=steps, dump_every=10, walls="shear", u_wall=uw, omega=1.0, omega_inner=0.5,
                                 plate_y0=0, plate_y1=ny - 1, body_x=body[0], body_y=body[1], body_z=body[2]),
                            verts=shape.verts, tris=shape.tris, params=params, particle_pos=pos)
    keep = dict(pop_init=out["pop_init"], vert_init=out["vert_init"], part_init=out["part_init"], pop_final_sample=out[f"pop_{steps}"][::2, ::2, ::2].copy(),
                omega=1.0, omega_inner=0.5, u_wall=uw, steps=steps, body=np.array(body), plates=np.array([0.0, ny - 1.0]))
    for s in range(10, steps + 1, 10):
        keep[f"vert_{s}"] = out[f"vert_{s}"]; keep[f"part_{s}"] = out[f"part_{s}"]
    np.savez_compressed(HERE / "ib_rbc_plates.npz", **keep)
    print("plates golden ok")


def make_semirigid():
    """SemiRigidParticle3D (the platelets of rbc_in_pipe) in a periodic box: sphere mesh as rest shape, started from an affinely
    deformed configuration so that radial and link springs are loaded."""
    shape, _, _, _ = kit.load_rbc_fixture()
    nx, ny, nz, steps = 24, 22, 20, 4
    c = np.array([11.3, 10.2, 9.6])
    pos = (shape.verts * np.array([1.1, 0.92, 0.98]))[None] + c
    sp = np.array([float(shape.link_len.mean()), 0.01, 0.02, 0.0, 0, 0, 0, 0])      # l0, k_in_plane, k_out_of_plane, k_bend
    out = kit.run_reference(dict(nx=nx, ny=ny, nz=nz, steps=steps, dump_every=1, periodicity="ppp", omega=1.0, omega_inner=0.5, u0=0.02,
                                 particle_type="semirigid"), verts=shape.verts, tris=shape.tris, params=sp, particle_pos=pos)
    keep = dict(pop_init=out["pop_init"], vert_init=out["vert_init"], part_init=out["part_init"], pop_final_sample=out[f"pop_{steps}"][::2, ::2, ::2].copy(),
                omega=1.0, omega_inner=0.5, steps=steps, sparams=sp, per=np.array([1, 1, 1], dtype=np.int32), body=np.zeros(3), regulator_p=0.0,
                regulator_target=np.zeros(3))
    for s in range(1, steps + 1):
        keep[f"vert_{s}"] = out[f"vert_{s}"]; keep[f"part_{s}"] = out[f"part_{s}"]
    np.savez_compressed(HERE / "ib_semirigid.npz", **keep)
    print("semirigid golden ok, max |F| =", np.abs(out[f"vert_{steps}"][..., 6:]).max())


def make_rigid():
    """RigidParticle3D (M3): a rigid sphere (626-vertex mesh, scale 0.8) with initial translation, spin and tilt in a periodic box with
    a sine flow: direct-forcing coupling + Newton-Euler integration with the quaternion exponential map, 8 steps."""
    shape, _, _, _ = kit.load_rbc_fixture()
    nx, ny, nz, steps = 24, 22, 20, 8
    rp = np.array([1.0, 0.0, 0.45, 0.8, 0, 0, 0, 0])            # density, damping, relaxation_factor, scale
    ang, ax = 0.7, np.array([0.3, -0.5, 0.8]); ax /= np.linalg.norm(ax)
    q = np.array([np.cos(ang / 2), *(np.sin(ang / 2) * ax)])
    rs13 = np.array([[11.3, 10.2, 9.6, 0.004, -0.002, 0.001, 0.002, 0.001, -0.003, *q]])
    out = kit.run_reference(dict(nx=nx, ny=ny, nz=nz, steps=steps, dump_every=1, periodicity="ppp", omega=1.0, omega_inner=0.5, u0=0.02),
                            verts=shape.verts, tris=shape.tris, params=rp, rigid_state=rs13)
    keep = dict(pop_init=out["pop_init"], vert_init=out["vert_init"], part_init=out["part_init"], rigid_init=out["rigid_init"], rigid_shape=out["rigid_shape"],
                pop_final_sample=out[f"pop_{steps}"][::2, ::2, ::2].copy(), omega=1.0, omega_inner=0.5, steps=steps, rparams=rp,
                per=np.array([1, 1, 1], dtype=np.int32), body=np.zeros(3))
    for s in range(1, steps + 1):
        keep[f"vert_{s}"] = out[f"vert_{s}"]; keep[f"part_{s}"] = out[f"part_{s}"]; keep[f"rigid_{s}"] = out[f"rigid_{s}"]
    np.savez_compressed(HERE / "ib_rigid.npz", **keep)
    print("rigid golden ok")


def make_sphere_ext_flow():
    """progs/sphere_in_ext_flow in miniature (oracle/harness/ref_sphere_harness.cpp includes the driver source in place): 25x27x16
    ForcedD3Q19 lattice, GuoExternalForceBGKdynamics, extensionalFlowSetup (velocity planes normal to x and y + four z edges, z periodic),
    one RigidParticle3D sphere (626 vertices, scale 0.6) advanced by the driver's MultiDirectForcingFunctional3D (4 sub-iterations), 10 steps."""
    shape, _, _, _ = kit.load_rbc_fixture()
    cfg = dict(geo_nx=24, geo_ny=26, geo_nz=15, tau=0.55, gamma=2e-3, relaxation_factor=0.3, relax_iterations=4, density=1.0, damping=0.0, scale=0.6,
               pos_x=12.3, pos_y=17.9, pos_z=7.4, steps=10)
    out = kit.run_reference_sphere(cfg, shape.verts, shape.tris)
    steps = cfg["steps"]
    keep = dict(pop_init=out["pop_init"], bvel=out["bvel"], vert_init=out["vert_init"], part_init=out["part_init"], rigid_init=out["rigid_init"],
                rigid_shape=out["rigid_shape"], pop_1=out["pop_1"], pop_final_sample=out[f"pop_{steps}"][::2, ::2, ::2].copy(), omega=1.0 / cfg["tau"],
                steps=steps, relax_iterations=cfg["relax_iterations"],
                rparams=np.array([cfg["density"], cfg["damping"], cfg["relaxation_factor"], cfg["scale"], 0, 0, 0, 0]), per=np.array([0, 0, 1], dtype=np.int32))
    for s in range(1, steps + 1):
        keep[f"vert_{s}"] = out[f"vert_{s}"]; keep[f"part_{s}"] = out[f"part_{s}"]; keep[f"rigid_{s}"] = out[f"rigid_{s}"]
    np.savez_compressed(HERE / "ib_sphere_ext_flow.npz", **keep)
    print("sphere_in_ext_flow golden ok, |F|max =", np.abs(out[f"vert_{steps}"][..., 6:]).max())


def pipe_case():
    """rbc_in_pipe-type set-up in miniature (x periodic; PipeBoundary; Morse wall interaction and Morse particle-particle
    interaction with the parameters of progs/rbc_in_pipe/input.xml): RBC A almost touching the pipe wall with the fluid initially
    moving towards the wall (vertices get reflected), RBC B 0.5 below A (vertex-vertex repulsion), RBC C across the periodic x face."""
    shape, params, rbc_pos, _ = kit.load_rbc_fixture()
    nx, ny, nz = 30, 28, 28
    y0 = z0 = 13.5; radius = 12.2
    base = rbc_pos - rbc_pos.mean(0)
    _, _, vt = np.linalg.svd(base, full_matrices=False)       # principal axes, minor axis last
    R = vt if np.linalg.det(vt) > 0 else vt * np.array([[1], [1], [-1]])
    flat = base @ R.T                                         # minor axis along z, a proper rotation
    def max_r(c):
        p = flat + c
        return np.sqrt((p[:, 1] - y0) ** 2 + (p[:, 2] - z0) ** 2).max()
    lo, hi = 0.0, radius
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        if max_r(np.array([10.2, y0 + 0.3, z0 + mid])) < radius - 0.004: lo = mid
        else: hi = mid
    ca = np.array([10.2, y0 + 0.3, z0 + lo])
    A = flat + ca
    def gap(dz):
        Bp = flat + ca - np.array([0.4, 0.2, dz])
        d = np.sqrt(((A[:, None, :] - Bp[None, :, :]) ** 2).sum(-1))
        return d.min()
    lo, hi = 0.5, 8.0
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if gap(mid) < 0.5: lo = mid
        else: hi = mid
    Bp = flat + ca - np.array([0.4, 0.2, hi])
    Cc = kit.place_particles(rbc_pos, [(29.2, 12.8, 9.0)], seed=5)[0]
    pos = np.stack([A, Bp, Cc])
    cfg = dict(nx=nx, ny=ny, nz=nz, steps=20, dump_every=5, periodicity="pff", omega=1.0, omega_inner=0.5, init="uniform", rho=1.0,
               ux=0.01, uy=0.0, uz=0.04, body_x=1e-5, body_y=0, body_z=0, pipe_y0=y0, pipe_z0=z0, pipe_radius=radius, pipe_length=nx - 1,
               wall_morse="2 2 1e-4", pp_morse="1 1 1e-3")
    return shape, params, pos, cfg


def make_pipe():
    shape, params, pos, cfg = pipe_case()
    out = kit.run_reference(cfg, verts=shape.verts, tris=shape.tris, params=params, particle_pos=pos)
    steps = cfg["steps"]
    keep = dict(pop_init=out["pop_init"], vert_init=out["vert_init"], part_init=out["part_init"], pop_final_sample=out[f"pop_{steps}"][::2, ::2, ::2].copy(),
                omega=1.0, omega_inner=0.5, steps=steps, body=np.array([1e-5, 0.0, 0.0]), per=np.array([1, 0, 0], dtype=np.int32),
                pipe=np.array([cfg["pipe_y0"], cfg["pipe_z0"], cfg["pipe_radius"]]), wall_morse=np.array([2.0, 2.0, 1e-4]), pp_morse=np.array([1.0, 1.0, 1e-3]))
    for s in range(5, steps + 1, 5):
        keep[f"vert_{s}"] = out[f"vert_{s}"]; keep[f"part_{s}"] = out[f"part_{s}"]
    np.savez_compressed(HERE / "ib_pipe_pp.npz", **keep)
    print("pipe golden ok")


def make_guo():
    """Guo off-lattice wall on a pipe (rbc_in_pipe.cpp:285-343: GuoRigidWallBoundary + NoDynamics outside), body-force driven flow
    starting from rest, tau = 0.55 as in progs/rbc_in_pipe/input.xml: LBM only, and the whole rbc_in_pipe set-up in miniature
    (pipe_case + Guo wall)."""
    nx, ny, nz = 12, 24, 24
    y0 = z0 = 11.5; R = 9.3
    cfg = dict(nx=nx, ny=ny, nz=nz, steps=6, dump_every=1, periodicity="pff", omega=1 / 0.55, omega_inner=0.5, init="uniform", rho=1.0,
               ux=0.0, uy=0.0, uz=0.0, body_x=2e-5, body_y=0, body_z=0, pipe_y0=y0, pipe_z0=z0, pipe_radius=R, pipe_length=nx - 1,
               guo_pipe=1, guo_dump=1)
    out = kit.run_reference(cfg)
    np.savez_compressed(HERE / "lbm_guo_pipe.npz", pop_init=out["pop_init"], pop_1=out["pop_1"], pop_2=out["pop_2"], pop_6=out["pop_6"],
                        nodes=out["guo_nodes"].reshape(-1, 7), pipe=np.array([y0, z0, R]), omega=1 / 0.55, omega_inner=0.5, body=np.array([2e-5, 0.0, 0.0]),
                        per=np.array([1, 0, 0], dtype=np.int32))
    shape, params, pos, cfg = pipe_case()
    cfg.update(guo_pipe=1, guo_dump=1, omega=1 / 0.55)
    out = kit.run_reference(cfg, verts=shape.verts, tris=shape.tris, params=params, particle_pos=pos)
    steps = cfg["steps"]
    keep = dict(pop_init=out["pop_init"], vert_init=out["vert_init"], part_init=out["part_init"], pop_final_sample=out[f"pop_{steps}"][::2, ::2, ::2].copy(),
                nodes=out["guo_nodes"].reshape(-1, 7), omega=1 / 0.55, omega_inner=0.5, steps=steps, body=np.array([1e-5, 0.0, 0.0]),
                per=np.array([1, 0, 0], dtype=np.int32), pipe=np.array([cfg["pipe_y0"], cfg["pipe_z0"], cfg["pipe_radius"]]),
                wall_morse=np.array([2.0, 2.0, 1e-4]), pp_morse=np.array([1.0, 1.0, 1e-3]))
    for s in range(5, steps + 1, 5):
        keep[f"vert_{s}"] = out[f"vert_{s}"]; keep[f"part_{s}"] = out[f"part_{s}"]
    np.savez_compressed(HERE / "ib_pipe_guo.npz", **keep)
    print("guo goldens ok")


if __name__ == "__main__":
    assert kit.have_reference(), "build the reference first: oracle/build_ref.sh"
    make_rbc_fixture()
    make_mesh_fixture("/root/reference/progs/capsule_in_shear/sphere_r6_TRImesh_1046n_2088el.msh", "sphere1046.npz")
    make_mesh_fixture("/root/reference/progs/rbc_in_pipe/sphere_2_2_2_TRImesh_54n_104el.msh", "platelet54.npz")
    make_lbm("lbm_periodic.npz", "ppp")
    make_lbm("lbm_walls.npz", "pfp")
    make_ib("ib_rbc1.npz", (24, 22, 20), "ppp", [(11.3, 10.1, 9.7)], steps=5, seed=1, regulator_p=1e-2)
    make_shear()
    make_capsule()
    make_plates()
    make_semirigid()
    make_rigid()
    make_sphere_ext_flow()
    make_pipe()
    make_guo()
    make_ib("ib_rbc2_pfp.npz", (36, 22, 20), "pfp", [(34.6, 10.8, 9.2), (16.2, 11.1, 10.3)], steps=3, seed=2)
