"""Pins the CPU oracle (oracle/biblbm_oracle.c) to the reference: every check is BIT-EXACT against dumps that the
unmodified reference produced (tests/golden/*.npz, made by tests/golden/make_golden.py) and, when the compiled
reference is present (oracle/_ref/ref_harness), against live runs of it on fresh seeded inputs."""
import numpy as np
import pytest

import kit


def test_shape_connectivity_matches_reference(orc, rbc):
    shape, _, _, z = rbc
    s = orc.shape_build(shape.verts, shape.tris)
    assert np.array_equal(s.links, z["links"]) and np.array_equal(s.link_tri, z["link_tri"])
    assert np.array_equal(s.link_len, z["link_len"])
    assert np.array_equal(s.tri_area, z["tri_area"])
    assert np.array_equal(s.pairs, z["pairs"]) and np.array_equal(s.pair_cs, z["pair_cs"])
    assert s.area == z["scalars"][0] and s.volume == z["scalars"][1]
    assert np.array_equal(s.center, z["scalars"][3:6])
    # Euler characteristic of a closed surface: V - E + F = 2
    assert s.nv - s.nl + s.nt == 2 and s.npairs == s.nl


@pytest.mark.parametrize("name", ["lbm_periodic", "lbm_walls"])
def test_collide_and_stream_bit_exact(orc, name):
    g = np.load(kit.GOLDEN / f"{name}.npz")
    c = g["pop_init"].copy()
    for s in range(1, 11):
        c = orc.collide_and_stream(c, g["per"], float(g["omega"]), float(g["omega_inner"]))
        if f"pop_{s}" in g:
            assert np.array_equal(c, g[f"pop_{s}"]), f"{name}: step {s} differs from the reference"


def test_envelope_swap_quirk(orc):
    """The reference collides periodic-envelope copies with (omega, omega_inner) swapped (see biblbm_oracle.c). Without the
    quirk only the cells fed through a periodic face differ; interior cells are identical."""
    g = np.load(kit.GOLDEN / "lbm_periodic.npz")
    c = orc.collide_and_stream(g["pop_1"].copy(), g["per"], float(g["omega"]), float(g["omega_inner"]), envelope_swap=False)
    d = np.abs(c - g["pop_2"])[..., :19].max(-1)
    assert d[1:-1, 1:-1, 1:-1].max() == 0.0
    assert d.max() > 1e-6


def test_mass_and_momentum_conservation(orc):
    """Size-independent property: without force a periodic BGK step conserves sum(f) and sum(c f) to round-off
    (with the reference's envelope quirk switched off: mixing two collisions of one cell breaks exact conservation)."""
    g = np.load(kit.GOLDEN / "lbm_periodic.npz")
    c = g["pop_init"].copy()
    c[..., 19:22] = 0.0
    m0 = c[..., :19].sum()
    j0 = np.einsum("xyzi,id->d", c[..., :19] + kit.T19, kit.C19.astype(float))
    for _ in range(5):
        c = orc.collide_and_stream(c, [1, 1, 1], 1.1, 0.4, envelope_swap=False)
    m1 = c[..., :19].sum()
    j1 = np.einsum("xyzi,id->d", c[..., :19] + kit.T19, kit.C19.astype(float))
    assert abs(m1 - m0) < 1e-12 and np.abs(j1 - j0).max() < 1e-12


@pytest.mark.parametrize("name", ["ib_rbc1", "ib_rbc2_pfp"])
def test_full_step_bit_exact(orc, rbc, name):
    shape, params, _, _ = rbc
    g = np.load(kit.GOLDEN / f"{name}.npz")
    cells, vert, state = g["pop_init"].copy(), g["vert_init"].copy(), g["part_init"].copy()
    assert np.array_equal(orc.particle_update(shape, vert[0, :, :3].copy()), state[0])
    vel = None
    for s in range(1, int(g["steps"]) + 1):
        vel = orc.full_step(cells, g["per"], float(g["omega"]), float(g["omega_inner"]), g["body"], shape, params, vert, state,
                            regulator_p=float(g["regulator_p"]), regulator_target=g["regulator_target"])
        assert np.array_equal(vert, g[f"vert_{s}"]), f"{name}: vertex state differs at step {s}"
        assert np.array_equal(state, g[f"part_{s}"]), f"{name}: particle scalars differ at step {s}"
    assert np.array_equal(cells[::2, ::2, ::2], g["pop_final_sample"])
    assert np.array_equal(vel[::2, ::2, ::2], g["vel_final_sample"])


def test_roma_kernel_properties(orc):
    """Roma 3-point delta (fsi/Dirac.h:73-83): partition of unity and zero first moment on any stencil."""
    rng = np.random.default_rng(0)
    for x in rng.uniform(-3, 3, 200):
        i0 = int(np.ceil(x - 1.5))
        w = np.array([orc.roma(i0 + k - x) for k in range(3)])
        assert abs(w.sum() - 1) < 1e-14
        assert abs((w * (np.arange(3) + i0 - x)).sum()) < 1e-14
        assert abs((w * w).sum() - 0.5) < 1e-14


def test_interpolate_spread_adjoint(orc):
    """<spread(F), u> == <F, interpolate(u)> on a periodic box (same weights both ways)."""
    rng = np.random.default_rng(1)
    dims = (10, 9, 11)
    vel = rng.normal(size=dims + (3,))
    pos = rng.uniform(-2, 13, size=(50, 3))
    F = rng.normal(size=(50, 3))
    cells = np.zeros(dims + (23,))
    orc.spread(cells, [1, 1, 1], pos, F)
    U = orc.interpolate(vel, [1, 1, 1], pos)
    assert abs((cells[..., 19:22] * vel).sum() - (F * U).sum()) < 1e-11


@pytest.mark.reference
@pytest.mark.skipif(not kit.have_reference(), reason="compiled reference (oracle/_ref/ref_harness) not present")
def test_live_reference_multi_rbc(orc, rbc):
    """4 RBCs drifting across the periodic faces of a 40x36x34 box for 60 steps: oracle == live reference, bit for bit."""
    shape, params, rbc_pos, _ = rbc
    dims, per = (40, 36, 34), [1, 1, 1]
    pos = kit.place_particles(rbc_pos, [(39.0, 10, 9), (19, 27.5, 9.3), (20.1, 9.0, 26.2), (1.0, 27.0, 33.0)], seed=5)
    steps, body = 60, [2e-5, 1e-5, -1e-5]
    ref = kit.run_reference(dict(nx=dims[0], ny=dims[1], nz=dims[2], steps=steps, dump_every=20, periodicity="ppp", omega=1.0,
                                 omega_inner=1 / 3, body_x=body[0], body_y=body[1], body_z=body[2], u0=0.04),
                            verts=shape.verts, tris=shape.tris, params=params, particle_pos=pos)
    cells, vert, state = ref["pop_init"].copy(), ref["vert_init"].copy(), ref["part_init"].copy()
    for s in range(1, steps + 1):
        orc.full_step(cells, per, 1.0, 1 / 3, body, shape, params, vert, state)
        if s % 20 == 0:
            assert np.array_equal(vert, ref[f"vert_{s}"]) and np.array_equal(cells, ref[f"pop_{s}"])


@pytest.mark.reference
@pytest.mark.skipif(not kit.have_reference(), reason="compiled reference (oracle/_ref/ref_harness) not present")
def test_live_reference_lbm_random_phase(orc):
    """Random 0/1 phase field + random forces, x,z periodic, y walls (free bounce-back), 17x13x15, 6 steps."""
    rng = np.random.default_rng(7)
    nx, ny, nz = 17, 13, 15
    ext = np.zeros((nx, ny, nz, 4))
    ext[..., :3] = rng.normal(scale=1e-4, size=(nx, ny, nz, 3))
    ext[..., 3] = rng.integers(0, 2, size=(nx, ny, nz))
    ref = kit.run_reference(dict(nx=nx, ny=ny, nz=nz, steps=6, dump_every=3, lbm_only=1, periodicity="pfp", omega=1.7,
                                 omega_inner=0.6, u0=0.08), externals=ext)
    c = ref["pop_init"].copy()
    for s in range(1, 7):
        c = orc.collide_and_stream(c, [1, 0, 1], 1.7, 0.6)
        if s % 3 == 0:
            assert np.array_equal(c, ref[f"pop_{s}"])


def _shear_walls(g):
    nx, ny, nz, _ = g["pop_init"].shape
    uw = float(g["u_wall"])
    lo = np.zeros((nx, nz, 3)); hi = np.zeros((nx, nz, 3))
    lo[..., 0] = -uw; hi[..., 0] = uw       # SimpleShearVelocity, progs/common/shear_flow_setup.h:5-30
    return lo, hi


def test_wall_planes_bit_exact(orc):
    """VelocityDirichlet + PlaneFd velocity planes of the shear drivers, frozen random force / phase field."""
    g = np.load(kit.GOLDEN / "lbm_shear.npz")
    lo, hi = _shear_walls(g)
    ny = g["pop_init"].shape[1]
    c = g["pop_init"].copy()
    for s in range(1, 11):
        c = orc.collide_and_stream(c, [1, 0, 1], float(g["omega"]), float(g["omega_inner"]))
        orc.wall_plane(c, 1, -1, 0, lo, float(g["omega"]))
        orc.wall_plane(c, 1, 1, ny - 1, hi, float(g["omega"]))
        if f"pop_{s}" in g:
            assert np.array_equal(c, g[f"pop_{s}"]), f"step {s}"


def test_rbc_in_shear_bit_exact(orc, rbc):
    """Miniature progs/rbc_in_shear (BASELINE.json configs[0]): Couette walls, one RBC, position regulator, lambda = 8.48."""
    shape, params, _, _ = rbc
    g = np.load(kit.GOLDEN / "ib_rbc_in_shear.npz")
    lo, hi = _shear_walls(g)
    cells, vert, state = g["pop_init"].copy(), g["vert_init"].copy(), g["part_init"].copy()
    vel = None
    for s in range(1, int(g["steps"]) + 1):
        cells, vel = orc.shear_step(cells, float(g["omega"]), float(g["omega_inner"]), (0, 0, 0), lo, hi, shape, params, vert, state,
                                    float(g["regulator_p"]), g["regulator_target"])
        assert np.array_equal(vert, g[f"vert_{s}"]) and np.array_equal(state, g[f"part_{s}"])
    assert np.array_equal(cells[::2, ::2, ::2], g["pop_final_sample"])
    assert np.array_equal(vel[::2, ::2, ::2], g["vel_final_sample"])


def test_capsule_in_shear_bit_exact(orc, rbc):
    """Miniature progs/capsule_in_shear (BASELINE.json configs[1]): neo-Hookean capsule between Couette walls."""
    shape, _, _, _ = rbc
    g = np.load(kit.GOLDEN / "ib_capsule_in_shear.npz")
    lo, hi = _shear_walls(g)
    cells, vert, state = g["pop_init"].copy(), g["vert_init"].copy(), g["part_init"].copy()
    orc.set_particle_model(1, shape.verts)
    try:
        for s in range(1, int(g["steps"]) + 1):
            cells, _ = orc.shear_step(cells, float(g["omega"]), float(g["omega_inner"]), (0, 0, 0), lo, hi, shape, g["cparams"], vert, state)
            assert np.array_equal(vert, g[f"vert_{s}"]) and np.array_equal(state, g[f"part_{s}"])
            assert np.abs(vert[..., 6:]).max() > 1e-4
    finally:
        orc.set_particle_model(0)
    assert np.array_equal(cells[::2, ::2, ::2], g["pop_final_sample"])


def test_parallel_plates_boundary_bit_exact(orc, rbc):
    """fsi.set_boundary(ParallelPlatesBoundary): near-wall list, Roma/top-hat blending, specular reflection (40 steps)."""
    shape, params, _, _ = rbc
    g = np.load(kit.GOLDEN / "ib_rbc_plates.npz")
    lo, hi = _shear_walls(g)
    cells, vert, state = g["pop_init"].copy(), g["vert_init"].copy(), g["part_init"].copy()
    orc.set_parallel_plates(1, float(g["plates"][0]), float(g["plates"][1]))
    try:
        classes = {orc.classify(p, cells.shape[:3], (1, 0, 1)) for p in vert[0, :, :3]}
        assert classes == {1, 3}
        for s in range(1, int(g["steps"]) + 1):
            cells, _ = orc.shear_step(cells, float(g["omega"]), float(g["omega_inner"]), g["body"], lo, hi, shape, params, vert, state)
            if s % 10 == 0:
                assert np.array_equal(vert, g[f"vert_{s}"]) and np.array_equal(state, g[f"part_{s}"]), f"step {s}"
    finally:
        orc.set_parallel_plates(0)
    assert np.array_equal(cells[::2, ::2, ::2], g["pop_final_sample"])


def test_pipe_wall_and_particle_interactions_bit_exact(orc, rbc):
    """rbc_in_pipe-type set-up (tests/golden/make_golden.py:pipe_case): PipeBoundary near-wall blending and reflection, Morse
    WallInteraction decorator and the cell-list particle-particle Morse force, 20 steps; oracle == reference bit for bit."""
    shape, params, _, _ = rbc
    g = np.load(kit.GOLDEN / "ib_pipe_pp.npz")
    cells, vert, state = g["pop_init"].copy(), g["vert_init"].copy(), g["part_init"].copy()
    orc.set_pipe(1, *[float(v) for v in g["pipe"]])
    orc.set_wall_interaction(1, *[float(v) for v in g["wall_morse"]])
    orc.set_pp_interaction(1, *[float(v) for v in g["pp_morse"]])
    try:
        classes = {orc.classify(p, cells.shape[:3], g["per"]) for p in vert[:, :, :3].reshape(-1, 3)}
        assert {1, 2, 3} <= classes
        # the fixture must exercise both interactions at the start
        f_pp = orc.pp_forces(cells.shape[:3], g["per"], vert[..., :3], np.repeat(np.arange(3), shape.nv))
        assert np.abs(f_pp).max() > 1e-6 and abs(f_pp.sum()) < 1e-12
        for s in range(1, int(g["steps"]) + 1):
            orc.full_step(cells, g["per"], float(g["omega"]), float(g["omega_inner"]), g["body"], shape, params, vert, state)
            if s % 5 == 0:
                assert np.array_equal(vert, g[f"vert_{s}"]) and np.array_equal(state, g[f"part_{s}"]), f"step {s}"
    finally:
        orc.reset_modes()
    assert np.array_equal(cells[::2, ::2, ::2], g["pop_final_sample"])


def test_semirigid_particle_bit_exact(orc, rbc):
    """SemiRigidParticle3D (platelets): radial + link springs, 4 coupled steps."""
    shape, _, _, _ = rbc
    g = np.load(kit.GOLDEN / "ib_semirigid.npz")
    cells, vert, state = g["pop_init"].copy(), g["vert_init"].copy(), g["part_init"].copy()
    orc.set_particle_model(2, shape.verts)
    try:
        for s in range(1, int(g["steps"]) + 1):
            orc.full_step(cells, g["per"], float(g["omega"]), float(g["omega_inner"]), g["body"], shape, g["sparams"], vert, state)
            assert np.array_equal(vert, g[f"vert_{s}"]) and np.array_equal(state, g[f"part_{s}"]), f"step {s}"
    finally:
        orc.reset_modes()
    assert np.array_equal(cells[::2, ::2, ::2], g["pop_final_sample"])
    assert np.abs(vert[..., 6:]).max() > 1e-5


def test_rigid_particle_bit_exact(orc, rbc):
    """RigidParticle3D: direct forcing, Newton-Euler, quaternion exponential map; oracle == reference bit for bit."""
    shape, _, _, _ = rbc
    g = np.load(kit.GOLDEN / "ib_rigid.npz")
    cells, vert, state = g["pop_init"].copy(), g["vert_init"].copy(), g["part_init"].copy()
    rs = np.zeros((1, 24)); rs[0, :13] = g["rigid_init"][0, :13]; rs[0, 19:23] = g["rparams"][:4]
    assert np.array_equal(orc.rigid_place(shape.verts, rs[0]), vert[0, :, :3])
    orc.set_particle_model(3, shape.verts)
    orc.set_rigid(rs, g["rigid_shape"][:9], float(g["rigid_shape"][9]))
    try:
        for s in range(1, int(g["steps"]) + 1):
            orc.full_step(cells, g["per"], float(g["omega"]), float(g["omega_inner"]), g["body"], shape, g["rparams"], vert, state)
            assert np.array_equal(vert, g[f"vert_{s}"]) and np.array_equal(state, g[f"part_{s}"]) and np.array_equal(rs[0, :19], g[f"rigid_{s}"][0]), s
    finally:
        orc.reset_modes()
    assert np.array_equal(cells[::2, ::2, ::2], g["pop_final_sample"])


def test_sphere_in_ext_flow_bit_exact(orc, rbc):
    """progs/sphere_in_ext_flow in miniature: single-rate GuoExternalForceBGKdynamics on ForcedD3Q19, velocity planes
    normal to x and y joined by OuterVelocityEdge cells, RigidParticle3D driven by the driver's own MultiDirectForcingFunctional3D
    (4 sub-iterations); oracle == reference bit for bit over 10 steps (lattice, vertices, rigid-body state)."""
    shape, _, _, _ = rbc
    g = np.load(kit.GOLDEN / "ib_sphere_ext_flow.npz")
    cells = np.zeros(g["pop_init"].shape[:3] + (23,)); cells[..., :22] = g["pop_init"]
    bvel = np.ascontiguousarray(g["bvel"])
    vert, state = g["vert_init"][0].copy(), g["part_init"][0].copy()
    rs = np.zeros((1, 24)); rs[0, :13] = g["rigid_init"][0, :13]; rs[0, 19:23] = g["rparams"][:4]
    assert np.array_equal(orc.rigid_place(shape.verts, rs[0]), vert[:, :3])
    orc.set_particle_model(3, shape.verts)
    orc.set_rigid(rs, g["rigid_shape"][:9], float(g["rigid_shape"][9]))
    try:
        for s in range(1, int(g["steps"]) + 1):
            cells = orc.ext_flow_step(cells, bvel, float(g["omega"]), int(g["relax_iterations"]), shape, vert, state)
            assert np.array_equal(vert, g[f"vert_{s}"][0]) and np.array_equal(state, g[f"part_{s}"][0]) and np.array_equal(rs[0, :19], g[f"rigid_{s}"][0]), s
            if s == 1:
                assert np.array_equal(cells[..., :22], g["pop_1"])
    finally:
        orc.reset_modes()
    assert np.abs(vert[:, 6:]).max() > 1e-4, "the forcing must act"
    assert np.array_equal(cells[::2, ::2, ::2, :22], g["pop_final_sample"])


def test_guo_pipe_wall_bit_exact(orc):
    """GuoRigidWallBoundary on a pipe + NoDynamics outside (rbc_in_pipe.cpp:285-343), LBM only: the wall-node list and 6 steps."""
    g = np.load(kit.GOLDEN / "lbm_guo_pipe.npz")
    cells = g["pop_init"].copy()
    dims = cells.shape[:3]
    orc.set_pipe(1, *[float(v) for v in g["pipe"]])
    try:
        nodes = orc.guo_build(dims)
        assert np.array_equal(nodes, g["nodes"]) and len(nodes) > 500
        kind = kit.pipe_outside_kind(dims, *[float(v) for v in g["pipe"]])
        assert kind[0][tuple(nodes[:, 1:3].astype(int).T)].any(), "some wall nodes must be NoDynamics cells (the margin-truncation quirk)"
        for s in range(1, 7):
            orc.full_step(cells, g["per"], float(g["omega"]), float(g["omega_inner"]), g["body"], None, None, None, None, kind=kind)
            orc.guo_apply(cells, g["per"], nodes, kind=kind)
            if f"pop_{s}" in g:
                assert np.array_equal(cells, g[f"pop_{s}"]), s
    finally:
        orc.reset_modes()


def test_rbc_in_pipe_miniature_bit_exact(orc, rbc):
    """The whole rbc_in_pipe set-up in miniature: Guo wall + NoDynamics outside, PipeBoundary for the particles, Morse wall and
    particle-particle interaction, 3 RBCs, 20 steps."""
    shape, params, _, _ = rbc
    g = np.load(kit.GOLDEN / "ib_pipe_guo.npz")
    cells, vert, state = g["pop_init"].copy(), g["vert_init"].copy(), g["part_init"].copy()
    dims = cells.shape[:3]
    orc.set_pipe(1, *[float(v) for v in g["pipe"]])
    orc.set_wall_interaction(1, *[float(v) for v in g["wall_morse"]])
    orc.set_pp_interaction(1, *[float(v) for v in g["pp_morse"]])
    try:
        nodes = orc.guo_build(dims)
        assert np.array_equal(nodes, g["nodes"])
        kind = kit.pipe_outside_kind(dims, *[float(v) for v in g["pipe"]])
        for s in range(1, int(g["steps"]) + 1):
            orc.full_step(cells, g["per"], float(g["omega"]), float(g["omega_inner"]), g["body"], shape, params, vert, state, kind=kind)
            orc.guo_apply(cells, g["per"], nodes, kind=kind)
            if s % 5 == 0:
                assert np.array_equal(vert, g[f"vert_{s}"]) and np.array_equal(state, g[f"part_{s}"]), f"step {s}"
    finally:
        orc.reset_modes()
    assert np.array_equal(cells[::2, ::2, ::2], g["pop_final_sample"])


def test_golden_membrane_forces_are_not_trivial():
    """Guards the fixtures themselves: the RBC goldens must carry real membrane stiffness and non-zero vertex forces."""
    z = np.load(kit.GOLDEN / "rbc626.npz")
    assert z["rbc_params"][0] > 1e-3 and z["rbc_params"][6] > 1e-3 and z["rbc_area_volume"][1] < 0.7 * z["scalars"][1]
    for name in ["ib_rbc1", "ib_rbc2_pfp", "ib_rbc_in_shear", "ib_rbc_plates"]:
        g = np.load(kit.GOLDEN / f"{name}.npz")
        last = max(int(k.split("_")[1]) for k in g.files if k.startswith("vert_") and k.split("_")[1].isdigit())
        assert np.abs(g[f"vert_{last}"][..., 6:]).max() > 1e-8, name
