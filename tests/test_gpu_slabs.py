"""Multi-slab parity on ONE GPU: bib_domain_create_multi with several x-slabs sharing device 0 runs the whole multi-slab protocol
(population halo, ghost particles, partial-sum exchanges, ownership hand-off, pair forces across faces, several particle groups,
x-dependent wall velocity) -- the code path of a multi-GPU run, only the transport differs (peer copies instead of NCCL) -- and is
compared with the CPU ORACLE stepping the undivided lattice. With >= 2 GPUs the same cases also run with one slab per GPU.

Tolerances: north_star (populations 1e-12 per step, positions / velocities 1e-9); the LBM-only cases are bit-exact."""
import numpy as np
import pytest

import biblbm_b200 as bib
import kit

pytestmark = pytest.mark.gpu


def device_sets(slabs):
    sets = [[0] * slabs]
    if bib.device_count() >= 2:
        sets.append([k % bib.device_count() for k in range(slabs)])
    return sets


@pytest.mark.parametrize("slabs", [2, 4])
@pytest.mark.parametrize("per", [(1, 1, 1), (0, 1, 0), (1, 0, 1)])
def test_slab_lattice_bit_identical_to_oracle(orc, slabs, per):
    nxl, ny, nz = 6, 9, 20
    gx = nxl * slabs
    rng = np.random.default_rng(21)
    cells = np.zeros((gx, ny, nz, 23))
    cells[..., :19] = 0.01 * rng.normal(size=(gx, ny, nz, 19))
    cells[..., 19:22] = 1e-4 * rng.normal(size=(gx, ny, nz, 3))
    cells[..., 22] = rng.integers(0, 2, size=(gx, ny, nz))
    ref = cells
    for _ in range(5):
        ref = orc.collide_and_stream(ref, per, 1.5, 0.6)
    for devs in device_sets(slabs):
        lat = bib.Lattice(gx, ny, nz, 1.5, 0.6, periodic=per, slabs=slabs, devices=devs)
        lat.upload(cells)
        assert np.array_equal(lat.download(), cells)
        assert np.array_equal(lat.download((nxl - 2, nxl + 1, 1, 3, 0, nz - 1)), cells[nxl - 2:nxl + 2, 1:4])
        lat.collideAndStream(5)
        assert np.array_equal(lat.download(), ref), (devs, per)
        lat.close()


def flat_rbc(rbc_pos, thin_axis):
    """The biconcave RBC centred at the origin, rotated (properly) so that its thin axis lies along `thin_axis`."""
    base = rbc_pos - rbc_pos.mean(0)
    _, _, vt = np.linalg.svd(base, full_matrices=False)       # rows: principal axes, the last one = thin axis
    R = np.roll(vt[::-1], thin_axis, axis=0)                  # thin axis -> row `thin_axis`
    if np.linalg.det(R) < 0:
        R[(thin_axis + 1) % 3] *= -1.0
    return base @ R.T


@pytest.mark.parametrize("slabs", [2, 3])
def test_slab_immersed_boundary_vs_oracle(orc, rbc, slabs):
    """RBCs straddling a slab face, sitting on the periodic seam and drifting across faces (ownership hand-off) in a periodic box."""
    shape, params, rbc_pos, _ = rbc
    nxl, ny, nz = 24, 24, 22
    gx = nxl * slabs
    per = (1, 1, 1)
    centers = [(nxl - 0.4, 11.7, 10.6), (gx - 1.3, 12.4, 11.2), (10.2, 11.1, 10.9)]
    pos = kit.place_particles(rbc_pos, centers, seed=9)
    body = (4e-5, 0.0, 1e-5)
    u0 = (0.05, 0.0, 0.004)
    steps = 150
    lat = bib.Lattice(gx, ny, nz, 1.0, 1 / 3, periodic=per, slabs=slabs)
    lat.initializeAtEquilibrium(1.0, u0)
    cells = lat.download()
    sid = lat.register_shape(shape.verts, shape.tris)
    lat.add_particles(sid, params, pos)
    assert lat.count() == (len(centers), len(centers) * shape.nv)
    assert np.array_equal(lat.vertices()[:, :3].reshape(pos.shape), pos)
    vert = np.zeros((len(centers), shape.nv, 9)); vert[..., :3] = pos
    state = np.array([orc.particle_update(shape, p) for p in pos])
    assert kit.rel_err(lat.particle_scalars(), state) <= 1e-14
    x_start = state[:, 2].copy()
    ref = cells.copy()
    for _ in range(steps):
        orc.full_step(ref, per, 1.0, 1 / 3, body, shape, params, vert, state)
    lat.step(body, n=steps)
    lat.sync()
    got_v = lat.vertices().reshape(vert.shape)
    got = lat.download()
    assert kit.rel_err(got_v[..., :3], vert[..., :3]) <= 1e-9
    assert kit.rel_err(got_v[..., 3:6], vert[..., 3:6]) <= 1e-9
    assert kit.rel_err(got[..., :19], ref[..., :19]) <= 1e-9
    assert np.array_equal(got[..., 22], ref[..., 22])
    assert kit.rel_err(lat.particle_scalars(), state) <= 1e-9
    # the particles really moved across a face / the seam (hand-off happened)
    moved = np.abs(state[:, 2] - x_start)
    assert moved.max() > 5.0 or (np.floor((state[:, 2] + 0.5) / nxl) != np.floor((x_start + 0.5) / nxl)).any()
    lat.close()


def channel_case(rbc_pos, gx, ny, nz):
    """rbcs_in_channel in miniature (progs/rbcs_in_channel/rbcs_in_channel.cpp:133-184): a dense row of RBCs, thin axis along x, stacked
    along x with near contacts (pair force active) and rims within the range of the wall potential."""
    disk = flat_rbc(rbc_pos, 0)
    thick = disk[:, 0].max() - disk[:, 0].min()
    pitch = thick + 0.55
    n = int(gx / pitch)
    pitch = gx / n
    pos = []
    for k in range(n):
        c = np.array([(k + 0.5) * pitch, 9.05 + 5.0 * (k % 2), 10.4 + 0.7 * (k % 3)])
        pos.append(disk + c)
    return np.array(pos)


@pytest.mark.parametrize("slabs", [2, 3])
def test_slab_channel_suspension_vs_oracle(orc, rbc, slabs):
    """BASELINE configs[3] class: velocity planes at y = 0 / ny-1 (x-dependent wall velocity), ParallelPlatesBoundary, body force, Morse wall
    force and Morse particle-particle force across slab faces, against the oracle on the undivided lattice."""
    shape, params, rbc_pos, _ = rbc
    nxl, ny, nz = 26, 27, 22
    gx = nxl * slabs
    pos = channel_case(rbc_pos, gx, ny, nz)
    n_p = pos.shape[0]
    # pair force and wall force are active in this configuration
    d01 = np.linalg.norm(pos[0][:, None, :] - pos[1][None, :, :], axis=-1).min()
    assert d01 < 1.0 and pos[..., 1].min() < 1.0 and pos[..., 1].min() > 0.0
    body = (4e-5, 0.0, 1e-5)
    wall_morse, pp_morse = (1.0, 2.0, 2e-4), (1.0, 1.5, 3e-4)
    x = np.arange(gx)
    lo = np.zeros((gx, nz, 3)); hi = np.zeros((gx, nz, 3))
    lo[..., 0] = -0.01 * (1.0 + 0.3 * np.sin(2 * np.pi * x / gx))[:, None]          # x-dependent wall velocity (periodic in x)
    hi[..., 0] = 0.012 * (1.0 + 0.2 * np.cos(2 * np.pi * x / gx))[:, None]
    steps = 60
    lat = bib.Lattice(gx, ny, nz, 1.0, 1 / 3, periodic=(1, 0, 1), slabs=slabs)
    lat.initializeAtEquilibrium(1.0, (0.0, 0.0, 0.0))
    cells = lat.download()
    lat.addVelocityBoundary((0, gx - 1, 0, 0, 0, nz - 1), 1, -1, lo)
    lat.addVelocityBoundary((0, gx - 1, ny - 1, ny - 1, 0, nz - 1), 1, +1, hi)
    lat.set_boundary(0.0, float(ny - 1))
    lat.set_wall_interaction(bib.POTENTIAL_MORSE, *wall_morse)
    lat.set_pp_interaction(bib.POTENTIAL_MORSE, *pp_morse)
    sid = lat.register_shape(shape.verts, shape.tris)
    lat.add_particles(sid, params, pos)
    vert = np.zeros((n_p, shape.nv, 9)); vert[..., :3] = pos
    state = np.array([orc.particle_update(shape, p) for p in pos])
    ref = cells.copy()
    orc.set_parallel_plates(1, 0.0, float(ny - 1))
    orc.set_wall_interaction(bib.POTENTIAL_MORSE, *wall_morse)
    orc.set_pp_interaction(bib.POTENTIAL_MORSE, *pp_morse)
    try:
        for _ in range(steps):
            ref, _vel = orc.shear_step(ref, 1.0, 1 / 3, body, lo, hi, shape, params, vert, state)
    finally:
        orc.reset_modes()
    lat.step(body, n=steps)
    lat.sync()
    got_v = lat.vertices().reshape(vert.shape)
    got = lat.download()
    assert kit.rel_err(got_v[..., :3], vert[..., :3]) <= 1e-9
    assert kit.rel_err(got_v[..., 6:], vert[..., 6:]) <= 1e-7          # forces: exp() of the Morse potentials, pair order
    assert kit.rel_err(got[..., :19], ref[..., :19]) <= 1e-9
    assert np.array_equal(got[..., 22], ref[..., 22])
    lat.close()


def test_slab_two_particle_groups_equal_single_slab(rbc):
    """Several particle groups (different parameter sets and particle types) on a multi-slab domain == the same set-up on one slab
    (itself checked against the reference): RBCs + stiffer RBCs + a semi-rigid 'platelet' group, pair force between all of them."""
    shape, params, rbc_pos, _ = rbc
    nxl, ny, nz, slabs = 26, 24, 22, 2
    gx = nxl * slabs
    per = (1, 1, 1)
    pa = kit.place_particles(rbc_pos, [(nxl - 0.8, 11.7, 10.6), (10.2, 11.1, 10.9)], seed=3)
    pb = kit.place_particles(rbc_pos, [(gx - 2.3, 12.4, 11.2)], seed=4)
    small = shape.verts * 0.3
    pc = kit.place_particles(small, [(nxl + 9.0, 5.0, 4.0), (nxl - 1.0, 19.0, 17.0)], seed=5)
    params_b = params.copy(); params_b[0] *= 3.0; params_b[6] *= 2.0
    sparams = np.array([1.0, 2e-3, 2e-3, 0.0, 0, 0, 0, 0])
    body = (5e-5, 0.0, 1e-5)
    out = []
    for n_slabs in (1, slabs):
        lat = bib.Lattice(gx, ny, nz, 1.0, 1 / 3, periodic=per, slabs=n_slabs)
        lat.set_option(bib.OPT_DETERMINISTIC_SPREAD, 1.0)
        lat.initializeAtEquilibrium(1.0, (0.04, 0.0, 0.003))
        sid = lat.register_shape(shape.verts, shape.tris)
        sid2 = lat.register_shape(small, shape.tris)
        lat.add_particles(sid, params, pa)
        lat.add_particles(sid, params_b, pb)
        lat.add_particles(sid2, sparams, pc, kind=bib.PARTICLE_SEMIRIGID)
        lat.set_pp_interaction(bib.POTENTIAL_MORSE, 1.0, 1.5, 3e-4)
        lat.step(body, n=80)
        lat.sync()
        out.append((lat.vertices(), lat.particle_scalars(), lat.download()))
        lat.close()
    assert kit.rel_err(out[1][0][:, :3], out[0][0][:, :3]) <= 1e-10
    assert kit.rel_err(out[1][1], out[0][1]) <= 1e-10
    assert kit.rel_err(out[1][2][..., :19], out[0][2][..., :19]) <= 1e-10
    assert np.array_equal(out[1][2][..., 22], out[0][2][..., 22])
