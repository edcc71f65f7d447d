"""In-place streaming (bib_set_default_inplace): ONE population buffer, a step pushes into the same memory
displaced by `shift` planes along x in launches of shift - 1 planes. The reference streams in place as well (its own way:
BlockLattice3D::bulkStream swaps with the neighbours inside one array, src/atomicBlock/blockLattice3D.hh:500-551); what has to hold
here is that the layout is INVISIBLE: every result is bit-identical to the two-buffer layout, i.e. to the reference.

The parity cases of test_gpu_parity.py / test_gpu_slabs.py are re-run unchanged with the in-place layout as the default of every
lattice they create (goldens of the unmodified reference, the CPU oracle, coupled steps, walls, Guo nodes, cell kinds, step graphs,
several slabs), for shifts 2 (one plane per launch), 5 (ragged last launch) and 64 (wider than the lattice: one launch)."""
import numpy as np
import pytest

import biblbm_b200 as bib
import kit
import test_gpu_parity as P
import test_gpu_slabs as S

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[2, 5])
def inplace(request):
    bib.set_default_inplace(request.param)
    yield request.param
    bib.set_default_inplace(-1)


def test_layout_selection_and_footprint():
    try:
        bib.set_default_inplace(0)
        ab = bib.Lattice(40, 16, 32, 1.0, 0.5)
        bib.set_default_inplace(5)
        ip = bib.Lattice(40, 16, 32, 1.0, 0.5)
        bib.set_default_inplace(-1)
        auto = bib.Lattice(40, 16, 32, 1.0, 0.5)           # fits with two buffers
        assert (ab.inplace_shift(), ip.inplace_shift(), auto.inplace_shift()) == (0, 5, 0)
        plane = 16 * 32
        assert ab.device_bytes() == 42 * plane * ((2 * 19 + 3) * 8 + 1)
        assert ip.device_bytes() == 47 * plane * ((19 + 3) * 8 + 1)
        with pytest.raises(bib.BibError, match="shift"):
            bib.set_default_inplace(1)
    finally:
        bib.set_default_inplace(-1)


@pytest.mark.parametrize("shift", [2, 3, 5, 64])
@pytest.mark.parametrize("per", [(1, 1, 1), (1, 0, 1), (0, 0, 0), (0, 1, 0)])
def test_inplace_lattice_vs_oracle(orc, per, shift):
    """Seven steps (both streaming directions, odd count) on a ragged lattice with random phases and forces: the reference's bits."""
    rng = np.random.default_rng(31)
    nx, ny, nz = 19, 9, 37
    cells = np.zeros((nx, ny, nz, 23))
    cells[..., :19] = 0.01 * rng.normal(size=(nx, ny, nz, 19))
    cells[..., 19:22] = 1e-4 * rng.normal(size=(nx, ny, nz, 3))
    cells[..., 22] = rng.integers(0, 2, size=(nx, ny, nz))
    try:
        bib.set_default_inplace(shift)
        lat = P.make_lattice(cells, per, 1.6, 0.7)
        assert lat.inplace_shift() == shift
        ref = cells
        for s in range(7):
            lat.collideAndStream()
            ref = orc.collide_and_stream(ref, per, 1.6, 0.7)
            if s in (0, 1, 6):
                assert np.array_equal(lat.download(), ref), f"step {s + 1}"
        patch = ref[3:6].copy(); patch[..., :19] *= 0.5     # upload / download address the current frame
        lat.upload(patch, (3, 5, 0, ny - 1, 0, nz - 1))
        ref[3:6] = patch
        lat.collideAndStream(2)
        for _ in range(2):
            ref = orc.collide_and_stream(ref, per, 1.6, 0.7)
        assert np.array_equal(lat.download(), ref)
    finally:
        bib.set_default_inplace(-1)


@pytest.mark.parametrize("name", ["lbm_periodic", "lbm_walls"])
def test_inplace_lattice_vs_golden(inplace, name):
    P.test_collide_and_stream_vs_golden(name)


def test_inplace_cell_kinds(orc, inplace):
    P.test_cell_kinds_vs_oracle(orc)


def test_inplace_velocity_and_external_vector(orc, inplace):
    P.test_velocity_and_external_vector(orc)


def test_inplace_wall_planes_and_edges(orc, inplace):
    P.test_wall_planes_vs_golden()
    P.test_ext_flow_boundaries_bit_exact_vs_oracle(orc)


def test_inplace_guo_wall(orc, inplace):
    P.test_guo_pipe_wall_vs_golden()
    P.test_guo_wall_chained_nodes_follow_the_reference_order(orc)


@pytest.mark.parametrize("name", ["ib_rbc1", "ib_rbc2_pfp"])
def test_inplace_coupled_step_bit_exact(rbc, inplace, name):
    P.test_full_step_deterministic_bit_exact(rbc, name)


def test_inplace_driver_cases_vs_golden(rbc, inplace):
    P.test_rbc_in_shear_vs_golden(rbc)
    P.test_rbc_in_pipe_miniature_vs_golden(rbc)
    P.test_sphere_in_ext_flow_vs_golden(rbc)


def test_inplace_step_graph_and_call_fusion(rbc, inplace):
    P.test_step_graph_equals_plain_launches(rbc)
    P.test_call_fusion_equals_separate_calls(rbc)
    P.test_fused_interpolation_equals_two_kernel_path(rbc, "ib_rbc2_pfp")


@pytest.mark.parametrize("per", [(1, 1, 1), (0, 1, 0), (1, 0, 1)])
def test_inplace_slabs_bit_identical_to_oracle(orc, inplace, per):
    S.test_slab_lattice_bit_identical_to_oracle(orc, 2, per)
    S.test_slab_lattice_bit_identical_to_oracle(orc, 4, per)


def test_inplace_slab_immersed_boundary_and_channel(orc, rbc, inplace):
    S.test_slab_immersed_boundary_vs_oracle(orc, rbc, 2)
    S.test_slab_channel_suspension_vs_oracle(orc, rbc, 3)


def test_inplace_equals_two_buffers_large():
    """Size-independent property at a size the oracle does not reach: 64 x 128 x 256 cells, 6 steps, the default shift (17 planes:
    four launches per step, the last one ragged) against the two-buffer layout - bit-identical populations."""
    out = []
    try:
        for shift in (0, 17):
            bib.set_default_inplace(shift)
            lat = bib.Lattice(64, 128, 256, 1.0, 1.0 / 3.0, periodic=(True, False, True))
            lat.initializeAtEquilibrium(1.0, (0.02, 0.01, -0.01))
            rng = np.random.default_rng(5)
            patch = np.zeros((8, 128, 256, 23)); patch[..., :19] = 0.01 * rng.normal(size=(8, 128, 256, 19)); patch[..., 22] = rng.integers(0, 2, size=(8, 128, 256))
            lat.upload(patch, (28, 35, 0, 127, 0, 255))
            lat.collideAndStream(6)
            out.append(lat.download((0, 63, 0, 127, 0, 255)))
            lat.close()
    finally:
        bib.set_default_inplace(-1)
    assert np.array_equal(out[0], out[1])
