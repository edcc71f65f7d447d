"""CPU-side checks of the C++ binding (overlay/): the reference's unchanged drivers link against the C-ABI library only
(no CPU compute path inside the binding), and fail loudly when there is no CUDA device."""
import os
import pathlib
import subprocess

import pytest

import biblbm_b200 as bib
import kit

ROOT = pathlib.Path(__file__).resolve().parent.parent
BUILD = ROOT / "overlay" / "_build"
DRIVERS = ["rbc_in_shear", "capsule_in_shear", "rbcs_in_shear", "rbc_in_pipe", "rigid_particle_in_shear", "deformation_test"]


@pytest.mark.parametrize("driver", DRIVERS)
def test_unchanged_driver_binds_to_the_cabi(driver):
    exe = BUILD / driver
    if not exe.exists():
        pytest.skip("overlay/build_drivers.sh has not run (needs the reference sources)")
    undefined = subprocess.run(["nm", "-D", "--undefined-only", str(exe)], capture_output=True, text=True, check=True).stdout
    used = sorted({line.split()[-1] for line in undefined.splitlines() if " bib_" in line})
    for sym in ("bib_domain_create", "bib_domain_upload", "bib_collide_and_stream", "bib_set_external_vector",
                "bib_ib_interpolate_velocity", "bib_ib_compute_and_spread_forces", "bib_ib_move_vertices_and_revoxelize"):
        assert sym in used, f"{driver} does not call {sym}"
    assert set(used) <= set(bib.SYMBOLS), "the binding uses symbols the header does not declare"
    if driver == "rbcs_in_shear":
        assert "bib_velocity_extrapolate_plane" in used          # CouetteVelocityExtrapolator runs on the device
    if driver == "deformation_test":
        assert "bib_ib_vertex_forces_preload" in used            # the driver's own OpticalTweezerForce is evaluated on the host and handed over
    if driver == "rbc_in_pipe":
        assert "bib_set_guo_wall_nodes" in used and "bib_ib_set_pp_interaction" in used


def test_sphere_in_ext_flow_binds_to_the_cabi():
    """The ForcedD3Q19Descriptor instance of the binding: the driver's own MultiDirectForcingFunctional3D resolves to one device call."""
    exe = BUILD / "sphere_in_ext_flow"
    if not exe.exists():
        pytest.skip("overlay/build_drivers.sh has not run (needs the reference sources)")
    undefined = subprocess.run(["nm", "-D", "--undefined-only", str(exe)], capture_output=True, text=True, check=True).stdout
    used = sorted({line.split()[-1] for line in undefined.splitlines() if " bib_" in line})
    for sym in ("bib_domain_create", "bib_collide_and_stream", "bib_add_velocity_boundary", "bib_add_velocity_edge", "bib_ib_rigid_particles_add",
                "bib_ib_rigid_state_download", "bib_ib_multi_direct_forcing", "bib_compute_velocity"):
        assert sym in used, f"sphere_in_ext_flow does not call {sym}"
    assert set(used) <= set(bib.SYMBOLS)


def test_driver_fails_loudly_without_gpu(tmp_path):
    exe = BUILD / "rbc_in_shear_short_b200"
    if not exe.exists():
        pytest.skip("overlay/build_drivers.sh has not run")
    if bib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    shape, _, _, _ = kit.load_rbc_fixture()
    kit.write_gmsh(tmp_path / "s.msh", shape.verts, shape.tris)
    env = dict(os.environ, BIB_NT="1", BIB_NCASES="1", BIB_MESH=str(tmp_path / "s.msh"), BIB_SHRINK_ITERS="10")
    r = subprocess.run([str(exe)], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stdout + r.stderr)
