"""The reference's own drivers, compiled UNCHANGED against the C++ binding in overlay/ (overlay/build_drivers.sh), running on
the GPU through the C-ABI, against the same driver compiled against the pure reference and run on the host.

Only the run length / mesh file / output cadence are parameterised (environment variables, see build_drivers.sh:patch_params);
the time loop, every call and every output statement are the driver's own. With BIB_DETERMINISTIC_SPREAD=1 the whole run is
bit-identical, so every file the driver writes (lightweight particle data, particle VTK, velocity VTK, Palabos lattice
checkpoint, fsi checkpoint) must be byte-identical and stdout must be equal line by line."""
import filecmp
import os
import pathlib
import re
import subprocess

import numpy as np
import pytest

import kit

ROOT = pathlib.Path(__file__).resolve().parent.parent
B200 = ROOT / "overlay" / "_build"
CPU = ROOT / "oracle" / "_ref"

pytestmark = pytest.mark.gpu


def run_driver(exe, cwd, env_extra, args=(), timeout=900):
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([str(exe), *args], cwd=cwd, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def all_files(root):
    return sorted(str(p.relative_to(root)) for p in pathlib.Path(root).rglob("*") if p.is_file())


def numbers(text):
    return np.array([float(x) for x in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", text)])


@pytest.fixture(scope="module")
def mesh(tmp_path_factory):
    shape, _, _, _ = kit.load_rbc_fixture()
    p = tmp_path_factory.mktemp("mesh") / "sphere626.msh"
    kit.write_gmsh(p, shape.verts, shape.tris)
    return p


def _need(*exes):
    for e in exes:
        if not e.exists():
            pytest.skip(f"{e} not built (overlay/build_drivers.sh needs the reference sources)")


def test_rbc_in_shear_driver_bit_identical(tmp_path, mesh):
    """progs/rbc_in_shear (BASELINE.json configs[0]) at the driver's own size (97^3, lambda = 8.48, Couette walls, position
    regulator): 31 steps, outputs every 10."""
    _need(B200 / "rbc_in_shear_short_b200", CPU / "rbc_in_shear_short_cpu")
    env = dict(BIB_NT="30", BIB_OUTPUT_INTERVAL="10", BIB_NCASES="1", BIB_MESH=str(mesh), BIB_SHRINK_ITERS="2000")
    d_cpu = tmp_path / "cpu"; d_gpu = tmp_path / "gpu"
    d_cpu.mkdir(); d_gpu.mkdir()
    out_cpu = run_driver(CPU / "rbc_in_shear_short_cpu", d_cpu, env)
    out_gpu = run_driver(B200 / "rbc_in_shear_short_b200", d_gpu, dict(env, BIB_DETERMINISTIC_SPREAD="1"))
    assert out_gpu == out_cpu
    files = all_files(d_cpu)
    assert files == all_files(d_gpu) and len(files) >= 8
    for f in files:
        assert filecmp.cmp(d_cpu / f, d_gpu / f, shallow=False), f"{f} differs"


def test_rbc_in_shear_driver_atomic_spread(tmp_path, mesh):
    """Same run with the production (fp64 atomics) spread: particle output within 1e-9 relative."""
    _need(B200 / "rbc_in_shear_short_b200", CPU / "rbc_in_shear_short_cpu")
    env = dict(BIB_NT="100", BIB_OUTPUT_INTERVAL="50", BIB_NCASES="1", BIB_MESH=str(mesh), BIB_SHRINK_ITERS="2000")
    d_cpu = tmp_path / "cpu"; d_gpu = tmp_path / "gpu"
    d_cpu.mkdir(); d_gpu.mkdir()
    run_driver(CPU / "rbc_in_shear_short_cpu", d_cpu, env)
    run_driver(B200 / "rbc_in_shear_short_b200", d_gpu, env)
    files = [f for f in all_files(d_cpu) if "objsP" in f]
    assert len(files) >= 3
    for f in files:
        a, b = numbers((d_cpu / f).read_text()), numbers((d_gpu / f).read_text())
        assert a.shape == b.shape
        assert np.max(np.abs(a - b)) <= 1e-9 * max(1.0, np.max(np.abs(a))), f


def _compare_runs(tmp_path, driver, env, args=(), meshes=(), min_files=4):
    _need(B200 / f"{driver}_short_b200", CPU / f"{driver}_short_cpu")
    d_cpu = tmp_path / "cpu"; d_gpu = tmp_path / "gpu"
    for d in (d_cpu, d_gpu):
        d.mkdir()
        for name, (verts, tris) in meshes:
            kit.write_gmsh(d / name, verts, tris)
    out_cpu = run_driver(CPU / f"{driver}_short_cpu", d_cpu, env, args)
    out_gpu = run_driver(B200 / f"{driver}_short_b200", d_gpu, dict(env, BIB_DETERMINISTIC_SPREAD="1"), args)
    assert out_gpu == out_cpu
    files = all_files(d_cpu)
    assert files == all_files(d_gpu) and len(files) >= min_files, files
    for f in files:
        if f.endswith(".msh") or "profile" in f:      # inputs; wall-clock timer report
            continue
        assert filecmp.cmp(d_cpu / f, d_gpu / f, shallow=False), f"{f} differs"
    return files


def test_capsule_in_shear_driver_bit_identical(tmp_path):
    """progs/capsule_in_shear -N 60 (BASELINE.json configs[1]): neo-Hookean capsule (1046 vertices) in a 61^3 Couette cell."""
    z = np.load(kit.GOLDEN / "sphere1046.npz")
    _compare_runs(tmp_path, "capsule_in_shear", dict(BIB_NT="40", BIB_OUTPUT_INTERVAL="20"), args=("-N", "60"),
                  meshes=[("sphere_r6_TRImesh_1046n_2088el.msh", (z["verts"], z["tris"]))])


def test_rbcs_in_shear_driver_bit_identical(tmp_path):
    """progs/rbcs_in_shear: 28 RBCs between parallel plates (ParallelPlatesBoundary, NoDynamics margin, velocity wall planes
    inside the box, CouetteVelocityExtrapolator applied to the velocity field on the host every step)."""
    shape, _, _, _ = kit.load_rbc_fixture()
    _compare_runs(tmp_path, "rbcs_in_shear", dict(BIB_NT="20", BIB_OUTPUT_INTERVAL="10", BIB_SHRINK_ITERS="2000"),
                  meshes=[("sphere_TRImesh_626n_1248el.msh", (shape.verts, shape.tris))])
