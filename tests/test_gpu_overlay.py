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


def same_file(a, b):
    """Byte equality, except for Palabos' XML block descriptions (*.plb): their <DynamicsDict> maps dynamics NAMES to the numeric
    ids the writing binary happened to assign at static-initialisation time (src/core/dynamicsIdentifiers.h), which differ between
    any two executables; names and everything else must still agree."""
    if filecmp.cmp(a, b, shallow=False):
        return True
    if not str(a).endswith(".plb"):
        return False
    try:
        ta, tb = a.read_text(), b.read_text()
    except UnicodeDecodeError:
        return False
    def norm(t):
        return re.sub(r"(<DynamicsDict>)(.*?)(</DynamicsDict>)", lambda m: m.group(1) + re.sub(r"> \d+ <", "> # <", m.group(2)) + m.group(3), t, flags=re.S)
    return "<DynamicsDict>" in ta and norm(ta) == norm(tb)


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


@pytest.mark.parametrize("inplace", ["0", "17"])
def test_rbc_in_shear_driver_bit_identical(tmp_path, mesh, inplace):
    """progs/rbc_in_shear (BASELINE.json configs[0]) at the driver's own size (97^3, lambda = 8.48, Couette walls, position
    regulator): 31 steps, outputs every 10. With two population buffers and with in-place streaming (BIB_INPLACE: one buffer, a step
    in 7 launches of 16 planes): the layout must be invisible in every output file."""
    _need(B200 / "rbc_in_shear_short_b200", CPU / "rbc_in_shear_short_cpu")
    env = dict(BIB_NT="30", BIB_OUTPUT_INTERVAL="10", BIB_NCASES="1", BIB_MESH=str(mesh), BIB_SHRINK_ITERS="2000")
    d_cpu = tmp_path / "cpu"; d_gpu = tmp_path / "gpu"
    d_cpu.mkdir(); d_gpu.mkdir()
    out_cpu = run_driver(CPU / "rbc_in_shear_short_cpu", d_cpu, env)
    out_gpu = run_driver(B200 / "rbc_in_shear_short_b200", d_gpu, dict(env, BIB_DETERMINISTIC_SPREAD="1", BIB_INPLACE=inplace))
    assert out_gpu == out_cpu
    files = all_files(d_cpu)
    assert files == all_files(d_gpu) and len(files) >= 8
    for f in files:
        assert same_file(d_cpu / f, d_gpu / f), f"{f} differs"


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


def _multi_env(n):
    """BIB_GPUS=n: one process, n x-slabs (overlay/biblbm_overlay/impl_desc.inc); one slab per GPU when the box has n GPUs, otherwise the
    slabs share GPU 0 (same multi-slab protocol, peer copies on one device)."""
    import biblbm_b200 as bib
    have = bib.device_count()
    return dict(BIB_GPUS=str(n), BIB_DEVICES=",".join(str(k if have >= n else 0) for k in range(n)))


def _numeric_files_close(d_ref, d_got, files, tol):
    for f in files:
        ta, tb = (d_ref / f).read_bytes(), (d_got / f).read_bytes()
        if ta == tb:
            continue
        if not ("objsP" in f or f.endswith(".txt") or f.endswith(".csv")):     # VTK (base64 float32) / check-points (raw doubles): last bits may differ
            assert len(ta) == len(tb), f
            continue
        a, b = numbers(ta.decode()), numbers(tb.decode())
        assert a.shape == b.shape, f
        assert np.max(np.abs(a - b)) <= tol * max(1.0, np.max(np.abs(a))), f


@pytest.mark.parametrize("n_slabs", [2, 3])
def test_rbc_in_shear_driver_on_several_slabs(tmp_path, mesh, n_slabs):
    """The UNCHANGED driver on a lattice split into x-slabs (BIB_GPUS=n; the RBC sits on the slab face of the 97-plane lattice: 48 | 49)
    against the CPU reference: the reference's `mpirun -np n` use (progs/rbc_in_pipe/batch.sh). Interpolation sums are split between
    the slabs (different association), so text output agrees to its printed digits and numerically to 1e-9."""
    _need(B200 / "rbc_in_shear_short_b200", CPU / "rbc_in_shear_short_cpu")
    env = dict(BIB_NT="30", BIB_OUTPUT_INTERVAL="10", BIB_NCASES="1", BIB_MESH=str(mesh), BIB_SHRINK_ITERS="2000")
    d_cpu = tmp_path / "cpu"; d_gpu = tmp_path / "gpu"
    d_cpu.mkdir(); d_gpu.mkdir()
    out_cpu = run_driver(CPU / "rbc_in_shear_short_cpu", d_cpu, env)
    out_gpu = run_driver(B200 / "rbc_in_shear_short_b200", d_gpu, dict(env, BIB_DETERMINISTIC_SPREAD="1", **_multi_env(n_slabs)))
    a, b = numbers(out_cpu), numbers(out_gpu)
    assert a.shape == b.shape and np.max(np.abs(a - b)) <= 1e-9 * max(1.0, np.max(np.abs(a)))
    files = all_files(d_cpu)
    assert files == all_files(d_gpu) and len(files) >= 8
    _numeric_files_close(d_cpu, d_gpu, files, 1e-9)


def test_rbcs_in_shear_driver_on_two_slabs(tmp_path):
    """progs/rbcs_in_shear (28 RBCs between plates, wall planes inside the box, NoDynamics margins, velocity extrapolation on the device
    field every step) on two x-slabs against the CPU reference."""
    shape, _, _, _ = kit.load_rbc_fixture()
    _need(B200 / "rbcs_in_shear_short_b200", CPU / "rbcs_in_shear_short_cpu")
    env = dict(BIB_NT="20", BIB_OUTPUT_INTERVAL="10", BIB_SHRINK_ITERS="2000")
    d_cpu = tmp_path / "cpu"; d_gpu = tmp_path / "gpu"
    for d in (d_cpu, d_gpu):
        d.mkdir()
        kit.write_gmsh(d / "sphere_TRImesh_626n_1248el.msh", shape.verts, shape.tris)
    out_cpu = run_driver(CPU / "rbcs_in_shear_short_cpu", d_cpu, env)
    out_gpu = run_driver(B200 / "rbcs_in_shear_short_b200", d_gpu, dict(env, BIB_DETERMINISTIC_SPREAD="1", **_multi_env(2)))
    a, b = numbers(out_cpu), numbers(out_gpu)
    assert a.shape == b.shape and np.max(np.abs(a - b)) <= 1e-9 * max(1.0, np.max(np.abs(a)))
    files = [f for f in all_files(d_cpu) if not f.endswith(".msh") and "profile" not in f]
    assert files == [f for f in all_files(d_gpu) if not f.endswith(".msh") and "profile" not in f]
    _numeric_files_close(d_cpu, d_gpu, files, 1e-9)


def _compare_runs(tmp_path, driver, env, args=(), meshes=(), min_files=4, texts=()):
    _need(B200 / f"{driver}_short_b200", CPU / f"{driver}_short_cpu")
    d_cpu = tmp_path / "cpu"; d_gpu = tmp_path / "gpu"
    for d in (d_cpu, d_gpu):
        d.mkdir()
        for name, (verts, tris) in meshes:
            kit.write_gmsh(d / name, verts, tris)
        for name, text in texts:
            (d / name).write_text(text)
    out_cpu = run_driver(CPU / f"{driver}_short_cpu", d_cpu, env, args)
    out_gpu = run_driver(B200 / f"{driver}_short_b200", d_gpu, dict(env, BIB_DETERMINISTIC_SPREAD="1"), args)
    assert out_gpu == out_cpu
    files = all_files(d_cpu)
    assert files == all_files(d_gpu) and len(files) >= min_files, files
    for f in files:
        if f.endswith(".msh") or "profile" in f or f in [t[0] for t in texts]:      # inputs; wall-clock timer report
            continue
        assert same_file(d_cpu / f, d_gpu / f), f"{f} differs"
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


def test_rigid_particle_in_shear_driver_bit_identical(tmp_path):
    """progs/rigid_particle_in_shear/ellipsoid_in_shear.cpp: three stiff particles (RBC membrane model with create_platelet_params)
    in a 33^3 Couette cell, one run per mesh. The driver's ellipsoid meshes are not in the reference repository; any closed
    triangulated surface will do, the platelet / RBC / capsule sphere meshes are written under the names it expects."""
    shape, _, _, _ = kit.load_rbc_fixture()
    plt = np.load(kit.GOLDEN / "platelet54.npz")
    big = np.load(kit.GOLDEN / "sphere1046.npz")
    meshes = [("ellipsoid_1.5_3_3_TRImesh_74n_144el.msh", (plt["verts"] * np.array([0.75, 1.5, 1.5]), plt["tris"])),
              ("ellipsoid_1.5_3_3_TRImesh_166n_328el.msh", (shape.verts * np.array([0.25, 0.5, 0.5]), shape.tris)),
              ("ellipsoid_1.5_3_3_TRImesh_374n_744el.msh", (big["verts"] * np.array([0.25, 0.5, 0.5]), big["tris"]))]
    files = _compare_runs(tmp_path, "rigid_particle_in_shear", dict(BIB_NT="25"), meshes=meshes, min_files=12)
    assert sum("checkpoint_lbm.dat" in f for f in files) == 3


PIPE_XML = """<?xml version="1.0" ?>
<geometry><diameter>30</diameter><length>24</length></geometry>
<physics><Re_d>1</Re_d></physics>
<particles>
  <rbc><mesh>sphere_TRImesh_626n_1248el.msh</mesh><concentration>0.12</concentration><poisson_ratio>0.995</poisson_ratio>
       <Xi_b>0.006</Xi_b><lambda>1</lambda><Ca>0.1</Ca></rbc>
  <platelet><mesh>sphere_2_2_2_TRImesh_54n_104el.msh</mesh><Ca>0.01</Ca><concentration>0.004</concentration></platelet>
  <wall_interaction><type>morse</type><r_min>2</r_min><beta>2</beta><strength>1e-4</strength></wall_interaction>
  <particle_interaction><type>morse</type><r_min>1</r_min><beta>1</beta><strength>1e-3</strength></particle_interaction>
</particles>
<numerics><max_iterations>40</max_iterations><tau>0.55</tau></numerics>
<checkpoints><folder>chk</folder><start_from_checkpoint>false</start_from_checkpoint><write_interval>40</write_interval></checkpoints>
<output><folder>out</folder><flow_vtk_interval>20</flow_vtk_interval><particle_vtk_interval>20</particle_vtk_interval>
        <particle_lw_interval>10</particle_lw_interval></output>
"""


def test_rbc_in_pipe_driver(tmp_path):
    """progs/rbc_in_pipe (BASELINE.json configs[2]) with its own input format, reduced to a 30-diameter pipe: RBCs + platelets
    (SemiRigidParticle3D) packed by the driver's own host-side growth procedure, Guo off-lattice wall, NoDynamics outside, pipe
    centred at y = z = 0 (negative lattice coordinates, envelope width 2), Morse wall and particle-particle forces. Morse exp() and
    the particle-particle summation order differ from the host in the last bits, so outputs are compared numerically:
    particle data 1e-9, fields 1e-9 absolute (|u| ~ 1e-3)."""
    import base64
    _need(B200 / "rbc_in_pipe_short_b200", CPU / "rbc_in_pipe_short_cpu")
    shape, _, _, _ = kit.load_rbc_fixture()
    plt = np.load(kit.GOLDEN / "platelet54.npz")
    env = dict(BIB_SHRINK_ITERS="2000", BIB_GROW_ITERS="2000")
    dirs = {}
    for name, exe, extra in (("cpu", CPU / "rbc_in_pipe_short_cpu", {}), ("gpu", B200 / "rbc_in_pipe_short_b200", {"BIB_DETERMINISTIC_SPREAD": "1"})):
        d = tmp_path / name
        d.mkdir()
        kit.write_gmsh(d / "sphere_TRImesh_626n_1248el.msh", shape.verts, shape.tris)
        kit.write_gmsh(d / "sphere_2_2_2_TRImesh_54n_104el.msh", plt["verts"], plt["tris"])
        (d / "input.xml").write_text(PIPE_XML)
        run_driver(exe, d, dict(env, **extra), args=("input.xml",))
        dirs[name] = d
    files = [f for f in all_files(dirs["cpu"] / "out")]
    assert files == all_files(dirs["gpu"] / "out") and len(files) >= 20
    checked = 0
    for f in files:
        a, b = dirs["cpu"] / "out" / f, dirs["gpu"] / "out" / f
        if "objsP" in f:
            x, y = numbers(a.read_text()), numbers(b.read_text())
            assert x.shape == y.shape and np.max(np.abs(x - y)) <= 1e-9 * max(1.0, np.max(np.abs(x))), f
            checked += 1
        elif f.endswith("checkpoint_lbm.dat"):
            x, y = np.fromfile(a), np.fromfile(b)
            assert x.shape == y.shape and np.max(np.abs(x - y)) <= 1e-9, f
            checked += 1
        elif f.endswith("checkpoint_fsi/0.plb"):
            assert a.stat().st_size == b.stat().st_size
        elif f.endswith(".vti"):
            def arrays(path):
                out = {}
                for m in re.finditer(r'<DataArray type="(\w+)" Name="(\w+)"[^>]*>\n([^<]*)\n</DataArray>', path.read_text()):
                    raw = m.group(3).strip()
                    n = np.frombuffer(base64.b64decode(raw[:8]), dtype=np.uint32)[0]
                    out[m.group(2)] = np.frombuffer(base64.b64decode(raw[8:])[:n], dtype=np.float32)
                return out
            xa, ya = arrays(a), arrays(b)
            assert set(xa) == set(ya) and "velocity" in xa
            for k in xa:
                assert xa[k].shape == ya[k].shape and np.max(np.abs(xa[k] - ya[k])) <= 1e-7, (f, k)
            checked += 1
    assert checked >= 8


def test_rbc_in_pipe_driver_on_two_slabs(tmp_path):
    """progs/rbc_in_pipe on two x-slabs (BIB_GPUS=2; the reference runs this driver with `mpirun -np 6`, progs/rbc_in_pipe/batch.sh): RBC +
    platelet particle groups, Guo wall nodes, Morse wall and particle-particle forces evaluated across the slab face. A 60-long pipe so
    that a slab (30 / 31 planes) is wider than a particle + 6 cells. Particle tables against the CPU reference to 1e-9."""
    _need(B200 / "rbc_in_pipe_short_b200", CPU / "rbc_in_pipe_short_cpu")
    shape, _, _, _ = kit.load_rbc_fixture()
    plt = np.load(kit.GOLDEN / "platelet54.npz")
    xml = PIPE_XML.replace("<length>24</length>", "<length>60</length>").replace("<max_iterations>40</max_iterations>", "<max_iterations>30</max_iterations>")
    env = dict(BIB_SHRINK_ITERS="2000", BIB_GROW_ITERS="2000")
    dirs = {}
    for name, exe, extra in (("cpu", CPU / "rbc_in_pipe_short_cpu", {}),
                             ("gpu", B200 / "rbc_in_pipe_short_b200", dict(BIB_DETERMINISTIC_SPREAD="1", **_multi_env(2)))):
        d = tmp_path / name
        d.mkdir()
        kit.write_gmsh(d / "sphere_TRImesh_626n_1248el.msh", shape.verts, shape.tris)
        kit.write_gmsh(d / "sphere_2_2_2_TRImesh_54n_104el.msh", plt["verts"], plt["tris"])
        (d / "input.xml").write_text(xml)
        run_driver(exe, d, dict(env, **extra), args=("input.xml",), timeout=1500)
        dirs[name] = d
    files = [f for f in all_files(dirs["cpu"] / "out")]
    assert files == all_files(dirs["gpu"] / "out")
    checked = 0
    for f in files:
        if "objsP" in f:
            x, y = numbers((dirs["cpu"] / "out" / f).read_text()), numbers((dirs["gpu"] / "out" / f).read_text())
            assert x.shape == y.shape and np.max(np.abs(x - y)) <= 1e-9 * max(1.0, np.max(np.abs(x))), f
            checked += 1
    assert checked >= 2


def _vti_arrays(path):
    import base64
    out = {}
    for m in re.finditer(r'<DataArray type="(\w+)" Name="(\w+)"[^>]*>\n([^<]*)\n</DataArray>', path.read_text()):
        raw = m.group(3).strip()
        n = np.frombuffer(base64.b64decode(raw[:8]), dtype=np.uint32)[0]
        out[m.group(2)] = np.frombuffer(base64.b64decode(raw[8:])[:n], dtype=np.float32)
    return out


def _compare_outputs_numeric(d_cpu, d_gpu, nv, min_checked, skip=()):
    """File-by-file numeric comparison of two driver output trees: text tables 1e-9 relative, Palabos lattice checkpoints 1e-9 absolute,
    velocity .vti (float32) 1e-7, particle .vtu (float32 appended data: positions 1e-5, forces 1e-9, connectivity byte-identical)."""
    files = [f for f in all_files(d_cpu) if not f.endswith((".gmsh", ".msh")) and f not in skip]
    assert files == [f for f in all_files(d_gpu) if not f.endswith((".gmsh", ".msh")) and f not in skip], files
    checked = 0
    for f in files:
        a, b = d_cpu / f, d_gpu / f
        if f.endswith((".csv", ".txt")):
            x, y = numbers(a.read_text()), numbers(b.read_text())
            assert x.size > 0 and x.shape == y.shape and np.max(np.abs(x - y)) <= 1e-9 * max(1.0, np.max(np.abs(x))), f
            checked += 1
        elif f.endswith("checkpoint_lbm.dat"):
            x, y = np.fromfile(a), np.fromfile(b)
            assert x.shape == y.shape and x.size > 0 and np.max(np.abs(x - y)) <= 1e-9, f
            checked += 1
        elif f.endswith(".vti"):
            xa, ya = _vti_arrays(a), _vti_arrays(b)
            assert set(xa) == set(ya) and "velocity" in xa
            for k in xa:
                assert xa[k].shape == ya[k].shape and np.max(np.abs(xa[k] - ya[k])) <= 1e-7, (f, k)
            checked += 1
        elif f.endswith(".vtu"):
            # appended raw data: [u32 length][nv x 3 float32] positions, the same for force_tot, then the (integer) cell arrays
            ra, rb = a.read_bytes(), b.read_bytes()
            ka, kb = ra.index(b"<AppendedData"), rb.index(b"<AppendedData")
            assert ra[:ka] == rb[:kb]
            da, db = ra[ra.index(b"_", ka) + 1:], rb[rb.index(b"_", kb) + 1:]
            nb = nv * 3 * 4
            for off, tol in ((0, 1e-5), (4 + nb, 1e-9)):           # positions ~ 50 in float32; forces ~ 1e-4
                x = np.frombuffer(da[off + 4:off + 4 + nb], dtype=np.float32); y = np.frombuffer(db[off + 4:off + 4 + nb], dtype=np.float32)
                assert np.max(np.abs(x - y)) <= tol, (f, off)
            assert da[2 * (4 + nb):] == db[2 * (4 + nb):], f
            checked += 1
        elif "checkpoint_fsi" in f and f.endswith("0.plb"):
            assert a.stat().st_size == b.stat().st_size, f
    assert checked >= min_checked, (checked, files)
    return files


def test_sphere_in_ext_flow_driver(tmp_path):
    """progs/sphere_in_ext_flow (BASELINE.json configs[1]) at the driver's own size: 97^3 ForcedD3Q19 lattice with the single-rate
    GuoExternalForceBGKdynamics, extensionalFlowSetup (velocity planes on the x / y faces, four external velocity edges along z), a
    rigid sphere of radius 8 (RigidParticle3D) coupled by the DRIVER'S OWN MultiDirectForcingFunctional3D (4 sub-iterations), which the
    binding runs as one device call. 21 steps; particle csv every 10, velocity / particle VTK every 10, both checkpoints at the end.
    The quaternion exponential map uses the device's sin / cos, so outputs are compared numerically (1e-9 relative)."""
    _need(B200 / "sphere_in_ext_flow_short_b200", CPU / "sphere_in_ext_flow_short_cpu")
    shape, _, _, _ = kit.load_rbc_fixture()
    c = shape.verts.mean(0)
    verts = (shape.verts - c) * (8.0 / float(np.linalg.norm(shape.verts - c, axis=1).mean()))
    env = dict(BIB_NT="20", BIB_OUTPUT_INTERVAL="10", BIB_MESH="sphere626_r8.gmsh")
    dirs = {}
    for name, exe, extra in (("cpu", CPU / "sphere_in_ext_flow_short_cpu", {}), ("gpu", B200 / "sphere_in_ext_flow_short_b200", {"BIB_DETERMINISTIC_SPREAD": "1"})):
        d = tmp_path / name
        d.mkdir()
        kit.write_gmsh(d / "sphere626_r8.gmsh", verts, shape.tris)
        out = run_driver(exe, d, dict(env, **extra))
        dirs[name] = (d, out)
    x, y = numbers(dirs["cpu"][1]), numbers(dirs["gpu"][1])
    assert x.shape == y.shape and np.max(np.abs(x - y)) <= 1e-9 * max(1.0, np.max(np.abs(x)))
    files = _compare_outputs_numeric(dirs["cpu"][0], dirs["gpu"][0], 626, min_checked=8)
    csv = [f for f in files if f.endswith(".csv")]
    assert len(csv) == 1 and len(files) >= 9
    rows = numbers((dirs["cpu"][0] / csv[0]).read_text()).reshape(-1, 13)
    assert rows.shape[0] == 3 and np.abs(rows[-1, 7:10]).max() > 0           # the coupling force is not trivially zero
    lbm = [f for f in files if f.endswith("checkpoint_lbm.dat")]
    assert len(lbm) == 1 and (dirs["cpu"][0] / lbm[0]).stat().st_size == 97 ** 3 * 22 * 8


DEFORMATION_XML = """<?xml version="1.0" ?>
<forcing><Nrelax>12</Nrelax><force_steps>2</force_steps><min_force>0</min_force><max_force>200e-12</max_force><node_fraction>0.02</node_fraction></forcing>
<units><tau>4</tau><dx>0.5e-6</dx><nu_plasma>1.46e-6</nu_plasma><rho_plasma>1025</rho_plasma></units>
<rbc><shear_modulus>6e-6</shear_modulus><bending_energy>4e-18</bending_energy><poisson_ratio>0.995</poisson_ratio><lambda>1</lambda>
<mesh>sphere_TRImesh_626n_1248el.msh</mesh><deflation_factor>0.59</deflation_factor></rbc>
<output><folder>out</folder></output>
"""


def test_deformation_test_driver(tmp_path):
    """progs/deformation_test (optical-tweezers stretching of one RBC in a fully periodic box at negative coordinates): two features no
    other driver uses -- a ForceDecorator class defined in the DRIVER (OpticalTweezerForce; evaluated by its own host code and handed
    over with bib_ib_vertex_forces_preload) and the call order computeVelocity -> collideAndStream -> setExternalVector -> IB (the
    binding evaluates the promised velocity field and interpolates before the collision). 26 steps, two force levels. stdout and the
    deformation log are identical; the lattice starts at (-34, -17, -17), so device coordinates = global - origin round differently in
    the last bit and the float32 VTK payloads are compared numerically (net vertex forces are sums of cancelling terms)."""
    _need(B200 / "deformation_test_short_b200", CPU / "deformation_test_short_cpu")
    shape, _, _, _ = kit.load_rbc_fixture()
    dirs = {}
    for name, exe, extra in (("cpu", CPU / "deformation_test_short_cpu", {}), ("gpu", B200 / "deformation_test_short_b200", {"BIB_DETERMINISTIC_SPREAD": "1"})):
        d = tmp_path / name
        d.mkdir()
        kit.write_gmsh(d / "sphere_TRImesh_626n_1248el.msh", shape.verts, shape.tris)
        (d / "input.xml").write_text(DEFORMATION_XML)
        dirs[name] = (d, run_driver(exe, d, dict(BIB_SHRINK_ITERS="2000", **extra), args=("input.xml",)))
    assert dirs["cpu"][1] == dirs["gpu"][1]
    files = _compare_outputs_numeric(dirs["cpu"][0], dirs["gpu"][0], 626, min_checked=5, skip=("input.xml",))
    assert filecmp.cmp(dirs["cpu"][0] / "out" / "deformation0.txt", dirs["gpu"][0] / "out" / "deformation0.txt", shallow=False)
    assert sum(f.endswith(".vti") for f in files) == 2 and sum(f.endswith(".vtu") for f in files) == 2
    dx = numbers((dirs["cpu"][0] / "out" / "deformation0.txt").read_text()).reshape(-1, 4)
    assert dx[-1, 2] > dx[0, 2]                          # the cell is being stretched along x
