"""debug helper: run the short rbcs_in_shear driver on CPU (reference) and GPU (binding) and compare the .vti arrays"""
import sys, os, pathlib, subprocess, numpy as np, base64, re
sys.path.insert(0, "tests")
import kit
ROOT = pathlib.Path(".").resolve()
shape, _, _, _ = kit.load_rbc_fixture()
tmp = pathlib.Path("/tmp/ovdbg2"); tmp.mkdir(exist_ok=True)
env = dict(os.environ, BIB_NT="12", BIB_OUTPUT_INTERVAL="10", BIB_SHRINK_ITERS="2000")
for name, exe, extra in (("cpu", ROOT / "oracle/_ref/rbcs_in_shear_short_cpu", {}), ("gpu", ROOT / "overlay/_build/rbcs_in_shear_short_b200", {"BIB_DETERMINISTIC_SPREAD": "1"})):
    d = tmp / name; d.mkdir(exist_ok=True)
    kit.write_gmsh(d / "sphere_TRImesh_626n_1248el.msh", shape.verts, shape.tris)
    r = subprocess.run([str(exe)], cwd=d, env=dict(env, **extra), capture_output=True, text=True)
    print(name, r.returncode, r.stderr[-500:])
def arrays(path):
    txt = open(path).read()
    out = {}
    for m in re.finditer(r'<DataArray type="(\w+)" Name="(\w+)"[^>]*>\n([^<]*)\n</DataArray>', txt):
        raw = m.group(3).strip()
        hdr = base64.b64decode(raw[:8]); n = np.frombuffer(hdr, dtype=np.uint32)[0]
        data = base64.b64decode(raw[8:])
        out[m.group(2)] = np.frombuffer(data[:n], dtype=np.float32)
    return out
for f in ("test_dirac2/vtk000000.vti", "test_dirac2/vtk000010.vti"):
    a, b = arrays(tmp / "cpu" / f), arrays(tmp / "gpu" / f)
    for k in a:
        d = np.nonzero(a[k] != b[k])[0]
        print(f, k, a[k].shape, "ndiff", d.size)
        if d.size:
            nc = 3 if k == "velocity" else 1
            cells = d // nc
            # vti order: x fastest? report raw indices and values
            print("  first", d[:12], a[k][d[:12]], b[k][d[:12]])
            nx, ny, nz = 37, 41, 37
            # palabos writes IndexOrdering::backward: z slowest? try both decodings
            print("  cell idx mod/div: i%37", np.unique(cells % 37)[:40], " (i//37)%41", np.unique((cells // 37) % 41)[:45], " i//(37*41)", np.unique(cells // (37 * 41))[:40])
