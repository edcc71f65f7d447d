import sys, os, pathlib, subprocess, numpy as np, filecmp
sys.path.insert(0, "tests")
import kit
ROOT = pathlib.Path(".").resolve()
shape, _, _, _ = kit.load_rbc_fixture()
tmp = pathlib.Path("/tmp/ovdbg"); tmp.mkdir(exist_ok=True)
kit.write_gmsh(tmp / "s.msh", shape.verts, shape.tris)
NT = sys.argv[1] if len(sys.argv) > 1 else "30"
env = dict(os.environ, BIB_NT=NT, BIB_OUTPUT_INTERVAL="10", BIB_NCASES="1", BIB_MESH=str(tmp / "s.msh"), BIB_SHRINK_ITERS="2000")
for name, exe, extra in (("cpu", ROOT / "oracle/_ref/rbc_in_shear_short_cpu", {}), ("gpu", ROOT / "overlay/_build/rbc_in_shear_short_b200", {"BIB_DETERMINISTIC_SPREAD": "1"})):
    d = tmp / name; d.mkdir(exist_ok=True)
    r = subprocess.run([str(exe)], cwd=d, env=dict(env, **extra), capture_output=True, text=True)
    print(name, r.returncode, r.stderr[-500:])
files = sorted(str(p.relative_to(tmp / "cpu")) for p in (tmp / "cpu").rglob("*") if p.is_file())
for f in files:
    same = filecmp.cmp(tmp / "cpu" / f, tmp / "gpu" / f, shallow=False)
    print(f, os.path.getsize(tmp / "cpu" / f), "same" if same else "DIFF")
    if not same:
        a = np.fromfile(tmp / "cpu" / f, dtype=np.uint8); b = np.fromfile(tmp / "gpu" / f, dtype=np.uint8)
        print("  sizes", a.size, b.size)
        n = min(a.size, b.size)
        idx = np.nonzero(a[:n] != b[:n])[0]
        print("  first diff byte offsets", idx[:20], "count", idx.size)
        if f.endswith("0.plb"):
            for off in range(0, 64):
                m = (n - off) // 8
                da = a[off:off + m * 8].view(np.float64); db = b[off:off + m * 8].view(np.float64)
                if np.all(np.isfinite(da[:100])) and np.max(np.abs(da[:100])) < 1e6 and np.max(np.abs(da[:100])) > 1e-30:
                    print("  offset", off, "max abs diff", np.max(np.abs(da - db)), "n diff", np.count_nonzero(da != db), "of", m)
                    k = np.nonzero(da != db)[0][:10]; print("  idx", k, (k % 9), da[k], db[k])
                    break
