"""Wall-clock per-step cost of the reference's own rbc_in_shear driver (97^3, 1 RBC): compiled against the B200 binding vs against the
pure reference, by differencing two run lengths (set-up and output cancel). Usage: python tools/driver_speed.py [--cpu]"""
import os, pathlib, subprocess, sys, tempfile, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
import kit


def run(exe, nt, extra=None):
    shape, _, _, _ = kit.load_rbc_fixture()
    with tempfile.TemporaryDirectory() as d:
        kit.write_gmsh(pathlib.Path(d) / "s.msh", shape.verts, shape.tris)
        env = dict(os.environ, BIB_NT=str(nt), BIB_OUTPUT_INTERVAL="1000000", BIB_NCASES="1", BIB_MESH=str(pathlib.Path(d) / "s.msh"), BIB_SHRINK_ITERS="2000")
        env.update(extra or {})
        t0 = time.perf_counter()
        r = subprocess.run([str(exe)], cwd=d, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-500:]
        return time.perf_counter() - t0


def per_step(exe, n1, n2, extra=None):
    t1, t2 = run(exe, n1, extra), run(exe, n2, extra)
    return (t2 - t1) / (n2 - n1), t1, t2


if __name__ == "__main__":
    gpu = ROOT / "overlay" / "_build" / "rbc_in_shear_short_b200"
    cpu = ROOT / "oracle" / "_ref" / "rbc_in_shear_short_cpu"
    if "--cpu" not in sys.argv:
        s, t1, t2 = per_step(gpu, 500, 4500)
        print(f"binding (GPU): {s * 1e3:.4f} ms/step  (runs {t1:.2f} s, {t2:.2f} s)")
        with tempfile.TemporaryDirectory() as d:
            rep = str(pathlib.Path(d) / "wall.txt")
            run(gpu, 8000, {"BIB_STEP_WALL_REPORT": rep})
            print("binding (GPU), BIB_STEP_WALL_REPORT after 8000 steps [calls, mean ms/step since step 100, median ms/step]:", open(rep).read().strip())
        if "--long" in sys.argv:
            s, t1, t2 = per_step(gpu, 500, 12500)
            print(f"binding (GPU), 12 000 steps: {s * 1e3:.4f} ms/step  (runs {t1:.2f} s, {t2:.2f} s)")
            s, t1, t2 = per_step(gpu, 4500, 12500)
            print(f"binding (GPU), steps 4500 .. 12500: {s * 1e3:.4f} ms/step  (runs {t1:.2f} s, {t2:.2f} s)")
    s, t1, t2 = per_step(cpu, 5, 35)
    print(f"reference (1 core): {s * 1e3:.2f} ms/step  (runs {t1:.2f} s, {t2:.2f} s)")
