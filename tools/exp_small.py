"""Few-particle regime: per-phase device times of a 97^3 periodic lattice with 1 RBC (the size of the reference's rbc_in_shear driver)."""
import json, sys, pathlib, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import biblbm_b200 as bib
import kit
shape, params, rbc_pos, _ = kit.load_rbc_fixture()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 97
n_rbc = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lat = bib.Lattice(n, n, n, 1.0, 1 / 3)
lat.initializeAtEquilibrium(1.0, (0.02, 0.0, 0.004))
centers = [(n / 2 + 20.0 * (k % 3 - 1), n / 2 + 20.0 * ((k // 3) % 3 - 1), n / 2 + 20.0 * (k // 9 - 1)) for k in range(n_rbc)] if n_rbc > 1 else [(n / 2, n / 2, n / 2)]
sid = lat.register_shape(shape.verts, shape.tris)
lat.add_particles(sid, params, kit.place_particles(rbc_pos, centers, seed=1))
body = (1e-6, 0, 0)
lat.step(body, n=20); lat.sync()
lat.time_begin(); lat.step(body, n=2000); ms = lat.time_end() / 2000
t0 = time.perf_counter()
for _ in range(2000): lat.step(body, n=1)
lat.sync(); wall = (time.perf_counter() - t0) / 2000 * 1e3
lat.phase_timing(True); lat.step(body, n=50); lat.phase_times(); lat.step(body, n=200); pt = lat.phase_times(); lat.phase_timing(False)
print(json.dumps({"n": n, "rbcs": n_rbc, "ms_per_step_graph": ms, "wall_ms_per_step_one_call_per_step": wall, "phases_us": {k: round(v / 200 * 1e3, 1) for k, v in pt.items() if v > 0}}))
