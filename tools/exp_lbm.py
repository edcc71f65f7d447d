"""Experiment driver: sustained timings of the lattice kernel / the full step for one library build and option set."""
import argparse, json, os, sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import biblbm_b200 as bib
sys.path.insert(0, str(ROOT))
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--tag", default="")
ap.add_argument("--face-swap", type=int, default=1)
ap.add_argument("--envelope", type=int, default=1)
ap.add_argument("--rbcs", type=int, default=1250)
ap.add_argument("--seconds", type=float, default=1.5)
ap.add_argument("--nx", type=int, default=128)
ap.add_argument("--prefetch", type=int, default=-1)
ap.add_argument("--inplace", type=int, default=-1)
ap.add_argument("--interp-compact", type=int, default=0)
ap.add_argument("--ny", type=int, default=1024)
ap.add_argument("--nz", type=int, default=1024)
a = ap.parse_args()
bib.set_default_inplace(a.inplace)
lat = bib.Lattice(a.nx, a.ny, a.nz, 1.0, 1 / 3)
lat.set_option(bib.OPT_FUSED_FACE_SWAP, float(a.face_swap))
lat.set_option(bib.OPT_ENVELOPE_SWAP, float(a.envelope))
lat.set_option(bib.OPT_INTERP_COMPACT, float(a.interp_compact))
if a.prefetch >= 0:
    lat.set_option(bib.OPT_LBM_PREFETCH, float(a.prefetch))
lat.initializeAtEquilibrium(1.0, (0.02, 0.01, -0.01))
body = (1e-6, 0.0, 0.0)
if a.rbcs:
    verts, tris, params, rbc_pos = bench.load_rbc()
    sid = lat.register_shape(verts, tris)
    lat.add_particles(sid, params, bench.rbc_positions(rbc_pos, a.rbcs, (a.nx, a.ny, a.nz), 0, seed=7))
lat.setExternalVector(body)      # production state: every cell holds the body force, force flags valid (sparse force reads)
lat.sync()
def timed(fn, K=20):
    fn(K); lat.sync()
    out = []
    t0 = time.time()
    while time.time() - t0 < a.seconds:
        lat.time_begin(); fn(K); out.append(lat.time_end() / K)
    return out
k = timed(lambda n: lat.collideAndStream(n))
res = {"tag": a.tag, "interp_compact": a.interp_compact, "inplace_shift": lat.inplace_shift(), "dims": [a.nx, a.ny, a.nz], "device_bytes": lat.device_bytes(), "lib": os.environ.get("BIB_LIB", "default"), "prefetch": a.prefetch, "face_swap": a.face_swap, "envelope": a.envelope,
       "lbm_ms_first": k[0], "lbm_ms_median": float(np.median(k)), "lbm_ms_min": float(np.min(k))}
if a.rbcs:
    s = timed(lambda n: lat.step(body, n=n))
    res.update({"step_ms_first": s[0], "step_ms_median": float(np.median(s)), "step_ms_min": float(np.min(s))})
    lat.phase_timing(True); lat.step(body, n=20); lat.phase_times(); lat.step(body, n=20); pt = lat.phase_times(); lat.phase_timing(False)
    res["phases"] = {k_: round(v / 20, 4) for k_, v in pt.items() if v > 0.01}
print(json.dumps(res))
