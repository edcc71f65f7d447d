#!/usr/bin/env python
"""bench.py -- headline benchmark of the biblbm hot path on B200 (contract: see the round brief).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[4], weak scaling along x): every GPU owns a 128 x 1024 x 1024 slab of a
(128*N) x 1024 x 1024 fp64 D3Q19 ForcedPhase lattice (N = 8 is the named 1024^3 lattice) carrying 1250 RBC meshes
(626 vertices each; N = 8: 10 000 RBCs). A step is one full driver step: velocity interpolation, membrane forces,
force spreading, vertex advection, voxelization and the fused collide-and-stream.
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
REF_HARNESS = ROOT / "oracle" / "_ref" / "ref_harness"
LBM_BYTES_PER_CELL = 304.0          # 19 populations x 8 B x (read + write): BASELINE.json / SURVEY.md section 8(d)
IB_BYTES_PER_VERTEX = 330.0


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.rows:
            if t < t0 - 0.1 or t > t1 + 0.3:
                continue
            parts = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(parts[1])); smax = float(parts[2])
            except Exception:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------- inputs
def rotation(axis, angle):
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def rbc_positions(rbc_pos, n, dims, x_offset, seed):
    """n randomly oriented RBCs on a jittered lattice of an nx x ny x nz slab starting at global x = x_offset."""
    rng = np.random.default_rng(seed)
    pitch = 18.0                                   # > RBC diameter (16.3) + jitter: no overlaps
    slots = [(i, j, k) for i in range(int(dims[0] // pitch)) for j in range(int(dims[1] // pitch)) for k in range(int(dims[2] // pitch))]
    if n > len(slots):
        raise ValueError(f"{n} RBCs do not fit a {dims} slab")
    pick = rng.choice(len(slots), size=n, replace=False)
    base = rbc_pos - rbc_pos.mean(0)
    out = np.empty((n,) + rbc_pos.shape)
    for q, s in enumerate(pick):
        c = (np.array(slots[s]) + 0.5) * pitch + rng.uniform(-0.6, 0.6, 3) + np.array([x_offset, 0, 0])
        out[q] = base @ rotation(rng.normal(size=3), rng.uniform(0, np.pi)).T + c
    return out


def load_rbc():
    z = np.load(ROOT / "tests" / "golden" / "rbc626.npz")
    return z["verts"], z["tris"], z["rbc_params"], z["rbc_pos"]


# ----------------------------------------------------------------------------------------------------- reference arm
def write_mesh(path, verts, tris):
    with open(path, "wb") as f:
        np.array([verts.shape[0], tris.shape[0]], dtype=np.int64).tofile(f)
        np.ascontiguousarray(verts, dtype=np.float64).tofile(f)
        np.ascontiguousarray(tris, dtype=np.int64).tofile(f)


def run_reference_sample(steps, warmup, replicas, box=96, n_rbc=8):
    """Times the UNMODIFIED reference (oracle/_ref/ref_harness) on `replicas` concurrent single-rank copies of a bounded
    sample of the workload: a box^3 periodic lattice with n_rbc RBCs at the bench's hematocrit. Returns aggregate numbers."""
    verts, tris, params, rbc_pos = load_rbc()
    procs, dirs = [], []
    for r in range(replicas):
        d = tempfile.mkdtemp(prefix="bibref_")
        dirs.append(d)
        write_mesh(os.path.join(d, "s.mesh"), verts, tris)
        pos = rbc_positions(rbc_pos, n_rbc, (box, box, box), 0, seed=100 + r)
        np.concatenate([params, pos.ravel()]).tofile(os.path.join(d, "p.f64"))
        cfg = dict(nx=box, ny=box, nz=box, steps=steps, warmup=warmup, dump_every=0, periodicity="ppp", omega=1.0, omega_inner=1.0 / 3.0,
                   init="uniform", rho=1.0, ux=0.02, uy=0.01, uz=-0.01, body_x=1e-6, body_y=0, body_z=0, timing=1, out=d,
                   mesh=os.path.join(d, "s.mesh"), particles=os.path.join(d, "p.f64"))
        with open(os.path.join(d, "c.txt"), "w") as f:
            f.write("".join(f"{k}={v}\n" for k, v in cfg.items()))
        procs.append(subprocess.Popen([str(REF_HARNESS), "run", os.path.join(d, "c.txt")], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True))
    t0 = time.time()
    outs = [p.communicate()[0] for p in procs]
    wall = time.time() - t0
    per = []
    for o in outs:
        line = [l for l in o.splitlines() if l.startswith("{")][-1]
        per.append(json.loads(line))
    t_step = max((p["t_compute_velocity"] + p["t_set_external"] + p["t_ib"] + p["t_collide_stream"]) / steps for p in per)
    t_lbm = max(p["t_collide_stream"] / steps for p in per)
    t_ib = max(p["t_ib"] / steps for p in per)
    cells, nvert = per[0]["cells"], per[0]["vertices"]
    for d in dirs:
        subprocess.run(["rm", "-rf", d])
    return dict(mlups=replicas * cells / t_step / 1e6, lbm_only_mlups=replicas * cells / t_lbm / 1e6,
                vertex_updates_per_s=replicas * nvert / t_ib if t_ib > 0 else None, ms_per_step=t_step * 1e3, wall_s=wall,
                sample=f"{replicas} concurrent single-rank replicas of a {box}^3 periodic box with {n_rbc} RBCs ({nvert} vertices) each, "
                       f"{steps} timed steps after {warmup}; no MPI exchange cost included")


def driver_speed():
    """The reference's OWN driver progs/rbc_in_shear (97^3 Couette cell, 1 RBC, unchanged source) compiled against the C++ binding in
    overlay/ vs against the pure reference: wall-clock per step from the difference of two run lengths (set-up and output cancel)."""
    sys.path.insert(0, str(ROOT / "tests"))
    import kit
    gpu, cpu = ROOT / "overlay" / "_build" / "rbc_in_shear_short_b200", ROOT / "oracle" / "_ref" / "rbc_in_shear_short_cpu"
    if not gpu.exists() or not cpu.exists():
        return None
    shape, _, _, _ = kit.load_rbc_fixture()

    def run(exe, nt):
        with tempfile.TemporaryDirectory() as d:
            kit.write_gmsh(os.path.join(d, "s.msh"), shape.verts, shape.tris)
            env = dict(os.environ, BIB_NT=str(nt), BIB_OUTPUT_INTERVAL="1000000", BIB_NCASES="1", BIB_MESH=os.path.join(d, "s.msh"), BIB_SHRINK_ITERS="2000")
            t0 = time.perf_counter()
            r = subprocess.run([str(exe)], cwd=d, env=env, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(r.stderr[-500:])
            return time.perf_counter() - t0
    try:
        run(gpu, 500)                                               # warm-up (driver + CUDA context start-up on this box)
        g = (run(gpu, 12500) - min(run(gpu, 500), run(gpu, 500))) / 12000.0       # start-up jitter of +-1 s -> +-0.08 ms per step
        c = (run(cpu, 35) - run(cpu, 5)) / 30.0
    except Exception as e:      # reported, never fatal for the bench line
        return {"error": str(e)[:200]}
    cells = 97 ** 3
    return {"workload": "progs/rbc_in_shear/rbc_in_shear.cpp unchanged (97^3 Couette cell, 1 RBC, regulator), linked against the binding vs the pure reference",
            "binding_ms_per_step": g * 1e3, "binding_mlups": cells / g / 1e6, "reference_ms_per_step": c * 1e3, "reference_mlups": cells / c / 1e6,
            "reference_cores": 1, "speedup": c / g}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not REF_HARNESS.exists():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_harness was not built (needs /root/reference at build time)"}))
        return
    cores = os.cpu_count() or 1
    replicas = max(1, min(cores, 32))
    r = run_reference_sample(args.steps, args.warmup, replicas)
    line = {"impl": "reference", "metric": "D3Q19+IB MLUPS (fp64, full driver step)", "value": r["mlups"], "unit": "MLUPS",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "lbm_only_mlups": r["lbm_only_mlups"], "ib_vertex_updates_per_s": r["vertex_updates_per_s"],
            "cpu_baseline": {"value": r["mlups"], "unit": "MLUPS", "cores": replicas, "kind": "reference", "sample": r["sample"]},
            "e2e": {"value": r["mlups"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(n):
    return {"workload": f"synthetic ({128 * n}x1024x1024) D3Q19 ForcedPhase lattice, {1250 * n} RBC meshes (626 vertices), x-slab per GPU 128x1024x1024",
            "global_cells": 128 * n * 1024 * 1024, "rbcs": 1250 * n, "vertices": 1250 * n * 626, "parallelism": f"x-slabs x{n}",
            "l2_policy": "inputs (41 GB of populations per GPU) exceed the 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------------------------------- our arm
def ours(args):
    import torch
    import torch.distributed as dist
    import biblbm_b200 as bib

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    nx, ny, nz = args.nx, args.ny, args.nz
    cells_local = nx * ny * nz
    lat = bib.Lattice(nx, ny, nz, 1.0, 1.0 / 3.0, periodic=(True, True, True), device=local)
    if world > 1:
        uid = [bib.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        lat.comm_init(rank, world, uid[0])
    lat.set_option(bib.OPT_FUSED_FACE_SWAP, float(args.face_swap))
    lat.set_option(bib.OPT_OVERLAP_FORCES, float(args.overlap_forces))
    lat.set_option(bib.OPT_PREFETCH_FORCES, float(args.prefetch_forces))
    lat.initializeAtEquilibrium(1.0, (0.02, 0.01, -0.01))
    body = (1e-6, 0.0, 0.0)
    n_rbc = args.rbcs                               # per GPU
    nvert = 0
    if n_rbc > 0:
        verts, tris, params, rbc_pos = load_rbc()
        sid = lat.register_shape(verts, tris)
        # every rank passes the same global particle list; the library keeps what the slab owns (+ ghosts of its neighbours)
        lat.add_particles(sid, params, rbc_positions(rbc_pos, n_rbc * world, (nx * world, ny, nz), 0, seed=7))
        nvert = n_rbc * verts.shape[0]
    lat.sync()

    def step(n=1):
        if n_rbc > 0:
            lat.step(body, n=n)
        else:
            lat.collideAndStream(n)

    # ---- headline: K full steps, state resident in HBM
    step(args.warmup)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    l0 = lat.launches()
    t_mark0 = time.time()
    lat.time_begin()
    w0 = time.perf_counter()
    step(args.steps)
    ms_dev = lat.time_end()
    barrier()
    w1 = time.perf_counter()
    t_mark1 = time.time()
    launches = lat.launches() - l0
    ms_step = max_over_ranks(ms_dev) / args.steps
    wall_ms_step = max_over_ranks((w1 - w0) * 1e3) / args.steps
    clocks = sampler.stop(t_mark0, t_mark1) if rank == 0 else None
    mlups = cells_local * world / (ms_step * 1e-3) / 1e6

    # ---- dominant kernel alone (collide-and-stream), CUDA events on the launching stream
    if args.face_swap != 1:                         # default (1): the periodic-face fix-up is part of k_collide_stream and is timed with it
        lat.set_option(bib.OPT_ENVELOPE_SWAP, 0.0)  # main kernel only (the separate k_face_swap pass is timed in the full step)
    lat.collideAndStream(2)
    barrier()
    lat.time_begin()
    lat.collideAndStream(args.steps)
    ms_kernel = max_over_ranks(lat.time_end()) / args.steps
    lat.set_option(bib.OPT_ENVELOPE_SWAP, 1.0)
    peak, peak_kind = measured_peaks()
    traffic = None
    tj = ROOT / "profiles" / "r1_k_collide_stream_traffic.json"
    if tj.exists():
        t = json.loads(tj.read_text())
        if t["workload"] == [nx, ny, nz]:
            traffic = t["dram_bytes_per_launch"]
    achieved = LBM_BYTES_PER_CELL * cells_local / (ms_kernel * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_collide_stream", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind}; burst figure, kernel timed alone)",
                "ms_per_launch": ms_kernel, "lbm_only_mlups": cells_local * world / (ms_kernel * 1e-3) / 1e6,
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "algorithmic_bytes_per_cell": LBM_BYTES_PER_CELL, "true_bytes_per_cell": 306.0,
                "true_bytes_note": "19 populations read + written, 1 phase byte, force-flag word; the 24 B/cell force field is read only in flagged (particle-band) cells"}

    # ---- IB phases alone (vertex-updates/s)
    ib = None
    if n_rbc > 0:
        barrier()
        lat.time_begin()
        for _ in range(args.steps):
            lat.ib_step()
        ms_ib = lat.time_end() / args.steps
        ms_ib = max_over_ranks(ms_ib)
        ib = {"vertex_updates_per_s": nvert * world / (ms_ib * 1e-3), "ms_per_step": ms_ib, "vertices": nvert * world,
              "achieved_gbs_at_330B_per_vertex": IB_BYTES_PER_VERTEX * nvert * world / (ms_ib * 1e-3) / 1e9}

    # ---- end to end through the C-ABI with host buffers: every step the host reads back what the reference drivers read
    #      (vertex state for write_lightweight_particle_data / VTK, particle scalars for the volume diagnostics)
    n_slots = lat.count()[0]
    d2h = lat.count()[1] * 72 + n_slots * 88
    # three sets of page-locked result buffers: two read-backs are in flight while the host consumes a third, so that the host always has
    # one whole step queued ahead of the copy it waits for (bib_ib_results_download_async / _wait: FIFO of depth 2)
    host_v = [torch.empty((max(1, lat.count()[1]), 9), dtype=torch.float64, pin_memory=True).numpy() for _ in range(3)]
    host_s = [torch.empty((max(1, n_slots), 11), dtype=torch.float64, pin_memory=True).numpy() for _ in range(3)]
    checksum = 0.0
    if n_rbc > 0:                                   # first use allocates the staging buffers and the copy stream: not part of a step
        for k in range(2):
            lat.results_download_async(host_v[k], host_s[k])
        lat.results_wait(); lat.results_wait()
    barrier()
    e0 = time.perf_counter()
    for it in range(args.steps):
        step(1)
        if n_rbc > 0:
            if it >= 2:
                lat.results_wait()                              # results of step it-2 are now on the host ...
                checksum += float(host_s[(it - 2) % 3][0, 1])    # ... and are consumed (a particle volume, like the drivers print)
            lat.results_download_async(host_v[it % 3], host_s[it % 3])
        else:
            lat.sync()
    if n_rbc > 0:
        for it in range(max(0, args.steps - 2), args.steps):
            lat.results_wait()
            checksum += float(host_s[it % 3][0, 1])
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - e0) * 1e3) / args.steps
    if n_rbc > 0:                                   # bytes actually delivered per step (multi-slab: the owned particles, compacted)
        n_del = lat.results_count()
        d2h = n_del * (nvert // n_rbc * 72 + 88)
    e2e = {"value": cells_local * world / (e2e_ms * 1e-3) / 1e6, "unit": "MLUPS", "h2d_bytes_per_step": 24, "d2h_bytes_per_step": d2h,
           "ms_per_step": e2e_ms,
           "note": "state stays resident in HBM (the point of the design); per step the host passes the body force (24 B, by value) and "
                   "reads all vertex states + particle scalars back into page-locked buffers (the data the reference drivers consume between steps); "
                   "the read-backs of steps t-1 and t overlap step t+1 on a copy stream (two in flight)", "checksum": checksum}

    if rank == 0:
        cpu = None
        if world == 1 and REF_HARNESS.exists() and not args.no_cpu_baseline:
            # the unmodified reference on ALL host cores of this box (one single-rank replica per core, bounded sample), and on one core
            n_rep = max(1, min(os.cpu_count() or 1, 32))
            r = run_reference_sample(10, 2, n_rep)
            r1 = run_reference_sample(5, 1, 1)
            cpu = {"value": r["mlups"], "unit": "MLUPS", "cores": n_rep, "kind": "reference", "sample": r["sample"],
                   "lbm_only_mlups": r["lbm_only_mlups"], "ib_vertex_updates_per_s": r["vertex_updates_per_s"],
                   "single_core": {"value": r1["mlups"], "lbm_only_mlups": r1["lbm_only_mlups"], "ib_vertex_updates_per_s": r1["vertex_updates_per_s"]}}
        drv = driver_speed() if (world == 1 and not args.no_cpu_baseline) else None
        cfg = workload_config(world) if (nx, ny, nz, args.rbcs) == (128, 1024, 1024, 1250) else {
            "workload": f"custom ({nx * world}x{ny}x{nz}) D3Q19 ForcedPhase lattice, {n_rbc * world} RBC meshes", "parallelism": f"x-slabs x{world}"}
        line = {"metric": "D3Q19+IB MLUPS (fp64, full driver step)", "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "wall_ms_per_step": wall_ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "roofline": roofline, "ib": ib, "cpu_baseline": cpu,
                "e2e": e2e, "driver_e2e": drv, "gpu_launches": launches, "clocks": clocks, "device_bytes": lat.device_bytes()}
        print(json.dumps(line))
    lat.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=128)
    ap.add_argument("--ny", type=int, default=1024)
    ap.add_argument("--nz", type=int, default=1024)
    ap.add_argument("--rbcs", type=int, default=1250)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--overlap-forces", type=int, default=1, choices=[0, 1], help="BIB_OPT_OVERLAP_FORCES (experiments)")
    ap.add_argument("--prefetch-forces", type=int, default=1, choices=[0, 1], help="BIB_OPT_PREFETCH_FORCES (experiments)")
    ap.add_argument("--face-swap", type=int, default=1, choices=[0, 1, 2], help="BIB_OPT_FUSED_FACE_SWAP (experiments)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
