#!/usr/bin/env python
"""Diagnostic (multi-GPU): where does the end-to-end loop of bench.py lose time relative to the device-timed step?
   torchrun ... tools/e2e_probe.py   -- prints per-variant ms/step on rank 0."""
import os, sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch, torch.distributed as dist
import biblbm_b200 as bib
import bench

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
nx, ny, nz = 128, 1024, 1024
lat = bib.Lattice(nx, ny, nz, 1.0, 1.0 / 3.0, periodic=(True, True, True), device=local)
if world > 1:
    uid = [bib.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    lat.comm_init(rank, world, uid[0])
lat.initializeAtEquilibrium(1.0, (0.02, 0.01, -0.01))
verts, tris, params, rbc_pos = bench.load_rbc()
sid = lat.register_shape(verts, tris)
lat.add_particles(sid, params, bench.rbc_positions(rbc_pos, 1250 * world, (nx * world, ny, nz), 0, seed=7))
body = (1e-6, 0.0, 0.0)
lat.step(body, n=5); lat.sync()
n_slots, n_vert = lat.count()
hv = [torch.empty((max(1, n_vert), 9), dtype=torch.float64, pin_memory=True).numpy() for _ in range(3)]
hs = [torch.empty((max(1, n_slots), 11), dtype=torch.float64, pin_memory=True).numpy() for _ in range(3)]
K = 10
lat.results_download_async(hv[0], hs[0]); lat.results_download_async(hv[1], hs[1]); lat.results_wait(); lat.results_wait()     # first use allocates: not timed
def run(name, fn):
    barrier(); t0 = time.perf_counter(); fn(); barrier(); dt = (time.perf_counter() - t0) * 1e3 / K
    if rank == 0: print(f"{name}: {dt:.3f} ms/step", flush=True)
run("A  step(K) one call", lambda: lat.step(body, n=K))
def b():
    for _ in range(K): lat.step(body, n=1)
run("B  K x step(1)", b)
def c():
    for it in range(K):
        lat.step(body, n=1); lat.sync()
run("C  K x (step(1); sync)", c)
def d():
    for it in range(K):
        lat.step(body, n=1)
        if it > 0: lat.results_wait()
        lat.results_download_async(hv[it % 2], hs[it % 2])
    lat.results_wait()
run("D  one read-back in flight", d)
def d2():
    for it in range(K):
        lat.step(body, n=1)
        if it > 1: lat.results_wait()
        lat.results_download_async(hv[it % 3], hs[it % 3])
    lat.results_wait(); lat.results_wait()
run("D2 two read-backs in flight (bench e2e loop)", d2)
def e():
    for it in range(K):
        lat.step(body, n=1)
        lat.results_download_async(hv[it % 2], hs[it % 2])
        lat.results_wait()
run("E  step; download; wait (no overlap)", e)
# where does the host wait? enqueue time of a step, and how long after it the copy of the PREVIOUS step has landed
barrier()
rows = []
for it in range(K):
    t0 = time.perf_counter(); lat.step(body, n=1); t1 = time.perf_counter()
    if it > 0: lat.results_wait()
    t2 = time.perf_counter(); lat.results_download_async(hv[it % 2], hs[it % 2]); t3 = time.perf_counter()
    rows.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
lat.results_wait(); barrier()
if rank == 0:
    for r in rows: print("enqueue step %.2f ms | wait for previous copy %.2f ms | enqueue download %.2f ms" % r)
if rank == 0: print("slots", n_slots, "vertices", n_vert, "d2h MB/step", (n_vert * 72 + n_slots * 88) / 1e6)
lat.close()
if world > 1: dist.destroy_process_group()
