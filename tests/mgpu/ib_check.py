"""Multi-GPU immersed-boundary check (torch.distributed.run, one rank per GPU): RBCs that straddle a slab face, sit on the
periodic seam and drift across faces; the sharded run must agree with the same library on ONE GPU holding the undivided
lattice (itself bit-checked against the reference oracle) within the north-star tolerance."""
import os
import sys
import pathlib

import numpy as np
import torch
import torch.distributed as dist

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import biblbm_b200 as bib  # noqa: E402
import kit  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    shape, params, rbc_pos, _ = kit.load_rbc_fixture()
    nxl, ny, nz = 24, 24, 22
    gx = nxl * world
    steps = int(os.environ.get("IB_CHECK_STEPS", "240"))
    centers = [(nxl - 0.4, 11.7, 10.6), (gx - 1.3, 12.4, 11.2), (10.2, 11.1, 10.9)]
    if world > 2:
        centers.append((2 * nxl + 3.1, 12.2, 11.0))
    # IB_CHECK_CASE=channel: the rbcs_in_channel set-up (BASELINE configs[3]) in miniature -- no-slip velocity planes at y = 0 / ny-1,
    # ParallelPlatesBoundary for the particles, body-force driven flow, a dense row of RBCs along x crossing every slab face
    channel = os.environ.get("IB_CHECK_CASE", "periodic") == "channel"
    per = (True, False, True) if channel else (True, True, True)
    if channel:
        centers = [(4.0 + 8.1 * k, 6.0 + 11.5 * (k % 2) + 0.3 * (k % 3), 10.4 + 0.5 * (k % 4)) for k in range(int(gx / 8.1))]
    pos = kit.place_particles(rbc_pos, centers, seed=9)
    body = (4e-5, 0.0, 1e-5)
    u0 = (0.05, 0.0, 0.004) if not channel else (0.0, 0.0, 0.0)

    def configure(l, n_x):
        if channel:
            zero = np.zeros((n_x, nz, 3))
            l.addVelocityBoundary((0, n_x - 1, 0, 0, 0, nz - 1), 1, -1, zero)
            l.addVelocityBoundary((0, n_x - 1, ny - 1, ny - 1, 0, nz - 1), 1, +1, zero)
            l.set_boundary(0.0, float(ny - 1))

    lat = bib.Lattice(nxl, ny, nz, 1.0, 1 / 3, periodic=per, device=local)
    uid = [bib.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    lat.comm_init(rank, world, uid[0])
    lat.initializeAtEquilibrium(1.0, u0)
    configure(lat, nxl)
    sid = lat.register_shape(shape.verts, shape.tris)
    lat.add_particles(sid, params, pos)          # same global list on every rank
    lat.step(body, n=steps)
    lat.sync()
    gids, roles = lat.local_slots()
    verts = lat.vertices().reshape(len(gids), shape.nv, 9)
    mine = {int(g): verts[s] for s, (g, r) in enumerate(zip(gids, roles)) if r == 1}
    # asynchronous read-back on a multi-slab domain: the owned particles, compacted in slot order, bit for bit what vertices() holds
    hv = torch.empty((len(gids) * shape.nv, 9), dtype=torch.float64, pin_memory=True).numpy()
    hs = torch.empty((len(gids), 11), dtype=torch.float64, pin_memory=True).numpy()
    lat.results_download_async(hv, hs)
    lat.results_wait()
    owned_slots = [s for s, r in enumerate(roles) if r == 1]
    n_del = lat.results_count()
    assert n_del == len(owned_slots), (n_del, len(owned_slots))
    scal = lat.particle_scalars()
    for j, s in enumerate(owned_slots):
        assert np.array_equal(hv.reshape(-1, shape.nv, 9)[j], verts[s]), (rank, j, s)
        assert np.array_equal(hs[j], scal[s]), (rank, j, s)
    # the asynchronous copy covers a prefix sized from the last count; a particle count beyond it is completed by results_wait
    os.environ["BIB_RESULTS_BOUND"] = "0" if len(owned_slots) < 2 else "1"
    hv[:] = 0.0; hs[:] = 0.0
    lat.results_download_async(hv, hs)
    lat.results_wait()
    del os.environ["BIB_RESULTS_BOUND"]
    assert lat.results_count() == len(owned_slots)
    for j, s in enumerate(owned_slots):
        assert np.array_equal(hv.reshape(-1, shape.nv, 9)[j], verts[s]), (rank, j, s)
        assert np.array_equal(hs[j], scal[s]), (rank, j, s)
    cells = torch.from_numpy(lat.download()).cuda()
    parts = [torch.empty_like(cells) for _ in range(world)]
    dist.all_gather(parts, cells)
    allv = [None] * world
    dist.all_gather_object(allv, mine)
    ok = True
    if rank == 0:
        owned = {}
        for r, m in enumerate(allv):
            for g, v in m.items():
                assert g not in owned, f"particle {g} owned twice"
                owned[g] = (r, v)
        assert sorted(owned) == list(range(len(centers))), f"owners: {sorted(owned)}"
        ref = bib.Lattice(gx, ny, nz, 1.0, 1 / 3, periodic=per, device=local)
        ref.initializeAtEquilibrium(1.0, u0)
        configure(ref, gx)
        rs = ref.register_shape(shape.verts, shape.tris)
        ref.add_particles(rs, params, pos)
        ref.step(body, n=steps)
        rv = ref.vertices().reshape(len(centers), shape.nv, 9)
        rc = ref.download()
        got = torch.cat(parts, 0).cpu().numpy()
        e_pos = max(kit.rel_err(owned[g][1][:, :3], rv[g][:, :3]) for g in owned)
        e_vel = max(kit.rel_err(owned[g][1][:, 3:6], rv[g][:, 3:6]) for g in owned)
        e_pop = kit.rel_err(got[..., :19], rc[..., :19])
        ph = np.array_equal(got[..., 22], rc[..., 22])
        owners = {g: owned[g][0] for g in sorted(owned)}
        com = {g: float(rv[g][:, 0].mean()) for g in sorted(owned)}
        print(f"case={'channel' if channel else 'periodic'} particles={len(centers)} world={world} steps={steps}: pos {e_pos:.2e} vel {e_vel:.2e} pop {e_pop:.2e} phase_equal={ph} owners={owners} com_x={com}")
        ok = e_pos <= 1e-9 and e_pop <= 1e-9 and ph
        ref.close()
    lat.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
