"""Multi-GPU slab check, launched by torch.distributed.run with one rank per GPU (see tests/test_gpu_multi.py):
every rank owns an x-slab of a seeded random lattice, steps it with the NCCL halo exchange and the result, gathered on
rank 0, must be BIT-IDENTICAL to the CPU oracle stepping the undivided lattice (envelope quirk at the global faces only)."""
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
    ok = True
    for per in [(1, 1, 1), (0, 1, 0), (1, 0, 1)]:
        nxl, ny, nz = 6, 9, 20
        rng = np.random.default_rng(21)
        gx = nxl * world
        cells = np.zeros((gx, ny, nz, 23))
        cells[..., :19] = 0.01 * rng.normal(size=(gx, ny, nz, 19))
        cells[..., 19:22] = 1e-4 * rng.normal(size=(gx, ny, nz, 3))
        cells[..., 22] = rng.integers(0, 2, size=(gx, ny, nz))
        lat = bib.Lattice(nxl, ny, nz, 1.5, 0.6, periodic=per, device=local)
        uid = [bib.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        lat.comm_init(rank, world, uid[0])
        lat.upload(cells[rank * nxl:(rank + 1) * nxl])
        steps = 5
        lat.collideAndStream(steps)
        mine = torch.from_numpy(lat.download()).cuda()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        if rank == 0:
            got = torch.cat(parts, 0).cpu().numpy()
            ref = cells
            orc = kit.oracle()
            for _ in range(steps):
                ref = orc.collide_and_stream(ref, per, 1.5, 0.6)
            same = np.array_equal(got, ref)
            print(f"per={per} world={world}: bit-identical={same} maxdiff={np.abs(got - ref).max():.3e}")
            ok = ok and same
        lat.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
