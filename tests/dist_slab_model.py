"""CPU model of the multi-slab algorithm (run by tests/test_dist_cpu.py under torch.distributed with the gloo backend,
world_size 2): every rank owns an x-slab with one ghost plane per side, collides its cells with the oracle, PUSHES the
post-collision populations (numpy) exactly like k_collide_stream does -- periodic wrap in y/z, ghost planes in x, the
envelope quirk only at the GLOBAL periodic faces -- and exchanges just the 5 populations that cross each face, received
in place into the boundary planes (the NCCL halo exchange of biblbm_b200/csrc/comm.cu). The gathered result must be
bit-identical to the oracle stepping the undivided lattice."""
import os
import sys
import pathlib

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent))
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import kit  # noqa: E402

PLUS_X = [10, 13, 14, 15, 16]
MINUS_X = [1, 4, 5, 6, 7]


def push_planes(orc, chunk, a, nx, dest, dest_plane0, omega0, omega1, rank, world, per):
    """chunk: the cells of the slab's planes [a, a + n) as [n][ny][nz][23]; their post-collision populations are pushed into
    dest[19][planes][ny][nz], where plane x of the slab is stored at dest_plane0 + x (the ghost planes at dest_plane0 - 1 and
    dest_plane0 + nx): periodic wrap in y / z, ghost planes in x, the envelope quirk only at the GLOBAL periodic faces."""
    import ctypes as C
    n, ny, nz, _ = chunk.shape
    post = chunk.copy()
    orc.lib.orc_collide_cells(post.ctypes.data_as(C.POINTER(C.c_double)), C.c_long(n * ny * nz), C.c_double(omega0), C.c_double(omega1), C.c_int(0))
    env = chunk.copy()
    orc.lib.orc_collide_cells(env.ctypes.data_as(C.POINTER(C.c_double)), C.c_long(n * ny * nz), C.c_double(omega0), C.c_double(omega1), C.c_int(1))
    first, last = rank == 0, rank == world - 1
    for i in range(19):
        cx, cy, cz = kit.C19[i]
        src_n = post[..., i]
        src_e = env[..., i]
        val = src_n.copy()
        # populations crossing a GLOBAL periodic face come from the swapped (envelope) collision
        if per[1]:
            if cy == 1: val[:, -1, :] = src_e[:, -1, :]
            if cy == -1: val[:, 0, :] = src_e[:, 0, :]
        if per[2]:
            if cz == 1: val[:, :, -1] = src_e[:, :, -1]
            if cz == -1: val[:, :, 0] = src_e[:, :, 0]
        if per[0]:
            if cx == 1 and last and a + n == nx: val[-1] = src_e[-1]
            if cx == -1 and first and a == 0: val[0] = src_e[0]
        assert per[1] and per[2], "model covers periodic y/z"
        shifted = np.roll(val, (cy, cz), axis=(1, 2))
        dest[i, dest_plane0 + a + cx:dest_plane0 + a + cx + n] = shifted        # x: into the ghost planes, no wrap


def push_step(orc, slab, omega0, omega1, rank, world, per):
    """Two buffers: slab [nx][ny][nz][23] -> fresh destination buffer f[19][nx + 2][ny][nz] (ghost planes 0 and nx + 1) before the halo exchange."""
    nx, ny, nz, _ = slab.shape
    f = np.zeros((19, nx + 2, ny, nz))
    push_planes(orc, slab, 0, nx, f, 1, omega0, omega1, rank, world, per)
    return f


class InPlaceSlab:
    """In-place streaming (biblbm_b200/csrc/lbm_kernels.cu: launch_inplace): ONE buffer of nx + 2 + S planes seen through two
    frames S planes apart (frame 1 at plane 0, frame 0 at plane S); a step pushes from frame `cur` into the other one in launches of S - 1 planes,
    ascending in x for 0 -> 1 and descending for 1 -> 0, so that no launch overwrites a plane that it or a later launch still reads."""

    def __init__(self, slab, shift):
        self.nx, self.ny, self.nz, _ = slab.shape
        self.S, self.cur = shift, 0
        self.mem = np.full((19, self.nx + 2 + shift, self.ny, self.nz), np.nan)
        self.ext = slab[..., 19:].copy()                      # force + phase (not streamed)
        self.mem[:, shift + 1:shift + 1 + self.nx] = np.moveaxis(slab[..., :19], -1, 0)

    def base(self, frame):
        return self.S if frame == 0 else 0

    def frame(self):                                          # the current frame incl. its two ghost planes: [19][nx + 2][ny][nz] (a view)
        b = self.base(self.cur)
        return self.mem[:, b:b + self.nx + 2]

    def step(self, orc, omega0, omega1, rank, world, per):
        import biblbm_b200 as bib
        nx = self.nx
        src, dst = self.base(self.cur), self.base(1 - self.cur)
        for a, n in bib.inplace_schedule(nx, self.S, self.cur):       # the library's own launch list (host-side)
            chunk = np.concatenate([np.moveaxis(self.mem[:, src + 1 + a:src + 1 + a + n], 0, -1), self.ext[a:a + n]], axis=-1)
            assert not np.isnan(chunk).any(), "a launch read a plane that an earlier launch of the step had overwritten with nothing valid"
            push_planes(orc, np.ascontiguousarray(chunk), a, nx, self.mem, dst + 1, omega0, omega1, rank, world, per)
        self.cur = 1 - self.cur
        return self.frame()


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    orc = kit.oracle()
    nxl, ny, nz = 5, 6, 7
    per = (1, 1, 1)
    rng = np.random.default_rng(5)
    gx = nxl * world
    cells = np.zeros((gx, ny, nz, 23))
    cells[..., :19] = 0.01 * rng.normal(size=(gx, ny, nz, 19))
    cells[..., 19:22] = 1e-4 * rng.normal(size=(gx, ny, nz, 3))
    cells[..., 22] = rng.integers(0, 2, size=(gx, ny, nz))
    slab = cells[rank * nxl:(rank + 1) * nxl].copy()
    right, left = (rank + 1) % world, (rank - 1) % world
    steps = 4
    shift = int(os.environ.get("SLAB_MODEL_INPLACE", "0"))    # 0: two buffers; S >= 2: every rank streams in place with shift S + rank (layouts may differ)
    inplace = InPlaceSlab(slab, shift + rank) if shift else None
    for _ in range(steps):
        f = inplace.step(orc, 1.5, 0.6, rank, world, per) if inplace else push_step(orc, slab, 1.5, 0.6, rank, world, per)
        # halo: ghost plane nx+1 (c_x=+1 pops) -> right neighbour's plane 1; ghost plane 0 (c_x=-1) -> left neighbour's plane nx
        send_r = torch.from_numpy(np.ascontiguousarray(f[PLUS_X, nxl + 1]))
        send_l = torch.from_numpy(np.ascontiguousarray(f[MINUS_X, 0]))
        recv_l = torch.empty_like(send_r); recv_r = torch.empty_like(send_l)
        reqs = [dist.isend(send_r, right), dist.isend(send_l, left), dist.irecv(recv_l, left), dist.irecv(recv_r, right)]
        for r in reqs:
            r.wait()
        f[PLUS_X, 1] = recv_l.numpy()
        f[MINUS_X, nxl] = recv_r.numpy()
        slab[..., :19] = np.moveaxis(f[:, 1:nxl + 1], 0, -1)
    parts = [torch.empty(slab.shape, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(slab))
    ok = True
    if rank == 0:
        got = torch.cat(parts, 0).numpy()
        ref = cells
        for _ in range(steps):
            ref = orc.collide_and_stream(ref, per, 1.5, 0.6)
        ok = bool(np.array_equal(got, ref))
        print(f"gloo slab model world={world} inplace_shift={shift}: bit-identical={ok}")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
