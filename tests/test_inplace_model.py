"""CPU model of the in-place streaming schedule (biblbm_b200/csrc/lbm_kernels.cu: launch_inplace, k_ghost_wrap).

One population buffer of nx + 2 + S planes is seen through two frames S planes apart; a step pushes from one frame into the other in
launches of at most S - 1 planes, ascending in x for frame 0 -> 1 and descending for 1 -> 0. The launch list is the library's own
(bib_inplace_schedule, host-side); the model replays the address arithmetic with numpy on a small lattice (pure streaming, every population value unique) and checks, for every launch, that
 (1) no launch writes a storage location that it or a LATER launch of the step still has to read (the threads of a launch run in
     any order, launches of a stream run in order),
 (2) every destination slot is written exactly once,
 (3) after the ghost-plane wrap the lattice equals plain two-buffer push streaming with the reference's boundary rule (a population
     that would leave through a non-periodic face stays in the cell's opposite slot: src/atomicBlock/blockLattice3D.hh:402-427)."""
import itertools

import numpy as np
import pytest

import biblbm_b200 as bib

CX = [0, -1, 0, 0, -1, -1, -1, -1, 0, 0, 1, 0, 0, 1, 1, 1, 1, 0, 0]
CY = [0, 0, -1, 0, -1, 1, 0, 0, -1, -1, 0, 1, 0, 1, -1, 0, 0, 1, 1]
CZ = [0, 0, 0, -1, 0, 0, -1, 1, -1, 1, 0, 0, 1, 0, 0, 1, -1, 1, -1]
OPP = [0] + [i + 9 for i in range(1, 10)] + [i - 9 for i in range(10, 19)]


def stream_two_buffers(f, per):
    """f[19][nx][ny][nz] -> streamed copy (push; wrap on periodic axes, stay in the opposite slot at non-periodic faces)."""
    n = f.shape[1:]
    out = np.full_like(f, np.nan)
    for i in range(19):
        for x, y, z in itertools.product(range(n[0]), range(n[1]), range(n[2])):
            t = [x + CX[i], y + CY[i], z + CZ[i]]
            blocked = False
            for k in range(3):
                if t[k] < 0 or t[k] >= n[k]:
                    if per[k]:
                        t[k] %= n[k]
                    else:
                        blocked = True
            if blocked:
                out[OPP[i], x, y, z] = f[i, x, y, z]
            else:
                out[i, t[0], t[1], t[2]] = f[i, x, y, z]
    return out


class InPlace:
    def __init__(self, f, per, shift, x_lo_wall=None, x_hi_wall=None, self_wrap=True):
        """x_lo_wall / x_hi_wall: is the low / high x face of THIS slab a wall (default: the lattice is one slab, walls iff x is not periodic);
        self_wrap: the slab is its own x-neighbour (k_ghost_wrap); False: halo_pull() fills the boundary planes from the neighbours."""
        self.nx, self.ny, self.nz = f.shape[1:]
        self.per, self.S = per, shift
        self.x_lo_wall = (not per[0]) if x_lo_wall is None else x_lo_wall
        self.x_hi_wall = (not per[0]) if x_hi_wall is None else x_hi_wall
        self.self_wrap = self_wrap
        self.mem = np.full((19, self.nx + 2 + shift, self.ny, self.nz), np.nan)
        self.cur = 0
        self.mem[:, self.base(0) + 1:self.base(0) + 1 + self.nx] = f

    def base(self, frame):                  # f[0] = f_base + shift planes, f[1] = f_base
        return self.S if frame == 0 else 0

    def state(self):
        b = self.base(self.cur)
        return self.mem[:, b + 1:b + 1 + self.nx].copy()

    def step(self):
        nx, ny, nz, S = self.nx, self.ny, self.nz, self.S
        src, dst = self.base(self.cur), self.base(1 - self.cur)
        # the launch list comes from the LIBRARY (bib_inplace_schedule = what launch_inplace iterates over; host-side, no device needed)
        chunks = [(a, a + n) for a, n in bib.inplace_schedule(nx, S, self.cur)]
        unread = np.zeros(self.mem.shape[1], dtype=bool)
        unread[src + 1:src + 1 + nx] = True
        written = np.zeros(self.mem.shape, dtype=np.int32)
        for a, b in chunks:
            writes = []
            for x in range(a, b):
                for i, y, z in itertools.product(range(19), range(ny), range(nz)):
                    v = self.mem[i, src + 1 + x, y, z]
                    tx, ty, tz = x + CX[i], y + CY[i], z + CZ[i]          # x: always into the ghost planes (x mode 0) unless a wall
                    blocked = False
                    if (tx < 0 and self.x_lo_wall) or (tx >= nx and self.x_hi_wall):
                        blocked = True
                    if ty < 0 or ty >= ny:
                        if self.per[1]: ty %= ny
                        else: blocked = True
                    if tz < 0 or tz >= nz:
                        if self.per[2]: tz %= nz
                        else: blocked = True
                    writes.append((OPP[i], dst + 1 + x, y, z, v) if blocked else (i, dst + 1 + tx, ty, tz, v))
            for i, p, y, z, v in writes:
                assert not unread[p], f"launch [{a},{b}) overwrites storage plane {p}, which is still to be read"
            for i, p, y, z, v in writes:
                self.mem[i, p, y, z] = v
                written[i, p, y, z] += 1
            unread[src + 1 + a:src + 1 + b] = False
        assert not unread.any()
        self.cur = 1 - self.cur
        self.written = written
        if self.per[0] and self.self_wrap:  # k_ghost_wrap
            b0 = self.base(self.cur)
            for i in range(19):
                if CX[i] == 0:
                    continue
                for y, z in itertools.product(range(ny), range(nz)):
                    sy, sz = y - CY[i], z - CZ[i]
                    if (not self.per[1] and (sy < 0 or sy >= ny)) or (not self.per[2] and (sz < 0 or sz >= nz)):
                        continue
                    if CX[i] > 0:
                        self.mem[i, b0 + 1, y, z] = self.mem[i, b0 + nx + 1, y, z]; written[i, b0 + 1, y, z] += 1
                    else:
                        self.mem[i, b0 + nx, y, z] = self.mem[i, b0, y, z]; written[i, b0 + nx, y, z] += 1
        if self.self_wrap:
            self.check_written_once()

    def check_written_once(self):
        b0 = self.base(self.cur)
        assert (self.written[:, b0 + 1:b0 + 1 + self.nx] == 1).all(), "every slot of the new state is written exactly once"

    def halo_pull(self, left, right):
        """k_halo_pull (comm.cu): c_x = +1 populations from the left neighbour's ghost plane nx_l + 1 into my plane 1, c_x = -1 populations
        from the right neighbour's ghost plane 0 into my plane nx; holes behind y / z walls are skipped."""
        ny, nz = self.ny, self.nz
        b0 = self.base(self.cur)
        for i in range(19):
            if CX[i] == 0:
                continue
            for y, z in itertools.product(range(ny), range(nz)):
                sy, sz = y - CY[i], z - CZ[i]
                if (not self.per[1] and (sy < 0 or sy >= ny)) or (not self.per[2] and (sz < 0 or sz >= nz)):
                    continue
                if CX[i] > 0 and left is not None:
                    self.mem[i, b0 + 1, y, z] = left.mem[i, left.base(left.cur) + left.nx + 1, y, z]; self.written[i, b0 + 1, y, z] += 1
                if CX[i] < 0 and right is not None:
                    self.mem[i, b0 + self.nx, y, z] = right.mem[i, right.base(right.cur), y, z]; self.written[i, b0 + self.nx, y, z] += 1


@pytest.mark.parametrize("shift", [2, 3, 4, 9])
@pytest.mark.parametrize("per", [(1, 1, 1), (0, 1, 0), (1, 0, 1), (0, 0, 0)])
def test_inplace_schedule_equals_two_buffer_streaming(per, shift):
    nx, ny, nz = 7, 3, 4
    f = np.arange(19 * nx * ny * nz, dtype=np.float64).reshape(19, nx, ny, nz) + 1.0
    model = InPlace(f, per, shift)
    ref = f
    for step in range(5):
        model.step()
        ref = stream_two_buffers(ref, per)
        assert np.array_equal(model.state(), ref), f"step {step + 1}"


def test_launches_wider_than_the_shift_would_clobber():
    """The bound the schedule relies on: with launches of S planes (instead of S - 1) a launch writes into planes it reads itself."""
    nx, ny, nz, S = 7, 2, 2, 3
    src, dst = S, 0
    a, b = 0, S                              # a launch of S planes, frame 0 -> 1
    read = set(range(src + 1 + a, src + 1 + b))
    wrote = set(range(dst + 1 + a - 1, dst + 1 + b + 1))
    assert read & wrote
    b = S - 1
    assert not (set(range(src + 1 + a, src + 1 + b)) & set(range(dst + 1 + a - 1, dst + 1 + b + 1)))


@pytest.mark.parametrize("per", [(1, 1, 1), (0, 1, 0), (1, 0, 1)])
@pytest.mark.parametrize("widths,shifts", [((4, 3), (3, 3)), ((3, 4, 3), (2, 4, 0 + 5))])
def test_inplace_slabs_with_pulled_halo_equal_the_undivided_lattice(per, widths, shifts):
    """Several x-slabs of unequal width, every slab in place with its OWN shift (the layouts of the ranks may differ), halo as a pull from the
    neighbours' ghost planes after all launches of the step: equal to two-buffer streaming of the undivided lattice."""
    gx, ny, nz = sum(widths), 3, 4
    f = np.arange(19 * gx * ny * nz, dtype=np.float64).reshape(19, gx, ny, nz) + 1.0
    x0 = np.concatenate([[0], np.cumsum(widths)])
    n = len(widths)
    slabs = [InPlace(f[:, x0[r]:x0[r + 1]], per, shifts[r], x_lo_wall=(r == 0 and not per[0]), x_hi_wall=(r == n - 1 and not per[0]), self_wrap=False)
             for r in range(n)]
    ref = f
    for step in range(4):
        for s in slabs:
            s.step()
        for r, s in enumerate(slabs):              # every slab has finished its launches ("ghost planes written") before anybody pulls
            left = slabs[(r - 1) % n] if (per[0] or r > 0) else None
            right = slabs[(r + 1) % n] if (per[0] or r < n - 1) else None
            s.halo_pull(left, right)
        for s in slabs:
            s.check_written_once()
        ref = stream_two_buffers(ref, per)
        got = np.concatenate([s.state() for s in slabs], axis=1)
        assert np.array_equal(got, ref), f"step {step + 1}"


def test_library_schedule_invariants_for_many_sizes():
    """For every (nx, shift, frame) of a sweep - including the bench slab (128, 17), the drivers' 97-plane lattices and shifts wider than the
    slab - the library's launch list covers every plane exactly once and no launch writes a storage plane that it or a LATER launch reads
    (storage plane of (frame 0, x) = x + 1 + S, of (frame 1, x) = x + 1; a launch over [a, b) writes the planes a - 1 .. b of the other frame)."""
    checked = 0
    for nx in list(range(1, 40)) + [61, 97, 121, 128, 241, 512, 768]:
        for S in [2, 3, 4, 5, 8, 9, 16, 17, 33, 64, 65, 129, 1000]:
            for frame in (0, 1):
                launches = bib.inplace_schedule(nx, S, frame)
                src, dst = (S, 0) if frame == 0 else (0, S)
                covered = sorted(x for a, n in launches for x in range(a, a + n))
                assert covered == list(range(nx)), (nx, S, frame)
                assert all(1 <= n <= S - 1 for _, n in launches)
                reads = [set(range(src + 1 + a, src + 1 + a + n)) for a, n in launches]
                for i, (a, n) in enumerate(launches):
                    writes = set(range(dst + 1 + a - 1, dst + 1 + a + n + 1))
                    still_needed = set().union(*reads[i:])
                    assert not (writes & still_needed), (nx, S, frame, i)
                checked += 1
    assert checked > 1000


def test_schedule_argument_checking():
    with pytest.raises(bib.BibError, match="inplace_schedule"):
        bib.inplace_schedule(10, 1, 0)
    with pytest.raises(bib.BibError, match="inplace_schedule"):
        bib.inplace_schedule(0, 5, 0)
    with pytest.raises(bib.BibError, match="inplace_schedule"):
        bib.inplace_schedule(10, 5, 2)
