"""Row-strip sharding of one world over several GPUs.

Host-side logic only: which global rows each rank owns, which rows cross a strip boundary after which
sub-step, and the per-tick schedule.  The schedule is executed either by libfss_b200.so itself
(fss_tick with a communicator attached: NCCL send/recv to rank +-1), by peer copies inside one process
(world.StripWorld), or -- in the CPU tests -- by any object with the World surface plus a row transport.

Boundary B between an upper strip (rows < B) and a lower strip (rows >= B), B = zone.y + k * 32:
  after tk 1 (phases 0,1 ran the chunk row just ABOVE B):  upper -> lower, rows [B-2, B+2)
  after tk 3 (phases 2,3 ran the chunk row just BELOW B):  lower -> upper, rows [B-2, B+10)
Reach, from the reference's rules: writes +-2 rows (fire ignition, world.cpp:1198-1206), reads up to 10 rows
below a chunk (pillar probe, world.cpp:1700-1702).  Temperature: one exchange of the temperature plane
after each tickTemperature (16 ghost rows either way).
"""
from typing import List, Tuple

from . import abi

HALO_ROWS = 16
DOWN = (-2, 2)   # rows relative to B sent upper -> lower after tk 1
UP = (-2, 10)    # rows relative to B sent lower -> upper after tk 3


def plan_strips(height: int, zone, n: int) -> List[Tuple[int, int]]:
    """[(first owned row, owned rows)] for n strips: interior boundaries at zone.y + k*32, the zone's 32-row
    bands dealt out as evenly as possible; strip 0 also owns the rows above the zone, strip n-1 those below."""
    zx, zy, zw, zh = zone
    bands = zh // abi.FSS_ZONE_ALIGN_Y
    if n < 1 or bands < n:
        raise ValueError(f"cannot cut {bands} bands of {abi.FSS_ZONE_ALIGN_Y} rows into {n} strips")
    cuts = [0]
    for r in range(1, n):
        cuts.append(zy + (bands * r // n) * abi.FSS_ZONE_ALIGN_Y)
    cuts.append(height)
    out = [(cuts[i], cuts[i + 1] - cuts[i]) for i in range(n)]
    for y0, rows in out:
        if rows < 2 * HALO_ROWS:
            raise ValueError("strip shorter than 32 rows")
    return out


def stored_rows(height: int, strip: Tuple[int, int]) -> Tuple[int, int]:
    """[lo, hi) global rows a strip world stores: owned rows plus HALO_ROWS ghost rows either side"""
    y0, rows = strip
    return max(0, y0 - HALO_ROWS), min(height, y0 + rows + HALO_ROWS)


def exchanges_after(tk: int):
    """(direction, (lo, hi) relative to the boundary) to run after sub-step tk, or None"""
    if tk == 1:
        return "down", DOWN
    if tk == 3:
        return "up", UP
    return None


def tick_schedule():
    """the per-tick sequence every execution mode follows: ('substep', iter, tk) and ('exchange', tk)"""
    for it in range(6):
        for tk in range(4):
            yield ("substep", it, tk)
            if exchanges_after(tk):
                yield ("exchange", it, tk)


def shift_rows(height: int, plan, rank: int, change_y: int):
    """World::tickChunks' grid shift (world.cpp:2617-2630) on strips: new[y] = old[y - change_y].  Only the rows a strip OWNS are
    current in its storage (the halo exchange refreshes just the ghost rows the rules look at), so the sources of its stored rows
    that lie outside its owned rows come from the neighbours.  Returns the global row ranges
      (fetch_from_upper, fetch_from_lower, give_to_lower, give_to_upper)
    as (lo, hi) pairs (empty: lo == hi), the same arithmetic as shift_rows() in csrc/fss_abi.cu; |change_y| must not exceed
    any strip's rows less HALO_ROWS."""
    n = len(plan)
    own_lo, own_hi = plan[rank][0], plan[rank][0] + plan[rank][1]
    s0, s1 = stored_rows(height, plan[rank])
    if abs(change_y) > min(r for _, r in plan) - HALO_ROWS:
        raise ValueError("shift longer than the shortest strip less its ghost rows")
    empty = (0, 0)
    u = (max(0, s0 - change_y), own_lo) if rank > 0 else empty
    gu = (own_lo, min(height, own_lo + HALO_ROWS - change_y)) if rank > 0 else empty
    lo_ = (own_hi, min(height, s1 - change_y)) if rank + 1 < n else empty
    gd = (max(0, own_hi - HALO_ROWS - change_y), own_hi) if rank + 1 < n else empty
    fix = lambda r: r if r[1] > r[0] else (r[0], r[0])
    return fix(u), fix(lo_), fix(gd), fix(gu)


class RankWorld:
    """One rank's strip of a world sharded over torch.distributed ranks (one process per GPU, NCCL).

    Construction is collective: rank 0 creates the NCCL unique id and broadcasts it over the existing
    torch.distributed group; the data path itself uses no torch collectives -- fss_tick exchanges halo rows
    with ncclSend/ncclRecv on the world's own stream."""

    def __init__(self, size, zone, rank, n_ranks, device, cells=None, **kw):
        import torch
        import torch.distributed as dist

        from . import world
        w, h = size
        self.plan = plan_strips(h, zone, n_ranks)
        self.strip = self.plan[rank]
        self.rank, self.n_ranks = rank, n_ranks
        self.world = world.World(cells=cells, zone=zone, size=size, strip=self.strip if n_ranks > 1 else None, device=device, **kw)
        if n_ranks > 1:
            uid = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                uid = torch.frombuffer(bytearray(world.comm_unique_id()), dtype=torch.uint8).clone()
            if dist.get_backend() == "nccl":
                uid = uid.cuda(device)
            dist.broadcast(uid, 0)
            self.world.comm_attach(bytes(uid.cpu().numpy().tobytes()), rank, n_ranks)

    def __getattr__(self, name):
        return getattr(self.world, name)
