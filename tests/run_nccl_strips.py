#!/usr/bin/env python3
"""Strip-sharded world over NCCL == the same world on one GPU (bit-exact).  Launch with
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/run_nccl_strips.py
Every rank owns one row strip on its own GPU; fss_tick exchanges ghost rows with ncclSend/ncclRecv.
Rank 0 also runs the undivided world and compares cells, digest and particle records."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers  # noqa: E402
from fallingsandsurvival_b200 import abi, strips, world  # noqa: E402


def main():
    rank, n, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    size = (520, 64 * 4 * n + 32)
    ticks = 12
    # a sand / water / stone world plus lava and hot gold ore; steam and fire arrive later, in ONE strip only (see below)
    from fallingsandsurvival_b200 import materials as M
    cells, zone = helpers.make_world(size, [(M.TEST_SAND, 25), (M.WATER, 20), (M.COBBLE_STONE, 10), (M.LAVA, 3), (M.GOLD_ORE, 3)], seed=31)
    cells["temperature"][cells["material"] == M.GOLD_ORE] = 600
    rw = strips.RankWorld(size, zone, rank, n, local, cells=cells, defs=helpers.full_table(), particles=True, track_dirty=True, track_flow=True)
    shifts = {3: (5, 40), 7: (-9, -70), 9: (0, 17)}  # World::tickChunks' camera shift, collective over the ranks (rows cross strips)

    def paint(w, t):
        # a host edit that only the LAST strip sees: steam and fire painted just below its upper boundary.  The steam rises into the
        # strip above, whose kernels were specialised for a world without GAS: the strips tell each other once per tick
        # (share_present in csrc/fss_abi.cu), so the upper strip switches kernels before the first steam cell reaches it
        if t != 4:
            return
        y = strips.plan_strips(size[1], zone, n)[-1][0] + 17  # (outside the 16 ghost rows the strip above stores: an edit inside them is made on both strips)
        patch = w.download(40, y, 200, 24)
        sel = (np.add.outer(np.arange(24), np.arange(200)) % 3) == 0
        patch["material"][sel] = M.STEAM
        patch["material"][~sel & ((np.add.outer(np.arange(24), np.arange(200)) % 7) == 0)] = M.FIRE
        w.upload(40, y, patch)

    for t in range(ticks):
        if rank == n - 1:
            paint(rw.world, t)
        rw.tick(t)
        if t % 4 == 2:
            rw.tick_temperature()
        if t in shifts:
            rw.shift_grid(*shifts[t])
        if t == 5:
            rw.render_dirty()  # consumes the marks so far: what is gathered below is what the later ticks left
    y0, rows = rw.strip
    # World::dirty on strips: the marks of the ghost rows travelled with the halo exchange; every strip renders the rows it owns
    marks = rw.world.dirty(0, y0, size[0], rows)
    moving, info = rw.render_dirty()
    layers = [rw.world.pixels(k, 0, y0, size[0], rows) for k in range(4)]
    mine = rw.world.download(0, y0, size[0], rows)
    digest = torch.tensor([rw.world.hash() & 0x7FFFFFFFFFFFFFFF, rw.world.hash() >> 63], dtype=torch.int64, device="cuda")
    parts = rw.world.drain_particles()
    gathered = [None] * n
    dist.all_gather_object(gathered, (y0, rows, mine.tobytes(), rw.world.hash(), parts.tobytes(), marks.tobytes(), moving, info, [a.tobytes() for a in layers]))
    ok = True
    if rank == 0:
        whole = world.World(cells, zone, defs=helpers.full_table(), particles=True, device=local, track_dirty=True, track_flow=True)
        for t in range(ticks):
            paint(whole, t)
            whole.tick(t)
            if t % 4 == 2:
                whole.tick_temperature()
            if t in shifts:
                whole.shift_grid(*shifts[t])
            if t == 5:
                whole.render_dirty()
        ref = whole.cells()
        ref_marks = whole.dirty()
        ref_moving, ref_info = whole.render_dirty()
        ref_layers = [whole.pixels(k) for k in range(4)]
        got_marks = np.zeros_like(ref_marks)
        got_moving, got_info = np.zeros_like(ref_moving), (0, 0, 0)
        got_layers = [np.zeros_like(a) for a in ref_layers]
        out = np.zeros_like(ref)
        hsum = 0
        recs = []
        for gy0, grows, raw, h, praw, mraw, mov, inf, lraw in gathered:
            out[gy0:gy0 + grows] = np.frombuffer(raw, dtype=abi.CELL_DTYPE).reshape(grows, size[0])
            got_marks[gy0:gy0 + grows] = np.frombuffer(mraw, dtype=np.uint8).reshape(grows, size[0])
            got_moving = got_moving + mov
            got_info = tuple(a | b for a, b in zip(got_info, inf))
            for k in range(4):
                got_layers[k][gy0:gy0 + grows] = np.frombuffer(lraw[k], dtype=np.uint8).reshape(grows, size[0], 4)
            hsum = (hsum + h) & 0xFFFFFFFFFFFFFFFF
            recs.append(np.frombuffer(praw, dtype=abi.PARTICLE_DTYPE))
        recs = np.concatenate(recs)
        recs = recs[abi.spawn_order(recs, zone)]
        dirty_ok = (got_marks.tobytes() == ref_marks.tobytes() and np.array_equal(got_moving, ref_moving) and got_info == ref_info
                    and all(a.tobytes() == b.tobytes() for a, b in zip(got_layers, ref_layers)))
        ok = out.tobytes() == ref.tobytes() and hsum == whole.hash() and recs.tobytes() == whole.drain_particles().tobytes() and dirty_ok
        print(f"World::dirty ({int(ref_marks.sum())} marks), pixel layers, counts over the strips {'==' if dirty_ok else '!='} single GPU", flush=True)
        print(f"nccl strips n={n}: cells {'==' if out.tobytes() == ref.tobytes() else '!='} single GPU, digest sum {hsum:016x} vs {whole.hash():016x}, "
              f"{len(recs)} particle records -> {'OK' if ok else 'MISMATCH'}", flush=True)
        whole.close()
    # the same world without particle emission: the register-window kernels (fire through its pre-pass) on strips, the edge band
    # beside the interior bands
    rw2 = strips.RankWorld(size, zone, rank, n, local, cells=cells, defs=helpers.full_table())
    if rank == n - 1:
        paint(rw2.world, 4)
    for t in range(8):
        rw2.tick(t)
    h2 = [None] * n
    dist.all_gather_object(h2, rw2.world.hash())
    if rank == 0:
        whole2 = world.World(cells, zone, defs=helpers.full_table(), device=local)
        paint(whole2, 4)
        for t in range(8):
            whole2.tick(t)
        ok2 = (sum(h2) & 0xFFFFFFFFFFFFFFFF) == whole2.hash()
        print(f"without particle emission (window kernels): digest sum {(sum(h2) & 0xFFFFFFFFFFFFFFFF):016x} vs {whole2.hash():016x} -> {'OK' if ok2 else 'MISMATCH'}", flush=True)
        ok = ok and ok2
        whole2.close()
    rw2.world.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    rw.world.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
