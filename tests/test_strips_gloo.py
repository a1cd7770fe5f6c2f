"""The N > 1 path on CPU: two gloo ranks run the strip schedule of fallingsandsurvival_b200/strips.py
(plan_strips, tick_schedule, exchanges_after) with the oracle as the compute stand-in and torch.distributed
send/recv as the row transport, and must reproduce the undivided world bit for bit.  The same schedule is
what fss_tick executes over NCCL (tests/run_nccl_strips.py checks that one on GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import full_table, make_world
from fallingsandsurvival_b200 import abi, strips

SIZE = (264, 336)
TICKS = 6


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rows_tensor(view, lo, hi):
    return torch.from_numpy(np.ascontiguousarray(view[lo:hi]).view(np.uint8).reshape(-1).copy())


def _exchange(view, rank, n, plan, tk):
    """what fss_exchange_halo does, with gloo: after tk 1 rows flow down, after tk 3 rows flow up"""
    direction, (lo, hi) = strips.exchanges_after(tk)
    w = view.shape[1]
    nbytes = (hi - lo) * w * abi.CELL_DTYPE.itemsize
    own_lo, own_hi = plan[rank][0], plan[rank][0] + plan[rank][1]
    reqs, recv = [], None
    if direction == "down":
        if rank + 1 < n:
            reqs.append(dist.isend(_rows_tensor(view, own_hi + lo, own_hi + hi), rank + 1))
        if rank > 0:
            recv = (torch.empty(nbytes, dtype=torch.uint8), own_lo + lo, own_lo + hi)
            reqs.append(dist.irecv(recv[0], rank - 1))
    else:
        if rank > 0:
            reqs.append(dist.isend(_rows_tensor(view, own_lo + lo, own_lo + hi), rank - 1))
        if rank + 1 < n:
            recv = (torch.empty(nbytes, dtype=torch.uint8), own_hi + lo, own_hi + hi)
            reqs.append(dist.irecv(recv[0], rank + 1))
    for r in reqs:
        r.wait()
    if recv is not None:
        buf, a, b = recv
        view[a:b] = buf.numpy().view(abi.CELL_DTYPE).reshape(b - a, w)


def _exchange_temperature(view, rank, n, plan):
    own_lo, own_hi = plan[rank][0], plan[rank][0] + plan[rank][1]
    H = strips.HALO_ROWS
    reqs, recvs = [], []
    if rank + 1 < n:
        reqs.append(dist.isend(torch.from_numpy(view["temperature"][own_hi - H:own_hi].copy()), rank + 1))
        t = torch.empty((H, view.shape[1]), dtype=torch.int32)
        recvs.append((t, own_hi, own_hi + H))
        reqs.append(dist.irecv(t, rank + 1))
    if rank > 0:
        reqs.append(dist.isend(torch.from_numpy(view["temperature"][own_lo:own_lo + H].copy()), rank - 1))
        t = torch.empty((H, view.shape[1]), dtype=torch.int32)
        recvs.append((t, own_lo - H, own_lo))
        reqs.append(dist.irecv(t, rank - 1))
    for r in reqs:
        r.wait()
    for t, a, b in recvs:
        view["temperature"][a:b] = t.numpy()


def _worker(rank, n, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=n)
    from oracle import oraclesim
    cells, zone = make_world(SIZE, "c3", seed=21)
    plan = strips.plan_strips(SIZE[1], zone, n)
    own_lo, own_hi = plan[rank][0], plan[rank][0] + plan[rank][1]
    s_lo, s_hi = strips.stored_rows(SIZE[1], plan[rank])
    # a rank only has valid data in its stored rows: poison everything else so that a read outside them shows up
    local = cells.copy()
    poison = np.ones(SIZE[1], dtype=bool)
    poison[s_lo:s_hi] = False
    local["material"][poison] = 1
    local["temperature"][poison] = 12345
    o = oraclesim.OracleWorld(local, zone, defs=full_table(), particles=True)
    view = o.view()
    for t in range(TICKS):
        for op, it, tk in strips.tick_schedule():
            if op == "substep":
                o.substep_rows(t, it, tk, own_lo, own_hi)
            else:
                _exchange(view, rank, n, plan, tk)
        if t % 4 == 2:
            o.tick_temperature_rows(own_lo, own_hi)
            _exchange_temperature(view, rank, n, plan)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), view[own_lo:own_hi])
    np.save(os.path.join(out_dir, f"parts{rank}.npy"), o.drain_particles())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [2, 3])
def test_gloo_ranks_reproduce_the_undivided_world(n, tmp_path):
    from oracle import oraclesim
    port = _free_port()
    mp.spawn(_worker, args=(n, port, str(tmp_path)), nprocs=n, join=True)
    cells, zone = make_world(SIZE, "c3", seed=21)
    whole = oraclesim.OracleWorld(cells, zone, defs=full_table(), particles=True)
    for t in range(TICKS):
        whole.tick(t)
        if t % 4 == 2:
            whole.tick_temperature()
    ref = whole.cells()
    got = np.concatenate([np.load(tmp_path / f"rank{r}.npy") for r in range(n)])
    assert got.shape == ref.shape
    assert got.tobytes() == ref.tobytes(), f"{int((got != ref).sum())} cells differ"
    parts = np.concatenate([np.load(tmp_path / f"parts{r}.npy") for r in range(n)])
    parts = parts[abi.spawn_order(parts, zone)]
    assert parts.tobytes() == whole.drain_particles().tobytes()


def _shift_worker(rank, n, port, out_dir, shifts):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=n)
    cells, zone = make_world(SIZE, "c3", seed=5)
    H, W = cells.shape
    plan = strips.plan_strips(H, zone, n)
    own_lo, own_hi = plan[rank][0], plan[rank][0] + plan[rank][1]
    s_lo, s_hi = strips.stored_rows(H, plan[rank])
    local = cells.copy()
    for k, (cx, cy) in enumerate(shifts):
        # only the owned rows are current: everything else (ghost rows included) is poisoned before every shift
        stale = np.ones(H, dtype=bool)
        stale[own_lo:own_hi] = False
        local["material"][stale] = 2
        local["color"][stale] = 0xDEAD0000 + k
        u, lw, gd, gu = strips.shift_rows(H, plan, rank, cy)
        reqs, recvs = [], []
        for (a, b), peer in ((gd, rank + 1), (gu, rank - 1)):
            if b > a:
                reqs.append(dist.isend(_rows_tensor(local, a, b), peer))
        for (a, b), peer in ((u, rank - 1), (lw, rank + 1)):
            if b > a:
                t = torch.empty((b - a) * W * abi.CELL_DTYPE.itemsize, dtype=torch.uint8)
                recvs.append((t, a, b))
                reqs.append(dist.irecv(t, peer))
        for r in reqs:
            r.wait()
        src = local.copy()
        for t, a, b in recvs:
            src[a:b] = t.numpy().view(abi.CELL_DTYPE).reshape(b - a, W)
        valid = np.zeros(H, dtype=bool)  # what k_shift may read: the owned rows and what was fetched
        valid[own_lo:own_hi] = True
        for _, a, b in recvs:
            valid[a:b] = True
        new = local.copy()
        ys, xs = np.mgrid[s_lo:s_hi, 0:W]
        oy, ox = ys - cy, xs - cx
        ok = (ox >= 0) & (ox < W) & (oy >= 0) & (oy < H)
        assert valid[oy[ok]].all(), f"rank {rank}, shift {(cx, cy)}: a source row is neither owned nor fetched"
        new[ys[ok], xs[ok]] = src[oy[ok], ox[ok]]
        local = new
        np.save(os.path.join(out_dir, f"shift{k}_rank{rank}.npy"), local[s_lo:s_hi])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [2, 3])
def test_gloo_shift_grid_across_strips(n, tmp_path):
    """fss_shift_grid on strips, the row arithmetic (strips.shift_rows == shift_rows() of csrc/fss_abi.cu) with gloo as the
    transport: every STORED row of every rank, ghost rows included, must come out as the undivided shift leaves it"""
    shifts = [(3, 5), (-7, -40), (0, 33), (11, -1), (-2, 64), (40, -70), (0, 16), (5, -16)]
    port = _free_port()
    mp.spawn(_shift_worker, args=(n, port, str(tmp_path), shifts), nprocs=n, join=True)
    cells, zone = make_world(SIZE, "c3", seed=5)
    H, W = cells.shape
    plan = strips.plan_strips(H, zone, n)
    ref = cells.copy()
    for k, (cx, cy) in enumerate(shifts):
        ys, xs = np.mgrid[0:H, 0:W]
        oy, ox = ys - cy, xs - cx
        ok = (ox >= 0) & (ox < W) & (oy >= 0) & (oy < H)
        new = ref.copy()
        new[ok] = ref[oy[ok], ox[ok]]
        ref = new
        for r in range(n):
            s_lo, s_hi = strips.stored_rows(H, plan[r])
            got = np.load(tmp_path / f"shift{k}_rank{r}.npy")
            own_lo, own_hi = plan[r][0], plan[r][0] + plan[r][1]
            # rows whose source lies outside the world keep their content: owned rows keep real cells, ghost rows keep the poison
            a, b = got[own_lo - s_lo:own_hi - s_lo], ref[own_lo:own_hi]
            assert a.tobytes() == b.tobytes(), f"shift {k} {(cx, cy)}, rank {r}: {int((a != b).sum())} owned cells differ"
            keep = np.zeros(H, dtype=bool)
            keep[max(0, cy):min(H, H + cy)] = True  # rows that have a source inside the world
            rows = np.arange(s_lo, s_hi)
            sel = keep[rows]
            a, b = got[sel][:, max(0, cx):min(W, W + cx)], ref[rows[sel]][:, max(0, cx):min(W, W + cx)]
            assert a.tobytes() == b.tobytes(), f"shift {k} {(cx, cy)}, rank {r}: {int((a != b).sum())} stored cells with a source differ"
    with pytest.raises(ValueError):
        strips.shift_rows(H, plan, 0, 400)


def test_plan_strips_properties():
    zone = (16, 16, 16384, 16384 * 8)
    plan = strips.plan_strips(16384 * 8 + 32, zone, 8)
    assert plan[0][0] == 0 and plan[-1][0] + plan[-1][1] == 16384 * 8 + 32
    for (a, ra), (b, rb) in zip(plan[:-1], plan[1:]):
        assert a + ra == b and (b - zone[1]) % abi.FSS_ZONE_ALIGN_Y == 0
    assert max(r for _, r in plan) - min(r for _, r in plan) <= 32 + 32
    with pytest.raises(ValueError):
        strips.plan_strips(96, (16, 16, 64, 64), 3)
    ops = list(strips.tick_schedule())
    assert len([o for o in ops if o[0] == "substep"]) == 24 and len([o for o in ops if o[0] == "exchange"]) == 12
