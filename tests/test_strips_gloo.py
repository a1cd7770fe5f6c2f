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
