#!/usr/bin/env python3
"""Benchmark of the World::tick hot path on B200 (contract: see the task statement, section 4).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c5] [--size S] [--impl b200|reference]

One "step" = one full World::tick (6 iter x 4 tk checkerboard sub-steps x 3 scans, world.cpp:1091-2041) over
the whole tick zone; workload c3 also runs World::tickTemperature every 4th step (Game.cpp:2759).
metric = cell-updates/s = zone cells x steps / seconds, whole job (all ranks).

N = 1   workload "c2" (BASELINE.json configs[1]): 16384 x 16384 zone, 25 % sand / 20 % water / 10 % stone; the line also
        carries `extra.c3` (configs[2], with roofline_k2 for k_temperature) and `extra.c5` (configs[4] at 32768 x 32768,
        early-out on and off), a few ticks each.
N > 1   workload "c4" (configs[3]): ONE 65536 x 65536 zone cut into N row strips (strong scaling), halo rows exchanged
        over NCCL after tk 1 / tk 3 of every iter (one process per GPU under torchrun).
Every line carries `digest`: fss_hash of the grid after the timed ticks, summed over the ranks -- the same for every N.
The grid is generated on the device (fss_fill_synthetic); "data": "synthetic".

Besides `value` (grid resident in HBM) the line carries
  e2e          the same metric through the C ABI with HOST buffers (fss_tick_host): every step the whole grid goes up
               from pinned host memory, is ticked, and comes back (the reference-side shim's worst case), pipelined
               in row bands so that both PCIe directions and the kernels overlap
  roofline     K1 (k_substep) against the measured HBM copy bandwidth, algorithmic bytes as DESIGN.md defines
  cpu_baseline the reference's own tick (oracle/_ref) or the oracle port on this box's host cores, N = 1 only
`--impl reference` times only that CPU implementation (rank 0), same metric and unit.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ZONE_PER_GPU = 16384
BORDER = 16
CELL_BYTES_PLANES = 20   # 5 planes x 4 B per cell in HBM
ALGO_BYTES_PER_CELL_UPDATE = 32.0  # 16-B packed cell read once + written once per tick
SUBSTEPS = 24
CPU_SAMPLE_ZONE = (2048, 1024)


def mixes():
    from fallingsandsurvival_b200 import worlds
    return {"c2": worlds.MIX_C2, "c3": worlds.MIX_C3, "c4": worlds.MIX_C2, "c5": worlds.MIX_C2}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pin_to_gpu_numa_node(local):
    """Bind this rank's host threads to the CPUs next to its GPU (NVML's ideal CPU affinity = the GPU's NUMA node) BEFORE any
    pinned host buffer is allocated: first touch then places the staging memory on that node, and N ranks stop sharing one
    socket's memory controllers and one set of cores (round 1: every rank ran on CPUs 0-31 / NUMA node 0)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        try:
            node = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            node = None
        cpus = sorted(os.sched_getaffinity(0))
        return {"numa_node": node, "cpus": f"{cpus[0]}-{cpus[-1]} ({len(cpus)})"}
    except Exception as e:  # no NVML / not permitted in this container: run unpinned and say so
        return {"numa_node": None, "cpus": f"unpinned: {type(e).__name__}"}


def ncu_traffic(workload, size):
    """dram__bytes_read.sum + dram__bytes_write.sum per movement launch (mean over the 24 launches of one tick) from the
    committed ncu capture of this very workload on this round's build (profiles/r02_k1_traffic.json; a number taken under a
    profiler cannot be produced inside a timed run), or None when the capture does not match the run"""
    if workload != "c2" or size != ZONE_PER_GPU:
        return None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_k1_traffic.json")) as f:
            d = json.load(f)
        return float(d["dram_bytes_per_launch_mean"]), d["source"]
    except (OSError, KeyError, ValueError):
        return None, None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "B200_PROFILING.md fallback"


# ---- the reference's CPU tick on a bounded sample ---------------------------------------------------------------
def cpu_tick_throughput(workload, steps, warmup, temperature, variants=False):
    """cell-updates/s of the reference's own World::tick (oracle/_ref, fastest variant) or of the oracle port on a
    2048x1024 zone of the same fill, all host threads."""
    from fallingsandsurvival_b200 import worlds
    from oracle import oraclesim, refsim
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    zw, zh = CPU_SAMPLE_ZONE
    size = (zw + 2 * BORDER, zh + 2 * BORDER)
    defs = helpers.full_table()
    cells = worlds.synthetic_world(size[0], size[1], mixes()[workload], seed=0xF55, defs=defs)
    zone = (BORDER, BORDER, zw, zh)
    cores = os.cpu_count() or 1
    if refsim.available("c128"):
        kind, w = "reference", refsim.RefWorld(cells, zone, threads=cores, variant="c128")
        what = "oracle/_ref c128: the reference's World::tick lines, CHUNK 128, keyed RNG"
    else:
        kind, w = "port", oraclesim.OracleWorld(cells, zone, defs=defs)
        what = "oracle/fss_oracle.c (OpenMP)"
    for t in range(warmup):
        w.tick(t)
    t0 = time.perf_counter()
    for t in range(warmup, warmup + steps):
        w.tick(t)
        if temperature and t % 4 == 2:
            w.tick_temperature()
    dt = time.perf_counter() - t0
    w.close()
    out = {"value": zw * zh * steps / dt, "unit": "cell-updates/s", "cores": cores, "kind": kind,
           "sample": f"{steps} ticks of a {zw}x{zh} zone, same fill as the workload, after {warmup} warm-up ticks; {what}",
           "ms_per_step": dt / steps * 1e3}
    if variants and kind == "reference":
        # the other two CPU variants, always labelled (few ticks each: they are slow)
        out["variants"] = {"counter-RNG CHUNK 128 (value above)": out["value"]}
        for name, variant, threads, nt in (("as shipped: libc rand(), CHUNK 128, 6 threads (world.cpp:64)", "shipped", 6, 3),
                                           (f"parity oracle schedule: counter-RNG CHUNK 4x16, {cores} threads", "c4x16", cores, 5)):
            if not refsim.available(variant):
                continue
            v = refsim.RefWorld(cells, zone, threads=threads, variant=variant)
            v.tick(0)
            t0 = time.perf_counter()
            for t in range(1, 1 + nt):
                v.tick(t)
            out["variants"][name] = zw * zh * nt / (time.perf_counter() - t0)
            v.close()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = default_workload(args)
    r = cpu_tick_throughput("c2" if workload == "c4" else workload, args.steps, args.warmup, workload == "c3")
    zw, zh = main_zone(workload, args.size, args.gpus)
    cfg = workload_config(workload, args.gpus, (zw, zh), early_out=not args.no_early_out)
    cfg["timed_zone_cells"] = list(CPU_SAMPLE_ZONE)  # what this arm's steps actually cover: a bounded sample of the workload
    line = {
        "impl": "reference", "metric": "cell-updates/sec", "value": r["value"], "unit": "cell-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if workload == "c4" else "weak", "vs_baseline": None, "dtype": "u32+f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


FILLS = {
    "c2": "25% TEST_SAND / 20% WATER / 10% COBBLE_STONE, rest AIR (hash(seed, x, y) % 100)",
    "c3": "c2 fill + 3% LAVA 2% FIRE 3% GOLD_ORE 3% STEAM 5% SOFT_DIRT",
    "c4": "25% TEST_SAND / 20% WATER / 10% COBBLE_STONE, rest AIR (hash(seed, x, y) % 100)",
    "c5": "the lower 90% of the rows COBBLE_STONE; 0.1% TEST_SAND scattered in the air above; 64 boxes of 256 x 256 cells with the c2 fill (pockets in the stone)",
}


def workload_config(workload, n, zone_wh, early_out=True):
    return {"workload": {"c4": "BASELINE.json configs[3]: 65536 x 65536 world strip-sharded with per-sub-step halo exchange over NVLink (NCCL)",
                         "c2": "BASELINE.json configs[1]: synthetic random sand/water/stone world, movement-only tick",
                         "c3": "BASELINE.json configs[2]: configs[1] + lava/fire/gold ore/steam/soft dirt, tickTemperature every 4th tick",
                         "c5": "BASELINE.json configs[4]: mostly settled world with sparse active boxes"}[workload],
            "zone_cells": list(zone_wh), "zone_cells_per_gpu": [zone_wh[0], zone_wh[1] // n], "strips": n,
            "fill": FILLS[workload], "particles": "off",
            "early_out": "on (chunks verified settled are skipped; results identical)" if early_out else "off (FSS_NO_EARLY_OUT: every micro-chunk visited in every sub-step)",
            "l2": "working set (5 planes x 4 B x cells) is far larger than the 126 MB L2; no flush needed"}


def default_workload(args):
    """N = 1: c2 (BASELINE configs[1]); N > 1: c4 (configs[3], ONE 65536 x 65536 world cut into N strips: strong scaling)"""
    if args.workload:
        return args.workload
    return "c2" if args.gpus == 1 else "c4"


def main_zone(workload, size, n):
    if workload == "c4":
        s = size or 65536
        return s, s
    s = size or (32768 if workload == "c5" else ZONE_PER_GPU)
    return s, s * n


def measure(args, workload, size, steps, warmup, early_out=True, e2e_steps=0, per_launch=True, rank=0, n=1, local=0):
    """Build the workload's world on the device(s), time `steps` ticks after `warmup`, return a dict of raw measurements."""
    import torch
    import torch.distributed as dist

    from fallingsandsurvival_b200 import materials as M
    from fallingsandsurvival_b200 import strips
    import helpers

    ZW, ZH = main_zone(workload, size, n)
    W, H = ZW + 2 * BORDER, ZH + 2 * BORDER
    zone = (BORDER, BORDER, ZW, ZH)
    defs = helpers.full_table()
    temperature = workload == "c3"
    rw = strips.RankWorld((W, H), zone, rank, n, local, defs=defs, early_out=early_out)
    g = rw.world
    stream = torch.cuda.Stream()  # a real stream: handle 0 would mean "the world's own stream" to fss_set_stream
    torch.cuda.set_stream(stream)
    g.set_stream(stream.cuda_stream)
    fill = dict(seed=0xF55, border=BORDER)
    if workload == "c5":
        # "rows below 90 % height COBBLE_STONE" = the lower nine tenths of the world are stone (mostly settled);
        # the active boxes are pockets of the c2 mix inside it, the sparse sand falls through the air above
        fill.update(solid_from_row=int(H * 0.1), n_boxes=64 if ZW >= 4096 else 4, box_size=256, sparse_material=M.TEST_SAND, sparse_per_mille=1)
    g.fill_synthetic(mixes()[workload], **fill)
    g.sync()

    def barrier():
        torch.cuda.synchronize()
        if n > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if n == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    tick_ct = 0

    def step():
        nonlocal tick_ct
        g.tick(tick_ct)
        if temperature and tick_ct % 4 == 2:
            g.tick_temperature()
        tick_ct += 1

    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = g.stats().kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    barrier()
    out = {"zone": (ZW, ZH), "ms": max_over_ranks(ev0.elapsed_time(ev1)), "launches": g.stats().kernel_launches - launches0,
           "clocks": sampler.stop() if rank == 0 else None, "stored_cells": (g.stored[1] - g.stored[0]) * g.stats().storage_pitch}

    # the grid after warmup + steps ticks, as one number: fss_hash per strip, summed (mod 2^64) over the ranks.  The same
    # world at a different GPU count must give the same digest: the driver-visible proof that the strips are bit-exact.
    digest = g.hash()
    if n > 1:
        t = torch.tensor([digest - (1 << 64) if digest >= (1 << 63) else digest], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)  # two's complement wrap-around = addition mod 2^64
        digest = int(t.item()) & 0xFFFFFFFFFFFFFFFF
    out["digest"] = f"{digest:016x}"
    out["digest_after_ticks"] = tick_ct

    # K1 alone, per launch, for the roofline: one more tick with an event pair around every sub-step launch
    out["k1_ms"] = []
    if n == 1 and per_launch:
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(SUBSTEPS + 1)]
        evs[0].record(stream)
        k = 0
        for it in range(6):
            for tk in range(4):
                g.substep(tick_ct, it, tk)
                k += 1
                evs[k].record(stream)
        tick_ct += 1
        torch.cuda.synchronize()
        out["k1_ms"] = [evs[i].elapsed_time(evs[i + 1]) for i in range(SUBSTEPS)]

    out["k2_ms"] = None
    if temperature and n == 1:
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.tick_temperature()
        torch.cuda.synchronize()
        t0e.record(stream)
        for _ in range(5):
            g.tick_temperature()
        t1e.record(stream)
        torch.cuda.synchronize()
        out["k2_ms"] = t0e.elapsed_time(t1e) / 5

    # ---- e2e: host buffers through the C ABI --------------------------------------------------------------
    y0, y1 = g.stored
    rows = y1 - y0
    out["e2e"] = None
    host = None
    if e2e_steps > 0:
        try:  # (c4 on two GPUs pins 43 GB per rank: a box without that much host memory still gets its `value` line)
            host = torch.empty((rows, W, 20), dtype=torch.uint8, pin_memory=True).numpy().view(helpers.abi.CELL_DTYPE).reshape(rows, W)
        except RuntimeError as e:
            out["e2e_skipped"] = f"pinned host buffer of {rows * W * 20 / 1e9:.1f} GB per rank: {str(e).splitlines()[0]}"
    if n > 1:  # every rank takes the same path
        ok = torch.tensor([1 if host is not None else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not int(ok.item()):
            host = None
    if host is not None:
        g.download(0, y0, W, rows, out=host)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            if temperature or n > 1:
                g.upload(0, y0, host)
                step()
                g.download(0, y0, W, rows, out=host)
            else:  # World::tick() with the grid on the host: upload, the 24 sub-steps and download overlap band by band
                g.tick_host(host, tick_ct)
                tick_ct += 1
        e1.record(stream)
        barrier()
        e2e_wall = time.perf_counter() - t0
        e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), e2e_wall * 1e3))

        # the same loop with the rectangle the reference game itself keeps loaded (1024 x 768 cells, Game.cpp:450-452)
        # instead of the whole grid: what a shim that tracks its edits pays per tick
        wx, wy, ww, wh = BORDER, max(y0, BORDER), min(1024, W - 2 * BORDER), min(768, rows - 2 * BORDER)
        win = torch.empty((wh, ww, 20), dtype=torch.uint8, pin_memory=True).numpy().view(helpers.abi.CELL_DTYPE).reshape(wh, ww)
        g.download(wx, wy, ww, wh, out=win)
        barrier()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record(stream)
        for _ in range(e2e_steps):
            g.upload(wx, wy, win)
            step()
            g.download(wx, wy, ww, wh, out=win)
        w1.record(stream)
        barrier()
        e2e_win_ms = max_over_ranks(w0.elapsed_time(w1))
        cells = ZW * ZH
        out["e2e"] = {"value": cells * e2e_steps / (e2e_ms / 1e3), "unit": "cell-updates/s",
                      "h2d_bytes_per_step": int(rows * W * 20), "d2h_bytes_per_step": int(rows * W * 20), "steps": e2e_steps,
                      "ms_per_step": e2e_ms / e2e_steps,
                      "what": ("per step: fss_upload_rect(whole stored grid of the rank, pinned host) + fss_tick + fss_download_rect(whole stored grid)"
                               if temperature or n > 1 else
                               "per step: fss_tick_host(whole grid in pinned host memory): the same upload + 24 sub-steps + download, "
                               "pipelined in row bands on three streams (both PCIe directions and the kernels overlap)")}
        out["e2e_window"] = {"value": cells * e2e_steps / (e2e_win_ms / 1e3), "unit": "cell-updates/s", "ms_per_step": e2e_win_ms / e2e_steps,
                             "h2d_bytes_per_step": int(ww * wh * 20), "d2h_bytes_per_step": int(ww * wh * 20),
                             "what": "same, but only a 1024 x 768 rectangle (the reference game's loaded window) crosses PCIe each step"}
        del host, win
    g.close()
    torch.cuda.empty_cache()
    return out


def k1_roofline(k1_ms, zone_cells, workload, size):
    peak, peak_src = measured_peak()
    avg = float(np.mean(k1_ms))
    algo = ALGO_BYTES_PER_CELL_UPDATE * zone_cells / SUBSTEPS  # bytes per k_substep launch
    sweep = 2.0 * CELL_BYTES_PLANES * zone_cells / 4            # active micro-chunks read + written once, all planes
    traffic, traffic_src = ncu_traffic(workload, size)
    return {
        "kernel": "K1 movement sub-step: k_substep_win<SAND> (iters 0-1 of the c2 fill) and k_substep_win<liquid> (iters 2-5), mean over the 24 launches of a tick", "bound": "hbm",
        "achieved": algo / (avg / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
        "frac": algo / (avg / 1e3) / 1e9 / peak, "traffic": traffic, "traffic_source": traffic_src,
        "dram_achieved": (traffic / (avg / 1e3) / 1e9) if traffic else None,
        "dram_frac": (traffic / (avg / 1e3) / 1e9 / peak) if traffic else None, "peak_source": peak_src,
        "launch_ms_avg": avg, "launch_ms_by_iter": [float(np.sum(k1_ms[4 * i:4 * i + 4])) for i in range(6)],
        "algorithmic_bytes_per_launch": algo,
        "note": "algorithmic = 32 B per cell-update (whole tick) x zone cells / 24 launches per tick; the shipped schedule "
                "sweeps the active quarter of the grid once per launch: see sweep_*",
        "sweep_bytes_per_launch": sweep, "sweep_achieved": sweep / (avg / 1e3) / 1e9, "sweep_frac": sweep / (avg / 1e3) / 1e9 / peak,
    }


def k2_roofline(k2_ms, stored_cells):
    peak, _ = measured_peak()
    return {"kernel": "k_temperature", "bound": "hbm", "launch_ms": k2_ms, "unit": "GB/s",
            "achieved": 12.0 * stored_cells / (k2_ms / 1e3) / 1e9, "peak": peak,
            "frac": 12.0 * stored_cells / (k2_ms / 1e3) / 1e9 / peak,
            "note": "12 B per cell-temperature-update (read temp + mf, write temp) x stored cells"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=["c2", "c3", "c4", "c5"],
                    help="default: c2 at --gpus 1, c4 (BASELINE configs[3]: ONE 65536 x 65536 zone cut into --gpus strips, strong scaling) above")
    ap.add_argument("--size", type=int, default=None, help="zone cells per GPU along each axis (multiple of 32); c4: the whole zone")
    ap.add_argument("--e2e-steps", type=int, default=None, help="default 5 (3 for c4, whose strips are tens of GB per rank)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra.c3 / extra.c5 records of the default N = 1 line")
    ap.add_argument("--no-early-out", action="store_true", help="visit every micro-chunk in every sub-step (FSS_NO_EARLY_OUT)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))

    rank = int(os.environ.get("RANK", "0"))
    n = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert n == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={n}: launch with torchrun --nproc-per-node {args.gpus}"
    affinity = pin_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    if n > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    workload = default_workload(args)
    explicit = args.workload is not None or args.size is not None
    e2e_steps = args.e2e_steps if args.e2e_steps is not None else (3 if workload == "c4" else 5)
    m = measure(args, workload, args.size, args.steps, args.warmup, early_out=not args.no_early_out, e2e_steps=e2e_steps,
                rank=rank, n=n, local=local)
    ZW, ZH = m["zone"]
    zone_cells = ZW * ZH
    line = None
    if rank == 0:
        line = {
            "metric": "cell-updates/sec", "value": zone_cells * args.steps / (m["ms"] / 1e3), "unit": "cell-updates/s", "n_gpus": n, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["ms"] / args.steps, "higher_is_better": True, "scaling": "strong" if workload == "c4" else "weak",
            "vs_baseline": None, "dtype": "u32+f32", "data": "synthetic", "config": workload_config(workload, n, (ZW, ZH), early_out=not args.no_early_out),
            "gpu_launches": int(m["launches"]), "clocks": m["clocks"], "host_affinity_rank0": affinity,
            "digest": m["digest"], "digest_what": f"fss_hash of the whole grid (sum over ranks mod 2^64) after {m['digest_after_ticks']} ticks; "
                                                  "identical for every --gpus N on the same workload",
        }
        if m["e2e"]:
            line["e2e"] = m["e2e"]
            line["e2e_window"] = m["e2e_window"]
        if m.get("e2e_skipped"):
            line["e2e_skipped"] = m["e2e_skipped"]
        if m["k1_ms"]:
            line["roofline"] = k1_roofline(m["k1_ms"], zone_cells, workload, ZW)
        if m["k2_ms"] is not None:
            line["roofline_k2"] = k2_roofline(m["k2_ms"], m["stored_cells"])

    # ---- the other single-GPU BASELINE configs, a few ticks each (default N = 1 run only) -------------------------
    if n == 1 and not explicit and not args.no_extras:
        extra = {}
        c3 = measure(args, "c3", None, 8, 4, e2e_steps=0)
        extra["c3"] = {"config": workload_config("c3", 1, c3["zone"]), "steps": 8, "warmup": 4, "ms_per_step": c3["ms"] / 8,
                       "value": c3["zone"][0] * c3["zone"][1] * 8 / (c3["ms"] / 1e3), "unit": "cell-updates/s", "digest": c3["digest"],
                       "launch_ms_by_iter": [float(np.sum(c3["k1_ms"][4 * i:4 * i + 4])) for i in range(6)],
                       "roofline_k2": k2_roofline(c3["k2_ms"], c3["stored_cells"]),
                       "note": "tickTemperature runs in the timed region every 4th tick (Game.cpp:2759); roofline_k2 = k_temperature alone, 5 launches"}
        line["roofline_k2"] = extra["c3"]["roofline_k2"]
        on = measure(args, "c5", None, 8, 6, early_out=True, e2e_steps=0, per_launch=False)
        off = measure(args, "c5", None, 4, 2, early_out=False, e2e_steps=0, per_launch=False)
        cells5 = on["zone"][0] * on["zone"][1]
        extra["c5"] = {"config": workload_config("c5", 1, on["zone"]), "steps": 8, "warmup": 6,
                       "ms_per_step": on["ms"] / 8, "value": cells5 * 8 / (on["ms"] / 1e3), "unit": "cell-updates/s", "digest": on["digest"],
                       "early_out_off": {"steps": 4, "warmup": 2, "ms_per_step": off["ms"] / 4, "value": cells5 * 4 / (off["ms"] / 1e3), "digest": off["digest"]},
                       "early_out_speedup": (off["ms"] / 4) / (on["ms"] / 8),
                       "note": "effective cell-updates/s (zone cells per tick, skipped or not); the two digests are taken after different tick counts"}
        line["extra"] = extra

    if rank == 0:
        if n == 1 and not args.no_cpu:
            try:
                c = cpu_tick_throughput(workload, 10, 2, workload == "c3", variants=True)
                line["cpu_baseline"] = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample", "variants") if k in c}
            except Exception as e:  # the checker libraries are optional on a box without them
                line["cpu_baseline"] = {"value": None, "unit": "cell-updates/s", "cores": os.cpu_count(), "kind": "port", "sample": f"unavailable: {e}"}
        print(json.dumps(line), flush=True)
    if n > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
