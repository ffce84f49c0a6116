#!/usr/bin/env python
"""bench.py -- fp64 cell-residual evaluations per second of the finite-volume hot path.

    python bench.py --gpus 1 --steps K --warmup W [--n N]        # ibis_b200 (CUDA) arm
    python bench.py --impl reference ...                         # reference Kokkos-OpenMP arm

Workload (BASELINE.json configs[4]): N^3 uniform hexes on the unit cube
"in unstructured form", ideal gas R=287, ausmdv, reconstruction order 2, Barth-Jespersen,
rho_u, inviscid, ssp-rk3, cfl 0.5, supersonic inflow on west/south/bottom, outflow on
east/north/top, perturbed initial field (ii).  One step = one RungeKutta::take_step =
3 residual evaluations + dt estimate + stage updates; value = cells * 3 * K / time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

ALG_BYTES_G, ALG_BYTES_F = 392, 638       # pass G, pass F per cell
ALG_BYTES_EVAL = ALG_BYTES_G + ALG_BYTES_F


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def box_config():
    import cases
    from ibis_b200 import config as cfg
    return cfg.make_config(3, cases.box3_bcs(), flux_calculator="ausmdv", reconstruction_order=2,
                           solver=cfg.runge_kutta(method="ssp-rk3", cfl=0.5))


def box_grid_host(n):
    """Reference-style path (SU2 description -> GridBlock arrays) for the CPU arm."""
    from ibis_b200 import grid as gridmod
    from ibis_b200 import su2
    mesh = su2.box_hex(n)
    return mesh, gridmod.build_grid(mesh, ["west", "east", "south", "north", "bottom", "top"])


def box_grid_fast(n, rank=0, world=1):
    """The same box (bit-identical arrays, tests/test_box_grid.py) from the closed-form C++
    generator; with world > 1 the global box is n x n x (n * world) with lengths (1, 1, world)
    and this rank owns the z-slab [rank * n, (rank + 1) * n) plus its two-ring halo."""
    from ibis_b200.boxgrid import BoxGrid
    if world == 1:
        return BoxGrid(n)
    return BoxGrid(n, n, n * world, lengths=(1.0, 1.0, float(world)), slab=(rank * n, (rank + 1) * n))


def cpu_reference_rate(n_cpu, min_seconds=10.0):
    """The reference's own Kokkos-OpenMP residual (oracle/_ref) on an n_cpu^3 box: evals/s."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cases
    import refharness as rh
    if not rh.available():
        return None
    mesh, g = box_grid_host(n_cpu)
    path = f"/tmp/bench_box_{n_cpu}.su2"
    mesh.write(path)
    rc = rh.RefCase(path, box_config())
    prim = cases.perturbed_state(g, 3)
    t1 = rc.time_dudt(prim, 1)
    nrep = max(3, int(min_seconds / max(t1, 1e-6)))
    nrep = min(nrep, 200)
    t = rc.time_dudt(prim, nrep)
    cores = rc.rl.num_threads
    rc.close()
    return {"value": g.n_cells * nrep / t, "unit": "cell-residual evals/s", "cores": cores,
            "kind": "reference",
            "sample": f"{nrep} FiniteVolume::compute_dudt calls of the compiled reference (Kokkos-OpenMP, "
                      f"{cores} threads) on the {n_cpu}^3 = {g.n_cells}-cell box, same config and field"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference_rate(args.n_cpu, min_seconds=max(2.0, 4.0 * args.steps))
    if cb is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libibis_ref.so not built"}))
        return
    n_cells = args.n_cpu ** 3
    line = {"impl": "reference", "metric": "fp64 cell-residual evals/sec", "value": cb["value"],
            "unit": "cell-residual evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 3.0 * n_cells / cb["value"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"box {args.n_cpu}^3 hex (bounded sample of the box workload), ausmdv, order 2, "
                                   "barth_jespersen, rho_u, inviscid"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "cell-residual evals/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_gpu(args):
    import torch
    import cases
    from ibis_b200 import config as cfg
    from ibis_b200.lib import Context

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n = args.n
    config = box_config()
    t0 = time.time()
    g = box_grid_fast(n, rank, world)
    prim = cases.perturbed_state(g, 3)
    t_grid = time.time() - t0
    t0 = time.time()
    ctx = Context(g, config, device=local)
    if world > 1:
        import torch.distributed as dist
        ids = [ctx.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0], rank, world)
        ctx.halo_setup([rank + d for d in g.peer_dir], g.send_offsets, g.send_cells, g.recv_offsets, g.recv_cells)
    n_owned = int(g.n_owned)
    fs, cq = ctx.fs_create(), ctx.cq_create()
    ctx.fs_upload(fs, prim)
    ctx.prim_to_cons(cq, fs)
    ctx.sync()
    t_setup = time.time() - t0
    tab = cfg.BUTCHER["ssp-rk3"]
    stream = torch.cuda.ExternalStream(ctx.stream_ptr())

    def step(dt_prev):
        dt, _ = ctx.rk_step(fs, cq, tab["a"], tab["b"], 0.5, dt_prev, 1.0)
        return dt

    dt_prev = sys.float_info.max
    for _ in range(args.warmup):
        dt_prev = step(dt_prev)
    ctx.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    ctx.profile_enable(True)
    ctx.profile_read()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            dt_prev = step(dt_prev)
        e1.record(stream)
    ctx.sync()
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    prof_ms, prof_n = ctx.profile_read()
    ctx.profile_enable(False)
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    bad = ctx.count_bad_cells(fs)

    # end-to-end: host-resident state in pinned memory, uploaded and read back every step
    pin_in = torch.from_numpy(prim).pin_memory()
    pin_out = torch.empty_like(pin_in).pin_memory()
    n_e2e = max(1, min(args.steps, 3))
    ctx.fs_upload_ptr(fs, pin_in.data_ptr())
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        ctx.fs_upload_ptr(fs, pin_in.data_ptr())
        ctx.prim_to_cons(cq, fs)
        ctx.rk_step(fs, cq, tab["a"], tab["b"], 0.5, sys.float_info.max, 1.0)
        ctx.fs_download_ptr(fs, pin_out.data_ptr())
    ctx.sync()
    t_e2e = (time.perf_counter() - t0) / n_e2e
    if world > 1:
        t = torch.tensor([t_e2e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())

    n_cells_total = n_owned * world
    ms_per_step = ms_total / args.steps
    value = n_cells_total * 3.0 * args.steps / (ms_total * 1e-3)
    peak, peak_kind = measured_peak()
    tF = prof_ms[1] / max(int(prof_n[1]), 1) * 1e-3
    tG = prof_ms[0] / max(int(prof_n[0]), 1) * 1e-3
    achieved_F = ALG_BYTES_F * n_owned / tF / 1e9
    line = {
        "metric": "fp64 cell-residual evals/sec", "value": value, "unit": "cell-residual evals/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"box {n}^3 = {n_owned} hex cells per GPU in unstructured form "
                               f"(BASELINE.json configs[4]{'' if n == 400 else ', reduced size'}), ausmdv, order 2, "
                               "barth_jespersen, rho_u, inviscid, ssp-rk3 (3 residual evaluations per step), "
                               "perturbed field; working set " + f"{n_owned * 1030 / 1e9:.1f} GB per evaluation, "
                               "larger than L2" if n_owned * 1030 > 2 * 126e6 else f"box {n}^3 (L2 resident)",
                   "cells_per_gpu": n_owned, "parallelism": f"dd{world} (z-slabs, two-ring halo)" if world > 1 else "dd1",
                   "fmad": False, "grid_build_s": round(t_grid, 2), "setup_s": round(t_setup, 2),
                   "bad_cells": bad},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_residual (pass F: reconstruct+flux+integral)",
                     "achieved": achieved_F, "peak": peak, "unit": "GB/s", "frac": achieved_F / peak,
                     "traffic": None, "peak_kind": peak_kind,
                     "algorithmic_bytes_per_cell": ALG_BYTES_F, "avg_launch_ms": tF * 1e3,
                     "pass_G": {"kernel": "k_grad_limit", "avg_launch_ms": tG * 1e3,
                                "achieved": ALG_BYTES_G * n_owned / tG / 1e9 if tG > 0 else None,
                                "algorithmic_bytes_per_cell": ALG_BYTES_G},
                     "evaluation": {"achieved": ALG_BYTES_EVAL * n_owned / (tF + tG) / 1e9,
                                    "frac": ALG_BYTES_EVAL * n_owned / (tF + tG) / 1e9 / peak,
                                    "algorithmic_bytes_per_cell": ALG_BYTES_EVAL}},
        "e2e": {"value": n_cells_total * 3.0 / t_e2e, "unit": "cell-residual evals/s",
                "h2d_bytes_per_step": int(prim.nbytes), "d2h_bytes_per_step": int(prim.nbytes),
                "note": "per step: fs_upload from pinned host memory, prim_to_cons, rk_step, fs_download"},
    }
    if rank == 0:
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_reference_rate(args.n_cpu)
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ibis_b200")
    ap.add_argument("--n", type=int, default=int(os.environ.get("IBIS_BENCH_N", "128")))
    ap.add_argument("--n-cpu", type=int, default=64)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
