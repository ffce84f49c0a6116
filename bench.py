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
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

# algorithmic bytes per cell of one residual evaluation = 1030 (pass G 392 +
# pass F 638).  The default (sides) path splits the work differently: k_grad_sides does pass G AND
# the MUSCL extrapolation + thermodynamic closure of pass F (its state / gradient / limiter /
# centroid / face-centre items: 224 B), k_flux_tile the rest of pass F (face frames 240, face and
# cell-face ids 54, volume 8, residual 40, face centres already counted: 414 B... see DESIGN.md section 3).
ALG_BYTES_EVAL = 1030
ALG_BYTES_SPLIT = {"sides": (616, 414), "gather": (392, 638)}
KERNEL_NAMES = {"sides": ("k_grad_sides (gradients + limiters + MUSCL sides)", "k_flux_tile (flux + surface integral)"),
                "gather": ("k_grad_limit (gradients + limiters)", "k_residual_tile (reconstruct + flux + surface integral)")}


def ncu_traffic_per_cell(path):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of pass G' / pass F', per cell, from the committed
    `ncu --set full` capture of the shipped kernels: profiles/traffic.json, written by tools/ncu_traffic.py from the
    raw page of the capture it names.  None when there is no capture for this path."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t[path]
    except Exception:                                                   # noqa: BLE001
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled through NVML from a thread of this
    process every 10 ms DURING the timed region (nvidia-smi -lms starts too slowly for a region
    of a few hundred milliseconds)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index=0):
        self.index, self.sm, self.mask, self.stop_flag, self.thread, self.max_mhz = index, [], 0, False, None, None
        self.power = []

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:                                          # noqa: BLE001
            self.err = str(e)
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(reasons(self.h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:                                           # noqa: BLE001
                pass
            time.sleep(0.01)

    def stop(self):
        if not self.thread:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")],
                    "samples": 0}
        self.stop_flag = True
        self.thread.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": [n for bit, n in self.REASONS.items() if self.mask & bit], "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None}


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
    generator.  With world > 1 the SAME n^3 box is decomposed by cells into z-slabs (strong
    scaling, BASELINE.json configs[4]): this rank owns the planes [n*rank/world, n*(rank+1)/world)
    and carries its two-ring halo and exchange lists."""
    from ibis_b200.boxgrid import BoxGrid
    if world == 1:
        return BoxGrid(n)
    return BoxGrid(n, n, n, slab=(n * rank // world, n * (rank + 1) // world))


def host_threads():
    """Every core this process may run on: the reference arm uses them all.  torch.distributed.run exports
    OMP_NUM_THREADS=1 to its workers; the reference's OpenMP runtime must not inherit that."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.setdefault("OMP_PROC_BIND", "false")
    return n


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_reference_rate(n_cpu, min_seconds=10.0):
    """The reference's own Kokkos-OpenMP residual (oracle/_ref) on an n_cpu^3 box: evals/s."""
    n_threads = host_threads()
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cases
    import refharness as rh
    if not rh.available():
        return cpu_port_rate(min(n_cpu, 32), min_seconds)
    mesh, g = box_grid_host(n_cpu)
    path = f"/tmp/bench_box_{n_cpu}.su2"
    mesh.write(path)
    rc = rh.RefCase(path, box_config(), num_threads=n_threads)
    prim = cases.perturbed_state(g, 3)
    t1 = rc.time_dudt(prim, 1)
    nrep = max(3, int(min_seconds / max(t1, 1e-6)))
    nrep = min(nrep, 200)
    t = rc.time_dudt(prim, nrep)
    cores = rc.rl.num_threads
    rc.close()
    return {"value": g.n_cells * nrep / t, "unit": "cell-residual evals/s", "cores": cores,
            "kind": "reference", "cpu": cpu_model(), "n_edge": n_cpu,
            "sample": f"{nrep} FiniteVolume::compute_dudt calls of the compiled reference (Kokkos-OpenMP, "
                      f"{cores} threads, {cpu_model()}) on the {n_cpu}^3 = {g.n_cells}-cell box, same config and field; "
                      "the bare residual evaluation, without the stage updates, dt estimate and conversions the GPU "
                      "arm's RungeKutta steps include (conservative for the GPU arm)"}


def cpu_port_rate(n_cpu, min_seconds=10.0):
    """Fallback when the compiled reference did not travel: the numpy restatement of the reference's residual
    (oracle/restatement.py, pinned bit for bit to the compiled reference) on one core, n_cpu^3 box."""
    try:
        import time
        import cases
        import restatement
        mesh, g = box_grid_host(n_cpu)
        config = box_config()
        prim = cases.perturbed_state(g, 3)
        restatement.compute_dudt(g, config, prim, True)
        t0, nrep = time.perf_counter(), 0
        while nrep < 3 or (time.perf_counter() - t0 < min_seconds and nrep < 200):
            restatement.compute_dudt(g, config, prim, True)
            nrep += 1
        t = time.perf_counter() - t0
        return {"value": g.n_cells * nrep / t, "unit": "cell-residual evals/s", "cores": 1, "kind": "port", "n_edge": n_cpu,
                "cpu": cpu_model(),
                "sample": f"{nrep} evaluations of the numpy restatement of FiniteVolume::compute_dudt (oracle/restatement.py, "
                          f"one core) on the {n_cpu}^3 = {g.n_cells}-cell box, same config and field"}
    except Exception:                                            # the baseline is a report, never a reason to fail
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference_rate(args.n_cpu, min_seconds=max(2.0, 4.0 * args.steps))
    if cb is None:
        print(json.dumps({"impl": "reference", "unavailable": "neither oracle/_ref/libibis_ref.so nor oracle/restatement.py usable"}))
        return
    n_edge = cb["n_edge"]
    n_cells = n_edge ** 3
    # the compiled reference did not travel: say so at top level, a numpy port on one core is not "the reference"
    impl = "reference" if cb["kind"] == "reference" else "reference-port(numpy restatement, 1 core)"
    line = {"impl": impl, "metric": "fp64 cell-residual evals/sec", "value": cb["value"],
            "unit": "cell-residual evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 3.0 * n_cells / cb["value"] * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"box {n_edge}^3 hex (bounded sample of the box workload), ausmdv, order 2, "
                                   "barth_jespersen, rho_u, inviscid"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "cell-residual evals/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---- JFNK half of the path: one GMRES iteration of the steady solver =
#   SteadyStateLinearisation::matrix_vector_product (steady_state.cpp:52-128): set the dual parts of the conserved
#       state to the Krylov vector, conserved -> primitive, one dual-number residual evaluation, result = v / dt* - dR.du
#   + one modified Gram-Schmidt step of Gmres::solve (gmres.cpp:144-150; dense_linear_algebra.h:185-309).
# Algorithmic bytes per cell (n_cons = 5, doubles; dual = 16 B):
#   dual second-order evaluation 1470 (the 1030 of the real one with 16-byte state / gradient / residual items)
#   set_dual 80 + 40 + 80, cons_to_prim 80 + 112, set_vector 80 + 40 + 40                               = 552
#   Gram-Schmidt step j: (j + 1) x [dot 80 + axpy 120] + norm 40 + scale 80 + copy into the basis 80     = 200 (j + 2)
JFNK_EVAL_BYTES, JFNK_GLUE_BYTES = 1470, 552


def jfnk_arnoldi_bytes(j):
    return 200 * (j + 2)


def run_jfnk_on(ctx, g, prim, m, cycles, warmup):
    """cycles x m GMRES iterations (matvec + Arnoldi step) on an existing context -> timings (CUDA events on the
    context's stream) and per-kernel-group roofline figures."""
    import torch
    n, nc = int(g.n_cells), 5
    stream = torch.cuda.ExternalStream(ctx.stream_ptr())
    fs, cq, tmp, dudt = ctx.fs_create(dual=True), ctx.cq_create(dual=True), ctx.cq_create(dual=True), ctx.cq_create(dual=True)
    ctx.fs_upload(fs, prim, 0.0 * prim)
    ctx.prim_to_cons(cq, fs)
    v, w = ctx.vec_create(n * nc), ctx.vec_create(n * nc)
    basis = ctx.basis_create(n * nc, m + 1)
    rng = np.random.default_rng(7)
    x0 = rng.standard_normal(n * nc)
    ctx.vec_upload(v, x0 / np.linalg.norm(x0))
    ctx.basis_set_col(basis, 0, v)
    dt_star = 1e-4

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        return e

    def iteration(j, marks=None):
        ctx.jfnk_set_dual(cq, v, tmp)
        ctx.cons_to_prim(tmp, fs)
        if marks is not None:
            marks.append(ev())
        ctx.compute_dudt(fs, dudt, True)
        if marks is not None:
            marks.append(ev())
        ctx.jfnk_set_vector(dudt, v, dt_star, w)
        if marks is not None:
            marks.append(ev())
        ctx.arnoldi_mgs(basis, j, w, v)
        if marks is not None:
            marks.append(ev())

    for _ in range(warmup):
        for j in range(m):
            iteration(j)
    ctx.sync()
    t_glue = t_eval = t_arn = 0.0
    l0 = ctx.launch_count()
    with torch.cuda.stream(stream):
        e0 = ev()
        all_marks = []
        for _ in range(cycles):
            for j in range(m):
                marks = [ev()]
                iteration(j, marks)
                all_marks.append(marks)
        e1 = ev()
    ctx.sync()
    torch.cuda.synchronize()
    launches = ctx.launch_count() - l0
    for mk in all_marks:
        t_glue += mk[0].elapsed_time(mk[1]) + mk[2].elapsed_time(mk[3])
        t_eval += mk[1].elapsed_time(mk[2])
        t_arn += mk[3].elapsed_time(mk[4])
    ms_total = e0.elapsed_time(e1)
    n_it = cycles * m
    peak, peak_kind = measured_peak()
    arn_bytes = sum(jfnk_arnoldi_bytes(j) for j in range(m)) * cycles * n

    def roof(bytes_, ms):
        a = bytes_ / (ms * 1e-3) / 1e9
        return {"achieved": a, "frac": a / peak, "ms": ms / n_it}

    # BLAS-1/2 of the Krylov solver alone (dense_linear_algebra.h): dot, axpy, gemv over the basis
    def timed(fn, reps=10):
        fn()
        ctx.sync()
        with torch.cuda.stream(stream):
            a = ev()
            for _ in range(reps):
                fn()
            b = ev()
        ctx.sync()
        return a.elapsed_time(b) / reps

    y = np.linspace(0.1, 1.0, m)
    blas = {
        "dot": roof(16.0 * n * nc * n_it, timed(lambda: ctx.vec_dot(v, w)) * n_it),
        "axpy": roof(24.0 * n * nc * n_it, timed(lambda: ctx.vec_axpy(w, v, 0.5)) * n_it),
        "gemv(m=%d)" % m: roof(8.0 * (m + 1) * n * nc * n_it, timed(lambda: ctx.basis_gemv(basis, m, y, w)) * n_it),
    }
    for k in blas.values():
        k.pop("ms")
    total_bytes = (JFNK_EVAL_BYTES + JFNK_GLUE_BYTES) * n * n_it + arn_bytes
    out = {"gmres_iterations": n_it, "krylov_dim": m, "ms_per_iteration": ms_total / n_it,
           "iterations_per_s": n_it / (ms_total * 1e-3), "dual_evals_per_s": n * n_it / (ms_total * 1e-3),
           "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                        "iteration": dict(roof(total_bytes, ms_total), algorithmic_bytes_per_cell=total_bytes / n / n_it),
                        "dual_evaluation": dict(roof(JFNK_EVAL_BYTES * n * n_it, t_eval), algorithmic_bytes_per_cell=JFNK_EVAL_BYTES),
                        "matvec_glue": dict(roof(JFNK_GLUE_BYTES * n * n_it, t_glue), algorithmic_bytes_per_cell=JFNK_GLUE_BYTES),
                        "arnoldi_mgs": dict(roof(arn_bytes, t_arn), algorithmic_bytes_per_cell=arn_bytes / n / n_it),
                        "blas": blas}}
    for h in (fs, cq, tmp, dudt, v, w, basis):
        ctx.free(h)
    return out


def run_jfnk(args, n=None, as_secondary=False):
    import torch
    import cases
    from ibis_b200.lib import Context
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    n = n or args.n
    g = box_grid_fast(n)
    prim = cases.perturbed_state(g, 3)
    ctx = Context(g, box_config(), device=int(os.environ.get("LOCAL_RANK", "0")))
    sampler = ClockSampler(0)
    sampler.start()
    r = run_jfnk_on(ctx, g, prim, args.krylov, max(1, args.steps // 2 if as_secondary else args.steps), 1 if as_secondary else max(1, args.warmup // 3))
    clocks = sampler.stop()
    ctx.close()
    workload = (f"JFNK half of the path on the box {n}^3 = {n ** 3} hex cells: GMRES({args.krylov}) iterations, each one exact "
                "Jacobian-vector product (dual-number second-order ausmdv residual, steady_state.cpp:52-128) + one modified "
                "Gram-Schmidt step (gmres.cpp:144-150); perturbed field, random unit start vector")
    r["config"] = {"workload": workload, "cells_total": n ** 3}
    if as_secondary:
        return r
    line = {"metric": "dual fp64 cell-residual evals/sec (GMRES iterations x cells)", "value": r["dual_evals_per_s"],
            "unit": "cell-residual evals/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_iteration"] * args.krylov, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 dual", "data": "synthetic", "config": r.pop("config"), "gpu_launches": r.pop("gpu_launches"),
            "clocks": clocks, "roofline": r.pop("roofline"), "jfnk": r}
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
        halo_mode = ctx.connect(dist, rank, world, [rank + d for d in g.peer_dir], g.send_offsets, g.send_cells,
                                g.recv_offsets, g.recv_cells, p2p=os.environ.get("IBIS_B200_HALO", "p2p") != "nccl")
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
            # asynchronous form of ibis_b200_rk_step: dt_prev chained on the device, nothing read back per step
            # (tests/test_properties.py: same steps bit for bit as the form that returns dt)
            ctx.rk_step_async(fs, cq, tab["a"], tab["b"], 0.5, 1.0)
        e1.record(stream)
    ctx.sync()
    torch.cuda.synchronize()
    dt_prev, _ = ctx.rk_last_dt()
    ms_total = e0.elapsed_time(e1)
    ms_local = ms_total
    launches = ctx.launch_count() - l0
    prof_ms, prof_n = ctx.profile_read()
    ctx.profile_enable(False)
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    bad = ctx.count_bad_cells(fs)

    # end-to-end: the flow state lives on the HOST, in pinned memory, as the reference's own
    # native_binary snapshot (io/native.cpp: one f64 array per variable T, p, vx, vy, vz over the valid
    # cells); every step reads the snapshot in (H2D + update_thermo_from_pT), converts, takes one
    # RungeKutta::take_step and writes the new snapshot back out (D2H) through the C ABI
    n_valid = int(g.n_cells)
    snap = np.ascontiguousarray(prim[:n_valid, [2, 1, 4, 5, 6]].T)          # T p vx vy vz
    host_in = ctx.pinned_array((5, n_valid))
    host_in[:] = snap
    host_out = [ctx.pinned_array((5, n_valid)) for _ in range(2)]
    n_e2e = max(10, args.steps)

    def e2e_loop(n_steps):
        # the snapshot ring (include/ibis_b200.h): snapshot k + 1 arrives and snapshot k - 1 leaves while step k runs
        ctx.snapshot_upload_begin(0, host_in)
        for k in range(n_steps):
            slot = k & 1
            ctx.snapshot_upload_end(slot, fs)                          # NativeBinaryInput::read + update_thermo_from_pT
            if k + 1 < n_steps:
                ctx.snapshot_upload_begin(slot ^ 1, host_in)
            ctx.prim_to_cons(cq, fs)
            ctx.rk_step_async(fs, cq, tab["a"], tab["b"], 0.5, 1.0)
            if k >= 2:
                ctx.snapshot_download_end(slot)                        # snapshot k - 2 has landed; its buffer is reused
            ctx.snapshot_download_begin(slot, fs, host_out[slot])      # NativeBinaryOutput::write's arrays
        for slot in range(min(2, n_steps)):
            ctx.snapshot_download_end(slot)
        ctx.sync()

    e2e_loop(2)                                                        # warm-up: ring allocation, first touches
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_loop(n_e2e)
    t_e2e = (time.perf_counter() - t0) / n_e2e
    assert np.isfinite(host_out[(n_e2e - 1) & 1]).all()
    if world > 1:
        t = torch.tensor([t_e2e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())

    n_cells_total = n ** 3
    ms_per_step = ms_total / args.steps
    value = n_cells_total * 3.0 * args.steps / (ms_total * 1e-3)
    peak, peak_kind = measured_peak()
    # per evaluation: pass G' is two launches (early / late tiles) when the halo exchange is overlapped
    n_eval = max(int(prof_n[1]), 1)
    tF = prof_ms[1] / n_eval * 1e-3
    tG = prof_ms[0] / n_eval * 1e-3
    path = "gather" if os.environ.get("IBIS_B200_SIDES", "1") == "0" else "sides"
    bytes_G, bytes_F = ALG_BYTES_SPLIT[path]
    name_G, name_F = KERNEL_NAMES[path]
    ach_G, ach_F = bytes_G * n_owned / tG / 1e9, bytes_F * n_owned / tF / 1e9
    dom = "G" if tG >= tF else "F"          # the dominant kernel = the one with the longer launches
    per_cell = ncu_traffic_per_cell(path)
    traffic = {k: per_cell[k] * n_owned for k in ("G", "F")} if per_cell else None
    line = {
        "metric": "fp64 cell-residual evals/sec", "value": value, "unit": "cell-residual evals/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"box {n}^3 = {n ** 3} hex cells in unstructured form "
                               f"(BASELINE.json configs[4]{'' if n == 400 else ', reduced size'}), ausmdv, order 2, "
                               "barth_jespersen, rho_u, inviscid, ssp-rk3 (3 residual evaluations per step), "
                               "perturbed field; working set " + f"{n_owned * 1030 / 1e9:.1f} GB per GPU and evaluation, "
                               + ("larger than L2" if n_owned * 1030 > 2 * 126e6 else "L2 RESIDENT (not a valid bench size)"),
                   "cells_total": n ** 3, "cells_per_gpu": n_owned,
                   "parallelism": f"dd{world} (cells decomposed into z-slabs, two-ring halo, NCCL)" if world > 1 else "dd1",
                   "fmad": False, "grid_build_s": round(t_grid, 2), "setup_s": round(t_setup, 2),
                   "bad_cells": bad},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": name_G if dom == "G" else name_F,
                     "achieved": ach_G if dom == "G" else ach_F, "peak": peak, "unit": "GB/s",
                     "frac": (ach_G if dom == "G" else ach_F) / peak,
                     "traffic": traffic[dom] if traffic else None,
                     "traffic_source": per_cell.get("source") if per_cell else None, "peak_kind": peak_kind,
                     "algorithmic_bytes_per_cell": bytes_G if dom == "G" else bytes_F,
                     "avg_launch_ms": (tG if dom == "G" else tF) * 1e3,
                     "path": path,
                     "pass_G": {"kernel": name_G, "avg_launch_ms": tG * 1e3, "achieved": ach_G, "frac": ach_G / peak,
                                "algorithmic_bytes_per_cell": bytes_G, "traffic": traffic["G"] if traffic else None,
                                "dram_gbs": traffic["G"] / tG / 1e9 if traffic else None},
                     "pass_F": {"kernel": name_F, "avg_launch_ms": tF * 1e3, "achieved": ach_F, "frac": ach_F / peak,
                                "algorithmic_bytes_per_cell": bytes_F, "traffic": traffic["F"] if traffic else None,
                                "dram_gbs": traffic["F"] / tF / 1e9 if traffic else None},
                     "evaluation": {"achieved": ALG_BYTES_EVAL * n_owned / (tF + tG) / 1e9,
                                    "frac": ALG_BYTES_EVAL * n_owned / (tF + tG) / 1e9 / peak,
                                    "algorithmic_bytes_per_cell": ALG_BYTES_EVAL},
                     "halo_exchange_ms": (prof_ms[2] / max(int(prof_n[2]), 1)) if world > 1 else None,
                     "halo_mode": halo_mode if world > 1 else None},
        "e2e": {"value": n_cells_total * 3.0 / t_e2e, "unit": "cell-residual evals/s",
                "h2d_bytes_per_step": int(snap.nbytes), "d2h_bytes_per_step": int(snap.nbytes),
                "steps": n_e2e,
                "note": "per step and rank: the reference's native_binary snapshot arrays T p vx vy vz from page-locked host "
                        "memory to the device (+ update_thermo_from_pT), prim_to_cons, one ssp-rk3 step, the new snapshot back "
                        "to page-locked host memory -- through the snapshot ring of the C ABI: the copies of step k + 1 / k - 1 "
                        "overlap the compute of step k (wall clock over all steps, copies included)"},
    }
    if world > 1:
        mine = {"rank": rank, "cells": n_owned, "step_ms": ms_local / args.steps, "pass_F_ms": tF * 1e3, "pass_G_ms": tG * 1e3,
                "halo_wait_ms": prof_ms[2] / max(int(prof_n[2]), 1)}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
        line["per_rank"] = per_rank
    ctx.close()
    if rank == 0:
        if world == 1 and not args.no_secondary:
            # the JFNK half of the path, measured on a box that fits beside nothing else (the headline context is closed)
            try:
                line["secondary"] = {"jfnk": run_jfnk(args, n=min(n, args.n_jfnk), as_secondary=True)}
            except Exception as e:                                       # noqa: BLE001  (a report, never a reason to fail)
                line["secondary"] = {"jfnk": {"error": str(e)[:300]}}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_reference_rate(args.n_cpu)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ibis_b200")
    ap.add_argument("--n", "--edge", dest="n", type=int, default=int(os.environ.get("IBIS_BENCH_N", "400")),
                    help="edge of the global hex box (400 -> the 64 M-cell grid of BASELINE.json)")
    ap.add_argument("--n-cpu", type=int, default=128,
                    help="edge of the box the reference arm / cpu_baseline leg runs (BASELINE.md section 4.3: 128)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="explicit", choices=["explicit", "jfnk"],
                    help="explicit: ssp-rk3 steps (the headline); jfnk: GMRES iterations of the steady solver")
    ap.add_argument("--krylov", type=int, default=10, help="GMRES iterations per cycle of the jfnk workload")
    ap.add_argument("--n-jfnk", type=int, default=200, help="edge of the box of the secondary jfnk measurement")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "jfnk":
        if args.n == 400 and "IBIS_BENCH_N" not in os.environ:
            args.n = 256            # dual state + Krylov basis of the 64 M-cell box do not fit beside each other
        run_jfnk(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
