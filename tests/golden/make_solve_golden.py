"""Generate tests/golden/solve_*.npz: final flow snapshots of the REAL reference's
make_solver(config)->solve() (oracle/_ref/libibis_ref.so) on tests/solve_cases.py.

    make -C oracle/ref_build -j8 && python tests/golden/make_solve_golden.py [case ...]
"""
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

import refharness as rh  # noqa: E402
import solve_cases as sc  # noqa: E402


def write_case(case_dir, mesh, config, snap, dim):
    """What `ibis prep` leaves behind (prep.py:373-412): config.json, grid, snapshot 0."""
    if os.path.exists(case_dir):
        shutil.rmtree(case_dir)
    for d in ("config", "log", "io/grid/0000", "io/flow/0000"):
        os.makedirs(os.path.join(case_dir, d))
    with open(os.path.join(case_dir, "config", "config.json"), "w") as f:
        json.dump(config, f, indent=1)
    mesh.write(os.path.join(case_dir, "io/grid/0000/block_0000.su2"))
    for k in ["T", "p", "vx", "vy"] + (["vz"] if dim == 3 else []):
        np.ascontiguousarray(snap[k], dtype=np.float64).tofile(os.path.join(case_dir, "io/flow/0000", k))
    with open(os.path.join(case_dir, "io/flow/0000/meta_data.json"), "w") as f:
        json.dump({"time": 0.0}, f)
    with open(os.path.join(case_dir, "io/flow/flows"), "w") as f:
        f.write("0000\n")


def main():
    pad = os.path.join(ROOT, "oracle", "_ref", "libpadmalloc.so")
    if os.path.exists(pad) and "libpadmalloc" not in os.environ.get("LD_PRELOAD", ""):
        # oracle/ref_build/padmalloc.c: the reference's CubicSpline constructor overruns two arrays
        os.environ["LD_PRELOAD"] = pad
        os.execv(sys.executable, [sys.executable] + sys.argv)
    names = sys.argv[1:] or list(sc.SOLVE_CASES)
    # ONE OpenMP thread: the reference's parallel_reduce then sums in index order, which makes the
    # fixtures independent of the machine they were generated on
    # IBIS_REF_THREADS=8 python tests/golden/make_solve_golden.py steady_wedge2d  ->  solve_steady_wedge2d_t8.npz:
    # the same run with the reference's reductions summed by eight threads, i.e. in another order.  The difference
    # between the two fixtures is the reference's OWN spread on this case (tests/test_gpu_solvers.py).
    n_threads = int(os.environ.get("IBIS_REF_THREADS", "1"))
    suffix = "" if n_threads == 1 else f"_t{n_threads}"
    rl = rh.RefLib.get(n_threads)
    for name in names:
        mesh, config, g, prim0, snap = sc.build_solve_case(name)
        case_dir = os.path.join("/tmp/ibis_solve_golden", name)
        write_case(case_dir, mesh, config, snap, mesh.dim)
        rc = rl.solve(case_dir)
        flow_dirs = sorted(d for d in os.listdir(os.path.join(case_dir, "io/flow")) if d.isdigit())
        last = int(flow_dirs[-1])
        out = {"rc": np.array(rc), "n_snapshots": np.array(len(flow_dirs))}
        for k, v in rh.read_flow(case_dir, last, mesh.dim).items():
            out["final_" + k] = v
        with open(os.path.join(case_dir, "io/flow", flow_dirs[-1], "meta_data.json")) as f:
            out["final_time"] = np.array(json.load(f)["time"])
        for log in ("residuals.dat", "absolute_residuals.dat", "gmres_diagnostics.dat"):
            p = os.path.join(case_dir, "log", log)
            if os.path.exists(p):
                rows = [ln.split() for ln in open(p).read().strip().split("\n")[1:] if ln.strip()]
                if rows:
                    out["log_" + log.split(".")[0]] = np.array(rows, dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, "solve_" + name + suffix + ".npz"), **out)
        print(name, "rc", rc, "cells", g.n_cells, "snapshots", flow_dirs, "t", float(out["final_time"]))


if __name__ == "__main__":
    main()
