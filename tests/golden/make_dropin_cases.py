"""Generate tests/golden/dropin/<case>/: the reference's own tests/minimal cases, prepared by the reference's own
`ibis prep` (job.py -> config/config.json, grid, snapshot 0; oracle/refharness.py::run_prep) and solved by the REAL
reference (oracle/_ref/libibis_ref.so, Solver::solve, one OpenMP thread) -- the final snapshot is what
tests/test_gpu_dropin.py holds the drop-in library (the same reference with its FiniteVolume bound to
libibis_b200.so) to.

    make -C oracle/ref_build -j8 && python tests/golden/make_dropin_cases.py
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

MINIMAL = "/root/reference/tests/minimal"
# case -> (directory with job.py, directory with the grid file the job names)
CASES = {
    "uniform_stream": ("uniform_stream", "uniform_stream"),
    "uniform_stream_3D": ("uniform_stream_3D", "uniform_stream_3D"),
    "steady_state_uniform_stream": ("steady_state_uniform_stream", "uniform_stream"),
}
KEEP = ["config/config.json", "io/grid/0000/block_0000.su2", "io/flow/flows", "io/flow/0000"]


def main():
    rl = rh.RefLib.get(1)
    for name, (job_dir, grid_dir) in CASES.items():
        work = os.path.join("/tmp/ibis_dropin_golden", name)
        if os.path.exists(work):
            shutil.rmtree(work)
        os.makedirs(work)
        for d in (grid_dir, job_dir):
            for f in os.listdir(os.path.join(MINIMAL, d)):
                shutil.copy(os.path.join(MINIMAL, d, f), os.path.join(work, f))
        config = rh.run_prep(work)
        dim = int(config["grid"]["dimensions"]) if "dimensions" in config["grid"] else (3 if "3D" in name else 2)
        dst = os.path.join(HERE, "dropin", name)
        if os.path.exists(dst):
            shutil.rmtree(dst)
        for k in KEEP:                                   # the prepared case, before the solve
            src = os.path.join(work, k)
            out = os.path.join(dst, k)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            if os.path.isdir(src):
                shutil.copytree(src, out)
            else:
                shutil.copy(src, out)
        rc = rl.solve(work)
        flow_dirs = sorted(d for d in os.listdir(os.path.join(work, "io/flow")) if d.isdigit())
        final = {"rc": np.array(rc), "n_snapshots": np.array(len(flow_dirs)), "dim": np.array(dim)}
        for k, v in rh.read_flow(work, int(flow_dirs[-1]), dim).items():
            final["final_" + k] = v
        with open(os.path.join(work, "io/flow", flow_dirs[-1], "meta_data.json")) as f:
            final["final_time"] = np.array(json.load(f)["time"])
        np.savez_compressed(os.path.join(dst, "expected.npz"), **final)
        print(name, "rc", rc, "snapshots", len(flow_dirs), "cells", len(final["final_T"]), "t", float(final["final_time"]))


if __name__ == "__main__":
    main()
